// Bring-up harness of wgrad3x3_kernel (MN-major tcgen05 operands straight from the planes): correctness against a CPU
// reference on two small plane segments, both operand modes, several split counts; then a timing of the FCOS-tower shape.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/test_wgrad csrc/test_wgrad.cu && ./build/test_wgrad
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "conv_gemm_host.cuh"
#include "wgrad3x3.cuh"

using namespace sylph;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

static uint32_t g_seed = 777;
static float frand() { g_seed = g_seed * 1664525u + 1013904223u; return ((g_seed >> 8) & 0xFFFF) / 32768.0f - 1.0f; }
static int round128(int x) { return (x + 127) / 128 * 128; }

static int run(const std::vector<std::pair<int, int>>& hw, bool split, int splits, bool timing) {
    std::vector<Seg> segs;
    int rows = 0;
    for (auto& s : hw) {
        Seg g;
        g.row0 = rows; g.H = s.first; g.W = s.second; g.pad = 1; g.Wp = g.W + 2; g.nrows = (g.H + 2) * g.Wp;
        rows += round128(g.nrows);
        segs.push_back(g);
    }
    std::vector<int> tile_seg(rows / 128, 0);
    for (size_t i = 0; i < segs.size(); ++i)
        for (int t = segs[i].row0 / 128; t < (segs[i].row0 + segs[i].nrows + 127) / 128; ++t) tile_seg[t] = static_cast<int>(i);
    const int ld = split ? 512 : 256;
    std::vector<__half> hx(static_cast<size_t>(rows) * ld, __float2half(0.f)), hdy(hx.size(), __float2half(0.f));
    std::vector<float> fx(static_cast<size_t>(rows) * 256, 0.f), fdy(fx.size(), 0.f), fxl(fx.size(), 0.f), fdyl(fx.size(), 0.f);
    for (auto& g : segs)
        for (int y = 0; y < g.H; ++y)
            for (int x = 0; x < g.W; ++x) {
                const size_t r = g.row0 + static_cast<size_t>(y + 1) * g.Wp + x + 1;
                for (int c = 0; c < 256; ++c) {
                    const float a = frand(), b = frand() * 0.5f;
                    const __half ah = __float2half_rn(a), bh = __float2half_rn(b);
                    hx[r * ld + c] = ah; hdy[r * ld + c] = bh;
                    fx[r * 256 + c] = __half2float(ah); fdy[r * 256 + c] = __half2float(bh);
                    if (split) {
                        const __half al = __float2half_rn(a - __half2float(ah)), bl = __float2half_rn(b - __half2float(bh));
                        hx[r * ld + 256 + c] = al; hdy[r * ld + 256 + c] = bl;
                        fxl[r * 256 + c] = __half2float(al); fdyl[r * 256 + c] = __half2float(bl);
                    }
                }
            }
    __half *dx, *ddy;
    Seg* dsegs; int* dts; float *dpart, *dw;
    CK(cudaMalloc(&dx, hx.size() * 2)); CK(cudaMalloc(&ddy, hdy.size() * 2));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ddy, hdy.data(), hdy.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dsegs, segs.size() * sizeof(Seg))); CK(cudaMemcpy(dsegs, segs.data(), segs.size() * sizeof(Seg), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dts, tile_seg.size() * 4)); CK(cudaMemcpy(dts, tile_seg.data(), tile_seg.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dpart, static_cast<size_t>(splits) * 9 * 256 * 256 * 4));
    CK(cudaMalloc(&dw, 9 * 256 * 256 * 4));
    CUtensorMap tdy, tx;
    std::string err;
    if (make_tmap_2d(&tdy, ddy, rows, ld, ld, kWgTileK, &err) || make_tmap_2d(&tx, dx, rows, ld, ld, kWgTileK, &err)) { printf("tmap: %s\n", err.c_str()); return 1; }
    WgradArgs a{dsegs, dts, rows / kWgTileK, dpart};
    auto launch = [&]() {
        if (split) {
            CK(cudaFuncSetAttribute(wgrad3x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgradSmem<true>::kTotal));
            CK(launch_k(wgrad3x3_kernel<true>, dim3(splits, 9, 2), dim3(kWgThreads), WgradSmem<true>::kTotal, 0, tdy, tx, a));
        } else {
            CK(cudaFuncSetAttribute(wgrad3x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgradSmem<false>::kTotal));
            CK(launch_k(wgrad3x3_kernel<false>, dim3(splits, 9, 2), dim3(kWgThreads), WgradSmem<false>::kTotal, 0, tdy, tx, a));
        }
        CK(launch_k(wgrad_reduce_kernel, dim3(9 * 256), dim3(256), 0, 0, static_cast<const float*>(dpart), splits, dw));
    };
    launch();
    CK(cudaDeviceSynchronize());
    if (timing) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < 3; ++i) launch();
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) launch();
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
        const double flops = 2.0 * rows * 256.0 * 256.0 * 9.0 * (split ? 3 : 1);
        printf("[bench wgrad rows=%d split=%d splits=%d] %.3f ms  %.1f TFLOP/s executed\n", rows, split, splits, ms, flops / ms * 1e-9);
        return 0;
    }
    std::vector<float> w(9 * 256 * 256);
    CK(cudaMemcpy(w.data(), dw, w.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, max_ref = 0;
    // reference on a sample of (co, ci) pairs, all taps
    for (int co = 0; co < 256; co += 7)
        for (int ci = 0; ci < 256; ci += 5)
            for (int tap = 0; tap < 9; ++tap) {
                const int dy = tap / 3 - 1, dxo = tap % 3 - 1;
                double s = 0;
                for (auto& g : segs)
                    for (int y = 0; y < g.H; ++y)
                        for (int x = 0; x < g.W; ++x) {
                            const size_t r = g.row0 + static_cast<size_t>(y + 1) * g.Wp + x + 1;
                            const size_t rb = r + dy * g.Wp + dxo;
                            const double ah = fdy[r * 256 + co], bh = fx[rb * 256 + ci];
                            s += ah * bh;
                            if (split) s += static_cast<double>(fdyl[r * 256 + co]) * bh + ah * static_cast<double>(fxl[rb * 256 + ci]);
                        }
                const double got = w[(co * 256 + ci) * 9 + tap];
                max_err = std::max(max_err, std::fabs(got - s));
                max_ref = std::max(max_ref, std::fabs(s));
            }
    const bool ok = max_err <= 2e-5 * std::max(max_ref, 1.0) + 1e-5;
    printf("[wgrad correctness segs=%zu rows=%d split=%d splits=%d] %s max_abs_err=%.3e max_ref=%.3e\n", segs.size(), rows, split, splits,
           ok ? "PASS" : "FAIL", max_err, max_ref);
    cudaFree(dx); cudaFree(ddy); cudaFree(dsegs); cudaFree(dts); cudaFree(dpart); cudaFree(dw);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    const bool quick = argc > 1 && std::string(argv[1]) == "quick";   // correctness cases only (compute-sanitizer runs)
    int fails = 0;
    const std::vector<std::pair<int, int>> small = {{20, 24}, {9, 11}};
    fails += run(small, false, 3, false);
    fails += run(small, true, 3, false);
    fails += run(small, true, 1, false);
    fails += run(small, true, 20, false);            // more splits than K tiles: idle CTAs write zeros
    fails += run({{13, 21}, {7, 11}, {25, 42}}, true, 5, false);
    // FCOS class tower, 3 query images at 800 x 1333: p3..p7 planes of each image
    std::vector<std::pair<int, int>> tower;
    for (int n = 0; n < 3; ++n) { tower.push_back({100, 168}); tower.push_back({50, 84}); tower.push_back({25, 42}); tower.push_back({13, 21}); tower.push_back({7, 11}); }
    if (fails == 0 && !quick) { run(tower, true, 8, true); run(tower, true, 16, true); run(tower, false, 8, true); }
    printf("%s\n", fails == 0 ? "ALL PASS" : "FAILURES");
    return fails;
}
