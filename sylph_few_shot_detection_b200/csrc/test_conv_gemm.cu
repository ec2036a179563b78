// Standalone bring-up harness for conv_gemm_f16_kernel (not part of the shipped library).
// Each case is checked against a double-precision CPU evaluation of the same flat-plane formula on the same
// fp16-rounded operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o test_conv_gemm test_conv_gemm.cu
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "conv_gemm_host.cuh"

using namespace sylph;

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)

static uint32_t g_seed = 12345;
static float frand() {
    g_seed = g_seed * 1664525u + 1013904223u;
    return ((g_seed >> 8) & 0xFFFF) / 65536.0f * 2.f - 1.f;
}
struct Case {
    const char* name;
    int bn;
    std::vector<Seg> segs;
    int total_rows;       // rows of the out/A buffers (multiple of 128)
    int cin_cols;         // A tensor-map inner dim (elements)
    int a_ld;             // A row pitch (elements); < cin_cols for overlapped rows
    int cout;             // multiple of bn
    int taps;
    int kpt;              // k-blocks (64 elements) per tap
    std::vector<int> dys, dxs;
    int flags;
    bool bias;
    int skew = 0, base_offset = 0;  // experiment: A box loaded `skew` rows early, descriptor started `skew` rows in
    bool pair = false;    // CTA-pair (cta_group::2) halo kernel
    bool nm = false;      // split mode, N-merged narrow 3x3 halo kernel: weights [taps][w_hi rows ; w_lo rows][cin]
    bool split = false;   // split-operand ("exact") mode: A rows [C hi | C lo], weights [w_hi | w_hi | w_lo], fp16 outputs hi | lo
    bool halo = false;    // 3x3 halo pipeline (one 130-row A box per vertical tap and k-block)
    bool staged = false;  // TMA-in / TMA-out epilogue, run IN PLACE (out == residual buffer) like the bottleneck conv3
    bool pair1x1 = false; // CTA-pair 1x1 convolution with the staged epilogue (conv1x1_pair.cuh)
    int sms_override = 0; // pretend the device has this many SMs: many tiles per CTA on a problem the CPU reference finishes quickly
    int pairsplit = 0;    // split mode, CTA-pair 1x1 kernel with the chunked staged epilogue: 1 = K' = 3C loop, 2 = quad stages
    bool qs = false;      // split mode, quad-stage 1x1 kernel (a_hi, a_lo, w_hi, w_lo of a k-block loaded once; conv_gemm.cuh QS)
    bool stem16 = false;  // stem as K = 16 taps: A map = [rows][16] with 32-byte swizzle, 131-row boxes (same math as a_ld = 16)
};

static std::string g_case_filter;   // ./test_conv_gemm case <substring>: run only the matching correctness cases
static int run_case(const Case& c, int num_sms) {
    if (c.sms_override > 0) num_sms = c.sms_override;
    if (!g_case_filter.empty() && std::string(c.name).find(g_case_filter) == std::string::npos) return 0;
    const int M = c.total_rows;
    const int Kt = c.kpt * kBlockK;             // logical K per tap
    const bool sp = c.split;
    const int a_ld = sp ? 2 * c.a_ld : c.a_ld;  // physical A pitch: [C hi | C lo] in split mode
    const int lo_a = sp ? c.a_ld : 0;
    const int Kp = sp ? 3 * Kt : Kt;            // physical K per tap of the weight rows: [w_hi | w_hi | w_lo]
    const int ldc = (sp && !(c.flags & kEpiOutF32)) ? 2 * c.cout : c.cout;   // fp16 outputs carry a lo half
    const int ld_r = sp ? 2 * c.cout : c.cout;
    const size_t a_elems = static_cast<size_t>(M) * a_ld + 128;
    const bool nm = c.nm && sp && c.halo;
    const int w_tiles = (sp && c.stem16) || nm ? 2 * c.taps : c.taps;   // split stem: tiles [hi taps | lo taps]; NM: [tap][hi ; lo]
    const int Kw = (c.stem16 || nm) ? Kt : Kp;
    std::vector<uint16_t> hA(a_elems, 0), hW(static_cast<size_t>(w_tiles) * c.cout * Kw, 0), hR;
    // the values the kernel is expected to multiply: hi (+ lo) of every operand, as doubles
    std::vector<double> vA(static_cast<size_t>(M) * c.a_ld + 128, 0.0), vW(static_cast<size_t>(c.taps) * c.cout * Kt, 0.0), vR;
    std::vector<float> hB(c.cout);
    auto put = [&](float v, uint16_t* hi_dst, uint16_t* lo_dst) -> double {
        const uint16_t hi = float_to_half_bits(v);
        *hi_dst = hi;
        double val = half_bits_to_float(hi);
        if (lo_dst) {
            const uint16_t lo = float_to_half_bits(v - half_bits_to_float(hi));
            *lo_dst = lo;
            val += half_bits_to_float(lo);
        }
        return val;
    };
    for (size_t r = 0; r * c.a_ld + c.a_ld <= vA.size(); ++r)
        for (int k = 0; k < c.a_ld; ++k) {
            const size_t at = r * a_ld + k;
            if (at + lo_a >= hA.size()) continue;
            vA[r * c.a_ld + k] = put(frand() * 1.0003f, &hA[at], sp ? &hA[at + lo_a] : nullptr);
        }
    for (int t = 0; t < c.taps; ++t)
        for (int n = 0; n < c.cout; ++n)
            for (int k = 0; k < Kt; ++k) {
                const float v = frand() * 0.10007f;
                double& val = vW[(static_cast<size_t>(t) * c.cout + n) * Kt + k];
                if (nm) {
                    val = put(v, &hW[(static_cast<size_t>(2 * t) * c.cout + n) * Kt + k], &hW[(static_cast<size_t>(2 * t + 1) * c.cout + n) * Kt + k]);
                } else if (c.stem16 && sp) {
                    val = put(v, &hW[(static_cast<size_t>(t) * c.cout + n) * Kt + k], &hW[(static_cast<size_t>(c.taps + t) * c.cout + n) * Kt + k]);
                } else if (sp) {
                    uint16_t* row = &hW[(static_cast<size_t>(t) * c.cout + n) * Kp];
                    val = put(v, &row[k], &row[2 * Kt + k]);
                    row[Kt + k] = row[k];
                } else {
                    val = put(v, &hW[(static_cast<size_t>(t) * c.cout + n) * Kt + k], nullptr);
                }
            }
    for (auto& v : hB) v = c.bias ? frand() : 0.f;
    if (c.flags & kEpiResidual) {
        hR.assign(static_cast<size_t>(M) * ld_r, 0);
        vR.assign(static_cast<size_t>(M) * c.cout, 0.0);
        for (int r = 0; r < M; ++r)
            for (int n = 0; n < c.cout; ++n)
                vR[static_cast<size_t>(r) * c.cout + n] = put(frand() * 1.0003f, &hR[static_cast<size_t>(r) * ld_r + n],
                                                              sp ? &hR[static_cast<size_t>(r) * ld_r + c.cout + n] : nullptr);
    }
    std::vector<int> tile_seg(M / 128, 0);
    for (size_t s = 0; s < c.segs.size(); ++s) {
        int t0 = c.segs[s].row0 / 128, t1 = (c.segs[s].row0 + c.segs[s].nrows + 127) / 128;
        for (int t = t0; t < t1; ++t) tile_seg[t] = static_cast<int>(s);
    }
    const bool f32out = c.flags & kEpiOutF32;
    const size_t out_bytes = static_cast<size_t>(M) * ldc * (f32out ? 4 : 2);
    __half *dA, *dW, *dR = nullptr;
    float *dB, *dG;
    void* dO;
    int* dTS;
    Seg* dS;
    CK(cudaMalloc(&dA, a_elems * 2));
    CK(cudaMalloc(&dW, hW.size() * 2));
    CK(cudaMalloc(&dB, hB.size() * 4));
    CK(cudaMalloc(&dO, out_bytes));
    CK(cudaMalloc(&dG, static_cast<size_t>(M / 128) * 64 * 4));
    CK(cudaMalloc(&dTS, tile_seg.size() * 4));
    CK(cudaMalloc(&dS, c.segs.size() * sizeof(Seg)));
    CK(cudaMemcpy(dA, hA.data(), a_elems * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dTS, tile_seg.data(), tile_seg.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dS, c.segs.data(), c.segs.size() * sizeof(Seg), cudaMemcpyHostToDevice));
    CK(cudaMemset(dO, 0xFF, out_bytes));
    if (!hR.empty()) {
        CK(cudaMalloc(&dR, hR.size() * 2));
        CK(cudaMemcpy(dR, hR.data(), hR.size() * 2, cudaMemcpyHostToDevice));
    }
    const uint64_t a_rows_dim = (c.a_ld < c.cin_cols) ? static_cast<uint64_t>(M) - (c.cin_cols / c.a_ld - 1) : M;
    CUtensorMap ta, tb;
    std::string err;
    if ((c.stem16 ? make_tmap_2d_k16(&ta, dA, M, 131, &err, sp ? 32 : 16)
                  : make_tmap_2d(&ta, dA, a_rows_dim, sp ? 2 * c.cin_cols : c.cin_cols, a_ld, (c.halo || c.pair) ? 130 : 128, &err)) ||
        make_tmap_2d(&tb, dW, static_cast<uint64_t>(w_tiles) * c.cout, Kw, Kw, (c.pair || c.pair1x1 || c.pairsplit) ? c.bn / 2 : c.bn, &err)) {
        printf("[%s] FAIL tensor map: %s\n", c.name, err.c_str());
        return 1;
    }
    GemmArgs g{};
    g.tile_begin = 0;
    g.num_m_tiles = M / 128;
    g.num_n_tiles = c.cout / c.bn;
    g.a_row_delta = 0;
    g.taps = c.taps;
    g.kblocks_per_tap = nm ? 2 * Kt / kBlockK : ((c.qs || c.pairsplit == 2) ? Kt / kBlockK : Kw / kBlockK);
    g.b_rows_per_tap = nm ? 2 * c.cout : c.cout;
    g.nm_lo_row = (c.stem16 && sp) ? c.taps * c.cout : c.cout;
    g.a_wrap = sp ? 2 * Kt : 0;
    g.out_lo = (sp && !f32out) ? c.cout : 0;
    g.res_lo = sp ? c.cout : 0;
    for (int t = 0; t < c.taps; ++t) { g.tap_dy[t] = c.dys[t]; g.tap_dx[t] = c.dxs[t]; }
    g.bias = c.bias ? dB : nullptr;
    g.residual = dR;
    g.ld_res = ld_r;
    g.out = dO;
    g.ldc = ldc;
    g.flags = c.flags;
    g.tile_seg = dTS;
    g.segs = dS;
    g.gn_partial = dG;
    g.dbg_a_row_skew = c.skew;
    g.dbg_base_offset = c.base_offset;
    if (c.staged) {
        if (!hR.empty()) CK(cudaMemcpy(dO, hR.data(), hR.size() * 2, cudaMemcpyHostToDevice));  // in place: out starts as the residual
        g.residual = static_cast<const __half*>(dO);
        CUtensorMap tio;
        if (make_tmap_2d(&tio, static_cast<const __half*>(dO), M, ldc, ldc, 128, &err)) {
            printf("[%s] FAIL epilogue tensor map: %s\n", c.name, err.c_str());
            return 1;
        }
        if (c.stem16) CK(launch_conv_gemm_stem16(ta, tb, tio, g, num_sms, 0, sp, sp && c.nm));
        else if (c.pair1x1) CK(launch_conv1x1_pair_staged(ta, tb, tio, tio, g, num_sms, 0));
        else if (c.pairsplit) CK(launch_conv1x1_pair_split(c.pairsplit == 2, ta, tb, tio, tio, g, num_sms, 0));
        else if (c.qs) CK(launch_conv_gemm_qs(c.bn, true, ta, tb, tio, tio, g, num_sms, 0));
        else CK(launch_conv_gemm_staged(c.bn, ta, tb, tio, tio, g, num_sms, 0, 0, sp));
    } else if (c.pair) {
        CK(launch_conv3x3_pair(ta, tb, g, num_sms, 0, c.bn, sp));
    } else if (nm) {
        CK(launch_conv_gemm_halo_nm(c.bn, ta, tb, g, num_sms, 0));
    } else if (c.halo) {
        CK(launch_conv_gemm_halo(c.bn, ta, tb, g, num_sms, 0, true, sp));
    } else if (c.qs) {
        CK(launch_conv_gemm_qs(c.bn, false, ta, tb, ta, ta, g, num_sms, 0));
    } else {
        CK(launch_conv_gemm(c.bn, ta, tb, g, num_sms, 0, sp));
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("[%s] FAIL kernel: %s\n", c.name, cudaGetErrorString(e));
        return 1;
    }
    std::vector<uint8_t> hO(out_bytes);
    CK(cudaMemcpy(hO.data(), dO, out_bytes, cudaMemcpyDeviceToHost));
    std::vector<float> hG(static_cast<size_t>(M / 128) * 64);
    CK(cudaMemcpy(hG.data(), dG, hG.size() * 4, cudaMemcpyDeviceToHost));

    double max_err = 0, max_ref = 0;
    std::vector<double> gref(hG.size(), 0.0);
    int bad = 0;
    // split: what is left is the tensor core's fp32 accumulation (it truncates when it aligns addends: up to ~7e-5 on
    // K' = 3 x 2304 sums of magnitude 7, measured) and the dropped lo.lo products (2^-22 relative)
    const double tol = sp ? 1e-4 : (f32out ? 1e-4 : 1.5e-3);
    for (int r = 0; r < M; ++r) {
        const Seg& sg = c.segs[tile_seg[r / 128]];
        int local = r - sg.row0, y = local / sg.Wp, x = local % sg.Wp;
        bool interior = local < sg.nrows && y >= sg.pad && y < sg.pad + sg.H && x >= sg.pad && x < sg.pad + sg.W;
        bool keep = interior || !(c.flags & kEpiMask);
        for (int n = 0; n < c.cout; ++n) {
            double acc = hB[n];
            for (int t = 0; t < c.taps; ++t) {
                long ar = static_cast<long>(r) + c.dys[t] * sg.Wp + c.dxs[t];
                if (ar < 0 || ar >= static_cast<long>(a_rows_dim)) continue;
                const double* arow = &vA[static_cast<size_t>(ar) * c.a_ld];   // overlapped rows (stem): arow[k] runs into the next rows
                const double* wrow = &vW[(static_cast<size_t>(t) * c.cout + n) * Kt];
                for (int k = 0; k < Kt && k < c.cin_cols; ++k) acc += arow[k] * wrow[k];
            }
            if ((c.flags & kEpiGnStats) && interior) {
                gref[(r / 128) * 64 + (n / 8) * 2] += acc;
                gref[(r / 128) * 64 + (n / 8) * 2 + 1] += acc * acc;
            }
            if (c.flags & kEpiResidual) acc += vR[static_cast<size_t>(r) * c.cout + n];
            if (c.flags & kEpiRelu) acc = acc > 0 ? acc : 0;
            if (!keep) acc = 0;
            const size_t idx = static_cast<size_t>(r) * ldc + n;
            double got = f32out ? reinterpret_cast<const float*>(hO.data())[idx]
                                : half_bits_to_float(reinterpret_cast<const uint16_t*>(hO.data())[idx]);
            if (sp && !f32out) got += half_bits_to_float(reinterpret_cast<const uint16_t*>(hO.data())[idx + c.cout]);
            double err2 = fabs(got - acc);
            if (!(err2 == err2)) err2 = 1e30;
            if (err2 > max_err) max_err = err2;
            if (fabs(acc) > max_ref) max_ref = fabs(acc);
            if (c.skew && (r % 128) >= 128 - c.skew) continue;  // those rows read past the 128-row box by construction
            if (err2 > tol * (1.0 + fabs(acc)) && bad < 5) {
                printf("  mismatch r=%d n=%d got=%g ref=%g\n", r, n, got, acc);
                ++bad;
            }
        }
    }
    double gn_err = 0;
    if (c.flags & kEpiGnStats)
        for (size_t i = 0; i < hG.size(); ++i) gn_err = fmax(gn_err, fabs(hG[i] - gref[i]) / (1.0 + fabs(gref[i])));
    bool ok = bad == 0 && gn_err < 1e-3;
    printf("[%s] %s max_abs_err=%.3e max_ref=%.3e gn_rel_err=%.3e\n", c.name, ok ? "PASS" : "FAIL", max_err, max_ref,
           gn_err);
    cudaFree(dA); cudaFree(dW); cudaFree(dB); cudaFree(dO); cudaFree(dG); cudaFree(dTS); cudaFree(dS);
    if (dR) cudaFree(dR);
    return ok ? 0 : 1;
}

static Seg mk_seg(int row0, int H, int W, int pad) {
    Seg s;
    s.row0 = row0; s.H = H; s.W = W; s.pad = pad; s.Wp = W + 2 * pad; s.nrows = (H + 2 * pad) * (W + 2 * pad);
    return s;
}
static int round128(int x) { return (x + 127) / 128 * 128; }

static int g_dbg_skip = 0;
static void bench_shape(const char* name, int bn, int m_tiles, int cin, int cout, int taps, int flags, int num_sms, int staged = 0, bool halo = false, bool pair = false, bool split = false, bool nm = false, bool qs = false) {
    // split: `cin` / `cout` stay the LOGICAL channel counts; the buffers hold [C hi | C lo] rows and [w_hi | w_hi | w_lo] weights
    const int M = m_tiles * 128;
    const int osz = (flags & kEpiOutF32) ? 4 : 2;
    const int cin_l = cin, cout_l = cout;
    const int ldo = (split && osz == 2) ? 2 * cout : cout;
    const int kw = nm ? cin : (split ? 3 * cin : cin);       // NM: [taps][2 cout][cin]
    const int w_rows_per_tap = nm ? 2 * cout : cout;
    if (split) cin *= 2;   // physical A pitch
    __half *dA, *dW, *dR = nullptr;
    void* dO;
    float *dG, *dB;
    int* dTS;
    Seg* dS;
    CK(cudaMalloc(&dA, static_cast<size_t>(M) * cin * 2));
    CK(cudaMalloc(&dW, static_cast<size_t>(taps) * w_rows_per_tap * kw * 2));
    CK(cudaMalloc(&dO, static_cast<size_t>(M) * ldo * osz));
    CK(cudaMalloc(&dB, cout * 4));
    CK(cudaMalloc(&dG, static_cast<size_t>(m_tiles) * 64 * 4));
    CK(cudaMalloc(&dTS, m_tiles * 4));
    CK(cudaMalloc(&dS, sizeof(Seg)));
    {   // non-trivial operand bits so the tensor pipes draw realistic power
        std::vector<uint16_t> h(static_cast<size_t>(M) * cin);
        for (auto& v : h) v = float_to_half_bits(frand());
        CK(cudaMemcpy(dA, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
        std::vector<uint16_t> w(static_cast<size_t>(taps) * w_rows_per_tap * kw);
        for (auto& v : w) v = float_to_half_bits(frand() * 0.05f);
        CK(cudaMemcpy(dW, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
    }
    CK(cudaMemset(dB, 0, cout * 4));
    CK(cudaMemset(dTS, 0, m_tiles * 4));
    Seg s = mk_seg(0, M / 170 - 2, 168, 1);  // p3-like plane width
    CK(cudaMemcpy(dS, &s, sizeof(Seg), cudaMemcpyHostToDevice));
    if (flags & kEpiResidual) {
        CK(cudaMalloc(&dR, static_cast<size_t>(M) * ldo * 2));
        CK(cudaMemset(dR, 0, static_cast<size_t>(M) * ldo * 2));
    }
    CUtensorMap ta, tb;
    std::string err;
    const bool pair1x1 = staged == 9;   // CTA-pair 1x1 kernel with the staged epilogue
    const int pairsplit = staged == 10 ? 1 : staged == 11 ? 2 : 0;   // split-mode CTA-pair 1x1 kernel (K' = 3C loop / quad stages)
    if (make_tmap_2d(&ta, dA, M, cin, cin, (halo || pair) ? 130 : 128, &err) || make_tmap_2d(&tb, dW, static_cast<uint64_t>(taps) * w_rows_per_tap, kw, kw, (pair || pair1x1 || pairsplit) ? bn / 2 : bn, &err)) {
        printf("[%s] tensor map failed: %s\n", name, err.c_str());
        return;
    }
    GemmArgs g{};
    g.num_m_tiles = m_tiles; g.num_n_tiles = cout / bn; g.taps = taps; g.kblocks_per_tap = nm ? 2 * cin_l / kBlockK : ((qs || pairsplit == 2) ? cin_l / kBlockK : kw / kBlockK); g.b_rows_per_tap = w_rows_per_tap; g.nm_lo_row = cout_l;
    g.a_wrap = split ? 2 * cin_l : 0; g.out_lo = (split && osz == 2) ? cout_l : 0; g.res_lo = split ? cout_l : 0;
    for (int t = 0; t < taps; ++t) { g.tap_dy[t] = taps == 9 ? (t / 3) - 1 : 0; g.tap_dx[t] = taps == 9 ? (t % 3) - 1 : 0; }
    g.bias = dB; g.residual = dR; g.ld_res = ldo; g.out = dO; g.ldc = ldo; g.flags = flags; g.tile_seg = dTS; g.segs = dS; g.gn_partial = dG;
    g.dbg_skip = g_dbg_skip;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CUtensorMap tio = ta;
    if (staged) {
        g.residual = static_cast<const __half*>(dO);
        if (make_tmap_2d(&tio, static_cast<const __half*>(dO), M, ldo, ldo, 128, &err)) { printf("tmap failed\n"); return; }
    }
    auto launch = [&]() { return pairsplit ? launch_conv1x1_pair_split(pairsplit == 2, ta, tb, tio, tio, g, num_sms, 0) : qs ? launch_conv_gemm_qs(bn, staged != 0, ta, tb, tio, tio, g, num_sms, 0) : pair1x1 ? launch_conv1x1_pair_staged(ta, tb, tio, tio, g, num_sms, 0) : staged ? launch_conv_gemm_staged(bn, ta, tb, tio, tio, g, num_sms, 0, staged == 100 ? 0 : staged, split) : (pair ? launch_conv3x3_pair(ta, tb, g, num_sms, 0, bn, split) : (nm ? launch_conv_gemm_halo_nm(bn, ta, tb, g, num_sms, 0) : halo ? launch_conv_gemm_halo(bn, ta, tb, g, num_sms, 0, true, split) : launch_conv_gemm(bn, ta, tb, g, num_sms, 0, split))); };
    for (int i = 0; i < 3; ++i) CK(launch());
    CK(cudaDeviceSynchronize());
    const int iters = 10;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) CK(launch());
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    double flop = 2.0 * M * cout * static_cast<double>(split ? 3 * cin_l : cin_l) * taps;   // executed (split: three products per multiply)
    double bytes = static_cast<double>(M) * cin * 2 + static_cast<double>(M) * ldo * (osz + ((flags & kEpiResidual) ? 2 : 0));
    printf("[bench %s] M=%d K=%d N=%d : %.3f ms  %.1f TFLOP/s  %.1f GB/s (algorithmic)\n", name, M, cin * taps, cout, ms,
           flop / ms * 1e-9, bytes / ms * 1e-6);
    cudaFree(dA); cudaFree(dW); cudaFree(dO); cudaFree(dB); cudaFree(dG); cudaFree(dTS); cudaFree(dS);
    if (dR) cudaFree(dR);
}

int main(int argc, char** argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    printf("device %s sm_%d%d SMs=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    const int sms = prop.multiProcessorCount;
    int fails = 0;
    if (argc > 2 && std::string(argv[1]) == "case") g_case_filter = argv[2];
    if (argc > 2 && std::string(argv[1]) == "only") {   // one shape only (ncu captures): ./test_conv_gemm only <name>
        const std::string n = argv[2];
        if (argc > 3) g_dbg_skip = atoi(argv[3]);
        printf("dbg_skip=%d ", g_dbg_skip);
        if (n == "res2_conv2") bench_shape("res2_conv2_3x3_64_64_HALO", 64, 4272, 64, 64, 9, kEpiRelu | kEpiMask, sms, 0, true);
        else if (n == "res2_conv2_pair") bench_shape("res2_conv2_3x3_64_64_PAIR", 64, 4272, 64, 64, 9, kEpiRelu | kEpiMask, sms, 0, false, true);
        else if (n == "res3_conv2_pair") bench_shape("res3_conv2_3x3_128_128_PAIR", 128, 1088, 128, 128, 9, kEpiRelu | kEpiMask, sms, 0, false, true);
        else if (n == "res3_conv2") bench_shape("res3_conv2_3x3_128_128_HALO", 128, 1088, 128, 128, 9, kEpiRelu | kEpiMask, sms, 0, true);
        else if (n == "res2_conv1") bench_shape("res2_conv1_1x1_256_64_STAGED", 64, 4272, 256, 64, 1, kEpiRelu | kEpiMask, sms, 1);
        else if (n == "res2_conv3") bench_shape("res2_conv3_1x1_64_256_res_STAGED_2x2", 256, 4272, 64, 256, 1, kEpiResidual | kEpiRelu | kEpiMask, sms, 1);
        else if (n == "res4_conv3") bench_shape("res4_conv3_1x1_256_1024_res_STAGED_2x2", 256, 280, 256, 1024, 1, kEpiResidual | kEpiRelu | kEpiMask, sms, 1);
        else printf("unknown shape %s\n", n.c_str());
        return 0;
    }
    if (argc > 1 && std::string(argv[1]) == "pairnarrow") {
        // 3x3 convolutions at the 33-image sizes: single-CTA halo pipeline against the CTA-pair kernel
        const int F = kEpiRelu | kEpiMask;
        bench_shape("res2_conv2_3x3_64_64_HALO_BRES", 64, 17622, 64, 64, 9, F, sms, 0, true);
        bench_shape("res2_conv2_3x3_64_64_PAIR", 64, 17622, 64, 64, 9, F, sms, 0, false, true);
        bench_shape("res3_conv2_3x3_128_128_HALO", 128, 4488, 128, 128, 9, F, sms, 0, true);
        bench_shape("res3_conv2_3x3_128_128_PAIR", 128, 4488, 128, 128, 9, F, sms, 0, false, true);
        bench_shape("res4_conv2_3x3_256_256_PAIR", 256, 1155, 256, 256, 9, F, sms, 0, false, true);
        bench_shape("res5_conv2_3x3_512_512_PAIR", 256, 330, 512, 512, 9, F, sms, 0, false, true);
        bench_shape("tower3x3_256_gn_f32out_PAIR", 256, 1480, 256, 256, 9, kEpiMask | kEpiGnStats | kEpiOutF32, sms, 0, false, true);
        return 0;
    }
    if (argc > 1 && std::string(argv[1]) == "pair1x1") {
        // single-CTA staged kernels against the CTA-pair 1x1 kernel on the res3..res5 shapes of a 33-image trunk pass
        struct Sh { const char* name; int m_tiles, cin, cout, flags; };
        const int RES = kEpiResidual | kEpiRelu | kEpiMask, C1 = kEpiRelu | kEpiMask, SC = kEpiMask;
        const Sh shapes[] = {{"res3_conv3_128_512", 4488, 128, 512, RES},   {"res4_conv3_256_1024", 1155, 256, 1024, RES},
                             {"res5_conv3_512_2048", 330, 512, 2048, RES},  {"res4_conv1_1024_256", 1155, 1024, 256, C1},
                             {"res5_conv1_2048_512", 330, 2048, 512, C1},   {"res3_shortcut_256_512", 4488, 256, 512, SC},
                             {"res4_shortcut_512_1024", 1155, 512, 1024, SC}, {"res5_shortcut_1024_2048", 330, 1024, 2048, SC},
                             {"res3_conv1_512_128(bn256)", 4488, 512, 256, C1}};
        for (const Sh& sh : shapes) {
            std::string a = std::string(sh.name) + "_SINGLE", b = std::string(sh.name) + "_PAIR";
            bench_shape(a.c_str(), 256, sh.m_tiles, sh.cin, sh.cout, 1, sh.flags, sms, 0 + 100);   // 100 -> variant 0 (auto)
            bench_shape(b.c_str(), 256, sh.m_tiles, sh.cin, sh.cout, 1, sh.flags, sms, 9);
        }
        return 0;
    }
    if (argc > 1 && std::string(argv[1]) == "bskip") {
        // A/B of the weight-tile reuse (dbg_skip bit 16 = always reload) on the bottleneck 1x1 shapes of one 33-image
        // 800x1344 trunk pass (25 support + 8 query images of the headline episode)
        for (int skip : {16, 0}) {
            g_dbg_skip = skip;
            printf("---- weight tiles %s\n", skip ? "reloaded for every output tile" : "kept in their ring slot");
            bench_shape("res2_conv3_1x1_64_256_res_STAGED_2x2", 256, 17622, 64, 256, 1, kEpiResidual | kEpiRelu | kEpiMask, sms, 1);
            bench_shape("res3_conv3_1x1_128_512_res_STAGED_2x2", 256, 4488, 128, 512, 1, kEpiResidual | kEpiRelu | kEpiMask, sms, 1);
            bench_shape("res3_conv3_1x1_128_512_res_STAGED_bn128", 128, 4488, 128, 512, 1, kEpiResidual | kEpiRelu | kEpiMask, sms, 1);
            bench_shape("res4_conv3_1x1_256_1024_res_STAGED_2x2", 256, 1155, 256, 1024, 1, kEpiResidual | kEpiRelu | kEpiMask, sms, 1);
            bench_shape("res4_conv3_1x1_256_1024_res_STAGED_bn128", 128, 1155, 256, 1024, 1, kEpiResidual | kEpiRelu | kEpiMask, sms, 1);
            bench_shape("res5_conv3_1x1_512_2048_res_STAGED_2x2", 256, 330, 512, 2048, 1, kEpiResidual | kEpiRelu | kEpiMask, sms, 1);
            bench_shape("res2_conv1_1x1_256_64_STAGED_6slots", 64, 17622, 256, 64, 1, kEpiRelu | kEpiMask, sms, 1);
            bench_shape("res2_conv1_1x1_256_64_STAGED_8slots", 64, 17622, 256, 64, 1, kEpiRelu | kEpiMask, sms, 3);
            bench_shape("res3_conv1_1x1_512_128_STAGED", 128, 4488, 512, 128, 1, kEpiRelu | kEpiMask, sms, 1);
            bench_shape("res3_shortcut_1x1_256_512_STAGED_2x2", 256, 4488, 256, 512, 1, kEpiMask, sms, 1);
            bench_shape("res3_shortcut_1x1_256_512_STAGED_bn128", 128, 4488, 256, 512, 1, kEpiMask, sms, 1);
        }
        return 0;
    }
    {   // host fp32 -> fp16 conversion against the CUDA reference conversion
        int bad = 0;
        for (int i = 0; i < 2000000; ++i) {
            float f = frand() * ((i % 7 == 0) ? 70000.f : (i % 5 == 0) ? 1e-5f : (i % 3 == 0) ? 3e-8f : 4.f);
            uint16_t a = float_to_half_bits(f);
            __half hh = __float2half_rn(fminf(fmaxf(f, -kHalfMax), kHalfMax));
            uint16_t b = *reinterpret_cast<uint16_t*>(&hh);
            if (a != b && bad++ < 5) printf("  half conversion mismatch %g: %04x vs %04x\n", f, a, b);
        }
        printf("[half_conversion] %s\n", bad ? "FAIL" : "PASS");
        fails += bad ? 1 : 0;
    }
    const std::vector<int> z1 = {0};
    std::vector<int> dy9, dx9, dy4, dx4;
    for (int t = 0; t < 9; ++t) { dy9.push_back(t / 3 - 1); dx9.push_back(t % 3 - 1); }
    for (int t = 0; t < 4; ++t) { dy4.push_back(t - 2); dx4.push_back(-2); }
    {
        Case c{"gemm_bn256_k64", 256, {mk_seg(0, 1, 638, 1)}, 640 * 3, 64, 64, 256, 1, 1, z1, z1, 0, true};
        c.segs[0].nrows = c.total_rows;
        fails += run_case(c, sms);
    }
    for (int skew = 1; skew <= 3; ++skew)
        for (int mode = 0; mode < 2; ++mode) {   // does a descriptor that starts `skew` rows into a swizzled tile work?
            Case c{mode ? "EXPERIMENT_skew_base_offset=skew" : "EXPERIMENT_skew_base_offset=0", 256, {mk_seg(0, 1, 638, 1)}, 640, 128, 128, 256, 1, 2, z1, z1, kEpiOutF32, true};
            c.segs[0].nrows = c.total_rows;
            c.skew = skew;
            c.base_offset = mode ? skew : 0;
            printf("skew=%d ", skew);
            run_case(c, sms);
        }
    {
        Case c{"gemm_bn256_k256_persistent_f32out", 256, {mk_seg(0, 1, 638, 1)}, 128 * 333, 256, 256, 256, 1, 4, z1, z1, kEpiRelu | kEpiOutF32, true};
        c.segs[0].nrows = c.total_rows;
        fails += run_case(c, sms);
    }
    {   // 3x3 conv over two planes of DIFFERENT padded width, mask + GN stats, fp32 raw output
        Seg s0 = mk_seg(0, 13, 21, 1);
        Seg s1 = mk_seg(round128(s0.nrows), 7, 11, 1);
        int total = s1.row0 + round128(s1.nrows);
        Case c{"conv3x3_bn256_mask_gn_f32out", 256, {s0, s1}, total, 64, 64, 256, 9, 1, dy9, dx9, kEpiMask | kEpiGnStats | kEpiOutF32, true};
        fails += run_case(c, sms);
    }
    {
        Case c{"gemm_bn64_res_relu", 64, {mk_seg(0, 1, 638, 1)}, 128 * 7, 64, 64, 64, 1, 1, z1, z1, kEpiResidual | kEpiRelu, true};
        c.segs[0].nrows = c.total_rows;
        fails += run_case(c, sms);
    }
    {
        Case c{"gemm_bn128_n256_res", 128, {mk_seg(0, 1, 638, 1)}, 128 * 9, 128, 128, 256, 1, 2, z1, z1, kEpiRelu | kEpiResidual, false};
        c.segs[0].nrows = c.total_rows;
        fails += run_case(c, sms);
    }
    {
        Case c{"gemm_bn256_n1024_res", 256, {mk_seg(0, 1, 638, 1)}, 128 * 3, 128, 128, 1024, 1, 2, z1, z1, kEpiRelu | kEpiResidual, true};
        c.segs[0].nrows = c.total_rows;
        fails += run_case(c, sms);
    }
    {   // staged epilogue, in place, 4 n-tiles, many tiles per CTA (buffer/phase cycling), masked borders
        Seg s0 = mk_seg(0, 150, 168, 1);
        Case c{"staged_inplace_res_relu_mask_n1024", 256, {s0}, round128(s0.nrows), 128, 128, 1024, 1, 2, z1, z1, kEpiRelu | kEpiResidual | kEpiMask, true};
        c.staged = true;
        fails += run_case(c, sms);
    }
    {   // staged epilogue without residual (store-only staging)
        Case c{"staged_noresidual_n256", 256, {mk_seg(0, 1, 638, 1)}, 128 * 301, 64, 64, 256, 1, 1, z1, z1, kEpiRelu, true};
        c.segs[0].nrows = c.total_rows;
        c.staged = true;
        fails += run_case(c, sms);
    }
    {
        Seg s0 = mk_seg(0, 13, 21, 1);
        Case c{"conv3x3_bn16_f32out", 16, {s0}, round128(s0.nrows), 256, 256, 16, 9, 4, dy9, dx9, kEpiMask | kEpiOutF32, true};
        fails += run_case(c, sms);
    }
    for (int bnv : {256, 128, 64, 16}) {   // halo pipeline, every BN, two planes of different width, many tiles per CTA
        Seg s0 = mk_seg(0, 140, 168, 1);
        Seg s1 = mk_seg(round128(s0.nrows), 37, 41, 1);
        int total = s1.row0 + round128(s1.nrows);
        const int cout = bnv == 16 ? 16 : (bnv == 256 ? 512 : bnv);
        Case c{"HALO_conv3x3", bnv, {s0, s1}, total, 128, 128, cout, 9, 2, dy9, dx9,
               kEpiMask | (bnv == 256 ? 0 : kEpiRelu) | (bnv == 16 ? kEpiOutF32 : 0), true};
        c.halo = true;
        printf("bn=%d ", bnv);
        fails += run_case(c, sms);
    }
    {   // halo pipeline with RESIDENT weights (64 -> 64 channels, res2 conv2), two planes, many tiles per CTA
        Seg s0 = mk_seg(0, 140, 168, 1);
        Seg s1 = mk_seg(round128(s0.nrows), 37, 41, 1);
        Case c{"HALO_BRES_conv3x3_64_64", 64, {s0, s1}, s1.row0 + round128(s1.nrows), 64, 64, 64, 9, 1, dy9, dx9, kEpiMask | kEpiRelu, true};
        c.halo = true;
        fails += run_case(c, sms);
    }
    {   // CTA pair, ODD number of M tiles (phantom tile), two planes, two N tiles, masked, fp16 out
        Seg s0 = mk_seg(0, 140, 168, 1);
        Seg s1 = mk_seg(round128(s0.nrows), 37, 41, 1);
        int total = s1.row0 + round128(s1.nrows);
        if ((total / 128) % 2 == 0) total += 128;
        Case c{"PAIR_conv3x3_n512_mask_relu", 256, {s0, s1}, total, 128, 128, 512, 9, 2, dy9, dx9, kEpiMask | kEpiRelu, true};
        c.pair = true;
        fails += run_case(c, sms);
    }
    for (int bnv : {64, 128}) {   // narrow CTA-pair tiles (res2 / res3 conv2): odd tile count, two planes, many tiles per pair
        Seg s0 = mk_seg(0, 140, 168, 1);
        Seg s1 = mk_seg(round128(s0.nrows), 37, 41, 1);
        int total = s1.row0 + round128(s1.nrows);
        if ((total / 128) % 2 == 0) total += 128;
        Case c{"PAIR_conv3x3_narrow_mask_relu", bnv, {s0, s1}, total, bnv, bnv, bnv, 9, bnv / 64, dy9, dx9, kEpiMask | kEpiRelu, true};
        c.pair = true;
        printf("bn=%d ", bnv);
        fails += run_case(c, sms);
    }
    {   // CTA pair + GroupNorm statistics + fp32 output (the tower configuration)
        Seg s0 = mk_seg(0, 13, 21, 1);
        Seg s1 = mk_seg(round128(s0.nrows), 7, 11, 1);
        int total = s1.row0 + round128(s1.nrows);
        Case c{"PAIR_conv3x3_bn256_mask_gn_f32out", 256, {s0, s1}, total, 256, 256, 256, 9, 4, dy9, dx9, kEpiMask | kEpiGnStats | kEpiOutF32, true};
        c.pair = true;
        fails += run_case(c, sms);
    }
    {   // halo + GroupNorm statistics + fp32 output (the tower configuration)
        Seg s0 = mk_seg(0, 13, 21, 1);
        Seg s1 = mk_seg(round128(s0.nrows), 7, 11, 1);
        int total = s1.row0 + round128(s1.nrows);
        Case c{"HALO_conv3x3_bn256_mask_gn_f32out", 256, {s0, s1}, total, 256, 256, 256, 9, 4, dy9, dx9, kEpiMask | kEpiGnStats | kEpiOutF32, true};
        c.halo = true;
        fails += run_case(c, sms);
    }
    {   // stem trick: 16-channel pixels, rows overlapped (pitch 16 halves, 64 visible), 4 vertical taps x K=64
        Seg s0 = mk_seg(0, 10, 12, 2);
        Case c{"stem_overlapped_rows_bn64", 64, {s0}, round128(s0.nrows), 64, 16, 64, 4, 1, dy4, dx4, kEpiMask | kEpiRelu, true};
        fails += run_case(c, sms);
    }
    {   // staged epilogue on the stem layout (BN = 64, overlapped rows, no residual)
        Seg s0 = mk_seg(0, 60, 70, 2);
        Case c{"STAGED_stem_overlapped_rows_bn64", 64, {s0}, round128(s0.nrows), 64, 16, 64, 4, 1, dy4, dx4, kEpiMask | kEpiRelu, true};
        c.staged = true;
        fails += run_case(c, sms);
    }
    {   // stem as 16 K = 16 taps over a 32-byte-swizzled [rows][16] map; many tiles per CTA, two planes
        Seg s0 = mk_seg(0, 150, 170, 2);
        Seg s1 = mk_seg(round128(s0.nrows), 37, 45, 2);
        Case c{"STEM16_sw32_k16_taps_bn64", 64, {s0, s1}, s1.row0 + round128(s1.nrows), 64, 16, 64, 4, 1, dy4, dx4, kEpiMask | kEpiRelu, true};
        c.staged = true;
        c.stem16 = true;
        fails += run_case(c, sms);
    }
    for (int bnv : {64, 128}) {   // staged epilogue for the narrow 1x1 convolutions (bottleneck conv1), many tiles per CTA
        Seg s0 = mk_seg(0, 150, 168, 1);
        Case c{"STAGED_conv1_1x1_relu_mask", bnv, {s0}, round128(s0.nrows), 256, 256, bnv, 1, 4, z1, z1, kEpiRelu | kEpiMask, true};
        c.staged = true;
        printf("bn=%d ", bnv);
        fails += run_case(c, sms);
    }
    {   // CTA-pair 1x1 kernel, staged epilogue in place: residual + ReLU + mask, 4 N tiles, ODD number of M tiles
        // (phantom tile in the last pair), many pair tiles per cluster (ring / staging-buffer phase cycling), K = 256
        Seg s0 = mk_seg(0, 30, 40, 1);   // 1344 rows = 11 M tiles (odd) -> 6 pair tiles x 4 N tiles on 8 clusters: 3 per cluster
        Case c{"PAIR1x1_staged_inplace_res_relu_mask_n1024", 256, {s0}, round128(s0.nrows), 256, 256, 1024, 1, 4, z1, z1, kEpiRelu | kEpiResidual | kEpiMask, true};
        c.staged = true;
        c.pair1x1 = true;
        c.sms_override = 16;
        fails += run_case(c, sms);
    }
    {   // same kernel without residual (conv1 / shortcut use), one N tile, K = 512, even M tiles, two planes
        Seg s0 = mk_seg(0, 30, 40, 1);
        Seg s1 = mk_seg(round128(s0.nrows), 13, 21, 1);
        Case c{"PAIR1x1_staged_noresidual_n256_k512", 256, {s0, s1}, s1.row0 + round128(s1.nrows), 512, 512, 256, 1, 8, z1, z1, kEpiRelu | kEpiMask, true};
        c.staged = true;
        c.pair1x1 = true;
        fails += run_case(c, sms);
    }
    {   // a single M tile (one real + one phantom tile in the only pair), 2 N tiles
        Seg s0 = mk_seg(0, 8, 10, 1);
        Case c{"PAIR1x1_single_tile_n512", 256, {s0}, round128(s0.nrows), 256, 256, 512, 1, 4, z1, z1, kEpiResidual | kEpiMask, true};
        c.staged = true;
        c.pair1x1 = true;
        fails += run_case(c, sms);
    }
    // ---------------------------------------------------------------- split-operand ("exact") mode of every kernel family
    {   // generic pipeline, direct epilogue: fp16 hi | lo output + residual, and fp32 output
        Case c{"SPLIT_gemm_bn256_n1024_res", 256, {mk_seg(0, 1, 638, 1)}, 128 * 3, 128, 128, 1024, 1, 2, z1, z1, kEpiRelu | kEpiResidual, true};
        c.segs[0].nrows = c.total_rows;
        c.split = true;
        fails += run_case(c, sms);
        Case d{"SPLIT_gemm_bn64_k256_f32out", 64, {mk_seg(0, 1, 638, 1)}, 128 * 37, 256, 256, 64, 1, 4, z1, z1, kEpiOutF32, true};
        d.segs[0].nrows = d.total_rows;
        d.split = true;
        fails += run_case(d, sms);
    }
    for (int bnv : {64, 128}) {   // staged epilogue, in place, residual, many tiles per CTA, several N tiles
        Seg s0 = mk_seg(0, 150, 168, 1);
        Case c{"SPLIT_staged_inplace_res_relu_mask", bnv, {s0}, round128(s0.nrows), 128, 128, bnv * 4, 1, 2, z1, z1, kEpiRelu | kEpiResidual | kEpiMask, true};
        c.staged = true;
        c.split = true;
        printf("bn=%d ", bnv);
        fails += run_case(c, sms);
        Case d{"SPLIT_staged_noresidual_k256", bnv, {s0}, round128(s0.nrows), 256, 256, bnv, 1, 4, z1, z1, kEpiRelu | kEpiMask, true};
        d.staged = true;
        d.split = true;
        printf("bn=%d ", bnv);
        fails += run_case(d, sms);
    }
    {   // 256-wide staged split tile (one staging buffer), residual in place and store-only
        Seg s0 = mk_seg(0, 150, 168, 1);
        Case c{"SPLIT_staged_bn256_inplace_res_relu_mask", 256, {s0}, round128(s0.nrows), 128, 128, 1024, 1, 2, z1, z1, kEpiRelu | kEpiResidual | kEpiMask, true};
        c.staged = true;
        c.split = true;
        fails += run_case(c, sms);
        Case d{"SPLIT_staged_bn256_noresidual_k512", 256, {s0}, round128(s0.nrows), 512, 512, 512, 1, 8, z1, z1, kEpiMask, true};
        d.staged = true;
        d.split = true;
        fails += run_case(d, sms);
    }
    for (int mode : {1, 2}) {   // split CTA-pair 1x1 kernel, chunked staged epilogue: K' = 3C loop (1) and quad stages (2)
        {   // residual in place + ReLU + mask, 4 N tiles, ODD number of M tiles (phantom tile), many pair tiles per cluster
            Seg s0 = mk_seg(0, 30, 40, 1);   // 11 M tiles
            Case c{"SPLIT_PAIR1x1_inplace_res_relu_mask_n1024", 256, {s0}, round128(s0.nrows), 256, 256, 1024, 1, 4, z1, z1, kEpiRelu | kEpiResidual | kEpiMask, true};
            c.staged = true; c.split = true; c.pairsplit = mode; c.sms_override = 16;
            printf("mode=%d ", mode);
            fails += run_case(c, sms);
        }
        {   // no residual (conv1 / shortcut use), one N tile, K = 512, two planes, all SMs
            Seg s0 = mk_seg(0, 150, 168, 1);
            Seg s1 = mk_seg(round128(s0.nrows), 13, 21, 1);
            Case c{"SPLIT_PAIR1x1_noresidual_n256_k512", 256, {s0, s1}, s1.row0 + round128(s1.nrows), 512, 512, 256, 1, 8, z1, z1, kEpiRelu | kEpiMask, true};
            c.staged = true; c.split = true; c.pairsplit = mode;
            printf("mode=%d ", mode);
            fails += run_case(c, sms);
        }
        {   // a single M tile (one real + one phantom tile in the only pair), 2 N tiles, residual
            Seg s0 = mk_seg(0, 8, 10, 1);
            Case c{"SPLIT_PAIR1x1_single_tile_n512", 256, {s0}, round128(s0.nrows), 256, 256, 512, 1, 4, z1, z1, kEpiResidual | kEpiMask, true};
            c.staged = true; c.split = true; c.pairsplit = mode;
            printf("mode=%d ", mode);
            fails += run_case(c, sms);
        }
        {   // many tiles per cluster without residual (buffer hand-back through epi_free), 2 N tiles
            Seg s0 = mk_seg(0, 150, 168, 1);
            Case c{"SPLIT_PAIR1x1_noresidual_n512_manytiles", 256, {s0}, round128(s0.nrows), 128, 128, 512, 1, 2, z1, z1, kEpiMask, true};
            c.staged = true; c.split = true; c.pairsplit = mode; c.sms_override = 8;
            printf("mode=%d ", mode);
            fails += run_case(c, sms);
        }
    }
    for (int bnv : {128, 256}) {   // quad-stage 1x1 kernels: staged in place with residual (many tiles per CTA), store-only, register epilogue
        Seg s0 = mk_seg(0, 150, 168, 1);
        Case c{"SPLIT_QS_staged_inplace_res_relu_mask_k256", bnv, {s0}, round128(s0.nrows), 256, 256, 1024, 1, 4, z1, z1, kEpiRelu | kEpiResidual | kEpiMask, true};
        c.staged = true; c.split = true; c.qs = true;
        printf("bn=%d ", bnv);
        fails += run_case(c, sms);
        Case d{"SPLIT_QS_staged_noresidual_k512", bnv, {s0}, round128(s0.nrows), 512, 512, 512, 1, 8, z1, z1, kEpiMask, true};
        d.staged = true; d.split = true; d.qs = true;
        printf("bn=%d ", bnv);
        fails += run_case(d, sms);
        Case e{"SPLIT_QS_direct_relu_mask_k1024", bnv, {s0}, round128(s0.nrows), 1024, 1024, 256, 1, 16, z1, z1, kEpiRelu | kEpiMask, true};
        e.split = true; e.qs = true;
        printf("bn=%d ", bnv);
        fails += run_case(e, sms);
        Case f{"SPLIT_QS_direct_f32out_k128", bnv, {mk_seg(0, 1, 638, 1)}, 128 * 5, 128, 128, 256, 1, 2, z1, z1, kEpiOutF32, true};
        f.segs[0].nrows = f.total_rows;
        f.split = true; f.qs = true;
        printf("bn=%d ", bnv);
        fails += run_case(f, sms);
    }
    for (int bnv : {256, 128, 64, 16}) {   // halo pipeline
        Seg s0 = mk_seg(0, 70, 84, 1);
        Seg s1 = mk_seg(round128(s0.nrows), 37, 41, 1);
        int total = s1.row0 + round128(s1.nrows);
        const int cout = bnv == 16 ? 16 : (bnv == 256 ? 512 : bnv);
        Case c{"SPLIT_HALO_conv3x3", bnv, {s0, s1}, total, 128, 128, cout, 9, 2, dy9, dx9,
               kEpiMask | (bnv == 256 ? 0 : kEpiRelu) | (bnv == 16 ? kEpiOutF32 : 0), true};
        c.halo = true;
        c.split = true;
        printf("bn=%d ", bnv);
        fails += run_case(c, sms);
    }
    for (int bnv : {128, 64, 16}) {   // N-merged split halo pipeline (two instructions per k-step), ring-streamed weights
        Seg s0 = mk_seg(0, 70, 84, 1);
        Seg s1 = mk_seg(round128(s0.nrows), 37, 41, 1);
        Case c{"SPLIT_NM_HALO_conv3x3", bnv, {s0, s1}, s1.row0 + round128(s1.nrows), 128, 128, bnv, 9, 2, dy9, dx9,
               kEpiMask | kEpiRelu | (bnv == 16 ? kEpiOutF32 : 0), true};
        c.halo = true;
        c.split = true;
        c.nm = true;
        printf("bn=%d ", bnv);
        fails += run_case(c, sms);
    }
    {   // N-merged with the nine resident 2 x 64-row weight tiles (64 -> 64 channels, res2 conv2)
        Seg s0 = mk_seg(0, 140, 168, 1);
        Seg s1 = mk_seg(round128(s0.nrows), 37, 41, 1);
        Case c{"SPLIT_NM_HALO_BRES_conv3x3_64_64", 64, {s0, s1}, s1.row0 + round128(s1.nrows), 64, 64, 64, 9, 1, dy9, dx9, kEpiMask | kEpiRelu, true};
        c.halo = true;
        c.split = true;
        c.nm = true;
        fails += run_case(c, sms);
    }
    {   // split halo pipeline with the 18 resident weight tiles (64 -> 64 channels, res2 conv2), two planes, many tiles per CTA
        Seg s0 = mk_seg(0, 140, 168, 1);
        Seg s1 = mk_seg(round128(s0.nrows), 37, 41, 1);
        Case c{"SPLIT_HALO_BRES_conv3x3_64_64", 64, {s0, s1}, s1.row0 + round128(s1.nrows), 64, 64, 64, 9, 1, dy9, dx9, kEpiMask | kEpiRelu, true};
        c.halo = true;
        c.split = true;
        fails += run_case(c, sms);
    }
    {   // CTA pair: fp16 hi | lo output (odd tile count) and the tower configuration (GN statistics, fp32 output)
        Seg s0 = mk_seg(0, 70, 84, 1);
        Seg s1 = mk_seg(round128(s0.nrows), 37, 41, 1);
        int total = s1.row0 + round128(s1.nrows);
        if ((total / 128) % 2 == 0) total += 128;
        Case c{"SPLIT_PAIR_conv3x3_n512_mask_relu", 256, {s0, s1}, total, 128, 128, 512, 9, 2, dy9, dx9, kEpiMask | kEpiRelu, true};
        c.pair = true;
        c.split = true;
        fails += run_case(c, sms);
        Seg t0 = mk_seg(0, 13, 21, 1);
        Seg t1 = mk_seg(round128(t0.nrows), 7, 11, 1);
        Case d{"SPLIT_PAIR_conv3x3_bn256_mask_gn_f32out", 256, {t0, t1}, t1.row0 + round128(t1.nrows), 256, 256, 256, 9, 4, dy9, dx9, kEpiMask | kEpiGnStats | kEpiOutF32, true};
        d.pair = true;
        d.split = true;
        fails += run_case(d, sms);
    }
    {   // stem: [16 hi | 16 lo] rows, eight resident weight tiles, three A boxes per vertical tap
        Seg s0 = mk_seg(0, 150, 170, 2);
        Seg s1 = mk_seg(round128(s0.nrows), 37, 45, 2);
        Case c{"SPLIT_STEM16", 64, {s0, s1}, s1.row0 + round128(s1.nrows), 64, 16, 64, 4, 1, dy4, dx4, kEpiMask | kEpiRelu, true};
        c.staged = true;
        c.stem16 = true;
        c.split = true;
        fails += run_case(c, sms);
        Case d = c;   // N-merged: four resident [w_hi ; w_lo] tiles, two A boxes per vertical tap
        d.name = "SPLIT_NM_STEM16";
        d.nm = true;
        fails += run_case(d, sms);
    }
    printf("correctness: %d failing case(s)\n", fails);
    if (argc > 1 && std::string(argv[1]) == "split") {
        // fast against split-operand mode on the shapes of a 33-image trunk pass and an 8-image head pass
        const int RES = kEpiResidual | kEpiRelu | kEpiMask, C1 = kEpiRelu | kEpiMask, SC = kEpiMask;
        for (int sp = 0; sp < 2; ++sp) {
            printf("---- %s\n", sp ? "split fp16 x3 (exact)" : "single fp16 (fast)");
            const bool b = sp != 0;
            bench_shape("res2_conv1_1x1_256_64_STAGED", 64, 17622, 256, 64, 1, C1, sms, 100, false, false, b);
            bench_shape("res2_conv2_3x3_64_64_HALO", 64, 17622, 64, 64, 9, C1, sms, 0, true, false, b);
            bench_shape("res2_conv3_1x1_64_256_res_STAGED", b ? 128 : 256, 17622, 64, 256, 1, RES, sms, 100, false, false, b);
            bench_shape("res3_conv1_1x1_512_128_STAGED", 128, 4488, 512, 128, 1, C1, sms, 100, false, false, b);
            bench_shape("res3_conv2_3x3_128_128_HALO", 128, 4488, 128, 128, 9, C1, sms, 0, true, false, b);
            bench_shape("res3_conv3_1x1_128_512_res_STAGED", b ? 128 : 256, 4488, 128, 512, 1, RES, sms, 100, false, false, b);
            bench_shape("res3_shortcut_1x1_256_512_STAGED", b ? 128 : 256, 4488, 256, 512, 1, SC, sms, 100, false, false, b);
            bench_shape("res4_conv1_1x1_1024_256_STAGED", b ? 128 : 256, 1155, 1024, 256, 1, C1, sms, 100, false, false, b);
            bench_shape("res4_conv2_3x3_256_256_PAIR", 256, 1155, 256, 256, 9, C1, sms, 0, false, true, b);
            bench_shape("res4_conv3_1x1_256_1024_res_STAGED", b ? 128 : 256, 1155, 256, 1024, 1, RES, sms, 100, false, false, b);
            bench_shape("res5_conv2_3x3_512_512_PAIR", 256, 330, 512, 512, 9, C1, sms, 0, false, true, b);
            bench_shape("res5_conv3_1x1_512_2048_res_STAGED", b ? 128 : 256, 330, 512, 2048, 1, RES, sms, 100, false, false, b);
            bench_shape("tower3x3_256_gn_f32out_PAIR", 256, 1480, 256, 256, 9, kEpiMask | kEpiGnStats | kEpiOutF32, sms, 0, false, true, b);
            bench_shape("fpn_lateral3_1x1_512_256", 256, 4488, 512, 256, 1, SC, sms, 0, false, false, b);
        }
        printf("---- split, narrow 3x3 layers: three instructions per k-step against the N-merged two\n");
        bench_shape("res2_conv2_3x3_64_64_HALO_BRES", 64, 17622, 64, 64, 9, C1, sms, 0, true, false, true);
        bench_shape("res2_conv2_3x3_64_64_NM_BRES", 64, 17622, 64, 64, 9, C1, sms, 0, true, false, true, true);
        bench_shape("res3_conv2_3x3_128_128_HALO", 128, 4488, 128, 128, 9, C1, sms, 0, true, false, true);
        bench_shape("res3_conv2_3x3_128_128_NM", 128, 4488, 128, 128, 9, C1, sms, 0, true, false, true, true);
        bench_shape("pred3x3_256_16_f32out_HALO", 16, 1480, 256, 16, 9, kEpiOutF32, sms, 0, true, false, true);
        bench_shape("pred3x3_256_16_f32out_NM", 16, 1480, 256, 16, 9, kEpiOutF32, sms, 0, true, false, true, true);
        printf("---- split, staged 1x1 layers: 128-wide tiles (2 staging buffers) against 256-wide tiles (1 staging buffer)\n");
        struct Sh { const char* name; int m_tiles, cin, cout, flags; };
        const Sh shapes[] = {{"res3_shortcut_256_512", 4488, 256, 512, SC},     {"res4_shortcut_512_1024", 1155, 512, 1024, SC},
                             {"res5_shortcut_1024_2048", 330, 1024, 2048, SC},  {"res4_conv1_1024_256", 1155, 1024, 256, C1},
                             {"res5_conv1_2048_512", 330, 2048, 512, C1},       {"res2_conv3_64_256", 17622, 64, 256, RES},
                             {"res3_conv3_128_512", 4488, 128, 512, RES},       {"res4_conv3_256_1024", 1155, 256, 1024, RES},
                             {"res5_conv3_512_2048", 330, 512, 2048, RES}};
        for (const Sh& sh : shapes) {
            std::string a = std::string(sh.name) + "_bn128", c = std::string(sh.name) + "_bn256", d = std::string(sh.name) + "_DIRECT_bn256";
            bench_shape(a.c_str(), 128, sh.m_tiles, sh.cin, sh.cout, 1, sh.flags, sms, 100, false, false, true);
            bench_shape(c.c_str(), 256, sh.m_tiles, sh.cin, sh.cout, 1, sh.flags, sms, 100, false, false, true);
            bench_shape(d.c_str(), 256, sh.m_tiles, sh.cin, sh.cout, 1, sh.flags, sms, 0, false, false, true);   // register epilogue
        }
        return 0;
    }
    if (argc > 1 && std::string(argv[1]) == "qs") {
        // split mode, deep 1x1 layers at the 33-image shapes: K' = 3C loop (six tiles per k-block) against quad stages (four)
        const int RES = kEpiResidual | kEpiRelu | kEpiMask, C1 = kEpiRelu | kEpiMask, SC = kEpiMask;
        struct Sh { const char* name; int m_tiles, cin, cout, flags; };
        const Sh shapes[] = {{"res3_conv3_128_512", 4488, 128, 512, RES},       {"res4_conv3_256_1024", 1155, 256, 1024, RES},
                             {"res5_conv3_512_2048", 330, 512, 2048, RES},      {"res3_shortcut_256_512", 4488, 256, 512, SC},
                             {"res4_shortcut_512_1024", 1155, 512, 1024, SC},   {"res5_shortcut_1024_2048", 330, 1024, 2048, SC},
                             {"res3_conv1_512_128", 4488, 512, 128, C1},        {"res4_conv1_1024_256", 1155, 1024, 256, C1},
                             {"res5_conv1_2048_512", 330, 2048, 512, C1},       {"fpn_lateral3_512_256", 4488, 512, 256, SC},
                             {"fpn_lateral4_1024_256", 1155, 1024, 256, SC},    {"fpn_lateral5_2048_256", 330, 2048, 256, SC}};
        for (const Sh& sh : shapes) {
            const std::string n = sh.name;
            const bool res = (sh.flags & kEpiResidual) != 0;
            if (sh.cout >= 256) {
                bench_shape((n + "_K3_staged_bn128").c_str(), 128, sh.m_tiles, sh.cin, sh.cout, 1, sh.flags, sms, 100, false, false, true);
                if (!res) bench_shape((n + "_K3_staged_bn256").c_str(), 256, sh.m_tiles, sh.cin, sh.cout, 1, sh.flags, sms, 100, false, false, true);
                if (!res) bench_shape((n + "_K3_direct_bn256").c_str(), 256, sh.m_tiles, sh.cin, sh.cout, 1, sh.flags, sms, 0, false, false, true);
                bench_shape((n + "_QS_staged_bn256").c_str(), 256, sh.m_tiles, sh.cin, sh.cout, 1, sh.flags, sms, 100, false, false, true, false, true);
                bench_shape((n + "_QS_direct_bn256").c_str(), 256, sh.m_tiles, sh.cin, sh.cout, 1, sh.flags, sms, 0, false, false, true, false, true);
            } else {
                bench_shape((n + "_K3_staged_bn128").c_str(), 128, sh.m_tiles, sh.cin, sh.cout, 1, sh.flags, sms, 100, false, false, true);
            }
            if (sh.cout % 256 == 0) {
                bench_shape((n + "_PAIR_K3").c_str(), 256, sh.m_tiles, sh.cin, sh.cout, 1, sh.flags, sms, 10, false, false, true);
                bench_shape((n + "_PAIR_QS").c_str(), 256, sh.m_tiles, sh.cin, sh.cout, 1, sh.flags, sms, 11, false, false, true);
            }
            bench_shape((n + "_QS_staged_bn128").c_str(), 128, sh.m_tiles, sh.cin, sh.cout, 1, sh.flags, sms, 100, false, false, true, false, true);
            bench_shape((n + "_QS_direct_bn128").c_str(), 128, sh.m_tiles, sh.cin, sh.cout, 1, sh.flags, sms, 0, false, false, true, false, true);
        }
        return fails ? 1 : 0;
    }
    if (argc > 1 && std::string(argv[1]) == "bench") {
        bench_shape("tower3x3_256_gn_f32out", 256, 1480, 256, 256, 9, kEpiMask | kEpiGnStats | kEpiOutF32, sms);
        bench_shape("tower3x3_256_plain", 256, 1480, 256, 256, 9, 0, sms);
        bench_shape("tower3x3_256_gn_f32out_HALO", 256, 1480, 256, 256, 9, kEpiMask | kEpiGnStats | kEpiOutF32, sms, 0, true);
        bench_shape("tower3x3_256_plain_HALO", 256, 1480, 256, 256, 9, 0, sms, 0, true);
        bench_shape("tower3x3_256_gn_f32out_PAIR", 256, 1480, 256, 256, 9, kEpiMask | kEpiGnStats | kEpiOutF32, sms, 0, false, true);
        bench_shape("tower3x3_256_plain_PAIR", 256, 1480, 256, 256, 9, 0, sms, 0, false, true);
        bench_shape("res5_conv2_3x3_512_512_PAIR", 256, 80, 512, 512, 9, kEpiRelu | kEpiMask, sms, 0, false, true);
        bench_shape("res2_conv2_3x3_64_64_HALO", 64, 4272, 64, 64, 9, kEpiRelu | kEpiMask, sms, 0, true);
        bench_shape("res3_conv2_3x3_128_128_HALO", 128, 1088, 128, 128, 9, kEpiRelu | kEpiMask, sms, 0, true);
        bench_shape("res5_conv2_3x3_512_512_HALO", 256, 80, 512, 512, 9, kEpiRelu | kEpiMask, sms, 0, true);
        bench_shape("pred3x3_256_16_f32out", 16, 1480, 256, 16, 9, kEpiOutF32, sms);
        bench_shape("pred3x3_256_16_f32out_HALO", 16, 1480, 256, 16, 9, kEpiOutF32, sms, 0, true);
        bench_shape("res2_conv3_1x1_64_256_res", 256, 4272, 64, 256, 1, kEpiResidual | kEpiRelu | kEpiMask, sms);
        bench_shape("res2_conv3_1x1_64_256_res_STAGED_2x2", 256, 4272, 64, 256, 1, kEpiResidual | kEpiRelu | kEpiMask, sms, 1);
        bench_shape("res3_conv3_1x1_128_512_res", 256, 1088, 128, 512, 1, kEpiResidual | kEpiRelu | kEpiMask, sms, 0);
        bench_shape("res3_conv3_1x1_128_512_res_STAGED_2x2", 256, 1088, 128, 512, 1, kEpiResidual | kEpiRelu | kEpiMask, sms, 1);
        bench_shape("res3_conv3_1x1_128_512_res_STAGED_3x1", 256, 1088, 128, 512, 1, kEpiResidual | kEpiRelu | kEpiMask, sms, 2);
        bench_shape("res4_conv3_1x1_256_1024_res_STAGED_2x2", 256, 280, 256, 1024, 1, kEpiResidual | kEpiRelu | kEpiMask, sms, 1);
        bench_shape("res4_conv3_1x1_256_1024_res_STAGED_3x1", 256, 280, 256, 1024, 1, kEpiResidual | kEpiRelu | kEpiMask, sms, 2);
        bench_shape("res5_conv3_1x1_512_2048_res", 256, 80, 512, 2048, 1, kEpiResidual | kEpiRelu | kEpiMask, sms, 0);
        bench_shape("res5_conv3_1x1_512_2048_res_STAGED_2x2", 256, 80, 512, 2048, 1, kEpiResidual | kEpiRelu | kEpiMask, sms, 1);
        bench_shape("res5_conv3_1x1_512_2048_res_STAGED_3x1", 256, 80, 512, 2048, 1, kEpiResidual | kEpiRelu | kEpiMask, sms, 2);
        bench_shape("res5_shortcut_1x1_1024_2048", 256, 80, 1024, 2048, 1, kEpiMask, sms, 0);
        bench_shape("res5_shortcut_1x1_1024_2048_STAGED_2x2", 256, 80, 1024, 2048, 1, kEpiMask, sms, 1);
        bench_shape("res5_shortcut_1x1_1024_2048_STAGED_3x1", 256, 80, 1024, 2048, 1, kEpiMask, sms, 2);
        bench_shape("res2_conv1_1x1_256_64", 64, 4272, 256, 64, 1, kEpiRelu | kEpiMask, sms);
        bench_shape("res2_conv1_1x1_256_64_STAGED", 64, 4272, 256, 64, 1, kEpiRelu | kEpiMask, sms, 1);
        bench_shape("res3_conv1_1x1_512_128", 128, 1088, 512, 128, 1, kEpiRelu | kEpiMask, sms);
        bench_shape("res3_conv1_1x1_512_128_STAGED", 128, 1088, 512, 128, 1, kEpiRelu | kEpiMask, sms, 1);
        bench_shape("res4_conv1_1x1_1024_256", 256, 280, 1024, 256, 1, kEpiRelu | kEpiMask, sms);
        bench_shape("res4_conv1_1x1_1024_256_STAGED", 256, 280, 1024, 256, 1, kEpiRelu | kEpiMask, sms, 2);
        bench_shape("stemlike_k256_n64", 64, 17072, 256, 64, 1, kEpiRelu | kEpiMask, sms);
        bench_shape("stemlike_k256_n64_STAGED", 64, 17072, 256, 64, 1, kEpiRelu | kEpiMask, sms, 1);
        bench_shape("res2_conv2_3x3_64_64", 64, 4272, 64, 64, 9, kEpiRelu | kEpiMask, sms);
        bench_shape("res3_conv2_3x3_128_128", 128, 1088, 128, 128, 9, kEpiRelu | kEpiMask, sms);
        bench_shape("res4_conv3_1x1_256_1024", 256, 280, 256, 1024, 1, kEpiResidual | kEpiRelu | kEpiMask, sms);
        bench_shape("res5_conv2_3x3_512_512", 256, 80, 512, 512, 9, kEpiRelu | kEpiMask, sms);
    }
    return fails ? 1 : 0;
}
