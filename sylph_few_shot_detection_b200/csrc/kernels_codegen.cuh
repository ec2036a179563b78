// Code-generation kernels around the tensor-core tower: multi-level ROIAlign gather, per-shot pooling + bias conv,
// K-shot mean, code normalisation.  These move kilobytes per class; they are latency/HBM-bound gather and
// warp-shuffle reduction kernels.
#pragma once
#include "kernels_misc.cuh"

namespace sylph {

struct PyramidGeom {
    PlaneGeom lv[5];
    float scale[5];  // 1 / stride
};

// FPN level of a box, the op sequence of detectron2 assign_boxes_to_levels:
// floor(4 + log2(sqrt(area) / 224 + 1e-8)) clamped to [3, 7], minus 3.  Explicit _rn intrinsics keep the compiler
// from contracting the fp32 steps.
__device__ __forceinline__ int assign_level(float x0, float y0, float x1, float y1) {
    const float area = __fmul_rn(__fsub_rn(x1, x0), __fsub_rn(y1, y0));
    const float s = sqrtf(area);
    const float t = __fadd_rn(__fdiv_rn(s, 224.f), 1e-8f);
    float l = floorf(__fadd_rn(4.f, log2f(t)));
    l = fminf(fmaxf(l, 3.f), 7.f);  // NaN (negative area) -> 3 like torch.clamp then cast
    return static_cast<int>(l) - 3;
}

// `ld` = row pitch, `lo` > 0 = offset of the lo half of a split-operand row (value = hi + lo, exact in fp32)
__device__ __forceinline__ float bilinear_tap(const __half* __restrict__ feat, const PlaneGeom& g, int n, int H, int W,
                                              float y, float x, int c, int ld, int lo) {
    if (y < -1.f || y > static_cast<float>(H) || x < -1.f || x > static_cast<float>(W)) return 0.f;
    y = fmaxf(y, 0.f);
    x = fmaxf(x, 0.f);
    int y_low = static_cast<int>(y), x_low = static_cast<int>(x), y_high, x_high;
    if (y_low >= H - 1) { y_high = y_low = H - 1; y = static_cast<float>(y_low); } else { y_high = y_low + 1; }
    if (x_low >= W - 1) { x_high = x_low = W - 1; x = static_cast<float>(x_low); } else { x_high = x_low + 1; }
    const float ly = y - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
    auto tap = [&](int yy, int xx) {
        const __half* q = feat + plane_row(g, n, yy, xx) * ld + c;
        float v = __half2float(__ldg(q));
        if (lo) v += __half2float(__ldg(q + lo));
        return v;
    };
    const float v1 = tap(y_low, x_low), v2 = tap(y_low, x_high), v3 = tap(y_high, x_low), v4 = tap(y_high, x_high);
    return hy * hx * v1 + hy * lx * v2 + ly * hx * v3 + ly * lx * v4;
}

// ROIAlignV2 (aligned=True, sampling_ratio=0, 7x7) over the assigned FPN level, all levels in ONE launch.
// grid = (n_rois, 7 output rows), block = 256 threads = 256 channels (each feature tap is one coalesced 1 KiB row).
// Output goes straight into the 9x9 zero-bordered ROI planes (128 rows per ROI) the tower GEMM reads.
// Replaces detectron2 ROIPooler.forward (torchvision roi_align per level + nonzero + index_put_),
// reference call site sylph/modeling/code_generator/code_generator.py:930.
__global__ void __launch_bounds__(256)
roi_align_kernel(const __half* __restrict__ pyramid, PyramidGeom pg, const float* __restrict__ boxes,
                 const int* __restrict__ roi_image, __half* __restrict__ roi_planes, long long* __restrict__ levels_out,
                 int split, int roi_stride) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int roi = blockIdx.x, ph = blockIdx.y, c = threadIdx.x;
    const float bx0 = boxes[roi * 4 + 0], by0 = boxes[roi * 4 + 1], bx1 = boxes[roi * 4 + 2], by1 = boxes[roi * 4 + 3];
    const int lvl = assign_level(bx0, by0, bx1, by1);
    if (ph == 0 && c == 0 && levels_out != nullptr) levels_out[roi] = lvl;
    const PlaneGeom g = pg.lv[lvl];
    const float sc = pg.scale[lvl];
    const int n = roi_image[roi];
    const float x0 = __fsub_rn(__fmul_rn(bx0, sc), 0.5f), y0 = __fsub_rn(__fmul_rn(by0, sc), 0.5f);
    const float x1 = __fsub_rn(__fmul_rn(bx1, sc), 0.5f), y1 = __fsub_rn(__fmul_rn(by1, sc), 0.5f);
    const float roi_w = __fsub_rn(x1, x0), roi_h = __fsub_rn(y1, y0);
    const float bin_h = __fdiv_rn(roi_h, 7.f), bin_w = __fdiv_rn(roi_w, 7.f);
    const int grid_h = static_cast<int>(ceilf(__fdiv_rn(roi_h, 7.f)));
    const int grid_w = static_cast<int>(ceilf(__fdiv_rn(roi_w, 7.f)));
    const float count = fmaxf(static_cast<float>(grid_h * grid_w), 1.f);
    for (int pw = 0; pw < 7; ++pw) {
        float acc = 0.f;
        for (int iy = 0; iy < grid_h; ++iy) {
            const float y = __fadd_rn(__fadd_rn(y0, __fmul_rn(static_cast<float>(ph), bin_h)),
                                      __fdiv_rn(__fmul_rn(static_cast<float>(iy) + 0.5f, bin_h), static_cast<float>(grid_h)));
            for (int ix = 0; ix < grid_w; ++ix) {
                const float x = __fadd_rn(__fadd_rn(x0, __fmul_rn(static_cast<float>(pw), bin_w)),
                                          __fdiv_rn(__fmul_rn(static_cast<float>(ix) + 0.5f, bin_w), static_cast<float>(grid_w)));
                acc += bilinear_tap(pyramid, g, n, g.H, g.W, y, x, c, split ? 512 : 256, split ? 256 : 0);
            }
        }
        const size_t row = static_cast<size_t>(roi) * roi_stride + (ph + 1) * 9 + (pw + 1);
        const float v = fminf(fmaxf(acc / count, -kHalfMax), kHalfMax);
        const __half h = __float2half_rn(v);
        if (split) {
            roi_planes[row * 512 + c] = h;
            roi_planes[row * 512 + 256 + c] = __float2half_rn(v - __half2float(h));
        } else {
            roi_planes[row * 256 + c] = h;
        }
    }
}

// The same ROIAlign, SEPARABLE: the sampling points of a bin form a tensor-product grid and bilinear interpolation
// factorises, so out(ph, pw) = sum_r Wy[ph][r] * sum_c Wx[pw][c] * f(r, c) with two small weight matrices per ROI
// (bilinear taps, the "outside the map" rule and the edge clamps are all per-axis, and 1 / (grid_h * grid_w) splits into
// the two factors).  Every pixel of the ROI footprint is then read about once (16-byte channel vectors) instead of
// once per sample tap: the per-sample kernel above issues 16-64 two-byte loads per output value, which is what bounds
// the 12 030-ROI class sweep of BASELINE configs[4] (4.0 ms, ~0.1 of HBM peak, profiles/r01_configs_3_4_5_s2.json).
// Sample coordinates are computed with exactly the expressions of the per-sample kernel; only the fp32 summation order
// differs (values agree to rounding; the FPN level per ROI stays bit-exact).
// grid = n_rois, block = 256 = 32 channel groups of 8 x 8 row slots (slot ph < 7 owns output row ph).
// Dynamic shared memory: 7 * (R + C) floats with R x C the footprint (rows x columns of the level the ROI reads).
constexpr int kRoiSepMaxSpan = 512;   // footprint rows / columns the separable kernel accepts (else: per-sample kernel)

__global__ void __launch_bounds__(256, 2)
roi_align_separable_kernel(const __half* __restrict__ pyramid, PyramidGeom pg, const float* __restrict__ boxes,
                           const int* __restrict__ roi_image, __half* __restrict__ roi_planes,
                           long long* __restrict__ levels_out, int split, int span_cap, int roi_stride) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    extern __shared__ float roi_w[];
    __shared__ int s_lo[14], s_hi[14];
    const int roi = blockIdx.x, t = threadIdx.x;
    const float bx0 = boxes[roi * 4 + 0], by0 = boxes[roi * 4 + 1], bx1 = boxes[roi * 4 + 2], by1 = boxes[roi * 4 + 3];
    const int lvl = assign_level(bx0, by0, bx1, by1);
    if (t == 0 && levels_out != nullptr) levels_out[roi] = lvl;
    const PlaneGeom g = pg.lv[lvl];
    const float sc = pg.scale[lvl];
    const int n = roi_image[roi];
    const float x0 = __fsub_rn(__fmul_rn(bx0, sc), 0.5f), y0 = __fsub_rn(__fmul_rn(by0, sc), 0.5f);
    const float x1 = __fsub_rn(__fmul_rn(bx1, sc), 0.5f), y1 = __fsub_rn(__fmul_rn(by1, sc), 0.5f);
    const float roi_w_ = __fsub_rn(x1, x0), roi_h = __fsub_rn(y1, y0);
    const float bin_h = __fdiv_rn(roi_h, 7.f), bin_w = __fdiv_rn(roi_w_, 7.f);
    const int grid_h = static_cast<int>(ceilf(__fdiv_rn(roi_h, 7.f)));
    const int grid_w = static_cast<int>(ceilf(__fdiv_rn(roi_w_, 7.f)));
    // footprint on the level: rows r0 .. r0 + R - 1, columns c0 .. c0 + C - 1 (clamped to the plane like the taps are)
    const int r0 = min(max(static_cast<int>(floorf(y0)), 0), g.H - 1);
    const int r1 = min(max(static_cast<int>(floorf(y1)) + 1, r0), g.H - 1);
    const int c0 = min(max(static_cast<int>(floorf(x0)), 0), g.W - 1);
    const int c1 = min(max(static_cast<int>(floorf(x1)) + 1, c0), g.W - 1);
    const int R = min(r1 - r0 + 1, span_cap), C = min(c1 - c0 + 1, span_cap);
    float* Wy = roi_w;            // [7][R]
    float* Wx = roi_w + 7 * R;    // [7][C]
    for (int i = t; i < 7 * (R + C); i += 256) roi_w[i] = 0.f;
    __syncthreads();
    if (t < 14) {   // thread = (axis, bin): accumulate the bilinear tap weights of the bin's sample rows / columns
        const bool is_x = t >= 7;
        const int b = is_x ? t - 7 : t;
        const int grid = is_x ? grid_w : grid_h, size = is_x ? g.W : g.H, span = is_x ? C : R, first = is_x ? c0 : r0;
        const float start = is_x ? x0 : y0, bin = is_x ? bin_w : bin_h;
        float* w = (is_x ? Wx : Wy) + b * span;
        const float inv = grid > 0 ? 1.f / static_cast<float>(grid) : 0.f;
        int lo = span, hi = -1;
        for (int i = 0; i < grid; ++i) {
            float p = __fadd_rn(__fadd_rn(start, __fmul_rn(static_cast<float>(b), bin)),
                                __fdiv_rn(__fmul_rn(static_cast<float>(i) + 0.5f, bin), static_cast<float>(grid)));
            if (p < -1.f || p > static_cast<float>(size)) continue;    // the sample contributes zero
            p = fmaxf(p, 0.f);
            int low = static_cast<int>(p), high;
            if (low >= size - 1) { high = low = size - 1; p = static_cast<float>(low); } else { high = low + 1; }
            const float l = p - low, h = 1.f - l;
            const int a = low - first, bb = high - first;
            if (a >= 0 && a < span) { w[a] += h * inv; lo = min(lo, a); hi = max(hi, a); }
            if (bb >= 0 && bb < span) { w[bb] += l * inv; lo = min(lo, bb); hi = max(hi, bb); }
        }
        s_lo[t] = lo;
        s_hi[t] = hi;
    }
    __syncthreads();
    const int cg = t & 31, ph = t >> 5;   // 8 channels starting at 8 * cg; output row ph (slot 7 idles)
    const int ld = split ? 512 : 256, lo_off = split ? 256 : 0;
    float acc[7][8];
#pragma unroll
    for (int pw = 0; pw < 7; ++pw)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[pw][j] = 0.f;
    if (ph < 7 && s_hi[ph] >= s_lo[ph]) {
        const int rlo = s_lo[ph], rhi = s_hi[ph];
        const float* wy = Wy + ph * R;
        for (int c = 0; c < C; ++c) {
            float col[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const __half* p = pyramid + plane_row(g, n, r0 + rlo, c0 + c) * ld;
            const size_t pitch = static_cast<size_t>(g.Wp) * ld;
            int r = rlo;
            for (; r + 1 <= rhi; r += 2) {     // two rows of loads in flight per thread
                float v[8], u[8];
                load8f(p, cg * 8, lo_off, v);
                load8f(p + pitch, cg * 8, lo_off, u);
                const float w0 = wy[r], w1 = wy[r + 1];
#pragma unroll
                for (int j = 0; j < 8; ++j) col[j] += w0 * v[j] + w1 * u[j];
                p += 2 * pitch;
            }
            if (r <= rhi) {
                float v[8];
                load8f(p, cg * 8, lo_off, v);
                const float w = wy[r];
#pragma unroll
                for (int j = 0; j < 8; ++j) col[j] += w * v[j];
            }
#pragma unroll
            for (int pw = 0; pw < 7; ++pw) {
                const float w = Wx[pw * C + c];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[pw][j] += w * col[j];
            }
        }
    }
    if (ph < 7) {
#pragma unroll
        for (int pw = 0; pw < 7; ++pw) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fminf(fmaxf(acc[pw][j], -kHalfMax), kHalfMax);
            const size_t row = static_cast<size_t>(roi) * roi_stride + (ph + 1) * 9 + (pw + 1);
            store8f(roi_planes + row * ld, cg * 8, lo_off, o, false);
        }
    }
}

// GroupNorm(32, 256) + ReLU over the 7x7 interior of every ROI plane of a raw (fp32) convolution output, written as the fp16
// (hi | lo) operand plane of the next layer.  One block per ROI, one channel per thread: a group is 8 consecutive lanes.  The ROI
// planes are PACKED (roi_stride = 81 rows, no alignment to the 128-row tiles), so a convolution tile holds rows of 2-3 ROIs and
// the per-tile GroupNorm partial sums of the big planes do not apply; 49 x 256 values per ROI need no partials anyway.
// Two-pass statistics in fp32 (reference: torch.nn.GroupNorm, eps 1e-5, biased variance).  Border rows are never written: they
// stay zero from the allocation.   grid = n_rois, block = 256.
__global__ void __launch_bounds__(256)
roi_gn_relu_kernel(const float* __restrict__ raw, __half* __restrict__ out, const float* __restrict__ gamma,
                   const float* __restrict__ beta, int split, int roi_stride) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int roi = blockIdx.x, t = threadIdx.x;
    const float* base = raw + static_cast<size_t>(roi) * roi_stride * 256;
    float v[49];
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < 49; ++p) {
        v[p] = base[static_cast<size_t>((p / 7 + 1) * 9 + (p % 7 + 1)) * 256 + t];
        s += v[p];
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / 392.f;
    float q = 0.f;
#pragma unroll
    for (int p = 0; p < 49; ++p) { const float d = v[p] - mean; q += d * d; }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / 392.f + 1e-5f);
    const float g = gamma[t] * rstd, b = beta[t] - mean * g;
    const int ld = split ? 512 : 256;
    __half* ob = out + static_cast<size_t>(roi) * roi_stride * ld;
#pragma unroll
    for (int p = 0; p < 49; ++p) {
        const float y = fminf(fmaxf(v[p] * g + b, 0.f), kHalfMax);
        const __half hi = __float2half_rn(y);
        __half* o = ob + static_cast<size_t>((p / 7 + 1) * 9 + (p % 7 + 1)) * ld;
        o[t] = hi;
        if (split) o[256 + t] = __float2half_rn(y - __half2float(hi));
    }
}

// Pool BEFORE the cls convolution: the reference's support_set_cls_conv is conv3x3(256 -> 256, bias) followed directly by a global
// average pool (CLS_LAYER ["", "", 1]: no norm, no ReLU in between; code_generator.py:954, utils.py:51-67), and the mean over the 49
// positions commutes with the convolution:  mean_p conv(x)[p] = sum_tap W_tap . (mean over the 7x7 window shifted by the tap) + b.
// This kernel writes, per ROI, the nine window means of the tower output (zero padding = the plane's zero border) as one row of
// 9 x 256 values (hi | lo halves in exact mode); the convolution becomes a [n_rois x 2304] x [2304 x 256] GEMM -- 1 / 128 of the
// tensor work of the per-pixel convolution over 128-row ROI tiles.   grid = n_rois, block = 256 (one channel per thread).
__global__ void __launch_bounds__(256)
roi_window_means_kernel(const __half* __restrict__ tower_out, __half* __restrict__ win /* [rows][9 * 256 (x2)] */, int split,
                        int roi_stride) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int roi = blockIdx.x, t = threadIdx.x;
    const int ld = split ? 512 : 256;
    const __half* base = tower_out + static_cast<size_t>(roi) * roi_stride * ld;
    float s[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) s[k] = 0.f;
#pragma unroll
    for (int y = 1; y <= 7; ++y)
#pragma unroll
        for (int x = 1; x <= 7; ++x) {
            const __half* a = base + static_cast<size_t>(y * 9 + x) * ld;
            float v = __half2float(a[t]);
            if (split) v += __half2float(a[256 + t]);
            // pixel (y, x) lies in the window of tap (dy, dx) iff 1 + dy <= y <= 7 + dy and 1 + dx <= x <= 7 + dx
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx)
                    if (y >= 1 + dy && y <= 7 + dy && x >= 1 + dx && x <= 7 + dx) s[(dy + 1) * 3 + (dx + 1)] += v;
        }
    __half* o = win + static_cast<size_t>(roi) * (split ? 4608 : 2304);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float m = s[k] / 49.f;
        const __half hi = __float2half_rn(m);
        o[k * 256 + t] = hi;
        if (split) o[2304 + k * 256 + t] = __float2half_rn(m - __half2float(hi));
    }
}

// Per support ROI ("shot"): global average pool of the cls-conv output (GlobalAdaptiveAvgPool2d, k_s = 1) and the
// 256 -> 1 3x3 bias convolution on the tower output, optional L2 normalisation over the 49 positions, then its pool.
// reference: code_generator.py:954-967, utils.py:51-67.   grid = n_rois, block = 256.
__global__ void __launch_bounds__(256)
shot_code_kernel(const float* __restrict__ cls_raw, const __half* __restrict__ tower_out,
                 const float* __restrict__ w_bias /* [9][256] tap-major */, const float* __restrict__ b_bias,
                 int has_bias_layer, int bias_l2_norm, float* __restrict__ shot_codes /* [n_rois][257] */, int split,
                 const float* __restrict__ cls_pooled = nullptr /* [n_rois][256]: the pooled cls-conv output (pool-before-conv path) */,
                 int roi_stride = 128,
                 const float* __restrict__ w_weight = nullptr /* [9][256]: WEIGHT_LAYER head */, const float* __restrict__ b_weight = nullptr,
                 float* __restrict__ shot_wlogit = nullptr /* [n_rois]: pooled output of the weight head */) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ float pix[49];
    const int roi = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const size_t base = static_cast<size_t>(roi) * roi_stride;
    if (cls_pooled != nullptr) {
        shot_codes[static_cast<size_t>(roi) * 257 + t] = cls_pooled[static_cast<size_t>(roi) * 256 + t];
    } else {
        float s = 0.f;
        for (int p = 0; p < 49; ++p) {
            const int row = (p / 7 + 1) * 9 + (p % 7 + 1);
            s += cls_raw[(base + row) * 256 + t];
        }
        shot_codes[static_cast<size_t>(roi) * 257 + t] = s / 49.f;
    }
    // a 256 -> 1 3x3 convolution of the tower output at the 49 positions (bias head, weight head): pix[p]
    auto head_conv = [&](const float* __restrict__ w9, const float* __restrict__ b1) {
        for (int p = warp; p < 49; p += 8) {
            const int row = (p / 7 + 1) * 9 + (p % 7 + 1);
            float acc = 0.f;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int r2 = row + (tap / 3 - 1) * 9 + (tap % 3 - 1);  // zero border supplies the padding
                const __half* a = tower_out + (base + r2) * (split ? 512 : 256);
                const float* w = w9 + tap * 256;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float av = __half2float(a[lane + 32 * j]);
                    if (split) av += __half2float(a[256 + lane + 32 * j]);
                    acc += av * __ldg(w + lane + 32 * j);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) pix[p] = acc + b1[0];
        }
        __syncthreads();
    };
    if (shot_wlogit != nullptr) {   // weight head: conv + GlobalAdaptiveAvgPool2d (code_generator.py:583-612)
        head_conv(w_weight, b_weight);
        if (t == 0) {
            float m = 0.f;
            for (int p = 0; p < 49; ++p) m += pix[p];
            shot_wlogit[roi] = m / 49.f;
        }
        __syncthreads();
    }
    if (!has_bias_layer) {
        if (t == 0) shot_codes[static_cast<size_t>(roi) * 257 + 256] = 0.f;
        return;
    }
    head_conv(w_bias, b_bias);
    if (t == 0) {
        float denom = 1.f;
        if (bias_l2_norm) {
            float ss = 0.f;
            for (int p = 0; p < 49; ++p) ss += pix[p] * pix[p];
            denom = fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps
        }
        float m = 0.f;
        for (int p = 0; p < 49; ++p) m += pix[p] / denom;
        shot_codes[static_cast<size_t>(roi) * 257 + 256] = m / 49.f;
    }
}

// K-shot combination per class with the reference's op order: sum_k w_k * x_k (compute_code, code_generator.py:805-817) with
// w_k = 1 / K, or -- WEIGHT_LAYER -- w = softmax over the class's shots of the weight head's logits (process_weight, :766-776).
__global__ void __launch_bounds__(288)
class_mean_kernel(const float* __restrict__ shot_codes, const int* __restrict__ class_offsets,
                  float* __restrict__ raw_codes, const float* __restrict__ shot_wlogit = nullptr) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int cls = blockIdx.x, t = threadIdx.x;
    if (t >= 257) return;
    const int k0 = class_offsets[cls], k1 = class_offsets[cls + 1];
    float s = 0.f;
    if (shot_wlogit != nullptr) {
        float mx = -3.4e38f;
        for (int k = k0; k < k1; ++k) mx = fmaxf(mx, shot_wlogit[k]);
        float den = 0.f;
        for (int k = k0; k < k1; ++k) den += expf(shot_wlogit[k] - mx);
        for (int k = k0; k < k1; ++k) s += (expf(shot_wlogit[k] - mx) / den) * shot_codes[static_cast<size_t>(k) * 257 + t];
    } else {
        const float w = 1.0f / static_cast<float>(k1 - k0);
        for (int k = k0; k < k1; ++k) s += w * shot_codes[static_cast<size_t>(k) * 257 + t];
    }
    raw_codes[static_cast<size_t>(cls) * 257 + t] = s;
}

// Base-class "all ground truths" path: every class arrives as several chunks of <= 10 support boxes, each chunk's code
// weighted by len / total_len and summed in arrival order (inference_on_support_set_dataset_base,
// sylph/evaluation/meta_learn_evaluation.py:190-203).  One block per class walks the chunk list in order, so the fp32
// rounding sequence is the reference's: acc = fl(acc + fl(code * w)) (no FMA contraction).
__global__ void __launch_bounds__(288)
accumulate_codes_kernel(const float* __restrict__ chunk_codes, const int* __restrict__ chunk_class,
                        const float* __restrict__ chunk_weight, int n_chunks, float* __restrict__ acc) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int cls = blockIdx.x, t = threadIdx.x;
    if (t >= 257) return;
    float s = acc[static_cast<size_t>(cls) * 257 + t];
    for (int k = 0; k < n_chunks; ++k) {
        if (chunk_class[k] != cls) continue;
        s = __fadd_rn(s, __fmul_rn(chunk_codes[static_cast<size_t>(k) * 257 + t], chunk_weight[k]));
    }
    acc[static_cast<size_t>(cls) * 257 + t] = s;
}

// reduce_class_code (sylph/modeling/code_generator/utils.py:397-427): sum the per-rank partial codes of a class in
// rank order (functools.reduce starting from 0), then divide by the accumulated weight where it differs from 1
// (`divisor[cls] != 0`; the host decides with the reference's |1 - acc_weight| > 1e-6 test in double precision).
__global__ void __launch_bounds__(288)
reduce_codes_kernel(const float* __restrict__ parts, int n_parts, int n_classes, const float* __restrict__ divisor,
                    float* __restrict__ out) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int cls = blockIdx.x, t = threadIdx.x;
    if (t >= 257) return;
    float s = 0.f;
    for (int r = 0; r < n_parts; ++r) s = __fadd_rn(s, parts[(static_cast<size_t>(r) * n_classes + cls) * 257 + t]);
    const float d = divisor[cls];
    out[static_cast<size_t>(cls) * 257 + t] = d != 0.f ? __fdiv_rn(s, d) : s;
}

// Code normalisation of one channel, called by all 256 threads of a block (code_process_module,
// code_generator.py:832-875): GroupNorm(32, 256) on the 1x1 map (8-lane shuffles) -> L2 normalise over the 256 channels.
// `red` is an 8-float shared scratch.
__device__ __forceinline__ float normalize_code_channel(float x, const float* __restrict__ gn_w, const float* __restrict__ gn_b,
                                                        int post_norm, int l2_norm, float* red, int t) {
    const int lane = t & 31, warp = t >> 5;
    if (post_norm) {
        float s = x;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        const float mean = s / 8.f;
        const float d = x - mean;
        float v = d * d;
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        x = d * rsqrtf(v / 8.f + 1e-5f) * gn_w[t] + gn_b[t];
    }
    if (l2_norm) {
        float ss = x * x;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += red[i];
        x = x / fmaxf(sqrtf(tot), 1e-12f);
    }
    return x;
}

// Code normalisation, one 256-thread block per class: normalised channel x conv_scale;
// bias x bias_scale + (-log((1 - p) / p)).
__global__ void __launch_bounds__(256)
normalize_codes_kernel(const float* __restrict__ raw, float* __restrict__ out, const float* __restrict__ gn_w,
                       const float* __restrict__ gn_b, int post_norm, int l2_norm, float conv_scale, float bias_scale,
                       float bias_value) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ float red[8];
    const int cls = blockIdx.x, t = threadIdx.x;
    float x = raw[static_cast<size_t>(cls) * 257 + t];
    const float b = raw[static_cast<size_t>(cls) * 257 + 256];
    x = normalize_code_channel(x, gn_w, gn_b, post_norm, l2_norm, red, t);
    out[static_cast<size_t>(cls) * 257 + t] = x * conv_scale;
    if (t == 0) out[static_cast<size_t>(cls) * 257 + 256] = b * bias_scale + bias_value;
}

// ---------------------------------------------------------------------------------------------------------------
// Class-code exchange over NVLink peer memory: the normalisation kernel IS the all-gather.  Every rank owns one
// exchange allocation (ExchangeState + two code buffers of max_classes rows, double-buffered by episode parity) that
// all ranks of the box map through CUDA IPC.  The producer normalises the rank's class shard and stores every row
// straight into ALL ranks' buffers (coalesced 1 KB rows over NVLink), then publishes the row with one system-scope
// atomic per peer; the consumer waits until the rows of all classes of this episode have landed locally and hands
// them to the caller.  Replaces MetaFCOSRunner._gather_class_code (sylph/runner/meta_fcos_runner.py:381-396:
// all_gather_object of pickled code dicts) + inference_normalization (meta_learn_evaluation.py:105-116) for the
// sharded episode; no NCCL call, no host synchronisation, two launches.
struct ExchangeState {
    unsigned long long arrived;      // rows that have landed in THIS rank's buffers, all episodes (remote atomics)
    unsigned long long pad0[15];     // keep the remotely updated word in its own 128-byte line
    unsigned long long expected;     // rows consumed by this rank so far (local; advanced by the collect kernel)
    unsigned int parity;             // buffer half of the CURRENT episode (local; same value on every rank)
    unsigned int error;              // sticky: a collect kernel gave up waiting (local)
    unsigned int blocks_done;        // blocks of the running collect kernel that have finished (local)
    unsigned int pad1[27];
};
static_assert(sizeof(ExchangeState) == 256, "ExchangeState is the 256-byte header of the exchange allocation");

constexpr int kMaxExchangePeers = 16;
struct ExchangePeers {
    float* codes[kMaxExchangePeers];                 // peer r's code buffers: [2][max_classes][257]
    unsigned long long* arrived[kMaxExchangePeers];  // &peer r's ExchangeState::arrived
    int world;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// One 256-thread block per local class.  do_norm = 0 forwards the rows unchanged (ROIEncoder codes are final).
__global__ void __launch_bounds__(256)
normalize_scatter_codes_kernel(const float* __restrict__ raw, ExchangePeers peers, const ExchangeState* self_state,
                               int class_offset, int max_classes, int do_norm, const float* __restrict__ gn_w,
                               const float* __restrict__ gn_b, int post_norm, int l2_norm, float conv_scale,
                               float bias_scale, float bias_value) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ float red[8];
    const int cls = blockIdx.x, t = threadIdx.x;
    float x = raw[static_cast<size_t>(cls) * 257 + t];
    float b = raw[static_cast<size_t>(cls) * 257 + 256];
    if (do_norm) {
        x = normalize_code_channel(x, gn_w, gn_b, post_norm, l2_norm, red, t) * conv_scale;
        b = b * bias_scale + bias_value;
    }
    const unsigned int parity = self_state->parity;   // written by the previous collect kernel of this stream
    const size_t row = (static_cast<size_t>(parity) * max_classes + class_offset + cls) * 257;
    for (int r = 0; r < peers.world; ++r) {
        peers.codes[r][row + t] = x;
        if (t == 0) peers.codes[r][row + 256] = b;
    }
    __syncthreads();
    // release: the block's stores (ordered before the barrier) become visible system-wide before the row is counted
    if (t < peers.world) {
        __threadfence_system();
        atomicAdd_system(peers.arrived[t], 1ULL);
    }
}

// Wait (bounded) until n_total rows of this episode have landed, copy them to the caller's buffer, then advance the
// rank's episode state.  A few blocks share the copy (grid-stride); each waits on the counter by itself and the last
// block to finish advances the state, so no block depends on another being resident.  The timeout only guards against
// ranks that disagree on the call sequence.
__global__ void __launch_bounds__(1024)
collect_codes_kernel(ExchangeState* state, const float* local_codes, int max_classes, int n_total, float* __restrict__ out,
                     unsigned long long timeout_ns) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ unsigned int s_parity;
    __shared__ unsigned long long s_want;
    if (threadIdx.x == 0) {
        const unsigned long long want = state->expected + static_cast<unsigned long long>(n_total);
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys_u64(&state->arrived) < want) {
            if (global_timer_ns() - t0 > timeout_ns) {
                state->error = 1u;
                break;
            }
            __nanosleep(200);
        }
        s_parity = state->parity;
        s_want = want;
    }
    __syncthreads();
    const float* src = local_codes + static_cast<size_t>(s_parity) * max_classes * 257;
    const int n = n_total * 257, step = gridDim.x * blockDim.x;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * step < n; i += 4 * step) {   // L2 is the coherence point for peer writes: ld.cg
        const float a = __ldcg(src + i), b = __ldcg(src + i + step), c = __ldcg(src + i + 2 * step), d = __ldcg(src + i + 3 * step);
        out[i] = a;
        out[i + step] = b;
        out[i + 2 * step] = c;
        out[i + 3 * step] = d;
    }
    for (; i < n; i += step) out[i] = __ldcg(src + i);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&state->blocks_done, 1u) == gridDim.x - 1) {   // every block has read expected / parity by now
            state->blocks_done = 0u;
            state->expected = s_want;
            state->parity = s_parity ^ 1u;
        }
    }
}

// Expand (n_classes, 257) codes into the K-major weight matrix [n_pad][256] + bias [n_pad] the logits GEMM reads
// (rows >= n_classes are zero); weights are rounded to fp16 here.
// `scale` is the learned output scale of CondConvBlock (head_utils.py:121-162, ROIEncoder head):
// scale * (conv(x, W) + b) = conv(x, scale * W) + scale * b; 1 for CondConvBasic.
// split mode: rows of [w_hi | w_hi | w_lo] (K' = 768) for the three-product k loop (GemmArgs::a_wrap).
__global__ void pack_code_weights_kernel(const float* __restrict__ codes, int n_classes, int n_pad, int use_bias, float scale,
                                         __half* __restrict__ w, float* __restrict__ bias, int split) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad * 256) return;
    const int r = i >> 8, c = i & 255;
    const float v = r < n_classes ? codes[static_cast<size_t>(r) * 257 + c] * scale : 0.f;
    const __half h = __float2half_rn(fminf(fmaxf(v, -kHalfMax), kHalfMax));
    if (split) {
        w[r * 768 + c] = h;
        w[r * 768 + 256 + c] = h;
        w[r * 768 + 512 + c] = __float2half_rn(v - __half2float(h));
    } else {
        w[i] = h;
    }
    if (c == 0) bias[r] = (r < n_classes && use_bias) ? codes[static_cast<size_t>(r) * 257 + 256] * scale : 0.f;
}

// (n_rois, 256, 7, 7) export of the pooled ROI planes (tests / plugin interop).
__global__ void export_roi_kernel(const __half* __restrict__ roi_planes, float* __restrict__ out, int n_rois, int split,
                                  int roi_stride) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const long long total = static_cast<long long>(n_rois) * 256 * 49;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(i % 49);
        const int c = static_cast<int>((i / 49) % 256);
        const int r = static_cast<int>(i / (49 * 256));
        const size_t row = static_cast<size_t>(r) * roi_stride + (p / 7 + 1) * 9 + (p % 7 + 1);
        out[i] = split ? __half2float(roi_planes[row * 512 + c]) + __half2float(roi_planes[row * 512 + 256 + c])
                       : __half2float(roi_planes[row * 256 + c]);
    }
}

}  // namespace sylph
