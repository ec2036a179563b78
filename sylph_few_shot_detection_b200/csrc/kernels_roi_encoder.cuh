// ROIEncoder code generator (the second registered CODE_GENERATOR plugin, sylph/modeling/code_generator/roi_encoder.py
// :118-281) -- the kernels around the tensor-core convolutions / GEMM:
//   context_pool_kernel   GlobalAdaptiveAvgPool2d(7) of every FPN level, averaged over levels
//                         (FeatureFusionModuleV2.forward, sylph/modeling/code_generator/utils.py:153-160)
//   ms_cam_kernel         MS_CAM local + global channel attention on the context, gate of the pooled ROI features
//                         (utils.py:70-103)
//   gather_tokens_kernel  7x7x256 ROI planes -> rows of the [tokens][12544] fp16 matrix the fc1 GEMM reads
//                         (Tokenizer "flatten", roi_encoder.py:59-60; the weight matrix is permuted at load time)
//   linear_kernel         small dense layers (tokenizer fc2.., folded self-attention, FFN, hyper-network heads)
//   add_layernorm_kernel  post-norm residual of nn.TransformerEncoderLayer
//   token_mean_kernel     box_tokens.mean(1) per class (roi_encoder.py:187)
// These move kilobytes per ROI (except the context pooling, which streams the support pyramid once per ROI): they are
// latency/HBM-bound CUDA-core kernels with coalesced channel-vector loads and warp-shuffle reductions.
//
// Eval-time transformer semantics: the reference builds nn.TransformerEncoderLayer with batch_first=False and feeds
// (bs, num_shots, C) -- the sequence axis is the CLASS axis.  forward_class_code runs one class per call (bs = 1), so
// every token attends to a sequence of length 1: softmax over one key is 1 and the layer reduces to
// x1 = LN1(x + W_o (W_v x + b_v) + b_o), x2 = LN2(x1 + W_2 relu(W_1 x1 + b_1) + b_2).  W_o W_v is folded on the host.
#pragma once
#include "kernels_codegen.cuh"

namespace sylph {

// ------------------------------------------------------------------------------------------------ context pooling
// grid = (n_images, 49 bins), block = 256 channels.  ctx[image][bin][c] = mean over the 5 levels of the adaptive-average
// bin (torch adaptive_avg_pool2d: rows floor(i*H/7) .. ceil((i+1)*H/7) - 1).  The context depends on the image only, so it
// is computed once per support image and shared by all ROIs of that image (the pyramid is streamed exactly once).
__global__ void __launch_bounds__(256)
context_pool_kernel(const __half* __restrict__ pyramid, PyramidGeom pg, float* __restrict__ ctx /* [n_images][49][256] */,
                    int split) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int n = blockIdx.x, bin = blockIdx.y, c = threadIdx.x;
    const int bi = bin / 7, bj = bin - bi * 7;
    float total = 0.f;
    if (split) {   // rows of [256 hi | 256 lo]: the value of a channel is hi + lo (exact in fp32)
#pragma unroll 1
        for (int l = 0; l < 5; ++l) {
            const PlaneGeom g = pg.lv[l];
            const int y0 = (bi * g.H) / 7, y1 = ((bi + 1) * g.H + 6) / 7;
            const int x0 = (bj * g.W) / 7, x1 = ((bj + 1) * g.W + 6) / 7;
            float s = 0.f;
            for (int y = y0; y < y1; ++y) {
                const __half* row = pyramid + plane_row(g, n, y, x0) * 512 + c;
                int x = x0;
                for (; x + 2 <= x1; x += 2) {
                    const float a = __half2float(__ldg(row)) + __half2float(__ldg(row + 256));
                    const float b = __half2float(__ldg(row + 512)) + __half2float(__ldg(row + 768));
                    s += a; s += b;
                    row += 1024;
                }
                for (; x < x1; ++x) { s += __half2float(__ldg(row)) + __half2float(__ldg(row + 256)); row += 512; }
            }
            total += s / static_cast<float>((y1 - y0) * (x1 - x0));
        }
        ctx[(static_cast<size_t>(n) * 49 + bin) * 256 + c] = total / 5.f;
        return;
    }
#pragma unroll 1
    for (int l = 0; l < 5; ++l) {
        const PlaneGeom g = pg.lv[l];
        const int y0 = (bi * g.H) / 7, y1 = ((bi + 1) * g.H + 6) / 7;
        const int x0 = (bj * g.W) / 7, x1 = ((bj + 1) * g.W + 6) / 7;
        float s = 0.f;
        for (int y = y0; y < y1; ++y) {
            const __half* row = pyramid + plane_row(g, n, y, x0) * 256 + c;
            int x = x0;
            for (; x + 4 <= x1; x += 4) {   // four independent 512-byte row reads in flight
                const float a = __half2float(__ldg(row)), b = __half2float(__ldg(row + 256));
                const float d = __half2float(__ldg(row + 512)), e = __half2float(__ldg(row + 768));
                s += a; s += b; s += d; s += e;
                row += 1024;
            }
            for (; x < x1; ++x) { s += __half2float(__ldg(row)); row += 256; }
        }
        total += s / static_cast<float>((y1 - y0) * (x1 - x0));
    }
    ctx[(static_cast<size_t>(n) * 49 + bin) * 256 + c] = total / 5.f;
}

// ------------------------------------------------------------------------------------------------ MS_CAM
struct MsCamWeights {
    const float* l_w1t;   // local_att.0  weight, transposed [256][64]
    const float* l_b1;    // [64]
    const float* l_g1w;   // local_att.1 GroupNorm(32, 64)
    const float* l_g1b;
    const float* l_w2t;   // local_att.3 weight, transposed [64][256]
    const float* l_b2;    // [256]
    const float* l_g2w;   // local_att.4 GroupNorm(32, 256)
    const float* l_g2b;
    const float* g_w1;    // global_att.1 weight [64][256]
    const float* g_b1;
    const float* g_g1w;   // global_att.2 GroupNorm(32, 64)
    const float* g_g1b;
    const float* g_w2t;   // global_att.4 weight, transposed [64][256]
    const float* g_b2;
    const float* g_g2w;   // global_att.5 GroupNorm(32, 256)
    const float* g_g2b;
};

constexpr int kMsCamSmem = (49 * 256 + 256 * 64 + 49 * 64 + 256 + 64 + 64) * 4;

// The attention gate sigmoid(local_att(ctx) + global_att(ctx)) depends on the IMAGE only (ctx is the image's pooled
// context), so it is computed once per support image -- grid = n_images, block = 256, gate[image][49][256] fp32 -- and
// applied to every ROI of that image by ms_cam_apply_kernel.  (Round 1 recomputed both branches per ROI: 6.6 ms of the
// 19.4 ms LVIS class sweep, where 12 030 ROIs share 16 images.)
__global__ void __launch_bounds__(256)
ms_cam_gate_kernel(const float* __restrict__ ctx, MsCamWeights w, float* __restrict__ gate) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    extern __shared__ uint8_t cam_smem_raw[];
    float* s_ctx = reinterpret_cast<float*>(cam_smem_raw);                        // [49][256]
    float* s_w = s_ctx + 49 * 256;            // [256][64] (conv1, transposed) then [64][256] (conv2, transposed)
    float* s_h1 = s_w + 256 * 64;             // [49][64]
    float* s_g0 = s_h1 + 49 * 64;             // [256]
    float* s_g1 = s_g0 + 256;                 // [64]
    float* s_st = s_g1 + 64;                  // [32][2] GroupNorm(32, 64) statistics of the local branch
    const int img = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float* cx = ctx + static_cast<size_t>(img) * 49 * 256;
    for (int i = t; i < 49 * 256; i += 256) s_ctx[i] = cx[i];
    for (int i = t; i < 256 * 64; i += 256) s_w[i] = w.l_w1t[i];
    __syncthreads();
    // ---- global branch, step 1: channel means of the context
    {
        float s = 0.f;
        for (int p = 0; p < 49; ++p) s += s_ctx[p * 256 + t];
        s_g0[t] = s / 49.f;
    }
    // ---- local branch, conv1 (256 -> 64) on the 49 positions: thread = (position group, output channel)
    {
        const int j = t & 63, pg0 = t >> 6;
        for (int p = pg0; p < 49; p += 4) {
            float acc = w.l_b1[j];
            const float* xr = s_ctx + p * 256;
#pragma unroll 8
            for (int c = 0; c < 256; ++c) acc += xr[c] * s_w[c * 64 + j];
            s_h1[p * 64 + j] = acc;
        }
    }
    __syncthreads();
    // ---- GroupNorm(32, 64) over (2 channels x 49 positions) + ReLU
    if (t < 32) {
        float s = 0.f, ss = 0.f;
        for (int p = 0; p < 49; ++p) {
            const float a = s_h1[p * 64 + 2 * t], b = s_h1[p * 64 + 2 * t + 1];
            s += a + b;
            ss += a * a + b * b;
        }
        const float mean = s / 98.f;
        const float var = fmaxf(ss / 98.f - mean * mean, 0.f);
        s_st[2 * t] = mean;
        s_st[2 * t + 1] = rsqrtf(var + 1e-5f);
    }
    for (int i = t; i < 64 * 256; i += 256) s_w[i] = w.l_w2t[i];   // conv1 weights are dead: stage conv2
    // ---- global branch, conv1 (256 -> 64) on the pooled vector: warp-cooperative dot products
    for (int j = warp; j < 64; j += 8) {
        float acc = 0.f;
        for (int c = lane; c < 256; c += 32) acc += w.g_w1[j * 256 + c] * s_g0[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_g1[j] = acc + w.g_b1[j];
    }
    __syncthreads();
    for (int i = t; i < 49 * 64; i += 256) {
        const int j = i & 63, g = j >> 1;
        const float v = (s_h1[i] - s_st[2 * g]) * s_st[2 * g + 1] * w.l_g1w[j] + w.l_g1b[j];
        s_h1[i] = fmaxf(v, 0.f);
    }
    if (t < 32) {   // GroupNorm(32, 64) of the global branch: 2 values per group (1x1 map) + ReLU
        const float a = s_g1[2 * t], b = s_g1[2 * t + 1];
        const float mean = (a + b) * 0.5f;
        const float var = ((a - mean) * (a - mean) + (b - mean) * (b - mean)) * 0.5f;
        const float r = rsqrtf(var + 1e-5f);
        s_g1[2 * t] = fmaxf((a - mean) * r * w.g_g1w[2 * t] + w.g_g1b[2 * t], 0.f);
        s_g1[2 * t + 1] = fmaxf((b - mean) * r * w.g_g1w[2 * t + 1] + w.g_g1b[2 * t + 1], 0.f);
    }
    __syncthreads();
    // ---- conv2 of both branches, thread = output channel; GroupNorm(32, 256) statistics through 8-lane shuffles
    float gl;
    {
        float acc = w.g_b2[t];
#pragma unroll 8
        for (int j = 0; j < 64; ++j) acc += s_g1[j] * w.g_w2t[j * 256 + t];
        float s = acc, ss;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        const float mean = s / 8.f;
        ss = (acc - mean) * (acc - mean);
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        ss += __shfl_xor_sync(0xffffffffu, ss, 4);
        gl = (acc - mean) * rsqrtf(ss / 8.f + 1e-5f) * w.g_g2w[t] + w.g_g2b[t];
    }
    float h2[49];
    float s = 0.f, ss = 0.f;
    const float b2 = w.l_b2[t];
#pragma unroll
    for (int p = 0; p < 49; ++p) {
        float acc = b2;
        const float* hr = s_h1 + p * 64;
#pragma unroll 8
        for (int j = 0; j < 64; ++j) acc += hr[j] * s_w[j * 256 + t];
        h2[p] = acc;
        s += acc;
        ss += acc * acc;
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    const float mean = s / 392.f;
    const float rstd = rsqrtf(fmaxf(ss / 392.f - mean * mean, 0.f) + 1e-5f);
    const float gw = w.l_g2w[t], gb = w.l_g2b[t];
#pragma unroll
    for (int p = 0; p < 49; ++p) {
        const float lg = (h2[p] - mean) * rstd * gw + gb + gl;
        gate[(static_cast<size_t>(img) * 49 + p) * 256 + t] = 1.f / (1.f + __expf(-lg));
    }
}

// out = pooled * gate[image of the ROI], written as a zero-bordered 9x9 plane (128 rows per ROI); one 8-channel vector
// per thread.  (FeatureFusionModuleV2.forward, sylph/modeling/code_generator/utils.py:153-165.)
__global__ void __launch_bounds__(256)
ms_cam_apply_kernel(const float* __restrict__ gate, const int* __restrict__ roi_image, const __half* __restrict__ pooled,
                    __half* __restrict__ out, int n_rois, int split, int roi_stride) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int ld = split ? 512 : 256, lo = split ? 256 : 0;
    const long long total = static_cast<long long>(n_rois) * roi_stride * 32;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c8 = static_cast<int>(i & 31);
        const long long row = i >> 5;
        const int roi = static_cast<int>(row / roi_stride), r = static_cast<int>(row - static_cast<long long>(roi) * roi_stride);
        const int y = r / 9, x = r - y * 9;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (r < 81 && y >= 1 && y <= 7 && x >= 1 && x <= 7) {
            load8f(pooled + row * ld, c8 * 8, lo, v);
            const float4* gp = reinterpret_cast<const float4*>(gate + (static_cast<size_t>(__ldg(roi_image + roi)) * 49 + (y - 1) * 7 + (x - 1)) * 256) + 2 * c8;
            const float4 a = __ldg(gp), b = __ldg(gp + 1);
            v[0] *= a.x; v[1] *= a.y; v[2] *= a.z; v[3] *= a.w; v[4] *= b.x; v[5] *= b.y; v[6] *= b.z; v[7] *= b.w;
        }
        store8f(out + row * ld, c8 * 8, lo, v, false);
    }
}

// ------------------------------------------------------------------------------------------------ tokens
// planes [n][128 rows][256] -> A[n][p * 256 + c] (p = 7 * y + x); one uint4 (8 channels) per thread.
// split mode: planes rows are [256 hi | 256 lo], token rows [12544 hi | 12544 lo].
__global__ void gather_tokens_kernel(const __half* __restrict__ planes, __half* __restrict__ tokens, int n_rois, int split,
                                     int roi_stride) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const long long total = static_cast<long long>(n_rois) * 49 * 32;
    if (split) {
        for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < 2 * total;
             i += static_cast<long long>(gridDim.x) * blockDim.x) {
            const int half_idx = static_cast<int>(i / total);           // 0 = hi, 1 = lo
            const long long j = i - half_idx * total;
            const int v = static_cast<int>(j & 31);
            const int p = static_cast<int>((j >> 5) % 49);
            const long long n = j / (49 * 32);
            const int row = (p / 7 + 1) * 9 + (p % 7 + 1);
            reinterpret_cast<uint4*>(tokens + n * 25088 + half_idx * 12544)[p * 32 + v] =
                __ldg(reinterpret_cast<const uint4*>(planes + (n * roi_stride + row) * 512 + half_idx * 256) + v);
        }
        return;
    }
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(i & 31);
        const int p = static_cast<int>((i >> 5) % 49);
        const long long n = i / (49 * 32);
        const int row = (p / 7 + 1) * 9 + (p % 7 + 1);
        reinterpret_cast<uint4*>(tokens)[i] = __ldg(reinterpret_cast<const uint4*>(planes + (n * roi_stride + row) * 256) + v);
    }
}

// y[t][o] = act(sum_k x[t][k] * W[o][k] + b[o]) (+ add_const).  grid = (ceil(T / 8), ceil(N / 64)), block = 256:
// 8 tokens of x live in shared memory, each warp owns 8 output neurons and streams their weight rows once
// (coalesced), keeping 8 token accumulators per lane; warp-shuffle reduction.  K <= 1024.
__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W, const float* __restrict__ b,
              float* __restrict__ y, int ldy, int T, int K, int N, int relu, float add_const) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ float xs[8][1024];
    const int t0 = blockIdx.x * 8, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * K; i += 256) {
        const int tt = i / K, k = i - tt * K;
        xs[tt][k] = (t0 + tt < T) ? x[static_cast<size_t>(t0 + tt) * ldx + k] : 0.f;
    }
    __syncthreads();
    for (int oo = 0; oo < 8; ++oo) {
        const int o = blockIdx.y * 64 + warp * 8 + oo;
        if (o >= N) break;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const float* wr = W + static_cast<size_t>(o) * K;
        for (int k = lane; k < K; k += 32) {
            const float wv = __ldg(wr + k);
#pragma unroll
            for (int tt = 0; tt < 8; ++tt) acc[tt] += wv * xs[tt][k];
        }
#pragma unroll
        for (int tt = 0; tt < 8; ++tt) {
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) acc[tt] += __shfl_xor_sync(0xffffffffu, acc[tt], s);
        }
        if (lane < 8 && t0 + lane < T) {
            float v = acc[0];
#pragma unroll
            for (int tt = 1; tt < 8; ++tt) v = (lane == tt) ? acc[tt] : v;
            v += b[o];
            if (relu) v = fmaxf(v, 0.f);
            y[static_cast<size_t>(t0 + lane) * ldy + o] = v + add_const;
        }
    }
}

// fp32 rows [T][K] -> fp16 operand rows for the tensor-core GEMM: [K hi | K lo] (exact mode) or [K] (fast mode); 8 values per thread.
__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ x, __half* __restrict__ out, int T, int K, int split) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int k8n = K / 8;
    const long long total = static_cast<long long>(T) * k8n;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int k8 = static_cast<int>(i % k8n);
        const long long t = i / k8n;
        const float4 a = __ldg(reinterpret_cast<const float4*>(x + t * K + k8 * 8));
        const float4 b = __ldg(reinterpret_cast<const float4*>(x + t * K + k8 * 8) + 1);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        __half* row = out + t * (split ? 2 * K : K);
        store8f(row, k8 * 8, split ? K : 0, v, false);
    }
}

// out[t] = LayerNorm(x[t] + r[t]) * g + b over 256 features (eps = 1e-5), one warp per token.
__global__ void __launch_bounds__(256)
add_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ r, const float* __restrict__ g,
                     const float* __restrict__ b, float* __restrict__ out, int T) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= T) return;
    float v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        v[i] = x[static_cast<size_t>(t) * 256 + lane + 32 * i] + r[static_cast<size_t>(t) * 256 + lane + 32 * i];
        s += v[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / 256.f;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) ss += (v[i] - mean) * (v[i] - mean);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss / 256.f + 1e-5f);
#pragma unroll
    for (int i = 0; i < 8; ++i)
        out[static_cast<size_t>(t) * 256 + lane + 32 * i] = (v[i] - mean) * rstd * g[lane + 32 * i] + b[lane + 32 * i];
}

// class token = mean over the class's shot tokens (box_tokens.mean(1)); grid = n_classes, block = 256.
__global__ void __launch_bounds__(256)
token_mean_kernel(const float* __restrict__ tokens, const int* __restrict__ class_offsets, float* __restrict__ out) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int cls = blockIdx.x, t = threadIdx.x;
    const int k0 = class_offsets[cls], k1 = class_offsets[cls + 1];
    float s = 0.f;
    for (int k = k0; k < k1; ++k) s += tokens[static_cast<size_t>(k) * 256 + t];
    out[static_cast<size_t>(cls) * 256 + t] = s / static_cast<float>(k1 - k0);
}

}  // namespace sylph
