// Backward of the episodic TRAINING step for what the shipped meta-training configurations train (SURVEY.md 8f-4;
// configs/{COCO,LVISv1}-Detection/Meta-FCOS/Meta-FCOS-finetune.yaml: BACKBONE.FREEZE and FREEZE_BBOX_BRANCH on, CODE_GENERATOR.FREEZE and
// FREEZE_CLS_TOWER off): the classification loss -> {the class codes -> the code generator} and {the FCOS class tower}.
//   d loss_fcos_cls / d logits   sigmoid focal loss (fvcore sigmoid_focal_loss_jit, called at fcos_outputs.py:525-537)
//   d / d codes                  CondConvBasic (meta_fcos/head_utils.py:60-81): logits = <tower output, cls_conv> + cls_bias
//   d / d raw codes, post_norm, conv_scale, bias_scale     code_process_module (code_generator.py:833-875)
//   d / d per-shot codes         compute_code (:778-831), uniform 1 / SHOT weights
//   d / d support_set_cls_conv / support_set_cls_bias (+ F.normalize over the 49 positions) / support_set_shared_tower
//                                (conv3x3 + GroupNorm(32) + ReLU per layer; :648-688, 941-967)
//   d / d cls_tower              (fcos.py:72-122; second half of this file + wgrad3x3.cuh)
// CODE GENERATOR (first half): fp32 on CUDA cores -- 15-50 ROIs of 7x7 pixels are 0.9 GFLOP per convolution and do not fill a tensor-core
// tile pipeline.  Its forward is re-evaluated in fp32 from the pooled ROI features the forward pass left in the ROI planes (the tensor-core
// forward kernels keep no activations of the ROI tower).  Convolutions are GEMMs over an explicit im2col matrix whose K index is
// ci * 9 + tap, i.e. PyTorch's OIHW order: the caller's parameter tensors are read, and the gradients written, in place in the
// state_dict layout.
// CLASS TOWER (second half): the convolutions' weight and input gradients run on the tensor cores (wgrad3x3.cuh; the forward convolution
// kernel on transposed weights); the kernels here are the GroupNorm / ReLU backward over the planes and the glue around them.
// No atomics anywhere: every reduction has a fixed order, so gradients are bit-reproducible run to run.
#pragma once
#include "kernels_loss.cuh"

namespace sylph {

// ------------------------------------------------------------------------------------------------ generic fp32 GEMM
// C[M, N] (row-major, ldc) = (accumulate ? C : 0) + A'[M, K] * B'[K, N] (+ bias[n]);  A'(m, k) = A[m * sam + k * sak],
// B'(k, n) = B[k * sbk + n * sbn].  (16 TM) x (16 TN) x 16 tiles, 256 threads, TM x TN outputs per thread; any M, N, K.
// <4, 4> = 64 x 64 tiles; <2, 2> = 32 x 32 tiles for problems whose 64 x 64 grid would leave most SMs idle.
constexpr int kSgK = 16;

template <int TM, int TN>
__global__ void __launch_bounds__(256)
sgemm_f32_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ B, long long sbk,
                 long long sbn, float* __restrict__ C, long long ldc, int M, int N, int K, const float* __restrict__ bias,
                 int accumulate) {
    constexpr int BM = 16 * TM, BN = 16 * TN;
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ float As[kSgK][BM + 4];
    __shared__ float Bs[kSgK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += kSgK) {
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int e = tid + i * 256;
            int m, k;
            if (sak == 1) { m = e >> 4; k = e & 15; } else { m = e % BM; k = e / BM; }
            const int gm = m0 + m, gk = k0 + k;
            As[k][m] = (gm < M && gk < K) ? __ldg(A + gm * sam + gk * sak) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < TN; ++i) {
            const int e = tid + i * 256;
            int n, kb;
            if (sbn == 1) { n = e % BN; kb = e / BN; } else { kb = e & 15; n = e >> 4; }
            const int gn = n0 + n, gkb = k0 + kb;
            Bs[kb][n] = (gn < N && gkb < K) ? __ldg(B + gkb * sbk + gn * sbn) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kSgK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int gm = m0 + ty * TM + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int gn = n0 + tx * TN + j;
            if (gn >= N) continue;
            float v = acc[i][j];
            if (bias) v += bias[gn];
            float* dst = C + gm * ldc + gn;
            *dst = accumulate ? *dst + v : v;
        }
    }
}

// out[c] = sum over r < rows of in[r * ld + c], c < cols: one block per 32 columns, 8 row lanes, fixed order.
__global__ void __launch_bounds__(256)
colsum_f32_kernel(const float* __restrict__ in, long long ld, int rows, int cols, float* __restrict__ out) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (c < cols)
        for (int r = rl; r < rows; r += 8) s += in[r * ld + c];
    red[rl][lane] = s;
    __syncthreads();
    if (rl == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][lane];
        out[c] = t;
    }
}

// ------------------------------------------------------------------------------------------------ ROI planes <-> fp32
// Interior 7x7 pixels of the pooled ROI planes (9x9 zero-bordered, `roi_stride` rows per ROI, rows [256 hi | 256 lo] in
// exact mode) -> X[(roi * 49 + p) * 256 + c] fp32.
__global__ void __launch_bounds__(256)
roi_planes_to_f32_kernel(const __half* __restrict__ planes, float* __restrict__ X, int n_rois, int split, int roi_stride) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const long long total = static_cast<long long>(n_rois) * 49 * 256;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i & 255);
        const long long rp = i >> 8;
        const int p = static_cast<int>(rp % 49);
        const int r = static_cast<int>(rp / 49);
        const size_t row = static_cast<size_t>(r) * roi_stride + (p / 7 + 1) * 9 + (p % 7 + 1);
        X[i] = split ? __half2float(planes[row * 512 + c]) + __half2float(planes[row * 512 + 256 + c])
                     : __half2float(planes[row * 256 + c]);
    }
}

// col[(roi * 49 + p) * 2304 + ci * 9 + tap] = X[roi, y + dy, x + dx, ci] (zero outside the 7x7 plane), tap = (dy+1)*3 + dx+1.
__global__ void __launch_bounds__(256)
im2col_roi_kernel(const float* __restrict__ X, float* __restrict__ col, int n_rois) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const long long total = static_cast<long long>(n_rois) * 49 * 2304;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i % 2304);
        const long long rp = i / 2304;
        const int p = static_cast<int>(rp % 49);
        const long long r = rp / 49;
        const int ci = k / 9, tap = k - ci * 9;
        const int yy = p / 7 + tap / 3 - 1, xx = p % 7 + tap % 3 - 1;
        col[i] = (yy >= 0 && yy < 7 && xx >= 0 && xx < 7) ? X[(r * 49 + yy * 7 + xx) * 256 + ci] : 0.f;
    }
}

// dX[roi, y, x, ci] = sum over taps of dcol[(roi, y - dy, x - dx), ci * 9 + tap] (the transpose of im2col, as a gather).
__global__ void __launch_bounds__(256)
col2im_roi_kernel(const float* __restrict__ dcol, float* __restrict__ dX, int n_rois) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const long long total = static_cast<long long>(n_rois) * 49 * 256;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i & 255);
        const long long rp = i >> 8;
        const int p = static_cast<int>(rp % 49);
        const long long r = rp / 49;
        const int y = p / 7, x = p % 7;
        float s = 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int yo = y - (tap / 3 - 1), xo = x - (tap % 3 - 1);   // the output pixel whose tap reads (y, x)
            if (yo >= 0 && yo < 7 && xo >= 0 && xo < 7) s += dcol[(r * 49 + yo * 7 + xo) * 2304 + ci * 9 + tap];
        }
        dX[i] = s;
    }
}

// ------------------------------------------------------------------------------------------------ GroupNorm(32) + ReLU
__device__ __forceinline__ float group8_sum(float v) {   // sum over the 8 channels of a group = 8 consecutive lanes
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

// One block per ROI, thread = channel.  Y, Xout: [(roi * 49 + p) * 256 + c].  mean / rstd: [roi * 256 + c] (the group's value,
// replicated per channel).  Two-pass variance over the group's 8 x 49 values, eps 1e-5 (torch.nn.GroupNorm).
__global__ void __launch_bounds__(256)
roi_gn_relu_fwd_f32_kernel(const float* __restrict__ Y, const float* __restrict__ gamma, const float* __restrict__ beta,
                           float* __restrict__ Xout, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int r = blockIdx.x, c = threadIdx.x;
    const float* y = Y + static_cast<size_t>(r) * 49 * 256 + c;
    float s = 0.f;
    for (int p = 0; p < 49; ++p) s += y[p * 256];
    const float mean = group8_sum(s) / 392.f;
    float v = 0.f;
    for (int p = 0; p < 49; ++p) { const float d = y[p * 256] - mean; v = fmaf(d, d, v); }
    const float rstd = 1.f / sqrtf(group8_sum(v) / 392.f + 1e-5f);
    const float g = gamma[c], b = beta[c];
    float* xo = Xout + static_cast<size_t>(r) * 49 * 256 + c;
    for (int p = 0; p < 49; ++p) xo[p * 256] = fmaxf(fmaf((y[p * 256] - mean) * rstd, g, b), 0.f);
    mean_out[r * 256 + c] = mean;
    rstd_out[r * 256 + c] = rstd;
}

// dXn: gradient with respect to the ReLU output.  dY receives the gradient with respect to the convolution output;
// dgamma_part / dbeta_part [roi * 256 + c] are summed over ROIs by colsum_f32_kernel.  dY may alias dXn.
__global__ void __launch_bounds__(256)
roi_gn_relu_bwd_f32_kernel(const float* dXn, const float* __restrict__ Y, const float* __restrict__ mean_in,
                           const float* __restrict__ rstd_in, const float* __restrict__ gamma, const float* __restrict__ beta,
                           float* dY, float* __restrict__ dgamma_part, float* __restrict__ dbeta_part) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int r = blockIdx.x, c = threadIdx.x;
    const size_t base = static_cast<size_t>(r) * 49 * 256 + c;
    const float mean = mean_in[r * 256 + c], rstd = rstd_in[r * 256 + c], g = gamma[c], b = beta[c];
    float dg = 0.f, db = 0.f, sa = 0.f, sb = 0.f;
    for (int p = 0; p < 49; ++p) {
        const float yh = (Y[base + p * 256] - mean) * rstd;
        const float dz = fmaf(yh, g, b) > 0.f ? dXn[base + p * 256] : 0.f;
        dg = fmaf(dz, yh, dg);
        db += dz;
        sa = fmaf(dz, g, sa);
        sb = fmaf(dz * g, yh, sb);
    }
    const float a = group8_sum(sa) / 392.f, bb = group8_sum(sb) / 392.f;
    for (int p = 0; p < 49; ++p) {
        const float yh = (Y[base + p * 256] - mean) * rstd;
        const float dz = fmaf(yh, g, b) > 0.f ? dXn[base + p * 256] : 0.f;
        dY[base + p * 256] = rstd * (dz * g - a - yh * bb);
    }
    dgamma_part[r * 256 + c] = dg;
    dbeta_part[r * 256 + c] = db;
}

// ------------------------------------------------------------------------------------------------ code processing
// Backward of code_process_module over all classes (code_generator.py:833-875): one block, classes in order, so the
// parameter gradients (post_norm weight / bias, conv_scale, bias_scale) have a fixed summation order.
//   raw[cls * 257 + k]: codes in front of the normalisation (k < 256 cls_conv, k == 256 cls_bias)
//   gout[cls * 257 + k]: gradient with respect to the FINAL codes;   draw: gradient with respect to raw.
// scalars_out[0] = d conv_scale, scalars_out[1] = d bias_scale.
__global__ void __launch_bounds__(256)
normalize_codes_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ gout, int n_classes,
                           const float* __restrict__ gn_w, const float* __restrict__ gn_b, int post_norm, int l2_norm,
                           const float* __restrict__ conv_scale, const float* __restrict__ bias_scale,
                           float* __restrict__ draw, float* __restrict__ d_gn_w, float* __restrict__ d_gn_b,
                           float* __restrict__ scalars_out) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ float red[8];
    __shared__ float tot_s;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    auto block_sum = [&](float v) -> float {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();                       // red / tot_s of the previous call have been read by everyone
        if (lane == 0) red[warp] = v;
        __syncthreads();
        if (t == 0) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += red[i];
            tot_s = s;
        }
        __syncthreads();
        return tot_s;
    };
    const float cs = conv_scale ? conv_scale[0] : 1.f, bs = bias_scale ? bias_scale[0] : 1.f;
    const float gw = post_norm ? gn_w[t] : 1.f, gb = post_norm ? gn_b[t] : 0.f;
    float acc_gw = 0.f, acc_gb = 0.f, acc_cs = 0.f, acc_bs = 0.f;
    for (int cls = 0; cls < n_classes; ++cls) {
        const float x = raw[static_cast<size_t>(cls) * 257 + t];
        float gh = x, rstd8 = 1.f, g = x;
        if (post_norm) {
            const float mean = group8_sum(x) / 8.f;
            const float d = x - mean;
            rstd8 = rsqrtf(group8_sum(d * d) / 8.f + 1e-5f);
            gh = d * rstd8;
            g = fmaf(gh, gw, gb);
        }
        float nrm = 1.f, l = g;
        if (l2_norm) {
            nrm = fmaxf(sqrtf(block_sum(g * g)), 1e-12f);
            l = g / nrm;
        }
        const float go = gout[static_cast<size_t>(cls) * 257 + t];
        acc_cs = fmaf(go, l, acc_cs);                      // summed over the block at the end
        const float dl = go * cs;
        float dg = dl;
        if (l2_norm) dg = (dl - l * block_sum(l * dl)) / nrm;
        float dx = dg;
        if (post_norm) {
            acc_gw = fmaf(dg, gh, acc_gw);
            acc_gb += dg;
            const float dgh = dg * gw;
            const float m1 = group8_sum(dgh) / 8.f, m2 = group8_sum(dgh * gh) / 8.f;
            dx = rstd8 * (dgh - m1 - gh * m2);
        }
        draw[static_cast<size_t>(cls) * 257 + t] = dx;
        if (t == 0) {
            const float gbias = gout[static_cast<size_t>(cls) * 257 + 256];
            acc_bs = fmaf(gbias, raw[static_cast<size_t>(cls) * 257 + 256], acc_bs);
            draw[static_cast<size_t>(cls) * 257 + 256] = gbias * bs;
        }
    }
    if (post_norm && d_gn_w) { d_gn_w[t] = acc_gw; d_gn_b[t] = acc_gb; }
    const float dcs = block_sum(acc_cs);
    if (t == 0) { scalars_out[0] = dcs; scalars_out[1] = acc_bs; }
}

// ------------------------------------------------------------------------------------------------ class mean + pools
// One block per ROI.  draw[cls * 257 + k]: gradient with respect to the raw class codes; the class of ROI r is
// roi_class[r] with K = shots_of_class.  Writes the gradient with respect to the cls convolution's output
// dYc[(r * 49 + p) * 256 + co] = draw / (K * 49) (uniform shot weights, global average pool) and, with a bias layer, the gradient
// dv[r * 49 + p] with respect to the bias convolution's output v (49 values per ROI):
//   BIAS_L2_NORM: u = v / max(|v|, 1e-12), b = mean(u)  ->  dv = (du - u <u, du>) / |v|,  du = draw_bias / (K * 49).
__global__ void __launch_bounds__(256)
shot_code_bwd_kernel(const float* __restrict__ draw, const int* __restrict__ roi_class, const int* __restrict__ class_off,
                     const float* __restrict__ v_bias, int bias_layer, int bias_l2_norm, float* __restrict__ dYc,
                     float* __restrict__ dv) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ float sv[49];
    __shared__ float s_nrm, s_dot;
    const int r = blockIdx.x, t = threadIdx.x;
    const int cls = roi_class[r];
    const float inv = 1.f / (static_cast<float>(class_off[cls + 1] - class_off[cls]) * 49.f);
    const float g = draw[static_cast<size_t>(cls) * 257 + t] * inv;
    float* dst = dYc + static_cast<size_t>(r) * 49 * 256 + t;
    for (int p = 0; p < 49; ++p) dst[p * 256] = g;
    if (!bias_layer) return;
    const float du = draw[static_cast<size_t>(cls) * 257 + 256] * inv;
    if (!bias_l2_norm) {
        if (t < 49) dv[r * 49 + t] = du;
        return;
    }
    if (t < 49) sv[t] = v_bias[r * 49 + t];
    __syncthreads();
    if (t == 0) {
        float ss = 0.f;
        for (int p = 0; p < 49; ++p) ss = fmaf(sv[p], sv[p], ss);
        const float nrm = fmaxf(sqrtf(ss), 1e-12f);
        float dot = 0.f;
        for (int p = 0; p < 49; ++p) dot += (sv[p] / nrm) * du;
        s_nrm = nrm;
        s_dot = dot;
    }
    __syncthreads();
    if (t < 49) {
        const float u = sv[t] / s_nrm;
        dv[r * 49 + t] = (du - u * s_dot) / s_nrm;
    }
}

// ------------------------------------------------------------------------------------------------ classification loss
// d loss_fcos_cls / d codes.  Locations in the level-first order of fcos_targets_loss_kernel; `labels` are that kernel's
// labels_out.  Per location and class: g = focal'(logit, target) * upstream / num_pos_avg; the block accumulates
// dcode[c][k] += g * tower[row][k] (thread = channel k) and dbias[c] += g over its share of the locations, then writes
// partials[block][c][257]; fcos_code_grad_reduce_kernel sums the blocks in order.
// d sigmoid_focal_loss / d logit (fvcore sigmoid_focal_loss_jit): with p = sigmoid(v), ce = softplus(-v) (target 1) or softplus(v)
// (target 0):  target 1: alpha (1 - p)^gamma (-gamma p ce - (1 - p));  target 0: (1 - alpha) p^gamma (gamma (1 - p) ce + p).
__device__ __forceinline__ float focal_grad(float v, bool pos, float alpha, float gamma) {
    const float p = 1.f / (1.f + expf(-v));
    const float ce = bce_with_logits(v, pos ? 1.f : 0.f);
    float gr;
    if (pos) {
        const float om = 1.f - p;
        gr = (gamma == 2.f ? om * om : powf(om, gamma)) * (-gamma * p * ce - om);
        if (alpha >= 0.f) gr *= alpha;
    } else {
        gr = (gamma == 2.f ? p * p : powf(p, gamma)) * (gamma * (1.f - p) * ce + p);
        if (alpha >= 0.f) gr *= 1.f - alpha;
    }
    return gr;
}

constexpr int kClsBwdRows = 32;   // locations per staging round
constexpr int kClsBwdTile = 8;    // classes per pass over the locations

__global__ void __launch_bounds__(256)
fcos_cls_loss_bwd_kernel(const float* __restrict__ logits, int logit_stride, const __half* __restrict__ tower, int ld, int lo,
                         PyramidGeom pg, int n_images, const long long* __restrict__ labels,
                         const long long* __restrict__ support_targets, int n_classes, float alpha, float gamma,
                         const double* __restrict__ global_pos, int world, const float* __restrict__ upstream,
                         float* __restrict__ partials) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ float sg[kClsBwdRows][kClsBwdTile];
    __shared__ unsigned long long srow[kClsBwdRows];
    const int t = threadIdx.x;
    long long total = 0;
    long long lvl_start[6];
    for (int l = 0; l < 5; ++l) { lvl_start[l] = total; total += static_cast<long long>(n_images) * pg.lv[l].H * pg.lv[l].W; }
    lvl_start[5] = total;
    const float num_pos_avg = fmaxf(static_cast<float>(global_pos[0] / world), 1.0f);
    const float scale = (upstream ? upstream[0] : 1.f) / num_pos_avg;
    const long long rounds = (total + kClsBwdRows - 1) / kClsBwdRows;
    for (int c0 = 0; c0 < n_classes; c0 += kClsBwdTile) {
        const int nc = min(kClsBwdTile, n_classes - c0);
        float acc[kClsBwdTile];
#pragma unroll
        for (int c = 0; c < kClsBwdTile; ++c) acc[c] = 0.f;
        float accb = 0.f;                                           // threads t < nc: bias gradient of class c0 + t
        for (long long rd = blockIdx.x; rd < rounds; rd += gridDim.x) {
            const long long i0 = rd * kClsBwdRows;
            __syncthreads();                                        // the previous round's sg / srow have been consumed
            {
                const int rr = t / kClsBwdTile, cc = t % kClsBwdTile;   // 32 x 8 = 256 (row, class) pairs
                const long long i = i0 + rr;
                float gval = 0.f;
                unsigned long long row = 0;
                if (i < total) {
                    int l = 0;
                    while (i >= lvl_start[l + 1]) ++l;
                    const PlaneGeom g = pg.lv[l];
                    const long long j = i - lvl_start[l];
                    const int hw = g.H * g.W;
                    const int n = static_cast<int>(j / hw);
                    const int loc = static_cast<int>(j - static_cast<long long>(n) * hw);
                    const int y = loc / g.W, x = loc - y * g.W;
                    row = plane_row(g, n, y, x);
                    if (cc < nc) {
                        const float v = logits[row * logit_stride + c0 + cc];
                        gval = focal_grad(v, support_targets[c0 + cc] == labels[i], alpha, gamma) * scale;
                    }
                }
                sg[rr][cc] = gval;
                if (cc == 0) srow[rr] = (i < total) ? row : ~0ull;
            }
            __syncthreads();
            // rows past the end carry g = 0 and read row 0: no branch, so that the loads of several rows are in flight together
#pragma unroll 8
            for (int rr = 0; rr < kClsBwdRows; ++rr) {
                const unsigned long long row = srow[rr];
                const __half* px = tower + (row == ~0ull ? 0ull : row) * ld;
                float x = __half2float(px[t]);
                if (lo) x += __half2float(px[lo + t]);
#pragma unroll
                for (int c = 0; c < kClsBwdTile; ++c) acc[c] = fmaf(sg[rr][c], x, acc[c]);
                if (t < nc) accb += sg[rr][t];
            }
        }
        float* out = partials + (static_cast<size_t>(blockIdx.x) * n_classes + c0) * 257;
        for (int c = 0; c < nc; ++c) out[static_cast<size_t>(c) * 257 + t] = acc[c];
        if (t < nc) out[static_cast<size_t>(t) * 257 + 256] = accb;
    }
}

// out[e] = sum over blocks of partials[b * n_elems + e], fp64: one warp per element, lane l sums blocks l, l + 32, ... in order, then a fixed
// shuffle tree (deterministic for a given grid size).
__global__ void __launch_bounds__(256)
fcos_code_grad_reduce_kernel(const float* __restrict__ partials, int n_blocks, int n_elems, float* __restrict__ out) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (e >= n_elems) return;                       // whole warps leave together
    double s = 0.0;
    for (int b = lane; b < n_blocks; b += 32) s += static_cast<double>(partials[static_cast<size_t>(b) * n_elems + e]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[e] = static_cast<float>(s);
}

// ------------------------------------------------------------------------------------------------ weight refresh
// The code generator's weights after an optimiser step, prepared on the device in the layouts engine.cu prep_conv builds on
// the host: fp32 OIHW [co][ci][taps] -> rows [(tap * cout_pad + o)][kp] of fp16, kp = ci (fast) or [w_hi | w_hi | w_lo] = 3 ci
// (exact: hi = rn16(w), lo = rn16(w - hi)).  pooled != 0: the same 3x3 layer as ONE [co][taps * ci] GEMM (K index tap * ci +
// ch; engine.cu "__cg_cls_pooled").
__global__ void __launch_bounds__(256)
pack_oihw_weights_kernel(const float* __restrict__ w, __half* __restrict__ out, int co, int ci, int taps, int cout_pad, int split,
                         int pooled) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const long long total = static_cast<long long>(co) * ci * taps;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        // e enumerates the OUTPUT order (tap, o, i) so that stores are contiguous along i
        const int i = static_cast<int>(e % ci);
        const int o = static_cast<int>((e / ci) % co);
        const int t = static_cast<int>(e / (static_cast<long long>(ci) * co));
        const float v = w[(static_cast<size_t>(o) * ci + i) * taps + t];
        const __half hi = __float2half_rn(v);
        size_t row, k, kdim;
        if (pooled) { row = o; k = static_cast<size_t>(t) * ci + i; kdim = static_cast<size_t>(taps) * ci; }
        else { row = static_cast<size_t>(t) * cout_pad + o; k = i; kdim = ci; }
        __half* dst = out + row * (split ? 3 * kdim : kdim);
        dst[k] = hi;
        if (split) {
            dst[kdim + k] = hi;
            dst[2 * kdim + k] = __float2half_rn(v - __half2float(hi));
        }
    }
}

// [1][256][9] bias / weight-head convolution -> tap-major [9][256] fp32 (shot_code_kernel's layout).
__global__ void __launch_bounds__(256)
transpose_taps_kernel(const float* __restrict__ w, float* __restrict__ out, int ci, int taps) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ci * taps) return;
    const int t = e / ci, ch = e - t * ci;
    out[e] = w[ch * taps + t];
}

// ------------------------------------------------------------------------------------------------ FCOS class tower
// Backward of the class tower over the query pyramid (fcos.py:72-122: NUM_CLS_CONVS x [conv3x3 + GroupNorm(32) + ReLU], weights
// shared by the levels, GroupNorm statistics per (image, level) plane), SURVEY.md 8f-4 with FREEZE_CLS_TOWER: False.
//   d loss / d tower output        = sum over classes of d loss / d logit[c] * cls_conv[c]   (tower_out_grad_kernel)
//   per layer, last first:  ReLU + GroupNorm backward over the planes (three kernels below) -> dY as fp16 (hi | lo) planes,
//   weight gradient = wgrad3x3_kernel(dY, layer input), input gradient = the FORWARD convolution kernel on dY with the
//   transposed, tap-reversed weights.
// Gradients span many orders of magnitude, fp16 does not: every layer's dY is stored as s * dY with a power of two s picked on
// the device from the largest |dz| of the layer (gn_bwd_scale_kernel); consumers divide by s in fp32 (the next layer's first
// pass, the weight-gradient reduce).

// dX[row][k] = sum_c g[row][c] * codes[c][k] on interior pixels (fp32, [rows][256]); thread = channel k.
__global__ void __launch_bounds__(256)
tower_out_grad_kernel(const float* __restrict__ logits, int logit_stride, const float* __restrict__ codes, PyramidGeom pg,
                      int n_images, const long long* __restrict__ labels, const long long* __restrict__ support_targets,
                      int n_classes, float alpha, float gamma, const double* __restrict__ global_pos, int world,
                      const float* __restrict__ upstream, float* __restrict__ dX) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ float sg[kClsBwdRows][kClsBwdTile];
    __shared__ unsigned long long srow[kClsBwdRows];
    const int t = threadIdx.x;
    long long total = 0;
    long long lvl_start[6];
    for (int l = 0; l < 5; ++l) { lvl_start[l] = total; total += static_cast<long long>(n_images) * pg.lv[l].H * pg.lv[l].W; }
    lvl_start[5] = total;
    const float num_pos_avg = fmaxf(static_cast<float>(global_pos[0] / world), 1.0f);
    const float scale = (upstream ? upstream[0] : 1.f) / num_pos_avg;
    const long long rounds = (total + kClsBwdRows - 1) / kClsBwdRows;
    for (long long rd = blockIdx.x; rd < rounds; rd += gridDim.x) {
        const long long i0 = rd * kClsBwdRows;
        float acc[kClsBwdRows];
#pragma unroll
        for (int rr = 0; rr < kClsBwdRows; ++rr) acc[rr] = 0.f;
        for (int c0 = 0; c0 < n_classes; c0 += kClsBwdTile) {
            const int nc = min(kClsBwdTile, n_classes - c0);
            __syncthreads();
            {
                const int rr = t / kClsBwdTile, cc = t % kClsBwdTile;
                const long long i = i0 + rr;
                float gval = 0.f;
                unsigned long long row = ~0ull;
                if (i < total) {
                    int l = 0;
                    while (i >= lvl_start[l + 1]) ++l;
                    const PlaneGeom g = pg.lv[l];
                    const long long j = i - lvl_start[l];
                    const int hw = g.H * g.W;
                    const int n = static_cast<int>(j / hw);
                    const int loc = static_cast<int>(j - static_cast<long long>(n) * hw);
                    const int y = loc / g.W, x = loc - y * g.W;
                    row = plane_row(g, n, y, x);
                    if (cc < nc)
                        gval = focal_grad(logits[row * logit_stride + c0 + cc], support_targets[c0 + cc] == labels[i], alpha, gamma) * scale;
                }
                sg[rr][cc] = gval;
                if (cc == 0) srow[rr] = row;
            }
            __syncthreads();
            float w[kClsBwdTile];
#pragma unroll
            for (int c = 0; c < kClsBwdTile; ++c) w[c] = c < nc ? codes[static_cast<size_t>(c0 + c) * 257 + t] : 0.f;
#pragma unroll
            for (int rr = 0; rr < kClsBwdRows; ++rr)
#pragma unroll
                for (int c = 0; c < kClsBwdTile; ++c) acc[rr] = fmaf(sg[rr][c], w[c], acc[rr]);
        }
#pragma unroll
        for (int rr = 0; rr < kClsBwdRows; ++rr) {
            const unsigned long long row = srow[rr];
            if (row != ~0ull) dX[row * 256 + t] = acc[rr];
        }
    }
}

constexpr int kGnBwdPartial = 3 * 256;   // per tile: sum dz [256], sum dz * yhat [256], max |dz| (replicated over the row)

// Pass 1, one CTA per 128-row tile at a time (lane = group of 8 channels, warp w takes rows w, w + 8, ...): per channel
// sum dz and sum dz * yhat over the tile's interior pixels, dz = (dX / in_scale) where the forward's pre-activation is > 0.
__global__ void __launch_bounds__(256)
gn_bwd_partial_kernel(const float* __restrict__ dX, const float* __restrict__ in_scale, const float* __restrict__ Y,
                      const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                      const int* __restrict__ tile_seg, const Seg* __restrict__ segs, int n_tiles, float* __restrict__ partial) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ float red[8][2][256];
    __shared__ float redmax[8];
    const int g = threadIdx.x & 31, w = threadIdx.x >> 5;
    float ga[8], be[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { ga[j] = gamma[8 * g + j]; be[j] = beta[8 * g + j]; }
    const float inv_in = in_scale ? 1.f / in_scale[0] : 1.f;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int s = __ldg(tile_seg + tile);
        const Seg sgm = segs[s];
        const float2 st = *reinterpret_cast<const float2*>(stats + (static_cast<size_t>(s) * 32 + g) * 2);
        float a0[8], a1[8], mx = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
        for (int r = w; r < kBlockM; r += 8) {
            const int row = tile * kBlockM + r;
            if (!row_is_interior(sgm, row)) continue;
            const float4* py = reinterpret_cast<const float4*>(Y + static_cast<size_t>(row) * 256) + 2 * g;
            const float4* pd = reinterpret_cast<const float4*>(dX + static_cast<size_t>(row) * 256) + 2 * g;
            const float4 ya = __ldg(py), yb = __ldg(py + 1), da = __ldg(pd), db = __ldg(pd + 1);
            const float yv[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
            const float dv[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float yh = (yv[j] - st.x) * st.y;
                const float dz = ((yv[j] - st.x) * st.y * ga[j] + be[j]) > 0.f ? dv[j] * inv_in : 0.f;   // the forward's expression
                a0[j] += dz;
                a1[j] = fmaf(dz, yh, a1[j]);
                mx = fmaxf(mx, fabsf(dz));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        __syncthreads();                      // red of the previous tile has been read
#pragma unroll
        for (int j = 0; j < 8; ++j) { red[w][0][8 * g + j] = a0[j]; red[w][1][8 * g + j] = a1[j]; }
        if (g == 0) redmax[w] = mx;
        __syncthreads();
        const int c = threadIdx.x;
        float s0 = 0.f, s1 = 0.f, m = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { s0 += red[k][0][c]; s1 += red[k][1][c]; m = fmaxf(m, redmax[k]); }
        float* dst = partial + static_cast<size_t>(tile) * kGnBwdPartial;
        dst[c] = s0;
        dst[256 + c] = s1;
        dst[512 + c] = m;
    }
}

// Per plane (one block, thread = channel): tile partials summed in order (fp64) -> seg_sums[seg][2][256] (sum dz, sum dz yhat) and
// the two group means of GroupNorm's backward, ab[(seg * 32 + g) * 2] = mean(dz gamma), [+1] = mean(dz gamma yhat) over the
// group's 8 x H x W values; seg_max[seg] = max |dz| of the plane.
__global__ void __launch_bounds__(1024)
gn_bwd_finalize_kernel(const float* __restrict__ partial, const Seg* __restrict__ segs, const float* __restrict__ gamma,
                       float* __restrict__ seg_sums, float* __restrict__ ab, float* __restrict__ seg_max) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    // 1024 threads: channel c = thread & 255, quarter q = thread >> 8 takes the plane's tiles t0 + q, t0 + q + 4, ...; the four partial sums
    // are combined in the order q = 0..3 (a p3 plane has 136 tiles: one thread walking them all was latency-bound)
    __shared__ double rs0[4][256], rs1[4][256];
    __shared__ float rm[4][256];
    const int s = blockIdx.x, c = threadIdx.x & 255, q = threadIdx.x >> 8;
    const Seg sg = segs[s];
    const int t0 = sg.row0 / kBlockM, t1 = (sg.row0 + sg.nrows + kBlockM - 1) / kBlockM;
    double s0 = 0.0, s1 = 0.0;
    float m = 0.f;
    for (int t = t0 + q; t < t1; t += 4) {
        const float* src = partial + static_cast<size_t>(t) * kGnBwdPartial;
        s0 += static_cast<double>(src[c]);
        s1 += static_cast<double>(src[256 + c]);
        m = fmaxf(m, src[512 + c]);
    }
    rs0[q][c] = s0; rs1[q][c] = s1; rm[q][c] = m;
    __syncthreads();
    if (q != 0) return;                              // warps 0..7 go on, complete
    s0 = rs0[0][c] + rs0[1][c] + rs0[2][c] + rs0[3][c];
    s1 = rs1[0][c] + rs1[1][c] + rs1[2][c] + rs1[3][c];
    m = fmaxf(fmaxf(rm[0][c], rm[1][c]), fmaxf(rm[2][c], rm[3][c]));
    seg_sums[(static_cast<size_t>(s) * 2) * 256 + c] = static_cast<float>(s0);
    seg_sums[(static_cast<size_t>(s) * 2 + 1) * 256 + c] = static_cast<float>(s1);
    const double gm = static_cast<double>(gamma[c]);
    double a = gm * s0, b = gm * s1;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const double cnt = static_cast<double>(sg.H) * sg.W * 8.0;
    if ((c & 7) == 0) {
        ab[(static_cast<size_t>(s) * 32 + (c >> 3)) * 2] = static_cast<float>(a / cnt);
        ab[(static_cast<size_t>(s) * 32 + (c >> 3)) * 2 + 1] = static_cast<float>(b / cnt);
    }
    if (c == 0) seg_max[s] = m;
}

// scale[0] = the power of two s with s * max|dz| * max(rstd) * max|gamma| ~ 256 (1 when the layer's gradient is all zero),
// scale[1] = 1 / s.  One block.
__global__ void __launch_bounds__(256)
gn_bwd_scale_kernel(const float* __restrict__ seg_max, const float* __restrict__ stats, int n_segs, const float* __restrict__ gamma,
                    float* __restrict__ scale) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ float red[256];
    const int t = threadIdx.x;
    float m = 0.f, r = 0.f;
    for (int i = t; i < n_segs; i += 256) m = fmaxf(m, seg_max[i]);
    for (int i = t; i < n_segs * 32; i += 256) r = fmaxf(r, stats[2 * i + 1]);
    const float gmx = fabsf(gamma[t]);
    auto block_max = [&](float v) -> float {
        __syncthreads();
        red[t] = v;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (t < o) red[t] = fmaxf(red[t], red[t + o]);
            __syncthreads();
        }
        return red[0];
    };
    const float mm = block_max(m), rr = block_max(r), gg = block_max(gmx);
    if (t == 0) {
        const float bound = mm * rr * gg;
        float s = 1.f;
        if (bound > 0.f && isfinite(bound)) {
            int e = static_cast<int>(floorf(log2f(256.f / bound)));
            e = max(-40, min(60, e));
            s = exp2f(static_cast<float>(e));
        }
        scale[0] = s;
        scale[1] = 1.f / s;
    }
}

// Pass 2: dY = rstd * (dz gamma - a - yhat b) on interior pixels, stored as scale * dY in fp16 (hi | lo) planes (0 elsewhere);
// per tile the channel sums of the UNscaled dY (the convolution bias gradient) go to bias_partial[tile][256].
__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const float* __restrict__ dX, const float* __restrict__ in_scale, const float* __restrict__ Y,
                    const float* __restrict__ stats, const float* __restrict__ ab, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ out_scale, const int* __restrict__ tile_seg,
                    const Seg* __restrict__ segs, int n_tiles, __half* __restrict__ dY, int split, float* __restrict__ bias_partial) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ float red[8][256];
    const int g = threadIdx.x & 31, w = threadIdx.x >> 5;
    float ga[8], be[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { ga[j] = gamma[8 * g + j]; be[j] = beta[8 * g + j]; }
    const float inv_in = in_scale ? 1.f / in_scale[0] : 1.f;
    const float osc = out_scale[0];
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int s = __ldg(tile_seg + tile);
        const Seg sgm = segs[s];
        const float2 st = *reinterpret_cast<const float2*>(stats + (static_cast<size_t>(s) * 32 + g) * 2);
        const float2 abv = *reinterpret_cast<const float2*>(ab + (static_cast<size_t>(s) * 32 + g) * 2);
        float bs[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bs[j] = 0.f;
        for (int r = w; r < kBlockM; r += 8) {
            const int row = tile * kBlockM + r;
            uint4 o = make_uint4(0u, 0u, 0u, 0u), ol = make_uint4(0u, 0u, 0u, 0u);
            if (row_is_interior(sgm, row)) {
                const float4* py = reinterpret_cast<const float4*>(Y + static_cast<size_t>(row) * 256) + 2 * g;
                const float4* pd = reinterpret_cast<const float4*>(dX + static_cast<size_t>(row) * 256) + 2 * g;
                const float4 ya = __ldg(py), yb = __ldg(py + 1), da = __ldg(pd), db = __ldg(pd + 1);
                const float yv[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
                const float dv[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float yh = (yv[j] - st.x) * st.y;
                    const float dz = ((yv[j] - st.x) * st.y * ga[j] + be[j]) > 0.f ? dv[j] * inv_in : 0.f;
                    const float dy = st.y * (dz * ga[j] - abv.x - yh * abv.y);
                    bs[j] += dy;
                    v[j] = dy * osc;
                }
                if (split) split8(v, false, true, o, ol);
                else o = pack8(v, false, true);
            }
            if (split) {
                reinterpret_cast<uint4*>(dY + static_cast<size_t>(row) * 512)[g] = o;
                reinterpret_cast<uint4*>(dY + static_cast<size_t>(row) * 512 + 256)[g] = ol;
            } else {
                reinterpret_cast<uint4*>(dY + static_cast<size_t>(row) * 256)[g] = o;
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) red[w][8 * g + j] = bs[j];
        __syncthreads();
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
        bias_partial[static_cast<size_t>(tile) * 256 + threadIdx.x] = t;
    }
}

// d gamma[c] = sum over planes of sum dz yhat, d beta[c] = sum over planes of sum dz, d conv bias[c] = sum over tiles of sum dY:
// fixed order, fp64.  grid 3 (which = blockIdx.x), thread = channel.
__global__ void __launch_bounds__(1024)
tower_param_grad_reduce_kernel(const float* __restrict__ seg_sums, int n_segs, const float* __restrict__ bias_partial, int n_tiles,
                               float* __restrict__ d_gamma, float* __restrict__ d_beta, float* __restrict__ d_bias) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    // 1024 threads: channel c = thread & 255, quarter q = thread >> 8 sums items q, q + 4, ...; quarters combined in the order 0..3
    __shared__ double red[4][256];
    const int c = threadIdx.x & 255, q = threadIdx.x >> 8;
    double s = 0.0;
    if (blockIdx.x == 0) {
        for (int i = q; i < n_segs; i += 4) s += static_cast<double>(seg_sums[(static_cast<size_t>(i) * 2 + 1) * 256 + c]);
    } else if (blockIdx.x == 1) {
        for (int i = q; i < n_segs; i += 4) s += static_cast<double>(seg_sums[(static_cast<size_t>(i) * 2) * 256 + c]);
    } else {
        for (int t = q; t < n_tiles; t += 4) s += static_cast<double>(bias_partial[static_cast<size_t>(t) * 256 + c]);
    }
    red[q][c] = s;
    __syncthreads();
    if (q != 0) return;
    const float v = static_cast<float>(red[0][c] + red[1][c] + red[2][c] + red[3][c]);
    if (blockIdx.x == 0) d_gamma[c] = v;
    else if (blockIdx.x == 1) d_beta[c] = v;
    else d_bias[c] = v;
}

// fp32 OIHW [co][ci][9] -> the operand rows of the INPUT-gradient convolution: Wt[o' = ci][i' = co][tap] = W[co][ci][8 - tap]
// in the layout of pack_oihw_weights_kernel (tap-major rows [(tap * 256 + o')][k], [hi | hi | lo] in exact mode).
__global__ void __launch_bounds__(256)
pack_oihw_weights_transposed_kernel(const float* __restrict__ w, __half* __restrict__ out, int split) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int total = 256 * 256 * 9;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int i = e & 255;            // input channel of the gradient convolution = co of the forward weights
        const int o = (e >> 8) & 255;     // output channel = ci of the forward weights
        const int t = e >> 16;
        const float v = w[(static_cast<size_t>(i) * 256 + o) * 9 + (8 - t)];
        const __half hi = __float2half_rn(v);
        __half* dst = out + (static_cast<size_t>(t) * 256 + o) * (split ? 768 : 256);
        dst[i] = hi;
        if (split) {
            dst[256 + i] = hi;
            dst[512 + i] = __float2half_rn(v - __half2float(hi));
        }
    }
}

// Scaled weight-gradient reduce: dW = (sum over splits of partial) * inv_scale[1] (wgrad_reduce_kernel with the layer's scale).
__global__ void __launch_bounds__(256)
wgrad_reduce_scaled_kernel(const float* __restrict__ partial, int splits, const float* __restrict__ scale, float* __restrict__ dW) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 9 * 256 * 256) return;
    const int ci = e & 255, co = (e >> 8) & 255, tap = e >> 16;
    double s = 0.0;
    for (int k = 0; k < splits; ++k) s += static_cast<double>(partial[(static_cast<size_t>(k * 9 + tap) * 256 + co) * 256 + ci]);
    dW[(co * 256 + ci) * 9 + tap] = static_cast<float>(s * static_cast<double>(scale[1]));
}

}  // namespace sylph
