// Episodic TRAINING forward of the FCOS outputs: ground-truth assignment of every location and the
// three loss sums, in one pass over the head outputs -- no (K, N_gt) intermediates, no nonzero() / .item() syncs.
// Replaces FCOSOutputs._get_ground_truth / compute_targets_for_locations / get_sample_region / losses /
// fcos_losses_episodic_learning (sylph/modeling/meta_fcos/fcos_outputs.py:140-349, 351-637), compute_ctrness_targets
// (:52-61), IOULoss (sylph/modeling/meta_fcos/iou_loss.py:26-86) and fvcore's sigmoid_focal_loss_jit.
// Integer outputs (labels, target_inds) and the regression targets are bit-exact (single fp32 subtractions / a division
// by a power of two); loss terms are evaluated in fp32 like the reference and accumulated in fp64 in a fixed order.
#pragma once
#include "kernels_detect.cuh"

namespace sylph {

constexpr float kFcosInf = 100000000.f;          // INF, fcos_outputs.py:29
constexpr long long kBackgroundId = 100000;      // back_ground_id, fcos_outputs.py:101
constexpr int kLossSums = 5;                     // focal, positives, sum of centre-ness targets, weighted loc loss, ctr BCE

struct LossParams {
    PyramidGeom pg;
    int n_images;
    int n_classes;
    int logit_stride;
    int stride[5];
    float level_scale[5];
    float soi_lo[5], soi_hi[5];   // sizes_of_interest, fcos_outputs.py:95-99
    float radius[5];              // strides[l] * POS_RADIUS, fcos_outputs.py:219
    int center_sample;
    float alpha, gamma;
    int loc_loss_type;            // 0 iou, 1 linear_iou, 2 giou
};

__device__ __forceinline__ float bce_with_logits(float x, float t) {
    // torch binary_cross_entropy_with_logits: (1 - t) * x + max(-x, 0) + log1p(exp(-|x|))
    return (1.f - t) * x + (fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x))));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One thread per (level, image, y, x) in the reference's level-first order (L, N, H, W; fcos_outputs.py:125-138).
// partials[blockIdx.x][kLossSums]; optional per-location outputs labels / target_inds (int64), reg_targets (fp32 x 4).
__global__ void __launch_bounds__(256)
fcos_targets_loss_kernel(const float* __restrict__ logits, const float* __restrict__ pred, LossParams p,
                         const float* __restrict__ gt_boxes, const long long* __restrict__ gt_classes,
                         const int* __restrict__ gt_offsets, const long long* __restrict__ support_targets,
                         double* __restrict__ partials, long long* __restrict__ labels_out,
                         long long* __restrict__ inds_out, float* __restrict__ reg_out) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    long long total = 0;
    long long lvl_start[6];
    for (int l = 0; l < 5; ++l) { lvl_start[l] = total; total += static_cast<long long>(p.n_images) * p.pg.lv[l].H * p.pg.lv[l].W; }
    lvl_start[5] = total;
    double acc[kLossSums] = {0, 0, 0, 0, 0};
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        int l = 0;
        while (i >= lvl_start[l + 1]) ++l;
        const PlaneGeom g = p.pg.lv[l];
        const long long j = i - lvl_start[l];
        const int hw = g.H * g.W;
        const int n = static_cast<int>(j / hw);
        const int loc = static_cast<int>(j - static_cast<long long>(n) * hw);
        const int y = loc / g.W, x = loc - y * g.W;
        const float xs = static_cast<float>(x * p.stride[l] + p.stride[l] / 2);   // compute_locations, fcos.py:270-282
        const float ys = static_cast<float>(y * p.stride[l] + p.stride[l] / 2);
        // ---- compute_targets_for_locations (fcos_outputs.py:250-349)
        const int g0 = gt_offsets[n], g1 = gt_offsets[n + 1];
        long long label = kBackgroundId, ind = -1;
        float rt[4] = {0.f, 0.f, 0.f, 0.f};
        if (g1 > g0) {
            // get_sample_region quirk (:213): a first box centred on x == 0 disables every location of the image
            const float cx_first = (gt_boxes[4 * g0] + gt_boxes[4 * g0 + 2]) * 0.5f;
            const bool sample_off = p.center_sample && cx_first == 0.f;
            float best = kFcosInf;
            int best_j = g0;
            for (int k = g0; k < g1; ++k) {
                const float bx0 = gt_boxes[4 * k], by0 = gt_boxes[4 * k + 1], bx1 = gt_boxes[4 * k + 2], by1 = gt_boxes[4 * k + 3];
                const float lft = xs - bx0, top = ys - by0, rgt = bx1 - xs, bot = by1 - ys;
                bool inside;
                if (p.center_sample) {
                    const float cx = (bx0 + bx1) * 0.5f, cy = (by0 + by1) * 0.5f, r = p.radius[l];
                    const float xmin = cx - r, ymin = cy - r, xmax = cx + r, ymax = cy + r;
                    const float c0 = xmin > bx0 ? xmin : bx0, c1 = ymin > by0 ? ymin : by0;
                    const float c2 = xmax > bx1 ? bx1 : xmax, c3 = ymax > by1 ? by1 : ymax;
                    inside = !sample_off && fminf(fminf(xs - c0, ys - c1), fminf(c2 - xs, c3 - ys)) > 0.f;
                } else {
                    inside = fminf(fminf(lft, top), fminf(rgt, bot)) > 0.f;
                }
                const float mx = fmaxf(fmaxf(lft, top), fmaxf(rgt, bot));
                const bool cared = mx >= p.soi_lo[l] && mx <= p.soi_hi[l];
                const float area = __fmul_rn(bx1 - bx0, by1 - by0);                 // Boxes.area()
                const float a = (inside && cared) ? area : kFcosInf;
                if (a < best) { best = a; best_j = k; }                            // first minimum (torch.min on CPU)
            }
            const float s = static_cast<float>(p.stride[l]);
            rt[0] = (xs - gt_boxes[4 * best_j]) / s;
            rt[1] = (ys - gt_boxes[4 * best_j + 1]) / s;
            rt[2] = (gt_boxes[4 * best_j + 2] - xs) / s;
            rt[3] = (gt_boxes[4 * best_j + 3] - ys) / s;
            ind = best_j;                                                          // + num_targets of earlier images
            label = best == kFcosInf ? kBackgroundId : gt_classes[best_j];
        }
        if (labels_out) labels_out[i] = label;
        if (inds_out) inds_out[i] = ind;
        if (reg_out) *reinterpret_cast<float4*>(reg_out + 4 * i) = make_float4(rt[0], rt[1], rt[2], rt[3]);
        // ---- focal loss over the episode's classes (fcos_losses_episodic_learning :525-537)
        const size_t row = plane_row(g, n, y, x);
        const float* lg = logits + row * p.logit_stride;
        for (int c = 0; c < p.n_classes; ++c) {
            const float v = lg[c];
            const float t = support_targets[c] == label ? 1.f : 0.f;
            const float pr = 1.f / (1.f + expf(-v));
            const float ce = bce_with_logits(v, t);
            const float p_t = __fadd_rn(__fmul_rn(pr, t), __fmul_rn(1.f - pr, 1.f - t));
            const float om = 1.f - p_t;
            float loss = __fmul_rn(ce, p.gamma == 2.f ? __fmul_rn(om, om) : powf(om, p.gamma));
            if (p.alpha >= 0.f) loss = __fmul_rn(__fadd_rn(__fmul_rn(p.alpha, t), __fmul_rn(1.f - p.alpha, 1.f - t)), loss);
            acc[0] += static_cast<double>(loss);
        }
        // ---- positives: centre-ness target, IoU loss, centre-ness BCE (:550-592)
        if (label != kBackgroundId && pred == nullptr) {
            acc[1] += 1.0;                                                     // box branch skipped: positives only
        } else if (label != kBackgroundId) {
            const float lr_min = fminf(rt[0], rt[2]), lr_max = fmaxf(rt[0], rt[2]);
            const float tb_min = fminf(rt[1], rt[3]), tb_max = fmaxf(rt[1], rt[3]);
            const float ctr_t = sqrtf(__fmul_rn(lr_min / lr_max, tb_min / tb_max));
            const float4 raw = *reinterpret_cast<const float4*>(pred + row * 16);
            const float sc = p.level_scale[l];
            const float pl = fmaxf(raw.x * sc, 0.f), pt = fmaxf(raw.y * sc, 0.f), prr = fmaxf(raw.z * sc, 0.f), pb = fmaxf(raw.w * sc, 0.f);
            const float t_area = __fmul_rn(rt[0] + rt[2], rt[1] + rt[3]);
            const float p_area = __fmul_rn(pl + prr, pt + pb);
            const float wi = fminf(pl, rt[0]) + fminf(prr, rt[2]);
            const float hi = fminf(pb, rt[3]) + fminf(pt, rt[1]);
            const float gw = fmaxf(pl, rt[0]) + fmaxf(prr, rt[2]);
            const float gh = fmaxf(pb, rt[3]) + fmaxf(pt, rt[1]);
            const float ac = __fmul_rn(gw, gh);
            const float inter = __fmul_rn(wi, hi);
            const float uni = __fadd_rn(t_area, p_area) - inter;
            const float iou = (inter + 1.f) / (uni + 1.f);
            float per;
            if (p.loc_loss_type == 0) per = -logf(iou);
            else if (p.loc_loss_type == 1) per = 1.f - iou;
            else per = 1.f - (iou - (ac - uni) / ac);
            acc[1] += 1.0;
            acc[2] += static_cast<double>(ctr_t);
            acc[3] += static_cast<double>(__fmul_rn(per, ctr_t));
            acc[4] += static_cast<double>(bce_with_logits(pred[row * 16 + 4], ctr_t));
        }
    }
    __shared__ double red[8][kLossSums];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kLossSums; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < kLossSums) {
        double v = 0;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        partials[static_cast<size_t>(blockIdx.x) * kLossSums + threadIdx.x] = v;
    }
}

// Fixed-order sum of the per-block partials -> sums[kLossSums] (one CTA, deterministic for a given grid size).
__global__ void __launch_bounds__(256)
fcos_loss_reduce_kernel(const double* __restrict__ partials, int n_blocks, double* __restrict__ sums) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ double red[8][kLossSums];
    double acc[kLossSums] = {0, 0, 0, 0, 0};
    for (int b = threadIdx.x; b < n_blocks; b += blockDim.x)
        for (int k = 0; k < kLossSums; ++k) acc[k] += partials[static_cast<size_t>(b) * kLossSums + k];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kLossSums; ++k) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < kLossSums) {
        double v = 0;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        sums[threadIdx.x] = v;
    }
}

// losses[0..2] = loss_fcos_cls, loss_fcos_loc, loss_fcos_ctr (fcos_losses_episodic_learning :520-592):
//   num_pos_avg = max(sum over ranks of positives / world, 1),  loss_denorm = max(sum over ranks of ctr targets / world, 1e-6);
// global = {positives, centre-ness target sum} summed over all ranks (== the local sums for one process).
__global__ void fcos_loss_finalize_kernel(const double* __restrict__ local, const double* __restrict__ global, int world,
                                          float* __restrict__ losses) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float num_pos_avg = fmaxf(static_cast<float>(global[0] / world), 1.0f);
    // reduce_sum(ctrness_targets_sum) is an fp32 tensor in the reference; .item() / num_gpus happens in Python doubles
    const double denorm = fmax(static_cast<double>(static_cast<float>(global[1])) / world, 1e-6);
    losses[0] = static_cast<float>(local[0]) / num_pos_avg;
    if (local[1] > 0) {
        losses[1] = static_cast<float>(local[3]) / static_cast<float>(denorm);
        losses[2] = static_cast<float>(local[4]) / num_pos_avg;
    } else {
        losses[1] = 0.f;
        losses[2] = 0.f;
    }
}

}  // namespace sylph
