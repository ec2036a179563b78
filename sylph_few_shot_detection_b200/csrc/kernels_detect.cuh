// FCOS proposal kernels: sigmoid / quality / threshold / candidate compaction, exact per-level top-k by radix
// select, per-image sort + class-aware NMS with early exit + post-NMS top-k + detector_postprocess.  No host syncs
// (the reference path has `.item()`, `.cpu()` and nonzero() syncs per level and image).
// Replaces FCOSOutputs.forward_for_single_feature_map / select_over_all_levels,
// sylph/modeling/meta_fcos/fcos_outputs.py:904-1028, and detector_postprocess (meta_one_stage_detector.py:288-295).
#pragma once
#include "kernels_codegen.cuh"

namespace sylph {

struct DetectParams {
    PyramidGeom pg;
    int n_images;
    int n_classes;
    int logit_stride;     // channels per row of the logits buffer (n_classes padded)
    int cap[5];           // candidate capacity per (image, level)
    long long cand_off[5];  // offset of level l block inside the candidate buffer: off[l] + image * cap[l]
    float thresh;
    int thresh_with_ctr;
    int box_quality;      // bit0 ctrness, bit1 iou
    int pre_topk;
    int post_topk;
    float nms_thresh;
    float level_scale[5];  // fcos_head.scales[l] (1 when USE_SCALE is off)
    int stride[5];
};

__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.f / (1.f + expf(-x)); }

// Candidate key: high word = score bits (score > 0, so integer order == float order), low word = ~(loc * C + cls),
// i.e. larger key = better score, ties broken towards the smaller (location, class) index.  Keys are unique.
__device__ __forceinline__ unsigned long long make_key(float score, unsigned idx) {
    return (static_cast<unsigned long long>(__float_as_uint(score)) << 32) | (0xFFFFFFFFu - idx);
}

// One thread per (image, level, pixel): sigmoid(logits) > thresh -> append (score = cls * quality) to the
// (image, level) candidate list.  counts[image * 5 + level] must be zero on entry.
__global__ void fcos_candidates_kernel(const float* __restrict__ logits, const float* __restrict__ pred,
                                       DetectParams p, unsigned long long* __restrict__ cand, int* __restrict__ counts,
                                       int* __restrict__ overflow) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    long long total = 0;
    long long lvl_start[6];
    for (int l = 0; l < 5; ++l) { lvl_start[l] = total; total += static_cast<long long>(p.n_images) * p.pg.lv[l].H * p.pg.lv[l].W; }
    lvl_start[5] = total;
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
        const size_t row = plane_row(g, n, y, x);
        const float ctr = sigmoidf_ref(pred[row * 16 + 4]);
        float q = ctr;
        if (p.box_quality == 2) q = sigmoidf_ref(pred[row * 16 + 5]);
        else if (p.box_quality == 3) q = sqrtf(sigmoidf_ref(pred[row * 16 + 5]) * ctr);
        const float* lg = logits + row * p.logit_stride;
        for (int c = 0; c < p.n_classes; ++c) {
            const float s = sigmoidf_ref(lg[c]);
            const float scored = s * q;
            const float tested = p.thresh_with_ctr ? scored : s;
            if (tested > p.thresh) {
                const int slot = atomicAdd(&counts[n * 5 + l], 1);
                if (slot < p.cap[l]) {
                    cand[p.cand_off[l] + static_cast<long long>(n) * p.cap[l] + slot] =
                        make_key(scored, static_cast<unsigned>(loc) * p.n_classes + c);
                } else {
                    *overflow = 1;
                }
            }
        }
    }
}

// Exact top-k (k = PRE_NMS_TOPK) of one (image, level) candidate list by MSB-first radix select on the unique
// 64-bit keys; result set is deterministic (topk(sorted=False) in the reference leaves ties unspecified).
// grid = n_images * 5, block = 1024.   sel[(image * 5 + level) * pre_topk + i], sel_count[image * 5 + level].
__global__ void __launch_bounds__(1024)
fcos_select_kernel(const unsigned long long* __restrict__ cand, const int* __restrict__ counts, DetectParams p,
                   unsigned long long* __restrict__ sel, int* __restrict__ sel_count) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ int hist[256];
    __shared__ unsigned long long prefix_s;
    __shared__ int remaining_s;
    __shared__ int out_pos;
    const int n = blockIdx.x / 5, l = blockIdx.x % 5;
    const unsigned long long* src = cand + p.cand_off[l] + static_cast<long long>(n) * p.cap[l];
    const int cnt = min(counts[blockIdx.x], p.cap[l]);
    unsigned long long* dst = sel + static_cast<size_t>(blockIdx.x) * p.pre_topk;
    const int k = p.pre_topk;
    if (cnt <= k) {
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) dst[i] = src[i];
        if (threadIdx.x == 0) sel_count[blockIdx.x] = cnt;
        return;
    }
    if (threadIdx.x == 0) { prefix_s = 0ull; remaining_s = k; out_pos = 0; }
    __syncthreads();
    // find the k-th largest key: after pass b the top (8 * (b + 1)) bits of it are known
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const unsigned long long prefix = prefix_s;
        const unsigned long long mask = pass == 0 ? 0ull : (~0ull << (shift + 8));
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            const unsigned long long v = src[i];
            if ((v & mask) == prefix) atomicAdd(&hist[(v >> shift) & 0xFF], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int rem = remaining_s, b = 255;
            for (; b > 0; --b) {
                if (hist[b] >= rem) break;
                rem -= hist[b];
            }
            prefix_s = prefix | (static_cast<unsigned long long>(b) << shift);
            remaining_s = rem;
        }
        __syncthreads();
    }
    const unsigned long long kth = prefix_s;  // keys are unique: exactly k keys are >= kth
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const unsigned long long v = src[i];
        if (v >= kth) {
            const int o = atomicAdd(&out_pos, 1);
            if (o < k) dst[o] = v;
        }
    }
    if (threadIdx.x == 0) sel_count[blockIdx.x] = k;
}

struct NmsImageArgs {
    float scale_x, scale_y;  // detector_postprocess scale factors (output / image size)
    float out_w, out_h;
};

// One CTA per image.  dynamic smem: keys[sort_n] (u64) | boxes[n_max] (float4) | class offset[n_max] (float) |
// suppressed bitmap[(n_max + 31) / 32] (u32).
__global__ void __launch_bounds__(1024)
fcos_nms_kernel(const unsigned long long* __restrict__ sel, const int* __restrict__ sel_count,
                const float* __restrict__ pred, DetectParams p, const NmsImageArgs* __restrict__ img_args, int sort_n,
                int n_max, float* __restrict__ dets, int* __restrict__ det_counts, int max_dets) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    extern __shared__ __align__(16) unsigned char sm[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(sm);
    float4* rbox = reinterpret_cast<float4*>(sm + static_cast<size_t>(sort_n) * 8);
    float* boff = reinterpret_cast<float*>(rbox + n_max);
    unsigned int* supp = reinterpret_cast<unsigned int*>(boff + n_max);
    __shared__ float red[32];
    __shared__ int kept_idx[1024];
    __shared__ int n_kept_s;
    const int n = blockIdx.x, tid = threadIdx.x;

    // ---- gather the per-level selections; re-key as (score | ~(level, loc * C + cls)) for a total order
    int offs[6];
    offs[0] = 0;
    for (int l = 0; l < 5; ++l) offs[l + 1] = offs[l] + sel_count[n * 5 + l];
    const int total = offs[5];
    for (int i = tid; i < sort_n; i += blockDim.x) {
        unsigned long long v = 0ull;
        if (i < total) {
            int l = 0;
            while (i >= offs[l + 1]) ++l;
            const unsigned long long k0 = sel[(static_cast<size_t>(n) * 5 + l) * p.pre_topk + (i - offs[l])];
            const unsigned idx = 0xFFFFFFFFu - static_cast<unsigned>(k0 & 0xFFFFFFFFu);
            v = (k0 & 0xFFFFFFFF00000000ull) | (0xFFFFFFFFu - ((static_cast<unsigned>(l) << 28) | idx));
        }
        keys[i] = v;
    }
    __syncthreads();
    // ---- bitonic sort, descending
    // every thread owns whole compare-exchange pairs (q -> i with bit j clear, partner i | j): no idle half
    for (int k = 2; k <= sort_n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int q = tid; q < (sort_n >> 1); q += blockDim.x) {
                const int i = ((q & ~(j - 1)) << 1) | (q & (j - 1));
                const int ixj = i | j;
                const unsigned long long a = keys[i], b = keys[ixj];
                const bool desc = (i & k) == 0;
                if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
            }
            __syncthreads();
        }
    }
    // ---- decode boxes (fcos_outputs.py:989-997): loc -/+ relu(reg * scale_l) * stride_l
    float local_max = -3.4e38f;
    for (int i = tid; i < total; i += blockDim.x) {
        const unsigned low = 0xFFFFFFFFu - static_cast<unsigned>(keys[i] & 0xFFFFFFFFu);
        const int l = low >> 28;
        const unsigned idx = low & 0x0FFFFFFFu;
        const int loc = idx / p.n_classes;
        const PlaneGeom g = p.pg.lv[l];
        const int y = loc / g.W, x = loc - y * g.W;
        const float* pr = pred + plane_row(g, n, y, x) * 16;
        const float st = static_cast<float>(p.stride[l]);
        const float lx = static_cast<float>(x * p.stride[l] + p.stride[l] / 2);
        const float ly = static_cast<float>(y * p.stride[l] + p.stride[l] / 2);
        const float r0 = fmaxf(pr[0] * p.level_scale[l], 0.f) * st, r1 = fmaxf(pr[1] * p.level_scale[l], 0.f) * st;
        const float r2 = fmaxf(pr[2] * p.level_scale[l], 0.f) * st, r3 = fmaxf(pr[3] * p.level_scale[l], 0.f) * st;
        const float4 b = make_float4(lx - r0, ly - r1, lx + r2, ly + r3);
        rbox[i] = b;
        local_max = fmaxf(local_max, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
    }
    // block max of all coordinates (torchvision batched_nms: offsets = cls * (boxes.max() + 1))
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if ((tid & 31) == 0) red[tid >> 5] = local_max;
    __syncthreads();
    float max_coord = red[0];
    for (int i = 1; i < static_cast<int>(blockDim.x >> 5); ++i) max_coord = fmaxf(max_coord, red[i]);
    const float off_unit = max_coord + 1.f;
    for (int i = tid; i < total; i += blockDim.x) {
        const unsigned low = 0xFFFFFFFFu - static_cast<unsigned>(keys[i] & 0xFFFFFFFFu);
        const unsigned idx = low & 0x0FFFFFFFu;
        boff[i] = static_cast<float>(idx % p.n_classes) * off_unit;
    }
    for (int i = tid; i < (n_max + 31) / 32; i += blockDim.x) supp[i] = 0u;
    if (tid == 0) n_kept_s = 0;
    __syncthreads();
    // ---- greedy NMS in score order; stops once POST_NMS_TOPK survivors (plus score ties, `>=`) are found.
    // Exactly the sequential greedy algorithm, evaluated 32 candidates at a time: (1) warp 0 collects the next <= 32
    // candidates that no earlier survivor suppressed, (2) the 32 x 32 IoU matrix of the chunk is one comparison per
    // thread (warp m ballots row m), (3) warp 0 resolves the chunk in order with bit operations and applies the stop
    // rules per survivor, (4) all threads test the remaining candidates against the chunk's survivors.  Four barriers per
    // chunk instead of one per survivor (the loop was barrier- and latency-bound: 0.22 ms for 8 images).
    const int keep_cap = min(max_dets, 1024);
    __shared__ int chunk_idx[32];
    __shared__ unsigned chunk_row[32];
    __shared__ int chunk_n_s, next_pos_s, done_s;
    __shared__ unsigned surv_s;
    __shared__ float last_score_s;
    if (tid == 0) { done_s = 0; last_score_s = 0.f; }
    int pos = 0;
    const int warp_id = tid >> 5, lane = tid & 31;
    while (true) {
        // (1) next chunk: the alive candidates among the 1024 positions from pos on, the first 32 of them
        if (warp_id == 0) {
            const int w = (pos >> 5) + lane;
            unsigned alive = 0u;
            if (w * 32 < total) {
                alive = ~supp[w];
                if (w == (pos >> 5)) alive &= 0xFFFFFFFFu << (pos & 31);
                if (w * 32 + 32 > total) alive &= (total & 31) ? ((1u << (total & 31)) - 1u) : 0xFFFFFFFFu;
            }
            const int cnt = __popc(alive);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            int excl = incl - cnt;
            const int all = __shfl_sync(0xffffffffu, incl, 31);
            // lane's alive bits fill chunk slots excl, excl + 1, ... while < 32
            unsigned a = alive;
            int last_taken = -1;
            while (a != 0u && excl < 32) {
                const int b = __ffs(a) - 1;
                a &= a - 1u;
                chunk_idx[excl++] = w * 32 + b;
                last_taken = w * 32 + b;
            }
            // next scan position: one past the last candidate taken when the chunk is full, else the end of the window
            int np = last_taken + 1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) np = max(np, __shfl_xor_sync(0xffffffffu, np, o));
            if (lane == 0) {
                chunk_n_s = min(all, 32);
                next_pos_s = all >= 32 ? np : min(total, ((pos >> 5) + 32) * 32);
            }
        }
        __syncthreads();
        const int cn = chunk_n_s;
        const int next_pos = next_pos_s;
        if (cn == 0) {
            if (next_pos >= total) break;
            pos = next_pos;
            __syncthreads();   // chunk_n_s / next_pos_s are rewritten by warp 0 in the next round
            continue;
        }
        // (2) IoU matrix of the chunk: thread (m, k) = (warp, lane); row m = candidates k > m that m would suppress
        {
            bool hit = false;
            if (warp_id < cn && lane < cn && lane > warp_id) {
                const int ia = chunk_idx[warp_id], ib = chunk_idx[lane];
                const float oa = boff[ia], ob = boff[ib];
                float4 a = rbox[ia], b = rbox[ib];
                a.x += oa; a.y += oa; a.z += oa; a.w += oa;
                b.x += ob; b.y += ob; b.z += ob; b.w += ob;
                const float area_a = (a.z - a.x) * (a.w - a.y);
                const float w = fmaxf(fminf(a.z, b.z) - fmaxf(a.x, b.x), 0.f);
                const float h = fmaxf(fminf(a.w, b.w) - fmaxf(a.y, b.y), 0.f);
                const float inter = w * h;
                const float area_b = (b.z - b.x) * (b.w - b.y);
                hit = inter / (area_a + area_b - inter) > p.nms_thresh;
            }
            const unsigned row = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) chunk_row[warp_id] = row;
        }
        __syncthreads();
        // (3) resolve the chunk in score order (warp 0, every lane runs the same uniform loop)
        if (warp_id == 0) {
            unsigned removed = 0u, surv = 0u;
            int nk = n_kept_s;
            float last_score = last_score_s;
            int done = 0;
            for (int k = 0; k < cn; ++k) {
                if ((removed >> k) & 1u) continue;
                const int i = chunk_idx[k];
                const float sc = sqrtf(__uint_as_float(static_cast<unsigned>(keys[i] >> 32)));
                if ((p.post_topk > 0 && nk >= p.post_topk && sc < last_score) || nk >= keep_cap) { done = 1; break; }
                if (lane == 0) kept_idx[nk] = i;
                ++nk;
                last_score = sc;
                surv |= 1u << k;
                removed |= chunk_row[k];
            }
            if (lane == 0) { n_kept_s = nk; last_score_s = last_score; done_s = done; surv_s = surv; }
        }
        __syncthreads();
        if (done_s) break;
        if (next_pos >= total) break;
        // (4) the chunk's survivors suppress the candidates after the chunk
        {
            const unsigned surv = surv_s;
            for (int j = next_pos + tid; j < total; j += blockDim.x) {
                if ((supp[j >> 5] >> (j & 31)) & 1u) continue;
                const float ob = boff[j];
                float4 b = rbox[j];
                b.x += ob; b.y += ob; b.z += ob; b.w += ob;
                const float area_b = (b.z - b.x) * (b.w - b.y);
                unsigned m = surv;
                while (m != 0u) {
                    const int k = __ffs(m) - 1;
                    m &= m - 1u;
                    const int ia = chunk_idx[k];
                    const float oa = boff[ia];
                    float4 a = rbox[ia];
                    a.x += oa; a.y += oa; a.z += oa; a.w += oa;
                    const float area_a = (a.z - a.x) * (a.w - a.y);
                    const float w = fmaxf(fminf(a.z, b.z) - fmaxf(a.x, b.x), 0.f);
                    const float h = fmaxf(fminf(a.w, b.w) - fmaxf(a.y, b.y), 0.f);
                    const float inter = w * h;
                    if (inter / (area_a + area_b - inter) > p.nms_thresh) {
                        atomicOr(&supp[j >> 5], 1u << (j & 31));
                        break;
                    }
                }
            }
        }
        pos = next_pos;
        __syncthreads();
    }
    __syncthreads();
    const int n_kept = n_kept_s;
    // ---- detector_postprocess: scale, clip, drop empty; one warp compacts in order
    if (tid < 32) {
        const NmsImageArgs ia = img_args[n];
        int out_n = 0;
        for (int base = 0; base < n_kept; base += 32) {
            const int k = base + tid;
            bool ok = false;
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            int i = 0;
            if (k < n_kept) {
                i = kept_idx[k];
                b = rbox[i];
                b.x = fminf(fmaxf(b.x * ia.scale_x, 0.f), ia.out_w);
                b.z = fminf(fmaxf(b.z * ia.scale_x, 0.f), ia.out_w);
                b.y = fminf(fmaxf(b.y * ia.scale_y, 0.f), ia.out_h);
                b.w = fminf(fmaxf(b.w * ia.scale_y, 0.f), ia.out_h);
                ok = (b.z - b.x) > 0.f && (b.w - b.y) > 0.f;
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const int o = out_n + __popc(m & ((1u << tid) - 1u));
                const unsigned low = 0xFFFFFFFFu - static_cast<unsigned>(keys[i] & 0xFFFFFFFFu);
                const int l = low >> 28;
                const unsigned idx = low & 0x0FFFFFFFu;
                const int loc = idx / p.n_classes, cls = idx % p.n_classes;
                const int W = p.pg.lv[l].W;
                const int y = loc / W, x = loc - y * W;
                float* d = dets + (static_cast<size_t>(n) * max_dets + o) * 9;
                d[0] = b.x; d[1] = b.y; d[2] = b.z; d[3] = b.w;
                d[4] = sqrtf(__uint_as_float(static_cast<unsigned>(keys[i] >> 32)));
                d[5] = static_cast<float>(cls);
                d[6] = static_cast<float>(x * p.stride[l] + p.stride[l] / 2);
                d[7] = static_cast<float>(y * p.stride[l] + p.stride[l] / 2);
                d[8] = static_cast<float>(l);
            }
            out_n += __popc(m);
        }
        if (tid == 0) det_counts[n] = out_n;
    }
}

}  // namespace sylph
