// Weight gradient of a 3x3 / stride-1 convolution over the flat zero-bordered planes, for the
// backward of the FCOS class tower (reference: autograd of the `cls_tower` convolutions,
// sylph/modeling/meta_fcos/fcos.py:72-122):
//     dW[co][ci][tap] = sum over pixel rows p of dY[p][co] * X[p + dy * Wp + dx][ci]
// i.e. per tap ONE GEMM  D[256 x 256] = dY^T * X_shifted  whose K dimension is the PIXEL axis.  Both operands are read
// straight from the row-major planes ([pixel][channel], channels contiguous): for the tensor core that is an MN-major
// operand -- a TMA box of 64 pixel rows x 64 channels (128-byte swizzle) is exactly the canonical MN-major SWIZZLE_128B atom
// ((8 x 16-byte units along MN) x 8 K rows, K groups 1024 bytes apart = SBO, 64-channel atoms one box apart = LBO), so no
// transpose pass is needed; the instruction descriptor carries a_major = b_major = MN.
// Work split: grid (S, 9 taps, 2 halves of co); CTA (s, tap, h) accumulates its share of the 64-row K tiles (t = s, s + S,
// ...) into one 128 x 256 fp32 TMEM accumulator and writes a partial [co 128][ci 256] tile; wgrad_reduce_kernel sums the S
// partials in order (deterministic) into the OIHW gradient.  Zero border rows of X make the shifted reads exact; dY must be 0
// on border / padding rows (the producer kernel writes interior pixels only into a zeroed buffer).
// Exact mode (SPLIT): rows are [C hi | C lo]; three products per k-step (hi.hi, lo.hi, hi.lo) into the same accumulator.
#pragma once
#include "conv_gemm.cuh"

namespace sylph {

constexpr int kWgTileK = 64;   // pixel rows per K tile (4 tcgen05.mma of K = 16)
constexpr int kWgStages = 2;
constexpr int kWgThreads = 192;   // warp 0: TMA producer, warp 1: MMA issue + TMEM allocation, warps 2..5: epilogue

struct WgradArgs {
    const Seg* segs;       // plane segments of the buffers (PlaneSet::d_segs)
    const int* tile_seg;   // 128-row tile -> segment (PlaneSet::d_tile_seg)
    int n_ktiles;          // total rows / 64
    float* partial;        // [gridDim.x][9][256][256]
};

template <bool SPLIT>
struct WgradSmem {
    static constexpr int kAtom = kWgTileK * 128;                 // 8 KiB: 64 rows x 64 channels
    static constexpr int kA = 2 * kAtom * (SPLIT ? 2 : 1);       // 128 output channels of dY (hi [, lo])
    static constexpr int kB = 4 * kAtom * (SPLIT ? 2 : 1);       // 256 input channels of X (hi [, lo])
    static constexpr int kStage = kA + kB;
    static constexpr int kBarOffset = kWgStages * kStage;
    static constexpr int kTotal = kBarOffset + 1024 + 1024;      // barriers + alignment slack
};

namespace ptx {
// MN-major operand, SWIZZLE_128B: start address | LBO = distance between 64-element atoms along M / N | SBO = 1024 B between
// groups of 8 K rows | descriptor version 1 | layout type 2.
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
}  // namespace ptx

template <bool SPLIT>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad3x3_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, const WgradArgs p) {
    using S = WgradSmem<SPLIT>;
    constexpr uint32_t kIdesc = ptx::make_idesc_f16(128, 256) | (1u << 15) | (1u << 16);   // A and B MN-major
    constexpr uint32_t kTmemCols = 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
    uint64_t* empty_bar = full_bar + kWgStages;
    uint64_t* acc_full = empty_bar + kWgStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    ptx::griddep_launch();
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_dy);
        ptx::prefetch_tensormap(&tmap_x);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kWgStages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        ptx::mbar_init(acc_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::griddep_wait();

    const int split = blockIdx.x, nsplit = gridDim.x, tap = blockIdx.y, half = blockIdx.z;
    const int dyo = tap / 3 - 1, dxo = tap % 3 - 1;
    const bool has_tiles = split < p.n_ktiles;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = split; t < p.n_ktiles; t += nsplit) {
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                const Seg sg = p.segs[p.tile_seg[t >> 1]];
                const int r0 = t * kWgTileK;
                const int rb = r0 + dyo * sg.Wp + dxo;        // rows of X this tap pairs with rows r0.. of dY
                uint8_t* sa = smem + stage * S::kStage;
                uint8_t* sb = sa + S::kA;
                ptx::mbar_arrive_expect_tx(&full_bar[stage], S::kStage);
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    ptx::tma_load_2d(sa + a * S::kAtom, &tmap_dy, &full_bar[stage], half * 128 + a * 64, r0);
                    if constexpr (SPLIT)
                        ptx::tma_load_2d(sa + (2 + a) * S::kAtom, &tmap_dy, &full_bar[stage], 256 + half * 128 + a * 64, r0);
                }
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    ptx::tma_load_2d(sb + b * S::kAtom, &tmap_x, &full_bar[stage], b * 64, rb);
                    if constexpr (SPLIT) ptx::tma_load_2d(sb + (4 + b) * S::kAtom, &tmap_x, &full_bar[stage], 256 + b * 64, rb);
                }
                if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t accumulate = 0;
            for (int t = split; t < p.n_ktiles; t += nsplit) {
                ptx::mbar_wait(&full_bar[stage], phase);
                ptx::tc_fence_after();
                const uint32_t a_hi = ptx::smem_u32(smem + stage * S::kStage);
                const uint32_t b_hi = a_hi + S::kA;
#pragma unroll
                for (int kk = 0; kk < kWgTileK / 16; ++kk) {
                    const uint32_t off = kk * 2048;                       // 16 K rows = two 8-row groups of 1024 bytes
                    const uint64_t da = ptx::make_sw128_mnmajor_desc(a_hi + off, S::kAtom);
                    const uint64_t db = ptx::make_sw128_mnmajor_desc(b_hi + off, S::kAtom);
                    ptx::umma_f16(tmem_base, da, db, kIdesc, accumulate);
                    accumulate = 1;
                    if constexpr (SPLIT) {
                        const uint64_t da_lo = ptx::make_sw128_mnmajor_desc(a_hi + 2 * S::kAtom + off, S::kAtom);
                        const uint64_t db_lo = ptx::make_sw128_mnmajor_desc(b_hi + 4 * S::kAtom + off, S::kAtom);
                        ptx::umma_f16(tmem_base, da_lo, db, kIdesc, 1);
                        ptx::umma_f16(tmem_base, da, db_lo, kIdesc, 1);
                    }
                }
                ptx::umma_commit(&empty_bar[stage]);
                if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
            }
            if (has_tiles) ptx::umma_commit(acc_full);
            else ptx::mbar_arrive(acc_full);
        }
    } else {
        // ------------------------------------------------------------ epilogue: TMEM -> partial tile (fp32)
        ptx::mbar_wait(acc_full, 0);
        ptx::tc_fence_after();
        const int quad = warp & 3;                                        // the TMEM lane quadrant this warp may read
        const int co = half * 128 + quad * 32 + lane;
        float* dst = p.partial + (static_cast<size_t>(split * 9 + tap) * 256 + co) * 256;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 32) {
            uint32_t v[32];
            if (has_tiles) {
                ptx::tmem_ld_32x32b_x32(t_row + c0, v);
                ptx::tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4*>(dst + c0 + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        ptx::tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// dW[(co * 256 + ci) * 9 + tap] = sum over splits (in order, fp64) of partial[((s * 9 + tap) * 256 + co) * 256 + ci]; OIHW.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int splits, float* __restrict__ dW) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 9 * 256 * 256) return;
    const int ci = e & 255, co = (e >> 8) & 255, tap = e >> 16;
    double s = 0.0;
    for (int k = 0; k < splits; ++k) s += static_cast<double>(partial[(static_cast<size_t>(k * 9 + tap) * 256 + co) * 256 + ci]);
    dW[(co * 256 + ci) * 9 + tap] = static_cast<float>(s);
}

}  // namespace sylph
