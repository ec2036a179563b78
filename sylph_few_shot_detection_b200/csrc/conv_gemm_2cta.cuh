// CTA-pair (cta_group::2) variant of the 3x3 halo convolution for the compute-bound N = 256 layers (FCOS towers, FPN
// output convolutions, res4/res5 conv2).
//
// Why: with one CTA per tile every tcgen05.mma of a 128 x 256 x 16 step reads 4 KB of A and 8 KB of B from shared
// memory while TMA refills the ring at the same rate -- together more than one SM's 128 B/clk of shared-memory
// bandwidth (profiles/r01_ncu_head_tower_kernel.md: tensor pipe 59 % active with L2 and DRAM far from saturated).
// A CTA pair computes a 256 x 256 tile: each CTA keeps ITS 128 rows of A and HALF of the B tile (128 of the 256
// output channels), the leader CTA issues `tcgen05.mma.cta_group::2` (M = 256) which reads A and B from both CTAs'
// shared memory and writes each CTA's 128 accumulator rows into that CTA's own TMEM.  Per SM: 4 KB A + 4 KB B per
// MMA and a 16 KB (not 32 KB) B refill per tap -- within the shared-memory budget.
//
// Protocol (both CTAs run the same code; r = %cluster_ctarank):
//   producer (warp 0)   : waits on its LOCAL empty barriers, issues its own `cp.async.bulk.tensor...cta_group::2`
//                         loads whose bytes complete on the LEADER's full barriers; the leader arms them with the
//                         byte count of both CTAs.
//   MMA (warp 1, r = 0) : waits on the leader's full barriers, issues the pair MMAs, and releases slots / publishes
//                         accumulators with `tcgen05.commit...multicast::cluster` into both CTAs' barriers.
//   epilogue (warps 2-9): identical to the single-CTA kernel on the CTA's own 128 rows; hands the accumulator back by
//                         arriving on the LEADER's tmem_empty barrier (remote arrive through mapa for r = 1).
#pragma once
#include "conv_gemm.cuh"

namespace sylph {

namespace ptx {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair; the bytes complete on the EVEN CTA's barrier (peer bit 24 cleared).
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0,
                                                 int32_t c1) {
    const uint32_t bar_addr = smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f16_pair_x4(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "add.s64 a1, %1, 2;\n\t"
        "add.s64 b1, %2, 2;\n\t"
        "add.s64 a2, %1, 4;\n\t"
        "add.s64 b2, %2, 4;\n\t"
        "add.s64 a3, %1, 6;\n\t"
        "add.s64 b3, %2, 6;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %3, 1;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %3, 1;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %3, 1;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (when all MMAs issued so far retire) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}

}  // namespace ptx

// BN < 256 (res2 / res3 conv2): every tcgen05.mma carries ~90 clocks of fixed cost on top of its N / 2 clocks of math
// (profiles/r01_mma_issue_experiments.md), so a single-CTA N = 64 instruction runs the tensor pipe at ~25 %; one pair
// instruction covers 256 rows, halving the instruction count per output tile.
template <int HALO, int BSLOTS, int BN_ = 256>
struct Gemm2Smem {
    static constexpr int kBN = BN_;                      // output channels of the pair tile
    static constexpr int kBHalf = BN_ / 2;               // B rows held by each CTA
    static constexpr int kAHaloTx = (kBlockM + 2) * 128;
    static constexpr int kAHaloBytes = 17 * 1024;
    static constexpr int kBBytes = kBHalf * kBlockK * 2;  // 16 KiB
    static constexpr int kRingOffset = HALO * kAHaloBytes;
    static constexpr int kBarOffset = kRingOffset + BSLOTS * kBBytes;
    static constexpr int kGnOffset = kBarOffset + 512;
    static constexpr int kTotal = kGnOffset + 4 * 32 * 2 * 4 + 1024;
    static constexpr int kThreads = 320;
};

// Tiles: pair tile index pt -> (pm, n_tile); CTA r handles output M tile 2 * pm + r.  p.num_m_tiles may be odd: the
// phantom tile of the last pair loads zero-filled rows and stores nothing.
template <int HALO, int BSLOTS, int BN_ = 256, bool SPLIT = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmArgs p) {
    using S = Gemm2Smem<HALO, BSLOTS, BN_>;
    constexpr int BN = S::kBN;
    constexpr int COLS = BN / 2;
    constexpr int CH = 32;
    constexpr int kAcc = BN <= 128 ? 4 : 2;              // accumulator buffers (see conv_gemm.cuh)
    static_assert(BN == 64 || BN == 128 || BN == 256, "pair tile widths");
    constexpr uint32_t kTmemCols = 512;
    constexpr uint32_t kIdesc = ptx::make_idesc_f16(256, BN);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);   // used in the leader only
    uint64_t* empty_bar = full_bar + BSLOTS;                                  // local, armed by multicast commits
    uint64_t* a_full = empty_bar + BSLOTS;                                    // leader only
    uint64_t* a_empty = a_full + HALO;                                        // local
    uint64_t* tmem_full = a_empty + HALO;                                     // local
    uint64_t* tmem_empty = tmem_full + 4;                                     // leader only (count: both CTAs)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 4);
    float* gn_smem = reinterpret_cast<float*>(smem + S::kGnOffset);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();

    ptx::griddep_launch();
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_a);
        ptx::prefetch_tensormap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < BSLOTS; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < HALO; ++a) {
            ptx::mbar_init(&a_full[a], 1);
            ptx::mbar_init(&a_empty[a], 1);
        }
        for (int a = 0; a < kAcc; ++a) {
            ptx::mbar_init(&tmem_full[a], 1);
            ptx::mbar_init(&tmem_empty[a], 2 * 8);   // one arrive per epilogue warp of both CTAs (512 per-thread
                                                     // arrivals, half of them remote, serialised on one barrier)
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_alloc_pair(tmem_slot, kTmemCols);
        ptx::tmem_relinquish_pair();
    }
    ptx::tc_fence_before();
    ptx::cluster_sync_all();   // barriers of BOTH CTAs are initialised before anyone signals across the pair
    ptx::tc_fence_after();
    if (*tmem_slot != 0u) __trap();   // all 512 columns: the allocation starts at column 0 (see conv_gemm.cuh)
    constexpr uint32_t tmem_base = 0u;
    ptx::griddep_wait();

    const int pair_m_tiles = (p.num_m_tiles + 1) >> 1;
    const int total_pairs = pair_m_tiles * p.num_n_tiles;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int groups = 3 * p.kblocks_per_tap;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int sb = 0, hs = 0;
            uint32_t bphase = 0, hphase = 0;
            for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
                const int pm = pt / p.num_n_tiles;
                const int n_tile = pt - pm * p.num_n_tiles;
                const int m_tile = 2 * pm + static_cast<int>(rank);
                const int m_seg = m_tile < p.num_m_tiles ? m_tile : p.num_m_tiles - 1;
                const int a_row_base = (p.tile_begin + m_tile) * kBlockM + p.a_row_delta;
                const int wp = p.segs[p.tile_seg[p.tile_begin + m_seg]].Wp;
                const int b_row_base = n_tile * BN + static_cast<int>(rank) * S::kBHalf;
                for (int dyi = 0; dyi < 3; ++dyi) {
                    for (int kb = 0; kb < p.kblocks_per_tap; ++kb) {
                        ptx::mbar_wait(&a_empty[hs], hphase ^ 1u);
                        if (rank == 0) ptx::mbar_arrive_expect_tx(&a_full[hs], 2 * S::kAHaloTx);
                        int acol = kb * kBlockK;   // SPLIT: [hi | lo] rows against [w_hi | w_hi | w_lo] (see GemmArgs::a_wrap)
                        if constexpr (SPLIT) { if (acol >= p.a_wrap) acol -= p.a_wrap; }
                        ptx::tma_load_2d_pair(smem + hs * S::kAHaloBytes, &tmap_a, &a_full[hs], acol,
                                              a_row_base + (dyi - 1) * wp - 1);
                        if (++hs == HALO) { hs = 0; hphase ^= 1u; }
                        for (int dxi = 0; dxi < 3; ++dxi) {
                            ptx::mbar_wait(&empty_bar[sb], bphase ^ 1u);
                            if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[sb], 2 * S::kBBytes);
                            ptx::tma_load_2d_pair(smem + S::kRingOffset + sb * S::kBBytes, &tmap_b, &full_bar[sb],
                                                  kb * kBlockK, (dyi * 3 + dxi) * p.b_rows_per_tap + b_row_base);
                            if (++sb == BSLOTS) { sb = 0; bphase ^= 1u; }
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        if (lane == 0 && rank == 0) {
            int sb = 0, hs = 0, acc = 0;
            uint32_t bphase = 0, hphase = 0, acc_phase = 0;
            for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
                ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
                uint32_t first = 0;
                for (int g = 0; g < groups; ++g) {
                    ptx::mbar_wait(&a_full[hs], hphase);
                    const uint32_t sa = ptx::smem_u32(smem + hs * S::kAHaloBytes);
                    for (int dxi = 0; dxi < 3; ++dxi) {
                        ptx::mbar_wait(&full_bar[sb], bphase);
                        ptx::tc_fence_after();
                        const uint64_t da = ptx::make_sw128_kmajor_desc(sa + dxi * 128);
                        const uint64_t db = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + S::kRingOffset + sb * S::kBBytes));
                        ptx::umma_f16_pair_x4(d_tmem, da, db, kIdesc, first);
                        first = 1;
                        ptx::umma_commit_pair(&empty_bar[sb]);
                        if (++sb == BSLOTS) { sb = 0; bphase ^= 1u; }
                    }
                    ptx::umma_commit_pair(&a_empty[hs]);
                    if (++hs == HALO) { hs = 0; hphase ^= 1u; }
                }
                ptx::umma_commit_pair(&tmem_full[acc]);
                if (++acc == kAcc) { acc = 0; acc_phase ^= 1u; }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
        const int quad = warp & 3;
        const int half_idx = (warp - 2) >> 2;
        const int col_begin = half_idx * COLS;
        const int et = (warp - 2) * 32 + lane;
        const int r_in_tile = quad * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        const bool need_seg = (p.flags & (kEpiMask | kEpiGnStats)) != 0;
        auto fetch_seg = [&](int t) -> Seg {   // plane of this CTA's M tile of pair tile t (clamped for the phantom tile)
            int pm_, n_;
            split_tile(t, p.num_n_tiles, pm_, n_);
            const int m_ = 2 * pm_ + static_cast<int>(rank);
            return p.segs[__ldg(p.tile_seg + p.tile_begin + (m_ < p.num_m_tiles ? m_ : p.num_m_tiles - 1))];
        };
        Seg sg_next{};
        if (need_seg && cluster_id < total_pairs) sg_next = fetch_seg(cluster_id);
        for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
            int pm, n_tile;
            split_tile(pt, p.num_n_tiles, pm, n_tile);
            const int m_tile = 2 * pm + static_cast<int>(rank);
            const bool valid = m_tile < p.num_m_tiles;
            const int abs_tile = p.tile_begin + m_tile;
            const int row = abs_tile * kBlockM + r_in_tile;
            const Seg sg = sg_next;
            if (need_seg && pt + num_clusters < total_pairs) sg_next = fetch_seg(pt + num_clusters);
            const bool interior = valid && (need_seg ? row_is_interior(sg, row) : true);
            const bool keep = interior || !(p.flags & kEpiMask);
            const size_t out_off = static_cast<size_t>(row) * p.ldc + static_cast<size_t>(n_tile) * BN + col_begin;

            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                   static_cast<uint32_t>(acc * BN + col_begin);
            if (valid) {
#pragma unroll
                for (int c0 = 0; c0 < COLS; c0 += CH) {
                    uint32_t v[CH];
                    ptx::tmem_ld_32x32b_x32(t_row + c0, v);
                    ptx::tmem_ld_wait();
                    float f[CH];
#pragma unroll
                    for (int j = 0; j < CH; ++j) f[j] = __uint_as_float(v[j]);
                    if (p.bias != nullptr) {
                        const float4* bp = reinterpret_cast<const float4*>(p.bias + n_tile * BN + col_begin + c0);
#pragma unroll
                        for (int j = 0; j < CH / 4; ++j) {
                            const float4 b = __ldg(bp + j);
                            f[4 * j + 0] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
                        }
                    }
                    if (BN == 256 && (p.flags & kEpiGnStats)) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float s = 0.f, ss = 0.f;
                            if (interior) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) { const float x = f[8 * g + j]; s += x; ss += x * x; }
                            }
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) {
                                s += __shfl_xor_sync(0xffffffffu, s, o);
                                ss += __shfl_xor_sync(0xffffffffu, ss, o);
                            }
                            if (lane == 0) {
                                const int grp = ((col_begin + c0) >> 3) + g;
                                gn_smem[(quad * 32 + grp) * 2 + 0] = s;
                                gn_smem[(quad * 32 + grp) * 2 + 1] = ss;
                            }
                        }
                    }
                    if (p.flags & kEpiOutF32) {
                        if (p.flags & kEpiRelu) {
#pragma unroll
                            for (int j = 0; j < CH; ++j) f[j] = fmaxf(f[j], 0.f);
                        }
                        if (!keep) {
#pragma unroll
                            for (int j = 0; j < CH; ++j) f[j] = 0.f;
                        }
                        float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + out_off + c0);
#pragma unroll
                        for (int j = 0; j < CH / 4; ++j) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    } else {
                        const bool relu = (p.flags & kEpiRelu) != 0;
                        uint4* op = reinterpret_cast<uint4*>(static_cast<__half*>(p.out) + out_off + c0);
                        if constexpr (SPLIT) {
                            uint4* oq = reinterpret_cast<uint4*>(static_cast<__half*>(p.out) + out_off + p.out_lo + c0);
#pragma unroll
                            for (int j = 0; j < CH / 8; ++j) {
                                uint4 hi, lo;
                                split8(f + 8 * j, relu, keep, hi, lo);
                                op[j] = hi;
                                oq[j] = lo;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < CH / 8; ++j) op[j] = pack8(f + 8 * j, relu, keep);
                        }
                    }
                }
            }
            // accumulator drained: tell the leader's MMA warp (remote arrive for the second CTA of the pair)
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_shared(ptx::smem_u32(&tmem_empty[acc]), 0));
            if (BN == 256 && (p.flags & kEpiGnStats)) {
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (valid && et < 64) {
                    const float t = gn_smem[et] + gn_smem[64 + et] + gn_smem[128 + et] + gn_smem[192 + et];
                    p.gn_partial[static_cast<size_t>(abs_tile) * 64 + et] = t;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            if (++acc == kAcc) { acc = 0; acc_phase ^= 1u; }
        }
    }

    ptx::tc_fence_before();
    ptx::cluster_sync_all();   // nobody leaves while the peer may still signal into / read from this CTA
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
    }
}

}  // namespace sylph
