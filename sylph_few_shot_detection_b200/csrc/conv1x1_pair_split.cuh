// CTA-pair (cta_group::2) 1x1 convolution in split-operand ("exact") mode with a CHUNKED staged epilogue: the deep bottleneck
// layers (res4 / res5 conv1 + conv3, the shortcut convolutions) of the exact precision mode.
//
// Why: three products per multiply make these layers shared-memory-bound on one SM -- a 128 x 128 x 16 tcgen05.mma reads
// 4 KB of A and 4 KB of B in 64 clocks (128 B/clk, ALL of an SM's shared-memory bandwidth) while TMA refills the ring and the
// staged epilogue moves 4 B per output element through the same memory (profiles/r02_quad_stage_ab.log: res4 conv3 runs at
// 0.42 of the tensor peak and 0.51 of the HBM peak at once).  A CTA pair computes a 256 x 256 tile: each CTA holds its 128 rows
// of A and HALF of the weight tile, one M = 256 instruction per k-step reads 4 KB + 4 KB per SM in 128 clocks (64 B/clk).
//
// Epilogue: a split staged tile of 128 x 256 outputs is 128 KB (hi + lo halves), which leaves no room for a second buffer, so
// the tile is staged in CHUNKS of 64 columns (one hi panel + one lo panel = 32 KB) through a ring of CB chunk buffers:
//   warp 10 (epilogue I/O): TMA-loads the residual chunk c + CB - 1 as soon as the store of chunk c - 1 has been read out,
//                           TMA-stores chunk c when the eight epilogue warps have rewritten it;
//   warps 2-9             : per chunk one tcgen05.ld of 32 columns, residual hi + lo from the chunk buffer, bias / ReLU / mask,
//                           split into hi / lo, rewrite in place.
// Residual loads never wait behind the mainloop's TMA queue (they did in the single-buffer 256-wide kernel, whose producer warp
// issues both), and the granularity of the overlap is a chunk, not a tile.
// Mainloop: K' = 3C loop over [hi | lo | hi] x [w_hi | w_hi | w_lo] (QS = false, 32 KB stages) or quad stages (QS = true: a_hi,
// a_lo, w_hi, w_lo of a k-block loaded once, 64 KB stages, three instructions per stage); protocol = conv1x1_pair_staged_kernel.
// Replaces the same cuDNN convolutions as conv_gemm.cuh (detectron2 BottleneckBlock conv1 / conv3 / shortcut).
#pragma once
#include "conv_gemm_2cta.cuh"

namespace sylph {

template <int STAGES, int CB, bool QS>
struct PairSplitSmem {
    static constexpr int kBN = 256;
    static constexpr int kABytes = kBlockM * kBlockK * 2;          // 16 KiB: this CTA's 128 rows of one A k-block (hi or lo)
    static constexpr int kBBytes = (kBN / 2) * kBlockK * 2;        // 16 KiB: this CTA's half of one weight k-block
    static constexpr int kStageBytes = (QS ? 2 : 1) * (kABytes + kBBytes);
    static constexpr int kChunkBytes = 2 * kBlockM * 128;          // 32 KiB: hi panel + lo panel of 64 output channels
    static constexpr int kEpiOffset = STAGES * kStageBytes;
    static constexpr int kBarOffset = kEpiOffset + CB * kChunkBytes;
    static constexpr int kTotal = kBarOffset + 1024 + 1024;        // barriers + alignment slack
    static constexpr int kThreads = 352;
};

template <int STAGES, int CB, bool QS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(352, 1)
conv1x1_pair_split_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                          const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_out,
                          const GemmArgs p) {
    using S = PairSplitSmem<STAGES, CB, QS>;
    constexpr int BN = S::kBN;
    constexpr int NCH = BN / 64;      // chunks per tile
    constexpr int kAcc = 2;
    constexpr uint32_t kTmemCols = 512;
    constexpr uint32_t kIdesc = ptx::make_idesc_f16(256, BN);
    static_assert(CB >= 2 && CB <= 8, "chunk ring");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);   // used in the leader only
    uint64_t* empty_bar = full_bar + STAGES;                                  // local, armed by multicast commits
    uint64_t* tmem_full = empty_bar + STAGES;                                 // local, armed by multicast commits
    uint64_t* tmem_empty = tmem_full + 2;                                     // leader only (8 warps of each CTA)
    uint64_t* res_full = tmem_empty + 2;                                      // local: residual chunk landed
    uint64_t* stage_ready = res_full + CB;                                    // local: epilogue rewrote the chunk buffer
    uint64_t* epi_free = stage_ready + CB;                                    // local: TMA store finished reading it
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_free + CB);
    uint8_t* epi_smem = smem + S::kEpiOffset;

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();

    ptx::griddep_launch();
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmap_a);
        ptx::prefetch_tensormap(&tmap_b);
        ptx::prefetch_tensormap(&tmap_res);
        ptx::prefetch_tensormap(&tmap_out);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tmem_full[a], 1);
            ptx::mbar_init(&tmem_empty[a], 2 * 8);     // one arrive per epilogue warp of BOTH CTAs
        }
        for (int b = 0; b < CB; ++b) {
            ptx::mbar_init(&res_full[b], 1);
            ptx::mbar_init(&stage_ready[b], 8);        // one arrive per epilogue warp
            ptx::mbar_init(&epi_free[b], 1);
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
    if (*tmem_slot != 0u) __trap();
    constexpr uint32_t tmem_base = 0u;
    ptx::griddep_wait();

    const int pair_m_tiles = (p.num_m_tiles + 1) >> 1;
    const int total_pairs = pair_m_tiles * p.num_n_tiles;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int ksteps = p.kblocks_per_tap;   // taps == 1; QS: C / 64, else 3C / 64
    const bool use_res = (p.flags & kEpiResidual) != 0;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs): operands only
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
                int pm, n_tile;
                split_tile(pt, p.num_n_tiles, pm, n_tile);
                const int m_tile = 2 * pm + static_cast<int>(rank);
                const int a_row_base = (p.tile_begin + m_tile) * kBlockM + p.a_row_delta;   // phantom tile: zero fill / ignored rows
                const int b_row_base = n_tile * BN + static_cast<int>(rank) * (BN / 2);
                for (int kb = 0; kb < ksteps; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                    if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
                    uint8_t* sa = smem + stage * S::kStageBytes;
                    if constexpr (QS) {
                        ptx::tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], kb * kBlockK, a_row_base);
                        ptx::tma_load_2d_pair(sa + S::kABytes, &tmap_a, &full_bar[stage], (p.a_wrap >> 1) + kb * kBlockK, a_row_base);
                        ptx::tma_load_2d_pair(sa + 2 * S::kABytes, &tmap_b, &full_bar[stage], kb * kBlockK, b_row_base);
                        ptx::tma_load_2d_pair(sa + 2 * S::kABytes + S::kBBytes, &tmap_b, &full_bar[stage], p.a_wrap + kb * kBlockK, b_row_base);
                    } else {
                        int acol = kb * kBlockK;
                        if (acol >= p.a_wrap) acol -= p.a_wrap;
                        ptx::tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], acol, a_row_base);
                        ptx::tma_load_2d_pair(sa + S::kABytes, &tmap_b, &full_bar[stage], kb * kBlockK, b_row_base);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        if (lane == 0 && rank == 0) {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
                ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
                for (int kb = 0; kb < ksteps; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * S::kStageBytes);
                    if constexpr (QS) {
                        const uint64_t da_hi = ptx::make_sw128_kmajor_desc(sa), da_lo = ptx::make_sw128_kmajor_desc(sa + S::kABytes);
                        const uint64_t db_hi = ptx::make_sw128_kmajor_desc(sa + 2 * S::kABytes);
                        const uint64_t db_lo = ptx::make_sw128_kmajor_desc(sa + 2 * S::kABytes + S::kBBytes);
                        ptx::umma_f16_pair_x4(d_tmem, da_hi, db_hi, kIdesc, kb ? 1u : 0u);
                        ptx::umma_f16_pair_x4(d_tmem, da_lo, db_hi, kIdesc, 1u);
                        ptx::umma_f16_pair_x4(d_tmem, da_hi, db_lo, kIdesc, 1u);
                    } else {
                        ptx::umma_f16_pair_x4(d_tmem, ptx::make_sw128_kmajor_desc(sa), ptx::make_sw128_kmajor_desc(sa + S::kABytes),
                                              kIdesc, kb ? 1u : 0u);
                    }
                    ptx::umma_commit_pair(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                ptx::umma_commit_pair(&tmem_full[acc]);
                if (++acc == kAcc) { acc = 0; acc_phase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp < 10) {
        // ------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
        const int quad = warp & 3;
        const int half_idx = (warp - 2) >> 2;           // which 32 of a chunk's 64 columns
        const int r_in_tile = quad * 32 + lane;
        int acc = 0, cit = 0;                            // cit: chunks of VALID tiles processed so far
        uint32_t acc_phase = 0;
        const bool need_seg = (p.flags & kEpiMask) != 0;
        const bool relu = (p.flags & kEpiRelu) != 0;
        const uint32_t leader_tmem_empty = ptx::mapa_shared(ptx::smem_u32(&tmem_empty[0]), 0);
        for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
            int pm, n_tile;
            split_tile(pt, p.num_n_tiles, pm, n_tile);
            const int m_tile = 2 * pm + static_cast<int>(rank);
            const bool valid = m_tile < p.num_m_tiles;
            const int abs_tile = p.tile_begin + m_tile;
            const int row = abs_tile * kBlockM + r_in_tile;
            bool keep = true;
            if (need_seg && valid) {
                const Seg sg = p.segs[__ldg(p.tile_seg + abs_tile)];
                keep = row_is_interior(sg, row);
            }
            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
            if (valid) {
#pragma unroll 1
                for (int ch = 0; ch < NCH; ++ch, ++cit) {
                    const int cb = cit % CB;
                    const uint32_t ph = (cit / CB) & 1;
                    if (use_res) ptx::mbar_wait(&res_full[cb], ph);
                    else ptx::mbar_wait(&epi_free[cb], ph ^ 1u);
                    uint32_t v[32];
                    ptx::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                                static_cast<uint32_t>(acc * BN + ch * 64 + half_idx * 32), v);
                    uint8_t* hi_row = epi_smem + cb * S::kChunkBytes + r_in_tile * 128;
                    uint8_t* lo_row = hi_row + kBlockM * 128;
                    const int ch16 = half_idx * 4;                 // first 16-byte piece of this warp's 32 columns
                    uint4 rr[4], rl[4];
                    if (use_res) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            rr[j] = *reinterpret_cast<const uint4*>(hi_row + (((ch16 + j) ^ (r_in_tile & 7)) << 4));
                            rl[j] = *reinterpret_cast<const uint4*>(lo_row + (((ch16 + j) ^ (r_in_tile & 7)) << 4));
                        }
                    }
                    ptx::tmem_ld_wait();
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    if (p.bias != nullptr) {
                        const float4* bp = reinterpret_cast<const float4*>(p.bias + n_tile * BN + ch * 64 + half_idx * 32);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = __ldg(bp + j);
                            f[4 * j + 0] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
                        }
                    }
                    if (use_res) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {   // (hi + lo) is exact in fp32; one rounding when it meets the accumulator
                            float r8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                            add8(r8, rr[j]);
                            add8(r8, rl[j]);
#pragma unroll
                            for (int q = 0; q < 8; ++q) f[8 * j + q] += r8[q];
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 hi, lo;
                        split8(f + 8 * j, relu, keep, hi, lo);
                        *reinterpret_cast<uint4*>(hi_row + (((ch16 + j) ^ (r_in_tile & 7)) << 4)) = hi;
                        *reinterpret_cast<uint4*>(lo_row + (((ch16 + j) ^ (r_in_tile & 7)) << 4)) = lo;
                    }
                    ptx::fence_proxy_async();          // generic-proxy smem writes -> visible to the TMA store
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&stage_ready[cb]);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            // accumulator drained: tell the leader's MMA warp (remote arrive for the second CTA of the pair)
            if (lane == 0) ptx::mbar_arrive_cluster(leader_tmem_empty + static_cast<uint32_t>(acc) * 8u);
            if (++acc == kAcc) { acc = 0; acc_phase ^= 1u; }
        }
    } else {
        // ------------------------------------------------------------ epilogue I/O warp (warp 10): residual chunk loads, chunk stores
        if (lane == 0) {
            // two cursors over the same sequence of (valid tile, chunk): L = next residual chunk to load, St = next chunk to store
            auto tile_valid = [&](int pt) {
                int pm, n_tile;
                split_tile(pt, p.num_n_tiles, pm, n_tile);
                return 2 * pm + static_cast<int>(rank) < p.num_m_tiles;
            };
            auto first_valid = [&](int pt) {
                while (pt < total_pairs && !tile_valid(pt)) pt += num_clusters;
                return pt;
            };
            auto coords = [&](int pt, int ch, int& col, int& row0) {
                int pm, n_tile;
                split_tile(pt, p.num_n_tiles, pm, n_tile);
                row0 = (p.tile_begin + 2 * pm + static_cast<int>(rank)) * kBlockM;
                col = n_tile * BN + ch * 64;
            };
            int l_pt = first_valid(cluster_id), l_ch = 0, n_load = 0;
            auto issue_load = [&]() {
                const int cb = n_load % CB;
                int col, row0;
                coords(l_pt, l_ch, col, row0);
                ptx::mbar_arrive_expect_tx(&res_full[cb], S::kChunkBytes);
                ptx::tma_load_2d(epi_smem + cb * S::kChunkBytes, &tmap_res, &res_full[cb], col, row0);
                ptx::tma_load_2d(epi_smem + cb * S::kChunkBytes + kBlockM * 128, &tmap_res, &res_full[cb], p.res_lo + col, row0);
                ++n_load;
                if (++l_ch == NCH) { l_ch = 0; l_pt = first_valid(l_pt + num_clusters); }
            };
            if (use_res) {
                for (int i = 0; i < CB && l_pt < total_pairs; ++i) issue_load();
            }
            int n_store = 0;
            for (int s_pt = first_valid(cluster_id); s_pt < total_pairs; s_pt = first_valid(s_pt + num_clusters)) {
                for (int ch = 0; ch < NCH; ++ch, ++n_store) {
                    const int cb = n_store % CB;
                    ptx::mbar_wait(&stage_ready[cb], (n_store / CB) & 1);
                    int col, row0;
                    coords(s_pt, ch, col, row0);
                    ptx::tma_store_2d(&tmap_out, epi_smem + cb * S::kChunkBytes, col, row0);
                    ptx::tma_store_2d(&tmap_out, epi_smem + cb * S::kChunkBytes + kBlockM * 128, p.out_lo + col, row0);
                    ptx::bulk_commit_group();
                    if (n_store >= 1) {
                        // the store of the PREVIOUS chunk has been read out of shared memory: its buffer is free
                        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        if (use_res) { if (l_pt < total_pairs) issue_load(); }     // chunk (n_store - 1) + CB goes into that buffer
                        else ptx::mbar_arrive(&epi_free[(n_store - 1) % CB]);
                    }
                }
            }
            ptx::bulk_wait_all();
        }
        __syncwarp();
    }

    ptx::tc_fence_before();
    ptx::cluster_sync_all();   // nobody leaves while the peer may still signal into / read from this CTA
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
    }
}

}  // namespace sylph
