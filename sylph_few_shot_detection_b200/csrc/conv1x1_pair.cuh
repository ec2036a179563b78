// CTA-pair (cta_group::2) 1x1 convolution with the staged TMA-in / TMA-out epilogue, N tiles of 256 channels: the
// bottleneck 1x1 convolutions of res4 / res5 (and the shortcut convolutions), K >= 256.
//
// Why: for these layers a 128 x 256 output tile pulls MORE weight bytes than activation bytes through TMA (res5 conv3,
// K = 512: A 128 KB + B 256 KB + residual 64 KB per tile), the tiles of all SMs together draw ~9.5 TB/s from L2 -- the
// L2 slice limit -- and every k-step costs a ~214-clock N = 256 tcgen05.mma (profiles/r01_weight_tile_reuse_ab.log,
// DESIGN.md section 4).  A CTA pair computes a 256 x 256 tile: each CTA loads ITS 128 rows of A and HALF of the B tile,
// the leader issues one M = 256 `tcgen05.mma.cta_group::2` per k-step.  Per SM and k-step: 16 KB A + 16 KB B (was
// 16 + 32) and half the MMA instructions.
//
// Mainloop protocol = conv3x3_pair_kernel (conv_gemm_2cta.cuh) without the halo; epilogue = the staged epilogue of
// conv_gemm_f16_kernel (conv_gemm.cuh), entirely CTA-local: the residual tile of the CTA's own 128 rows arrives by TMA
// in a swizzled staging buffer while the previous tile is in its epilogue, is rewritten in place (bias, residual, ReLU,
// border mask) and leaves by TMA store from warp 10.
#pragma once
#include "conv_gemm_2cta.cuh"

namespace sylph {

template <int STAGES, int EPI_BUFS>
struct Pair1x1Smem {
    static constexpr int kBN = 256;
    static constexpr int kABytes = kBlockM * kBlockK * 2;          // 16 KiB: this CTA's 128 rows of A
    static constexpr int kBBytes = (kBN / 2) * kBlockK * 2;        // 16 KiB: this CTA's half of the B tile
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kEpiBytes = kBlockM * kBN * 2;            // 64 KiB staged output tile (4 swizzled panels)
    static constexpr int kEpiOffset = STAGES * kStageBytes;
    static constexpr int kBarOffset = kEpiOffset + EPI_BUFS * kEpiBytes;
    static constexpr int kTotal = kBarOffset + 1024 + 1024;        // barriers + alignment slack
    static constexpr int kThreads = 352;
};

// Tiles: pair tile pt -> (pm, n_tile), n fastest; CTA r of the pair owns output M tile 2 * pm + r.  An odd number of
// M tiles leaves a phantom tile in the last pair: its loads still run (zero fill / ignored rows) so that the pair's
// barrier protocol stays symmetric, its residual load, epilogue arithmetic and store are skipped.
template <int STAGES, int EPI_BUFS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(352, 1)
conv1x1_pair_staged_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                           const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_out,
                           const GemmArgs p) {
    using S = Pair1x1Smem<STAGES, EPI_BUFS>;
    constexpr int BN = S::kBN;
    constexpr int COLS = BN / 2;      // columns per epilogue warp (two warps per TMEM lane quadrant)
    constexpr int CH = 32;
    constexpr int kAcc = 2;
    constexpr int kBufs = EPI_BUFS;
    constexpr uint32_t kTmemCols = 512;
    constexpr uint32_t kIdesc = ptx::make_idesc_f16(256, BN);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);   // used in the leader only
    uint64_t* empty_bar = full_bar + STAGES;                                  // local, armed by multicast commits
    uint64_t* tmem_full = empty_bar + STAGES;                                 // local, armed by multicast commits
    uint64_t* tmem_empty = tmem_full + 2;                                     // leader only (8 warps of each CTA)
    uint64_t* res_full = tmem_empty + 2;                                      // local: residual tile landed
    uint64_t* stage_ready = res_full + 2;                                     // local: epilogue wrote the staging buffer
    uint64_t* epi_free = stage_ready + 2;                                     // local: TMA store finished reading it
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_free + 2);
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
            ptx::mbar_init(&res_full[a], 1);
            ptx::mbar_init(&stage_ready[a], 8);        // one arrive per epilogue warp
            ptx::mbar_init(&epi_free[a], 1);
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
    const int ksteps = p.kblocks_per_tap;   // taps == 1
    const bool use_res_tile = (p.flags & kEpiResidual) != 0;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0, vit = 0;
            uint32_t phase = 0;
            for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
                int pm, n_tile;
                split_tile(pt, p.num_n_tiles, pm, n_tile);
                const int m_tile = 2 * pm + static_cast<int>(rank);
                const bool valid = m_tile < p.num_m_tiles;
                const int out_row_base = (p.tile_begin + m_tile) * kBlockM;
                const int a_row_base = out_row_base + p.a_row_delta;
                const int b_row_base = n_tile * BN + static_cast<int>(rank) * (BN / 2);
                if (valid) {
                    if (use_res_tile) {
                        const int buf = vit % kBufs;
                        ptx::mbar_wait(&epi_free[buf], ((vit / kBufs) & 1) ^ 1u);
                        ptx::mbar_arrive_expect_tx(&res_full[buf], S::kEpiBytes);
#pragma unroll
                        for (int pn = 0; pn < BN / 64; ++pn)
                            ptx::tma_load_2d(epi_smem + buf * S::kEpiBytes + pn * (kBlockM * 128), &tmap_res, &res_full[buf],
                                             n_tile * BN + pn * 64, out_row_base);
                    }
                    ++vit;
                }
                for (int kb = 0; kb < ksteps; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
                    if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
                    uint8_t* sa = smem + stage * S::kStageBytes;
                    ptx::tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], kb * kBlockK, a_row_base);
                    ptx::tma_load_2d_pair(sa + S::kABytes, &tmap_b, &full_bar[stage], kb * kBlockK, b_row_base);
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
                    ptx::umma_f16_pair_x4(d_tmem, ptx::make_sw128_kmajor_desc(sa), ptx::make_sw128_kmajor_desc(sa + S::kABytes),
                                          kIdesc, kb ? 1u : 0u);
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
        const int half_idx = (warp - 2) >> 2;
        const int col_begin = half_idx * COLS;
        const int r_in_tile = quad * 32 + lane;
        int acc = 0, vit = 0;
        uint32_t acc_phase = 0;
        const bool need_seg = (p.flags & kEpiMask) != 0;
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
            if (valid) {
                const int buf = vit % kBufs;
                const uint32_t ph = (vit / kBufs) & 1;
                uint8_t* tile_smem = epi_smem + buf * S::kEpiBytes;
                if (use_res_tile) ptx::mbar_wait(&res_full[buf], ph);
                else ptx::mbar_wait(&epi_free[buf], ph ^ 1u);
                ptx::tc_fence_after();
                const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                       static_cast<uint32_t>(acc * BN + col_begin);
#pragma unroll
                for (int c0 = 0; c0 < COLS; c0 += CH) {
                    uint32_t v[CH];
                    ptx::tmem_ld_32x32b_x32(t_row + c0, v);
                    const int col = col_begin + c0;
                    uint8_t* panel = tile_smem + (col >> 6) * (kBlockM * 128) + r_in_tile * 128;
                    const int ch16 = (col & 63) >> 3;
                    uint4 rr[CH / 8];
                    if (use_res_tile) {
#pragma unroll
                        for (int j = 0; j < CH / 8; ++j)
                            rr[j] = *reinterpret_cast<const uint4*>(panel + (((ch16 + j) ^ (r_in_tile & 7)) << 4));
                    }
                    ptx::tmem_ld_wait();
                    float f[CH];
#pragma unroll
                    for (int j = 0; j < CH; ++j) f[j] = __uint_as_float(v[j]);
                    if (p.bias != nullptr) {
                        const float4* bp = reinterpret_cast<const float4*>(p.bias + n_tile * BN + col);
#pragma unroll
                        for (int j = 0; j < CH / 4; ++j) {
                            const float4 b = __ldg(bp + j);
                            f[4 * j + 0] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
                        }
                    }
                    if (use_res_tile) {
#pragma unroll
                        for (int j = 0; j < CH / 8; ++j) {
                            const float2 a = unpack_half2(rr[j].x), b = unpack_half2(rr[j].y), c = unpack_half2(rr[j].z),
                                         d = unpack_half2(rr[j].w);
                            f[8 * j + 0] += a.x; f[8 * j + 1] += a.y; f[8 * j + 2] += b.x; f[8 * j + 3] += b.y;
                            f[8 * j + 4] += c.x; f[8 * j + 5] += c.y; f[8 * j + 6] += d.x; f[8 * j + 7] += d.y;
                        }
                    }
                    const bool relu = (p.flags & kEpiRelu) != 0;
#pragma unroll
                    for (int j = 0; j < CH / 8; ++j)
                        *reinterpret_cast<uint4*>(panel + (((ch16 + j) ^ (r_in_tile & 7)) << 4)) = pack8(f + 8 * j, relu, keep);
                }
                ptx::fence_proxy_async();          // generic-proxy smem writes -> visible to the TMA store
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                // accumulator drained: tell the leader's MMA warp (remote arrive for the second CTA of the pair)
                ptx::mbar_arrive_cluster(leader_tmem_empty + static_cast<uint32_t>(acc) * 8u);
                if (valid) ptx::mbar_arrive(&stage_ready[vit % kBufs]);
            }
            if (valid) ++vit;
            if (++acc == kAcc) { acc = 0; acc_phase ^= 1u; }
        }
    } else {
        // ------------------------------------------------------------ TMA store warp (warp 10)
        if (lane == 0) {
            int vit = 0;
            for (int pt = cluster_id; pt < total_pairs; pt += num_clusters) {
                int pm, n_tile;
                split_tile(pt, p.num_n_tiles, pm, n_tile);
                const int m_tile = 2 * pm + static_cast<int>(rank);
                if (m_tile >= p.num_m_tiles) continue;
                const int out_row_base = (p.tile_begin + m_tile) * kBlockM;
                const int buf = vit % kBufs;
                ptx::mbar_wait(&stage_ready[buf], (vit / kBufs) & 1);
#pragma unroll
                for (int pn = 0; pn < BN / 64; ++pn)
                    ptx::tma_store_2d(&tmap_out, epi_smem + buf * S::kEpiBytes + pn * (kBlockM * 128), n_tile * BN + pn * 64,
                                      out_row_base);
                ptx::bulk_commit_group();
                ptx::bulk_wait_read_all();
                ptx::mbar_arrive(&epi_free[buf]);
                ++vit;
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
