// Host-side launcher for conv_gemm_f16_kernel: TMA descriptor encoding (driver entry point fetched at run time,
// so the library links against cudart only) and per-BN dispatch.
#pragma once
#include <cstdio>
#include <cstring>
#include <string>

#include "conv_gemm.cuh"
#include "conv_gemm_2cta.cuh"
#include "conv1x1_pair.cuh"
#include "conv1x1_pair_split.cuh"

namespace sylph {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled get_tmap_encode() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
        fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }
    return fn;
}

// 2-D fp16 matrix [rows][cols] with row pitch `ld_elems`; box = [box_rows][64 cols], 128-byte swizzle, zero OOB fill.
// `ld_elems` may be smaller than `cols` (overlapping rows) -- used by the stem convolution.
inline int make_tmap_2d(CUtensorMap* m, const __half* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                        uint32_t box_rows, std::string* err) {
    PFN_tmapEncodeTiled enc = get_tmap_encode();
    if (enc == nullptr) {
        if (err) *err = "cuTensorMapEncodeTiled entry point unavailable";
        return 1;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_elems * sizeof(__half)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r));
        return 2;
    }
    return 0;
}

// 2-D fp16 matrix [rows][16] (32-byte rows): box = [box_rows][16], 32-byte swizzle -- the stem's space-to-depth input.
// `cols` = 32 for the split-operand layout ([16 hi | 16 lo] per row): the box stays 16 wide and starts at column 0 or 16.
inline int make_tmap_2d_k16(CUtensorMap* m, const __half* base, uint64_t rows, uint32_t box_rows, std::string* err,
                            uint64_t cols = 16) {
    PFN_tmapEncodeTiled enc = get_tmap_encode();
    if (enc == nullptr) {
        if (err) *err = "cuTensorMapEncodeTiled entry point unavailable";
        return 1;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * sizeof(__half)};
    cuuint32_t box[2] = {16, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (err) *err = "cuTensorMapEncodeTiled (k16) failed with CUresult " + std::to_string(static_cast<int>(r));
        return 2;
    }
    return 0;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: one flag per (kernel instantiation, device), so a
// second context on another GPU of the same process opts in as well.
struct PerDeviceOnce {
    bool done[64] = {};
    template <typename F>
    cudaError_t run(F&& f) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 64) return f();
        if (done[dev]) return cudaSuccess;
        e = f();
        if (e == cudaSuccess) done[dev] = true;
        return e;
    }
};

template <int BN, int STAGES, int EPI_BUFS, int HALO = 0, bool STEM16 = false, bool BRES = false, bool SPLIT = false, bool NM = false, bool QS = false>
inline cudaError_t launch_conv_gemm_bn(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tres,
                                       const CUtensorMap& tout, const GemmArgs& args, int num_sms, cudaStream_t stream) {
    using S = GemmSmem<BN, STAGES, EPI_BUFS, HALO, STEM16, BRES, SPLIT, NM, QS>;
    static_assert(S::kTotal <= 227 * 1024, "shared memory of this instantiation exceeds the 227 KB of a CTA");
    static PerDeviceOnce once;
    {
        cudaError_t e = once.run([] {
            return cudaFuncSetAttribute(conv_gemm_f16_kernel<BN, STAGES, EPI_BUFS, HALO, STEM16, BRES, SPLIT, NM, QS>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
        });
        if (e != cudaSuccess) return e;
    }
    const int total = args.num_m_tiles * args.num_n_tiles;
    if (total <= 0) return cudaSuccess;
    int grid = total < num_sms ? total : num_sms;
    // Generic pipeline with k-steps | STAGES: ring slot s always carries k-step (s mod ksteps), so a CTA that also keeps
    // its N tile never reloads a weight tile.  tile = blockIdx + j * grid and n_tile = tile mod num_n_tiles: the N tile
    // is fixed per CTA iff grid is a multiple of num_n_tiles -- give up at most num_n_tiles - 1 SMs for that.
    const int ksteps = args.taps * args.kblocks_per_tap;
    if (HALO == 0 && !STEM16 && !QS && args.num_n_tiles > 1 && ksteps <= STAGES && STAGES % ksteps == 0 && grid == num_sms &&
        grid > 4 * args.num_n_tiles && !(args.dbg_skip & 16))
        grid -= grid % args.num_n_tiles;
    return launch_k(conv_gemm_f16_kernel<BN, STAGES, EPI_BUFS, HALO, STEM16, BRES, SPLIT, NM, QS>, dim3(grid), dim3(S::kThreads), S::kTotal, stream, ta, tb, tres, tout, args);
}

constexpr int kStagedTwoBufMaxKSteps = 8;
constexpr bool kStagedConv1EightSlots = true;   // res2 conv1 (K = 256): 8 ring slots so that weight tiles stay put

// Direct-epilogue variants: BN in {16, 64, 128, 256}.
inline cudaError_t launch_conv_gemm(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args,
                                    int num_sms, cudaStream_t stream, bool split = false) {
    if (split) {
        switch (bn) {
            case 16: return launch_conv_gemm_bn<16, 8, 0, 0, false, false, true>(ta, tb, ta, ta, args, num_sms, stream);
            case 64: return launch_conv_gemm_bn<64, 8, 0, 0, false, false, true>(ta, tb, ta, ta, args, num_sms, stream);
            case 128: return launch_conv_gemm_bn<128, 6, 0, 0, false, false, true>(ta, tb, ta, ta, args, num_sms, stream);
            case 256: return launch_conv_gemm_bn<256, 4, 0, 0, false, false, true>(ta, tb, ta, ta, args, num_sms, stream);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (bn) {
        case 16: return launch_conv_gemm_bn<16, 8, 0>(ta, tb, ta, ta, args, num_sms, stream);
        case 64: return launch_conv_gemm_bn<64, 8, 0>(ta, tb, ta, ta, args, num_sms, stream);
        case 128: return launch_conv_gemm_bn<128, 6, 0>(ta, tb, ta, ta, args, num_sms, stream);
        case 256: return launch_conv_gemm_bn<256, 4, 0>(ta, tb, ta, ta, args, num_sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

// 3x3 halo pipeline (taps in the standard (dy, dx) row-major order): `ta` must be a map whose box has kBlockM + 2 rows.
// <BN, B slots, 0, A halo slots>: 3 x 17 KB + 5 x 32 KB = 211 KB for BN = 256.
inline cudaError_t launch_conv_gemm_halo(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args,
                                         int num_sms, cudaStream_t stream, bool allow_resident = true, bool split = false) {
    if (split) {
        // 64 -> 64 channels (res2 conv2) in split mode: K' = 192 per tap = 3 k-blocks over 18 distinct weight tiles (w_hi, w_lo of
        // the nine taps), all resident (144 KB) next to 4 halo A slots (68 KB)
        if (allow_resident && bn == 64 && args.kblocks_per_tap == 3 && args.num_n_tiles == 1 && args.a_wrap == 128)
            return launch_conv_gemm_bn<64, 18, 0, 4, false, true, true>(ta, tb, ta, ta, args, num_sms, stream);
        switch (bn) {
            case 16: return launch_conv_gemm_bn<16, 12, 0, 4, false, false, true>(ta, tb, ta, ta, args, num_sms, stream);
            case 64: return launch_conv_gemm_bn<64, 12, 0, 4, false, false, true>(ta, tb, ta, ta, args, num_sms, stream);
            case 128: return launch_conv_gemm_bn<128, 8, 0, 4, false, false, true>(ta, tb, ta, ta, args, num_sms, stream);
            case 256: return launch_conv_gemm_bn<256, 5, 0, 3, false, false, true>(ta, tb, ta, ta, args, num_sms, stream);
            default: return cudaErrorInvalidValue;
        }
    }
    // 64 -> 64 channels (res2 conv2): all nine 8 KB weight tiles stay resident, 8 halo A slots (72 + 136 KB)
    if (allow_resident && bn == 64 && args.kblocks_per_tap == 1 && args.num_n_tiles == 1)
        return launch_conv_gemm_bn<64, 9, 0, 8, false, true>(ta, tb, ta, ta, args, num_sms, stream);
    switch (bn) {
        case 16: return launch_conv_gemm_bn<16, 12, 0, 4>(ta, tb, ta, ta, args, num_sms, stream);
        case 64: return launch_conv_gemm_bn<64, 12, 0, 4>(ta, tb, ta, ta, args, num_sms, stream);
        case 128: return launch_conv_gemm_bn<128, 8, 0, 4>(ta, tb, ta, ta, args, num_sms, stream);
        case 256: return launch_conv_gemm_bn<256, 5, 0, 3>(ta, tb, ta, ta, args, num_sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

// N-merged split halo pipeline for the narrow 3x3 layers (conv_gemm.cuh, NM): `tb` is the map over the [taps][2 cout_pad][cin]
// weight matrix with BN-row boxes, args.kblocks_per_tap = 2 cin / 64, args.nm_lo_row = cout_pad.
// B slots hold 2 BN rows: BN = 64: 4 halo slots + 12 x 16 KB = 260 KB is too much -> 8 slots (196 KB); resident form for
// 64 -> 64 channels: 9 x 16 KB + 4 halo slots = 212 KB; BN = 128: 4 x 17 KB + 4 x 32 KB = 196 KB; BN = 16: 12 x 4 KB.
inline cudaError_t launch_conv_gemm_halo_nm(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args,
                                            int num_sms, cudaStream_t stream, bool allow_resident = true) {
    if (allow_resident && bn == 64 && args.kblocks_per_tap == 2 && args.num_n_tiles == 1)
        return launch_conv_gemm_bn<64, 9, 0, 4, false, true, true, true>(ta, tb, ta, ta, args, num_sms, stream);
    switch (bn) {
        case 16: return launch_conv_gemm_bn<16, 12, 0, 4, false, false, true, true>(ta, tb, ta, ta, args, num_sms, stream);
        case 64: return launch_conv_gemm_bn<64, 8, 0, 4, false, false, true, true>(ta, tb, ta, ta, args, num_sms, stream);
        case 128: return launch_conv_gemm_bn<128, 4, 0, 4, false, false, true, true>(ta, tb, ta, ta, args, num_sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

// Quad-stage split 1x1 kernels (conv_gemm.cuh, QS): args.kblocks_per_tap = cin / 64 (LOGICAL k-blocks), args.a_wrap = 2 cin,
// `tb` the usual map over the [w_hi | w_hi | w_lo] rows with BN-row boxes.  A stage is 32 KB of A + 2 BN x 128 B of weights:
//   staged, BN = 128: 2 stages x 64 KB + ONE 64 KB staging buffer (192 KB); staged, BN = 256: 1 stage x 96 KB + 128 KB;
//   register epilogue, BN = 256: 2 stages x 96 KB; BN = 128: 3 stages x 64 KB.
inline cudaError_t launch_conv_gemm_qs(int bn, bool staged, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tres,
                                       const CUtensorMap& tout, const GemmArgs& args, int num_sms, cudaStream_t stream) {
    if (staged) {
        if (bn == 128) return launch_conv_gemm_bn<128, 2, 1, 0, false, false, true, false, true>(ta, tb, tres, tout, args, num_sms, stream);
        if (bn == 256) return launch_conv_gemm_bn<256, 1, 1, 0, false, false, true, false, true>(ta, tb, tres, tout, args, num_sms, stream);
        return cudaErrorInvalidValue;
    }
    if (bn == 128) return launch_conv_gemm_bn<128, 3, 0, 0, false, false, true, false, true>(ta, tb, ta, ta, args, num_sms, stream);
    if (bn == 256) return launch_conv_gemm_bn<256, 2, 0, 0, false, false, true, false, true>(ta, tb, ta, ta, args, num_sms, stream);
    return cudaErrorInvalidValue;
}

// CTA-pair (cta_group::2) 3x3 halo convolution, N tiles of BN in {64, 128, 256}: `ta` box = kBlockM + 2 rows,
// `tb` box = BN / 2 rows (each CTA of the pair holds half of the output channels of a B tile).
template <int BN, bool SPLIT = false>
inline cudaError_t launch_conv3x3_pair_bn(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args, int num_sms,
                                          cudaStream_t stream) {
    constexpr int kHalo = BN == 256 ? 3 : 6;
    constexpr int kBSlots = BN == 256 ? 8 : 12;
    using S = Gemm2Smem<kHalo, kBSlots, BN>;
    static PerDeviceOnce once;
    {
        cudaError_t e = once.run([] {
            return cudaFuncSetAttribute(conv3x3_pair_kernel<kHalo, kBSlots, BN, SPLIT>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
        });
        if (e != cudaSuccess) return e;
    }
    const int pairs = ((args.num_m_tiles + 1) / 2) * args.num_n_tiles;
    if (pairs <= 0) return cudaSuccess;
    const int clusters = pairs < num_sms / 2 ? pairs : num_sms / 2;
    return launch_k(conv3x3_pair_kernel<kHalo, kBSlots, BN, SPLIT>, dim3(clusters * 2), dim3(S::kThreads), S::kTotal, stream, ta, tb, args);
}

inline cudaError_t launch_conv3x3_pair(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& args, int num_sms,
                                       cudaStream_t stream, int bn = 256, bool split = false) {
    if (split) {
        switch (bn) {
            case 64: return launch_conv3x3_pair_bn<64, true>(ta, tb, args, num_sms, stream);
            case 128: return launch_conv3x3_pair_bn<128, true>(ta, tb, args, num_sms, stream);
            case 256: return launch_conv3x3_pair_bn<256, true>(ta, tb, args, num_sms, stream);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (bn) {
        case 64: return launch_conv3x3_pair_bn<64>(ta, tb, args, num_sms, stream);
        case 128: return launch_conv3x3_pair_bn<128>(ta, tb, args, num_sms, stream);
        case 256: return launch_conv3x3_pair_bn<256>(ta, tb, args, num_sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

// Staged (TMA in / TMA out) epilogue, fp16 output: `tres` / `tout` are [rows][C] maps with 128 x 64 boxes.
// Every fp16 store of the direct epilogue is a 16-byte piece of a different row (a thread owns a row), which tops out
// near 1.3 TB/s of write bandwidth; staging the tile in swizzled shared memory and leaving by TMA writes full lines.
inline cudaError_t launch_conv_gemm_staged(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tres,
                                           const CUtensorMap& tout, const GemmArgs& args, int num_sms,
                                           cudaStream_t stream, int variant = 0, bool split = false) {
    // Split-operand mode: a staged tile is a hi half and a lo half, so N tiles are at most 128 channels wide
    // (2 buffers x 64 KB next to a 3-stage ring) and the epilogue moves 4 B per output element either way.
    if (split) {
        if (bn == 64) return launch_conv_gemm_bn<64, 6, 2, 0, false, false, true>(ta, tb, tres, tout, args, num_sms, stream);
        if (bn == 128) return launch_conv_gemm_bn<128, 3, 2, 0, false, false, true>(ta, tb, tres, tout, args, num_sms, stream);
        // 256-wide split tile: ONE 128 KB staging buffer (hi + lo) next to a 2-stage ring -- halves the A re-reads from L2 of
        // the wide layers (N >= 512), at the price of no residual prefetch under the previous tile's epilogue
        if (bn == 256) return launch_conv_gemm_bn<256, 2, 1, 0, false, false, true>(ta, tb, tres, tout, args, num_sms, stream);
        return cudaErrorInvalidValue;
    }
    // BN = 256, variant 1: 2 mainloop stages + 2 staging buffers (the residual of tile i+1 streams in during the
    // epilogue of tile i); variant 2: 3 stages + 1 staging buffer for long K.  0 = pick by the number of k-steps.
    if (bn == 64) {
        // K = 256 (bottleneck conv1 of res2): 4 k-steps; 8 slots = two tiles in flight with resident weight tiles
        if (variant == 3 || (variant == 0 && args.taps * args.kblocks_per_tap == 4 && kStagedConv1EightSlots))
            return launch_conv_gemm_bn<64, 8, 2>(ta, tb, tres, tout, args, num_sms, stream);
        return launch_conv_gemm_bn<64, 6, 2>(ta, tb, tres, tout, args, num_sms, stream);
    }
    if (bn == 128) return launch_conv_gemm_bn<128, 4, 2>(ta, tb, tres, tout, args, num_sms, stream);
    if (bn != 256) return cudaErrorInvalidValue;
    if (variant == 0) variant = (args.taps * args.kblocks_per_tap <= kStagedTwoBufMaxKSteps) ? 1 : 2;
    if (variant == 1) return launch_conv_gemm_bn<256, 2, 2>(ta, tb, tres, tout, args, num_sms, stream);
    return launch_conv_gemm_bn<256, 3, 1>(ta, tb, tres, tout, args, num_sms, stream);
}

// CTA-pair 1x1 convolution with the staged epilogue (BN = 256, taps = 1): `tb` box = 128 rows (each CTA of the pair holds
// half of the 256 output channels of a B tile), `ta` / `tres` / `tout` as for launch_conv_gemm_staged.
// 3 ring slots x 32 KB + 2 staging buffers x 64 KB = 224 KB.
inline cudaError_t launch_conv1x1_pair_staged(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tres,
                                              const CUtensorMap& tout, const GemmArgs& args, int num_sms, cudaStream_t stream) {
    using S = Pair1x1Smem<3, 2>;
    static PerDeviceOnce once;
    {
        cudaError_t e = once.run([] {
            return cudaFuncSetAttribute(conv1x1_pair_staged_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
        });
        if (e != cudaSuccess) return e;
    }
    if (args.taps != 1) return cudaErrorInvalidValue;
    const int pairs = ((args.num_m_tiles + 1) / 2) * args.num_n_tiles;
    if (pairs <= 0) return cudaSuccess;
    const int clusters = pairs < num_sms / 2 ? pairs : num_sms / 2;
    return launch_k(conv1x1_pair_staged_kernel<3, 2>, dim3(clusters * 2), dim3(S::kThreads), S::kTotal, stream, ta, tb, tres, tout, args);
}

// Split-mode CTA-pair 1x1 convolution with the chunked staged epilogue (conv1x1_pair_split.cuh; BN = 256, taps = 1): `tb` box =
// 128 rows, `ta` / `tres` / `tout` as for launch_conv_gemm_staged.  qs = false: K' = 3C loop (args.kblocks_per_tap = 3C / 64),
// 3 stages x 32 KB + 4 chunk buffers x 32 KB; qs = true: quad stages (args.kblocks_per_tap = C / 64), 2 x 64 KB + 3 x 32 KB.
template <bool QS>
inline cudaError_t launch_conv1x1_pair_split_t(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tres,
                                               const CUtensorMap& tout, const GemmArgs& args, int num_sms, cudaStream_t stream) {
    constexpr int kStages = QS ? 2 : 3, kChunkBufs = QS ? 3 : 4;
    using S = PairSplitSmem<kStages, kChunkBufs, QS>;
    static_assert(S::kTotal <= 227 * 1024, "shared memory of this instantiation exceeds the 227 KB of a CTA");
    static PerDeviceOnce once;
    {
        cudaError_t e = once.run([] {
            return cudaFuncSetAttribute(conv1x1_pair_split_kernel<kStages, kChunkBufs, QS>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
        });
        if (e != cudaSuccess) return e;
    }
    if (args.taps != 1) return cudaErrorInvalidValue;
    const int pairs = ((args.num_m_tiles + 1) / 2) * args.num_n_tiles;
    if (pairs <= 0) return cudaSuccess;
    const int clusters = pairs < num_sms / 2 ? pairs : num_sms / 2;
    return launch_k(conv1x1_pair_split_kernel<kStages, kChunkBufs, QS>, dim3(clusters * 2), dim3(S::kThreads), S::kTotal, stream, ta, tb, tres, tout, args);
}
inline cudaError_t launch_conv1x1_pair_split(bool qs, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tres,
                                             const CUtensorMap& tout, const GemmArgs& args, int num_sms, cudaStream_t stream) {
    return qs ? launch_conv1x1_pair_split_t<true>(ta, tb, tres, tout, args, num_sms, stream)
              : launch_conv1x1_pair_split_t<false>(ta, tb, tres, tout, args, num_sms, stream);
}

// Stem: 4 vertical taps x (one 131 x 16 A box, 4 horizontal K = 16 MMAs), staged epilogue, Cout = 64; the weights
// (4 x 8 KB) stay resident in shared memory and 16 A slots keep four output tiles of loads in flight.
// `ta` from make_tmap_2d_k16 (box kBlockM + 3 rows), `tb` = [4 * 64][64] weights with 64-row boxes.
inline cudaError_t launch_conv_gemm_stem16(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout,
                                           const GemmArgs& args, int num_sms, cudaStream_t stream, bool split = false, bool nm = false) {
    // split + nm: four resident [w_hi ; w_lo] tiles of 128 rows, two A boxes (hi, lo) per vertical tap, 32 instead of 48
    // instructions per tile (args.nm_lo_row = row offset of the w_lo tiles in the weight matrix)
    if (split && nm) return launch_conv_gemm_bn<64, 4, 2, 16, true, false, true, true>(ta, tb, tout, tout, args, num_sms, stream);
    // split: 8 resident weight tiles (w_hi, w_lo of the four vertical taps), three A boxes per vertical tap
    if (split) return launch_conv_gemm_bn<64, 8, 2, 16, true, false, true>(ta, tb, tout, tout, args, num_sms, stream);
    return launch_conv_gemm_bn<64, 4, 2, 16, true>(ta, tb, tout, tout, args, num_sms, stream);
}

// IEEE fp32 -> fp16, round to nearest even, saturating to the finite range (host-side weight preparation).
inline uint16_t float_to_half_bits(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7FFFFFFFu;
    if (x >= 0x7F800000u) return static_cast<uint16_t>(sign | (x > 0x7F800000u ? 0x7E00u : 0x7BFFu));
    if (x >= 0x477FF000u) return static_cast<uint16_t>(sign | 0x7BFFu);  // >= 65520 rounds past the largest finite half
    if (x < 0x33000001u) return static_cast<uint16_t>(sign);              // below half of the smallest subnormal
    int e = static_cast<int>(x >> 23) - 127;
    uint32_t m = (x & 0x7FFFFFu) | 0x800000u;
    int shift;
    uint32_t base;
    if (e < -14) { shift = 13 + (-14 - e); base = 0; }           // subnormal half
    else { shift = 13; base = static_cast<uint32_t>(e + 15) << 10; m &= 0x7FFFFFu; }
    const uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
    uint32_t h = base + q;
    if (rem > halfway || (rem == halfway && (h & 1u))) ++h;
    return static_cast<uint16_t>(sign | h);
}

// fp16 bits -> fp32 (exact), for the lo half of split weights: w_lo = rn16(w - float(w_hi)).
inline float half_bits_to_float(uint16_t h) {
    const uint32_t sign = static_cast<uint32_t>(h & 0x8000u) << 16;
    const uint32_t e = (h >> 10) & 0x1Fu, m = h & 0x3FFu;
    uint32_t x;
    if (e == 0) {
        if (m == 0) { x = sign; }
        else {   // subnormal half: normalise
            int sh = 0;
            uint32_t mm = m;
            while (!(mm & 0x400u)) { mm <<= 1; ++sh; }
            x = sign | (static_cast<uint32_t>(127 - 15 - sh + 1) << 23) | ((mm & 0x3FFu) << 13);
        }
    } else if (e == 31) {
        x = sign | 0x7F800000u | (m << 13);
    } else {
        x = sign | ((e + 112u) << 23) | (m << 13);
    }
    float f;
    memcpy(&f, &x, 4);
    return f;
}

}  // namespace sylph
