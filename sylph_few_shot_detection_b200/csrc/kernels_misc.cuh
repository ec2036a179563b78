// HBM-bound helper kernels around the tensor-core convolutions: image normalise/pad/space-to-depth, max-pool,
// stride-2 gathers, FPN top-down add, GroupNorm finalize/apply, NCHW import/export.  All operate on the flat
// zero-bordered NHWC fp16 planes described in conv_gemm.cuh; channel vectors move as 16-byte lanes (8 halves).
#pragma once
#include "conv_gemm.cuh"

namespace sylph {

// Regular plane geometry shared by all images of a batch at one resolution.
struct PlaneGeom {
    int row_base;      // first row of image 0 inside the buffer
    int rows_per_img;  // multiple of 128
    int Wp, pad, H, W;
};

__device__ __forceinline__ size_t plane_row(const PlaneGeom& g, int n, int y, int x) {
    return static_cast<size_t>(g.row_base) + static_cast<size_t>(n) * g.rows_per_img +
           static_cast<size_t>(y + g.pad) * g.Wp + (x + g.pad);
}

// Split-operand ("exact") storage: a row of C logical channels is [C hi | C lo] fp16 (pitch 2C).  Kernels below take
// the row pitch `ld` and the offset `lo` of the lo half (0 = plain fp16 rows of the fast mode).
__device__ __forceinline__ void load8f(const __half* row, int c, int lo, float* v) {
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(row + c));
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    add8(v, h);
    if (lo) add8(v, __ldg(reinterpret_cast<const uint4*>(row + lo + c)));   // hi + lo is exact in fp32
}
__device__ __forceinline__ void store8f(__half* row, int c, int lo, const float* v, bool relu) {
    if (lo) {
        uint4 h, l;
        split8(v, relu, true, h, l);
        *reinterpret_cast<uint4*>(row + c) = h;
        *reinterpret_cast<uint4*>(row + lo + c) = l;
    } else {
        *reinterpret_cast<uint4*>(row + c) = pack8(v, relu, true);
    }
}

__device__ __forceinline__ uint4 hmax8(uint4 a, uint4 b) {
    uint4 r;
    *reinterpret_cast<__half2*>(&r.x) = __hmax2(*reinterpret_cast<__half2*>(&a.x), *reinterpret_cast<__half2*>(&b.x));
    *reinterpret_cast<__half2*>(&r.y) = __hmax2(*reinterpret_cast<__half2*>(&a.y), *reinterpret_cast<__half2*>(&b.y));
    *reinterpret_cast<__half2*>(&r.z) = __hmax2(*reinterpret_cast<__half2*>(&a.z), *reinterpret_cast<__half2*>(&b.z));
    *reinterpret_cast<__half2*>(&r.w) = __hmax2(*reinterpret_cast<__half2*>(&a.w), *reinterpret_cast<__half2*>(&b.w));
    return r;
}

// ------------------------------------------------------------------------------------------------ image prep
// Replaces (x - pixel_mean) / pixel_std + ImageList.from_tensors zero padding
// (sylph/modeling/meta_arch/meta_one_stage_detector.py:174-178), fused with the 2x2 space-to-depth re-layout that
// turns the 7x7/2 stem convolution into a 4x4/1 convolution over 12 (padded to 16) channels.
struct ImageDesc {
    const void* ptr;  // (3, h, w) fp32 or uint8
    int h, w;
    int is_u8;
};

__global__ void prep_stem_input_kernel(const ImageDesc* __restrict__ imgs, __half* __restrict__ out, PlaneGeom g,
                                       int n_images, float m0, float m1, float m2, float s0, float s1, float s2, int split) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const long long total = static_cast<long long>(n_images) * g.H * g.W;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int x2 = static_cast<int>(i % g.W);
        const int y2 = static_cast<int>((i / g.W) % g.H);
        const int n = static_cast<int>(i / (static_cast<long long>(g.W) * g.H));
        const ImageDesc im = imgs[n];
        const float mean[3] = {m0, m1, m2};
        const float stdv[3] = {s0, s1, s2};
        float v[16];
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int iy = 2 * y2 + dy, ix = 2 * x2 + dx;
                const bool in = iy < im.h && ix < im.w;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float t = 0.f;
                    if (in) {
                        const size_t off = (static_cast<size_t>(c) * im.h + iy) * im.w + ix;
                        const float px = im.is_u8 ? static_cast<float>(__ldg(static_cast<const unsigned char*>(im.ptr) + off))
                                                  : __ldg(static_cast<const float*>(im.ptr) + off);
                        t = (px - mean[c]) / stdv[c];
                    }
                    v[(dy * 2 + dx) * 3 + c] = t;
                }
            }
        v[12] = v[13] = v[14] = v[15] = 0.f;
        if (split) {   // rows of [16 hi | 16 lo]
            __half* row = out + plane_row(g, n, y2, x2) * 32;
            store8f(row, 0, 16, v, false);
            store8f(row, 8, 16, v + 8, false);
            continue;
        }
        uint4* o = reinterpret_cast<uint4*>(out + plane_row(g, n, y2, x2) * 16);
        o[0] = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
        o[1] = make_uint4(pack_half2(v[8], v[9]), pack_half2(v[10], v[11]), 0u, 0u);
    }
}

// uint8 images (what the data mapper hands over): the same arithmetic through a 3 x 256 table of the fp16 results
// ((v - mean) / std rounded to nearest -- bit-identical to the kernel above, no divisions in the pixel loop), and the
// six input rows of a 256-pixel output segment staged in shared memory with aligned 4-byte loads (3 coalesced word
// loads per thread instead of 12 scattered byte loads: the byte loads, not HBM, bounded the kernel at 1.8 TB/s).
// grid = (ceil(W2 / 256), rows of work); block = 256.
__global__ void __launch_bounds__(256)
prep_stem_input_u8_kernel(const ImageDesc* __restrict__ imgs, __half* __restrict__ out, PlaneGeom g, int n_images,
                          float m0, float m1, float m2, float s0, float s1, float s2, int split) {
    ptx::griddep_launch();
    __shared__ unsigned short lut[3][256];
    __shared__ unsigned short lut_lo[3][256];   // split mode: rn16(x - hi), x = (v - mean) / std in fp32
    __shared__ __align__(16) unsigned char rows[6][528];   // 512 pixels + up to 3 bytes of misalignment, padded
    {
        const float mean[3] = {m0, m1, m2};
        const float stdv[3] = {s0, s1, s2};
        for (int i = threadIdx.x; i < 768; i += 256) {
            const int c = i >> 8, v = i & 255;
            const float x = (static_cast<float>(v) - mean[c]) / stdv[c];
            const uint32_t hi = pack_half2(x, 0.f);
            lut[c][v] = static_cast<unsigned short>(hi & 0xFFFFu);
            lut_lo[c][v] = static_cast<unsigned short>(pack_half2(x - unpack_half2(hi).x, 0.f) & 0xFFFFu);
        }
    }
    ptx::griddep_wait();
    const int x0 = blockIdx.x * 256;              // first output pixel of this block's segment
    const int n_rows = n_images * g.H;
    for (int ry = blockIdx.y; ry < n_rows; ry += gridDim.y) {
        const int n = ry / g.H, y2 = ry - n * g.H;
        const ImageDesc im = imgs[n];
        const unsigned char* base = static_cast<const unsigned char*>(im.ptr);
        const size_t img_bytes = static_cast<size_t>(3) * im.h * im.w;
        const int px0 = 2 * x0;                   // first input pixel
        const int npx = min(512, im.w - px0);     // input pixels of this segment that exist
        __syncthreads();                          // previous iteration's readers are done with `rows` (and the table is built)
        int shift[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const int c = r >> 1, iy = 2 * y2 + (r & 1);
            shift[r] = 0;
            if (iy < im.h && npx > 0) {
                const size_t a = (static_cast<size_t>(c) * im.h + iy) * im.w + px0;
                const size_t a4 = (reinterpret_cast<size_t>(base) + a) & ~static_cast<size_t>(3);
                shift[r] = static_cast<int>(reinterpret_cast<size_t>(base) + a - a4);
                const int nwords = (shift[r] + npx + 3) >> 2;
                const size_t lo = reinterpret_cast<size_t>(base), hi = lo + img_bytes;
                for (int wd = threadIdx.x; wd < nwords; wd += 256) {
                    const size_t wa = a4 + 4 * static_cast<size_t>(wd);
                    uint32_t word;
                    if (wa >= lo && wa + 4 <= hi) {
                        word = __ldg(reinterpret_cast<const uint32_t*>(wa));
                    } else {   // the aligned word straddles an end of the image tensor: byte loads of what exists
                        word = 0;
                        for (int b = 0; b < 4; ++b) {
                            const size_t ba = wa + b;
                            if (ba >= lo && ba < hi)
                                word |= static_cast<uint32_t>(__ldg(reinterpret_cast<const unsigned char*>(ba))) << (8 * b);
                        }
                    }
                    *reinterpret_cast<uint32_t*>(&rows[r][4 * wd]) = word;
                }
            }
        }
        __syncthreads();
        const int x2 = x0 + threadIdx.x;
        if (x2 < g.W) {
            unsigned short h[12], l[12];
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const int iy = 2 * y2 + dy, ix = 2 * x2 + dx;
                    const bool in = iy < im.h && ix < im.w;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int r = c * 2 + dy;
                        const int px = in ? rows[r][shift[r] + 2 * threadIdx.x + dx] : 0;
                        h[(dy * 2 + dx) * 3 + c] = in ? lut[c][px] : static_cast<unsigned short>(0);
                        l[(dy * 2 + dx) * 3 + c] = (in && split) ? lut_lo[c][px] : static_cast<unsigned short>(0);
                    }
                }
            uint4* o = reinterpret_cast<uint4*>(out + plane_row(g, n, y2, x2) * (split ? 32 : 16));
            auto pk = [](unsigned short lo, unsigned short hi) { return static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16); };
            o[0] = make_uint4(pk(h[0], h[1]), pk(h[2], h[3]), pk(h[4], h[5]), pk(h[6], h[7]));
            o[1] = make_uint4(pk(h[8], h[9]), pk(h[10], h[11]), 0u, 0u);
            if (split) {
                o[2] = make_uint4(pk(l[0], l[1]), pk(l[2], l[3]), pk(l[4], l[5]), pk(l[6], l[7]));
                o[3] = make_uint4(pk(l[8], l[9]), pk(l[10], l[11]), 0u, 0u);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ image prep, exact mode
// Exact (split-operand) mode feeds the stem with RE-CENTRED INTEGER pixels instead of normalised ones:
//   u = (v - m_r) / 256,  m_r = round(pixel_mean)   -- for uint8 images an 8-bit integer times 2^-8: EXACT in one fp16, lo = 0
//   y = sum_taps (256 w s / std) u + [b_bn - sum_taps (w s / std)(mean - m_r)]      (weights and bias folded in double, engine.cu)
// which equals the reference's conv((v - mean) / std) up to the fp32 rounding of either side.  The reference pads the NORMALISED
// image with zeros (ImageList.from_tensors + the convolution's own padding), i.e. with pixel value `mean`: every position outside
// the image -- including the 2-pixel border ring of the plane that the convolution's padding reads -- holds
// pad_c = rn16((mean_c - m_r_c) / 256) instead of 0, so no position-dependent correction is needed (the ring is written here on
// every pass; the error of rounding pad_c to fp16 is < 3e-7 of the pixel range, at padded taps only).
// The stem then needs ONE pass over a_hi for uint8 images (a_lo == 0): 16 instead of 32 instructions per tile.
// Float images run both passes; integer-valued float images give bit-identical features (the a_lo pass adds exact zeros).
struct CentreParams {
    float m_r[3];     // round(pixel_mean)
    float mean[3], stdv[3];
    float pad[3];     // (mean - m_r) / 256
    int denorm;       // float input is already (x - mean) / std (plugin boundary): v = x * std + mean first
};

__device__ __forceinline__ void write_ring_pixels(__half* __restrict__ out, const PlaneGeom& g, int n_images, const CentreParams& cp,
                                                  long long first, long long stride) {
    // border ring of every plane: (H + 4) x (W + 4) minus the interior, rows of [16 hi | 16 lo]
    const int ring_w = g.W + 4, ring_h = g.H + 4;
    const long long per_img = 2LL * 2 * ring_w + 2LL * 2 * g.H;   // two full rows above + below, two columns left + right
    uint4 v0, v1;
    {
        const uint32_t p0 = pack_half2(cp.pad[0], cp.pad[1]) , p1 = pack_half2(cp.pad[2], cp.pad[0]), p2 = pack_half2(cp.pad[1], cp.pad[2]);
        v0 = make_uint4(p0, p1, p2, p0);   // channels (dy, dx, c) = 12 values: c cycles 0 1 2 0 1 2 ...
        v1 = make_uint4(p1, p2, 0u, 0u);
    }
    for (long long i = first; i < per_img * n_images; i += stride) {
        const int n = static_cast<int>(i / per_img);
        long long r = i - n * per_img;
        int y, x;
        if (r < 2LL * ring_w) { y = static_cast<int>(r / ring_w) - 2; x = static_cast<int>(r % ring_w) - 2; }
        else if (r < 4LL * ring_w) { r -= 2LL * ring_w; y = g.H + static_cast<int>(r / ring_w); x = static_cast<int>(r % ring_w) - 2; }
        else { r -= 4LL * ring_w; y = static_cast<int>(r / 4); const int q = static_cast<int>(r % 4); x = q < 2 ? q - 2 : g.W + q - 2; }
        (void)ring_h;
        uint4* o = reinterpret_cast<uint4*>(out + plane_row(g, n, y, x) * 32);
        o[0] = v0; o[1] = v1;
        o[2] = make_uint4(0u, 0u, 0u, 0u); o[3] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// float (or uint8 through the slow path) images -> [16 hi | 16 lo] rows of u = (v - m_r) / 256
__global__ void prep_stem_centred_kernel(const ImageDesc* __restrict__ imgs, __half* __restrict__ out, PlaneGeom g, int n_images,
                                         CentreParams cp) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const long long total = static_cast<long long>(n_images) * g.H * g.W;
    const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = tid; i < total; i += nthreads) {
        const int x2 = static_cast<int>(i % g.W);
        const int y2 = static_cast<int>((i / g.W) % g.H);
        const int n = static_cast<int>(i / (static_cast<long long>(g.W) * g.H));
        const ImageDesc im = imgs[n];
        float v[16];
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int iy = 2 * y2 + dy, ix = 2 * x2 + dx;
                const bool in = iy < im.h && ix < im.w;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float t = __half2float(__float2half_rn(cp.pad[c]));   // the fp16 pad value, lo = 0, as in the uint8 kernel
                    if (in) {
                        const size_t off = (static_cast<size_t>(c) * im.h + iy) * im.w + ix;
                        float px = im.is_u8 ? static_cast<float>(__ldg(static_cast<const unsigned char*>(im.ptr) + off))
                                            : __ldg(static_cast<const float*>(im.ptr) + off);
                        if (cp.denorm) px = __fmaf_rn(px, cp.stdv[c], cp.mean[c]);
                        t = (px - cp.m_r[c]) * (1.f / 256.f);
                    }
                    v[(dy * 2 + dx) * 3 + c] = t;
                }
            }
        v[12] = v[13] = v[14] = v[15] = 0.f;
        __half* row = out + plane_row(g, n, y2, x2) * 32;
        store8f(row, 0, 16, v, false);
        store8f(row, 8, 16, v + 8, false);
    }
    write_ring_pixels(out, g, n_images, cp, tid, nthreads);
}

// uint8 images: table of the (exact) fp16 values, six input rows of a 256-pixel output segment staged in shared memory (see
// prep_stem_input_u8_kernel); writes the hi half of every row only -- the stem never reads the lo half on this path.
__global__ void __launch_bounds__(256)
prep_stem_centred_u8_kernel(const ImageDesc* __restrict__ imgs, __half* __restrict__ out, PlaneGeom g, int n_images, CentreParams cp,
                            int zero_lo) {   // zero_lo: the stem variant in use reads the lo half (SYLPH_NM=0): keep it zero
    ptx::griddep_launch();
    __shared__ unsigned short lut[3][256];
    __shared__ unsigned short padv[3];
    __shared__ __align__(16) unsigned char rows[2][6][528];   // double-buffered: the loads of segment i + 1 fly while i is converted
    for (int i = threadIdx.x; i < 768; i += 256) {
        const int c = i >> 8, v = i & 255;
        lut[c][v] = static_cast<unsigned short>(pack_half2((static_cast<float>(v) - cp.m_r[c]) * (1.f / 256.f), 0.f) & 0xFFFFu);
    }
    if (threadIdx.x < 3) padv[threadIdx.x] = static_cast<unsigned short>(pack_half2(cp.pad[threadIdx.x], 0.f) & 0xFFFFu);
    ptx::griddep_wait();
    const int x0 = blockIdx.x * 256;
    const int px0 = 2 * x0;                           // first input pixel of this block's 512-pixel segment
    const int n_rows = n_images * g.H;
    // input row r (0..5 = channel r / 2, image row 2 y2 + (r & 1)) of work item ry: aligned word range and byte shift
    auto row_span = [&](int ry, int r, size_t& a4, int& shift, int& nwords, size_t& lo_b, size_t& hi_b) {
        const int n = ry / g.H, y2 = ry - n * g.H;
        const ImageDesc im = imgs[n];
        const int c = r >> 1, iy = 2 * y2 + (r & 1);
        const int npx = min(512, im.w - px0);
        nwords = 0; shift = 0; a4 = 0;
        lo_b = reinterpret_cast<size_t>(im.ptr);
        hi_b = lo_b + static_cast<size_t>(3) * im.h * im.w;
        if (iy < im.h && npx > 0) {
            const size_t a = lo_b + (static_cast<size_t>(c) * im.h + iy) * im.w + px0;
            a4 = a & ~static_cast<size_t>(3);
            shift = static_cast<int>(a - a4);
            nwords = (shift + npx + 3) >> 2;
        }
    };
    // every row has at most (3 + 512 + 3) / 4 = 130 words: thread t < 130 carries word t of each of the six rows
    auto fetch = [&](int ry, uint32_t* w) {
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            size_t a4, lo_b, hi_b;
            int shift, nwords;
            row_span(ry, r, a4, shift, nwords, lo_b, hi_b);
            w[r] = 0;
            if (static_cast<int>(threadIdx.x) < nwords) {
                const size_t wa = a4 + 4 * static_cast<size_t>(threadIdx.x);
                if (wa >= lo_b && wa + 4 <= hi_b) {
                    w[r] = __ldg(reinterpret_cast<const uint32_t*>(wa));
                } else {   // the aligned word straddles an end of the image tensor: byte loads of what exists
                    for (int b = 0; b < 4; ++b) {
                        const size_t ba = wa + b;
                        if (ba >= lo_b && ba < hi_b)
                            w[r] |= static_cast<uint32_t>(__ldg(reinterpret_cast<const unsigned char*>(ba))) << (8 * b);
                    }
                }
            }
        }
    };
    auto stash = [&](int buf, const uint32_t* w) {
        if (threadIdx.x < 132) {
#pragma unroll
            for (int r = 0; r < 6; ++r) *reinterpret_cast<uint32_t*>(&rows[buf][r][4 * threadIdx.x]) = w[r];
        }
    };
    int cur = 0;
    uint32_t w[6];
    if (static_cast<int>(blockIdx.y) < n_rows) {
        fetch(blockIdx.y, w);
        stash(0, w);
    }
    __syncthreads();
    for (int ry = blockIdx.y; ry < n_rows; ry += gridDim.y) {
        const int nxt = ry + gridDim.y;
        if (nxt < n_rows) fetch(nxt, w);              // global loads in flight during the conversion below
        const int n = ry / g.H, y2 = ry - n * g.H;
        const ImageDesc im = imgs[n];
        const int x2 = x0 + threadIdx.x;
        if (x2 < g.W) {
            int shift[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                size_t a4, lo_b, hi_b;
                int nwords;
                row_span(ry, r, a4, shift[r], nwords, lo_b, hi_b);
            }
            unsigned short h[12];
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const int iy = 2 * y2 + dy, ix = 2 * x2 + dx;
                    const bool in = iy < im.h && ix < im.w;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int r = c * 2 + dy;
                        const int px = in ? rows[cur][r][shift[r] + 2 * threadIdx.x + dx] : 0;
                        h[(dy * 2 + dx) * 3 + c] = in ? lut[c][px] : padv[c];
                    }
                }
            uint4* o = reinterpret_cast<uint4*>(out + plane_row(g, n, y2, x2) * 32);
            auto pk = [](unsigned short lo, unsigned short hi) { return static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16); };
            o[0] = make_uint4(pk(h[0], h[1]), pk(h[2], h[3]), pk(h[4], h[5]), pk(h[6], h[7]));
            o[1] = make_uint4(pk(h[8], h[9]), pk(h[10], h[11]), 0u, 0u);
            if (zero_lo) { o[2] = make_uint4(0u, 0u, 0u, 0u); o[3] = make_uint4(0u, 0u, 0u, 0u); }
        }
        if (nxt < n_rows) stash(cur ^ 1, w);          // the other buffer: its readers finished before the previous barrier
        __syncthreads();
        cur ^= 1;
    }
    const long long tid = (static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    write_ring_pixels(out, g, n_images, cp, tid, static_cast<long long>(gridDim.x) * gridDim.y * blockDim.x);
}

// ------------------------------------------------------------------------------------------------ max-pool 3x3 / 2
// detectron2 BasicStem max_pool2d(kernel 3, stride 2, padding 1) on post-ReLU (>= 0) activations: the zero border of
// the input plane stands in for the -inf padding.
// `lo` > 0 (split mode): rows are [C hi | C lo]; the maximum is taken over hi + lo (exact in fp32) and split again.
__global__ void maxpool3x3s2_kernel(const __half* __restrict__ in, __half* __restrict__ out, PlaneGeom gi, PlaneGeom go,
                                    int n_images, int C, int lo) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int c8n = C / 8;
    if (lo) {
        // a thread walks a strip of kStrip output rows of one (column, 8-channel group): the horizontal 3-tap maximum of input row
        // 2 oy + 1 is also the first row of output oy + 1, so an output costs 6 taps (12 loads) instead of 9 (18)
        constexpr int kStrip = 8;
        const int ld = 2 * C;
        const int strips = (go.H + kStrip - 1) / kStrip;
        const long long total = static_cast<long long>(n_images) * strips * go.W * c8n;
        auto rowmax = [&](int n, int iy, int ox, int c8, float* m) {
#pragma unroll
            for (int j = 0; j < 8; ++j) m[j] = 0.f;
            if (iy >= gi.H + gi.pad) return;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = 2 * ox - 1 + kx;
                if (ix < gi.W + gi.pad) {
                    float v[8];
                    load8f(in + plane_row(gi, n, iy, ix) * ld, c8 * 8, lo, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
                }
            }
        };
        for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
             i += static_cast<long long>(gridDim.x) * blockDim.x) {
            const int c8 = static_cast<int>(i % c8n);
            long long r = i / c8n;
            const int ox = static_cast<int>(r % go.W);
            r /= go.W;
            const int oy0 = static_cast<int>(r % strips) * kStrip;
            const int n = static_cast<int>(r / strips);
            float carry[8];
            rowmax(n, 2 * oy0 - 1, ox, c8, carry);   // row -1 of the first strip is the zero border
            for (int k = 0; k < kStrip; ++k) {
                const int oy = oy0 + k;
                if (oy >= go.H) break;
                float a[8], b[8];
                rowmax(n, 2 * oy, ox, c8, a);
                rowmax(n, 2 * oy + 1, ox, c8, b);
                float m[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = fmaxf(fmaxf(carry[j], a[j]), b[j]);
                store8f(out + plane_row(go, n, oy, ox) * ld, c8 * 8, lo, m, false);
#pragma unroll
                for (int j = 0; j < 8; ++j) carry[j] = b[j];
            }
        }
        return;
    }
    const long long total = static_cast<long long>(n_images) * go.H * go.W * c8n;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c8 = static_cast<int>(i % c8n);
        long long r = i / c8n;
        const int ox = static_cast<int>(r % go.W);
        r /= go.W;
        const int oy = static_cast<int>(r % go.H);
        const int n = static_cast<int>(r / go.H);
        uint4 m = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx;  // >= -1: inside the zero border (pad >= 1)
                if (iy < gi.H + gi.pad && ix < gi.W + gi.pad)
                    m = hmax8(m, __ldg(reinterpret_cast<const uint4*>(in + plane_row(gi, n, iy, ix) * C) + c8));
            }
        reinterpret_cast<uint4*>(out + plane_row(go, n, oy, ox) * C)[c8] = m;
    }
}

// ------------------------------------------------------------------------------------------------ stride-2 gather
// out(oy, ox) = in(2 oy, 2 ox): the input side of every stride-2 1x1 convolution (STRIDE_IN_1X1) and the output side
// of the stride-2 3x3 convolutions p6/p7 (computed at stride 1).
__global__ void subsample2_kernel(const __half* __restrict__ in, __half* __restrict__ out, PlaneGeom gi, PlaneGeom go,
                                  int n_images, int C) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int c8n = C / 8;
    const long long total = static_cast<long long>(n_images) * go.H * go.W * c8n;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c8 = static_cast<int>(i % c8n);
        long long r = i / c8n;
        const int ox = static_cast<int>(r % go.W);
        r /= go.W;
        const int oy = static_cast<int>(r % go.H);
        const int n = static_cast<int>(r / go.H);
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + plane_row(gi, n, 2 * oy, 2 * ox) * C) + c8);
        reinterpret_cast<uint4*>(out + plane_row(go, n, oy, ox) * C)[c8] = v;
    }
}

// ------------------------------------------------------------------------------------------------ FPN top-down
// lateral(y, x) += coarser(y / 2, x / 2)  (F.interpolate(scale_factor=2, mode="nearest") + add), summed in fp32.
__global__ void upsample_add_kernel(__half* fine, const __half* coarse, PlaneGeom gf, PlaneGeom gc, int n_images, int C) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int c8n = C / 8;
    const long long total = static_cast<long long>(n_images) * gf.H * gf.W * c8n;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c8 = static_cast<int>(i % c8n);
        long long r = i / c8n;
        const int x = static_cast<int>(r % gf.W);
        r /= gf.W;
        const int y = static_cast<int>(r % gf.H);
        const int n = static_cast<int>(r / gf.H);
        uint4* fp = reinterpret_cast<uint4*>(fine + plane_row(gf, n, y, x) * C) + c8;
        const uint4 c = *(reinterpret_cast<const uint4*>(coarse + plane_row(gc, n, y >> 1, x >> 1) * C) + c8);
        const uint4 f = *fp;
        const float2 f0 = unpack_half2(f.x), f1 = unpack_half2(f.y), f2 = unpack_half2(f.z), f3 = unpack_half2(f.w);
        const float2 c0 = unpack_half2(c.x), c1 = unpack_half2(c.y), c2 = unpack_half2(c.z), c3 = unpack_half2(c.w);
        *fp = make_uint4(pack_half2(f0.x + c0.x, f0.y + c0.y), pack_half2(f1.x + c1.x, f1.y + c1.y),
                         pack_half2(f2.x + c2.x, f2.y + c2.y), pack_half2(f3.x + c3.x, f3.y + c3.y));
    }
}

__global__ void relu_copy_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long long n8) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = hmax8(__ldg(in + i), z);
}

// split mode: rows of [C hi | C lo]; relu(hi + lo) keeps both halves where the value is positive
__global__ void relu_copy_split_kernel(const __half* __restrict__ in, __half* __restrict__ out, long long rows, int C) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int c8n = C / 8;
    const long long total = rows * c8n;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c8 = static_cast<int>(i % c8n);
        const long long r = i / c8n;
        float v[8];
        load8f(in + r * 2 * C, c8 * 8, C, v);
        store8f(out + r * 2 * C, c8 * 8, C, v, true);
    }
}

// ------------------------------------------------------------------------------------------------ GroupNorm(32, 256)
// Finalize: per (plane, group) reduce the per-tile partial sums the conv epilogue wrote, in double, in a fixed
// order (deterministic).  stats[(seg * 32 + g) * 2] = mean, [+1] = rstd.   eps = 1e-5 (nn.GroupNorm default).
__global__ void __launch_bounds__(256)
gn_finalize_kernel(const float* __restrict__ partial, const Seg* __restrict__ segs, int seg_begin, int n_segs,
                   float* __restrict__ stats) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    // one block per plane: 32 groups x 8 lanes; each lane strides over the plane's tiles, then an 8-lane shuffle
    // tree in a fixed order (deterministic), all in double
    const int s = seg_begin + blockIdx.x;
    const int g = threadIdx.x >> 3, l = threadIdx.x & 7;
    const Seg sg = segs[s];
    const int t0 = sg.row0 / kBlockM, t1 = (sg.row0 + sg.nrows + kBlockM - 1) / kBlockM;
    double sum = 0.0, sq = 0.0;
    for (int t = t0 + l; t < t1; t += 8) {
        const float2 v = *reinterpret_cast<const float2*>(partial + static_cast<size_t>(t) * 64 + g * 2);
        sum += v.x;
        sq += v.y;
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (l == 0) {
        const double cnt = static_cast<double>(sg.H) * sg.W * 8.0;
        const double mean = sum / cnt;
        double var = sq / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        stats[(static_cast<size_t>(s) * 32 + g) * 2] = static_cast<float>(mean);
        stats[(static_cast<size_t>(s) * 32 + g) * 2 + 1] = static_cast<float>(1.0 / sqrt(var + 1e-5));
    }
}

// Apply: y = relu((x - mean) * rstd * gamma + beta) on interior pixels, 0 elsewhere; reads the fp32 conv output,
// writes the fp16 activation the next convolution consumes.  One CTA per 128-row tile at a time: the plane descriptor
// and the (plane, group) statistics are looked up once per tile (they were a three-deep dependent load chain in front
// of every row), lane = group of 8 channels, warp w takes rows w, w + 8, ... with four rows of loads in flight.
__global__ void __launch_bounds__(256)
gn_apply_relu_kernel(const float* __restrict__ x, __half* __restrict__ y, const float* __restrict__ stats,
                     const float* __restrict__ gamma, const float* __restrict__ beta, const int* __restrict__ tile_seg,
                     const Seg* __restrict__ segs, int tile_begin, int n_tiles, int relu, int split) {
    ptx::griddep_launch();
    const int g = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * g), gb = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * g + 1);
    const float4 ba = __ldg(reinterpret_cast<const float4*>(beta) + 2 * g), bb = __ldg(reinterpret_cast<const float4*>(beta) + 2 * g + 1);
    ptx::griddep_wait();
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int tile = tile_begin + t;
        const int s = __ldg(tile_seg + tile);
        const Seg sg = segs[s];
        const float2 st = *reinterpret_cast<const float2*>(stats + (static_cast<size_t>(s) * 32 + g) * 2);
        const float mean = st.x, rstd = st.y;
#pragma unroll 4
        for (int r = w; r < kBlockM; r += 8) {
            const int row = tile * kBlockM + r;
            uint4 o = make_uint4(0u, 0u, 0u, 0u), ol = make_uint4(0u, 0u, 0u, 0u);
            if (row_is_interior(sg, row)) {
                const float4* p = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * 256) + 2 * g;
                const float4 a = __ldg(p), b = __ldg(p + 1);
                float v[8] = {(a.x - mean) * rstd * ga.x + ba.x, (a.y - mean) * rstd * ga.y + ba.y,
                              (a.z - mean) * rstd * ga.z + ba.z, (a.w - mean) * rstd * ga.w + ba.w,
                              (b.x - mean) * rstd * gb.x + bb.x, (b.y - mean) * rstd * gb.y + bb.y,
                              (b.z - mean) * rstd * gb.z + bb.z, (b.w - mean) * rstd * gb.w + bb.w};
                if (split) split8(v, relu != 0, true, o, ol);
                else o = pack8(v, relu != 0, true);
            }
            if (split) {   // rows of [256 hi | 256 lo]
                reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * 512)[g] = o;
                reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * 512 + 256)[g] = ol;
            } else {
                reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * 256)[g] = o;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ NCHW <-> planes
// (n, C, H, W) fp32 <-> plane rows; used by the plugin-level import and by tests/exports, not by the episode path.
// `lo` > 0 (fp16 planes in split mode): the value is plane[.. + c] + plane[.. + lo + c].
template <typename T>
__global__ void export_nchw_kernel(const T* __restrict__ plane, float* __restrict__ out, PlaneGeom g, int n_images,
                                   int C, int cstride, int coff, float scale, int relu, int lo) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const long long total = static_cast<long long>(n_images) * C * g.H * g.W;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % g.W);
        long long r = i / g.W;
        const int y = static_cast<int>(r % g.H);
        r /= g.H;
        const int c = static_cast<int>(r % C);
        const int n = static_cast<int>(r / C);
        const size_t at = plane_row(g, n, y, x) * cstride + coff + c;
        float v = static_cast<float>(plane[at]);
        if (lo) v += static_cast<float>(plane[at + lo]);
        v *= scale;
        if (relu) v = fmaxf(v, 0.f);
        out[i] = v;
    }
}

__global__ void import_nchw_kernel(const float* __restrict__ in, __half* __restrict__ plane, PlaneGeom g, int n_images,
                                   int C, int lo) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const long long total = static_cast<long long>(n_images) * C * g.H * g.W;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        long long r = i / C;
        const int x = static_cast<int>(r % g.W);
        r /= g.W;
        const int y = static_cast<int>(r % g.H);
        const int n = static_cast<int>(r / g.H);
        const float v = __ldg(in + ((static_cast<size_t>(n) * C + c) * g.H + y) * g.W + x);
        const __half h = __float2half_rn(fminf(fmaxf(v, -kHalfMax), kHalfMax));
        if (lo) {   // rows of [C hi | C lo]
            plane[plane_row(g, n, y, x) * 2 * C + c] = h;
            plane[plane_row(g, n, y, x) * 2 * C + lo + c] = __float2half_rn(v - __half2float(h));
        } else {
            plane[plane_row(g, n, y, x) * C + c] = h;
        }
    }
}

}  // namespace sylph
