// HBM-bound helper kernels around the tensor-core convolutions: image normalise/pad/space-to-depth, max-pool,
// stride-2 gathers, FPN top-down add, GroupNorm finalize/apply, NCHW import/export.  All operate on the flat
// zero-bordered NHWC planes described in conv_gemm.cuh; channel vectors are moved as float4 (coalesced 16-byte lanes).
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

// ------------------------------------------------------------------------------------------------ image prep
// Replaces (x - pixel_mean) / pixel_std + ImageList.from_tensors zero padding
// (sylph/modeling/meta_arch/meta_one_stage_detector.py:174-178), fused with the 2x2 space-to-depth re-layout that
// turns the 7x7/2 stem convolution into a 4x4/1 convolution over 12 (padded to 16) channels.
struct ImageDesc {
    const void* ptr;  // (3, h, w) fp32 or uint8
    int h, w;
    int is_u8;
};

__global__ void prep_stem_input_kernel(const ImageDesc* __restrict__ imgs, float* __restrict__ out, PlaneGeom g,
                                       int n_images, float m0, float m1, float m2, float s0, float s1, float s2) {
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
                    v[(dy * 2 + dx) * 3 + c] = ptx::round_tf32(t);
                }
            }
        v[12] = v[13] = v[14] = v[15] = 0.f;
        float4* o = reinterpret_cast<float4*>(out + plane_row(g, n, y2, x2) * 16);
        o[0] = make_float4(v[0], v[1], v[2], v[3]);
        o[1] = make_float4(v[4], v[5], v[6], v[7]);
        o[2] = make_float4(v[8], v[9], v[10], v[11]);
        o[3] = make_float4(v[12], v[13], v[14], v[15]);
    }
}

// ------------------------------------------------------------------------------------------------ max-pool 3x3 / 2
// detectron2 BasicStem max_pool2d(kernel 3, stride 2, padding 1) on post-ReLU (>= 0) activations: the zero border of
// the input plane stands in for the -inf padding.
__global__ void maxpool3x3s2_kernel(const float* __restrict__ in, float* __restrict__ out, PlaneGeom gi, PlaneGeom go,
                                    int n_images, int C) {
    const int c4n = C / 4;
    const long long total = static_cast<long long>(n_images) * go.H * go.W * c4n;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c4 = static_cast<int>(i % c4n);
        long long r = i / c4n;
        const int ox = static_cast<int>(r % go.W);
        r /= go.W;
        const int oy = static_cast<int>(r % go.H);
        const int n = static_cast<int>(r / go.H);
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx;  // >= -1: inside the zero border (pad >= 1)
                if (iy < gi.H + gi.pad && ix < gi.W + gi.pad) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(in + plane_row(gi, n, iy, ix) * C) + c4);
                    m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
                }
            }
        reinterpret_cast<float4*>(out + plane_row(go, n, oy, ox) * C)[c4] = m;
    }
}

// ------------------------------------------------------------------------------------------------ stride-2 gather
// out(oy, ox) = in(2 oy, 2 ox): the input side of every stride-2 1x1 convolution (STRIDE_IN_1X1) and the output side
// of the stride-2 3x3 convolutions p6/p7 (computed at stride 1).
__global__ void subsample2_kernel(const float* __restrict__ in, float* __restrict__ out, PlaneGeom gi, PlaneGeom go,
                                  int n_images, int C) {
    const int c4n = C / 4;
    const long long total = static_cast<long long>(n_images) * go.H * go.W * c4n;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c4 = static_cast<int>(i % c4n);
        long long r = i / c4n;
        const int ox = static_cast<int>(r % go.W);
        r /= go.W;
        const int oy = static_cast<int>(r % go.H);
        const int n = static_cast<int>(r / go.H);
        const float4 v = __ldg(reinterpret_cast<const float4*>(in + plane_row(gi, n, 2 * oy, 2 * ox) * C) + c4);
        reinterpret_cast<float4*>(out + plane_row(go, n, oy, ox) * C)[c4] = v;
    }
}

// ------------------------------------------------------------------------------------------------ FPN top-down
// lateral(y, x) += coarser(y / 2, x / 2)  (F.interpolate(scale_factor=2, mode="nearest") + add), rounded to TF32
// because the sum feeds the 3x3 output convolution.
__global__ void upsample_add_kernel(float* __restrict__ fine, const float* __restrict__ coarse, PlaneGeom gf,
                                    PlaneGeom gc, int n_images, int C) {
    const int c4n = C / 4;
    const long long total = static_cast<long long>(n_images) * gf.H * gf.W * c4n;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c4 = static_cast<int>(i % c4n);
        long long r = i / c4n;
        const int x = static_cast<int>(r % gf.W);
        r /= gf.W;
        const int y = static_cast<int>(r % gf.H);
        const int n = static_cast<int>(r / gf.H);
        float4* fp = reinterpret_cast<float4*>(fine + plane_row(gf, n, y, x) * C) + c4;
        const float4 c = __ldg(reinterpret_cast<const float4*>(coarse + plane_row(gc, n, y >> 1, x >> 1) * C) + c4);
        float4 f = *fp;
        f.x = ptx::round_tf32(f.x + c.x); f.y = ptx::round_tf32(f.y + c.y);
        f.z = ptx::round_tf32(f.z + c.z); f.w = ptx::round_tf32(f.w + c.w);
        *fp = f;
    }
}

__global__ void relu_copy_kernel(const float4* __restrict__ in, float4* __restrict__ out, long long n4) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float4 v = __ldg(in + i);
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        out[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------ GroupNorm(32, 256)
// Finalize: per (plane, group) reduce the per-tile partial sums the conv epilogue wrote, in double, in a fixed
// order (deterministic).  stats[(seg * 32 + g) * 2] = mean, [+1] = rstd.   eps = 1e-5 (nn.GroupNorm default).
__global__ void gn_finalize_kernel(const float* __restrict__ partial, const Seg* __restrict__ segs, int seg_begin,
                                   int n_segs, float* __restrict__ stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_segs * 32) return;
    const int s = seg_begin + i / 32, g = i % 32;
    const Seg sg = segs[s];
    const int t0 = sg.row0 / kBlockM, t1 = (sg.row0 + sg.nrows + kBlockM - 1) / kBlockM;
    double sum = 0.0, sq = 0.0;
    for (int t = t0; t < t1; ++t) {
        sum += partial[static_cast<size_t>(t) * 64 + g * 2];
        sq += partial[static_cast<size_t>(t) * 64 + g * 2 + 1];
    }
    const double cnt = static_cast<double>(sg.H) * sg.W * 8.0;
    const double mean = sum / cnt;
    double var = sq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[(static_cast<size_t>(s) * 32 + g) * 2] = static_cast<float>(mean);
    stats[(static_cast<size_t>(s) * 32 + g) * 2 + 1] = static_cast<float>(1.0 / sqrt(var + 1e-5));
}

// Apply: y = relu((x - mean) * rstd * gamma + beta) on interior pixels, 0 elsewhere, rounded to TF32; in place.
// One thread per (row, 4 channels); C = 256 -> 64 threads per row, coalesced 1 KiB rows.
__global__ void gn_apply_relu_kernel(float* __restrict__ x, const float* __restrict__ stats,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const int* __restrict__ tile_seg, const Seg* __restrict__ segs, int row_begin,
                                     long long n_rows, int relu) {
    const long long total = n_rows * 64;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c4 = static_cast<int>(i & 63);
        const long long row = row_begin + (i >> 6);
        const int s = tile_seg[row / kBlockM];
        const Seg sg = segs[s];
        const int local = static_cast<int>(row - sg.row0);
        const int y = local / sg.Wp, xx = local - y * sg.Wp;
        const bool interior = local < sg.nrows && y >= sg.pad && y < sg.pad + sg.H && xx >= sg.pad && xx < sg.pad + sg.W;
        float4* p = reinterpret_cast<float4*>(x + row * 256) + c4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (interior) {
            v = *p;
            const int g = c4 >> 1;  // 8 channels per group
            const float mean = stats[(static_cast<size_t>(s) * 32 + g) * 2];
            const float rstd = stats[(static_cast<size_t>(s) * 32 + g) * 2 + 1];
            const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
            const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
            v.x = (v.x - mean) * rstd * ga.x + be.x;
            v.y = (v.y - mean) * rstd * ga.y + be.y;
            v.z = (v.z - mean) * rstd * ga.z + be.z;
            v.w = (v.w - mean) * rstd * ga.w + be.w;
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            v.x = ptx::round_tf32(v.x); v.y = ptx::round_tf32(v.y); v.z = ptx::round_tf32(v.z); v.w = ptx::round_tf32(v.w);
        }
        *p = v;
    }
}

// ------------------------------------------------------------------------------------------------ NCHW <-> planes
// (n, C, H, W) fp32 <-> plane rows; used by the plugin-level import and by tests/exports, not by the episode path.
// `cstride` = channels per plane row, `coff` = first channel, `scale_relu`: export-time transform for bbox_reg.
__global__ void export_nchw_kernel(const float* __restrict__ plane, float* __restrict__ out, PlaneGeom g, int n_images,
                                   int C, int cstride, int coff, float scale, int relu) {
    const long long total = static_cast<long long>(n_images) * C * g.H * g.W;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % g.W);
        long long r = i / g.W;
        const int y = static_cast<int>(r % g.H);
        r /= g.H;
        const int c = static_cast<int>(r % C);
        const int n = static_cast<int>(r / C);
        float v = plane[plane_row(g, n, y, x) * cstride + coff + c] * scale;
        if (relu) v = fmaxf(v, 0.f);
        out[i] = v;
    }
}

__global__ void import_nchw_kernel(const float* __restrict__ in, float* __restrict__ plane, PlaneGeom g, int n_images,
                                   int C, int round) {
    const long long total = static_cast<long long>(n_images) * C * g.H * g.W;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        long long r = i / C;
        const int x = static_cast<int>(r % g.W);
        r /= g.W;
        const int y = static_cast<int>(r % g.H);
        const int n = static_cast<int>(r / g.H);
        float v = __ldg(in + ((static_cast<size_t>(n) * C + c) * g.H + y) * g.W + x);
        if (round) v = ptx::round_tf32(v);
        plane[plane_row(g, n, y, x) * C + c] = v;
    }
}

}  // namespace sylph
