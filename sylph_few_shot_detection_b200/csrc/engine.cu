// libsylph_b200: host-side engine + C ABI (include/sylph_b200.h) of the B200-native Meta-FCOS inference path.
// Owns the prepared weights, the flat-plane activation buffers and the launch sequence; all arithmetic runs in the
// CUDA kernels of this directory.  There is no CPU fallback: without a CUDA device every entry point fails.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/sylph_b200.h"
#include "conv_gemm_host.cuh"
#include "kernels_detect.cuh"
#include "kernels_roi_encoder.cuh"
#include "kernels_loss.cuh"
#include "kernels_backward.cuh"
#include "wgrad3x3.cuh"

namespace sylph {

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }


struct HostTensor {
    std::vector<int64_t> shape;
    std::vector<float> data;
};

// A convolution prepared for conv_gemm: fp16 weights [taps][cout_pad][k_per_tap], fp32 bias [cout_pad].
// Fast mode: k_per_tap = cin, one fp16 value per weight.  Split-operand (exact) mode: k_per_tap = 3 * cin, every row is
// [w_hi | w_hi | w_lo] with w_hi = rn16(w), w_lo = rn16(w - w_hi) (conv_gemm.cuh, GemmArgs::a_wrap).
struct ConvW {
    __half* w = nullptr;
    __half* w_nm = nullptr;   // split mode, 3x3 layers with N <= 128: [taps][w_hi rows ; w_lo rows][cin] for the N-merged kernel
    float* bias = nullptr;
    int taps = 0, k_per_tap = 0, cin = 0, cout = 0, cout_pad = 0, bn = 0, ksize = 0;
    int b_rows = 0;   // rows of the weight matrix (taps * cout_pad; the split-mode stem has 2 x 4 tiles)
};

struct PlaneSet {
    std::vector<Seg> segs;
    int total_rows = 0;
    Seg* d_segs = nullptr;
    int* d_tile_seg = nullptr;
};

struct Buffer {
    void* p = nullptr;
    size_t cap = 0;
    std::string sig;
};

struct Timing {
    std::string name;
    cudaEvent_t e0, e1;
    double flops, bytes;
};

struct Slot {
    bool valid = false;
    int n = 0, hpad = 0, wpad = 0;
    int lh[5] = {0}, lw[5] = {0};
    std::vector<int> img_h, img_w;
    PyramidGeom pg;
    std::shared_ptr<PlaneSet> ps;  // level-major pyramid plane set
    long long level_row0[6] = {0};
    __half* pyr = nullptr;
};

}  // namespace sylph

using namespace sylph;

struct sylph_ctx {
    int device = 0;
    int num_sms = 0;
    sylph_model_config cfg;
    std::string err;
    std::map<std::string, HostTensor> staged;
    bool finalized = false;
    int64_t launches = 0;
    bool profiling = false;
    int staged_epilogue = 1;  // SYLPH_STAGED_EPILOGUE=0 falls back to the register epilogue for conv3
    int halo_pipeline = 1;    // SYLPH_HALO=0 falls back to one A box per tap for the 3x3 convolutions
    int pair_kernel = 1;      // SYLPH_PAIR: bit 0 = CTA-pair kernel for the N = 256 3x3 convolutions, bit 1 = also for N = 64 /
                              // 128 (measured SLOWER than the single-CTA halo kernel: profiles/r01_mma_issue_experiments.md)
    int fuse_upsample = 1;    // SYLPH_FUSE_UPSAMPLE=0: separate upsample_add kernels after the FPN lateral convolutions
    int pair1x1 = 1;          // SYLPH_PAIR1X1=0 keeps the single-CTA staged kernel for every 1x1 convolution; 2 = pair kernel
                              // for every staged 1x1 convolution with 256-channel N tiles and K >= 256 (experiments)
    int stem16 = 1;           // SYLPH_STEM16=0 runs the stem over 64-wide overlapped rows instead of K = 16 taps
    int roi_stride = 81;      // rows per ROI plane: 81 = packed 9x9 planes (SYLPH_ROI_PACKED=0: one 128-row tile per ROI, 37 % more conv tiles)
    int linear_gemm_min = 512; // SYLPH_LINEAR_GEMM_MIN: dense layers of the ROIEncoder run as tensor-core GEMMs from this many rows on
    int cls_pooled = 1;       // SYLPH_CLS_POOLED=0: per-pixel cls convolution over the ROI planes, pooled afterwards (round-1 form)
    int sync_each = 0;        // SYLPH_SYNC_EACH=1: synchronise after every convolution launch and name the one that faults
    int quad = 1;             // SYLPH_QS=0: K' = 3C loop for every split 1x1 layer instead of quad stages on the deep ones
    int nmerge = 1;           // SYLPH_NM=0: three instructions per k-step for the narrow split 3x3 layers instead of the N-merged two
    int roi_separable = 1;    // SYLPH_ROI_ALIGN=sample: the per-sample ROIAlign kernel instead of the separable one
    int split = 1;            // precision mode (sylph_set_precision / SYLPH_PRECISION): 1 = "exact", split fp16 operands
                              // (hi + lo pairs, three tensor-core products per multiply: fp32-level results); 0 = "fast",
                              // single fp16 operands (10-bit mantissa, 1-2.5e-3 max-norm error on deep activations)
    int ld(int channels) const { return split ? 2 * channels : channels; }   // row pitch of an activation tensor
    int lo(int channels) const { return split ? channels : 0; }             // offset of the lo half inside a row
    std::vector<Timing> timings;

    // prepared weights
    ConvW stem;
    struct Block { ConvW c1, c2, c3, sc; bool has_sc = false; };
    std::vector<std::vector<Block>> stages;
    ConvW lat[3], outc[3], p6, p7;
    std::vector<ConvW> cls_tower, box_tower, cg_tower;
    std::vector<float*> cls_gn_w, cls_gn_b, box_gn_w, box_gn_b, cg_gn_w, cg_gn_b;
    ConvW pred, cg_cls, cg_cls_pooled;
    float level_scale[5] = {1, 1, 1, 1, 1};
    float* cg_wbias = nullptr;  // [9][256]
    float* cg_bbias = nullptr;  // [1]
    float* cg_wweight = nullptr;  // [9][256]: per-shot weight head (CODE_GENERATOR.WEIGHT_LAYER), tap-major like cg_wbias
    float* cg_bweight = nullptr;  // [1]
    float* post_gn_w = nullptr;
    float* post_gn_b = nullptr;
    float conv_scale = 1.f, bias_scale = 1.f, bias_value = 0.f;
    // ROIEncoder generator (cfg.generator == 1)
    struct Dense { float* w = nullptr; float* b = nullptr; int in = 0, out = 0; ConvW gemm; bool has_gemm = false; };   // gemm: the same layer for the tensor cores (many rows)
    struct EncLayer { Dense attn, ff1, ff2; float *n1w = nullptr, *n1b = nullptr, *n2w = nullptr, *n2b = nullptr; };
    ConvW re_pool_conv, re_fc1;
    float *re_pool_gn_w = nullptr, *re_pool_gn_b = nullptr;
    std::vector<ConvW> re_tok_conv;
    std::vector<float*> re_tok_gn_w, re_tok_gn_b;
    MsCamWeights re_cam{};
    std::vector<Dense> re_tok_fc;      // tokenizer fc2.. (fc1 is the tensor-core GEMM re_fc1)
    std::vector<EncLayer> re_enc;
    std::vector<Dense> re_whead, re_bhead;
    float cond_scale = 1.f;            // fcos_head.cond_cls_logits.scales.0.scale (CondConvBlock, head_utils.py:121-162)

    // class-code exchange over NVLink peer memory (sylph_exchange_*): one allocation per rank, mapped by all ranks
    struct Exchange {
        int world = 0, rank = 0, max_classes = 0;
        void* local = nullptr;                       // ExchangeState header + [2][max_classes][257] floats
        void* peer_base[sylph::kMaxExchangePeers] = {};
        sylph::ExchangePeers peers{};
        bool connected = false;
        unsigned long long timeout_ns = 5000000000ull;   // SYLPH_EXCHANGE_TIMEOUT_MS
        unsigned int* host_err = nullptr;            // pinned copy of ExchangeState::error, refreshed after every collect
    } xch;
    std::map<std::string, Buffer> bufs;
    // pinned host ring for small host->device argument arrays: copies from it are truly asynchronous, so no entry
    // point has to drain the stream (a pageable cudaMemcpyAsync synchronises the stream first)
    uint8_t* pinned = nullptr;
    size_t pinned_cap = 0, pinned_head = 0;
    unsigned int* detect_overflow_host = nullptr;   // pinned copy of the candidate-overflow word of the last detect calls (sticky)
    uint8_t* graph_arena = nullptr;  // pinned argument blocks referenced by captured CUDA graphs (bump-allocated)
    size_t graph_arena_cap = 0, graph_arena_head = 0;
    std::map<std::string, std::shared_ptr<PlaneSet>> plane_sets;
    Slot slots[SYLPH_NUM_SLOTS];
    // state of the last generate_codes / detect call (for exports)
    int last_n_rois = 0;
    int last_detect_slot = -1, last_detect_classes = 0, last_logit_stride = 0;
    const __half* last_cls_tower = nullptr;   // output planes of the class tower of the last head pass (backward of the cls loss)
    bool weights_ready = false;               // sylph_finalize_weights has succeeded once (sylph_update_code_generator needs it)
    // training of the FCOS class tower (sylph_set_training): the head pass keeps every layer's input planes, pre-GroupNorm
    // convolution output and GroupNorm statistics for sylph_cls_tower_backward
    bool train_save = false;
    bool loss_box_branch = true;              // sylph_set_loss_box_branch(0): the training forward skips the box tower and its predictors
    std::vector<const __half*> saved_x;       // input planes of class-tower layer i
    std::vector<const float*> saved_raw, saved_stats;
    std::vector<ConvW> cls_tower_t;           // transposed, tap-reversed class-tower weights (input-gradient convolutions)
    bool cls_tower_t_ready = false;
    bool inplace_uploads = false;             // sylph_update_code_generator: re-upload into the existing device buffers

    int fail(const char* fmt, ...) {
        char b[1024];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(b, sizeof(b), fmt, ap);
        va_end(ap);
        err = b;
        return 1;
    }
};

#define CU_TRY(ctx, expr)                                                                            \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) return (ctx)->fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define TRY(expr)              \
    do {                       \
        int r__ = (expr);      \
        if (r__ != 0) return r__; \
    } while (0)

namespace sylph {

// ------------------------------------------------------------------------------------------------ memory helpers
static int ensure(sylph_ctx* c, const std::string& name, size_t bytes, const std::string& sig, void** out,
                  cudaStream_t st, bool zero_on_change) {
    Buffer& b = c->bufs[name];
    bool fresh = false;
    if (b.cap < bytes) {
        if (b.p) {
            CU_TRY(c, cudaDeviceSynchronize());
            CU_TRY(c, cudaFree(b.p));
        }
        size_t cap = bytes + (bytes >> 3) + 65536;
        CU_TRY(c, cudaMalloc(&b.p, cap));
        b.cap = cap;
        fresh = true;
    }
    if (zero_on_change && (fresh || b.sig != sig)) CU_TRY(c, cudaMemsetAsync(b.p, 0, b.cap, st));
    b.sig = sig;
    *out = b.p;
    return 0;
}

// Copy `bytes` of host data to device memory through the pinned ring, stream-ordered, without blocking the host.
// While the stream is being CAPTURED into a CUDA graph the copy node would re-read the ring slot at every replay, so
// captured copies take a block of a separate pinned arena that is never reused (released with the context).
static int stage_h2d(sylph_ctx* c, void* dst_dev, const void* src_host, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return 0;
    const size_t need = (bytes + 255) & ~size_t(255);
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (c->graph_arena == nullptr) {   // allocated outside any capture (allocation calls are illegal while capturing)
        c->graph_arena_cap = size_t(2) << 20;
        CU_TRY(c, cudaMallocHost(reinterpret_cast<void**>(&c->graph_arena), c->graph_arena_cap));
    }
    if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusActive) {
        if (c->graph_arena_head + need > c->graph_arena_cap)
            return c->fail("pinned argument arena for captured CUDA graphs is exhausted (%zu bytes)", c->graph_arena_cap);
        uint8_t* block = c->graph_arena + c->graph_arena_head;   // never reused: the graph re-reads it at every replay
        c->graph_arena_head += need;
        memcpy(block, src_host, bytes);
        CU_TRY(c, cudaMemcpyAsync(dst_dev, block, bytes, cudaMemcpyHostToDevice, st));
        return 0;
    }
    if (c->pinned == nullptr || need > c->pinned_cap) {
        if (c->pinned) { CU_TRY(c, cudaDeviceSynchronize()); CU_TRY(c, cudaFreeHost(c->pinned)); }
        c->pinned_cap = std::max<size_t>(need * 2, size_t(4) << 20);
        CU_TRY(c, cudaMallocHost(reinterpret_cast<void**>(&c->pinned), c->pinned_cap));
        c->pinned_head = 0;
    }
    if (c->pinned_head + need > c->pinned_cap) {
        CU_TRY(c, cudaDeviceSynchronize());  // wrap-around: earlier copies out of the ring (on ANY stream) must have executed
        c->pinned_head = 0;
    }
    uint8_t* slot = c->pinned + c->pinned_head;
    c->pinned_head += need;
    memcpy(slot, src_host, bytes);
    CU_TRY(c, cudaMemcpyAsync(dst_dev, slot, bytes, cudaMemcpyHostToDevice, st));
    return 0;
}

static int upload(sylph_ctx* c, const std::vector<float>& h, float** d) {
    if (!(c->inplace_uploads && *d != nullptr)) CU_TRY(c, cudaMalloc(d, std::max<size_t>(h.size(), 1) * sizeof(float)));
    CU_TRY(c, cudaMemcpy(*d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

static int upload_half(sylph_ctx* c, const std::vector<uint16_t>& h, __half** d) {
    if (!(c->inplace_uploads && *d != nullptr)) CU_TRY(c, cudaMalloc(d, std::max<size_t>(h.size(), 1) * sizeof(uint16_t)));
    CU_TRY(c, cudaMemcpy(*d, h.data(), h.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    return 0;
}

static int make_plane_set(sylph_ctx* c, const std::string& key, const std::vector<Seg>& segs, int total_rows,
                          std::shared_ptr<PlaneSet>* out) {
    auto it = c->plane_sets.find(key);
    if (it != c->plane_sets.end()) {
        *out = it->second;
        return 0;
    }
    auto ps = std::make_shared<PlaneSet>();
    ps->segs = segs;
    ps->total_rows = total_rows;
    std::vector<int> tile_seg(total_rows / kBlockM, 0);
    for (size_t s = 0; s < segs.size(); ++s) {
        const int t0 = segs[s].row0 / kBlockM, t1 = (segs[s].row0 + segs[s].nrows + kBlockM - 1) / kBlockM;
        for (int t = t0; t < t1; ++t) tile_seg[t] = static_cast<int>(s);
    }
    CU_TRY(c, cudaMalloc(&ps->d_segs, segs.size() * sizeof(Seg)));
    CU_TRY(c, cudaMalloc(&ps->d_tile_seg, tile_seg.size() * sizeof(int)));
    CU_TRY(c, cudaMemcpy(ps->d_segs, segs.data(), segs.size() * sizeof(Seg), cudaMemcpyHostToDevice));
    CU_TRY(c, cudaMemcpy(ps->d_tile_seg, tile_seg.data(), tile_seg.size() * sizeof(int), cudaMemcpyHostToDevice));
    c->plane_sets[key] = ps;
    *out = ps;
    return 0;
}

// n planes of identical geometry, each starting on a 128-row boundary.
static PlaneGeom regular_geom(int row_base, int H, int W, int pad) {
    PlaneGeom g;
    g.row_base = row_base;
    g.Wp = W + 2 * pad;
    g.pad = pad;
    g.H = H;
    g.W = W;
    g.rows_per_img = round_up((H + 2 * pad) * (W + 2 * pad), kBlockM);
    return g;
}

static std::vector<Seg> geom_segs(const PlaneGeom& g, int n) {
    std::vector<Seg> v;
    for (int i = 0; i < n; ++i) {
        Seg s;
        s.row0 = g.row_base + i * g.rows_per_img;
        s.nrows = (g.H + 2 * g.pad) * g.Wp;
        s.Wp = g.Wp;
        s.pad = g.pad;
        s.H = g.H;
        s.W = g.W;
        v.push_back(s);
    }
    return v;
}

static int grid_for(long long work_items, int threads, int num_sms) {
    long long blocks = (work_items + threads - 1) / threads;
    long long cap = static_cast<long long>(num_sms) * 16;
    return static_cast<int>(std::max<long long>(1, std::min(blocks, cap)));
}

// ------------------------------------------------------------------------------------------------ weight prep
static const HostTensor* find_t(sylph_ctx* c, const std::string& k) {
    auto it = c->staged.find(k);
    return it == c->staged.end() ? nullptr : &it->second;
}

static int pick_bn(int cout) {
    if (cout <= 16) return 16;
    if (cout <= 64) return 64;
    if (cout <= 128) return 128;
    return 256;
}

// OIHW conv (+ optional FrozenBN fold, + optional conv bias) -> ConvW.
static int prep_conv(sylph_ctx* c, const std::string& prefix, bool frozen_bn, bool has_bias, ConvW* out) {
    const HostTensor* w = find_t(c, prefix + ".weight");
    if (!w || w->shape.size() != 4) return c->fail("missing conv weight %s.weight", prefix.c_str());
    const int co = static_cast<int>(w->shape[0]), ci = static_cast<int>(w->shape[1]);
    const int kh = static_cast<int>(w->shape[2]), kw = static_cast<int>(w->shape[3]);
    if (ci % kBlockK != 0) return c->fail("%s: Cin=%d is not a multiple of %d", prefix.c_str(), ci, kBlockK);
    std::vector<float> scale(co, 1.f), shift(co, 0.f);
    if (frozen_bn) {
        const HostTensor *g = find_t(c, prefix + ".norm.weight"), *b = find_t(c, prefix + ".norm.bias"),
                         *m = find_t(c, prefix + ".norm.running_mean"), *v = find_t(c, prefix + ".norm.running_var");
        if (!g || !b || !m || !v) return c->fail("missing FrozenBN tensors for %s", prefix.c_str());
        for (int o = 0; o < co; ++o) {
            // detectron2 FrozenBatchNorm2d: scale = weight * rsqrt(var + eps); bias = bias - mean * scale
            const float s = g->data[o] * (1.0f / std::sqrt(v->data[o] + 1e-5f));
            scale[o] = s;
            shift[o] = b->data[o] - m->data[o] * s;
        }
    }
    if (has_bias) {
        const HostTensor* b = find_t(c, prefix + ".bias");
        if (!b) return c->fail("missing %s.bias", prefix.c_str());
        for (int o = 0; o < co; ++o) shift[o] += b->data[o];
    }
    out->taps = kh * kw;
    out->ksize = kh;
    out->cin = ci;
    out->k_per_tap = c->split ? 3 * ci : ci;
    out->cout = co;
    out->bn = pick_bn(co);
    out->cout_pad = round_up(co, out->bn);
    out->b_rows = out->taps * out->cout_pad;
    const int kp = out->k_per_tap;
    std::vector<uint16_t> hw(static_cast<size_t>(out->taps) * out->cout_pad * kp, 0);
    std::vector<float> hb(out->cout_pad, 0.f);
    for (int o = 0; o < co; ++o) {
        hb[o] = shift[o];
        for (int i = 0; i < ci; ++i)
            for (int t = 0; t < out->taps; ++t) {
                const float v = w->data[(static_cast<size_t>(o) * ci + i) * out->taps + t] * scale[o];
                const uint16_t hi = float_to_half_bits(v);
                uint16_t* row = &hw[(static_cast<size_t>(t) * out->cout_pad + o) * kp];
                row[i] = hi;
                if (c->split) {
                    row[ci + i] = hi;
                    row[2 * ci + i] = float_to_half_bits(v - half_bits_to_float(hi));
                }
            }
    }
    TRY(upload_half(c, hw, &out->w));
    TRY(upload(c, hb, &out->bias));
    if (c->split && out->taps == 9 && out->bn <= 128) {
        // the same weights for the N-merged kernel (conv_gemm.cuh, NM): per tap the cout_pad rows of w_hi, then those of w_lo
        std::vector<uint16_t> nm(static_cast<size_t>(out->taps) * 2 * out->cout_pad * ci, 0);
        for (int t = 0; t < out->taps; ++t)
            for (int o = 0; o < co; ++o) {
                const uint16_t* src = &hw[(static_cast<size_t>(t) * out->cout_pad + o) * kp];
                memcpy(&nm[(static_cast<size_t>(2 * t) * out->cout_pad + o) * ci], src, static_cast<size_t>(ci) * 2);
                memcpy(&nm[(static_cast<size_t>(2 * t + 1) * out->cout_pad + o) * ci], src + 2 * ci, static_cast<size_t>(ci) * 2);
            }
        TRY(upload_half(c, nm, &out->w_nm));
    }
    return 0;
}

// 7x7 / 2 stem over 3 channels -> 4 vertical taps x (4 horizontal taps x 16 space-to-depth channels).
static int prep_stem(sylph_ctx* c, ConvW* out) {
    const std::string prefix = "backbone.bottom_up.stem.conv1";
    const HostTensor* w = find_t(c, prefix + ".weight");
    if (!w || w->shape.size() != 4 || w->shape[1] != 3 || w->shape[2] != 7) return c->fail("bad stem weight");
    const int co = static_cast<int>(w->shape[0]);
    const HostTensor *g = find_t(c, prefix + ".norm.weight"), *b = find_t(c, prefix + ".norm.bias"),
                     *m = find_t(c, prefix + ".norm.running_mean"), *v = find_t(c, prefix + ".norm.running_var");
    if (!g || !b || !m || !v) return c->fail("missing stem FrozenBN");
    out->taps = 4;
    out->ksize = 7;
    out->k_per_tap = 64;
    out->cin = 64;
    out->cout = co;
    out->bn = pick_bn(co);
    out->cout_pad = round_up(co, out->bn);
    // split mode: tiles 0..3 = w_hi of the four vertical taps, tiles 4..7 = w_lo (conv_gemm.cuh, STEM16)
    const int n_tiles = c->split ? 8 : 4;
    out->b_rows = n_tiles * out->cout_pad;
    std::vector<uint16_t> hw(static_cast<size_t>(n_tiles) * out->cout_pad * 64, 0);
    std::vector<float> hb(out->cout_pad, 0.f);
    // Exact mode: the stem reads re-centred integer pixels u = (v - round(mean)) / 256 (kernels_misc.cuh, "image prep, exact
    // mode"), so the weights carry 256 / std of their input channel and the bias the -(mean - round(mean)) / std term of all 49 taps;
    // folded in double, then split into hi + lo.
    const sylph_model_config& f = c->cfg;
    for (int o = 0; o < co; ++o) {
        const float s = g->data[o] * (1.0f / std::sqrt(v->data[o] + 1e-5f));
        hb[o] = b->data[o] - m->data[o] * s;
        if (c->split) {
            double corr = 0.0;
            for (int ch = 0; ch < 3; ++ch) {
                const double delta = static_cast<double>(f.pixel_mean[ch]) - std::nearbyint(static_cast<double>(f.pixel_mean[ch]));
                for (int kk = 0; kk < 49; ++kk)
                    corr += static_cast<double>(w->data[(static_cast<size_t>(o) * 3 + ch) * 49 + kk]) * s / f.pixel_std[ch] * delta;
            }
            hb[o] = static_cast<float>(static_cast<double>(b->data[o]) - static_cast<double>(m->data[o]) * s - corr);
            for (int ty = 0; ty < 4; ++ty)
                for (int tx = 0; tx < 4; ++tx)
                    for (int dy = 0; dy < 2; ++dy)
                        for (int dx = 0; dx < 2; ++dx) {
                            const int ky = 2 * ty + dy - 1, kx = 2 * tx + dx - 1;
                            if (ky < 0 || ky > 6 || kx < 0 || kx > 6) continue;
                            for (int ch = 0; ch < 3; ++ch) {
                                const int k = tx * 16 + (dy * 2 + dx) * 3 + ch;
                                const double wv = static_cast<double>(w->data[((static_cast<size_t>(o) * 3 + ch) * 7 + ky) * 7 + kx]) * s *
                                                  256.0 / f.pixel_std[ch];
                                const uint16_t hi = float_to_half_bits(static_cast<float>(wv));
                                hw[(static_cast<size_t>(ty) * out->cout_pad + o) * 64 + k] = hi;
                                hw[(static_cast<size_t>(4 + ty) * out->cout_pad + o) * 64 + k] =
                                    float_to_half_bits(static_cast<float>(wv - static_cast<double>(half_bits_to_float(hi))));
                            }
                        }
            continue;
        }
        for (int ty = 0; ty < 4; ++ty)
            for (int tx = 0; tx < 4; ++tx)
                for (int dy = 0; dy < 2; ++dy)
                    for (int dx = 0; dx < 2; ++dx) {
                        const int ky = 2 * ty + dy - 1, kx = 2 * tx + dx - 1;
                        if (ky < 0 || ky > 6 || kx < 0 || kx > 6) continue;
                        for (int ch = 0; ch < 3; ++ch) {
                            const int k = tx * 16 + (dy * 2 + dx) * 3 + ch;
                            const float v = w->data[((static_cast<size_t>(o) * 3 + ch) * 7 + ky) * 7 + kx] * s;
                            const uint16_t hi = float_to_half_bits(v);
                            hw[(static_cast<size_t>(ty) * out->cout_pad + o) * 64 + k] = hi;
                            if (c->split)
                                hw[(static_cast<size_t>(4 + ty) * out->cout_pad + o) * 64 + k] =
                                    float_to_half_bits(v - half_bits_to_float(hi));
                        }
                    }
    }
    TRY(upload_half(c, hw, &out->w));
    TRY(upload(c, hb, &out->bias));
    return 0;
}

static int upload_vec(sylph_ctx* c, const std::string& key, float** d, size_t expect) {
    const HostTensor* t = find_t(c, key);
    if (!t || t->data.size() != expect) return c->fail("missing or mis-sized tensor %s", key.c_str());
    return upload(c, t->data, d);
}

static int scalar_of(sylph_ctx* c, const std::string& key, float* v) {
    const HostTensor* t = find_t(c, key);
    if (!t || t->data.size() != 1) return c->fail("missing scalar %s", key.c_str());
    *v = t->data[0];
    return 0;
}

// ------------------------------------------------------------------------------------------------ ROIEncoder weights
static int upload_t(sylph_ctx* c, const std::string& key, float** d, size_t expect, bool transpose = false, int rows = 0,
                    int cols = 0) {
    const HostTensor* t = find_t(c, key);
    if (!t || t->data.size() != expect) return c->fail("missing or mis-sized tensor %s", key.c_str());
    if (!transpose) return upload(c, t->data, d);
    std::vector<float> h(expect);
    for (int r = 0; r < rows; ++r)
        for (int q = 0; q < cols; ++q) h[static_cast<size_t>(q) * rows + r] = t->data[static_cast<size_t>(r) * cols + q];
    return upload(c, h, d);
}

// The same dense layer as a 1x1 "convolution" for the tensor-core GEMM kernels ([rows][in] x [in][out], split operands in exact
// mode): used when the layer runs over thousands of rows (class sweeps); `w` is [out][in] row-major.
static int prep_dense_gemm(sylph_ctx* c, const std::string& tag, const std::vector<float>& w, const std::vector<float>& b,
                           sylph_ctx::Dense* d) {
    if (d->in % kBlockK != 0 || d->out < 16) return 0;
    HostTensor hw, hb;
    hw.shape = {d->out, d->in, 1, 1};
    hw.data = w;
    hb.shape = {d->out};
    hb.data = b;
    c->staged["__dense_" + tag + ".weight"] = std::move(hw);
    c->staged["__dense_" + tag + ".bias"] = std::move(hb);
    TRY(prep_conv(c, "__dense_" + tag, false, true, &d->gemm));
    d->has_gemm = true;
    return 0;
}

static int prep_dense(sylph_ctx* c, const std::string& prefix, int in, int out, sylph_ctx::Dense* d) {
    d->in = in;
    d->out = out;
    TRY(upload_t(c, prefix + ".weight", &d->w, static_cast<size_t>(in) * out));
    TRY(upload_t(c, prefix + ".bias", &d->b, static_cast<size_t>(out)));
    const HostTensor *w = find_t(c, prefix + ".weight"), *b = find_t(c, prefix + ".bias");
    if (w && b) TRY(prep_dense_gemm(c, prefix, w->data, b->data, d));
    return 0;
}

// ROIEncoder (sylph/modeling/code_generator/roi_encoder.py:206-281): FeatureFusionModuleV2 conv + MS_CAM, Tokenizer,
// TransformerEncoder (sequence length 1 at inference: W_o W_v folded), weight / bias HyperNetworkHead.
static int prep_roi_encoder(sylph_ctx* c) {
    const sylph_model_config& f = c->cfg;
    const std::string cg = "code_generator.";
    if (f.re_tok_convs < 1 || f.re_tok_fcs < 1 || f.re_layers < 0 || f.re_head_fcs < 1 || f.re_head_dim > 1024)
        return c->fail("unsupported ROIEncoder dimensions");
    TRY(prep_conv(c, cg + "box_pooler.conv.0", false, true, &c->re_pool_conv));
    TRY(upload_vec(c, cg + "box_pooler.conv.1.weight", &c->re_pool_gn_w, 256));
    TRY(upload_vec(c, cg + "box_pooler.conv.1.bias", &c->re_pool_gn_b, 256));
    const std::string cam = cg + "box_pooler.context_attention_module.";
    MsCamWeights& m = c->re_cam;
    float* p;
    TRY(upload_t(c, cam + "local_att.0.weight", &p, 64 * 256, true, 64, 256)); m.l_w1t = p;
    TRY(upload_t(c, cam + "local_att.0.bias", &p, 64)); m.l_b1 = p;
    TRY(upload_t(c, cam + "local_att.1.weight", &p, 64)); m.l_g1w = p;
    TRY(upload_t(c, cam + "local_att.1.bias", &p, 64)); m.l_g1b = p;
    TRY(upload_t(c, cam + "local_att.3.weight", &p, 256 * 64, true, 256, 64)); m.l_w2t = p;
    TRY(upload_t(c, cam + "local_att.3.bias", &p, 256)); m.l_b2 = p;
    TRY(upload_t(c, cam + "local_att.4.weight", &p, 256)); m.l_g2w = p;
    TRY(upload_t(c, cam + "local_att.4.bias", &p, 256)); m.l_g2b = p;
    TRY(upload_t(c, cam + "global_att.1.weight", &p, 64 * 256)); m.g_w1 = p;
    TRY(upload_t(c, cam + "global_att.1.bias", &p, 64)); m.g_b1 = p;
    TRY(upload_t(c, cam + "global_att.2.weight", &p, 64)); m.g_g1w = p;
    TRY(upload_t(c, cam + "global_att.2.bias", &p, 64)); m.g_g1b = p;
    TRY(upload_t(c, cam + "global_att.4.weight", &p, 256 * 64, true, 256, 64)); m.g_w2t = p;
    TRY(upload_t(c, cam + "global_att.4.bias", &p, 256)); m.g_b2 = p;
    TRY(upload_t(c, cam + "global_att.5.weight", &p, 256)); m.g_g2w = p;
    TRY(upload_t(c, cam + "global_att.5.bias", &p, 256)); m.g_g2b = p;
    // tokenizer convolutions: Conv2d(bias = not norm) + GN + ReLU; the GEMM epilogue adds a zero bias
    c->re_tok_conv.assign(f.re_tok_convs, ConvW());
    c->re_tok_gn_w.assign(f.re_tok_convs, nullptr);
    c->re_tok_gn_b.assign(f.re_tok_convs, nullptr);
    for (int i = 0; i < f.re_tok_convs; ++i) {
        const std::string k = cg + "tokenizer.conv" + std::to_string(i + 1);
        TRY(prep_conv(c, k, false, false, &c->re_tok_conv[i]));
        TRY(upload_vec(c, k + ".norm.weight", &c->re_tok_gn_w[i], 256));
        TRY(upload_vec(c, k + ".norm.bias", &c->re_tok_gn_b[i], 256));
    }
    {   // fc1: (256, 256 * 49) over nn.Flatten order c * 49 + p  ->  1x1 "convolution" over K = p * 256 + c
        const HostTensor *w = find_t(c, cg + "tokenizer.fc1.weight"), *b = find_t(c, cg + "tokenizer.fc1.bias");
        if (!w || !b || w->data.size() != static_cast<size_t>(256) * 12544 || b->data.size() != 256)
            return c->fail("tokenizer.fc1 must map 256 x 7 x 7 to 256 features");
        HostTensor pw;
        pw.shape = {256, 12544, 1, 1};
        pw.data.resize(w->data.size());
        for (int o = 0; o < 256; ++o)
            for (int ch = 0; ch < 256; ++ch)
                for (int px = 0; px < 49; ++px)
                    pw.data[static_cast<size_t>(o) * 12544 + px * 256 + ch] = w->data[static_cast<size_t>(o) * 12544 + ch * 49 + px];
        c->staged["__re_fc1.weight"] = std::move(pw);
        c->staged["__re_fc1.bias"] = *b;
        TRY(prep_conv(c, "__re_fc1", false, true, &c->re_fc1));
    }
    c->re_tok_fc.assign(f.re_tok_fcs - 1, sylph_ctx::Dense());
    for (int i = 1; i < f.re_tok_fcs; ++i) TRY(prep_dense(c, cg + "tokenizer.fc" + std::to_string(i + 1), 256, 256, &c->re_tok_fc[i - 1]));
    c->re_enc.assign(f.re_layers, sylph_ctx::EncLayer());
    for (int l = 0; l < f.re_layers; ++l) {
        const std::string k = cg + "transformer_encoder.layers." + std::to_string(l) + ".";
        const HostTensor *ipw = find_t(c, k + "self_attn.in_proj_weight"), *ipb = find_t(c, k + "self_attn.in_proj_bias"),
                         *ow = find_t(c, k + "self_attn.out_proj.weight"), *ob = find_t(c, k + "self_attn.out_proj.bias");
        if (!ipw || !ipb || !ow || !ob || ipw->data.size() != 768 * 256 || ow->data.size() != 256 * 256)
            return c->fail("transformer layer %d: d_model must be 256", l);
        // sequence length 1: attention output = out_proj(v_proj(x)); fold in double precision
        std::vector<float> wf(256 * 256), bf(256);
        for (int o = 0; o < 256; ++o) {
            double bacc = ob->data[o];
            for (int j = 0; j < 256; ++j) bacc += static_cast<double>(ow->data[o * 256 + j]) * ipb->data[512 + j];
            bf[o] = static_cast<float>(bacc);
            for (int i = 0; i < 256; ++i) {
                double acc = 0.0;
                for (int j = 0; j < 256; ++j) acc += static_cast<double>(ow->data[o * 256 + j]) * ipw->data[(512 + j) * 256 + i];
                wf[o * 256 + i] = static_cast<float>(acc);
            }
        }
        sylph_ctx::EncLayer& L = c->re_enc[l];
        L.attn.in = L.attn.out = 256;
        TRY(upload(c, wf, &L.attn.w));
        TRY(upload(c, bf, &L.attn.b));
        TRY(prep_dense_gemm(c, k + "attn_folded", wf, bf, &L.attn));
        TRY(prep_dense(c, k + "linear1", 256, 1024, &L.ff1));
        TRY(prep_dense(c, k + "linear2", 1024, 256, &L.ff2));
        TRY(upload_vec(c, k + "norm1.weight", &L.n1w, 256));
        TRY(upload_vec(c, k + "norm1.bias", &L.n1b, 256));
        TRY(upload_vec(c, k + "norm2.weight", &L.n2w, 256));
        TRY(upload_vec(c, k + "norm2.bias", &L.n2b, 256));
    }
    auto prep_head = [&](const std::string& name, int out_dim, std::vector<sylph_ctx::Dense>* hd) -> int {
        hd->assign(f.re_head_fcs, sylph_ctx::Dense());
        int din = 256;
        for (int i = 0; i < f.re_head_fcs; ++i) {
            const int dout = (i == f.re_head_fcs - 1) ? out_dim : f.re_head_dim;
            TRY(prep_dense(c, cg + name + ".fc" + std::to_string(i + 1), din, dout, &(*hd)[i]));
            din = dout;
        }
        return 0;
    };
    TRY(prep_head("weight_head", 256, &c->re_whead));
    TRY(prep_head("bias_head", 1, &c->re_bhead));
    TRY(scalar_of(c, "proposal_generator.fcos_head.cond_cls_logits.scales.0.scale", &c->cond_scale));
    return 0;
}

// ------------------------------------------------------------------------------------------------ conv launch
struct ConvCall {
    const ConvW* W;
    const __half* A;
    long long a_rows;   // rows of the A buffer visible to TMA
    int a_cols, a_ld;   // inner dim / pitch (elements)
    const PlaneSet* ps; // plane set of the OUTPUT buffer
    int tile_begin, n_tiles;
    int a_row_delta;
    void* out;
    int ldc;
    int flags;
    const __half* residual = nullptr;
    int ld_res = 0;
    float* gn_partial = nullptr;
    const float* bias_override = nullptr;
    const __half* w_override = nullptr;
    int stem = 0;            // 1 = the stem convolution; 2 = exact mode with uint8 images (a_lo == 0: the a_lo pass is skipped)
    int up_seg_delta = 0;    // kEpiUpsample: coarser plane = segment of the output tile + this
    int staged = 0;          // TMA-in / TMA-out epilogue (fp16 output, BN = 256); needs out_rows
    int pairsplit = 0;       // split mode: CTA-pair 1x1 kernel with the chunked staged epilogue and quad stages (conv1x1_pair_split.cuh)
    int qs = 0;              // split mode: quad-stage 1x1 kernel (conv_gemm.cuh QS) with N tiles of this width (128 / 256), 0 = K' = 3C loop
    long long out_rows = 0;  // rows of the output (and residual) buffer, for the staged epilogue's tensor maps
    const char* name = "conv";
};

static int run_conv_impl(sylph_ctx* c, const ConvCall& k, cudaStream_t st, const bool has_res);

static int run_conv(sylph_ctx* c, const ConvCall& k, cudaStream_t st) {
    const ConvW& W = *k.W;
    // Split-operand mode: a_cols / a_ld / ldc / ld_res are PHYSICAL pitches (2C: [C hi | C lo]); the weight matrix has
    // K' = 3 cin per tap, the A column of a k-block wraps at 2 cin, fp16 outputs and residuals carry a lo half at +C.
    if (c->split && !k.stem && (k.a_cols != 2 * W.cin || W.k_per_tap != 3 * W.cin))
        return c->fail("split-operand convolution %s: A has %d columns, weights K' = %d for cin = %d", k.name, k.a_cols, W.k_per_tap, W.cin);
    const bool has_res = (k.flags & kEpiResidual) != 0;
    // Split mode, deep 1x1 layers without a residual (conv1 of res4 / res5, the res5 shortcut: cin >= 1024, K' >= 3072): bound by
    // L2->SM operand traffic, not by stores -- the 256-wide register-epilogue kernel (4-stage ring, half the A re-reads) beats the
    // staged 128-wide one: 0.304 -> 0.253, 0.320 -> 0.247, 0.540 -> 0.486 ms (profiles/r02_split_tile_width_ab.log)
    const bool deep_direct = c->split && k.staged && !k.stem && W.taps == 1 && !has_res && W.bn == 256 && W.cin >= 1024;
    ConvCall kk = k;
    if (deep_direct) kk.staged = 0;
    // Split mode, 1x1 layers with cin >= 512 (and the res3 / res4 shortcuts): quad stages -- a_hi, a_lo, w_hi, w_lo of a k-block are
    // loaded once (four tiles instead of six per k-block).  Measured at the 33-image shapes (profiles/r02_quad_stage_ab.log):
    // conv1 res3 0.344 -> 0.288, res4 0.256 -> 0.228, res5 0.251 -> 0.223; shortcut res3 0.639 -> 0.568, res4 0.542 -> 0.458,
    // res5 0.495 -> 0.424; laterals 0.572 / 0.257 / 0.158 -> 0.512 / 0.229 / 0.125 ms.  conv3 (staged, residual) stays on the
    // K' = 3C kernel: quad stages leave room for ONE staging buffer only, which costs more than the operand traffic saved
    // (res4 0.404 -> 0.498), except on res5 (K = 512) where the register epilogue wins (0.359 -> 0.325).
    if (c->split && c->quad && !k.stem && W.taps == 1 && k.w_override == nullptr && W.cout_pad % 128 == 0 && W.cin % 64 == 0 &&
        !(k.flags & kEpiGnStats)) {
        const bool upsample = (k.flags & kEpiUpsample) != 0;
        // CTA-pair kernel with the chunked staged epilogue (conv1x1_pair_split.cuh) wherever the output leaves by TMA in
        // 256-channel tiles: one M = 256 instruction per k-step halves every SM's shared-memory operand reads -- at 33 images
        // conv3 res4 0.406 -> 0.299, res5 0.360 -> 0.234; shortcut res3 0.639 -> 0.444, res4 0.542 -> 0.380, res5 0.495 -> 0.383;
        // conv1 res4 0.256 -> 0.210, res5 0.251 -> 0.217 ms (profiles/r02_pair_split_ab.log); res3 conv3 (K = 128) is HBM-bound
        // either way and stays on the single-CTA kernel.
        const bool pair_ok = c->pair1x1 && k.staged && k.out_rows > 0 && W.cout_pad % 256 == 0 && !upsample;
        if (has_res && !upsample) {
            if (pair_ok && W.cin >= 256) { kk.pairsplit = 1; kk.staged = 1; }                   // res4 / res5 conv3
            else if (W.cin >= 512) { kk.qs = 128; kk.staged = 0; }
        } else if (k.staged && !(k.flags & kEpiRelu) && W.cout_pad >= 512) {                    // shortcut convolutions
            if (pair_ok && W.cin >= 256) { kk.pairsplit = 1; kk.staged = 1; }   // (deep_direct above may have cleared `staged`)
            else if (W.cin >= 1024) { kk.qs = 256; kk.staged = 0; }
            else if (W.cin >= 256) { kk.qs = 128; }
        } else if (W.cin >= 512) {                                                              // conv1, FPN laterals
            if (pair_ok) { kk.pairsplit = 1; kk.staged = 1; }
            else {
                kk.qs = (W.cin == 512 && W.cout_pad % 256 == 0 && !k.staged) ? 256 : 128;
                kk.staged = 0;
            }
        }
    }
    return run_conv_impl(c, kk, st, has_res);
}

static int run_conv_impl(sylph_ctx* c, const ConvCall& k, cudaStream_t st, const bool has_res) {
    const ConvW& W = *k.W;
    CUtensorMap ta, tb;
    std::string err;
    const bool split = c->split != 0;
    const bool out_f16 = !(k.flags & kEpiOutF32);
    const bool halo = c->halo_pipeline && W.taps == 9 && !k.stem && !k.staged;
    const bool pair = halo && c->pair_kernel && (W.bn == 256 || ((c->pair_kernel & 2) && W.bn >= 64)) && !(k.flags & kEpiResidual);
    const bool stem16 = k.stem && c->stem16 && k.staged && W.bn == 64 && W.taps == 4 && k.a_ld == (split ? 32 : 16);
    if (k.stem && split && !stem16) return c->fail("the split-operand stem needs SYLPH_STEM16=1 and the staged epilogue");
    // Staged 1x1 convolutions with 256-channel N tiles: the CTA-pair kernel where it measured faster on the 33-image
    // shapes (profiles/r01_pair1x1_shapes.log): shortcut convolutions (no residual, K >= 256, N >= 512: res4 0.205 ->
    // 0.161 ms) and conv3 with K >= 512 (res5 0.126 -> 0.111 ms); res3 / res4 conv3 and the conv1 layers stay single-CTA
    // (lock-stepped epilogues of the pair cost more than the halved weight traffic buys there).
    const bool pair1x1 = !split && k.staged && !k.stem && W.taps == 1 && W.bn == 256 && k.n_tiles >= 2 && W.k_per_tap >= 256 &&
                         (c->pair1x1 == 2 || (c->pair1x1 == 1 && (has_res ? W.k_per_tap >= 512 : W.cout_pad >= 512)));
    // N tile: a staged split tile holds a hi and a lo half, so it is at most 128 channels wide
    // (two 64 KB staging buffers); shortcut convolutions (no residual to prefetch, N >= 512) run 256-wide tiles over ONE 128 KB
    // staging buffer, which halves their A re-reads from L2: 0.72 -> 0.64 / 0.60 -> 0.55 / 0.60 -> 0.54 ms on res3..res5 at 33
    // images; every other staged layer measured slower that way (profiles/r02_split_tile_width_ab.log)
    const bool wide_split = split && k.staged && W.bn == 256 && !has_res && !(k.flags & kEpiRelu) && W.cout_pad >= 512;
    const int bn = k.pairsplit ? 256 : k.qs ? k.qs : (split && k.staged && W.bn == 256 && !wide_split) ? 128 : W.bn;
    if (stem16) {
        if (make_tmap_2d_k16(&ta, k.A, static_cast<uint64_t>(k.a_rows), kBlockM + 3, &err, split ? 32 : 16))
            return c->fail("A tensor map (%s): %s", k.name, err.c_str());
    } else if (make_tmap_2d(&ta, k.A, static_cast<uint64_t>(k.a_rows), k.a_cols, k.a_ld, halo ? kBlockM + 2 : kBlockM, &err))
        return c->fail("A tensor map (%s): %s", k.name, err.c_str());
    // narrow split 3x3 layers (res2 / res3 conv2, the predictor): N-merged kernel over the [taps][w_hi ; w_lo][cin] weights --
    // res2 conv2 0.60 -> 0.47 ms, res3 conv2 0.47 -> 0.42 ms, predictor 0.26 -> 0.18 ms (test_conv_gemm split)
    const bool nm = split && halo && !pair && c->nmerge && W.w_nm != nullptr && k.w_override == nullptr && W.cout_pad == bn;
    if (nm) {
        if (make_tmap_2d(&tb, W.w_nm, static_cast<uint64_t>(W.taps) * 2 * W.cout_pad, W.cin, W.cin, bn, &err))
            return c->fail("B tensor map (%s): %s", k.name, err.c_str());
    } else if (make_tmap_2d(&tb, k.w_override ? k.w_override : W.w, static_cast<uint64_t>(W.b_rows), W.k_per_tap,
                            W.k_per_tap, (pair || pair1x1 || k.pairsplit) ? bn / 2 : bn, &err))
        return c->fail("B tensor map (%s): %s", k.name, err.c_str());
    GemmArgs g{};
    g.tile_begin = k.tile_begin;
    g.num_m_tiles = k.n_tiles;
    g.num_n_tiles = W.cout_pad / bn;
    g.a_row_delta = k.a_row_delta;
    g.taps = W.taps;
    g.kblocks_per_tap = W.k_per_tap / kBlockK;
    g.b_rows_per_tap = W.cout_pad;
    g.a_wrap = split ? 2 * W.cin : 0;
    g.nm_passes = (k.stem == 2) ? 1 : 2;
    if (stem16 && split) g.nm_lo_row = 4 * W.cout_pad;   // stem weight matrix: the four w_hi tiles, then the four w_lo tiles
    if (k.qs || k.pairsplit) g.kblocks_per_tap = W.cin / kBlockK;   // logical k-blocks: a stage carries a_hi, a_lo, w_hi, w_lo
    if (nm) {
        g.kblocks_per_tap = 2 * W.cin / kBlockK;
        g.b_rows_per_tap = 2 * W.cout_pad;
        g.nm_lo_row = W.cout_pad;
    }
    g.out_lo = (split && out_f16) ? k.ldc / 2 : 0;
    g.res_lo = split ? k.ld_res / 2 : 0;
    for (int t = 0; t < W.taps; ++t) {
        if (k.stem) { g.tap_dy[t] = static_cast<signed char>(t - 2); g.tap_dx[t] = -2; }
        else if (W.taps == 9) { g.tap_dy[t] = static_cast<signed char>(t / 3 - 1); g.tap_dx[t] = static_cast<signed char>(t % 3 - 1); }
        else { g.tap_dy[t] = 0; g.tap_dx[t] = 0; }
    }
    g.bias = k.bias_override ? k.bias_override : W.bias;
    g.residual = k.residual;
    g.ld_res = k.ld_res;
    g.out = k.out;
    g.ldc = k.ldc;
    g.flags = k.flags;
    g.tile_seg = k.ps->d_tile_seg;
    g.segs = k.ps->d_segs;
    g.gn_partial = k.gn_partial;
    g.up_seg_delta = k.up_seg_delta;
    Timing tm;
    if (c->profiling) {
        tm.name = k.name;
        cudaEventCreate(&tm.e0);
        cudaEventCreate(&tm.e1);
        const double rows = static_cast<double>(k.n_tiles) * kBlockM;
        // executed tensor-core work (split mode: three products per multiply) and bytes of the padded planes moved
        tm.flops = 2.0 * rows * W.cout_pad * W.k_per_tap * W.taps;
        tm.bytes = rows * (static_cast<double>(k.a_ld < k.a_cols ? k.a_ld : k.a_cols) * 2.0 +
                           static_cast<double>(out_f16 ? k.ldc * 2.0 : std::min(k.ldc, W.cout_pad) * 4.0) +
                           ((k.flags & kEpiResidual) ? k.ld_res * 2.0 : 0.0));
        cudaEventRecord(tm.e0, st);
    }
    if (k.staged) {
        if (bn < 64 || (k.flags & (kEpiOutF32 | kEpiGnStats)) || k.out_rows <= 0)
            return c->fail("staged epilogue needs BN >= 64, fp16 output and out_rows (%s)", k.name);
        CUtensorMap tres, tout;
        const __half* rsrc = (k.flags & kEpiResidual) ? k.residual : static_cast<const __half*>(k.out);
        if (make_tmap_2d(&tres, rsrc, static_cast<uint64_t>(k.out_rows), (k.flags & kEpiResidual) ? k.ld_res : k.ldc,
                         (k.flags & kEpiResidual) ? k.ld_res : k.ldc, kBlockM, &err) ||
            make_tmap_2d(&tout, static_cast<const __half*>(k.out), static_cast<uint64_t>(k.out_rows), k.ldc, k.ldc, kBlockM, &err))
            return c->fail("epilogue tensor maps (%s): %s", k.name, err.c_str());
        if (stem16) CU_TRY(c, launch_conv_gemm_stem16(ta, tb, tout, g, c->num_sms, st, split, split && c->nmerge));
        else if (pair1x1) CU_TRY(c, launch_conv1x1_pair_staged(ta, tb, tres, tout, g, c->num_sms, st));
        else if (k.pairsplit) CU_TRY(c, launch_conv1x1_pair_split(true, ta, tb, tres, tout, g, c->num_sms, st));
        else if (k.qs) CU_TRY(c, launch_conv_gemm_qs(bn, true, ta, tb, tres, tout, g, c->num_sms, st));
        else CU_TRY(c, launch_conv_gemm_staged(bn, ta, tb, tres, tout, g, c->num_sms, st, 0, split));
    } else if (k.qs) {
        CU_TRY(c, launch_conv_gemm_qs(bn, false, ta, tb, ta, ta, g, c->num_sms, st));
    } else if (pair) {
        CU_TRY(c, launch_conv3x3_pair(ta, tb, g, c->num_sms, st, bn, split));
    } else if (nm) {
        CU_TRY(c, launch_conv_gemm_halo_nm(bn, ta, tb, g, c->num_sms, st));
    } else if (halo) {
        CU_TRY(c, launch_conv_gemm_halo(bn, ta, tb, g, c->num_sms, st, true, split));
    } else {
        CU_TRY(c, launch_conv_gemm(bn, ta, tb, g, c->num_sms, st, split));
    }
    c->launches++;
    if (c->sync_each) {   // SYLPH_SYNC_EACH=1 (debugging): name the convolution whose kernel faults
        const cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess)
            return c->fail("convolution %s (cin %d, cout %d, taps %d, tiles %d, staged %d, qs %d, pairsplit %d): %s", k.name, W.cin,
                           W.cout, W.taps, k.n_tiles, k.staged, k.qs, k.pairsplit, cudaGetErrorString(e));
    }
    if (c->profiling) {
        cudaEventRecord(tm.e1, st);
        c->timings.push_back(tm);
    }
    return 0;
}

struct StageTimer {
    sylph_ctx* c;
    cudaStream_t st;
    Timing tm;
    bool on;
    StageTimer(sylph_ctx* c_, const char* name, cudaStream_t s, double bytes = 0) : c(c_), st(s), on(c_->profiling) {
        if (on) {
            tm.name = name;
            tm.flops = 0;
            tm.bytes = bytes;
            cudaEventCreate(&tm.e0);
            cudaEventCreate(&tm.e1);
            cudaEventRecord(tm.e0, st);
        }
    }
    ~StageTimer() {
        if (on) {
            cudaEventRecord(tm.e1, st);
            c->timings.push_back(tm);
        }
    }
};

}  // namespace sylph

// ================================================================================================== C ABI
extern "C" {

const char* sylph_version(void) { return "sylph_b200 0.2 (sm_100a, tcgen05 implicit-GEMM; split-fp16 x3 'exact' and single-fp16 'fast' operands, FP32 accumulate)"; }

int sylph_create(sylph_ctx** out, int device, const sylph_model_config* cfg) {
    if (!out || !cfg) return 1;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device >= n) return 2;  // no CPU fallback
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 2;
    if (prop.major != 10) return 3;  // sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return 2;
    sylph_ctx* c = new sylph_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    c->cfg = *cfg;
    if (const char* e = getenv("SYLPH_STAGED_EPILOGUE")) c->staged_epilogue = atoi(e);
    if (const char* e = getenv("SYLPH_HALO")) c->halo_pipeline = atoi(e);
    if (const char* e = getenv("SYLPH_PAIR")) c->pair_kernel = atoi(e);
    if (const char* e = getenv("SYLPH_STEM16")) c->stem16 = atoi(e);
    if (const char* e = getenv("SYLPH_PAIR1X1")) c->pair1x1 = atoi(e);
    if (const char* e = getenv("SYLPH_FUSE_UPSAMPLE")) c->fuse_upsample = atoi(e);
    if (const char* e = getenv("SYLPH_NM")) c->nmerge = atoi(e);
    if (const char* e = getenv("SYLPH_QS")) c->quad = atoi(e);
    if (const char* e = getenv("SYLPH_SYNC_EACH")) c->sync_each = atoi(e);
    if (const char* e = getenv("SYLPH_CLS_POOLED")) c->cls_pooled = atoi(e);
    if (const char* e = getenv("SYLPH_LINEAR_GEMM_MIN")) c->linear_gemm_min = atoi(e);
    if (const char* e = getenv("SYLPH_ROI_PACKED")) c->roi_stride = atoi(e) ? 81 : 128;
    if (const char* e = getenv("SYLPH_ROI_ALIGN")) c->roi_separable = strcmp(e, "sample") == 0 ? 0 : 1;
    if (const char* e = getenv("SYLPH_PRECISION")) c->split = (strcmp(e, "fast") == 0 || strcmp(e, "0") == 0) ? 0 : 1;
    *out = c;
    if (cfg->pre_nms_topk * 5 > 8192) { c->fail("pre_nms_topk * 5 must be <= 8192"); }
    return 0;
}

void sylph_destroy(sylph_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    sylph_exchange_destroy(c);
    for (auto& kv : c->bufs) if (kv.second.p) cudaFree(kv.second.p);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->graph_arena) cudaFreeHost(c->graph_arena);
    if (c->detect_overflow_host) cudaFreeHost(c->detect_overflow_host);
    for (auto& kv : c->plane_sets) { cudaFree(kv.second->d_segs); cudaFree(kv.second->d_tile_seg); }
    delete c;  // prepared weights are released with the CUDA context
}

const char* sylph_last_error(const sylph_ctx* c) { return c ? c->err.c_str() : "null context"; }

int sylph_set_precision(sylph_ctx* c, int exact) {
    if (!c) return 1;
    if (c->finalized || !c->bufs.empty()) return c->fail("sylph_set_precision must be called before sylph_finalize_weights");
    c->split = exact ? 1 : 0;
    return 0;
}

int sylph_get_precision(const sylph_ctx* c) { return c ? c->split : -1; }

int sylph_load_tensor(sylph_ctx* c, const char* key, const float* host_data, const int64_t* shape, int ndim) {
    if (!c || !key || !host_data) return 1;
    HostTensor t;
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= static_cast<size_t>(shape[i]); }
    t.data.assign(host_data, host_data + n);
    c->staged[key] = std::move(t);
    c->finalized = false;
    return 0;
}

}  // extern "C"

// CodeGenerator weights (code_generator.py:520-646).  With c->inplace_uploads the prepared tensors are refreshed in their
// existing device buffers (sylph_update_code_generator: the weights after an optimiser step, same shapes).
static int prep_code_generator(sylph_ctx* c) {
    const sylph_model_config& f = c->cfg;
    const std::string cg = "code_generator.code_generator_head.";
    if (!c->inplace_uploads) {
        c->cg_tower.assign(f.cg_tower_layers, ConvW());
        c->cg_gn_w.assign(f.cg_tower_layers, nullptr);
        c->cg_gn_b.assign(f.cg_tower_layers, nullptr);
    }
    for (int i = 0; i < f.cg_tower_layers; ++i) {
        TRY(prep_conv(c, cg + "support_set_shared_tower." + std::to_string(3 * i), false, true, &c->cg_tower[i]));
        TRY(upload_vec(c, cg + "support_set_shared_tower." + std::to_string(3 * i + 1) + ".weight", &c->cg_gn_w[i], 256));
        TRY(upload_vec(c, cg + "support_set_shared_tower." + std::to_string(3 * i + 1) + ".bias", &c->cg_gn_b[i], 256));
    }
    TRY(prep_conv(c, cg + "support_set_cls_conv.0", false, true, &c->cg_cls));
    if (c->cg_cls.cout != 256) return c->fail("CODE_GENERATOR.OUT_CHANNEL must be 256 on this path");
    {   // the same layer as a [2304 -> 256] GEMM over the nine window means of a ROI (kernels_codegen.cuh, roi_window_means_kernel)
        const HostTensor *w = find_t(c, cg + "support_set_cls_conv.0.weight"), *b = find_t(c, cg + "support_set_cls_conv.0.bias");
        if (!w || !b || w->data.size() != static_cast<size_t>(256) * 256 * 9) return c->fail("support_set_cls_conv must be a 3x3 256 -> 256 convolution");
        HostTensor pw;
        pw.shape = {256, 2304, 1, 1};
        pw.data.resize(w->data.size());
        for (int o = 0; o < 256; ++o)
            for (int ch = 0; ch < 256; ++ch)
                for (int tap = 0; tap < 9; ++tap)
                    pw.data[static_cast<size_t>(o) * 2304 + tap * 256 + ch] = w->data[(static_cast<size_t>(o) * 256 + ch) * 9 + tap];
        c->staged["__cg_cls_pooled.weight"] = std::move(pw);
        c->staged["__cg_cls_pooled.bias"] = *b;
        TRY(prep_conv(c, "__cg_cls_pooled", false, true, &c->cg_cls_pooled));
    }
    if (f.cg_bias_layer) {
        const HostTensor *w = find_t(c, cg + "support_set_cls_bias.0.weight"), *b = find_t(c, cg + "support_set_cls_bias.0.bias");
        if (!w || !b || w->data.size() != 256 * 9) return c->fail("missing support_set_cls_bias tensors");
        std::vector<float> hw(9 * 256);
        for (int ch = 0; ch < 256; ++ch)
            for (int t = 0; t < 9; ++t) hw[t * 256 + ch] = w->data[ch * 9 + t];
        TRY(upload(c, hw, &c->cg_wbias));
        TRY(upload(c, b->data, &c->cg_bbias));
        TRY(scalar_of(c, cg + "bias_scale.scale", &c->bias_scale));
    } else {
        c->bias_scale = 1.f;
    }
    if (f.cg_weight_layer) {
        const HostTensor *w = find_t(c, cg + "support_set_cls_weight.0.weight"), *b = find_t(c, cg + "support_set_cls_weight.0.bias");
        if (!w || !b || w->data.size() != 256 * 9) return c->fail("missing support_set_cls_weight tensors");
        std::vector<float> hw(9 * 256);
        for (int ch = 0; ch < 256; ++ch)
            for (int t = 0; t < 9; ++t) hw[t * 256 + ch] = w->data[ch * 9 + t];
        TRY(upload(c, hw, &c->cg_wweight));
        TRY(upload(c, b->data, &c->cg_bweight));
    }
    if (f.cg_post_norm) {
        TRY(upload_vec(c, cg + "post_norm.weight", &c->post_gn_w, 256));
        TRY(upload_vec(c, cg + "post_norm.bias", &c->post_gn_b, 256));
    }
    c->conv_scale = 1.f;
    if (f.cg_has_conv_scale) TRY(scalar_of(c, cg + "conv_scale.scale", &c->conv_scale));
    return 0;
}

extern "C" {

int sylph_finalize_weights(sylph_ctx* c) {
    if (!c) return 1;
    CU_TRY(c, cudaSetDevice(c->device));
    const sylph_model_config& f = c->cfg;
    TRY(prep_stem(c, &c->stem));
    int nblocks[4];
    if (f.resnet_depth == 50) { int b[4] = {3, 4, 6, 3}; memcpy(nblocks, b, sizeof(b)); }
    else if (f.resnet_depth == 101) { int b[4] = {3, 4, 23, 3}; memcpy(nblocks, b, sizeof(b)); }
    else if (f.resnet_depth == 152) { int b[4] = {3, 8, 36, 3}; memcpy(nblocks, b, sizeof(b)); }
    else return c->fail("unsupported RESNETS.DEPTH %d", f.resnet_depth);
    c->stages.clear();
    for (int s = 0; s < 4; ++s) {
        std::vector<sylph_ctx::Block> blocks(nblocks[s]);
        for (int b = 0; b < nblocks[s]; ++b) {
            const std::string p = "backbone.bottom_up.res" + std::to_string(s + 2) + "." + std::to_string(b);
            blocks[b].has_sc = find_t(c, p + ".shortcut.weight") != nullptr;
            if (blocks[b].has_sc) TRY(prep_conv(c, p + ".shortcut", true, false, &blocks[b].sc));
            TRY(prep_conv(c, p + ".conv1", true, false, &blocks[b].c1));
            TRY(prep_conv(c, p + ".conv2", true, false, &blocks[b].c2));
            TRY(prep_conv(c, p + ".conv3", true, false, &blocks[b].c3));
        }
        c->stages.push_back(std::move(blocks));
    }
    for (int i = 0; i < 3; ++i) {
        TRY(prep_conv(c, "backbone.fpn_lateral" + std::to_string(i + 3), false, true, &c->lat[i]));
        TRY(prep_conv(c, "backbone.fpn_output" + std::to_string(i + 3), false, true, &c->outc[i]));
    }
    TRY(prep_conv(c, "backbone.top_block.p6", false, true, &c->p6));
    TRY(prep_conv(c, "backbone.top_block.p7", false, true, &c->p7));
    const std::string head = "proposal_generator.fcos_head.";
    auto prep_tower = [&](const std::string& name, int n, std::vector<ConvW>* tw, std::vector<float*>* gw,
                          std::vector<float*>* gb) -> int {
        tw->assign(n, ConvW());
        gw->assign(n, nullptr);
        gb->assign(n, nullptr);
        for (int i = 0; i < n; ++i) {
            TRY(prep_conv(c, name + std::to_string(3 * i), false, true, &(*tw)[i]));
            TRY(upload_vec(c, name + std::to_string(3 * i + 1) + ".weight", &(*gw)[i], 256));
            TRY(upload_vec(c, name + std::to_string(3 * i + 1) + ".bias", &(*gb)[i], 256));
        }
        return 0;
    };
    TRY(prep_tower(head + "cls_tower.", f.num_cls_convs, &c->cls_tower, &c->cls_gn_w, &c->cls_gn_b));
    TRY(prep_tower(head + "bbox_tower.", f.num_box_convs, &c->box_tower, &c->box_gn_w, &c->box_gn_b));
    {   // bbox_pred (4) + ctrness (1) + iou_overlap (1) fused into one 16-wide 3x3 convolution
        const HostTensor *wb = find_t(c, head + "bbox_pred.weight"), *wc = find_t(c, head + "ctrness.weight"),
                         *wi = find_t(c, head + "iou_overlap.weight"), *bb = find_t(c, head + "bbox_pred.bias"),
                         *bc = find_t(c, head + "ctrness.bias"), *bi = find_t(c, head + "iou_overlap.bias");
        if (!wb || !wc || !wi || !bb || !bc || !bi) return c->fail("missing FCOS predictor tensors");
        HostTensor w, b;
        w.shape = {6, 256, 3, 3};
        w.data = wb->data;
        w.data.insert(w.data.end(), wc->data.begin(), wc->data.end());
        w.data.insert(w.data.end(), wi->data.begin(), wi->data.end());
        b.shape = {6};
        b.data = bb->data;
        b.data.insert(b.data.end(), bc->data.begin(), bc->data.end());
        b.data.insert(b.data.end(), bi->data.begin(), bi->data.end());
        c->staged["__pred.weight"] = w;
        c->staged["__pred.bias"] = b;
        TRY(prep_conv(c, "__pred", false, true, &c->pred));
    }
    for (int l = 0; l < 5; ++l) {
        c->level_scale[l] = 1.f;
        if (f.use_scale) TRY(scalar_of(c, head + "scales." + std::to_string(l) + ".scale", &c->level_scale[l]));
    }
    if (f.generator == 2) {
        // base detector (EPISODIC_LEARNING off): no code generator; the class codes of sylph_detect are cls_logits
    } else if (f.generator == 1) {
        TRY(prep_roi_encoder(c));
    } else {
        TRY(prep_code_generator(c));
    }
    c->bias_value = -std::log((1.f - f.prior_prob) / f.prior_prob);
    c->staged.clear();
    c->finalized = true;
    c->weights_ready = true;
    c->cls_tower_t_ready = false;   // the transposed class-tower copies of the backward follow the new weights at their next use
    return 0;
}

int sylph_update_code_generator(sylph_ctx* c) {
    if (!c) return 1;
    if (!c->weights_ready) return c->fail("sylph_update_code_generator: call sylph_finalize_weights once first");
    if (c->cfg.generator != 0) return c->fail("sylph_update_code_generator: only the CodeGenerator plugin is trainable on this path");
    for (const auto& kv : c->staged)
        if (kv.first.rfind("code_generator.", 0) != 0) {
            c->staged.clear();   // the context keeps serving the weights it had
            c->finalized = true;
            return c->fail("sylph_update_code_generator: staged tensor %s is not a code-generator tensor "
                           "(sylph_finalize_weights reloads a whole model)", kv.first.c_str());
        }
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaDeviceSynchronize());   // no launch may still read the tensors that are about to change
    c->inplace_uploads = true;
    const int rc = prep_code_generator(c);
    c->inplace_uploads = false;
    c->staged.clear();
    c->finalized = true;
    return rc;
}

}  // extern "C"

namespace sylph {

// ------------------------------------------------------------------------------------------------ pyramid slot
static int setup_slot(sylph_ctx* c, int slot, int n, int hpad, int wpad, const int* lh, const int* lw, cudaStream_t st) {
    Slot& S = c->slots[slot];
    S.n = n;
    S.hpad = hpad;
    S.wpad = wpad;
    std::vector<Seg> segs;
    int row = 0;
    for (int l = 0; l < 5; ++l) {
        S.lh[l] = lh[l];
        S.lw[l] = lw[l];
        PlaneGeom g = regular_geom(row, lh[l], lw[l], 1);
        S.pg.lv[l] = g;
        S.pg.scale[l] = 1.0f / static_cast<float>(8 << l);
        S.level_row0[l] = row;
        auto v = geom_segs(g, n);
        segs.insert(segs.end(), v.begin(), v.end());
        row += n * g.rows_per_img;
    }
    S.level_row0[5] = row;
    std::string key = "pyr:" + std::to_string(n);
    for (int l = 0; l < 5; ++l) key += ":" + std::to_string(lh[l]) + "x" + std::to_string(lw[l]);
    TRY(make_plane_set(c, key, segs, row, &S.ps));
    void* p;
    TRY(ensure(c, "pyr" + std::to_string(slot), (static_cast<size_t>(row) + kBlockM) * c->ld(256) * 2, key, &p, st, true));
    S.pyr = static_cast<__half*>(p);
    S.valid = true;
    return 0;
}

// Output of the bottom-up trunk (stem, res2..res5) over a batch of images: the res3..res5 activations stay in the
// context buffers "bb.<key>.resN.x" (image-major planes) for the FPN of one or more pyramid slots.
struct TrunkOut {
    std::string key;
    int n = 0, hpad = 0, wpad = 0;
    PlaneGeom gs[4];
};

static int run_trunk(sylph_ctx* c, const std::string& key, int n, const void* const* images_dev, int is_u8, const int* hs,
                     const int* ws, int hpad, int wpad, cudaStream_t st, TrunkOut* out, bool normalized = false) {
    sylph_model_config f = c->cfg;
    CentreParams cp{};   // exact mode: re-centred integer pixels (kernels_misc.cuh); `normalized` input is de-normalised first
    for (int i = 0; i < 3; ++i) {
        cp.m_r[i] = std::nearbyint(f.pixel_mean[i]);
        cp.mean[i] = f.pixel_mean[i];
        cp.stdv[i] = f.pixel_std[i];
        cp.pad[i] = static_cast<float>((static_cast<double>(f.pixel_mean[i]) - std::nearbyint(static_cast<double>(f.pixel_mean[i]))) / 256.0);
    }
    cp.denorm = normalized ? 1 : 0;
    if (normalized) {   // the caller's batch is already (x - mean) / std: the fused preparation only re-lays it out
        for (int i = 0; i < 3; ++i) { f.pixel_mean[i] = 0.f; f.pixel_std[i] = 1.f; }
    }
    int hmax = 0, wmax = 0;
    for (int i = 0; i < n; ++i) { hmax = std::max(hmax, hs[i]); wmax = std::max(wmax, ws[i]); }
    const std::string sig = std::to_string(n) + ":" + std::to_string(hpad) + "x" + std::to_string(wpad);
    const std::string bb = "bb." + key + ".";
    // geometries
    const PlaneGeom g0 = regular_geom(0, hpad / 2, wpad / 2, 2);      // space-to-depth input / stem output
    PlaneGeom gs[4];                                                 // res2..res5
    for (int s = 0; s < 4; ++s) gs[s] = regular_geom(0, hpad >> (s + 2), wpad >> (s + 2), 1);
    out->key = key; out->n = n; out->hpad = hpad; out->wpad = wpad;
    for (int s = 0; s < 4; ++s) out->gs[s] = gs[s];

    std::shared_ptr<PlaneSet> ps0, pss[4];
    TRY(make_plane_set(c, "stem:" + sig, geom_segs(g0, n), n * g0.rows_per_img, &ps0));
    for (int s = 0; s < 4; ++s)
        TRY(make_plane_set(c, "res" + std::to_string(s + 2) + ":" + sig, geom_segs(gs[s], n), n * gs[s].rows_per_img, &pss[s]));

    auto buf = [&](const std::string& name, long long rows, int ch, bool zero, __half** out) -> int {
        void* p;
        TRY(ensure(c, name, (static_cast<size_t>(rows) + kBlockM) * c->ld(ch) * 2, sig, &p, st, zero));
        *out = static_cast<__half*>(p);
        return 0;
    };
    const long long rows0 = static_cast<long long>(n) * g0.rows_per_img;
    __half *S0, *S1;
    TRY(buf(bb + "s0", rows0, 16, true, &S0));
    TRY(buf(bb + "s1", rows0, 64, false, &S1));
    // ---- image descriptors
    std::vector<ImageDesc> descs(n);
    for (int i = 0; i < n; ++i) { descs[i].ptr = images_dev[i]; descs[i].h = hs[i]; descs[i].w = ws[i]; descs[i].is_u8 = is_u8; }
    void* d_desc;
    TRY(ensure(c, bb + "desc", n * sizeof(ImageDesc), "", &d_desc, st, false));
    TRY(stage_h2d(c, d_desc, descs.data(), n * sizeof(ImageDesc), st));
    {   // every layer over the whole batch
        {
            StageTimer t(c, "prep_stem_input", st, static_cast<double>(n) * (3.0 * hmax * wmax + 32.0 * g0.H * g0.W));
            if (c->split && is_u8)
                CU_TRY(c, launch_k(prep_stem_centred_u8_kernel, dim3((g0.W + 255) / 256, std::min(n * g0.H, c->num_sms * 8)), dim3(256), 0, st,
                    static_cast<const ImageDesc*>(d_desc), S0, g0, n, cp, c->nmerge ? 0 : 1));
            else if (c->split)
                CU_TRY(c, launch_k(prep_stem_centred_kernel, dim3(grid_for(static_cast<long long>(n) * g0.H * g0.W, 256, c->num_sms)), dim3(256), 0, st,
                    static_cast<const ImageDesc*>(d_desc), S0, g0, n, cp));
            else if (is_u8)
                CU_TRY(c, launch_k(prep_stem_input_u8_kernel, dim3((g0.W + 255) / 256, std::min(n * g0.H, c->num_sms * 8)), dim3(256), 0, st,
                    static_cast<const ImageDesc*>(d_desc), S0, g0, n, f.pixel_mean[0], f.pixel_mean[1], f.pixel_mean[2],
                    f.pixel_std[0], f.pixel_std[1], f.pixel_std[2], c->split));
            else
                CU_TRY(c, launch_k(prep_stem_input_kernel, dim3(grid_for(static_cast<long long>(n) * g0.H * g0.W, 256, c->num_sms)), dim3(256), 0, st,
                    static_cast<const ImageDesc*>(d_desc), S0, g0, n, f.pixel_mean[0], f.pixel_mean[1], f.pixel_mean[2],
                    f.pixel_std[0], f.pixel_std[1], f.pixel_std[2], c->split));
            CU_TRY(c, cudaGetLastError());
            c->launches++;
        }
        {   // stem: 7x7/2 conv + FrozenBN + ReLU as a 4-tap GEMM over overlapped 64-float rows
            ConvCall k{};
            k.W = &c->stem; k.A = S0; k.a_rows = rows0; k.a_cols = 64; k.a_ld = c->ld(16); k.ps = ps0.get();
            k.tile_begin = 0; k.n_tiles = static_cast<int>(rows0 / kBlockM); k.a_row_delta = 0; k.out = S1; k.ldc = c->ld(64);
            k.flags = kEpiRelu | kEpiMask; k.stem = (c->split && is_u8) ? 2 : 1; k.name = "stem7x7";   // 2: a_lo == 0, one A pass
            k.staged = c->staged_epilogue; k.out_rows = rows0;
            TRY(run_conv(c, k, st));
        }
        // ---- res2..res5
        __half* X = nullptr;  // running stage output
        int x_ch = 64;
        for (int s = 0; s < 4; ++s) {
            const PlaneGeom& g = gs[s];
            const long long rows = static_cast<long long>(n) * g.rows_per_img;
            const int tiles = static_cast<int>(rows / kBlockM);
            const int out_ch = 256 << s, bott = 64 << s, in_ch = x_ch;
            __half *IN, *Y, *T1, *T2;
            const std::string sn = bb + "res" + std::to_string(s + 2);
            TRY(buf(sn + ".in", rows, in_ch, true, &IN));
            TRY(buf(sn + ".x", rows, out_ch, false, &Y));
            TRY(buf(sn + ".t1", rows, bott, false, &T1));
            TRY(buf(sn + ".t2", rows, bott, false, &T2));
            {
                StageTimer t(c, s == 0 ? "maxpool3x3s2" : "subsample2", st,
                             static_cast<double>(n) * g.H * g.W * in_ch * 2 * (s == 0 ? 5.0 : 2.0));
                const long long work = static_cast<long long>(n) * g.H * g.W * (in_ch / 8);
                // split mode: a thread walks a strip of 8 output rows (kernels_misc.cuh)
                const long long work_mp = c->split ? static_cast<long long>(n) * ((g.H + 7) / 8) * g.W * (in_ch / 8) : work;
                if (s == 0) CU_TRY(c, launch_k(maxpool3x3s2_kernel, dim3(grid_for(work_mp, 256, c->num_sms)), dim3(256), 0, st, S1, IN, g0, g, n, in_ch, c->lo(in_ch)));
                else CU_TRY(c, launch_k(subsample2_kernel, dim3(grid_for(work * (c->split ? 2 : 1), 256, c->num_sms)), dim3(256), 0, st, X, IN, gs[s - 1], g, n, c->ld(in_ch)));
                CU_TRY(c, cudaGetLastError());
                c->launches++;
            }
            const auto& blocks = c->stages[s];
            for (size_t b = 0; b < blocks.size(); ++b) {
                const sylph_ctx::Block& B = blocks[b];
                const __half* bin = (b == 0) ? IN : Y;
                const int bin_ch = (b == 0) ? in_ch : out_ch;
                ConvCall k{};
                k.ps = pss[s].get(); k.tile_begin = 0; k.n_tiles = tiles; k.a_row_delta = 0; k.a_rows = rows;
                if (B.has_sc) {
                    if (b != 0) return c->fail("shortcut conv on a non-first block is not supported");
                    k.W = &B.sc; k.A = bin; k.a_cols = k.a_ld = c->ld(bin_ch); k.out = Y; k.ldc = c->ld(out_ch); k.flags = kEpiMask;
                    k.name = "res.shortcut1x1";
                    k.staged = c->staged_epilogue; k.out_rows = rows;
                    TRY(run_conv(c, k, st));
                    k.staged = 0;
                } else if (b == 0) {
                    return c->fail("identity shortcut on the first block of a stage is not supported");
                }
                k.residual = nullptr;
                k.W = &B.c1; k.A = bin; k.a_cols = k.a_ld = c->ld(bin_ch); k.out = T1; k.ldc = c->ld(bott);
                k.flags = kEpiRelu | kEpiMask; k.name = "res.conv1_1x1";
                k.staged = c->staged_epilogue; k.out_rows = rows;
                TRY(run_conv(c, k, st));
                k.staged = 0;
                k.W = &B.c2; k.A = T1; k.a_cols = k.a_ld = c->ld(bott); k.out = T2; k.ldc = c->ld(bott); k.name = "res.conv2_3x3";
                TRY(run_conv(c, k, st));
                k.W = &B.c3; k.A = T2; k.a_cols = k.a_ld = c->ld(bott); k.out = Y; k.ldc = c->ld(out_ch); k.residual = Y; k.ld_res = c->ld(out_ch);
                k.flags = kEpiRelu | kEpiMask | kEpiResidual; k.name = "res.conv3_1x1";
                k.staged = c->staged_epilogue; k.out_rows = rows;
                TRY(run_conv(c, k, st));
                k.staged = 0;
            }
            X = Y;
            x_ch = out_ch;
        }
        return 0;
    }
}

// FPN + P6/P7 of `n` images of a trunk batch (images first .. first + n - 1) into pyramid slot `slot`.
static int run_fpn(sylph_ctx* c, int slot, const TrunkOut& T, int first, int n, const int* hs, const int* ws, cudaStream_t st) {
    const PlaneGeom* gs = T.gs;
    int lh[5], lw[5];
    for (int l = 0; l < 3; ++l) { lh[l] = gs[l + 1].H; lw[l] = gs[l + 1].W; }
    lh[3] = (lh[2] + 1) / 2; lw[3] = (lw[2] + 1) / 2;
    lh[4] = (lh[3] + 1) / 2; lw[4] = (lw[3] + 1) / 2;
    TRY(setup_slot(c, slot, n, T.hpad, T.wpad, lh, lw, st));
    Slot& S = c->slots[slot];
    S.img_h.assign(hs, hs + n);
    S.img_w.assign(ws, ws + n);
    const std::string sig = std::to_string(n) + ":" + std::to_string(T.hpad) + "x" + std::to_string(T.wpad);
    const std::string fb = "fpn" + std::to_string(slot) + ".";
    auto buf = [&](const std::string& name, long long rows, int ch, bool zero, __half** out) -> int {
        void* p;
        TRY(ensure(c, name, (static_cast<size_t>(rows) + kBlockM) * c->ld(ch) * 2, sig, &p, st, zero));
        *out = static_cast<__half*>(p);
        return 0;
    };
    // ---- FPN (top-down): lateral buffers share the pyramid row indexing
    __half* LAT;
    {
        void* p;
        TRY(ensure(c, fb + "lat", (static_cast<size_t>(S.level_row0[5]) + kBlockM) * c->ld(256) * 2, sig, &p, st, true));
        LAT = static_cast<__half*>(p);
    }
    for (int l = 2; l >= 0; --l) {
        const PlaneGeom& g = S.pg.lv[l];
        const long long rows = static_cast<long long>(n) * g.rows_per_img;
        __half* XS = static_cast<__half*>(c->bufs["bb." + T.key + ".res" + std::to_string(l + 3) + ".x"].p);
        ConvCall k{};
        k.W = &c->lat[l]; k.A = XS; k.a_rows = static_cast<long long>(T.n) * g.rows_per_img; k.a_cols = k.a_ld = c->ld(512 << l); k.ps = S.ps.get();
        k.tile_begin = static_cast<int>(S.level_row0[l] / kBlockM); k.n_tiles = static_cast<int>(rows / kBlockM);
        k.a_row_delta = -static_cast<int>(S.level_row0[l]) + first * g.rows_per_img; k.out = LAT; k.ldc = c->ld(256);
        k.flags = kEpiMask; k.name = "fpn.lateral1x1";   // direct epilogue: the staged variant measured slower here (K >= 512)
        const bool fuse_up = (c->fuse_upsample || c->split) && l < 2;   // the separate add kernel exists for plain fp16 rows only
        if (fuse_up) {   // top-down add in the lateral convolution's epilogue: + LAT(level l + 1)(y / 2, x / 2), summed in fp32
            k.flags |= kEpiResidual | kEpiUpsample;
            k.residual = LAT; k.ld_res = c->ld(256); k.up_seg_delta = n;
        }
        TRY(run_conv(c, k, st));
        k.flags = kEpiMask; k.residual = nullptr; k.up_seg_delta = 0;
        if (l < 2 && !fuse_up) {
            StageTimer t(c, "fpn.upsample_add", st, static_cast<double>(n) * g.H * g.W * 256 * 2 * 2.25);
            CU_TRY(c, launch_k(upsample_add_kernel, dim3(grid_for(static_cast<long long>(n) * g.H * g.W * 32, 256, c->num_sms)), dim3(256), 0, st, 
                LAT, LAT, g, S.pg.lv[l + 1], n, 256));
            CU_TRY(c, cudaGetLastError());
            c->launches++;
        }
        k.W = &c->outc[l]; k.A = LAT; k.a_rows = S.level_row0[5]; k.a_cols = k.a_ld = c->ld(256); k.a_row_delta = 0;
        k.out = S.pyr; k.name = "fpn.output3x3";
        TRY(run_conv(c, k, st));
    }
    // ---- p6 = conv3x3/2(p5), p7 = conv3x3/2(relu(p6)): stride-1 conv, then the even positions
    {
        const PlaneGeom& g5 = S.pg.lv[2];
        const long long rows5 = static_cast<long long>(n) * g5.rows_per_img;
        __half *TMP, *R6;
        TRY(buf(fb + "p6tmp", S.level_row0[5], 256, false, &TMP));
        TRY(buf(fb + "p6relu", S.level_row0[5], 256, true, &R6));
        ConvCall k{};
        k.W = &c->p6; k.A = S.pyr; k.a_rows = S.level_row0[5]; k.a_cols = k.a_ld = c->ld(256); k.ps = S.ps.get();
        k.tile_begin = static_cast<int>(S.level_row0[2] / kBlockM); k.n_tiles = static_cast<int>(rows5 / kBlockM);
        k.a_row_delta = 0; k.out = TMP; k.ldc = c->ld(256); k.flags = kEpiMask; k.name = "fpn.p6_3x3";
        TRY(run_conv(c, k, st));
        const PlaneGeom& g6 = S.pg.lv[3];
        const int ld = c->ld(256);   // stride-2 gathers copy whole rows: the hi and lo halves travel together
        CU_TRY(c, launch_k(subsample2_kernel, dim3(grid_for(static_cast<long long>(n) * g6.H * g6.W * (ld / 8), 256, c->num_sms)), dim3(256), 0, st,
            TMP, S.pyr, g5, g6, n, ld));
        CU_TRY(c, cudaGetLastError());
        const long long rows6 = static_cast<long long>(n) * g6.rows_per_img;
        if (c->split)
            CU_TRY(c, launch_k(relu_copy_split_kernel, dim3(grid_for(rows6 * 32, 256, c->num_sms)), dim3(256), 0, st,
                static_cast<const __half*>(S.pyr + S.level_row0[3] * ld), R6 + S.level_row0[3] * ld, rows6, 256));
        else
            CU_TRY(c, launch_k(relu_copy_kernel, dim3(grid_for(rows6 * 32, 256, c->num_sms)), dim3(256), 0, st,
                reinterpret_cast<const uint4*>(S.pyr + S.level_row0[3] * 256), reinterpret_cast<uint4*>(R6 + S.level_row0[3] * 256),
                rows6 * 32));
        CU_TRY(c, cudaGetLastError());
        c->launches += 2;
        k.W = &c->p7; k.A = R6; k.tile_begin = static_cast<int>(S.level_row0[3] / kBlockM);
        k.n_tiles = static_cast<int>(rows6 / kBlockM); k.name = "fpn.p7_3x3";
        TRY(run_conv(c, k, st));
        const PlaneGeom& g7 = S.pg.lv[4];
        CU_TRY(c, launch_k(subsample2_kernel, dim3(grid_for(static_cast<long long>(n) * g7.H * g7.W * (ld / 8), 256, c->num_sms)), dim3(256), 0, st,
            TMP, S.pyr, g6, g7, n, ld));
        CU_TRY(c, cudaGetLastError());
        c->launches++;
    }
    return 0;
}

static int run_backbone(sylph_ctx* c, int slot, int n, const void* const* images_dev, int is_u8, const int* hs,
                        const int* ws, cudaStream_t st, bool normalized = false) {
    int hmax = 0, wmax = 0;
    for (int i = 0; i < n; ++i) { hmax = std::max(hmax, hs[i]); wmax = std::max(wmax, ws[i]); }
    TrunkOut T;
    TRY(run_trunk(c, "s" + std::to_string(slot), n, images_dev, is_u8, hs, ws, round_up(hmax, 32), round_up(wmax, 32), st, &T,
                  normalized));
    return run_fpn(c, slot, T, 0, n, hs, ws, st);
}

// conv3x3 + GroupNorm(32) + ReLU over `tiles` tiles of a plane set: conv epilogue accumulates the per-tile partial
// sums, finalize reduces them per plane, apply normalises in place.
static int conv_gn_relu(sylph_ctx* c, const ConvW& W, const float* gn_w, const float* gn_b, const __half* in,
                        long long a_rows, float* raw, __half* out, const PlaneSet* ps, int tile_begin, int tiles,
                        int seg_begin, int n_segs, float* gn_partial, float* gn_stats, const char* name, cudaStream_t st) {
    ConvCall k{};
    k.W = &W; k.A = in; k.a_rows = a_rows; k.a_cols = k.a_ld = c->ld(256); k.ps = ps; k.tile_begin = tile_begin; k.n_tiles = tiles;
    k.a_row_delta = 0; k.out = raw; k.ldc = 256; k.flags = kEpiGnStats | kEpiOutF32; k.gn_partial = gn_partial; k.name = name;
    TRY(run_conv(c, k, st));
    CU_TRY(c, launch_k(gn_finalize_kernel, dim3(n_segs), dim3(256), 0, st, gn_partial, ps->d_segs, seg_begin, n_segs, gn_stats));
    CU_TRY(c, cudaGetLastError());
    {
        const long long rows = static_cast<long long>(tiles) * kBlockM;
        StageTimer t(c, "gn_apply_relu", st, static_cast<double>(rows) * 256 * 6);
        CU_TRY(c, launch_k(gn_apply_relu_kernel, dim3(std::max(1, std::min(tiles, c->num_sms * 8))), dim3(256), 0, st,
            raw, out, gn_stats, gn_w, gn_b, ps->d_tile_seg, ps->d_segs, tile_begin, tiles, 1, c->split));
        CU_TRY(c, cudaGetLastError());
    }
    c->launches += 2;
    return 0;
}

// conv3x3 (256 -> 256) + GroupNorm + ReLU over ROI planes (code-generator tower, ROIEncoder pool / tokenizer convolutions).  The
// planes are packed at c->roi_stride rows, so the convolution runs over ceil(n_rois * stride / 128) tiles without per-tile
// GroupNorm sums (a tile holds rows of several ROIs) and roi_gn_relu_kernel normalises the 49 interior pixels of every ROI.
static int roi_conv_gn_relu(sylph_ctx* c, const ConvW& W, const float* gn_w, const float* gn_b, const __half* in, long long a_rows,
                            float* raw, __half* out, const PlaneSet* ps, int n_rois, const char* name, cudaStream_t st) {
    ConvCall k{};
    k.W = &W; k.A = in; k.a_rows = a_rows; k.a_cols = k.a_ld = c->ld(256); k.ps = ps; k.tile_begin = 0;
    k.n_tiles = static_cast<int>((static_cast<long long>(n_rois) * c->roi_stride + kBlockM - 1) / kBlockM);
    k.a_row_delta = 0; k.out = raw; k.ldc = 256; k.flags = kEpiOutF32; k.name = name;
    TRY(run_conv(c, k, st));
    StageTimer t(c, "roi_gn_relu", st, static_cast<double>(n_rois) * 49 * 256 * (4 + (c->split ? 4 : 2)));
    CU_TRY(c, launch_k(roi_gn_relu_kernel, dim3(n_rois), dim3(256), 0, st, static_cast<const float*>(raw), out, gn_w, gn_b, c->split, c->roi_stride));
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return 0;
}

static int run_conv(sylph_ctx* c, const ConvCall& k, cudaStream_t st);

static int run_linear(sylph_ctx* c, const sylph_ctx::Dense& d, const float* x, int ldx, float* y, int ldy, int T, int relu,
                      float add_const, cudaStream_t st, const PlaneSet* ps = nullptr) {
    if (d.in > 1024) return c->fail("linear layer wider than 1024 inputs");
    // Thousands of rows (class sweeps): fp32 rows -> fp16 (hi | lo) rows, then the tensor-core GEMM with an fp32 epilogue (bias,
    // ReLU).  `y` must have room for T rounded up to 128 rows; `ps` supplies the tile -> plane table the kernel's producer reads.
    if (ps != nullptr && d.has_gemm && T >= c->linear_gemm_min && add_const == 0.f && ldy == d.out && ldx == d.in) {
        const int t_pad = round_up(T, kBlockM);
        void* pa;
        TRY(ensure(c, "re.lin_a", (static_cast<size_t>(t_pad) + kBlockM) * c->ld(1024) * 2, "lin", &pa, st, true));
        const long long work = static_cast<long long>(T) * (d.in / 8);
        CU_TRY(c, launch_k(split_rows_kernel, dim3(grid_for(work, 256, c->num_sms)), dim3(256), 0, st, x, static_cast<__half*>(pa), T, d.in, c->split));
        c->launches++;
        ConvCall k{};
        k.W = &d.gemm; k.A = static_cast<const __half*>(pa); k.a_rows = t_pad; k.a_cols = k.a_ld = c->ld(d.in); k.ps = ps;
        k.tile_begin = 0; k.n_tiles = t_pad / kBlockM; k.a_row_delta = 0; k.out = y; k.ldc = ldy;
        k.flags = kEpiOutF32 | (relu ? kEpiRelu : 0); k.name = "roienc.dense_gemm";
        return run_conv(c, k, st);
    }
    CU_TRY(c, launch_k(linear_kernel, dim3(ceil_div(T, 8), ceil_div(d.out, 64)), dim3(256), 0, st, x, ldx,
                       static_cast<const float*>(d.w), static_cast<const float*>(d.b), y, ldy, T, d.in, d.out, relu, add_const));
    c->launches++;
    return 0;
}

// ROIEncoder.forward at inference (sylph/modeling/code_generator/roi_encoder.py:146-204) after the shared ROIAlign:
// r0 holds the pooled ROI planes; r1 / r2 are scratch planes of the same shape.
static int roi_encoder_codes(sylph_ctx* c, const Slot& S, int n_rois, int n_classes, const int* d_roi_image,
                             const int* d_class_off, const PlaneSet* ps, __half* r0, __half* r1, __half* r2, float* raw,
                             float* gp, float* gs, float* codes_out_dev, cudaStream_t st) {
    const sylph_model_config& f = c->cfg;
    const long long rows = round_up(static_cast<int>(static_cast<long long>(n_rois) * c->roi_stride), kBlockM);
    const int t_pad = round_up(n_rois, kBlockM);
    void *pctx, *ptok, *px0, *px1, *pxa, *ph, *pcls, *phd;
    TRY(ensure(c, "re.ctx", static_cast<size_t>(S.n) * 49 * 256 * 4, "", &pctx, st, false));
    TRY(ensure(c, "re.tokens", (static_cast<size_t>(t_pad) + kBlockM) * c->ld(12544) * 2, "tok", &ptok, st, true));
    TRY(ensure(c, "re.x0", static_cast<size_t>(t_pad) * 256 * 4, "", &px0, st, false));
    TRY(ensure(c, "re.x1", static_cast<size_t>(t_pad) * 256 * 4, "", &px1, st, false));
    TRY(ensure(c, "re.xa", static_cast<size_t>(t_pad) * 256 * 4, "", &pxa, st, false));
    TRY(ensure(c, "re.h", static_cast<size_t>(t_pad) * 1024 * 4, "", &ph, st, false));
    TRY(ensure(c, "re.cls", static_cast<size_t>(n_classes) * 256 * 4, "", &pcls, st, false));
    TRY(ensure(c, "re.hd", static_cast<size_t>(n_classes) * 1024 * 4 * 2, "", &phd, st, false));
    // FeatureFusionModuleV2: conv3x3 + GN + ReLU on the pooled features, then the MS_CAM context gate
    TRY(roi_conv_gn_relu(c, c->re_pool_conv, c->re_pool_gn_w, c->re_pool_gn_b, r0, rows, raw, r1, ps, n_rois, "roienc.pool_conv3x3", st));
    {
        StageTimer t(c, "roienc.context_pool", st, static_cast<double>(S.n) * 22400 * 256 * 2);
        CU_TRY(c, launch_k(context_pool_kernel, dim3(S.n, 49), dim3(256), 0, st, static_cast<const __half*>(S.pyr), S.pg,
                           static_cast<float*>(pctx), c->split));
        c->launches++;
    }
    {
        static PerDeviceOnce once;
        CU_TRY(c, once.run([] { return cudaFuncSetAttribute(ms_cam_gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMsCamSmem); }));
        void* pgate;
        TRY(ensure(c, "re.gate", static_cast<size_t>(S.n) * 49 * 256 * 4, "", &pgate, st, false));
        StageTimer t(c, "roienc.ms_cam", st, static_cast<double>(n_rois) * 49 * 256 * (c->split ? 8 : 4) + static_cast<double>(S.n) * 49 * 256 * 8);
        CU_TRY(c, launch_k(ms_cam_gate_kernel, dim3(S.n), dim3(256), static_cast<size_t>(kMsCamSmem), st,
                           static_cast<const float*>(pctx), c->re_cam, static_cast<float*>(pgate)));
        CU_TRY(c, launch_k(ms_cam_apply_kernel, dim3(grid_for(static_cast<long long>(n_rois) * c->roi_stride * 32, 256, c->num_sms)), dim3(256), 0, st,
                           static_cast<const float*>(pgate), d_roi_image, static_cast<const __half*>(r1), r2, n_rois, c->split, c->roi_stride));
        c->launches += 2;
    }
    // Tokenizer: NUM_CONV x (conv3x3 + GN + ReLU), flatten, fc1 (tensor-core GEMM over K = 12544), fc2.. (+ ReLU)
    __half* cur = r2;
    __half* nxt = r0;
    for (int i = 0; i < f.re_tok_convs; ++i) {
        TRY(roi_conv_gn_relu(c, c->re_tok_conv[i], c->re_tok_gn_w[i], c->re_tok_gn_b[i], cur, rows, raw, nxt, ps, n_rois, "roienc.tok_conv3x3", st));
        __half* done = cur;
        cur = nxt;
        nxt = (done == r2) ? r1 : done;
    }
    CU_TRY(c, launch_k(gather_tokens_kernel, dim3(grid_for(static_cast<long long>(n_rois) * 49 * 32, 256, c->num_sms)), dim3(256),
                       0, st, static_cast<const __half*>(cur), static_cast<__half*>(ptok), n_rois, c->split, c->roi_stride));
    c->launches++;
    {
        ConvCall k{};
        k.W = &c->re_fc1; k.A = static_cast<const __half*>(ptok); k.a_rows = t_pad; k.a_cols = k.a_ld = c->ld(12544); k.ps = ps;
        k.tile_begin = 0; k.n_tiles = t_pad / kBlockM; k.a_row_delta = 0; k.out = px0; k.ldc = 256;
        k.flags = kEpiRelu | kEpiOutF32; k.name = "roienc.fc1_gemm";
        TRY(run_conv(c, k, st));
    }
    float* x = static_cast<float*>(px0);
    float* y = static_cast<float*>(px1);
    float* xa = static_cast<float*>(pxa);
    float* h = static_cast<float*>(ph);
    StageTimer t(c, "roienc.token_mlp", st, 0);
    for (const auto& d : c->re_tok_fc) {
        TRY(run_linear(c, d, x, 256, y, 256, n_rois, 1, 0.f, st, ps));
        std::swap(x, y);
    }
    // TransformerEncoder, sequence length 1 per class (see kernels_roi_encoder.cuh)
    for (const auto& L : c->re_enc) {
        TRY(run_linear(c, L.attn, x, 256, xa, 256, n_rois, 0, 0.f, st, ps));
        CU_TRY(c, launch_k(add_layernorm_kernel, dim3(ceil_div(n_rois, 8)), dim3(256), 0, st, static_cast<const float*>(x),
                           static_cast<const float*>(xa), static_cast<const float*>(L.n1w), static_cast<const float*>(L.n1b), y, n_rois));
        TRY(run_linear(c, L.ff1, y, 256, h, 1024, n_rois, 1, 0.f, st, ps));
        TRY(run_linear(c, L.ff2, h, 1024, xa, 256, n_rois, 0, 0.f, st, ps));
        CU_TRY(c, launch_k(add_layernorm_kernel, dim3(ceil_div(n_rois, 8)), dim3(256), 0, st, static_cast<const float*>(y),
                           static_cast<const float*>(xa), static_cast<const float*>(L.n2w), static_cast<const float*>(L.n2b), x, n_rois));
        c->launches += 2;
    }
    CU_TRY(c, launch_k(token_mean_kernel, dim3(n_classes), dim3(256), 0, st, static_cast<const float*>(x), d_class_off,
                       static_cast<float*>(pcls)));
    c->launches++;
    // hyper-network heads: weights -> columns 0..255, prior + delta bias -> column 256 of the code rows
    const float prior = -std::log((1.f - 0.01f) / 0.01f);   // roi_encoder.py:141-142 (prior_prob = 0.01, hard-coded)
    for (int which = 0; which < 2; ++which) {
        const auto& hd = which == 0 ? c->re_whead : c->re_bhead;
        const float* in = static_cast<const float*>(pcls);
        int ld = 256;
        float* tmp[2] = {static_cast<float*>(phd), static_cast<float*>(phd) + static_cast<size_t>(n_classes) * 1024};
        for (size_t i = 0; i < hd.size(); ++i) {
            const bool last = i + 1 == hd.size();
            float* out = last ? (codes_out_dev + (which == 0 ? 0 : 256)) : tmp[i & 1];
            TRY(run_linear(c, hd[i], in, ld, out, last ? 257 : 1024, n_classes, last ? 0 : 1, (last && which == 1) ? prior : 0.f, st));
            in = out;
            ld = 1024;
        }
    }
    return 0;
}

}  // namespace sylph

// Towers + code-conditioned classifier + box / centre-ness / IoU predictors on a feature slot (MetaFCOSHead.forward,
// fcos.py:582-667): fills the "det.logits" ([rows][cout_pad] fp32) and "det.pred" ([rows][16] fp32) buffers.
struct HeadOut { float* logits; float* pred; int logit_stride; };

static int run_head(sylph_ctx* c, int slot, const float* codes_dev, int n_classes, cudaStream_t st, HeadOut* out,
                    cudaEvent_t codes_ready = nullptr, bool with_box_branch = true) {
    const sylph_model_config& f = c->cfg;
    const Slot& S = c->slots[slot];
    const long long rows = S.level_row0[5];
    const int tiles = static_cast<int>(rows / kBlockM);
    const int n_segs = 5 * S.n;
    // ---- code-conditioned classifier weights
    ConvW CW;
    CW.taps = 1; CW.ksize = 1; CW.cin = 256; CW.k_per_tap = c->split ? 768 : 256; CW.cout = n_classes; CW.bn = pick_bn(n_classes);
    CW.cout_pad = round_up(n_classes, CW.bn);
    CW.b_rows = CW.cout_pad;
    void *cw, *cb, *ta, *tb, *rawp, *lg, *pr, *gp, *gs;
    TRY(ensure(c, "det.code_w", static_cast<size_t>(CW.cout_pad) * CW.k_per_tap * 2, "", &cw, st, false));
    TRY(ensure(c, "det.code_b", static_cast<size_t>(CW.cout_pad) * 4, "", &cb, st, false));
    TRY(ensure(c, "det.ta", (static_cast<size_t>(rows) + kBlockM) * c->ld(256) * 2, "", &ta, st, false));
    TRY(ensure(c, "det.tb", (static_cast<size_t>(rows) + kBlockM) * c->ld(256) * 2, "", &tb, st, false));
    TRY(ensure(c, "det.raw", (static_cast<size_t>(rows) + kBlockM) * 256 * 4, "", &rawp, st, false));
    TRY(ensure(c, "det.logits", (static_cast<size_t>(rows) + kBlockM) * CW.cout_pad * 4, "", &lg, st, false));
    TRY(ensure(c, "det.pred", (static_cast<size_t>(rows) + kBlockM) * 16 * 4, "", &pr, st, false));
    TRY(ensure(c, "det.gn_partial", static_cast<size_t>(tiles) * 64 * 4, "", &gp, st, false));
    TRY(ensure(c, "det.gn_stats", static_cast<size_t>(n_segs) * 64 * 4, "", &gs, st, false));
    CW.w = static_cast<__half*>(cw);
    CW.bias = static_cast<float*>(cb);
    auto tower = [&](const std::vector<ConvW>& tw, const std::vector<float*>& gw, const std::vector<float*>& gb,
                     const char* name, __half** result, bool save = false) -> int {
        const __half* cur = S.pyr;
        __half* bufs2[2] = {static_cast<__half*>(ta), static_cast<__half*>(tb)};
        if (save) {
            c->saved_x.assign(tw.size(), nullptr);
            c->saved_raw.assign(tw.size(), nullptr);
            c->saved_stats.assign(tw.size(), nullptr);
        }
        for (size_t i = 0; i < tw.size(); ++i) {
            __half* o = bufs2[i & 1];
            float* raw_i = static_cast<float*>(rawp);
            float* gs_i = static_cast<float*>(gs);
            if (save) {   // training: every layer keeps its own output planes, convolution output and statistics
                void *po, *pr, *pg;
                const std::string tag = std::to_string(i);
                TRY(ensure(c, "det.cls_x" + tag, (static_cast<size_t>(rows) + kBlockM) * c->ld(256) * 2, "", &po, st, false));
                TRY(ensure(c, "det.cls_raw" + tag, (static_cast<size_t>(rows) + kBlockM) * 256 * 4, "", &pr, st, false));
                TRY(ensure(c, "det.cls_gs" + tag, static_cast<size_t>(n_segs) * 64 * 4, "", &pg, st, false));
                o = static_cast<__half*>(po);
                raw_i = static_cast<float*>(pr);
                gs_i = static_cast<float*>(pg);
                c->saved_x[i] = cur;
                c->saved_raw[i] = raw_i;
                c->saved_stats[i] = gs_i;
            }
            TRY(conv_gn_relu(c, tw[i], gw[i], gb[i], cur, rows, raw_i, o, S.ps.get(), 0, tiles, 0, n_segs,
                             static_cast<float*>(gp), gs_i, name, st));
            cur = o;
        }
        *result = const_cast<__half*>(cur);
        return 0;
    };
    // Everything that does not depend on the class codes first (box tower + predictors, class tower): a caller that
    // generates the codes on another stream overlaps that work with these 9 tensor-bound launches (sylph_detect_after).
    __half* x;
    if (with_box_branch) {   // skipped by the training forward when the box losses are off (FREEZE_BBOX_BRANCH, sylph_set_loss_box_branch)
        TRY(tower(c->box_tower, c->box_gn_w, c->box_gn_b, "head.bbox_tower3x3", &x));
        ConvCall k{};
        k.W = &c->pred; k.A = x; k.a_rows = rows; k.a_cols = k.a_ld = c->ld(256); k.ps = S.ps.get(); k.tile_begin = 0; k.n_tiles = tiles;
        k.a_row_delta = 0; k.out = pr; k.ldc = 16; k.flags = kEpiOutF32; k.name = "head.pred3x3";
        TRY(run_conv(c, k, st));
    }
    TRY(tower(c->cls_tower, c->cls_gn_w, c->cls_gn_b, "head.cls_tower3x3", &x, c->train_save));
    if (codes_ready != nullptr) CU_TRY(c, cudaStreamWaitEvent(st, codes_ready, 0));
    CU_TRY(c, launch_k(pack_code_weights_kernel, dim3(ceil_div(static_cast<long long>(CW.cout_pad) * 256, 256)), dim3(256), 0, st,
        codes_dev, n_classes, CW.cout_pad, f.generator != 0 ? 1 : f.cg_use_bias, f.generator == 1 ? c->cond_scale : 1.f, CW.w, CW.bias, c->split));
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    {
        ConvCall k{};
        k.W = &CW; k.A = x; k.a_rows = rows; k.a_cols = k.a_ld = c->ld(256); k.ps = S.ps.get(); k.tile_begin = 0; k.n_tiles = tiles;
        k.a_row_delta = 0; k.out = lg; k.ldc = CW.cout_pad; k.flags = kEpiOutF32; k.name = "head.cond_cls1x1";
        TRY(run_conv(c, k, st));
    }
    c->last_detect_slot = slot;
    c->last_detect_classes = n_classes;
    c->last_logit_stride = CW.cout_pad;
    c->last_cls_tower = x;
    out->logits = static_cast<float*>(lg);
    out->pred = with_box_branch ? static_cast<float*>(pr) : nullptr;
    out->logit_stride = CW.cout_pad;
    return 0;
}

extern "C" {

int sylph_extract_features(sylph_ctx* c, int slot, int n_images, const float* const* images_dev, const int* heights,
                           const int* widths, void* stream) {
    if (!c) return 1;
    if (!c->finalized) return c->fail("weights not finalized");
    if (slot < 0 || slot >= SYLPH_NUM_SLOTS || n_images <= 0) return c->fail("bad slot / image count");
    CU_TRY(c, cudaSetDevice(c->device));
    return run_backbone(c, slot, n_images, reinterpret_cast<const void* const*>(images_dev), 0, heights, widths,
                        static_cast<cudaStream_t>(stream));
}

int sylph_extract_features_u8(sylph_ctx* c, int slot, int n_images, const uint8_t* const* images_dev, const int* heights,
                              const int* widths, void* stream) {
    if (!c) return 1;
    if (!c->finalized) return c->fail("weights not finalized");
    if (slot < 0 || slot >= SYLPH_NUM_SLOTS || n_images <= 0) return c->fail("bad slot / image count");
    CU_TRY(c, cudaSetDevice(c->device));
    return run_backbone(c, slot, n_images, reinterpret_cast<const void* const*>(images_dev), 1, heights, widths,
                        static_cast<cudaStream_t>(stream));
}

int sylph_extract_features_normalized(sylph_ctx* c, int slot, int n_images, const float* batch_dev, int height, int width,
                                      void* stream) {
    if (!c) return 1;
    if (!c->finalized) return c->fail("weights not finalized");
    if (slot < 0 || slot >= SYLPH_NUM_SLOTS || n_images <= 0 || !batch_dev || height <= 0 || width <= 0)
        return c->fail("bad slot / batch");
    CU_TRY(c, cudaSetDevice(c->device));
    std::vector<const void*> ptrs(n_images);
    std::vector<int> hs(n_images, height), ws(n_images, width);
    for (int i = 0; i < n_images; ++i) ptrs[i] = batch_dev + static_cast<size_t>(i) * 3 * height * width;
    return run_backbone(c, slot, n_images, ptrs.data(), 0, hs.data(), ws.data(), static_cast<cudaStream_t>(stream), true);
}

int sylph_extract_features_multi(sylph_ctx* c, int n_groups, const int* slots, const int* counts,
                                 const void* const* images_dev, int is_u8, const int* heights, const int* widths, void* stream) {
    if (!c) return 1;
    if (!c->finalized) return c->fail("weights not finalized");
    if (n_groups <= 0 || n_groups > SYLPH_NUM_SLOTS || !slots || !counts) return c->fail("bad group list");
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int total = 0, hp = -1, wp = -1;
    bool same = true;
    for (int g = 0; g < n_groups; ++g) {
        if (slots[g] < 0 || slots[g] >= SYLPH_NUM_SLOTS || counts[g] <= 0) return c->fail("bad slot / image count in group %d", g);
        for (int h = 0; h < g; ++h)
            if (slots[h] == slots[g]) return c->fail("slot %d listed twice", slots[g]);
        int hmax = 0, wmax = 0;
        for (int i = 0; i < counts[g]; ++i) { hmax = std::max(hmax, heights[total + i]); wmax = std::max(wmax, widths[total + i]); }
        const int ghp = round_up(hmax, 32), gwp = round_up(wmax, 32);
        if (g > 0 && (ghp != hp || gwp != wp)) same = false;
        hp = ghp; wp = gwp;
        total += counts[g];
    }
    // ImageList.from_tensors pads every reference call to ITS OWN batch maximum: the groups may only share one trunk
    // batch when they pad to the same size; otherwise they run one after the other exactly like separate calls.
    if (!same || n_groups == 1) {
        int off = 0;
        for (int g = 0; g < n_groups; ++g) {
            TRY(run_backbone(c, slots[g], counts[g], images_dev + off, is_u8, heights + off, widths + off, st));
            off += counts[g];
        }
        return 0;
    }
    TrunkOut T;
    TRY(run_trunk(c, "m", total, images_dev, is_u8, heights, widths, hp, wp, st, &T));
    int off = 0;
    for (int g = 0; g < n_groups; ++g) {
        TRY(run_fpn(c, slots[g], T, off, counts[g], heights + off, widths + off, st));
        off += counts[g];
    }
    return 0;
}

int sylph_import_features(sylph_ctx* c, int slot, int n_images, int padded_h, int padded_w,
                          const float* const* level_ptrs_dev, const int* level_h, const int* level_w, void* stream) {
    if (!c) return 1;
    if (slot < 0 || slot >= SYLPH_NUM_SLOTS || n_images <= 0) return c->fail("bad slot / image count");
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TRY(setup_slot(c, slot, n_images, padded_h, padded_w, level_h, level_w, st));
    Slot& S = c->slots[slot];
    S.img_h.assign(n_images, padded_h);
    S.img_w.assign(n_images, padded_w);
    for (int l = 0; l < 5; ++l) {
        const long long work = static_cast<long long>(n_images) * 256 * level_h[l] * level_w[l];
        CU_TRY(c, launch_k(import_nchw_kernel, dim3(grid_for(work, 256, c->num_sms)), dim3(256), 0, st, level_ptrs_dev[l], S.pyr, S.pg.lv[l], n_images, 256, c->lo(256)));
        CU_TRY(c, cudaGetLastError());
        c->launches++;
    }
    return 0;
}

int sylph_set_image_sizes(sylph_ctx* c, int slot, int n_images, const int* heights, const int* widths) {
    if (!c) return 1;
    if (slot < 0 || slot >= SYLPH_NUM_SLOTS || !c->slots[slot].valid) return c->fail("slot %d holds no features", slot);
    Slot& S = c->slots[slot];
    if (n_images != S.n || !heights || !widths) return c->fail("sylph_set_image_sizes: the slot holds %d images", S.n);
    for (int i = 0; i < n_images; ++i)
        if (heights[i] <= 0 || widths[i] <= 0 || heights[i] > S.hpad || widths[i] > S.wpad)
            return c->fail("image %d: size %d x %d outside the padded batch %d x %d", i, heights[i], widths[i], S.hpad, S.wpad);
    S.img_h.assign(heights, heights + n_images);
    S.img_w.assign(widths, widths + n_images);
    return 0;
}

int sylph_feature_shape(sylph_ctx* c, int slot, int* n_images, int* padded_h, int* padded_w, int* level_h, int* level_w) {
    if (!c) return 1;
    if (slot < 0 || slot >= SYLPH_NUM_SLOTS || !c->slots[slot].valid) return c->fail("slot %d holds no features", slot);
    const Slot& S = c->slots[slot];
    if (n_images) *n_images = S.n;
    if (padded_h) *padded_h = S.hpad;
    if (padded_w) *padded_w = S.wpad;
    for (int l = 0; l < 5; ++l) {
        if (level_h) level_h[l] = S.lh[l];
        if (level_w) level_w[l] = S.lw[l];
    }
    return 0;
}

int sylph_export_features(sylph_ctx* c, int slot, int level, float* out_dev, void* stream) {
    if (!c) return 1;
    if (slot < 0 || slot >= SYLPH_NUM_SLOTS || !c->slots[slot].valid || level < 0 || level > 4) return c->fail("bad slot/level");
    const Slot& S = c->slots[slot];
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long work = static_cast<long long>(S.n) * 256 * S.lh[level] * S.lw[level];
    CU_TRY(c, launch_k(export_nchw_kernel<__half>, dim3(grid_for(work, 256, c->num_sms)), dim3(256), 0, st, S.pyr, out_dev, S.pg.lv[level], S.n, 256, c->ld(256), 0, 1.f, 0, c->lo(256)));
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return 0;
}

int sylph_generate_codes(sylph_ctx* c, int slot, int n_rois, const float* boxes_host, const int* roi_image,
                         int n_classes, const int* class_offsets, float* codes_out_dev, int64_t* levels_out_dev,
                         void* stream) {
    if (!c) return 1;
    if (!c->finalized) return c->fail("weights not finalized");
    if (slot < 0 || slot >= SYLPH_NUM_SLOTS || !c->slots[slot].valid) return c->fail("slot %d holds no features", slot);
    if (c->cfg.generator == 2) return c->fail("this model has no code generator (MODEL.META_LEARN.EPISODIC_LEARNING is off)");
    if (n_rois <= 0 || n_classes <= 0) return c->fail("empty support set");  // select_a_mask raises ValueError
    const Slot& S = c->slots[slot];
    for (int i = 0; i < n_rois; ++i)
        if (roi_image[i] < 0 || roi_image[i] >= S.n) return c->fail("roi_image[%d]=%d out of range", i, roi_image[i]);
    if (class_offsets[0] != 0 || class_offsets[n_classes] != n_rois) return c->fail("class_offsets must span all ROIs");
    for (int k = 0; k < n_classes; ++k)
        if (class_offsets[k + 1] <= class_offsets[k]) return c->fail("class %d has no support ROI", k);
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const sylph_model_config& f = c->cfg;
    const long long rows = round_up(static_cast<int>(static_cast<long long>(n_rois) * c->roi_stride), kBlockM);   // ROI planes, packed
    void *pb, *pi, *po, *r0, *r1, *r2, *rawp, *gp, *gs, *sc;
    TRY(ensure(c, "cg.boxes", n_rois * 16, "", &pb, st, false));
    TRY(ensure(c, "cg.roi_image", n_rois * 4, "", &pi, st, false));
    TRY(ensure(c, "cg.class_off", (n_classes + 1) * 4, "", &po, st, false));
    TRY(ensure(c, "cg.r0", (rows + kBlockM) * c->ld(256) * 2, "roi", &r0, st, true));
    TRY(ensure(c, "cg.r1", (rows + kBlockM) * c->ld(256) * 2, "roi", &r1, st, true));
    TRY(ensure(c, "cg.r2", (rows + kBlockM) * c->ld(256) * 2, "roi", &r2, st, true));
    TRY(ensure(c, "cg.raw", (rows + kBlockM) * 256 * 4, "roi", &rawp, st, false));
    TRY(ensure(c, "cg.gn_partial", static_cast<size_t>(n_rois) * 64 * 4, "", &gp, st, false));
    TRY(ensure(c, "cg.gn_stats", static_cast<size_t>(n_rois) * 64 * 4, "", &gs, st, false));
    TRY(ensure(c, "cg.shot", static_cast<size_t>(n_rois) * 257 * 4, "", &sc, st, false));
    {   // the boxes may already live on the device (a static buffer refreshed between CUDA-graph replays)
        cudaPointerAttributes pa{};
        const bool on_device = cudaPointerGetAttributes(&pa, boxes_host) == cudaSuccess && pa.type == cudaMemoryTypeDevice;
        if (!on_device) cudaGetLastError();   // plain host memory is reported as an error by older drivers: clear it
        if (on_device) pb = const_cast<float*>(boxes_host);
        else TRY(stage_h2d(c, pb, boxes_host, static_cast<size_t>(n_rois) * 16, st));
    }
    TRY(stage_h2d(c, pi, roi_image, static_cast<size_t>(n_rois) * 4, st));
    TRY(stage_h2d(c, po, class_offsets, static_cast<size_t>(n_classes + 1) * 4, st));
    // plane set of the ROI planes (one 9x9 plane = one tile per ROI); grows with the largest ROI count seen
    std::shared_ptr<PlaneSet> ps;
    {
        const int cap = std::max(64, 1 << static_cast<int>(std::ceil(std::log2(static_cast<double>(n_rois)))));
        const PlaneGeom g = regular_geom(0, 7, 7, 1);
        TRY(make_plane_set(c, "roi:" + std::to_string(cap), geom_segs(g, cap), cap * 128, &ps));
    }
    {
        StageTimer t(c, "roi_align", st, static_cast<double>(n_rois) * (50176.0 + 0.5e6) * (c->split ? 2 : 1));
        // separable form (every footprint pixel read once); SYLPH_ROI_ALIGN=sample keeps the per-sample kernel, which also
        // serves planes wider than the separable kernel's weight tables
        int span = 1;
        for (int l = 0; l < 5; ++l) span = std::max(span, std::max(S.lh[l], S.lw[l]));
        if (c->roi_separable && span <= kRoiSepMaxSpan) {
            const size_t smem = static_cast<size_t>(14) * span * sizeof(float);
            CU_TRY(c, launch_k(roi_align_separable_kernel, dim3(n_rois), dim3(256), smem, st, static_cast<const __half*>(S.pyr), S.pg,
                               static_cast<const float*>(pb), static_cast<const int*>(pi), static_cast<__half*>(r0),
                               reinterpret_cast<long long*>(levels_out_dev), c->split, span, c->roi_stride));
        } else {
            CU_TRY(c, launch_k(roi_align_kernel, dim3(n_rois, 7), dim3(256), 0, st, S.pyr, S.pg, static_cast<const float*>(pb), static_cast<const int*>(pi),
                                                             static_cast<__half*>(r0), reinterpret_cast<long long*>(levels_out_dev), c->split, c->roi_stride));
        }
        CU_TRY(c, cudaGetLastError());
        c->launches++;
    }
    c->last_n_rois = n_rois;
    if (f.generator == 1)
        return roi_encoder_codes(c, S, n_rois, n_classes, static_cast<const int*>(pi), static_cast<const int*>(po), ps.get(),
                                 static_cast<__half*>(r0), static_cast<__half*>(r1), static_cast<__half*>(r2),
                                 static_cast<float*>(rawp), static_cast<float*>(gp), static_cast<float*>(gs), codes_out_dev, st);
    __half* cur = static_cast<__half*>(r0);
    __half* nxt = static_cast<__half*>(r1);
    float* raw = static_cast<float*>(rawp);
    for (int i = 0; i < f.cg_tower_layers; ++i) {
        TRY(roi_conv_gn_relu(c, c->cg_tower[i], c->cg_gn_w[i], c->cg_gn_b[i], cur, rows, raw, nxt, ps.get(), n_rois, "codegen.tower3x3", st));
        cur = nxt;
        nxt = (cur == r1) ? static_cast<__half*>(r2) : static_cast<__half*>(r1);
    }
    const float* pooled = nullptr;
    if (c->cls_pooled) {
        // pool before the cls convolution (exact algebra, kernels_codegen.cuh): nine window means per ROI, then one GEMM with
        // M = n_rois, K = 2304 -- the per-pixel convolution over one 128-row tile per ROI did 128x the tensor work
        const int t_pad = round_up(n_rois, kBlockM);
        void *pw, *pp;
        TRY(ensure(c, "cg.win", static_cast<size_t>(t_pad + kBlockM) * c->ld(2304) * 2, "win", &pw, st, true));
        TRY(ensure(c, "cg.pooled", static_cast<size_t>(t_pad + kBlockM) * 256 * 4, "", &pp, st, false));
        {
            StageTimer t(c, "codegen.window_means", st, static_cast<double>(n_rois) * (49.0 * 256 + 2304) * (c->split ? 4 : 2));
            CU_TRY(c, launch_k(roi_window_means_kernel, dim3(n_rois), dim3(256), 0, st, static_cast<const __half*>(cur), static_cast<__half*>(pw), c->split, c->roi_stride));
            CU_TRY(c, cudaGetLastError());
            c->launches++;
        }
        ConvCall k{};
        k.W = &c->cg_cls_pooled; k.A = static_cast<const __half*>(pw); k.a_rows = t_pad; k.a_cols = k.a_ld = c->ld(2304); k.ps = ps.get();
        k.tile_begin = 0; k.n_tiles = t_pad / kBlockM; k.a_row_delta = 0; k.out = pp; k.ldc = 256; k.flags = kEpiOutF32;
        k.name = "codegen.cls_gemm";
        TRY(run_conv(c, k, st));
        pooled = static_cast<const float*>(pp);
    } else {
        ConvCall k{};
        k.W = &c->cg_cls; k.A = cur; k.a_rows = rows; k.a_cols = k.a_ld = c->ld(256); k.ps = ps.get(); k.tile_begin = 0;
        k.n_tiles = static_cast<int>(rows / kBlockM); k.a_row_delta = 0; k.out = raw; k.ldc = 256; k.flags = kEpiOutF32; k.name = "codegen.cls_conv3x3";
        TRY(run_conv(c, k, st));
    }
    {
        StageTimer t(c, "codegen.tail", st, static_cast<double>(n_rois) * 2 * 49 * 256 * 4);
        void* pwl = nullptr;   // per-shot weight logits (WEIGHT_LAYER): softmax over the shots of a class replaces 1 / K
        if (f.cg_weight_layer) TRY(ensure(c, "cg.wlogit", static_cast<size_t>(n_rois) * 4, "", &pwl, st, false));
        CU_TRY(c, launch_k(shot_code_kernel, dim3(n_rois), dim3(256), 0, st, raw, cur, c->cg_wbias, c->cg_bbias, f.cg_bias_layer, f.cg_bias_l2_norm,
                                                 static_cast<float*>(sc), c->split, pooled, c->roi_stride,
                                                 static_cast<const float*>(c->cg_wweight), static_cast<const float*>(c->cg_bweight), static_cast<float*>(pwl)));
        CU_TRY(c, cudaGetLastError());
        CU_TRY(c, launch_k(class_mean_kernel, dim3(n_classes), dim3(288), 0, st, static_cast<const float*>(sc), static_cast<const int*>(po), codes_out_dev,
                           static_cast<const float*>(pwl)));
        CU_TRY(c, cudaGetLastError());
        c->launches += 2;
    }
    return 0;
}

int sylph_export_roi_features(sylph_ctx* c, float* out_dev, void* stream) {
    if (!c) return 1;
    if (c->last_n_rois <= 0) return c->fail("no ROI features: call sylph_generate_codes first");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CU_TRY(c, launch_k(export_roi_kernel, dim3(grid_for(static_cast<long long>(c->last_n_rois) * 256 * 49, 256, c->num_sms)), dim3(256), 0, st, 
        static_cast<const __half*>(c->bufs["cg.r0"].p), out_dev, c->last_n_rois, c->split, c->roi_stride));
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return 0;
}

int sylph_normalize_codes(sylph_ctx* c, const float* raw_codes_dev, float* out_codes_dev, int n_classes, void* stream) {
    if (!c) return 1;
    if (!c->finalized) return c->fail("weights not finalized");
    if (n_classes <= 0) return 0;  // forward_normalize_code returns an empty list unchanged
    const sylph_model_config& f = c->cfg;
    if (f.generator == 1) return c->fail("ROIEncoder codes are final: there is no code normalisation for this generator");
    if (f.generator == 2) return c->fail("this model has no code generator (MODEL.META_LEARN.EPISODIC_LEARNING is off)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CU_TRY(c, launch_k(normalize_codes_kernel, dim3(n_classes), dim3(256), 0, st, raw_codes_dev, out_codes_dev, c->post_gn_w, c->post_gn_b, f.cg_post_norm,
                                                      f.cg_conv_l2_norm, c->conv_scale, c->bias_scale, c->bias_value));
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return 0;
}

// ------------------------------------------------------------------------------------------------ code exchange
static size_t exchange_bytes(int max_classes) {
    return sizeof(ExchangeState) + static_cast<size_t>(2) * max_classes * SYLPH_CODE_STRIDE * sizeof(float);
}

int sylph_exchange_create(sylph_ctx* c, int world, int rank, int max_classes, uint8_t* handle_out) {
    if (!c) return 1;
    static_assert(sizeof(cudaIpcMemHandle_t) == SYLPH_IPC_HANDLE_BYTES, "IPC handle size");
    if (world < 1 || world > kMaxExchangePeers) return c->fail("exchange: world size %d outside 1..%d", world, kMaxExchangePeers);
    if (rank < 0 || rank >= world || max_classes < 1 || !handle_out) return c->fail("exchange: bad arguments");
    if (c->xch.local) return c->fail("exchange: already created (sylph_exchange_destroy first)");
    CU_TRY(c, cudaSetDevice(c->device));
    void* p = nullptr;
    CU_TRY(c, cudaMalloc(&p, exchange_bytes(max_classes)));   // cudaMalloc (not a pool) so that the block can be IPC-exported
    CU_TRY(c, cudaMemset(p, 0, exchange_bytes(max_classes)));
    CU_TRY(c, cudaDeviceSynchronize());
    memset(handle_out, 0, SYLPH_IPC_HANDLE_BYTES);
    if (world > 1) {
        cudaIpcMemHandle_t h;
        cudaError_t e = cudaIpcGetMemHandle(&h, p);
        if (e != cudaSuccess) {
            cudaFree(p);
            return c->fail("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        }
        memcpy(handle_out, &h, SYLPH_IPC_HANDLE_BYTES);
    }
    unsigned int* host_err = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void**>(&host_err), sizeof(unsigned int), cudaHostAllocDefault) != cudaSuccess) {
        cudaFree(p);
        return c->fail("exchange: cudaHostAlloc failed");
    }
    *host_err = 0u;
    c->xch = sylph_ctx::Exchange();
    c->xch.world = world;
    c->xch.rank = rank;
    c->xch.max_classes = max_classes;
    c->xch.local = p;
    c->xch.host_err = host_err;
    if (const char* e = getenv("SYLPH_EXCHANGE_TIMEOUT_MS")) c->xch.timeout_ns = static_cast<unsigned long long>(atoll(e)) * 1000000ull;
    return 0;
}

int sylph_exchange_connect(sylph_ctx* c, const uint8_t* handles_all) {
    if (!c) return 1;
    sylph_ctx::Exchange& x = c->xch;
    if (!x.local) return c->fail("exchange: sylph_exchange_create first");
    if (x.connected) return c->fail("exchange: already connected");
    if (x.world > 1 && !handles_all) return c->fail("exchange: handles of all ranks required");
    CU_TRY(c, cudaSetDevice(c->device));
    for (int r = 0; r < x.world; ++r) {
        void* base = x.local;
        if (r != x.rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, handles_all + static_cast<size_t>(r) * SYLPH_IPC_HANDLE_BYTES, SYLPH_IPC_HANDLE_BYTES);
            cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                for (int q = 0; q < r; ++q)
                    if (q != x.rank && x.peer_base[q]) { cudaIpcCloseMemHandle(x.peer_base[q]); x.peer_base[q] = nullptr; }
                return c->fail("cudaIpcOpenMemHandle(rank %d) failed: %s (ranks must be GPUs of one box with peer access)", r,
                               cudaGetErrorString(e));
            }
        }
        x.peer_base[r] = base;
        x.peers.arrived[r] = &reinterpret_cast<ExchangeState*>(base)->arrived;
        x.peers.codes[r] = reinterpret_cast<float*>(static_cast<uint8_t*>(base) + sizeof(ExchangeState));
    }
    x.peers.world = x.world;
    x.connected = true;
    return 0;
}

int sylph_normalize_codes_exchange(sylph_ctx* c, const float* raw_codes_dev, int n_local, int class_offset, int n_total,
                                   float* all_codes_out_dev, void* stream) {
    if (!c) return 1;
    if (!c->finalized) return c->fail("weights not finalized");
    sylph_ctx::Exchange& x = c->xch;
    if (!x.connected) return c->fail("exchange: not connected");
    if (n_local < 0 || class_offset < 0 || n_total < 1 || class_offset + n_local > n_total || n_total > x.max_classes)
        return c->fail("exchange: shard [%d, %d) of %d classes does not fit max_classes %d", class_offset, class_offset + n_local,
                       n_total, x.max_classes);
    if (!all_codes_out_dev || (n_local > 0 && !raw_codes_dev)) return c->fail("exchange: null buffer");
    TRY(sylph_exchange_poll(c));
    const sylph_model_config& f = c->cfg;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ExchangeState* state = reinterpret_cast<ExchangeState*>(x.local);
    const float* local_codes = reinterpret_cast<const float*>(static_cast<uint8_t*>(x.local) + sizeof(ExchangeState));
    if (n_local > 0) {
        CU_TRY(c, launch_k(normalize_scatter_codes_kernel, dim3(n_local), dim3(256), 0, st, raw_codes_dev, x.peers,
                           static_cast<const ExchangeState*>(state), class_offset, x.max_classes, f.generator == 0 ? 1 : 0,
                           static_cast<const float*>(c->post_gn_w), static_cast<const float*>(c->post_gn_b), f.cg_post_norm,
                           f.cg_conv_l2_norm, c->conv_scale, c->bias_scale, c->bias_value));
        c->launches++;
    }
    const int collect_blocks = std::max(1, std::min(32, ceil_div(static_cast<long long>(n_total) * SYLPH_CODE_STRIDE, 4096)));
    CU_TRY(c, launch_k(collect_codes_kernel, dim3(collect_blocks), dim3(1024), 0, st, state, local_codes, x.max_classes, n_total,
                       all_codes_out_dev, x.timeout_ns));
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    // the sticky error word follows the rows to the host (asynchronous 4-byte copy): sylph_exchange_poll reads it without
    // touching the device, so a caller that has synchronised on the episode's results knows whether the codes were complete
    CU_TRY(c, cudaMemcpyAsync(x.host_err, &state->error, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    return 0;
}

int sylph_exchange_poll(sylph_ctx* c) {
    if (!c) return 1;
    if (!c->xch.local || !c->xch.host_err) return c->fail("exchange: not created");
    if (*static_cast<volatile unsigned int*>(c->xch.host_err) != 0u)
        return c->fail("exchange: a rank waited longer than %llu ms for the class codes of its peers and gave up (ranks out of "
                       "step, or SYLPH_EXCHANGE_TIMEOUT_MS too small): the codes of that episode were incomplete",
                       c->xch.timeout_ns / 1000000ull);
    return 0;
}

int sylph_exchange_status(sylph_ctx* c, int* timed_out, int64_t* rows_arrived) {
    if (!c) return 1;
    if (!c->xch.local) return c->fail("exchange: not created");
    ExchangeState h;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaMemcpy(&h, c->xch.local, sizeof(h), cudaMemcpyDeviceToHost));   // synchronous: drains the device first
    if (timed_out) *timed_out = static_cast<int>(h.error);
    if (rows_arrived) *rows_arrived = static_cast<int64_t>(h.arrived);
    return 0;
}

void sylph_exchange_destroy(sylph_ctx* c) {
    if (!c || !c->xch.local) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->xch.world; ++r)
        if (r != c->xch.rank && c->xch.peer_base[r]) cudaIpcCloseMemHandle(c->xch.peer_base[r]);
    cudaFree(c->xch.local);
    if (c->xch.host_err) cudaFreeHost(c->xch.host_err);
    c->xch = sylph_ctx::Exchange();
}

int sylph_accumulate_codes(sylph_ctx* c, const float* chunk_codes_dev, int n_chunks, const int* chunk_class_host,
                           const float* chunk_weight_host, float* acc_dev, int n_classes, void* stream) {
    if (!c) return 1;
    if (n_chunks <= 0) return 0;
    if (n_classes <= 0 || !chunk_codes_dev || !chunk_class_host || !chunk_weight_host || !acc_dev)
        return c->fail("sylph_accumulate_codes: null argument or no classes");
    for (int k = 0; k < n_chunks; ++k)
        if (chunk_class_host[k] < 0 || chunk_class_host[k] >= n_classes)
            return c->fail("chunk %d has class id %d outside [0, %d)", k, chunk_class_host[k], n_classes);
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    void *pc, *pw;
    TRY(ensure(c, "acc.chunk_class", static_cast<size_t>(n_chunks) * 4, "", &pc, st, false));
    TRY(ensure(c, "acc.chunk_weight", static_cast<size_t>(n_chunks) * 4, "", &pw, st, false));
    TRY(stage_h2d(c, pc, chunk_class_host, static_cast<size_t>(n_chunks) * 4, st));
    TRY(stage_h2d(c, pw, chunk_weight_host, static_cast<size_t>(n_chunks) * 4, st));
    CU_TRY(c, launch_k(accumulate_codes_kernel, dim3(n_classes), dim3(288), 0, st, chunk_codes_dev, static_cast<const int*>(pc),
                       static_cast<const float*>(pw), n_chunks, acc_dev));
    c->launches++;
    return 0;
}

int sylph_reduce_codes(sylph_ctx* c, const float* parts_dev, int n_parts, int n_classes, const float* divisor_host,
                       float* codes_out_dev, void* stream) {
    if (!c) return 1;
    if (n_classes <= 0) return 0;  // reduce_class_code returns an empty list unchanged
    if (n_parts <= 0 || !parts_dev || !divisor_host || !codes_out_dev) return c->fail("sylph_reduce_codes: bad arguments");
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    void* pd;
    TRY(ensure(c, "acc.divisor", static_cast<size_t>(n_classes) * 4, "", &pd, st, false));
    TRY(stage_h2d(c, pd, divisor_host, static_cast<size_t>(n_classes) * 4, st));
    CU_TRY(c, launch_k(reduce_codes_kernel, dim3(n_classes), dim3(288), 0, st, parts_dev, n_parts, n_classes,
                       static_cast<const float*>(pd), codes_out_dev));
    c->launches++;
    return 0;
}

int sylph_detect(sylph_ctx* c, int slot, const float* codes_dev, int n_classes, const int* out_sizes_host,
                 float* dets_out_dev, int* counts_out_dev, int max_dets, void* stream) {
    return sylph_detect_after(c, slot, codes_dev, n_classes, out_sizes_host, dets_out_dev, counts_out_dev, max_dets, nullptr, stream);
}

int sylph_detect_after(sylph_ctx* c, int slot, const float* codes_dev, int n_classes, const int* out_sizes_host,
                       float* dets_out_dev, int* counts_out_dev, int max_dets, void* codes_ready_event, void* stream) {
    if (!c) return 1;
    if (!c->finalized) return c->fail("weights not finalized");
    if (slot < 0 || slot >= SYLPH_NUM_SLOTS || !c->slots[slot].valid) return c->fail("slot %d holds no features", slot);
    if (n_classes <= 0) return c->fail("no class codes");
    const sylph_model_config& f = c->cfg;
    if (max_dets < f.post_nms_topk || max_dets > 1024) return c->fail("max_dets must be in [post_nms_topk, 1024]");
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Slot& S = c->slots[slot];
    const int n_segs = 5 * S.n;
    HeadOut H{};
    TRY(run_head(c, slot, codes_dev, n_classes, st, &H, static_cast<cudaEvent_t>(codes_ready_event)));
    float* lg = H.logits;
    float* pr = H.pred;
    struct { int cout_pad; } CW{H.logit_stride};
    // ---- proposals
    DetectParams P{};
    P.pg = S.pg;
    P.n_images = S.n;
    P.n_classes = n_classes;
    P.logit_stride = CW.cout_pad;
    long long off = 0;
    for (int l = 0; l < 5; ++l) {
        const long long full = static_cast<long long>(S.lh[l]) * S.lw[l] * n_classes;
        // the NMS key packs (level << 28 | location * classes + class): larger planes / class counts would corrupt it silently
        if (full >= (1LL << 28))
            return c->fail("level %d: %d x %d locations x %d classes = %lld candidate slots exceed the 2^28 key space of the "
                           "proposal kernels (split the class list over several detect calls)", l, S.lh[l], S.lw[l], n_classes, full);
        P.cap[l] = static_cast<int>(std::min<long long>(full, 1 << 22));
        P.cand_off[l] = off;
        off += static_cast<long long>(S.n) * P.cap[l];
        P.level_scale[l] = c->level_scale[l];
        P.stride[l] = 8 << l;
    }
    P.thresh = f.inference_thresh;
    P.thresh_with_ctr = f.thresh_with_ctr;
    P.box_quality = f.box_quality;
    P.pre_topk = f.pre_nms_topk;
    P.post_topk = f.post_nms_topk;
    P.nms_thresh = f.nms_thresh;
    void *cand, *cnt, *sel, *selc, *ia;
    TRY(ensure(c, "det.cand", static_cast<size_t>(off) * 8, "", &cand, st, false));
    TRY(ensure(c, "det.counts", static_cast<size_t>(n_segs + 1) * 4, "", &cnt, st, false));
    TRY(ensure(c, "det.sel", static_cast<size_t>(n_segs) * P.pre_topk * 8, "", &sel, st, false));
    TRY(ensure(c, "det.sel_count", static_cast<size_t>(n_segs) * 4, "", &selc, st, false));
    TRY(ensure(c, "det.img_args", static_cast<size_t>(S.n) * sizeof(NmsImageArgs), "", &ia, st, false));
    std::vector<NmsImageArgs> args(S.n);
    for (int i = 0; i < S.n; ++i) {
        const int oh = out_sizes_host ? out_sizes_host[2 * i] : S.img_h[i];
        const int ow = out_sizes_host ? out_sizes_host[2 * i + 1] : S.img_w[i];
        args[i].scale_x = static_cast<float>(static_cast<double>(ow) / S.img_w[i]);
        args[i].scale_y = static_cast<float>(static_cast<double>(oh) / S.img_h[i]);
        args[i].out_w = static_cast<float>(ow);
        args[i].out_h = static_cast<float>(oh);
    }
    TRY(stage_h2d(c, ia, args.data(), S.n * sizeof(NmsImageArgs), st));
    CU_TRY(c, cudaMemsetAsync(cnt, 0, static_cast<size_t>(n_segs + 1) * 4, st));
    {
        StageTimer t(c, "proposals", st, static_cast<double>(S.n) * 22400 * (CW.cout_pad + 16) * 4);
        long long total_px = 0;
        for (int l = 0; l < 5; ++l) total_px += static_cast<long long>(S.n) * S.lh[l] * S.lw[l];
        CU_TRY(c, launch_k(fcos_candidates_kernel, dim3(grid_for(total_px, 256, c->num_sms)), dim3(256), 0, st, 
            static_cast<const float*>(lg), static_cast<const float*>(pr), P, static_cast<unsigned long long*>(cand),
            static_cast<int*>(cnt), static_cast<int*>(cnt) + n_segs));
        CU_TRY(c, cudaGetLastError());
        CU_TRY(c, launch_k(fcos_select_kernel, dim3(n_segs), dim3(1024), 0, st, static_cast<const unsigned long long*>(cand), static_cast<const int*>(cnt), P,
                                                    static_cast<unsigned long long*>(sel), static_cast<int*>(selc)));
        CU_TRY(c, cudaGetLastError());
        const int n_max = 5 * P.pre_topk;
        int sort_n = 32;
        while (sort_n < n_max) sort_n <<= 1;
        const size_t smem = static_cast<size_t>(sort_n) * 8 + static_cast<size_t>(n_max) * 21 + 16;
        static PerDeviceOnce once;
        CU_TRY(c, once.run([] { return cudaFuncSetAttribute(fcos_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); }));
        if (smem > 200 * 1024) return c->fail("NMS shared memory budget exceeded (pre_nms_topk too large)");
        CU_TRY(c, launch_k(fcos_nms_kernel, dim3(S.n), dim3(1024), smem, st, static_cast<const unsigned long long*>(sel), static_cast<const int*>(selc),
                                                 static_cast<const float*>(pr), P, static_cast<const NmsImageArgs*>(ia), sort_n,
                                                 n_max, dets_out_dev, counts_out_dev, max_dets));
        CU_TRY(c, cudaGetLastError());
        c->launches += 3;
        // the overflow word (an (image, level) list had more candidates above the threshold than its 2^22 slots) follows the
        // detections to pinned host memory: sylph_detect_poll reports it without touching the device.  Not while capturing
        // a CUDA graph (the pinned word is allocated lazily, and allocation is illegal during capture).
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        const bool capturing = cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusActive;
        if (!capturing) {
            if (c->detect_overflow_host == nullptr) {
                CU_TRY(c, cudaHostAlloc(reinterpret_cast<void**>(&c->detect_overflow_host), 2 * sizeof(unsigned int), cudaHostAllocDefault));
                c->detect_overflow_host[0] = c->detect_overflow_host[1] = 0u;
            }
            CU_TRY(c, cudaMemcpyAsync(c->detect_overflow_host, static_cast<int*>(cnt) + n_segs, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        }
    }
    return 0;
}

int sylph_detect_poll(sylph_ctx* c) {
    if (!c) return 1;
    if (c->detect_overflow_host == nullptr) return 0;
    if (*static_cast<volatile unsigned int*>(c->detect_overflow_host) != 0u) {
        c->detect_overflow_host[1] = 1u;
        c->detect_overflow_host[0] = 0u;
        return c->fail("a detect call dropped candidates: more than 2^22 (location, class) pairs of one (image, level) passed "
                       "INFERENCE_TH_TEST (lower the class count per call or raise the threshold)");
    }
    return 0;
}

int sylph_export_head_output(sylph_ctx* c, int which, int level, float* out_dev, void* stream) {
    if (!c) return 1;
    if (c->last_detect_slot < 0 || level < 0 || level > 4 || which < 0 || which > 3) return c->fail("no detect call to export from");
    const Slot& S = c->slots[c->last_detect_slot];
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float* src = static_cast<const float*>(c->bufs[which == 0 ? "det.logits" : "det.pred"].p);
    int C = 1, cstride = 16, coff = 0, relu = 0;
    float scale = 1.f;
    if (which == 0) { C = c->last_detect_classes; cstride = c->last_logit_stride; }
    else if (which == 1) { C = 4; scale = c->level_scale[level]; relu = 1; }
    else if (which == 2) { coff = 4; }
    else { coff = 5; }
    const long long work = static_cast<long long>(S.n) * C * S.lh[level] * S.lw[level];
    CU_TRY(c, launch_k(export_nchw_kernel<float>, dim3(grid_for(work, 256, c->num_sms)), dim3(256), 0, st, src, out_dev, S.pg.lv[level], S.n, C, cstride, coff, scale, relu, 0));
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return 0;
}

int sylph_fcos_loss_sums(sylph_ctx* c, int slot, const float* codes_dev, int n_classes, const int64_t* support_targets_host,
                         const sylph_loss_config* lc, int n_gt, const float* gt_boxes_host, const int64_t* gt_classes_host,
                         const int* gt_offsets_host, double* sums_out_dev, int64_t* labels_out_dev,
                         int64_t* target_inds_out_dev, float* reg_targets_out_dev, void* stream) {
    if (!c) return 1;
    if (!c->finalized) return c->fail("weights not finalized");
    if (slot < 0 || slot >= SYLPH_NUM_SLOTS || !c->slots[slot].valid) return c->fail("slot %d holds no features", slot);
    if (n_classes <= 0 || !support_targets_host) return c->fail("no class codes / support_set_targets");
    if (!lc || !gt_offsets_host || !sums_out_dev) return c->fail("null argument");
    if (lc->loc_loss_type < 0 || lc->loc_loss_type > 2) return c->fail("LOC_LOSS_TYPE must be iou | linear_iou | giou");
    if (n_gt < 0 || (n_gt > 0 && (!gt_boxes_host || !gt_classes_host))) return c->fail("ground-truth arrays missing");
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Slot& S = c->slots[slot];
    if (gt_offsets_host[0] != 0 || gt_offsets_host[S.n] != n_gt) return c->fail("gt_offsets must run from 0 to n_gt over %d images", S.n);
    for (int i = 0; i < S.n; ++i)
        if (gt_offsets_host[i + 1] < gt_offsets_host[i]) return c->fail("gt_offsets must be non-decreasing");
    HeadOut H{};
    TRY(run_head(c, slot, codes_dev, n_classes, st, &H, nullptr, c->loss_box_branch));
    LossParams P{};
    P.pg = S.pg;
    P.n_images = S.n;
    P.n_classes = n_classes;
    P.logit_stride = H.logit_stride;
    long long total_px = 0;
    for (int l = 0; l < 5; ++l) {
        P.stride[l] = 8 << l;
        P.level_scale[l] = c->level_scale[l];
        P.soi_lo[l] = l == 0 ? -1.f : static_cast<float>(lc->sizes_of_interest[l - 1]);
        P.soi_hi[l] = l == 4 ? kFcosInf : static_cast<float>(lc->sizes_of_interest[l]);
        P.radius[l] = static_cast<float>(static_cast<double>(P.stride[l]) * static_cast<double>(lc->pos_radius));
        total_px += static_cast<long long>(S.n) * S.lh[l] * S.lw[l];
    }
    P.center_sample = lc->center_sample;
    P.alpha = lc->focal_alpha;
    P.gamma = lc->focal_gamma;
    P.loc_loss_type = lc->loc_loss_type;
    const int blocks = grid_for(total_px, 256, c->num_sms);
    void *gb, *gc, *go, *stg, *part;
    TRY(ensure(c, "loss.gt_boxes", static_cast<size_t>(std::max(n_gt, 1)) * 16, "", &gb, st, false));
    TRY(ensure(c, "loss.gt_classes", static_cast<size_t>(std::max(n_gt, 1)) * 8, "", &gc, st, false));
    TRY(ensure(c, "loss.gt_offsets", static_cast<size_t>(S.n + 1) * 4, "", &go, st, false));
    TRY(ensure(c, "loss.support_targets", static_cast<size_t>(n_classes) * 8, "", &stg, st, false));
    TRY(ensure(c, "loss.partials", static_cast<size_t>(blocks) * kLossSums * 8, "", &part, st, false));
    TRY(stage_h2d(c, gb, gt_boxes_host, static_cast<size_t>(n_gt) * 16, st));
    TRY(stage_h2d(c, gc, gt_classes_host, static_cast<size_t>(n_gt) * 8, st));
    TRY(stage_h2d(c, go, gt_offsets_host, static_cast<size_t>(S.n + 1) * 4, st));
    TRY(stage_h2d(c, stg, support_targets_host, static_cast<size_t>(n_classes) * 8, st));
    {
        StageTimer t(c, "loss.targets+sums", st, static_cast<double>(total_px) * (H.logit_stride + 16) * 4);
        CU_TRY(c, launch_k(fcos_targets_loss_kernel, dim3(blocks), dim3(256), 0, st, static_cast<const float*>(H.logits),
                           static_cast<const float*>(H.pred), P, static_cast<const float*>(gb), static_cast<const long long*>(gc),
                           static_cast<const int*>(go), static_cast<const long long*>(stg), static_cast<double*>(part),
                           reinterpret_cast<long long*>(labels_out_dev), reinterpret_cast<long long*>(target_inds_out_dev),
                           reg_targets_out_dev));
        CU_TRY(c, cudaGetLastError());
        CU_TRY(c, launch_k(fcos_loss_reduce_kernel, dim3(1), dim3(256), 0, st, static_cast<const double*>(part), blocks, sums_out_dev));
        CU_TRY(c, cudaGetLastError());
        c->launches += 2;
    }
    return 0;
}

int sylph_fcos_loss_finalize(sylph_ctx* c, const double* local_sums_dev, const double* global_pos_ctr_dev, int world_size,
                             float* losses_out_dev, void* stream) {
    if (!c) return 1;
    if (!local_sums_dev || !losses_out_dev) return c->fail("null argument");
    if (world_size < 1) return c->fail("world_size must be >= 1");
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CU_TRY(c, launch_k(fcos_loss_finalize_kernel, dim3(1), dim3(32), 0, st, local_sums_dev,
                       global_pos_ctr_dev ? global_pos_ctr_dev : local_sums_dev + 1, world_size, losses_out_dev));
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return 0;
}

// ------------------------------------------------------------------------------------------------ training backward
int sylph_fcos_cls_loss_backward(sylph_ctx* c, int slot, int n_classes, const int64_t* support_targets_host,
                                 const sylph_loss_config* lc, const int64_t* labels_dev, const double* local_sums_dev,
                                 const double* global_pos_ctr_dev, int world_size, const float* grad_loss_dev,
                                 float* grad_codes_out_dev, void* stream) {
    if (!c) return 1;
    if (!c->finalized) return c->fail("weights not finalized");
    if (slot < 0 || slot >= SYLPH_NUM_SLOTS || !c->slots[slot].valid) return c->fail("slot %d holds no features", slot);
    if (!support_targets_host || !lc || !labels_dev || !local_sums_dev || !grad_codes_out_dev) return c->fail("null argument");
    if (world_size < 1) return c->fail("world_size must be >= 1");
    if (c->last_detect_slot != slot || c->last_detect_classes != n_classes || c->last_cls_tower == nullptr)
        return c->fail("sylph_fcos_cls_loss_backward: the last head pass was not sylph_fcos_loss_sums on slot %d with %d classes", slot, n_classes);
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Slot& S = c->slots[slot];
    auto lg = c->bufs.find("det.logits");
    if (lg == c->bufs.end() || !lg->second.p) return c->fail("no logits buffer");
    void *stg, *part;
    long long total_px = 0;
    for (int l = 0; l < 5; ++l) total_px += static_cast<long long>(S.n) * S.lh[l] * S.lw[l];
    const int blocks = static_cast<int>(std::min<long long>(8LL * c->num_sms, (total_px + kClsBwdRows - 1) / kClsBwdRows));
    TRY(ensure(c, "bwd.support_targets", static_cast<size_t>(n_classes) * 8, "", &stg, st, false));
    TRY(ensure(c, "bwd.code_partials", static_cast<size_t>(blocks) * n_classes * 257 * 4, "", &part, st, false));
    TRY(stage_h2d(c, stg, support_targets_host, static_cast<size_t>(n_classes) * 8, st));
    {
        StageTimer t(c, "bwd.cls_loss_codes", st, static_cast<double>(total_px) * (c->ld(256) * 2 + c->last_logit_stride * 4));
        CU_TRY(c, launch_k(fcos_cls_loss_bwd_kernel, dim3(blocks), dim3(256), 0, st, static_cast<const float*>(lg->second.p),
                           c->last_logit_stride, c->last_cls_tower, c->ld(256), c->lo(256), S.pg, S.n,
                           reinterpret_cast<const long long*>(labels_dev), static_cast<const long long*>(stg), n_classes,
                           lc->focal_alpha, lc->focal_gamma, global_pos_ctr_dev ? global_pos_ctr_dev : local_sums_dev + 1,
                           world_size, grad_loss_dev, static_cast<float*>(part)));
        CU_TRY(c, cudaGetLastError());
        const int n_elems = n_classes * 257;
        CU_TRY(c, launch_k(fcos_code_grad_reduce_kernel, dim3(ceil_div(static_cast<long long>(n_elems) * 32, 256)), dim3(256), 0, st,
                           static_cast<const float*>(part), blocks, n_elems, grad_codes_out_dev));
        CU_TRY(c, cudaGetLastError());
        c->launches += 2;
    }
    return 0;
}

static int sgemm(sylph_ctx* c, cudaStream_t st, const float* A, long long sam, long long sak, const float* B, long long sbk,
                 long long sbn, float* C, long long ldc, int M, int N, int K, const float* bias, int accumulate) {
    // 64 x 64 tiles when they fill the machine, 32 x 32 tiles otherwise (the forward re-evaluation of 15 ROIs is 12 x 4 large tiles)
    if (static_cast<long long>(ceil_div(N, 64)) * ceil_div(M, 64) >= 2LL * c->num_sms)
        CU_TRY(c, launch_k(sgemm_f32_kernel<4, 4>, dim3(ceil_div(N, 64), ceil_div(M, 64)), dim3(256), 0, st, A, sam, sak, B, sbk, sbn,
                           C, ldc, M, N, K, bias, accumulate));
    else
        CU_TRY(c, launch_k(sgemm_f32_kernel<2, 2>, dim3(ceil_div(N, 32), ceil_div(M, 32)), dim3(256), 0, st, A, sam, sak, B, sbk, sbn,
                           C, ldc, M, N, K, bias, accumulate));
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return 0;
}

int sylph_codegen_backward(sylph_ctx* c, int n_rois, int n_classes, const int* class_offsets_host, const float* raw_codes_dev,
                           const float* grad_codes_dev, const sylph_codegen_tensors* params, const sylph_codegen_tensors* grads,
                           void* stream) {
    if (!c) return 1;
    if (!c->finalized) return c->fail("weights not finalized");
    const sylph_model_config& f = c->cfg;
    if (f.generator != 0) return c->fail("sylph_codegen_backward: only the CodeGenerator plugin is trainable on this path");
    if (f.cg_weight_layer) return c->fail("sylph_codegen_backward: CODE_GENERATOR.WEIGHT_LAYER is not differentiated on this path");
    if (!class_offsets_host || !raw_codes_dev || !grad_codes_dev || !params || !grads) return c->fail("null argument");
    if (n_rois <= 0 || n_rois != c->last_n_rois)
        return c->fail("sylph_codegen_backward: %d ROIs, but the last sylph_generate_codes call pooled %d", n_rois, c->last_n_rois);
    if (n_rois > 4096) return c->fail("sylph_codegen_backward: at most 4096 support ROIs per call");
    if (f.cg_tower_layers > SYLPH_CG_MAX_TOWER) return c->fail("more than %d tower layers", SYLPH_CG_MAX_TOWER);
    if (class_offsets_host[0] != 0 || class_offsets_host[n_classes] != n_rois) return c->fail("class_offsets must span all ROIs");
    const int L = f.cg_tower_layers;
    for (int i = 0; i < L; ++i)
        if (!params->tower_w[i] || !params->tower_b[i] || !params->tower_gn_w[i] || !params->tower_gn_b[i] ||
            !grads->tower_w[i] || !grads->tower_b[i] || !grads->tower_gn_w[i] || !grads->tower_gn_b[i])
            return c->fail("tower layer %d: missing parameter / gradient tensor", i);
    if (!params->cls_w || !params->cls_b || !grads->cls_w || !grads->cls_b) return c->fail("support_set_cls_conv tensors missing");
    if (f.cg_bias_layer && (!params->bias_w || !params->bias_b || !grads->bias_w || !grads->bias_b))
        return c->fail("support_set_cls_bias tensors missing");
    if (f.cg_post_norm && (!params->post_norm_w || !params->post_norm_b || !grads->post_norm_w || !grads->post_norm_b))
        return c->fail("post_norm tensors missing");
    if (f.cg_has_conv_scale && (!params->conv_scale || !grads->conv_scale)) return c->fail("conv_scale tensors missing");
    if (f.cg_bias_layer && (!params->bias_scale || !grads->bias_scale)) return c->fail("bias_scale tensors missing");
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::vector<int> roi_class(n_rois);
    for (int k = 0; k < n_classes; ++k) {
        if (class_offsets_host[k + 1] <= class_offsets_host[k]) return c->fail("class %d has no support ROI", k);
        for (int r = class_offsets_host[k]; r < class_offsets_host[k + 1]; ++r) roi_class[r] = k;
    }
    const int P = n_rois * 49;
    const size_t act = static_cast<size_t>(P) * 256 * 4;
    void *xbuf, *ybuf, *stats, *col, *dcol, *da, *db, *yc, *vb, *dv, *draw, *parts, *scal, *prc, *pco;
    TRY(ensure(c, "bwd.x", act * (L + 1), "", &xbuf, st, false));              // inputs of every layer (X_0 .. X_L)
    TRY(ensure(c, "bwd.y", act * std::max(L, 1), "", &ybuf, st, false));       // convolution outputs in front of the GroupNorms
    TRY(ensure(c, "bwd.stats", static_cast<size_t>(std::max(L, 1)) * 2 * n_rois * 256 * 4, "", &stats, st, false));
    TRY(ensure(c, "bwd.col", static_cast<size_t>(P) * 2304 * 4, "", &col, st, false));
    TRY(ensure(c, "bwd.dcol", static_cast<size_t>(P) * 2304 * 4, "", &dcol, st, false));
    TRY(ensure(c, "bwd.da", act, "", &da, st, false));
    TRY(ensure(c, "bwd.db", act, "", &db, st, false));
    TRY(ensure(c, "bwd.yc", act, "", &yc, st, false));
    TRY(ensure(c, "bwd.vb", static_cast<size_t>(P) * 4, "", &vb, st, false));
    TRY(ensure(c, "bwd.dv", static_cast<size_t>(P) * 4, "", &dv, st, false));
    TRY(ensure(c, "bwd.draw", static_cast<size_t>(n_classes) * 257 * 4, "", &draw, st, false));
    TRY(ensure(c, "bwd.parts", static_cast<size_t>(2) * n_rois * 256 * 4, "", &parts, st, false));
    TRY(ensure(c, "bwd.scalars", 64, "", &scal, st, false));
    TRY(ensure(c, "bwd.roi_class", static_cast<size_t>(n_rois) * 4, "", &prc, st, false));
    TRY(ensure(c, "bwd.class_off", static_cast<size_t>(n_classes + 1) * 4, "", &pco, st, false));
    TRY(stage_h2d(c, prc, roi_class.data(), static_cast<size_t>(n_rois) * 4, st));
    TRY(stage_h2d(c, pco, class_offsets_host, static_cast<size_t>(n_classes + 1) * 4, st));
    auto r0 = c->bufs.find("cg.r0");
    if (r0 == c->bufs.end() || !r0->second.p) return c->fail("no ROI planes");
    auto X = [&](int i) { return static_cast<float*>(xbuf) + static_cast<size_t>(i) * P * 256; };
    auto Y = [&](int i) { return static_cast<float*>(ybuf) + static_cast<size_t>(i) * P * 256; };
    auto MEAN = [&](int i) { return static_cast<float*>(stats) + static_cast<size_t>(2 * i) * n_rois * 256; };
    auto RSTD = [&](int i) { return static_cast<float*>(stats) + static_cast<size_t>(2 * i + 1) * n_rois * 256; };
    float* colp = static_cast<float*>(col);
    float* dcolp = static_cast<float*>(dcol);
    const int g_act = grid_for(static_cast<long long>(P) * 256, 256, c->num_sms);
    const int g_col = grid_for(static_cast<long long>(P) * 2304, 256, c->num_sms);
    StageTimer t(c, "bwd.code_generator", st, 0.0);
    // ---- forward, fp32 (code_generator.py:941-967)
    CU_TRY(c, launch_k(roi_planes_to_f32_kernel, dim3(g_act), dim3(256), 0, st, static_cast<const __half*>(r0->second.p), X(0), n_rois,
                       c->split, c->roi_stride));
    c->launches++;
    for (int i = 0; i < L; ++i) {
        CU_TRY(c, launch_k(im2col_roi_kernel, dim3(g_col), dim3(256), 0, st, static_cast<const float*>(X(i)), colp, n_rois));
        c->launches++;
        TRY(sgemm(c, st, colp, 2304, 1, params->tower_w[i], 1, 2304, Y(i), 256, P, 256, 2304, params->tower_b[i], 0));
        CU_TRY(c, launch_k(roi_gn_relu_fwd_f32_kernel, dim3(n_rois), dim3(256), 0, st, static_cast<const float*>(Y(i)),
                           static_cast<const float*>(params->tower_gn_w[i]), static_cast<const float*>(params->tower_gn_b[i]), X(i + 1),
                           MEAN(i), RSTD(i)));
        c->launches++;
    }
    CU_TRY(c, launch_k(im2col_roi_kernel, dim3(g_col), dim3(256), 0, st, static_cast<const float*>(X(L)), colp, n_rois));
    c->launches++;
    if (f.cg_bias_layer)   // v = support_set_cls_bias(x): one output channel
        TRY(sgemm(c, st, colp, 2304, 1, params->bias_w, 1, 2304, static_cast<float*>(vb), 1, P, 1, 2304, params->bias_b, 0));
    // ---- code processing, class mean, pools (code_generator.py:778-875)
    CU_TRY(c, launch_k(normalize_codes_bwd_kernel, dim3(1), dim3(256), 0, st, raw_codes_dev, grad_codes_dev, n_classes,
                       static_cast<const float*>(params->post_norm_w), static_cast<const float*>(params->post_norm_b), f.cg_post_norm,
                       f.cg_conv_l2_norm, static_cast<const float*>(f.cg_has_conv_scale ? params->conv_scale : nullptr),
                       static_cast<const float*>(f.cg_bias_layer ? params->bias_scale : nullptr), static_cast<float*>(draw),
                       f.cg_post_norm ? grads->post_norm_w : nullptr, f.cg_post_norm ? grads->post_norm_b : nullptr,
                       static_cast<float*>(scal)));
    c->launches++;
    if (f.cg_has_conv_scale) CU_TRY(c, cudaMemcpyAsync(grads->conv_scale, scal, 4, cudaMemcpyDeviceToDevice, st));
    if (f.cg_bias_layer) CU_TRY(c, cudaMemcpyAsync(grads->bias_scale, static_cast<float*>(scal) + 1, 4, cudaMemcpyDeviceToDevice, st));
    float* dyc = static_cast<float*>(da);
    CU_TRY(c, launch_k(shot_code_bwd_kernel, dim3(n_rois), dim3(256), 0, st, static_cast<const float*>(draw), static_cast<const int*>(prc),
                       static_cast<const int*>(pco), static_cast<const float*>(vb), f.cg_bias_layer, f.cg_bias_l2_norm, dyc,
                       static_cast<float*>(dv)));
    c->launches++;
    // ---- support_set_cls_conv / support_set_cls_bias: weight, bias and input gradients
    auto colsum = [&](const float* in, int ld_in, int rows, int cols, float* out) -> int {
        CU_TRY(c, launch_k(colsum_f32_kernel, dim3(ceil_div(cols, 32)), dim3(256), 0, st, in, static_cast<long long>(ld_in), rows, cols, out));
        c->launches++;
        return 0;
    };
    TRY(sgemm(c, st, dyc, 1, 256, colp, 2304, 1, grads->cls_w, 2304, 256, 2304, P, nullptr, 0));          // dW = dY^T col
    TRY(colsum(dyc, 256, P, 256, grads->cls_b));
    if (L > 0) TRY(sgemm(c, st, dyc, 256, 1, params->cls_w, 2304, 1, dcolp, 2304, P, 2304, 256, nullptr, 0));  // dcol = dY W
    if (f.cg_bias_layer) {
        const float* dvp = static_cast<const float*>(dv);
        TRY(sgemm(c, st, dvp, 1, 1, colp, 2304, 1, grads->bias_w, 2304, 1, 2304, P, nullptr, 0));
        TRY(colsum(dvp, 1, P, 1, grads->bias_b));
        if (L > 0) TRY(sgemm(c, st, dvp, 1, 1, params->bias_w, 2304, 1, dcolp, 2304, P, 2304, 1, nullptr, 1));
    }
    // ---- support_set_shared_tower, last layer first
    float* dx = static_cast<float*>(db);
    const char* dbg = getenv("SYLPH_BWD_DEBUG_STOP");   // debugging aid: return with the intermediates of this layer in place
    const int dbg_stop = dbg ? atoi(dbg) : -1;
    for (int i = L - 1; i >= 0; --i) {
        CU_TRY(c, launch_k(col2im_roi_kernel, dim3(g_act), dim3(256), 0, st, static_cast<const float*>(dcolp), dx, n_rois));
        c->launches++;
        if (dbg_stop == 10 + i) return 0;
        float* dgp = static_cast<float*>(parts);
        float* dbp = dgp + static_cast<size_t>(n_rois) * 256;
        CU_TRY(c, launch_k(roi_gn_relu_bwd_f32_kernel, dim3(n_rois), dim3(256), 0, st, static_cast<const float*>(dx),
                           static_cast<const float*>(Y(i)), static_cast<const float*>(MEAN(i)), static_cast<const float*>(RSTD(i)),
                           static_cast<const float*>(params->tower_gn_w[i]), static_cast<const float*>(params->tower_gn_b[i]), dx, dgp, dbp));
        c->launches++;
        if (dbg_stop == i) return 0;
        TRY(colsum(dgp, 256, n_rois, 256, grads->tower_gn_w[i]));
        TRY(colsum(dbp, 256, n_rois, 256, grads->tower_gn_b[i]));
        CU_TRY(c, launch_k(im2col_roi_kernel, dim3(g_col), dim3(256), 0, st, static_cast<const float*>(X(i)), colp, n_rois));
        c->launches++;
        TRY(sgemm(c, st, dx, 1, 256, colp, 2304, 1, grads->tower_w[i], 2304, 256, 2304, P, nullptr, 0));
        TRY(colsum(dx, 256, P, 256, grads->tower_b[i]));
        if (i > 0) TRY(sgemm(c, st, dx, 256, 1, params->tower_w[i], 2304, 1, dcolp, 2304, P, 2304, 256, nullptr, 0));
    }
    CU_TRY(c, cudaGetLastError());
    return 0;
}

int sylph_update_code_generator_device(sylph_ctx* c, const sylph_codegen_tensors* params, void* stream) {
    if (!c) return 1;
    if (!c->weights_ready || !c->finalized) return c->fail("sylph_update_code_generator_device: call sylph_finalize_weights once first");
    const sylph_model_config& f = c->cfg;
    if (f.generator != 0) return c->fail("sylph_update_code_generator_device: only the CodeGenerator plugin is trainable on this path");
    if (f.cg_weight_layer) return c->fail("sylph_update_code_generator_device: CODE_GENERATOR.WEIGHT_LAYER is not covered");
    if (!params) return c->fail("null argument");
    const int L = f.cg_tower_layers;
    if (L > SYLPH_CG_MAX_TOWER) return c->fail("more than %d tower layers", SYLPH_CG_MAX_TOWER);
    for (int i = 0; i < L; ++i)
        if (!params->tower_w[i] || !params->tower_b[i] || !params->tower_gn_w[i] || !params->tower_gn_b[i])
            return c->fail("tower layer %d: missing parameter tensor", i);
    if (!params->cls_w || !params->cls_b) return c->fail("support_set_cls_conv tensors missing");
    if (f.cg_bias_layer && (!params->bias_w || !params->bias_b || !params->bias_scale)) return c->fail("support_set_cls_bias / bias_scale tensors missing");
    if (f.cg_post_norm && (!params->post_norm_w || !params->post_norm_b)) return c->fail("post_norm tensors missing");
    if (f.cg_has_conv_scale && !params->conv_scale) return c->fail("conv_scale tensor missing");
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto pack = [&](const float* w, const float* b, ConvW& W, int pooled) -> int {
        if (W.w_nm != nullptr) return c->fail("N-merged weight copies are not refreshed on the device");
        const int ci = pooled ? W.cin / 9 : W.cin;
        CU_TRY(c, launch_k(pack_oihw_weights_kernel, dim3(grid_for(static_cast<long long>(W.cout) * ci * 9, 256, c->num_sms)), dim3(256), 0, st,
                           w, W.w, W.cout, ci, 9, W.cout_pad, c->split, pooled));
        CU_TRY(c, cudaMemcpyAsync(W.bias, b, static_cast<size_t>(W.cout) * 4, cudaMemcpyDeviceToDevice, st));
        c->launches++;
        return 0;
    };
    for (int i = 0; i < L; ++i) {
        if (c->cg_tower[i].taps != 9 || c->cg_tower[i].cin != 256) return c->fail("tower layer %d is not a 3x3 256-channel convolution", i);
        TRY(pack(params->tower_w[i], params->tower_b[i], c->cg_tower[i], 0));
        CU_TRY(c, cudaMemcpyAsync(c->cg_gn_w[i], params->tower_gn_w[i], 1024, cudaMemcpyDeviceToDevice, st));
        CU_TRY(c, cudaMemcpyAsync(c->cg_gn_b[i], params->tower_gn_b[i], 1024, cudaMemcpyDeviceToDevice, st));
    }
    TRY(pack(params->cls_w, params->cls_b, c->cg_cls, 0));
    TRY(pack(params->cls_w, params->cls_b, c->cg_cls_pooled, 1));
    if (f.cg_bias_layer) {
        CU_TRY(c, launch_k(transpose_taps_kernel, dim3(9), dim3(256), 0, st, static_cast<const float*>(params->bias_w), c->cg_wbias, 256, 9));
        c->launches++;
        CU_TRY(c, cudaMemcpyAsync(c->cg_bbias, params->bias_b, 4, cudaMemcpyDeviceToDevice, st));
    }
    if (f.cg_post_norm) {
        CU_TRY(c, cudaMemcpyAsync(c->post_gn_w, params->post_norm_w, 1024, cudaMemcpyDeviceToDevice, st));
        CU_TRY(c, cudaMemcpyAsync(c->post_gn_b, params->post_norm_b, 1024, cudaMemcpyDeviceToDevice, st));
    }
    // the two scalars are kernel arguments of the normalisation kernels: 8 bytes back to the host (waits for the stream)
    float cs = 1.f, bs = 1.f;
    if (f.cg_has_conv_scale) CU_TRY(c, cudaMemcpyAsync(&cs, params->conv_scale, 4, cudaMemcpyDeviceToHost, st));
    if (f.cg_bias_layer) CU_TRY(c, cudaMemcpyAsync(&bs, params->bias_scale, 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(c, cudaStreamSynchronize(st));
    c->conv_scale = cs;
    c->bias_scale = bs;
    return 0;
}

// ------------------------------------------------------------------------------------------------ class-tower training
int sylph_set_loss_box_branch(sylph_ctx* c, int enabled) {
    if (!c) return 1;
    c->loss_box_branch = enabled != 0;
    return 0;
}

int sylph_set_training(sylph_ctx* c, int enabled) {
    if (!c) return 1;
    c->train_save = enabled != 0;
    if (!c->train_save) { c->saved_x.clear(); c->saved_raw.clear(); c->saved_stats.clear(); }
    return 0;
}

// Transposed, tap-reversed copies of the class-tower weights for the input-gradient convolutions (device-side packing).
static int prep_cls_tower_transposed(sylph_ctx* c, const sylph_tower_tensors* params, cudaStream_t st) {
    const int L = static_cast<int>(c->cls_tower.size());
    if (static_cast<int>(c->cls_tower_t.size()) != L) c->cls_tower_t.assign(L, ConvW());
    for (int i = 0; i < L; ++i) {
        const ConvW& F = c->cls_tower[i];
        if (F.taps != 9 || F.cin != 256 || F.cout != 256 || F.w_nm != nullptr) return c->fail("class-tower layer %d is not a 3x3 256 -> 256 convolution", i);
        ConvW& T = c->cls_tower_t[i];
        if (T.w == nullptr) {
            T = F;
            T.w = nullptr; T.bias = nullptr; T.w_nm = nullptr;
            CU_TRY(c, cudaMalloc(&T.w, static_cast<size_t>(F.b_rows) * F.k_per_tap * 2));
            CU_TRY(c, cudaMalloc(&T.bias, static_cast<size_t>(F.cout_pad) * 4));
            CU_TRY(c, cudaMemsetAsync(T.bias, 0, static_cast<size_t>(F.cout_pad) * 4, st));
        }
        CU_TRY(c, launch_k(pack_oihw_weights_transposed_kernel, dim3(grid_for(256 * 256 * 9, 256, c->num_sms)), dim3(256), 0, st,
                           static_cast<const float*>(params->conv_w[i]), T.w, c->split));
        c->launches++;
    }
    c->cls_tower_t_ready = true;
    return 0;
}

int sylph_update_cls_tower_device(sylph_ctx* c, const sylph_tower_tensors* params, void* stream) {
    if (!c) return 1;
    if (!c->weights_ready || !c->finalized) return c->fail("sylph_update_cls_tower_device: call sylph_finalize_weights once first");
    if (!params) return c->fail("null argument");
    const int L = static_cast<int>(c->cls_tower.size());
    if (L > SYLPH_CG_MAX_TOWER) return c->fail("more than %d class-tower layers", SYLPH_CG_MAX_TOWER);
    for (int i = 0; i < L; ++i)
        if (!params->conv_w[i] || !params->conv_b[i] || !params->gn_w[i] || !params->gn_b[i]) return c->fail("class-tower layer %d: missing parameter tensor", i);
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int i = 0; i < L; ++i) {
        ConvW& W = c->cls_tower[i];
        if (W.taps != 9 || W.cin != 256 || W.cout != 256 || W.w_nm != nullptr) return c->fail("class-tower layer %d is not a 3x3 256 -> 256 convolution", i);
        CU_TRY(c, launch_k(pack_oihw_weights_kernel, dim3(grid_for(256 * 256 * 9, 256, c->num_sms)), dim3(256), 0, st,
                           static_cast<const float*>(params->conv_w[i]), W.w, 256, 256, 9, W.cout_pad, c->split, 0));
        CU_TRY(c, cudaMemcpyAsync(W.bias, params->conv_b[i], 1024, cudaMemcpyDeviceToDevice, st));
        CU_TRY(c, cudaMemcpyAsync(c->cls_gn_w[i], params->gn_w[i], 1024, cudaMemcpyDeviceToDevice, st));
        CU_TRY(c, cudaMemcpyAsync(c->cls_gn_b[i], params->gn_b[i], 1024, cudaMemcpyDeviceToDevice, st));
        c->launches++;
    }
    return prep_cls_tower_transposed(c, params, st);
}

int sylph_cls_tower_backward(sylph_ctx* c, int slot, int n_classes, const float* codes_dev, const int64_t* support_targets_host,
                             const sylph_loss_config* lc, const int64_t* labels_dev, const double* local_sums_dev,
                             const double* global_pos_ctr_dev, int world_size, const float* grad_loss_dev,
                             const sylph_tower_tensors* params, const sylph_tower_tensors* grads, void* stream) {
    if (!c) return 1;
    if (!c->finalized) return c->fail("weights not finalized");
    if (slot < 0 || slot >= SYLPH_NUM_SLOTS || !c->slots[slot].valid) return c->fail("slot %d holds no features", slot);
    if (!codes_dev || !support_targets_host || !lc || !labels_dev || !local_sums_dev || !params || !grads) return c->fail("null argument");
    if (world_size < 1) return c->fail("world_size must be >= 1");
    const int L = static_cast<int>(c->cls_tower.size());
    if (L < 1 || L > SYLPH_CG_MAX_TOWER) return c->fail("class tower of %d layers", L);
    if (c->last_detect_slot != slot || c->last_detect_classes != n_classes || static_cast<int>(c->saved_x.size()) != L || !c->train_save)
        return c->fail("sylph_cls_tower_backward: the last head pass was not sylph_fcos_loss_sums on slot %d with %d classes in training mode "
                       "(sylph_set_training)", slot, n_classes);
    for (int i = 0; i < L; ++i)
        if (!params->conv_w[i] || !params->gn_w[i] || !params->gn_b[i] || !grads->conv_w[i] || !grads->conv_b[i] || !grads->gn_w[i] || !grads->gn_b[i])
            return c->fail("class-tower layer %d: missing parameter / gradient tensor", i);
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Slot& S = c->slots[slot];
    if (!c->cls_tower_t_ready) TRY(prep_cls_tower_transposed(c, params, st));
    auto lg = c->bufs.find("det.logits");
    if (lg == c->bufs.end() || !lg->second.p) return c->fail("no logits buffer");
    const long long rows = S.level_row0[5];
    const int tiles = static_cast<int>(rows / kBlockM);
    const int n_segs = 5 * S.n;
    long long total_px = 0;
    for (int l = 0; l < 5; ++l) total_px += static_cast<long long>(S.n) * S.lh[l] * S.lw[l];
    const std::string sig = std::to_string(rows) + ":" + std::to_string(S.n) + ":" + std::to_string(S.hpad) + "x" + std::to_string(S.wpad);
    void *stg, *dxa, *dxb, *dyp, *part, *segsum, *ab, *segmax, *scales, *biasp, *wpart;
    const int splits = std::max(1, c->num_sms / 18);
    TRY(ensure(c, "tbw.support_targets", static_cast<size_t>(n_classes) * 8, "", &stg, st, false));
    TRY(ensure(c, "tbw.dx_a", (static_cast<size_t>(rows) + kBlockM) * 256 * 4, sig, &dxa, st, true));   // gradient entering a layer's ReLU (fp32)
    TRY(ensure(c, "tbw.dx_b", (static_cast<size_t>(rows) + kBlockM) * 256 * 4, sig, &dxb, st, true));
    TRY(ensure(c, "tbw.dy", (static_cast<size_t>(rows) + kBlockM) * c->ld(256) * 2, sig, &dyp, st, true));  // scaled dY planes
    TRY(ensure(c, "tbw.partial", static_cast<size_t>(tiles) * kGnBwdPartial * 4, "", &part, st, false));
    TRY(ensure(c, "tbw.seg_sums", static_cast<size_t>(n_segs) * 2 * 256 * 4, "", &segsum, st, false));
    TRY(ensure(c, "tbw.ab", static_cast<size_t>(n_segs) * 64 * 4, "", &ab, st, false));
    TRY(ensure(c, "tbw.seg_max", static_cast<size_t>(n_segs) * 4, "", &segmax, st, false));
    TRY(ensure(c, "tbw.scales", static_cast<size_t>(SYLPH_CG_MAX_TOWER) * 2 * 4, "", &scales, st, false));
    TRY(ensure(c, "tbw.bias_partial", static_cast<size_t>(tiles) * 256 * 4, "", &biasp, st, false));
    TRY(ensure(c, "tbw.wgrad_partial", static_cast<size_t>(splits) * 9 * 256 * 256 * 4, "", &wpart, st, false));
    TRY(stage_h2d(c, stg, support_targets_host, static_cast<size_t>(n_classes) * 8, st));
    StageTimer timer(c, "bwd.cls_tower", st, 0.0);
    // ---- gradient of the loss with respect to the tower's output (through the conditional convolution)
    {
        const int blocks = static_cast<int>(std::min<long long>(8LL * c->num_sms, (total_px + kClsBwdRows - 1) / kClsBwdRows));
        CU_TRY(c, launch_k(tower_out_grad_kernel, dim3(blocks), dim3(256), 0, st, static_cast<const float*>(lg->second.p), c->last_logit_stride,
                           codes_dev, S.pg, S.n, reinterpret_cast<const long long*>(labels_dev), static_cast<const long long*>(stg), n_classes,
                           lc->focal_alpha, lc->focal_gamma, global_pos_ctr_dev ? global_pos_ctr_dev : local_sums_dev + 1, world_size,
                           grad_loss_dev, static_cast<float*>(dxa)));
        c->launches++;
    }
    float* dx_in = static_cast<float*>(dxa);
    float* dx_out = static_cast<float*>(dxb);
    const float* in_scale = nullptr;          // the scale of the gradient in dx_in (NULL = 1)
    const int g_tiles = std::max(1, std::min(tiles, c->num_sms * 8));
    for (int i = L - 1; i >= 0; --i) {
        float* sc = static_cast<float*>(scales) + 2 * i;
        CU_TRY(c, launch_k(gn_bwd_partial_kernel, dim3(g_tiles), dim3(256), 0, st, static_cast<const float*>(dx_in), in_scale, c->saved_raw[i],
                           c->saved_stats[i], static_cast<const float*>(params->gn_w[i]), static_cast<const float*>(params->gn_b[i]),
                           static_cast<const int*>(S.ps->d_tile_seg), static_cast<const Seg*>(S.ps->d_segs), tiles, static_cast<float*>(part)));
        CU_TRY(c, launch_k(gn_bwd_finalize_kernel, dim3(n_segs), dim3(1024), 0, st, static_cast<const float*>(part), static_cast<const Seg*>(S.ps->d_segs),
                           static_cast<const float*>(params->gn_w[i]), static_cast<float*>(segsum), static_cast<float*>(ab), static_cast<float*>(segmax)));
        CU_TRY(c, launch_k(gn_bwd_scale_kernel, dim3(1), dim3(256), 0, st, static_cast<const float*>(segmax), c->saved_stats[i], n_segs,
                           static_cast<const float*>(params->gn_w[i]), sc));
        CU_TRY(c, launch_k(gn_bwd_apply_kernel, dim3(g_tiles), dim3(256), 0, st, static_cast<const float*>(dx_in), in_scale, c->saved_raw[i],
                           c->saved_stats[i], static_cast<const float*>(ab), static_cast<const float*>(params->gn_w[i]),
                           static_cast<const float*>(params->gn_b[i]), static_cast<const float*>(sc), static_cast<const int*>(S.ps->d_tile_seg),
                           static_cast<const Seg*>(S.ps->d_segs), tiles, static_cast<__half*>(dyp), c->split, static_cast<float*>(biasp)));
        CU_TRY(c, launch_k(tower_param_grad_reduce_kernel, dim3(3), dim3(1024), 0, st, static_cast<const float*>(segsum), n_segs,
                           static_cast<const float*>(biasp), tiles, grads->gn_w[i], grads->gn_b[i], grads->conv_b[i]));
        CU_TRY(c, cudaGetLastError());
        c->launches += 5;
        // ---- weight gradient: dW = dY^T X_i per tap, straight from the planes (wgrad3x3.cuh)
        {
            CUtensorMap tdy, tx;
            std::string err;
            if (make_tmap_2d(&tdy, static_cast<const __half*>(dyp), static_cast<uint64_t>(rows), c->ld(256), c->ld(256), kWgTileK, &err) ||
                make_tmap_2d(&tx, c->saved_x[i], static_cast<uint64_t>(rows), c->ld(256), c->ld(256), kWgTileK, &err))
                return c->fail("wgrad tensor maps: %s", err.c_str());
            WgradArgs wa{static_cast<const Seg*>(S.ps->d_segs), static_cast<const int*>(S.ps->d_tile_seg), static_cast<int>(rows / kWgTileK),
                         static_cast<float*>(wpart)};
            if (c->split) {
                CU_TRY(c, cudaFuncSetAttribute(wgrad3x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgradSmem<true>::kTotal));   // per device: set on every call
                CU_TRY(c, launch_k(wgrad3x3_kernel<true>, dim3(splits, 9, 2), dim3(kWgThreads), WgradSmem<true>::kTotal, st, tdy, tx, wa));
            } else {
                CU_TRY(c, cudaFuncSetAttribute(wgrad3x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgradSmem<false>::kTotal));
                CU_TRY(c, launch_k(wgrad3x3_kernel<false>, dim3(splits, 9, 2), dim3(kWgThreads), WgradSmem<false>::kTotal, st, tdy, tx, wa));
            }
            CU_TRY(c, launch_k(wgrad_reduce_scaled_kernel, dim3(9 * 256), dim3(256), 0, st, static_cast<const float*>(wpart), splits,
                               static_cast<const float*>(sc), grads->conv_w[i]));
            CU_TRY(c, cudaGetLastError());
            c->launches += 2;
        }
        // ---- input gradient: the forward convolution kernel on dY with the transposed, tap-reversed weights -> fp32, still scaled
        if (i > 0) {
            ConvCall k{};
            k.W = &c->cls_tower[i]; k.w_override = c->cls_tower_t[i].w; k.bias_override = c->cls_tower_t[i].bias;
            k.A = static_cast<const __half*>(dyp); k.a_rows = rows; k.a_cols = k.a_ld = c->ld(256); k.ps = S.ps.get(); k.tile_begin = 0;
            k.n_tiles = tiles; k.a_row_delta = 0; k.out = dx_out; k.ldc = 256; k.flags = kEpiOutF32; k.name = "bwd.cls_tower_dgrad3x3";
            TRY(run_conv(c, k, st));
            std::swap(dx_in, dx_out);
            in_scale = sc;
        }
    }
    return 0;
}

int sylph_debug_read_buffer(sylph_ctx* c, const char* name, void* out_dev, size_t bytes, void* stream) {
    if (!c || !name || !out_dev) return 1;
    auto it = c->bufs.find(name);
    if (it == c->bufs.end() || !it->second.p) return c->fail("no buffer named %s", name);
    if (bytes > it->second.cap) return c->fail("buffer %s holds %zu bytes, %zu requested", name, it->second.cap, bytes);
    CU_TRY(c, cudaMemcpyAsync(out_dev, it->second.p, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return 0;
}

int64_t sylph_launch_count(const sylph_ctx* c) { return c ? c->launches : 0; }

int sylph_set_profiling(sylph_ctx* c, int enabled) {
    if (!c) return 1;
    for (auto& t : c->timings) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
    c->timings.clear();
    c->profiling = enabled != 0;
    return 0;
}

int sylph_get_timings(sylph_ctx* c, char (*names)[48], float* ms, double* flops, double* bytes, int cap) {
    if (!c) return -1;
    cudaDeviceSynchronize();
    const int n = static_cast<int>(c->timings.size());
    for (int i = 0; i < n && i < cap; ++i) {
        strncpy(names[i], c->timings[i].name.c_str(), 47);
        names[i][47] = 0;
        cudaEventElapsedTime(&ms[i], c->timings[i].e0, c->timings[i].e1);
        flops[i] = c->timings[i].flops;
        bytes[i] = c->timings[i].bytes;
    }
    return n;
}

}  // extern "C"
