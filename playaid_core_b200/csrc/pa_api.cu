// C-ABI of the B200-native fighter action-recognition path (see include/playaid_b200.h).
// Host-side orchestration only: weight packing, BatchNorm folding, TMA descriptor construction,
// workspace carving and kernel sequencing. No PyTorch types cross this boundary.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "pa_internal.cuh"

using namespace pa;

// ------------------------------------------------------------------------------------------------
struct pa_ctx {
    int device = 0;
    int num_sms = 148;
    int64_t launches = 0;
    std::string last_error;
    decltype(&cuTensorMapEncodeTiled) encode_tiled = nullptr;
    // Scratch of pa_preprocess / pa_stage_windows, one set PER STREAM (calls on different streams may overlap on the
    // device; calls on one stream are ordered by it). Sized for 1024 crops on a stream's first call and grown (one
    // cudaMalloc + cudaFree, which synchronise the device) only when a call brings more crops than any before it.
    struct Scratch {
        int* stage_sched = nullptr;    // pa_stage_windows scheduling counters
        int32_t* pp_status = nullptr;  // per-crop status when the caller passes none
        int* pp_deferred = nullptr;    // device counter: slabs deferred to the large-window pass
        uint8_t* pp_plan = nullptr;    // per-crop geometry + coefficient tables (preprocess_plan_kernel)
        int2* tc_items = nullptr;      // work items of the tensor-core preprocess kernel
        int* tc_counters = nullptr;    // [0] enqueued, [1] taken, [2] CTAs done, [3] tiles reserved
        uint8_t* tc_tiles = nullptr;   // vertical coefficient tiles of the tensor-core preprocess kernel (built by the plan kernel)
        int* tc_tile_rec = nullptr;
        int cap = 0;                   // crops pp_status / pp_plan / tc_items are sized for
    };
    std::map<cudaStream_t, Scratch> scratch;
    std::mutex scratch_mu;
    // optional per-kernel timing (CUDA events on the launching stream)
    bool profiling = false;
    struct Span { std::string name; cudaEvent_t e0, e1; };
    std::vector<Span> spans;
};

static int cuda_fail(pa_ctx* ctx, cudaError_t e, const char* what);

// Experiment switches (PA_NO_PATCH, PA_NO_PAIR, PA_PP_*, PA_CONV_DEBUG, PA_ST_*) are compiled out of the shipped
// library: they exist only when it is built with -DPA_EXPERIMENT, and are then read once.
static bool exp_flag(const char* name) {
#ifdef PA_EXPERIMENT
    static std::map<std::string, bool> cache;
    auto it = cache.find(name);
    if (it == cache.end()) it = cache.emplace(name, getenv(name) != nullptr).first;
    return it->second;
#else
    (void)name;
    return false;
#endif
}

// RAII span: records an event pair around one launch when profiling is on
struct ProfSpan {
    pa_ctx* ctx; cudaStream_t st; int idx = -1;
    ProfSpan(pa_ctx* c, const std::string& name, cudaStream_t s) : ctx(c), st(s) {
        if (!ctx || !ctx->profiling) return;
        pa_ctx::Span sp; sp.name = name;
        cudaEventCreate(&sp.e0); cudaEventCreate(&sp.e1);
        cudaEventRecord(sp.e0, st);
        ctx->spans.push_back(sp);
        idx = (int)ctx->spans.size() - 1;
    }
    ~ProfSpan() { if (idx >= 0) cudaEventRecord(ctx->spans[idx].e1, st); }
};

extern "C" int pa_profile_begin(pa_ctx* ctx) {
    if (!ctx) return PA_ERR_INVALID_ARG;
    for (auto& s : ctx->spans) { cudaEventDestroy(s.e0); cudaEventDestroy(s.e1); }
    ctx->spans.clear();
    ctx->profiling = true;
    return PA_OK;
}

// Stops profiling, synchronises the recorded events and writes "name\tlaunches\ttotal_ms\n" lines.
extern "C" int pa_profile_end(pa_ctx* ctx, char* buf, size_t buflen) {
    if (!ctx || !buf || buflen == 0) return PA_ERR_INVALID_ARG;
    ctx->profiling = false;
    std::map<std::string, std::pair<int, double>> acc;
    std::vector<std::string> order;
    for (auto& s : ctx->spans) {
        if (cudaEventSynchronize(s.e1) != cudaSuccess) return cuda_fail(ctx, cudaGetLastError(), "profile sync");
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s.e0, s.e1);
        if (!acc.count(s.name)) order.push_back(s.name);
        acc[s.name].first += 1;
        acc[s.name].second += ms;
        cudaEventDestroy(s.e0); cudaEventDestroy(s.e1);
    }
    ctx->spans.clear();
    std::string out;
    for (auto& n : order) {
        char line[256];
        snprintf(line, sizeof(line), "%s\t%d\t%.6f\n", n.c_str(), acc[n].first, acc[n].second);
        out += line;
    }
    if (out.size() + 1 > buflen) return PA_ERR_WORKSPACE;
    memcpy(buf, out.c_str(), out.size() + 1);
    return PA_OK;
}

static int cuda_fail(pa_ctx* ctx, cudaError_t e, const char* what) {
    if (ctx) ctx->last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return PA_ERR_CUDA;
}
#define PA_CUDA(ctx, expr)                                            \
    do {                                                              \
        cudaError_t _e = (expr);                                      \
        if (_e != cudaSuccess) return cuda_fail((ctx), _e, #expr);    \
    } while (0)

extern "C" int pa_abi_version(void) { return PA_ABI_VERSION; }

extern "C" const char* pa_status_string(int s) {
    switch (s) {
        case PA_OK: return "ok";
        case PA_ERR_INVALID_ARG: return "invalid argument";
        case PA_ERR_CUDA: return "CUDA error (see pa_last_error)";
        case PA_ERR_UNSUPPORTED: return "unsupported configuration";
        case PA_ERR_NOT_READY: return "model not finalized";
        case PA_ERR_WORKSPACE: return "workspace too small";
        case PA_ERR_MISSING_TENSOR: return "state_dict tensor missing or wrong shape";
        default: return "unknown status";
    }
}

extern "C" const char* pa_last_error(pa_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

extern "C" int pa_ctx_create(int device, pa_ctx** out) {
    if (!out) return PA_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return PA_ERR_CUDA;
    pa_ctx* ctx = new pa_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return PA_ERR_CUDA;
    }
    if (prop.major != 10) {  // sm_100a only: tcgen05 / TMEM
        delete ctx;
        return PA_ERR_UNSUPPORTED;
    }
    ctx->num_sms = prop.multiProcessorCount;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
        delete ctx;
        return PA_ERR_CUDA;
    }
    ctx->encode_tiled = (decltype(&cuTensorMapEncodeTiled))fn;
    *out = ctx;
    return PA_OK;
}

extern "C" int pa_ctx_destroy(pa_ctx* ctx) {
    if (ctx)
        for (auto& kv : ctx->scratch) {
            pa_ctx::Scratch& sc = kv.second;
            if (sc.pp_status) cudaFree(sc.pp_status);
            if (sc.pp_deferred) cudaFree(sc.pp_deferred);
            if (sc.pp_plan) cudaFree(sc.pp_plan);
            if (sc.stage_sched) cudaFree(sc.stage_sched);
            if (sc.tc_items) cudaFree(sc.tc_items);
            if (sc.tc_counters) cudaFree(sc.tc_counters);
            if (sc.tc_tiles) cudaFree(sc.tc_tiles);
            if (sc.tc_tile_rec) cudaFree(sc.tc_tile_rec);
        }
    delete ctx;
    return PA_OK;
}

extern "C" int64_t pa_launch_count(pa_ctx* ctx) { return ctx ? ctx->launches : 0; }

static const int kTableStride = 40960;  // int32 per crop (160 KB) of coefficient tables: enough for a full-frame 1080p window

// the calling stream's scratch, sized for at least n_crops
static int get_scratch(pa_ctx* ctx, cudaStream_t st, int n_crops, pa_ctx::Scratch** out) {
    std::lock_guard<std::mutex> lock(ctx->scratch_mu);
    pa_ctx::Scratch& sc = ctx->scratch[st];
    if (!sc.stage_sched) PA_CUDA(ctx, cudaMalloc((void**)&sc.stage_sched, PA_STAGE_SCHED_INTS * sizeof(int)));
    if (!sc.pp_deferred) { PA_CUDA(ctx, cudaMalloc((void**)&sc.pp_deferred, sizeof(int))); PA_CUDA(ctx, cudaMemset(sc.pp_deferred, 0, sizeof(int))); }
    if (!sc.tc_counters) { PA_CUDA(ctx, cudaMalloc((void**)&sc.tc_counters, 4 * sizeof(int))); PA_CUDA(ctx, cudaMemset(sc.tc_counters, 0, 4 * sizeof(int))); }
    if (sc.cap < n_crops) {
        if (sc.pp_status) cudaFree(sc.pp_status);
        if (sc.pp_plan) cudaFree(sc.pp_plan);
        if (sc.tc_items) cudaFree(sc.tc_items);
        if (sc.tc_tiles) cudaFree(sc.tc_tiles);
        if (sc.tc_tile_rec) cudaFree(sc.tc_tile_rec);
        sc.pp_status = nullptr; sc.pp_plan = nullptr; sc.tc_items = nullptr; sc.tc_tiles = nullptr; sc.tc_tile_rec = nullptr; sc.cap = 0;
        const int cap = n_crops < 1024 ? 1024 : n_crops;
        const size_t geom_b = preprocess_geom_bytes();
        PA_CUDA(ctx, cudaMalloc((void**)&sc.pp_status, (size_t)cap * sizeof(int32_t)));
        PA_CUDA(ctx, cudaMalloc((void**)&sc.pp_plan, (((size_t)cap * geom_b + 255) & ~(size_t)255) + (size_t)cap * kTableStride * 4));
        PA_CUDA(ctx, cudaMalloc((void**)&sc.tc_items, (size_t)cap * PA_TC_ITEMS_PER_CROP * sizeof(int2)));
        PA_CUDA(ctx, cudaMalloc((void**)&sc.tc_tiles, (size_t)cap * PA_TC_TILES_PER_CROP * PA_TC_TILE_BYTES));
        PA_CUDA(ctx, cudaMalloc((void**)&sc.tc_tile_rec, (size_t)cap * PA_TC_TILES_PER_CROP * sizeof(int)));
        sc.cap = cap;
    }
    *out = &sc;
    return PA_OK;
}

// ------------------------------------------------------------------------------------------------ preprocess
extern "C" int pa_stage_windows(pa_ctx* ctx, const uint8_t* host_frames, int n_frames, int H, int W, int64_t pitch_bytes,
                                int64_t frame_stride_bytes, const int32_t* boxes, int n_crops, int padding, int frame_base,
                                uint8_t* dev_frames, void* stream) {
    if (!ctx || n_frames <= 0 || H <= 0 || W <= 0 || n_crops < 0 || padding < 0) return PA_ERR_INVALID_ARG;
    if (pitch_bytes < (int64_t)W * 3 || frame_stride_bytes < pitch_bytes * H) return PA_ERR_INVALID_ARG;
    if (n_crops == 0) return PA_OK;      // an empty record list is a valid no-op (its pointers may be null)
    if (!host_frames || !dev_frames || !boxes) return PA_ERR_INVALID_ARG;
    StageParams p;
    p.src = host_frames; p.dst = dev_frames;
    p.frames_bytes = frame_stride_bytes * (int64_t)(n_frames - 1) + pitch_bytes * (int64_t)(H - 1) + (int64_t)W * 3;
    p.n_frames = n_frames; p.H = H; p.W = W; p.pitch = pitch_bytes; p.fstride = frame_stride_bytes;
    p.boxes = boxes; p.n_crops = n_crops; p.padding = padding; p.frame_base = frame_base;
    pa_ctx::Scratch* sc = nullptr;
    { int rc = get_scratch(ctx, (cudaStream_t)stream, 0, &sc); if (rc != PA_OK) return rc; }
    p.sched = sc->stage_sched;
    ProfSpan sp(ctx, "stage_windows", (cudaStream_t)stream);
    if (launch_stage_windows(p, ctx->num_sms, (cudaStream_t)stream) != PA_OK) return cuda_fail(ctx, cudaGetLastError(), "stage_windows launch");
    ctx->launches += 1;
    return PA_OK;
}

extern "C" int pa_preprocess(pa_ctx* ctx, const uint8_t* frames, int n_frames, int H, int W, int64_t pitch_bytes,
                             int64_t frame_stride_bytes, const int32_t* boxes, int n_crops, int out_size, int padding,
                             int swap_rb, const float* mean3, const float* std3, void* out, int out_dtype,
                             int out_layout, int32_t* status, void* stream) {
    if (!ctx || n_frames <= 0 || H <= 0 || W <= 0 || n_crops < 0) return PA_ERR_INVALID_ARG;
    if (out_size <= 0 || out_size > 1024 || padding < 0) return PA_ERR_INVALID_ARG;
    if (out_dtype < PA_DTYPE_U8 || out_dtype > PA_DTYPE_F16_U8 || out_layout < PA_LAYOUT_NHWC || out_layout > PA_LAYOUT_NHWC4P)
        return PA_ERR_INVALID_ARG;
    if (pitch_bytes < (int64_t)W * 3 || frame_stride_bytes < pitch_bytes * H) return PA_ERR_INVALID_ARG;
    if (n_crops == 0) return PA_OK;      // an empty record list is a valid no-op (records / out / status may be null)
    if (!frames || !boxes || !out) return PA_ERR_INVALID_ARG;
    PPParams p;
    p.frames = frames;
    p.frames_bytes = frame_stride_bytes * (int64_t)(n_frames - 1) + pitch_bytes * (int64_t)(H - 1) + (int64_t)W * 3;
    p.n_frames = n_frames; p.H = H; p.W = W;
    p.pitch = pitch_bytes; p.fstride = frame_stride_bytes;
    p.boxes = boxes; p.n_crops = n_crops; p.out = out_size; p.padding = padding; p.swap_rb = swap_rb ? 1 : 0;
    for (int c = 0; c < 3; c++) {
        p.mean[c] = mean3 ? mean3[c] : 0.f;
        p.stdv[c] = std3 ? std3[c] : 1.f;
    }
    p.outp = out; p.out_dtype = out_dtype; p.out_layout = out_layout;
    p.out_f16 = (out_dtype == PA_DTYPE_F16 || out_dtype == PA_DTYPE_F16X2 || out_dtype == PA_DTYPE_F16_U8) ? 1 : 0;
    p.out_split = (out_dtype == PA_DTYPE_BF16X2 || out_dtype == PA_DTYPE_F16X2) ? 1 : 0;
    p.out_raw = (out_dtype == PA_DTYPE_BF16_U8 || out_dtype == PA_DTYPE_F16_U8) ? 1 : 0;
    if (out_layout == PA_LAYOUT_NHWC4P && (out_dtype == PA_DTYPE_U8 || out_dtype == PA_DTYPE_F32)) return PA_ERR_UNSUPPORTED;
    const int ch = (out_layout == PA_LAYOUT_NHWC4 || out_layout == PA_LAYOUT_NHWC4P) ? 4 : 3;
    p.plane_elems = (int64_t)n_crops * out_size * (out_size + (out_layout == PA_LAYOUT_NHWC4P ? 8 : 0)) * ch;
    pa_ctx::Scratch* sc = nullptr;
    { int rc = get_scratch(ctx, (cudaStream_t)stream, n_crops, &sc); if (rc != PA_OK) return rc; }
    if (!status) status = sc->pp_status;
    p.status = status;
    // pass 1: 108 KB of shared memory per 384-thread CTA (2 CTAs / SM) covers the usual fighter windows;
    // pass 2: the few crops that did not fit are redone with the whole carve-out (1 CTA / SM).
    // No memsets in front of the plan kernel: it initialises the per-crop status and the deferred-slab counter itself, and the
    // tensor-core kernel's last CTA re-zeroes the work-item counters. The plan kernel is launched with programmatic stream
    // serialization and never waits on its predecessor (it reads only the crop records), so it runs under the tail of
    // whatever kernel precedes it in the stream.
    p.deferred = sc->pp_deferred;
    // geometry + coefficient tables once per crop (L2-resident scratch of this stream)
    const size_t geom_b = preprocess_geom_bytes();
    p.geoms = sc->pp_plan;
    p.tables = (int*)(sc->pp_plan + (((size_t)sc->cap * geom_b + 255) & ~(size_t)255));
    p.table_stride = kTableStride;
    // Tensor-core path (preprocess_tc.inc): the raw window bytes go from the frame to the tensor core by TMA, which needs
    // the frames in DEVICE memory with 16-byte aligned rows. Everything else (pinned host frames read in place, unaligned
    // pitches, crops the plan kernel does not admit) stays on the streaming CUDA-core kernel.
    CUtensorMap frames_map;
    p.tc_enable = 0; p.tc_items = sc->tc_items; p.tc_counters = sc->tc_counters;
    p.tc_tiles = sc->tc_tiles; p.tc_tile_rec = sc->tc_tile_rec; p.tc_pool_tiles = sc->cap * PA_TC_TILES_PER_CROP;
    if (!exp_flag("PA_NO_TC") && (pitch_bytes & 15) == 0 && (frame_stride_bytes & 15) == 0 && ((uintptr_t)frames & 15) == 0 &&
        (int64_t)W * 3 >= 128 && out_size <= 128) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, frames) == cudaSuccess && attr.type == cudaMemoryTypeDevice) {
            cuuint64_t dims[3] = {(cuuint64_t)W * 3, (cuuint64_t)H, (cuuint64_t)n_frames};
            cuuint64_t strides[2] = {(cuuint64_t)pitch_bytes, (cuuint64_t)frame_stride_bytes};
            cuuint32_t box[3] = {128, 128, 1}, estr[3] = {1, 1, 1};
            CUresult r = ctx->encode_tiled(&frames_map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)frames, dims, strides, box, estr,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r == CUDA_SUCCESS) p.tc_enable = 1;
        } else {
            cudaGetLastError();   // a plain host pointer makes cudaPointerGetAttributes fail on old drivers: not an error here
        }
    }
    {
        ProfSpan sp(ctx, "preprocess_plan", (cudaStream_t)stream);
        if (launch_preprocess_plan(p, (cudaStream_t)stream) != PA_OK) return cuda_fail(ctx, cudaGetLastError(), "preprocess plan launch");
    }
    if (p.tc_enable) {
        ProfSpan sp(ctx, "preprocess_tc", (cudaStream_t)stream);
        if (launch_preprocess_tc(p, frames_map, ctx->num_sms, (cudaStream_t)stream) != PA_OK) return cuda_fail(ctx, cudaGetLastError(), "preprocess tc launch");
        ctx->launches += 1;
    }
    // tunables (defaults measured on B200; PA_PP_* environment variables override for experiments)
    static int cfg_threads = 0, cfg_smem_kb = 0, cfg_xb = 0;
    if (!cfg_threads) {
        const char* e;
        cfg_threads = 256; cfg_smem_kb = 72; cfg_xb = 0;
#ifdef PA_EXPERIMENT   // tuning switches exist only in experiment builds (python -m playaid_core_b200.build --experiment)
        cfg_threads = (e = getenv("PA_PP_THREADS")) ? atoi(e) : 256;
        cfg_smem_kb = (e = getenv("PA_PP_SMEM_KB")) ? atoi(e) : 72;
        cfg_xb = (e = getenv("PA_PP_XB")) ? atoi(e) : 0;
#else
        (void)e;
#endif
        if (cfg_threads != 256 && cfg_threads != 384) cfg_threads = 256;
        if (cfg_smem_kb < 48 || cfg_smem_kb > 224) cfg_smem_kb = 72;
    }
    p.threads = cfg_threads;
    p.num_sms = ctx->num_sms;
    p.use_xb = cfg_xb;
    p.smem_bytes = cfg_smem_kb * 1024;
    p.first_pass_smem = 0;
    p.defer_too_large = 1;
    p.overlap_prev = p.tc_enable;
    int rc;
    {
        ProfSpan sp(ctx, "preprocess", (cudaStream_t)stream);
        rc = launch_preprocess(p, (cudaStream_t)stream);
    }
    if (rc != PA_OK) return cuda_fail(ctx, cudaGetLastError(), "preprocess launch");
    p.smem_bytes = 216 * 1024;   // not the whole carve-out: a window-staging worker (6.7 KB) may sit on the SM and must not block this launch
    p.first_pass_smem = cfg_smem_kb * 1024;
    p.defer_too_large = 0;
    p.overlap_prev = 0;
    {
        ProfSpan sp(ctx, "preprocess_large_windows", (cudaStream_t)stream);
        rc = launch_preprocess(p, (cudaStream_t)stream);
    }
    if (rc != PA_OK) return cuda_fail(ctx, cudaGetLastError(), "preprocess launch (large windows)");
    ctx->launches += 3;
    return PA_OK;
}

extern "C" int pa_boxes_from_log(pa_ctx* ctx, const double* log_records, int n, int W, int H, double* boxes, int32_t* crop_records, void* stream) {
    if (!ctx || n < 0 || W <= 0 || H <= 0) return PA_ERR_INVALID_ARG;
    if (n == 0) return PA_OK;
    if (!log_records || (!boxes && !crop_records)) return PA_ERR_INVALID_ARG;
    ProfSpan sp(ctx, "boxes_from_log", (cudaStream_t)stream);
    if (launch_boxes(log_records, n, W, H, boxes, crop_records, (cudaStream_t)stream) != PA_OK) return cuda_fail(ctx, cudaGetLastError(), "boxes launch");
    ctx->launches += 1;
    return PA_OK;
}

extern "C" size_t pa_crop_elems(int out_size) { return (size_t)out_size * (out_size + 8) * 4; }

// ------------------------------------------------------------------------------------------------ model
struct HostTensor {
    std::vector<float> data;
    std::vector<int64_t> shape;
};

struct ConvLayer {
    std::string w_key, bn_key;  // bn_key empty -> bias_key used
    std::string bias_key;
    int cin = 0, cout = 0, k = 1, stride = 1, pad = 0;
    int hin = 0;                 // input spatial size (square)
    bool relu = false;
    // device
    bf16 *w_hi = nullptr, *w_lo = nullptr;
    float *scale = nullptr, *shift = nullptr;
    int k_total = 0;
    int block_n = 64;
};

struct PlanOp {
    char name[64];
    int kind;  // 0 stem, 2 conv gemm, 3 avgpool, 4 conv 3x3 patch mode, 5/6 their CTA-pair variants, 10..14 transformer pieces
    int patch_ht; bool patch_wres; size_t patch_smem;
    Conv1Args c1;
    Conv1Maps c1maps;
    ConvMaps maps;
    ConvArgs args;
    int block_n, n_a, n_b;
    // pools
    const bf16 *pin_hi, *pin_lo;
    bf16 *pout_hi, *pout_lo;
    int pn, ph, pw, pc, pf16;
    // transformer ops (kinds 10..14: tokens, fp32 -> 16-bit, attention, residual + LayerNorm, log-softmax)
    const float *fa, *fb, *fg, *fbeta;
    float* fx;
    int i0, i1, i2;
    float eps;
};

struct TfLayer {   // one post-norm TransformerEncoderLayer
    ConvLayer in_proj, out_proj, lin1, lin2;
    float *g1 = nullptr, *b1 = nullptr, *g2 = nullptr, *b2 = nullptr;
    float eps1 = 1e-5f, eps2 = 1e-5f;
};

struct pa_model {
    pa_ctx* ctx = nullptr;
    int arch = 0;                  // 0 CNNActionDetector, 1 ResnetTransformerDetector (ResFormer)
    ConvLayer ffn, cls;            // ResFormer: Linear(2048, hidden), Linear(256, A)
    std::vector<TfLayer> tf;
    float* enc = nullptr;          // ResFormer time encoding [seq][256 - hidden]
    int hidden = 0;
    int n_actions = 0, seq = 0;
    int precision = -1;
    bool ready = false;
    std::map<std::string, HostTensor> tensors;
    ConvLayer stem;
    std::vector<ConvLayer> convs;  // resnet body in execution order (incl. downsample)
    ConvLayer fc, proj;
    float *b1d = nullptr, *w1t = nullptr, *b1 = nullptr, *w2t = nullptr, *b2 = nullptr;
    bf16 *stem_w_hi = nullptr, *stem_w_lo = nullptr;
    float* stem_scale_u8 = nullptr;   // BN scale / 255: the stem epilogue of pa_features_u8 (crops hold byte values)
    std::vector<void*> dev_allocs;
    // cached forward plan
    std::vector<PlanOp> plan;
    const void* plan_crops = nullptr;
    void* plan_ws = nullptr;
    float* plan_feat = nullptr;
    int plan_n = -1;
    bool plan_u8 = false;
    // cached head plan
    PlanOp head_gemm;
    const float* hplan_feat = nullptr;
    void* hplan_ws = nullptr;
    int hplan_n = -1;
};

static inline uint16_t f2bf(float f) {  // round-to-nearest-even, like __float2bfloat16_rn
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline float bf2f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// IEEE binary16, round-to-nearest-even, with subnormals and saturation to +-65504
static inline uint16_t f2h(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t sign = (u >> 16) & 0x8000u;
    u &= 0x7FFFFFFFu;
    if (u > 0x7F800000u) return (uint16_t)(sign | 0x7E00u);  // NaN
    if (u >= 0x477FF000u) return (uint16_t)(sign | 0x7BFFu);  // >= 65520 rounds past max: saturate
    if (u < 0x38800000u) {  // subnormal half (|f| < 2^-14)
        if (u < 0x33000000u) return (uint16_t)sign;  // < 2^-25: rounds to zero
        const int e = (int)(u >> 23);  // biased float exponent
        uint32_t m = (u & 0x7FFFFFu) | 0x800000u;
        const int shift = 126 - e;  // bits to drop so that 2^-24 is the unit
        const uint32_t half = 1u << (shift - 1);
        const uint32_t rem = m & ((1u << shift) - 1);
        m >>= shift;
        if (rem > half || (rem == half && (m & 1))) m++;
        return (uint16_t)(sign | m);
    }
    uint32_t v = u - 0x38000000u;  // rebias exponent 127 -> 15
    const uint32_t rem = v & 0x1FFFu;
    v >>= 13;
    if (rem > 0x1000u || (rem == 0x1000u && (v & 1))) v++;
    return (uint16_t)(sign | v);
}
static inline float h2f(uint16_t h) {
    const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1F, m = h & 0x3FF, u;
    if (e == 0) {
        if (m == 0) u = sign;
        else {
            int s = 0;
            while (!(m & 0x400)) { m <<= 1; s++; }
            m &= 0x3FF;
            u = sign | ((uint32_t)(113 - s) << 23) | (m << 13);
        }
    } else if (e == 31) u = sign | 0x7F800000u | (m << 13);
    else u = sign | ((e + 112) << 23) | (m << 13);
    float f;
    memcpy(&f, &u, 4);
    return f;
}

static inline bool prec_f16(int p) { return p >= PA_PREC_F16; }
static inline bool prec_split(int p) { return p == PA_PREC_BF16X2 || p == PA_PREC_BF16X3 || p == PA_PREC_F16X2 || p == PA_PREC_F16X3; }
static inline bool prec_split_w(int p) { return p == PA_PREC_BF16X3 || p == PA_PREC_F16X3; }

extern "C" int pa_model_create(pa_ctx* ctx, int n_actions, int seq_len, pa_model** out) {
    if (!ctx || !out || n_actions <= 0 || n_actions > 128 || seq_len <= 0 || seq_len > 15 || (seq_len % 2) == 0) return PA_ERR_INVALID_ARG;
    pa_model* m = new pa_model();
    m->ctx = ctx; m->n_actions = n_actions; m->seq = seq_len;
    *out = m;
    return PA_OK;
}

extern "C" int pa_model_destroy(pa_model* m) {
    if (!m) return PA_OK;
    for (void* p : m->dev_allocs) cudaFree(p);
    delete m;
    return PA_OK;
}

extern "C" int pa_model_set_tensor(pa_model* m, const char* name, const float* host, const int64_t* shape, int ndim) {
    if (!m || !name || !host || ndim < 0 || ndim > 8) return PA_ERR_INVALID_ARG;
    std::string key(name);
    if (key.rfind("model.", 0) == 0) key = key.substr(6);
    if (key.size() >= 19 && key.compare(key.size() - 19, 19, "num_batches_tracked") == 0) return PA_OK;
    HostTensor t;
    int64_t n = 1;
    for (int i = 0; i < ndim; i++) { t.shape.push_back(shape[i]); n *= shape[i]; }
    t.data.assign(host, host + n);
    m->tensors[key] = std::move(t);
    m->ready = false;
    return PA_OK;
}

static const HostTensor* get_tensor(pa_model* m, const std::string& key, std::initializer_list<int64_t> shape) {
    auto it = m->tensors.find(key);
    if (it == m->tensors.end()) { m->ctx->last_error = "missing tensor " + key; return nullptr; }
    if (it->second.shape != std::vector<int64_t>(shape)) { m->ctx->last_error = "bad shape for " + key; return nullptr; }
    return &it->second;
}

template <typename T>
static int upload(pa_model* m, const std::vector<T>& host, T** dev) {
    void* p = nullptr;
    PA_CUDA(m->ctx, cudaMalloc(&p, host.size() * sizeof(T) + 16));
    m->dev_allocs.push_back(p);
    PA_CUDA(m->ctx, cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dev = (T*)p;
    return PA_OK;
}

static void split_weights(const std::vector<float>& w, std::vector<uint16_t>& hi, std::vector<uint16_t>& lo, bool f16) {
    hi.resize(w.size()); lo.resize(w.size());
    for (size_t i = 0; i < w.size(); i++) {
        if (f16) {
            hi[i] = f2h(w[i]);
            lo[i] = f2h(w[i] - h2f(hi[i]));
        } else {
            hi[i] = f2bf(w[i]);
            lo[i] = f2bf(w[i] - bf2f(hi[i]));
        }
    }
}

// stem weights [64][3][7][7] -> [64][256] with k = ky*32 + (kx+1)*4 + c (zeros elsewhere): the K order of conv1.cu's im2col
static void pack_stem_weights(const float* w, std::vector<float>& packed) {
    packed.assign((size_t)64 * 256, 0.f);
    for (int o = 0; o < 64; o++)
        for (int c = 0; c < 3; c++)
            for (int ky = 0; ky < 7; ky++)
                for (int kx = 0; kx < 7; kx++)
                    packed[(size_t)o * 256 + ky * 32 + (kx + 1) * 4 + c] = w[(((size_t)o * 3 + c) * 7 + ky) * 7 + kx];
}

// pack [cout][cin][k][k] -> [cout][tap][cin], fold BN (or bias) into scale/shift, upload
static int prepare_conv(pa_model* m, ConvLayer& L) {
    const HostTensor* w = get_tensor(m, L.w_key, {L.cout, L.cin, L.k, L.k});
    if (!w) return PA_ERR_MISSING_TENSOR;
    const int taps = L.k * L.k;
    L.k_total = taps * L.cin;
    std::vector<float> packed((size_t)L.cout * L.k_total);
    for (int o = 0; o < L.cout; o++)
        for (int c = 0; c < L.cin; c++)
            for (int t = 0; t < taps; t++)
                packed[((size_t)o * taps + t) * L.cin + c] = w->data[((size_t)o * L.cin + c) * taps + t];
    std::vector<uint16_t> hi, lo;
    split_weights(packed, hi, lo, prec_f16(m->precision));
    int rc = upload(m, hi, (uint16_t**)&L.w_hi);
    if (rc != PA_OK) return rc;
    if (prec_split_w(m->precision)) { rc = upload(m, lo, (uint16_t**)&L.w_lo); if (rc != PA_OK) return rc; }
    std::vector<float> scale(L.cout, 1.f), shift(L.cout, 0.f);
    if (!L.bn_key.empty()) {
        const HostTensor *g = get_tensor(m, L.bn_key + ".weight", {L.cout}), *b = get_tensor(m, L.bn_key + ".bias", {L.cout}),
                         *mu = get_tensor(m, L.bn_key + ".running_mean", {L.cout}), *var = get_tensor(m, L.bn_key + ".running_var", {L.cout});
        if (!g || !b || !mu || !var) return PA_ERR_MISSING_TENSOR;
        for (int o = 0; o < L.cout; o++) {
            const double s = (double)g->data[o] / std::sqrt((double)var->data[o] + 1e-5);
            scale[o] = (float)s;
            shift[o] = (float)((double)b->data[o] - (double)mu->data[o] * s);
        }
    } else if (!L.bias_key.empty()) {
        const HostTensor* b = get_tensor(m, L.bias_key, {L.cout});
        if (!b) return PA_ERR_MISSING_TENSOR;
        for (int o = 0; o < L.cout; o++) shift[o] = b->data[o];
    }
    rc = upload(m, scale, &L.scale); if (rc != PA_OK) return rc;
    rc = upload(m, shift, &L.shift); if (rc != PA_OK) return rc;
    return PA_OK;
}

static ConvLayer make_conv(const std::string& w, const std::string& bn, int cin, int cout, int k, int stride, int pad, int hin, bool relu) {
    ConvLayer L;
    L.w_key = w; L.bn_key = bn; L.cin = cin; L.cout = cout; L.k = k; L.stride = stride; L.pad = pad; L.hin = hin; L.relu = relu;
    L.block_n = cout >= 256 ? 256 : cout;
    return L;
}

extern "C" int pa_model_finalize(pa_model* m, int precision) {
    if (!m) return PA_ERR_INVALID_ARG;
    if (precision < PA_PREC_BF16 || precision > PA_PREC_F16X3) return PA_ERR_INVALID_ARG;
    pa_ctx* ctx = m->ctx;
    PA_CUDA(ctx, cudaSetDevice(ctx->device));
    for (void* p : m->dev_allocs) cudaFree(p);
    m->dev_allocs.clear();
    m->convs.clear();
    m->plan_n = -1; m->hplan_n = -1;
    m->precision = precision;
    const std::string P = "cnn2d.";
    int rc;
    // ---- stem: [64][3][7][7] -> [64][256] with k = ky*32 + (kx+1)*4 + c
    {
        const HostTensor* w = get_tensor(m, P + "conv1.weight", {64, 3, 7, 7});
        if (!w) return PA_ERR_MISSING_TENSOR;
        std::vector<float> packed;
        pack_stem_weights(w->data.data(), packed);
        std::vector<uint16_t> hi, lo;
        split_weights(packed, hi, lo, prec_f16(precision));
        rc = upload(m, hi, (uint16_t**)&m->stem_w_hi); if (rc != PA_OK) return rc;
        m->stem_w_lo = nullptr;
        if (prec_split_w(precision)) { rc = upload(m, lo, (uint16_t**)&m->stem_w_lo); if (rc != PA_OK) return rc; }
        m->stem = make_conv("", P + "bn1", 3, 64, 7, 2, 3, 128, true);
        ConvLayer& L = m->stem;
        const HostTensor *g = get_tensor(m, L.bn_key + ".weight", {64}), *b = get_tensor(m, L.bn_key + ".bias", {64}),
                         *mu = get_tensor(m, L.bn_key + ".running_mean", {64}), *var = get_tensor(m, L.bn_key + ".running_var", {64});
        if (!g || !b || !mu || !var) return PA_ERR_MISSING_TENSOR;
        std::vector<float> scale(64), shift(64);
        for (int o = 0; o < 64; o++) {
            const double s = (double)g->data[o] / std::sqrt((double)var->data[o] + 1e-5);
            scale[o] = (float)s;
            shift[o] = (float)((double)b->data[o] - (double)mu->data[o] * s);
        }
        rc = upload(m, scale, &L.scale); if (rc != PA_OK) return rc;
        rc = upload(m, shift, &L.shift); if (rc != PA_OK) return rc;
        // x = v / 255 (ai_runner.py:463) folded into the scale for crops that carry the byte value v itself
        std::vector<float> scale_u8(64);
        for (int o = 0; o < 64; o++) scale_u8[o] = (float)((double)g->data[o] / std::sqrt((double)var->data[o] + 1e-5) / 255.0);
        rc = upload(m, scale_u8, &m->stem_scale_u8); if (rc != PA_OK) return rc;
    }
    // ---- residual stages (execution order per block: conv1, [downsample], conv2)
    const int chans[4] = {64, 128, 256, 512};
    int hin = 32, cin = 64;
    for (int s = 0; s < 4; s++) {
        const int cout = chans[s];
        for (int b = 0; b < 2; b++) {
            const std::string B = P + "layer" + std::to_string(s + 1) + "." + std::to_string(b) + ".";
            const int stride = (s > 0 && b == 0) ? 2 : 1;
            m->convs.push_back(make_conv(B + "conv1.weight", B + "bn1", cin, cout, 3, stride, 1, hin, true));
            if (stride == 2) m->convs.push_back(make_conv(B + "downsample.0.weight", B + "downsample.1", cin, cout, 1, 2, 0, hin, false));
            m->convs.push_back(make_conv(B + "conv2.weight", B + "bn2", cout, cout, 3, 1, 1, hin / stride, true));
            hin /= stride;
            cin = cout;
        }
    }
    for (ConvLayer& L : m->convs) { rc = prepare_conv(m, L); if (rc != PA_OK) return rc; }
    // ---- fc as a 1x1 "conv" over pooled features
    m->fc = make_conv(P + "fc.weight", "", 512, 1000, 1, 1, 0, 1, false);
    m->fc.bias_key = P + "fc.bias";
    {
        // fc.weight is [1000][512]: view as [1000][512][1][1]
        auto it = m->tensors.find(m->fc.w_key);
        if (it == m->tensors.end()) { ctx->last_error = "missing tensor " + m->fc.w_key; return PA_ERR_MISSING_TENSOR; }
        if (it->second.shape.size() == 2) { it->second.shape.push_back(1); it->second.shape.push_back(1); }
    }
    rc = prepare_conv(m, m->fc); if (rc != PA_OK) return rc;
    // ---- temporal Conv1d as per-frame projections: rows n = t*512 + o, K = 1000
    {
        const int S = m->seq;
        const HostTensor* w = get_tensor(m, "cnn1d.0.weight", {512, 1000, S});
        const HostTensor* b = get_tensor(m, "cnn1d.0.bias", {512});
        if (!w || !b) return PA_ERR_MISSING_TENSOR;
        std::vector<float> packed((size_t)S * 512 * 1000);
        for (int o = 0; o < 512; o++)
            for (int i = 0; i < 1000; i++)
                for (int t = 0; t < S; t++) packed[((size_t)t * 512 + o) * 1000 + i] = w->data[((size_t)o * 1000 + i) * S + t];
        std::vector<uint16_t> hi, lo;
        split_weights(packed, hi, lo, prec_f16(precision));
        ConvLayer& L = m->proj;
        L = make_conv("", "", 1000, S * 512, 1, 1, 0, 1, false);
        L.k_total = 1000;
        rc = upload(m, hi, (uint16_t**)&L.w_hi); if (rc != PA_OK) return rc;
        if (prec_split_w(precision)) { rc = upload(m, lo, (uint16_t**)&L.w_lo); if (rc != PA_OK) return rc; }
        rc = upload(m, b->data, &m->b1d); if (rc != PA_OK) return rc;
    }
    // ---- classifier MLP (fp32 SIMT in the head kernel), weights transposed for coalesced reads
    {
        const int A = m->n_actions;
        const HostTensor *w1 = get_tensor(m, "classifier.0.weight", {128, 512}), *b1 = get_tensor(m, "classifier.0.bias", {128}),
                         *w2 = get_tensor(m, "classifier.2.weight", {A, 128}), *b2 = get_tensor(m, "classifier.2.bias", {A});
        if (!w1 || !b1 || !w2 || !b2) return PA_ERR_MISSING_TENSOR;
        std::vector<float> w1t((size_t)512 * 128), w2t((size_t)128 * A);
        for (int o = 0; o < 128; o++) for (int i = 0; i < 512; i++) w1t[(size_t)i * 128 + o] = w1->data[(size_t)o * 512 + i];
        for (int o = 0; o < A; o++) for (int i = 0; i < 128; i++) w2t[(size_t)i * A + o] = w2->data[(size_t)o * 128 + i];
        rc = upload(m, w1t, &m->w1t); if (rc != PA_OK) return rc;
        rc = upload(m, b1->data, &m->b1); if (rc != PA_OK) return rc;
        rc = upload(m, w2t, &m->w2t); if (rc != PA_OK) return rc;
        rc = upload(m, b2->data, &m->b2); if (rc != PA_OK) return rc;
    }
    m->ready = true;
    return PA_OK;
}

extern "C" int pa_model_precision(const pa_model* m) { return m ? m->precision : -1; }

// activation arena: four [n][32][32][64]-sized buffers (the stem kernel emits the pooled map directly), per plane
static const size_t kSmallElems = (size_t)32 * 32 * 64;  // per crop
static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static size_t features_ws_bytes(const pa_model* m, int n) {
    const int planes = prec_split(m->precision) ? 2 : 1;
    size_t per_plane = 4 * align256(kSmallElems * n * 2) + align256((size_t)n * 512 * 2);
    return per_plane * planes + 1024;
}
static size_t head_ws_bytes(const pa_model* m, int n_feat) {
    return 2 * align256((size_t)n_feat * 1000 * 2) + align256((size_t)n_feat * m->seq * 512 * 4) + 1024;
}

extern "C" int pa_model_workspace_bytes(const pa_model* m, int n_crops, size_t* bytes) {
    if (!m || !bytes || n_crops <= 0) return PA_ERR_INVALID_ARG;
    if (m->precision < 0) return PA_ERR_NOT_READY;
    size_t a = features_ws_bytes(m, n_crops), b = head_ws_bytes(m, n_crops);
    *bytes = a > b ? a : b;
    return PA_OK;
}

// ---- TMA descriptors
static int make_map_a(pa_ctx* ctx, CUtensorMap* map, const bf16* base, int C, int W, int H, int N, int parity_stride,
                      int py, int px, int wt, int ht, int nt) {
    // parity_stride == 1: plain NHWC view. == 2: every other row / column starting at (py, px).
    const int s = parity_stride;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)(W / s), (cuuint64_t)(H / s), (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)s * C * 2, (cuuint64_t)s * W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)wt, (cuuint32_t)ht, (cuuint32_t)nt};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const bf16* p = base + ((size_t)py * W + px) * C;
    CUresult r = ctx->encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)p, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ctx->last_error = "cuTensorMapEncodeTiled(A) failed: " + std::to_string((int)r); return PA_ERR_CUDA; }
    return PA_OK;
}
static int make_map_b(pa_ctx* ctx, CUtensorMap* map, const bf16* base, int k_total, int cout, int block_n) {
    cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)cout};
    cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ctx->last_error = "cuTensorMapEncodeTiled(B) failed: " + std::to_string((int)r); return PA_ERR_CUDA; }
    return PA_OK;
}

// conv1 A operand: overlapping 64-byte windows of the padded NHWC4P crop rows of one parity
static int make_map_c1a(pa_ctx* ctx, CUtensorMap* map, const bf16* base, int n, int py) {
    const cuuint64_t row_bytes = 136 * 4 * 2;
    cuuint64_t dims[4] = {32, 64, 64, (cuuint64_t)n};
    cuuint64_t strides[3] = {16, 2 * row_bytes, 128 * row_bytes};
    cuuint32_t box[4] = {32, 64, 2, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const bf16* p = base + (size_t)py * 136 * 4;
    CUresult r = ctx->encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)p, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ctx->last_error = "cuTensorMapEncodeTiled(conv1 A) failed: " + std::to_string((int)r); return PA_ERR_CUDA; }
    return PA_OK;
}
static int make_map_c1b(pa_ctx* ctx, CUtensorMap* map, const bf16* w) {
    cuuint64_t dims[2] = {256, 64};
    cuuint64_t strides[1] = {512};
    cuuint32_t box[2] = {32, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ctx->last_error = "cuTensorMapEncodeTiled(conv1 B) failed: " + std::to_string((int)r); return PA_ERR_CUDA; }
    return PA_OK;
}
static int plan_stem(pa_ctx* ctx, PlanOp& op, const bf16* in_hi, const bf16* in_lo, const bf16* w_hi, const bf16* w_lo,
                     const float* scale, const float* shift, bf16* out_hi, bf16* out_lo, int n, int f16) {
    memset(&op, 0, sizeof(op));
    op.kind = 0;
    snprintf(op.name, sizeof(op.name), "conv1_stem+maxpool");
    for (int py = 0; py < 2; py++) {
        int rc = make_map_c1a(ctx, &op.c1maps.a[0][py], in_hi, n, py);
        if (rc != PA_OK) return rc;
        if (in_lo) { rc = make_map_c1a(ctx, &op.c1maps.a[1][py], in_lo, n, py); if (rc != PA_OK) return rc; }
    }
    int rc = make_map_c1b(ctx, &op.c1maps.b[0], w_hi);
    if (rc != PA_OK) return rc;
    if (w_lo) { rc = make_map_c1b(ctx, &op.c1maps.b[1], w_lo); if (rc != PA_OK) return rc; }
    op.c1.scale = scale; op.c1.shift = shift;
    op.c1.out_hi = out_hi; op.c1.out_lo = out_lo;
    op.c1.n_crops = n; op.c1.f16 = f16;
    op.c1.split_a = in_lo ? 1 : 0; op.c1.split_w = w_lo ? 1 : 0;
    return PA_OK;
}

struct Act {  // an activation tensor: hi plane (+ lo plane)
    bf16* hi = nullptr;
    bf16* lo = nullptr;
};

// Build the launch record of one conv/linear layer. `hin` = input spatial size, N images.
static int plan_conv(pa_model* m, const ConvLayer& L, const Act& in, int N, const Act* res, const Act* out, float* out_f32, PlanOp& op) {
    pa_ctx* ctx = m->ctx;
    memset(&op.maps, 0, sizeof(op.maps));
    op.kind = 2;
    {
        std::string nm = L.w_key.empty() ? std::string("temporal_proj") : L.w_key;
        const std::string pre = "cnn2d.", suf = ".weight";
        if (nm.rfind(pre, 0) == 0) nm = nm.substr(pre.size());
        if (nm.size() > suf.size() && nm.compare(nm.size() - suf.size(), suf.size(), suf) == 0) nm = nm.substr(0, nm.size() - suf.size());
        snprintf(op.name, sizeof(op.name), "conv_gemm:%s", nm.c_str());
    }
    const int n_a = in.lo ? 2 : 1, n_b = L.w_lo ? 2 : 1;
    const int hout = L.hin / L.stride;
    int wt, ht, nt;
    if (hout >= 32) { wt = hout; ht = 128 / hout; nt = 1; }
    else if (hout == 16) { wt = 16; ht = 8; nt = 1; }
    else if (hout == 8) { wt = 8; ht = 8; nt = 2; }
    else if (hout == 4) { wt = 4; ht = 4; nt = 8; }
    else if (hout == 1) { wt = 1; ht = 1; nt = 128; }
    else return PA_ERR_UNSUPPORTED;
    // stride-1 3x3 layers on full-width tiles: patch staging (one box of ht+2 rows per horizontal shift)
    bool patch = (L.k == 3 && L.stride == 1 && L.pad == 1 && (hout == 32 || hout == 16) && (n_b == 1 || n_a == 2) && (L.cin % 64) == 0 &&
                  L.cout == L.block_n && !exp_flag("PA_NO_PATCH"));
    int patch_stages = 0;
    bool patch_pair = false;
    if (patch) {
        const int64_t m_tiles_ = ((int64_t)N * hout * hout + 127) / 128;
        if (m_tiles_ >= 2 && !exp_flag("PA_NO_PAIR")) {   // CTA pair: half of every weight tile per SM (conv_patch2.cu)
            patch_stages = conv_patch2_plan(L.block_n, n_a, n_b, wt, ht, L.cin / 64, &op.patch_wres, &op.patch_smem);
            patch_pair = patch_stages >= 2;
        }
        if (!patch_pair) patch_stages = n_b == 1 ? conv_patch_plan(L.block_n, n_a, wt, ht, L.cin / 64, &op.patch_wres, &op.patch_smem) : 0;   // split weights: pair kernel only
        if (patch_stages < 2) patch = false;
    }
    for (int pl = 0; pl < n_a; pl++) {
        const bf16* base = pl == 0 ? in.hi : in.lo;
        if (patch) {
            int rc = make_map_a(ctx, &op.maps.a[pl][0], base, L.cin, L.hin, L.hin, N, 1, 0, 0, wt, ht + 2, nt);
            if (rc != PA_OK) return rc;
        } else if (L.stride == 1) {
            int rc = make_map_a(ctx, &op.maps.a[pl][0], base, L.cin, L.hin, L.hin, N, 1, 0, 0, wt, ht, nt);
            if (rc != PA_OK) return rc;
        } else {
            for (int q = 0; q < 4; q++) {
                int rc = make_map_a(ctx, &op.maps.a[pl][q], base, L.cin, L.hin, L.hin, N, 2, q >> 1, q & 1, wt, ht, nt);
                if (rc != PA_OK) return rc;
            }
        }
    }
    int rc = make_map_b(ctx, &op.maps.b[0], L.w_hi, L.k_total, L.cout, L.block_n);
    if (rc != PA_OK) return rc;
    if (n_b == 2) { rc = make_map_b(ctx, &op.maps.b[1], L.w_lo, L.k_total, L.cout, L.block_n); if (rc != PA_OK) return rc; }
    ConvArgs& a = op.args;
    memset(&a, 0, sizeof(a));
    a.m_total = N * hout * hout;
    a.m_tiles = (a.m_total + 127) / 128;
    a.n_tiles = (L.cout + L.block_n - 1) / L.block_n;
    a.cout = L.cout;
    a.taps_h = L.k; a.taps_w = L.k; a.stride = L.stride; a.pad = L.pad;
    a.kb_per_tap = (L.cin + 63) / 64;
    a.k_per_tap = L.cin;
    a.ho = hout; a.wo = hout;
    op.block_n = L.block_n; op.n_a = n_a; op.n_b = n_b;
    // wide layers: a CTA pair per 256-row tile, each CTA staging half of the weight tile (conv_gemm2.cu)
    const bool pair = !patch && L.block_n == 256 && (n_b == 1 || n_a == 2) && a.m_tiles >= 2 && !exp_flag("PA_NO_PAIR");
    if (pair || (patch && patch_pair)) {
        rc = make_map_b(ctx, &op.maps.bh[0], L.w_hi, L.k_total, L.cout, L.block_n / 2);
        if (rc != PA_OK) return rc;
        if (n_b == 2) { rc = make_map_b(ctx, &op.maps.bh[1], L.w_lo, L.k_total, L.cout, L.block_n / 2); if (rc != PA_OK) return rc; }
    }
    a.n_b = n_b;
    // 1-CTA kernel: separate passes over K for the residual products only where the accumulator's truncation bias (linear in
    // K) matters or the weights are split as well; the pair kernels always run them as passes (it costs them nothing)
    a.multipass = (n_a + n_b > 2) && (n_b == 2 || L.k_total >= 1152) ? 1 : 0;
    a.num_stages = patch ? patch_stages : (pair ? conv_gemm2_pick_stages(L.block_n, n_a) : conv_gemm_pick_stages(L.block_n, n_a, n_b, a.multipass != 0));
    if (a.num_stages < 2) return PA_ERR_UNSUPPORTED;
    if (patch) { op.kind = patch_pair ? 6 : 4; op.patch_ht = ht; }
    if (pair) op.kind = 5;
    a.scale = L.scale; a.shift = L.shift;
    a.res_hi = res ? res->hi : nullptr;
    a.res_lo = res ? res->lo : nullptr;
    a.relu = L.relu ? 1 : 0;
    a.out_hi = out ? out->hi : nullptr;
    a.out_lo = out ? out->lo : nullptr;
    a.out_f32 = out_f32;
    a.f16 = prec_f16(m->precision) ? 1 : 0;
    return PA_OK;
}

// launch one planned convolution (kinds 2: 1-CTA GEMM, 4: patch mode, 5: CTA-pair GEMM, 6: CTA-pair patch mode)
static int launch_conv_op(pa_ctx* ctx, const PlanOp& op, cudaStream_t st) {
    if (op.kind == 4) return launch_conv_patch(op.maps, op.args, op.block_n, op.n_a, op.patch_ht, op.patch_wres, op.patch_smem, ctx->num_sms, st);
    if (op.kind == 5) return launch_conv_gemm2(op.maps, op.args, op.block_n, op.n_a, ctx->num_sms, st);
    if (op.kind == 6) return launch_conv_patch2(op.maps, op.args, op.block_n, op.n_a, op.patch_ht, op.patch_wres, op.patch_smem, ctx->num_sms, st);
    return launch_conv_gemm(op.maps, op.args, op.block_n, op.n_a, op.n_b, ctx->num_sms, st);
}

static int build_feature_plan(pa_model* m, const void* crops, int n, float* feat, void* ws, size_t ws_bytes, bool u8) {
    if (features_ws_bytes(m, n) > ws_bytes) return PA_ERR_WORKSPACE;
    const bool split = prec_split(m->precision);
    const int f16 = prec_f16(m->precision) ? 1 : 0;
    uint8_t* p = (uint8_t*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    auto carve = [&](size_t elems) {
        Act a;
        a.hi = (bf16*)p; p += align256(elems * 2);
        if (split) { a.lo = (bf16*)p; p += align256(elems * 2); }
        return a;
    };
    Act sm[4];
    for (int i = 0; i < 4; i++) sm[i] = carve(kSmallElems * n);
    Act pooled = carve((size_t)n * 512);
    m->plan.clear();
    // stem
    {
        PlanOp op;
        const bf16* in_hi = (const bf16*)crops;
        const bf16* in_lo = (split && !u8) ? in_hi + (size_t)n * 128 * 136 * 4 : nullptr;   // byte values are exact: no lo plane
        int rc = plan_stem(m->ctx, op, in_hi, in_lo, m->stem_w_hi, m->stem_w_lo, u8 ? m->stem_scale_u8 : m->stem.scale, m->stem.shift,
                           sm[0].hi, sm[0].lo, n, f16);
        if (rc != PA_OK) return rc;
        m->plan.push_back(op);   // conv1 + BN + ReLU + max-pool in one kernel
    }
    int x = 0;  // index of the buffer holding the block input
    size_t ci = 0;
    for (int s = 0; s < 4; s++) {
        for (int b = 0; b < 2; b++) {
            const bool down = (s > 0 && b == 0);
            const int t = (x + 1) & 3, r = (x + 2) & 3, y = (x + 3) & 3;
            PlanOp op;
            int rc = plan_conv(m, m->convs[ci++], sm[x], n, nullptr, &sm[t], nullptr, op);
            if (rc != PA_OK) return rc;
            m->plan.push_back(op);
            const Act* res = &sm[x];
            if (down) {
                rc = plan_conv(m, m->convs[ci++], sm[x], n, nullptr, &sm[r], nullptr, op);
                if (rc != PA_OK) return rc;
                m->plan.push_back(op);
                res = &sm[r];
            }
            rc = plan_conv(m, m->convs[ci++], sm[t], n, res, &sm[y], nullptr, op);
            if (rc != PA_OK) return rc;
            m->plan.push_back(op);
            x = y;
        }
    }
    {
        PlanOp op; memset(&op, 0, sizeof(op));
        op.kind = 3;
        snprintf(op.name, sizeof(op.name), "avgpool");
        op.pin_hi = sm[x].hi; op.pin_lo = sm[x].lo; op.pout_hi = pooled.hi; op.pout_lo = pooled.lo;
        op.pn = n; op.ph = 16; op.pc = 512; op.pf16 = f16;
        m->plan.push_back(op);
    }
    {
        PlanOp op;
        int rc = plan_conv(m, m->fc, pooled, n, nullptr, nullptr, feat, op);
        if (rc != PA_OK) return rc;
        m->plan.push_back(op);
    }
    m->plan_crops = crops; m->plan_ws = ws; m->plan_feat = feat; m->plan_n = n; m->plan_u8 = u8;
    return PA_OK;
}

static int features_impl(pa_model* m, const void* crops, int n_crops, float* feat, void* workspace, size_t workspace_bytes, void* stream, bool u8) {
    if (!m || !crops || !feat || !workspace || n_crops <= 0) return PA_ERR_INVALID_ARG;
    if (!m->ready || m->arch != 0) return PA_ERR_NOT_READY;
    pa_ctx* ctx = m->ctx;
    cudaStream_t st = (cudaStream_t)stream;
    if (m->plan_n != n_crops || m->plan_crops != crops || m->plan_ws != workspace || m->plan_feat != feat || m->plan_u8 != u8) {
        m->plan_n = -1;
        int rc = build_feature_plan(m, crops, n_crops, feat, workspace, workspace_bytes, u8);
        if (rc != PA_OK) return rc;
    }
    for (const PlanOp& op : m->plan) {
        int rc = PA_OK;
        ProfSpan sp(ctx, op.name, st);
        switch (op.kind) {
            case 0: rc = launch_conv1(op.c1maps, op.c1, ctx->num_sms, st); break;
            case 2: case 4: case 5: case 6: rc = launch_conv_op(ctx, op, st); break;
            case 3: rc = launch_avgpool(op.pin_hi, op.pin_lo, op.pout_hi, op.pout_lo, op.pn, op.ph, op.pc, op.pf16, st); break;
        }
        if (rc != PA_OK) return rc == PA_ERR_CUDA ? cuda_fail(ctx, cudaGetLastError(), "feature kernel launch") : rc;
        ctx->launches += 1;
    }
    return PA_OK;
}

extern "C" int pa_features(pa_model* m, const void* crops, int n_crops, float* feat, void* workspace, size_t workspace_bytes, void* stream) {
    return features_impl(m, crops, n_crops, feat, workspace, workspace_bytes, stream, false);
}
extern "C" int pa_features_u8(pa_model* m, const void* crops, int n_crops, float* feat, void* workspace, size_t workspace_bytes, void* stream) {
    return features_impl(m, crops, n_crops, feat, workspace, workspace_bytes, stream, true);
}

extern "C" int pa_head(pa_model* m, const float* feat, int n_feat, const int32_t* feat_status, const int32_t* win_idx, int n_win,
                       float* logp, int32_t* label, float* conf, void* workspace, size_t workspace_bytes, void* stream) {
    if (!m || !feat || !win_idx || !logp || !label || !conf || !workspace || n_feat <= 0 || n_win <= 0) return PA_ERR_INVALID_ARG;
    if (!m->ready) return PA_ERR_NOT_READY;
    pa_ctx* ctx = m->ctx;
    cudaStream_t st = (cudaStream_t)stream;
    if (head_ws_bytes(m, n_feat) > workspace_bytes) return PA_ERR_WORKSPACE;
    const bool split = prec_split(m->precision);
    uint8_t* p = (uint8_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    Act fb;
    fb.hi = (bf16*)p; p += align256((size_t)n_feat * 1000 * 2);
    fb.lo = split ? (bf16*)p : nullptr; p += align256((size_t)n_feat * 1000 * 2);
    float* proj = (float*)p;
    if (m->hplan_n != n_feat || m->hplan_feat != feat || m->hplan_ws != workspace) {
        int rc = plan_conv(m, m->proj, fb, n_feat, nullptr, nullptr, proj, m->head_gemm);
        if (rc != PA_OK) return rc;
        m->head_gemm.args.scale = nullptr; m->head_gemm.args.shift = nullptr;
        m->hplan_n = n_feat; m->hplan_feat = feat; m->hplan_ws = workspace;
    }
    int rc;
    {
        ProfSpan sp(ctx, "feat_to_bf16", st);
        rc = launch_split_f32(feat, fb.hi, fb.lo, (int64_t)n_feat * 1000, prec_f16(m->precision) ? 1 : 0, st);
    }
    if (rc != PA_OK) return cuda_fail(ctx, cudaGetLastError(), "split launch");
    const PlanOp& g = m->head_gemm;
    {
        ProfSpan sp(ctx, g.name, st);
        rc = launch_conv_op(ctx, g, st);
    }
    if (rc != PA_OK) return rc == PA_ERR_CUDA ? cuda_fail(ctx, cudaGetLastError(), "projection launch") : rc;
    HeadArgs h;
    h.proj = proj; h.feat_status = feat_status; h.win_idx = win_idx; h.n_win = n_win; h.n_feat = n_feat; h.seq = m->seq; h.n_actions = m->n_actions;
    h.b1d = m->b1d; h.w1t = m->w1t; h.b1 = m->b1; h.w2t = m->w2t; h.b2 = m->b2;
    h.logp = logp; h.label = label; h.conf = conf;
    {
        ProfSpan sp(ctx, "head_mlp_softmax", st);
        rc = launch_head(h, st);
    }
    if (rc != PA_OK) return rc == PA_ERR_CUDA ? cuda_fail(ctx, cudaGetLastError(), "head launch") : rc;
    ctx->launches += 3;
    return PA_OK;
}

// ------------------------------------------------------------------------------------------------ single-layer entry points
// One convolution / the stem on caller-provided NHWC bf16 activations and host fp32 weights.
// Used by the layer-level parity tests (tests/test_gpu_layers.py); synchronous, allocates scratch.
// flags: bit 0 = split the weights too (3 MMAs), bit 1 = IEEE-half operands instead of bf16
static int layer_model(pa_ctx* ctx, int flags, pa_model& m) {
    m.ctx = ctx;
    const bool f16 = (flags & 2) != 0, sw = (flags & 1) != 0;
    m.precision = f16 ? (sw ? PA_PREC_F16X3 : PA_PREC_F16) : (sw ? PA_PREC_BF16X3 : PA_PREC_BF16);
    return PA_OK;
}

extern "C" int pa_conv2d(pa_ctx* ctx, const void* in_hi, const void* in_lo, int n, int hin, int cin, const float* w_host,
                         int cout, int k, int stride, int pad, const float* scale_host, const float* shift_host,
                         const void* res_hi, const void* res_lo, int relu, void* out_hi, void* out_lo, float* out_f32,
                         int split_w, void* stream) {
    if (!ctx || !in_hi || !w_host || n <= 0 || (!out_hi && !out_f32)) return PA_ERR_INVALID_ARG;
    if ((split_w & 1) && !in_lo) return PA_ERR_INVALID_ARG;
    pa_model m;
    layer_model(ctx, split_w, m);
    HostTensor t;
    t.shape = {cout, cin, k, k};
    t.data.assign(w_host, w_host + (size_t)cout * cin * k * k);
    m.tensors["w"] = t;
    ConvLayer L = make_conv("w", "", cin, cout, k, stride, pad, hin, relu != 0);
    int rc = prepare_conv(&m, L);
    if (rc == PA_OK) {
        if (scale_host) cudaMemcpy(L.scale, scale_host, cout * sizeof(float), cudaMemcpyHostToDevice);
        if (shift_host) cudaMemcpy(L.shift, shift_host, cout * sizeof(float), cudaMemcpyHostToDevice);
        Act in, res, out;
        in.hi = (bf16*)in_hi; in.lo = (bf16*)in_lo;
        res.hi = (bf16*)res_hi; res.lo = (bf16*)res_lo;
        out.hi = (bf16*)out_hi; out.lo = (bf16*)out_lo;
        PlanOp op;
        rc = plan_conv(&m, L, in, n, res_hi ? &res : nullptr, out_hi ? &out : nullptr, out_f32, op);
        if (rc == PA_OK) rc = launch_conv_op(ctx, op, (cudaStream_t)stream);
        if (rc == PA_OK) {
            ctx->launches += 1;
            cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
            if (e != cudaSuccess) rc = cuda_fail(ctx, e, "pa_conv2d");
        } else if (rc == PA_ERR_CUDA) {
            cuda_fail(ctx, cudaGetLastError(), "pa_conv2d launch");
        }
    }
    for (void* p : m.dev_allocs) cudaFree(p);
    m.dev_allocs.clear();
    return rc;
}

extern "C" int pa_stem(pa_ctx* ctx, const void* in_hi, const void* in_lo, int n, const float* w_host /*[64][3][7][7]*/,
                       const float* scale_host, const float* shift_host, void* out_hi, void* out_lo, int split_w, void* stream) {
    if (!ctx || !in_hi || !w_host || !scale_host || !shift_host || !out_hi || n <= 0) return PA_ERR_INVALID_ARG;
    if ((split_w & 1) && !in_lo) return PA_ERR_INVALID_ARG;
    pa_model m;
    layer_model(ctx, split_w, m);
    std::vector<float> packed;
    pack_stem_weights(w_host, packed);
    std::vector<uint16_t> hi, lo;
    split_weights(packed, hi, lo, prec_f16(m.precision));
    std::vector<float> sc(scale_host, scale_host + 64), sh(shift_host, shift_host + 64);
    uint16_t *dw_hi = nullptr, *dw_lo = nullptr;
    int rc = upload(&m, hi, &dw_hi);
    if (rc == PA_OK && (split_w & 1)) rc = upload(&m, lo, &dw_lo);
    float *dsc = nullptr, *dsh = nullptr;
    if (rc == PA_OK) rc = upload(&m, sc, &dsc);
    if (rc == PA_OK) rc = upload(&m, sh, &dsh);
    if (rc == PA_OK) {
        PlanOp op;
        rc = plan_stem(ctx, op, (const bf16*)in_hi, (const bf16*)in_lo, (const bf16*)dw_hi, (const bf16*)dw_lo, dsc, dsh,
                       (bf16*)out_hi, (bf16*)out_lo, n, prec_f16(m.precision) ? 1 : 0);
        if (rc == PA_OK) rc = launch_conv1(op.c1maps, op.c1, ctx->num_sms, (cudaStream_t)stream);
        if (rc == PA_OK) {
            ctx->launches += 1;
            cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
            if (e != cudaSuccess) rc = cuda_fail(ctx, e, "pa_stem");
        } else if (rc == PA_ERR_CUDA && ctx->last_error.empty()) {
            cuda_fail(ctx, cudaGetLastError(), "pa_stem launch");
        }
    }
    for (void* p : m.dev_allocs) cudaFree(p);
    m.dev_allocs.clear();
    return rc;
}

#include "resformer.inc"
