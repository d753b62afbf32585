// Engine, batches and the C-ABI entry points that touch the GPU.
//
// Host orchestration only: every stage of the network and both decoders run as CUDA
// kernels (kernels_v1.cu, kernels_tc.cu).  There is deliberately no CPU fallback; when
// CUDA is unusable the entry points fail with NULL / NAN / -1 and sb2_last_error().
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <time.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "kernels.h"
#include "sb2_internal.h"

using namespace sb2;

#define CUDA_OK(call)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            sb2_set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__,    \
                          __LINE__, #call);                                                   \
            return -1;                                                                        \
        }                                                                                     \
    } while (0)

// ------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------

struct DevModel {
    bool loaded = false;
    sb2_host_model host{};
    float *d_all = nullptr;
    float *conv_taps = nullptr, *conv_b = nullptr;
    float *iW[SB2_NLAYER]{}, *b[SB2_NLAYER]{}, *sW[SB2_NLAYER]{}, *sW2[SB2_NLAYER]{};
    float *FF_W = nullptr, *FF_b = nullptr;
    float *comb_Wf[2]{}, *comb_Wb[2]{}, *comb_b[2]{};    // raw_r94: feedforward2_tanh layers
    uint8_t *iw_img[SB2_NLAYER]{};       // tensor-core affine: per-layer input-transform image
    uint8_t *d_gemm_img_all = nullptr;
    uint8_t *head_img = nullptr;         // fused output head (1025-state models, K = 96)
};

struct sb2_engine {
    int device = 0;
    DevModel models[SB2_NMODEL + 1];            // [SB2_NMODEL] = the events (LSTM) model of nanonet_posterior
    char weights_dir[1024]{};
    std::atomic<uint64_t> launches{0};
    std::atomic<uint64_t> reallocs{0};          // (re)allocations of batch workspaces: 0 in steady state
    float *flush_buf = nullptr;
    size_t flush_n = 0;
    std::mutex mu;
    int scan_impl = 5;      // gate math of the tcgen05 scan (5 / 2 / 0), or -1 = fp32 CUDA-core scan
    int gemm_impl = 0;
    long long *d_trace = nullptr;   // diagnostic: hand-over timestamps of the scan kernel (SCRAPPIE_B200_TRACE=1)
    int head_exact = 0;     // 1: cephes exp / log in the fused head (bit-level mirror of the reference's maths)
    // idle workspaces of sb2_basecall_batch / sb2_basecall_raw_batch, per model: device buffers, pinned staging and
    // CUDA graphs survive between calls, so the documented drop-in call allocates nothing in steady state
    std::vector<struct sb2_batch *> pool[SB2_NMODEL];
    // largest shape any pooled call of a model has needed so far: pooled workspaces are sized to it, so that after one
    // pass over a mixed workload every workspace fits every batch and nothing is re-allocated any more (cudaFree
    // synchronises the whole device -- with long-read batches in flight on other streams that stalls every caller)
    std::atomic<size_t> hw_reads[SB2_NMODEL]{}, hw_cols[SB2_NMODEL]{}, hw_samples[SB2_NMODEL]{}, hw_bases[SB2_NMODEL]{}, hw_xrows[SB2_NMODEL]{}, hw_stage[SB2_NMODEL]{};
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// NVTX range around a host-side stage (header-only NVTX3: a no-op unless a profiler is attached).  Kernel launches are
// asynchronous, so a range brackets the ENQUEUE of a stage; in an Nsight timeline the kernels of the stage line up
// under it through the CUDA correlation ids.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

static int upload_model(sb2_engine *eng, DevModel *dm) {
    const sb2_host_model &h = dm->host;
    const size_t H = h.H;
    // compact device copies; each tensor starts on a 256-byte boundary
    std::vector<std::pair<float **, std::vector<float>>> items;
    auto compact = [](const sb2_tensor &t) {
        std::vector<float> v((size_t)t.nc * t.nr);
        for (uint32_t c = 0; c < t.nc; c++) memcpy(&v[(size_t)c * t.nr], t.data + (size_t)c * t.stride, t.nr * sizeof(float));
        return v;
    };
    if (h.arch != 2) {   // conv taps: reference stores filter f as a column of winlen*4 floats with the tap at
        // every 4th slot (src/layers.c:155-157); device layout is [tap][filter]
        const size_t NF = h.nfilter;
        std::vector<float> taps((size_t)h.winlen * NF);
        for (uint32_t f = 0; f < NF; f++)
            for (uint32_t k = 0; k < h.winlen; k++) taps[(size_t)k * NF + f] = h.conv_W.data[(size_t)f * h.conv_W.stride + 4 * k];
        items.push_back({&dm->conv_taps, taps});
        items.push_back({&dm->conv_b, compact(h.conv_b)});
    }
    const int nlayer = (h.arch == 0) ? SB2_NLAYER : 4;
    for (int l = 0; l < nlayer; l++) {
        items.push_back({&dm->iW[l], compact(h.iW[l])});
        items.push_back({&dm->b[l], compact(h.b[l])});
        items.push_back({&dm->sW[l], compact(h.sW[l])});
        items.push_back({&dm->sW2[l], compact(h.sW2[l])});
    }
    if (h.arch >= 1)
        for (int i = 0; i < 2; i++) {
            items.push_back({&dm->comb_Wf[i], compact(h.comb_Wf[i])});
            items.push_back({&dm->comb_Wb[i], compact(h.comb_Wb[i])});
            items.push_back({&dm->comb_b[i], compact(h.comb_b[i])});
        }
    items.push_back({&dm->FF_W, compact(h.FF_W)});
    items.push_back({&dm->FF_b, compact(h.FF_b)});
    size_t total = 0;
    for (auto &it : items) total += align_up(it.second.size() * sizeof(float), 256);
    CUDA_OK(cudaSetDevice(eng->device));
    CUDA_OK(cudaMalloc(&dm->d_all, total));
    size_t off = 0;
    for (auto &it : items) {
        float *dst = reinterpret_cast<float *>(reinterpret_cast<char *>(dm->d_all) + off);
        CUDA_OK(cudaMemcpy(dst, it.second.data(), it.second.size() * sizeof(float), cudaMemcpyHostToDevice));
        *it.first = dst;
        off += align_up(it.second.size() * sizeof(float), 256);
    }
    if (h.arch >= 1) {      // raw_r94 / events: the scans read fp32 weights; the odd-shaped affine maps use the fp32 kernel
        dm->loaded = true;
        return 0;
    }
    {   // tensor-core affine images (input transforms)
        const size_t nb = align_up(gemm_image_bytes(3, (int)H, (int)H), 256);
        std::vector<uint8_t> img(nb * SB2_NLAYER, 0);
        for (int l = 0; l < SB2_NLAYER; l++) {
            const std::vector<float> iw = compact(h.iW[l]);
            build_gemm_image(iw.data(), (int)H, 3 * (int)H, (int)H, (int)H, 3, img.data() + nb * l);
        }
        CUDA_OK(cudaMalloc(&dm->d_gemm_img_all, img.size()));
        CUDA_OK(cudaMemcpy(dm->d_gemm_img_all, img.data(), img.size(), cudaMemcpyHostToDevice));
        for (int l = 0; l < SB2_NLAYER; l++) dm->iw_img[l] = dm->d_gemm_img_all + nb * l;
    }
    if (h.head == 0 && h.nstate == 1025 && H == 96) {
        std::vector<uint8_t> img(head_image_bytes((int)H), 0);
        const std::vector<float> ff = compact(h.FF_W);
        build_gemm_image(ff.data(), (int)H, 1024, (int)H, 128, 8, img.data());
        CUDA_OK(cudaMalloc(&dm->head_img, img.size()));
        CUDA_OK(cudaMemcpy(dm->head_img, img.data(), img.size(), cudaMemcpyHostToDevice));
    }
    dm->loaded = true;
    return 0;
}

// Free everything a (possibly partial) model load allocated; the slot can be loaded again afterwards.
static void release_model(sb2_engine *eng, DevModel *dm) {
    cudaSetDevice(eng->device);
    if (dm->d_all) cudaFree(dm->d_all);
    if (dm->d_gemm_img_all) cudaFree(dm->d_gemm_img_all);
    if (dm->head_img) cudaFree(dm->head_img);
    sb2_host_model_free(&dm->host);
    *dm = DevModel{};
}

extern "C" int sb2_engine_load_blob(sb2_engine *eng, enum raw_model_type model, const void *blob, size_t nbytes) {
    if (nullptr == eng || model < 0 || model >= SCRAPPIE_MODEL_INVALID) return -1;
    std::lock_guard<std::mutex> lock(eng->mu);
    DevModel *dm = &eng->models[model];
    if (dm->loaded) return 0;
    if (0 != sb2_host_model_parse(blob, nbytes, &dm->host)) return -1;
    int rc = 0;
    if (dm->host.H != 96 && dm->host.H != 112) {
        sb2_set_error("unsupported GRU width %u", dm->host.H);
        rc = -1;
    }
    if (0 == rc) rc = upload_model(eng, dm);
    if (0 != rc) release_model(eng, dm);             // a retry must not find half a model (or leak it)
    return rc;
}

static DevModel *get_model(sb2_engine *eng, enum raw_model_type model) {
    if (nullptr == eng || model < 0 || model >= SCRAPPIE_MODEL_INVALID) return nullptr;
    DevModel *dm = &eng->models[model];
    {
        std::lock_guard<std::mutex> lock(eng->mu);  // `loaded` is published under the engine's mutex
        if (dm->loaded) return dm;
    }
    char path[1200];
    snprintf(path, sizeof(path), "%s/%s.bin", eng->weights_dir, sb2_model_file_stem(model));
    void *blob = nullptr;
    size_t nbytes = 0;
    if (0 != sb2_read_file(path, &blob, &nbytes)) return nullptr;
    const int rc = sb2_engine_load_blob(eng, model, blob, nbytes);
    free(blob);
    return (0 == rc) ? dm : nullptr;
}

extern "C" sb2_engine *sb2_engine_create(int device, const char *weights_dir) {
    // Streams are mapped onto CUDA_DEVICE_MAX_CONNECTIONS hardware work queues (default 8); streams that share a queue
    // are served in order, so one batch's kernels wait behind another batch's dependent chain (measured: with every
    // batch in flight on its own stream, 32 queues give +7 % end to end on equal-length reads and 1.3-2x on mixed
    // lengths, where a long read's scan blocks its queue for tens of ms).  The variable is read when the process creates
    // its CUDA context: it only takes effect if nothing in the process has used CUDA before (a host application that
    // initialises CUDA itself should export it).  An explicit setting of the caller is kept.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        sb2_set_error("no usable CUDA device (%s); libscrappie_b200 has no CPU fallback",
                      e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        fprintf(stderr, "scrappie_b200: %s\n", sb2_last_error());
        return nullptr;
    }
    if (device < 0 || device >= ndev) {
        sb2_set_error("device %d out of range (%d visible)", device, ndev);
        return nullptr;
    }
    sb2_engine *eng = new sb2_engine();
    eng->device = device;
    if (nullptr != weights_dir) snprintf(eng->weights_dir, sizeof(eng->weights_dir), "%s", weights_dir);
    else if (0 != sb2_default_weights_dir(eng->weights_dir, sizeof(eng->weights_dir))) eng->weights_dir[0] = '\0';
    const char *scan = getenv("SCRAPPIE_B200_SCAN");
    const char *gemm = getenv("SCRAPPIE_B200_GEMM");
    // GRU scan: tcgen05 kernels (weights in TMEM) with gate math 5 = SFU ex2 + Newton-refined reciprocal (default),
    // 2 = polynomial exp2, 0 = cephes-identical gates; -1 = fp32 CUDA cores (debug cross-check).  Read once per engine.
    eng->scan_impl = 5;
    if (scan && (0 == strcmp(scan, "poly") || 0 == strcmp(scan, "v4_poly"))) eng->scan_impl = 2;
    if (scan && (0 == strcmp(scan, "cephes") || 0 == strcmp(scan, "v4_cephes"))) eng->scan_impl = 0;
    if (scan && 0 == strcmp(scan, "ffma")) eng->scan_impl = -1;
    eng->gemm_impl = (gemm && 0 == strcmp(gemm, "ffma")) ? 0 : 1;    // 1 = tcgen05 (default), 0 = fp32 CUDA cores
    const char *headm = getenv("SCRAPPIE_B200_HEAD");
    eng->head_exact = (headm && 0 == strcmp(headm, "exact")) ? 1 : 0;
    if (cudaSetDevice(device) != cudaSuccess) { delete eng; return nullptr; }
    // dynamic shared-memory limits are a per-device attribute of each kernel: set them for THIS device now, so that
    // no launch path ever configures anything (first use may be under stream capture, or from several host threads)
    if (0 != configure_scan_kernels() || 0 != configure_gemm_kernels() || 0 != configure_v1_kernels()) {
        sb2_set_error("device %d: cannot configure kernels (%s); sm_100a required", device, cudaGetErrorString(cudaGetLastError()));
        fprintf(stderr, "scrappie_b200: %s\n", sb2_last_error());
        delete eng;
        return nullptr;
    }
    {   // stream-ordered allocator: keep freed workspace memory in the device's pool instead of returning it to the driver
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep_all = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep_all);
        }
        cudaGetLastError();
    }
    if (getenv("SCRAPPIE_B200_TRACE") && cudaMalloc(&eng->d_trace, 512 * sizeof(long long)) == cudaSuccess)
        cudaMemset(eng->d_trace, 0, 512 * sizeof(long long));
    return eng;
}

extern "C" void sb2_engine_destroy(sb2_engine *eng) {
    if (nullptr == eng) return;
    cudaSetDevice(eng->device);
    for (auto &free_list : eng->pool) {
        for (sb2_batch *b : free_list) sb2_batch_destroy(b);
        free_list.clear();
    }
    for (auto &dm : eng->models) release_model(eng, &dm);
    if (eng->flush_buf) cudaFree(eng->flush_buf);
    if (eng->d_trace) cudaFree(eng->d_trace);
    delete eng;
}

// Diagnostic: clock64() stamps recorded by CTA 0 of the layer-2 scan (steps 100..103, 16 slots each).
extern "C" int sb2_engine_read_trace(sb2_engine *eng, long long *out, int n) {
    if (nullptr == eng || nullptr == eng->d_trace || nullptr == out) return -1;
    CUDA_OK(cudaSetDevice(eng->device));
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpy(out, eng->d_trace, sizeof(long long) * (size_t)std::min(n, 512), cudaMemcpyDeviceToHost));
    return 0;
}

// Number of times a batch workspace (device buffers, pinned staging, base-string area) had to be (re)allocated: grows
// while the pool warms up, stays constant in steady state.
extern "C" uint64_t sb2_engine_realloc_count(const sb2_engine *eng) { return eng ? eng->reallocs.load() : 0; }

extern "C" uint64_t sb2_engine_launch_count(const sb2_engine *eng) { return eng ? eng->launches.load() : 0; }

extern "C" sb2_params sb2_default_params(void) {
    sb2_params p;
    p.min_prob = 1e-5f; p.tempW = 1.0f; p.tempb = 1.0f;
    p.stay_pen = 0.0f; p.skip_pen = 0.0f; p.local_pen = 2.0f;
    p.allow_slip = 0;
    p.homopolymer = HOMOPOLYMER_MEAN;
    return p;
}

extern "C" void *sb2_host_alloc_pinned(size_t nbytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, nbytes) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void sb2_host_free_pinned(void *p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------------------------
// batch
// ------------------------------------------------------------------------------------

// stage order of one pass: conv, (affine_l, scan_l) x 5, head GEMM, head finish, decode
enum { ST_CONV = 0, ST_HEAD = 11, ST_FINISH = 12, ST_DECODE = 13, ST_COUNT = 14 };
static inline int ST_AFFINE(int l) { return 1 + 2 * l; }
static inline int ST_SCAN(int l) { return 2 + 2 * l; }

struct sb2_batch {
    sb2_engine *eng = nullptr;
    enum raw_model_type model_type = SCRAPPIE_MODEL_INVALID;
    DevModel *m = nullptr;
    int nread = 0, total_cols = 0, max_cols = 0;
    int64_t total_samples = 0;
    std::vector<int> nsample, nblock, col_off;
    std::vector<int64_t> samp_off;
    // device
    float *d_raw = nullptr, *d_X[2]{}, *d_Xin = nullptr, *d_post = nullptr, *d_score = nullptr, *d_layers = nullptr;
    float *d_Xin2 = nullptr, *d_FF = nullptr;          // raw_r94 only
    float *d_bprob = nullptr;                           // posterior_crf output, (total_cols + nread) x 8, on first use
    int *d_nsample = nullptr, *d_nblock = nullptr, *d_coloff = nullptr, *d_tbE = nullptr, *d_path = nullptr;
    int64_t *d_sampoff = nullptr;
    uint8_t *d_tb = nullptr;
    sb2_conv_tail *d_tails = nullptr;
    int final_x = 0;
    bool keep_layers = false;
    bool timing = false;
    // persistent buffers of the basecall path (allocated on first use, never in the hot loop)
    int *h_paths = nullptr;          // pinned: total_cols + nread ints
    float *h_scores = nullptr;       // pinned: nread
    int *h_gidx = nullptr;           // pinned: 2 ints (column, state) per gathered posterior entry
    float *h_gval = nullptr;         // pinned
    int *d_gidx = nullptr;
    float *d_gval = nullptr;
    size_t gcap = 0;                 // capacity in entries
    // device-side finishing (homopolymer + overlapper on the GPU): only base strings come back
    int *d_path2 = nullptr, *d_nbase = nullptr, *h_nbase = nullptr;
    char *d_bases = nullptr, *h_bases = nullptr;
    int bases_stride = 0;
    // forward + decode captured once as a CUDA graph and replayed (13 launches -> 1)
    cudaGraphExec_t graph = nullptr;
    sb2_params graph_params{};
    uint64_t graph_launches = 0;
    int eager_runs = 0;
    // capacities of the device buffers: a pooled workspace is re-shaped for every call and only ever grows
    size_t cap_reads = 0, cap_cols = 0, cap_samples = 0, cap_bases = 0, cap_finish_cols = 0, cap_finish_reads = 0;
    bool pooled = false;
    uint8_t *h_meta = nullptr, *d_meta = nullptr;       // nsample | nblock | col_off | samp_off | conv tails: one H2D copy
    size_t meta_bytes = 0;
    float *h_stage = nullptr;                           // pinned staging of the signals in the padded layout
    size_t stage_cap = 0;
    int64_t *d_src = nullptr;                           // raw-signal basecall: source offset of every kept read
    // Scan-ordered Xin (kernels_tc.cu): reads are grouped `rpg` at a time; group g owns rows [xgrp[g], xgrp[g] + rpg *
    // Tmax_g) of Xin, step-major.  d_xrow[0 / 1][row] = input column of Xin row `row` for forward / backward layers.
    int rpg = 4;
    bool xil = false;
    // Column stride of the posterior ON THE DEVICE.  The reference's stride (4 * ceil(nstate / 4) = 1028 floats = 4112
    // bytes) leaves every other column misaligned to the 128-byte lines the head's stores and the decoder's loads move;
    // the 1025-state models therefore use 1056 floats (33 lines) in HBM and the download re-strides to the reference's.
    int pstride = 0;
    std::vector<long long> xgrp;
    size_t xrows = 0, cap_xrows = 0;                    // rows of Xin the layout needs / the buffer holds
    long long *d_xgrp = nullptr;
    int *d_xrow[2]{};
    int graph_dims[3]{};                                // (nread, total_cols, max_cols) the captured graph was made for
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[ST_COUNT + 1]{};
    cudaEvent_t ev_done = nullptr;          // blocking-sync event the basecall path waits on
    float stage_ms[ST_COUNT]{};
    float stage_at[ST_COUNT + 1]{};          // start of each stage relative to the first batch of a timed multi-batch run
    BatchDims dims{};
};

template <typename T>
static int dev_alloc(T **p, size_t n) {
    CUDA_OK(cudaMalloc(reinterpret_cast<void **>(p), (n ? n : 1) * sizeof(T)));
    return 0;
}

static void batch_free_device(sb2_batch *b) {
    void *ptrs[] = {b->d_raw, b->d_X[0], b->d_X[1], b->d_Xin, b->d_post, b->d_score, b->d_layers, b->d_tbE, b->d_path,
                    b->d_tb, b->d_meta, b->d_path2, b->d_nbase, b->d_bases, b->d_bprob, b->d_Xin2, b->d_FF, b->d_gidx,
                    b->d_gval, b->d_src, b->d_xrow[0], b->d_xrow[1]};
    for (void *p : ptrs) if (p) cudaFreeAsync(p, b->stream);
    b->d_xrow[0] = b->d_xrow[1] = nullptr; b->d_xgrp = nullptr;
    b->d_raw = b->d_X[0] = b->d_X[1] = b->d_Xin = b->d_post = b->d_score = b->d_layers = nullptr;
    b->d_tbE = b->d_path = nullptr; b->d_tb = nullptr; b->d_meta = nullptr; b->d_path2 = b->d_nbase = nullptr;
    b->d_bases = nullptr; b->d_bprob = b->d_Xin2 = b->d_FF = nullptr; b->d_gidx = nullptr; b->d_gval = nullptr; b->d_src = nullptr;
    void *hptrs[] = {b->h_paths, b->h_scores, b->h_gidx, b->h_gval, b->h_nbase, b->h_bases, b->h_meta};
    for (void *p : hptrs) if (p) cudaFreeHost(p);
    b->h_paths = nullptr; b->h_scores = nullptr; b->h_gidx = nullptr; b->h_gval = nullptr; b->h_nbase = nullptr;
    b->h_bases = nullptr; b->h_meta = nullptr;
    if (b->graph) { cudaGraphExecDestroy(b->graph); b->graph = nullptr; }
    b->cap_reads = b->cap_cols = b->cap_samples = b->cap_bases = b->cap_finish_cols = b->cap_finish_reads = b->cap_xrows = 0;
    b->gcap = 0;
    b->eager_runs = 0;
}

// Fault injection (sb2_debug_fail_alloc): countdown to the workspace allocation that is made to fail; < 0 = off.
static std::atomic<long> g_fault_countdown{-1};
extern "C" void sb2_debug_fail_alloc(long nth) { g_fault_countdown.store(nth); }
static bool fault_now() {
    if (g_fault_countdown.load(std::memory_order_relaxed) < 0) return false;
    return g_fault_countdown.fetch_sub(1) == 0;
}
#define FAULT_POINT()                                                                             \
    do {                                                                                          \
        if (fault_now()) { sb2_set_error("injected allocation failure (%s:%d)", __FILE__, __LINE__); return -1; } \
    } while (0)

// pinned host memory of a batch workspace
template <typename T>
static int pinned_alloc(T **p, size_t n) {
    *p = nullptr;
    FAULT_POINT();
    CUDA_OK(cudaMallocHost(reinterpret_cast<void **>(p), (n ? n : 1) * sizeof(T)));
    return 0;
}

// Batch buffers come from the device's stream-ordered memory pool (cudaMallocAsync / cudaFreeAsync on the batch's own
// stream; the pool keeps what is freed, see sb2_engine_create): re-sizing a workspace then never synchronises the
// device.  cudaFree would -- it waits for every kernel in flight on every stream, and with long-read batches of other
// callers running that is hundreds of milliseconds per freed buffer.
template <typename T>
static int batch_alloc(sb2_batch *b, T **p, size_t n) {
    *p = nullptr;
    FAULT_POINT();
    CUDA_OK(cudaMallocAsync(reinterpret_cast<void **>(p), (n ? n : 1) * sizeof(T), b->stream));
    return 0;
}

extern "C" void sb2_batch_destroy(sb2_batch *b) {
    if (nullptr == b) return;
    cudaSetDevice(b->eng->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    batch_free_device(b);
    if (b->stream) cudaStreamSynchronize(b->stream);    // the stream-ordered frees above
    if (b->h_stage) cudaFreeHost(b->h_stage);
    for (auto &e : b->ev) if (e) cudaEventDestroy(e);
    if (b->ev_done) cudaEventDestroy(b->ev_done);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

// Host-side layout of a batch: offsets of every read in the sample and column dimensions, conv tail plans.
static int batch_layout(sb2_batch *b, const size_t *nsample, size_t nread, std::vector<sb2_conv_tail> &tails) {
    const sb2_host_model &h = b->m->host;
    b->nread = (int)nread;
    b->max_cols = 0;
    b->nsample.resize(nread); b->nblock.resize(nread); b->samp_off.resize(nread); b->col_off.resize(nread + 1);
    tails.resize(nread);
    int64_t so = 0;
    int64_t co = 0;
    for (size_t r = 0; r < nread; r++) {
        if (nsample[r] > (size_t)1 << 30) { sb2_set_error("read %zu too long", r); return -1; }
        if (0 != sb2_conv_plan(nsample[r], h.winlen, h.conv_stride, &tails[r])) return -1;
        b->nsample[r] = (int)nsample[r];
        b->nblock[r] = tails[r].ncol;
        b->samp_off[r] = so;
        b->col_off[r] = (int)co;
        so += (int64_t)((nsample[r] + 3) / 4 * 4);          // keep each read 16-byte aligned
        co += tails[r].ncol;
        b->max_cols = std::max(b->max_cols, (int)tails[r].ncol);
        if (co > (int64_t)1 << 30) { sb2_set_error("batch too large"); return -1; }
    }
    b->col_off[nread] = (int)co;
    b->total_cols = (int)co;
    b->total_samples = so;
    b->pstride = (h.head == 0 && h.nstate == 1025) ? 1056 : (int)h.ostride;
    // read groups of the GRU scan and their rows in the scan-ordered Xin
    b->rpg = scan_reads_per_group((int)nread);
    const size_t ngroup = (nread + b->rpg - 1) / b->rpg;
    b->xgrp.assign(ngroup + 1, 0);
    for (size_t g = 0; g < ngroup; g++) {
        int tmax = 0;
        for (size_t r = g * b->rpg; r < std::min(nread, (g + 1) * b->rpg); r++) tmax = std::max(tmax, b->nblock[r]);
        b->xgrp[g + 1] = b->xgrp[g] + (long long)b->rpg * tmax;
    }
    b->xrows = (size_t)b->xgrp[ngroup];
    // tensor-core scan fed by the tensor-core affine map (the rgrgr / rnnrf topology); row numbers are 32-bit
    b->xil = (h.arch == 0) && b->eng->gemm_impl != 0 && b->eng->scan_impl >= 0 && b->xrows < ((size_t)1 << 31);
    if (!b->xil) b->xrows = (size_t)co;
    return 0;
}

static size_t meta_offsets(size_t cap_reads, size_t off[6]) {
    size_t o = 0;
    off[0] = o; o += align_up(cap_reads * sizeof(int), 16);               // nsample
    off[1] = o; o += align_up(cap_reads * sizeof(int), 16);               // nblock
    off[2] = o; o += align_up((cap_reads + 1) * sizeof(int), 16);         // col_off
    off[3] = o; o += align_up(cap_reads * sizeof(int64_t), 16);           // samp_off
    off[4] = o; o += align_up(cap_reads * sizeof(sb2_conv_tail), 16);     // conv tails
    off[5] = o; o += align_up((cap_reads / 4 + 2) * sizeof(long long), 16);  // first Xin row of every read group
    return o;
}

// Make the device buffers large enough for the current layout.  Fresh batches get exactly what they need; pooled
// workspaces grow with 1/8 head room and keep what they have when the next call is smaller.
static int batch_reserve(sb2_batch *b) {
    const sb2_host_model &h = b->m->host;
    const size_t H = h.H;
    const size_t nread = (size_t)b->nread, ncol_need = (size_t)b->total_cols, nsamp_need = (size_t)b->total_samples;
    if (nread <= b->cap_reads && ncol_need <= b->cap_cols && nsamp_need <= b->cap_samples && b->xrows <= b->cap_xrows) return 0;
    if (b->stream) CUDA_OK(cudaStreamSynchronize(b->stream));
    const bool keep_layers = b->keep_layers;
    auto raise_to = [](std::atomic<size_t> &hw, size_t v) {
        size_t cur = hw.load();
        while (cur < v && !hw.compare_exchange_weak(cur, v)) { }
        return std::max(cur, v);
    };
    size_t cap_reads = std::max(nread, b->cap_reads), ncol = std::max(ncol_need, b->cap_cols), nsamp = std::max(nsamp_need, b->cap_samples);
    const size_t xrows_need = std::max(b->xrows, ncol_need);
    size_t xrows = std::max(xrows_need, b->cap_xrows);
    if (b->pooled) {
        // The 1/16 head room goes on what this call NEEDS, never on the capacity the workspace already has: a re-reserve
        // caused by one dimension (say, more reads) must not push the other dimensions' high-water marks up by another
        // 6 % -- every other workspace of the pool would fall below the marks and be re-made when it is handed back.
        sb2_engine *eng = b->eng;
        const int mt = (int)b->model_type;
        cap_reads = raise_to(eng->hw_reads[mt], cap_reads);
        ncol = raise_to(eng->hw_cols[mt], std::max(ncol_need + ncol_need / 16, b->cap_cols));
        nsamp = raise_to(eng->hw_samples[mt], std::max(nsamp_need + nsamp_need / 16, b->cap_samples));
        xrows = raise_to(eng->hw_xrows[mt], std::max(xrows_need + xrows_need / 16, b->cap_xrows));
    }
    xrows = std::max(xrows, ncol);
    batch_free_device(b);
    b->eng->reallocs += 1;
    // raw_r94 keeps both directions of a bidirectional pair alive and merges them into `ffw` features
    if (h.arch == 1 && (batch_alloc(b, &b->d_Xin2, ncol * 3 * H) || batch_alloc(b, &b->d_FF, ncol * std::max(h.ffw, h.nfilter)))) return -1;
    if (batch_alloc(b, &b->d_raw, nsamp) || batch_alloc(b, &b->d_X[0], ncol * H) || batch_alloc(b, &b->d_X[1], ncol * H) ||
        batch_alloc(b, &b->d_Xin, xrows * 3 * H) || batch_alloc(b, &b->d_post, ncol * std::max((size_t)h.ostride, (size_t)1056)) ||
        batch_alloc(b, &b->d_score, cap_reads) || batch_alloc(b, &b->d_path, ncol + cap_reads))
        return -1;
    if (h.head == 0) {
        if (batch_alloc(b, &b->d_tb, ncol * (h.nstate - 1)) || batch_alloc(b, &b->d_tbE, ncol)) return -1;
    } else {
        if (batch_alloc(b, &b->d_tb, ncol * 8)) return -1;
    }
    size_t off[6];
    b->meta_bytes = meta_offsets(cap_reads, off);
    if (batch_alloc(b, &b->d_meta, b->meta_bytes)) return -1;
    if (pinned_alloc(&b->h_meta, b->meta_bytes)) return -1;
    b->d_nsample = reinterpret_cast<int *>(b->d_meta + off[0]);
    b->d_nblock = reinterpret_cast<int *>(b->d_meta + off[1]);
    b->d_coloff = reinterpret_cast<int *>(b->d_meta + off[2]);
    b->d_sampoff = reinterpret_cast<int64_t *>(b->d_meta + off[3]);
    b->d_tails = reinterpret_cast<sb2_conv_tail *>(b->d_meta + off[4]);
    b->d_xgrp = reinterpret_cast<long long *>(b->d_meta + off[5]);
    if (batch_alloc(b, &b->d_xrow[0], xrows) || batch_alloc(b, &b->d_xrow[1], xrows)) return -1;
    CUDA_OK(cudaMemsetAsync(b->d_raw, 0, nsamp * sizeof(float), b->stream));
    CUDA_OK(cudaMemsetAsync(b->d_Xin, 0, xrows * 3 * H * sizeof(float), b->stream));   // rows of a ragged group's shorter reads are read, never written
    b->cap_reads = cap_reads; b->cap_cols = ncol; b->cap_samples = nsamp; b->cap_xrows = xrows;
    if (keep_layers && batch_alloc(b, &b->d_layers, (size_t)6 * ncol * H)) return -1;
    return 0;
}

// One H2D copy of the batch's dimension tables (asynchronous on the batch's stream: h_meta is pinned and is not
// touched again before the next re-shape, which only happens after the stream has been synchronised).
static int batch_upload_meta(sb2_batch *b, const std::vector<sb2_conv_tail> &tails) {
    size_t off[6];
    meta_offsets(b->cap_reads, off);
    const size_t nread = (size_t)b->nread;
    memcpy(b->h_meta + off[5], b->xgrp.data(), b->xgrp.size() * sizeof(long long));
    memcpy(b->h_meta + off[0], b->nsample.data(), nread * sizeof(int));
    memcpy(b->h_meta + off[1], b->nblock.data(), nread * sizeof(int));
    memcpy(b->h_meta + off[2], b->col_off.data(), (nread + 1) * sizeof(int));
    memcpy(b->h_meta + off[3], b->samp_off.data(), nread * sizeof(int64_t));
    memcpy(b->h_meta + off[4], tails.data(), nread * sizeof(sb2_conv_tail));
    CUDA_OK(cudaMemcpyAsync(b->d_meta, b->h_meta, b->meta_bytes, cudaMemcpyHostToDevice, b->stream));
    b->dims.nread = b->nread;
    b->dims.total_cols = b->total_cols;
    b->dims.max_cols = b->max_cols;
    b->dims.nsample = b->d_nsample;
    b->dims.nblock = b->d_nblock;
    b->dims.samp_off = b->d_sampoff;
    b->dims.col_off = b->d_coloff;
    b->dims.col_read = nullptr;
    return 0;
}

// (Re-)shape a batch for `nread` reads of the given lengths.
static int batch_shape(sb2_batch *b, const size_t *nsample, size_t nread) {
    CUDA_OK(cudaSetDevice(b->eng->device));             // callers are arbitrary host threads: their current device is not ours
    // same lengths as this workspace's last use (a caller streaming equal-sized batches): every table on the device is
    // still valid -- nothing to compute, nothing to upload, and the captured graph stays
    if (nullptr != b->stream && b->cap_reads > 0 && (size_t)b->nread == nread) {
        bool same = true;
        for (size_t r = 0; r < nread && same; r++) same = (size_t)b->nsample[r] == nsample[r];
        if (same) return 0;
    }
    std::vector<sb2_conv_tail> tails;
    const int prev[3] = {b->nread, b->total_cols, b->max_cols};
    if (0 != batch_layout(b, nsample, nread, tails)) return -1;
    CUDA_OK(cudaSetDevice(b->eng->device));
    if (nullptr == b->stream) {
        CUDA_OK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
        for (auto &e : b->ev) CUDA_OK(cudaEventCreate(&e));
    }
    if (0 != batch_reserve(b)) return -1;
    if (0 != batch_upload_meta(b, tails)) return -1;
    if (b->xil) {
        if (0 != launch_scan_rows(b->dims, b->d_xgrp, b->rpg, b->xrows, b->d_xrow[0], b->d_xrow[1], b->stream)) return -1;
        b->eng->launches += 1;
        CUDA_OK(cudaGetLastError());
    }
    // A captured graph bakes in grid sizes and the dimension arguments: it is kept only for an identical shape, and a
    // new shape runs eagerly once before it is captured (a stream of differently shaped batches never pays for
    // capture + instantiation).
    if (prev[0] != b->nread || prev[1] != b->total_cols || prev[2] != b->max_cols) {
        if (b->graph) { cudaGraphExecDestroy(b->graph); b->graph = nullptr; }
        b->eager_runs = 0;
    }
    return 0;
}

static int batch_init(sb2_batch *b, const size_t *nsample, size_t nread) {
    if (0 != batch_shape(b, nsample, nread)) return -1;
    CUDA_OK(cudaStreamSynchronize(b->stream));
    return 0;
}

extern "C" sb2_batch *sb2_batch_create(sb2_engine *eng, enum raw_model_type model, const size_t *nsample, size_t nread) {
    if (nullptr == eng || nullptr == nsample || 0 == nread) { sb2_set_error("batch: bad arguments"); return nullptr; }
    DevModel *dm = get_model(eng, model);
    if (nullptr == dm) return nullptr;
    sb2_batch *b = new sb2_batch();
    b->eng = eng;
    b->model_type = model;
    b->m = dm;
    if (0 != batch_init(b, nsample, nread)) { sb2_batch_destroy(b); return nullptr; }
    return b;
}

extern "C" size_t sb2_batch_nblock(const sb2_batch *b, size_t read) { return (b && read < (size_t)b->nread) ? (size_t)b->nblock[read] : 0; }
extern "C" size_t sb2_batch_total_blocks(const sb2_batch *b) { return b ? (size_t)b->total_cols : 0; }
extern "C" size_t sb2_batch_nstate(const sb2_batch *b) { return b ? b->m->host.nstate : 0; }
extern "C" size_t sb2_batch_total_samples_padded(const sb2_batch *b) { return b ? (size_t)b->total_samples : 0; }
extern "C" size_t sb2_batch_sample_offset(const sb2_batch *b, size_t read) { return (b && read < (size_t)b->nread) ? (size_t)b->samp_off[read] : 0; }

extern "C" int sb2_batch_keep_layers(sb2_batch *b, int keep) {
    if (nullptr == b) return -1;
    if (b->m->host.arch != 0) { sb2_set_error("per-layer dumps are only available for the rgrgr / rnnrf topology"); return -1; }
    CUDA_OK(cudaSetDevice(b->eng->device));
    if (keep && nullptr == b->d_layers && batch_alloc(b, &b->d_layers, (size_t)6 * b->cap_cols * b->m->host.H)) return -1;
    b->keep_layers = keep != 0;
    return 0;
}

extern "C" int sb2_batch_upload(sb2_batch *b, const float *const *signals) {
    if (nullptr == b || nullptr == signals) return -1;
    CUDA_OK(cudaSetDevice(b->eng->device));
    for (int r = 0; r < b->nread; r++)
        CUDA_OK(cudaMemcpyAsync(b->d_raw + b->samp_off[r], signals[r], (size_t)b->nsample[r] * sizeof(float),
                                cudaMemcpyHostToDevice, b->stream));
    CUDA_OK(cudaStreamSynchronize(b->stream));
    return 0;
}

// `concat` uses the batch's padded layout: read r at sb2_batch_sample_offset(b, r).
extern "C" int sb2_batch_upload_concat(sb2_batch *b, const float *concat, int pinned_async) {
    if (nullptr == b || nullptr == concat) return -1;
    CUDA_OK(cudaSetDevice(b->eng->device));
    CUDA_OK(cudaMemcpyAsync(b->d_raw, concat, (size_t)b->total_samples * sizeof(float), cudaMemcpyHostToDevice, b->stream));
    if (!pinned_async) CUDA_OK(cudaStreamSynchronize(b->stream));
    return 0;
}

static inline void stage_mark(sb2_batch *b, int i) { if (b->timing) cudaEventRecord(b->ev[i], b->stream); }

// Launch-configuration errors surface at the launch site (cudaPeekAtLastError costs nothing: no synchronisation), so a
// failure names the stage instead of the end of the pass.
#define LAUNCH_OK(what)                                                                           \
    do {                                                                                          \
        cudaError_t e_ = cudaPeekAtLastError();                                                   \
        if (e_ != cudaSuccess) {                                                                  \
            sb2_set_error("CUDA error %s after launching %s (%s:%d)", cudaGetErrorString(e_), what, __FILE__, __LINE__); \
            cudaGetLastError();                                                                   \
            return -1;                                                                            \
        }                                                                                         \
    } while (0)

// One GRU layer scan with the engine's selected kernel generation.
static int run_scan(sb2_batch *b, const float *Xin, const long long *xgrp, const DevModel &m, int l, const float *resid,
                    float *out, int backward, long long *trace) {
    const int H = (int)b->m->host.H;
    cudaStream_t s = b->stream;
    const int impl = b->eng->scan_impl;
    if (impl < 0) {
        launch_gru_scan_ffma(Xin, m.sW[l], m.sW2[l], resid, out, b->dims, H, backward, s);
        return 0;
    }
    const int rc = launch_gru_scan_tc(Xin, xgrp, m.sW[l], m.sW2[l], resid, out, b->dims, H, backward, impl, trace, s);
    if (0 != rc) sb2_set_error("tensor-core scan: unsupported configuration");
    return rc;
}

// nanonet_raw_posterior (src/networks.c:196-247): conv+tanh, two bidirectional GRU pairs each merged by
// feedforward2_tanh, softmax head.  The GRU scans run on the tensor-core kernel; the affine maps of this
// legacy model (K = 32 / 128, M = 128) use the fp32 kernel.
static int forward_raw_r94(sb2_batch *b, const sb2_params *p, bool return_log) {
    const sb2_host_model &h = b->m->host;
    const DevModel &m = *b->m;
    const int H = (int)h.H, NF = (int)h.nfilter, FW = (int)h.ffw, ncol = b->total_cols;
    cudaStream_t s = b->stream;
    uint64_t nl = 0;
    stage_mark(b, ST_CONV);
    launch_conv_act(b->d_raw, b->dims, b->d_tails, m.conv_taps, m.conv_b, (int)h.winlen, NF, (int)h.conv_stride,
                    (int)h.conv_act, b->d_FF, s);
    nl++;
    const float *in = b->d_FF;
    int K = NF;
    for (int pair = 0; pair < 2; pair++) {
        const int lf = 2 * pair, lb = 2 * pair + 1;
        stage_mark(b, ST_AFFINE(2 * pair));
        launch_affine(in, ncol, K, m.iW[lf], K, m.b[lf], 3 * H, b->d_Xin, 3 * H, 1.0f, 1.0f, 0, 0, s);
        launch_affine(in, ncol, K, m.iW[lb], K, m.b[lb], 3 * H, b->d_Xin2, 3 * H, 1.0f, 1.0f, 0, 0, s);
        stage_mark(b, ST_SCAN(2 * pair));
        if (0 != run_scan(b, b->d_Xin, nullptr, m, lf, nullptr, b->d_X[0], 0, nullptr)) return -1;
        stage_mark(b, ST_AFFINE(2 * pair + 1));
        stage_mark(b, ST_SCAN(2 * pair + 1));
        if (0 != run_scan(b, b->d_Xin2, nullptr, m, lb, nullptr, b->d_X[1], 1, nullptr)) return -1;
        // feedforward2_tanh: tanh(b + Wf gruF + Wb gruB)
        launch_affine(b->d_X[0], ncol, H, m.comb_Wf[pair], H, m.comb_b[pair], FW, b->d_FF, FW, 1.0f, 1.0f, 0, 0, s);
        launch_affine(b->d_X[1], ncol, H, m.comb_Wb[pair], H, nullptr, FW, b->d_FF, FW, 1.0f, 1.0f, 2, 1, s);
        nl += 6;
        in = b->d_FF;
        K = FW;
    }
    stage_mark(b, ST_AFFINE(4));
    stage_mark(b, ST_SCAN(4));
    stage_mark(b, ST_HEAD);
    launch_affine(b->d_FF, ncol, FW, m.FF_W, FW, m.FF_b, (int)h.nstate, b->d_post, b->pstride,
                  p->tempW / p->tempb, p->tempb, 1, 0, s);
    stage_mark(b, ST_FINISH);
    launch_softmax_finish(b->d_post, ncol, (int)h.nstate, b->pstride, p->min_prob, return_log ? 1 : 0, s);
    nl += 2;
    stage_mark(b, ST_DECODE);
    b->eng->launches += nl;
    CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int sb2_batch_forward(sb2_batch *b, const sb2_params *p, bool return_log) {
    if (nullptr == b || nullptr == p) return -1;
    const sb2_host_model &h = b->m->host;
    const DevModel &m = *b->m;
    if (h.head == 1 && !return_log) { sb2_set_error("rnnrf: return_log = false is unsupported (src/networks.c:569)"); return -1; }
    if (!(p->min_prob >= 0.0f && p->min_prob <= 1.0f) || !(p->tempW > 0.0f) || !(p->tempb > 0.0f)) {
        sb2_set_error("invalid min_prob / temperature");
        return -1;
    }
    CUDA_OK(cudaSetDevice(b->eng->device));
    NvtxRange range("sb2_batch_forward");
    cudaGetLastError();     // a stale error of an unrelated earlier runtime call on this thread must not fail this pass
    if (h.arch == 1) return forward_raw_r94(b, p, return_log);
    const int H = (int)h.H;
    cudaStream_t s = b->stream;
    uint64_t nl = 0;

    stage_mark(b, ST_CONV);
    launch_conv_act(b->d_raw, b->dims, b->d_tails, m.conv_taps, m.conv_b, (int)h.winlen, H, (int)h.conv_stride,
                    (int)h.conv_act, b->d_X[0], s);
    LAUNCH_OK("conv_act");
    nl++;
    int cur = 0;
    const size_t layer_bytes = (size_t)b->total_cols * H * sizeof(float);
    if (b->keep_layers) CUDA_OK(cudaMemcpyAsync(b->d_layers, b->d_X[0], layer_bytes, cudaMemcpyDeviceToDevice, s));
    for (int l = 0; l < SB2_NLAYER; l++) {
        NvtxRange layer_range("gru layer: affine + scan");
        stage_mark(b, ST_AFFINE(l));
        if (b->eng->gemm_impl == 0) {
            launch_affine(b->d_X[cur], b->total_cols, H, m.iW[l], H, m.b[l], 3 * H, b->d_Xin, 3 * H, 1.0f, 1.0f, 0, 0, s);
        } else if (0 != launch_affine_tc(b->d_X[cur], b->xil ? (int)b->xrows : b->total_cols, H, m.iw_img[l], m.b[l], b->d_Xin,
                                         b->xil ? b->d_xrow[(l % 2) == 0 ? 1 : 0] : nullptr, s)) {
            sb2_set_error("tensor-core affine kernel could not be configured");
            return -1;
        }
        LAUNCH_OK("affine (GRU input transform)");
        stage_mark(b, ST_SCAN(l));
        if (0 != run_scan(b, b->d_Xin, b->xil ? b->d_xgrp : nullptr, m, l, h.residual ? b->d_X[cur] : nullptr, b->d_X[cur ^ 1], (l % 2) == 0,
                          (l == 1) ? b->eng->d_trace : nullptr))
            return -1;
        LAUNCH_OK("gru_scan");
        nl += 2;
        cur ^= 1;
        if (b->keep_layers)
            CUDA_OK(cudaMemcpyAsync(b->d_layers + (size_t)(l + 1) * b->total_cols * H, b->d_X[cur], layer_bytes,
                                    cudaMemcpyDeviceToDevice, s));
    }
    b->final_x = cur;
    NvtxRange head_range("head: FF + softmax / globalnorm");
    stage_mark(b, ST_HEAD);
    if (h.head == 0 && b->eng->gemm_impl != 0 && nullptr != m.head_img) {
        if (0 != launch_head_softmax_tc(b->d_X[cur], b->total_cols, H, m.head_img, m.FF_W + (size_t)1024 * H, m.FF_b,
                                        b->d_post, b->pstride, p->tempW / p->tempb, p->tempb, p->min_prob,
                                        return_log ? 1 : 0, b->eng->head_exact, s)) {
            sb2_set_error("tensor-core head kernel could not be configured");
            return -1;
        }
        stage_mark(b, ST_FINISH);
        nl += 1;
    } else if (h.head == 0) {
        launch_affine(b->d_X[cur], b->total_cols, H, m.FF_W, H, m.FF_b, (int)h.nstate, b->d_post, b->pstride,
                      p->tempW / p->tempb, p->tempb, 1, 0, s);
        stage_mark(b, ST_FINISH);
        launch_softmax_finish(b->d_post, b->total_cols, (int)h.nstate, b->pstride, p->min_prob, return_log ? 1 : 0, s);
        nl += 2;
    } else {
        if (h.nstate <= 32 && h.ostride <= 32 && b->eng->gemm_impl != 0) {
            launch_small_head(b->d_X[cur], b->total_cols, H, m.FF_W, m.FF_b, (int)h.nstate, b->d_post, b->pstride, s);
        } else {
            CUDA_OK(cudaMemsetAsync(b->d_post, 0, (size_t)b->total_cols * b->pstride * sizeof(float), s));
            launch_affine(b->d_X[cur], b->total_cols, H, m.FF_W, H, m.FF_b, (int)h.nstate, b->d_post, b->pstride,
                          1.0f, 1.0f, 0, 0, s);
        }
        stage_mark(b, ST_FINISH);
        launch_globalnorm(b->d_post, b->dims, b->pstride, s);
        nl += 2;
    }
    stage_mark(b, ST_DECODE);
    b->eng->launches += nl;
    CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int sb2_batch_decode(sb2_batch *b, const sb2_params *p) {
    if (nullptr == b || nullptr == p) return -1;
    const sb2_host_model &h = b->m->host;
    CUDA_OK(cudaSetDevice(b->eng->device));
    NvtxRange range("sb2_batch_decode");
    if (h.head == 0)
        launch_decode_transducer(b->d_post, b->dims, (int)h.nstate, b->pstride, p->stay_pen, p->skip_pen,
                                 p->local_pen, p->allow_slip, b->d_tb, b->d_tbE, b->d_path, b->d_score, b->stream);
    else
        launch_decode_crf(b->d_post, b->dims, b->pstride, b->d_tb, b->d_path, b->d_score, b->stream);
    b->eng->launches += 1;
    stage_mark(b, ST_COUNT);
    CUDA_OK(cudaGetLastError());
    return 0;
}

// forward (log posterior) + decode.  After one eager run the launch sequence of a batch is captured into a CUDA
// graph and replayed while the parameters stay the same; SCRAPPIE_B200_GRAPH=0 keeps every launch eager.
extern "C" int sb2_batch_run(sb2_batch *b, const sb2_params *p) {
    if (nullptr == b || nullptr == p) return -1;
    static const bool use_graph = !(getenv("SCRAPPIE_B200_GRAPH") && 0 == strcmp(getenv("SCRAPPIE_B200_GRAPH"), "0"));
    if (!use_graph || b->timing || b->keep_layers)
        return (0 == sb2_batch_forward(b, p, true) && 0 == sb2_batch_decode(b, p)) ? 0 : -1;
    CUDA_OK(cudaSetDevice(b->eng->device));
    if (nullptr != b->graph && 0 == memcmp(&b->graph_params, p, sizeof(*p))) {
        CUDA_OK(cudaGraphLaunch(b->graph, b->stream));
        b->eng->launches += b->graph_launches;
        return 0;
    }
    if (b->eager_runs++ == 0)       // first run eager: one-time kernel attribute set-up must not happen under capture
        return (0 == sb2_batch_forward(b, p, true) && 0 == sb2_batch_decode(b, p)) ? 0 : -1;
    if (b->graph) { cudaGraphExecDestroy(b->graph); b->graph = nullptr; }
    const uint64_t l0 = b->eng->launches.load();
    CUDA_OK(cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = (0 == sb2_batch_forward(b, p, true) && 0 == sb2_batch_decode(b, p)) ? 0 : -1;
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(b->stream, &g);
    if (0 != rc || e != cudaSuccess || nullptr == g) {
        if (g) cudaGraphDestroy(g);
        if (0 == rc) sb2_set_error("CUDA graph capture failed: %s", cudaGetErrorString(e));
        return -1;
    }
    // per-node launch priorities (the scan's, when enabled) are honoured only with this flag
    static const unsigned long long gflags = (scan_launch_priority() != 0) ? cudaGraphInstantiateFlagUseNodePriority : 0;
    const cudaError_t ei = cudaGraphInstantiate(&b->graph, g, gflags);
    cudaGraphDestroy(g);
    if (ei != cudaSuccess) { b->graph = nullptr; sb2_set_error("CUDA graph instantiation failed: %s", cudaGetErrorString(ei)); return -1; }
    b->graph_params = *p;
    b->graph_dims[0] = b->nread; b->graph_dims[1] = b->total_cols; b->graph_dims[2] = b->max_cols;
    // the launches counted while capturing were recorded, not executed: they run now, with the graph
    b->graph_launches = b->eng->launches.load() - l0;
    CUDA_OK(cudaGraphLaunch(b->graph, b->stream));
    return 0;
}

extern "C" int sb2_batch_sync(sb2_batch *b) {
    if (nullptr == b) return -1;
    CUDA_OK(cudaSetDevice(b->eng->device));
    CUDA_OK(cudaStreamSynchronize(b->stream));
    return 0;
}

extern "C" int sb2_batch_download_posterior(sb2_batch *b, size_t read, float *dst, size_t dst_stride) {
    if (nullptr == b || nullptr == dst || read >= (size_t)b->nread) return -1;
    const sb2_host_model &h = b->m->host;
    CUDA_OK(cudaSetDevice(b->eng->device));
    CUDA_OK(cudaStreamSynchronize(b->stream));
    const float *src = b->d_post + (size_t)b->col_off[read] * b->pstride;
    CUDA_OK(cudaMemcpy2D(dst, dst_stride * sizeof(float), src, b->pstride * sizeof(float),
                         std::min((size_t)h.ostride, dst_stride) * sizeof(float), (size_t)b->nblock[read],
                         cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int sb2_batch_download_paths(sb2_batch *b, int *paths_concat, float *scores) {
    if (nullptr == b) return -1;
    CUDA_OK(cudaSetDevice(b->eng->device));
    if (paths_concat)
        CUDA_OK(cudaMemcpyAsync(paths_concat, b->d_path, ((size_t)b->total_cols + b->nread) * sizeof(int),
                                cudaMemcpyDeviceToHost, b->stream));
    if (scores)
        CUDA_OK(cudaMemcpyAsync(scores, b->d_score, (size_t)b->nread * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    CUDA_OK(cudaStreamSynchronize(b->stream));
    return 0;
}

extern "C" int sb2_batch_download_layer(sb2_batch *b, int layer, size_t read, float *dst) {
    if (nullptr == b || nullptr == dst || layer < 0 || layer > 5 || read >= (size_t)b->nread || nullptr == b->d_layers) {
        sb2_set_error("layer download: enable sb2_batch_keep_layers before forward");
        return -1;
    }
    const size_t H = b->m->host.H;
    CUDA_OK(cudaSetDevice(b->eng->device));
    CUDA_OK(cudaStreamSynchronize(b->stream));
    const float *src = b->d_layers + ((size_t)layer * b->total_cols + b->col_off[read]) * H;
    CUDA_OK(cudaMemcpy(dst, src, (size_t)b->nblock[read] * H * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int sb2_batch_time(sb2_batch *b, const sb2_params *p, int nrep, int flush_l2, float *ms_out,
                              float *ms_forward_out, float *ms_decode_out) {
    if (nullptr == b || nullptr == p || nrep <= 0) return -1;
    sb2_engine *eng = b->eng;
    CUDA_OK(cudaSetDevice(eng->device));
    if (flush_l2 && nullptr == eng->flush_buf) {
        eng->flush_n = (size_t)96 << 20;                 // 384 MB of floats > 126 MB L2
        if (dev_alloc(&eng->flush_buf, eng->flush_n)) return -1;
    }
    cudaEvent_t e0, e1, e2;
    CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1)); CUDA_OK(cudaEventCreate(&e2));
    int rc = 0;
    for (int i = 0; i < nrep && 0 == rc; i++) {
        if (flush_l2) launch_flush(eng->flush_buf, eng->flush_n, b->stream);
        b->timing = (i == nrep - 1);
        cudaEventRecord(e0, b->stream);
        rc |= sb2_batch_forward(b, p, true);
        cudaEventRecord(e1, b->stream);
        rc |= sb2_batch_decode(b, p);
        cudaEventRecord(e2, b->stream);
        if (cudaStreamSynchronize(b->stream) != cudaSuccess) rc = -1;
        if (rc) break;
        float f = 0, d = 0;
        cudaEventElapsedTime(&f, e0, e1);
        cudaEventElapsedTime(&d, e1, e2);
        if (ms_out) ms_out[i] = f + d;
        if (ms_forward_out) ms_forward_out[i] = f;
        if (ms_decode_out) ms_decode_out[i] = d;
    }
    if (0 == rc) {
        for (int st = 0; st < ST_COUNT; st++) cudaEventElapsedTime(&b->stage_ms[st], b->ev[st], b->ev[st + 1]);
    }
    b->timing = false;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    if (rc) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) sb2_set_error("CUDA error %s in timed run", cudaGetErrorString(e)); }
    return rc;
}

extern "C" int sb2_batch_stage_ms(const sb2_batch *b, float *stage_ms, int nstage_max) {
    if (nullptr == b || nullptr == stage_ms) return -1;
    const int n = std::min(nstage_max, (int)ST_COUNT);
    for (int i = 0; i < n; i++) stage_ms[i] = b->stage_ms[i];
    return n;
}

// stage boundaries of the last timed multi-batch run, in ms after the first batch's first launch
extern "C" int sb2_batch_stage_offsets(const sb2_batch *b, float *at, int n_max) {
    if (nullptr == b || nullptr == at) return -1;
    const int n = std::min(n_max, (int)ST_COUNT + 1);
    for (int i = 0; i < n; i++) at[i] = b->stage_at[i];
    return n;
}

// ------------------------------------------------------------------------------------
// whole-read basecalling for a batch (calculate_post, src/scrappie_raw.c:265-315)
// ------------------------------------------------------------------------------------

static double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// Number of host threads for the per-read post-processing of one batch (SCRAPPIE_B200_HOST_THREADS, default 4).
static int host_threads() {
    static int n = 0;
    if (n == 0) {
        const char *e = getenv("SCRAPPIE_B200_HOST_THREADS");
        n = e ? atoi(e) : 4;
        if (n < 1) n = 1;
    }
    return n;
}

static int basecall_buffers(sb2_batch *b) {
    // (all of these are freed together with the device buffers when a workspace grows; each is checked on its own so
    // that a call that failed half-way through can simply be repeated)
    const size_t np = b->cap_cols + b->cap_reads;
    // every run position needs (stay, repeat k-mer) of one column; runs can overlap, so leave head room
    const size_t gcap = b->gcap ? b->gcap : 4 * b->cap_cols + 64;
    if (nullptr == b->h_paths && pinned_alloc(&b->h_paths, np)) return -1;
    if (nullptr == b->h_scores && pinned_alloc(&b->h_scores, b->cap_reads)) return -1;
    if (nullptr == b->h_gidx && pinned_alloc(&b->h_gidx, gcap * 2)) return -1;
    if (nullptr == b->h_gval && pinned_alloc(&b->h_gval, gcap)) return -1;
    if (nullptr == b->d_gidx && batch_alloc(b, &b->d_gidx, gcap * 2)) return -1;
    if (nullptr == b->d_gval && batch_alloc(b, &b->d_gval, gcap)) return -1;
    b->gcap = gcap;
    return 0;
}

static int finish_buffers(sb2_batch *b) {
    const sb2_host_model &h = b->m->host;
    const int klen = (h.head == 0) ? (int)(logf((float)h.nstate) / logf(4.0f)) : 1;
    // worst case: every block moves by a full k-mer
    b->bases_stride = (int)align_up((size_t)klen * ((size_t)b->max_cols + 1) + 1, 16);
    const size_t nbytes = (size_t)b->nread * b->bases_stride;
    if (nullptr == b->d_path2 && batch_alloc(b, &b->d_path2, b->cap_cols + b->cap_reads)) return -1;
    if (nullptr == b->d_nbase && batch_alloc(b, &b->d_nbase, b->cap_reads)) return -1;
    if (nullptr == b->h_nbase && pinned_alloc(&b->h_nbase, b->cap_reads)) return -1;
    if (nullptr == b->h_scores && pinned_alloc(&b->h_scores, b->cap_reads)) return -1;
    if (nbytes > b->cap_bases) {
        b->eng->reallocs += 1;
        if (b->d_bases) { CUDA_OK(cudaStreamSynchronize(b->stream)); cudaFreeAsync(b->d_bases, b->stream); cudaFreeHost(b->h_bases); b->d_bases = nullptr; b->h_bases = nullptr; }
        size_t cap = nbytes;
        if (b->pooled) {                                // engine-wide high-water mark, as for the activations
            std::atomic<size_t> &hw = b->eng->hw_bases[(int)b->model_type];
            size_t cur = hw.load();
            cap = nbytes + nbytes / 4;
            while (cur < cap && !hw.compare_exchange_weak(cur, cap)) { }
            cap = std::max(cur, cap);
        }
        if (batch_alloc(b, &b->d_bases, cap)) return -1;
        if (pinned_alloc(&b->h_bases, cap)) { cudaFreeAsync(b->d_bases, b->stream); b->d_bases = nullptr; return -1; }
        b->cap_bases = cap;
    }
    return 0;
}

// The tail of calculate_post (src/scrappie_raw.c:290-312) for a whole batch on the device.
static int finish_on_device(sb2_batch *b, const sb2_params *p, sb2_call *out) {
    NvtxRange range("finish reads: homopolymer + bases, D2H");
    const sb2_host_model &h = b->m->host;
    if (0 != finish_buffers(b)) return -1;
    const int klen = (h.head == 0) ? (int)(logf((float)h.nstate) / logf(4.0f)) : 1;
    // small batches fetch the whole base-string area in one copy (below): its unused tails are defined bytes
    if ((size_t)b->nread * b->bases_stride <= ((size_t)2 << 20))
        CUDA_OK(cudaMemsetAsync(b->d_bases, 0, (size_t)b->nread * b->bases_stride, b->stream));
    launch_finish_reads(b->d_post, b->dims, (int)h.nstate, b->pstride, (int)h.head,
                        (h.head == 0 && p->homopolymer == HOMOPOLYMER_MEAN) ? 1 : 0, klen, b->d_path, b->d_path2,
                        b->d_bases, b->bases_stride, b->d_nbase, b->stream);
    b->eng->launches += 1;
    CUDA_OK(cudaMemcpyAsync(b->h_nbase, b->d_nbase, (size_t)b->nread * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CUDA_OK(cudaMemcpyAsync(b->h_scores, b->d_score, (size_t)b->nread * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    // Host threads wait on a blocking-sync event (they sleep instead of spinning: with one thread per batch and
    // several ranks per host the spinning waits starved each other).  Small batches fetch the whole base-string
    // area in the same pass; large ones first learn the longest call, then copy only that width.
    if (nullptr == b->ev_done) CUDA_OK(cudaEventCreateWithFlags(&b->ev_done, cudaEventBlockingSync | cudaEventDisableTiming));
    const size_t all_bytes = (size_t)b->nread * b->bases_stride;
    if (all_bytes <= ((size_t)2 << 20)) {
        CUDA_OK(cudaMemcpyAsync(b->h_bases, b->d_bases, all_bytes, cudaMemcpyDeviceToHost, b->stream));
        CUDA_OK(cudaEventRecord(b->ev_done, b->stream));
        CUDA_OK(cudaEventSynchronize(b->ev_done));
    } else {
        CUDA_OK(cudaEventRecord(b->ev_done, b->stream));
        CUDA_OK(cudaEventSynchronize(b->ev_done));
        int maxlen = 0;
        for (int r = 0; r < b->nread; r++) maxlen = std::max(maxlen, b->h_nbase[r]);
        const size_t width = std::min((size_t)b->bases_stride, align_up((size_t)maxlen + 1, 16));
        CUDA_OK(cudaMemcpy2DAsync(b->h_bases, b->bases_stride, b->d_bases, b->bases_stride, width, (size_t)b->nread,
                                  cudaMemcpyDeviceToHost, b->stream));
        CUDA_OK(cudaEventRecord(b->ev_done, b->stream));
        CUDA_OK(cudaEventSynchronize(b->ev_done));
    }
    int ncalled = 0;
    for (int r = 0; r < b->nread; r++) {
        const int nbase = b->h_nbase[r];
        out[r].score = b->h_scores[r];
        out[r].nblock = (size_t)b->nblock[r];
        if (nbase < 0) continue;                         // no k-mer in the path: NULL like the host overlapper
        char *bases = static_cast<char *>(calloc((size_t)nbase + 1, 1));
        if (nullptr == bases) continue;
        memcpy(bases, b->h_bases + (size_t)r * b->bases_stride, (size_t)nbase);
        out[r].bases = bases;
        out[r].nbase = (size_t)nbase;
        ncalled++;
    }
    return ncalled;
}

struct HpJob { int read; sb2_hp_run run; size_t off; };

// Homopolymer fix-up: runs are found on the host from the Viterbi path; only the two posterior
// entries per run position (stay, repeat k-mer) are fetched from HBM (src/homopolymer.c:205-217).
static int homopolymer_fixup(sb2_batch *b, int *paths) {
    const sb2_host_model &h = b->m->host;
    const int klen = (int)(logf((float)h.nstate) / logf(4.0f));
    const int nread = b->nread;
    std::vector<sb2_hp_run *> runs(nread, nullptr);
    std::vector<int> nrun(nread, 0);
    std::vector<size_t> need(nread + 1, 0);
    int bad = 0;
#pragma omp parallel for schedule(static) num_threads(host_threads()) reduction(| : bad)
    for (int r = 0; r < nread; r++) {
        nrun[r] = sb2_find_homopolymer_runs(paths + b->col_off[r] + r, b->nblock[r], klen, &runs[r]);
        if (nrun[r] < 0) { bad |= 1; nrun[r] = 0; }
        size_t n = 0;
        for (int i = 0; i < nrun[r]; i++) n += (size_t)runs[r][i].length;
        need[r + 1] = n;
    }
    for (int r = 0; r < nread; r++) need[r + 1] += need[r];
    const size_t nent = 2 * need[nread];
    int rc = bad ? -1 : 0;
    if (0 == rc && nent > b->gcap) {                    // overlapping runs can exceed the estimate: grow, do not fail
        cudaStreamSynchronize(b->stream);
        cudaFreeHost(b->h_gidx); cudaFreeHost(b->h_gval); cudaFreeAsync(b->d_gidx, b->stream); cudaFreeAsync(b->d_gval, b->stream);
        b->h_gidx = nullptr; b->h_gval = nullptr; b->d_gidx = nullptr; b->d_gval = nullptr;
        b->gcap = nent + nent / 2;
        if (cudaMallocHost(reinterpret_cast<void **>(&b->h_gidx), b->gcap * 2 * sizeof(int)) != cudaSuccess ||
            cudaMallocHost(reinterpret_cast<void **>(&b->h_gval), b->gcap * sizeof(float)) != cudaSuccess ||
            batch_alloc(b, &b->d_gidx, b->gcap * 2) || batch_alloc(b, &b->d_gval, b->gcap)) {
            sb2_set_error("homopolymer gather buffers: out of memory");
            rc = -1;
        }
    }
    if (0 == rc && nent > 0) {
#pragma omp parallel for schedule(static) num_threads(host_threads())
        for (int r = 0; r < nread; r++) {
            size_t k = 2 * need[r];
            for (int i = 0; i < nrun[r]; i++)
                for (int j = 0; j < runs[r][i].length; j++) {
                    const int col = b->col_off[r] + runs[r][i].start + j - 1;
                    b->h_gidx[2 * k] = col; b->h_gidx[2 * k + 1] = (int)h.nstate - 1; k++;
                    b->h_gidx[2 * k] = col; b->h_gidx[2 * k + 1] = runs[r][i].state; k++;
                }
        }
        if (cudaMemcpyAsync(b->d_gidx, b->h_gidx, nent * 2 * sizeof(int), cudaMemcpyHostToDevice, b->stream) != cudaSuccess) rc = -1;
        launch_gather(b->d_post, b->pstride, b->d_gidx, (int)nent, b->d_gval, b->stream);
        b->eng->launches += 1;
        if (cudaMemcpyAsync(b->h_gval, b->d_gval, nent * sizeof(float), cudaMemcpyDeviceToHost, b->stream) != cudaSuccess) rc = -1;
        if (cudaStreamSynchronize(b->stream) != cudaSuccess) rc = -1;
        if (rc) sb2_set_error("homopolymer gather failed");
    }
    if (0 == rc && nent > 0) {
#pragma omp parallel for schedule(static) num_threads(host_threads())
        for (int r = 0; r < nread; r++) {
            std::vector<float> ps, pr;
            size_t k = 2 * need[r];
            for (int i = 0; i < nrun[r]; i++) {
                const int len = runs[r][i].length;
                ps.resize(len); pr.resize(len);
                for (int j = 0; j < len; j++) { ps[j] = b->h_gval[k++]; pr[j] = b->h_gval[k++]; }
                sb2_apply_homopolymer_run(paths + b->col_off[r] + r, &runs[r][i], ps.data(), pr.data());
            }
        }
    }
    for (int r = 0; r < nread; r++) free(runs[r]);
    return rc;
}

// Basecall on an existing batch workspace: upload (from `concat` in the padded layout if
// given, else the signals must already be resident), forward, decode, download paths,
// homopolymer fix-up, overlapper.
extern "C" int sb2_batch_basecall(sb2_batch *b, const float *concat, int pinned, const sb2_params *p, sb2_call *out) {
    if (nullptr == b || nullptr == p || nullptr == out) return -1;
    const int nread = b->nread;
    for (int r = 0; r < nread; r++) out[r] = sb2_call{nullptr, NAN, 0, 0};
    static const bool timing = getenv("SCRAPPIE_B200_TIMING") != nullptr;
    const double t0 = now_ms();
    NvtxRange range("sb2_batch_basecall");
    CUDA_OK(cudaSetDevice(b->eng->device));
    if (nullptr != concat && 0 != sb2_batch_upload_concat(b, concat, pinned)) return -1;
    // SCRAPPIE_B200_FINISH=host keeps the homopolymer fix-up and the overlapper on the CPU (cross-check path)
    static const bool finish_host = getenv("SCRAPPIE_B200_FINISH") && 0 == strcmp(getenv("SCRAPPIE_B200_FINISH"), "host");
    if (!finish_host) {
        if (0 != sb2_batch_run(b, p)) return -1;
        const double tg = now_ms();
        const int n = finish_on_device(b, p, out);
        if (timing) fprintf(stderr, "scrappie_b200: basecall %d reads: launch %.2f ms, gpu + finishing + copies %.2f ms\n", nread, tg - t0, now_ms() - tg);
        return n;
    }
    if (0 != basecall_buffers(b)) return -1;
    if (0 != sb2_batch_run(b, p) || 0 != sb2_batch_download_paths(b, b->h_paths, b->h_scores)) return -1;
    const double t1 = now_ms();
    const sb2_host_model &h = b->m->host;
    int *paths = b->h_paths;
    if (h.head == 0 && p->homopolymer == HOMOPOLYMER_MEAN && 0 != homopolymer_fixup(b, paths)) return -1;
    const double t2 = now_ms();
    int ncalled = 0;
#pragma omp parallel for schedule(static) num_threads(host_threads()) reduction(+ : ncalled)
    for (int r = 0; r < nread; r++) {
        const int *path = paths + b->col_off[r] + r;
        const size_t nb = (size_t)b->nblock[r];
        std::vector<int> pos(nb + 1, 0);
        char *bases = (h.head == 0) ? overlapper(path, nb + 1, (int)h.nstate - 1, pos.data())
                                    : crfpath_to_basecall(path, nb, pos.data());
        out[r].bases = bases;
        out[r].score = b->h_scores[r];
        out[r].nblock = nb;
        out[r].nbase = bases ? strlen(bases) : 0;
        if (bases) ncalled++;
    }
    if (timing)
        fprintf(stderr, "scrappie_b200: basecall %d reads: gpu+copies %.2f ms, homopolymer %.2f ms, bases %.2f ms\n", nread,
                t1 - t0, t2 - t1, now_ms() - t2);
    return ncalled;
}

// Release the base strings of an array of calls (bases are calloc'd, as in the reference).
extern "C" void sb2_calls_free(sb2_call *calls, size_t n) {
    if (nullptr == calls) return;
    for (size_t i = 0; i < n; i++) { free(calls[i].bases); calls[i].bases = nullptr; }
}

// ---- workspace pool -------------------------------------------------------------------------------------------
// sb2_basecall_batch / sb2_basecall_raw_batch take an idle workspace of the model (or make one), re-shape it for the
// call and hand it back afterwards: device buffers, pinned staging areas, the stream and -- while consecutive calls
// have the same shape -- the captured CUDA graph are reused, so a steady stream of calls allocates nothing.
// Concurrent callers each get their own workspace.
static sb2_batch *pool_acquire(sb2_engine *eng, enum raw_model_type model, const size_t *nsample, size_t nread) {
    DevModel *dm = get_model(eng, model);
    if (nullptr == dm) return nullptr;
    // what the call will need, to pick the idle workspace that fits best (a stream of differently sized batches must
    // not make every workspace grow to the largest, nor re-allocate on every call)
    size_t need_cols = 0, need_samples = 0;
    for (size_t r = 0; r < nread; r++) {
        need_cols += (nsample[r] + dm->host.conv_stride - 1) / dm->host.conv_stride;
        need_samples += (nsample[r] + 3) / 4 * 4;
    }
    {
        std::lock_guard<std::mutex> lock(eng->mu);
        auto &free_list = eng->pool[model];
        int best = -1, largest = -1;
        for (int i = 0; i < (int)free_list.size(); i++) {
            const sb2_batch *w = free_list[i];
            const bool fits = w->cap_cols >= need_cols && w->cap_samples >= need_samples && w->cap_reads >= nread;
            if (fits && (best < 0 || w->cap_cols < free_list[best]->cap_cols)) best = i;
            if (largest < 0 || w->cap_cols > free_list[largest]->cap_cols) largest = i;
        }
        const int pick = best >= 0 ? best : largest;
        if (pick >= 0) {
            sb2_batch *b = free_list[pick];
            free_list.erase(free_list.begin() + pick);
            return b;
        }
    }
    sb2_batch *b = new sb2_batch();
    b->eng = eng;
    b->model_type = model;
    b->m = dm;
    b->pooled = true;
    return b;
}

static void pool_release(sb2_engine *eng, sb2_batch *b, bool healthy) {
    if (nullptr == b) return;
    if (!healthy) { sb2_batch_destroy(b); return; }    // never recycle a workspace a CUDA error went through
    // A workspace made before the high-water marks reached their present values would be passed over by the best-fit
    // search until, much later, nothing else is idle and it has to grow in the middle of steady state: let it go now
    // (stream-ordered frees: cheap), the next caller that finds the pool short makes a full-sized one.
    const int mt = (int)b->model_type;
    if (b->cap_cols < eng->hw_cols[mt] || b->cap_samples < eng->hw_samples[mt] || b->cap_xrows < eng->hw_xrows[mt] ||
        b->cap_reads < eng->hw_reads[mt]) {
        sb2_batch_destroy(b);
        return;
    }
    std::lock_guard<std::mutex> lock(eng->mu);
    eng->pool[b->model_type].push_back(b);
}

// Release every idle pooled workspace (device memory, pinned staging, graphs).  Workspaces in use by a concurrent call are
// unaffected.  Returns the number released.
extern "C" int sb2_engine_trim_pool(sb2_engine *eng) {
    if (nullptr == eng) return -1;
    std::vector<sb2_batch *> idle;
    {
        std::lock_guard<std::mutex> lock(eng->mu);
        for (auto &free_list : eng->pool) {
            idle.insert(idle.end(), free_list.begin(), free_list.end());
            free_list.clear();
        }
    }
    for (sb2_batch *b : idle) sb2_batch_destroy(b);
    {   // hand the pooled device memory back to the driver (other allocators of the process may want it)
        cudaMemPool_t pool = nullptr;
        if (cudaSetDevice(eng->device) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, eng->device) == cudaSuccess) {
            cudaDeviceSynchronize();
            cudaMemPoolTrimTo(pool, 0);
        }
        cudaGetLastError();
    }
    for (int m = 0; m < SB2_NMODEL; m++) { eng->hw_reads[m] = 0; eng->hw_cols[m] = 0; eng->hw_samples[m] = 0; eng->hw_bases[m] = 0; eng->hw_xrows[m] = 0; eng->hw_stage[m] = 0; }
    return (int)idle.size();
}

// Shortest signal the model's convolution accepts; shorter reads are dropped from a batch (bases NULL, score NAN)
// instead of failing it -- the reference handles reads one at a time, so a bad read never loses the others.
static size_t min_read_samples(sb2_engine *eng, enum raw_model_type model) {
    DevModel *dm = get_model(eng, model);
    return dm ? (size_t)dm->host.winlen : 0;
}

// Signals from ordinary (pageable) host memory: gathered into the workspace's pinned staging area in the batch's
// padded layout (pads zeroed), then one asynchronous H2D copy.
static int stage_signals(sb2_batch *b, const float *const *signals, const std::vector<size_t> &keep) {
    NvtxRange range("stage signals: pageable -> pinned -> device");
    CUDA_OK(cudaSetDevice(b->eng->device));
    const size_t total = (size_t)b->total_samples;
    if (total > b->stage_cap) {
        b->eng->reallocs += 1;
        if (b->h_stage) cudaFreeHost(b->h_stage);
        b->h_stage = nullptr;
        size_t cap = total + total / 8;
        {   // engine-wide high-water mark: a workspace that starts with a small batch is sized for the largest seen
            std::atomic<size_t> &hw = b->eng->hw_stage[(int)b->model_type];
            size_t cur = hw.load();
            while (cur < cap && !hw.compare_exchange_weak(cur, cap)) { }
            cap = std::max(cur, cap);
        }
        b->stage_cap = 0;
        if (pinned_alloc(&b->h_stage, cap)) return -1;
        b->stage_cap = cap;
    }
    // plain loop: the callers are already one host thread per batch in flight; an OpenMP team per caller would leave
    // dozens of spinning workers behind every call
    const int n = b->nread;
    for (int r = 0; r < n; r++) {
        float *dst = b->h_stage + b->samp_off[r];
        const size_t len = (size_t)b->nsample[r], padded = (len + 3) / 4 * 4;
        memcpy(dst, signals[keep[r]], len * sizeof(float));
        for (size_t i = len; i < padded; i++) dst[i] = 0.0f;
    }
    CUDA_OK(cudaMemcpyAsync(b->d_raw, b->h_stage, total * sizeof(float), cudaMemcpyHostToDevice, b->stream));
    return 0;
}

extern "C" int sb2_basecall_batch(sb2_engine *eng, enum raw_model_type model, const float *const *signals,
                                  const size_t *nsample, size_t nread, const sb2_params *p, sb2_call *out) {
    if (nullptr == eng || nullptr == signals || nullptr == nsample || nullptr == p || nullptr == out) return -1;
    for (size_t r = 0; r < nread; r++) out[r] = sb2_call{nullptr, NAN, 0, 0};
    const size_t min_len = min_read_samples(eng, model);
    if (0 == min_len) return -1;
    std::vector<size_t> keep, len;
    for (size_t r = 0; r < nread; r++)
        if (nullptr != signals[r] && nsample[r] >= min_len) { keep.push_back(r); len.push_back(nsample[r]); }
    if (keep.empty()) return 0;
    sb2_batch *b = pool_acquire(eng, model, len.data(), len.size());
    if (nullptr == b) return -1;
    int ncalled = -1;
    std::vector<sb2_call> calls(keep.size(), sb2_call{nullptr, NAN, 0, 0});
    if (0 == batch_shape(b, len.data(), len.size()) && 0 == stage_signals(b, signals, keep)) {
        ncalled = sb2_batch_basecall(b, nullptr, 0, p, calls.data());
        if (ncalled >= 0) for (size_t i = 0; i < keep.size(); i++) out[keep[i]] = calls[i];
    }
    pool_release(eng, b, ncalled >= 0);
    return ncalled;
}

// ------------------------------------------------------------------------------------
// signal preparation on the device (trim_and_segment_raw + medmad_normalise_array) and the raw-signal basecall
// ------------------------------------------------------------------------------------
extern "C" sb2_trim sb2_default_trim(void) {
    sb2_trim t;
    t.trim_start = 200; t.trim_end = 10; t.varseg_chunk = 100; t.varseg_thresh = 0.0f;   // src/scrappie_raw.c:98-121
    return t;
}

namespace {
struct RawStage {
    float *d_raw = nullptr, *d_mads = nullptr;
    int64_t *d_off = nullptr, *d_moff = nullptr;
    int *d_n = nullptr, *d_se = nullptr;
    std::vector<int64_t> off;
    ~RawStage() {
        void *ptrs[] = {d_raw, d_mads, d_off, d_moff, d_n, d_se};
        for (void *p : ptrs) if (p) cudaFree(p);
    }
};
}  // namespace

// Upload the untrimmed signals and run the trimmer; start/end per read (end = 0: nothing left).
static int stage_and_trim(sb2_engine *eng, const float *const *raws, const size_t *nsample, size_t nread,
                          const sb2_trim *t, RawStage &st, std::vector<int> &start, std::vector<int> &end) {
    NvtxRange range("upload raw signals + trim");
    CUDA_OK(cudaSetDevice(eng->device));
    st.off.resize(nread);
    std::vector<int64_t> moff(nread);
    std::vector<int> n(nread);
    int64_t so = 0, mo = 0;
    const size_t chunk = t->varseg_chunk;
    for (size_t r = 0; r < nread; r++) {
        if (nullptr == raws[r] || nsample[r] > (size_t)1 << 30) { sb2_set_error("read %zu: bad signal", r); return -1; }
        st.off[r] = so; moff[r] = mo; n[r] = (int)nsample[r];
        so += (int64_t)((nsample[r] + 3) / 4 * 4);
        mo += (int64_t)(chunk >= 2 ? nsample[r] / chunk : 0);
    }
    start.assign(nread, 0); end.assign(nread, 0);
    if (chunk < 2 || chunk > (size_t)1 << 20) return 0;                 // trim_raw_by_mad fails: every read is dropped
    if (dev_alloc(&st.d_raw, (size_t)so) || dev_alloc(&st.d_mads, (size_t)mo) || dev_alloc(&st.d_off, nread) ||
        dev_alloc(&st.d_moff, nread) || dev_alloc(&st.d_n, nread) || dev_alloc(&st.d_se, 2 * nread))
        return -1;
    for (size_t r = 0; r < nread; r++)
        CUDA_OK(cudaMemcpyAsync(st.d_raw + st.off[r], raws[r], nsample[r] * sizeof(float), cudaMemcpyHostToDevice, 0));
    CUDA_OK(cudaMemcpyAsync(st.d_off, st.off.data(), nread * sizeof(int64_t), cudaMemcpyHostToDevice, 0));
    CUDA_OK(cudaMemcpyAsync(st.d_moff, moff.data(), nread * sizeof(int64_t), cudaMemcpyHostToDevice, 0));
    CUDA_OK(cudaMemcpyAsync(st.d_n, n.data(), nread * sizeof(int), cudaMemcpyHostToDevice, 0));
    const int ts = (int)std::min(t->trim_start, (size_t)1 << 30), te = (int)std::min(t->trim_end, (size_t)1 << 30);
    launch_trim(st.d_raw, st.d_off, st.d_n, (int)nread, (int)chunk, t->varseg_thresh, ts, te, st.d_mads, st.d_moff, st.d_se, 0);
    eng->launches += 1;
    CUDA_OK(cudaGetLastError());
    std::vector<int> se(2 * nread);
    CUDA_OK(cudaMemcpy(se.data(), st.d_se, 2 * nread * sizeof(int), cudaMemcpyDeviceToHost));
    for (size_t r = 0; r < nread; r++) { start[r] = se[2 * r]; end[r] = se[2 * r + 1]; }
    return 0;
}

extern "C" int sb2_prepare_reads(sb2_engine *eng, const float *const *raws, const size_t *nsample, size_t nread,
                                 const sb2_trim *t, size_t *start_out, size_t *end_out, float *const *normalised) {
    if (nullptr == eng || nullptr == raws || nullptr == nsample || nullptr == t || 0 == nread) return -1;
    RawStage st;
    std::vector<int> start, end;
    if (0 != stage_and_trim(eng, raws, nsample, nread, t, st, start, end)) return -1;
    for (size_t r = 0; r < nread; r++) {
        if (start_out) start_out[r] = (size_t)start[r];
        if (end_out) end_out[r] = (size_t)end[r];
    }
    if (nullptr == normalised || nullptr == st.d_raw) return 0;
    std::vector<int64_t> src(nread);
    std::vector<int> n(nread);
    for (size_t r = 0; r < nread; r++) { src[r] = st.off[r] + start[r]; n[r] = end[r] - start[r]; }
    float *d_out = nullptr;
    int64_t *d_src = nullptr;
    int *d_n = nullptr;
    const size_t total = (size_t)(st.off[nread - 1] + (int64_t)((nsample[nread - 1] + 3) / 4 * 4));
    int rc = (dev_alloc(&d_out, total) || dev_alloc(&d_src, nread) || dev_alloc(&d_n, nread)) ? -1 : 0;
    if (0 == rc && (cudaMemcpy(d_src, src.data(), nread * sizeof(int64_t), cudaMemcpyHostToDevice) != cudaSuccess ||
                    cudaMemcpy(d_n, n.data(), nread * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess))
        rc = -1;
    if (0 == rc) {
        launch_medmad(st.d_raw, d_src, d_out, d_src, d_n, (int)nread, 0);
        eng->launches += 1;
        for (size_t r = 0; r < nread && 0 == rc; r++)
            if (n[r] > 0 && nullptr != normalised[r] &&
                cudaMemcpy(normalised[r], d_out + src[r], (size_t)n[r] * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess)
                rc = -1;
    }
    if (d_out) cudaFree(d_out);
    if (d_src) cudaFree(d_src);
    if (d_n) cudaFree(d_n);
    if (rc) sb2_set_error("prepare_reads: CUDA error %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

// calculate_post (src/scrappie_raw.c:265-315) for a batch of untrimmed pA signals: trim, normalise, network, decode,
// homopolymer fix-up and base strings, all on the device.  Reads that trim to nothing get out[r].bases = NULL.
extern "C" int sb2_basecall_raw_batch(sb2_engine *eng, enum raw_model_type model, const float *const *raws,
                                      const size_t *nsample, size_t nread, const sb2_trim *t, const sb2_params *p,
                                      sb2_call *out, size_t *start_out, size_t *end_out) {
    if (nullptr == eng || nullptr == raws || nullptr == nsample || nullptr == t || nullptr == p || nullptr == out || 0 == nread) return -1;
    for (size_t r = 0; r < nread; r++) out[r] = sb2_call{nullptr, NAN, 0, 0};
    RawStage st;
    std::vector<int> start, end;
    if (0 != stage_and_trim(eng, raws, nsample, nread, t, st, start, end)) return -1;
    const size_t min_len = min_read_samples(eng, model);
    if (0 == min_len) return -1;
    std::vector<size_t> keep, len;
    std::vector<int64_t> src;
    for (size_t r = 0; r < nread; r++) {
        if (start_out) start_out[r] = (size_t)start[r];
        if (end_out) end_out[r] = (size_t)end[r];
        // reads that trim to nothing, or to less than one convolution window, are dropped (bases NULL, score NAN):
        // one bad read must not lose the rest of the batch
        if (end[r] > start[r] && (size_t)(end[r] - start[r]) >= min_len) {
            keep.push_back(r); len.push_back((size_t)(end[r] - start[r])); src.push_back(st.off[r] + start[r]);
        }
    }
    if (keep.empty()) return 0;
    sb2_batch *b = pool_acquire(eng, model, len.data(), len.size());
    if (nullptr == b) return -1;
    int ncalled = -1;
    std::vector<sb2_call> calls(keep.size(), sb2_call{nullptr, NAN, 0, 0});
    bool ok = 0 == batch_shape(b, len.data(), len.size());
    if (ok && (nullptr == b->d_src || keep.size() > b->cap_reads)) ok = 0 == batch_alloc(b, &b->d_src, b->cap_reads);
    // the trimmer ran on the legacy stream: order this workspace's stream after it
    ok = ok && cudaStreamSynchronize(0) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(b->d_src, src.data(), keep.size() * sizeof(int64_t), cudaMemcpyHostToDevice, b->stream) == cudaSuccess;
    if (ok) {
        launch_medmad(st.d_raw, b->d_src, b->d_raw, b->d_sampoff, b->d_nsample, b->nread, b->stream);
        eng->launches += 1;
        ok = cudaStreamSynchronize(b->stream) == cudaSuccess;      // src (pageable) and st.d_raw must outlive the copies
    }
    if (ok) {
        ncalled = sb2_batch_basecall(b, nullptr, 0, p, calls.data());
        if (ncalled >= 0) for (size_t i = 0; i < keep.size(); i++) out[keep[i]] = calls[i];
    }
    pool_release(eng, b, ncalled >= 0);
    return ncalled;
}

// Time forward+decode of several batches running concurrently, each on its own stream
// (the way a job larger than one batch is executed).  The timed region is bracketed by
// events on the first batch's stream: every other stream waits for the start event and
// the end event waits for every stream.
extern "C" int sb2_multi_time(sb2_batch **batches, int nbatch, const sb2_params *p, int nrep, int flush_l2,
                              float *ms_out) {
    if (nullptr == batches || nbatch <= 0 || nullptr == p || nrep <= 0) return -1;
    sb2_engine *eng = batches[0]->eng;
    CUDA_OK(cudaSetDevice(eng->device));
    if (flush_l2 && nullptr == eng->flush_buf) {
        eng->flush_n = (size_t)96 << 20;
        if (dev_alloc(&eng->flush_buf, eng->flush_n)) return -1;
    }
    cudaStream_t main_s = batches[0]->stream;
    cudaEvent_t e0, e1;
    std::vector<cudaEvent_t> done(nbatch);
    CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
    for (auto &e : done) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    int rc = 0;
    for (int i = 0; i < nrep && 0 == rc; i++) {
        for (int k = 0; k < nbatch; k++) if (cudaStreamSynchronize(batches[k]->stream) != cudaSuccess) rc = -1;
        if (flush_l2) launch_flush(eng->flush_buf, eng->flush_n, main_s);
        cudaEventRecord(e0, main_s);
        for (int k = 1; k < nbatch; k++) cudaStreamWaitEvent(batches[k]->stream, e0, 0);
        for (int k = 0; k < nbatch; k++) {
            batches[k]->timing = (i == nrep - 1);
            rc |= sb2_batch_run(batches[k], p);
            batches[k]->timing = false;
            cudaEventRecord(done[k], batches[k]->stream);
        }
        for (int k = 1; k < nbatch; k++) cudaStreamWaitEvent(main_s, done[k], 0);
        cudaEventRecord(e1, main_s);
        if (cudaStreamSynchronize(main_s) != cudaSuccess) rc = -1;
        if (rc) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms_out) ms_out[i] = ms;
    }
    if (0 == rc)
        for (int k = 0; k < nbatch; k++)
            for (int st = 0; st < ST_COUNT; st++)
                cudaEventElapsedTime(&batches[k]->stage_ms[st], batches[k]->ev[st], batches[k]->ev[st + 1]);
    if (0 == rc)
        for (int k = 0; k < nbatch; k++)
            for (int st = 0; st <= ST_COUNT; st++)
                cudaEventElapsedTime(&batches[k]->stage_at[st], batches[0]->ev[0], batches[k]->ev[st]);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    for (auto &e : done) cudaEventDestroy(e);
    if (rc) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) sb2_set_error("CUDA error %s in timed run", cudaGetErrorString(e)); }
    return rc;
}

// Streaming throughput: every batch runs `nrep` steps back to back on its own stream with no host or
// cross-stream synchronisation between steps, so a batch's decode overlaps the other batches' next network
// pass the way a continuously fed basecaller runs.  One event pair brackets all nbatch * nrep runs; the
// per-step working set (GBs) is far larger than L2, so no flush is needed between steps.
extern "C" int sb2_multi_stream_time(sb2_batch **batches, int nbatch, const sb2_params *p, const int *nrep_per_batch,
                                     float *ms_total) {
    if (nullptr == batches || nbatch <= 0 || nullptr == p || nullptr == nrep_per_batch || nullptr == ms_total) return -1;
    int nrep = 0;
    for (int k = 0; k < nbatch; k++) nrep = std::max(nrep, nrep_per_batch[k]);
    if (nrep <= 0) return -1;
    sb2_engine *eng = batches[0]->eng;
    CUDA_OK(cudaSetDevice(eng->device));
    cudaStream_t main_s = batches[0]->stream;
    cudaEvent_t e0, e1;
    std::vector<cudaEvent_t> done(nbatch);
    CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
    for (auto &e : done) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    int rc = 0;
    for (int k = 0; k < nbatch; k++) if (cudaStreamSynchronize(batches[k]->stream) != cudaSuccess) rc = -1;
    cudaEventRecord(e0, main_s);
    for (int k = 1; k < nbatch; k++) cudaStreamWaitEvent(batches[k]->stream, e0, 0);
    for (int i = 0; i < nrep && 0 == rc; i++)
        for (int k = 0; k < nbatch; k++)
            if (i < nrep_per_batch[k]) rc |= sb2_batch_run(batches[k], p);
    for (int k = 1; k < nbatch; k++) {
        cudaEventRecord(done[k], batches[k]->stream);
        cudaStreamWaitEvent(main_s, done[k], 0);
    }
    cudaEventRecord(e1, main_s);
    if (cudaStreamSynchronize(main_s) != cudaSuccess) rc = -1;
    if (0 == rc) cudaEventElapsedTime(ms_total, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    for (auto &e : done) cudaEventDestroy(e);
    if (rc) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) sb2_set_error("CUDA error %s in timed run", cudaGetErrorString(e)); }
    return rc;
}

// ------------------------------------------------------------------------------------
// libscrappie single-read entry points (batches of one on a process-wide engine)
// ------------------------------------------------------------------------------------

static sb2_engine *default_engine() {
    static std::mutex mu;
    static sb2_engine *eng = nullptr;
    static bool tried = false;
    std::lock_guard<std::mutex> lock(mu);
    if (!tried) {
        tried = true;
        const char *dev = getenv("SCRAPPIE_B200_DEVICE");
        eng = sb2_engine_create(dev ? atoi(dev) : 0, nullptr);
    }
    return eng;
}

static scrappie_matrix posterior_single(enum raw_model_type model, const raw_table signal, float min_prob,
                                        float tempW, float tempb, bool return_log) {
    if (0 == signal.n || nullptr == signal.raw || signal.end <= signal.start) return nullptr;
    sb2_engine *eng = default_engine();
    if (nullptr == eng) return nullptr;
    const size_t n = signal.end - signal.start;
    sb2_batch *b = sb2_batch_create(eng, model, &n, 1);
    if (nullptr == b) return nullptr;
    sb2_params p = sb2_default_params();
    p.min_prob = min_prob; p.tempW = tempW; p.tempb = tempb;
    const float *sig = signal.raw + signal.start;
    scrappie_matrix post = nullptr;
    if (0 == sb2_batch_upload(b, &sig) && 0 == sb2_batch_forward(b, &p, return_log)) {
        post = make_scrappie_matrix(b->m->host.nstate, (size_t)b->nblock[0]);
        if (post && 0 != sb2_batch_download_posterior(b, 0, post->data.f, post->stride)) post = free_scrappie_matrix(post);
    }
    sb2_batch_destroy(b);
    return post;
}

extern "C" scrappie_matrix nanonet_rgrgr_r94_posterior(const raw_table signal, float min_prob, float tempW, float tempb, bool return_log) {
    return posterior_single(SCRAPPIE_MODEL_RGRGR_R9_4, signal, min_prob, tempW, tempb, return_log);
}
extern "C" scrappie_matrix nanonet_rgrgr_r941_posterior(const raw_table signal, float min_prob, float tempW, float tempb, bool return_log) {
    return posterior_single(SCRAPPIE_MODEL_RGRGR_R9_4_1, signal, min_prob, tempW, tempb, return_log);
}
extern "C" scrappie_matrix nanonet_rgrgr_r10_posterior(const raw_table signal, float min_prob, float tempW, float tempb, bool return_log) {
    return posterior_single(SCRAPPIE_MODEL_RGRGR_R10, signal, min_prob, tempW, tempb, return_log);
}
extern "C" scrappie_matrix nanonet_rnnrf_r94_transitions(const raw_table signal, float min_prob, float tempW, float tempb, bool return_log) {
    return posterior_single(SCRAPPIE_MODEL_RNNRF_R9_4, signal, min_prob, tempW, tempb, return_log);
}

// interface/scrappie.h:49-51
extern "C" scrappie_matrix nanonet_raw_posterior(const raw_table signal, float min_prob, float tempW, float tempb, bool return_log) {
    return posterior_single(SCRAPPIE_MODEL_RAW, signal, min_prob, tempW, tempb, return_log);
}

extern "C" posterior_function_ptr get_posterior_function(const enum raw_model_type model) {
    switch (model) {
    case SCRAPPIE_MODEL_RAW: return nanonet_raw_posterior;
    case SCRAPPIE_MODEL_RGRGR_R9_4: return nanonet_rgrgr_r94_posterior;
    case SCRAPPIE_MODEL_RGRGR_R9_4_1: return nanonet_rgrgr_r941_posterior;
    case SCRAPPIE_MODEL_RGRGR_R10: return nanonet_rgrgr_r10_posterior;
    case SCRAPPIE_MODEL_RNNRF_R9_4: return nanonet_rnnrf_r94_transitions;
    default:
        fprintf(stderr, "Invalid scrappie model %s:%d\n", __FILE__, __LINE__);
        exit(EXIT_FAILURE);
    }
}

// Decoders on a host matrix: upload, run the kernel, download the path.
struct DecodeScratch {
    float *d_post = nullptr, *d_score = nullptr;
    int *d_meta = nullptr, *d_tbE = nullptr, *d_path = nullptr;
    uint8_t *d_tb = nullptr;
    ~DecodeScratch() {
        void *ptrs[] = {d_post, d_score, d_meta, d_tbE, d_path, d_tb};
        for (void *p : ptrs) if (p) cudaFree(p);
    }
};

static int decode_single(const_scrappie_matrix m, bool crf, float stay_pen, float skip_pen, float local_pen,
                         bool allow_slip, int *seq, float *score_out) {
    sb2_engine *eng = default_engine();
    if (nullptr == eng) return -1;
    CUDA_OK(cudaSetDevice(eng->device));
    const size_t nb = m->nc, stride = m->stride;
    const int nstate = (int)m->nr;
    if (!crf && nstate != 1025 && nstate != 4097) { sb2_set_error("decode_transducer: %d states unsupported", nstate); return -1; }
    if (crf && nstate != 25) { sb2_set_error("decode_crf: expected 25 transition rows"); return -1; }
    DecodeScratch w;
    const size_t tb_bytes = crf ? nb * 8 : nb * (size_t)(nstate - 1);
    if (dev_alloc(&w.d_post, nb * stride) || dev_alloc(&w.d_score, 1) || dev_alloc(&w.d_meta, 4) ||
        dev_alloc(&w.d_tbE, nb) || dev_alloc(&w.d_path, nb + 1) || dev_alloc(&w.d_tb, tb_bytes))
        return -1;
    const int meta[4] = {(int)nb, 0, (int)nb, 0};        // nblock[0]; col_off[0..1]
    CUDA_OK(cudaMemcpy(w.d_meta, meta, sizeof(meta), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(w.d_post, m->data.f, nb * stride * sizeof(float), cudaMemcpyHostToDevice));
    BatchDims d{};
    d.nread = 1; d.total_cols = (int)nb; d.max_cols = (int)nb;
    d.nblock = w.d_meta; d.col_off = w.d_meta + 1;
    if (crf) launch_decode_crf(w.d_post, d, (int)stride, w.d_tb, w.d_path, w.d_score, 0);
    else launch_decode_transducer(w.d_post, d, nstate, (int)stride, stay_pen, skip_pen, local_pen, allow_slip ? 1 : 0,
                                  w.d_tb, w.d_tbE, w.d_path, w.d_score, 0);
    eng->launches += 1;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpy(seq, w.d_path, (nb + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(score_out, w.d_score, sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" float decode_transducer(const_scrappie_matrix logpost, float stay_pen, float skip_pen, float local_pen,
                                   int *seq, bool allow_slip) {
    if (nullptr == logpost || nullptr == seq) return NAN;
    float score = NAN;
    if (0 != decode_single(logpost, false, stay_pen, skip_pen, local_pen, allow_slip, seq, &score)) return NAN;
    return score;
}

extern "C" float decode_crf(const_scrappie_matrix trans, int *path) {
    if (nullptr == trans || nullptr == path) return NAN;
    float score = NAN;
    if (0 != decode_single(trans, true, 0.f, 0.f, 0.f, false, path, &score)) return NAN;
    return score;
}

// posterior_crf (src/decode.c:928-1012) on a host matrix of 25 transition energies per block
extern "C" scrappie_matrix posterior_crf(const_scrappie_matrix trans) {
    if (nullptr == trans) return nullptr;
    sb2_engine *eng = default_engine();
    if (nullptr == eng) return nullptr;
    if (cudaSetDevice(eng->device) != cudaSuccess) return nullptr;
    if (trans->nr != 25) { sb2_set_error("posterior_crf: expected 25 transition rows"); return nullptr; }
    const size_t nb = trans->nc, stride = trans->stride;
    scrappie_matrix post = make_scrappie_matrix(5, nb + 1);
    if (nullptr == post) return nullptr;
    float *d_trans = nullptr, *d_out = nullptr;
    int *d_meta = nullptr;
    bool ok = 0 == dev_alloc(&d_trans, nb * stride) && 0 == dev_alloc(&d_out, (nb + 1) * 8) && 0 == dev_alloc(&d_meta, 4);
    const int meta[4] = {(int)nb, 0, (int)nb, 0};        // nblock[0]; col_off[0..1]
    ok = ok && cudaMemcpy(d_meta, meta, sizeof(meta), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(d_trans, trans->data.f, nb * stride * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok) {
        BatchDims d{};
        d.nread = 1; d.total_cols = (int)nb; d.max_cols = (int)nb;
        d.nblock = d_meta; d.col_off = d_meta + 1;
        launch_posterior_crf(d_trans, d, (int)stride, d_out, 0);
        eng->launches += 1;
        ok = cudaGetLastError() == cudaSuccess &&
             cudaMemcpy2D(post->data.f, post->stride * sizeof(float), d_out, 8 * sizeof(float),
                          std::min((size_t)8, (size_t)post->stride) * sizeof(float), nb + 1, cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    if (d_trans) cudaFree(d_trans);
    if (d_out) cudaFree(d_out);
    if (d_meta) cudaFree(d_meta);
    if (!ok) {
        sb2_set_error("posterior_crf: CUDA error %s", cudaGetErrorString(cudaGetLastError()));
        return free_scrappie_matrix(post);
    }
    return post;
}

// The same for every read of an rnnrf_r94 batch, from the transitions the forward pass left on the device.
extern "C" int sb2_batch_posterior_crf(sb2_batch *b) {
    if (nullptr == b) return -1;
    const sb2_host_model &h = b->m->host;
    if (h.head != 1 || h.nstate != 25) { sb2_set_error("posterior_crf needs a CRF model (rnnrf_r94)"); return -1; }
    CUDA_OK(cudaSetDevice(b->eng->device));
    if (nullptr == b->d_bprob && batch_alloc(b, &b->d_bprob, (b->cap_cols + b->cap_reads) * 8)) return -1;
    launch_posterior_crf(b->d_post, b->dims, b->pstride, b->d_bprob, b->stream);
    b->eng->launches += 1;
    CUDA_OK(cudaGetLastError());
    return 0;
}

// dst: (nblock + 1) x 8 floats (the reference's 5-row matrix, stride 8)
extern "C" int sb2_batch_download_base_probs(sb2_batch *b, size_t read, float *dst) {
    if (nullptr == b || nullptr == dst || read >= (size_t)b->nread || nullptr == b->d_bprob) {
        sb2_set_error("base probabilities: run sb2_batch_posterior_crf first");
        return -1;
    }
    CUDA_OK(cudaSetDevice(b->eng->device));
    CUDA_OK(cudaStreamSynchronize(b->stream));
    CUDA_OK(cudaMemcpy(dst, b->d_bprob + ((size_t)b->col_off[read] + read) * 8, ((size_t)b->nblock[read] + 1) * 8 * sizeof(float),
                       cudaMemcpyDeviceToHost));
    return 0;
}

namespace {
struct EventsScratch {
    std::vector<void *> ptrs;
    template <typename T> T *get(size_t n) {
        T *p = nullptr;
        if (cudaMalloc(reinterpret_cast<void **>(&p), (n ? n : 1) * sizeof(T)) != cudaSuccess) return nullptr;
        ptrs.push_back(p);
        return p;
    }
    ~EventsScratch() { for (void *p : ptrs) cudaFree(p); }
};
}  // namespace

// ------------------------------------------------------------------------------------
// map_to_sequence_* (src/decode.c:1420-1964) on a host posterior
// ------------------------------------------------------------------------------------
static bool bounds_sane(const size_t *low, const size_t *high, size_t nblock, size_t seqlen) {
    // are_bounds_sane, src/decode.c:1638-1691
    if (nullptr == low || nullptr == high || 0 == nblock) return false;
    bool ok = (low[0] == 0) && (high[nblock - 1] == seqlen);
    for (size_t i = 0; i < nblock && ok; i++) ok = low[i] <= seqlen && high[i] <= seqlen && low[i] <= high[i];
    for (size_t i = 1; i < nblock && ok; i++) ok = low[i] <= high[i - 1] && low[i] >= low[i - 1] && high[i] >= high[i - 1];
    return ok;
}

extern "C" bool are_bounds_sane(size_t const *low, size_t const *high, size_t nblock, size_t seqlen) {
    const bool ok = bounds_sane(low, high, nblock, seqlen);
    if (!ok) fprintf(stderr, "scrappie_b200: banding structure is not valid\n");
    return ok;
}

static float map_to_sequence_run(const_scrappie_matrix logpost, float stay_pen, float skip_pen, float local_pen,
                                 const int *seq, size_t seqlen, bool forward, const size_t *low, const size_t *high,
                                 int *path) {
    if (nullptr == logpost || nullptr == seq || seqlen < 3 || 0 == logpost->nc || seqlen > (size_t)1 << 24) return NAN;
    const bool banded = (nullptr != low || nullptr != high);
    const size_t nb = logpost->nc, stride = logpost->stride;
    if (banded && !are_bounds_sane(low, high, nb, seqlen)) return NAN;
    for (size_t i = 0; i < seqlen; i++)
        if (seq[i] < 0 || (size_t)seq[i] + 1 >= logpost->nr) { sb2_set_error("map_to_sequence: state out of range"); return NAN; }
    sb2_engine *eng = default_engine();
    if (nullptr == eng || cudaSetDevice(eng->device) != cudaSuccess) return NAN;
    EventsScratch sc;
    float *d_lp = sc.get<float>(nb * stride), *d_buf = sc.get<float>(2 * (seqlen + 2)), *d_score = sc.get<float>(1);
    int *d_seq = sc.get<int>(seqlen), *d_low = nullptr, *d_high = nullptr, *d_path = nullptr;
    uint8_t *d_tb = nullptr, *d_tbe = nullptr;
    const bool want_path = (!forward && !banded && nullptr != path);
    if (want_path) { d_tb = sc.get<uint8_t>(nb * seqlen); d_tbe = sc.get<uint8_t>(nb); d_path = sc.get<int>(nb); }
    std::vector<int> lo32, hi32;
    if (banded) {
        lo32.assign(low, low + nb); hi32.assign(high, high + nb);
        d_low = sc.get<int>(nb); d_high = sc.get<int>(nb);
    }
    if (!d_lp || !d_buf || !d_score || !d_seq || (want_path && (!d_tb || !d_tbe || !d_path)) || (banded && (!d_low || !d_high))) {
        sb2_set_error("map_to_sequence: out of device memory");
        return NAN;
    }
    bool ok = cudaMemcpy(d_lp, logpost->data.f, nb * stride * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(d_seq, seq, seqlen * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok && banded)
        ok = cudaMemcpy(d_low, lo32.data(), nb * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(d_high, hi32.data(), nb * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
    float score = NAN;
    if (ok) {
        launch_map_to_sequence(d_lp, (int)nb, (int)logpost->nr, (int)stride, stay_pen, skip_pen, local_pen, d_seq, (int)seqlen,
                               d_low, d_high, forward ? 1 : 0, d_buf, d_tb, d_tbe, d_score, d_path, 0);
        eng->launches += 1;
        ok = cudaGetLastError() == cudaSuccess && cudaMemcpy(&score, d_score, sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess;
        if (ok && want_path) ok = cudaMemcpy(path, d_path, nb * sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    if (!ok) { sb2_set_error("map_to_sequence: CUDA error %s", cudaGetErrorString(cudaGetLastError())); return NAN; }
    return score;
}

extern "C" float map_to_sequence_viterbi(const_scrappie_matrix logpost, float stay_pen, float skip_pen, float local_pen,
                                         int const *seq, size_t seqlen, int *path) {
    return map_to_sequence_run(logpost, stay_pen, skip_pen, local_pen, seq, seqlen, false, nullptr, nullptr, path);
}
extern "C" float map_to_sequence_forward(const_scrappie_matrix logpost, float stay_pen, float skip_pen, float local_pen,
                                         int const *seq, size_t seqlen) {
    return map_to_sequence_run(logpost, stay_pen, skip_pen, local_pen, seq, seqlen, true, nullptr, nullptr, nullptr);
}
extern "C" float map_to_sequence_viterbi_banded(const_scrappie_matrix logpost, float stay_pen, float skip_pen, float local_pen,
                                                int const *seq, size_t seqlen, size_t const *poslow, size_t const *poshigh) {
    if (nullptr == poslow || nullptr == poshigh) return NAN;
    return map_to_sequence_run(logpost, stay_pen, skip_pen, local_pen, seq, seqlen, false, poslow, poshigh, nullptr);
}
extern "C" float map_to_sequence_forward_banded(const_scrappie_matrix logpost, float stay_pen, float skip_pen, float local_pen,
                                                int const *seq, size_t seqlen, size_t const *poslow, size_t const *poshigh) {
    if (nullptr == poslow || nullptr == poshigh) return NAN;
    return map_to_sequence_run(logpost, stay_pen, skip_pen, local_pen, seq, seqlen, true, poslow, poshigh, nullptr);
}

// ------------------------------------------------------------------------------------
// events (LSTM) model: nanonet_posterior (interface/scrappie.h:47-48, src/networks.c:146-194)
// ------------------------------------------------------------------------------------
static DevModel *get_events_model(sb2_engine *eng) {
    DevModel *dm = &eng->models[SB2_NMODEL];
    std::lock_guard<std::mutex> lock(eng->mu);
    if (dm->loaded) return dm;
    char path[1200];
    snprintf(path, sizeof(path), "%s/nanonet_events.bin", eng->weights_dir);
    void *blob = nullptr;
    size_t nbytes = 0;
    if (0 != sb2_read_file(path, &blob, &nbytes)) return nullptr;
    int rc = sb2_host_model_parse(blob, nbytes, &dm->host);
    free(blob);
    if (0 == rc && (dm->host.arch != 2 || dm->host.H != 96)) { sb2_set_error("events model: unsupported shape"); rc = -1; }
    if (0 == rc) rc = upload_model(eng, dm);
    if (0 != rc) { release_model(eng, dm); return nullptr; }
    return dm;
}


extern "C" int sb2_events_posterior_batch(sb2_engine *eng, const event_table *tables, size_t ntable, float min_prob,
                                          float tempW, float tempb, bool return_log, scrappie_matrix *out) {
    if (nullptr == eng || nullptr == tables || nullptr == out || 0 == ntable) return -1;
    if (!(min_prob >= 0.0f && min_prob <= 1.0f) || !(tempW > 0.0f) || !(tempb > 0.0f)) {
        sb2_set_error("invalid min_prob / temperature");
        return -1;
    }
    for (size_t i = 0; i < ntable; i++) out[i] = nullptr;
    CUDA_OK(cudaSetDevice(eng->device));
    DevModel *dm = get_events_model(eng);
    if (nullptr == dm) return -1;
    const sb2_host_model &h = dm->host;
    const int H = (int)h.H, FW = (int)h.ffw, NF = (int)h.nfilter, NS = (int)h.nstate, OS = (int)h.ostride;

    // features on the host, as the reference does (4 floats per event)
    std::vector<size_t> keep;
    std::vector<int> nblock, col_off(1, 0);
    std::vector<float> feats;
    for (size_t i = 0; i < ntable; i++) {
        if (0 == tables[i].n || nullptr == tables[i].event || tables[i].end <= tables[i].start) continue;   // NULL like the reference
        scrappie_matrix f = nanonet_features_from_events(tables[i], true);
        if (nullptr == f) continue;
        keep.push_back(i);
        nblock.push_back((int)f->nc);
        col_off.push_back(col_off.back() + (int)f->nc);
        feats.insert(feats.end(), f->data.f, f->data.f + f->nc * 4);
        free_scrappie_matrix(f);
    }
    if (keep.empty()) return 0;
    const size_t N = (size_t)col_off.back();
    EventsScratch sc;
    float *d_feat = sc.get<float>(N * 4), *d_f3 = sc.get<float>(N * NF);
    float *d_xf = sc.get<float>(N * 4 * H), *d_xb = sc.get<float>(N * 4 * H);
    float *d_hf = sc.get<float>(N * H), *d_hb = sc.get<float>(N * H);
    float *d_ff[2] = {sc.get<float>(N * FW), sc.get<float>(N * FW)};
    float *d_post = sc.get<float>(N * OS);
    int *d_nblock = sc.get<int>(keep.size()), *d_coloff = sc.get<int>(keep.size() + 1);
    if (!d_feat || !d_f3 || !d_xf || !d_xb || !d_hf || !d_hb || !d_ff[0] || !d_ff[1] || !d_post || !d_nblock || !d_coloff) {
        sb2_set_error("events posterior: out of device memory");
        return -1;
    }
    CUDA_OK(cudaMemcpy(d_feat, feats.data(), N * 4 * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d_nblock, nblock.data(), keep.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(d_coloff, col_off.data(), (keep.size() + 1) * sizeof(int), cudaMemcpyHostToDevice));
    BatchDims d{};
    d.nread = (int)keep.size(); d.total_cols = (int)N;
    d.max_cols = *std::max_element(nblock.begin(), nblock.end());
    d.nblock = d_nblock; d.col_off = d_coloff;
    cudaStream_t s = 0;
    launch_window3(d_feat, d_f3, d, s);
    const float *in = d_f3;
    int K = NF;
    uint64_t nl = 1;
    for (int pair = 0; pair < 2; pair++) {
        const int lf = 2 * pair, lb = 2 * pair + 1;
        launch_affine(in, (int)N, K, dm->iW[lf], K, dm->b[lf], 4 * H, d_xf, 4 * H, 1.0f, 1.0f, 0, 0, s);
        launch_affine(in, (int)N, K, dm->iW[lb], K, dm->b[lb], 4 * H, d_xb, 4 * H, 1.0f, 1.0f, 0, 0, s);
        if (0 != launch_lstm_scan(d_xf, dm->sW[lf], dm->sW2[lf], d_hf, d, H, 0, s) ||
            0 != launch_lstm_scan(d_xb, dm->sW[lb], dm->sW2[lb], d_hb, d, H, 1, s)) {
            sb2_set_error("events posterior: unsupported LSTM width");
            return -1;
        }
        // feedforward2_tanh: tanh(b + Wf lstmF + Wb lstmB)
        launch_affine(d_hf, (int)N, H, dm->comb_Wf[pair], H, dm->comb_b[pair], FW, d_ff[pair], FW, 1.0f, 1.0f, 0, 0, s);
        launch_affine(d_hb, (int)N, H, dm->comb_Wb[pair], H, nullptr, FW, d_ff[pair], FW, 1.0f, 1.0f, 2, 1, s);
        nl += 6;
        in = d_ff[pair];
        K = FW;
    }
    launch_affine(in, (int)N, FW, dm->FF_W, FW, dm->FF_b, NS, d_post, OS, tempW / tempb, tempb, 1, 0, s);
    launch_softmax_finish(d_post, (int)N, NS, OS, min_prob, return_log ? 1 : 0, s);
    eng->launches += nl + 2;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(s));
    int nmade = 0;
    for (size_t i = 0; i < keep.size(); i++) {
        scrappie_matrix post = make_scrappie_matrix((size_t)NS, (size_t)nblock[i]);
        if (nullptr == post) continue;
        if (cudaMemcpy2D(post->data.f, post->stride * sizeof(float), d_post + (size_t)col_off[i] * OS, OS * sizeof(float),
                         std::min((size_t)OS, (size_t)post->stride) * sizeof(float), (size_t)nblock[i],
                         cudaMemcpyDeviceToHost) != cudaSuccess) {
            free_scrappie_matrix(post);
            continue;
        }
        out[keep[i]] = post;
        nmade++;
    }
    return nmade;
}

extern "C" scrappie_matrix nanonet_posterior(const event_table events, float min_prob, float tempW, float tempb,
                                             bool return_log) {
    if (0 == events.n || nullptr == events.event) return nullptr;
    sb2_engine *eng = default_engine();
    if (nullptr == eng) return nullptr;
    scrappie_matrix post = nullptr;
    if (1 != sb2_events_posterior_batch(eng, &events, 1, min_prob, tempW, tempb, return_log, &post)) return nullptr;
    return post;
}

// ------------------------------------------------------------------------------------
// tensor-core self test (descriptor / layout validation and latency probe)
// ------------------------------------------------------------------------------------
extern "C" int sb2_tc_selftest(int K, int N, int reps, float *max_abs_err, long long *cycles3) {
    sb2_engine *eng = default_engine();
    if (nullptr == eng) return -1;
    CUDA_OK(cudaSetDevice(eng->device));
    std::vector<float> A((size_t)128 * K), B((size_t)N * K), D((size_t)128 * N);
    uint32_t seed = 12345u;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xFFFF) / 32768.0f - 1.0f; };
    for (auto &v : A) v = rnd() * 2.0f;
    for (auto &v : B) v = rnd();
    float *dA = nullptr, *dB = nullptr, *dD = nullptr;
    long long *dC = nullptr;
    if (dev_alloc(&dA, A.size()) || dev_alloc(&dB, B.size()) || dev_alloc(&dD, D.size()) || dev_alloc(&dC, 4)) return -1;
    CUDA_OK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemset(dD, 0, D.size() * 4));
    CUDA_OK(cudaMemset(dC, 0, 4 * sizeof(long long)));
    if (0 != launch_tc_selftest(dA, dB, dD, K, N, reps, dC, 0)) { sb2_set_error("selftest: cannot configure kernel"); return -1; }
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    long long cyc[4];
    CUDA_OK(cudaMemcpy(cyc, dC, sizeof(cyc), cudaMemcpyDeviceToHost));
    double worst = 0.0;
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < N; n++) {
            double acc = 0.0;
            for (int k = 0; k < K; k++) acc += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
            worst = std::max(worst, std::fabs(acc - (double)D[(size_t)m * N + n]));
        }
    if (max_abs_err) *max_abs_err = (float)worst;
    if (cycles3) { cycles3[0] = cyc[0]; cycles3[1] = cyc[1]; cycles3[2] = cyc[2]; }
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
    return 0;
}
