// extern "C" surface of libjat_b200.so (see include/jat_b200.h): argument checking, TMA descriptor
// construction, kernel dispatch, and the host-side launch plan of the whole DiT forward.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <map>
#include <functional>
#include <new>
#include <queue>
#include <set>
#include <vector>

#include "../../include/jat_b200.h"
#include "attention_gqa.cuh"
#include "attention_bwd.cuh"
#include "backward_elementwise.cuh"
#include "chunks.cuh"
#include "elementwise.cuh"
#include "gemm_tcgen05.cuh"
#include "optimizer.cuh"
#include "train_step.cuh"

using namespace jat;

// ------------------------------------------------------------------------------------------------ ctx
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct jat_ctx {
    int device;
    int sm_count;
    int gemm_sms;          // SMs the persistent GEMM grids occupy (sm_count minus the reserve for concurrent NCCL kernels)
    PFN_encodeTiled encode;
    std::atomic<long long> launches;
    int gemm_cta_pair;  // default tile configuration (overridable per call)
    int gemm_block_n;
    // optional per-launch CUDA-event profiling (jat_profile_*): one (start, stop) pair per launch
    bool profiling;
    std::vector<cudaEvent_t> ev_start, ev_stop;
    std::vector<int> ev_tag;
    size_t ev_used;
    cudaStream_t cur_stream;
    long long* att_trace;  // debug: device buffer of 128 clock64 slots for the attention kernel, or NULL
    // GEMM tail split (see GemmParams): partial-accumulator workspace (one 128x256 f32 tile per SM) + arrival counters
    float* tail_ws;
    int* tail_cnt;
    int tail_split;        // 0 = off; 1 = cut the tiles of a partial last wave along K, in-order fix-up (any epilogue);
                           // 2 = the same cut for the reduce-add epilogues only, parts added straight into the output
    long long* gemm_trace; // debug: device buffer of 64 x 8 clock64 slots for the GEMM kernel, or NULL
    // function attributes are per DEVICE: which kernels already carry their dynamic shared-memory limit on ctx->device, and
    // the co-resident cluster limit of the 4-CTA multicast GEMM instantiations there
    std::set<const void*> smem_configured;
    std::map<const void*, int> max_clusters;
    std::map<long long, int> attn_gs;   // (KV-group tiles, G) -> query heads per attention CTA
};
static const size_t kTailWsBytesPerSM = 128 * 256 * sizeof(float);

static const char* const kKernelTags[] = {"gemm_bias_act", "gemm_qkv_rope", "gemm_gate_residual", "gemm_unpatchify",
                                          "adaln_norm_modulate", "patchify_cast", "timestep_features",
                                          "cfg_euler_update", "gqa_attention_fwd", "chunk_normalize", "crossfade_denorm",
                                          "gemm_accum", "gemm_dact", "adaln_bwd", "gate_bwd", "colsum_cast", "attention_bwd",
                                          "train_glue", "optimizer"};
enum { TAG_GEMM0 = 0, TAG_ADALN = 4, TAG_PATCHIFY = 5, TAG_TSTEP = 6, TAG_EULER = 7, TAG_ATTN = 8, TAG_CHUNKN = 9,
       TAG_XFADE = 10, TAG_GEMM_ACCUM = 11, TAG_GEMM_DACT = 12, TAG_ADALN_BWD = 13,
       TAG_GATE_BWD = 14, TAG_COLSUM = 15, TAG_ATTN_BWD = 16, TAG_TRAIN_GLUE = 17, TAG_OPTIMIZER = 18, TAG_COUNT = 19 };

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}
#define JAT_CUDA(call)                                        \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)
#define JAT_TRY(call)              \
    do {                           \
        int r__ = (call);          \
        if (r__ != 0) return r__;  \
    } while (0)

// Every launching entry point makes ctx->device current for its duration (streams, function attributes and launches are
// per device; a process may hold one ctx per device) and restores the caller's device on return.
struct DeviceGuard {
    int prev;
    explicit DeviceGuard(const jat_ctx* c) : prev(-1) {
        int cur = -1;
        if (c != nullptr && cudaGetDevice(&cur) == cudaSuccess && cur != c->device) {
            prev = cur;
            cudaSetDevice(c->device);
        }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// cudaFuncAttributeMaxDynamicSharedMemorySize once per (ctx = device, kernel)
static int ensure_dyn_smem(jat_ctx* ctx, const void* kern, int bytes) {
    if (ctx->smem_configured.count(kern)) return 0;
    JAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    ctx->smem_configured.insert(kern);
    return 0;
}

extern "C" int jat_abi_version(void) { return JAT_ABI_VERSION; }
extern "C" const char* jat_last_error(void) { return g_err; }

extern "C" int jat_create(int device, jat_ctx** out) {
    if (!out) return fail(JAT_ERR_BAD_ARG, "jat_create: out == NULL");
    *out = nullptr;
    int prev_device = -1;
    cudaGetDevice(&prev_device);
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_device};  // leave the caller's device current
    JAT_CUDA(cudaSetDevice(device));
    JAT_CUDA(cudaFree(0));
    cudaDeviceProp prop;
    JAT_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(JAT_ERR_BAD_ARG, "jat_create: device %d is sm_%d%d, this library is sm_100a only",
                                      device, prop.major, prop.minor);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    JAT_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(JAT_ERR_NO_DRIVER, "cuTensorMapEncodeTiled not found");
    jat_ctx* c = new (std::nothrow) jat_ctx();
    if (!c) return fail(JAT_ERR_BAD_ARG, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->gemm_sms = c->sm_count;
    if (getenv("JAT_SM_RESERVE")) {
        const int r = atoi(getenv("JAT_SM_RESERVE"));
        if (r > 0 && r < c->sm_count - 8) c->gemm_sms = (c->sm_count - r) & ~1;
    }
    c->encode = (PFN_encodeTiled)fn;
    c->launches.store(0);
    c->gemm_cta_pair = 1;  // CTA pairs (cta_group::2) by default: half the B-operand smem traffic per SM
    if (getenv("JAT_GEMM_CLUSTER")) c->gemm_cta_pair = atoi(getenv("JAT_GEMM_CLUSTER")) >= 2 ? 2 : 1;
    c->gemm_block_n = 0;
    c->profiling = false;
    c->ev_used = 0;
    c->cur_stream = nullptr;
    c->att_trace = nullptr;
    c->tail_ws = nullptr;
    c->tail_cnt = nullptr;
    // default 2: the tiles of a partial last wave of the inference out_proj / fc2 GEMMs are cut along K and every part reduce-adds
    // its partial sum into the residual stream (gate-residual class 0.84 -> 0.91 of the sustained bf16 peak); the f32 adds of
    // one tile's parts land in arrival order, i.e. ~3 % of the output tiles can differ in the last bit from run to run (and,
    // through the bf16 rounding of the next GEMM operand, the model output by a few 1e-5 relative).
    // JAT_GEMM_TAIL=0 / jat_set_gemm_tail_split(ctx, 0) / torch.use_deterministic_algorithms(True) (honoured by the Python
    // engine) give the bit-reproducible schedule.
    c->tail_split = getenv("JAT_GEMM_TAIL") ? atoi(getenv("JAT_GEMM_TAIL")) : 2;
    if (c->tail_split < 0 || c->tail_split > 2) c->tail_split = 0;
    c->gemm_trace = nullptr;
    const size_t cnt_bytes = (size_t)c->sm_count * GEMM_EPI_WARPS * sizeof(int);
    if (cudaMalloc(&c->tail_ws, kTailWsBytesPerSM * c->sm_count) != cudaSuccess ||
        cudaMalloc(&c->tail_cnt, cnt_bytes) != cudaSuccess || cudaMemset(c->tail_cnt, 0, cnt_bytes) != cudaSuccess) {
        cudaFree(c->tail_ws);
        cudaFree(c->tail_cnt);
        delete c;
        return cuda_fail(cudaGetLastError(), "jat_create: GEMM tail-split workspace");
    }
    *out = c;
    return 0;
}
extern "C" void jat_destroy(jat_ctx* ctx) {
    if (!ctx) return;
    cudaFree(ctx->tail_ws);
    cudaFree(ctx->tail_cnt);
    for (cudaEvent_t e : ctx->ev_start) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_stop) cudaEventDestroy(e);
    delete ctx;
}
extern "C" int jat_sm_count(const jat_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

extern "C" int jat_set_gemm_sm_reserve(jat_ctx* ctx, int reserve) {
    if (!ctx) return fail(JAT_ERR_BAD_ARG, "ctx == NULL");
    if (reserve < 0 || reserve > ctx->sm_count - 8) return fail(JAT_ERR_BAD_ARG, "jat_set_gemm_sm_reserve: reserve out of range");
    ctx->gemm_sms = (ctx->sm_count - reserve) & ~1;
    return 0;
}
extern "C" int64_t jat_launch_count(const jat_ctx* ctx) { return ctx ? (int64_t)ctx->launches.load() : 0; }
extern "C" int jat_set_gemm_config(jat_ctx* ctx, int cta_pair, int block_n) {
    if (!ctx) return fail(JAT_ERR_BAD_ARG, "ctx == NULL");
    if (block_n != 0 && block_n != 128 && block_n != 256) return fail(JAT_ERR_BAD_ARG, "block_n must be 0/128/256");
    ctx->gemm_cta_pair = cta_pair < 0 ? 1 : (cta_pair > 2 ? 2 : cta_pair);
    ctx->gemm_block_n = block_n;
    return 0;
}

extern "C" int jat_debug_set_gemm_trace(jat_ctx* ctx, void* buf) {
    if (!ctx) return fail(JAT_ERR_BAD_ARG, "ctx == NULL");
    ctx->gemm_trace = (long long*)buf;
    return 0;
}

extern "C" int jat_set_gemm_tail_split(jat_ctx* ctx, int enable) {
    if (!ctx) return fail(JAT_ERR_BAD_ARG, "ctx == NULL");
    ctx->tail_split = enable == 2 ? 2 : (enable ? 1 : 0);
    return 0;
}

// Debug aid: timestamps (clock64) of the attention kernel's pipeline events, CTA (0,0,0); buf = 128 x int64 or NULL.
extern "C" int jat_debug_set_attention_trace(jat_ctx* ctx, void* buf) {
    if (!ctx) return fail(JAT_ERR_BAD_ARG, "ctx == NULL");
    ctx->att_trace = (long long*)buf;
    return 0;
}

// Profiling mode: bracket the launch with a CUDA-event pair on the launching stream.
static void pre_launch(jat_ctx* ctx, int tag, cudaStream_t s) {
    if (!ctx->profiling) return;
    if (ctx->ev_used == ctx->ev_start.size()) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { ctx->profiling = false; return; }
        ctx->ev_start.push_back(a);
        ctx->ev_stop.push_back(b);
        ctx->ev_tag.push_back(tag);
    }
    ctx->ev_tag[ctx->ev_used] = tag;
    ctx->cur_stream = s;
    cudaEventRecord(ctx->ev_start[ctx->ev_used], s);
}
static int post_launch(jat_ctx* ctx, const char* name) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "launch %s: %s", name, cudaGetErrorString(e));
        return (int)e;
    }
    if (ctx->profiling) {
        cudaEventRecord(ctx->ev_stop[ctx->ev_used], ctx->cur_stream);
        ctx->ev_used++;
    }
    ctx->launches.fetch_add(1);
    return 0;
}

extern "C" int jat_profile_begin(jat_ctx* ctx) {
    if (!ctx) return fail(JAT_ERR_BAD_ARG, "ctx == NULL");
    ctx->profiling = true;
    ctx->ev_used = 0;
    return 0;
}
// Synchronises the device, then fills per-kernel-class totals; returns the number of classes.
extern "C" int jat_profile_end(jat_ctx* ctx, int max_tags, const char** names, double* total_ms, int64_t* counts) {
    if (!ctx || !names || !total_ms || !counts) return fail(JAT_ERR_BAD_ARG, "jat_profile_end: null argument");
    ctx->profiling = false;
    JAT_CUDA(cudaDeviceSynchronize());
    const int n = max_tags < TAG_COUNT ? max_tags : TAG_COUNT;
    for (int i = 0; i < n; ++i) { names[i] = kKernelTags[i]; total_ms[i] = 0.0; counts[i] = 0; }
    for (size_t i = 0; i < ctx->ev_used; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev_start[i], ctx->ev_stop[i]) != cudaSuccess) continue;
        const int t = ctx->ev_tag[i];
        if (t < n) { total_ms[t] += ms; counts[t]++; }
    }
    ctx->ev_used = 0;
    return n;
}

// Launch with programmatic stream serialization (see pdl_wait() in common.cuh): ONLY for kernels that call pdl_wait()
// before their first global-memory access.  OFF by default: measured on the sampling step it is 1-2 % SLOWER than plain
// stream order (the successor's early-resident CTAs cost more than the ~3 us launch gaps they hide); JAT_PDL=1 enables it.
static bool pdl_enabled() {
    static const bool on = getenv("JAT_PDL") && atoi(getenv("JAT_PDL")) != 0;
    return on;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// 2D tensor map: `rows` x `cols` (cols contiguous), row pitch ld elements; box = box_rows x box_cols with
// box_cols * elem_bytes == 128 bytes, 128-byte swizzle, out-of-bounds elements read as zero / not written.
static int make_tmap(jat_ctx* ctx, CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                     uint32_t box_rows, bool f32 = false) {
    const uint64_t esz = f32 ? 4 : 2;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * esz) % 16 != 0)
        return fail(JAT_ERR_BAD_ARG, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {ld * esz};
    cuuint32_t box[2] = {(cuuint32_t)(128 / esz), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                             const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(JAT_ERR_TENSORMAP, "cuTensorMapEncodeTiled failed (%d): rows %llu cols %llu ld %llu box_rows %u",
                                       (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows);
    return 0;
}

// ------------------------------------------------------------------------------------------------ dropout
static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
extern "C" uint32_t jat_dropout_site_seed(uint64_t seed, int block, int site) {
    const uint64_t h = splitmix64(splitmix64(seed) ^ (((uint64_t)(uint32_t)block << 8) | (uint64_t)(uint32_t)(site & 0xff)));
    return (uint32_t)(h >> 32) ^ (uint32_t)h;
}
// p -> DropCfg; returns false if p is outside [0, 1)
static bool make_drop(float p, uint32_t seed, DropCfg* d) {
    d->thresh = 0u; d->seed = seed; d->inv_keep = 1.0f;
    if (!(p >= 0.0f) || p >= 1.0f) return false;
    if (p == 0.0f) return true;
    double th = (double)p * 65536.0 + 0.5;  // 16-bit lanes: two mask elements per hash (common.cuh: drop_scale)
    d->thresh = th >= 65535.0 ? 65535u : (th < 1.0 ? 1u : (uint32_t)th);
    d->inv_keep = (float)(1.0 / (1.0 - (double)d->thresh / 65536.0));
    return true;
}

extern "C" int jat_dropout_scale_mask(jat_ctx* ctx, float* out, int64_t rows, int cols, float p, uint32_t site_seed,
                                      void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !out || rows <= 0 || cols <= 0 || rows > 0xffffffffll) return fail(JAT_ERR_BAD_ARG, "jat_dropout_scale_mask: bad argument");
    DropCfg d;
    if (!make_drop(p, site_seed, &d)) return fail(JAT_ERR_BAD_ARG, "jat_dropout_scale_mask: p must be in [0, 1)");
    const long long n = (long long)rows * cols;
    pre_launch(ctx, TAG_COLSUM, (cudaStream_t)stream);
    dropout_scale_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out, n, cols, d);
    return post_launch(ctx, "dropout_scale_mask");
}

extern "C" int jat_drop_path_scales(jat_ctx* ctx, float* out, const float* rates, int depth, int B, uint64_t seed,
                                    void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !out || !rates || depth <= 0 || B <= 0) return fail(JAT_ERR_BAD_ARG, "jat_drop_path_scales: bad argument");
    const uint32_t s32 = jat_dropout_site_seed(seed, -1, JAT_DROP_SITE_PATH);
    const int n = depth * 2 * B;
    pre_launch(ctx, TAG_COLSUM, (cudaStream_t)stream);
    drop_path_scales_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(out, rates, depth, B, s32);
    return post_launch(ctx, "drop_path_scales");
}

// ------------------------------------------------------------------------------------------------ GEMM
template <int BN, int CG, int EPI, int ACT, int OUT_BF16, int A_MN = 0, int B_MN = 0, int MC = 1, int EW = GEMM_EPI_WARPS>
static int launch_gemm(jat_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                       const GemmParams& p, cudaStream_t s) {
    using Cfg = GemmCfg<BN, CG>;
    auto kern = gemm_tcgen05_kernel<BN, CG, EPI, ACT, OUT_BF16, A_MN, B_MN, MC, EW>;
    JAT_TRY(ensure_dyn_smem(ctx, (const void*)kern, Cfg::SMEM_BYTES));
    int clusters = ctx->gemm_sms / (CG * MC);
    const int num_work = p.head_tiles * p.k_splits + (p.num_tiles - p.head_tiles) * p.tail_splits;
    if (clusters > num_work) clusters = num_work;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * CG * MC));
    cfg.blockDim = dim3(128 + EW * 32);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG * MC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if constexpr (MC > 1) {
        // clusters of 4 must sit inside one GPC: fewer than sm_count / 4 may be co-resident; a persistent grid larger than
        // that would run its surplus clusters as a second wave
        int& max_clusters = ctx->max_clusters.emplace((const void*)kern, -1).first->second;
        if (max_clusters < 0) {
            cfg.gridDim = dim3((unsigned)(ctx->sm_count / (CG * MC) * CG * MC));
            if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters <= 0) {
                cudaGetLastError();
                max_clusters = ctx->sm_count / (CG * MC);
            }
            if (getenv("JAT_DEBUG")) fprintf(stderr, "jat: max active clusters of %d CTAs: %d\n", CG * MC, max_clusters);
        }
        if (clusters > max_clusters) clusters = max_clusters;
        cfg.gridDim = dim3((unsigned)(clusters * CG * MC));
    }
    pre_launch(ctx, EPI <= EPI_UNPATCHIFY ? TAG_GEMM0 + EPI : (EPI == EPI_ACCUM ? TAG_GEMM_ACCUM : TAG_GEMM_DACT), s);
    JAT_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, p));
    return post_launch(ctx, "gemm_tcgen05");
}

// forward (K-major) GEMMs on clusters of two CTA pairs with the W tile multicast between the pairs
static int dispatch_gemm_mc2(jat_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                             const GemmParams& p, const jat_gemm_epilogue* e, cudaStream_t s) {
    switch (e->kind) {
        case JAT_EPI_BIAS_ACT:
            if (e->act == JAT_ACT_NONE && e->out_dtype == JAT_DTYPE_F32)
                return launch_gemm<256, 2, EPI_BIAS_ACT, ACT_NONE, 0, 0, 0, 2>(ctx, ta, tb, to, p, s);
            if (e->act == JAT_ACT_NONE && e->out_dtype == JAT_DTYPE_BF16)
                return launch_gemm<256, 2, EPI_BIAS_ACT, ACT_NONE, 1, 0, 0, 2>(ctx, ta, tb, to, p, s);
            if (e->act == JAT_ACT_GELU_ERF && e->out_dtype == JAT_DTYPE_BF16)
                return launch_gemm<256, 2, EPI_BIAS_ACT, ACT_GELU, 1, 0, 0, 2>(ctx, ta, tb, to, p, s);
            if (e->act == JAT_ACT_SILU && e->out_dtype == JAT_DTYPE_BF16)
                return launch_gemm<256, 2, EPI_BIAS_ACT, ACT_SILU, 1, 0, 0, 2>(ctx, ta, tb, to, p, s);
            break;
        case JAT_EPI_QKV_ROPE:
            return launch_gemm<256, 2, EPI_QKV_ROPE, ACT_NONE, 1, 0, 0, 2>(ctx, ta, tb, to, p, s);
        case JAT_EPI_GATE_RESIDUAL:
            return launch_gemm<256, 2, EPI_GATE_RESIDUAL, ACT_NONE, 0, 0, 0, 2>(ctx, ta, tb, to, p, s);
        case JAT_EPI_UNPATCHIFY:
            return launch_gemm<256, 2, EPI_UNPATCHIFY, ACT_NONE, 0, 0, 0, 2>(ctx, ta, tb, to, p, s);
    }
    return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: epilogue (%d, act %d, dtype %d) has no multicast-cluster variant", e->kind, e->act,
                e->out_dtype);
}

// 16 epilogue warps (see gemm_tcgen05_kernel) for the bias + GELU and the GELU' dgrad epilogues: JAT_GEMM_EW16 selects when -- 0 never, 1 when the
// epilogue carries the training extras (pre-activation copy or dropout: ~33 instructions per element; fc1 forward of the
// training step 135 -> 113 us, class 8.75 -> 8.14 ms per step), 2 always.
// Not combined with the workspace tail split (its per-warp regions are laid out for 8 warps).
static bool gemm_ew16(const GemmParams& p) {
    static const int mode = getenv("JAT_GEMM_EW16") ? atoi(getenv("JAT_GEMM_EW16")) : 1;
    if (mode <= 0 || (p.tail_splits > 1 && !p.tail_direct)) return false;
    return mode >= 2 || p.aux != nullptr || p.drop.thresh != 0u;
}

template <int BN, int CG>
static int dispatch_gemm_epi(jat_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                             const GemmParams& p, const jat_gemm_epilogue* e, cudaStream_t s) {
    const int amn = e->a_transposed ? 1 : 0, bmn = e->w_transposed ? 1 : 0;
    if (amn && bmn) {  // wgrad
        if (e->kind == JAT_EPI_ACCUM) return launch_gemm<BN, CG, EPI_ACCUM, ACT_NONE, 0, 1, 1>(ctx, ta, tb, to, p, s);
        return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: both operands transposed is supported with JAT_EPI_ACCUM only");
    }
    if (bmn) {  // dgrad
        switch (e->kind) {
            case JAT_EPI_BIAS_ACT:
                if (e->act == JAT_ACT_NONE && e->out_dtype == JAT_DTYPE_BF16)
                    return launch_gemm<BN, CG, EPI_BIAS_ACT, ACT_NONE, 1, 0, 1>(ctx, ta, tb, to, p, s);
                if (e->act == JAT_ACT_NONE && e->out_dtype == JAT_DTYPE_F32)
                    return launch_gemm<BN, CG, EPI_BIAS_ACT, ACT_NONE, 0, 0, 1>(ctx, ta, tb, to, p, s);
                break;
            case JAT_EPI_ACCUM:
                return launch_gemm<BN, CG, EPI_ACCUM, ACT_NONE, 0, 0, 1>(ctx, ta, tb, to, p, s);
            case JAT_EPI_DACT:
                if constexpr (BN == 256 && CG == 2) {
                    // (GELU' dgrad class 3.90 -> 3.61 ms per training step; pre-activations loaded per 32-column half)
                    if (e->act == JAT_ACT_GELU_ERF && gemm_ew16(p))
                        return launch_gemm<256, 2, EPI_DACT, ACT_GELU, 1, 0, 1, 1, 16>(ctx, ta, tb, to, p, s);
                }
                if (e->act == JAT_ACT_GELU_ERF) return launch_gemm<BN, CG, EPI_DACT, ACT_GELU, 1, 0, 1>(ctx, ta, tb, to, p, s);
                if (e->act == JAT_ACT_SILU) return launch_gemm<BN, CG, EPI_DACT, ACT_SILU, 1, 0, 1>(ctx, ta, tb, to, p, s);
                break;
        }
        return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: unsupported epilogue (%d, act %d, dtype %d) with w_transposed", e->kind,
                    e->act, e->out_dtype);
    }
    if (amn) return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: a_transposed needs w_transposed");
    switch (e->kind) {
        case JAT_EPI_BIAS_ACT:
            if (e->act == JAT_ACT_NONE && e->out_dtype == JAT_DTYPE_F32)
                return launch_gemm<BN, CG, EPI_BIAS_ACT, ACT_NONE, 0>(ctx, ta, tb, to, p, s);
            if (e->act == JAT_ACT_NONE && e->out_dtype == JAT_DTYPE_BF16)
                return launch_gemm<BN, CG, EPI_BIAS_ACT, ACT_NONE, 1>(ctx, ta, tb, to, p, s);
            if constexpr (BN == 256 && CG == 2) {
                if (e->act == JAT_ACT_GELU_ERF && e->out_dtype == JAT_DTYPE_BF16 && gemm_ew16(p))
                    return launch_gemm<256, 2, EPI_BIAS_ACT, ACT_GELU, 1, 0, 0, 1, 16>(ctx, ta, tb, to, p, s);
            }
            if (e->act == JAT_ACT_GELU_ERF && e->out_dtype == JAT_DTYPE_BF16)
                return launch_gemm<BN, CG, EPI_BIAS_ACT, ACT_GELU, 1>(ctx, ta, tb, to, p, s);
            if (e->act == JAT_ACT_SILU && e->out_dtype == JAT_DTYPE_BF16)
                return launch_gemm<BN, CG, EPI_BIAS_ACT, ACT_SILU, 1>(ctx, ta, tb, to, p, s);
            return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: unsupported act/out_dtype combination (%d, %d)", e->act,
                        e->out_dtype);
        case JAT_EPI_QKV_ROPE:
            return launch_gemm<BN, CG, EPI_QKV_ROPE, ACT_NONE, 1>(ctx, ta, tb, to, p, s);
        case JAT_EPI_GATE_RESIDUAL:
            return launch_gemm<BN, CG, EPI_GATE_RESIDUAL, ACT_NONE, 0>(ctx, ta, tb, to, p, s);
        case JAT_EPI_UNPATCHIFY:
            return launch_gemm<BN, CG, EPI_UNPATCHIFY, ACT_NONE, 0>(ctx, ta, tb, to, p, s);
    }
    return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: unsupported epilogue kind %d for untransposed operands", e->kind);
}

extern "C" int jat_gemm_bf16(jat_ctx* ctx, const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                             const jat_gemm_epilogue* e, int cta_pair, int block_n, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !A || !W || !e || !e->out) return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: null argument");
    if (M <= 0 || N <= 0 || K <= 0) return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: non-positive size");
    const bool a_mn = e->a_transposed != 0, w_mn = e->w_transposed != 0;
    if ((K % 64 != 0 && !(a_mn && w_mn)) || N % 128 != 0)
        return fail(JAT_ERR_BAD_SHAPE, "jat_gemm_bf16: need K %% 64 == 0 and N %% 128 == 0 (got N=%d K=%d)", N, K);
    const bool auto_cfg = cta_pair < 0 && block_n == 0 && ctx->gemm_block_n == 0 && ctx->gemm_cta_pair == 1 && !(a_mn && w_mn);
    int auto_ks = 0;   // split-K factor chosen by the small-M model below (0: the caller's)
    if (cta_pair < 0) cta_pair = ctx->gemm_cta_pair;
    if (block_n == 0) block_n = ctx->gemm_block_n;
    if (block_n == 0) block_n = (N % 256 == 0) ? 256 : 128;
    if (auto_cfg && block_n == 256) {
        // Small-M regime (B = 1 inference: M = 345-690 token rows, BASELINE configs[1]): the default 256 x 256 CTA-pair tiles
        // leave most of the machine idle (out_proj / fc2 of the v2 model: 12 tiles on 74 pairs) and the launch lasts as long
        // as ONE tile's K loop.  Makespan model per candidate: rounds of the persistent schedule x (k-blocks x MMA time of a
        // k-block + epilogue); small tiles pay ~15 % (128-wide: operand stream per FLOP doubles) / ~10 % (single CTA: no
        // W-tile sharing).  The default is kept unless a candidate is at least 20 % cheaper -- large problems never switch.
        const int kblocks = (K + GEMM_BK - 1) / GEMM_BK;
        const int ks0 = e->k_splits > 1 ? e->k_splits : 1;
        // split-K is offered to the inference gate-residual GEMMs (out_proj / fc2 at small M: one tile's K loop IS the
        // launch) under the same policy as the tail split: allowed unless the bit-reproducible schedule was asked for
        // (the parts' f32 adds into the residual stream land in arrival order)
        const bool may_split = ks0 == 1 && ctx->tail_split == 2 && e->kind == JAT_EPI_GATE_RESIDUAL && e->aux == nullptr &&
                               !(e->drop_p > 0.0f) && !w_mn;
        auto cost = [&](int cg_, int bn_, int ks) -> double {
            const long long items = (long long)((M + 128 * cg_ - 1) / (128 * cg_)) * (N / bn_) * ks;
            const long long workers = ctx->gemm_sms / cg_;
            const long long rounds = (items + workers - 1) / workers;
            const double kb_clk = 2.0 * bn_ * (bn_ == 128 ? 1.15 : 1.0) * (cg_ == 1 ? 1.10 : 1.0);
            return (double)rounds * ((double)((kblocks + ks - 1) / ks) * kb_clk + 8.0 * bn_);
        };
        const double base = cost(2, 256, ks0);
        double best = base * 0.8;
        const int cand[4][2] = {{2, 256}, {2, 128}, {1, 256}, {1, 128}};
        for (int i = 0; i < 4; ++i) {
            for (int ks = ks0; ks <= (may_split ? 4 : ks0); ++ks) {
                if (i == 0 && ks == ks0) continue;   // the default itself
                if (ks > 1 && kblocks / ks < 8) break;
                const double c = cost(cand[i][0], cand[i][1], ks);
                if (c < best) { best = c; cta_pair = cand[i][0] == 2 ? 1 : 0; block_n = cand[i][1]; auto_ks = ks; }
            }
        }
    }
    if (block_n != 128 && block_n != 256) return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: block_n must be 128 or 256");
    if (N % block_n != 0) block_n = 128;
    const int cg = cta_pair ? 2 : 1;
    // cta_pair == 2: clusters of two CTA pairs sharing the W tile by TMA multicast (forward layouts, 256-wide tiles, no
    // split-K); anything else falls back to independent CTA pairs
    const bool dbg_skip_on = getenv("JAT_DBG_GEMM_SKIP") != nullptr && atoi(getenv("JAT_DBG_GEMM_SKIP")) != 0;
    const int mc = (cta_pair == 2 && block_n == 256 && !a_mn && !w_mn && e->k_splits <= 1 && e->kind != JAT_EPI_ACCUM &&
                    e->kind != JAT_EPI_DACT && M > 256 * (ctx->sm_count / 4) && !dbg_skip_on) ? 2 : 1;

    GemmParams p = {};
    p.M = M; p.N = N; p.K = K;
    p.num_n_blocks = N / block_n;
    p.num_k_blocks = (K + GEMM_BK - 1) / GEMM_BK;
    const int rows_per_tile = GEMM_BM * cg * mc;
    p.num_tiles = ((M + rows_per_tile - 1) / rows_per_tile) * p.num_n_blocks;
    p.bias = e->bias;
    p.out = e->out;
    p.ldo = e->ldo;
    p.gate = e->gate;
    p.gate_bstride = e->gate_batch_stride;
    p.gate_rowscale = e->gate_rowscale;
    if (!make_drop(e->drop_p, e->drop_seed, &p.drop)) return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: drop_p must be in [0, 1)");
    if (p.drop.thresh != 0u && !(e->kind == JAT_EPI_GATE_RESIDUAL || e->kind == JAT_EPI_DACT ||
                                 (e->kind == JAT_EPI_BIAS_ACT && e->out_dtype == JAT_DTYPE_BF16)))
        return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: dropout is fused into the BIAS_ACT(bf16) / GATE_RESIDUAL / DACT epilogues only");
    if (p.drop.thresh != 0u && e->k_splits > 1) return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: dropout cannot be combined with k_splits");
    p.tokens_per_batch = e->tokens_per_batch > 0 ? e->tokens_per_batch : M;
    p.rope_cos = e->rope_cos;
    p.rope_sin = e->rope_sin;
    p.rope_cols = e->rope_cols;
    p.t_out = e->t_out;
    p.aux = e->aux;
    p.ld_aux = e->ld_aux;
    p.k_splits = auto_ks > 0 ? auto_ks : (e->k_splits > 1 ? e->k_splits : 1);
    if (p.k_splits > 1 && e->kind != JAT_EPI_GATE_RESIDUAL && e->kind != JAT_EPI_ACCUM)
        return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: k_splits needs a reduce-add epilogue (GATE_RESIDUAL / ACCUM)");
    if (p.k_splits > p.num_k_blocks) p.k_splits = p.num_k_blocks;
    // Tail split: when the last wave of the persistent schedule is partial, its tiles are cut along K so that the
    // wave fills the machine (e.g. out_proj / fc2: 380 tiles on 74 CTA pairs = 5 waves + 10 tiles -> 10 x 7 parts).
    {
        static const int dbg = getenv("JAT_DBG_GEMM_SKIP") ? atoi(getenv("JAT_DBG_GEMM_SKIP")) : 0;
        p.dbg_skip = dbg;
        p.trace = ctx->gemm_trace;
    }
    p.head_tiles = p.num_tiles;
    p.tail_splits = 1;
    p.tail_direct = 0;
    p.tail_ws = ctx->tail_ws;
    p.tail_cnt = ctx->tail_cnt;
    {
        const int clusters = ctx->gemm_sms / cg;
        const int rem = p.num_tiles % clusters;
        // mode 2: reduce-add epilogues without a pre-gate copy only -- every part runs the ordinary epilogue on its
        // partial sum (bias with part 0), so there is no workspace round trip; the f32 adds of the parts of one tile land
        // in arrival order (last-bit run-to-run differences in those tiles)
        const bool direct = ctx->tail_split == 2;
        const bool direct_ok = (e->kind == JAT_EPI_GATE_RESIDUAL || e->kind == JAT_EPI_ACCUM) && e->aux == nullptr;
        if (ctx->tail_split && (!direct || direct_ok) && mc == 1 && p.k_splits == 1 && p.num_tiles > clusters && rem > 0) {
            int splits = clusters / rem;
            if (splits > 8) splits = 8;
            if (splits > p.num_k_blocks / 2) splits = p.num_k_blocks / 2;
            if (splits >= 2) { p.head_tiles = p.num_tiles - rem; p.tail_splits = splits; p.tail_direct = direct ? 1 : 0; }
        }
    }

    switch (e->kind) {
        case JAT_EPI_BIAS_ACT:
            if (e->ldo % 8 != 0) return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: ldo must be a multiple of 8");
            break;
        case JAT_EPI_QKV_ROPE:
            if (!e->rope_cos || !e->rope_sin || e->rope_cols % 64 != 0 || e->ldo % 8 != 0)
                return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: QKV_ROPE needs cos/sin tables and rope_cols %% 64 == 0");
            break;
        case JAT_EPI_GATE_RESIDUAL:
            if (!e->gate || e->ldo % 4 != 0 || e->gate_batch_stride % 4 != 0)
                return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: GATE_RESIDUAL needs gate and 16B-aligned pitches");
            break;
        case JAT_EPI_ACCUM:
            if (e->ldo % 4 != 0) return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: ACCUM needs a 16B-aligned output pitch");
            break;
        case JAT_EPI_DACT:
            if (!e->aux || e->ld_aux % 8 != 0 || e->ldo % 8 != 0 || (reinterpret_cast<uintptr_t>(e->aux) & 15) != 0)
                return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: DACT needs a 16B-aligned aux (pre-activation) operand");
            break;
        case JAT_EPI_UNPATCHIFY:
            if (e->patch_len != 4) return fail(JAT_ERR_BAD_SHAPE, "jat_gemm_bf16: UNPATCHIFY supports patch_len 4 only");
            if (e->t_out <= 0 || e->t_out > p.tokens_per_batch * 4)
                return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: UNPATCHIFY t_out out of range");
            break;
        default:
            return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: unknown epilogue kind %d", e->kind);
    }

    if (e->kind == JAT_EPI_BIAS_ACT && e->aux != nullptr &&
        (e->out_dtype != JAT_DTYPE_BF16 || e->ld_aux % 8 != 0 || (reinterpret_cast<uintptr_t>(e->aux) & 15) != 0))
        return fail(JAT_ERR_BAD_ARG, "jat_gemm_bf16: the pre-activation copy (aux) needs bf16 output and 16B alignment");
    CUtensorMap ta, tb, to;
    // transposed operands: rows = reduction index, boxes of 64 reduction rows x 64 M/N elements
    if (a_mn) JAT_TRY(make_tmap(ctx, &ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64));
    else JAT_TRY(make_tmap(ctx, &ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM));
    if (w_mn) JAT_TRY(make_tmap(ctx, &tb, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, 64));
    else JAT_TRY(make_tmap(ctx, &tb, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)(block_n / cg / mc)));
    if (e->kind == JAT_EPI_UNPATCHIFY) {
        to = ta;  // unused by that epilogue
    } else {
        // epilogue slabs: 32 rows x 128 bytes, stored (or reduce-added) by the TMA unit; rows >= M are clipped
        const bool out_f32 = e->kind == JAT_EPI_GATE_RESIDUAL || e->kind == JAT_EPI_ACCUM ||
                             (e->kind == JAT_EPI_BIAS_ACT && e->out_dtype == JAT_DTYPE_F32);
        JAT_TRY(make_tmap(ctx, &to, e->out, (uint64_t)M, (uint64_t)N, (uint64_t)e->ldo, 32, out_f32));
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (mc == 2) return dispatch_gemm_mc2(ctx, ta, tb, to, p, e, s);
    if (block_n == 256) {
        if (cg == 1) return dispatch_gemm_epi<256, 1>(ctx, ta, tb, to, p, e, s);
        return dispatch_gemm_epi<256, 2>(ctx, ta, tb, to, p, e, s);
    }
    if (cg == 1) return dispatch_gemm_epi<128, 1>(ctx, ta, tb, to, p, e, s);
    return dispatch_gemm_epi<128, 2>(ctx, ta, tb, to, p, e, s);
}

// ------------------------------------------------------------------------------------------------ AdaLN
template <int NORM>
static int launch_adaln(jat_ctx* ctx, const float* x, __nv_bfloat16* out, const float* shift, const float* scale,
                        long long bstride, const float* weight, float eps, int M, int D, int ntok, cudaStream_t s,
                        float2* rowstats = nullptr, float* x_copy = nullptr) {
    const int nvec = D / 4;
    const int nv = (nvec + 31) / 32;
    dim3 grid((M + ADALN_WARPS - 1) / ADALN_WARPS), block(ADALN_WARPS * 32);
    pre_launch(ctx, TAG_ADALN, s);
#define JAT_ADALN_CASE(NV) \
    launch_pdl(adaln_norm_modulate_kernel<NV, NORM>, grid, block, 0, s, x, out, shift, scale, bstride, weight, eps, M, D, ntok, \
               rowstats, x_copy)
    if (nv <= 4) JAT_ADALN_CASE(4);
    else if (nv <= 8) JAT_ADALN_CASE(8);
    else if (nv <= 10) JAT_ADALN_CASE(10);
    else if (nv <= 16) JAT_ADALN_CASE(16);
    else JAT_ADALN_CASE(32);
#undef JAT_ADALN_CASE
    return post_launch(ctx, "adaln_norm_modulate");
}

static int adaln_norm_modulate_stats(jat_ctx* ctx, const float* x, void* out_bf16, const float* shift, const float* scale,
                                     int64_t mod_batch_stride, const float* weight, int norm_kind, float eps, int M, int D,
                                     int tokens_per_batch, float* rowstats, float* x_copy, void* stream);

extern "C" int jat_adaln_norm_modulate(jat_ctx* ctx, const float* x, void* out_bf16, const float* shift,
                                       const float* scale, int64_t mod_batch_stride, const float* weight, int norm_kind,
                                       float eps, int M, int D, int tokens_per_batch, void* stream) {
    DeviceGuard dev_guard__(ctx);
    return adaln_norm_modulate_stats(ctx, x, out_bf16, shift, scale, mod_batch_stride, weight, norm_kind, eps, M, D,
                                     tokens_per_batch, nullptr, nullptr, stream);
}

// training forward: rowstats = optional f32 [M, 2] (row mean, rstd), x_copy = optional f32 [M, D] copy of x, both for the backward
static int adaln_norm_modulate_stats(jat_ctx* ctx, const float* x, void* out_bf16, const float* shift, const float* scale,
                                     int64_t mod_batch_stride, const float* weight, int norm_kind, float eps, int M, int D,
                                     int tokens_per_batch, float* rowstats, float* x_copy, void* stream) {
    if (!ctx || !x || !out_bf16) return fail(JAT_ERR_BAD_ARG, "jat_adaln_norm_modulate: null argument");
    if ((shift == nullptr) != (scale == nullptr))
        return fail(JAT_ERR_BAD_ARG, "jat_adaln_norm_modulate: shift and scale must both be given or both NULL");
    if (M <= 0 || D <= 0 || D % 4 != 0 || D > 4096)
        return fail(JAT_ERR_BAD_SHAPE, "jat_adaln_norm_modulate: need D %% 4 == 0 and D <= 4096 (got %d)", D);
    if (mod_batch_stride % 4 != 0) return fail(JAT_ERR_BAD_ARG, "jat_adaln_norm_modulate: mod_batch_stride %% 4 != 0");
    if (tokens_per_batch <= 0) tokens_per_batch = M;
    cudaStream_t s = (cudaStream_t)stream;
    if (norm_kind == JAT_NORM_LAYERNORM)
        return launch_adaln<0>(ctx, x, (__nv_bfloat16*)out_bf16, shift, scale, mod_batch_stride, nullptr, eps, M, D,
                               tokens_per_batch, s, (float2*)rowstats, x_copy);
    if (norm_kind == JAT_NORM_RMSNORM) {
        if (!weight) return fail(JAT_ERR_BAD_ARG, "jat_adaln_norm_modulate: RMSNorm needs a weight");
        return launch_adaln<1>(ctx, x, (__nv_bfloat16*)out_bf16, shift, scale, mod_batch_stride, weight, eps, M, D,
                               tokens_per_batch, s, (float2*)rowstats, x_copy);
    }
    return fail(JAT_ERR_BAD_ARG, "jat_adaln_norm_modulate: unknown norm_kind %d", norm_kind);
}

// ------------------------------------------------------------------------------------------------ misc
extern "C" int jat_patchify_cast2(jat_ctx* ctx, const float* x_t, int xt_batch, const float* x_cond, int cond_batch,
                                  void* out_bf16, int B, int C, int Cc, int T, int P, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !x_t || !out_bf16) return fail(JAT_ERR_BAD_ARG, "jat_patchify_cast: null argument");
    if (P != 4) return fail(JAT_ERR_BAD_SHAPE, "jat_patchify_cast: patch_len must be 4 (got %d)", P);
    if (B <= 0 || C <= 0 || Cc <= 0 || T <= 0 || C % PATCH_TC != 0 || Cc % PATCH_TC != 0 || xt_batch <= 0 || B > 65535)
        return fail(JAT_ERR_BAD_SHAPE, "jat_patchify_cast: need C %% 32 == 0, Cc %% 32 == 0, 0 < B <= 65535");
    const int N = (T + P - 1) / P;
    dim3 grid((N + PATCH_TN - 1) / PATCH_TN, (C + Cc) / PATCH_TC, B);
    pre_launch(ctx, TAG_PATCHIFY, (cudaStream_t)stream);
    launch_pdl(patchify_cast_kernel, grid, dim3(256), 0, (cudaStream_t)stream, x_t, xt_batch, x_cond, cond_batch,
               (__nv_bfloat16*)out_bf16, C, Cc, T, N, (C + Cc) * 4);
    return post_launch(ctx, "patchify_cast");
}

extern "C" int jat_patchify_cast(jat_ctx* ctx, const float* x_t, int xt_batch, const float* x_cond, int cond_batch,
                                 void* out_bf16, int B, int C, int T, int P, void* stream) {
    return jat_patchify_cast2(ctx, x_t, xt_batch, x_cond, cond_batch, out_bf16, B, C, C, T, P, stream);
}

extern "C" int jat_patchify_single(jat_ctx* ctx, const float* x, void* out_bf16, int B, int C, int T, int P, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !x || !out_bf16) return fail(JAT_ERR_BAD_ARG, "jat_patchify_single: null argument");
    if (P != 4) return fail(JAT_ERR_BAD_SHAPE, "jat_patchify_single: patch_len must be 4 (got %d)", P);
    if (B <= 0 || C <= 0 || T <= 0 || C % PATCH_TC != 0 || B > 65535) return fail(JAT_ERR_BAD_SHAPE, "jat_patchify_single: bad sizes");
    const int N = (T + P - 1) / P;
    dim3 grid((N + PATCH_TN - 1) / PATCH_TN, C / PATCH_TC, B);
    pre_launch(ctx, TAG_PATCHIFY, (cudaStream_t)stream);
    launch_pdl(patchify_cast_kernel, grid, dim3(256), 0, (cudaStream_t)stream, x, B, (const float*)nullptr, 0,
               (__nv_bfloat16*)out_bf16, C, 0, T, N, C * 4);
    return post_launch(ctx, "patchify_single");
}

extern "C" int jat_timestep_features(jat_ctx* ctx, const float* t, void* out_bf16, int B, int D, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !t || !out_bf16) return fail(JAT_ERR_BAD_ARG, "jat_timestep_features: null argument");
    if (B <= 0 || D < 4 || D % 2 != 0 || B > 65535) return fail(JAT_ERR_BAD_SHAPE, "jat_timestep_features: bad B/D");
    dim3 grid((D / 2 + 127) / 128, B);
    pre_launch(ctx, TAG_TSTEP, (cudaStream_t)stream);
    launch_pdl(timestep_features_kernel, grid, dim3(128), 0, (cudaStream_t)stream, t, (__nv_bfloat16*)out_bf16, B, D);
    return post_launch(ctx, "timestep_features");
}

extern "C" int jat_cfg_euler_update(jat_ctx* ctx, float* z, const float* x_c, const float* x_u, float cfg_scale,
                                    const float* t_dt, int step, int64_t numel, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !z || !x_c || !t_dt) return fail(JAT_ERR_BAD_ARG, "jat_cfg_euler_update: null argument");
    if (numel <= 0 || step < 0) return fail(JAT_ERR_BAD_ARG, "jat_cfg_euler_update: bad numel/step");
    long long want = (numel / 4 + 255) / 256;
    long long cap = (long long)ctx->sm_count * 8;
    int blocks = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    pre_launch(ctx, TAG_EULER, (cudaStream_t)stream);
    launch_pdl(cfg_euler_update_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, z, x_c, x_u, cfg_scale, t_dt, step,
               (long long)numel);
    return post_launch(ctx, "cfg_euler_update");
}

// ------------------------------------------------------------------------------------------------ backward (elementwise)
template <int NORM>
static int launch_adaln_bwd(jat_ctx* ctx, const __nv_bfloat16* dh, const float* x, const float* scale, long long bstride,
                            const float* weight, float eps, float* dx, int accumulate, float* dshift, float* dscale,
                            long long dbstride, float* dweight, float2* rowstats, int B, int ntok, int D, cudaStream_t s) {
    const int M = B * ntok;
    const int nv = (D / 4 + 31) / 32;
    dim3 grid((M + BWD_WARPS - 1) / BWD_WARPS), block(BWD_WARPS * 32);
    pre_launch(ctx, TAG_ADALN_BWD, s);
#define JAT_CASE(NV) \
    adaln_bwd_dx_kernel<NV, NORM><<<grid, block, 0, s>>>(dh, x, scale, bstride, weight, eps, dx, accumulate, rowstats, M, D, ntok)
    if (nv <= 4) JAT_CASE(4);
    else if (nv <= 8) JAT_CASE(8);
    else if (nv <= 10) JAT_CASE(10);
    else JAT_CASE(16);
#undef JAT_CASE
    JAT_TRY(post_launch(ctx, "adaln_bwd_dx"));
    if (scale == nullptr && dweight == nullptr) return 0;
    dim3 grid2((ntok + BWD_ROWS_PER_CTA - 1) / BWD_ROWS_PER_CTA, B), block2((D / 4 + 31) / 32 * 32);
    pre_launch(ctx, TAG_ADALN_BWD, s);
    adaln_bwd_colsum_kernel<NORM><<<grid2, block2, 0, s>>>(dh, x, rowstats, scale, bstride, weight, dshift, dscale, dbstride,
                                                           dweight, D, ntok);
    return post_launch(ctx, "adaln_bwd_colsum");
}

extern "C" int jat_adaln_bwd(jat_ctx* ctx, const void* dh_bf16, const float* x, const float* scale, int64_t mod_batch_stride,
                             const float* weight, int norm_kind, float eps, float* dx, int accumulate, float* dshift,
                             float* dscale, int64_t dmod_batch_stride, float* dweight, float* rowstats_scratch, int B,
                             int tokens_per_batch, int D, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !dh_bf16 || !x || !dx) return fail(JAT_ERR_BAD_ARG, "jat_adaln_bwd: null argument");
    if (scale != nullptr && (!dshift || !dscale)) return fail(JAT_ERR_BAD_ARG, "jat_adaln_bwd: dshift/dscale missing");
    if ((scale != nullptr || dweight != nullptr) && !rowstats_scratch)
        return fail(JAT_ERR_BAD_ARG, "jat_adaln_bwd: rowstats_scratch (f32 [M, 2]) missing");
    if (B <= 0 || B > 65535 || tokens_per_batch <= 0 || D <= 0 || D % 4 != 0 || D > 2048)
        return fail(JAT_ERR_BAD_SHAPE, "jat_adaln_bwd: need D %% 4 == 0, D <= 2048, 0 < B <= 65535");
    if (mod_batch_stride % 4 != 0 || dmod_batch_stride % 4 != 0) return fail(JAT_ERR_BAD_ARG, "jat_adaln_bwd: strides %% 4 != 0");
    cudaStream_t s = (cudaStream_t)stream;
    if (norm_kind == JAT_NORM_LAYERNORM)
        return launch_adaln_bwd<0>(ctx, (const __nv_bfloat16*)dh_bf16, x, scale, mod_batch_stride, nullptr, eps, dx, accumulate,
                                   dshift, dscale, dmod_batch_stride, nullptr, (float2*)rowstats_scratch, B, tokens_per_batch, D, s);
    if (norm_kind == JAT_NORM_RMSNORM) {
        if (!weight) return fail(JAT_ERR_BAD_ARG, "jat_adaln_bwd: RMSNorm needs a weight");
        return launch_adaln_bwd<1>(ctx, (const __nv_bfloat16*)dh_bf16, x, scale, mod_batch_stride, weight, eps, dx, accumulate,
                                   dshift, dscale, dmod_batch_stride, dweight, (float2*)rowstats_scratch, B, tokens_per_batch, D, s);
    }
    return fail(JAT_ERR_BAD_ARG, "jat_adaln_bwd: unknown norm_kind %d", norm_kind);
}

extern "C" int jat_gate_bwd(jat_ctx* ctx, const float* dx, const void* y_bf16, const float* gate, int64_t mod_batch_stride,
                            void* dy_bf16, float* dgate, int64_t dmod_batch_stride, float* dxsum_scratch, float* dbias, int B,
                            int tokens_per_batch, int D, void* stream) {
    DeviceGuard dev_guard__(ctx);
    return jat_gate_bwd_dropout(ctx, dx, y_bf16, gate, mod_batch_stride, dy_bf16, dgate, dmod_batch_stride, dxsum_scratch,
                                dbias, B, tokens_per_batch, D, 0.0f, 0u, nullptr, stream);
}

extern "C" int jat_gate_bwd_dropout(jat_ctx* ctx, const float* dx, const void* y_bf16, const float* gate,
                                    int64_t mod_batch_stride, void* dy_bf16, float* dgate, int64_t dmod_batch_stride,
                                    float* dxsum_scratch, float* dbias, int B, int tokens_per_batch, int D, float drop_p,
                                    uint32_t drop_seed, const float* gate_rowscale, void* stream) {
    DeviceGuard dev_guard__(ctx);
    DropCfg drop;
    if (!make_drop(drop_p, drop_seed, &drop)) return fail(JAT_ERR_BAD_ARG, "jat_gate_bwd: drop_p must be in [0, 1)");
    if (!ctx || !dx || !y_bf16 || !gate || !dy_bf16 || !dgate) return fail(JAT_ERR_BAD_ARG, "jat_gate_bwd: null argument");
    if (B <= 0 || B > 65535 || tokens_per_batch <= 0 || D <= 0 || D % 4 != 0 || D > 2048)
        return fail(JAT_ERR_BAD_SHAPE, "jat_gate_bwd: need D %% 4 == 0, D <= 2048, 0 < B <= 65535");
    cudaStream_t s = (cudaStream_t)stream;
    (void)dxsum_scratch;  // no longer used: the bias gradient is accumulated by the kernel itself
    dim3 grid((tokens_per_batch + BWD_ROWS_PER_CTA - 1) / BWD_ROWS_PER_CTA, B), block((D / 4 + 31) / 32 * 32);
    pre_launch(ctx, TAG_GATE_BWD, s);
    gate_bwd_kernel<<<grid, block, 0, s>>>(dx, (const __nv_bfloat16*)y_bf16, gate, mod_batch_stride, (__nv_bfloat16*)dy_bf16,
                                           dgate, dmod_batch_stride, dbias, D, tokens_per_batch, drop, gate_rowscale);
    return post_launch(ctx, "gate_bwd");
}

extern "C" int jat_adaln_gate_bwd(jat_ctx* ctx, const void* dh_bf16, const float* x, const float* rowstats, const float* scale,
                                  int64_t mod_batch_stride, const float* weight, int norm_kind, float* dx, float* dshift,
                                  float* dscale, int64_t dmod_batch_stride, float* dweight, const void* y_bf16,
                                  const float* gate, void* dy_bf16, float* dgate, float* dxsum_scratch, float* dbias, int B,
                                  int tokens_per_batch, int D, float drop_p, uint32_t drop_seed, const float* gate_rowscale,
                                  void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !dh_bf16 || !x || !rowstats || !scale || !dx || !dshift || !dscale)
        return fail(JAT_ERR_BAD_ARG, "jat_adaln_gate_bwd: null argument");
    const bool has_gate = y_bf16 != nullptr;
    if (has_gate && (!gate || !dy_bf16 || !dgate)) return fail(JAT_ERR_BAD_ARG, "jat_adaln_gate_bwd: gate / dy / dgate missing");
    if (dbias != nullptr && !has_gate) return fail(JAT_ERR_BAD_ARG, "jat_adaln_gate_bwd: dbias needs the gate part");
    (void)dxsum_scratch;
    if (B <= 0 || B > 65535 || tokens_per_batch <= 0 || D <= 0 || D % 4 != 0 || D > 2048)
        return fail(JAT_ERR_BAD_SHAPE, "jat_adaln_gate_bwd: need D %% 4 == 0, D <= 2048, 0 < B <= 65535");
    if (mod_batch_stride % 4 != 0 || dmod_batch_stride % 4 != 0)
        return fail(JAT_ERR_BAD_ARG, "jat_adaln_gate_bwd: strides %% 4 != 0");
    if (norm_kind != JAT_NORM_LAYERNORM && norm_kind != JAT_NORM_RMSNORM)
        return fail(JAT_ERR_BAD_ARG, "jat_adaln_gate_bwd: unknown norm_kind %d", norm_kind);
    if (norm_kind == JAT_NORM_RMSNORM && !weight) return fail(JAT_ERR_BAD_ARG, "jat_adaln_gate_bwd: RMSNorm needs a weight");
    DropCfg drop;
    if (!make_drop(drop_p, drop_seed, &drop)) return fail(JAT_ERR_BAD_ARG, "jat_adaln_gate_bwd: drop_p must be in [0, 1)");
    cudaStream_t s = (cudaStream_t)stream;
    float* xs = dbias;
    // one wave: two CTAs of D/4 threads per SM, every batch item cut into the same number of row ranges
    int per_batch = (2 * ctx->sm_count) / B;
    if (per_batch < 1) per_batch = 1;
    int rows = (tokens_per_batch + per_batch - 1) / per_batch;
    rows = (rows + AGB_ROWS - 1) / AGB_ROWS * AGB_ROWS;
    dim3 grid((tokens_per_batch + rows - 1) / rows, B), block((D / 4 + 31) / 32 * 32);
    const float2* rsp = (const float2*)rowstats;
    const __nv_bfloat16 *dh = (const __nv_bfloat16*)dh_bf16, *y = (const __nv_bfloat16*)y_bf16;
    __nv_bfloat16* dy = (__nv_bfloat16*)dy_bf16;
    pre_launch(ctx, TAG_ADALN_BWD, s);
#define JAT_AGB(NORM, GATE)                                                                                              \
    do {                                                                                                                 \
        if (block.x <= 320)                                                                                              \
            adaln_gate_bwd_kernel<NORM, GATE, 320><<<grid, block, 0, s>>>(dh, x, rsp, scale, mod_batch_stride, weight, dx, dshift, \
                dscale, dmod_batch_stride, dweight, y, gate, dy, dgate, xs, drop, gate_rowscale, D, tokens_per_batch, rows); \
        else                                                                                                             \
            adaln_gate_bwd_kernel<NORM, GATE, 512><<<grid, block, 0, s>>>(dh, x, rsp, scale, mod_batch_stride, weight, dx, dshift, \
                dscale, dmod_batch_stride, dweight, y, gate, dy, dgate, xs, drop, gate_rowscale, D, tokens_per_batch, rows); \
    } while (0)
    // staged form (inputs through shared memory by the bulk-copy engine) whenever the rows are 16-byte granular
    static const int staged_env = [] { const char* v = getenv("JAT_AGB_STAGED"); return v ? atoi(v) : 1; }();
    const size_t stage_smem = (size_t)AGS_STAGES * AGS_ROWS * D * (has_gate ? 12 : 10) + AGS_PAD;
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(dh) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx) |
                             reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    const bool staged = staged_env != 0 && D % 8 == 0 && aligned16 && stage_smem <= (block.x <= 320 ? 110u : 220u) * 1024u;
#define JAT_AGS(NORM, GATE, T)                                                                                           \
    do {                                                                                                                 \
        auto kern = adaln_gate_bwd_staged_kernel<NORM, GATE, T>;                                                         \
        JAT_TRY(ensure_dyn_smem(ctx, (const void*)kern, (T <= 320 ? 110 : 220) * 1024));                                 \
        int rows2 = (tokens_per_batch + per_batch - 1) / per_batch;                                                      \
        rows2 = (rows2 + AGS_ROWS - 1) / AGS_ROWS * AGS_ROWS;                                                            \
        dim3 grid2((tokens_per_batch + rows2 - 1) / rows2, B);                                                           \
        kern<<<grid2, block, stage_smem, s>>>(dh, x, rsp, scale, mod_batch_stride, weight, dx, dshift, dscale,           \
            dmod_batch_stride, dweight, y, gate, dy, dgate, xs, drop, gate_rowscale, D, tokens_per_batch, rows2);        \
    } while (0)
#define JAT_AGX(NORM, GATE)                                                                                              \
    do {                                                                                                                 \
        if (!staged) JAT_AGB(NORM, GATE);                                                                                \
        else if (block.x <= 320) JAT_AGS(NORM, GATE, 320);                                                               \
        else JAT_AGS(NORM, GATE, 512);                                                                                   \
    } while (0)
    if (norm_kind == JAT_NORM_LAYERNORM) { if (has_gate) JAT_AGX(0, 1); else JAT_AGX(0, 0); }
    else { if (has_gate) JAT_AGX(1, 1); else JAT_AGX(1, 0); }
#undef JAT_AGX
#undef JAT_AGS
#undef JAT_AGB
    return post_launch(ctx, "adaln_gate_bwd");
}

extern "C" int jat_colsum_bf16(jat_ctx* ctx, const void* a_bf16, int64_t lda, int M, int cols, float* out, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !a_bf16 || !out) return fail(JAT_ERR_BAD_ARG, "jat_colsum_bf16: null argument");
    if (M <= 0 || cols <= 0 || cols % 4 != 0 || lda % 4 != 0 || (reinterpret_cast<uintptr_t>(a_bf16) & 7) != 0)
        return fail(JAT_ERR_BAD_SHAPE, "jat_colsum_bf16: need cols %% 4 == 0, lda %% 4 == 0, 8-byte aligned input");
    int chunks = (M + 31) / 32;
    // one wave: 16 CTAs of 128 threads per SM; a grid two CTAs over it (the former "+ 1") ran a second wave for them
    int cap = (ctx->sm_count * 16) / ((cols + 511) / 512);
    if (cap < 1) cap = 1;
    if (chunks > cap) chunks = cap;
    dim3 grid((cols + 511) / 512, chunks);
    pre_launch(ctx, TAG_COLSUM, (cudaStream_t)stream);
    colsum_bf16_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a_bf16, (long long)lda, M, cols, out);
    return post_launch(ctx, "colsum_bf16");
}

extern "C" int jat_cast_f32_bf16(jat_ctx* ctx, const float* in, void* out_bf16, int64_t n, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !in || !out_bf16 || n <= 0) return fail(JAT_ERR_BAD_ARG, "jat_cast_f32_bf16: bad argument");
    if ((reinterpret_cast<uintptr_t>(in) & 15) != 0 || (reinterpret_cast<uintptr_t>(out_bf16) & 7) != 0)
        return fail(JAT_ERR_BAD_ARG, "jat_cast_f32_bf16: misaligned");
    const long long blocks = (n / 4 + 255) / 256 + 1;
    pre_launch(ctx, TAG_COLSUM, (cudaStream_t)stream);
    cast_f32_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out_bf16, (long long)n);
    return post_launch(ctx, "cast_f32_bf16");
}

// Gradient buckets <-> bf16 all-reduce payload (jat_b200.ddp.bf16_allreduce_hook)
extern "C" int jat_grad_compress(jat_ctx* ctx, const float* in, void* out_bf16, int64_t n, float scale, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !in || !out_bf16 || n <= 0) return fail(JAT_ERR_BAD_ARG, "jat_grad_compress: bad argument");
    if (((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out_bf16)) & 15) != 0)
        return fail(JAT_ERR_BAD_ARG, "jat_grad_compress: buffers must be 16-byte aligned");
    long long blocks = (n / 8 + 255) / 256;
    const long long cap = (long long)ctx->sm_count * 8;
    blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
    pre_launch(ctx, TAG_COLSUM, (cudaStream_t)stream);
    grad_compress_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out_bf16, (long long)n, scale);
    return post_launch(ctx, "grad_compress");
}
extern "C" int jat_grad_decompress(jat_ctx* ctx, const void* in_bf16, float* out, int64_t n, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !in_bf16 || !out || n <= 0) return fail(JAT_ERR_BAD_ARG, "jat_grad_decompress: bad argument");
    if (((reinterpret_cast<uintptr_t>(in_bf16) | reinterpret_cast<uintptr_t>(out)) & 15) != 0)
        return fail(JAT_ERR_BAD_ARG, "jat_grad_decompress: buffers must be 16-byte aligned");
    long long blocks = (n / 8 + 255) / 256;
    const long long cap = (long long)ctx->sm_count * 8;
    blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
    pre_launch(ctx, TAG_COLSUM, (cudaStream_t)stream);
    grad_decompress_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in_bf16, out, (long long)n);
    return post_launch(ctx, "grad_decompress");
}

// ------------------------------------------------------------------------------------------------ training-step glue
extern "C" int jat_train_inputs(jat_ctx* ctx, const float* hr, const float* lr, const float* hr_mean, const float* hr_std,
                                const float* lr_mean, const float* lr_std, const float* noise, const float* cond_noise,
                                const float* cond_scale_dev, float cond_scale, const float* keep, const float* t, float* hr_norm,
                                float* lr_cond, float* z_t, int B, int C, int T, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !hr || !lr || !hr_mean || !hr_std || !lr_mean || !lr_std || !noise || !t || !hr_norm || !lr_cond || !z_t)
        return fail(JAT_ERR_BAD_ARG, "jat_train_inputs: null argument");
    if (B <= 0 || C <= 0 || T <= 0 || B > 65535 || C > 65535) return fail(JAT_ERR_BAD_SHAPE, "jat_train_inputs: bad B/C/T");
    const uintptr_t al = (uintptr_t)hr | (uintptr_t)lr | (uintptr_t)noise | (uintptr_t)cond_noise | (uintptr_t)hr_norm |
                         (uintptr_t)lr_cond | (uintptr_t)z_t;
    if ((T & 3) == 0 && (al & 15) != 0) return fail(JAT_ERR_BAD_ARG, "jat_train_inputs: tensors must be 16-byte aligned");
    dim3 grid((T / 4 + 255) / 256 > 0 ? (T / 4 + 255) / 256 : 1, C, B);
    pre_launch(ctx, TAG_TRAIN_GLUE, (cudaStream_t)stream);
    train_inputs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(hr, lr, hr_mean, hr_std, lr_mean, lr_std, noise, cond_noise,
                                                                cond_scale_dev, cond_scale, keep, t, hr_norm, lr_cond, z_t, C, T);
    return post_launch(ctx, "train_inputs");
}

static int recon_loss(jat_ctx* ctx, int kind, const float* pred, const float* target, float* d_pred, double* stats, int64_t n,
                      float eps, void* stream) {
    if (!ctx || !pred || !target || !stats || n <= 0) return fail(JAT_ERR_BAD_ARG, "jat_mse_loss / jat_charbonnier_loss: bad argument");
    if ((((uintptr_t)pred | (uintptr_t)target | (uintptr_t)d_pred) & 15) != 0)
        return fail(JAT_ERR_BAD_ARG, "jat_mse_loss / jat_charbonnier_loss: tensors must be 16-byte aligned");
    if (kind == 1 && !(eps > 0.0f)) return fail(JAT_ERR_BAD_ARG, "jat_charbonnier_loss: eps must be positive");
    cudaStream_t s = (cudaStream_t)stream;
    JAT_CUDA(cudaMemsetAsync(stats, 0, (kind == 0 ? 4 : 5) * sizeof(double), s));
    long long want = (n / 4 + 255) / 256;
    const long long cap = (long long)ctx->sm_count * 8;
    const int blocks = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    pre_launch(ctx, TAG_TRAIN_GLUE, s);
    if (kind == 0) recon_loss_kernel<0><<<blocks, 256, 0, s>>>(pred, target, d_pred, stats, (long long)n, (float)(2.0 / (double)n), 0.0f);
    else recon_loss_kernel<1><<<blocks, 256, 0, s>>>(pred, target, d_pred, stats, (long long)n, (float)(1.0 / (double)n), eps);
    return post_launch(ctx, kind == 0 ? "mse_loss" : "charbonnier_loss");
}

extern "C" int jat_mse_loss(jat_ctx* ctx, const float* pred, const float* target, float* d_pred, double* stats4, int64_t n,
                            void* stream) {
    DeviceGuard dev_guard__(ctx);
    return recon_loss(ctx, 0, pred, target, d_pred, stats4, n, 0.0f, stream);
}

extern "C" int jat_charbonnier_loss(jat_ctx* ctx, const float* pred, const float* target, float* d_pred, double* stats5, int64_t n,
                                    float eps, void* stream) {
    DeviceGuard dev_guard__(ctx);
    return recon_loss(ctx, 1, pred, target, d_pred, stats5, n, eps, stream);
}

// ------------------------------------------------------------------------------------------------ optimizer step
static_assert(sizeof(jat_adamw_tensor) == sizeof(OptTensor), "jat_adamw_tensor layout");

extern "C" int jat_adamw_chunk_elems(void) { return OPT_CHUNK; }

static int opt_check(const char* fn, jat_ctx* ctx, const void* table, const void* chunk_first, int n_tensors, int total_chunks) {
    if (!ctx || !table || !chunk_first) return fail(JAT_ERR_BAD_ARG, "%s: null argument", fn);
    if (n_tensors <= 0 || total_chunks < n_tensors) return fail(JAT_ERR_BAD_ARG, "%s: bad tensor / chunk count", fn);
    return 0;
}

extern "C" int jat_grad_sumsq(jat_ctx* ctx, const jat_adamw_tensor* table_dev, const int32_t* chunk_first_dev, int n_tensors,
                              int total_chunks, float* partials_dev, double* sumsq_dev, int accumulate, void* stream) {
    DeviceGuard dev_guard__(ctx);
    JAT_TRY(opt_check("jat_grad_sumsq", ctx, table_dev, chunk_first_dev, n_tensors, total_chunks));
    if (!partials_dev || !sumsq_dev) return fail(JAT_ERR_BAD_ARG, "jat_grad_sumsq: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    pre_launch(ctx, TAG_OPTIMIZER, s);
    grad_sumsq_kernel<<<(unsigned)total_chunks, OPT_THREADS, 0, s>>>((const OptTensor*)table_dev, chunk_first_dev, n_tensors,
                                                                      partials_dev);
    JAT_TRY(post_launch(ctx, "grad_sumsq"));
    pre_launch(ctx, TAG_OPTIMIZER, s);
    grad_sumsq_final_kernel<<<1, 1024, 0, s>>>(partials_dev, total_chunks, sumsq_dev, accumulate);
    return post_launch(ctx, "grad_sumsq_final");
}

extern "C" int jat_adamw_step(jat_ctx* ctx, const jat_adamw_tensor* table_dev, const int32_t* chunk_first_dev, int n_tensors,
                              int total_chunks, double lr, double beta1, double beta2, double eps, double weight_decay,
                              float max_norm, const double* sumsq_dev, void* stream) {
    DeviceGuard dev_guard__(ctx);
    JAT_TRY(opt_check("jat_adamw_step", ctx, table_dev, chunk_first_dev, n_tensors, total_chunks));
    if (!(lr >= 0.0) || !(eps >= 0.0) || !(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(weight_decay >= 0.0))
        return fail(JAT_ERR_BAD_ARG, "jat_adamw_step: hyper-parameter out of range");
    if (max_norm > 0.f && !sumsq_dev) return fail(JAT_ERR_BAD_ARG, "jat_adamw_step: clipping needs the gradient sum of squares");
    AdamWArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
    a.max_norm = max_norm;
    a.sumsq = sumsq_dev;
    a.lr_wd = (float)(lr * weight_decay);
    a.b1 = (float)beta1; a.one_minus_b1 = (float)(1.0 - beta1);
    a.b2 = (float)beta2; a.one_minus_b2 = (float)(1.0 - beta2);
    a.eps_f = (float)eps;
    cudaStream_t s = (cudaStream_t)stream;
    pre_launch(ctx, TAG_OPTIMIZER, s);
    static const bool f64_math = getenv("JAT_ADAMW_F64") && atoi(getenv("JAT_ADAMW_F64")) != 0;  // ATen's operand types
    if (f64_math) adamw_kernel<1><<<(unsigned)total_chunks, OPT_THREADS, 0, s>>>((const OptTensor*)table_dev, chunk_first_dev, n_tensors, a);
    else adamw_kernel<0><<<(unsigned)total_chunks, OPT_THREADS, 0, s>>>((const OptTensor*)table_dev, chunk_first_dev, n_tensors, a);
    return post_launch(ctx, "adamw");
}

// ------------------------------------------------------------------------------------------------ chunks
extern "C" int jat_chunk_normalize(jat_ctx* ctx, const float* latent, int64_t total_frames, int64_t ld, const float* mean,
                                   const float* std, float* out, int n_chunks, int first_chunk, int chunk_step, int C,
                                   int chunk_frames, int stride, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !latent || !out) return fail(JAT_ERR_BAD_ARG, "jat_chunk_normalize: null argument");
    if ((mean == nullptr) != (std == nullptr)) return fail(JAT_ERR_BAD_ARG, "jat_chunk_normalize: mean and std go together");
    if (n_chunks <= 0 || C <= 0 || chunk_frames <= 0 || stride <= 0 || first_chunk < 0 || chunk_step <= 0 ||
        total_frames <= 0 || ld < total_frames || n_chunks > 65535 || C > 65535)
        return fail(JAT_ERR_BAD_SHAPE, "jat_chunk_normalize: bad sizes");
    dim3 grid((chunk_frames + 255) / 256, C, n_chunks);
    pre_launch(ctx, TAG_CHUNKN, (cudaStream_t)stream);
    chunk_normalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(latent, (long long)total_frames, (long long)ld, mean, std,
                                                                   out, C, chunk_frames, stride, first_chunk, chunk_step);
    return post_launch(ctx, "chunk_normalize");
}

extern "C" int jat_crossfade_denorm(jat_ctx* ctx, const float* chunks, int n_chunks, int C, int chunk_frames, int overlap,
                                    const float* fade_in, const float* fade_out, const float* mean, const float* std,
                                    float* out, int64_t total_frames, int64_t ldo, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !chunks || !out) return fail(JAT_ERR_BAD_ARG, "jat_crossfade_denorm: null argument");
    if ((mean == nullptr) != (std == nullptr)) return fail(JAT_ERR_BAD_ARG, "jat_crossfade_denorm: mean and std go together");
    if (n_chunks <= 0 || C <= 0 || C > 65535 || chunk_frames <= 0 || overlap < 0 || total_frames <= 0 || ldo < total_frames)
        return fail(JAT_ERR_BAD_SHAPE, "jat_crossfade_denorm: bad sizes");
    if (overlap > 0 && (!fade_in || !fade_out)) return fail(JAT_ERR_BAD_ARG, "jat_crossfade_denorm: fade tables missing");
    if (n_chunks > 1 && chunk_frames < 2 * overlap)
        return fail(JAT_ERR_BAD_SHAPE, "jat_crossfade_denorm: chunk_frames (%d) < 2 * overlap (%d)", chunk_frames, overlap);
    const int64_t stride = chunk_frames - overlap;
    if (total_frames > (int64_t)(n_chunks - 1) * stride + chunk_frames || total_frames <= (int64_t)(n_chunks - 1) * stride + overlap)
        if (!(n_chunks == 1 && total_frames <= chunk_frames))
            return fail(JAT_ERR_BAD_SHAPE, "jat_crossfade_denorm: total_frames %lld does not match %d chunks",
                        (long long)total_frames, n_chunks);
    dim3 grid((unsigned)((total_frames + 255) / 256), C);
    pre_launch(ctx, TAG_XFADE, (cudaStream_t)stream);
    crossfade_denorm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(chunks, n_chunks, C, chunk_frames, overlap, fade_in,
                                                                    fade_out, mean, std, out, (long long)total_frames,
                                                                    (long long)ldo);
    return post_launch(ctx, "crossfade_denorm");
}

// ------------------------------------------------------------------------------------------------ attention
template <int NKH, bool DROP = false>
static int launch_attention(jat_ctx* ctx, const CUtensorMap& tq, const void* qkv, uint64_t rows, uint64_t cols,
                            const AttnParams& p, dim3 grid, cudaStream_t s) {
    using Cfg = AttCfg<NKH>;
    CUtensorMap tkv;
    JAT_TRY(make_tmap(ctx, &tkv, qkv, rows, cols, cols, (uint32_t)NKH));
    JAT_TRY(ensure_dyn_smem(ctx, (const void*)gqa_attention_fwd_kernel<NKH, DROP>, Cfg::SMEM_BYTES));
    pre_launch(ctx, TAG_ATTN, s);
    launch_pdl(gqa_attention_fwd_kernel<NKH, DROP>, grid, dim3(ATT_THREADS), Cfg::SMEM_BYTES, s, tq, tkv, p);
    return post_launch(ctx, "gqa_attention_fwd");
}

extern "C" int jat_gqa_attention_fwd(jat_ctx* ctx, const void* qkv, void* out, float* lse, int B, int N, int Hq, int Hkv,
                                     int head_dim, void* stream) {
    DeviceGuard dev_guard__(ctx);
    return jat_gqa_attention_fwd_dropout(ctx, qkv, out, lse, B, N, Hq, Hkv, head_dim, 0.0f, 0u, stream);
}

// Query heads per CTA.  A CTA serves Gs heads of one (query tile, KV group) after staging the group's K / V once; the
// launch is `groups` x ceil(G / Gs) CTAs, one per SM at a time (512 TMEM columns), handed out full parts first.  Whole
// groups (Gs = G) stage K / V least often, but 336 CTAs of 5 heads on 148 SMs (training, B = 28) take 3 rounds for 2.27
// rounds of work.  Pick the Gs with the shortest list-schedule makespan, a CTA costing (heads + 0.6) head-times -- 0.6 for
// barrier / TMEM set-up and the K / V + first Q load before the first MMA, fitted to the training step's measured class
// times (Gs = 5: 3.19 ms, 4: 2.87, 3: 2.92, 2: 3.05; JAT_ATTN_GS forces a value for such sweeps).
static int attention_heads_per_cta(jat_ctx* ctx, long long groups, int G) {
    static const int forced = [] { const char* v = getenv("JAT_ATTN_GS"); return v ? atoi(v) : 0; }();   // experiments
    if (forced > 0) return forced < G ? forced : G;
    const long long key = groups * 64 + G;
    auto it = ctx->attn_gs.find(key);
    if (it != ctx->attn_gs.end()) return it->second;
    const int sms = ctx->sm_count;
    const double setup = 0.6;
    int best = G;
    double best_t = -1.0;
    for (int gs = G; gs >= 1; --gs) {
        const int parts = (G + gs - 1) / gs, last = G - (parts - 1) * gs;
        // list scheduling: CTAs in launch order (full parts first) to the SM that frees up first
        std::priority_queue<double, std::vector<double>, std::greater<double>> pq;
        for (int i = 0; i < sms; ++i) pq.push(0.0);
        double makespan = 0.0;
        const long long n_full = groups * (parts - 1), n_last = groups;
        if (n_full + n_last > 200000) { continue; }
        for (long long c = 0; c < n_full + n_last; ++c) {
            const double w = (c < n_full ? gs : last) + setup;
            const double t = pq.top() + w;
            pq.pop();
            pq.push(t);
            if (t > makespan) makespan = t;
        }
        if (best_t < 0.0 || makespan < best_t - 1e-9) { best_t = makespan; best = gs; }
    }
    ctx->attn_gs[key] = best;
    return best;
}

// one launch of the single-pass kernel on keys [key0, key0 + nkeys) of every batch item
static int attention_pass(jat_ctx* ctx, const void* qkv, void* out, float* lse, int B, int N, int Hq, int Hkv, int key0,
                          int nkeys, const DropCfg& drop, cudaStream_t s) {
    AttnParams p = {};
    p.B = B; p.N = N; p.Hq = Hq; p.Hkv = Hkv; p.G = Hq / Hkv;
    p.key0 = key0; p.NKeys = nkeys;
    p.out = (__nv_bfloat16*)out;
    p.lse = lse;
    p.scale_log2e = 0.125f * 1.4426950408889634f;
    p.trace = ctx->att_trace;
    p.drop = drop;
    const uint64_t rows = (uint64_t)B * N, cols = (uint64_t)(Hq + 2 * Hkv) * ATT_HD;
    CUtensorMap tq;
    JAT_TRY(make_tmap(ctx, &tq, qkv, rows, cols, cols, ATT_BQ));
    const int qtiles = (N + ATT_BQ - 1) / ATT_BQ;
    p.Gs = attention_heads_per_cta(ctx, (long long)qtiles * Hkv * B, p.G);
    if ((long long)B * ((p.G + p.Gs - 1) / p.Gs) > 65535) p.Gs = p.G;   // grid.z limit: whole groups
    const int parts = (p.G + p.Gs - 1) / p.Gs;
    if ((long long)B * parts > 65535) return fail(JAT_ERR_BAD_SHAPE, "jat_gqa_attention_fwd: batch %d too large", B);
    dim3 grid(qtiles, Hkv, B * parts);
    // key range padded to NK = 2*NKH columns (two softmax warpgroups); padded keys are masked in-kernel
    if (p.drop.thresh != 0u) {
        if (nkeys <= 64) return launch_attention<32, true>(ctx, tq, qkv, rows, cols, p, grid, s);
        if (nkeys <= 128) return launch_attention<64, true>(ctx, tq, qkv, rows, cols, p, grid, s);
        if (nkeys <= 192) return launch_attention<96, true>(ctx, tq, qkv, rows, cols, p, grid, s);
        if (nkeys <= 256) return launch_attention<128, true>(ctx, tq, qkv, rows, cols, p, grid, s);
        return launch_attention<176, true>(ctx, tq, qkv, rows, cols, p, grid, s);
    }
    if (nkeys <= 64) return launch_attention<32>(ctx, tq, qkv, rows, cols, p, grid, s);
    if (nkeys <= 128) return launch_attention<64>(ctx, tq, qkv, rows, cols, p, grid, s);
    if (nkeys <= 192) return launch_attention<96>(ctx, tq, qkv, rows, cols, p, grid, s);
    if (nkeys <= 256) return launch_attention<128>(ctx, tq, qkv, rows, cols, p, grid, s);
    return launch_attention<176>(ctx, tq, qkv, rows, cols, p, grid, s);
}

extern "C" int jat_gqa_attention_fwd_dropout(jat_ctx* ctx, const void* qkv, void* out, float* lse, int B, int N, int Hq,
                                             int Hkv, int head_dim, float drop_p, uint32_t drop_seed, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (N > ATT_MAX_NK)
        return fail(JAT_ERR_BAD_SHAPE, "jat_gqa_attention_fwd: N = %d tokens > %d needs the chunked form jat_gqa_attention_fwd_long "
                    "(scratch for the per-chunk partial results)", N, ATT_MAX_NK);
    return jat_gqa_attention_fwd_long(ctx, qkv, out, lse, nullptr, nullptr, B, N, Hq, Hkv, head_dim, drop_p, drop_seed, stream);
}

extern "C" int jat_attention_passes(int N) { return N <= 0 ? 0 : (N + ATT_MAX_NK - 1) / ATT_MAX_NK; }

extern "C" int jat_gqa_attention_fwd_long(jat_ctx* ctx, const void* qkv, void* out, float* lse, void* part_o, float* part_lse,
                                          int B, int N, int Hq, int Hkv, int head_dim, float drop_p, uint32_t drop_seed,
                                          void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !qkv || !out) return fail(JAT_ERR_BAD_ARG, "jat_gqa_attention_fwd: null argument");
    if (head_dim != ATT_HD) return fail(JAT_ERR_BAD_SHAPE, "jat_gqa_attention_fwd: head_dim must be 64 (got %d)", head_dim);
    if (B <= 0 || N <= 0 || Hq <= 0 || Hkv <= 0 || Hq % Hkv != 0 || B > 65535 || Hkv > 65535)
        return fail(JAT_ERR_BAD_SHAPE, "jat_gqa_attention_fwd: bad B/N/heads");
    if ((long long)B * Hq * N > 0xffffffffll) return fail(JAT_ERR_BAD_SHAPE, "jat_gqa_attention_fwd: B*Hq*N exceeds 2^32");
    DropCfg drop;
    if (!make_drop(drop_p, drop_seed, &drop)) return fail(JAT_ERR_BAD_ARG, "jat_gqa_attention_fwd: drop_p must be in [0, 1)");
    cudaStream_t s = (cudaStream_t)stream;
    const int passes = jat_attention_passes(N);
    if (passes == 1) return attention_pass(ctx, qkv, out, lse, B, N, Hq, Hkv, 0, N, drop, s);
    if (!part_o || !part_lse)
        return fail(JAT_ERR_BAD_ARG, "jat_gqa_attention_fwd_long: N = %d needs part_o (bf16 [%d, B*N, Hq*64]) and part_lse "
                    "(f32 [%d, B, Hq, N])", N, passes, passes);
    const long long o_pass = (long long)B * N * Hq * ATT_HD, l_pass = (long long)B * Hq * N;
    for (int c = 0; c < passes; ++c) {
        const int key0 = c * ATT_MAX_NK, nk = N - key0 < ATT_MAX_NK ? N - key0 : ATT_MAX_NK;
        JAT_TRY(attention_pass(ctx, qkv, (__nv_bfloat16*)part_o + c * o_pass, part_lse + c * l_pass, B, N, Hq, Hkv, key0, nk, drop, s));
    }
    const long long warps = (long long)B * N * Hq;
    pre_launch(ctx, TAG_ATTN, s);
    attention_combine_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, s>>>((const __nv_bfloat16*)part_o, part_lse,
                                                                                 (__nv_bfloat16*)out, lse, B, N, Hq, passes);
    return post_launch(ctx, "attention_combine");
}

extern "C" int jat_gqa_attention_bwd(jat_ctx* ctx, const void* qkv, const void* d_out, const void* out, const float* lse,
                                     float* dsum_scratch, float* dq_acc_scratch, void* dqkv, const float* rope_cos,
                                     const float* rope_sin, int B, int N, int Hq, int Hkv, int head_dim, void* stream) {
    DeviceGuard dev_guard__(ctx);
    return jat_gqa_attention_bwd_dropout(ctx, qkv, d_out, out, lse, dsum_scratch, dq_acc_scratch, dqkv, rope_cos, rope_sin, B, N,
                                         Hq, Hkv, head_dim, 0.0f, 0u, stream);
}

extern "C" int jat_gqa_attention_bwd_dropout(jat_ctx* ctx, const void* qkv, const void* d_out, const void* out, const float* lse,
                                             float* dsum_scratch, float* dq_acc_scratch, void* dqkv, const float* rope_cos,
                                             const float* rope_sin, int B, int N, int Hq, int Hkv, int head_dim, float drop_p,
                                             uint32_t drop_seed, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !qkv || !d_out || !out || !lse || !dsum_scratch || !dq_acc_scratch || !dqkv || !rope_cos || !rope_sin)
        return fail(JAT_ERR_BAD_ARG, "jat_gqa_attention_bwd: null argument");
    if (head_dim != ATT_HD) return fail(JAT_ERR_BAD_SHAPE, "jat_gqa_attention_bwd: head_dim must be 64 (got %d)", head_dim);
    if (B <= 0 || N <= 0 || Hq <= 0 || Hkv <= 0 || Hq % Hkv != 0 || B > 65535 || Hkv > 65535)
        return fail(JAT_ERR_BAD_SHAPE, "jat_gqa_attention_bwd: bad B/N/heads");
    cudaStream_t s = (cudaStream_t)stream;
    const uint64_t rows = (uint64_t)B * N, qcols = (uint64_t)Hq * ATT_HD, cols = (uint64_t)(Hq + 2 * Hkv) * ATT_HD;
    JAT_CUDA(cudaMemsetAsync(dq_acc_scratch, 0, rows * qcols * sizeof(float), s));
    {
        const long long n = (long long)rows * Hq * 8;
        pre_launch(ctx, TAG_ATTN_BWD, s);
        attn_bwd_rowdot_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const __nv_bfloat16*)d_out, (const __nv_bfloat16*)out,
                                                                         dsum_scratch, B, N, Hq);
        JAT_TRY(post_launch(ctx, "attn_bwd_rowdot"));
    }
    CUtensorMap tqkv, tdo, tdq;
    JAT_TRY(make_tmap(ctx, &tqkv, qkv, rows, cols, cols, ATTB_TILE));
    JAT_TRY(make_tmap(ctx, &tdo, d_out, rows, qcols, qcols, ATTB_TILE));
    JAT_TRY(make_tmap(ctx, &tdq, dq_acc_scratch, rows, qcols, qcols, ATTB_TILE, /*f32=*/true));
    AttnBwdParams p = {};
    p.B = B; p.N = N; p.Hq = Hq; p.Hkv = Hkv; p.G = Hq / Hkv;
    p.lse = lse; p.dsum = dsum_scratch; p.dqkv = (__nv_bfloat16*)dqkv;
    p.rope_cos = rope_cos; p.rope_sin = rope_sin;
    p.scale = 0.125f;
    p.scale_log2e = 0.125f * 1.4426950408889634f;
    if (!make_drop(drop_p, drop_seed, &p.drop)) return fail(JAT_ERR_BAD_ARG, "jat_gqa_attention_bwd: drop_p must be in [0, 1)");
    p.trace = ctx->att_trace;
    // v2 (default): query-row threads, score MMAs overlapped with the math; JAT_ATTN_BWD=1 selects the first version
    static const bool use_v1 = getenv("JAT_ATTN_BWD") && atoi(getenv("JAT_ATTN_BWD")) == 1;
    dim3 grid((N + ATTB_TILE - 1) / ATTB_TILE, Hkv, B);
    const dim3 grid1d((unsigned)(grid.x * grid.y * grid.z));   // v2 decodes (key tile, KV head, batch item) itself, heavy tiles first
    pre_launch(ctx, TAG_ATTN_BWD, s);
    if (use_v1) {
        JAT_TRY(ensure_dyn_smem(ctx, (const void*)gqa_attention_bwd_v1_kernel, ATTB_SMEM_BYTES));
        gqa_attention_bwd_v1_kernel<<<grid, ATTB_THREADS, ATTB_SMEM_BYTES, s>>>(tqkv, tdo, tdq, p);
    } else if (p.drop.thresh != 0u) {
        JAT_TRY(ensure_dyn_smem(ctx, (const void*)gqa_attention_bwd_kernel<true>, ATTB2_SMEM_BYTES));
        gqa_attention_bwd_kernel<true><<<grid1d, ATTB2_THREADS, ATTB2_SMEM_BYTES, s>>>(tqkv, tdo, tdq, p);
    } else {
        JAT_TRY(ensure_dyn_smem(ctx, (const void*)gqa_attention_bwd_kernel<false>, ATTB2_SMEM_BYTES));
        gqa_attention_bwd_kernel<false><<<grid1d, ATTB2_THREADS, ATTB2_SMEM_BYTES, s>>>(tqkv, tdo, tdq, p);
    }
    JAT_TRY(post_launch(ctx, "gqa_attention_bwd"));
    {
        const long long n = (long long)rows * Hq * 8;
        pre_launch(ctx, TAG_ATTN_BWD, s);
        attn_bwd_dq_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dq_acc_scratch, (__nv_bfloat16*)dqkv, rope_cos,
                                                                              rope_sin, B, N, Hq, Hkv, use_v1 ? 1.0f : p.scale);
        return post_launch(ctx, "attn_bwd_dq_finalize");
    }
}

// ------------------------------------------------------------------------------------------------ DiT forward plan
static jat_gemm_epilogue epi_bias_act(const float* bias, void* out, int64_t ldo, int act, int dtype) {
    jat_gemm_epilogue e;
    memset(&e, 0, sizeof(e));
    e.kind = JAT_EPI_BIAS_ACT; e.act = act; e.out_dtype = dtype; e.bias = bias; e.out = out; e.ldo = ldo;
    return e;
}

extern "C" int jat_dit_modulation(jat_ctx* ctx, const jat_dit_weights* w, const jat_dit_workspace* ws, const float* t,
                                  int Bt, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !w || !ws || !t) return fail(JAT_ERR_BAD_ARG, "jat_dit_modulation: null argument");
    if (!ws->t_feat || !ws->t_hid || !ws->t_act || !ws->mod) return fail(JAT_ERR_BAD_ARG, "jat_dit_modulation: workspace incomplete");
    const int D = w->hidden;
    const int NM = w->depth * 6 * D;
    JAT_TRY(jat_timestep_features(ctx, t, ws->t_feat, Bt, D, stream));
    jat_gemm_epilogue e1 = epi_bias_act(w->te_b1, ws->t_hid, D, JAT_ACT_SILU, JAT_DTYPE_BF16);
    JAT_TRY(jat_gemm_bf16(ctx, ws->t_feat, D, w->te_w1, D, Bt, D, D, &e1, -1, 0, stream));
    // adaLN_modulation = SiLU -> Linear: t_emb is only ever consumed through that SiLU, so fuse it here.
    jat_gemm_epilogue e2 = epi_bias_act(w->te_b2, ws->t_act, D, JAT_ACT_SILU, JAT_DTYPE_BF16);
    JAT_TRY(jat_gemm_bf16(ctx, ws->t_hid, D, w->te_w2, D, Bt, D, D, &e2, -1, 0, stream));
    jat_gemm_epilogue e3 = epi_bias_act(w->ada_b, ws->mod, NM, JAT_ACT_NONE, JAT_DTYPE_F32);
    JAT_TRY(jat_gemm_bf16(ctx, ws->t_act, D, w->ada_w, D, Bt, NM, D, &e3, -1, 0, stream));
    return 0;
}

extern "C" int jat_dit_forward_tokens(jat_ctx* ctx, const jat_dit_weights* w, const jat_dit_workspace* ws,
                                      const float* x_t, int xt_batch, const float* x_cond, int cond_batch,
                                      const float* mod, int64_t mod_batch_stride, float* out, int B, int T,
                                      void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !w || !ws || !x_t || !mod || !out) return fail(JAT_ERR_BAD_ARG, "jat_dit_forward_tokens: null argument");
    const int D = w->hidden, P = w->patch_len, C = w->channels, F = w->mlp_hidden, BD = w->bottleneck;
    const int N = (T + P - 1) / P;
    if (N > w->max_len) return fail(JAT_ERR_SEQ_TOO_LONG, "Sequence length %d exceeds max_len %d", N, w->max_len);
    if (N > w->rope_max_pos) return fail(JAT_ERR_SEQ_TOO_LONG, "Sequence length %d exceeds RoPE table %d", N, w->rope_max_pos);
    if (w->head_dim != 64) return fail(JAT_ERR_BAD_SHAPE, "head_dim must be 64");
    const int M = B * N;
    const int QKV = (w->n_q_heads + 2 * w->n_kv_heads) * 64;
    const int Cc = w->cond_channels > 0 ? w->cond_channels : C;
    const int KIN = (C + Cc) * P;
    cudaStream_t s = (cudaStream_t)stream;

    JAT_TRY(jat_patchify_cast2(ctx, x_t, xt_batch, x_cond, cond_batch, ws->patches, B, C, Cc, T, P, stream));
    jat_gemm_epilogue e = epi_bias_act(w->pe_b1, ws->pe_hid, BD, JAT_ACT_GELU_ERF, JAT_DTYPE_BF16);
    JAT_TRY(jat_gemm_bf16(ctx, ws->patches, KIN, w->pe_w1, KIN, M, BD, KIN, &e, -1, 0, stream));
    e = epi_bias_act(w->pe_b2, ws->x, D, JAT_ACT_NONE, JAT_DTYPE_F32);
    JAT_TRY(jat_gemm_bf16(ctx, ws->pe_hid, BD, w->pe_w2, BD, M, D, BD, &e, -1, 0, stream));

    for (int i = 0; i < w->depth; ++i) {
        const float* m = mod + (int64_t)i * 6 * D;  // shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp
        JAT_TRY(jat_adaln_norm_modulate(ctx, ws->x, ws->h, m, m + D, mod_batch_stride,
                                        w->norm1_w ? w->norm1_w[i] : nullptr, w->norm_kind, w->norm_eps, M, D, N, stream));
        memset(&e, 0, sizeof(e));
        e.kind = JAT_EPI_QKV_ROPE; e.out = ws->qkv; e.ldo = QKV; e.tokens_per_batch = N;
        e.rope_cos = w->rope_cos; e.rope_sin = w->rope_sin; e.rope_cols = (w->n_q_heads + w->n_kv_heads) * 64;
        JAT_TRY(jat_gemm_bf16(ctx, ws->h, D, w->wqkv[i], D, M, QKV, D, &e, -1, 0, stream));
        JAT_TRY(jat_gqa_attention_fwd_long(ctx, ws->qkv, ws->attn, nullptr, ws->attn_part, ws->lse_part, B, N, w->n_q_heads,
                                           w->n_kv_heads, 64, 0.0f, 0u, stream));
        memset(&e, 0, sizeof(e));
        e.kind = JAT_EPI_GATE_RESIDUAL; e.out = ws->x; e.ldo = D; e.tokens_per_batch = N;
        e.gate = m + 2 * D; e.gate_batch_stride = mod_batch_stride;
        JAT_TRY(jat_gemm_bf16(ctx, ws->attn, D, w->wo[i], D, M, D, D, &e, -1, 0, stream));

        JAT_TRY(jat_adaln_norm_modulate(ctx, ws->x, ws->h, m + 3 * D, m + 4 * D, mod_batch_stride,
                                        w->norm2_w ? w->norm2_w[i] : nullptr, w->norm_kind, w->norm_eps, M, D, N, stream));
        e = epi_bias_act(w->b1[i], ws->mlp_hid, F, JAT_ACT_GELU_ERF, JAT_DTYPE_BF16);
        JAT_TRY(jat_gemm_bf16(ctx, ws->h, D, w->w1[i], D, M, F, D, &e, -1, 0, stream));
        memset(&e, 0, sizeof(e));
        e.kind = JAT_EPI_GATE_RESIDUAL; e.out = ws->x; e.ldo = D; e.tokens_per_batch = N; e.bias = w->b2[i];
        e.gate = m + 5 * D; e.gate_batch_stride = mod_batch_stride;
        JAT_TRY(jat_gemm_bf16(ctx, ws->mlp_hid, F, w->w2[i], F, M, D, F, &e, -1, 0, stream));
        if (ws->block_out)
            JAT_CUDA(cudaMemcpyAsync(ws->block_out + (int64_t)i * M * D, ws->x, (size_t)M * D * sizeof(float),
                                     cudaMemcpyDeviceToDevice, s));
    }
    JAT_TRY(jat_adaln_norm_modulate(ctx, ws->x, ws->h, nullptr, nullptr, 0, w->final_norm_w, w->norm_kind, w->norm_eps, M,
                                    D, N, stream));
    memset(&e, 0, sizeof(e));
    e.kind = JAT_EPI_UNPATCHIFY; e.out = out; e.bias = w->final_b; e.tokens_per_batch = N; e.patch_len = P; e.t_out = T;
    JAT_TRY(jat_gemm_bf16(ctx, ws->h, D, w->final_w, D, M, C * P, D, &e, -1, 0, stream));
    return 0;
}


// ------------------------------------------------------------------------------------------------ training step
// Training-mode forward: the same kernels as jat_dit_forward_tokens, but every activation
// the backward pass needs is kept in `sv` (per-block slabs, leading dimension = depth) instead of being overwritten.
static jat_gemm_epilogue epi_plain(int kind, void* out, int64_t ldo) {
    jat_gemm_epilogue e;
    memset(&e, 0, sizeof(e));
    e.kind = kind; e.out = out; e.ldo = ldo; e.out_dtype = JAT_DTYPE_BF16;
    return e;
}
static char* at(void* base, int64_t index, int64_t elems, int64_t esz) { return (char*)base + index * elems * esz; }

extern "C" int jat_dit_forward_train(jat_ctx* ctx, const jat_dit_weights* w, const jat_dit_workspace* ws,
                                     const jat_dit_saved* sv, const float* x_t, const float* x_cond, const float* t,
                                     float* out, int B, int T, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !w || !ws || !sv || !x_t || !x_cond || !t || !out) return fail(JAT_ERR_BAD_ARG, "jat_dit_forward_train: null argument");
    const int D = w->hidden, P = w->patch_len, C = w->channels, F = w->mlp_hidden, BD = w->bottleneck;
    const int N = (T + P - 1) / P;
    if (N > w->max_len) return fail(JAT_ERR_SEQ_TOO_LONG, "Sequence length %d exceeds max_len %d", N, w->max_len);
    if (w->head_dim != 64) return fail(JAT_ERR_BAD_SHAPE, "head_dim must be 64");
    const int M = B * N, Hq = w->n_q_heads, Hkv = w->n_kv_heads;
    const int Cc = w->cond_channels > 0 ? w->cond_channels : C;
    const int QKV = (Hq + 2 * Hkv) * 64, KIN = (C + Cc) * P, NM = w->depth * 6 * D;
    jat_gemm_epilogue e;
    // ---- train-mode regularisers: Dropout(p) on attention probabilities / GELU output / mlp.3 output, DropPath per block
    const float pd = sv->dropout_p;
    if (!(pd >= 0.0f) || pd >= 1.0f) return fail(JAT_ERR_BAD_ARG, "jat_dit_forward_train: dropout_p must be in [0, 1)");
    const float* dps = nullptr;  // [depth, 2, B] per-sample DropPath factors
    if (sv->drop_path_rates != nullptr) {
        if (!sv->dp_scale) return fail(JAT_ERR_BAD_ARG, "jat_dit_forward_train: drop_path_rates given without dp_scale");
        JAT_TRY(jat_drop_path_scales(ctx, sv->dp_scale, sv->drop_path_rates, w->depth, B, sv->seed, stream));
        dps = sv->dp_scale;
    }

    // ---- timestep path with the pre-activations kept (t_embedder + adaLN_modulation of every block)
    JAT_TRY(jat_timestep_features(ctx, t, ws->t_feat, B, D, stream));
    e = epi_bias_act(w->te_b1, ws->t_hid, D, JAT_ACT_SILU, JAT_DTYPE_BF16); e.aux = sv->t_u1; e.ld_aux = D;
    JAT_TRY(jat_gemm_bf16(ctx, ws->t_feat, D, w->te_w1, D, B, D, D, &e, -1, 0, stream));
    e = epi_bias_act(w->te_b2, ws->t_act, D, JAT_ACT_SILU, JAT_DTYPE_BF16); e.aux = sv->t_u2; e.ld_aux = D;
    JAT_TRY(jat_gemm_bf16(ctx, ws->t_hid, D, w->te_w2, D, B, D, D, &e, -1, 0, stream));
    e = epi_bias_act(w->ada_b, ws->mod, NM, JAT_ACT_NONE, JAT_DTYPE_F32);
    JAT_TRY(jat_gemm_bf16(ctx, ws->t_act, D, w->ada_w, D, B, NM, D, &e, -1, 0, stream));

    // ---- patch embed
    JAT_TRY(jat_patchify_cast2(ctx, x_t, B, x_cond, B, ws->patches, B, C, Cc, T, P, stream));
    e = epi_bias_act(w->pe_b1, ws->pe_hid, BD, JAT_ACT_GELU_ERF, JAT_DTYPE_BF16); e.aux = sv->pe_u; e.ld_aux = BD;
    JAT_TRY(jat_gemm_bf16(ctx, ws->patches, KIN, w->pe_w1, KIN, M, BD, KIN, &e, -1, 0, stream));
    e = epi_bias_act(w->pe_b2, ws->x, D, JAT_ACT_NONE, JAT_DTYPE_F32);
    JAT_TRY(jat_gemm_bf16(ctx, ws->pe_hid, BD, w->pe_w2, BD, M, D, BD, &e, -1, 0, stream));

    const int64_t MD = (int64_t)M * D;
    for (int i = 0; i < w->depth; ++i) {
        const float* m = ws->mod + (int64_t)i * 6 * D;
        void* h1 = at(sv->h1, i, MD, 2);
        void* qkv = at(sv->qkv, i, (int64_t)M * QKV, 2);
        void* attn = at(sv->attn, i, MD, 2);
        void* h2 = at(sv->h2, i, MD, 2);
        void* mact = at(sv->mact, i, (int64_t)M * F, 2);
        // the norm kernel also keeps x (f32) and the row statistics for the backward pass
        JAT_TRY(adaln_norm_modulate_stats(ctx, ws->x, h1, m, m + D, NM, w->norm1_w ? w->norm1_w[i] : nullptr, w->norm_kind,
                                          w->norm_eps, M, D, N, sv->rs1 ? sv->rs1 + (int64_t)i * M * 2 : nullptr,
                                          (float*)at(sv->x_in, i, MD, 4), stream));
        memset(&e, 0, sizeof(e));
        e.kind = JAT_EPI_QKV_ROPE; e.out = qkv; e.ldo = QKV; e.tokens_per_batch = N;
        e.rope_cos = w->rope_cos; e.rope_sin = w->rope_sin; e.rope_cols = (Hq + Hkv) * 64;
        JAT_TRY(jat_gemm_bf16(ctx, h1, D, w->wqkv[i], D, M, QKV, D, &e, -1, 0, stream));
        JAT_TRY(jat_gqa_attention_fwd_long(ctx, qkv, attn, (float*)at(sv->lse, i, (int64_t)B * Hq * N, 4), ws->attn_part,
                                           ws->lse_part, B, N, Hq, Hkv, 64, pd,
                                           jat_dropout_site_seed(sv->seed, i, JAT_DROP_SITE_ATTN), stream));
        memset(&e, 0, sizeof(e));
        e.kind = JAT_EPI_GATE_RESIDUAL; e.out = ws->x; e.ldo = D; e.tokens_per_batch = N;
        e.gate = m + 2 * D; e.gate_batch_stride = NM; e.aux = at(sv->y1, i, MD, 2); e.ld_aux = D;
        e.gate_rowscale = dps ? dps + (int64_t)(2 * i) * B : nullptr;
        JAT_TRY(jat_gemm_bf16(ctx, attn, D, w->wo[i], D, M, D, D, &e, -1, 0, stream));

        JAT_TRY(adaln_norm_modulate_stats(ctx, ws->x, h2, m + 3 * D, m + 4 * D, NM, w->norm2_w ? w->norm2_w[i] : nullptr,
                                          w->norm_kind, w->norm_eps, M, D, N, sv->rs2 ? sv->rs2 + (int64_t)i * M * 2 : nullptr,
                                          (float*)at(sv->x_mid, i, MD, 4), stream));
        e = epi_bias_act(w->b1[i], mact, F, JAT_ACT_GELU_ERF, JAT_DTYPE_BF16);
        e.aux = at(sv->u, i, (int64_t)M * F, 2); e.ld_aux = F;
        e.drop_p = pd; e.drop_seed = jat_dropout_site_seed(sv->seed, i, JAT_DROP_SITE_MLP_HIDDEN);
        JAT_TRY(jat_gemm_bf16(ctx, h2, D, w->w1[i], D, M, F, D, &e, -1, 0, stream));
        memset(&e, 0, sizeof(e));
        e.kind = JAT_EPI_GATE_RESIDUAL; e.out = ws->x; e.ldo = D; e.tokens_per_batch = N; e.bias = w->b2[i];
        e.gate = m + 5 * D; e.gate_batch_stride = NM; e.aux = at(sv->y2, i, MD, 2); e.ld_aux = D;
        e.drop_p = pd; e.drop_seed = jat_dropout_site_seed(sv->seed, i, JAT_DROP_SITE_MLP_OUT);
        e.gate_rowscale = dps ? dps + (int64_t)(2 * i + 1) * B : nullptr;
        JAT_TRY(jat_gemm_bf16(ctx, mact, F, w->w2[i], F, M, D, F, &e, -1, 0, stream));
    }
    // final layer: ws->x (input of the final norm) and ws->h (its output) stay valid until the backward pass
    JAT_TRY(jat_adaln_norm_modulate(ctx, ws->x, ws->h, nullptr, nullptr, 0, w->final_norm_w, w->norm_kind, w->norm_eps, M,
                                    D, N, stream));
    memset(&e, 0, sizeof(e));
    e.kind = JAT_EPI_UNPATCHIFY; e.out = out; e.bias = w->final_b; e.tokens_per_batch = N; e.patch_len = P; e.t_out = T;
    JAT_TRY(jat_gemm_bf16(ctx, ws->h, D, w->final_w, D, M, C * P, D, &e, -1, 0, stream));
    return 0;
}

// weight gradient  g[Nout, Kin] += dY^T[Nout, Mtok] X[Mtok, Kin]  (both operands as stored), reduction split so that
// the grid holds a few waves of work items
static int wgrad(jat_ctx* ctx, const void* dY, int64_t ld_dy, const void* X, int64_t ld_x, float* g, int Nout, int Kin,
                 int Mtok, void* stream) {
    jat_gemm_epilogue e = epi_plain(JAT_EPI_ACCUM, g, Kin);
    e.a_transposed = 1; e.w_transposed = 1;
    const int bn = (Kin % 256 == 0) ? 256 : 128;
    const long long tiles = (long long)((Nout + 255) / 256) * (Kin / bn);
    const int kblocks = (Mtok + GEMM_BK - 1) / GEMM_BK;
    // split the token reduction s ways so that the persistent schedule's makespan is shortest: every CTA pair runs
    // ceil(tiles * s / pairs) items of ceil(kblocks / s) k-blocks each, plus ~4 k-blocks' worth of f32 reduce-add epilogue
    const long long pairs = ctx->gemm_sms / 2;
    long long best = 1, best_cost = -1;
    for (long long sp = 1; sp <= 8 && sp <= kblocks; ++sp) {
        const long long cost = ((tiles * sp + pairs - 1) / pairs) * ((kblocks + sp - 1) / sp + 4);
        if (best_cost < 0 || cost < best_cost) { best = sp; best_cost = cost; }
    }
    e.k_splits = (int)best;
    return jat_gemm_bf16(ctx, dY, ld_dy, X, ld_x, Nout, Kin, Mtok, &e, -1, 0, stream);
}
// input gradient  dX[Mtok, Kin] = dY[Mtok, Nout] W[Nout, Kin]  (W as stored), optional * act'(u)
static int dgrad(jat_ctx* ctx, const void* dY, int64_t ld_dy, const void* W, int Kin, int Nout, int Mtok, void* dX, int act,
                 const void* u, void* stream, float drop_p = 0.0f, uint32_t drop_seed = 0u) {
    jat_gemm_epilogue e = epi_plain(act == JAT_ACT_NONE ? JAT_EPI_BIAS_ACT : JAT_EPI_DACT, dX, Kin);
    e.act = act; e.w_transposed = 1; e.aux = const_cast<void*>(u); e.ld_aux = Kin;
    e.drop_p = drop_p; e.drop_seed = drop_seed;
    return jat_gemm_bf16(ctx, dY, ld_dy, W, Kin, Mtok, Kin, Nout, &e, -1, 0, stream);
}

// The backward pass in three stages, so that a framework can hand the gradients of a stage to its gradient
// all-reduce (DDP buckets) while the next stage runs:  begin (final layer) -> block depth-1 ... block 0 -> end
// (patch embed + timestep path).  dmod is block-major: [depth][B][6*hidden].
struct BwdDims {
    int D, P, C, F, BD, N, M, Hq, Hkv, QKV, KIN, SIXD, CP;
    int64_t MD;
};
static BwdDims bwd_dims(const jat_dit_weights* w, int B, int T) {
    BwdDims d;
    d.D = w->hidden; d.P = w->patch_len; d.C = w->channels; d.F = w->mlp_hidden; d.BD = w->bottleneck;
    d.N = (T + d.P - 1) / d.P; d.M = B * d.N; d.Hq = w->n_q_heads; d.Hkv = w->n_kv_heads;
    d.QKV = (d.Hq + 2 * d.Hkv) * 64; d.KIN = (d.C + (w->cond_channels > 0 ? w->cond_channels : d.C)) * d.P;
    d.SIXD = 6 * d.D; d.CP = d.C * d.P;
    d.MD = (int64_t)d.M * d.D;
    return d;
}

extern "C" int jat_dit_backward_begin(jat_ctx* ctx, const jat_dit_weights* w, const jat_dit_workspace* ws,
                                      const jat_dit_saved* sv, const jat_dit_bwd_scratch* sc, const jat_dit_weights* gr,
                                      const float* d_out, int B, int T, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !w || !ws || !sv || !sc || !gr || !d_out) return fail(JAT_ERR_BAD_ARG, "jat_dit_backward_begin: null argument");
    const BwdDims d = bwd_dims(w, B, T);
    cudaStream_t s = (cudaStream_t)stream;
    const bool rms = w->norm_kind == JAT_NORM_RMSNORM;
    JAT_CUDA(cudaMemsetAsync(sc->dmod, 0, (size_t)w->depth * B * d.SIXD * sizeof(float), s));
    JAT_CUDA(cudaMemsetAsync(sc->dt_acc, 0, (size_t)B * d.D * sizeof(float), s));
    // final layer: out = unpatchify(hf Wf^T + bf), hf = norm(x_L)
    JAT_TRY(jat_patchify_single(ctx, d_out, sc->dout_p, B, d.C, T, d.P, stream));
    JAT_TRY(wgrad(ctx, sc->dout_p, d.CP, ws->h, d.D, (float*)gr->final_w, d.CP, d.D, d.M, stream));
    JAT_TRY(jat_colsum_bf16(ctx, sc->dout_p, d.CP, d.M, d.CP, (float*)gr->final_b, stream));
    JAT_TRY(dgrad(ctx, sc->dout_p, d.CP, w->final_w, d.D, d.CP, d.M, sc->dh, JAT_ACT_NONE, nullptr, stream));
    return jat_adaln_bwd(ctx, sc->dh, ws->x, nullptr, 0, w->final_norm_w, w->norm_kind, w->norm_eps, sc->dx, 0, nullptr,
                         nullptr, 0, rms ? (float*)gr->final_norm_w : nullptr, sc->rowstats, B, d.N, d.D, stream);
}

extern "C" int jat_dit_backward_block(jat_ctx* ctx, const jat_dit_weights* w, const jat_dit_workspace* ws,
                                      const jat_dit_saved* sv, const jat_dit_bwd_scratch* sc, const jat_dit_weights* gr,
                                      int i, int B, int T, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !w || !ws || !sv || !sc || !gr) return fail(JAT_ERR_BAD_ARG, "jat_dit_backward_block: null argument");
    if (i < 0 || i >= w->depth) return fail(JAT_ERR_BAD_ARG, "jat_dit_backward_block: block index %d out of range", i);
    const BwdDims d = bwd_dims(w, B, T);
    const int D = d.D, F = d.F, M = d.M, N = d.N, NM = w->depth * d.SIXD;
    const int64_t MD = d.MD;
    const bool rms = w->norm_kind == JAT_NORM_RMSNORM;
    const float* m = ws->mod + (int64_t)i * d.SIXD;               // forward modulation: [B][depth*6D], row stride NM
    float* dm = sc->dmod + (int64_t)i * B * d.SIXD;               // its gradient: block-major, row stride 6D
    const void* h1 = at(sv->h1, i, MD, 2);
    const void* qkv = at(sv->qkv, i, (int64_t)M * d.QKV, 2);
    const void* attn = at(sv->attn, i, MD, 2);
    const void* h2 = at(sv->h2, i, MD, 2);
    const void* u = at(sv->u, i, (int64_t)M * F, 2);
    const void* mact = at(sv->mact, i, (int64_t)M * F, 2);
    const float pd = sv->dropout_p;
    const float* dps = sv->drop_path_rates != nullptr ? sv->dp_scale : nullptr;
    const bool fused = sv->rs1 != nullptr && sv->rs2 != nullptr;
    // ---- MLP branch: x2 = x1 + drop_path(gate_mlp * drop(drop(gelu(h2 W1^T + b1)) W2^T + b2))
    // (fused path: the norm1 backward of block i + 1 has already done this gate backward on the dx rows it produced)
    if (!fused || i == w->depth - 1)
    JAT_TRY(jat_gate_bwd_dropout(ctx, sc->dx, at(sv->y2, i, MD, 2), m + 5 * D, NM, sc->dy, dm + 5 * D, d.SIXD, sc->dxsum,
                                 (float*)gr->b2[i], B, N, D, pd, jat_dropout_site_seed(sv->seed, i, JAT_DROP_SITE_MLP_OUT),
                                 dps ? dps + (int64_t)(2 * i + 1) * B : nullptr, stream));
    JAT_TRY(wgrad(ctx, sc->dy, D, mact, F, (float*)gr->w2[i], D, F, M, stream));
    JAT_TRY(dgrad(ctx, sc->dy, D, w->w2[i], F, D, M, sc->du, JAT_ACT_GELU_ERF, u, stream, pd,
                  jat_dropout_site_seed(sv->seed, i, JAT_DROP_SITE_MLP_HIDDEN)));
    JAT_TRY(wgrad(ctx, sc->du, F, h2, D, (float*)gr->w1[i], F, D, M, stream));
    JAT_TRY(jat_colsum_bf16(ctx, sc->du, F, M, F, (float*)gr->b1[i], stream));
    JAT_TRY(dgrad(ctx, sc->du, F, w->w1[i], D, F, M, sc->dh, JAT_ACT_NONE, nullptr, stream));
    if (fused) {
        // norm2 backward + the gate backward of the attention branch below it (x1 = x + drop_path(gate_msa * (attn(h1) Wo^T)))
        JAT_TRY(jat_adaln_gate_bwd(ctx, sc->dh, (const float*)at(sv->x_mid, i, MD, 4), sv->rs2 + (int64_t)i * M * 2, m + 4 * D, NM,
                                   w->norm2_w ? w->norm2_w[i] : nullptr, w->norm_kind, sc->dx, dm + 3 * D, dm + 4 * D, d.SIXD,
                                   rms ? (float*)gr->norm2_w[i] : nullptr, at(sv->y1, i, MD, 2), m + 2 * D, sc->dy, dm + 2 * D,
                                   nullptr, nullptr, B, N, D, 0.0f, 0u, dps ? dps + (int64_t)(2 * i) * B : nullptr, stream));
    } else {
    JAT_TRY(jat_adaln_bwd(ctx, sc->dh, (const float*)at(sv->x_mid, i, MD, 4), m + 4 * D, NM, w->norm2_w ? w->norm2_w[i] : nullptr,
                          w->norm_kind, w->norm_eps, sc->dx, 1, dm + 3 * D, dm + 4 * D, d.SIXD,
                          rms ? (float*)gr->norm2_w[i] : nullptr, sc->rowstats, B, N, D, stream));
    // ---- attention branch: x1 = x + drop_path(gate_msa * (attn(h1) Wo^T))
    JAT_TRY(jat_gate_bwd_dropout(ctx, sc->dx, at(sv->y1, i, MD, 2), m + 2 * D, NM, sc->dy, dm + 2 * D, d.SIXD, nullptr, nullptr,
                                 B, N, D, 0.0f, 0u, dps ? dps + (int64_t)(2 * i) * B : nullptr, stream));
    }
    JAT_TRY(wgrad(ctx, sc->dy, D, attn, D, (float*)gr->wo[i], D, D, M, stream));
    JAT_TRY(dgrad(ctx, sc->dy, D, w->wo[i], D, D, M, sc->da, JAT_ACT_NONE, nullptr, stream));
    JAT_TRY(jat_gqa_attention_bwd_dropout(ctx, qkv, sc->da, attn, (const float*)at(sv->lse, i, (int64_t)B * d.Hq * N, 4), sc->dsum,
                                          sc->dq_acc, sc->dqkv, w->rope_cos, w->rope_sin, B, N, d.Hq, d.Hkv, 64, pd,
                                          jat_dropout_site_seed(sv->seed, i, JAT_DROP_SITE_ATTN), stream));
    JAT_TRY(wgrad(ctx, sc->dqkv, d.QKV, h1, D, (float*)gr->wqkv[i], d.QKV, D, M, stream));
    JAT_TRY(dgrad(ctx, sc->dqkv, d.QKV, w->wqkv[i], D, d.QKV, M, sc->dh, JAT_ACT_NONE, nullptr, stream));
    if (fused && i > 0) {
        // norm1 backward + the gate backward of the MLP branch of block i - 1 (the next stage), whose modulation row and
        // gradient slab sit 6D before this block's: gate_mlp(i-1) = m - D, dgate_mlp(i-1) = dm - B*6D + 5D.  The kernel
        // addresses gate / dgate with the strides of scale / dshift, so both are passed relative to this block's views.
        JAT_TRY(jat_adaln_gate_bwd(ctx, sc->dh, (const float*)at(sv->x_in, i, MD, 4), sv->rs1 + (int64_t)i * M * 2, m + D, NM,
                                   w->norm1_w ? w->norm1_w[i] : nullptr, w->norm_kind, sc->dx, dm, dm + D, d.SIXD,
                                   rms ? (float*)gr->norm1_w[i] : nullptr, at(sv->y2, i - 1, MD, 2), m - D, sc->dy,
                                   dm - (int64_t)B * d.SIXD + 5 * D, nullptr, (float*)gr->b2[i - 1], B, N, D, pd,
                                   jat_dropout_site_seed(sv->seed, i - 1, JAT_DROP_SITE_MLP_OUT),
                                   dps ? dps + (int64_t)(2 * (i - 1) + 1) * B : nullptr, stream));
    } else if (fused) {
        JAT_TRY(jat_adaln_gate_bwd(ctx, sc->dh, (const float*)at(sv->x_in, i, MD, 4), sv->rs1 + (int64_t)i * M * 2, m + D, NM,
                                   w->norm1_w ? w->norm1_w[i] : nullptr, w->norm_kind, sc->dx, dm, dm + D, d.SIXD,
                                   rms ? (float*)gr->norm1_w[i] : nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, B,
                                   N, D, 0.0f, 0u, nullptr, stream));
    } else
    JAT_TRY(jat_adaln_bwd(ctx, sc->dh, (const float*)at(sv->x_in, i, MD, 4), m + D, NM, w->norm1_w ? w->norm1_w[i] : nullptr,
                          w->norm_kind, w->norm_eps, sc->dx, 1, dm, dm + D, d.SIXD, rms ? (float*)gr->norm1_w[i] : nullptr,
                          sc->rowstats, B, N, D, stream));
    // ---- this block's adaLN_modulation Linear: mod_i = t_act Wada_i^T + bada_i  (dmod_i is complete now)
    void* dmb = at(sc->dmod_bf16, i, (int64_t)B * d.SIXD, 2);
    JAT_TRY(jat_cast_f32_bf16(ctx, dm, dmb, (int64_t)B * d.SIXD, stream));
    float* g_ada_w = (float*)gr->ada_w + (int64_t)i * d.SIXD * D;
    const char* ada_w = (const char*)w->ada_w + (int64_t)i * d.SIXD * D * 2;
    JAT_TRY(wgrad(ctx, dmb, d.SIXD, ws->t_act, D, g_ada_w, d.SIXD, D, B, stream));
    JAT_TRY(jat_colsum_bf16(ctx, dmb, d.SIXD, B, d.SIXD, (float*)gr->ada_b + (int64_t)i * d.SIXD, stream));
    {   // d t_act (f32, summed over blocks) += dmod_i Wada_i
        jat_gemm_epilogue e = epi_plain(JAT_EPI_ACCUM, sc->dt_acc, D);
        e.w_transposed = 1; e.k_splits = 16;
        JAT_TRY(jat_gemm_bf16(ctx, dmb, d.SIXD, ada_w, D, B, D, d.SIXD, &e, -1, 0, stream));
    }
    return 0;
}

extern "C" int jat_dit_backward_end(jat_ctx* ctx, const jat_dit_weights* w, const jat_dit_workspace* ws,
                                    const jat_dit_saved* sv, const jat_dit_bwd_scratch* sc, const jat_dit_weights* gr, int B,
                                    int T, void* stream) {
    DeviceGuard dev_guard__(ctx);
    if (!ctx || !w || !ws || !sv || !sc || !gr) return fail(JAT_ERR_BAD_ARG, "jat_dit_backward_end: null argument");
    const BwdDims d = bwd_dims(w, B, T);
    const int D = d.D, BD = d.BD, M = d.M;
    // ---- patch embed: x0 = gelu(patches W1^T + b1) W2^T + b2   (no gradient flows to the inputs)
    JAT_TRY(jat_cast_f32_bf16(ctx, sc->dx, sc->dy, d.MD, stream));
    JAT_TRY(wgrad(ctx, sc->dy, D, ws->pe_hid, BD, (float*)gr->pe_w2, D, BD, M, stream));
    JAT_TRY(jat_colsum_bf16(ctx, sc->dy, D, M, D, (float*)gr->pe_b2, stream));
    JAT_TRY(dgrad(ctx, sc->dy, D, w->pe_w2, BD, D, M, sc->dpe, JAT_ACT_GELU_ERF, sv->pe_u, stream));
    JAT_TRY(wgrad(ctx, sc->dpe, BD, ws->patches, d.KIN, (float*)gr->pe_w1, BD, d.KIN, M, stream));
    JAT_TRY(jat_colsum_bf16(ctx, sc->dpe, BD, M, BD, (float*)gr->pe_b1, stream));
    // ---- timestep path: t_act = silu(t_emb), t_emb = silu(feat W1^T + b1) W2^T + b2
    pre_launch(ctx, TAG_COLSUM, (cudaStream_t)stream);
    dact_mul_kernel<ACT_SILU><<<(B * D + 255) / 256, 256, 0, (cudaStream_t)stream>>>(sc->dt_acc, (const __nv_bfloat16*)sv->t_u2,
                                                                                   (__nv_bfloat16*)sc->dt_a, (long long)B * D);
    JAT_TRY(post_launch(ctx, "dact_mul"));
    JAT_TRY(wgrad(ctx, sc->dt_a, D, ws->t_hid, D, (float*)gr->te_w2, D, D, B, stream));
    JAT_TRY(jat_colsum_bf16(ctx, sc->dt_a, D, B, D, (float*)gr->te_b2, stream));
    JAT_TRY(dgrad(ctx, sc->dt_a, D, w->te_w2, D, D, B, sc->dt_b, JAT_ACT_SILU, sv->t_u1, stream));
    JAT_TRY(wgrad(ctx, sc->dt_b, D, ws->t_feat, D, (float*)gr->te_w1, D, D, B, stream));
    return jat_colsum_bf16(ctx, sc->dt_b, D, B, D, (float*)gr->te_b1, stream);
}

extern "C" int jat_dit_backward(jat_ctx* ctx, const jat_dit_weights* w, const jat_dit_workspace* ws, const jat_dit_saved* sv,
                                const jat_dit_bwd_scratch* sc, const jat_dit_weights* gr, const float* d_out, int B, int T,
                                void* stream) {
    DeviceGuard dev_guard__(ctx);
    JAT_TRY(jat_dit_backward_begin(ctx, w, ws, sv, sc, gr, d_out, B, T, stream));
    for (int i = w->depth - 1; i >= 0; --i) JAT_TRY(jat_dit_backward_block(ctx, w, ws, sv, sc, gr, i, B, T, stream));
    return jat_dit_backward_end(ctx, w, ws, sv, sc, gr, B, T, stream);
}
