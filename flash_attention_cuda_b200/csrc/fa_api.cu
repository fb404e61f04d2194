// fa_api.cu -- host launcher behind the C ABI in include/flash_attn.h.
//
// Replaces the reference's flash_attention_v9_dispatch (flash_attention.cu:606-663): argument
// validation (the reference has none), TMA descriptors for Q/K/V, one persistent launch.
// The reference's four-tier (causal x seq>=2048) template dispatch (FA.cu:620-661) collapses to
// one kernel per head_dim: tile shapes are fixed by the tcgen05 instruction shape (M=128) and the
// work loop adapts to the sequence length at run time.
#include "../../include/flash_attn.h"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "fa_fwd_sm100.cuh"        // default kernel: one CTA per work item, two Q tiles per CTA
#include "fa_fwd_pair_sm100.cuh"   // experimental CTA-pair kernel (cta_group::2), opt-in

namespace {

constexpr int kMaxDevices = 64;
constexpr int kSchedSlots = 4096;
constexpr int kHostChunks = 32;    // most head chunks flash_attn_fwd_host can pipeline over PCIe (events are per chunk)
constexpr int kHostChunksDefault = 8;   // FLASH_ATTN_B200_HOST_CHUNKS overrides (A/B runs)
constexpr int kGroupMB = 32;       // K+V bytes of one scheduling group of heads (make_params)
// exp2 on the FMA pipe (fa::poly_pair): share of element pairs for D = 128 with >= 32 KV tiles / for D = 64.
// Build-time so that A/B variants are one -D away; the defaults are the measured optimum (profiles/).
#ifndef FA_POLY_LONG
#define FA_POLY_LONG 1
#endif
#ifndef FA_POLY_D64
#define FA_POLY_D64 0
#endif
constexpr int kPolyLong = FA_POLY_LONG, kPolyD64 = FA_POLY_D64;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

std::atomic<unsigned long long> g_launches{0};
std::atomic<int> g_sm_margin{0};     // SMs the persistent grid leaves free (flash_attn_set_sm_margin)
std::once_flag g_encode_once;
EncodeTiledFn g_encode = nullptr;

struct DeviceState {
    std::once_flag once;
    int ok = 0;           // cudaSuccess when attributes are set
    int num_sms = 0;
    int cc_major = 0;
    // dynamic tile scheduler state: kSchedSlots x {next, done} ints, zero when idle; a launch uses
    // slot (sequence number % kSchedSlots) and its last CTA re-zeroes it
    int* sched = nullptr;
    std::atomic<unsigned> sched_seq{0};
    // staging buffers for flash_attn_fwd_host
    std::mutex host_mu;
    void* stage = nullptr;
    size_t stage_bytes = 0;
    // three streams (H2D, kernels, D2H) and per-chunk events: the host entry point pipelines head chunks
    cudaStream_t host_stream = nullptr, host_in = nullptr, host_out = nullptr;
    cudaEvent_t ev_in[kHostChunks] = {}, ev_k[kHostChunks] = {};
    bool host_ready = false;
};
DeviceState g_dev[kMaxDevices];

void load_encode_fn() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        g_encode = reinterpret_cast<EncodeTiledFn>(fn);
}

template <int D, int kPoly, bool kBF16 = false>
int set_kernel_attrs() {
    return (int)cudaFuncSetAttribute(fa::fa_fwd_kernel<D, kPoly, kBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     fa::Cfg<D>::kSmemBytes);
}
// exp2 on the FMA pipe for 1 pair in 4 only where it pays: D = 128 and at least 32 KV tiles (fa_fwd_sm100.cuh)
bool use_poly(int D, int Nkv) { return D == 128 && Nkv >= 4096; }
template <int D, int CG>
int set_pair_kernel_attrs() {
    return (int)cudaFuncSetAttribute(fa_pair::fa_fwd_kernel<D, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     fa_pair::Cfg<D, CG>::kSmemBytes);
}

// FLASH_ATTN_B200_KERNEL=pair selects the experimental CTA-pair kernel for the whole
// process; anything else runs the default kernel.  Read once.
bool use_pair_kernel() {
    static const bool pair = []() {
        const char* e = getenv("FLASH_ATTN_B200_KERNEL");
        return e && strcmp(e, "pair") == 0;
    }();
    return pair;
}
// CTAs per work unit of the pair kernel: D = 128 runs as CTA pairs (tcgen05.mma.cta_group::2, M = 256);
// D = 64 keeps one CTA per unit (its V half would be narrower than a 128-byte swizzle panel).
// FLASH_ATTN_B200_CG=1 forces single CTAs for D = 128 as well (A/B measurements).
int pair_cta_group_for(int D) {
    static const int forced = []() {
        const char* e = getenv("FLASH_ATTN_B200_CG");
        return (e && e[0] == '1') ? 1 : 0;
    }();
    return (D == 64 || forced == 1) ? 1 : 2;
}

// Per-device one-time setup.  Re-entrant from several host threads (one per GPU in the
// multi-GPU harness): everything is guarded by the device's once_flag.
DeviceState* device_state(int* err) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { *err = (int)e; return nullptr; }
    if (dev < 0 || dev >= kMaxDevices) { *err = FA_ERR_UNSUPPORTED_ARCH; return nullptr; }
    DeviceState* st = &g_dev[dev];
    std::call_once(st->once, [st, dev]() {
        cudaDeviceProp prop;
        cudaError_t e2 = cudaGetDeviceProperties(&prop, dev);
        if (e2 != cudaSuccess) { st->ok = (int)e2; return; }
        st->num_sms = prop.multiProcessorCount;
        st->cc_major = prop.major;
        if (prop.major != 10) { st->ok = FA_ERR_UNSUPPORTED_ARCH; return; }
        int r = set_kernel_attrs<128, kPolyLong>();
        if (r == 0) r = set_kernel_attrs<128, 0>();
        if (r == 0) r = set_kernel_attrs<64, kPolyD64>();
        if (r == 0) r = set_kernel_attrs<128, kPolyLong, true>();
        if (r == 0) r = set_kernel_attrs<128, 0, true>();
        if (r == 0) r = set_kernel_attrs<64, kPolyD64, true>();
        if (r == 0 && use_pair_kernel()) {
            r = set_pair_kernel_attrs<128, 1>();
            if (r == 0) r = set_pair_kernel_attrs<128, 2>();
            if (r == 0) r = set_pair_kernel_attrs<64, 1>();
        }
        if (r == 0) r = (int)cudaMalloc(&st->sched, kSchedSlots * 2 * sizeof(int));
        if (r == 0) r = (int)cudaMemset(st->sched, 0, kSchedSlots * 2 * sizeof(int));
        st->ok = r;
    });
    if (st->ok != 0) { *err = st->ok; return nullptr; }
    return st;
}

// [BH, N, D] fp16, box = 64 halves x `rows` rows x 1 head, 128-byte swizzle; rows past N read as zero
// (and are dropped on a TMA store).
int make_tmap(CUtensorMap* tm, const void* base, int BH, int N, int D, int rows = fa::kBlockN, bool bf16 = false) {
    std::call_once(g_encode_once, load_encode_fn);
    if (!g_encode) return FA_ERR_TENSORMAP;
    cuuint64_t gdim[3] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)BH};
    cuuint64_t gstride[2] = {(cuuint64_t)D * 2, (cuuint64_t)N * D * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? FA_OK : FA_ERR_TENSORMAP;
}

int validate(const void* q, const void* k, const void* v, const void* o, int B, int H, int Nq, int Nkv, int D) {
    if (D != 64 && D != 128) return FA_ERR_BAD_HEAD_DIM;
    if (!q || !k || !v || !o) return FA_ERR_NULL_PTR;
    if (B < 1 || H < 1 || Nq < 1 || Nkv < 1) return FA_ERR_BAD_SHAPE;
    if ((long long)B * H > 0x7fffffffLL) return FA_ERR_BAD_SHAPE;
    if ((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o) & 15u) != 0) return FA_ERR_MISALIGNED;
    return FA_OK;
}

fa::Params make_params(int BH, int Nq, int Nkv, int D, int causal, long long shift) {
    fa::Params p;
    memset(&p, 0, sizeof p);
    p.Nq = Nq; p.Nkv = Nkv; p.BH = BH;
    p.causal = causal ? 1 : 0;
    if (shift > 0x3fffffffLL) shift = 0x3fffffffLL;
    if (shift < -0x3fffffffLL) shift = -0x3fffffffLL;
    p.shift = (int)shift;
    // experimental: one Q tile per work item (shorter dependency chains for grids that cannot fill the machine);
    // only in -DFA_SINGLE_TILE_MODE builds and with FLASH_ATTN_B200_ITEM_TILES=1 -- not yet measured on a GPU
#ifdef FA_SINGLE_TILE_MODE
    static const int single = [] {
        const char* e = getenv("FLASH_ATTN_B200_ITEM_TILES");
        return e && atoi(e) == 1 ? 1 : 0;
    }();
#else
    const int single = 0;
#endif
#ifdef FA_SINGLE_TILE_MODE
    p.single = single;
#endif
    p.nqp = single ? (Nq + fa::kBlockM - 1) / fa::kBlockM : (Nq + 2 * fa::kBlockM - 1) / (2 * fa::kBlockM);
    const long long tw = (long long)BH * p.nqp;
    p.total_work = (int)tw;
    // heads per scheduling group: K+V of the group <= kGroupMB (B200's 126 MB L2 is two 63 MB halves, and a line
    // read from the far half is also kept in the near one); FLASH_ATTN_B200_L2_GROUP_MB overrides for A/B runs
    static const long long group_mb = [] {
        const char* e = getenv("FLASH_ATTN_B200_L2_GROUP_MB");
        const long long v = e ? atoll(e) : 0;
        return v > 0 ? v : (long long)kGroupMB;
    }();
    const long long kv_bytes = 2LL * Nkv * D * 2;
    long long gh = (group_mb << 20) / (kv_bytes > 0 ? kv_bytes : 1);
    if (gh < 1) gh = 1;
    if (gh > BH) gh = BH;
    // equal groups: a short last group would start its heaviest items when the launch is almost over
    const long long n_groups = (BH + gh - 1) / gh;
    gh = (BH + n_groups - 1) / n_groups;
    p.group_heads = (int)gh;
    p.scale = 1.0f / sqrtf((float)D);             // FA.cu:612
    p.scale_log2 = p.scale * 1.4426950408889634f;
    return p;
}

template <int D, int kPoly, bool kBF16 = false>
int launch(DeviceState* st, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
           const CUtensorMap& to, fa::Params p, cudaStream_t stream) {
    int avail = st->num_sms - g_sm_margin.load(std::memory_order_relaxed);
    if (avail < 1) avail = 1;
    int grid = p.total_work < avail ? p.total_work : avail;
    if (grid < 1) grid = 1;
    p.sched = st->sched + 2 * (st->sched_seq.fetch_add(1, std::memory_order_relaxed) % kSchedSlots);
    // launched with programmatic stream serialization (PDL): the kernel's prologue overlaps the tail
    // of its predecessor in the stream; it executes griddepcontrol.wait before touching global memory
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(fa::kNumThreads);
    cfg.dynamicSmemBytes = fa::Cfg<D>::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, fa::fa_fwd_kernel<D, kPoly, kBF16>, tq, tk, tv, to, p);
    if (le != cudaSuccess) return (int)le;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();   // FA.cu:662
}

// ---- experimental CTA-pair kernel (fa_fwd_pair_sm100.cuh) ----
fa_pair::Params make_pair_params(const fa::Params& b, int D) {
    fa_pair::Params p;
    memset(&p, 0, sizeof p);
    p.o = b.o; p.o_partial = b.o_partial; p.ml = b.ml;
    p.Nq = b.Nq; p.Nkv = b.Nkv; p.BH = b.BH;
    p.causal = b.causal; p.shift = b.shift;
    p.cg = pair_cta_group_for(D);
    p.nqu = (b.Nq + p.cg * fa_pair::kBlockM - 1) / (p.cg * fa_pair::kBlockM);
    p.total_work = (int)((long long)b.BH * p.nqu);
    p.group_heads = b.group_heads;
    p.partial_mode = b.partial_mode; p.accumulate = b.accumulate;
    p.scale = b.scale; p.scale_log2 = b.scale_log2;
    return p;
}

template <int D, int CG>
int launch_pair(DeviceState* st, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                const CUtensorMap& to, fa_pair::Params p, cudaStream_t stream) {
    int avail = (st->num_sms - g_sm_margin.load(std::memory_order_relaxed)) / CG;
    if (avail < 1) avail = 1;
    int units = p.total_work < avail ? p.total_work : avail;
    if (units < 1) units = 1;
    p.sched = st->sched + 2 * (st->sched_seq.fetch_add(1, std::memory_order_relaxed) % kSchedSlots);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)(units * CG));
    cfg.blockDim = dim3(fa_pair::kNumThreads);
    cfg.dynamicSmemBytes = fa_pair::Cfg<D, CG>::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = CG;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CG == 2 ? 2 : 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, fa_pair::fa_fwd_kernel<D, CG>, tq, tk, tv, to, p);
    if (le != cudaSuccess) return (int)le;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int run_pair(DeviceState* st, const void* q, const void* k, const void* v, const fa::Params& base, int D,
             cudaStream_t stream) {
    fa_pair::Params p = make_pair_params(base, D);
    if ((long long)p.BH * p.nqu > 0x7fffffffLL) return FA_ERR_BAD_SHAPE;
    CUtensorMap tq, tk, tv, to;
    int rc;
    if ((rc = make_tmap(&tq, q, p.BH, p.Nq, D)) != FA_OK) return rc;
    if ((rc = make_tmap(&tk, k, p.BH, p.Nkv, D, fa_pair::kBlockN / p.cg)) != FA_OK) return rc;
    if ((rc = make_tmap(&tv, v, p.BH, p.Nkv, D)) != FA_OK) return rc;
    // O store map (unused in partial mode: describe Q's extent on a valid pointer)
    if ((rc = make_tmap(&to, p.o ? (const void*)p.o : q, p.BH, p.Nq, D)) != FA_OK) return rc;
    if (D == 64) return launch_pair<64, 1>(st, tq, tk, tv, to, p, stream);
    return p.cg == 2 ? launch_pair<128, 2>(st, tq, tk, tv, to, p, stream)
                     : launch_pair<128, 1>(st, tq, tk, tv, to, p, stream);
}

int run(const void* q, const void* k, const void* v, fa::Params& p, int D, cudaStream_t stream, bool bf16 = false) {
    int err = 0;
    DeviceState* st = device_state(&err);
    if (!st) return err;
    if (use_pair_kernel() && !bf16) return run_pair(st, q, k, v, p, D, stream);   // the experimental kernel is FP16 only
    if ((long long)p.BH * p.nqp > 0x7fffffffLL) return FA_ERR_BAD_SHAPE;
    CUtensorMap tq, tk, tv, to;
    int rc;
    if ((rc = make_tmap(&tq, q, p.BH, p.Nq, D, fa::kBlockN, bf16)) != FA_OK) return rc;
    if ((rc = make_tmap(&tk, k, p.BH, p.Nkv, D, fa::kBlockN, bf16)) != FA_OK) return rc;
    if ((rc = make_tmap(&tv, v, p.BH, p.Nkv, D, fa::kBlockN, bf16)) != FA_OK) return rc;
    // O store map (unused in partial mode: describe Q's extent on a valid pointer)
    if ((rc = make_tmap(&to, p.o ? (const void*)p.o : q, p.BH, p.Nq, D, fa::kBlockN, bf16)) != FA_OK) return rc;
    if (bf16) {
        if (D == 64) return launch<64, kPolyD64, true>(st, tq, tk, tv, to, p, stream);
        return use_poly(D, p.Nkv) ? launch<128, kPolyLong, true>(st, tq, tk, tv, to, p, stream)
                                  : launch<128, 0, true>(st, tq, tk, tv, to, p, stream);
    }
    if (D == 64) return launch<64, kPolyD64>(st, tq, tk, tv, to, p, stream);
    return use_poly(D, p.Nkv) ? launch<128, kPolyLong>(st, tq, tk, tv, to, p, stream) : launch<128, 0>(st, tq, tk, tv, to, p, stream);
}

}  // namespace

extern "C" {

int flash_attn_fwd(const void* q, const void* k, const void* v, void* o, int B, int H, int N, int D, int causal,
                   void* stream) {
    int rc = validate(q, k, v, o, B, H, N, N, D);
    if (rc != FA_OK) return rc;
    fa::Params p = make_params(B * H, N, N, D, causal, 0);
    p.o = static_cast<__half*>(o);
    return run(q, k, v, p, D, static_cast<cudaStream_t>(stream));
}

int flash_attn_fwd_bf16(const void* q, const void* k, const void* v, void* o, int B, int H, int N, int D, int causal,
                        void* stream) {
    int rc = validate(q, k, v, o, B, H, N, N, D);
    if (rc != FA_OK) return rc;
    fa::Params p = make_params(B * H, N, N, D, causal, 0);
    p.o = static_cast<__half*>(o);       // 16-bit elements either way; the kernel instantiation decides the format
    return run(q, k, v, p, D, static_cast<cudaStream_t>(stream), /*bf16=*/true);
}

int flash_attn_fwd_ex(const void* q, const void* k, const void* v, float* o_partial, float* ml, int B, int H,
                      int Nq, int Nkv, int D, int causal, long long q_offset, long long kv_offset, int accumulate,
                      void* stream) {
    int rc = validate(q, k, v, o_partial, B, H, Nq, Nkv, D);
    if (rc != FA_OK) return rc;
    if (!ml) return FA_ERR_NULL_PTR;
    fa::Params p = make_params(B * H, Nq, Nkv, D, causal, q_offset - kv_offset);
    p.o_partial = o_partial;
    p.ml = ml;
    p.partial_mode = 1;
    p.accumulate = accumulate ? 1 : 0;
    return run(q, k, v, p, D, static_cast<cudaStream_t>(stream));
}

int flash_attn_finalize(const float* o_partial, const float* ml, void* o, long long rows, int D, void* stream) {
    if (!o_partial || !ml || !o) return FA_ERR_NULL_PTR;
    if (D != 64 && D != 128) return FA_ERR_BAD_HEAD_DIM;
    if (rows < 1) return FA_ERR_BAD_SHAPE;
    const long long threads = rows * (D / 4);
    const int block = 256;
    const long long grid = (threads + block - 1) / block;
    if (grid > 0x7fffffffLL) return FA_ERR_BAD_SHAPE;
    fa::fa_finalize_kernel<<<(unsigned)grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
        o_partial, ml, static_cast<__half*>(o), rows, D);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int flash_attn_merge(const float* o_partial, const float* ml, void* o, int splits, long long rows, int D,
                     void* stream) {
    if (!o_partial || !ml || !o) return FA_ERR_NULL_PTR;
    if (D != 64 && D != 128) return FA_ERR_BAD_HEAD_DIM;
    if (rows < 1 || splits < 1) return FA_ERR_BAD_SHAPE;
    const long long threads = rows * (D / 4);
    const int block = 256;
    const long long grid = (threads + block - 1) / block;
    if (grid > 0x7fffffffLL) return FA_ERR_BAD_SHAPE;
    fa::fa_merge_kernel<<<(unsigned)grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
        o_partial, ml, static_cast<__half*>(o), rows, D, splits);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int flash_attn_fwd_host(const void* hq, const void* hk, const void* hv, void* ho, int B, int H, int N, int D,
                        int causal) {
    if (!hq || !hk || !hv || !ho) return FA_ERR_NULL_PTR;
    if (D != 64 && D != 128) return FA_ERR_BAD_HEAD_DIM;
    if (B < 1 || H < 1 || N < 1) return FA_ERR_BAD_SHAPE;
    int err = 0;
    DeviceState* st = device_state(&err);
    if (!st) return err;
    const size_t bytes = (size_t)B * H * N * D * sizeof(__half);
    std::lock_guard<std::mutex> lock(st->host_mu);
    cudaError_t e;
    if (!st->host_ready) {
        if ((e = cudaStreamCreateWithFlags(&st->host_stream, cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
        if ((e = cudaStreamCreateWithFlags(&st->host_in, cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
        if ((e = cudaStreamCreateWithFlags(&st->host_out, cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
        for (int i = 0; i < kHostChunks; i++) {
            if ((e = cudaEventCreateWithFlags(&st->ev_in[i], cudaEventDisableTiming)) != cudaSuccess) return (int)e;
            if ((e = cudaEventCreateWithFlags(&st->ev_k[i], cudaEventDisableTiming)) != cudaSuccess) return (int)e;
        }
        st->host_ready = true;
    }
    if (st->stage_bytes < 4 * bytes) {
        if (st->stage) cudaFree(st->stage);
        st->stage = nullptr;
        st->stage_bytes = 0;
        if ((e = cudaMalloc(&st->stage, 4 * bytes)) != cudaSuccess) return (int)e;
        st->stage_bytes = 4 * bytes;
    }
    char* base = static_cast<char*>(st->stage);
    char *dq = base, *dk = base + bytes, *dv = base + 2 * bytes, *dout = base + 3 * bytes;
    // The reference copies everything in, dispatches, copies everything out (FA.cu:774-780).  Heads are
    // independent, so the same work is cut into head chunks and pipelined: while chunk c computes,
    // chunk c+1 is on its way in and chunk c-1 on its way out (PCIe is full duplex; three streams,
    // one event pair per chunk).  The wire time of Q, K, V dominates; kernels and O hide under it.
    const int BH = B * H;
    static const int want_chunks = [] {
        const char* e = getenv("FLASH_ATTN_B200_HOST_CHUNKS");
        const int v = e ? atoi(e) : 0;
        return v >= 1 && v <= kHostChunks ? v : kHostChunksDefault;
    }();
    const int chunks = BH < want_chunks ? BH : want_chunks;
    const size_t head_bytes = (size_t)N * D * sizeof(__half);
    int h0 = 0;
    for (int c = 0; c < chunks; c++) {
        const int h1 = (int)((long long)BH * (c + 1) / chunks);
        const int nh = h1 - h0;
        const size_t off = (size_t)h0 * head_bytes, len = (size_t)nh * head_bytes;
        const char *hq8 = static_cast<const char*>(hq), *hk8 = static_cast<const char*>(hk),
                   *hv8 = static_cast<const char*>(hv);
        if ((e = cudaMemcpyAsync(dq + off, hq8 + off, len, cudaMemcpyHostToDevice, st->host_in)) != cudaSuccess) return (int)e;
        if ((e = cudaMemcpyAsync(dk + off, hk8 + off, len, cudaMemcpyHostToDevice, st->host_in)) != cudaSuccess) return (int)e;
        if ((e = cudaMemcpyAsync(dv + off, hv8 + off, len, cudaMemcpyHostToDevice, st->host_in)) != cudaSuccess) return (int)e;
        if ((e = cudaEventRecord(st->ev_in[c], st->host_in)) != cudaSuccess) return (int)e;
        if ((e = cudaStreamWaitEvent(st->host_stream, st->ev_in[c], 0)) != cudaSuccess) return (int)e;
        int rc = flash_attn_fwd(dq + off, dk + off, dv + off, dout + off, 1, nh, N, D, causal, st->host_stream);   // FA.cu:777
        if (rc != FA_OK) return rc;
        if ((e = cudaEventRecord(st->ev_k[c], st->host_stream)) != cudaSuccess) return (int)e;
        if ((e = cudaStreamWaitEvent(st->host_out, st->ev_k[c], 0)) != cudaSuccess) return (int)e;
        if ((e = cudaMemcpyAsync(static_cast<char*>(ho) + off, dout + off, len, cudaMemcpyDeviceToHost, st->host_out)) !=
            cudaSuccess)
            return (int)e;
        h0 = h1;
    }
    return (int)cudaStreamSynchronize(st->host_out);
}

int flash_attn_get_kernel_info(int B, int H, int N, int D, int causal, flash_attn_kernel_info* info) {
    (void)causal;
    if (!info) return FA_ERR_NULL_PTR;
    if (D != 64 && D != 128) return FA_ERR_BAD_HEAD_DIM;
    if (B < 1 || H < 1 || N < 1) return FA_ERR_BAD_SHAPE;
    memset(info, 0, sizeof *info);
    cudaFuncAttributes attr;
    cudaError_t e = D == 64            ? cudaFuncGetAttributes(&attr, fa::fa_fwd_kernel<64, kPolyD64>)
                    : use_poly(D, N) ? cudaFuncGetAttributes(&attr, fa::fa_fwd_kernel<128, kPolyLong>)
                                     : cudaFuncGetAttributes(&attr, fa::fa_fwd_kernel<128, 0>);
    if (e != cudaSuccess) return (int)e;
    int err = 0;
    DeviceState* st = device_state(&err);
    if (!st) return err;
    fa::Params p = make_params(B * H, N, N, D, causal, 0);
    info->cta_group = 1;
    if (use_pair_kernel()) {
        const int cg = pair_cta_group_for(D);
        e = D == 64    ? cudaFuncGetAttributes(&attr, fa_pair::fa_fwd_kernel<64, 1>)
            : cg == 2 ? cudaFuncGetAttributes(&attr, fa_pair::fa_fwd_kernel<128, 2>)
                      : cudaFuncGetAttributes(&attr, fa_pair::fa_fwd_kernel<128, 1>);
        if (e != cudaSuccess) return (int)e;
        const fa_pair::Params pp = make_pair_params(p, D);
        info->regs_per_thread = attr.numRegs;
        info->local_bytes_per_thread = (int)attr.localSizeBytes;
        info->static_smem_bytes = (int)attr.sharedSizeBytes;
        info->dynamic_smem_bytes = D == 64    ? fa_pair::Cfg<64, 1>::kSmemBytes
                                   : cg == 2 ? fa_pair::Cfg<128, 2>::kSmemBytes
                                             : fa_pair::Cfg<128, 1>::kSmemBytes;
        info->threads_per_cta = fa_pair::kNumThreads;
        int units = pp.total_work < st->num_sms / cg ? pp.total_work : st->num_sms / cg;
        info->ctas = (units < 1 ? 1 : units) * cg;
        info->tmem_columns = fa_pair::kTmemCols;
        info->kv_stages = D == 64    ? fa_pair::Cfg<64, 1>::kKStages + fa_pair::Cfg<64, 1>::kVStages
                          : cg == 2 ? fa_pair::Cfg<128, 2>::kKStages + fa_pair::Cfg<128, 2>::kVStages
                                    : fa_pair::Cfg<128, 1>::kKStages + fa_pair::Cfg<128, 1>::kVStages;
        info->work_items = pp.total_work;
        info->num_sms = st->num_sms;
        info->cta_group = cg;
        return FA_OK;
    }
    info->regs_per_thread = attr.numRegs;
    info->local_bytes_per_thread = (int)attr.localSizeBytes;
    info->static_smem_bytes = (int)attr.sharedSizeBytes;
    info->dynamic_smem_bytes = D == 128 ? fa::Cfg<128>::kSmemBytes : fa::Cfg<64>::kSmemBytes;
    info->threads_per_cta = fa::kNumThreads;
    info->ctas = p.total_work < st->num_sms ? p.total_work : st->num_sms;
    info->tmem_columns = fa::kTmemCols;
    info->kv_stages = D == 128 ? fa::Cfg<128>::kStages : fa::Cfg<64>::kStages;
    info->work_items = p.total_work;
    info->num_sms = st->num_sms;
    return FA_OK;
}

int flash_attn_set_sm_margin(int sms) {
    if (sms < 0) sms = 0;
    return g_sm_margin.exchange(sms, std::memory_order_relaxed);
}

// ---- peer-readable blocks (include/flash_attn.h): cudaMalloc + legacy CUDA IPC.  cudaMalloc rather than a
// caller-provided pointer: a handle of a sub-allocation of somebody's caching allocator would expose the whole
// segment, and cuMemMap-based ("expandable") segments cannot be exported this way at all.
int flash_attn_peer_alloc(size_t bytes, void** ptr, unsigned char* handle) {
    if (!ptr || !handle) return FA_ERR_NULL_PTR;
    if (bytes == 0) return FA_ERR_BAD_SHAPE;
    static_assert(sizeof(cudaIpcMemHandle_t) == FA_PEER_HANDLE_BYTES, "handle size");
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, bytes);
    if (e != cudaSuccess) return (int)e;
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, d);
    if (e != cudaSuccess) { cudaFree(d); return (int)e; }
    memcpy(handle, &h, sizeof h);
    *ptr = d;
    return FA_OK;
}

int flash_attn_peer_open(const unsigned char* handle, void** ptr) {
    if (!ptr || !handle) return FA_ERR_NULL_PTR;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    // maps the owner's allocation into this process and enables peer access to the owner's device
    return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

int flash_attn_peer_close(void* ptr) { return ptr ? (int)cudaIpcCloseMemHandle(ptr) : FA_ERR_NULL_PTR; }
int flash_attn_peer_free(void* ptr) { return ptr ? (int)cudaFree(ptr) : FA_ERR_NULL_PTR; }

int flash_attn_peer_copy(void* dst, const void* src, size_t bytes, void* stream) {
    if (!dst || !src) return FA_ERR_NULL_PTR;
    if (bytes == 0) return FA_OK;
    // unified addressing: the runtime sees that src lives on another device and programs a copy engine
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream);
}

unsigned long long flash_attn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void flash_attn_destroy(void) {
    int saved = 0;
    if (cudaGetDevice(&saved) != cudaSuccess) return;
    for (int d = 0; d < kMaxDevices; d++) {
        DeviceState* st = &g_dev[d];
        std::lock_guard<std::mutex> lock(st->host_mu);
        if (st->stage || st->host_ready) {
            cudaSetDevice(d);
            if (st->stage) cudaFree(st->stage);
            if (st->host_stream) cudaStreamDestroy(st->host_stream);
            if (st->host_in) cudaStreamDestroy(st->host_in);
            if (st->host_out) cudaStreamDestroy(st->host_out);
            for (int i = 0; i < kHostChunks; i++) {
                if (st->ev_in[i]) cudaEventDestroy(st->ev_in[i]);
                if (st->ev_k[i]) cudaEventDestroy(st->ev_k[i]);
                st->ev_in[i] = st->ev_k[i] = nullptr;
            }
            st->stage = nullptr;
            st->stage_bytes = 0;
            st->host_stream = st->host_in = st->host_out = nullptr;
            st->host_ready = false;
        }
    }
    cudaSetDevice(saved);
}

const char* flash_attn_error_string(int code) {
    switch (code) {
        case FA_OK: return "success";
        case FA_ERR_BAD_HEAD_DIM: return "head_dim must be 64 or 128";
        case FA_ERR_NULL_PTR: return "null pointer argument";
        case FA_ERR_MISALIGNED: return "tensor base pointers must be 16-byte aligned";
        case FA_ERR_BAD_SHAPE: return "invalid shape (B, H, N must be >= 1 and within tensor-map limits)";
        case FA_ERR_UNSUPPORTED_ARCH: return "device is not compute capability 10.x (B200, sm_100a)";
        case FA_ERR_TENSORMAP: return "cuTensorMapEncodeTiled failed or is unavailable";
        case FA_ERR_WORKSPACE: return "workspace missing or too small";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

const char* flash_attn_version(void) { return use_pair_kernel() ? "flashattn_b200 0.2 (sm_100a tcgen05/TMA forward, experimental CTA-pair kernel)"
                             : "flashattn_b200 0.2 (sm_100a tcgen05/TMA forward)"; }

}  // extern "C"

// Watchdog record of the current device: {aborted, barrier tag, block, thread} (see sm100_ptx.cuh).
// Synchronises the device.  Returns FA_OK or a cudaError_t.
extern "C" int flash_attn_debug_status(unsigned int* out4) {
    if (!out4) return FA_ERR_NULL_PTR;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return (int)e;
    return (int)cudaMemcpyFromSymbol(out4, sm100::g_watchdog, 4 * sizeof(unsigned int));
}

#ifdef FA_TIMING
// debug builds only (-DFA_TIMING): in-kernel clock64 probes, see tests/harness/timing.py
extern "C" int flash_attn_debug_timing(unsigned long long* out32, int reset) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return (int)e;
    unsigned long long z[64] = {0};
    if (use_pair_kernel()) {   // 64 counters
        e = cudaMemcpyFromSymbol(out32, fa_pair::g_timing_pair, 64 * sizeof(unsigned long long));
        if (e == cudaSuccess && reset) e = cudaMemcpyToSymbol(fa_pair::g_timing_pair, z, 64 * sizeof(unsigned long long));
        return (int)e;
    }
    e = cudaMemcpyFromSymbol(out32, fa::g_timing, 32 * sizeof(unsigned long long));
    if (e == cudaSuccess && reset) e = cudaMemcpyToSymbol(fa::g_timing, z, 32 * sizeof(unsigned long long));
    return (int)e;
}
#endif

// Host-side mirror of the device work decomposition, exported for the scheduler tests
// (every (bh, q-tile) exactly once, heavy-first, masked tiles skipped).
extern "C" int flash_attn_debug_work_item(int w, int B, int H, int Nq, int Nkv, int D, int causal, long long shift,
                                          int* total, int* bh, int* q0, int* n0, int* n1) {
    if (D != 64 && D != 128) return FA_ERR_BAD_HEAD_DIM;
    fa::Params p = make_params(B * H, Nq, Nkv, D, causal, shift);
    if (use_pair_kernel()) {
        const fa_pair::Params pp = make_pair_params(p, D);
        *total = pp.total_work;
        if (w < 0 || w >= pp.total_work) return FA_ERR_BAD_SHAPE;
        fa_pair::WorkItem it = fa_pair::decode_work(w, pp);
        *bh = it.bh; *q0 = it.q0; *n0 = it.n0; *n1 = it.n1;
        return FA_OK;
    }
    *total = p.total_work;
    if (w < 0 || w >= p.total_work) return FA_ERR_BAD_SHAPE;
    fa::WorkItem it = fa::decode_work(w, p);
    *bh = it.bh; *q0 = it.q0; *n0 = it.n0; *n1 = it.n1;
    return FA_OK;
}
// 128-row Q tiles a work item covers: 2 for the default kernel (one CTA, two tiles); for the pair
// kernel the CTAs per unit (1 or 2), each holding one tile
extern "C" int flash_attn_debug_tiles_per_item(int D) {
    if (D != 64 && D != 128) return FA_ERR_BAD_HEAD_DIM;
    if (use_pair_kernel()) return pair_cta_group_for(D);
#ifdef FA_SINGLE_TILE_MODE
    return make_params(1, 1, 1, D, 0, 0).single ? 1 : 2;
#else
    return 2;
#endif
}
