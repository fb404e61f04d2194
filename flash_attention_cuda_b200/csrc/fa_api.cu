// fa_api.cu -- host launcher behind the C ABI in include/flash_attn.h.
//
// Replaces the reference's flash_attention_v9_dispatch (flash_attention.cu:606-663): argument
// validation (the reference has none), TMA descriptors for Q/K/V, one persistent launch.
// The reference's four-tier (causal x seq>=2048) template dispatch (FA.cu:620-661) collapses to
// one kernel per head_dim: tile shapes are fixed by the tcgen05 instruction shape (M=128) and the
// work loop adapts to the sequence length at run time.
#include "../../include/flash_attn.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "fa_fwd_sm100.cuh"        // the kernel: one CTA per work item, two Q tiles per CTA

namespace {

constexpr int kMaxDevices = 64;
constexpr int kSchedSlots = 4096;
constexpr int kHostChunks = 32;    // most head chunks flash_attn_fwd_host can pipeline over PCIe (events are per chunk)
constexpr int kHostChunksDefault = 8;   // FLASH_ATTN_B200_HOST_CHUNKS overrides (A/B runs)
#ifndef FA_HOST_FIRST_DEFAULT
#define FA_HOST_FIRST_DEFAULT 0
#endif
#ifndef FA_HOST_CTAS_DEFAULT
#define FA_HOST_CTAS_DEFAULT 0
#endif
#ifndef FA_HOST_ZEROCOPY_DEFAULT
#define FA_HOST_ZEROCOPY_DEFAULT 1
#endif
constexpr int kGroupMB = 32;       // K+V bytes of one scheduling group of heads (make_params)
// exp2 on the FMA pipe (fa::poly_pair): share of element pairs for D = 128 with >= 32 KV tiles / for D = 64.
// Build-time so that A/B variants are one -D away; the defaults are the measured optimum (profiles/).
#ifndef FA_POLY_LONG
#define FA_POLY_LONG 1
#endif
#ifndef FA_POLY_D64
#define FA_POLY_D64 1
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

constexpr int kCaptureSlots = 1024;  // scheduler words reserved for launches recorded into CUDA graphs (one each, never reused)
constexpr int kTmapCache = 16;       // cached descriptor sets per device (flash_attn_fwd re-encodes nothing for a repeated call;
                                     // flash_attn_fwd_host alone cycles through 8 of them, one per head chunk)

struct TmapSet {
    const void *q = nullptr, *k = nullptr, *v = nullptr, *o = nullptr;
    int BH = 0, Nq = 0, Nkv = 0, D = 0, bf16 = 0;
    long long kv_head_rows = 0;
    unsigned long long stamp = 0;    // 0 = empty; otherwise the LRU clock of its last use
    CUtensorMap tq, tk, tv, to;
};

struct DeviceState {
    std::mutex init_mu;
    bool inited = false;
    int ok = 0;           // cudaSuccess when attributes are set
    int num_sms = 0;
    int cc_major = 0;
    // dynamic tile scheduler state: {next, done} int pairs, zero when idle.  An eager launch uses pair
    // (sequence number % kSchedSlots) and its last CTA re-zeroes it; a launch recorded into a CUDA graph gets
    // a pair of its own behind those (a replay may run next to eager launches that wrapped around to any
    // of the shared pairs, and next to other graphs)
    int* sched = nullptr;
    std::atomic<unsigned> sched_seq{0};
    std::atomic<int> capture_seq{0};
    // kernel watchdog record {aborted, tag, block, thread} in zero-copy host memory: the kernel writes it, the
    // launcher reads it at the start of every call without synchronising (sm100_ptx.cuh)
    unsigned int* wd_host = nullptr;
    unsigned int wd_last[4] = {0, 0, 0, 0};
    // descriptor cache
    std::mutex tmap_mu;
    TmapSet tmaps[kTmapCache];
    unsigned long long tmap_clock = 0;
    // staging buffers for flash_attn_fwd_host
    std::mutex host_mu;
    void* stage = nullptr;
    size_t stage_bytes = 0;
    // three streams (H2D, kernels, D2H) and per-chunk events: the host entry point pipelines head chunks
    cudaStream_t host_stream = nullptr, host_in = nullptr, host_out = nullptr;
    cudaStream_t host_in2[2] = {nullptr, nullptr};            // K and V travel next to Q on streams of their own
    cudaEvent_t ev_in[kHostChunks] = {}, ev_k[kHostChunks] = {};
    cudaEvent_t ev_in2[2][kHostChunks] = {};
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
#ifndef FA_POLY_MIN_N
#define FA_POLY_MIN_N 4096
#endif
// (without a causal mask it already pays from N = 512 on: +1-3 %, profiles/r02_c6_polyall_ab.log)
bool use_poly(int D, int Nkv, int causal = 1) { return D == 128 && Nkv >= (causal ? FA_POLY_MIN_N : 512); }
// Per-device setup, on the first call on that device (and again after flash_attn_destroy).  Re-entrant from
// several host threads (one per GPU in a multi-GPU harness): guarded by the device's mutex.
DeviceState* device_state(int* err) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { *err = (int)e; return nullptr; }
    if (dev < 0 || dev >= kMaxDevices) { *err = FA_ERR_UNSUPPORTED_ARCH; return nullptr; }
    DeviceState* st = &g_dev[dev];
    std::lock_guard<std::mutex> lock(st->init_mu);
    if (!st->inited) {
        st->inited = true;
        st->ok = [st, dev]() -> int {
            cudaDeviceProp prop;
            cudaError_t e2 = cudaGetDeviceProperties(&prop, dev);
            if (e2 != cudaSuccess) return (int)e2;
            st->num_sms = prop.multiProcessorCount;
            st->cc_major = prop.major;
            if (prop.major != 10) return FA_ERR_UNSUPPORTED_ARCH;
            int r = set_kernel_attrs<128, kPolyLong>();
            if (r == 0) r = set_kernel_attrs<128, 0>();
            if (r == 0) r = set_kernel_attrs<64, kPolyD64>();
            if (r == 0) r = set_kernel_attrs<128, kPolyLong, true>();
            if (r == 0) r = set_kernel_attrs<128, 0, true>();
            if (r == 0) r = set_kernel_attrs<64, kPolyD64, true>();
            const size_t sched_bytes = (size_t)(kSchedSlots + kCaptureSlots) * 2 * sizeof(int);
            if (r == 0) r = (int)cudaMalloc(&st->sched, sched_bytes);
            if (r == 0) r = (int)cudaMemset(st->sched, 0, sched_bytes);
            // watchdog mirror: 4 words of mapped, pinned host memory; the kernel gets its device alias through a symbol
            if (r == 0) r = (int)cudaHostAlloc(reinterpret_cast<void**>(&st->wd_host), 4 * sizeof(unsigned int), cudaHostAllocMapped);
            if (r == 0) {
                memset(st->wd_host, 0, 4 * sizeof(unsigned int));
                unsigned int* dalias = nullptr;
                r = (int)cudaHostGetDevicePointer(reinterpret_cast<void**>(&dalias), st->wd_host, 0);
                if (r == 0) r = (int)cudaMemcpyToSymbol(sm100::g_watchdog_host, &dalias, sizeof dalias);
                const unsigned int z[4] = {0, 0, 0, 0};
                if (r == 0) r = (int)cudaMemcpyToSymbol(sm100::g_watchdog, z, sizeof z);
            }
            return r;
        }();
    }
    if (st->ok != 0) { *err = st->ok; return nullptr; }
    return st;
}

// A kernel whose mbarrier protocol timed out (sm100_ptx.cuh) has drained with garbage results and left a record.
// It is reported ONCE, as FA_ERR_WATCHDOG from the next call on that device, and then cleared, so that one bad launch
// (or a legitimate wait stretched past the timeout by a debugger / time-slicing) does not poison the process.
int take_watchdog(DeviceState* st) {
    if (!st->wd_host || *reinterpret_cast<volatile unsigned int*>(st->wd_host) == 0u) return FA_OK;
    cudaDeviceSynchronize();      // everything queued behind the aborted launch has drained as well
    for (int i = 0; i < 4; i++) st->wd_last[i] = reinterpret_cast<volatile unsigned int*>(st->wd_host)[i];
    const unsigned int z[4] = {0, 0, 0, 0};
    cudaMemcpyToSymbol(sm100::g_watchdog, z, sizeof z);
    memset(st->wd_host, 0, 4 * sizeof(unsigned int));
    // scheduler words of the aborted launches may be mid-count
    cudaMemset(st->sched, 0, (size_t)(kSchedSlots + kCaptureSlots) * 2 * sizeof(int));
    return FA_ERR_WATCHDOG;
}

// [BH, N, D] fp16, box = 64 halves x `rows` rows x 1 head, 128-byte swizzle; rows past N read as zero
// (and are dropped on a TMA store).
int make_tmap(CUtensorMap* tm, const void* base, int BH, int N, int D, int rows = fa::kBlockN, bool bf16 = false,
              long long head_rows = 0) {
    std::call_once(g_encode_once, load_encode_fn);
    if (!g_encode) return FA_ERR_TENSORMAP;
    if (head_rows < N) head_rows = N;         // rows between two heads (gathered K/V: the buffer holds more rows than are used)
    cuuint64_t gdim[3] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)BH};
    cuuint64_t gstride[2] = {(cuuint64_t)D * 2, (cuuint64_t)head_rows * D * 2};
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

// Split mode (fa_fwd_sm100.cuh): one Q tile per work item, its KV tiles alternating between the CTA's two tile slots.
// Used where pair items are too few to keep the machine busy (use_split below).
// FLASH_ATTN_B200_SPLIT = 0 / 1 forces it off / on (A/B runs); flash_attn_debug_set_split overrides at run time.
std::atomic<int> g_split_override{-2};     // -2: read the environment; -1: automatic; 0 / 1: forced
bool use_split(int BH, int Nq, int causal, int num_sms) {
    int o = g_split_override.load(std::memory_order_relaxed);
    if (o == -2) {
        const char* e = getenv("FLASH_ATTN_B200_SPLIT");
        o = e && (e[0] == '0' || e[0] == '1') ? e[0] - '0' : -1;
        g_split_override.store(o, std::memory_order_relaxed);
    }
    if (o >= 0) return o == 1;
    // measured on B200 (profiles/r02_c4_split_ab.log, H = 32, D = 128): with at most one single-tile item per SM split mode
    // wins by 8-14 % (N = 512); up to two per SM it still wins under a causal mask (N = 768: +9 %, N = 1024: +5 % hot /
    // +13 % cold L2), where item lengths differ, and loses 10 % without one; beyond that pair items win (a K/V tile then
    // serves two Q tiles and the machine is full either way)
    const long long items = (long long)BH * ((Nq + fa::kBlockM - 1) / fa::kBlockM);
    return items <= num_sms || (causal && items <= 2LL * num_sms);
}

fa::Params make_params(int BH, int Nq, int Nkv, int D, int causal, long long shift, bool split = false) {
    fa::Params p;
    memset(&p, 0, sizeof p);
    p.Nq = Nq; p.Nkv = Nkv; p.BH = BH;
    p.causal = causal ? 1 : 0;
    if (shift > 0x3fffffffLL) shift = 0x3fffffffLL;
    if (shift < -0x3fffffffLL) shift = -0x3fffffffLL;
    p.shift = (int)shift;
    p.split = split ? 1 : 0;
    p.nqp = split ? (Nq + fa::kBlockM - 1) / fa::kBlockM : (Nq + 2 * fa::kBlockM - 1) / (2 * fa::kBlockM);
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
           const CUtensorMap& to, fa::Params p, cudaStream_t stream, int max_ctas = 0) {
    int avail = st->num_sms - g_sm_margin.load(std::memory_order_relaxed);
    if (max_ctas > 0 && avail > max_ctas) avail = max_ctas;     // a deliberately narrow grid (flash_attn_fwd_host)
    if (avail < 1) avail = 1;
    int grid = p.total_work < avail ? p.total_work : avail;
    if (grid < 1) grid = 1;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess) cap = cudaStreamCaptureStatusNone;
    if (cap == cudaStreamCaptureStatusActive) {
        // the pair is baked into the graph: give it one nobody else will ever count on
        const int c = st->capture_seq.fetch_add(1, std::memory_order_relaxed);
        if (c >= kCaptureSlots) return FA_ERR_WORKSPACE;
        p.sched = st->sched + 2 * (kSchedSlots + c);
    } else {
        p.sched = st->sched + 2 * (st->sched_seq.fetch_add(1, std::memory_order_relaxed) % kSchedSlots);
    }
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

// The four descriptors of a call (Q, K, V loads, O store), from the device's cache when the same pointers and shape
// were seen before: cuTensorMapEncodeTiled x 4 is most of the host time of a launch, and short sequences
// (BASELINE config 1: ~20 us of GPU time) are called in loops on the same buffers.
int get_tmaps(DeviceState* st, const void* q, const void* k, const void* v, const void* o, const fa::Params& p, int D,
              bool bf16, TmapSet* out, long long kv_head_rows = 0) {
    {
        std::lock_guard<std::mutex> lock(st->tmap_mu);
        for (TmapSet& c : st->tmaps)
            if (c.stamp && c.q == q && c.k == k && c.v == v && c.o == o && c.BH == p.BH && c.Nq == p.Nq && c.Nkv == p.Nkv &&
                c.D == D && c.bf16 == (int)bf16 && c.kv_head_rows == kv_head_rows) {
                c.stamp = ++st->tmap_clock;
                *out = c;
                return FA_OK;
            }
    }
    TmapSet t;
    t.q = q; t.k = k; t.v = v; t.o = o;
    t.BH = p.BH; t.Nq = p.Nq; t.Nkv = p.Nkv; t.D = D; t.bf16 = (int)bf16;
    t.kv_head_rows = kv_head_rows;
    int rc;
    if ((rc = make_tmap(&t.tq, q, p.BH, p.Nq, D, fa::kBlockN, bf16)) != FA_OK) return rc;
    if ((rc = make_tmap(&t.tk, k, p.BH, p.Nkv, D, fa::kBlockN, bf16, kv_head_rows)) != FA_OK) return rc;
    if ((rc = make_tmap(&t.tv, v, p.BH, p.Nkv, D, fa::kBlockN, bf16, kv_head_rows)) != FA_OK) return rc;
    // O store map (unused in partial mode: describe Q's extent on a valid pointer)
    if ((rc = make_tmap(&t.to, o ? o : q, p.BH, p.Nq, D, fa::kBlockN, bf16)) != FA_OK) return rc;
    {
        std::lock_guard<std::mutex> lock(st->tmap_mu);
        TmapSet* victim = &st->tmaps[0];
        for (TmapSet& c : st->tmaps)
            if (c.stamp < victim->stamp) victim = &c;
        t.stamp = ++st->tmap_clock;
        *victim = t;
    }
    *out = t;
    return FA_OK;
}

int run(const void* q, const void* k, const void* v, fa::Params& p, int D, cudaStream_t stream, bool bf16 = false,
        long long kv_head_rows = 0, int max_ctas = 0) {
    int err = 0;
    DeviceState* st = device_state(&err);
    if (!st) return err;
    if ((err = take_watchdog(st)) != FA_OK) return err;
    if ((long long)p.BH * p.nqp > 0x7fffffffLL) return FA_ERR_BAD_SHAPE;
    TmapSet t;
    int rc = get_tmaps(st, q, k, v, p.partial_mode ? nullptr : (const void*)p.o, p, D, bf16, &t, kv_head_rows);
    if (rc != FA_OK) return rc;
    const bool poly = use_poly(D, p.Nkv, p.causal);
    if (bf16) {
        if (D == 64) return launch<64, kPolyD64, true>(st, t.tq, t.tk, t.tv, t.to, p, stream, max_ctas);
        return poly ? launch<128, kPolyLong, true>(st, t.tq, t.tk, t.tv, t.to, p, stream)
                    : launch<128, 0, true>(st, t.tq, t.tk, t.tv, t.to, p, stream, max_ctas);
    }
    if (D == 64) return launch<64, kPolyD64>(st, t.tq, t.tk, t.tv, t.to, p, stream, max_ctas);
    return poly ? launch<128, kPolyLong>(st, t.tq, t.tk, t.tv, t.to, p, stream)
                : launch<128, 0>(st, t.tq, t.tk, t.tv, t.to, p, stream, max_ctas);
}

}  // namespace

extern "C" {

// flash_attn_fwd on at most max_ctas CTAs (0 = every SM).  Same work items, same arithmetic, same bits: the persistent
// CTAs just claim more items each.
static int fwd_on_ctas(const void* q, const void* k, const void* v, void* o, int B, int H, int N, int D, int causal,
                       void* stream, int max_ctas) {
    int rc = validate(q, k, v, o, B, H, N, N, D);
    if (rc != FA_OK) return rc;
    DeviceState* st = device_state(&rc);
    if (!st) return rc;
    fa::Params p = make_params(B * H, N, N, D, causal, 0, use_split(B * H, N, causal, st->num_sms));
    p.o = static_cast<__half*>(o);
    return run(q, k, v, p, D, static_cast<cudaStream_t>(stream), /*bf16=*/false, /*kv_head_rows=*/0, max_ctas);
}

int flash_attn_fwd(const void* q, const void* k, const void* v, void* o, int B, int H, int N, int D, int causal,
                   void* stream) {
    return fwd_on_ctas(q, k, v, o, B, H, N, D, causal, stream, 0);
}

int flash_attn_fwd_bf16(const void* q, const void* k, const void* v, void* o, int B, int H, int N, int D, int causal,
                        void* stream) {
    int rc = validate(q, k, v, o, B, H, N, N, D);
    if (rc != FA_OK) return rc;
    DeviceState* st = device_state(&rc);
    if (!st) return rc;
    fa::Params p = make_params(B * H, N, N, D, causal, 0, use_split(B * H, N, causal, st->num_sms));
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

int flash_attn_fwd_gathered(const void* q, const void* k, const void* v, void* o, int B, int H, int Nq, int Nkv, int D,
                            int causal, long long q_offset, long long kv_head_rows, const int* ready, int ready_rows,
                            void* stream) {
    int rc = validate(q, k, v, o, B, H, Nq, Nkv, D);
    if (rc != FA_OK) return rc;
    if (kv_head_rows < Nkv) return FA_ERR_BAD_SHAPE;
    if (ready && (ready_rows < fa::kBlockN || ready_rows % fa::kBlockN != 0)) return FA_ERR_BAD_SHAPE;
    fa::Params p = make_params(B * H, Nq, Nkv, D, causal, q_offset, /*split=*/false);
    p.o = static_cast<__half*>(o);
    p.ready = ready;
    p.ready_rows = ready ? ready_rows : 0;
    return run(q, k, v, p, D, static_cast<cudaStream_t>(stream), /*bf16=*/false, kv_head_rows);
}

typedef CUresult (*StreamWriteValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
int flash_attn_stream_write_flag(int* flag, int value, void* stream) {
    if (!flag) return FA_ERR_NULL_PTR;
    static StreamWriteValue32Fn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &f, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<StreamWriteValue32Fn>(f);
    }();
    if (!fn) return FA_ERR_WORKSPACE;
    // a stream memory operation: ordered behind the copies queued on `stream`, executed without a kernel (no SM is free
    // next to the persistent attention grid)
    return fn(static_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(flag), (cuuint32_t)value, 0) == CUDA_SUCCESS
               ? FA_OK : (int)cudaErrorUnknown;
}

int flash_attn_peer_copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                            void* stream) {
    if (!dst || !src) return FA_ERR_NULL_PTR;
    if (width == 0 || height == 0) return FA_OK;
    if (width > dpitch || width > spitch) return FA_ERR_BAD_SHAPE;
    return (int)cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDefault, (cudaStream_t)stream);
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

// Head chunks of one flash_attn_fwd_host call: bound[0] = 0 < bound[1] < ... < bound[chunks] = BH.  A chunk should carry
// enough bytes to amortise its copies, events and launch (>= 16 MiB of input), every chunk carries at least one head, and with
// `taper` the chunks shrink towards the end of the call (weights chunks+2, chunks+1, ..., 3: the last ~1/4 of the first)
// because the last chunk's kernel and its way back are the part nothing hides; first_w > 0 replaces the first weight.
static int host_chunk_bounds(int BH, size_t bytes, int want_chunks, bool taper, int first_w, int* bound) {
    int chunks = (int)((3 * bytes) >> 24);
    if (chunks > want_chunks) chunks = want_chunks;
    if (chunks > kHostChunks) chunks = kHostChunks;
    if (chunks > BH) chunks = BH;
    if (chunks < 1) chunks = 1;
    auto weight = [&](int c) -> long long { return !taper ? 1 : (c == 0 && first_w > 0) ? first_w : chunks + 2 - c; };
    long long wsum = 0, acc = 0;
    for (int c = 0; c < chunks; c++) wsum += weight(c);
    bound[0] = 0;
    for (int c = 0; c < chunks; c++) {
        acc += weight(c);
        int b = (int)((long long)BH * acc / wsum);
        if (b <= bound[c]) b = bound[c] + 1;                  // every chunk carries at least one head (chunks <= BH)
        if (b > BH - (chunks - 1 - c)) b = BH - (chunks - 1 - c);
        bound[c + 1] = b;
    }
    bound[chunks] = BH;
    return chunks;
}

static std::atomic<int> g_host_first{-1};         // -1: read FLASH_ATTN_B200_HOST_FIRST once; 0: plain taper; w > 0: weight of the first chunk
static std::atomic<int> g_host_zerocopy{-1};      // -1: read FLASH_ATTN_B200_HOST_ZEROCOPY once; 0 / 1: staged / direct O store
static std::atomic<int> g_host_ctas{-1};          // -1: read FLASH_ATTN_B200_HOST_CTAS once; 0: every SM; n > 0: CTAs of a chunk's kernel under the direct store

// Body of flash_attn_fwd_host once the streams, events and the staging buffer exist.
static int host_pipeline(DeviceState* st, const void* hq, const void* hk, const void* hv, void* ho, int BH, int N, int D,
                         int causal, size_t bytes) {
    cudaError_t e;
    char* base = static_cast<char*>(st->stage);
    char *dq = base, *dk = base + bytes, *dv = base + 2 * bytes, *dout = base + 3 * bytes;
    // The reference copies everything in, dispatches, copies everything out (FA.cu:774-780).  Heads are
    // independent, so the same work is cut into head chunks and pipelined: while chunk c computes,
    // chunk c+1 is on its way in and chunk c-1 on its way out (PCIe is full duplex; three streams,
    // one event pair per chunk).  The wire time of Q, K, V dominates; kernels and O hide under it.
    // The overlap needs PINNED host buffers (cudaHostAlloc / cudaHostRegister): with pageable memory every
    // cudaMemcpyAsync stages through the driver's bounce buffer and blocks this thread, chunk by chunk.
    static const int want_chunks = [] {
        const char* env = getenv("FLASH_ATTN_B200_HOST_CHUNKS");
        const int v = env ? atoi(env) : 0;
        return v >= 1 && v <= kHostChunks ? v : kHostChunksDefault;
    }();
    // FLASH_ATTN_B200_HOST_FIRST=w (A/B runs): weight of the first chunk -- until its kernel starts nothing flows back
    int first_w = g_host_first.load(std::memory_order_relaxed);      // flash_attn_debug_set_host_first (A/B runs)
    if (first_w < 0) {
        const char* env = getenv("FLASH_ATTN_B200_HOST_FIRST");
        first_w = env ? atoi(env) : FA_HOST_FIRST_DEFAULT;
        g_host_first.store(first_w, std::memory_order_relaxed);
    }
    static const bool taper = [] { const char* env = getenv("FLASH_ATTN_B200_HOST_TAPER"); return !(env && atoi(env) == 0); }();
    int bound[kHostChunks + 1];
    const int chunks = host_chunk_bounds(BH, bytes, want_chunks, taper, first_w, bound);
    const size_t head_bytes = (size_t)N * D * sizeof(__half);
    // Q, K, V of a chunk travel one after the other on one stream.  FLASH_ATTN_B200_HOST_STREAMS=3 gives each tensor a
    // stream of its own (the idea: a copy costs ~12 us of set-up on its engine, during which the link idles if nothing else
    // is in flight) -- measured the same or slower (profiles/r02_c34_host_pipeline.log).  What does pay, 1.5-3 %: the
    // chunks shrink towards the end of the call (FLASH_ATTN_B200_HOST_TAPER=0: equal chunks), because the last chunk's
    // kernel and its copy back are the part nothing hides.
    static const int in_streams = [] { const char* env = getenv("FLASH_ATTN_B200_HOST_STREAMS"); return env && atoi(env) == 3 ? 3 : 1; }();
    // FLASH_ATTN_B200_HOST_ZEROCOPY=1: when the caller's O buffer is pinned and mapped into the device's address space, the
    // kernel's epilogue TMA-stores the finished O tiles straight into it -- the copy back is part of the kernel, there is no
    // staging buffer for O, no D2H copy and no copy engine behind the last kernel.  Pageable buffers keep the staged path.
    int zc_mode = g_host_zerocopy.load(std::memory_order_relaxed);      // flash_attn_debug_set_host_zerocopy (tests, A/B runs)
    if (zc_mode < 0) {
        const char* env = getenv("FLASH_ATTN_B200_HOST_ZEROCOPY");
        zc_mode = env ? atoi(env) : FA_HOST_ZEROCOPY_DEFAULT;
        g_host_zerocopy.store(zc_mode, std::memory_order_relaxed);
    }
    int host_ctas = g_host_ctas.load(std::memory_order_relaxed);        // flash_attn_debug_set_host_ctas (A/B runs)
    if (host_ctas < 0) {
        const char* env = getenv("FLASH_ATTN_B200_HOST_CTAS");
        host_ctas = env ? atoi(env) : FA_HOST_CTAS_DEFAULT;
        if (host_ctas < 0) host_ctas = 0;
        g_host_ctas.store(host_ctas, std::memory_order_relaxed);
    }
    char* o_direct = nullptr;
    if (zc_mode == 1) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, ho) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer)
            o_direct = static_cast<char*>(pa.devicePointer);
        else
            cudaGetLastError();       // pageable memory is not an error here
    }
    // FLASH_ATTN_B200_HOST_TRACE=1: device timestamps of every chunk's three stages (last H2D byte, kernel start / end, last
    // D2H byte) relative to the call's first copy, printed to stderr after the call -- where the call's time goes
    // (profiles/r02_c40_host_trace.log).  Timing events exist only in this mode.
    static const bool trace = [] { const char* env = getenv("FLASH_ATTN_B200_HOST_TRACE"); return env && atoi(env) == 1; }();
    cudaEvent_t tr_start = nullptr, tr_in[kHostChunks] = {}, tr_k0[kHostChunks] = {}, tr_k1[kHostChunks] = {}, tr_out[kHostChunks] = {};
    if (trace) {
        cudaEventCreate(&tr_start);
        for (int c = 0; c < chunks; c++) {
            cudaEventCreate(&tr_in[c]); cudaEventCreate(&tr_k0[c]); cudaEventCreate(&tr_k1[c]); cudaEventCreate(&tr_out[c]);
        }
        cudaEventRecord(tr_start, st->host_in);
    }
    for (int c = 0; c < chunks; c++) {
        const int h0 = bound[c], h1 = bound[c + 1];
        const int nh = h1 - h0;
        const size_t off = (size_t)h0 * head_bytes, len = (size_t)nh * head_bytes;
        const char *hq8 = static_cast<const char*>(hq), *hk8 = static_cast<const char*>(hk),
                   *hv8 = static_cast<const char*>(hv);
        cudaStream_t sk = in_streams == 3 ? st->host_in2[0] : st->host_in, sv = in_streams == 3 ? st->host_in2[1] : st->host_in;
        if ((e = cudaMemcpyAsync(dq + off, hq8 + off, len, cudaMemcpyHostToDevice, st->host_in)) != cudaSuccess) return (int)e;
        if ((e = cudaMemcpyAsync(dk + off, hk8 + off, len, cudaMemcpyHostToDevice, sk)) != cudaSuccess) return (int)e;
        if ((e = cudaMemcpyAsync(dv + off, hv8 + off, len, cudaMemcpyHostToDevice, sv)) != cudaSuccess) return (int)e;
        if ((e = cudaEventRecord(st->ev_in[c], st->host_in)) != cudaSuccess) return (int)e;
        if ((e = cudaStreamWaitEvent(st->host_stream, st->ev_in[c], 0)) != cudaSuccess) return (int)e;
        if (in_streams == 3) {
            for (int s2 = 0; s2 < 2; s2++) {
                if ((e = cudaEventRecord(st->ev_in2[s2][c], st->host_in2[s2])) != cudaSuccess) return (int)e;
                if ((e = cudaStreamWaitEvent(st->host_stream, st->ev_in2[s2][c], 0)) != cudaSuccess) return (int)e;
            }
        }
        if (trace) { cudaEventRecord(tr_in[c], st->host_in); cudaEventRecord(tr_k0[c], st->host_stream); }
        // With the direct store the kernel IS the way back, and PCIe carries its writes upstream next to the read requests
        // of the H2D copies: a kernel on every SM emits O as fast as the link takes it and the copies in fall from 54 to
        // ~35 GB/s while it runs (profiles/r02_c41_host_trace_zerocopy.log).  FLASH_ATTN_B200_HOST_CTAS=n (A/B switch, off by
        // default) runs every chunk's kernel but the last on n CTAs, stretching it over the next chunk's way in so that O
        // trickles back instead: bit-identical output and the SAME call time for n = 64 ... 14 (4.206-4.212 ms,
        // profiles/r02_c48_host_ab_ctas.log) -- the link charges the bytes of either direction whenever they travel.
        const int lim = (o_direct && c + 1 < chunks) ? host_ctas : 0;
        int rc = fwd_on_ctas(dq + off, dk + off, dv + off, o_direct ? o_direct + off : dout + off, 1, nh, N, D, causal,
                             st->host_stream, lim);   // FA.cu:777
        if (rc != FA_OK) return rc;
        if (trace) cudaEventRecord(tr_k1[c], st->host_stream);
        if (o_direct) {
            if (trace) cudaEventRecord(tr_out[c], st->host_stream);
            continue;
        }
        if ((e = cudaEventRecord(st->ev_k[c], st->host_stream)) != cudaSuccess) return (int)e;
        if ((e = cudaStreamWaitEvent(st->host_out, st->ev_k[c], 0)) != cudaSuccess) return (int)e;
        if ((e = cudaMemcpyAsync(static_cast<char*>(ho) + off, dout + off, len, cudaMemcpyDeviceToHost, st->host_out)) !=
            cudaSuccess)
            return (int)e;
        if (trace) cudaEventRecord(tr_out[c], st->host_out);
    }
    e = cudaStreamSynchronize(o_direct ? st->host_stream : st->host_out);
    if (trace) {
        fprintf(stderr, "host_trace chunks=%d (ms after the first copy was queued: heads | H2D done | kernel start, end | D2H done)\n", chunks);
        for (int c = 0; c < chunks; c++) {
            float a = 0.f, b = 0.f, k1 = 0.f, d = 0.f;
            cudaEventElapsedTime(&a, tr_start, tr_in[c]);
            cudaEventElapsedTime(&b, tr_start, tr_k0[c]);
            cudaEventElapsedTime(&k1, tr_start, tr_k1[c]);
            cudaEventElapsedTime(&d, tr_start, tr_out[c]);
            fprintf(stderr, "host_trace %2d: %2d | %.3f | %.3f %.3f | %.3f\n", c, bound[c + 1] - bound[c], a, b, k1, d);
            cudaEventDestroy(tr_in[c]); cudaEventDestroy(tr_k0[c]); cudaEventDestroy(tr_k1[c]); cudaEventDestroy(tr_out[c]);
        }
        cudaEventDestroy(tr_start);
    }
    return (int)e;
}

int flash_attn_fwd_host(const void* hq, const void* hk, const void* hv, void* ho, int B, int H, int N, int D,
                        int causal) {
    if (!hq || !hk || !hv || !ho) return FA_ERR_NULL_PTR;
    if (D != 64 && D != 128) return FA_ERR_BAD_HEAD_DIM;
    if (B < 1 || H < 1 || N < 1) return FA_ERR_BAD_SHAPE;
    if ((long long)B * H > 0x7fffffffLL) return FA_ERR_BAD_SHAPE;
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
        for (int s2 = 0; s2 < 2; s2++)
            if ((e = cudaStreamCreateWithFlags(&st->host_in2[s2], cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
        for (int i = 0; i < kHostChunks; i++) {
            if ((e = cudaEventCreateWithFlags(&st->ev_in[i], cudaEventDisableTiming)) != cudaSuccess) return (int)e;
            if ((e = cudaEventCreateWithFlags(&st->ev_k[i], cudaEventDisableTiming)) != cudaSuccess) return (int)e;
            for (int s2 = 0; s2 < 2; s2++)
                if ((e = cudaEventCreateWithFlags(&st->ev_in2[s2][i], cudaEventDisableTiming)) != cudaSuccess) return (int)e;
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
    int rc = host_pipeline(st, hq, hk, hv, ho, B * H, N, D, causal, bytes);
    if (rc != FA_OK) {
        // an error in the middle of the chunk loop leaves copies and kernels queued on the cached staging buffer
        // (and on the caller's host memory): nothing may still be in flight when the caller sees the error
        cudaStreamSynchronize(st->host_in);
        for (int s2 = 0; s2 < 2; s2++) cudaStreamSynchronize(st->host_in2[s2]);
        cudaStreamSynchronize(st->host_stream);
        cudaStreamSynchronize(st->host_out);
        return rc;
    }
    return take_watchdog(st);   // O is on the host: say so if a kernel of this call gave up on a barrier
}

int flash_attn_get_kernel_info(int B, int H, int N, int D, int causal, flash_attn_kernel_info* info) {
    if (!info) return FA_ERR_NULL_PTR;
    if (D != 64 && D != 128) return FA_ERR_BAD_HEAD_DIM;
    if (B < 1 || H < 1 || N < 1) return FA_ERR_BAD_SHAPE;
    memset(info, 0, sizeof *info);
    cudaFuncAttributes attr;
    cudaError_t e = D == 64            ? cudaFuncGetAttributes(&attr, fa::fa_fwd_kernel<64, kPolyD64>)
                    : use_poly(D, N, causal) ? cudaFuncGetAttributes(&attr, fa::fa_fwd_kernel<128, kPolyLong>)
                                     : cudaFuncGetAttributes(&attr, fa::fa_fwd_kernel<128, 0>);
    if (e != cudaSuccess) return (int)e;
    int err = 0;
    DeviceState* st = device_state(&err);
    if (!st) return err;
    if ((long long)B * H > 0x7fffffffLL) return FA_ERR_BAD_SHAPE;
    fa::Params p = make_params(B * H, N, N, D, causal, 0, use_split(B * H, N, causal, st->num_sms));
    if ((long long)p.BH * p.nqp > 0x7fffffffLL) return FA_ERR_BAD_SHAPE;
    info->cta_group = 1;
    info->regs_per_thread = attr.numRegs;
    info->local_bytes_per_thread = (int)attr.localSizeBytes;
    info->static_smem_bytes = (int)attr.sharedSizeBytes;
    info->dynamic_smem_bytes = D == 128 ? fa::Cfg<128>::kSmemBytes : fa::Cfg<64>::kSmemBytes;
    info->threads_per_cta = fa::kNumThreads;
    int avail = st->num_sms - g_sm_margin.load(std::memory_order_relaxed);   // flash_attn_set_sm_margin
    if (avail < 1) avail = 1;
    info->ctas = p.total_work < avail ? p.total_work : avail;
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
        std::lock_guard<std::mutex> init_lock(st->init_mu);
        std::lock_guard<std::mutex> lock(st->host_mu);
        if (!st->inited) continue;
        cudaSetDevice(d);
        cudaDeviceSynchronize();
        if (st->stage) cudaFree(st->stage);
        if (st->host_stream) cudaStreamDestroy(st->host_stream);
        if (st->host_in) cudaStreamDestroy(st->host_in);
        for (int s2 = 0; s2 < 2; s2++) {
            if (st->host_in2[s2]) cudaStreamDestroy(st->host_in2[s2]);
            st->host_in2[s2] = nullptr;
        }
        if (st->host_out) cudaStreamDestroy(st->host_out);
        for (int i = 0; i < kHostChunks; i++) {
            if (st->ev_in[i]) cudaEventDestroy(st->ev_in[i]);
            if (st->ev_k[i]) cudaEventDestroy(st->ev_k[i]);
            st->ev_in[i] = st->ev_k[i] = nullptr;
            for (int s2 = 0; s2 < 2; s2++) {
                if (st->ev_in2[s2][i]) cudaEventDestroy(st->ev_in2[s2][i]);
                st->ev_in2[s2][i] = nullptr;
            }
        }
        st->stage = nullptr;
        st->stage_bytes = 0;
        st->host_stream = st->host_in = st->host_out = nullptr;
        st->host_ready = false;
        if (st->sched) cudaFree(st->sched);
        st->sched = nullptr;
        if (st->wd_host) {
            unsigned int* none = nullptr;
            cudaMemcpyToSymbol(sm100::g_watchdog_host, &none, sizeof none);
            cudaFreeHost(st->wd_host);
        }
        st->wd_host = nullptr;
        {
            std::lock_guard<std::mutex> tl(st->tmap_mu);
            for (TmapSet& c : st->tmaps) c.stamp = 0;
        }
        st->inited = false;      // the next call on this device sets it up again
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
        case FA_ERR_WATCHDOG: return "an earlier kernel gave up waiting on a barrier and produced garbage (flash_attn_status has the record)";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

const char* flash_attn_version(void) { return "flashattn_b200 0.3 (sm_100a tcgen05/TMA forward)"; }

}  // extern "C"

// Watchdog record of the current device, {aborted, barrier tag, block, thread} (see sm100_ptx.cuh): a pending one
// (aborted = 1: the next call will return FA_ERR_WATCHDOG) or else the last one that was reported (aborted = 0 and
// the fields of that report; all zero if there never was one).  Does not synchronise, does not clear.
extern "C" int flash_attn_status(unsigned int* out4) {
    if (!out4) return FA_ERR_NULL_PTR;
    int err = 0;
    DeviceState* st = device_state(&err);
    if (!st) return err;
    const volatile unsigned int* h = st->wd_host;
    if (h && h[0]) {
        for (int i = 0; i < 4; i++) out4[i] = h[i];
    } else {
        out4[0] = 0;
        for (int i = 1; i < 4; i++) out4[i] = st->wd_last[i];
    }
    return FA_OK;
}
// Test hooks: the same after a device synchronise / a one-thread kernel that raises the record the way a timed-out wait does.
extern "C" int flash_attn_debug_status(unsigned int* out4) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return (int)e;
    return flash_attn_status(out4);
}
__global__ void fa_debug_trip_kernel(unsigned int tag) {
    sm100::watchdog_raise((int)tag);
    sm100::watchdog_publish();
}
extern "C" int flash_attn_debug_trip_watchdog(unsigned int tag, void* stream) {
    int err = 0;
    if (!device_state(&err)) return err;
    fa_debug_trip_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(tag);
    return (int)cudaGetLastError();
}

#ifdef FA_TIMING
// debug builds only (-DFA_TIMING): in-kernel clock64 probes, see tests/harness/timing.py
extern "C" int flash_attn_debug_timing(unsigned long long* out32, int reset) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return (int)e;
    unsigned long long z[64] = {0};
    e = cudaMemcpyFromSymbol(out32, fa::g_timing, 64 * sizeof(unsigned long long));
    if (e == cudaSuccess && reset) e = cudaMemcpyToSymbol(fa::g_timing, z, 64 * sizeof(unsigned long long));
    return (int)e;
}
#endif

// Host-side mirror of the device work decomposition, exported for the scheduler tests
// (every (bh, q-tile) exactly once, heavy-first, masked tiles skipped).  split: 0 = pair items, 1 = split mode.
extern "C" int flash_attn_debug_work_item(int w, int B, int H, int Nq, int Nkv, int D, int causal, long long shift,
                                          int* total, int* bh, int* q0, int* n0, int* n1) {
    if (D != 64 && D != 128) return FA_ERR_BAD_HEAD_DIM;
    fa::Params p = make_params(B * H, Nq, Nkv, D, causal, shift, false);
    *total = p.total_work;
    if (w < 0 || w >= p.total_work) return FA_ERR_BAD_SHAPE;
    fa::WorkItem it = fa::decode_work(w, p);
    *bh = it.bh; *q0 = it.q0; *n0 = it.n0; *n1 = it.n1;
    return FA_OK;
}
extern "C" int flash_attn_debug_work_item_split(int w, int B, int H, int Nq, int Nkv, int D, int causal, long long shift,
                                                int* total, int* bh, int* q0, int* n0, int* n1) {
    if (D != 64 && D != 128) return FA_ERR_BAD_HEAD_DIM;
    fa::Params p = make_params(B * H, Nq, Nkv, D, causal, shift, true);
    *total = p.total_work;
    if (w < 0 || w >= p.total_work) return FA_ERR_BAD_SHAPE;
    fa::WorkItem it = fa::decode_work(w, p);
    *bh = it.bh; *q0 = it.q0; *n0 = it.n0; *n1 = it.n1;
    return FA_OK;
}
// 128-row Q tiles a pair-mode work item covers
extern "C" int flash_attn_debug_tiles_per_item(int D) {
    if (D != 64 && D != 128) return FA_ERR_BAD_HEAD_DIM;
    return 2;
}
// -1 automatic (use_split), 0 never, 1 always: which work decomposition flash_attn_fwd uses from now on (A/B runs, tests)
extern "C" void flash_attn_debug_set_split(int mode) { g_split_override.store(mode < -1 || mode > 1 ? -1 : mode, std::memory_order_relaxed); }
// 0: flash_attn_fwd_host copies O back with a copy engine; 1: the kernel stores O straight into a pinned, mapped host buffer
extern "C" void flash_attn_debug_set_host_zerocopy(int mode) { g_host_zerocopy.store(mode ? 1 : 0, std::memory_order_relaxed); }
// Test hook: the head-chunk boundaries flash_attn_fwd_host would use (bound must hold 33 ints); returns the chunk count
extern "C" int flash_attn_debug_host_chunks(int BH, long long bytes_per_tensor, int want_chunks, int taper, int first_w, int* bound) {
    if (!bound || BH < 1 || bytes_per_tensor < 1 || want_chunks < 1) return FA_ERR_BAD_SHAPE;
    return host_chunk_bounds(BH, (size_t)bytes_per_tensor, want_chunks, taper != 0, first_w, bound);
}
extern "C" void flash_attn_debug_set_host_ctas(int n) { g_host_ctas.store(n < 0 ? 0 : n, std::memory_order_relaxed); }
extern "C" void flash_attn_debug_set_host_first(int w) { g_host_first.store(w < 0 ? 0 : w, std::memory_order_relaxed); }
// 1 when flash_attn_fwd would run this shape in split mode on the current device
extern "C" int flash_attn_debug_uses_split(int B, int H, int N, int causal) {
    int err = 0;
    DeviceState* st = device_state(&err);
    if (!st) return err;
    return use_split(B * H, N, causal, st->num_sms) ? 1 : 0;
}
