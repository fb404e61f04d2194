// fa_fwd_pair_sm100.cuh -- experimental CTA-pair FlashAttention forward for B200 (sm_100a), "v9":
// one Q tile per CTA, cta_group::2 MMAs, decoupled S/P buffers, TMA-store epilogue.
// Opt-in (FLASH_ATTN_B200_KERNEL=pair); the default kernel is fa_fwd_sm100.cuh.  Parity-tested;
// measured 1-2 % (D=128) to 12 % (D=64) behind the default -- profiles/r01_v9_experiments.txt.
//
// Replaces the reference's device path flash_attention_v9<...> (flash_attention.cu:67-554):
//   FA.cu:103-112  block->(bh, q-block) mapping, GRID_SWAP    -> persistent work loop, heavy-first
//   FA.cu:145-159  Q fragments in registers                   -> Q tile resident in smem (TMA), double-buffered
//   FA.cu:417-447  synchronous K/V tile load + 2 barriers     -> producer warp, separate K and V mbarrier rings
//   FA.cu:188-233  DO_QK_MATMUL (mma.sync m16n8k16)           -> tcgen05.mma SS, S in TMEM
//   FA.cu:235-288  DO_SOFTMAX (quad shuffles, eager rescale)  -> one thread per row, lazy rescale
//   FA.cu:290-334  DO_PV_MATMUL (P in registers)              -> P fp16 in TMEM, tcgen05.mma TS
//   FA.cu:497-553  multi-pass smem output staging             -> TMEM -> registers -> swizzled smem -> TMA store
//   FA.cu:460-496  split-K partial epilogue (dead code there) -> partial mode used by ring CP
//
// Work unit = one head x CG consecutive 128-row Q tiles, CG = 1 (one CTA) or 2 (a CTA pair driving
// tcgen05.mma.cta_group::2 with M = 256: each CTA owns 128 Q rows and supplies half of every K and V
// tile, so K/V smem fill, smem operand reads and L2 traffic per SM are halved).
//
// CTA = 384 threads:  warps 0-3  softmax set A: KV tiles with even global index
//                     warps 4-7  softmax set B: odd tiles -- SAME 128 query rows as set A
//                     warp 8     TMEM allocator + tcgen05.mma issuer for S = Q K^T (leader CTA only)
//                     warp 9     TMA producer (Q, K and V tiles) + dynamic tile scheduler
//                     warp 10    TMA store of finished O tiles
//                     warp 11    tcgen05.mma issuer for O += P V (leader CTA only)
// Two issuer warps because one warp's instruction latency (~6 cycles per dependent instruction
// when it runs alone) made a single issuer the bottleneck at ~280 instructions per tile
// (profiles/r01_v9d_*); the two MMA streams only meet through mbarriers anyway.
//
// TMEM (512 columns x 128 lanes x 32 bit):  S_a [0,128)  S_b [128,256)  P_a [256,320)  P_b [320,384)
//                                           O [384,384+D)
// S and P do not alias, and each S buffer belongs to one softmax set, so the tensor pipe never waits
// for a whole S -> P -> PV -> QK round trip of one tile (the limiter of the previous two-Q-tile
// design, profiles/r01_v4_*): QK(g+2) is issued as soon as set (g&1) has READ S(g) into registers,
// PV(g) as soon as P(g) lands, and the two sets run their softmax of consecutive tiles concurrently.
// Because both sets feed one O accumulator they share one lazily-updated reference max per row:
// a set publishes its m_ref after every tile (smem + mbarrier, per 32-row quadrant) and the other
// set picks it up before deciding its own; partial row sums are merged in the epilogue.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "sm100_ptx.cuh"

namespace fa_pair {

using namespace sm100;

constexpr int kBlockM = 128;      // Q rows per CTA tile (UMMA M per CTA)
constexpr int kBlockN = 128;      // K/V rows per tile (UMMA N of QK^T, K extent of PV)
constexpr int kNumThreads = 384;
constexpr int kMmaWarp = 8;
constexpr int kLoadWarp = 9;
constexpr int kStoreWarp = 10;
constexpr int kPvWarp = 11;
constexpr int kTmemCols = 512;
constexpr int kRegsSoftmax = 208;   // setmaxnreg: softmax warpgroups grow, the producer/MMA warpgroup shrinks
constexpr int kRegsOther = 80;      // 2*128*208 + 128*80 = 63488 <= 168 (launch) * 384
constexpr float kRescaleThreshold = 8.0f;   // lazy rescale: tolerate P up to 2^8 before moving the reference max
// Of every 4 element pairs, this many take exp2 on the FMA pipe (Cody-Waite + degree-3 minimax)
// instead of MUFU.EX2: at 16 MUFU/clk/SM the 16384 exponentials of a 128x128 tile cost as many
// cycles as its two MMAs, so the SFU -- not the tensor core -- would set the pace.
// Where the m_ready arrival / wait of the shared reference max sit: 0 = arrive at the publish, wait at
// the decision; 1 = arrive after the first half of P, wait at tile start (hard anti-phase: measured
// slower); 2 = arrive after the first half of P, wait at the decision (soft anti-phase: +2 %).
#ifndef FA_ANTIPHASE
#define FA_ANTIPHASE 2
#endif
#ifndef FA_POLY_PAIRS
#define FA_POLY_PAIRS 1
#endif
constexpr int kPolyPairs = FA_POLY_PAIRS;

template <int D, int CG>
struct Cfg {
    static_assert(D == 64 || D == 128, "head_dim must be 64 or 128");
    static_assert(CG == 1 || CG == 2, "CTA group size is 1 or 2");
    static_assert(D / CG >= 64, "the V half of a CTA pair must be at least one 128-byte panel wide");
    static constexpr int kPanels = D / 64;                      // 128-byte swizzle panels per Q/K row
    static constexpr int kQPanelBytes = kBlockM * 128;          // 128 rows x 128 B
    static constexpr int kQTileBytes = kPanels * kQPanelBytes;
    static constexpr int kKRows = kBlockN / CG;                 // keys of a tile this CTA loads
    static constexpr int kKPanelBytes = kKRows * 128;
    static constexpr int kKBytes = kPanels * kKPanelBytes;      // ring entry: this CTA's part of a K tile
    static constexpr int kVPanels = D / (64 * CG);              // this CTA's d-columns of V, in panels
    static constexpr int kVPanelBytes = kBlockN * 128;
    static constexpr int kVBytes = kVPanels * kVPanelBytes;     // ring entry: this CTA's part of a V tile
    static constexpr int kKStages = (D == 128) ? (CG == 1 ? 2 : 4) : 5;
    static constexpr int kVStages = (D == 128) ? (CG == 1 ? 2 : 5) : 6;
    static constexpr int kOffK = 2 * kQTileBytes;               // Q is double-buffered (and stages O)
    static constexpr int kOffV = kOffK + kKStages * kKBytes;
    static constexpr int kOffBar = kOffV + kVStages * kVBytes;
    // barrier slots (8 B each)
    static constexpr int kBarQFull = 0;                          // [2]   leader
    static constexpr int kBarQEmpty = kBarQFull + 2;             // [2]   per CTA
    static constexpr int kBarKFull = kBarQEmpty + 2;             // [kKStages] leader
    static constexpr int kBarKEmpty = kBarKFull + kKStages;      // per CTA
    static constexpr int kBarVFull = kBarKEmpty + kKStages;      // [kVStages] leader
    static constexpr int kBarVEmpty = kBarVFull + kVStages;      // per CTA
    static constexpr int kBarSFull = kBarVEmpty + kVStages;      // [2]   per CTA (multicast commit)
    static constexpr int kBarSFree = kBarSFull + 2;              // [2]   leader
    static constexpr int kBarPFull = kBarSFree + 2;              // [2][2] leader, index 2*b + half
    static constexpr int kBarPvDone = kBarPFull + 4;             // [2]   per CTA (multicast commit)
    static constexpr int kBarOFree = kBarPvDone + 2;             // [1]   leader
    static constexpr int kBarOStaged = kBarOFree + 1;            // [2]   per CTA, one per Q slot
    static constexpr int kBarMReady = kBarOStaged + 2;           // [2][4] per CTA, index 4*set + quadrant
    static constexpr int kBarSchedFull = kBarMReady + 8;         // [2]   per CTA
    static constexpr int kBarSchedEmpty = kBarSchedFull + 2;     // [2]   leader
    static constexpr int kNumBars = kBarSchedEmpty + 2;
    static constexpr int kOffMisc = kOffBar + kNumBars * 8;      // tmem slot (4) + pad (4) + mailbox 2 x int
    static constexpr int kOffMref = (kOffMisc + 16 + 15) & ~15;  // float [2 sets][128 rows]
    static constexpr int kOffFin = kOffMref + 2 * kBlockM * 4;   // float2 [2 parities][2 sets][128 rows]
    static constexpr int kSmemBytes = kOffFin + 2 * 2 * kBlockM * 8;
    static_assert(kSmemBytes <= 232448, "exceeds the 227 KB opt-in shared memory of sm_100");
    static constexpr int kTmemS = 0, kTmemP = 256, kTmemO = 384;   // S_b at kTmemS + 128 b, P_b at kTmemP + 64 b
    static constexpr uint32_t kIdescQK = umma_idesc_f16(kBlockM * CG, kBlockN, 0, 0);
    static constexpr uint32_t kIdescPV = umma_idesc_f16(kBlockM * CG, D, 0, 1);  // V is MN-major ([kv][d], d contiguous)
    static constexpr int kSchedConsumers = CG == 1 ? 11 : 21;   // leader: 8 softmax warps + store warp + 2 MMA warps; peer: 8 + store + its producer
};

struct Params {
    float* o_partial;   // fp32 un-normalised [BH*Nq, D]       (partial_mode == 1; FA.cu:460-496 format)
    float* ml;          // (m, l) pairs [BH*Nq, 2]
    __half* o;          // fp16 output [BH, Nq, D]: only used for rows without any visible key (zeros)
    int Nq, Nkv, BH;
    int causal;
    int shift;          // q_offset - kv_offset: key c visible to query r iff c <= r + shift
    int cg;             // CTAs per work unit (1 or 2)
    int nqu;            // work units per head = ceil(Nq / (128 * cg))
    int total_work;     // BH * nqu
    int group_heads;    // heads per scheduling group (their K/V working set is sized to stay in L2)
    int partial_mode;
    int accumulate;
    int* sched;         // {next work index, finished CTAs}: dynamic tile scheduler state, self-resetting
    float scale;        // 1/sqrt(D)
    float scale_log2;   // scale * log2(e)
};

// ---- work decomposition (shared by host tests and every warp role) ----
struct WorkItem {
    int bh, q0;      // head index, first local query row of the unit
    int n0, n1;      // KV tiles the unit's Q tiles need (0 = nothing visible / tile absent); n1 = 0 when cg == 1
    int n;           // KV tiles the unit streams = max(n0, n1)
};
__host__ __device__ inline int kv_trip_count(int q_start, int Nq, int Nkv, int causal, int shift) {
    if (q_start >= Nq) return 0;
    const int nkv_tiles = (Nkv + kBlockN - 1) / kBlockN;
    if (!causal) return nkv_tiles;
    int last_row = q_start + kBlockM - 1;
    if (last_row > Nq - 1) last_row = Nq - 1;
    long long vis = (long long)last_row + shift + 1;  // keys [0, vis) visible to the last row
    if (vis <= 0) return 0;
    if (vis > Nkv) vis = Nkv;
    return (int)((vis + kBlockN - 1) / kBlockN);
}
// Work order (replaces GRID_SWAP / reversed q-blocks, FA.cu:103-112).  Heads are taken in groups whose
// K/V fit comfortably in L2; inside a group the order is heaviest Q unit first ACROSS the group's
// heads (causal: the last unit sees the most keys), so the dynamic scheduler hands out long items
// early and the tail of the launch is made of the lightest ones, while the CTAs running at any moment
// still share a few heads' K/V through L2.
__host__ __device__ inline WorkItem decode_work(int w, const Params& p) {
    WorkItem it;
    const int per_group = p.group_heads * p.nqu;
    const int g = w / per_group;
    const int r = w - g * per_group;
    int heads = p.BH - g * p.group_heads;
    if (heads > p.group_heads) heads = p.group_heads;
    const int u = p.nqu - 1 - r / heads;
    it.bh = g * p.group_heads + r % heads;
    it.q0 = u * p.cg * kBlockM;
    it.n0 = kv_trip_count(it.q0, p.Nq, p.Nkv, p.causal, p.shift);
    it.n1 = p.cg == 2 ? kv_trip_count(it.q0 + kBlockM, p.Nq, p.Nkv, p.causal, p.shift) : 0;
    it.n = it.n0 > it.n1 ? it.n0 : it.n1;
    return it;
}

#ifdef FA_TIMING
__device__ unsigned long long g_timing_pair[64];
// phase probe: accumulates clock deltas of one sampled warp into g_timing_pair[base + i]
#define FA_PROBE_DECL long long _pt = clock64(); const bool _ps = (threadIdx.x == 0) && ((gk & 7u) == 3u);
#define FA_PROBE(i) { const long long _n = clock64(); if (_ps) atomicAdd(&g_timing_pair[32 + (i)], (unsigned long long)(_n - _pt)); _pt = _n; }
#ifdef FA_TIMING_MMA   // distorts the MMA warp (an atomic per probe): separate switch
#define FA_MPROBE_DECL long long _mt = clock64();
#define FA_MPROBE(i) { const long long _n = clock64(); if (lane == 0) atomicAdd(&g_timing_pair[48 + (i)], (unsigned long long)(_n - _mt)); _mt = _n; }
#else
#define FA_MPROBE_DECL
#define FA_MPROBE(i)
#endif
#else
#define FA_PROBE_DECL
#define FA_PROBE(i)
#define FA_MPROBE_DECL
#define FA_MPROBE(i)
#endif

struct Ring {
    uint32_t idx, phase;
    template <int kStages>
    __device__ __forceinline__ void advance() {
        if (++idx == (uint32_t)kStages) { idx = 0; phase ^= 1u; }
    }
};

// ---- exp2 of a pair of (already scaled and shifted) scores ----
// kPoly = false: two MUFU.EX2.  kPoly = true: FMA/ALU pipes only.  x = n + f, n = round(x),
// f in [-0.5, 0.5]; 2^f by a degree-3 minimax polynomial (max relative error 7.5e-5, below the
// 4.9e-4 of the fp16 rounding P gets anyway); 2^n by adding n into the exponent field.
template <bool kPoly>
__device__ __forceinline__ void exp2_pair(uint64_t x2, float& p0, float& p1) {
    if (!kPoly) {
        float x0, x1;
        unpack_f32x2(x2, x0, x1);
        p0 = ex2_approx(x0);
        p1 = ex2_approx(x1);
    } else {
        float x0, x1;
        unpack_f32x2(x2, x0, x1);
        x0 = fmaxf(x0, -126.0f);                       // masked (-inf) and far-away scores -> 2^-126 ~ 0
        x1 = fmaxf(x1, -126.0f);
        x2 = pack_f32x2(x0, x1);
        const uint64_t magic = pack_f32x2(12582912.0f, 12582912.0f);        // 1.5 * 2^23: rounds to integer
        const uint64_t t2 = add_f32x2(x2, magic);
        const uint64_t n2 = add_f32x2(t2, pack_f32x2(-12582912.0f, -12582912.0f));
        const uint64_t f2 = fma_f32x2(n2, pack_f32x2(-1.0f, -1.0f), x2);
        uint64_t q2 = fma_f32x2(pack_f32x2(0.05517143756151199f, 0.05517143756151199f), f2,
                                pack_f32x2(0.24261081218719482f, 0.24261081218719482f));
        q2 = fma_f32x2(q2, f2, pack_f32x2(0.6932609677314758f, 0.6932609677314758f));
        q2 = fma_f32x2(q2, f2, pack_f32x2(0.9999281167984009f, 0.9999281167984009f));
        float t0, t1, q0, q1;
        unpack_f32x2(t2, t0, t1);
        unpack_f32x2(q2, q0, q1);
        p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
        p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
    }
}

// exponentials + fp16 packing of 64 consecutive columns (one half of the tile)
__device__ __forceinline__ void exp_half(const uint32_t* s, uint32_t* pk, uint64_t scale2, uint64_t neg2,
                                         uint64_t& sum_a, uint64_t& sum_b) {
#pragma unroll
    for (int i = 0; i < 64; i += 8) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int e = i + 2 * q;
            const uint64_t x2 =
                fma_f32x2(pack_f32x2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), scale2, neg2);
            float p0, p1;
#ifdef FA_SKELETON
            unpack_f32x2(x2, p0, p1);
#else
            if (q < kPolyPairs) exp2_pair<true>(x2, p0, p1);
            else exp2_pair<false>(x2, p0, p1);
#endif
            if (q & 1) sum_b = add_f32x2(sum_b, pack_f32x2(p0, p1));   // row sum of the un-rounded p (FA.cu:273-279)
            else sum_a = add_f32x2(sum_a, pack_f32x2(p0, p1));
            __half2 h = __floats2half2_rn(p0, p1);                     // low half = even column
            pk[e / 2] = *reinterpret_cast<uint32_t*>(&h);
        }
    }
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// Everything a softmax thread needs to know about where things live.
struct SoftmaxCtx {
    uint32_t tS, tP, tO;            // TMEM addresses of this set's S / P buffer and of O, lane field included
    uint32_t bar_s_free;            // leader's (shared::cluster address when CG == 2)
    uint32_t bar_p_full;            // leader's, + 8 * half
    uint32_t bar_pv_done_mine;      // local: the PV that read this set's P buffer has retired
    uint32_t bar_pv_done_other;     // local: same for the other set's buffer
    uint32_t bar_m_ready_mine;      // local, this warp's quadrant
    uint32_t bar_m_ready_other;
    float* mref_mine;               // &mref[set][row]
    const float* mref_other;        // &mref[1-set][row]
};

template <int CG>
__device__ __forceinline__ void arrive_leader(uint32_t bar) {
    if (CG == 1) mbar_arrive(bar);
    else mbar_arrive_cluster(bar);
}

// ---- softmax of one 128x128 S tile; one thread owns one row ----
//   gk     = index of this tile among the tiles of this set's buffers (global tile index >> 1)
//   first  = first KV tile of the work unit (no running state yet)
template <int D, int CG, bool kMask>
__device__ __forceinline__ void softmax_tile(const Params& p, const SoftmaxCtx& c, int lim_local, bool first,
                                             uint32_t gk, uint32_t g_prev_k, float& m_ref, float& l_run) {
    FA_PROBE_DECL
    uint32_t s[kBlockN];
    tmem_ld_x32(c.tS + 0, s + 0);
    tmem_ld_x32(c.tS + 32, s + 32);
    tmem_ld_x32(c.tS + 64, s + 64);
    tmem_ld_x32(c.tS + 96, s + 96);
    tmem_wait_ld();
    // S is in registers: the tensor core may overwrite this buffer with the tile after next
    tc_fence_before();
    __syncwarp();
    if (lane_id() == 0) arrive_leader<CG>(c.bar_s_free);
    FA_PROBE(0)

    if (kMask) {
#pragma unroll
        for (int i = 0; i < kBlockN; i++)
            if (i >= lim_local) s[i] = 0xff800000u;  // -inf
    }

    // row max: 3-input max (FMNMX3), four independent chains
    float mx0 = fmaxf(__uint_as_float(s[0]), __uint_as_float(s[1]));
    float mx1 = fmaxf(__uint_as_float(s[2]), __uint_as_float(s[3]));
    float mx2 = fmaxf(__uint_as_float(s[4]), __uint_as_float(s[5]));
    float mx3 = fmaxf(__uint_as_float(s[6]), __uint_as_float(s[7]));
#pragma unroll
    for (int i = 8; i < kBlockN; i += 8) {
        mx0 = fmax3(mx0, __uint_as_float(s[i + 0]), __uint_as_float(s[i + 1]));
        mx1 = fmax3(mx1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        mx2 = fmax3(mx2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
        mx3 = fmax3(mx3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
    }
    const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
    FA_PROBE(1)

    // The other set handled the previous tile of this row: adopt its reference max (it has already
    // rescaled O; only this set's partial row sum has to follow).
    if (!first) {
#if FA_ANTIPHASE != 1
        mbar_wait(c.bar_m_ready_other, g_prev_k & 1u, 24);
        const float m_prev = *reinterpret_cast<const volatile float*>(c.mref_other);
        if (m_prev > m_ref) {
            l_run = (m_ref == -INFINITY) ? 0.f : l_run * ex2_approx((m_ref - m_prev) * p.scale_log2);
            m_ref = m_prev;
        }
    }
    FA_PROBE(2)
    const float m_new = fmaxf(m_ref, m_tile);

    // Lazy rescale (replaces the reference's every-tile O *= alpha, FA.cu:267-270): the reference
    // max only moves when the true max has outgrown it by 2^kRescaleThreshold.
    const bool need = (m_new - m_ref) * p.scale_log2 > kRescaleThreshold;  // NaN (-inf - -inf) -> false
    if (__any_sync(0xffffffffu, need)) {
        if (!first) {
            const float alpha = (m_new == -INFINITY) ? 1.0f : ex2_approx((m_ref - m_new) * p.scale_log2);
            const uint64_t alpha2 = pack_f32x2(alpha, alpha);
            // O holds PV(0..g-1); the last of them (issued from the other set's P) must have retired
            mbar_wait(c.bar_pv_done_other, g_prev_k & 1u, 40);
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < D; cc += 32) {
                uint32_t o[32];
                tmem_ld_x32(c.tO + cc, o);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float lo, hi;
                    unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(o[i]), __uint_as_float(o[i + 1])), alpha2), lo, hi);
                    o[i] = __float_as_uint(lo);
                    o[i + 1] = __float_as_uint(hi);
                }
                tmem_st_x32(c.tO + cc, o);
            }
            tmem_wait_st();
            l_run *= alpha;
        }
        m_ref = m_new;
    }
    // publish the reference max this row now uses (with FA_ANTIPHASE the other set is released only
    // after the first half of P has gone out -- see the note in the softmax loop)
    *reinterpret_cast<volatile float*>(c.mref_mine) = m_ref;
#if !FA_ANTIPHASE
    __syncwarp();
    if (lane_id() == 0) mbar_arrive(c.bar_m_ready_mine);
#endif
    FA_PROBE(3)

    const float m_used = (m_ref == -INFINITY) ? 0.0f : m_ref;
    const float neg = -m_used * p.scale_log2;
    const uint64_t scale2 = pack_f32x2(p.scale_log2, p.scale_log2);
    const uint64_t neg2 = pack_f32x2(neg, neg);
    uint64_t sum_a = 0ull, sum_b = 0ull;     // (0.f, 0.f)
    uint32_t pk[32];
    // P (fp16 A operand of PV) goes to this set's own 64-column buffer in two halves:
    // keys 0-63 -> columns [0,32) -> barrier half 0, keys 64-127 -> columns [32,64) -> half 1
#pragma unroll
    for (int h = 0; h < 2; h++) {
        exp_half(s + 64 * h, pk, scale2, neg2, sum_a, sum_b);
        FA_PROBE(4 + 3 * h)
        if (h == 0 && gk > 0) {
            // the PV that consumed this buffer's previous contents has retired (long ago, normally)
            mbar_wait(c.bar_pv_done_mine, (gk - 1u) & 1u, 41);
            tc_fence_after();
        }
        FA_PROBE(5 + 3 * h)
        tmem_st_x32(c.tP + 32 * h, pk);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) {
            arrive_leader<CG>(c.bar_p_full + 8 * h);   // one arrival per warp
#if FA_ANTIPHASE
            if (h == 0) mbar_arrive(c.bar_m_ready_mine);
#endif
        }
        FA_PROBE(6 + 3 * h)
    }
#endif
    float a0, a1;
    unpack_f32x2(add_f32x2(sum_a, sum_b), a0, a1);
    l_run += a0 + a1;
}

template <int D, int CG>
__global__ void __launch_bounds__(kNumThreads, 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const Params p) {
    using C = Cfg<D, CG>;
    // SWIZZLE_128B tiles need 1024-byte alignment; no static shared memory is declared, so the
    // dynamic window starts at the CTA's (1024-aligned) shared base
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    const uint32_t sQ = smem_base;
    const uint32_t sK = smem_base + C::kOffK;
    const uint32_t sV = smem_base + C::kOffV;
    const uint32_t bars = smem_base + C::kOffBar;
    auto bar = [&](int slot) -> uint32_t { return bars + 8u * (uint32_t)slot; };
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + C::kOffMisc);
    volatile int* sched_w = reinterpret_cast<volatile int*>(smem_raw + C::kOffMisc + 8);   // [2]
    float* mref = reinterpret_cast<float*>(smem_raw + C::kOffMref);                        // [2][128]
    float2* fin = reinterpret_cast<float2*>(smem_raw + C::kOffFin);                        // [2][2][128]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool leader = rank == 0u;
    // address of the leader CTA's copy of a barrier / mailbox word of this CTA
    const uint32_t lead_delta = (CG == 2) ? (mapa_shared(smem_base, 0u) - smem_base) : 0u;
    auto lbar = [&](int slot) -> uint32_t { return bar(slot) + lead_delta; };
#ifdef FA_TIMING
    long long k_c0 = 0;
    unsigned long long k_t0 = 0;
    if (threadIdx.x == 0) {
        k_c0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(k_t0));
    }
#endif

    if (threadIdx.x == 0) {
        if ((smem_base & 1023u) != 0u && atomicExch(&g_watchdog[0], 1u) == 0u) {
            g_watchdog[1] = 99u;   // misaligned dynamic shared memory: results are garbage, the host reports it
            g_watchdog[2] = blockIdx.x;
            g_watchdog[3] = smem_base;
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(bar(C::kBarQFull + i), 1);
            mbar_init(bar(C::kBarQEmpty + i), 2);                  // last QK^T retired + O tile stored
            mbar_init(bar(C::kBarSFull + i), 1);
            mbar_init(bar(C::kBarSFree + i), 4 * CG);              // one arrival per softmax warp of the set
            mbar_init(bar(C::kBarPFull + 2 * i), 4 * CG);
            mbar_init(bar(C::kBarPFull + 2 * i + 1), 4 * CG);
            mbar_init(bar(C::kBarPvDone + i), 1);
            mbar_init(bar(C::kBarSchedFull + i), 1);
            mbar_init(bar(C::kBarSchedEmpty + i), C::kSchedConsumers);
        }
        for (int i = 0; i < C::kKStages; i++) {
            mbar_init(bar(C::kBarKFull + i), 1);
            mbar_init(bar(C::kBarKEmpty + i), 1);
        }
        for (int i = 0; i < C::kVStages; i++) {
            mbar_init(bar(C::kBarVFull + i), 1);
            mbar_init(bar(C::kBarVEmpty + i), 1);
        }
        mbar_init(bar(C::kBarOFree), 8 * CG);
        mbar_init(bar(C::kBarOStaged), 8);
        mbar_init(bar(C::kBarOStaged + 1), 8);
        for (int i = 0; i < 8; i++) mbar_init(bar(C::kBarMReady + i), 1);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) {
        if (CG == 1) {
            tmem_alloc(smem_u32(tmem_slot_ptr), kTmemCols);
            tmem_relinquish();
        } else {
            tmem_alloc_2sm(smem_u32(tmem_slot_ptr), kTmemCols);
            tmem_relinquish_2sm();
        }
    }
    if (warp == kLoadWarp && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        tma_prefetch_desc(&tmO);
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) {   // the peer's barriers must be initialised before anything arrives on them
        cluster_arrive_release();
        cluster_wait_acquire();
    }
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor
    // prefetch) may overlap the tail of the previous kernel in the stream; global memory is only
    // touched below this point.  The next kernel's prologue may start as soon as our CTAs retire.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // Dynamic tile scheduler (replaces the reference's static blockIdx mapping, FA.cu:103-112): the
    // leader's producer warp claims work indices (first one static, the rest from a global counter)
    // and publishes them through a 2-slot smem mailbox (in both CTAs of a pair); every consumer warp
    // reads slot i&1 for its i-th unit.  Returns -1 when the grid has run out of work.
    auto next_work = [&](uint32_t i) -> int {
        const uint32_t slot = i & 1u;
        if (CG == 1) mbar_wait(bar(C::kBarSchedFull + slot), (i >> 1) & 1u, 50);
        else mbar_wait_cluster(bar(C::kBarSchedFull + slot), (i >> 1) & 1u, 50);
        const int w = sched_w[slot];
        __syncwarp();
        if (lane == 0) arrive_leader<CG>(lbar(C::kBarSchedEmpty + slot));
        return w;
    };
    // barriers the peer CTA arrives on too; what they guard travels through TMEM / the async proxy
    auto wait_lead = [&](uint32_t b, uint32_t parity, int tag) { mbar_wait(b, parity, tag); };

    // The producer and MMA warps run their loops converged (all 32 lanes take the same branches
    // and waits); the instructions with side effects sit under elect_one().  Warp-uniform control
    // flow keeps descriptors and barrier addresses in uniform registers.
    if (warp >= 8) {
    setmaxnreg_dec<kRegsOther>();   // each role's code must be dominated by its own setmaxnreg
    if (warp == kLoadWarp) {
        // =============================== TMA producer (Q, K, V) + scheduler ===============================
        Ring rk{0u, 0u}, rv{0u, 0u};
        const uint32_t num_units = gridDim.x / CG;     // CTAs (pairs) in flight
        const uint32_t unit_id = blockIdx.x / CG;
        for (uint32_t it = 0;; ++it) {
            const uint32_t slot = it & 1u;
            int w = 0;
            if (leader) {
                // claim the next work unit and publish it
                wait_lead(bar(C::kBarSchedEmpty + slot), ((it >> 1) & 1u) ^ 1u, 3);
                if (lane == 0) w = (it == 0) ? (int)unit_id : (int)num_units + atomicAdd(p.sched, 1);
                w = __shfl_sync(0xffffffffu, w, 0);
                if (w >= p.total_work) w = -1;
                if (lane == 0) {
                    sched_w[slot] = w;
                    mbar_arrive(bar(C::kBarSchedFull + slot));   // release: the slot write is visible to waiters
                    if (CG == 2) {
                        st_shared_cluster_u32(mapa_shared(smem_u32(const_cast<int*>(sched_w + slot)), 1u), (uint32_t)w);
                        mbar_arrive_cluster_release(mapa_shared(bar(C::kBarSchedFull + slot), 1u));
                    }
                }
                __syncwarp();
            } else {
                w = next_work(it);
            }
            if (w < 0) break;
            const WorkItem wi = decode_work(w, p);
            const uint32_t qslot = it & 1u;
            auto load = [&](uint32_t dst, const CUtensorMap* tm, uint32_t full, int c0, int c1) {
                if (CG == 1) tma_load_3d(dst, tm, full, c0, c1, wi.bh);
                else tma_load_3d_2sm(dst, tm, full, c0, c1, wi.bh);
            };
            // Q tile of this CTA: the slot is free once the QK^T MMAs of the unit two back have retired
            // and its O tile (staged in the same buffer) has been read by the TMA store
            mbar_wait(bar(C::kBarQEmpty + qslot), ((it >> 1) & 1u) ^ 1u, 1);
            if (elect_one()) {
                if (leader) mbar_arrive_expect_tx(bar(C::kBarQFull + qslot), CG * C::kQTileBytes);
#pragma unroll
                for (int pn = 0; pn < C::kPanels; pn++)
                    load(sQ + qslot * C::kQTileBytes + pn * C::kQPanelBytes, &tmQ, lbar(C::kBarQFull + qslot), pn * 64,
                         wi.q0 + (int)rank * kBlockM);
            }
            __syncwarp();
            for (int j = 0; j < wi.n; j++) {
                // K_j: this CTA's kKRows keys
                mbar_wait(bar(C::kBarKEmpty + rk.idx), rk.phase ^ 1u, 2);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(bar(C::kBarKFull + rk.idx), CG * C::kKBytes);
#pragma unroll
                    for (int pn = 0; pn < C::kPanels; pn++)
                        load(sK + rk.idx * C::kKBytes + pn * C::kKPanelBytes, &tmK, lbar(C::kBarKFull + rk.idx), pn * 64,
                             j * kBlockN + (int)rank * C::kKRows);
                }
                __syncwarp();
                rk.advance<C::kKStages>();
                // V_j: this CTA's D/CG columns.  K and V have separate rings and separate consumers
                // (the two issuer warps), so this wait can only delay later loads, never deadlock.
                mbar_wait(bar(C::kBarVEmpty + rv.idx), rv.phase ^ 1u, 4);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(bar(C::kBarVFull + rv.idx), CG * C::kVBytes);
#pragma unroll
                    for (int pn = 0; pn < C::kVPanels; pn++)
                        load(sV + rv.idx * C::kVBytes + pn * C::kVPanelBytes, &tmV, lbar(C::kBarVFull + rv.idx),
                             (int)rank * (D / CG) + pn * 64, j * kBlockN);
                }
                __syncwarp();
                rv.advance<C::kVStages>();
            }
        }
    } else if (warp == kMmaWarp && leader) {
        // =============================== tcgen05.mma issuer: S_b = Q K_j^T ===============================
        // Runs as far ahead as the two S buffers allow: QK(g) goes out as soon as K_g has landed and
        // the softmax set of tile g-2 has pulled S(g-2) into registers.
        Ring rk{0u, 0u};
        uint32_t gq = 0;                                   // global tile index
        const uint64_t qdesc0 = umma_smem_desc(sQ, 16, 1024);
        const uint64_t kdesc0 = umma_smem_desc(sK, 16, 1024);
        for (uint32_t it = 0;; ++it) {
            const int w = next_work(it);
            if (w < 0) break;
            const WorkItem wi = decode_work(w, p);
            const uint32_t qslot = it & 1u;
            mbar_wait(bar(C::kBarQFull + qslot), (it >> 1) & 1u, 10);
            const uint64_t qdesc = qdesc0 + (uint64_t)((qslot * C::kQTileBytes) >> 4);
            if (wi.n == 0) {                               // nothing will read this Q tile
                if (elect_one()) {
                    if (CG == 1) umma_commit(bar(C::kBarQEmpty + qslot));
                    else umma_commit_2sm(bar(C::kBarQEmpty + qslot));
                }
                __syncwarp();
            }
            for (int j = 0; j < wi.n; j++) {
                const uint32_t b = gq & 1u, k = gq >> 1;
                mbar_wait(bar(C::kBarKFull + rk.idx), rk.phase, 11);
                if (k > 0) mbar_wait(bar(C::kBarSFree + b), (k - 1u) & 1u, 12);   // the set has read S_b(previous)
                tc_fence_after();
                const uint32_t tS = tmem_base + C::kTmemS + 128u * b;
                const uint64_t kdesc = kdesc0 + (uint64_t)((rk.idx * C::kKBytes) >> 4);
                if (elect_one()) {
                    // D/16 k-steps; k-step ks lives in panel ks/4 at byte offset (ks%4)*32
#pragma unroll
                    for (int ks = 0; ks < D / 16; ks++) {
                        const uint64_t qoff = (uint64_t)(((ks >> 2) * C::kQPanelBytes + (ks & 3) * 32) >> 4);
                        const uint64_t koff = (uint64_t)(((ks >> 2) * C::kKPanelBytes + (ks & 3) * 32) >> 4);
                        if (CG == 1) umma_ss(tS, qdesc + qoff, kdesc + koff, C::kIdescQK, ks > 0 ? 1u : 0u);
                        else umma_ss_2sm(tS, qdesc + qoff, kdesc + koff, C::kIdescQK, ks > 0 ? 1u : 0u);
                    }
                    if (CG == 1) {
                        umma_commit(bar(C::kBarSFull + b));
                        umma_commit(bar(C::kBarKEmpty + rk.idx));
                        if (j == wi.n - 1) umma_commit(bar(C::kBarQEmpty + qslot));   // last reader of this Q tile
                    } else {
                        umma_commit_2sm(bar(C::kBarSFull + b));
                        umma_commit_2sm(bar(C::kBarKEmpty + rk.idx));
                        if (j == wi.n - 1) umma_commit_2sm(bar(C::kBarQEmpty + qslot));
                    }
                }
                __syncwarp();
                rk.advance<C::kKStages>();
                ++gq;
            }
        }
    } else if (warp == kPvWarp && leader) {
        // =============================== tcgen05.mma issuer: O (+)= P_b V_j ===============================
        // 8 k-steps of 16 kv rows (P k-step = 8 TMEM columns, V k-step = 16 rows * 128 B), issued in two
        // halves of 4 k-steps as the two halves of P arrive
        Ring rv{0u, 0u};
        uint32_t gp = 0;                                   // global tile index
        uint32_t nz_units = 0;                             // units with tiles whose PVs have all been issued
        const uint64_t vdesc0 = umma_smem_desc(sV, C::kVPanelBytes, 1024);
        const uint32_t tO = tmem_base + C::kTmemO;
        for (uint32_t it = 0;; ++it) {
            const int w = next_work(it);
            if (w < 0) break;
            const WorkItem wi = decode_work(w, p);
            for (int j = 0; j < wi.n; j++) {
                const uint32_t b = gp & 1u, k = gp >> 1;
                mbar_wait(bar(C::kBarVFull + rv.idx), rv.phase, 13);
                if (j == 0 && nz_units > 0) mbar_wait(bar(C::kBarOFree), (nz_units - 1u) & 1u, 14);   // epilogue has read O
                const uint32_t tP = tmem_base + C::kTmemP + 64u * b;
                const uint64_t vdesc = vdesc0 + (uint64_t)((rv.idx * C::kVBytes) >> 4);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    mbar_wait(bar(C::kBarPFull + 2 * b + h), k & 1u, 15 + h);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 4 * h; ks < 4 * h + 4; ks++) {
                            const uint64_t voff = (uint64_t)((ks * 16 * 128) >> 4);
                            const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
                            if (CG == 1) umma_ts(tO, tP + ks * 8, vdesc + voff, C::kIdescPV, acc);
                            else umma_ts_2sm(tO, tP + ks * 8, vdesc + voff, C::kIdescPV, acc);
                        }
                        if (h == 1) {
                            if (CG == 1) {
                                umma_commit(bar(C::kBarPvDone + b));
                                umma_commit(bar(C::kBarVEmpty + rv.idx));
                            } else {
                                umma_commit_2sm(bar(C::kBarPvDone + b));
                                umma_commit_2sm(bar(C::kBarVEmpty + rv.idx));
                            }
                        }
                    }
                    __syncwarp();
                }
                rv.advance<C::kVStages>();
                ++gp;
            }
            if (wi.n > 0) ++nz_units;
        }
    } else if (warp == kStoreWarp) {
        // =============================== O tile store ===============================
        // Softmax warps stage O_t / l as fp16 in the (now idle) Q buffer of the unit, 128B-swizzled;
        // this warp hands it to TMA, which clips rows past Nq, and then returns the buffer.
        uint32_t staged_parity = 0u;
        for (uint32_t it = 0;; ++it) {
            const int w = next_work(it);
            if (w < 0) break;
            const WorkItem wi = decode_work(w, p);
            const uint32_t qslot = it & 1u;
            const int q_start = wi.q0 + (int)rank * kBlockM;
            // one barrier per Q slot, waited only for units with tiles: it cannot complete twice before this
            // warp has seen the first completion (the slot's next use needs the Q load this warp releases)
            if (wi.n > 0) {
                mbar_wait(bar(C::kBarOStaged + qslot), (staged_parity >> qslot) & 1u, 60);
                staged_parity ^= 1u << qslot;
            }
            const bool do_store = !p.partial_mode && wi.n > 0 && q_start < p.Nq;
            if (lane == 0) {   // one fixed lane: bulk-group state is per thread
                if (do_store) {
#pragma unroll
                    for (int pn = 0; pn < C::kPanels; pn++)
                        tma_store_3d(&tmO, sQ + qslot * C::kQTileBytes + pn * C::kQPanelBytes, pn * 64, q_start, wi.bh);
                    tma_store_commit();
                    tma_store_wait_read<0>();
                }
                mbar_arrive(bar(C::kBarQEmpty + qslot));
            }
            __syncwarp();
        }
        if (lane == 0) tma_store_wait_all<0>();
        __syncwarp();
    }
    } else {
        setmaxnreg_inc<kRegsSoftmax>();
        // =============================== softmax / correction / epilogue ===============================
        const int set = warp >> 2;                             // 0: even global tiles, 1: odd
        const int quad = warp & 3;
        const int row_in_tile = quad * 32 + lane;              // TMEM lane == S/O row
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        SoftmaxCtx c;
        c.tS = tmem_base + lane_base + C::kTmemS + 128u * set;
        c.tP = tmem_base + lane_base + C::kTmemP + 64u * set;
        c.tO = tmem_base + lane_base + C::kTmemO;
        c.bar_s_free = lbar(C::kBarSFree + set);
        c.bar_p_full = lbar(C::kBarPFull + 2 * set);
        c.bar_pv_done_mine = bar(C::kBarPvDone + set);
        c.bar_pv_done_other = bar(C::kBarPvDone + (set ^ 1));
        c.bar_m_ready_mine = bar(C::kBarMReady + 4 * set + quad);
        c.bar_m_ready_other = bar(C::kBarMReady + 4 * (set ^ 1) + quad);
        c.mref_mine = mref + set * kBlockM + row_in_tile;
        c.mref_other = mref + (set ^ 1) * kBlockM + row_in_tile;
        const uint32_t bar_s_full = bar(C::kBarSFull + set);
        uint32_t g0 = 0;       // global index of the unit's first tile
        uint32_t nz_units = 0;

        for (uint32_t it = 0;; ++it) {
            const int w = next_work(it);
            if (w < 0) break;
            const WorkItem wi = decode_work(w, p);
            const int q_start = wi.q0 + (int)rank * kBlockM;
            const int row = q_start + row_in_tile;             // local query row
            // keys [0, lim) are visible to this row
            long long lim_ll = p.causal ? (long long)row + p.shift + 1 : (long long)p.Nkv;
            if (lim_ll > p.Nkv) lim_ll = p.Nkv;
            if (lim_ll < 0) lim_ll = 0;
            const int lim = (int)lim_ll;

            float m_ref = -INFINITY, l_run = 0.f;
            for (int j = (int)((g0 ^ (uint32_t)set) & 1u); j < wi.n; j += 2) {   // this set's tiles: (g0 + j) & 1 == set
                const uint32_t g = g0 + (uint32_t)j;
                const uint32_t gk = g >> 1;
#ifdef FA_TIMING
                const long long tw0 = clock64();
#endif
                // Anti-phase: this set starts tile g only after the other set is half-way through tile
                // g-1 (its m_ready arrival: reference max published, first half of P delivered).  Left
                // alone the two sets run in lock-step -- both in the MUFU-bound exponential phase at
                // the same time on the same scheduler, both idle-waiting at the same time -- which
                // costs ~25 % (profiles/r01_v9d_*).  Half a tile apart, one set's exponentials cover
                // the other's TMEM loads, row max and barrier round trips.
#if FA_ANTIPHASE == 1
                if (g > 0) mbar_wait(c.bar_m_ready_other, ((g - 1u) >> 1) & 1u, 24);
#endif
                mbar_wait(bar_s_full, gk & 1u, 20 + set);
                tc_fence_after();
#ifdef FA_TIMING
                const long long tw1 = clock64();
#endif
                const int k0 = j * kBlockN;
                const bool need_mask = (k0 + kBlockN > p.Nkv) || (p.causal && k0 + kBlockN - 1 > q_start + p.shift);
                const uint32_t g_prev_k = (g - 1u) >> 1;       // only used when j > 0
                if (need_mask)
                    softmax_tile<D, CG, true>(p, c, lim - k0, j == 0, gk, g_prev_k, m_ref, l_run);
                else
                    softmax_tile<D, CG, false>(p, c, kBlockN, j == 0, gk, g_prev_k, m_ref, l_run);
#ifdef FA_TIMING
                if (lane == 0 && quad == 0 && j > 1 && (j & 7) < 2) {   // sampled
                    const long long tw2 = clock64();
                    atomicAdd(&g_timing_pair[set * 3 + 0], (unsigned long long)(tw1 - tw0));
                    atomicAdd(&g_timing_pair[set * 3 + 1], (unsigned long long)(tw2 - tw1));
                    atomicAdd(&g_timing_pair[set * 3 + 2], 1ull);
                }
#endif
            }

            // ---- epilogue: both sets share it; set s takes O columns [s*D/2, (s+1)*D/2) ----
            const bool have = wi.n > 0;
            const bool row_ok = row < p.Nq;
            const size_t grow = (size_t)wi.bh * p.Nq + (row_ok ? row : 0);
            constexpr int kHalf = D / 2;
            const int col0 = set * kHalf;
            float m_old = -FLT_MAX, l_old = 0.f;
            if (p.partial_mode && p.accumulate && row_ok) {     // read before set 0 overwrites them below
                m_old = p.ml[grow * 2 + 0];
                l_old = p.ml[grow * 2 + 1];
            }
            float m_fin = -INFINITY, l_tot = 0.f;
            if (have) {
                const uint32_t g_last = g0 + (uint32_t)wi.n - 1u;
                mbar_wait(bar(C::kBarPvDone + (g_last & 1u)), (g_last >> 1) & 1u, 30 + set);
                tc_fence_after();
                // merge the two sets' partial row sums (each relative to the reference max its set last saw)
                float2* my_fin = fin + ((it & 1u) * 2 + set) * kBlockM + row_in_tile;
                const float2* ot_fin = fin + ((it & 1u) * 2 + (set ^ 1)) * kBlockM + row_in_tile;
                reinterpret_cast<volatile float*>(my_fin)[0] = m_ref;
                reinterpret_cast<volatile float*>(my_fin)[1] = l_run;
                bar_sync(1, 256);
                float2 ot;
                ot.x = reinterpret_cast<const volatile float*>(ot_fin)[0];
                ot.y = reinterpret_cast<const volatile float*>(ot_fin)[1];
                m_fin = fmaxf(m_ref, ot.x);
                const float a_me = (m_ref == -INFINITY) ? 0.f : ex2_approx((m_ref - m_fin) * p.scale_log2);
                const float a_ot = (ot.x == -INFINITY) ? 0.f : ex2_approx((ot.x - m_fin) * p.scale_log2);
                l_tot = l_run * a_me + ot.y * a_ot;
            }
            uint32_t o[kHalf];
            if (have) {
#pragma unroll
                for (int cc = 0; cc < kHalf; cc += 32) tmem_ld_x32(c.tO + col0 + cc, o + cc);
                tmem_wait_ld();
                // O is in registers: the first PV of the next unit may overwrite the accumulator
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_leader<CG>(lbar(C::kBarOFree));
            } else {
#pragma unroll
                for (int i = 0; i < kHalf; i++) o[i] = 0u;
            }
            if (!p.partial_mode) {
                const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;   // FA.cu:502-503
                if (have) {
                    // fp16 row half -> staging tile (the unit's Q buffer), 16-byte chunks XOR-swizzled by row & 7
                    const uint32_t stage = sQ + (it & 1u) * C::kQTileBytes;
                    constexpr int kChunks = kHalf / 8;          // 16-byte chunks this thread writes
#pragma unroll
                    for (int ch = 0; ch < kChunks; ch++) {
                        const int col = col0 + ch * 8;          // first of 8 output columns
                        const int panel = col >> 6;
                        const int chunk = (col & 63) >> 3;
                        const uint32_t addr = stage + panel * C::kQPanelBytes + row_in_tile * 128 +
                                              ((chunk ^ (row_in_tile & 7)) << 4);
                        const uint32_t* oo = o + ch * 8;
                        st_shared_v4(addr,
                                     pack_half2(__uint_as_float(oo[0]) * inv, __uint_as_float(oo[1]) * inv),
                                     pack_half2(__uint_as_float(oo[2]) * inv, __uint_as_float(oo[3]) * inv),
                                     pack_half2(__uint_as_float(oo[4]) * inv, __uint_as_float(oo[5]) * inv),
                                     pack_half2(__uint_as_float(oo[6]) * inv, __uint_as_float(oo[7]) * inv));
                    }
                    fence_proxy_async_smem();
                } else if (row_ok) {
                    // no key visible to any row of the unit: zeros, straight to global
                    __half* orow = p.o + grow * D + col0;
#pragma unroll
                    for (int i = 0; i < kHalf; i += 8) *reinterpret_cast<uint4*>(orow + i) = make_uint4(0u, 0u, 0u, 0u);
                }
            } else {
                // partial state (FA.cu:460-496): un-normalised fp32 O, (m, l) with m in the
                // scaled-score (natural-log) domain; merge algebra of FA.cu:575-597 when accumulating
                float m_out = (m_fin == -INFINITY) ? -FLT_MAX : m_fin * p.scale;
                float l_out = l_tot;
                float w_new = 1.f, w_old = 0.f;
                if (p.accumulate) {
                    const float m_max = fmaxf(m_old, m_out);
                    const float kLog2e = 1.4426950408889634f;
                    w_old = (m_old <= -FLT_MAX) ? 0.f : ex2_approx((m_old - m_max) * kLog2e);
                    w_new = (m_out <= -FLT_MAX) ? 0.f : ex2_approx((m_out - m_max) * kLog2e);
                    l_out = l_old * w_old + l_tot * w_new;
                    m_out = m_max;
                }
                if (row_ok) {
                    float* prow = p.o_partial + grow * D + col0;
#pragma unroll
                    for (int i = 0; i < kHalf; i += 4) {
                        float4 v = make_float4(__uint_as_float(o[i]) * w_new, __uint_as_float(o[i + 1]) * w_new,
                                               __uint_as_float(o[i + 2]) * w_new, __uint_as_float(o[i + 3]) * w_new);
                        if (p.accumulate) {
                            const float4 old = *reinterpret_cast<const float4*>(prow + i);
                            v.x += old.x * w_old; v.y += old.y * w_old;
                            v.z += old.z * w_old; v.w += old.w * w_old;
                        }
                        *reinterpret_cast<float4*>(prow + i) = v;
                    }
                    if (set == 0) {
                        // both sets have read the old (m, l) before the fin exchange barrier above when
                        // the unit has tiles; without tiles nothing changes (w_new = 0 or l = 0)
                        if (have || !p.accumulate) {
                            p.ml[grow * 2 + 0] = m_out;
                            p.ml[grow * 2 + 1] = l_out;
                        }
                    }
                }
            }
            __syncwarp();
            if (have && lane == 0) mbar_arrive(bar(C::kBarOStaged + (it & 1u)));
            if (have) ++nz_units;
            g0 += (uint32_t)wi.n;
        }
        (void)nz_units;
    }

    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (CG == 2) {   // no CTA of the pair may exit while the other can still signal it or read its smem/TMEM
        cluster_arrive_release();
        cluster_wait_acquire();
    }
#ifdef FA_TIMING
    if (threadIdx.x == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        atomicAdd(&g_timing_pair[20], (unsigned long long)(clock64() - k_c0));   // CTA lifetime, SM cycles
        atomicAdd(&g_timing_pair[21], t1 - k_t0);                                // CTA lifetime, ns
        atomicAdd(&g_timing_pair[22], 1ull);
    }
#endif
    if (threadIdx.x == 0) {
        // last CTA out re-arms the scheduler state for the launch that reuses this slot
        __threadfence();
        if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
            p.sched[0] = 0;
            p.sched[1] = 0;
            __threadfence();
        }
    }
    if (warp == kMmaWarp) {
        __syncwarp();
        tc_fence_after();
        if (CG == 1) tmem_dealloc(tmem_base, kTmemCols);
        else tmem_dealloc_2sm(tmem_base, kTmemCols);
    }
}

}  // namespace fa_pair
