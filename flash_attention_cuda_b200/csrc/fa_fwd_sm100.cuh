// fa_fwd_sm100.cuh -- FlashAttention forward for B200 (sm_100a): persistent, warp-specialised,
// TMA -> 128B-swizzled smem -> tcgen05.mma with S/P/O in tensor memory.
//
// Replaces the reference's device path flash_attention_v9<...> (flash_attention.cu:67-554):
//   FA.cu:103-112  block->(bh, q-block) mapping, GRID_SWAP    -> persistent work loop, heavy-first
//   FA.cu:145-159  Q fragments in registers                   -> Q tile pair resident in smem (TMA), two slots
//   FA.cu:417-447  synchronous K/V tile load + 2 barriers     -> producer warp, 3-entry mbarrier ring
//   FA.cu:188-233  DO_QK_MATMUL (mma.sync m16n8k16)           -> tcgen05.mma SS, S in TMEM
//   FA.cu:235-288  DO_SOFTMAX (quad shuffles, eager rescale)  -> one thread per row, lazy rescale
//   FA.cu:290-334  DO_PV_MATMUL (P in registers)              -> P fp16 in TMEM, tcgen05.mma TS
//   FA.cu:497-553  multi-pass smem output staging             -> TMEM -> registers -> swizzled smem -> TMA store
//   FA.cu:460-496  split-K partial epilogue (dead code there) -> partial mode used by ring CP
//
// CTA = 384 threads:  warps 0-3  softmax/correction/epilogue for Q tile 0 (128 rows)
//                     warps 4-7  same for Q tile 1
//                     warp 8     TMEM allocator + tcgen05.mma issuer (one lane)
//                     warp 9     TMA producer (one lane)
//                     warp 10    TMA store of finished O tiles
//                     warp 11    idle (pads the third warpgroup)
// One CTA per SM; each CTA loops over work items (bh, pair of 128-row Q tiles).
//
// Shared memory (D=128): Q 2 slots x 2 tiles x 32 KB, K/V ring 3 x 32 KB.  Items alternate between the
// Q slots: the next item's Q pair is loaded while the current one runs, and its first QK^T is issued
// right behind the current item's last PV (the softmax warps are then still in the epilogue), so the
// tensor pipe does not drain at item boundaries -- worth 2 % at N=8192, 10-19 % at N<=1024.  After its
// last QK^T a slot stages the item's O tiles (fp16, 128B-swizzled) for the TMA store.
//
// TMEM (512 columns x 128 lanes x 32 bit): S0 [0,128) S1 [128,256) O0 [256,256+D) O1 [256+D,256+2D).
// P_t (fp16, two per column) overwrites columns [0,64) of S_t once the owning thread has read
// its S row, and is handed to the MMA warp in two halves (keys 0-63, 64-127) so that the first four
// PV k-steps run while the second half of the exponentials is still being computed.
// Tensor-pipe order per KV tile j:  PV0(j) QK0(j+1) PV1(j) QK1(j+1), so each softmax warpgroup works
// on S_t(j+1) while the tensor core runs the other tile's two MMAs.
//
// Why not 64-wide KV sub-tiles with S double-buffered (tried, profiles/r01_v3_subtile_*): QK^T with
// N=64 in SS mode re-reads the Q slice from shared memory for half the math (192 B/clk > the 128 B/clk
// smem port).  N=128 sits exactly at the port limit, so the S->P->PV->QK chain is shortened instead.
//
// Build-time switches (all off in the product build; each one is an experiment recorded under profiles/):
//   -DFA_TIMING        in-kernel clock64 probes (tests/harness/timing.py)
//   -DFA_SKELETON / -DFA_NO_EXP   exponentials replaced by the identity (what the tensor side alone can do)
//   -DFA_SUM_GUARD     softmax without a row max, guarded by the row sum (r01_v4c_sumguard_experiment.txt)
//   -DFA_SCHED_FENCE   data-dependency fence that makes ptxas store the first piece of P before the second
//                      piece's exponentials (r01_accumulate_race.txt)
//   -DFA_STREAM        streamed softmax (softmax_tile_stream: S read in chunks, exponentials against the row's current
//                      reference, lazy rescale at the end of the tile): parity-clean, 1 % slower (r02_stream_experiment.txt)
//   -DFA_STREAM_S      second half of S re-read from TMEM instead of held in registers (r01_softmax_schedule.txt)
//   -DFA_P_PARTS=3     P in three pieces;  -DFA_REGS_SOFTMAX / -DFA_REGS_OTHER  setmaxnreg budgets
//                      (r01_v4b_defer_group_ab.log)
//   -DFA_EPI_WG=1      epilogue warpgroup (512 threads; r02_c14_*: parity-clean, 4-8 % slower)
//   -DFA_NO_WAIT_IN_VARIANT   the wait for S in front of the masked / unmasked branch instead of inside each (r02_c15_*)
//   -DFA_NO_SCALE_REG / -DFA_NO_PIN_ADDR   scale, S address and p_full barrier address re-materialised per tile (r02_c21_*, r02_c22_*)
//   -DFA_SPEC          piece 0's exponentials against the row's current reference, vote afterwards (softmax_tile_spec;
//                      parity-clean, 0 to -4 %: r02_c21_*)
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "sm100_ptx.cuh"

namespace fa {

using namespace sm100;

constexpr int kBlockM = 128;      // Q rows per tile  (UMMA M)
constexpr int kBlockN = 128;      // K/V rows per tile (UMMA N of QK^T, K extent of PV)
constexpr int kMmaWarp = 8;
constexpr int kLoadWarp = 9;
constexpr int kStoreWarp = 10;
constexpr int kTmemCols = 512;
// Epilogue warpgroup (experiment, -DFA_EPI_WG=1; off in the product build): warps 12-15 drain a finished O tile from TMEM
// (O / l -> fp16 -> swizzled smem for the TMA store) while the softmax warps are already on the next work item, so an
// item's epilogue (~2000 cycles: wait for the last PV, 128 columns through registers) leaves the softmax warps' chain.
// Parity-clean on the whole GPU suite and 4-8 % SLOWER at D=128 when first built (profiles/r02_c14_*); on the final kernel
// level at N=8192 without a mask, -1.2 % with one, -5 % at causal N=2048 (profiles/r02_c33_*; its warps sleeping between
// polls, -DFA_EPI_SLEEP=ns, changes nothing).  With 512 threads the register file only allows 200/72/40 or 208/56/40
// registers for softmax / MMA+producer / epilogue warps (spills in the softmax or the MMA warp), and what the softmax warps
// gain is small because the next item's first S only exists ~1000 cycles after the item's last PV anyway.
#ifndef FA_EPI_WG
#define FA_EPI_WG 0
#endif
constexpr bool kEpiWg = FA_EPI_WG != 0;
constexpr int kEpiWarp0 = 12;
constexpr int kNumThreads = kEpiWg ? 512 : 384;
#ifndef FA_REGS_SOFTMAX
#define FA_REGS_SOFTMAX (FA_EPI_WG ? 200 : 208)
#endif
#ifndef FA_REGS_OTHER
#define FA_REGS_OTHER (FA_EPI_WG ? 56 : 80)
#endif
#ifndef FA_REGS_EPI
#define FA_REGS_EPI 56
#endif
constexpr int kRegsSoftmax = FA_REGS_SOFTMAX;   // setmaxnreg: softmax warpgroups grow, the other warpgroups shrink
constexpr int kRegsOther = FA_REGS_OTHER;       // without the epilogue warpgroup: 2*128*208 + 128*80 = 63488 <= 168 (launch) * 384
constexpr int kRegsEpi = FA_REGS_EPI;           // with it: 2*128*200 + 128*56 + 128*56 = 65536 = 128 (launch) * 512
static_assert(2 * 128 * kRegsSoftmax + 128 * kRegsOther + (kEpiWg ? 128 * kRegsEpi : 0) <= 65536, "register file");
// P_t reaches the MMA warp in kPParts pieces (k-steps of 16 keys: [0,4) [4,8) or [0,4) [4,6) [6,8)): the PV k-steps of
// a piece run under the exponentials of the next one, and only the last piece's k-steps sit between the end of the
// softmax and the next QK^T of the tile.
#ifndef FA_P_PARTS
#define FA_P_PARTS 2
#endif
constexpr int kPParts = FA_P_PARTS;
static_assert(kPParts == 2 || kPParts == 3, "P is delivered in 2 or 3 pieces");
__host__ __device__ constexpr int p_part_ks(int part) {   // first k-step of a piece
    return kPParts == 2 ? part * 4 : (part == 0 ? 0 : part == 1 ? 4 : part == 2 ? 6 : 8);
}
constexpr float kRescaleThreshold = 8.0f;   // lazy rescale: tolerate P up to 2^8 before moving the reference max
// Of every 4 element pairs, this many take exp2 on the FMA pipe (Cody-Waite + degree-3 minimax)
// instead of MUFU.EX2: at 16 MUFU/clk/SM the 16384 exponentials of a 128x128 tile cost as many
// cycles as its two MMAs, so the SFU -- not the tensor core -- would set the pace.
// Template parameter kPoly of the kernel: of every 4 element pairs, this many take exp2 on the FMA pipe.
// 1 pays off only when the loop runs long enough for MUFU throughput to matter (D=128, N >= 4096: +1.4 %);
// for shorter sequences and D=64 the extra ~140 instructions per tile cost more than the 32 MUFU they
// save (causal N=2048: 722 vs 664 TFLOPS, D=64: 689 vs 676 -- profiles/r01_v4b_poly_ab.log), and 2 loses everywhere.

template <int D>
struct Cfg {
    static_assert(D == 64 || D == 128, "head_dim must be 64 or 128");
    static constexpr int kPanels = D / 64;                 // 128-byte swizzle panels per row
    static constexpr int kPanelBytes = 128 * 128;          // 128 rows x 128 B
    static constexpr int kTileBytes = kPanels * kPanelBytes;
    // K/V ring entries (one tile each).  Depth 3, 4 and 5 measure the same (profiles/r01_v4b_*): the
    // ring only has to cover one TMA round trip, so the shared memory goes to a second Q buffer instead.
    static constexpr int kStages = (D == 128) ? 3 : 8;
    // Q: two slots (work items alternate) x two tiles.  While item i runs, item i+1's Q pair is already
    // resident, so its first QK^T goes out right behind item i's last PV; after item i's last QK^T its
    // slot doubles as the staging buffer of its O tiles until the TMA store has read them.
    static constexpr int kSmemQ = 4 * kTileBytes;
    static constexpr int kSmemKV = kStages * kTileBytes;
    static constexpr int kBarOffset = kSmemQ + kSmemKV;
    static constexpr int kNumBars = 4 + 2 * kStages + 14 + 4 + 4 + 4;
    static constexpr int kMlOffset = kBarOffset + kNumBars * 8 + 32;           // +32: tmem slot, scheduler slots
    // hand-over area: split mode (m, l) of tile slot 1, one pair per row; pair mode with the epilogue warpgroup: the row
    // sums l of the two tiles, one float per row
    static constexpr int kXchBytes = kBlockM * 8;
    // SWIZZLE_128B tiles need a 1024-byte aligned base; the dynamic window is that aligned on sm_100 in practice, the
    // slack (whatever is left of the 227 KB, at most 1 KB) covers a base that is not, and the kernel checks
    static constexpr int kSmemMax = 232448;
    static constexpr int kSmemBytes = (kMlOffset + kXchBytes + 1024 <= kSmemMax) ? kMlOffset + kXchBytes + 1024 : kSmemMax;
    static_assert(kMlOffset + kXchBytes <= kSmemMax, "exceeds the 227 KB opt-in shared memory of sm_100");
    static constexpr int kTmemS0 = 0, kTmemS1 = 128, kTmemO0 = 256, kTmemO1 = 256 + D;
    static constexpr uint32_t kIdescQK = umma_idesc_f16(kBlockM, kBlockN, 0, 0);
    static constexpr uint32_t kIdescPV = umma_idesc_f16(kBlockM, D, 0, 1);  // V is MN-major ([kv][d], d contiguous)
    // BF16 operands (Q, K, V and therefore P): A and B format fields [7,10) / [10,13) = 1
    static constexpr uint32_t kBf16Operands = (1u << 7) | (1u << 10);
};

struct Params {
    __half* o;          // fp16 (or bf16: kernel template) output [BH, Nq, D]   (partial_mode == 0)
    float* o_partial;   // fp32 un-normalised [BH*Nq, D]       (partial_mode == 1; FA.cu:460-496 format)
    float* ml;          // (m, l) pairs [BH*Nq, 2]
    int Nq, Nkv, BH;
    int causal;
    int shift;          // q_offset - kv_offset: key c visible to query r iff c <= r + shift
    int nqp;            // work items per head: Q tile pairs, ceil(Nq / 256) (split mode: Q tiles, ceil(Nq / 128))
    int split;          // 1: a work item is ONE Q tile and its KV tiles alternate between the CTA's two tile slots (below)
    int total_work;     // BH * nqp
    int group_heads;    // heads per scheduling group (their K/V working set is sized to stay in L2)
    int partial_mode;
    int accumulate;
    int zero;           // always 0: trip count of the empty loops that fence ptxas' instruction scheduler (FA_SCHED_FENCE)
    int* sched;         // {next work index, finished CTAs}: dynamic tile scheduler state, self-resetting
    // Gathered K/V (context parallelism, flash_attn_fwd_gathered): K/V rows [i * ready_rows, (i + 1) * ready_rows) of every
    // head may be loaded once ready[i] != 0 -- another stream's copy engine is still filling the buffer while the kernel runs
    const int* ready;
    int ready_rows;
    float scale;        // 1/sqrt(D)
    float scale_log2;   // scale * log2(e)
};

// Split mode (short sequences, fa_api.cu: use_split).  With few and short work items the kernel is bound by the length of one
// item's chain of KV steps, not by throughput (BASELINE config 1: 128 pair items for 148 SMs, up to 8 steps of ~1.4 us).
// The reference's answer to that is a split-K grid axis with a separate merge kernel (FA.cu:170-176, 460-496, 559-598;
// never dispatched).  Here the split stays inside the CTA: a work item is ONE 128-row Q tile, the two tile slots
// (S0/O0 + softmax warps 0-3, S1/O1 + softmax warps 4-7) take its even / odd KV tiles, and the two partial states are merged
// with the algebra of FA.cu:575-597 through shared memory in the epilogue.  Twice the items, half the chain, the tensor pipe
// still fed by two independent chains; the price is that a K/V tile now serves one Q tile instead of two.
// ---- work decomposition (shared by host tests and every warp role) ----
struct WorkItem {
    int bh, q0;      // head index, first local query row of the item
    int n0, n1;      // KV tiles the two tile slots process (pair mode: per Q tile; split mode: even / odd tiles of the one Q tile)
    int nkv;         // KV tiles the producer loads for the item
    int tile1;       // the item has a second Q tile (rows q0+128..): false past the end of Q and in split mode
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
// K/V fit comfortably in L2; inside a group the order is heaviest item first ACROSS the group's
// heads (causal: the last Q tiles see the most keys), so the dynamic scheduler hands out long items
// early and the tail of the launch is made of the lightest ones, while the CTAs running at any moment
// still share a few heads' K/V through L2.
__host__ __device__ inline WorkItem decode_work(int w, const Params& p) {
    WorkItem it;
    const int per_group = p.group_heads * p.nqp;
    const int g = w / per_group;
    const int r = w - g * per_group;
    int heads = p.BH - g * p.group_heads;
    if (heads > p.group_heads) heads = p.group_heads;
    const int qp = p.nqp - 1 - r / heads;
    it.bh = g * p.group_heads + r % heads;
    if (p.split) {
        it.q0 = qp * kBlockM;
        it.tile1 = 0;
        it.nkv = kv_trip_count(it.q0, p.Nq, p.Nkv, p.causal, p.shift);
        it.n0 = (it.nkv + 1) / 2;      // KV tiles 0, 2, 4, ...
        it.n1 = it.nkv / 2;            // KV tiles 1, 3, 5, ...
        return it;
    }
    it.q0 = qp * 2 * kBlockM;
    it.tile1 = it.q0 + kBlockM < p.Nq;
    it.n0 = kv_trip_count(it.q0, p.Nq, p.Nkv, p.causal, p.shift);
    it.n1 = it.tile1 ? kv_trip_count(it.q0 + kBlockM, p.Nq, p.Nkv, p.causal, p.shift) : 0;
    it.nkv = it.n0 > it.n1 ? it.n0 : it.n1;
    return it;
}

#ifdef FA_TIMING
__device__ unsigned long long g_timing[64];
#endif

// ptxas schedules a basic block as a whole and gives the F2FP + tcgen05.st of the first piece of P the lowest priority
// (nothing in the block depends on them), so without a fence the first piece is stored after ~85 % of ALL the tile's
// exponentials and the "PV of piece 0 under the exponentials of piece 1" overlap mostly disappears.  Loops and
// volatile asm do not stop it (the exponentials are speculatable and get hoisted across them); a data dependency
// does: the second piece's shift operand is OR-ed with (%clock & p.zero) -- p.zero is a kernel parameter that is
// always 0 -- read behind the first piece's store.
__device__ __forceinline__ uint64_t sched_fence(uint64_t v, int zero) {
#ifndef FA_SCHED_FENCE
    (void)zero;
    return v;
#else
    uint32_t c;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(c)::"memory");
    c &= (uint32_t)zero;
    return v | ((uint64_t)c << 32) | (uint64_t)c;
#endif
}

// A value ptxas may not rematerialise (it would rather recompute an address or re-load a kernel parameter in every tile
// than hold a register): OR-ed with (%clock & zero), zero being the kernel parameter that is always 0.
__device__ __forceinline__ uint32_t pin_u32(uint32_t v, int zero) {
    uint32_t c;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
    return v | (c & (uint32_t)zero);
}

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

// two fp32 -> one 32-bit word of the tensors' 16-bit format, round to nearest (low half = first value)
template <bool kBF16>
__device__ __forceinline__ uint32_t pack_16x2(float lo, float hi) {
    if (kBF16) {
        __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&b);
    }
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// Which element pairs take exp2 on the FMA pipe: kPoly = 0 none, 1 one pair in 4, 2 two in 4, 3 three in 8
__host__ __device__ constexpr bool poly_pair(int kPoly, int pair) {
    return kPoly == 3 ? (pair % 8 == 0 || pair % 8 == 3 || pair % 8 == 6) : (pair % 4) < kPoly;
}
// exponentials + fp16 packing of kCols consecutive columns (64 = one half of the tile)
template <int kPoly, bool kBF16, int kCols = 64>
__device__ __forceinline__ void exp_half(const uint32_t* s, uint32_t* pk, uint64_t scale2, uint64_t neg2,
                                         uint64_t& sum_a, uint64_t& sum_b) {
#pragma unroll
    for (int e = 0; e < kCols; e += 2) {
        const int q = e / 2;                                           // element pair
        const uint64_t x2 =
            fma_f32x2(pack_f32x2(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), scale2, neg2);
        float p0, p1;
#ifdef FA_SKELETON
        unpack_f32x2(x2, p0, p1);
#else
        if (poly_pair(kPoly, q)) exp2_pair<true>(x2, p0, p1);
        else exp2_pair<false>(x2, p0, p1);
#endif
        if (q & 1) sum_b = add_f32x2(sum_b, pack_f32x2(p0, p1));       // row sum of the un-rounded p (FA.cu:273-279)
        else sum_a = add_f32x2(sum_a, pack_f32x2(p0, p1));
        pk[q] = pack_16x2<kBF16>(p0, p1);                              // low half = even column
    }
}

#ifdef FA_SUM_GUARD
// ---- softmax of one 128x128 S tile without a row max (one thread owns one row) ----
// The reference kernel takes the row max of every tile and rescales O every tile (FA.cu:256-270); the lazy form
// above still pays 61 FMNMX3 + a vote per tile on the critical path S -> P.  But the max is only needed to keep
// exp2(s - m_ref) inside the FP16 range, and the row SUM of the exponentials -- which the algorithm needs anyway
// (FA.cu:273-279) -- bounds every one of them: sum <= 2^13 means every p <= 2^13 < 65504, with FP16's relative
// precision intact.  So: exponentials against the current m_ref, check the running tile sum before a piece of P is
// stored, and only when it trips (first tile of a row, or scores that outgrew m_ref by ~2^13) take the slow path:
// reload S from TMEM (P has not overwritten those columns yet), take the true max, rescale O and l, redo the piece.
// A trip in the second piece finds the first one already published: O is then rescaled between the two pieces' PV
// MMAs, behind the o_half commit the MMA warp issues after the first piece's k-steps.
constexpr float kTripSum = 8192.0f;

template <int kFrom, int kTo>
__device__ __forceinline__ float row_max(const uint32_t* s) {
    float mx0 = fmaxf(__uint_as_float(s[kFrom + 0]), __uint_as_float(s[kFrom + 1]));
    float mx1 = fmaxf(__uint_as_float(s[kFrom + 2]), __uint_as_float(s[kFrom + 3]));
    float mx2 = fmaxf(__uint_as_float(s[kFrom + 4]), __uint_as_float(s[kFrom + 5]));
    float mx3 = fmaxf(__uint_as_float(s[kFrom + 6]), __uint_as_float(s[kFrom + 7]));
#pragma unroll
    for (int i = kFrom + 8; i < kTo; i += 8) {
        mx0 = fmax3(mx0, __uint_as_float(s[i + 0]), __uint_as_float(s[i + 1]));
        mx1 = fmax3(mx1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        mx2 = fmax3(mx2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
        mx3 = fmax3(mx3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
    }
    return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
}

template <int D>
__device__ __forceinline__ void rescale_o(uint32_t tO, float alpha) {
    const uint64_t alpha2 = pack_f32x2(alpha, alpha);
#pragma unroll 1
    for (int c = 0; c < D; c += 32) {
        uint32_t o[32];
        tmem_ld_x32(tO + c, o);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            float lo, hi;
            unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(o[i]), __uint_as_float(o[i + 1])), alpha2), lo, hi);
            o[i] = __float_as_uint(lo);
            o[i + 1] = __float_as_uint(hi);
        }
        tmem_st_x32(tO + c, o);
    }
}

__device__ __forceinline__ float hsum_f32x2(uint64_t a, uint64_t b) {
    float x, y;
    unpack_f32x2(add_f32x2(a, b), x, y);
    return x + y;
}

template <int D, bool kMask, int kPoly, bool kBF16>
__device__ __forceinline__ void softmax_tile(const Params& p, uint32_t tS, uint32_t tO, uint32_t bar_p_full,
                                             uint32_t bar_o_full, uint32_t bar_o_half, int lim_local, bool have_o,
                                             uint32_t pv_count, float& m_ref, float& l_run, float /*sl2*/ = 0.f) {
    static_assert(kPParts == 2, "the sum-guarded softmax delivers P in two pieces");
    uint32_t s[kBlockN];
    tmem_ld_x32(tS + 0, s + 0);
    tmem_ld_x32(tS + 32, s + 32);
    tmem_ld_x32(tS + 64, s + 64);
    tmem_ld_x32(tS + 96, s + 96);
    tmem_wait_ld();

    if (kMask) {
#pragma unroll
        for (int i = 0; i < kBlockN; i++)
            if (i >= lim_local) s[i] = 0xff800000u;  // -inf
    }
    const uint64_t scale2 = pack_f32x2(p.scale_log2, p.scale_log2);
    auto publish = [&](int part) {
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) mbar_arrive(bar_p_full + 8 * part);   // one arrival per warp (barrier count 4)
    };
    auto shift2 = [&]() {
        const float m_used = (m_ref == -INFINITY) ? 0.0f : m_ref;
        const float neg = -m_used * p.scale_log2;
        return pack_f32x2(neg, neg);
    };

    // ---- piece 0: keys 0-63 -> TMEM columns [0,32)
    uint64_t sum_a, sum_b;
    uint32_t pk[32];
    float sum0;
#pragma unroll 1
    for (int pass = 0;; ++pass) {
        if (pass == 1 || have_o) {      // first tile of an item: no reference yet, nothing to speculate on
            sum_a = 0ull;
            sum_b = 0ull;
            exp_half<kPoly, kBF16>(s, pk, scale2, shift2(), sum_a, sum_b);
            sum0 = hsum_f32x2(sum_a, sum_b);
            if (pass == 1) break;
            const bool bad = !(sum0 <= kTripSum) || m_ref == -INFINITY;   // NaN trips as well
            if (!__any_sync(0xffffffffu, bad)) break;
        }
        // slow path (warp-uniform): true max of the whole tile, O and l follow the new reference
        tmem_ld_x32(tS + 0, s + 0);
        tmem_ld_x32(tS + 32, s + 32);
        tmem_wait_ld();
        if (kMask) {
#pragma unroll
            for (int i = 0; i < 64; i++)
                if (i >= lim_local) s[i] = 0xff800000u;
        }
        const float m_new = fmaxf(m_ref, row_max<0, kBlockN>(s));
        if (have_o) {
            const float alpha = (m_new == -INFINITY) ? 1.0f : ex2_approx((m_ref - m_new) * p.scale_log2);
            // O_t holds PV(0..j-1); the last of them must have retired before we touch it
            mbar_wait(bar_o_full, (pv_count - 1u) & 1u, 40);
            tc_fence_after();
            rescale_o<D>(tO, alpha);
            l_run *= alpha;
        }
        m_ref = m_new;
    }
    tmem_st_x32(tS, pk);

    // ---- piece 1: keys 64-127 -> columns [32,64); piece 0 is published a quarter tile in
    uint32_t pk2[32];
    float tile_sum;
    {
        const uint64_t neg2 = shift2();
        sum_a = 0ull;
        sum_b = 0ull;
        exp_half<kPoly, kBF16, 32>(s + 64, pk2, scale2, neg2, sum_a, sum_b);
        publish(0);
        exp_half<kPoly, kBF16, 32>(s + 96, pk2 + 16, scale2, neg2, sum_a, sum_b);
        tile_sum = sum0 + hsum_f32x2(sum_a, sum_b);
    }
    if (__any_sync(0xffffffffu, !(tile_sum <= kTripSum))) {
        // slow path: piece 0 is already with the MMA warp under the old reference
        tmem_ld_x32(tS + 64, s + 64);       // S columns [64,128) are untouched by P
        tmem_ld_x32(tS + 96, s + 96);
        tmem_wait_ld();
        if (kMask) {
#pragma unroll
            for (int i = 64; i < kBlockN; i++)
                if (i >= lim_local) s[i] = 0xff800000u;
        }
        const float m_new = fmaxf(m_ref, row_max<64, kBlockN>(s));
        const float alpha = (m_new == -INFINITY) ? 1.0f : ex2_approx((m_ref - m_new) * p.scale_log2);
        mbar_wait(bar_o_half, pv_count & 1u, 41);      // PV of piece 0 of THIS tile (and everything before it) retired
        tc_fence_after();
        rescale_o<D>(tO, alpha);
        l_run *= alpha;
        sum0 *= alpha;
        m_ref = m_new;
        sum_a = 0ull;
        sum_b = 0ull;
        exp_half<kPoly, kBF16>(s + 64, pk2, scale2, shift2(), sum_a, sum_b);
        tile_sum = sum0 + hsum_f32x2(sum_a, sum_b);
    }
    tmem_st_x32(tS + 32, pk2);
    publish(1);
    l_run += tile_sum;
}
#else
// ---- softmax of one 128x128 S tile; one thread owns one row ----
template <int D, bool kMask, int kPoly, bool kBF16>
__device__ __forceinline__ void softmax_tile(const Params& p, uint32_t tS, uint32_t tO, uint32_t bar_p_full,
                                             uint32_t bar_o_full, uint32_t /*bar_o_half*/, int lim_local, bool have_o,
                                             uint32_t pv_count, float& m_ref, float& l_run, float sl2) {
    // sl2 = sl2: an argument so that the caller can pin it in a register (-DFA_SCALE_REG) instead of the tile
    // re-loading it from the constant bank between the row max and the first exponential
    uint32_t s[kBlockN];
    tmem_ld_x32(tS + 0, s + 0);
    tmem_ld_x32(tS + 32, s + 32);
    tmem_ld_x32(tS + 64, s + 64);
    tmem_ld_x32(tS + 96, s + 96);
    tmem_wait_ld();

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
    const float m_new = fmaxf(m_ref, m_tile);

    // Lazy rescale (replaces the reference's every-tile O *= alpha, FA.cu:267-270): the reference
    // max only moves when the true max has outgrown it by 2^kRescaleThreshold.  Rare after a row's first tiles, so the
    // block lives BEHIND the tile's straight-line code (label `rescale` below): inline it sat in the middle of the hot
    // path, which then jumped 2.6 KB across it on every tile (stall_no_inst at the jump target, r02_final ncu capture).
    const bool need = (m_new - m_ref) * sl2 > kRescaleThreshold;  // NaN (-inf - -inf) -> false
    if (__any_sync(0xffffffffu, need)) goto rescale;
resume:
    {
    const float m_used = (m_ref == -INFINITY) ? 0.0f : m_ref;
    const float neg = -m_used * sl2;
    const uint64_t scale2 = pack_f32x2(sl2, sl2);
    const uint64_t neg2 = pack_f32x2(neg, neg);
    uint64_t sum_a = 0ull, sum_b = 0ull;     // (0.f, 0.f)
    uint32_t pk[32];
    // P_t (fp16 A operand of PV) overwrites columns [0,64) of S_t (key c -> column c/2), delivered in kPParts pieces:
    // keys 0-63 -> piece 0, keys 64-127 -> piece 1 (or 64-95 -> piece 1, 96-127 -> piece 2)
    // The wait::st + arrive of a piece are issued under the exponentials of the next piece (the store's latency
    // would otherwise sit on this thread's critical path: +1-3 % at N <= 2048, profiles/r01_v4b_defer_group_ab.log).
    auto publish = [&](int part) {
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) mbar_arrive(bar_p_full + 8 * part);   // one arrival per warp (barrier count 4)
    };
    uint32_t pk2[32];
#ifdef FA_STREAM_S
    // Variant: the second half of S is not kept in registers across the first half's exponentials but read again
    // from TMEM (P only ever covers columns [0,64) of S_t, so columns [64,128) stay intact for the whole tile).
    // 64 fewer live registers while the first piece is computed: room for ptxas to keep MUFU results in flight
    // instead of consuming a third of them within 4-7 instructions (profiles/r01_softmax_schedule.txt).
    {
        static_assert(kPParts == 2, "");
        uint32_t hi[64];
        exp_half<kPoly, kBF16, 48>(s, pk, scale2, neg2, sum_a, sum_b);
        tmem_ld_x32(tS + 64, hi);                 // lands under the last quarter of the first piece
        tmem_ld_x32(tS + 96, hi + 32);
        exp_half<kPoly, kBF16, 16>(s + 48, pk + 24, scale2, neg2, sum_a, sum_b);
        tmem_st_x32(tS, pk);
        tmem_wait_ld();
        if (kMask) {
#pragma unroll
            for (int i = 0; i < 64; i++)
                if (64 + i >= lim_local) hi[i] = 0xff800000u;
        }
        exp_half<kPoly, kBF16, 32>(hi, pk2, scale2, neg2, sum_a, sum_b);
        publish(0);
        exp_half<kPoly, kBF16, 32>(hi + 32, pk2 + 16, scale2, neg2, sum_a, sum_b);
        tmem_st_x32(tS + 32, pk2);
        publish(1);
        float a0, a1;
        unpack_f32x2(add_f32x2(sum_a, sum_b), a0, a1);
        l_run += a0 + a1;
        return;
    }
#endif
    exp_half<kPoly, kBF16>(s, pk, scale2, neg2, sum_a, sum_b);
    tmem_st_x32(tS, pk);
    const uint64_t neg2b = sched_fence(neg2, p.zero);      // piece 1 may not start before piece 0 is on its way
    exp_half<kPoly, kBF16, 32>(s + 64, pk2, scale2, neg2b, sum_a, sum_b);
    publish(0);
    if (kPParts == 2) {
        exp_half<kPoly, kBF16, 32>(s + 96, pk2 + 16, scale2, neg2b, sum_a, sum_b);
        tmem_st_x32(tS + 32, pk2);
        publish(1);
    } else {
        tmem_st_x16(tS + 32, pk2);
        exp_half<kPoly, kBF16, 16>(s + 96, pk2 + 16, scale2, neg2, sum_a, sum_b);
        publish(1);
        exp_half<kPoly, kBF16, 16>(s + 112, pk2 + 24, scale2, neg2, sum_a, sum_b);
        tmem_st_x16(tS + 48, pk2 + 16);
        publish(2);
    }
    float a0, a1;
    unpack_f32x2(add_f32x2(sum_a, sum_b), a0, a1);
    l_run += a0 + a1;
    return;
    }
rescale:
    if (have_o) {
        const float alpha = (m_new == -INFINITY) ? 1.0f : ex2_approx((m_ref - m_new) * sl2);
        const uint64_t alpha2 = pack_f32x2(alpha, alpha);
        // O_t holds PV(0..j-1); the last of them must have retired before we touch it
        mbar_wait(bar_o_full, (pv_count - 1u) & 1u, 40);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < D; c += 32) {
            uint32_t o[32];
            tmem_ld_x32(tO + c, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                float lo, hi;
                unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(o[i]), __uint_as_float(o[i + 1])), alpha2), lo, hi);
                o[i] = __float_as_uint(lo);
                o[i + 1] = __float_as_uint(hi);
            }
            tmem_st_x32(tO + c, o);
        }
        l_run *= alpha;
    }
    m_ref = m_new;
    goto resume;
}
#endif  // FA_SUM_GUARD

// ---- softmax_tile with the reference decision off the critical path (-DFA_SPEC) ----
// softmax_tile cannot start an exponential before the 61-deep FMNMX3 row max, the compare against the reference, a warp vote
// and a branch have all resolved: ~250 cycles of ALU-pipe work plus ~100 cycles of serial latency per tile during which the
// FMA pipe and the SFU idle (profiles/r02_softmax_hot_loop.sass.txt).  After a row's first tiles the vote says "keep the
// reference" almost always, so this form starts the first piece's exponentials against the reference the row already has,
// with the row max computed in the same basic block (ptxas interleaves the two), and votes afterwards: a tile whose max has
// outgrown the reference by 2^kRescaleThreshold rescales O as before and runs the piece again (the loop below re-enters the
// same code: nothing is duplicated).  Same instructions on the hot path, same results bit for bit.
template <int kFrom, int kTo>
__device__ __forceinline__ float row_max_of(const uint32_t* s) {
    float mx0 = fmaxf(__uint_as_float(s[kFrom + 0]), __uint_as_float(s[kFrom + 1]));
    float mx1 = fmaxf(__uint_as_float(s[kFrom + 2]), __uint_as_float(s[kFrom + 3]));
    float mx2 = fmaxf(__uint_as_float(s[kFrom + 4]), __uint_as_float(s[kFrom + 5]));
    float mx3 = fmaxf(__uint_as_float(s[kFrom + 6]), __uint_as_float(s[kFrom + 7]));
#pragma unroll
    for (int i = kFrom + 8; i < kTo; i += 8) {
        mx0 = fmax3(mx0, __uint_as_float(s[i + 0]), __uint_as_float(s[i + 1]));
        mx1 = fmax3(mx1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        mx2 = fmax3(mx2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
        mx3 = fmax3(mx3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
    }
    return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
}

template <int D, bool kMask, int kPoly, bool kBF16>
__device__ __forceinline__ void softmax_tile_spec(const Params& p, uint32_t tS, uint32_t tO, uint32_t bar_p_full,
                                                  uint32_t bar_o_full, int lim_local, uint32_t pv_count,
                                                  float& m_ref, float& l_run, float sl2) {
    static_assert(kPParts == 2, "two pieces of P");
    uint32_t s[kBlockN];
    tmem_ld_x32(tS + 0, s + 0);
    tmem_ld_x32(tS + 32, s + 32);
    tmem_ld_x32(tS + 64, s + 64);
    tmem_ld_x32(tS + 96, s + 96);
    tmem_wait_ld();
    if (kMask) {
#pragma unroll
        for (int i = 0; i < kBlockN; i++)
            if (i >= lim_local) s[i] = 0xff800000u;  // -inf
    }
    // (a row's first tile has no reference to speculate on and no O to rescale: the caller runs it through softmax_tile)
    const uint64_t scale2 = pack_f32x2(sl2, sl2);
    uint64_t neg2, sum_a, sum_b;
    uint32_t pk[32];
    float m_new = m_ref;
    bool again = false;
    // bottom-tested on purpose: the body (piece 0's exponentials + the row max) exists once; a pass that follows a rescale
    // ends at the vote because then m_ref == m_new
#pragma unroll 1
    do {
        if (again) {
            // rare: the reference moves, O and l follow (FA.cu:267-270), the piece is computed again
            const float alpha = (m_new == -INFINITY) ? 1.0f : ex2_approx((m_ref - m_new) * sl2);
            const uint64_t alpha2 = pack_f32x2(alpha, alpha);
            mbar_wait(bar_o_full, (pv_count - 1u) & 1u, 40);   // O_t holds PV(0..j-1); the last of them must have retired
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < D; c += 32) {
                uint32_t o[32];
                tmem_ld_x32(tO + c, o);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float lo, hi;
                    unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(o[i]), __uint_as_float(o[i + 1])), alpha2), lo, hi);
                    o[i] = __float_as_uint(lo);
                    o[i + 1] = __float_as_uint(hi);
                }
                tmem_st_x32(tO + c, o);
            }
            l_run *= alpha;
            m_ref = m_new;
        }
        const float m_used = (m_ref == -INFINITY) ? 0.0f : m_ref;
        const float neg = -m_used * sl2;
        neg2 = pack_f32x2(neg, neg);
        sum_a = 0ull;
        sum_b = 0ull;
        exp_half<kPoly, kBF16>(s, pk, scale2, neg2, sum_a, sum_b);
        m_new = fmaxf(m_ref, row_max_of<0, kBlockN>(s));
        const bool need = (m_new - m_ref) * sl2 > kRescaleThreshold;  // NaN (-inf - -inf) -> false
        again = __any_sync(0xffffffffu, need);
    } while (again);
    auto publish = [&](int part) {
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) mbar_arrive(bar_p_full + 8 * part);   // one arrival per warp (barrier count 4)
    };
    tmem_st_x32(tS, pk);
    uint32_t pk2[32];
    exp_half<kPoly, kBF16, 32>(s + 64, pk2, scale2, neg2, sum_a, sum_b);
    publish(0);
    exp_half<kPoly, kBF16, 32>(s + 96, pk2 + 16, scale2, neg2, sum_a, sum_b);
    tmem_st_x32(tS + 32, pk2);
    publish(1);
    float a0, a1;
    unpack_f32x2(add_f32x2(sum_a, sum_b), a0, a1);
    l_run += a0 + a1;
}

// ---- streamed softmax of one 128x128 S tile (one thread per row) ----
// softmax_tile above cannot start an exponential before all 128 columns of the row have crossed the TMEM read
// port (64 B/clk per sub-partition: 256 cycles per warp and tile) and the 61-deep max has run -- ~400 cycles of
// the S -> P -> PV -> QK^T chain during which the SFU idles.  Here S is read in four chunks of 32 columns and the
// exponentials of a chunk run against the reference max the row ALREADY has (first tile of an item: the max of
// chunk 0), under the loads and the max of the following chunks:
//   * P may exceed 1 by up to 2^kHardThreshold (fp16/bf16 hold 2^15 with full relative precision, l and O are fp32);
//     only a row whose scores outgrow the reference by more than that makes the warp drop the tile's work and
//     redo it with softmax_tile (nothing has been written to TMEM at that point: the vote sits before the
//     first tcgen05.st);
//   * the lazy update of the reference (threshold 2^kRescaleThreshold, as above) moves to the END of the tile:
//     O is rescaled behind this tile's PV, while the tensor core runs the tile's next QK^T -- off the chain.
constexpr float kHardThreshold = 15.0f;

template <bool kMask>
__device__ __forceinline__ void mask_chunk(uint32_t* c, int base, int lim_local) {
    if (kMask) {
#pragma unroll
        for (int i = 0; i < 32; i++)
            if (base + i >= lim_local) c[i] = 0xff800000u;  // -inf
    }
}
__device__ __forceinline__ float max_chunk(const uint32_t* c) {
    float m0 = fmax3(__uint_as_float(c[0]), __uint_as_float(c[1]), __uint_as_float(c[2]));
    float m1 = fmax3(__uint_as_float(c[3]), __uint_as_float(c[4]), __uint_as_float(c[5]));
    float m2 = fmax3(__uint_as_float(c[6]), __uint_as_float(c[7]), __uint_as_float(c[8]));
    float m3 = fmax3(__uint_as_float(c[9]), __uint_as_float(c[10]), __uint_as_float(c[11]));
    m0 = fmax3(m0, __uint_as_float(c[12]), __uint_as_float(c[13]));
    m1 = fmax3(m1, __uint_as_float(c[14]), __uint_as_float(c[15]));
    m2 = fmax3(m2, __uint_as_float(c[16]), __uint_as_float(c[17]));
    m3 = fmax3(m3, __uint_as_float(c[18]), __uint_as_float(c[19]));
    m0 = fmax3(m0, __uint_as_float(c[20]), __uint_as_float(c[21]));
    m1 = fmax3(m1, __uint_as_float(c[22]), __uint_as_float(c[23]));
    m2 = fmax3(m2, __uint_as_float(c[24]), __uint_as_float(c[25]));
    m3 = fmax3(m3, __uint_as_float(c[26]), __uint_as_float(c[27]));
    m0 = fmax3(m0, __uint_as_float(c[28]), __uint_as_float(c[29]));
    m1 = fmax3(m1, __uint_as_float(c[30]), __uint_as_float(c[31]));
    return fmaxf(fmax3(m0, m1, m2), m3);
}

// Returns false (and leaves TMEM, m_ref, l_run untouched) when the tile has to be redone by softmax_tile.
template <int D, bool kMask, int kPoly, bool kBF16>
__device__ __forceinline__ bool softmax_tile_stream(const Params& p, uint32_t tS, uint32_t tO, uint32_t bar_p_full,
                                                    uint32_t bar_o_full, int lim_local, bool last, uint32_t pv_count,
                                                    float& m_ref, float& l_run) {
    static_assert(kPParts == 2, "the streamed softmax delivers P in two pieces");
    uint32_t a[32], b[32], c[32];
    tmem_ld_x32(tS, a);
    tmem_wait_ld();
    tmem_ld_x32(tS + 32, b);                   // lands under chunk 0's max and exponentials
    mask_chunk<kMask>(a, 0, lim_local);
    const float mx0 = max_chunk(a);
    // reference for this tile: the row's current one; a row without one (first tile) takes the max of chunk 0
    const float m_use = (m_ref == -INFINITY) ? mx0 : m_ref;
    const float neg = -((m_use == -INFINITY) ? 0.0f : m_use) * p.scale_log2;
    const uint64_t scale2 = pack_f32x2(p.scale_log2, p.scale_log2);
    const uint64_t neg2 = pack_f32x2(neg, neg);
    uint64_t sum_a = 0ull, sum_b = 0ull;
    uint32_t pk[32], pk2[32];
    exp_half<kPoly, kBF16, 32>(a, pk, scale2, neg2, sum_a, sum_b);
    tmem_wait_ld();
    tmem_ld_x32(tS + 64, a);                   // chunks 2 and 3 land under chunk 1
    tmem_ld_x32(tS + 96, c);
    mask_chunk<kMask>(b, 32, lim_local);
    const float mx1 = max_chunk(b);
    exp_half<kPoly, kBF16, 32>(b, pk + 16, scale2, neg2, sum_a, sum_b);
    tmem_wait_ld();
    mask_chunk<kMask>(a, 64, lim_local);
    mask_chunk<kMask>(c, 96, lim_local);
    const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(max_chunk(a), max_chunk(c)));
    // the exponentials above (and below) are only good if nothing outgrew the reference by more than 2^kHard
    const bool bad = (m_tile - m_use) * p.scale_log2 > kHardThreshold;    // NaN (-inf - -inf) -> false
    if (__any_sync(0xffffffffu, bad)) return false;

    auto publish = [&](int part) {
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) mbar_arrive(bar_p_full + 8 * part);   // one arrival per warp (barrier count 4)
    };
    tmem_st_x32(tS, pk);                       // piece 0: keys 0-63 -> columns [0,32) (all of S is in registers)
    exp_half<kPoly, kBF16, 32>(a, pk2, scale2, neg2, sum_a, sum_b);
    publish(0);
    exp_half<kPoly, kBF16, 32>(c, pk2 + 16, scale2, neg2, sum_a, sum_b);
    tmem_st_x32(tS + 32, pk2);                 // piece 1: keys 64-127 -> columns [32,64)
    publish(1);
    float a0, a1;
    unpack_f32x2(add_f32x2(sum_a, sum_b), a0, a1);
    l_run += a0 + a1;
    m_ref = m_use;

    // Lazy rescale for the NEXT tile (replaces the reference's every-tile O *= alpha, FA.cu:267-270)
    const float m_new = fmaxf(m_use, m_tile);
    const bool need = (m_new - m_use) * p.scale_log2 > kRescaleThreshold;
    if (!last && __any_sync(0xffffffffu, need)) {
        const float alpha = (m_new == -INFINITY) ? 1.0f : ex2_approx((m_use - m_new) * p.scale_log2);
        mbar_wait(bar_o_full, pv_count & 1u, 42);          // PV of THIS tile retired: O_t is quiescent
        tc_fence_after();
        const uint64_t alpha2 = pack_f32x2(alpha, alpha);
#pragma unroll 1
        for (int cc = 0; cc < D; cc += 32) {
            uint32_t o[32];
            tmem_ld_x32(tO + cc, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                float lo, hi;
                unpack_f32x2(mul_f32x2(pack_f32x2(__uint_as_float(o[i]), __uint_as_float(o[i + 1])), alpha2), lo, hi);
                o[i] = __float_as_uint(lo);
                o[i + 1] = __float_as_uint(hi);
            }
            tmem_st_x32(tO + cc, o);
        }
        l_run *= alpha;
        m_ref = m_new;
    }
    return true;
}

// ---- plain epilogue of one O tile: TMEM -> registers -> O / l -> fp16 -> staging tile in shared memory ----
// One thread per row (TMEM lane).  The staging tile is the item's idle Q tile buffer, 16-byte chunks XOR-swizzled by
// row & 7 as TMA expects for SWIZZLE_128B; the caller orders the writes before the TMA store (fence.proxy.async).
template <int D, bool kBF16, int kChunk = 32>
__device__ __forceinline__ void stage_o_tile(uint32_t tO, uint32_t tile_smem, int row_in_tile, float l_run, bool have_o) {
    using C = Cfg<D>;
    static_assert(kChunk == 32 || kChunk == 16, "columns per TMEM load");
    const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;   // FA.cu:502-503
    const uint32_t stage = tile_smem + row_in_tile * 128;
#pragma unroll
    for (int c = 0; c < D; c += kChunk) {
        uint32_t o[kChunk];
        if (have_o) {
            if (kChunk == 32) tmem_ld_x32(tO + c, o);
            else tmem_ld_x16(tO + c, o);
            tmem_wait_ld();
        } else {
#pragma unroll
            for (int i = 0; i < kChunk; i++) o[i] = 0u;
        }
#pragma unroll
        for (int i = 0; i < kChunk; i += 8) {
            const int col = c + i;
            const uint32_t addr = stage + (col >> 6) * C::kPanelBytes + ((((col & 63) >> 3) ^ (row_in_tile & 7)) << 4);
            const uint32_t v0 = pack_16x2<kBF16>(__uint_as_float(o[i + 0]) * inv, __uint_as_float(o[i + 1]) * inv);
            const uint32_t v1 = pack_16x2<kBF16>(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
            const uint32_t v2 = pack_16x2<kBF16>(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
            const uint32_t v3 = pack_16x2<kBF16>(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
            st_shared_v4(addr, v0, v1, v2, v3);
        }
    }
    fence_proxy_async_smem();   // generic-proxy writes -> visible to the TMA store
}

template <int D, int kPoly, bool kBF16 = false>
__global__ void __launch_bounds__(kNumThreads, 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const Params p) {
    using C = Cfg<D>;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    if (smem_base - smem_u32(smem_raw) + (uint32_t)(C::kMlOffset + C::kXchBytes) > (uint32_t)C::kSmemBytes) {
        if (threadIdx.x == 0) {                       // dynamic window less aligned than the slack covers: refuse to run
            watchdog_raise(99);
            watchdog_publish();
        }
        return;
    }
    const uint32_t sQ = smem_base;
    const uint32_t sKV = smem_base + C::kSmemQ;
    const uint32_t bars = smem_base + C::kBarOffset;
    const uint32_t bar_q_full = bars + 0;                         // [slot]
    const uint32_t bar_q_empty = bars + 16;                       // [slot]
    const uint32_t bar_kv_full = bars + 32;                       // [kStages]
    const uint32_t bar_kv_empty = bar_kv_full + 8 * C::kStages;   // [kStages]
    const uint32_t bar_s_full = bar_kv_empty + 8 * C::kStages;    // [tile]
    const uint32_t bar_p_full = bar_s_full + 16;                  // [tile][piece]  index 4*t + piece
    const uint32_t bar_o_full = bar_p_full + 64;                  // [tile]
    const uint32_t bar_o_half = bar_o_full + 16;                  // [tile] PV of the first piece of P retired (sum-guard slow path)
    const uint32_t bar_o_staged = bar_o_half + 16;                // [Q slot][tile] softmax warps -> store warp
    const uint32_t bar_sched_full = bar_o_staged + 32;            // [2] work-index slots, producer -> everyone
    const uint32_t bar_sched_empty = bar_sched_full + 16;         // [2]
    const uint32_t bar_epi_full = bar_sched_empty + 16;           // [tile] softmax warps -> epilogue warpgroup: the item's row sums are in smem
    const uint32_t bar_o_free = bar_epi_full + 16;                // [tile] epilogue warpgroup -> MMA warp / softmax warps: O_t and the l slots are drained
    const uint32_t tmem_slot = bar_o_free + 16;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    volatile int* sched_w = reinterpret_cast<volatile int*>(tmem_slot_ptr + 2);   // [2]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // the epilogue warpgroup takes the plain fp16 epilogue; the merge of split mode and the partial-state format stay with
    // the softmax warps (their rows' (m, l) live in those warps' registers and both are short-sequence / per-hop paths)
    const bool epi = kEpiWg && !p.split && !p.partial_mode;
#ifdef FA_TIMING
    long long k_c0 = 0;
    unsigned long long k_t0 = 0;
    const long long k_c0_all = clock64();
    if (threadIdx.x == 0) {
        k_c0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(k_t0));
    }
#endif

    if (threadIdx.x == 0) {
        *watchdog_block_flag() = 0u;
        for (int i = 0; i < 2; i++) {
            mbar_init(bar_q_full + 8 * i, 1);
            mbar_init(bar_q_empty + 8 * i, 2);        // last QK^T of the item retired + its O tiles stored
        }
        for (int i = 0; i < C::kStages; i++) {
            mbar_init(bar_kv_full + 8 * i, 1);
            mbar_init(bar_kv_empty + 8 * i, 1);
        }
        for (int t = 0; t < 2; t++) {
            mbar_init(bar_s_full + 8 * t, 1);
            for (int part = 0; part < 4; part++)
                mbar_init(bar_p_full + 32 * t + 8 * part, 4);   // one arrival per softmax warp of the tile, per piece of P
            mbar_init(bar_o_full + 8 * t, 1);
            mbar_init(bar_o_half + 8 * t, 1);
            mbar_init(bar_o_staged + 8 * t, 4);       // one arrival per softmax (or epilogue) warp of the tile,
            mbar_init(bar_o_staged + 8 * (2 + t), 4); // per Q slot
            mbar_init(bar_epi_full + 8 * t, 4);       // one arrival per softmax warp of the tile
            mbar_init(bar_o_free + 8 * t, 4);         // one arrival per epilogue warp
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(bar_sched_full + 8 * i, 1);
            mbar_init(bar_sched_empty + 8 * i, epi ? 14 : 10);   // MMA warp + 8 softmax warps + store warp (+ 4 epilogue warps)
        }
        fence_mbar_init();
    }
    if (warp == kMmaWarp) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    if (warp == kLoadWarp && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        tma_prefetch_desc(&tmO);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor
    // prefetch) may overlap the tail of the previous kernel in the stream; global memory is only
    // touched below this point.  The next kernel's prologue may start as soon as our CTAs retire.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // Dynamic tile scheduler (replaces the reference's static blockIdx mapping, FA.cu:103-112): the
    // producer warp claims work indices (first one static, the rest from a global counter) and
    // publishes them through a 2-slot smem mailbox; every consumer warp reads slot i&1 for its i-th
    // item.  Work order is heads-outermost, heaviest Q pair first (decode_work), so the causal tail is
    // made of the lightest items.  Returns -1 when the grid has run out of work.
    auto next_work = [&](uint32_t i) -> int {
        const uint32_t slot = i & 1u;
        mbar_wait(bar_sched_full + 8 * slot, (i >> 1) & 1u, 50);
        const int w = sched_w[slot];
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sched_empty + 8 * slot);
        return w;
    };

    // The producer and MMA warps run their loops converged (all 32 lanes take the same branches
    // and waits); the instructions with side effects sit under elect_one().  Warp-uniform control
    // flow keeps descriptors and barrier addresses in uniform registers -- in a lane-divergent
    // region every UTCHMMA costs an extra ELECT / R2UR.BROADCAST sequence and the single issuing
    // thread becomes the bottleneck of the whole CTA (profiles/r01_v1_full_n8192_summary.txt).
    if (kEpiWg && warp >= kEpiWarp0) {
        setmaxnreg_dec<kRegsEpi>();
        // =============================== epilogue warpgroup ===============================
        // Warp 12 + q owns TMEM lanes [32q, 32q + 32) = rows 32q.. of whichever tile it drains.  Per work item and tile:
        // wait until the tile's softmax warps have left their row sums in shared memory (epi_full) and the tile's last PV
        // has retired (o_full), stage the fp16 rows for the store warp (o_staged), and give O_t and the l slots back
        // (o_free: the MMA warp waits for it before the next item's first PV overwrites O_t, the softmax warps before
        // they write the next item's row sums).  Tile 0 then tile 1: in steady state they finish half a period apart.
        if (epi) {
            const int q = warp & 3;
            const int row_in_tile = q * 32 + lane;
            const uint32_t lane_base = (uint32_t)(q * 32) << 16;
            uint32_t cnt0 = 0u, cnt1 = 0u;        // items drained so far, per tile
            uint32_t pv0 = 0u, pv1 = 0u;          // PV MMAs issued so far, per tile == o_full completions to expect
            for (uint32_t it = 0;; ++it) {
                const int w = next_work(it);
                if (w < 0) break;
                const WorkItem wi = decode_work(w, p);
                const uint32_t slot = it & 1u;
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    if (t == 1 && !wi.tile1) continue;                  // tile absent: nobody arrives, nothing to stage
                    const int n_t = t ? wi.n1 : wi.n0;
                    uint32_t& cnt = t ? cnt1 : cnt0;
                    uint32_t& pv = t ? pv1 : pv0;
#ifdef FA_EPI_SLEEP
                    mbar_wait_sleepy(bar_epi_full + 8 * t, cnt & 1u, 36 + t, FA_EPI_SLEEP);
#else
                    mbar_wait(bar_epi_full + 8 * t, cnt & 1u, 36 + t);
#endif
                    const float l_run = ld_shared_f32(smem_base + C::kMlOffset + (t * kBlockM + row_in_tile) * 4);
                    if (n_t > 0) {
                        pv += (uint32_t)n_t;
                        mbar_wait(bar_o_full + 8 * t, (pv - 1u) & 1u, 30 + t);
                        tc_fence_after();
                    } else {
                        // no key visible to the tile: nothing else has waited for the item's Q pair to land, and the rows
                        // are staged in that very buffer (see the softmax warps' epilogue below)
                        mbar_wait(bar_q_full + 8 * slot, (it >> 1) & 1u, 32 + t);
                    }
                    stage_o_tile<D, kBF16, kRegsEpi >= 56 ? 32 : 16>(tmem_base + lane_base + (t ? C::kTmemO1 : C::kTmemO0),
                                           sQ + (slot * 2 + t) * C::kTileBytes, row_in_tile, l_run, n_t > 0);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(bar_o_free + 8 * t);
                        mbar_arrive(bar_o_staged + 8 * (slot * 2 + t));
                    }
                    ++cnt;
                }
            }
        }
    } else if (warp >= 8) {
    setmaxnreg_dec<kRegsOther>();   // each role's code must be dominated by its own setmaxnreg
    if (warp == kLoadWarp) {
        // =============================== TMA producer ===============================
        Ring ring{0u, 0u};
        int ready_upto = 0;            // gathered K/V: chunks [0, ready_upto) are known to have landed
        for (uint32_t it = 0;; ++it) {
            // claim the next work item and publish it
            const uint32_t slot = it & 1u;
            mbar_wait(bar_sched_empty + 8 * slot, ((it >> 1) & 1u) ^ 1u, 3);
            int w = 0;
            if (lane == 0) w = (it == 0) ? (int)blockIdx.x : (int)gridDim.x + atomicAdd(p.sched, 1);
            w = __shfl_sync(0xffffffffu, w, 0);
            if (w >= p.total_work) w = -1;
            if (lane == 0) {
                sched_w[slot] = w;
                mbar_arrive(bar_sched_full + 8 * slot);   // release: the slot write is visible to waiters
                // This CTA will not claim again.  The last CTA to get here re-arms the scheduler words for the launch that
                // reuses this slot (nobody claims any more: every claim above returned before its CTA counted itself).
                // Done here, under the last item's work, and not on the CTA's exit path: a fence + atomic round trip there
                // cost every launch ~1 us of tail.
                if (w < 0 && atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
                    p.sched[0] = 0;
                    p.sched[1] = 0;
                }
            }
            __syncwarp();
            if (w < 0) break;
            const WorkItem wi = decode_work(w, p);
            const int nmax = wi.nkv;
            const bool have_q1 = wi.tile1 != 0;
            // slot it&1 is free once the item two back has issued its last QK^T and stored its O tiles
            mbar_wait(bar_q_empty + 8 * slot, ((it >> 1) & 1u) ^ 1u, 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_q_full + 8 * slot, (have_q1 ? 2 : 1) * C::kTileBytes);
                for (int t = 0; t < (have_q1 ? 2 : 1); t++)
                    for (int pn = 0; pn < C::kPanels; pn++)
                        tma_load_3d(sQ + (slot * 2 + t) * C::kTileBytes + pn * C::kPanelBytes, &tmQ,
                                    bar_q_full + 8 * slot, pn * 64, wi.q0 + t * kBlockM, wi.bh);
            }
            __syncwarp();
            auto load_kv = [&](const CUtensorMap* tm, int j) {
                if (p.ready != nullptr) {
                    // chunk of the gathered K/V this tile lies in: wait until its copy has landed.  Flags only ever turn
                    // on during a launch, so what has been seen ready stays ready for the later items.
                    const int chunk = (j * kBlockN) / p.ready_rows;
                    if (chunk >= ready_upto) {
                        const unsigned long long t0 = global_timer_ns();
                        while (ld_acquire_sys(p.ready + chunk) == 0) {
                            __nanosleep(200);
                            if (watchdog_aborted()) break;
                            if (global_timer_ns() - t0 > kWatchdogNs) { watchdog_raise(70); break; }
                        }
                        ready_upto = chunk + 1;
                    }
                }
                const uint32_t full = bar_kv_full + 8 * ring.idx;
                mbar_wait(bar_kv_empty + 8 * ring.idx, ring.phase ^ 1u, 2);
                if (elect_one()) {
                    mbar_arrive_expect_tx(full, C::kTileBytes);
                    const uint32_t dst = sKV + ring.idx * C::kTileBytes;
#pragma unroll
                    for (int pn = 0; pn < C::kPanels; pn++)
                        tma_load_3d(dst + pn * C::kPanelBytes, tm, full, pn * 64, j * kBlockN, wi.bh);
                }
                __syncwarp();
                ring.advance<C::kStages>();
            };
            if (!p.split) {
                // ring order K_0 V_0 K_1 V_1 ... (the order the MMA warp releases them in)
                for (int j = 0; j < nmax; j++) {
                    load_kv(&tmK, j);
                    load_kv(&tmV, j);
                }
            } else {
                // split mode: K_0 K_1 V_0 K_2 V_1 K_3 V_2 ... -- both tile slots get their first S at once
                if (nmax > 0) load_kv(&tmK, 0);
                if (nmax > 1) load_kv(&tmK, 1);
                for (int j = 0; j < nmax; j++) {
                    load_kv(&tmV, j);
                    if (j + 2 < nmax) load_kv(&tmK, j + 2);
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // =============================== tcgen05.mma issuer ===============================
        Ring rk{0u, 0u};              // ring entry holding K_j
        Ring rv{1u % C::kStages, 0u}; // ring entry holding V_j
        uint32_t it = 0;
        uint32_t p_phase0 = 0u, p_phase1 = 0u;
        const uint32_t tS0 = tmem_base + C::kTmemS0, tS1 = tmem_base + C::kTmemS1;
        const uint32_t tO0 = tmem_base + C::kTmemO0, tO1 = tmem_base + C::kTmemO1;

        // S_t = Q_t K_j^T : D/16 k-steps; k-step ks lives in panel ks/4 at byte offset (ks%4)*32
        auto issue_qk = [&](uint32_t tS, uint64_t qdesc, uint32_t k_smem, uint32_t bar) {
            const uint64_t kdesc = umma_smem_desc(k_smem, 16, 1024);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < D / 16; ks++) {
                    const uint64_t off = (uint64_t)(((ks >> 2) * C::kPanelBytes + (ks & 3) * 32) >> 4);
                    umma_ss(tS, qdesc + off, kdesc + off, C::kIdescQK | (kBF16 ? C::kBf16Operands : 0u), ks > 0 ? 1u : 0u);
                }
                umma_commit(bar);
            }
            __syncwarp();
        };
        // O_t (+)= P_t V_j : 8 k-steps of 16 kv rows (P k-step = 8 TMEM columns, V k-step = 16 rows * 128 B),
        // issued piece by piece as the pieces of P arrive
        auto issue_pv = [&](uint32_t tO, uint32_t tP, uint32_t v_smem, bool accumulate, uint32_t bar_p,
                            uint32_t parity, uint32_t bar_o, uint32_t bar_oh, int tag) {
            const uint64_t vdesc = umma_smem_desc(v_smem, C::kPanelBytes, 1024);
#pragma unroll
            for (int part = 0; part < kPParts; part++) {
                mbar_wait(bar_p + 8 * part, parity, tag);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int ks = p_part_ks(part); ks < p_part_ks(part + 1); ks++)
                        umma_ts(tO, tP + ks * 8, vdesc + (uint64_t)((ks * 16 * 128) >> 4), C::kIdescPV | (kBF16 ? C::kBf16Operands : 0u),
                                (accumulate || ks > 0) ? 1u : 0u);
#ifdef FA_SUM_GUARD
                    if (part == 0) umma_commit(bar_oh);
#endif
                    if (part == kPParts - 1) umma_commit(bar_o);
                }
                __syncwarp();
            }
        };
        auto commit = [&](uint32_t bar) {
            if (elect_one()) umma_commit(bar);
            __syncwarp();
        };

        // First QK^T of an item (both tiles).  Issued one item ahead: right behind the previous item's
        // last PV, while the softmax warps are still in that item's epilogue.
        auto prologue = [&](uint32_t itn, const WorkItem& wn) {
            const uint32_t slot = itn & 1u;
            const int nm = wn.n0 > wn.n1 ? wn.n0 : wn.n1;
            mbar_wait(bar_q_full + 8 * slot, (itn >> 1) & 1u, 10);
            tc_fence_after();
            if (nm > 0) {
                mbar_wait(bar_kv_full + 8 * rk.idx, rk.phase, 11);
                tc_fence_after();
                const uint32_t k_smem = sKV + rk.idx * C::kTileBytes;
                if (wn.n0 > 0) issue_qk(tS0, umma_smem_desc(sQ + (slot * 2 + 0) * C::kTileBytes, 16, 1024), k_smem, bar_s_full);
                if (wn.n1 > 0) issue_qk(tS1, umma_smem_desc(sQ + (slot * 2 + 1) * C::kTileBytes, 16, 1024), k_smem, bar_s_full + 8);
                commit(bar_kv_empty + 8 * rk.idx);
                rk.advance<C::kStages>(); rk.advance<C::kStages>();
            }
            // the Q slot is released when the last QK^T of the item has retired (and its O tiles are stored)
            if (nm <= 1) commit(bar_q_empty + 8 * slot);
        };

        if (p.split) {
            // One Q tile per item; slot t = j & 1 takes KV tile j.  Ring entries are consumed in the producer's order
            // K_0 K_1 V_0 K_2 V_1 K_3 ...; tensor-pipe order PV_t(j) QK_t(j+2), t alternating, as in pair mode.
            Ring r{0u, 0u};
            uint32_t pph[2] = {0u, 0u};
            auto split_prologue = [&](uint32_t itn, const WorkItem& wn) {
                const uint32_t slot = itn & 1u;
                const uint64_t qd = umma_smem_desc(sQ + (slot * 2) * C::kTileBytes, 16, 1024);
                mbar_wait(bar_q_full + 8 * slot, (itn >> 1) & 1u, 10);
                tc_fence_after();
                for (int t = 0; t < 2 && t < wn.nkv; t++) {
                    mbar_wait(bar_kv_full + 8 * r.idx, r.phase, 11);
                    tc_fence_after();
                    issue_qk(t ? tS1 : tS0, qd, sKV + r.idx * C::kTileBytes, bar_s_full + 8 * t);
                    commit(bar_kv_empty + 8 * r.idx);
                    r.advance<C::kStages>();
                }
                if (wn.nkv <= 2) commit(bar_q_empty + 8 * slot);   // the last QK^T of the item has been issued
            };
            int w = next_work(0);
            WorkItem wi;
            if (w >= 0) {
                wi = decode_work(w, p);
                split_prologue(0u, wi);
            }
            for (; w >= 0; ++it) {
                const uint32_t slot = it & 1u;
                const uint64_t qd = umma_smem_desc(sQ + (slot * 2) * C::kTileBytes, 16, 1024);
                const int n = wi.nkv;
                for (int j = 0; j < n; j++) {
                    const int t = j & 1;
                    mbar_wait(bar_kv_full + 8 * r.idx, r.phase, 12);      // V_j
                    issue_pv(t ? tO1 : tO0, t ? tS1 : tS0, sKV + r.idx * C::kTileBytes, j >= 2, bar_p_full + 32 * t,
                             pph[t], bar_o_full + 8 * t, bar_o_half + 8 * t, 13 + t);
                    pph[t] ^= 1u;
                    commit(bar_kv_empty + 8 * r.idx);
                    r.advance<C::kStages>();
                    if (j + 2 < n) {
                        mbar_wait(bar_kv_full + 8 * r.idx, r.phase, 15);  // K_{j+2}
                        tc_fence_after();
                        issue_qk(t ? tS1 : tS0, qd, sKV + r.idx * C::kTileBytes, bar_s_full + 8 * t);
                        commit(bar_kv_empty + 8 * r.idx);
                        r.advance<C::kStages>();
                        if (j + 3 == n) commit(bar_q_empty + 8 * slot);   // that was the last reader of Q
                    }
                }
                w = next_work(it + 1u);
                if (w >= 0) {
                    wi = decode_work(w, p);
                    split_prologue(it + 1u, wi);
                }
            }
        } else {
        int w = next_work(0);
        WorkItem wi;
        if (w >= 0) {
            wi = decode_work(w, p);
            prologue(0u, wi);
        }
        uint32_t ocnt0 = 0u, ocnt1 = 0u;      // items so far in which tile t exists == o_free completions before this item
        for (; w >= 0; ++it) {
            const uint32_t slot = it & 1u;
            const int n0 = wi.n0, n1 = wi.n1;
            const int nmax = n0 > n1 ? n0 : n1;
            const uint64_t qdesc0 = umma_smem_desc(sQ + (slot * 2 + 0) * C::kTileBytes, 16, 1024);
            const uint64_t qdesc1 = umma_smem_desc(sQ + (slot * 2 + 1) * C::kTileBytes, 16, 1024);
            for (int j = 0; j < nmax; j++) {
                const bool has_next = j + 1 < nmax;
                mbar_wait(bar_kv_full + 8 * rv.idx, rv.phase, 12);
                const uint32_t v_smem = sKV + rv.idx * C::kTileBytes;
                const uint32_t k_smem = sKV + rk.idx * C::kTileBytes;
                // ---- tile 0: PV0(j), QK0(j+1)
                if (j < n0) {
                    // the epilogue warpgroup must have drained the previous item's O_0 before this PV overwrites it
                    if (epi && j == 0 && ocnt0 > 0u) mbar_wait(bar_o_free, (ocnt0 - 1u) & 1u, 16);
                    issue_pv(tO0, tS0, v_smem, j > 0, bar_p_full, p_phase0, bar_o_full, bar_o_half, 13);
                    p_phase0 ^= 1u;
                }
                if (has_next) {
                    mbar_wait(bar_kv_full + 8 * rk.idx, rk.phase, 15);
                    tc_fence_after();
                }
                if (j + 1 < n0) issue_qk(tS0, qdesc0, k_smem, bar_s_full);
                // ---- tile 1: PV1(j), QK1(j+1)
                if (j < n1) {
                    if (epi && j == 0 && ocnt1 > 0u) mbar_wait(bar_o_free + 8, (ocnt1 - 1u) & 1u, 17);
                    issue_pv(tO1, tS1, v_smem, j > 0, bar_p_full + 32, p_phase1, bar_o_full + 8, bar_o_half + 8, 14);
                    p_phase1 ^= 1u;
                }
                commit(bar_kv_empty + 8 * rv.idx);
                rv.advance<C::kStages>(); rv.advance<C::kStages>();
                if (j + 1 < n1) issue_qk(tS1, qdesc1, k_smem, bar_s_full + 8);
                if (has_next) {
                    commit(bar_kv_empty + 8 * rk.idx);
                    rk.advance<C::kStages>(); rk.advance<C::kStages>();
                    if (j + 2 == nmax) commit(bar_q_empty + 8 * slot);   // QK^T(nmax-1) was the last reader of Q
                }
            }
            ++ocnt0;
            if (wi.tile1) ++ocnt1;
            // next item: its Q pair is already resident in the other slot
            w = next_work(it + 1u);
            if (w >= 0) {
                wi = decode_work(w, p);
                prologue(it + 1u, wi);
            }
        }
        }   // pair mode
    } else if (warp == kStoreWarp) {
        // =============================== O tile store ===============================
        // The softmax warps stage O_t / l as fp16 in the item's (now idle) Q tile buffers, 128B-swizzled;
        // this warp hands them to TMA, which clips rows past Nq, and then returns the Q slot.
        uint32_t staged_parity = 0u;      // bit (slot*2 + tile): parity of that barrier's next completion
        for (uint32_t it = 0;; ++it) {
            const int w = next_work(it);
            if (w < 0) break;
            const WorkItem wi = decode_work(w, p);
            const uint32_t slot = it & 1u;
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const int q_start = wi.q0 + t * kBlockM;
                if (t == 1 && !wi.tile1) continue;                      // tile absent: nobody arrives, nothing to store
                const uint32_t b = slot * 2 + t;
                mbar_wait(bar_o_staged + 8 * b, (staged_parity >> b) & 1u, 60 + t);
                staged_parity ^= 1u << b;
                if (lane == 0 && !p.partial_mode) {                     // one fixed lane: bulk-group state is per thread
#pragma unroll
                    for (int pn = 0; pn < C::kPanels; pn++)
                        tma_store_3d(&tmO, sQ + (slot * 2 + t) * C::kTileBytes + pn * C::kPanelBytes, pn * 64, q_start, wi.bh);
                    tma_store_commit();
                }
            }
            if (lane == 0) {
                tma_store_wait_read<0>();
                mbar_arrive(bar_q_empty + 8 * slot);
            }
            __syncwarp();
        }
        // (exiting on wait_group.read alone -- the writes complete with the grid -- measured the same: profiles/r02_c26_*)
        if (lane == 0) tma_store_wait_all<0>();
        __syncwarp();
    }
    } else {
        setmaxnreg_inc<kRegsSoftmax>();
        // =============================== softmax / correction / epilogue ===============================
        const int t = warp >> 2;                               // which Q tile of the pair
        const int row_in_tile = (warp & 3) * 32 + lane;        // TMEM lane == S/O row
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
#ifndef FA_NO_PIN_ADDR
        // pinned (pin_u32): ptxas otherwise rebuilds both addresses in every tile from %tid / the shared-window base
        // (S2R, LEA, LOP3 ... ~10 instructions per tile on each softmax warp): +0.6 % at N=8192, +2.8 % at causal N=2048
        const uint32_t tS = pin_u32(tmem_base + lane_base + (t ? C::kTmemS1 : C::kTmemS0), p.zero);
        const uint32_t my_p_full = pin_u32(bar_p_full + 32 * t, p.zero);        // + 8 * piece
#else
        const uint32_t tS = tmem_base + lane_base + (t ? C::kTmemS1 : C::kTmemS0);
        const uint32_t my_p_full = bar_p_full + 32 * t;        // + 8 * piece
#endif
        const uint32_t tO = tmem_base + lane_base + (t ? C::kTmemO1 : C::kTmemO0);
        const uint32_t my_s_full = bar_s_full + 8 * t;
        const uint32_t my_o_full = bar_o_full + 8 * t;
        const uint32_t my_o_half = bar_o_half + 8 * t;
        uint32_t s_phase = 0;
        uint32_t pv_count = 0;   // P tiles handed to the MMA warp so far == o_full completions expected
        // p.scale_log2 pinned in a register: ptxas otherwise re-loads it from the constant bank in every tile, between the
        // row max and the first exponential (LDC + scoreboard wait on the S -> P chain).  The OR with (clock & p.zero) --
        // p.zero is always 0 -- makes the value opaque, so it cannot be rematerialised: +0.6-1.3 % (profiles/r02_c21_*)
        float sl2 = p.scale_log2;
#ifndef FA_NO_SCALE_REG
        sl2 = __uint_as_float(pin_u32(__float_as_uint(sl2), p.zero));
#endif
        uint32_t epi_count = 0;  // items handed to the epilogue warpgroup so far

        for (uint32_t it = 0;; ++it) {
            const int w = next_work(it);
            if (w < 0) break;
            const WorkItem wi = decode_work(w, p);
            const int q_start = p.split ? wi.q0 : wi.q0 + t * kBlockM;   // split mode: both slots work on the one Q tile
            if (t == 1 && !wi.tile1 && !p.split) continue;     // this Q tile does not exist (the store warp knows)
            const int n_t = t ? wi.n1 : wi.n0;
            const int row = q_start + row_in_tile;             // local query row
            // keys [0, lim) are visible to this row
            long long lim_ll = p.causal ? (long long)row + p.shift + 1 : (long long)p.Nkv;
            if (lim_ll > p.Nkv) lim_ll = p.Nkv;
            if (lim_ll < 0) lim_ll = 0;
            const int lim = (int)lim_ll;

            float m_ref = -INFINITY, l_run = 0.f;
#ifdef FA_TIMING
            const long long ti0 = clock64();
#endif
            for (int j = 0; j < n_t; j++) {
#if !defined(FA_NO_WAIT_IN_VARIANT) && !defined(FA_STREAM) && !defined(FA_SUM_GUARD)
                // The wait for S sits inside each variant's branch, directly in front of that variant's code: the warp spins
                // in the cache lines that precede the body it is about to run instead of jumping into a cold line once S
                // arrives (stall_no_inst at the head of softmax_tile: 4 % of the softmax warps' samples, r02_final ncu capture).
                {
                    const int k0 = (p.split ? 2 * j + t : j) * kBlockN;
                    const bool need_mask = (k0 + kBlockN > p.Nkv) || (p.causal && k0 + kBlockN - 1 > q_start + p.shift);
#ifdef FA_TIMING
                    const long long tw0 = clock64();
                    long long tw1 = 0;
#define FA_TW1 tw1 = clock64();
#else
#define FA_TW1
#endif
#ifdef FA_SPEC
                    // masked tiles and a row's first tile take the classic form (mask limit kBlockN = no mask), every other
                    // tile the speculative one: two softmax bodies in the kernel, as before
                    if (need_mask || j == 0) {
                        mbar_wait(my_s_full, s_phase, 20 + t);
                        tc_fence_after();
                        FA_TW1
                        softmax_tile<D, true, kPoly, kBF16>(p, tS, tO, my_p_full, my_o_full, my_o_half, need_mask ? lim - k0 : kBlockN,
                                                            j > 0, pv_count, m_ref, l_run, sl2);
                    } else {
                        mbar_wait(my_s_full, s_phase, 22 + t);
                        tc_fence_after();
                        FA_TW1
                        softmax_tile_spec<D, false, kPoly, kBF16>(p, tS, tO, my_p_full, my_o_full, kBlockN, pv_count, m_ref, l_run, sl2);
                    }
#else
                    if (need_mask) {
                        mbar_wait(my_s_full, s_phase, 20 + t);
                        tc_fence_after();
                        FA_TW1
                        softmax_tile<D, true, kPoly, kBF16>(p, tS, tO, my_p_full, my_o_full, my_o_half, lim - k0, j > 0, pv_count, m_ref, l_run, sl2);
                    } else {
                        mbar_wait(my_s_full, s_phase, 22 + t);
                        tc_fence_after();
                        FA_TW1
                        softmax_tile<D, false, kPoly, kBF16>(p, tS, tO, my_p_full, my_o_full, my_o_half, kBlockN, j > 0, pv_count, m_ref, l_run, sl2);
                    }
#endif
#undef FA_TW1
#ifdef FA_TIMING
                    if (j == 0 && lane == 0 && (warp & 3) == 0) {     // item start -> first S of the item
                        atomicAdd(&g_timing[8 + t * 2], (unsigned long long)(tw1 - ti0));
                        atomicAdd(&g_timing[9 + t * 2], 1ull);
                        if (it == 0) atomicAdd(&g_timing[12 + t], (unsigned long long)(tw1 - k_c0_all));   // kernel start -> first S
                    }
#ifdef FA_TIMING_ALL
                    if (lane == 0 && (warp & 3) == 0 && j > 0) {                   // short sequences: every tile but the first
#else
                    if (lane == 0 && (warp & 3) == 0 && j > 0 && (j & 7) == 0) {   // sampled: 1 tile in 8
#endif
                        const long long tw2 = clock64();
                        atomicAdd(&g_timing[t * 3 + 0], (unsigned long long)(tw1 - tw0));
                        atomicAdd(&g_timing[t * 3 + 1], (unsigned long long)(tw2 - tw1));
                        atomicAdd(&g_timing[t * 3 + 2], 1ull);
                    }
#endif
                    s_phase ^= 1u;
                    ++pv_count;
                    continue;
                }
#endif
#ifdef FA_TIMING
                const long long tw0 = clock64();
#endif
                mbar_wait(my_s_full, s_phase, 20 + t);
                s_phase ^= 1u;
                tc_fence_after();
#ifdef FA_TIMING
                const long long tw1 = clock64();
                if (j == 0 && lane == 0 && (warp & 3) == 0) {     // item start -> first S of the item
                    atomicAdd(&g_timing[8 + t * 2], (unsigned long long)(tw1 - ti0));
                    atomicAdd(&g_timing[9 + t * 2], 1ull);
                    if (it == 0) atomicAdd(&g_timing[12 + t], (unsigned long long)(tw1 - k_c0_all));   // kernel start -> first S
                }
#endif
                const int k0 = (p.split ? 2 * j + t : j) * kBlockN;        // split mode: slot t owns KV tiles t, t+2, ...
                const bool need_mask = (k0 + kBlockN > p.Nkv) || (p.causal && k0 + kBlockN - 1 > q_start + p.shift);
#if !defined(FA_STREAM) || defined(FA_SUM_GUARD)
                if (need_mask)
                    softmax_tile<D, true, kPoly, kBF16>(p, tS, tO, my_p_full, my_o_full, my_o_half, lim - k0, j > 0, pv_count, m_ref, l_run, sl2);
                else
                    softmax_tile<D, false, kPoly, kBF16>(p, tS, tO, my_p_full, my_o_full, my_o_half, kBlockN, j > 0, pv_count, m_ref, l_run, sl2);
#else
                // streamed softmax; a tile whose scores outgrew the reference max by more than 2^15 (warp vote)
                // is redone by the classic form, which takes the whole row's max first
                const bool last = j + 1 == n_t;
                const bool done = need_mask
                    ? softmax_tile_stream<D, true, kPoly, kBF16>(p, tS, tO, my_p_full, my_o_full, lim - k0, last, pv_count, m_ref, l_run)
                    : softmax_tile_stream<D, false, kPoly, kBF16>(p, tS, tO, my_p_full, my_o_full, kBlockN, last, pv_count, m_ref, l_run);
                if (!done)
                    softmax_tile<D, true, kPoly, kBF16>(p, tS, tO, my_p_full, my_o_full, my_o_half,
                                                        need_mask ? lim - k0 : kBlockN, j > 0, pv_count, m_ref, l_run, sl2);
#endif
                ++pv_count;
#ifdef FA_TIMING
#ifdef FA_TIMING_ALL
                if (lane == 0 && (warp & 3) == 0 && j > 0) {                   // short sequences: every tile but the first
#else
                if (lane == 0 && (warp & 3) == 0 && j > 0 && (j & 7) == 0) {   // sampled: 1 tile in 8
#endif
                    const long long tw2 = clock64();
                    atomicAdd(&g_timing[t * 3 + 0], (unsigned long long)(tw1 - tw0));
                    atomicAdd(&g_timing[t * 3 + 1], (unsigned long long)(tw2 - tw1));
                    atomicAdd(&g_timing[t * 3 + 2], 1ull);
                }
#endif
            }

            // ---- epilogue: O_t / l -> fp16 -> global (or the partial-state format) ----
#ifdef FA_TIMING
            const long long ti_ep = clock64();
#endif
            if (epi) {
                // The epilogue warpgroup drains O_t; this row's part is its sum.  The slot is free once that warpgroup has
                // finished the tile's previous item (always the case by now unless this item had no visible key at all).
                if (epi_count > 0u) mbar_wait(bar_o_free + 8 * t, (epi_count - 1u) & 1u, 34 + t);
                ++epi_count;
                st_shared_f32(smem_base + C::kMlOffset + (t * kBlockM + row_in_tile) * 4, l_run);
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_epi_full + 8 * t);
#ifdef FA_TIMING
                if (lane == 0 && (warp & 3) == 0) {
                    atomicAdd(&g_timing[14 + t], (unsigned long long)(clock64() - ti_ep));
                    atomicAdd(&g_timing[16 + t], (unsigned long long)(clock64() - ti0));   // whole item
                }
#endif
                continue;
            }
            if (n_t > 0) {
                mbar_wait(my_o_full, (pv_count - 1u) & 1u, 30 + t);
                tc_fence_after();
            } else {
                // No key visible to this tile (a K/V block entirely in its future): nothing above waited for the
                // item's Q pair to land, and the epilogue stages its rows in that very buffer.  Without this wait the
                // tail of the TMA load overwrites the staged rows of the tile's last warp (seen as a few wrong rows in
                // the in-place accumulate sequence, tests/harness/accumulate_stress.py).
                mbar_wait(bar_q_full + 8 * (it & 1u), (it >> 1) & 1u, 32 + t);
            }
            const bool row_ok = row < p.Nq;
            const size_t grow = (size_t)wi.bh * p.Nq + row;
            if (p.split) {
                // ---- split mode: merge the two slots' partial states (FA.cu:575-597) through shared memory ----
                // Row r of slot 1's un-normalised fp32 O travels in the four (D=64: two) 128-byte lines "row r" of the
                // item's two Q tile buffers, 16-byte chunks XOR-swizzled by r & 7 like every other tile here; slot 0's
                // thread r reads them back and then writes its fp16 output row over lines it has already consumed.
                const uint32_t buf = sQ + ((it & 1u) * 2) * C::kTileBytes;
                const uint32_t ml_slot = smem_base + C::kMlOffset + row_in_tile * 8;
                auto line_addr = [&](int c4) {      // c4: index of a 16-byte chunk of the row's fp32 data (4 floats)
                    const int line = c4 >> 3;       // 8 chunks per 128-byte line; lines: buf0.panel0, buf0.panel1, buf1.panel0, ...
                    return buf + (line / C::kPanels) * C::kTileBytes + (line % C::kPanels) * C::kPanelBytes + row_in_tile * 128 +
                           (((c4 & 7) ^ (row_in_tile & 7)) << 4);
                };
                if (t == 1) {
                    if (n_t > 0) {
#pragma unroll
                        for (int c = 0; c < D; c += 32) {
                            uint32_t o[32];
                            tmem_ld_x32(tO + c, o);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; i += 4) st_shared_v4(line_addr((c + i) >> 2), o[i], o[i + 1], o[i + 2], o[i + 3]);
                        }
                        st_shared_v2(ml_slot, __float_as_uint(m_ref), __float_as_uint(l_run));
                        tc_fence_before();
                        bar_arrive(1 + (warp & 3), 64);          // hand-over to warp (warp & 3) of slot 0: same 32 rows
                    }
                    continue;                                    // slot 1 neither stores nor stages
                }
                float w0 = 1.f, w1 = 0.f, l_tot = l_run;
                const bool have1 = wi.n1 > 0;
                if (have1) {
                    bar_sync(1 + (warp & 3), 64);
                    float m1, l1;
                    ld_shared_v2f(ml_slot, m1, l1);
                    const float m_max = fmaxf(m_ref, m1);
                    w0 = (m_ref == -INFINITY) ? 0.f : ex2_approx((m_ref - m_max) * p.scale_log2);
                    w1 = (m1 == -INFINITY) ? 0.f : ex2_approx((m1 - m_max) * p.scale_log2);
                    l_tot = l_run * w0 + l1 * w1;
                }
                const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;      // FA.cu:502-503, 590
                const float s0 = w0 * inv, s1 = w1 * inv;
                const uint32_t stage = buf + row_in_tile * 128;
#pragma unroll
                for (int c = 0; c < D; c += 32) {
                    uint32_t o[32];
                    if (n_t > 0) {
                        tmem_ld_x32(tO + c, o);
                        tmem_wait_ld();
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i++) o[i] = 0u;
                    }
                    float f[32];
#pragma unroll
                    for (int i = 0; i < 32; i++) f[i] = __uint_as_float(o[i]) * s0;
                    if (have1) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 x = ld_shared_v4f(line_addr((c + i) >> 2));
                            f[i] += x.x * s1; f[i + 1] += x.y * s1; f[i + 2] += x.z * s1; f[i + 3] += x.w * s1;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        const int col = c + i;
                        const uint32_t addr = stage + (col >> 6) * C::kPanelBytes + ((((col & 63) >> 3) ^ (row_in_tile & 7)) << 4);
                        st_shared_v4(addr, pack_16x2<kBF16>(f[i], f[i + 1]), pack_16x2<kBF16>(f[i + 2], f[i + 3]),
                                     pack_16x2<kBF16>(f[i + 4], f[i + 5]), pack_16x2<kBF16>(f[i + 6], f[i + 7]));
                    }
                }
                fence_proxy_async_smem();   // generic-proxy writes -> visible to the TMA store
            } else if (!p.partial_mode) {
                // fp16 row -> staging tile = this item's Q tile buffer (every QK^T of the item has retired: o_full covers them)
                if (!kEpiWg)
                    stage_o_tile<D, kBF16>(tO, sQ + ((it & 1u) * 2 + t) * C::kTileBytes, row_in_tile, l_run, n_t > 0);
            } else {
                // partial state (FA.cu:460-496): un-normalised fp32 O, (m, l) with m in the
                // scaled-score (natural-log) domain; merge algebra of FA.cu:575-597 when accumulating
                float m_out = (m_ref == -INFINITY) ? -FLT_MAX : m_ref * p.scale;
                float l_out = l_run;
                float w_new = 1.f, w_old = 0.f;
                if (p.accumulate && row_ok) {
                    // written by an earlier launch, possibly from another SM: read at L2 like the O partials below
                    const float2 ml_old = __ldcg(reinterpret_cast<const float2*>(p.ml + grow * 2));
                    const float m_old = ml_old.x;
                    const float l_old = ml_old.y;
                    const float m_max = fmaxf(m_old, m_out);
                    const float kLog2e = 1.4426950408889634f;
                    w_old = (m_old <= -FLT_MAX) ? 0.f : ex2_approx((m_old - m_max) * kLog2e);
                    w_new = (m_out <= -FLT_MAX) ? 0.f : ex2_approx((m_out - m_max) * kLog2e);
                    l_out = l_old * w_old + l_run * w_new;
                    m_out = m_max;
                }
                // The fp32 rows go through this warp's slice of the item's idle Q tile buffer and leave it
                // transposed: one thread per row would touch 16 bytes every 4*D bytes of global memory (the
                // accumulate form cost +25 % per ring hop kernel that way), whereas 4*D/2-byte row segments
                // per group of lanes are whole sectors.  D/2 columns per pass; 16-byte chunks XOR-swizzled by
                // row so that both the row-wise writes and the transposed reads are bank-conflict free.
                constexpr int kW = D / 2;                  // fp32 columns per pass
                constexpr int kCPR = kW / 4;               // 16-byte chunks per staged row = lanes per row on the way out
                constexpr int kRPI = 32 / kCPR;            // rows one warp instruction covers on the way out
                const uint32_t wstage = sQ + ((it & 1u) * 2 + t) * C::kTileBytes + (uint32_t)(warp & 3) * (32 * kW * 4);
                const int warp_row0 = q_start + (warp & 3) * 32;
                float* const gbase = p.o_partial + ((size_t)wi.bh * p.Nq + warp_row0) * D;
#pragma unroll
                for (int ps = 0; ps < 2; ps++) {
#pragma unroll
                    for (int c = 0; c < kW; c += 32) {
                        uint32_t o[32];
                        if (n_t > 0) {
                            tmem_ld_x32(tO + ps * kW + c, o);
                            tmem_wait_ld();
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; i++) o[i] = 0u;
                        }
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const int ch = (c + i) >> 2;
                            st_shared_v4(wstage + lane * (kW * 4) + ((ch ^ (lane & (kCPR - 1))) << 4),
                                         __float_as_uint(__uint_as_float(o[i]) * w_new), __float_as_uint(__uint_as_float(o[i + 1]) * w_new),
                                         __float_as_uint(__uint_as_float(o[i + 2]) * w_new), __float_as_uint(__uint_as_float(o[i + 3]) * w_new));
                        }
                    }
                    __syncwarp();
                    const int ch = lane & (kCPR - 1);
                    // all loads of the old partial first (they are independent: one round trip, not kCPR)
                    float4 old[kCPR];
#pragma unroll
                    for (int i = 0; i < kCPR; i++) {
                        const int rr = i * kRPI + lane / kCPR;          // row within this warp's 32
                        old[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.accumulate && warp_row0 + rr < p.Nq)
                            old[i] = __ldcg(reinterpret_cast<const float4*>(gbase + (size_t)rr * D + ps * kW + ch * 4));
                    }
#pragma unroll
                    for (int i = 0; i < kCPR; i++) {
                        const int rr = i * kRPI + lane / kCPR;
                        float4 v = ld_shared_v4f(wstage + rr * (kW * 4) + ((ch ^ (rr & (kCPR - 1))) << 4));
                        const float wo = __shfl_sync(0xffffffffu, w_old, rr);
                        v.x += old[i].x * wo; v.y += old[i].y * wo;
                        v.z += old[i].z * wo; v.w += old[i].w * wo;
                        if (warp_row0 + rr < p.Nq)
                            *reinterpret_cast<float4*>(gbase + (size_t)rr * D + ps * kW + ch * 4) = v;
                    }
                    __syncwarp();   // the slice is rewritten by the next pass
                }
                if (row_ok) {
                    p.ml[grow * 2 + 0] = m_out;
                    p.ml[grow * 2 + 1] = l_out;
                }
            }
            // O_t / S_t are free again: the next item's first P arrival orders after these reads
            tc_fence_before();
            __syncwarp();
            // store warp: tile staged (or written out, in partial mode).  One barrier per (Q slot, tile):
            // it cannot complete twice before the store warp has seen the first completion, because the
            // slot's next use needs the Q load that the store warp itself releases.
            if (lane == 0) mbar_arrive(bar_o_staged + 8 * ((it & 1u) * 2 + t));
#ifdef FA_TIMING
            if (lane == 0 && (warp & 3) == 0) {
                atomicAdd(&g_timing[14 + t], (unsigned long long)(clock64() - ti_ep));
                atomicAdd(&g_timing[16 + t], (unsigned long long)(clock64() - ti0));   // whole item
            }
#endif
        }
    }

    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
#ifdef FA_TIMING
    if (threadIdx.x == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        atomicAdd(&g_timing[20], (unsigned long long)(clock64() - k_c0));   // CTA lifetime, SM cycles
        atomicAdd(&g_timing[21], t1 - k_t0);                                // CTA lifetime, ns
        atomicAdd(&g_timing[22], 1ull);
    }
#endif
    // a waiter of this CTA gave up: leave the record where the launcher's next call finds it
    if (threadIdx.x == 0 && *reinterpret_cast<volatile unsigned int*>(watchdog_block_flag()) != 0u) watchdog_publish();
    if (warp == kMmaWarp) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// fp16 O = o_partial / l  (final step of the FA.cu:575-597 merge)
__global__ void fa_finalize_kernel(const float* __restrict__ o_partial, const float* __restrict__ ml,
                                   __half* __restrict__ o, long long rows, int D) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 4 elements
    const int per_row = D / 4;
    const long long r = idx / per_row;
    if (r >= rows) return;
    const int c = (int)(idx % per_row) * 4;
    const float l = ml[r * 2 + 1];
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    const float4 v = *reinterpret_cast<const float4*>(o_partial + r * D + c);
    __half2 a = __floats2half2_rn(v.x * inv, v.y * inv);
    __half2 b = __floats2half2_rn(v.z * inv, v.w * inv);
    uint2 out;
    out.x = *reinterpret_cast<uint32_t*>(&a);
    out.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(o + r * D + c) = out;
}

// fp16 O = sum_s w_s O_s / sum_s w_s l_s with w_s = exp(m_s - max_s m_s): the reference's
// flash_attention_splitk_merge (FA.cu:559-598, never launched there) over `splits` partial states laid
// out [split][row][D] / [split][row][2].  Used by ring context parallelism: every chunk pair writes its
// own partial (write-only) and one pass merges them, instead of a read-modify-write per hop.
__global__ void fa_merge_kernel(const float* __restrict__ o_partial, const float* __restrict__ ml,
                                __half* __restrict__ o, long long rows, int D, int splits) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 4 elements
    const int per_row = D / 4;
    const long long r = idx / per_row;
    if (r >= rows) return;
    const int c = (int)(idx % per_row) * 4;
    // both loops unrolled by 4: the loads of four partial states are in flight together (the merge is pure HBM traffic)
    float m_max = -FLT_MAX;
#pragma unroll 4
    for (int s = 0; s < splits; s++) m_max = fmaxf(m_max, ml[((long long)s * rows + r) * 2]);
    float l = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int s = 0; s < splits; s++) {
        const float m_s = ml[((long long)s * rows + r) * 2];
        const float l_s = ml[((long long)s * rows + r) * 2 + 1];
        const float w = (m_s <= -FLT_MAX) ? 0.f : exp2f((m_s - m_max) * 1.4426950408889634f);   // FA.cu:583-586
        const float4 v = *reinterpret_cast<const float4*>(o_partial + ((long long)s * rows + r) * D + c);
        l += w * l_s;
        acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
    }
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    __half2 a = __floats2half2_rn(acc.x * inv, acc.y * inv);
    __half2 b = __floats2half2_rn(acc.z * inv, acc.w * inv);
    uint2 out;
    out.x = *reinterpret_cast<uint32_t*>(&a);
    out.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(o + r * D + c) = out;
}

}  // namespace fa
