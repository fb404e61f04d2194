// fa_fwd_sm100.cuh -- FlashAttention forward for B200 (sm_100a): persistent, warp-specialised,
// TMA -> 128B-swizzled smem -> tcgen05.mma with S/P/O in tensor memory.
//
// Replaces the reference's device path flash_attention_v9<...> (flash_attention.cu:67-554):
//   FA.cu:103-112  block->(bh, q-block) mapping, GRID_SWAP    -> persistent work loop, heavy-first
//   FA.cu:145-159  Q fragments in registers                   -> Q tile pair resident in smem (TMA)
//   FA.cu:417-447  synchronous K/V tile load + 2 barriers     -> producer warp, mbarrier ring
//   FA.cu:188-233  DO_QK_MATMUL (mma.sync m16n8k16)           -> tcgen05.mma SS, S in TMEM
//   FA.cu:235-288  DO_SOFTMAX (quad shuffles, eager rescale)  -> one thread per row, lazy rescale
//   FA.cu:290-334  DO_PV_MATMUL (P in registers)              -> P fp16 in TMEM, tcgen05.mma TS
//   FA.cu:497-553  epilogue via smem                          -> TMEM -> registers -> global
//   FA.cu:460-496  split-K partial epilogue (dead code there) -> partial mode used by ring CP
//
// CTA = 384 threads:  warps 0-3  softmax/correction/epilogue for Q tile 0 (128 rows)
//                     warps 4-7  same for Q tile 1
//                     warp 8     TMEM allocator + tcgen05.mma issuer (one lane)
//                     warp 9     TMA producer (one lane)
//                     warps 10,11 idle (pad the third warpgroup)
// One CTA per SM; each CTA loops over work items (bh, pair of 128-row Q tiles).
//
// TMEM (512 columns x 128 lanes x 32 bit): S0 [0,128) S1 [128,256) O0 [256,256+D) O1 [256+D,256+2D).
// The KV axis is processed in 64-column sub-tiles u = 0,1,2,...: sub-tile u of Q tile t lives in half
// (u & 1) of S_t, so S is double-buffered per Q tile: QK_t(u+2) is issued right behind PV_t(u) and a
// softmax warpgroup never waits for "its own" MMAs -- while it turns S_t(u) into P_t(u), S_t(u+1) is
// already in TMEM.  P_t(u) (fp16, two per column) overwrites columns [0,32) of its S half once the
// owning thread has read its S row.  Tensor-pipe order: PV0(u) QK0(u+2) PV1(u) QK1(u+2).
// K/V stay 128-row smem tiles (one TMA ring entry each); a sub-tile is half of one.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "sm100_ptx.cuh"

namespace fa {

using namespace sm100;

constexpr int kBlockM = 128;      // Q rows per tile  (UMMA M)
constexpr int kBlockN = 128;      // K/V rows per smem tile (one TMA ring entry)
constexpr int kSubN = 64;         // K/V rows per MMA/softmax sub-tile (UMMA N of QK^T, K extent of PV)
constexpr int kNumThreads = 384;
constexpr int kMmaWarp = 8;
constexpr int kLoadWarp = 9;
constexpr int kTmemCols = 512;
constexpr int kRegsSoftmax = 208;   // setmaxnreg: softmax warpgroups grow, the producer/MMA warpgroup shrinks
constexpr int kRegsOther = 80;      // 2*128*208 + 128*80 = 63488 <= 168 (launch) * 384
constexpr float kRescaleThreshold = 8.0f;  // lazy rescale: tolerate P up to 2^8 before moving the reference max

template <int D>
struct Cfg {
    static_assert(D == 64 || D == 128, "head_dim must be 64 or 128");
    static constexpr int kPanels = D / 64;                 // 128-byte swizzle panels per row
    static constexpr int kPanelBytes = 128 * 128;          // 128 rows x 128 B
    static constexpr int kTileBytes = kPanels * kPanelBytes;
    static constexpr int kStages = (D == 128) ? 5 : 8;     // K/V ring entries (one tile each)
    static constexpr int kSmemQ = 2 * kTileBytes;
    static constexpr int kSmemKV = kStages * kTileBytes;
    static constexpr int kBarOffset = kSmemQ + kSmemKV;
    static constexpr int kNumBars = 2 + 2 * kStages + 12;
    static constexpr int kSmemBytes = kBarOffset + kNumBars * 8 + 16 + 1024;  // +1024: manual alignment slack
    static constexpr int kTmemS0 = 0, kTmemS1 = 128, kTmemO0 = 256, kTmemO1 = 256 + D;
    static constexpr uint32_t kIdescQK = umma_idesc_f16(kBlockM, kSubN, 0, 0);
    static constexpr uint32_t kIdescPV = umma_idesc_f16(kBlockM, D, 0, 1);  // V is MN-major ([kv][d], d contiguous)
};

struct Params {
    __half* o;          // fp16 output [BH, Nq, D]            (partial_mode == 0)
    float* o_partial;   // fp32 un-normalised [BH*Nq, D]       (partial_mode == 1; FA.cu:460-496 format)
    float* ml;          // (m, l) pairs [BH*Nq, 2]
    int Nq, Nkv, BH;
    int causal;
    int shift;          // q_offset - kv_offset: key c visible to query r iff c <= r + shift
    int nqp;            // Q tile pairs per head = ceil(Nq / 256)
    int total_work;     // BH * nqp
    int partial_mode;
    int accumulate;
    float scale;        // 1/sqrt(D)
    float scale_log2;   // scale * log2(e)
};

// ---- work decomposition (shared by host tests and every warp role) ----
struct WorkItem {
    int bh, q0;      // head index, first local query row of the pair
    int n0, n1;      // 64-wide KV sub-tiles the two Q tiles need (0 = nothing visible / tile absent)
};
__host__ __device__ inline int kv_trip_count(int q_start, int Nq, int Nkv, int causal, int shift) {
    if (q_start >= Nq) return 0;
    long long vis = Nkv;                              // keys [0, vis) are visible to the tile's last row
    if (causal) {
        int last_row = q_start + kBlockM - 1;
        if (last_row > Nq - 1) last_row = Nq - 1;
        vis = (long long)last_row + shift + 1;
        if (vis <= 0) return 0;
        if (vis > Nkv) vis = Nkv;
    }
    return (int)((vis + kSubN - 1) / kSubN);
}
// Work order: heads outermost (the CTAs running concurrently share a few heads' K/V in L2),
// heaviest Q pair first inside a head (causal: the last pair sees the most keys).
__host__ __device__ inline WorkItem decode_work(int w, const Params& p) {
    WorkItem it;
    it.bh = w / p.nqp;
    const int qp = p.nqp - 1 - (w % p.nqp);
    it.q0 = qp * 2 * kBlockM;
    it.n0 = kv_trip_count(it.q0, p.Nq, p.Nkv, p.causal, p.shift);
    it.n1 = kv_trip_count(it.q0 + kBlockM, p.Nq, p.Nkv, p.causal, p.shift);
    return it;
}

struct Ring {
    uint32_t idx, phase;
    template <int kStages>
    __device__ __forceinline__ void advance() {
        if (++idx == (uint32_t)kStages) { idx = 0; phase ^= 1u; }
    }
};

// ---- softmax of one 128x64 S sub-tile; one thread owns one row ----
template <int D, bool kMask>
__device__ __forceinline__ void softmax_subtile(const Params& p, uint32_t tS, uint32_t tO, uint32_t bar_p_full,
                                                uint32_t bar_o_full, int lim_local, bool have_o,
                                                uint32_t pv_count, float& m_ref, float& l_run) {
    uint32_t s[kSubN];
    tmem_ld_x32(tS + 0, s + 0);
    tmem_ld_x32(tS + 32, s + 32);
    tmem_wait_ld();

    if (kMask) {
#pragma unroll
        for (int i = 0; i < kSubN; i++)
            if (i >= lim_local) s[i] = 0xff800000u;  // -inf
    }

    // row max: 3-input max (FMNMX3), four independent chains
    float mx0 = fmaxf(__uint_as_float(s[0]), __uint_as_float(s[1]));
    float mx1 = fmaxf(__uint_as_float(s[2]), __uint_as_float(s[3]));
    float mx2 = fmaxf(__uint_as_float(s[4]), __uint_as_float(s[5]));
    float mx3 = fmaxf(__uint_as_float(s[6]), __uint_as_float(s[7]));
#pragma unroll
    for (int i = 8; i < kSubN; i += 8) {
        mx0 = fmax3(mx0, __uint_as_float(s[i + 0]), __uint_as_float(s[i + 1]));
        mx1 = fmax3(mx1, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        mx2 = fmax3(mx2, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
        mx3 = fmax3(mx3, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
    }
    const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
    const float m_new = fmaxf(m_ref, m_tile);

    // Lazy rescale (replaces the reference's every-tile O *= alpha, FA.cu:267-270): the reference
    // max only moves when the true max has outgrown it by 2^kRescaleThreshold.
    const bool need = (m_new - m_ref) * p.scale_log2 > kRescaleThreshold;  // NaN (-inf - -inf) -> false
    if (__any_sync(0xffffffffu, need)) {
        if (have_o) {
            const float alpha = (m_new == -INFINITY) ? 1.0f : ex2_approx((m_ref - m_new) * p.scale_log2);
            const uint64_t alpha2 = pack_f32x2(alpha, alpha);
            // O_t holds PV(0..u-1); the last of them must have retired before we touch it.  PV(u-2)
            // and older are known to be complete (S(u) is ready and the tensor pipe is in order), so
            // the barrier of PV(u-1)'s parity is at most one phase behind: the parity test is exact.
            mbar_wait(bar_o_full, (pv_count - 1u) & 1u, 40);
            tc_fence_after();
#pragma unroll
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
    }

    const float m_used = (m_ref == -INFINITY) ? 0.0f : m_ref;
    const float neg = -m_used * p.scale_log2;
    const uint64_t scale2 = pack_f32x2(p.scale_log2, p.scale_log2);
    const uint64_t neg2 = pack_f32x2(neg, neg);
    uint64_t sum_a = 0ull, sum_b = 0ull;     // (0.f, 0.f)
    uint32_t pk[kSubN / 2];
#pragma unroll
    for (int i = 0; i < kSubN; i += 4) {
        float x0, x1, x2, x3;
        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(s[i + 0]), __uint_as_float(s[i + 1])), scale2, neg2), x0, x1);
        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(s[i + 2]), __uint_as_float(s[i + 3])), scale2, neg2), x2, x3);
        const float p0 = ex2_approx(x0), p1 = ex2_approx(x1), p2 = ex2_approx(x2), p3 = ex2_approx(x3);
        sum_a = add_f32x2(sum_a, pack_f32x2(p0, p1));      // row sum of the un-rounded p (FA.cu:273-279)
        sum_b = add_f32x2(sum_b, pack_f32x2(p2, p3));
        __half2 h01 = __floats2half2_rn(p0, p1);           // low half = even column
        __half2 h23 = __floats2half2_rn(p2, p3);
        pk[i / 2 + 0] = *reinterpret_cast<uint32_t*>(&h01);
        pk[i / 2 + 1] = *reinterpret_cast<uint32_t*>(&h23);
    }
    // P (fp16 A operand of PV) overwrites columns [0,32) of this S half
    tmem_st_x32(tS, pk);
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane_id() == 0) mbar_arrive(bar_p_full);           // one arrival per warp (barrier count 4)
    float a0, a1;
    unpack_f32x2(add_f32x2(sum_a, sum_b), a0, a1);
    l_run += a0 + a1;
}

template <int D>
__global__ void __launch_bounds__(kNumThreads, 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const Params p) {
    using C = Cfg<D>;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = smem_base;
    const uint32_t sKV = smem_base + C::kSmemQ;
    const uint32_t bars = smem_base + C::kBarOffset;
    const uint32_t bar_q_full = bars + 0;
    const uint32_t bar_q_empty = bars + 8;
    const uint32_t bar_kv_full = bars + 16;                       // [kStages]
    const uint32_t bar_kv_empty = bar_kv_full + 8 * C::kStages;   // [kStages]
    const uint32_t bar_s_full = bar_kv_empty + 8 * C::kStages;    // [tile][half]  index 2*t + h
    const uint32_t bar_p_full = bar_s_full + 32;                  // [tile][half]
    const uint32_t bar_o_full = bar_p_full + 32;                  // [tile][half]: PV_t(u) commits to half u & 1
    const uint32_t tmem_slot = bar_o_full + 32;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(bar_q_full, 1);
        mbar_init(bar_q_empty, 1);
        for (int i = 0; i < C::kStages; i++) {
            mbar_init(bar_kv_full + 8 * i, 1);
            mbar_init(bar_kv_empty + 8 * i, 1);
        }
        for (int i = 0; i < 4; i++) {
            mbar_init(bar_s_full + 8 * i, 1);
            mbar_init(bar_p_full + 8 * i, 4);         // one arrival per softmax warp of the tile
        }
        for (int i = 0; i < 4; i++) mbar_init(bar_o_full + 8 * i, 1);
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
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // The producer and MMA warps run their loops converged (all 32 lanes take the same branches
    // and waits); the instructions with side effects sit under elect_one().  Warp-uniform control
    // flow keeps descriptors and barrier addresses in uniform registers -- in a lane-divergent
    // region every UTCHMMA costs an extra ELECT / R2UR.BROADCAST sequence and the single issuing
    // thread becomes the bottleneck of the whole CTA (profiles/r01_v1_full_n8192_summary.txt).
    if (warp >= 8) {
    setmaxnreg_dec<kRegsOther>();   // each role's code must be dominated by its own setmaxnreg
    if (warp == kLoadWarp) {
        // =============================== TMA producer ===============================
        Ring ring{0u, 0u};
        uint32_t it = 0;
        for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++it) {
            const WorkItem wi = decode_work(w, p);
            const int nmax = ((wi.n0 > wi.n1 ? wi.n0 : wi.n1) + 1) >> 1;   // 128-row K/V tiles to stream
            const bool have_q1 = wi.q0 + kBlockM < p.Nq;
            mbar_wait(bar_q_empty, (it & 1u) ^ 1u, 1);   // previous item's QK^T MMAs retired
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_q_full, (have_q1 ? 2 : 1) * C::kTileBytes);
                for (int t = 0; t < (have_q1 ? 2 : 1); t++)
                    for (int pn = 0; pn < C::kPanels; pn++)
                        tma_load_3d(sQ + t * C::kTileBytes + pn * C::kPanelBytes, &tmQ, bar_q_full, pn * 64,
                                    wi.q0 + t * kBlockM, wi.bh);
            }
            __syncwarp();
            for (int j = 0; j < nmax; j++) {
                // ring order K_0 V_0 K_1 V_1 ... (the order the MMA warp releases them in)
#pragma unroll
                for (int kv = 0; kv < 2; kv++) {
                    const uint32_t full = bar_kv_full + 8 * ring.idx;
                    mbar_wait(bar_kv_empty + 8 * ring.idx, ring.phase ^ 1u, 2);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(full, C::kTileBytes);
                        const uint32_t dst = sKV + ring.idx * C::kTileBytes;
#pragma unroll
                        for (int pn = 0; pn < C::kPanels; pn++)
                            tma_load_3d(dst + pn * C::kPanelBytes, kv == 0 ? &tmK : &tmV, full, pn * 64,
                                        j * kBlockN, wi.bh);
                    }
                    __syncwarp();
                    ring.advance<C::kStages>();
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // =============================== tcgen05.mma issuer ===============================
        Ring rk{0u, 0u};              // ring entry holding K_j
        Ring rv{1u % C::kStages, 0u}; // ring entry holding V_j
        uint32_t it = 0;
        const uint32_t tS0 = tmem_base + C::kTmemS0, tS1 = tmem_base + C::kTmemS1;
        const uint32_t tO0 = tmem_base + C::kTmemO0, tO1 = tmem_base + C::kTmemO1;
        const uint64_t qdesc0 = umma_smem_desc(sQ, 16, 1024);
        const uint64_t qdesc1 = umma_smem_desc(sQ + C::kTileBytes, 16, 1024);

        // S_t[half h] = Q_t K_j[64h..64h+63]^T : N = 64, D/16 k-steps; k-step ks lives in panel ks/4 at
        // byte offset (ks%4)*32; the K sub-tile starts 64 rows (8192 B) into each panel
        auto issue_qk = [&](uint32_t tS, uint64_t qdesc, uint32_t k_smem, int h, uint32_t bar) {
            const uint64_t kdesc = umma_smem_desc(k_smem + h * (kSubN * 128), 16, 1024);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < D / 16; ks++) {
                    const uint64_t off = (uint64_t)(((ks >> 2) * C::kPanelBytes + (ks & 3) * 32) >> 4);
                    umma_ss(tS + h * kSubN, qdesc + off, kdesc + off, C::kIdescQK, ks > 0 ? 1u : 0u);
                }
                umma_commit(bar);
            }
            __syncwarp();
        };
        // O_t (+)= P_t[half h] V_j[64h..64h+63] : 4 k-steps of 16 kv rows; P k-step = 8 TMEM columns,
        // V k-step = 16 rows * 128 B
        auto issue_pv = [&](uint32_t tO, uint32_t tS, uint32_t v_smem, int h, bool accumulate, uint32_t bar) {
            const uint64_t vdesc = umma_smem_desc(v_smem + h * (kSubN * 128), C::kPanelBytes, 1024);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < kSubN / 16; ks++)
                    umma_ts(tO, tS + h * kSubN + ks * 8, vdesc + (uint64_t)((ks * 16 * 128) >> 4), C::kIdescPV,
                            (accumulate || ks > 0) ? 1u : 0u);
                umma_commit(bar);
            }
            __syncwarp();
        };
        auto commit = [&](uint32_t bar) {
            if (elect_one()) umma_commit(bar);
            __syncwarp();
        };

        uint32_t p_phase = 0u;   // bit (2*t + h): parity of the next P_t[half h] arrival
        for (int w = blockIdx.x; w < p.total_work; w += gridDim.x, ++it) {
            const WorkItem wi = decode_work(w, p);
            const int n0 = wi.n0, n1 = wi.n1;                 // sub-tiles per Q tile
            const int nsub = n0 > n1 ? n0 : n1;
            const int ntiles = (nsub + 1) >> 1;
            mbar_wait(bar_q_full, it & 1u, 10);
            tc_fence_after();
            if (ntiles > 0) {
                // prologue: sub-tiles 0 and 1 of both Q tiles come from K_0
                mbar_wait(bar_kv_full + 8 * rk.idx, rk.phase, 11);
                tc_fence_after();
                const uint32_t k_smem = sKV + rk.idx * C::kTileBytes;
                if (0 < n0) issue_qk(tS0, qdesc0, k_smem, 0, bar_s_full + 0);
                if (0 < n1) issue_qk(tS1, qdesc1, k_smem, 0, bar_s_full + 16);
                if (1 < n0) issue_qk(tS0, qdesc0, k_smem, 1, bar_s_full + 8);
                if (1 < n1) issue_qk(tS1, qdesc1, k_smem, 1, bar_s_full + 24);
                commit(bar_kv_empty + 8 * rk.idx);
                rk.advance<C::kStages>(); rk.advance<C::kStages>();
            }
            // Q is free for the next item as soon as the last QK^T of this one has retired
            if (ntiles <= 1) commit(bar_q_empty);
            for (int j = 0; j < ntiles; j++) {
                const bool has_next = j + 1 < ntiles;
                mbar_wait(bar_kv_full + 8 * rv.idx, rv.phase, 12);
                const uint32_t v_smem = sKV + rv.idx * C::kTileBytes;
                const uint32_t k_smem = sKV + rk.idx * C::kTileBytes;   // K_{j+1}
                if (has_next) {
                    mbar_wait(bar_kv_full + 8 * rk.idx, rk.phase, 15);
                    tc_fence_after();
                }
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int u = 2 * j + h;
                    // ---- Q tile 0: PV0(u), QK0(u+2)
                    if (u < n0) {
                        mbar_wait(bar_p_full + 8 * h, (p_phase >> h) & 1u, 13);
                        p_phase ^= 1u << h;
                        tc_fence_after();
                        issue_pv(tO0, tS0, v_smem, h, u > 0, bar_o_full + 8 * h);
                    }
                    if (u + 2 < n0) issue_qk(tS0, qdesc0, k_smem, h, bar_s_full + 8 * h);
                    // ---- Q tile 1: PV1(u), QK1(u+2)
                    if (u < n1) {
                        mbar_wait(bar_p_full + 16 + 8 * h, (p_phase >> (2 + h)) & 1u, 14);
                        p_phase ^= 1u << (2 + h);
                        tc_fence_after();
                        issue_pv(tO1, tS1, v_smem, h, u > 0, bar_o_full + 16 + 8 * h);
                    }
                    if (u + 2 < n1) issue_qk(tS1, qdesc1, k_smem, h, bar_s_full + 16 + 8 * h);
                }
                commit(bar_kv_empty + 8 * rv.idx);
                rv.advance<C::kStages>(); rv.advance<C::kStages>();
                if (has_next) {
                    commit(bar_kv_empty + 8 * rk.idx);
                    rk.advance<C::kStages>(); rk.advance<C::kStages>();
                    if (j + 2 == ntiles) commit(bar_q_empty);   // K_{ntiles-1} fed the last QK^T of this item
                }
            }
        }
    }
    } else {
        setmaxnreg_inc<kRegsSoftmax>();
        // =============================== softmax / correction / epilogue ===============================
        const int t = warp >> 2;                               // which Q tile of the pair
        const int row_in_tile = (warp & 3) * 32 + lane;        // TMEM lane == S/O row
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem_base + lane_base + (t ? C::kTmemS1 : C::kTmemS0);
        const uint32_t tO = tmem_base + lane_base + (t ? C::kTmemO1 : C::kTmemO0);
        const uint32_t my_s_full = bar_s_full + 16 * t;        // + 8 * half
        const uint32_t my_p_full = bar_p_full + 16 * t;        // + 8 * half
        const uint32_t my_o_full = bar_o_full + 16 * t;        // + 8 * half
        uint32_t s_phase = 0;    // bit h: parity of the next S_t[half h] completion
        // P sub-tiles of parity h handed to the MMA warp so far == completions expected on o_full[t][h].
        // Two barriers because two PVs can be outstanding at once (sub-tiles u-1 and u): with a single
        // barrier a parity wait could not tell "both done" from "neither done".
        uint32_t pv_count[2] = {0u, 0u};

        for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
            const WorkItem wi = decode_work(w, p);
            const int q_start = wi.q0 + t * kBlockM;
            if (q_start >= p.Nq) continue;                     // this Q tile does not exist
            const int n_t = t ? wi.n1 : wi.n0;
            const int row = q_start + row_in_tile;             // local query row
            // keys [0, lim) are visible to this row
            long long lim_ll = p.causal ? (long long)row + p.shift + 1 : (long long)p.Nkv;
            if (lim_ll > p.Nkv) lim_ll = p.Nkv;
            if (lim_ll < 0) lim_ll = 0;
            const int lim = (int)lim_ll;

            float m_ref = -INFINITY, l_run = 0.f;
#pragma unroll 1
            for (int u0 = 0; u0 < n_t; u0 += 2) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int u = u0 + h;
                    if (u < n_t) {
                        mbar_wait(my_s_full + 8 * h, (s_phase >> h) & 1u, 20 + t);
                        s_phase ^= 1u << h;
                        tc_fence_after();
                        const int k0 = u * kSubN;
                        const bool need_mask =
                            (k0 + kSubN > p.Nkv) || (p.causal && k0 + kSubN - 1 > q_start + p.shift);
                        // a rescale at sub-tile u waits for PV(u-1): parity 1-h
                        if (need_mask)
                            softmax_subtile<D, true>(p, tS + h * kSubN, tO, my_p_full + 8 * h, my_o_full + 8 * (1 - h),
                                                     lim - k0, u > 0, pv_count[1 - h], m_ref, l_run);
                        else
                            softmax_subtile<D, false>(p, tS + h * kSubN, tO, my_p_full + 8 * h, my_o_full + 8 * (1 - h),
                                                      kSubN, u > 0, pv_count[1 - h], m_ref, l_run);
                        ++pv_count[h];
                    }
                }
            }

            // ---- epilogue: O_t / l -> fp16 -> global (or the partial-state format) ----
            // the last PV of each parity (sub-tiles n_t-1 and n_t-2) must have retired
            if (n_t > 1) {
                const int h2 = n_t & 1;              // parity of sub-tile n_t - 2
                mbar_wait(my_o_full + 8 * h2, (pv_count[h2] - 1u) & 1u, 30 + t);
            }
            if (n_t > 0) {
                const int h1 = (n_t - 1) & 1;
                mbar_wait(my_o_full + 8 * h1, (pv_count[h1] - 1u) & 1u, 32 + t);
                tc_fence_after();
            }
            const bool row_ok = row < p.Nq;
            const size_t grow = (size_t)wi.bh * p.Nq + row;
            if (!p.partial_mode) {
                const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;   // FA.cu:502-503
                __half* orow = p.o + grow * D;
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
                    if (row_ok) {
#pragma unroll
                        for (int i = 0; i < 32; i += 8) {
                            uint4 v;
                            __half2 h;
                            h = __floats2half2_rn(__uint_as_float(o[i + 0]) * inv, __uint_as_float(o[i + 1]) * inv);
                            v.x = *reinterpret_cast<uint32_t*>(&h);
                            h = __floats2half2_rn(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
                            v.y = *reinterpret_cast<uint32_t*>(&h);
                            h = __floats2half2_rn(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
                            v.z = *reinterpret_cast<uint32_t*>(&h);
                            h = __floats2half2_rn(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
                            v.w = *reinterpret_cast<uint32_t*>(&h);
                            *reinterpret_cast<uint4*>(orow + c + i) = v;
                        }
                    }
                }
            } else {
                // partial state (FA.cu:460-496): un-normalised fp32 O, (m, l) with m in the
                // scaled-score (natural-log) domain; merge algebra of FA.cu:575-597 when accumulating
                float m_out = (m_ref == -INFINITY) ? -FLT_MAX : m_ref * p.scale;
                float l_out = l_run;
                float w_new = 1.f, w_old = 0.f;
                float* prow = p.o_partial + grow * D;
                if (p.accumulate && row_ok) {
                    const float m_old = p.ml[grow * 2 + 0];
                    const float l_old = p.ml[grow * 2 + 1];
                    const float m_max = fmaxf(m_old, m_out);
                    const float kLog2e = 1.4426950408889634f;
                    w_old = (m_old <= -FLT_MAX) ? 0.f : ex2_approx((m_old - m_max) * kLog2e);
                    w_new = (m_out <= -FLT_MAX) ? 0.f : ex2_approx((m_out - m_max) * kLog2e);
                    l_out = l_old * w_old + l_run * w_new;
                    m_out = m_max;
                }
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
                    if (row_ok) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            float4 v = make_float4(__uint_as_float(o[i]) * w_new, __uint_as_float(o[i + 1]) * w_new,
                                                   __uint_as_float(o[i + 2]) * w_new, __uint_as_float(o[i + 3]) * w_new);
                            if (p.accumulate) {
                                const float4 old = *reinterpret_cast<const float4*>(prow + c + i);
                                v.x += old.x * w_old; v.y += old.y * w_old;
                                v.z += old.z * w_old; v.w += old.w * w_old;
                            }
                            *reinterpret_cast<float4*>(prow + c + i) = v;
                        }
                    }
                }
                if (row_ok) {
                    p.ml[grow * 2 + 0] = m_out;
                    p.ml[grow * 2 + 1] = l_out;
                }
            }
            // O_t / S_t are free again: the next item's first P arrival orders after these reads
            tc_fence_before();
        }
    }

    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
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

}  // namespace fa
