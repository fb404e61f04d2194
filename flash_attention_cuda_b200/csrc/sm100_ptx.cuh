// sm100_ptx.cuh -- thin inline-PTX layer for Blackwell (sm_100a): mbarrier, TMA, tcgen05
// (TMEM alloc / ld / st / mma / commit) and UMMA descriptor encoders.
//
// Everything here is a 1:1 wrapper over one PTX instruction, so the kernels read as the
// hardware protocol they implement.  Compile with -gencode arch=compute_100a,code=sm_100a.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace sm100 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() {
    uint32_t l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

// One lane of a converged warp (elect.sync).  Code under `if (elect_one())` keeps warp-uniform
// operands in uniform registers, which is what UTCHMMA / UTMALDG / UTCBAR consume.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
template <int kRegs>
__device__ __forceinline__ void setmaxnreg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs));
}
template <int kRegs>
__device__ __forceinline__ void setmaxnreg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Blocking wait with a watchdog.  A protocol bug must not hang the GPU: after kWatchdogNs of wall time (no
// legitimate wait here exceeds a few ms; the margin covers time-slicing and debuggers) the waiter records who it
// was in a device word that also releases every other wait in its slow path, and the kernel drains and exits with
// garbage results instead of spinning.  On its way out thread 0 of every CTA mirrors the record into zero-copy host
// memory (g_watchdog_host, installed by the launcher), where the launcher finds it at the start of its next call
// without synchronising, reports it once as FA_ERR_WATCHDOG and clears both copies (fa_api.cu: take_watchdog).  No
// function call / printf / trap here on purpose: a call in the kernel makes ptxas ignore the per-role setmaxnreg
// budgets and spill the softmax warps.
__device__ unsigned int g_watchdog[4];         // [0]: 0, or the packed record of the first waiter that gave up (below)
__device__ unsigned int* g_watchdog_host;      // device alias of the launcher's pinned mirror {abort flag, barrier tag, block, thread}, or null
constexpr unsigned long long kWatchdogNs = 10ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ bool watchdog_aborted() {
    return *reinterpret_cast<volatile unsigned int*>(&g_watchdog[0]) != 0u;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// The slow path of every wait is inlined at ~40 sites, most of them inside the kernel's hot loops, whose code has to
// stay small (L0 instruction cache: ~6 KB per sub-partition, 32 KB per SM behind it): raising the record is ONE
// compare-and-swap of a packed word {1, tag:7, block:14, thread:10}; thread 0 of every CTA copies it to the host mirror
// when the kernel drains (watchdog_publish) -- the CTA of the waiter that gave up does; the others never read the word.
// block-local "a waiter of this CTA gave up": lets the CTA's exit path skip the global read of the record (a ~1 us round trip
// that would otherwise sit on the tail of every CTA of every launch)
// (The 16 bytes of static shared memory shift the kernel's dynamic window off its 1024-byte alignment; fa_fwd_kernel's 1 KB of
// alignment slack absorbs exactly that -- with 16 bytes to spare at D = 128 -- and its entry check refuses to run otherwise.
// Keeping the word at the end of the dynamic window instead measured the same: profiles/r02_c32_ab_merge2.log.)
__device__ __forceinline__ unsigned int* watchdog_block_flag() {
    __shared__ unsigned int raised;
    return &raised;
}
__device__ __forceinline__ void watchdog_raise(int tag) {
    *reinterpret_cast<volatile unsigned int*>(watchdog_block_flag()) = 1u;
    atomicCAS(&g_watchdog[0], 0u,
              0x80000000u | ((unsigned)tag & 0x7fu) << 24 | (blockIdx.x & 0x3fffu) << 10 | (threadIdx.x & 0x3ffu));
}
__device__ __forceinline__ void watchdog_publish() {
    const unsigned int rec = *reinterpret_cast<volatile unsigned int*>(&g_watchdog[0]);
    if (rec == 0u) return;
    volatile unsigned int* h = g_watchdog_host;
    if (h && h[0] == 0u) {
        h[1] = (rec >> 24) & 0x7fu;
        h[2] = (rec >> 10) & 0x3fffu;
        h[3] = rec & 0x3ffu;
        __threadfence_system();
        h[0] = 1u;
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = global_timer_ns();
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++polls & 63u) == 0u) {
            if (watchdog_aborted()) return;
            if (global_timer_ns() - t0 > kWatchdogNs) {
                watchdog_raise(tag);
                return;
            }
        }
    }
}

// mbar_wait for a warp that has a long time to wait and shares its sub-partition with latency-critical warps: sleeps between
// polls instead of re-issuing try_wait back to back (epilogue warpgroup experiment, -DFA_EPI_WG=1)
__device__ __forceinline__ void mbar_wait_sleepy(uint32_t bar, uint32_t parity, int tag, unsigned ns) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = global_timer_ns();
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(ns);
        if ((++polls & 63u) == 0u) {
            if (watchdog_aborted()) return;
            if (global_timer_ns() - t0 > kWatchdogNs) {
                watchdog_raise(tag);
                return;
            }
        }
    }
}

// ------------------------------------------------------------------ clusters / CTA pairs
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_arrive_release() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on an mbarrier that lives in another CTA of the cluster (address from mapa_shared).
// Default semantics (release at CTA scope): what is handed over lives in tensor memory and is
// ordered by tcgen05.fence::before/after_thread_sync, so no cluster-wide release of generic
// memory is needed -- the .release.cluster form below stalls the arriving warp for ~1000 cycles.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// same, publishing this thread's earlier st.shared::cluster writes to the waiter (scheduler mailbox)
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void st_shared_cluster_u32(uint32_t cluster_addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// mbar_wait for barriers that also receive arrivals from the peer CTA (acquire at cluster scope)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int tag) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const unsigned long long t0 = global_timer_ns();
    uint32_t polls = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if ((++polls & 63u) == 0u) {
            if (watchdog_aborted()) return;
            if (global_timer_ns() - t0 > kWatchdogNs) {
                watchdog_raise(tag);
                return;
            }
        }
    }
}
// a flag another engine of the system (copy engine, stream memory op) sets while the kernel runs
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// named barrier among `count` threads of the CTA (id 0 is __syncthreads)
__device__ __forceinline__ void bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// producer side of a named barrier: counts this warp in without waiting (pairs with bar_sync on the consumer side)
__device__ __forceinline__ void bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void ld_shared_v2f(uint32_t addr, float& a, float& b) {
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(addr) : "memory");
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_shared_v4f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ------------------------------------------------------------------ proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ------------------------------------------------------------------ TMA (cp.async.bulk.tensor)
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const void* tmap, uint32_t bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
// CTA-pair form: the data lands in this CTA's shared memory, the transaction bytes are counted on
// an mbarrier of the pair's leader CTA (`bar` is a shared::cluster address)
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst_smem, const void* tmap, uint32_t bar,
                                                int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst_smem, const void* tmap, uint32_t bar,
                                                 int c0, int c1, int c2, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(bar),
        "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t src_smem, int c0, int c1,
                                             int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
        ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int kPending>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
template <int kPending>
__device__ __forceinline__ void tma_store_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(kPending) : "memory");
}

// ------------------------------------------------------------------ TMEM allocation
// One full warp executes alloc/dealloc (.sync.aligned).  The base address is written to smem.
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}

// CTA-pair allocation: the same warp of both CTAs executes these with the same smem_dst offset
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}

// ------------------------------------------------------------------ tcgen05.mma / commit
// D[tmem] (+)= A[smem desc] * B[smem desc]; one thread issues for the CTA.
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                        uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc,
                                        uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive (count 1) on an mbarrier once every tcgen05.mma issued so far by this thread has
// completed.  Implies tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}

// CTA-pair MMA (M = 256 over two SMs): issued by one thread of the leader CTA; descriptors and
// TMEM addresses are CTA-relative and are applied in both CTAs, each of which supplies its own
// 128 rows of A and half of B's N extent.
__device__ __forceinline__ void umma_ss_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
__device__ __forceinline__ void umma_ts_2sm(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
// commit of a CTA-pair MMA stream: arrives on the mbarrier at this smem offset in both CTAs
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
    const uint16_t mask = 3;
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"(mask)
        : "memory");
}

// ------------------------------------------------------------------ tcgen05.ld / st (32 lanes x 32b, xN columns)
// Thread `lane` of warp w touches TMEM lane 32*(w%4)+lane and N consecutive 32-bit columns.
#define FA_R4(a, i) a[i], a[i + 1], a[i + 2], a[i + 3]
#define FA_OUT4(a, i) "=r"(a[i]), "=r"(a[i + 1]), "=r"(a[i + 2]), "=r"(a[i + 3])
#define FA_IN4(a, i) "r"(a[i]), "r"(a[i + 1]), "r"(a[i + 2]), "r"(a[i + 3])

__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : FA_OUT4(r, 0), FA_OUT4(r, 4), FA_OUT4(r, 8), FA_OUT4(r, 12), FA_OUT4(r, 16),
          FA_OUT4(r, 20), FA_OUT4(r, 24), FA_OUT4(r, 28)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : FA_OUT4(r, 0), FA_OUT4(r, 4), FA_OUT4(r, 8), FA_OUT4(r, 12)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), FA_IN4(r, 0), FA_IN4(r, 4), FA_IN4(r, 8), FA_IN4(r, 12), FA_IN4(r, 16),
        FA_IN4(r, 20), FA_IN4(r, 24), FA_IN4(r, 28)
        : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), FA_IN4(r, 0), FA_IN4(r, 4), FA_IN4(r, 8), FA_IN4(r, 12)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ UMMA descriptors
// Shared-memory matrix descriptor (64 bit), SWIZZLE_128B, Blackwell version field = 1.
//   [0,14)  start address >> 4      [16,30) leading byte offset >> 4
//   [32,46) stride byte offset >> 4 [46,48) version = 1   [61,64) layout type (2 = 128B swizzle)
// K-major operand (rows x 64 halves panels, 128 B per row): SBO = 8 rows * 128 B = 1024,
// LBO unused (1).  MN-major operand (K rows x 64 halves panels): SBO = 1024 between groups of
// 8 K-rows, LBO = byte distance between successive 64-element MN chunks (the panel stride).
__host__ __device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                            uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}
// Instruction descriptor (32 bit) for kind::f16 with fp16 A/B and fp32 accumulate.
//   [4,6) D format (1 = f32)  [7,10) A format (0 = f16)  [10,13) B format (0 = f16)
//   [15] A major (0 = K)      [16] B major (0 = K, 1 = MN)
//   [17,23) N >> 3            [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (static_cast<uint32_t>(a_mn_major) << 15) |
           (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}

// ------------------------------------------------------------------ math
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
#ifdef FA_NO_EXP
    y = x;
#else
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
#endif
    return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {   // FMNMX3
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
// packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2): two lanes per instruction on the FMA pipe
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

}  // namespace sm100
