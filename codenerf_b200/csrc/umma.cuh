// umma.cuh -- hand-written sm_100a primitives: mbarrier, bulk async copy (TMA engine,
// 1-D form), tcgen05 tensor-core MMA with TMEM accumulators, descriptors.
// Inline PTX only; compile with -gencode arch=compute_100a,code=sm_100a.
#pragma once
#include <cstdio>
#include <cuda_bf16.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// CNB_MBAR_HINT_NS > 0: suspend-time hint of the potentially blocking try_wait (the thread still resumes as soon as the
// phase completes; a longer limit only means fewer wake-ups of a waiting warp, i.e. fewer spin-loop instructions).
#ifndef CNB_MBAR_HINT_NS
#define CNB_MBAR_HINT_NS 0
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
#if CNB_MBAR_HINT_NS
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "n"(CNB_MBAR_HINT_NS)
        : "memory");
#else
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
#endif
    return ok != 0;
}
// Bounded spin: a protocol bug must not hang the GPU box, and must not be silent either.  A wait that lasts longer
// than the watchdog (~4 s of SM clocks, three orders of magnitude above the longest legitimate wait, so time-slicing
// or a sanitizer does not trip it) records a code and TRAPS: the launch fails, every later CUDA call on the context
// returns an error, and the host wrappers (which check the status of every call) raise.  The flag is per translation
// unit (it only serves post-mortem inspection under cuda-gdb); cnb_debug_pipeline_timeouts reports it, or the
// context error, to tests.
static __device__ unsigned int g_umma_timeout = 0;
constexpr long long kWatchdogCycles = 8000000000LL;
static __device__ __noinline__ void mbar_watchdog_trip(unsigned int code) {
    atomicExch(&g_umma_timeout, code);
    __threadfence_system();
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
#ifdef CNB_WATCHDOG_PRINT      // debugging aid: report the wait that timed out and carry on (results are garbage)
        if (clock64() - t0 > 400000000LL) {
            if ((threadIdx.x & 31) == 0) printf("TIMEOUT blk %d warp %d bar+%d parity %u\n", (int)blockIdx.x, (int)(threadIdx.x >> 5), (int)(smem_u32(bar) & 0xfff), parity);
            return;
        }
#else
        if (clock64() - t0 > kWatchdogCycles) mbar_watchdog_trip(1u);
#endif
    }
}

// ---- register reallocation between warpgroups (all 4 warps of a warpgroup must execute the same one) ----
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// ---- proxies / fences -------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads, bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t threads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---- bulk async copies (TMA engine, no tensor map) ---------------------------
// global -> shared, completion on an mbarrier (complete_tx::bytes).  16-B aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// ... with an L2 eviction-priority policy (weights: keep them resident while a streaming store floods L2)
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
// L2 eviction-priority policies (createpolicy) and hinted accesses
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ void st_global_hint(uint32_t* ptr, uint32_t v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(ptr), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_global_v4_hint(void* ptr, uint4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
// (ReLU bit words: 28 KB per tile would sweep the ~27 KB of L1 left beside the shared memory and evict the few lines that
// do get re-used -- z values, target colours, bias rows; CNB_MASK_L1_NOALLOC=0 restores the allocating form)
#ifndef CNB_MASK_L1_NOALLOC
#define CNB_MASK_L1_NOALLOC 1
#endif
__device__ __forceinline__ uint32_t ld_global_hint(const uint32_t* ptr, uint64_t pol) {
    uint32_t v;
#if CNB_MASK_L1_NOALLOC
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(ptr), "l"(pol) : "memory");
#else
    asm volatile("ld.global.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(ptr), "l"(pol) : "memory");
#endif
    return v;
}
// shared -> global bulk store with an L2 policy (streaming data that should not displace the working set)
__device__ __forceinline__ void bulk_s2g_hint(void* gmem_dst, const void* smem_src, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's committed bulk stores have finished READING shared memory
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- TMEM -------------------------------------------------------------------
// one full warp; writes the base address of `cols` (power of two >= 32) columns to *smem_out
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_out, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base_lane + i).
// A warp may only touch TMEM lanes [32*(warp_id%4), +32).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));      // volatile (ordered against the barrier waits / fences), but no memory clobber
}
// 32 lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- A operand in tensor memory ("TS" MMAs) ------------------------------------------------------
// TMEM image of a K-major bf16 A tile [128 x K]: row m in lane m, elements (2c, 2c+1) of the row packed in the
// 32-bit column c (tests/probe_ts.cu checks this against the shared-memory form bit for bit).
// 32 lanes x 16 consecutive 32-bit columns from registers; same lane-quarter rule as tcgen05.ld.
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// ---- CTA pairs (cta_group::2): cluster helpers ------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank)); return r;
}
// Remote arrive on an mbarrier of another CTA of the cluster (address from mapa).  Default semantics (release at CTA
// scope), as CUTLASS's ClusterBarrier::arrive does: an explicit `.release.cluster` compiles to MEMBAR.ALL.GPU +
// CCTL.IVALL -- a GPU-scope fence that waits for every outstanding global store of the thread (the ReLU mask words,
// the stash) and an invalidation of the SM's whole L1 -- on EVERY operand hand-off; that, not shared-memory
// bandwidth, is what made the CTA-pair kernels' epilogues slow in round 1.  What the hand-off needs is already
// there: the operand stores are fenced into the async proxy (fence.proxy.async) and read by this CTA's own tensor
// core; the arrive only tells the leader that it may issue.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
#if CNB_MBAR_HINT_NS
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "n"(CNB_MBAR_HINT_NS) : "memory");
#else
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"      // default acquire.cta: `.acquire.cluster` costs a CCTL.IVALL per wait
        "selp.b32 %0, 1, 0, P;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > kWatchdogCycles) mbar_watchdog_trip(3u);
    }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_out, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem, both CTAs] (+)= A . B^T with M = 256 split over the CTA pair; issued by ONE thread of the leader CTA.
__device__ __forceinline__ void mma_bf16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in every CTA of `mask` once all prior MMAs completed
__device__ __forceinline__ void mma_commit_2cta(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// ---- weight multicast inside a cluster (independent cta_group::1 MMAs per CTA) -----------------
// One bulk copy lands at the same shared-memory offset of every CTA in `mask` and performs complete_tx on
// the mbarrier at the same offset in each of them.
__device__ __forceinline__ void bulk_g2s_mcast(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mcast_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                                    uint16_t mask, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint [%0], [%1], %2, [%3], %4, %5;"
        ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask), "l"(pol) : "memory");
}
// cta_group::1 commit that arrives on the same-offset mbarrier of every CTA in `mask`
__device__ __forceinline__ void mma_commit_mcast(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// Tensor-map TMA load of one 2-D box into THIS CTA's shared memory, completing on an mbarrier of the pair's
// leader CTA (`leader_bar` = mapa(bar, 0)): the cta_group::2 form that feeds 2-CTA MMAs without a forwarding hop.
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, int c0, int c1, uint32_t leader_bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}

// ... with an L2 eviction-priority policy (the weights must stay resident while the stash streams through L2)
__device__ __forceinline__ void tma_load_2d_pair_hint(void* smem_dst, const void* tmap, int c0, int c1, uint32_t leader_bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(leader_bar), "r"(c0), "r"(c1), "l"(pol) : "memory");
}

// ---- descriptors --------------------------------------------------------------
// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32.
//   bits [4,6) D format (1 = f32), [7,10) A format (1 = bf16), [10,13) B format (1 = bf16),
//   bit 15 A major (0 = K, 1 = MN), bit 16 B major, [17,23) N >> 3, [24,29) M >> 4.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Shared-memory matrix descriptor (sm_100: version field = 1).
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte offset >> 4,
//   [46,48) version = 1, [61,64) swizzle: 0 none, 2 = 128B, 4 = 64B, 6 = 32B.
// K-major operand (rows x K, K contiguous), swizzled: rows of one swizzle span; 8-row groups
//   SBO bytes apart; LBO unused.  Advancing K inside the span = adding bytes to the start address.
// MN-major operand (MN contiguous): 64-element (128-B) MN spans LBO bytes apart, 8-row K groups SBO apart.
enum : uint32_t { SWZ_NONE = 0, SWZ_128B = 2, SWZ_64B = 4, SWZ_32B = 6 };
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                               uint32_t swizzle) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)swizzle << 61;
    return d;
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives once every previously issued MMA of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---- swizzled operand addressing ----------------------------------------------
// Byte offset of element (row, col) inside a [rows x 64] bf16 block stored with the 128-B swizzle
// (row pitch 128 B, 16-B chunk index XOR (row & 7)).  The same bytes are a K-major operand
// (rows = M/N, cols = K) and an MN-major operand (rows = K, cols = M/N).
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t col) {
    return row * 128u + ((((col >> 3) ^ row) & 7u) << 4) + ((col & 7u) << 1);
}
// [rows x 32] bf16 block with the 64-B swizzle (row pitch 64 B, 16-B chunk index XOR ((row >> 1) & 3)).
__host__ __device__ __forceinline__ uint32_t sw64_offset(uint32_t row, uint32_t col) {
    return row * 64u + ((((col >> 3) ^ (row >> 1)) & 3u) << 4) + ((col & 7u) << 1);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace umma
