// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05/TMEM.
// Hand-written for this kernel set; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace reve {

// Diagnostic word written (to mapped pinned host memory) before a watchdog trap.
struct DebugBlock {
    unsigned int code;      // tag of the wait that timed out (0 = none)
    unsigned int block;     // blockIdx.x
    unsigned int aux0;
    unsigned int aux1;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
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
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

// Non-blocking poll (try_wait may suspend the thread for a system-dependent time before returning false).
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

// Bounded wait: a protocol bug must end as a trapped launch with a diagnostic, never as a hung
// GPU.  ~2^32 cycles (> 2 s at any clock) is far beyond any legitimate wait in these kernels.
static __device__ __noinline__ void watchdog_fail(DebugBlock* dbg, uint32_t tag, uint32_t a0, uint32_t a1) {
    if (dbg) {
        dbg->block = blockIdx.x;
        dbg->aux0 = a0;
        dbg->aux1 = a1;
        dbg->code = tag;
        __threadfence_system();
    }
    asm volatile("trap;");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, DebugBlock* dbg,
                                          uint32_t tag, uint32_t aux = 0) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > (1ll << 32)) watchdog_fail(dbg, tag, aux, parity);
    }
}
// Same, for warps whose wait is long and off the critical path in a kernel that is bound by instruction issue: every
// failed try_wait is followed by a nanosleep, so the waiting warp stops competing for issue slots with the working ones.
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, DebugBlock* dbg,
                                                  uint32_t tag, uint32_t aux = 0) {
    if constexpr (SLEEP_NS <= 0) {
        mbar_wait(bar, parity, dbg, tag, aux);
    } else {
        if (mbar_try_wait(bar, parity)) return;
        const long long t0 = clock64();
        while (!mbar_try_wait(bar, parity)) {
            asm volatile("nanosleep.u32 %0;" ::"n"(SLEEP_NS));
            if (clock64() - t0 > (1ll << 32)) watchdog_fail(dbg, tag, aux, parity);
        }
    }
}

// ---------------------------------------------------------------- async proxy / TMA
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// global (tensor map, 3-D coords innermost first) -> shared, completes on an mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// same, with an L2 cache-policy operand (createpolicy-style 64-bit constant)
__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                                 int c0, int c1, int c2, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}
constexpr uint64_t kPolicyEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kPolicyEvictLast = 0x14F0000000000000ull;
constexpr uint64_t kPolicyEvictNormal = 0x1000000000000000ull;
// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
// 32-byte (one full sector) global store
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void st_global_v8_hint(void* p, const uint32_t (&v)[8], uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8}, %9;" ::"l"(p), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// shared -> global (tensor map), bulk-group completion
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1,
                                             int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
        ::"l"(m), "r"(src), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m),
                 "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// plain 1-D bulk copy global -> shared (16-byte aligned, size multiple of 16)
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes,
                                             uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

// generic <-> async proxy ordering for global memory (TMA stores/loads vs flag loads/stores)
__device__ __forceinline__ void fence_proxy_async_global() {
    asm volatile("fence.proxy.async.global;" ::: "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void ld_acquire_gpu_v2(const unsigned int* p, int& a, int& b) {
    asm volatile("ld.acquire.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_relaxed_gpu(unsigned int* p, unsigned int v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Bounded wait for a monotonic counter in global memory to reach `want` (`cache` = last value seen).  Returns
// whether memory had to be read.
__device__ __forceinline__ bool flag_wait_ge(const unsigned int* p, int want, int& cache, DebugBlock* dbg, uint32_t tag) {
    if (cache >= want) return false;
    const long long t0 = clock64();
    for (;;) {
        cache = static_cast<int>(ld_acquire_gpu(p));
        if (cache >= want) return true;
        if (clock64() - t0 > (1ll << 32)) watchdog_fail(dbg, tag, static_cast<uint32_t>(want), static_cast<uint32_t>(cache));
    }
}

// ---------------------------------------------------------------- tcgen05 / TMEM
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
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 (fp16/bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same with an A-operand collector hint: consecutive MMAs that read the same A tile keep it in the tensor
// core's collector buffer instead of re-reading shared memory.  FILL = read A and keep it, USE = reuse and
// keep, LASTUSE = reuse then drop (SASS: A_KEEP / A_REUSE).
enum class ACollector { FILL, USE, LASTUSE };
template <ACollector C>
__device__ __forceinline__ void umma_f16_a(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
    if constexpr (C == ACollector::FILL)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                     "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    else if constexpr (C == ACollector::USE)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::use [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                     "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                     "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// all previously issued tcgen05 async ops of this thread -> one arrival on the mbarrier
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp receives lane (base_lane+i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// ---------------------------------------------------------------- TMEM stores / CTA pairs (cta_group::2)
// 32 lanes x 16 consecutive 32-bit columns, all set to the same value
__device__ __forceinline__ void tmem_st16_fill(uint32_t taddr, uint32_t v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
        ::"r"(taddr), "r"(v)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default .release.cta semantics: a cluster-scope release here costs ~1000 cycles per arrive (measured), and
    // the TMEM hand-over is ordered by tcgen05.fence::before_thread_sync, not by this arrive
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// In a CTA pair the shared-window address of the odd CTA differs from the even one's in this bit; clearing it
// makes a TMA load of either CTA report its bytes to the even (leader) CTA's mbarrier.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_3d_hint_pair(uint32_t dst, const CUtensorMap* m, uint32_t leader_bar,
                                                      int c0, int c1, int c2, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(m), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// M = 256 across the CTA pair: each CTA supplies its own 128-row A tile and half of B's N rows, and receives
// its 128 lanes x N columns of D.  Issued by one thread of the even CTA.
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued tcgen05 async ops -> one arrival on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask)
                 : "memory");
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle (one 64 x fp16 row = 128 B,
// 8-row groups 1024 B apart).  `addr` is the shared-space byte address of row 0 / k-offset;
// `base_offset` is the 128-byte-row phase of a start address that is not 1024-aligned.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t addr, uint32_t base_offset) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);            // [0,14)  start address >> 4
    d |= static_cast<uint64_t>(1) << 16;                          // [16,30) LBO (unused for SW128 K-major)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                  // [32,46) SBO = 1024 B
    d |= static_cast<uint64_t>(1) << 46;                          // [46,48) descriptor version (sm_100)
    d |= static_cast<uint64_t>(base_offset & 7) << 49;            // [49,52) matrix base offset
    d |= static_cast<uint64_t>(2) << 61;                          // [61,64) SWIZZLE_128B
    return d;
}
// Instruction descriptor for kind::f16: fp16 A/B (K-major both), fp32 D, M x N tile.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n) {
    return (1u << 4)                              // D format: F32
           | (0u << 7) | (0u << 10)               // A, B format: F16
           | (0u << 15) | (0u << 16)              // A, B K-major
           | (static_cast<uint32_t>(n >> 3) << 17)
           | (static_cast<uint32_t>(m >> 4) << 24);
}

}  // namespace reve
