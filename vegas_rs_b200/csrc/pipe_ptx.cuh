// pipe_ptx.cuh -- PTX wrappers shared by the persistent, phase-pipelined kernels (heis_pipe.cu, basis_pipe.cu): mbarrier,
// TMA (cp.async.bulk[.tensor]), proxy fences, scoped loads, bounded waits.  sm_100a.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vg {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity), "r"(2000u)   // suspend-time hint (ns): sleep in hardware instead of re-polling
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, uint32_t c0, uint32_t c1, uint32_t c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// The same with an L2 eviction-priority hint (createpolicy): a phase-pipelined kernel knows which tiles the next phase is
// about to read again (keep: evict_last) and which it touches for the last time in this step (evict_first).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_3d_hint(void* dst, const CUtensorMap* map, uint64_t* bar, uint32_t c0, uint32_t c1, uint32_t c2, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap* map, const void* src, uint32_t c0, uint32_t c1, uint32_t c2, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
                 ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
                 : "memory");
}

// TMA store of a box from shared memory (bulk async-group completion)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, uint32_t c0, uint32_t c1, uint32_t c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr unsigned long long PIPE_TIMEOUT_NS = 4000000000ull;   // a wait that long means a broken dependency: give up, never hang

enum PipeError : unsigned int { PIPE_ERR_FULL = 1, PIPE_ERR_EMPTY = 2, PIPE_ERR_GATE = 3, PIPE_ERR_PEER = 4 };

// Waits for phase `parity` of an mbarrier.  False when the launch is being abandoned (abort flag) or on time-out.
// backoff_ns: the hardware suspend of try_wait is ended by EVERY mbarrier event of the CTA, so a waiting warp re-polls about
// every 50 ns and the polls of a phase-pipelined kernel add up to a quarter of its issued instructions
// (profiles/r02/README.md section 8); an unconditional sleep between polls trades wake-up latency for issue slots.
__device__ __forceinline__ bool wait_bar(uint64_t* b, uint32_t parity, volatile uint32_t* abort_flag, unsigned int* gerr, unsigned int code,
                                         uint32_t backoff_ns = 0u) {
    if (mbar_try_wait(b, parity)) return true;
    const unsigned long long t0 = global_timer();
    uint32_t n = 0;
    while (!mbar_try_wait(b, parity)) {
        if (backoff_ns) __nanosleep(backoff_ns);
        if ((++n & 63u) == 0) {
            if (*abort_flag) return false;
            if (global_timer() - t0 > PIPE_TIMEOUT_NS) { *abort_flag = 1u; atomicExch(gerr, code); return false; }
        }
    }
    return true;
}

// Spins until *p >= target (SYS: the word is written by another GPU); `seen` receives the last value read.
template <bool SYS>
__device__ __forceinline__ bool wait_counter(const unsigned long long* p, unsigned long long target, unsigned long long& seen,
                                             volatile uint32_t* abort_flag, unsigned int* gerr, unsigned int code) {
    seen = SYS ? ld_acquire_sys(p) : ld_acquire_gpu(p);
    if (seen >= target) return true;
    const unsigned long long t0 = global_timer();
    uint32_t n = 0;
    while ((seen = (SYS ? ld_acquire_sys(p) : ld_acquire_gpu(p))) < target) {
        __nanosleep(200);     // the helper warps outrank the consumers in the issue arbiter: do not burn their slots
        if ((++n & 63u) == 0) {
            if (*abort_flag) return false;
            if (global_timer() - t0 > PIPE_TIMEOUT_NS) { *abort_flag = 1u; atomicExch(gerr, code); return false; }
        }
    }
    return true;
}

// A position in a ring of `n` mbarrier-guarded slots: slot index and the parity of its current use.
struct RingPos {
    uint32_t slot = 0, parity = 0;
    __device__ __forceinline__ void advance(uint32_t n) { if (++slot == n) { slot = 0; parity ^= 1u; } }
};


// 1-D bulk copies (contiguous bytes; 16-byte aligned, size a multiple of 16)
__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}

}  // namespace
}  // namespace vg
