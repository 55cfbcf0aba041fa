// pb_async.cuh — sm_100 asynchronous data movement used by the F0 kernels: 1-D bulk copies global -> shared memory
// issued by ONE thread to the TMA unit (cp.async.bulk, SASS UBLKCP) and completed on an mbarrier, so the SM's
// load/store pipe and the warps' issue slots stay free for arithmetic while the next work item's data lands.
// Under the SIMT emulator (tests only) the copy is a synchronous memcpy and the wait is the caller's barrier.
#pragma once
#include "pb_rt.h"

typedef unsigned long long pbMbar;      // one 64-bit mbarrier object in shared memory

#ifdef PB_SIMT_EMU
#include <cstring>
static inline void pb_mbar_init(pbMbar* bar, int) { *bar = 0; }
static inline void pb_mbar_init_fence() {}
static inline void pb_mbar_expect_tx(pbMbar*, unsigned) {}
static inline void pb_bulk_g2s(void* dst, const void* src, unsigned bytes, pbMbar*) { memcpy(dst, src, bytes); }
static inline void pb_mbar_wait(pbMbar*, unsigned) {}
#else
__device__ __forceinline__ unsigned pb_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pb_mbar_init(pbMbar* bar, int arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pb_smem_u32(bar)), "r"(arrivals) : "memory");
}
// makes the initialised barriers visible to the async proxy (the TMA unit completes transactions on them)
__device__ __forceinline__ void pb_mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// one arrival that also announces `bytes` of asynchronous writes the current phase has to wait for
__device__ __forceinline__ void pb_mbar_expect_tx(pbMbar* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pb_smem_u32(bar)), "r"(bytes) : "memory");
}
// bytes: multiple of 16; dst and src 16-byte aligned
__device__ __forceinline__ void pb_bulk_g2s(void* dst, const void* src, unsigned bytes, pbMbar* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(pb_smem_u32(dst)), "l"(src), "r"(bytes), "r"(pb_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pb_mbar_wait(pbMbar* bar, unsigned parity) {
    const unsigned a = pb_smem_u32(bar);
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
#endif
