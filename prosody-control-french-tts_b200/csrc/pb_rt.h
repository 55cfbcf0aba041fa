// pb_rt.h — the only place that knows whether we are compiled by nvcc (the product) or by g++ against the
// SIMT emulator (tests/simt_emu, test infrastructure).  Kernel sources use plain CUDA constructs plus the
// handful of macros below.
#pragma once
#include <cstddef>
#include <cstdint>

#ifdef PB_SIMT_EMU
#include "pb_simt_emu.h"
typedef void* pbStream_t;
typedef struct { double t; } pbEvent_t;
#define PB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    pb_emu::launch((grid), (block), (smem), [&] { kernel(__VA_ARGS__); })
#define PB_DYN_SMEM(name) unsigned char* name = (unsigned char*)pb_emu::g_blk->dyn_smem
#define PB_GROUP_SYNC(id, nthreads) pb_emu::named_barrier((id), (nthreads))
#define PB_UNROLL
#define PB_NOINLINE __attribute__((noinline))
#else
#define PB_NOINLINE __noinline__
#include <cuda_runtime.h>
typedef cudaStream_t pbStream_t;
typedef cudaEvent_t pbEvent_t;
#define PB_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define PB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define PB_GROUP_SYNC(id, nthreads) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory")
#define PB_UNROLL _Pragma("unroll")
#endif

#define PB_FULL_MASK 0xffffffffu

// ---- NVTX ranges (header-only NVTX3: no library to link; a no-op unless a profiler is attached).  The reference has no tracing
// at all; these mark the host phases and kernel groups of a batch call on an Nsight timeline.
#ifdef PB_SIMT_EMU
struct PbRange { explicit PbRange(const char*) {} };
#else
#include <nvtx3/nvToolsExt.h>
struct PbRange {
    explicit PbRange(const char* name) { nvtxRangePushA(name); }
    ~PbRange() { nvtxRangePop(); }
    PbRange(const PbRange&) = delete; PbRange& operator=(const PbRange&) = delete;
};
#endif

// ---- thin runtime wrappers (return 0 on success; the message of a failure is fetched with pbrt_error())
#ifdef PB_SIMT_EMU
#include <chrono>
static inline int pbrt_device_count() { return 1; }
static inline int pbrt_set_device(int) { return 0; }
static inline int pbrt_malloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 1; }
static inline int pbrt_free(void* p) { free(p); return 0; }
static inline int pbrt_malloc_host(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 1; }
static inline int pbrt_free_host(void* p) { free(p); return 0; }
static inline const void* pbrt_host_device_ptr(const void* p) { return p; }
static inline int pbrt_h2d(void* d, const void* s, size_t n, pbStream_t) { memcpy(d, s, n); return 0; }
static inline int pbrt_d2h(void* d, const void* s, size_t n, pbStream_t) { memcpy(d, s, n); return 0; }
static inline int pbrt_memset(void* d, int v, size_t n, pbStream_t) { memset(d, v, n); return 0; }
static inline int pbrt_stream_create(pbStream_t* s) { *s = nullptr; return 0; }
static inline int pbrt_stream_destroy(pbStream_t) { return 0; }
static inline int pbrt_stream_sync(pbStream_t) { return 0; }
static inline int pbrt_stream_wait_event(pbStream_t, pbEvent_t) { return 0; }
static inline int pbrt_event_create(pbEvent_t* e) { e->t = 0; return 0; }
static inline int pbrt_event_destroy(pbEvent_t) { return 0; }
static inline int pbrt_event_record(pbEvent_t* e, pbStream_t) {
    e->t = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    return 0;
}
static inline float pbrt_event_ms(pbEvent_t a, pbEvent_t b) { return (float)(b.t - a.t); }
static inline int pbrt_last_error() { return 0; }
static inline const char* pbrt_error() { return "emulator"; }
static inline int pbrt_props(int, int* sms, int* maj, int* min_, long long* mem) { *sms = 2; *maj = 0; *min_ = 0; *mem = 0; return 0; }
#else
static inline int pbrt_device_count() { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
static inline int pbrt_set_device(int d) { return cudaSetDevice(d) != cudaSuccess; }
static inline int pbrt_malloc(void** p, size_t n) { return cudaMalloc(p, n ? n : 1) != cudaSuccess; }
static inline int pbrt_free(void* p) { return cudaFree(p) != cudaSuccess; }
static inline int pbrt_malloc_host(void** p, size_t n) { return cudaMallocHost(p, n ? n : 1) != cudaSuccess; }
static inline int pbrt_free_host(void* p) { return cudaFreeHost(p) != cudaSuccess; }
// device-side address of pinned host memory (identical under unified addressing)
static inline const void* pbrt_host_device_ptr(const void* p) {
    void* d = nullptr;
    if (cudaHostGetDevicePointer(&d, const_cast<void*>(p), 0) != cudaSuccess) { cudaGetLastError(); return p; }
    return d;
}
static inline int pbrt_h2d(void* d, const void* s, size_t n, pbStream_t st) { return cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, st) != cudaSuccess; }
static inline int pbrt_d2h(void* d, const void* s, size_t n, pbStream_t st) { return cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, st) != cudaSuccess; }
static inline int pbrt_memset(void* d, int v, size_t n, pbStream_t st) { return cudaMemsetAsync(d, v, n, st) != cudaSuccess; }
static inline int pbrt_stream_create(pbStream_t* s) { return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking) != cudaSuccess; }
static inline int pbrt_stream_destroy(pbStream_t s) { return cudaStreamDestroy(s) != cudaSuccess; }
static inline int pbrt_stream_sync(pbStream_t s) { return cudaStreamSynchronize(s) != cudaSuccess; }
static inline int pbrt_stream_wait_event(pbStream_t s, pbEvent_t e) { return cudaStreamWaitEvent(s, e, 0) != cudaSuccess; }
static inline int pbrt_event_create(pbEvent_t* e) { return cudaEventCreate(e) != cudaSuccess; }
static inline int pbrt_event_destroy(pbEvent_t e) { return cudaEventDestroy(e) != cudaSuccess; }
static inline int pbrt_event_record(pbEvent_t* e, pbStream_t s) { return cudaEventRecord(*e, s) != cudaSuccess; }
static inline float pbrt_event_ms(pbEvent_t a, pbEvent_t b) { float ms = 0.f; cudaEventElapsedTime(&ms, a, b); return ms; }
static inline int pbrt_last_error() { return cudaPeekAtLastError() != cudaSuccess; }
static inline const char* pbrt_error() { return cudaGetErrorString(cudaGetLastError()); }
static inline int pbrt_props(int dev, int* sms, int* maj, int* min_, long long* mem) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 1;
    *sms = p.multiProcessorCount; *maj = p.major; *min_ = p.minor; *mem = (long long)p.totalGlobalMem; return 0;
}
#endif
