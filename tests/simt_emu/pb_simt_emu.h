// pb_simt_emu.h — TEST INFRASTRUCTURE ONLY.
//
// A tiny SIMT emulator: lets the *unmodified* kernel sources under
// prosody-control-french-tts_b200/csrc/ be compiled with g++ (-DPB_SIMT_EMU) and executed on the CPU,
// one std::thread per CUDA thread, blocks run one after another.  It exists because the build container
// has no GPU: indexing / synchronisation bugs are caught here before GPU minutes are spent.
//
// It is NOT a CPU fallback.  The product (libprosody_b200.so) is built by nvcc only; the emulated build
// goes to tests/simt_emu/_build/ and is loaded only by tests/ (see tests/simt_emu/build_emu.py).
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

struct uint3_ { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
struct double2 { double x, y; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
// sm_100 packed-pair arithmetic (FADD2 / FMUL2 / FFMA2)
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{__builtin_fmaf(a.x, b.x, c.x), __builtin_fmaf(a.y, b.y, c.y)}; }
static inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }
static inline int2 make_int2(int a, int b) { return int2{a, b}; }

namespace pb_emu {

// A reusable barrier (generation counting) for n participants.
class Barrier {
public:
    explicit Barrier(int n = 1) : n_(n), count_(0), gen_(0) {}
    void reset(int n) { n_ = n; count_ = 0; }
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        unsigned g = gen_;
        if (++count_ == n_) { count_ = 0; ++gen_; cv_.notify_all(); }
        else cv_.wait(lk, [&] { return g != gen_; });
    }
private:
    std::mutex m_;
    std::condition_variable cv_;
    int n_, count_;
    unsigned gen_;
};

struct BlockCtx {
    int nthreads = 0;
    Barrier block_bar;
    std::vector<Barrier*> warp_bar;
    std::vector<Barrier*> named_bar;          // indexed by id, sized lazily by first use
    std::mutex named_m;
    std::vector<uint64_t> xchg;               // per-thread exchange slot (shuffles / ballots)
    char* dyn_smem = nullptr;
};

extern BlockCtx* g_blk;
extern thread_local uint3_ t_threadIdx, t_blockIdx;
extern thread_local dim3 t_blockDim, t_gridDim;

inline int lane() { return (int)(t_threadIdx.x & 31); }
inline int warp() { return (int)(t_threadIdx.x >> 5); }
inline int warp_width() { int rem = g_blk->nthreads - warp() * 32; return rem < 32 ? rem : 32; }
inline void warp_sync() { g_blk->warp_bar[warp()]->wait(); }

template <typename T> inline T shfl(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    uint64_t raw = 0; std::memcpy(&raw, &v, sizeof(T));
    int base = warp() * 32;
    g_blk->xchg[base + lane()] = raw;
    warp_sync();
    int s = src & 31; if (s >= warp_width()) s = lane();
    uint64_t got = g_blk->xchg[base + s];
    warp_sync();
    T out; std::memcpy(&out, &got, sizeof(T)); return out;
}
inline unsigned ballot(int pred) {
    int base = warp() * 32;
    g_blk->xchg[base + lane()] = pred ? 1u : 0u;
    warp_sync();
    unsigned m = 0;
    for (int i = 0; i < warp_width(); i++) if (g_blk->xchg[base + i]) m |= 1u << i;
    warp_sync();
    return m;
}
inline void named_barrier(int id, int nthreads) {
    Barrier* b;
    {
        std::lock_guard<std::mutex> lk(g_blk->named_m);
        if ((int)g_blk->named_bar.size() <= id) g_blk->named_bar.resize(id + 1, nullptr);
        if (!g_blk->named_bar[id]) g_blk->named_bar[id] = new Barrier(nthreads);
        b = g_blk->named_bar[id];
    }
    b->wait();
}

// Run `body` as a grid of blocks. Blocks execute sequentially; threads of a block are real threads.
void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()>& body);

}  // namespace pb_emu

#define threadIdx (pb_emu::t_threadIdx)
#define blockIdx (pb_emu::t_blockIdx)
#define blockDim (pb_emu::t_blockDim)
#define gridDim (pb_emu::t_gridDim)

// ---- warp / block primitives
#define __syncthreads() (pb_emu::g_blk->block_bar.wait())
#define __syncwarp(...) (pb_emu::warp_sync())
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return pb_emu::shfl(v, src); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return pb_emu::shfl(v, pb_emu::lane() ^ m); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) {
    int s = pb_emu::lane() + d; return pb_emu::shfl(v, s > 31 ? pb_emu::lane() : s);
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) {
    int s = pb_emu::lane() - d; return pb_emu::shfl(v, s < 0 ? pb_emu::lane() : s);
}
static inline unsigned __ballot_sync(unsigned, int p) { return pb_emu::ballot(p); }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) {
    const int base = pb_emu::warp() * 32;
    pb_emu::g_blk->xchg[base + pb_emu::lane()] = v;
    pb_emu::warp_sync();
    unsigned r = 0;
    for (int i = 0; i < pb_emu::warp_width(); i++) r |= (unsigned)pb_emu::g_blk->xchg[base + i];
    pb_emu::warp_sync();
    return r;
}
// packed 16-bit SIMD intrinsics used by the kernels
static inline int __dp2a_lo(int a, int b, int c) {
    return c + (int)(short)(a & 0xffff) * (int)(signed char)(b & 0xff) + (int)(short)((unsigned)a >> 16) * (int)(signed char)((b >> 8) & 0xff);
}
static inline unsigned __vmins2(unsigned a, unsigned b) {
    const short al = (short)(a & 0xffff), ah = (short)(a >> 16), bl = (short)(b & 0xffff), bh = (short)(b >> 16);
    return (unsigned)(unsigned short)(al < bl ? al : bl) | ((unsigned)(unsigned short)(ah < bh ? ah : bh) << 16);
}
static inline unsigned __vmaxs2(unsigned a, unsigned b) {
    const short al = (short)(a & 0xffff), ah = (short)(a >> 16), bl = (short)(b & 0xffff), bh = (short)(b >> 16);
    return (unsigned)(unsigned short)(al > bl ? al : bl) | ((unsigned)(unsigned short)(ah > bh ? ah : bh) << 16);
}
// sm_80+ integer warp reductions (REDUX)
static inline int pb_emu_reduce_i(int v, int op) {
    const int base = pb_emu::warp() * 32;
    pb_emu::g_blk->xchg[base + pb_emu::lane()] = (uint64_t)(uint32_t)v;
    pb_emu::warp_sync();
    int r = (int)(uint32_t)pb_emu::g_blk->xchg[base];
    for (int i = 1; i < pb_emu::warp_width(); i++) {
        const int o = (int)(uint32_t)pb_emu::g_blk->xchg[base + i];
        r = op == 0 ? r + o : op == 1 ? (o < r ? o : r) : (o > r ? o : r);
    }
    pb_emu::warp_sync();
    return r;
}
static inline int __reduce_add_sync(unsigned, int v) { return pb_emu_reduce_i(v, 0); }
static inline int __reduce_min_sync(unsigned, int v) { return pb_emu_reduce_i(v, 1); }
static inline int __reduce_max_sync(unsigned, int v) { return pb_emu_reduce_i(v, 2); }
static inline int __any_sync(unsigned, int p) { return pb_emu::ballot(p) != 0; }
static inline int __all_sync(unsigned, int p) {
    unsigned full = pb_emu::warp_width() == 32 ? 0xffffffffu : ((1u << pb_emu::warp_width()) - 1);
    return pb_emu::ballot(p) == full;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }

// ---- atomics (blocks are sequential, threads are real: use GCC atomics)
template <typename T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline float atomicAdd(float* p, float v) {
    float old = *p, nw;
    do { nw = old + v; } while (!__atomic_compare_exchange(p, &old, &nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    return old;
}
static inline double atomicAdd(double* p, double v) {
    double old = *p, nw;
    do { nw = old + v; } while (!__atomic_compare_exchange(p, &old, &nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    return old;
}
template <typename T> static inline T atomicMax(T* p, T v) {
    T old = *p;
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
template <typename T> static inline T atomicMin(T* p, T v) {
    T old = *p;
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}

// ---- loads / math intrinsics
template <typename T> static inline T __ldg(const T* p) { return *p; }
// glibc already declares __cosf & co; route the CUDA fast-math names through macros
static inline float pb_emu_cosf(float x) { return cosf(x); }
static inline float pb_emu_sinf(float x) { return sinf(x); }
static inline float pb_emu_expf(float x) { return expf(x); }
static inline float pb_emu_log2f(float x) { return log2f(x); }
#define __cosf(x) pb_emu_cosf(x)
#define __sinf(x) pb_emu_sinf(x)
#define __expf(x) pb_emu_expf(x)
#define __log2f(x) pb_emu_log2f(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float sinpif(float x) { return (float)sin(M_PI * (double)x); }
static inline float cospif(float x) { return (float)cos(M_PI * (double)x); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
static inline unsigned __float_as_uint(float f) { unsigned v; std::memcpy(&v, &f, 4); return v; }
static inline float __uint_as_float(unsigned v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline int __float2int_rd(float f) { return (int)floorf(f); }
static inline int __double2int_rz(double d) { return (int)d; }
static inline long long __double2ll_rd(double d) { return (long long)floor(d); }
static inline long long __double2ll_ru(double d) { return (long long)ceil(d); }
using std::max;
using std::min;
