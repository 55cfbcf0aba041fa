// pb_lufs.cuh — device code of the loudness path (K4) and Praat frame intensity (K4b).
//
// get_lufs (/root/reference/Code/audioPipeline.py:338-358): samples / max|samples| -> pyloudnorm
// Meter(rate).integrated_loudness: K-weighting (RBJ high-shelf then high-pass biquad, scipy.lfilter, float64),
// 400 ms blocks every 100 ms, absolute (-70 LUFS) and relative (-10 LU) gates.
//
// The IIR cascade is a 4-state linear recurrence.  Instead of one sequential pass per unit it is evaluated as a
// chunk-parallel scan: the 100 ms hops of pyloudnorm's block grid are the chunks; (1) every chunk is filtered
// from a zero state to get its state contribution, (2) a tiny per-unit pass chains the chunk states with the
// precomputed transition matrix A^L, (3) every chunk is filtered again from its true initial state while
// accumulating the energy of the weighted signal.  Block j is exactly chunks j..j+3 because pyloudnorm's upper
// bound int(T_g*(j*step+1)*rate) is the lower bound of block j+4 evaluated on the same float64 operands.
// All filter state is float64 (the 38 Hz high-pass has a double pole at radius ~0.99).
#pragma once
#include "pb_rt.h"
#include "pb_stream.cuh"
#include <math.h>

#define PB_LUFS_NM 6          // transition matrices A^(L0) .. A^(L0+NM-1) kept per meter rate

struct PbMeterDev {           // one per distinct pyln.Meter(rate)
    double b1[3], a1[3];      // high shelf  (a[0] == 1)
    double b2[3], a2[3];      // high pass
    double rate;
    int32_t L0;               // smallest chunk length with a precomputed transition matrix
    int32_t pad;
    double M[PB_LUFS_NM][16]; // row-major 4x4, M[k] = A^(L0+k) on the state (p0,p1,q0,q1)
};

struct PbLufsUnitDev {
    int64_t pcm_off;          // file start in the pcm buffer
    int64_t a, b;             // real samples [a, b) of the file ...
    int64_t npad;             // ... followed by npad zeros (pydub's missing-frame padding)
    int64_t chunk_off;        // first chunk of this unit in the chunk arrays
    int32_t n_chunks;         // numBlocks + 3
    int32_t n_blocks;
    int32_t meter;            // index into the meter table
    int32_t out_index;
    double inv_peak;          // filled by the peak kernel: 1 / (max|x| or 1.0)
};

// chunk boundary c of a unit: int(T_g * (c * step) * rate) clipped to n (numpy slice clipping)
__device__ __forceinline__ long long pb_lufs_bound(int c, double rate, long long n) {
    const double v = __dmul_rn(__dmul_rn(0.4, __dmul_rn((double)c, 0.25)), rate);
    long long l = (long long)v;
    return l > n ? n : l;
}

// Peak of every unit (the reference divides by max|samples| before metering). One warp per unit, 16-byte loads.
__global__ void __launch_bounds__(256) pb_lufs_peak_kernel(const int16_t* __restrict__ pcm, PbLufsUnitDev* __restrict__ units, int n_units, int long_chunks) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int u = blockIdx.x * wpb + (threadIdx.x >> 5); u < n_units; u += gridDim.x * wpb) {
        const PbLufsUnitDev ud = units[u];
        if (ud.n_chunks > long_chunks) continue;             // long units: pb_lufs_peak_long_kernel
        int mx = 0;
        pb_warp_foreach_s16(pcm + ud.pcm_off, ud.a, ud.b, lane, [&](int v) { mx = max(mx, v < 0 ? -v : v); });
        PB_UNROLL for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(PB_FULL_MASK, mx, o));
        if (lane == 0) units[u].inv_peak = mx > 0 ? 1.0 / (double)mx : 1.0;
    }
}

__device__ __forceinline__ int pb_lufs_find_unit(const PbLufsUnitDev* __restrict__ units, int n, long long chunk) {
    int lo = 0, hi = n;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (units[mid].chunk_off <= chunk) lo = mid; else hi = mid; }
    return lo;
}

// (1) and (3): one thread per chunk runs the K-weighting recurrence over its samples, eight at a time through aligned
// 16-byte loads.  ENERGY = false: from rest, keeping only the final state (the chunk's zero-state contribution);
// ENERGY = true: from the chunk's true initial state, accumulating the energy of the weighted signal.
// Both passes work on the INTEGER sample values: the filter is linear, so the reference's division by the peak becomes a factor
// inv_peak^2 on the block energies (applied by the gate kernels) and costs nothing per sample.  Every state update is written as
// t = b x + s_next (off the critical path), s = -a y + t — scipy.signal.lfilter's own association (direct form II transposed):
// 10 (+1 for the energy) FP64 operations per sample, two of them on the dependent chain; round 1 spent 13 (+1) with a chain of
// four.  The kernels are FP64-pipe bound (profiles/r02b_lufs_chunk_ncu_summary.csv), so the operation count is the time.
// Formulations of pass (1) without the chain were tried and lost: four dot products against A^j B with one warp per chunk and
// 2-byte strided loads (round 1: no memory-level parallelism, 6.6 ms vs 2.4 ms); thread per chunk against a whole-chunk table in
// global memory (round 2: L1 misses, 555 us vs 353 us); 64-sample sub-blocks against a 2 KB shared-memory table chained through
// A^64 (round 2: 4.25 FP64 operations per sample, but two broadcast 16-byte shared loads per sample and the I2F.F64 conversions
// — 8 lanes per clock — bound it: no faster than the recurrence).
#define PB_LUFS_STEP(xv)                                                  \
    {                                                                     \
        const double x_ = (xv);                                           \
        const double y1_ = fma(b10, x_, p0);                              \
        p0 = fma(-a11, y1_, fma(b11, x_, p1));                            \
        p1 = fma(-a12, y1_, b12 * x_);                                    \
        const double y2_ = fma(b20, y1_, q0);                             \
        q0 = fma(-a21, y2_, fma(b21, y1_, q1));                           \
        q1 = fma(-a22, y2_, b22 * y1_);                                   \
        e = fma(y2_, y2_, e);                                             \
    }
template <bool ENERGY>
__global__ void __launch_bounds__(128, 6)
pb_lufs_chunk_kernel(const int16_t* __restrict__ pcm, const PbLufsUnitDev* __restrict__ units, int n_units,
                     const PbMeterDev* __restrict__ meters, long long n_chunks_total,
                     double* __restrict__ state /* [n_chunks][4] */, double* __restrict__ energy /* [n_chunks] */) {
    for (long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x; ch < n_chunks_total; ch += (long long)gridDim.x * blockDim.x) {
        const int u = pb_lufs_find_unit(units, n_units, ch);
        const PbLufsUnitDev ud = units[u];
        const PbMeterDev* __restrict__ mt = meters + ud.meter;
        const int c = (int)(ch - ud.chunk_off);
        const long long nreal = ud.b - ud.a, n = nreal + ud.npad;
        const long long lo = pb_lufs_bound(c, mt->rate, n), hi = pb_lufs_bound(c + 1, mt->rate, n);
        const long long real_hi = hi < nreal ? hi : nreal;
        const double b10 = mt->b1[0], b11 = mt->b1[1], b12 = mt->b1[2], a11 = mt->a1[1], a12 = mt->a1[2];
        const double b20 = mt->b2[0], b21 = mt->b2[1], b22 = mt->b2[2], a21 = mt->a2[1], a22 = mt->a2[2];
        double p0 = 0.0, p1 = 0.0, q0 = 0.0, q1 = 0.0, e = 0.0;
        if (ENERGY) { p0 = state[ch * 4 + 0]; p1 = state[ch * 4 + 1]; q0 = state[ch * 4 + 2]; q1 = state[ch * 4 + 3]; }
        const int16_t* __restrict__ p = pcm + ud.pcm_off + ud.a;
        long long i = lo;
        while (i < real_hi && (((size_t)(p + i)) & 15)) { PB_LUFS_STEP((double)p[i]); i++; }
        while (i + 8 <= real_hi) {
            const int4 v = *reinterpret_cast<const int4*>(p + i);
            PB_LUFS_STEP((double)(short)(v.x & 0xffff)); PB_LUFS_STEP((double)(v.x >> 16));
            PB_LUFS_STEP((double)(short)(v.y & 0xffff)); PB_LUFS_STEP((double)(v.y >> 16));
            PB_LUFS_STEP((double)(short)(v.z & 0xffff)); PB_LUFS_STEP((double)(v.z >> 16));
            PB_LUFS_STEP((double)(short)(v.w & 0xffff)); PB_LUFS_STEP((double)(v.w >> 16));
            i += 8;
        }
        while (i < real_hi) { PB_LUFS_STEP((double)p[i]); i++; }
        while (i < hi) { PB_LUFS_STEP(0.0); i++; }          // pydub's silent padding
        if (ENERGY) energy[ch] = e;
        else { state[ch * 4 + 0] = p0; state[ch * 4 + 1] = p1; state[ch * 4 + 2] = q0; state[ch * 4 + 3] = q1; }
    }
}

// s <- A^len s for the state whose component r this lane holds (the other three sit in the lanes base_lane .. base_lane + 3).
// Every lane of the warp must call it (shuffles); `on` = this lane's result is wanted.
__device__ __forceinline__ double pb_lufs_advance(const PbMeterDev* __restrict__ mt, int len, int r, int base_lane, bool on, double s) {
    const double s0 = __shfl_sync(PB_FULL_MASK, s, base_lane + 0), s1 = __shfl_sync(PB_FULL_MASK, s, base_lane + 1);
    const double s2 = __shfl_sync(PB_FULL_MASK, s, base_lane + 2), s3 = __shfl_sync(PB_FULL_MASK, s, base_lane + 3);
    if (!on) return s;
    const int k = len - mt->L0;
    if (k >= 0 && k < PB_LUFS_NM) {
        const double* M = mt->M[k] + r * 4;
        return M[0] * s0 + M[1] * s1 + M[2] * s2 + M[3] * s3;
    }
    // a chunk clipped by the end of the unit: step the homogeneous recurrence (only empty chunks follow)
    double a0 = s0, a1 = s1, a2 = s2, a3 = s3;
    for (int i = 0; i < len; i++) {
        const double y1 = a0;
        const double np0 = -mt->a1[1] * y1 + a1, np1 = -mt->a1[2] * y1;
        const double y2 = mt->b2[0] * y1 + a2;
        const double nq0 = mt->b2[1] * y1 - mt->a2[1] * y2 + a3, nq1 = mt->b2[2] * y1 - mt->a2[2] * y2;
        a0 = np0; a1 = np1; a2 = nq0; a3 = nq1;
    }
    return r == 0 ? a0 : r == 1 ? a1 : r == 2 ? a2 : a3;
}

// (2) Per unit: turn the zero-state chunk contributions into true initial states: s_in[c+1] = A^len(c) s_in[c] + s_zs[c].
// Four lanes per unit (one state component each, the 4x4 product via shuffles), eight units per warp; the chunk
// contributions are prefetched one chunk ahead so the sequential chain is only the four FMAs.
__global__ void __launch_bounds__(128)
pb_lufs_scan_kernel(const PbLufsUnitDev* __restrict__ units, int n_units, const PbMeterDev* __restrict__ meters, double* __restrict__ state, int long_chunks) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int r = lane & 3, slot = lane >> 2, base_lane = lane & ~3;
    for (int ubase = (blockIdx.x * wpb + (threadIdx.x >> 5)) * 8; ubase < n_units; ubase += gridDim.x * wpb * 8) {
        const int u = ubase + slot;
        const bool active = u < n_units;
        const PbLufsUnitDev ud = units[active ? u : n_units - 1];
        const PbMeterDev* __restrict__ mt = meters + ud.meter;
        const int nch = (active && ud.n_chunks <= long_chunks) ? ud.n_chunks : 0;      // long units: pb_lufs_scan_long_kernel
        int maxch = nch;
        PB_UNROLL for (int o = 16; o > 0; o >>= 1) maxch = max(maxch, __shfl_xor_sync(PB_FULL_MASK, maxch, o));
        const long long n = (ud.b - ud.a) + ud.npad;
        const double rate = mt->rate;
        double* st = state + ud.chunk_off * 4 + r;
        double s = 0.0;
        double znext = nch > 0 ? st[0] : 0.0;
        long long lo = pb_lufs_bound(0, rate, n);
        for (int c = 0; c < maxch; c++) {
            const bool on = c < nch;
            const double z = znext;                     // zero-state contribution of chunk c (component r)
            if (on) st[(size_t)c * 4] = s;              // becomes its initial state
            if (c + 1 < nch) znext = st[(size_t)(c + 1) * 4];
            const long long hi = pb_lufs_bound(c + 1, rate, n);
            const int len = (int)(hi - lo);
            lo = hi;
            s = pb_lufs_advance(mt, len, r, base_lane, on, s);
            if (on) s += z;
        }
    }
}

// ------------------------------------------------------------------------------------------------ K4, long units
// A one-hour recording measured as one unit has 36 000 chunks: one warp finding its peak, four lanes chaining its states and
// one thread gating its blocks take ~100 ms while the rest of the GPU idles.  Units with more than `long_chunks` chunks go
// through these instead (the chunk kernels (1) and (3) are chunk-parallel already):
//   peak:  one warp per 64 Ki-sample piece, atomicMax into an int per long unit;
//   scan:  the chain of affine maps s -> A^len s + z is cut into groups of `group` chunks; phase 1 composes every group's
//          map (the matrix by running the four basis vectors through it), phase 2 walks the groups of a unit, phase 3 redoes
//          every group from its true entering state;
//   gate:  one CTA per long unit, block sums through a shared-memory tree.
// The states reach a chunk through a different association of the same float64 products (~1e-13 relative).
struct PbLufsLong { int unit; int job_off; int n_groups; int peak; };

__global__ void pb_lufs_long_index_kernel(const PbLufsUnitDev* __restrict__ units, int n_units, int long_chunks, int group,
                                          PbLufsLong* __restrict__ longs, int2* __restrict__ jobs, int* __restrict__ counters) {
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += gridDim.x * blockDim.x) {
        const int nch = units[u].n_chunks;
        if (nch <= long_chunks) continue;
        const int ng = (nch + group - 1) / group;
        const int li = atomicAdd(&counters[0], 1), j0 = atomicAdd(&counters[1], ng);
        PbLufsLong lu; lu.unit = u; lu.job_off = j0; lu.n_groups = ng; lu.peak = 0;
        longs[li] = lu;
        for (int g = 0; g < ng; g++) jobs[j0 + g] = make_int2(li, g);
    }
}

#define PB_LUFS_PEAK_PIECE 65536
__global__ void __launch_bounds__(256)
pb_lufs_peak_long_kernel(const int16_t* __restrict__ pcm, const PbLufsUnitDev* __restrict__ units, PbLufsLong* __restrict__ longs, const int* __restrict__ counters) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int li = 0; li < counters[0]; li++) {
        const PbLufsUnitDev ud = units[longs[li].unit];
        const long long pieces = (ud.b - ud.a + PB_LUFS_PEAK_PIECE - 1) / PB_LUFS_PEAK_PIECE;
        int mx = 0;
        for (long long pc = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); pc < pieces; pc += (long long)gridDim.x * wpb) {
            const long long lo = ud.a + pc * PB_LUFS_PEAK_PIECE, hi = lo + PB_LUFS_PEAK_PIECE < ud.b ? lo + PB_LUFS_PEAK_PIECE : ud.b;
            pb_warp_foreach_s16(pcm + ud.pcm_off, lo, hi, lane, [&](int v) { mx = max(mx, v < 0 ? -v : v); });
        }
        PB_UNROLL for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(PB_FULL_MASK, mx, o));
        if (lane == 0 && mx > 0) atomicMax(&longs[li].peak, mx);
    }
}
__global__ void pb_lufs_peak_fin_kernel(PbLufsUnitDev* __restrict__ units, const PbLufsLong* __restrict__ longs, const int* __restrict__ counters) {
    for (int li = blockIdx.x * blockDim.x + threadIdx.x; li < counters[0]; li += gridDim.x * blockDim.x) {
        const int mx = longs[li].peak;
        units[longs[li].unit].inv_peak = mx > 0 ? 1.0 / (double)mx : 1.0;
    }
}

// PHASE 1: the group's map (P: 4x4, column k = image of basis vector k; Z: image of the zero state) from the chunks' zero-state
// contributions.  PHASE 3: every chunk's true initial state from the group's entering state S.  Four lanes per group, eight per warp.
template <int PHASE>
__global__ void __launch_bounds__(128)
pb_lufs_scan_group_kernel(const PbLufsUnitDev* __restrict__ units, const PbMeterDev* __restrict__ meters, const PbLufsLong* __restrict__ longs,
                          const int2* __restrict__ jobs, const int* __restrict__ counters, int group, double* __restrict__ state,
                          double* __restrict__ P, double* __restrict__ Z, const double* __restrict__ S) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int r = lane & 3, slot = lane >> 2, base_lane = lane & ~3;
    const int n_jobs = counters[1];
    for (int jbase = (blockIdx.x * wpb + (threadIdx.x >> 5)) * 8; jbase < n_jobs; jbase += gridDim.x * wpb * 8) {
        const int job = jbase + slot;
        const bool active = job < n_jobs;
        const int2 jb = jobs[active ? job : n_jobs - 1];
        const PbLufsUnitDev ud = units[longs[jb.x].unit];
        const PbMeterDev* __restrict__ mt = meters + ud.meter;
        const int c0 = jb.y * group, c1 = min(ud.n_chunks, c0 + group);
        const long long n = (ud.b - ud.a) + ud.npad;
        const double rate = mt->rate;
        double* st = state + ud.chunk_off * 4 + r;
        double v0 = r == 0 ? 1.0 : 0.0, v1 = r == 1 ? 1.0 : 0.0, v2 = r == 2 ? 1.0 : 0.0, v3 = r == 3 ? 1.0 : 0.0;
        double s = PHASE == 3 ? S[(size_t)(active ? job : n_jobs - 1) * 4 + r] : 0.0;
        long long lo = pb_lufs_bound(c0, rate, n);
        for (int c = c0; c < c0 + group; c++) {           // uniform trip count: every lane takes part in the shuffles
            const bool on = active && c < c1;
            const double z = on ? st[(size_t)c * 4] : 0.0;
            const long long hi = pb_lufs_bound(c + 1, rate, n);
            const int len = (int)(hi - lo);
            lo = hi;
            if (PHASE == 1) {
                v0 = pb_lufs_advance(mt, len, r, base_lane, on, v0); v1 = pb_lufs_advance(mt, len, r, base_lane, on, v1);
                v2 = pb_lufs_advance(mt, len, r, base_lane, on, v2); v3 = pb_lufs_advance(mt, len, r, base_lane, on, v3);
            } else if (on) st[(size_t)c * 4] = s;
            s = pb_lufs_advance(mt, len, r, base_lane, on, s);
            if (on) s += z;
        }
        if (PHASE == 1 && active) {
            double* p = P + (size_t)job * 16;
            p[0 * 4 + r] = v0; p[1 * 4 + r] = v1; p[2 * 4 + r] = v2; p[3 * 4 + r] = v3;
            Z[(size_t)job * 4 + r] = s;
        }
    }
}

// PHASE 2: entering state of every group, in order: S_{g+1} = P_g S_g + Z_g.  Four lanes per long unit.
__global__ void __launch_bounds__(32)
pb_lufs_scan_link_kernel(const PbLufsLong* __restrict__ longs, const int* __restrict__ counters, const double* __restrict__ P,
                         const double* __restrict__ Z, double* __restrict__ S) {
    const int lane = threadIdx.x & 31, r = lane & 3;
    for (int li = blockIdx.x; li < counters[0]; li += gridDim.x) {
        const PbLufsLong lu = longs[li];
        double s = 0.0;
        for (int g = 0; g < lu.n_groups; g++) {
            const size_t job = (size_t)lu.job_off + g;
            const double* p = P + job * 16;
            const double p0 = p[0 * 4 + r], p1 = p[1 * 4 + r], p2 = p[2 * 4 + r], p3 = p[3 * 4 + r], z = Z[job * 4 + r];   // independent of s
            if (lane < 4) S[job * 4 + r] = s;
            const double s0 = __shfl_sync(PB_FULL_MASK, s, 0), s1 = __shfl_sync(PB_FULL_MASK, s, 1);
            const double s2 = __shfl_sync(PB_FULL_MASK, s, 2), s3 = __shfl_sync(PB_FULL_MASK, s, 3);
            s = p0 * s0 + p1 * s1 + p2 * s2 + p3 * s3 + z;
        }
    }
}

// Gates of a long unit: one CTA, the two passes of pb_lufs_gate_kernel with block-wide sums.
__global__ void __launch_bounds__(256)
pb_lufs_gate_long_kernel(const PbLufsUnitDev* __restrict__ units, const PbLufsLong* __restrict__ longs, const int* __restrict__ counters,
                         const PbMeterDev* __restrict__ meters, const double* __restrict__ energy, double* __restrict__ lufs_out) {
    __shared__ double s_sum[256];
    __shared__ int s_cnt[256];
    const int tid = threadIdx.x;
    const double NEG_INF = -(double)INFINITY;
    for (int li = blockIdx.x; li < counters[0]; li += gridDim.x) {
        const PbLufsUnitDev ud = units[longs[li].unit];
        const double inv = ud.inv_peak * ud.inv_peak / (0.4 * meters[ud.meter].rate);   // the chunk energies are those of the integer samples
        const double* e = energy + ud.chunk_off;
        double gamma_r = 0.0, out = NEG_INF;
        for (int pass = 0; pass < 2; pass++) {
            double sum = 0.0; int cnt = 0;
            for (int j = tid; j < ud.n_blocks; j += 256) {
                const double z = inv * (e[j] + e[j + 1] + e[j + 2] + e[j + 3]);
                const double l = z > 0.0 ? -0.691 + 10.0 * log10(z) : NEG_INF;
                if (pass == 0 ? l >= -70.0 : (l > gamma_r && l > -70.0)) { sum += z; cnt++; }
            }
            s_sum[tid] = sum; s_cnt[tid] = cnt;
            __syncthreads();
            for (int o = 128; o > 0; o >>= 1) {
                if (tid < o) { s_sum[tid] += s_sum[tid + o]; s_cnt[tid] += s_cnt[tid + o]; }
                __syncthreads();
            }
            sum = s_sum[0]; cnt = s_cnt[0];
            __syncthreads();
            if (cnt == 0) break;                                       // uniform: every thread read the same totals
            if (pass == 0) gamma_r = -0.691 + 10.0 * log10(sum / (double)cnt) - 10.0;
            else { const double za = sum / (double)cnt; out = za > 0.0 ? -0.691 + 10.0 * log10(za) : NEG_INF; }
        }
        if (tid == 0) lufs_out[ud.out_index] = out;
    }
}

// Per unit: block energies from 4 consecutive chunks, the two gates, LUFS (pyloudnorm meter.py).
__global__ void __launch_bounds__(128)
pb_lufs_gate_kernel(const PbLufsUnitDev* __restrict__ units, int n_units, const PbMeterDev* __restrict__ meters,
                    const double* __restrict__ energy, double* __restrict__ lufs_out, int long_chunks) {
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += gridDim.x * blockDim.x) {
        const PbLufsUnitDev ud = units[u];
        if (ud.n_chunks > long_chunks) continue;             // long units: pb_lufs_gate_long_kernel
        const double rate = meters[ud.meter].rate;
        const double inv = ud.inv_peak * ud.inv_peak / (0.4 * rate);      // the chunk energies are those of the integer samples
        const double* e = energy + ud.chunk_off;
        const double NEG_INF = -(double)INFINITY;
        // absolute gate
        double sum = 0.0; int cnt = 0;
        for (int j = 0; j < ud.n_blocks; j++) {
            const double z = inv * (e[j] + e[j + 1] + e[j + 2] + e[j + 3]);
            const double l = z > 0.0 ? -0.691 + 10.0 * log10(z) : NEG_INF;
            if (l >= -70.0) { sum += z; cnt++; }
        }
        double out = NEG_INF;
        if (cnt > 0) {
            const double gamma_r = -0.691 + 10.0 * log10(sum / (double)cnt) - 10.0;
            sum = 0.0; cnt = 0;
            for (int j = 0; j < ud.n_blocks; j++) {
                const double z = inv * (e[j] + e[j + 1] + e[j + 2] + e[j + 3]);
                const double l = z > 0.0 ? -0.691 + 10.0 * log10(z) : NEG_INF;
                if (l > gamma_r && l > -70.0) { sum += z; cnt++; }
            }
            if (cnt > 0) { const double za = sum / (double)cnt; out = za > 0.0 ? -0.691 + 10.0 * log10(za) : NEG_INF; }
        }
        lufs_out[ud.out_index] = out;
    }
}

// ------------------------------------------------------------------------------------------------ K4b: Praat Sound_to_Intensity
// parselmouth Sound.to_intensity() (Code/visualisation/Compare_speech_noenhanced.py:19-26) = Praat Sound_to_Intensity:
// per frame, the samples within +-halfWindow of the frame centre (clipped to the sound), mean-subtracted, weighted by
// Praat's Bessel window, mean energy re 4e-10 in dB.  One warp per frame, float64 accumulation; the integer sum for the
// mean is exact.  (Frames overlap 8x; the 2 B/sample re-reads come from L1/L2.)
struct PbIntensityUnitDev {
    int64_t pcm_off;      // file start in the pcm buffer
    int64_t frame_off;    // first frame of this unit in the output
    double t1;            // time of the first frame
    int32_t nx;           // samples in the file
    int32_t n_frames;
};
__global__ void __launch_bounds__(128)
pb_intensity_kernel(const int16_t* __restrict__ pcm, const PbIntensityUnitDev* __restrict__ units, int n_units, long long n_frames_total,
                    const double* __restrict__ window /* [2*half+1] */, int half, double dx, double dt, int subtract_mean,
                    float* __restrict__ out_db) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (long long fr = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); fr < n_frames_total; fr += (long long)gridDim.x * wpb) {
        int lo = 0, hi = n_units;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (units[mid].frame_off <= fr) lo = mid; else hi = mid; }
        const PbIntensityUnitDev ud = units[lo];
        const int f = (int)(fr - ud.frame_off);
        const double x1 = 0.5 * dx;
        const double t = __dadd_rn(ud.t1, __dmul_rn((double)f, dt));
        // Sampled_xToNearestIndex: round((t - x1)/dx + 1)
        const long long mid = (long long)floor(__dadd_rn(__dadd_rn(__ddiv_rn(__dsub_rn(t, x1), dx), 1.0), 0.5));
        long long l = mid - half, r = mid + half;
        if (l < 1) l = 1;
        if (r > ud.nx) r = ud.nx;
        const int16_t* __restrict__ p = pcm + ud.pcm_off;
        long long s = 0;
        for (long long i = l + lane; i <= r; i += 32) s += p[i - 1];
        PB_UNROLL for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(PB_FULL_MASK, s, o);
        const double mean = (subtract_mean && r >= l) ? ((double)s / 32768.0) / (double)(r - l + 1) : 0.0;
        double sumxw = 0.0, sumw = 0.0;
        for (long long i = l + lane; i <= r; i += 32) {
            const double w = window[(int)(i - mid + half)];
            const double x = (double)p[i - 1] / 32768.0 - mean;
            sumxw += x * x * w; sumw += w;
        }
        PB_UNROLL for (int o = 16; o > 0; o >>= 1) {
            sumxw += __shfl_xor_sync(PB_FULL_MASK, sumxw, o); sumw += __shfl_xor_sync(PB_FULL_MASK, sumw, o);
        }
        if (lane == 0) {
            double I = sumw > 0.0 ? sumxw / sumw : 0.0;
            I /= 4.0e-10;
            out_db[fr] = I < 1.0e-30 ? -300.0f : (float)(10.0 * log10(I));
        }
    }
}

// ------------------------------------------------------------------------------------------------ legacy RMS loudness
// _calculate_loudness (Code/Pipeline/compute_loudness_adjustments.py:8-25) squares the samples in int16, so every
// square wraps modulo 2^16 before the mean is taken.  One warp per unit: exact int64 sum of the wrapped squares.
struct PbRangeDev { int64_t first; int64_t count; };
__global__ void __launch_bounds__(256)
pb_wrapped_square_sum_kernel(const int16_t* __restrict__ pcm, const PbRangeDev* __restrict__ ranges, int n, long long* __restrict__ out) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int u = blockIdx.x * wpb + (threadIdx.x >> 5); u < n; u += gridDim.x * wpb) {
        const PbRangeDev rg = ranges[u];
        long long sum = 0;
        pb_warp_foreach_s16(pcm, rg.first, rg.first + rg.count, lane, [&](int v) { sum += (long long)(short)(v * v); });
        PB_UNROLL for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(PB_FULL_MASK, sum, o);
        if (lane == 0) out[u] = sum;
    }
}
