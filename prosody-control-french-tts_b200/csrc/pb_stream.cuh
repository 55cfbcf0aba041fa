// pb_stream.cuh — warp-cooperative streaming over a range of s16 samples with 16-byte loads.
// The HBM-bound reductions (K0 unit stats, loudness peak) are one warp per unit built on this.
#pragma once
#include "pb_rt.h"

// Calls f(sample) exactly once for every sample of p[lo, hi), spread over the lanes of the calling warp:
// a scalar head up to the first 16-byte boundary, int4 loads (8 samples) in the middle, a scalar tail.
template <class F>
__device__ __forceinline__ void pb_warp_foreach_s16(const int16_t* __restrict__ p, long long lo, long long hi, int lane, F f) {
    if (hi <= lo) return;
    const size_t mis = ((size_t)(p + lo)) & 15;
    long long head = mis ? (long long)((16 - mis) >> 1) : 0;
    if (head > hi - lo) head = hi - lo;
    if (lane < head) f((int)p[lo + lane]);
    const long long i0 = lo + head;
    const long long nvec = (hi - i0) >> 3;
    const int4* __restrict__ vp = reinterpret_cast<const int4*>(p + i0);
    for (long long k = lane; k < nvec; k += 32) {
        const int4 v = vp[k];
        f((int)(short)(v.x & 0xffff)); f(v.x >> 16); f((int)(short)(v.y & 0xffff)); f(v.y >> 16);
        f((int)(short)(v.z & 0xffff)); f(v.z >> 16); f((int)(short)(v.w & 0xffff)); f(v.w >> 16);
    }
    const long long t0 = i0 + (nvec << 3);
    if (t0 + lane < hi) f((int)p[t0 + lane]);
}
