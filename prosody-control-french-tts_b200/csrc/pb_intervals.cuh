// pb_intervals.cuh — K6: segmented reduction of per-frame tracks over word / syntagme intervals.
//
// "Aggregates both [F0 and intensity] over the word intervals from the aligners" (BASELINE.json north_star; SURVEY.md §8b
// pb_reduce_intervals).  One warp per interval: the frames whose centre lies inside [tmin, tmax] (the host resolves that to a
// frame range in float64, Praat's Sampled_getWindowSamples rule) are reduced to
//     n_frames, n_voiced (f0 > 0), median and mean of the voiced f0, mean of the second track (intensity).
// The median is np.median (mean of the two middle order statistics), found by bisection on the float bit patterns —
// positive floats order like their bits — so no sorting and no scratch memory.  HBM-bound: 8 B per frame, each frame read
// once per bisection step from L1/L2 (word intervals hold a few dozen frames).
#pragma once
#include "pb_pitch.cuh"

struct PbIntervalDev { long long first; int count; int out_index; };   // frame range of one interval in the track arrays

__global__ void __launch_bounds__(256)
pb_interval_reduce_kernel(const float* __restrict__ f0, const float* __restrict__ track2, const PbIntervalDev* __restrict__ iv, int n_iv,
                          int* __restrict__ n_voiced, double* __restrict__ median_f0, double* __restrict__ mean_f0, double* __restrict__ mean_2) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n_iv; i += gridDim.x * wpb) {
        const PbIntervalDev d = iv[i];
        const float* __restrict__ v = f0 + d.first;
        int nv = 0; double sf = 0.0, s2 = 0.0;
        for (int k = lane; k < d.count; k += 32) {
            const float x = v[k];
            if (x > 0.0f) { nv++; sf += (double)x; }
            if (track2) s2 += (double)track2[d.first + k];
        }
        nv = pb_warp_sum_i(nv);
        PB_UNROLL for (int o = 16; o > 0; o >>= 1) {
            sf += __shfl_xor_sync(PB_FULL_MASK, sf, o);
            s2 += __shfl_xor_sync(PB_FULL_MASK, s2, o);
        }
        double med = 0.0;
        if (nv > 0) {
            const int r = (nv - 1) >> 1;                        // 0-based rank of the lower middle
            unsigned lo = 0u, hi = 0x7f800000u;                 // smallest bit pattern with count(<= pattern) >= r + 1
            while (lo < hi) {
                const unsigned mid = lo + ((hi - lo) >> 1);
                int c = 0;
                for (int k = lane; k < d.count; k += 32) { const float x = v[k]; c += (x > 0.0f && __float_as_uint(x) <= mid); }
                c = pb_warp_sum_i(c);
                if (c >= r + 1) hi = mid; else lo = mid + 1;
            }
            const float lower = __uint_as_float(lo);
            float upper = lower;
            if ((nv & 1) == 0) {                                // the next order statistic
                int c = 0; float nxt = 3.0e38f;
                for (int k = lane; k < d.count; k += 32) { const float x = v[k]; if (x > 0.0f) { if (x <= lower) c++; else nxt = fminf(nxt, x); } }
                c = pb_warp_sum_i(c);
                PB_UNROLL for (int o = 16; o > 0; o >>= 1) nxt = fminf(nxt, __shfl_xor_sync(PB_FULL_MASK, nxt, o));
                upper = (c >= r + 2) ? lower : nxt;
            }
            med = ((double)lower + (double)upper) / 2.0;
        }
        if (lane == 0) {
            n_voiced[d.out_index] = nv;
            median_f0[d.out_index] = med;                                               // 0.0 when nothing is voiced, like get_median_pitch
            mean_f0[d.out_index] = nv > 0 ? sf / (double)nv : 0.0;
            mean_2[d.out_index] = d.count > 0 ? s2 / (double)d.count : 0.0;
        }
    }
}
