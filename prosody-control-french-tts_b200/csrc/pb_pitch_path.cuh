// pb_pitch_path.cuh — K3: Viterbi path finder + median of the voiced frames (see pb_pitch.cuh for the overview).
#pragma once
#include "pb_pitch.cuh"

// ------------------------------------------------------------------------------------------------ K3: path finder + median
// One warp per unit; lane c2 owns candidate c2 of the current frame.  Praat Pitch_pathFinder (fon/Pitch.cpp):
// Viterbi over the candidate lattice in float64, strict '>' so the lowest-index predecessor wins ties.
// Back-pointers: with <= 16 candidates (Praat's default is 15) the 4-bit predecessors of a frame are OR-reduced into one
// 64-bit word; the backtrack then pulls 256 frames of words into shared memory at a time, one lane chases the chain
// there (shared-memory latency instead of a dependent global load per frame) and all lanes gather the selected
// frequency / strength (parselmouth's selected_array).  Finally np.median of the frequencies > 0 (mean of the two middle
// values) by bisection on the float bit patterns.
__device__ __forceinline__ double pb_shfl_d(double v, int src) { return __shfl_sync(PB_FULL_MASK, v, src); }

#define PB_PATH_WARPS 4
#define PB_PATH_CHUNK 256

__global__ void __launch_bounds__(PB_PATH_WARPS * 32)
pb_pitch_path_kernel(const PbUnitDev* __restrict__ units, PbPitchGeomDev gm, const float* __restrict__ cand_f,
                     const float* __restrict__ cand_s, const uint8_t* __restrict__ ncand, const float* __restrict__ intensity,
                     uint8_t* __restrict__ psi, float* __restrict__ sel_f, float* __restrict__ sel_s,
                     double* __restrict__ median_out, int32_t* __restrict__ nvoiced_out) {
    __shared__ unsigned long long s_psi[PB_PATH_WARPS][PB_PATH_CHUNK];
    __shared__ uint8_t s_pl[PB_PATH_WARPS][PB_PATH_CHUNK];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int maxc = gm.max_cand;
    const bool packed = maxc <= 16;
    unsigned long long* __restrict__ psi64 = reinterpret_cast<unsigned long long*>(psi);
    const double tcorr = 0.01 / gm.dt;
    const double ojc = gm.octave_jump_cost * tcorr, vuc = gm.voiced_unvoiced_cost * tcorr;
    const double l2_ceiling = log2(gm.ceiling);
    const double us_scale = gm.silence_threshold > 0.0 ? 1.0 / (gm.silence_threshold / (1.0 + gm.voicing_threshold)) : 0.0;
    for (int u = blockIdx.x * wpb + w; u < gm.n_units; u += gridDim.x * wpb) {
        const PbUnitDev ud = units[u];
        const int nF = ud.n_frames;
        const int64_t f0 = ud.frame_off;
        if (ud.global_peak == 0.0) {
            // Praat returns before the path finder: every frame voiceless
            for (int f = lane; f < nF; f += 32) { sel_f[f0 + f] = 0.0f; sel_s[f0 + f] = 0.0f; }
            if (lane == 0) { median_out[ud.out_index] = 0.0; nvoiced_out[ud.out_index] = 0; }
            continue;
        }
        double delta_prev = 0.0, l2_prev = 0.0;     // of candidate `lane` in the previous frame
        int voiced_prev = 0, nc_prev = 0;
        // two frames of lookahead on the candidate loads
        int nc_a = ncand[f0], nc_b = nF > 1 ? ncand[f0 + 1] : 0;
        float cf_a = lane < nc_a ? cand_f[f0 * maxc + lane] : 0.0f, cs_a = lane < nc_a ? cand_s[f0 * maxc + lane] : 0.0f, in_a = intensity[f0];
        float cf_b = 0.0f, cs_b = 0.0f, in_b = 0.0f;
        if (nF > 1) { cf_b = lane < nc_b ? cand_f[(f0 + 1) * maxc + lane] : 0.0f; cs_b = lane < nc_b ? cand_s[(f0 + 1) * maxc + lane] : 0.0f; in_b = intensity[f0 + 1]; }
        for (int f = 0; f < nF; f++) {
            const int nc = nc_a; const float cf = cf_a, cs = cs_a, inten = in_a;
            nc_a = nc_b; cf_a = cf_b; cs_a = cs_b; in_a = in_b;
            if (f + 2 < nF) {
                const int64_t fr = f0 + f + 2;
                nc_b = ncand[fr];
                cf_b = lane < nc_b ? cand_f[fr * maxc + lane] : 0.0f; cs_b = lane < nc_b ? cand_s[fr * maxc + lane] : 0.0f;
                in_b = intensity[fr];
            }
            const double fr_d = (double)cf;
            const int voiced = fr_d > 0.0 && fr_d < gm.ceiling;
            double us = gm.silence_threshold <= 0.0 ? 0.0 : 2.0 - (double)inten * us_scale;
            us = gm.voicing_threshold + (us > 0.0 ? us : 0.0);
            // the candidate frequencies are float32 (relative error ~1e-4 against Praat's): a float32 log2 (error ~1e-7) costs
            // a tenth of the software float64 one and changes no decision
            const double l2 = voiced ? (double)log2f(cf) : 0.0;
            const double local = voiced ? (double)cs - gm.octave_cost_d * (l2_ceiling - l2) : us;
            double best = local; int place = 0;
            if (f > 0) {
                best = -1.0e30; place = -1;
                for (int c1 = 0; c1 < nc_prev; c1++) {
                    const double dp = pb_shfl_d(delta_prev, c1), lp = pb_shfl_d(l2_prev, c1);
                    const int vp = __shfl_sync(PB_FULL_MASK, voiced_prev, c1);
                    double cost;
                    if (!voiced) cost = vp ? vuc : 0.0;
                    else cost = vp ? ojc * fabs(lp - l2) : vuc;
                    const double value = __dadd_rn(__dsub_rn(dp, cost), local);
                    if (value > best) { best = value; place = c1; }
                }
                if (packed) {
                    const unsigned nib = (lane < nc && place >= 0) ? (unsigned)place : 0u;
                    const unsigned lo = __reduce_or_sync(PB_FULL_MASK, lane < 8 ? nib << (4 * lane) : 0u);
                    const unsigned hi = __reduce_or_sync(PB_FULL_MASK, (lane >= 8 && lane < 16) ? nib << (4 * (lane - 8)) : 0u);
                    if (lane == 0) psi64[f0 + f] = ((unsigned long long)hi << 32) | lo;
                } else if (lane < nc) psi[(f0 + f) * maxc + lane] = (uint8_t)place;
            }
            delta_prev = best; l2_prev = l2; voiced_prev = voiced; nc_prev = nc;
        }
        // terminal candidate: first maximum
        double bv = lane < nc_prev ? delta_prev : -1.0e300; int bi = lane;
        PB_UNROLL for (int o = 16; o > 0; o >>= 1) {
            const double ov = pb_shfl_d(bv, lane ^ o); const int oi = __shfl_xor_sync(PB_FULL_MASK, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        __syncwarp();
        if (packed) {
            // ---- backtrack in chunks of PB_PATH_CHUNK frames, newest first
            int place = bi;                                              // candidate chosen at the top frame of the chunk
            for (int c1 = nF; c1 > 0; c1 -= PB_PATH_CHUNK) {
                const int c0 = c1 > PB_PATH_CHUNK ? c1 - PB_PATH_CHUNK : 0;      // frames [c0, c1)
                for (int f = c0 + lane; f < c1; f += 32) s_psi[w][f - c0] = f > 0 ? psi64[f0 + f] : 0ull;
                __syncwarp();
                if (lane == 0) {
                    for (int f = c1 - 1; f >= c0; f--) {
                        s_pl[w][f - c0] = (uint8_t)place;
                        place = (int)((s_psi[w][f - c0] >> (4 * place)) & 15ull);   // predecessor in frame f-1
                    }
                }
                place = __shfl_sync(PB_FULL_MASK, place, 0);
                __syncwarp();
                for (int f = c0 + lane; f < c1; f += 32) {
                    const int64_t fr = f0 + f;
                    const int pl = s_pl[w][f - c0];
                    sel_f[fr] = cand_f[fr * maxc + pl]; sel_s[fr] = cand_s[fr * maxc + pl];
                }
                __syncwarp();
            }
        } else if (lane == 0) {
            int place = bi;
            for (int f = nF - 1; f >= 0; f--) {
                const int64_t fr = f0 + f;
                sel_f[fr] = cand_f[fr * maxc + place]; sel_s[fr] = cand_s[fr * maxc + place];
                if (f > 0) place = psi[fr * maxc + place];
            }
        }
        __syncwarp();
        // ---- np.median(freqs[freqs > 0]) : positive floats order like their bit patterns
        int nv = 0;
        for (int f = lane; f < nF; f += 32) nv += sel_f[f0 + f] > 0.0f;
        nv = pb_warp_sum_i(nv);
        double med = 0.0;
        if (nv > 0) {
            const int k = (nv - 1) >> 1;                       // 0-based rank of the lower middle
            unsigned lo = 0u, hi = 0x7f800000u;                // smallest pattern with count(<= pattern) >= k+1
            while (lo < hi) {
                const unsigned mid = lo + ((hi - lo) >> 1);
                int c = 0;
                for (int f = lane; f < nF; f += 32) { const float v = sel_f[f0 + f]; c += (v > 0.0f && __float_as_uint(v) <= mid); }
                c = pb_warp_sum_i(c);
                if (c >= k + 1) hi = mid; else lo = mid + 1;
            }
            const float lower = __uint_as_float(lo);
            float upper = lower;
            if ((nv & 1) == 0) {
                // the next order statistic: lower again if enough duplicates, else the smallest value above it
                int c = 0; float nxt = 3.0e38f;
                for (int f = lane; f < nF; f += 32) {
                    const float v = sel_f[f0 + f];
                    if (v > 0.0f) { if (v <= lower) c++; else nxt = fminf(nxt, v); }
                }
                c = pb_warp_sum_i(c);
                PB_UNROLL for (int o = 16; o > 0; o >>= 1) nxt = fminf(nxt, __shfl_xor_sync(PB_FULL_MASK, nxt, o));
                upper = (c >= k + 2) ? lower : nxt;
            }
            med = ((double)lower + (double)upper) / 2.0;
        }
        if (lane == 0) { median_out[ud.out_index] = med; nvoiced_out[ud.out_index] = nv; }
    }
}
