// pb_pitch_path.cuh — K3: Viterbi path finder + median of the voiced frames (see pb_pitch.cuh for the overview).
#pragma once
#include "pb_async.cuh"
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

// The constants of the recursion (Praat: timeStepCorrection = 0.01 / dt scales the two transition costs).
struct PbPathConsts { double ojc, vuc, l2_ceiling, us_scale, ceiling, voicing_threshold, silence_threshold, octave_cost; };
__device__ __forceinline__ PbPathConsts pb_path_consts(const PbPitchGeomDev& gm) {
    PbPathConsts k;
    const double tcorr = 0.01 / gm.dt;
    k.ojc = gm.octave_jump_cost * tcorr; k.vuc = gm.voiced_unvoiced_cost * tcorr;
    k.l2_ceiling = log2(gm.ceiling);
    k.us_scale = gm.silence_threshold > 0.0 ? 1.0 / (gm.silence_threshold / (1.0 + gm.voicing_threshold)) : 0.0;
    k.ceiling = gm.ceiling; k.voicing_threshold = gm.voicing_threshold; k.silence_threshold = gm.silence_threshold; k.octave_cost = gm.octave_cost_d;
    return k;
}

// What the recursion carries from one frame to the next, for this lane's candidate.
struct PbPathState { double delta, l2; int voiced, nc; };

// voiced flag and log2 frequency of a candidate.  The candidate frequencies are float32 (relative error ~1e-4 against Praat's): a
// float32 log2 (error ~1e-7) costs a tenth of the software float64 one and changes no decision.
__device__ __forceinline__ void pb_path_cand_props(const PbPathConsts& k, float cf, int& voiced, double& l2) {
    const double fr_d = (double)cf;
    voiced = fr_d > 0.0 && fr_d < k.ceiling;
    l2 = voiced ? (double)log2f(cf) : 0.0;
}

// The forward recursion over frames [fb, fe) of the unit whose first frame is f0.  c is this lane's candidate and hb the first
// lane of its half-warp (0 and the lane itself unless two recursions share a warp, pb_path_block_entry_kernel).  On entry `st`
// describes frame fb - 1 (ignored when fb == 0: Praat starts from the local values); on exit, frame fe - 1.  STORE: the
// back-pointers go to psi64 / psi (and, when s_psi is given, the packed words of this range to shared memory as well).
template <bool STORE>
__device__ __forceinline__ void pb_path_forward(const PbPathConsts& k, int maxc, bool packed, const float* __restrict__ cand_f,
                                                const float* __restrict__ cand_s, const uint8_t* __restrict__ ncand, const float* __restrict__ intensity,
                                                int64_t f0, int fb, int fe, int c, int hb, int lane, PbPathState& st,
                                                unsigned long long* __restrict__ psi64, uint8_t* __restrict__ psi, unsigned long long* s_psi) {
    double delta_prev = st.delta, l2_prev = st.l2;     // of candidate c in the previous frame
    int voiced_prev = st.voiced, nc_prev = st.nc;
    // two frames of lookahead on the candidate loads
    int nc_a = ncand[f0 + fb], nc_b = fb + 1 < fe ? ncand[f0 + fb + 1] : 0;
    float cf_a = c < nc_a ? cand_f[(f0 + fb) * maxc + c] : 0.0f, cs_a = c < nc_a ? cand_s[(f0 + fb) * maxc + c] : 0.0f, in_a = intensity[f0 + fb];
    float cf_b = 0.0f, cs_b = 0.0f, in_b = 0.0f;
    if (fb + 1 < fe) { cf_b = c < nc_b ? cand_f[(f0 + fb + 1) * maxc + c] : 0.0f; cs_b = c < nc_b ? cand_s[(f0 + fb + 1) * maxc + c] : 0.0f; in_b = intensity[f0 + fb + 1]; }
    for (int f = fb; f < fe; f++) {
        const int nc = nc_a; const float cf = cf_a, cs = cs_a, inten = in_a;
        nc_a = nc_b; cf_a = cf_b; cs_a = cs_b; in_a = in_b;
        if (f + 2 < fe) {
            const int64_t fr = f0 + f + 2;
            nc_b = ncand[fr];
            cf_b = c < nc_b ? cand_f[fr * maxc + c] : 0.0f; cs_b = c < nc_b ? cand_s[fr * maxc + c] : 0.0f;
            in_b = intensity[fr];
        }
        int voiced; double l2;
        pb_path_cand_props(k, cf, voiced, l2);
        double us = k.silence_threshold <= 0.0 ? 0.0 : 2.0 - (double)inten * k.us_scale;
        us = k.voicing_threshold + (us > 0.0 ? us : 0.0);
        const double local = voiced ? (double)cs - k.octave_cost * (k.l2_ceiling - l2) : us;
        double best = local; int place = 0;
        if (f > 0) {
            best = -1.0e30; place = -1;
            for (int c1 = 0; c1 < nc_prev; c1++) {
                const double dp = pb_shfl_d(delta_prev, hb + c1), lp = pb_shfl_d(l2_prev, hb + c1);
                const int vp = __shfl_sync(PB_FULL_MASK, voiced_prev, hb + c1);
                double cost;
                if (!voiced) cost = vp ? k.vuc : 0.0;
                else cost = vp ? k.ojc * fabs(lp - l2) : k.vuc;
                const double value = __dadd_rn(__dsub_rn(dp, cost), local);
                if (value > best) { best = value; place = c1; }
            }
        }
        if (STORE) {
            if (packed) {
                const unsigned nib = (f > 0 && lane < nc && place >= 0) ? (unsigned)place : 0u;
                const unsigned lo = __reduce_or_sync(PB_FULL_MASK, lane < 8 ? nib << (4 * lane) : 0u);
                const unsigned hi = __reduce_or_sync(PB_FULL_MASK, (lane >= 8 && lane < 16) ? nib << (4 * (lane - 8)) : 0u);
                if (lane == 0) {
                    const unsigned long long word = ((unsigned long long)hi << 32) | lo;
                    if (f > 0) psi64[f0 + f] = word;
                    if (s_psi) s_psi[f - fb] = word;
                }
            } else if (f > 0 && lane < nc) psi[(f0 + f) * maxc + lane] = (uint8_t)place;
        }
        delta_prev = best; l2_prev = l2; voiced_prev = voiced; nc_prev = nc;
    }
    st.delta = delta_prev; st.l2 = l2_prev; st.voiced = voiced_prev; st.nc = nc_prev;
}

// np.median(freqs[freqs > 0]) of sel[0..n) by the threads of one warp (NT = 32) or one CTA (NT = blockDim.x, `scratch` = two ints of
// shared memory): positive floats order like their bit patterns, so the lower middle is found by bisection on the pattern and
// the upper middle is either the same value (duplicates) or the smallest value above it.  Returns the count of positives in nv.
// f(v) for v = sel[tid], sel[tid + nt], ...: eight independent loads in flight per thread (the bisection below streams the
// selected frequencies ~32 times; one load at a time it ran at L2 latency)
template <typename F>
__device__ __forceinline__ void pb_foreach_strided(const float* __restrict__ sel, int n, int tid, int nt, F f) {
    int i = tid;
    for (; i + 7 * nt < n; i += 8 * nt) {
        float v[8];
        PB_UNROLL for (int q = 0; q < 8; q++) v[q] = sel[i + q * nt];
        PB_UNROLL for (int q = 0; q < 8; q++) f(v[q]);
    }
    for (; i < n; i += nt) f(sel[i]);
}

template <bool CTA>
__device__ __forceinline__ double pb_median_positive(const float* __restrict__ sel, int n, int tid, int nt, int* scratch, int& nv_out) {
    auto total = [&](int v) -> int {
        v = pb_warp_sum_i(v);
        if (!CTA) return v;
        __syncthreads();
        if (tid == 0) scratch[0] = 0;
        __syncthreads();
        if ((tid & 31) == 0 && v) atomicAdd(&scratch[0], v);
        __syncthreads();
        return scratch[0];
    };
    int nv = 0;
    pb_foreach_strided(sel, n, tid, nt, [&](float v) { nv += v > 0.0f; });
    nv = total(nv);
    nv_out = nv;
    if (nv <= 0) return 0.0;
    const int kk = (nv - 1) >> 1;                      // 0-based rank of the lower middle
    unsigned lo = 0u, hi = 0x7f800000u;                // smallest pattern with count(<= pattern) >= kk+1
    while (lo < hi) {
        const unsigned mid = lo + ((hi - lo) >> 1);
        int cnt = 0;
        pb_foreach_strided(sel, n, tid, nt, [&](float v) { cnt += (v > 0.0f && __float_as_uint(v) <= mid); });
        cnt = total(cnt);
        if (cnt >= kk + 1) hi = mid; else lo = mid + 1;
    }
    const float lower = __uint_as_float(lo);
    float upper = lower;
    if ((nv & 1) == 0) {
        // the next order statistic: lower again if enough duplicates, else the smallest value above it
        int cnt = 0; float nxt = 3.0e38f;
        pb_foreach_strided(sel, n, tid, nt, [&](float v) { if (v > 0.0f) { if (v <= lower) cnt++; else nxt = fminf(nxt, v); } });
        cnt = total(cnt);
        PB_UNROLL for (int o = 16; o > 0; o >>= 1) nxt = fminf(nxt, __shfl_xor_sync(PB_FULL_MASK, nxt, o));
        if (CTA) {
            __syncthreads();
            if (tid == 0) scratch[1] = 0x7f7fffff;      // bit pattern of FLT_MAX: positive floats compare like their patterns
            __syncthreads();
            if ((tid & 31) == 0) atomicMin(&scratch[1], (int)__float_as_uint(nxt));
            __syncthreads();
            nxt = __uint_as_float((unsigned)scratch[1]);
        }
        upper = (cnt >= kk + 2) ? lower : nxt;
    }
    return ((double)lower + (double)upper) / 2.0;
}

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
    const PbPathConsts k = pb_path_consts(gm);
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
        if (packed && nF > gm.path_long) continue;          // long chains: the blocked path finder below (pb_path_block_*_kernel)
        PbPathState st; st.delta = 0.0; st.l2 = 0.0; st.voiced = 0; st.nc = 0;
        pb_path_forward<true>(k, maxc, packed, cand_f, cand_s, ncand, intensity, f0, 0, nF, lane, 0, lane, st, psi64, psi, nullptr);
        const double delta_prev = st.delta; const int nc_prev = st.nc;
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
        int nv = 0;
        const double med = pb_median_positive<false>(sel_f + f0, nF, lane, 32, nullptr, nv);
        if (lane == 0) { median_out[ud.out_index] = med; nvoiced_out[ud.out_index] = nv; }
    }
}

// ------------------------------------------------------------------------------------------------ K3, long chains
// One warp walking a 360 000-frame chain (a one-hour recording analysed as one sound: BASELINE config 4) takes ~0.6 us per frame:
// 230 ms during which the rest of the GPU idles.  The recursion is a product of (max, +) matrices, which is associative, so a
// long unit is cut into blocks of `path_block` frames and
//   (a) pb_path_block_entry_kernel: for every block and every candidate e of the frame before it, the recursion is run over the
//       block from "e has score 0, every other candidate -inf": row e of the block's transfer matrix T (two rows per warp, one per
//       half-warp).  Block 0 starts the way Praat does and yields the scores themselves;
//   (b) pb_path_block_scan_kernel: one warp per long unit folds the matrices in order, D_b[c] = max_e D_{b-1}[e] + T_b[e][c]
//       (a few hundred steps instead of a few hundred thousand), keeps every D_b and picks the terminal candidate;
//   (c) pb_path_block_final_kernel: every block is run again from its true entering scores D_{b-1}, exactly the sequential
//       recursion, now storing the back-pointers; each lane then chases "leaves the block through candidate c" back to the frame
//       before the block: the block's back-map (16 nibbles);
//   (d) pb_path_block_link_kernel: one warp per long unit walks the back-maps from the terminal candidate: the candidate every
//       block is left through;
//   (e) pb_path_block_select_kernel: every block backtracks from that candidate and gathers selected_array;
//   (f) pb_path_median_long_kernel: np.median of the voiced frequencies, one CTA per long unit.
// The scores reach a block through sums associated differently from the sequential walk (D + T instead of frame by frame): equal
// in exact arithmetic, ~1e-10 apart in float64 at scores of ~1e5, so a decision could differ only where two paths tie to that
// precision.  Units up to path_long frames keep the one-warp kernel above.
#define PB_PATHL_BLOCK_MAX 512
#define PB_NEG_INF_D (-__builtin_huge_val())
#define PB_PATHL_WARPS 4

struct PbLongUnit { int unit; int job_off; int n_blocks; int terminal; };    // terminal: candidate of the last frame (filled by the scan)

// Lists the long units of a launch group and their blocks.  counters[0] = long units, counters[1] = block jobs (zeroed before).
__global__ void pb_path_long_index_kernel(const PbUnitDev* __restrict__ units, PbPitchGeomDev gm, PbLongUnit* __restrict__ longs,
                                          int2* __restrict__ jobs, int* __restrict__ counters) {
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < gm.n_units; u += gridDim.x * blockDim.x) {
        const int nF = units[u].n_frames;
        if (nF <= gm.path_long || units[u].global_peak == 0.0) continue;
        const int nb = (nF + gm.path_block - 1) / gm.path_block;
        const int li = atomicAdd(&counters[0], 1), j0 = atomicAdd(&counters[1], nb);
        PbLongUnit lu; lu.unit = u; lu.job_off = j0; lu.n_blocks = nb; lu.terminal = 0;
        longs[li] = lu;
        for (int b = 0; b < nb; b++) jobs[j0 + b] = make_int2(li, b);
    }
}

// (a) rows of the transfer matrices: T[job][e][c], 16 x 16 doubles per block
__global__ void __launch_bounds__(PB_PATHL_WARPS * 32)
pb_path_block_entry_kernel(const PbUnitDev* __restrict__ units, PbPitchGeomDev gm, const float* __restrict__ cand_f,
                           const float* __restrict__ cand_s, const uint8_t* __restrict__ ncand, const float* __restrict__ intensity,
                           const PbLongUnit* __restrict__ longs, const int2* __restrict__ jobs, const int* __restrict__ counters, double* __restrict__ T) {
    const int lane = threadIdx.x & 31, c = lane & 15, hb = lane & 16;
    const int maxc = gm.max_cand;
    const PbPathConsts k = pb_path_consts(gm);
    const int n_jobs = counters[1];
    const long long n_warps = (long long)gridDim.x * PB_PATHL_WARPS;
    for (long long wi = (long long)blockIdx.x * PB_PATHL_WARPS + (threadIdx.x >> 5); wi < 8LL * n_jobs; wi += n_warps) {
        const int job = (int)(wi >> 3), ep = (int)(wi & 7);
        const int2 jb = jobs[job];
        const PbUnitDev* ud = units + longs[jb.x].unit;
        const int nF = ud->n_frames; const int64_t f0 = ud->frame_off;
        const int fb = jb.y * gm.path_block, fe = min(nF, fb + gm.path_block);
        PbPathState st; st.delta = 0.0; st.l2 = 0.0; st.voiced = 0; st.nc = 0;
        int e = 2 * ep + (hb >> 4);
        bool row_ok;
        if (jb.y == 0) {
            if (ep != 0) continue;
            row_ok = hb == 0;                                    // the natural start: one row, the scores themselves
        } else {
            const int ncp = ncand[f0 + fb - 1];
            if (2 * ep >= ncp) continue;
            row_ok = e < ncp;
            if (!row_ok) e = 2 * ep;                             // odd candidate count: the upper half-warp repeats the lower one's row
            const float cfp = c < ncp ? cand_f[(f0 + fb - 1) * maxc + c] : 0.0f;
            pb_path_cand_props(k, cfp, st.voiced, st.l2);
            st.nc = ncp;
            st.delta = c == e ? 0.0 : PB_NEG_INF_D;
        }
        pb_path_forward<false>(k, maxc, true, cand_f, cand_s, ncand, intensity, f0, fb, fe, c, hb, lane, st, nullptr, nullptr, nullptr);
        if (row_ok) T[((size_t)job * 16 + e) * 16 + c] = c < st.nc ? st.delta : PB_NEG_INF_D;
    }
}

// (b) D_b = D_{b-1} (x) T_b in order; D[job][c] = score of candidate c of the block's last frame.  The matrices (2 KB each,
// independent of the running scores) are pulled into a ring of shared-memory slots by bulk copies issued PB_PATHL_RING blocks ahead,
// so a step is the sixteen shuffle / add / compare triples and not an L2 round trip.
#define PB_PATHL_RING 4
__global__ void __launch_bounds__(32)
pb_path_block_scan_kernel(const PbUnitDev* __restrict__ units, PbPitchGeomDev gm, const uint8_t* __restrict__ ncand,
                          PbLongUnit* __restrict__ longs, const int* __restrict__ counters, const double* __restrict__ T, double* __restrict__ D) {
    __shared__ __align__(16) double s_T[PB_PATHL_RING][256];
    __shared__ pbMbar s_bar[PB_PATHL_RING];
    const int lane = threadIdx.x & 31, c = lane & 15;
    const double NEG_INF = PB_NEG_INF_D;
    if (lane == 0) for (int i = 0; i < PB_PATHL_RING; i++) pb_mbar_init(&s_bar[i], 1);
    pb_mbar_init_fence();
    __syncwarp();
    unsigned parity = 0;                                         // bit i: the phase slot i is waited on next
    for (int li = blockIdx.x; li < counters[0]; li += gridDim.x) {
        const PbLongUnit lu = longs[li];
        const PbUnitDev* ud = units + lu.unit;
        const int nF = ud->n_frames; const int64_t f0 = ud->frame_off;
        const int nb = lu.n_blocks;
        auto issue = [&](int b) {                                // block b -> slot b % RING (lane 0)
            pbMbar* bar = &s_bar[b % PB_PATHL_RING];
            pb_mbar_expect_tx(bar, 2048u);
            pb_bulk_g2s(s_T[b % PB_PATHL_RING], T + ((size_t)lu.job_off + b) * 256, 2048u, bar);
        };
        if (lane == 0) for (int b = 1; b < nb && b <= PB_PATHL_RING; b++) issue(b);
        double d = T[(size_t)lu.job_off * 256 + c];              // block 0: row 0
        if (lane < 16) D[(size_t)lu.job_off * 16 + c] = d;
        int ncp_next = nb > 1 ? ncand[f0 + gm.path_block - 1] : 0;
        for (int b = 1; b < nb; b++) {
            const size_t job = (size_t)lu.job_off + b;
            const int slot = b % PB_PATHL_RING;
            const int ncp = ncp_next;
            if (b + 1 < nb) ncp_next = ncand[f0 + (int64_t)(b + 1) * gm.path_block - 1];
            pb_mbar_wait(&s_bar[slot], (parity >> slot) & 1u); parity ^= 1u << slot;
            __syncwarp();
            const double* t = s_T[slot] + c;
            double best = NEG_INF;
            PB_UNROLL for (int e = 0; e < 16; e++) {
                const double v = pb_shfl_d(d, e) + t[e * 16];
                if (e < ncp && v > best) best = v;
            }
            d = best;
            if (lane < 16) D[job * 16 + c] = d;
            __syncwarp();                                        // every lane has read the slot before it is refilled
            if (lane == 0 && b + PB_PATHL_RING < nb) issue(b + PB_PATHL_RING);
        }
        // terminal candidate: first maximum
        const int ncl = ncand[f0 + nF - 1];
        double bv = lane < ncl ? d : -1.0e300; int bi = lane;
        PB_UNROLL for (int o = 16; o > 0; o >>= 1) {
            const double ov = pb_shfl_d(bv, lane ^ o); const int oi = __shfl_xor_sync(PB_FULL_MASK, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) longs[li].terminal = bi;
    }
}

// (c) the sequential recursion inside every block from its true entering scores; back-pointers out, and the block's back-map
__global__ void __launch_bounds__(PB_PATHL_WARPS * 32)
pb_path_block_final_kernel(const PbUnitDev* __restrict__ units, PbPitchGeomDev gm, const float* __restrict__ cand_f,
                           const float* __restrict__ cand_s, const uint8_t* __restrict__ ncand, const float* __restrict__ intensity,
                           const PbLongUnit* __restrict__ longs, const int2* __restrict__ jobs, const int* __restrict__ counters,
                           const double* __restrict__ D, uint8_t* __restrict__ psi, unsigned long long* __restrict__ maps) {
    __shared__ unsigned long long s_psi[PB_PATHL_WARPS][PB_PATHL_BLOCK_MAX];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int maxc = gm.max_cand;
    unsigned long long* __restrict__ psi64 = reinterpret_cast<unsigned long long*>(psi);
    const PbPathConsts k = pb_path_consts(gm);
    const int n_jobs = counters[1];
    for (int job = blockIdx.x * PB_PATHL_WARPS + w; job < n_jobs; job += gridDim.x * PB_PATHL_WARPS) {
        const int2 jb = jobs[job];
        const PbUnitDev* ud = units + longs[jb.x].unit;
        const int nF = ud->n_frames; const int64_t f0 = ud->frame_off;
        const int fb = jb.y * gm.path_block, fe = min(nF, fb + gm.path_block);
        PbPathState st; st.delta = 0.0; st.l2 = 0.0; st.voiced = 0; st.nc = 0;
        if (jb.y > 0) {
            const int ncp = ncand[f0 + fb - 1];
            const float cfp = lane < ncp ? cand_f[(f0 + fb - 1) * maxc + lane] : 0.0f;
            pb_path_cand_props(k, cfp, st.voiced, st.l2);
            st.nc = ncp;
            st.delta = lane < 16 ? D[((size_t)job - 1) * 16 + lane] : 0.0;
        }
        __syncwarp();
        pb_path_forward<true>(k, maxc, true, cand_f, cand_s, ncand, intensity, f0, fb, fe, lane, 0, lane, st, psi64, nullptr, s_psi[w]);
        __syncwarp();
        // leaves the block through candidate `lane` -> candidate of frame fb - 1 (every lane reads the same word: a broadcast)
        int place = lane & 15;
        for (int f = fe - 1; f >= fb; f--) place = (int)((s_psi[w][f - fb] >> (4 * place)) & 15ull);
        const unsigned lo = __reduce_or_sync(PB_FULL_MASK, lane < 8 ? (unsigned)place << (4 * lane) : 0u);
        const unsigned hi = __reduce_or_sync(PB_FULL_MASK, (lane >= 8 && lane < 16) ? (unsigned)place << (4 * (lane - 8)) : 0u);
        if (lane == 0) maps[job] = ((unsigned long long)hi << 32) | lo;
        __syncwarp();
    }
}

// (d) the candidate every block is left through, newest block first
__global__ void __launch_bounds__(32)
pb_path_block_link_kernel(const PbLongUnit* __restrict__ longs, const int* __restrict__ counters, const unsigned long long* __restrict__ maps,
                          uint8_t* __restrict__ exits) {
    __shared__ unsigned long long s_map[1024];
    __shared__ uint8_t s_exit[1024];
    const int lane = threadIdx.x & 31;
    for (int li = blockIdx.x; li < counters[0]; li += gridDim.x) {
        const PbLongUnit lu = longs[li];
        int place = lu.terminal;
        for (int c1 = lu.n_blocks; c1 > 0; c1 -= 1024) {
            const int c0 = c1 > 1024 ? c1 - 1024 : 0;
            for (int b = c0 + lane; b < c1; b += 32) s_map[b - c0] = maps[(size_t)lu.job_off + b];
            __syncwarp();
            if (lane == 0) for (int b = c1 - 1; b >= c0; b--) { s_exit[b - c0] = (uint8_t)place; place = (int)((s_map[b - c0] >> (4 * place)) & 15ull); }
            place = __shfl_sync(PB_FULL_MASK, place, 0);
            __syncwarp();
            for (int b = c0 + lane; b < c1; b += 32) exits[(size_t)lu.job_off + b] = s_exit[b - c0];
            __syncwarp();
        }
    }
}

// (e) backtrack inside every block from the candidate it is left through; selected_array
__global__ void __launch_bounds__(PB_PATHL_WARPS * 32)
pb_path_block_select_kernel(const PbUnitDev* __restrict__ units, PbPitchGeomDev gm, const float* __restrict__ cand_f, const float* __restrict__ cand_s,
                            const PbLongUnit* __restrict__ longs, const int2* __restrict__ jobs, const int* __restrict__ counters,
                            const uint8_t* __restrict__ psi, const uint8_t* __restrict__ exits, float* __restrict__ sel_f, float* __restrict__ sel_s) {
    __shared__ unsigned long long s_psi[PB_PATHL_WARPS][PB_PATHL_BLOCK_MAX];
    __shared__ uint8_t s_pl[PB_PATHL_WARPS][PB_PATHL_BLOCK_MAX];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int maxc = gm.max_cand;
    const unsigned long long* __restrict__ psi64 = reinterpret_cast<const unsigned long long*>(psi);
    const int n_jobs = counters[1];
    for (int job = blockIdx.x * PB_PATHL_WARPS + w; job < n_jobs; job += gridDim.x * PB_PATHL_WARPS) {
        const int2 jb = jobs[job];
        const PbUnitDev* ud = units + longs[jb.x].unit;
        const int nF = ud->n_frames; const int64_t f0 = ud->frame_off;
        const int fb = jb.y * gm.path_block, fe = min(nF, fb + gm.path_block);
        for (int f = fb + lane; f < fe; f += 32) s_psi[w][f - fb] = f > 0 ? psi64[f0 + f] : 0ull;
        __syncwarp();
        if (lane == 0) {
            int place = exits[job];
            for (int f = fe - 1; f >= fb; f--) { s_pl[w][f - fb] = (uint8_t)place; place = (int)((s_psi[w][f - fb] >> (4 * place)) & 15ull); }
        }
        __syncwarp();
        for (int f = fb + lane; f < fe; f += 32) {
            const int64_t fr = f0 + f;
            const int pl = s_pl[w][f - fb];
            sel_f[fr] = cand_f[fr * maxc + pl]; sel_s[fr] = cand_s[fr * maxc + pl];
        }
        __syncwarp();
    }
}

// (f) median of the voiced frequencies of every long unit, one CTA each
__global__ void __launch_bounds__(1024)
pb_path_median_long_kernel(const PbUnitDev* __restrict__ units, const PbLongUnit* __restrict__ longs, const int* __restrict__ counters,
                           const float* __restrict__ sel_f, double* __restrict__ median_out, int32_t* __restrict__ nvoiced_out) {
    __shared__ int scratch[2];
    for (int li = blockIdx.x; li < counters[0]; li += gridDim.x) {
        const PbUnitDev* ud = units + longs[li].unit;
        int nv = 0;
        const double med = pb_median_positive<true>(sel_f + ud->frame_off, ud->n_frames, threadIdx.x, blockDim.x, scratch, nv);
        if (threadIdx.x == 0) { median_out[ud->out_index] = med; nvoiced_out[ud->out_index] = nv; }
        __syncthreads();
    }
}
