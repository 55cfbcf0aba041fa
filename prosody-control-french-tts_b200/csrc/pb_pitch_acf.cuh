// pb_pitch_acf.cuh — K1: frames -> normalised autocorrelation, written to a global scratch for K2 (pb_pitch_cand.cuh).
//
// Follows the first half of Sound_into_PitchFrame (Praat fon/Sound_to_Pitch.cpp, AC_HANNING): local mean, Hanning window,
// autocorrelation by FFT, division by the window's autocorrelation.  See pb_pitch.cuh for the layout of the F0 path.
//
// Round 1 ran this and the candidate search as ONE kernel (pb_pitch_frames_kernel): its loop body was twice the 32 KB
// instruction cache, its 128 registers (the radix-32 butterflies) capped residency at 4 warps per scheduler for the
// latency-bound candidate search as well, and 40 % of its executed instructions were integer / control overhead
// (profiles/r01_frames_v6_*).  Split in two, K1 is a compact FFT kernel that fits the instruction cache and K2 runs at
// 3x the residency; r goes through HBM once (1.3 KB per frame at 16 kHz / 75 Hz), far below what the memory system can
// absorb beside the arithmetic (profiles/r02_*).
//
// Per work item a GROUP of G warps (G = 1 for FFT sizes <= 1024) owns a PAIR of consecutive frames:
//   * the samples both frames read are staged by ONE cp.async.bulk (TMA, pb_async.cuh) issued as soon as the previous
//     pair has been windowed; they land while that pair is transformed;
//   * both real frames are windowed into one complex sequence a + i b and transformed together; the two power spectra
//     are separated with the conjugate-symmetry identity, packed again and transformed back (the spectra are real and
//     even, so a forward transform returns both autocorrelations);
//   * r[lag] = ac[lag] / (ac[0] windowR[lag]) for lags 0..B+1 of both frames goes to racf[slot], slot = 2 * item + frame.
// Everything index-like is hoisted: shared-memory addresses of the FFT passes are one base register plus compile-time
// offsets, the pass-specific stores are two straight-line variants instead of one with a runtime stride.
#pragma once
#include "pb_async.cuh"
#include "pb_pitch.cuh"

// How interior frame pairs are windowed: 1 = into the shared FFT buffer by a separate loop (conflict-free rows n, n + GT), scale
// folded in; 2 = in registers, inside the first FFT pass's loads (no windowed copy in shared memory).  Mode 2 moves 20 % fewer
// shared-memory wavefronts but its unrolled, predicated first pass makes the loop body miss the instruction cache and exposes the
// window-table loads: 5.8 ms against 5.0 ms for mode 1 on 2 M frames (profiles/r02_acf_modes.txt), so mode 1 ships.
#ifndef PB_K1_MODE
#define PB_K1_MODE 1
#endif

// Where a frame pair's samples sit: part index of frame A's sample 0, distance to frame B, and the offset of frame A's
// sample `span_lo` inside the staged (16-byte aligned) range.
struct PbPairPos { int start0; int hop; int shift; int edge; };

// Frame position, Praat's Sampled_indexToX / Sampled_xToLowIndex in float64 with explicit rounding (no FMA contraction):
// returns the 1-based part index of frame sample n = 0.
__device__ __forceinline__ long long pb_frame_start(double t1, double x1, int frame, const PbPitchGeomDev& gm) {
    const double t = __dadd_rn(t1, __dmul_rn((double)frame, gm.dt));
    const long long left = (long long)floor(__ddiv_rn(__dsub_rn(t, x1), gm.dx)) + 1;
    return left + 1 - gm.half_nw;
}

// Everything K1 needs to stage a pair, computed once by a small kernel ahead of it (float64 frame positions, 64-bit address
// arithmetic): {start0, hop, packed, src16}.  packed = shift | edge << 4 | dst16 << 8 | n16 << 16: the bulk copy moves n16
// 16-byte chunks from pcm_base16 + 16 * src16 to the staging buffer + 16 * dst16; `edge` marks the pairs whose staged range is not
// wholly inside the pcm buffer (first / last file of a call) or does not fit the packing: K1 stages those the long way.
__device__ __forceinline__ int pb_span_lo(const PbPitchGeomDev& gm) { const int m = gm.half_nw - gm.nsamp_period; return m < 0 ? m : 0; }
__device__ __forceinline__ int pb_span_hi(const PbPitchGeomDev& gm) { const int m = gm.half_nw + gm.nsamp_period; return m > gm.nw ? m : gm.nw; }

__global__ void __launch_bounds__(256)
pb_pair_pos_kernel(const int16_t* __restrict__ pcm, const PbUnitDev* __restrict__ units, const int32_t* __restrict__ pair_off, PbPitchGeomDev gm, int4* __restrict__ out) {
    const int span_lo = pb_span_lo(gm), span_len = pb_span_hi(gm) - span_lo;
    const long long buf_lo = (long long)(size_t)pcm, buf_hi = buf_lo + 2 * gm.pcm_len;
    const long long base16 = buf_lo & ~15LL, lo16 = (buf_lo + 15) & ~15LL, hi16 = buf_hi & ~15LL;
    for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < gm.n_pairs; item += gridDim.x * blockDim.x) {
        const int u = pb_upper_unit(pair_off, gm.n_units, item);
        const PbUnitDev* ud = units + u;
        const int fA = 2 * (item - ud->pair_off);
        const long long s0 = pb_frame_start(ud->t1, ud->x1, fA, gm);
        const int hop = fA + 1 < ud->n_frames ? (int)(pb_frame_start(ud->t1, ud->x1, fA + 1, gm) - s0) : 0;
        // frame A sample n is part sample start0+n = pcm sample pcm_off + ix1 + start0 + n - 2
        const long long byte0 = buf_lo + 2 * (ud->pcm_off + ud->ix1 + s0 - 2 + span_lo);
        const long long a0 = byte0 & ~15LL;
        const int shift = (int)((byte0 - a0) >> 1);
        const long long a1 = a0 + 16LL * ((shift + span_len + hop + 7) >> 3);
        const long long src16 = (a0 - base16) >> 4, n16 = (a1 - a0) >> 4;
        const bool edge = a0 < lo16 || a1 > hi16 || src16 < 0 || src16 > 0x7fffffffLL || n16 > 0xffff;
        out[item] = make_int4((int)s0, hop, shift | (edge ? 16 : 0) | (int)((edge ? 0 : n16) << 16), (int)(edge ? 0 : src16));
    }
}

// The long way (edge pairs only, out of line).  Stage the samples the pair (frames fA, fA+1 of unit ud) will read into `dst`: every 16-byte chunk that lies entirely
// inside the pcm buffer goes in ONE bulk copy issued by thread 0 of the group (completion on `bar`); the (at most two)
// chunks straddling the ends of the buffer are filled sample by sample and flagged in pos.edge so that the group
// synchronises before reading them.  Values outside the unit's part / file are masked at read time.
template <int GT>
__device__ PB_NOINLINE PbPairPos pb_stage_pair_edge(const int16_t* __restrict__ pcm, const PbUnitDev& ud, int2 sp, long long pcm_len,
                                                   int span_lo, int span_len, int16_t* dst, int g, pbMbar* bar) {
    PbPairPos pos;
    pos.start0 = sp.x; pos.hop = sp.y; pos.edge = 0;
    // frame A sample n is part sample start0+n = pcm sample pcm_off + ix1 + start0 + n - 2
    const long long gs = ud.pcm_off + ud.ix1 + (long long)pos.start0 - 2 + span_lo;
    const long long byte0 = (long long)(size_t)pcm + 2 * gs;            // may lie outside the buffer: never dereferenced there
    const long long a0 = byte0 & ~15LL;
    pos.shift = (int)((byte0 - a0) >> 1);
    const long long a1 = a0 + 16LL * ((pos.shift + span_len + pos.hop + 7) >> 3);      // end of the staged range
    const long long buf_lo = (long long)(size_t)pcm, buf_hi = buf_lo + 2 * pcm_len;
    const long long lo16 = (buf_lo + 15) & ~15LL, hi16 = buf_hi & ~15LL;                // whole chunks inside the buffer
    const long long c0 = a0 > lo16 ? a0 : lo16, c1 = a1 < hi16 ? a1 : hi16;
    if (g == 0) {
        const unsigned bytes = c1 > c0 ? (unsigned)(c1 - c0) : 0u;
        pb_mbar_expect_tx(bar, bytes);
        if (bytes) pb_bulk_g2s((char*)dst + (c0 - a0), (const void*)(size_t)c0, bytes, bar);
    }
    if (a0 < c0 || a1 > c1) {
        // chunks that are not wholly inside the buffer (first / last file of the call only)
        pos.edge = 1;
        const long long e0 = c1 > c0 ? c0 : a1, e1 = c1 > c0 ? c1 : a1;    // [a0, e0) and [e1, a1) are filled here
        for (long long b = a0 + 2 * g; b < a1; b += 2 * GT) {
            if (b >= e0 && b < e1) continue;
            *(int16_t*)((char*)dst + (b - a0)) = (b >= buf_lo && b + 2 <= buf_hi) ? *(const int16_t*)(size_t)b : (int16_t)0;
        }
    }
    return pos;
}

// The short way: the descriptor already holds the copy; thread 0 of the group issues it.
template <int GT>
__device__ __forceinline__ PbPairPos pb_stage_pair(const int16_t* __restrict__ pcm, const PbUnitDev* __restrict__ units, int u, int4 d, long long pcm_len,
                                                   int span_lo, int span_len, int16_t* dst, int g, pbMbar* bar) {
    if (d.z & 16) return pb_stage_pair_edge<GT>(pcm, units[u], make_int2(d.x, d.y), pcm_len, span_lo, span_len, dst, g, bar);
    PbPairPos pos;
    pos.start0 = d.x; pos.hop = d.y; pos.shift = d.z & 15; pos.edge = 0;
    if (g == 0) {
        const unsigned bytes = ((unsigned)d.z >> 16) << 4;
        pb_mbar_expect_tx(bar, bytes);
        pb_bulk_g2s(dst, (const char*)((size_t)pcm & ~(size_t)15) + ((size_t)(unsigned)d.w << 4), bytes, bar);
    }
    return pos;
}

// ------------------------------------------------------------------------------------------------ K1
// MINB: resident CTAs per SM the register allocation targets; WSYNC: one-warp groups synchronise with __syncwarp() instead of a
// named barrier (both chosen at launch: PB_ACF_CTAS / PB_ACF_WSYNC, defaults from the measurements in DESIGN.md)
template <int LOG2N, int MINB, bool WSYNC>
__global__ void __launch_bounds__(PbFftCfg<LOG2N>::WARPS_PER_CTA * 32, MINB)
pb_pitch_acf_kernel(const int16_t* __restrict__ pcm, const PbUnitDev* __restrict__ units, const int32_t* __restrict__ pair_off,
                    const int4* __restrict__ pairpos, PbPitchGeomDev gm, int item0, int n_items, int rstride_g,
                    float* __restrict__ racf, long long* __restrict__ slot_fr,
                    float* __restrict__ cand_f, float* __restrict__ cand_s, uint8_t* __restrict__ ncand, float* __restrict__ intensity) {
    typedef PbFftCfg<LOG2N> C;
    constexpr int R = C::R, LR = C::LR, N = C::N, G = C::G, GT = C::GT;
    constexpr bool FAST = (R == 32);                    // compile-time shared-memory offsets (all sizes >= 1024)
    constexpr bool REG_OUT = FAST && G == 1 && C::F == 1 && C::RPL <= R;   // N = 1024: r leaves from the registers of the last pass
    PB_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = warp / G, wg = warp % G;          // group in CTA, warp in group
    const int g = wg * 32 + lane;                       // thread in group = butterfly index
    const int bar_id = 1 + group;
#define PB_K1_SYNC() do { if (WSYNC && G == 1) __syncwarp(); else pb_group_sync<G>(bar_id); } while (0)
    // per-group shared memory: FFT buffer, a small reduction scratch, the sample staging buffer; CTA-wide: one mbarrier per group
    const size_t group_bytes = (size_t)(C::BUF + 8 * G) * sizeof(float2) + (size_t)gm.pre_cap * sizeof(int16_t);
    unsigned char* gbase = smem_raw + (size_t)group * group_bytes;
    float2* buf = (float2*)gbase;
    float* red = (float*)(buf + C::BUF);                // [G][4] floats
    int16_t* pre = (int16_t*)(buf + C::BUF + 8 * G);    // [pre_cap] staged samples of one pair
    pbMbar* mbar = (pbMbar*)(smem_raw + (size_t)C::GROUPS_PER_CTA * group_bytes) + group;
    if (g == 0) pb_mbar_init(mbar, 1);
    pb_mbar_init_fence();
    __syncthreads();
    unsigned phase = 0;
    const int B = gm.brent_ixmax;
    const int nw = gm.nw;
    const int pk_lo = max(0, gm.half_nw - gm.half_period), pk_n = min(nw, gm.half_nw + gm.half_period) - pk_lo;   // [lo, lo + n)
    const int mean_n0 = gm.half_nw - gm.nsamp_period;   // local mean spans frame samples [mean_n0, mean_n0 + 2 P)
    const int mean_len = 2 * gm.nsamp_period;
    const int span_lo = min(0, mean_n0);                // staged range of frame samples: [span_lo, span_lo + span_len)
    const int span_hi = max(nw, mean_n0 + mean_len);
    const int span_len = span_hi - span_lo;
    const float mean_scale = (float)(1.0 / (32768.0 * (double)mean_len));
    const bool fuse_ok = mean_n0 >= 0 && mean_n0 + mean_len <= nw;     // the local-mean span lies inside the window (periods_per_window >= 2)
    // per thread, loop-invariant: which of its R first-pass inputs n = g + GT t lie inside the window / inside the local-peak span
    unsigned vmask = 0, pkmask = 0;
    PB_UNROLL for (int t = 0; t < R; t++) {
        const int n = g + t * GT;
        if (n < nw) vmask |= 1u << t;
        if ((unsigned)(n - pk_lo) < (unsigned)pk_n) pkmask |= 1u << t;
    }

    // Blocked distribution: a group walks a contiguous range of frame pairs, so consecutive iterations stay in the same
    // unit and the next pair's samples are staged while the current pair is transformed.
    const int n_groups = gridDim.x * C::GROUPS_PER_CTA;
    const int per_group = (n_items + n_groups - 1) / n_groups;
    const int it_begin = (blockIdx.x * C::GROUPS_PER_CTA + group) * per_group;
    const int it_end = min(n_items, it_begin + per_group);
    if (it_begin >= it_end) return;
    int u_next = pb_upper_unit(pair_off, gm.n_units, item0 + it_begin);
    int u_next_end = pair_off[u_next + 1];
    int u = -1;
    PbPairPos pos_next = pb_stage_pair<GT>(pcm, units, u_next, pairpos[item0 + it_begin], gm.pcm_len, span_lo, span_len, pre, g, mbar);
    // what the loop needs of the current unit, refreshed when the unit changes (the 80-byte descriptor itself stays in L1 / L2):
    // first pair, frame count, first frame, global peak, and the part indices that hold samples [pmin, pmax1)
    int u_pair_off = 0, u_nframes = 0, u_pmin = 0, u_pmax1 = 0;
    long long u_frame_off = 0;
    float gpk = 0.0f;

    for (int li = it_begin; li < it_end; li++) {
        const int item = item0 + li;
        const PbPairPos pos = pos_next;
        // the next pair's descriptor is fetched now and used after the first pass's loads: its latency hides behind the windowing
        // (fetched where it is used, the `edge` test on it was the kernel's single hottest stall: profiles/r02b_acf_ncu_summary.csv)
        const int4 d_next = li + 1 < it_end ? __ldg(pairpos + item + 1) : make_int4(0, 0, 0, 0);
        if (u != u_next) {
            u = u_next;
            const PbUnitDev* up = units + u;
            u_pair_off = up->pair_off; u_nframes = up->n_frames; u_frame_off = up->frame_off; gpk = (float)up->global_peak;
            const long long ix1 = up->ix1, e = (long long)up->file_nx - ix1 + 1;
            const long long pmin = 2 - ix1 > 1 ? 2 - ix1 : 1, pmax1 = (up->nx < e ? up->nx : e) + 1;
            const long long BIG = 1LL << 30;                    // frame positions are ints: saturate
            u_pmin = (int)(pmin > BIG ? BIG : pmin); u_pmax1 = (int)(pmax1 > BIG ? BIG : (pmax1 < -BIG ? -BIG : pmax1));
        }
        const int fA = 2 * (item - u_pair_off);
        const bool hasB = fA + 1 < u_nframes;
        const bool global_silent = gpk == 0.0f;
        const long long frA = u_frame_off + fA;
        pb_mbar_wait(mbar, phase); phase ^= 1u;          // this pair's samples (requested during the previous pair) have landed
#ifdef PB_SIMT_EMU
        PB_K1_SYNC();
#else
        if (pos.edge) PB_K1_SYNC();          // edge chunks were filled with ordinary stores
#endif
        const int16_t* sm = pre;
        const int sb0 = pos.shift - span_lo, sb1 = sb0 + pos.hop;       // staged index of frame sample n is sb + n
        float pkA = 0.0f, pkB = 0.0f;       // local peaks (Praat: max |windowed sample| around the frame centre), in SCALED units
        float sA = 1.0f, sB = 1.0f;         // power-of-two scales of the two frames
        float2 q15s = make_float2(0.0f, 0.0f), nms = make_float2(0.0f, 0.0f);     // fused windowing: value = (sample * q15s + nms) * window
        bool any_signal = false;
        // valid frame samples n: [nlo, nhi)
        const long long loA = (long long)u_pmin - pos.start0, hiA = (long long)u_pmax1 - pos.start0;
        // interior pair: every sample both frames touch exists.  Then (FAST sizes) nothing is windowed into shared memory at all:
        // the first FFT pass reads the staged samples directly and windows them in registers (below); here only the exact
        // integer statistics of the two frames are taken: sum over the local-mean span, minimum and maximum over the window.
        const bool interior = FAST && fuse_ok && hasB && loA <= span_lo && hiA - pos.hop >= span_hi;      // frame B's range is frame A's shifted by hop
        const bool fused = PB_K1_MODE == 2 && interior;
        if (interior) {
            // Both statistics come from ONE sweep over the local-mean span (the central two thirds of the window, where the
            // Hanning window is large: its extremes decide the frame's magnitude for the purpose of the scale), two samples per
            // 32-bit load: dp2a adds both halves exactly, the packed min / max keep both.  The span's first / last sample may sit in
            // a word that is only half inside: lane 0 adds those two samples on their own.
            int s0, s1, mn0, mx0, mn1, mx1;
            {
                int sum[2]; unsigned vmn[2], vmx[2];
                PB_UNROLL for (int f = 0; f < 2; f++) {
                    const int M0 = (f ? sb1 : sb0) + mean_n0, M1 = M0 + mean_len;      // staged sample range of the span
                    const int k0 = (M0 + 1) >> 1, k1 = M1 >> 1;                         // whole words inside it
                    const unsigned* smw = reinterpret_cast<const unsigned*>(sm);
                    int acc = 0; unsigned lo = 0x7fff7fffu, hi = 0x80008000u;
                    for (int k = k0 + g; k < k1; k += GT) {
                        const unsigned w = smw[k];
                        acc = __dp2a_lo((int)w, 0x0101, acc);
                        lo = __vmins2(lo, w); hi = __vmaxs2(hi, w);
                    }
                    if (g == 0) {
                        if (M0 & 1) acc += (int)sm[M0];
                        if (M1 & 1) acc += (int)sm[M1 - 1];
                    }
                    sum[f] = acc; vmn[f] = lo; vmx[f] = hi;
                }
                s0 = sum[0]; s1 = sum[1];
                mn0 = min((int)(short)(vmn[0] & 0xffff), (int)(short)(vmn[0] >> 16)); mx0 = max((int)(short)(vmx[0] & 0xffff), (int)(short)(vmx[0] >> 16));
                mn1 = min((int)(short)(vmn[1] & 0xffff), (int)(short)(vmn[1] >> 16)); mx1 = max((int)(short)(vmx[1] & 0xffff), (int)(short)(vmx[1] >> 16));
            }
            s0 = __reduce_add_sync(PB_FULL_MASK, s0); s1 = __reduce_add_sync(PB_FULL_MASK, s1);
            mn0 = __reduce_min_sync(PB_FULL_MASK, mn0); mx0 = __reduce_max_sync(PB_FULL_MASK, mx0);
            mn1 = __reduce_min_sync(PB_FULL_MASK, mn1); mx1 = __reduce_max_sync(PB_FULL_MASK, mx1);
            if (G > 1) {
                int* redi = (int*)red;
                if (lane == 0) { redi[wg * 6 + 0] = s0; redi[wg * 6 + 1] = s1; redi[wg * 6 + 2] = mn0; redi[wg * 6 + 3] = mx0; redi[wg * 6 + 4] = mn1; redi[wg * 6 + 5] = mx1; }
                PB_K1_SYNC();
                s0 = 0; s1 = 0;
                for (int k = 0; k < G; k++) {
                    s0 += redi[k * 6 + 0]; s1 += redi[k * 6 + 1];
                    mn0 = min(mn0, redi[k * 6 + 2]); mx0 = max(mx0, redi[k * 6 + 3]); mn1 = min(mn1, redi[k * 6 + 4]); mx1 = max(mx1, redi[k * 6 + 5]);
                }
            }
            const float q15 = 1.0f / 32768.0f;
            const float meanA = (float)s0 * mean_scale, meanB = (float)s1 * mean_scale;
            // the larger distance of the span's extremes from the mean: the frame's magnitude to within the window's taper
            const float mA = fmaxf(fabsf((float)mx0 * q15 - meanA), fabsf((float)mn0 * q15 - meanA));
            const float mB = fmaxf(fabsf((float)mx1 * q15 - meanB), fabsf((float)mn1 * q15 - meanB));
            // bring both frames to comparable magnitude (power-of-two scales are exact and cancel in r = ac/ac[0]):
            // keeps the weaker frame of a pair out of the stronger one's rounding noise
            sA = mA > 0.0f ? __int_as_float((254 - ((__float_as_int(mA) >> 23) & 0xff)) << 23) : 1.0f;
            sB = mB > 0.0f ? __int_as_float((254 - ((__float_as_int(mB) >> 23) & 0xff)) << 23) : 1.0f;
            q15s = make_float2(q15 * sA, q15 * sB);
            nms = make_float2(-meanA * sA, -meanB * sB);
            any_signal = mA > 0.0f || mB > 0.0f;
            if (PB_K1_MODE != 2) {
                // ---- window both frames into the FFT buffer, z = a + i b (natural order, already scaled), rows n and n + GT per step
                const int16_t* pa = sm + sb0; const int16_t* pb = sm + sb1;
                for (int n = g; n < nw; n += 2 * GT) {
                    const float w0 = __ldg(gm.window + n);
                    const float2 x0 = __fmul2_rn(__ffma2_rn(make_float2((float)pa[n], (float)pb[n]), q15s, nms), make_float2(w0, w0));
                    if ((unsigned)(n - pk_lo) < (unsigned)pk_n) { pkA = fmaxf(pkA, fabsf(x0.x)); pkB = fmaxf(pkB, fabsf(x0.y)); }
                    buf[pb_pad5(n)] = x0;
                    const int n1 = n + GT;
                    if (n1 < nw) {
                        const float w1 = __ldg(gm.window + n1);
                        const float2 x1 = __fmul2_rn(__ffma2_rn(make_float2((float)pa[n1], (float)pb[n1]), q15s, nms), make_float2(w1, w1));
                        if ((unsigned)(n1 - pk_lo) < (unsigned)pk_n) { pkA = fmaxf(pkA, fabsf(x1.x)); pkB = fmaxf(pkB, fabsf(x1.y)); }
                        buf[pb_pad5(n1)] = x1;
                    }
                }
                for (int n = nw + g; n < N; n += GT) buf[pb_pad5(n)] = make_float2(0.0f, 0.0f);      // zero padding
                PB_K1_SYNC();                   // the windowed frames are in the buffer
            }
        } else {
            // ---- first / last frames of a slice that Praat zero-fills beyond the file, a unit with an odd frame count, or a small
            //      FFT geometry: window both frames into the FFT buffer, z = a + i b (natural order, zero padded)
            float mxA = 0.0f, mxB = 0.0f;
            int nlo[2], nhi[2], sb[2]; float lmean[2];
            PB_UNROLL for (int f = 0; f < 2; f++) {
                long long lo = loA - (f ? pos.hop : 0), hi = hiA - (f ? pos.hop : 0);
                const long long BIG = 1 << 30;
                nlo[f] = (int)(lo < -BIG ? -BIG : (lo > BIG ? BIG : lo));
                nhi[f] = (int)(hi < -BIG ? -BIG : (hi > BIG ? BIG : hi));
                sb[f] = f ? sb1 : sb0;
                int s = 0;
                for (int q = lane; q < mean_len; q += 32) {
                    const int n = mean_n0 + q;
                    s += (n >= nlo[f] && n < nhi[f]) ? (int)sm[sb[f] + n] : 0;
                }
                s = pb_warp_sum_i(s);
                lmean[f] = (float)s * mean_scale;
            }
            if (!hasB) { nlo[1] = 0; nhi[1] = 0; sb[1] = sb[0]; }
            const float2 nmean = make_float2(-lmean[0], -lmean[1]), q15 = make_float2(1.0f / 32768.0f, 1.0f / 32768.0f);
            const float hb = hasB ? 1.0f : 0.0f;
            for (int n = g; n < nw; n += GT) {
                const float w = __ldg(&gm.window[n]);
                const int sa = (n >= nlo[0] && n < nhi[0]) ? (int)sm[sb[0] + n] : 0;
                const int sbv = (n >= nlo[1] && n < nhi[1]) ? (int)sm[sb[1] + n] : 0;
                const float2 ab = __fmul2_rn(__ffma2_rn(make_float2((float)sa, (float)sbv), q15, nmean), make_float2(w, w * hb));
                const float aa = fabsf(ab.x), bb = fabsf(ab.y);
                mxA = fmaxf(mxA, aa); mxB = fmaxf(mxB, bb);
                if ((unsigned)(n - pk_lo) < (unsigned)pk_n) { pkA = fmaxf(pkA, aa); pkB = fmaxf(pkB, bb); }
                buf[pb_pad5(n)] = ab;
            }
            for (int n = nw + g; n < N; n += GT) buf[pb_pad5(n)] = make_float2(0.0f, 0.0f);      // zero padding
            mxA = pb_warp_max(mxA); mxB = pb_warp_max(mxB);
            if (G > 1) {
                if (lane == 0) { red[wg * 4 + 0] = mxA; red[wg * 4 + 1] = mxB; }
                PB_K1_SYNC();
                for (int k = 0; k < G; k++) { mxA = fmaxf(mxA, red[k * 4 + 0]); mxB = fmaxf(mxB, red[k * 4 + 1]); }
            }
            sA = mxA > 0.0f ? __int_as_float((254 - ((__float_as_int(mxA) >> 23) & 0xff)) << 23) : 1.0f;
            sB = mxB > 0.0f ? __int_as_float((254 - ((__float_as_int(mxB) >> 23) & 0xff)) << 23) : 1.0f;
            pkA *= sA; pkB *= sB;               // scaled units, like the fused path
            any_signal = mxA > 0.0f || mxB > 0.0f;
            PB_K1_SYNC();                       // the windowed frames are in the buffer
        }
        const bool active = !global_silent && any_signal;
        bool staged_next = false;
        // the next pair's samples are requested as soon as this pair's have been consumed (after the first pass's loads when
        // the windowing is fused into them): they land while this pair is transformed
#define PB_K1_STAGE_NEXT() do { \
            staged_next = true; \
            if (li + 1 < it_end) { \
                if (item + 1 >= u_next_end) { do { u_next++; u_next_end = pair_off[u_next + 1]; } while (item + 1 >= u_next_end); } \
                pos_next = pb_stage_pair<GT>(pcm, units, u_next, d_next, gm.pcm_len, span_lo, span_len, pre, g, mbar); \
            } } while (0)

        // ---- two FFTs x two passes through ONE copy of the butterfly code: FFT (step >> 1), pass (step & 1).  The step
        //      index is made opaque so the optimiser neither peels nor unswitches the loop (either duplicates the
        //      ~300-instruction network).
        float2 v[R];
        int n_steps = active ? 4 : 0;
        asm volatile("" : "+r"(n_steps));
#ifndef PB_SIMT_EMU
#pragma unroll 1
#endif
        for (int s_it = 0; s_it < n_steps; s_it++) {
            int step = s_it;
            asm volatile("" : "+r"(step));
            int pass = step & 1;
            asm volatile("" : "+r"(pass));
            // pass 1 reads natural order: the frames (step 0) or the spectra (step 2);
            // pass 2 reads the pass-1 layout and applies the inter-pass twiddles
            if (FAST && step == 0 && fused) {
                // frame sample n = g + GT t straight from the staging buffer: ((a, b) q15 s - mean s) w, zero beyond the window
                const int16_t* pa = sm + sb0 + g; const int16_t* pb = sm + sb1 + g;
                const float* win = gm.window + g;
                PB_UNROLL for (int t = 0; t < R; t++) {
                    float2 x = make_float2(0.0f, 0.0f);
                    if (vmask & (1u << t)) {
                        const float w = __ldg(win + t * GT);
                        x = __fmul2_rn(__ffma2_rn(make_float2((float)pa[t * GT], (float)pb[t * GT]), q15s, nms), make_float2(w, w));
                        if (pkmask & (1u << t)) { pkA = fmaxf(pkA, fabsf(x.x)); pkB = fmaxf(pkB, fabsf(x.y)); }
                    }
                    v[t] = x;
                }
            } else if (FAST) {
                // i = g + 32 G t  ->  i + (i >> 5) = g + (g >> 5) + 33 G t: one base, compile-time offsets
                const float2* src = buf + (g + (g >> 5));
                PB_UNROLL for (int t = 0; t < R; t++) v[t] = src[t * (33 * G)];
                if (step == 0 && !interior) { const float2 sc = make_float2(sA, sB); PB_UNROLL for (int t = 0; t < R; t++) v[t] = __fmul2_rn(v[t], sc); }
            } else {
                const int sh = pass ? LR : 5;
                PB_UNROLL for (int t = 0; t < R; t++) { const int i = g + t * (N / R); v[t] = buf[i + (i >> sh)]; }
                if (step == 0) { const float2 sc = make_float2(sA, sB); PB_UNROLL for (int t = 0; t < R; t++) v[t] = __fmul2_rn(v[t], sc); }
            }
            if (pass) {
                const float2* tw = gm.tw_a + (g & (R - 1));
                PB_UNROLL for (int t = 1; t < R; t++) {
                    const float2 w = __ldg(tw + t * R);
                    const float2 x = v[t];
                    v[t] = __ffma2_rn(make_float2(x.y, x.y), make_float2(-w.y, w.x), __fmul2_rn(make_float2(x.x, x.x), w));   // x * w
                }
            }
            PB_K1_SYNC();                   // every load of this step is done before any store
            if (step == 0) PB_K1_STAGE_NEXT();
            pb_dft<R>(v);
            // pass 1 (Ns = 1): out[g*R + t], skew (index >> LR);  pass 2 (Ns = R): out[(g/R) R^2 + g%R + t R], skew 5
            if (REG_OUT && step == 3) {
                // ---- the last pass of the second transform leaves lags g + 32 t in this lane's registers: the lags 0..B+1 go
                //      straight to the global scratch, r[lag] = ac[lag] / (ac[0] windowR[lag]); nothing returns to shared memory
                const float2 a0v = v[0];
                const float2 ac0 = make_float2(__shfl_sync(PB_FULL_MASK, a0v.x, 0), __shfl_sync(PB_FULL_MASK, a0v.y, 0));
                const float2 inv0 = make_float2(ac0.x > 0.0f ? 1.0f / ac0.x : 0.0f, ac0.y > 0.0f ? 1.0f / ac0.y : 0.0f);
                float* ra = racf + (size_t)(2 * li) * rstride_g + g;
                float* rb = ra + rstride_g;
                const float* iwp = gm.inv_wr + g;                  // inv_wr[B + 1] = 0
                PB_UNROLL for (int q = 0; q < C::RPL; q++) {
                    if (g + q * GT <= B + 1) {
                        const float iw = __ldg(iwp + q * GT);
                        float2 r2 = __fmul2_rn(v[pb_bitrev(q, LR)], __fmul2_rn(inv0, make_float2(iw, iw)));
                        if (q == 0 && g == 0) r2 = make_float2(1.0f, 1.0f);
                        ra[q * GT] = r2.x; rb[q * GT] = r2.y;
                    }
                }
            } else if (FAST) {
                // pass 1: 32 g + t -> 33 g + t;  pass 2: (g>>5) 1024 + (g&31) + 32 t -> (g>>5) 1056 + (g&31) + 33 t
                if (pass) {
                    float2* dst = buf + ((g >> 5) * 1056 + (g & 31));
                    PB_UNROLL for (int t = 0; t < R; t++) dst[t * 33] = v[pb_bitrev(t, LR)];
                } else {
                    float2* dst = buf + 33 * g;
                    PB_UNROLL for (int t = 0; t < R; t++) dst[t] = v[pb_bitrev(t, LR)];
                }
            } else {
                const int ob = pass ? ((g >> LR) * (R * R) + (g & (R - 1))) : g * R;
                const int os = pass ? R : 1;
                const int osh = pass ? 5 : LR;
                PB_UNROLL for (int t = 0; t < R; t++) { const int o = ob + t * os; buf[o + (o >> osh)] = v[pb_bitrev(t, LR)]; }
            }
            if (!(REG_OUT && step == 3)) PB_K1_SYNC();
            if (pass && C::F > 1) {
                // ---- final pass (radix F, Ns = R*R): butterflies are in place
                constexpr int F = C::F > 1 ? C::F : 2, LF = pb_ilog2(F);
                PB_UNROLL for (int b = 0; b < C::FB; b++) {
                    const int j = g + b * GT;             // 0 .. N/F-1 = R*R-1
                    float2 a[F];
                    PB_UNROLL for (int t = 0; t < F; t++) a[t] = buf[pb_pad5(j + t * (R * R))];
                    PB_UNROLL for (int t = 1; t < F; t++) {
                        const float2 w = __ldg(&gm.tw_b[t * (R * R) + j]);
                        const float2 x = a[t];
                        a[t] = make_float2(x.x * w.x - x.y * w.y, x.x * w.y + x.y * w.x);
                    }
                    pb_dft<F>(a);
                    PB_UNROLL for (int t = 0; t < F; t++) buf[pb_pad5(j + t * (R * R))] = a[pb_bitrev(t, LF)];
                }
                PB_K1_SYNC();
            }
            if (step == 1) {
                // ---- power spectra of both frames from Z = FFT(a + i b):  4 P_a = |Z_k + conj Z_-k|^2,  4 P_b = |Z_k - conj Z_-k|^2,
                //      packed again as P_a + i P_b (real and even: written to k and N - k)
                if (FAST && G == 1) {
                    // k = g + 32 i -> slot g + 33 i;  N - k = 32 (31 - i) + (32 - g) -> slot (32 - g) + 33 (31 - i)   (g > 0)
                    // lane 0: N - k = 32 (32 - i) -> slot 33 (32 - i); k = 0 pairs with itself
                    const float2* pk_ = buf + g;
                    float2* qk = buf + (g ? 32 - g : 33);
                    PB_UNROLL for (int i = 0; i < 16; i++) {
                        const float2 za = pk_[33 * i];
                        float2 zb = qk[33 * (31 - i)];
                        if (i == 0 && g == 0) zb = za;
                        const float2 p = make_float2(za.x + zb.x, za.y - zb.y), q = make_float2(za.x - zb.x, za.y + zb.y);
                        const float2 w = make_float2(p.x * p.x + p.y * p.y, q.x * q.x + q.y * q.y);
                        ((float2*)pk_)[33 * i] = w;
                        if (!(i == 0 && g == 0)) qk[33 * (31 - i)] = w;
                    }
                    if (g == 0) {                          // k = 512 pairs with itself
                        const float2 za = buf[pb_pad5(N / 2)];
                        buf[pb_pad5(N / 2)] = make_float2(4.0f * za.x * za.x, 4.0f * za.y * za.y);
                    }
                } else {
                    for (int k = g; k <= N / 2; k += GT) {
                        const int k2 = (N - k) & (N - 1);
                        const float2 za = buf[pb_pad5(k)], zb = buf[pb_pad5(k2)];
                        const float2 p = make_float2(za.x + zb.x, za.y - zb.y), q = make_float2(za.x - zb.x, za.y + zb.y);
                        const float2 w = make_float2(p.x * p.x + p.y * p.y, q.x * q.x + q.y * q.y);
                        buf[pb_pad5(k)] = w; buf[pb_pad5(k2)] = w;
                    }
                }
                PB_K1_SYNC();
            }
        }
        if (!staged_next) PB_K1_STAGE_NEXT();          // silent pair: nothing was transformed
#undef PB_K1_STAGE_NEXT
        // local peaks of the two frames: over the lanes (and warps) of the group, back to unscaled units (exact: powers of two)
        pkA = pb_warp_max(pkA); pkB = pb_warp_max(pkB);
        if (G > 1) {
            PB_K1_SYNC();
            if (lane == 0) { red[wg * 4 + 2] = pkA; red[wg * 4 + 3] = pkB; }
            PB_K1_SYNC();
            for (int k = 0; k < G; k++) { pkA = fmaxf(pkA, red[k * 4 + 2]); pkB = fmaxf(pkB, red[k * 4 + 3]); }
        }
        pkA = pkA / sA; pkB = pkB / sB;

        // ---- outputs: r[lag] = ac[lag] / (ac[0] * windowR[lag]) for lags 0..B+1 of the active frames, to the global scratch
        const int slot = 2 * li;
        const bool actA = active && pkA > 0.0f, actB = active && hasB && pkB > 0.0f;
        if (active && !REG_OUT) {
            // both rows of the slot pair are written whenever the pair is active (K2 only reads the rows slot_fr marks)
            const float2 ac0 = buf[0];
            const float2 inv0 = make_float2(ac0.x > 0.0f ? 1.0f / ac0.x : 0.0f, ac0.y > 0.0f ? 1.0f / ac0.y : 0.0f);
            float* ra = racf + (size_t)slot * rstride_g + g;
            float* rb = ra + rstride_g;
            const float2* src = buf + (g + (g >> 5));
            const float* iwp = gm.inv_wr + g;                  // inv_wr[B + 1] = 0
            PB_UNROLL for (int q = 0; q < C::RPL; q++) {
                const int lag = g + q * GT;
                if (lag <= B + 1) {
                    const float2 a = FAST ? src[q * (33 * G)] : buf[pb_pad5(lag)];
                    const float iw = __ldg(iwp + q * GT);
                    float2 r2 = __fmul2_rn(a, __fmul2_rn(inv0, make_float2(iw, iw)));
                    if (q == 0 && g == 0) r2 = make_float2(1.0f, 1.0f);
                    ra[q * GT] = r2.x; rb[q * GT] = r2.y;
                }
            }
        }
        if (g == 0) {
            // frames that cannot have a voiced candidate are finished here: one voiceless candidate, intensity as Praat's
            const int mc = gm.max_cand;
            slot_fr[slot] = actA ? frA : -1;
            slot_fr[slot + 1] = actB ? frA + 1 : -1;
            if (!actA) { cand_f[frA * mc] = 0.0f; cand_s[frA * mc] = 0.0f; ncand[frA] = 1; intensity[frA] = 0.0f; }
            else { const float t = pkA / gpk; intensity[frA] = t > 1.0f ? 1.0f : t; }
            if (hasB) {
                if (!actB) { cand_f[(frA + 1) * mc] = 0.0f; cand_s[(frA + 1) * mc] = 0.0f; ncand[frA + 1] = 1; intensity[frA + 1] = 0.0f; }
                else { const float t = pkB / gpk; intensity[frA + 1] = t > 1.0f ? 1.0f : t; }
            }
        }
        if (!REG_OUT) PB_K1_SYNC();     // buf is reused by the next iteration (REG_OUT: its last reads were the loads of the last pass, already fenced)
    }
}
#undef PB_K1_SYNC

// ------------------------------------------------------------------------------------------------ K1 at 2048 / 4096 points, split
// 24 kHz and 22.05 kHz at a 75 Hz floor (BASELINE configs 3 and 4) need a 2048-point transform and 44.1 kHz a 4096-point one, but
// their windows fill less than half of it (958 / 880 / 1764 samples) and fewer than a quarter of the lags are read (481 / 442 /
// 884).  With N = 1024 NP (NP = 2 or 4) and W = exp(-2 pi i / N) both facts prune radix-2 stages:
//   * zero-padded input: the bins k = NP j + q are a 1024-point transform of their own,
//         X[NP j + q] = DFT_1024(y_q)[j],   y_q[n] = W^(n q) sum_{m < NP/2} x[n + 1024 m] exp(-2 pi i m q / NP),
//     (NP = 2: y_q = x W^(nq); NP = 4: y_q = W^(nq) (x[n] + (-i)^q x[n + 1024]));
//   * lags below 1024 only:  r[t] = sum_q W^(q t) DFT_1024(P_q)[t],  P_q[j] = P[NP j + q] — the power spectrum of exactly those bins.
// So a pair of frames is NP independent 1024-point pipelines — one warp each, the register-blocked radix-32 x 32 code of the
// 1024-point kernel — that meet three times: when the window is written (every sample goes to all NP buffers, rotated), at the
// power spectrum (the conjugate partner of bin NP j + q is bin NP (1023 - j) + (NP - q): the q = 0 and q = NP/2 pipelines pair
// inside their own buffer, the others read their partner's; every lane computes the 32 power values of its own column straight
// into the registers the second transform's first pass wants, so nothing is written back) and when the lags are combined
// (NP - 1 complex multiply-adds per lag instead of log2 NP radix-2 passes through shared memory).
// The general kernel above runs these sizes as 32 x 32 x NP with NP warps sharing one buffer and a named barrier around every
// pass; it stays for geometries whose window exceeds N/2 or whose lag range exceeds 1024.  Measured (599 k frames): NP = 2 3.94 ->
// 3.24 ms (24 kHz), 3.86 -> 3.10 ms (22.05 kHz): the default at N = 2048.  NP = 4 8.08 -> 7.89 ms (44.1 kHz): two of its four
// pipelines read each other's spectra, the window is rotated four ways and a pair still takes four group barriers — 2 %, so N = 4096 keeps the general kernel unless
// PB_ACF_SPLIT=4 asks for this one.
// The power-of-two frame scales must be applied BEFORE the rotation by W^(nq), which mixes the two frames' components.
template <int NP, int MINB>
__global__ void __launch_bounds__((NP > PB_WPC ? NP : PB_WPC) * 32, MINB)
pb_pitch_acf_split_kernel(const int16_t* __restrict__ pcm, const PbUnitDev* __restrict__ units, const int32_t* __restrict__ pair_off,
                          const int4* __restrict__ pairpos, PbPitchGeomDev gm, int item0, int n_items, int rstride_g,
                          float* __restrict__ racf, long long* __restrict__ slot_fr,
                          float* __restrict__ cand_f, float* __restrict__ cand_s, uint8_t* __restrict__ ncand, float* __restrict__ intensity) {
    constexpr int R = 32, LR = 5, NH = 1024, G = NP, GT = 32 * NP;
    constexpr int BUFH = NH + (NH >> 5) + 8;            // float2 slots of one pipeline's buffer (skew padding as in the 1024-point kernel)
    constexpr int WARPS = NP > PB_WPC ? NP : PB_WPC;
    constexpr int GROUPS = WARPS / G;
    PB_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = warp / G, wg = warp % G;          // wg = q: the residue class of this warp's bins
    const int g = wg * 32 + lane;                       // thread in group
    const int bar_id = 1 + group;
#define PB_KS_SYNC() pb_group_sync<G>(bar_id)
    const size_t group_bytes = (size_t)(NP * BUFH + 8 * G) * sizeof(float2) + (size_t)gm.pre_cap * sizeof(int16_t);
    unsigned char* gbase = smem_raw + (size_t)group * group_bytes;
    float2* const bufs = (float2*)gbase;                // NP buffers
    float2* buf = bufs + wg * BUFH;                     // this warp's
    const int wq = (NP - wg) % NP;                      // the pipeline that holds this one's conjugate partners
    const float2* buf_partner = bufs + wq * BUFH;
    float* red = (float*)(bufs + NP * BUFH);            // [G][4] floats / [G][6] ints
    int16_t* pre = (int16_t*)(bufs + NP * BUFH + 8 * G);
    pbMbar* mbar = (pbMbar*)(smem_raw + (size_t)GROUPS * group_bytes) + group;
    if (g == 0) pb_mbar_init(mbar, 1);
    pb_mbar_init_fence();
    __syncthreads();
    unsigned phase = 0;
    const int B = gm.brent_ixmax;
    const int nw = gm.nw;
    constexpr int RPL = NP == 2 ? 16 : 32;              // rows of 32 lags a lane may have to deliver (lags 0 .. B + 1 < 32 RPL)
    const int pk_lo = max(0, gm.half_nw - gm.half_period), pk_n = min(nw, gm.half_nw + gm.half_period) - pk_lo;
    const int mean_n0 = gm.half_nw - gm.nsamp_period;
    const int mean_len = 2 * gm.nsamp_period;
    const int span_lo = min(0, mean_n0);
    const int span_hi = max(nw, mean_n0 + mean_len);
    const int span_len = span_hi - span_lo;
    const float mean_scale = (float)(1.0 / (32768.0 * (double)mean_len));
    const bool fuse_ok = mean_n0 >= 0 && mean_n0 + mean_len <= nw;
    const float2* __restrict__ tw_q = gm.tw_b;          // tw_q[q * 1024 + n] = W^(n q)

    const int n_groups = gridDim.x * GROUPS;
    const int per_group = (n_items + n_groups - 1) / n_groups;
    const int it_begin = (blockIdx.x * GROUPS + group) * per_group;
    const int it_end = min(n_items, it_begin + per_group);
    if (it_begin >= it_end) return;
    int u_next = pb_upper_unit(pair_off, gm.n_units, item0 + it_begin);
    int u_next_end = pair_off[u_next + 1];
    int u = -1;
    PbPairPos pos_next = pb_stage_pair<GT>(pcm, units, u_next, pairpos[item0 + it_begin], gm.pcm_len, span_lo, span_len, pre, g, mbar);
    int u_pair_off = 0, u_nframes = 0, u_pmin = 0, u_pmax1 = 0;
    long long u_frame_off = 0;
    float gpk = 0.0f;

    // one windowed, scaled sample pair (x = a + i b at n, and at n + 1024 when NP = 4) into the NP buffers, rotated
    auto scatter = [&](int n, float2 x0, float2 x1) {
        bufs[pb_pad5(n)] = NP == 2 ? x0 : make_float2(x0.x + x1.x, x0.y + x1.y);
        PB_UNROLL for (int q = 1; q < NP; q++) {
            float2 y;
            if (NP == 2) y = x0;
            else if (q == 1) y = make_float2(x0.x + x1.y, x0.y - x1.x);          // x0 - i x1
            else if (q == 2) y = make_float2(x0.x - x1.x, x0.y - x1.y);          // x0 - x1
            else y = make_float2(x0.x - x1.y, x0.y + x1.x);                      // x0 + i x1
            const float2 tw = __ldg(tw_q + q * NH + n);
            bufs[q * BUFH + pb_pad5(n)] = make_float2(y.x * tw.x - y.y * tw.y, y.x * tw.y + y.y * tw.x);
        }
    };

    for (int li = it_begin; li < it_end; li++) {
        const int item = item0 + li;
        const PbPairPos pos = pos_next;
        const int4 d_next = li + 1 < it_end ? __ldg(pairpos + item + 1) : make_int4(0, 0, 0, 0);
        if (u != u_next) {
            u = u_next;
            const PbUnitDev* up = units + u;
            u_pair_off = up->pair_off; u_nframes = up->n_frames; u_frame_off = up->frame_off; gpk = (float)up->global_peak;
            const long long ix1 = up->ix1, e = (long long)up->file_nx - ix1 + 1;
            const long long pmin = 2 - ix1 > 1 ? 2 - ix1 : 1, pmax1 = (up->nx < e ? up->nx : e) + 1;
            const long long BIG = 1LL << 30;
            u_pmin = (int)(pmin > BIG ? BIG : pmin); u_pmax1 = (int)(pmax1 > BIG ? BIG : (pmax1 < -BIG ? -BIG : pmax1));
        }
        const int fA = 2 * (item - u_pair_off);
        const bool hasB = fA + 1 < u_nframes;
        const bool global_silent = gpk == 0.0f;
        const long long frA = u_frame_off + fA;
        pb_mbar_wait(mbar, phase); phase ^= 1u;
#ifdef PB_SIMT_EMU
        PB_KS_SYNC();
#else
        if (pos.edge) PB_KS_SYNC();
#endif
        const int16_t* sm = pre;
        const int sb0 = pos.shift - span_lo, sb1 = sb0 + pos.hop;
        float pkA = 0.0f, pkB = 0.0f;
        float sA = 1.0f, sB = 1.0f;
        bool any_signal = false;
        const long long loA = (long long)u_pmin - pos.start0, hiA = (long long)u_pmax1 - pos.start0;
        const bool interior = fuse_ok && hasB && loA <= span_lo && hiA - pos.hop >= span_hi;
        if (interior) {
            // ---- exact integer statistics of the local-mean span (see the general kernel), then the scaled window into the buffers
            int s0, s1, mn0, mx0, mn1, mx1;
            {
                int sum[2]; unsigned vmn[2], vmx[2];
                PB_UNROLL for (int f = 0; f < 2; f++) {
                    const int M0 = (f ? sb1 : sb0) + mean_n0, M1 = M0 + mean_len;
                    const int k0 = (M0 + 1) >> 1, k1 = M1 >> 1;
                    const unsigned* smw = reinterpret_cast<const unsigned*>(sm);
                    int acc = 0; unsigned lo = 0x7fff7fffu, hi = 0x80008000u;
                    for (int k = k0 + g; k < k1; k += GT) {
                        const unsigned w = smw[k];
                        acc = __dp2a_lo((int)w, 0x0101, acc);
                        lo = __vmins2(lo, w); hi = __vmaxs2(hi, w);
                    }
                    if (g == 0) {
                        if (M0 & 1) acc += (int)sm[M0];
                        if (M1 & 1) acc += (int)sm[M1 - 1];
                    }
                    sum[f] = acc; vmn[f] = lo; vmx[f] = hi;
                }
                s0 = sum[0]; s1 = sum[1];
                mn0 = min((int)(short)(vmn[0] & 0xffff), (int)(short)(vmn[0] >> 16)); mx0 = max((int)(short)(vmx[0] & 0xffff), (int)(short)(vmx[0] >> 16));
                mn1 = min((int)(short)(vmn[1] & 0xffff), (int)(short)(vmn[1] >> 16)); mx1 = max((int)(short)(vmx[1] & 0xffff), (int)(short)(vmx[1] >> 16));
            }
            s0 = __reduce_add_sync(PB_FULL_MASK, s0); s1 = __reduce_add_sync(PB_FULL_MASK, s1);
            mn0 = __reduce_min_sync(PB_FULL_MASK, mn0); mx0 = __reduce_max_sync(PB_FULL_MASK, mx0);
            mn1 = __reduce_min_sync(PB_FULL_MASK, mn1); mx1 = __reduce_max_sync(PB_FULL_MASK, mx1);
            {
                int* redi = (int*)red;
                if (lane == 0) { redi[wg * 6 + 0] = s0; redi[wg * 6 + 1] = s1; redi[wg * 6 + 2] = mn0; redi[wg * 6 + 3] = mx0; redi[wg * 6 + 4] = mn1; redi[wg * 6 + 5] = mx1; }
                PB_KS_SYNC();
                s0 = 0; s1 = 0;
                PB_UNROLL for (int k = 0; k < G; k++) {
                    s0 += redi[k * 6 + 0]; s1 += redi[k * 6 + 1];
                    mn0 = min(mn0, redi[k * 6 + 2]); mx0 = max(mx0, redi[k * 6 + 3]); mn1 = min(mn1, redi[k * 6 + 4]); mx1 = max(mx1, redi[k * 6 + 5]);
                }
            }
            const float q15 = 1.0f / 32768.0f;
            const float meanA = (float)s0 * mean_scale, meanB = (float)s1 * mean_scale;
            const float mA = fmaxf(fabsf((float)mx0 * q15 - meanA), fabsf((float)mn0 * q15 - meanA));
            const float mB = fmaxf(fabsf((float)mx1 * q15 - meanB), fabsf((float)mn1 * q15 - meanB));
            sA = mA > 0.0f ? __int_as_float((254 - ((__float_as_int(mA) >> 23) & 0xff)) << 23) : 1.0f;
            sB = mB > 0.0f ? __int_as_float((254 - ((__float_as_int(mB) >> 23) & 0xff)) << 23) : 1.0f;
            const float2 q15s = make_float2(q15 * sA, q15 * sB), nms = make_float2(-meanA * sA, -meanB * sB);
            any_signal = mA > 0.0f || mB > 0.0f;
            const int16_t* pa = sm + sb0; const int16_t* pb = sm + sb1;
            auto windowed = [&](int n) -> float2 {
                if (n >= nw) return make_float2(0.0f, 0.0f);
                const float w0 = __ldg(gm.window + n);
                const float2 x = __fmul2_rn(__ffma2_rn(make_float2((float)pa[n], (float)pb[n]), q15s, nms), make_float2(w0, w0));
                if ((unsigned)(n - pk_lo) < (unsigned)pk_n) { pkA = fmaxf(pkA, fabsf(x.x)); pkB = fmaxf(pkB, fabsf(x.y)); }
                return x;
            };
            for (int n = g; n < NH; n += GT) scatter(n, windowed(n), NP > 2 ? windowed(n + NH) : make_float2(0.0f, 0.0f));
        } else {
            // ---- frames that Praat zero-fills beyond the file, or a unit with an odd frame count: masked samples; two sweeps, the
            //      magnitudes first (the scales are wanted before the rotation), then the scaled window into the buffers
            float mxA = 0.0f, mxB = 0.0f;
            int nlo[2], nhi[2], sb[2]; float lmean[2];
            PB_UNROLL for (int f = 0; f < 2; f++) {
                long long lo = loA - (f ? pos.hop : 0), hi = hiA - (f ? pos.hop : 0);
                const long long BIG = 1 << 30;
                nlo[f] = (int)(lo < -BIG ? -BIG : (lo > BIG ? BIG : lo));
                nhi[f] = (int)(hi < -BIG ? -BIG : (hi > BIG ? BIG : hi));
                sb[f] = f ? sb1 : sb0;
                int s = 0;
                for (int q = lane; q < mean_len; q += 32) {
                    const int n = mean_n0 + q;
                    s += (n >= nlo[f] && n < nhi[f]) ? (int)sm[sb[f] + n] : 0;
                }
                s = pb_warp_sum_i(s);
                lmean[f] = (float)s * mean_scale;
            }
            if (!hasB) { nlo[1] = 0; nhi[1] = 0; sb[1] = sb[0]; }
            const float2 nmean = make_float2(-lmean[0], -lmean[1]), q15 = make_float2(1.0f / 32768.0f, 1.0f / 32768.0f);
            const float hb = hasB ? 1.0f : 0.0f;
            auto windowed = [&](int n) -> float2 {
                if (n >= nw) return make_float2(0.0f, 0.0f);
                const float w = __ldg(&gm.window[n]);
                const int sa = (n >= nlo[0] && n < nhi[0]) ? (int)sm[sb[0] + n] : 0;
                const int sbv = (n >= nlo[1] && n < nhi[1]) ? (int)sm[sb[1] + n] : 0;
                return __fmul2_rn(__ffma2_rn(make_float2((float)sa, (float)sbv), q15, nmean), make_float2(w, w * hb));
            };
            for (int n = g; n < nw; n += GT) {
                const float2 x = windowed(n);
                const float aa = fabsf(x.x), bb = fabsf(x.y);
                mxA = fmaxf(mxA, aa); mxB = fmaxf(mxB, bb);
                if ((unsigned)(n - pk_lo) < (unsigned)pk_n) { pkA = fmaxf(pkA, aa); pkB = fmaxf(pkB, bb); }
            }
            mxA = pb_warp_max(mxA); mxB = pb_warp_max(mxB);
            if (lane == 0) { red[wg * 4 + 0] = mxA; red[wg * 4 + 1] = mxB; }
            PB_KS_SYNC();
            PB_UNROLL for (int k = 0; k < G; k++) { mxA = fmaxf(mxA, red[k * 4 + 0]); mxB = fmaxf(mxB, red[k * 4 + 1]); }
            sA = mxA > 0.0f ? __int_as_float((254 - ((__float_as_int(mxA) >> 23) & 0xff)) << 23) : 1.0f;
            sB = mxB > 0.0f ? __int_as_float((254 - ((__float_as_int(mxB) >> 23) & 0xff)) << 23) : 1.0f;
            pkA *= sA; pkB *= sB;
            const float2 sc = make_float2(sA, sB);
            for (int n = g; n < NH; n += GT)
                scatter(n, __fmul2_rn(windowed(n), sc), NP > 2 ? __fmul2_rn(windowed(n + NH), sc) : make_float2(0.0f, 0.0f));
            any_signal = mxA > 0.0f || mxB > 0.0f;
        }
        PB_KS_SYNC();                       // every buffer holds the windowed pair; the staging buffer is free
        const bool active = !global_silent && any_signal;
        if (li + 1 < it_end) {
            if (item + 1 >= u_next_end) { do { u_next++; u_next_end = pair_off[u_next + 1]; } while (item + 1 >= u_next_end); }
            pos_next = pb_stage_pair<GT>(pcm, units, u_next, d_next, gm.pcm_len, span_lo, span_len, pre, g, mbar);
        }
        // ---- this warp's 1024-point pipeline: two transforms x two passes through one copy of the butterfly code
        float2 v[R];
        int n_steps = active ? 4 : 0;
        asm volatile("" : "+r"(n_steps));
#ifndef PB_SIMT_EMU
#pragma unroll 1
#endif
        for (int s_it = 0; s_it < n_steps; s_it++) {
            int step = s_it;
            asm volatile("" : "+r"(step));
            int pass = step & 1;
            asm volatile("" : "+r"(pass));
            // ---- step 2 starts from the power spectra of both frames, packed as P_a + i P_b.  Bin j = lane + 32 i of pipeline q pairs
            //      with bin 1023 - j of pipeline NP - q (q > 0), with bin (1024 - j) mod 1024 of its own (q = 0).
            constexpr bool CROSS = NP > 2;
            const bool own_pair = wg == 0 || 2 * wg == NP;            // the partner bins sit in this warp's own buffer
            if (step == 2 && own_pair) {
                // in place, one evaluation per PAIR of bins (written to both), as in the 1024-point kernel
                if (wg == 0) {
                    const float2* pk_ = buf + lane;
                    float2* qk = buf + (lane ? 32 - lane : 33);
                    PB_UNROLL for (int i = 0; i < 16; i++) {
                        const float2 za = pk_[33 * i];
                        float2 zb = qk[33 * (31 - i)];
                        if (i == 0 && lane == 0) zb = za;
                        const float2 p = make_float2(za.x + zb.x, za.y - zb.y), q = make_float2(za.x - zb.x, za.y + zb.y);
                        const float2 w = make_float2(p.x * p.x + p.y * p.y, q.x * q.x + q.y * q.y);
                        ((float2*)pk_)[33 * i] = w;
                        if (!(i == 0 && lane == 0)) qk[33 * (31 - i)] = w;
                    }
                    if (lane == 0) {
                        const float2 za = buf[pb_pad5(NH / 2)];
                        buf[pb_pad5(NH / 2)] = make_float2(4.0f * za.x * za.x, 4.0f * za.y * za.y);
                    }
                } else {
                    // j = lane + 32 i -> slot lane + 33 i;  1023 - j = (31 - lane) + 32 (31 - i) -> slot (31 - lane) + 33 (31 - i)
                    float2* pk_ = buf + lane;
                    float2* qk = buf + (31 - lane);
                    PB_UNROLL for (int i = 0; i < 16; i++) {
                        const float2 za = pk_[33 * i], zb = qk[33 * (31 - i)];
                        const float2 p = make_float2(za.x + zb.x, za.y - zb.y), q = make_float2(za.x - zb.x, za.y + zb.y);
                        const float2 w = make_float2(p.x * p.x + p.y * p.y, q.x * q.x + q.y * q.y);
                        pk_[33 * i] = w; qk[33 * (31 - i)] = w;
                    }
                }
                __syncwarp();
            }
            if (!(CROSS && step == 2 && !own_pair)) {
                const float2* src = buf + lane;                       // i = lane + 32 t  ->  lane + 33 t
                PB_UNROLL for (int t = 0; t < R; t++) v[t] = src[t * 33];
            } else {
                // the partner bins sit in ANOTHER pipeline's buffer (which that warp must not lose to an in-place update): every lane
                // computes the 32 power values of its own column straight into the registers of the first pass
                const float2* own = buf + lane;
                const float2* qk = buf_partner + (31 - lane);
                PB_UNROLL for (int i = 0; i < R; i++) {
                    const float2 za = own[33 * i], zb = qk[33 * (31 - i)];
                    const float2 p = make_float2(za.x + zb.x, za.y - zb.y), q = make_float2(za.x - zb.x, za.y + zb.y);
                    v[i] = make_float2(p.x * p.x + p.y * p.y, q.x * q.x + q.y * q.y);
                }
            }
            // every partner has been read before any buffer is written again (the stores of this step): only the pipelines that read
            // each other's buffers meet (64 threads on a barrier of their own)
            if (CROSS && step == 2 && !own_pair) PB_GROUP_SYNC(8 + group, 64);
            if (pass) {
                const float2* tw = gm.tw_a + lane;
                PB_UNROLL for (int t = 1; t < R; t++) {
                    const float2 w = __ldg(tw + t * R);
                    const float2 x = v[t];
                    v[t] = __ffma2_rn(make_float2(x.y, x.y), make_float2(-w.y, w.x), __fmul2_rn(make_float2(x.x, x.x), w));
                }
            }
            __syncwarp();
            pb_dft<R>(v);
            if (step == 3) break;                                     // the lags stay in registers
            if (pass) { float2* dst = buf + lane; PB_UNROLL for (int t = 0; t < R; t++) dst[t * 33] = v[pb_bitrev(t, LR)]; }
            else { float2* dst = buf + 33 * lane; PB_UNROLL for (int t = 0; t < R; t++) dst[t] = v[pb_bitrev(t, LR)]; }
            // the first transform's spectrum is read across pipelines when NP > 2 (by the cross-paired ones only)
            if (NP > 2 && step == 1 && !own_pair) PB_GROUP_SYNC(8 + group, 64); else __syncwarp();
        }
        // local peaks over the group, back to unscaled units
        pkA = pb_warp_max(pkA); pkB = pb_warp_max(pkB);
        if (lane == 0) { red[wg * 4 + 2] = pkA; red[wg * 4 + 3] = pkB; }
        // ---- r[t] = sum_q W^(q t) F_q[t]: the pipelines q > 0 hand their lags (rotated) over through their own buffers
        if (active && wg > 0) {
            float2* dst = buf + lane;
            const float2* tw = tw_q + wg * NH + lane;
            PB_UNROLL for (int q = 0; q < RPL; q++) {
                if (32 * q <= B + 1) {
                    const float2 w = __ldg(tw + 32 * q);
                    const float2 x = v[pb_bitrev(q, LR)];
                    dst[32 * q] = make_float2(x.x * w.x - x.y * w.y, x.x * w.y + x.y * w.x);
                }
            }
        }
        PB_KS_SYNC();
        PB_UNROLL for (int k = 0; k < G; k++) { if (k != wg) { pkA = fmaxf(pkA, red[k * 4 + 2]); pkB = fmaxf(pkB, red[k * 4 + 3]); } }
        pkA = pkA / sA; pkB = pkB / sB;
        const int slot = 2 * li;
        const bool actA = active && pkA > 0.0f, actB = active && hasB && pkB > 0.0f;
        if (active && wg == 0) {
            auto lag = [&](int q) -> float2 {
                float2 a = v[pb_bitrev(q, LR)];
                PB_UNROLL for (int k = 1; k < NP; k++) { const float2 o = bufs[k * BUFH + lane + 32 * q]; a.x += o.x; a.y += o.y; }
                return a;
            };
            const float2 a0v = lag(0);
            const float2 ac0 = make_float2(__shfl_sync(PB_FULL_MASK, a0v.x, 0), __shfl_sync(PB_FULL_MASK, a0v.y, 0));
            const float2 inv0 = make_float2(ac0.x > 0.0f ? 1.0f / ac0.x : 0.0f, ac0.y > 0.0f ? 1.0f / ac0.y : 0.0f);
            float* ra = racf + (size_t)slot * rstride_g + lane;
            float* rb = ra + rstride_g;
            const float* iwp = gm.inv_wr + lane;               // inv_wr[B + 1] = 0
            PB_UNROLL for (int q = 0; q < RPL; q++) {
                if (lane + 32 * q <= B + 1) {
                    const float iw = __ldg(iwp + 32 * q);
                    float2 r2 = __fmul2_rn(lag(q), __fmul2_rn(inv0, make_float2(iw, iw)));
                    if (q == 0 && lane == 0) r2 = make_float2(1.0f, 1.0f);
                    ra[32 * q] = r2.x; rb[32 * q] = r2.y;
                }
            }
        }
        if (g == 0) {
            const int mc = gm.max_cand;
            slot_fr[slot] = actA ? frA : -1;
            slot_fr[slot + 1] = actB ? frA + 1 : -1;
            if (!actA) { cand_f[frA * mc] = 0.0f; cand_s[frA * mc] = 0.0f; ncand[frA] = 1; intensity[frA] = 0.0f; }
            else { const float t = pkA / gpk; intensity[frA] = t > 1.0f ? 1.0f : t; }
            if (hasB) {
                if (!actB) { cand_f[(frA + 1) * mc] = 0.0f; cand_s[(frA + 1) * mc] = 0.0f; ncand[frA + 1] = 1; intensity[frA + 1] = 0.0f; }
                else { const float t = pkB / gpk; intensity[frA + 1] = t > 1.0f ? 1.0f : t; }
            }
        }
        PB_KS_SYNC();     // the buffers and the reduction scratch are reused by the next pair
    }
}
#undef PB_KS_SYNC
