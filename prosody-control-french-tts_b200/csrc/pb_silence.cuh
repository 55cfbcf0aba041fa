// pb_silence.cuh — K5: silence detection with pydub.silence semantics.
//
// pydub's detect_silence slides a min_silence_len window over the audio at a 1 ms step and calls audioop.rms on every
// slice: O(samples x window) on the CPU.  Here one CTA owns a tile of window starts: its warps square-sum the PCM into
// one integer energy per millisecond bin (bin i = frames [int(i*rate/1000), int((i+1)*rate/1000)), pydub's own frame
// arithmetic), a block scan turns the bins into prefix sums, and every window is then one subtraction and one integer
// comparison.  PCM is read once (plus the window-length halo between tiles): the kernel is HBM-bound.
//
//   rms <= thresh   <=>   (unsigned) sqrt(S / n) <= floor(thresh)   <=>   S < (floor(thresh) + 1)^2 * n
//
// S (sum of squares) and n (slice length incl. the < 2 ms of zeros pydub pads at the end of the data) are integers, so
// the decision is exact; the host refuses parameters where n * (floor(thresh)+1)^2 could approach 2^53 (where audioop's
// double arithmetic itself would start rounding).
//
// Output: the boundaries of the runs of silent window starts, as sortable 64-bit keys appended through one atomic
// counter; the host sorts them and applies pydub's merge / invert / keep_silence / midpoint rules (a few entries per file).
#pragma once
#include "pb_rt.h"

struct PbSilFileDev {
    long long pcm_off;     // first sample of the file in the PCM buffer
    long long tile_off;    // first tile of this file in the launch's tile list
    double per_ms;         // frame_rate / 1000.0
    int nx;                // samples in the file
    int len_ms;            // len(audio_segment) = round(1000 * nx / rate)
    int n_win;             // window starts: len_ms - min_silence_len + 1 (0 if the file is shorter than the window)
    int file_id;
};

#define PB_SIL_WARPS 16
#define PB_SIL_PADW(w) ((w) + ((w) >> 5))       // one pad word per 32: lane-strided bin reads stay (nearly) conflict-free

// pydub frame_count(ms) = ms * (frame_rate / 1000.0), truncated by int(); files are < 2^31 - 2^16 samples (host check)
__device__ __forceinline__ int pb_sil_frame32(int ms, double per_ms) { return __double2int_rz(__dmul_rn((double)ms, per_ms)); }

// Sum of squares of the two s16 halves of wd, split as x * x = 256 * x * (x >> 8) + x * (x & 255) so that each half is one
// dp2a (16-bit x 8-bit dot product with accumulate): acc_hi += x0 * hi0 + x1 * hi1, acc_lo += x0 * lo0 + x1 * lo1.
__device__ __forceinline__ void pb_sil_sq2(uint32_t wd, int& acc_hi, int& acc_lo) {
#ifdef PB_SIMT_EMU
    const int x0 = (int)(short)(wd & 0xffff), x1 = (int)wd >> 16;
    acc_hi += x0 * (x0 >> 8) + x1 * (x1 >> 8);
    acc_lo += x0 * (x0 & 255) + x1 * (x1 & 255);
#else
    const uint32_t b = __byte_perm(wd, 0, 0x3120);            // bytes (lo0, lo1, hi0, hi1)
    asm("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(acc_lo) : "r"(wd), "r"(b));
    asm("dp2a.hi.s32.s32 %0, %1, %2, %0;" : "+r"(acc_hi) : "r"(wd), "r"(b));
#endif
}

__device__ __forceinline__ int4 pb_sil_load_vec(const int16_t* __restrict__ pcm, const int4* __restrict__ pal, long long mis, long long pcm_len, long long v) {
    const long long s0 = (v << 3) - mis;
    if (s0 >= 0 && s0 + 8 <= pcm_len) return pal[v];
    int x[8];                                                   // first / last vector of a buffer that is not 16-byte aligned
    PB_UNROLL for (int k = 0; k < 8; k++) x[k] = (s0 + k >= 0 && s0 + k < pcm_len) ? (int)pcm[s0 + k] & 0xffff : 0;
    return make_int4(x[0] | (x[1] << 16), x[2] | (x[3] << 16), x[4] | (x[5] << 16), x[6] | (x[7] << 16));
}

// tile_windows: window starts per CTA tile; the tile needs tile_windows + win_ms + 1 bins (+1 window of halo each side).
// smem layout: u64 bins[nb_cap + 1] | u64 warp_tot[PB_SIL_WARPS] | u8 flags[tile_windows + 2] | u32 stage[PB_SIL_WARPS][words_per_warp]
// NV: 16-byte vectors one lane keeps in flight per 32-bin group (32 * NV * 8 samples cover 32 ms at the batch's highest rate)
template <int NV>
__global__ void __launch_bounds__(PB_SIL_WARPS * 32, 2)
pb_silence_runs_kernel(const int16_t* __restrict__ pcm, long long pcm_len, const PbSilFileDev* __restrict__ files, int n_files,
                       long long n_tiles, int tile_windows, int win_ms, int nb_cap, int words_per_warp, long long limit_per_sample,
                       unsigned long long* __restrict__ run_keys, unsigned long long run_cap, unsigned long long* __restrict__ run_count) {
    PB_DYN_SMEM(smem);
    unsigned long long* bins = reinterpret_cast<unsigned long long*>(smem);
    unsigned long long* warp_tot = bins + nb_cap + 1;
    unsigned char* flags = reinterpret_cast<unsigned char*>(warp_tot + PB_SIL_WARPS);
    uint32_t* stage_all = reinterpret_cast<uint32_t*>(flags + ((tile_windows + 2 + 15) & ~15));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* stage = stage_all + (size_t)warp * words_per_warp;
    const long long mis = (long long)((((size_t)pcm) & 15) >> 1);            // samples between the last 16-byte boundary and pcm
    const int4* __restrict__ pal = reinterpret_cast<const int4*>(pcm - mis);

    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        int lo = 0, hi = n_files - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (files[mid].tile_off <= t) lo = mid; else hi = mid - 1; }
        const PbSilFileDev F = files[lo];
        const int j0 = (int)(t - F.tile_off) * tile_windows;                  // first window start of the tile
        const int jn = min(tile_windows, F.n_win - j0);                       // windows in the tile
        const int base = max(j0 - 1, 0);                                      // first bin (one window of halo to the left)
        const int wend = min(j0 + jn + 1, F.n_win);                           // one past the last window whose flag we need
        const int nb = (wend - 1 + win_ms) - base;                            // bins [base, base + nb)
        // ---- 1. per-millisecond energies: each warp stages the samples of 32 consecutive bins, each lane sums one bin.
        // The 16-byte loads of the warp's NEXT group are issued before the current group is summed (registers q).
        {
            // frames [a, b) of this lane's bin, first vector and vector count of the group
            auto geom = [&](int gg, int& ga, int& gb, long long& gv0, int& gnvec) {
                const int i0 = base + gg * 32, iend = min(i0 + 32, base + nb);
                const int fa = min(pb_sil_frame32(min(i0 + lane, iend), F.per_ms), F.nx);
                const int fB = min(pb_sil_frame32(iend, F.per_ms), F.nx);
                int fb = __shfl_down_sync(PB_FULL_MASK, fa, 1);
                if (lane == 31) fb = fB;
                const int fA = __shfl_sync(PB_FULL_MASK, fa, 0);
                ga = fa; gb = fb;
                gv0 = (F.pcm_off + fA + mis) >> 3;
                gnvec = (int)(((F.pcm_off + fB + mis + 7) >> 3) - gv0);
            };
            int4 q[NV];
            auto load = [&](long long gv0, int gnvec) {
                PB_UNROLL for (int c = 0; c < NV; c++) {
                    const int vi = c * 32 + lane;
                    if (vi < gnvec) q[c] = pb_sil_load_vec(pcm, pal, mis, pcm_len, gv0 + vi);
                }
            };
            int g = warp, a = 0, b = 0, nvec = 0;
            long long v0 = 0;
            if (g * 32 < nb) { geom(g, a, b, v0, nvec); load(v0, nvec); }
            while (g * 32 < nb) {
                PB_UNROLL for (int c = 0; c < NV; c++) {
                    const int vi = c * 32 + lane;
                    if (vi < nvec) {
                        uint32_t* d = stage + (vi * 4 + (vi >> 3));            // = PADW(4 vi): the four words never straddle a pad
                        d[0] = (uint32_t)q[c].x; d[1] = (uint32_t)q[c].y; d[2] = (uint32_t)q[c].z; d[3] = (uint32_t)q[c].w;
                    }
                }
                for (int vi = NV * 32 + lane; vi < nvec; vi += 32) {           // a rate above what NV covers: the rest, unpipelined
                    const int4 r = pb_sil_load_vec(pcm, pal, mis, pcm_len, v0 + vi);
                    uint32_t* d = stage + (vi * 4 + (vi >> 3));
                    d[0] = (uint32_t)r.x; d[1] = (uint32_t)r.y; d[2] = (uint32_t)r.z; d[3] = (uint32_t)r.w;
                }
                __syncwarp();
                const int gn = g + PB_SIL_WARPS;
                int an = 0, bn = 0, nvecn = 0;
                long long v0n = 0;
                if (gn * 32 < nb) { geom(gn, an, bn, v0n, nvecn); load(v0n, nvecn); }
                int k = (int)(F.pcm_off + a + mis - (v0 << 3));
                const int e = k + (b - a);
                unsigned long long sum = 0;
                if (k < e) {
                    if (k & 1) { const int x = (int)stage[PB_SIL_PADW(k >> 1)] >> 16; sum = (unsigned long long)((long long)x * x); k++; }
                    const int wl = e >> 1;
                    int acc_hi = 0, acc_lo = 0;
                    for (int w = k >> 1; w < wl;) {                              // runs of words between two pad slots
                        const int w_end = min(wl, (w | 31) + 1);
                        const uint32_t* sp = stage + (w + (w >> 5));
                        const int n = w_end - w;
#pragma unroll 4
                        for (int c = 0; c < n; c++) pb_sil_sq2(sp[c], acc_hi, acc_lo);
                        w = w_end;
                    }
                    sum += (unsigned long long)((long long)acc_hi * 256 + (long long)acc_lo);
                    if (e & 1) { const int x = (int)(short)(stage[PB_SIL_PADW(wl)] & 0xffff); sum += (unsigned long long)((long long)x * x); }
                }
                if (g * 32 + lane < nb) bins[g * 32 + lane] = sum;
                __syncwarp();
                g = gn; a = an; b = bn; v0 = v0n; nvec = nvecn;
            }
        }
        if (threadIdx.x == 0) bins[nb] = 0;
        __syncthreads();
        // ---- 2. exclusive prefix sums over bins[0 .. nb]: every warp owns one contiguous segment (total first, then scan)
        {
            const int seg = (((nb + 1) + PB_SIL_WARPS - 1) / PB_SIL_WARPS + 31) & ~31;
            const int s_lo = warp * seg, s_hi = min(s_lo + seg, nb + 1);
            unsigned long long tot = 0;
            for (int k = s_lo + lane; k < s_hi; k += 32) tot += bins[k];
            PB_UNROLL for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(PB_FULL_MASK, tot, o);
            if (lane == 0) warp_tot[warp] = tot;
            __syncthreads();
            unsigned long long carry = 0;
            for (int w = 0; w < warp; w++) carry += warp_tot[w];
            for (int k0 = s_lo; k0 < s_hi; k0 += 32) {
                const int k = k0 + lane;
                const unsigned long long v = k < s_hi ? bins[k] : 0;
                unsigned long long inc = v;
                PB_UNROLL for (int o = 1; o < 32; o <<= 1) { const unsigned long long up = __shfl_up_sync(PB_FULL_MASK, inc, o); if (lane >= o) inc += up; }
                if (k < s_hi) bins[k] = carry + inc - v;
                carry += __shfl_sync(PB_FULL_MASK, inc, 31);
            }
        }
        __syncthreads();
        // ---- 3. one flag per window start j in [base .. wend): silent <=> S < limit * n  (or an empty slice: rms 0)
        for (int k = threadIdx.x; k < wend - base; k += blockDim.x) {
            const int j = base + k;
            const unsigned long long S = bins[k + win_ms] - bins[k];
            const int sf = pb_sil_frame32(j, F.per_ms), ef = pb_sil_frame32(j + win_ms, F.per_ms);
            const int real = min(ef, F.nx) - min(sf, F.nx);
            const long long cnt = real > 0 ? (long long)(ef - sf) : 0;
            flags[k + 1 - (j0 - base)] = (cnt == 0 || (long long)S < limit_per_sample * cnt) ? 1 : 0;   // flags[0] <-> window j0 - 1
        }
        if (threadIdx.x == 0) {
            if (j0 == 0) flags[0] = 0;                                        // no window before the first one
            if (j0 + jn >= F.n_win) flags[jn + 1] = 0;                        // nor after the last
        }
        __syncthreads();
        // ---- 4. boundaries of the runs of silent window starts
        for (int k = threadIdx.x; k < jn; k += blockDim.x) {
            if (!flags[k + 1]) continue;
            const unsigned long long key = ((unsigned long long)F.file_id << 33) | ((unsigned long long)(j0 + k) << 1);
            if (!flags[k]) { const unsigned long long slot = atomicAdd(run_count, 1ULL); if (slot < run_cap) run_keys[slot] = key; }
            if (!flags[k + 2]) { const unsigned long long slot = atomicAdd(run_count, 1ULL); if (slot < run_cap) run_keys[slot] = key | 1ULL; }
        }
        __syncthreads();
    }
}
