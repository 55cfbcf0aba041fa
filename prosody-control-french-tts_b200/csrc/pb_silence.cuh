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

// pydub frame_count(ms) = ms * (frame_rate / 1000.0), truncated by int()
__device__ __forceinline__ long long pb_sil_frame(int ms, double per_ms) { return (long long)__dmul_rn((double)ms, per_ms); }

// tile_windows: window starts per CTA tile; the tile needs tile_windows + win_ms + 1 bins (+1 window of halo each side).
// smem layout: u64 bins[nb_cap + 1] | u64 warp_tot[PB_SIL_WARPS] | u8 flags[tile_windows + 2] | u32 stage[PB_SIL_WARPS][words_per_warp]
__global__ void __launch_bounds__(PB_SIL_WARPS * 32)
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
        // ---- 1. per-millisecond energies: each warp stages the samples of 32 consecutive bins, each lane sums one bin
        for (int g = warp; g * 32 < nb; g += PB_SIL_WARPS) {
            const int i = base + g * 32 + lane;
            const int ic = min(i, base + nb), ie = min(i + 1, base + nb);
            long long a = pb_sil_frame(ic, F.per_ms), b = pb_sil_frame(ie, F.per_ms);
            a = a < F.nx ? a : F.nx; b = b < F.nx ? b : F.nx;
            const long long A = __shfl_sync(PB_FULL_MASK, a, 0);
            long long B = pb_sil_frame(min(base + g * 32 + 32, base + nb), F.per_ms);
            B = B < F.nx ? B : F.nx;
            const long long gA = F.pcm_off + A + mis, gB = F.pcm_off + B + mis; // sample positions counted from the aligned base
            const long long v0 = gA >> 3, v1 = (gB + 7) >> 3;
            for (long long v = v0 + lane; v < v1; v += 32) {
                const long long s0 = (v << 3) - mis;
                int4 q;
                if (s0 >= 0 && s0 + 8 <= pcm_len) q = pal[v];
                else {
                    int x[8];
                    PB_UNROLL for (int k = 0; k < 8; k++) x[k] = (s0 + k >= 0 && s0 + k < pcm_len) ? (int)pcm[s0 + k] & 0xffff : 0;
                    q.x = x[0] | (x[1] << 16); q.y = x[2] | (x[3] << 16); q.z = x[4] | (x[5] << 16); q.w = x[6] | (x[7] << 16);
                }
                const int w = (int)(v - v0) * 4;
                stage[PB_SIL_PADW(w)] = (uint32_t)q.x; stage[PB_SIL_PADW(w + 1)] = (uint32_t)q.y;
                stage[PB_SIL_PADW(w + 2)] = (uint32_t)q.z; stage[PB_SIL_PADW(w + 3)] = (uint32_t)q.w;
            }
            __syncwarp();
            int k = (int)(F.pcm_off + a + mis - (v0 << 3));
            const int e = k + (int)(b - a);
            unsigned long long sum = 0;
            if (k < e && (k & 1)) { const int x = (int)stage[PB_SIL_PADW(k >> 1)] >> 16; sum += (unsigned)(x * x); k++; }
            for (; k + 1 < e; k += 2) {
                const uint32_t wd = stage[PB_SIL_PADW(k >> 1)];
                const int x0 = (int)(short)(wd & 0xffff), x1 = (int)wd >> 16;
                sum += (unsigned)(x0 * x0); sum += (unsigned)(x1 * x1);
            }
            if (k < e) { const int x = (int)(short)(stage[PB_SIL_PADW(k >> 1)] & 0xffff); sum += (unsigned)(x * x); }
            if (g * 32 + lane < nb) bins[g * 32 + lane] = sum;
            __syncwarp();
        }
        if (threadIdx.x == 0) bins[nb] = 0;
        __syncthreads();
        // ---- 2. exclusive prefix sums over bins[0 .. nb]: every warp scans one contiguous segment, 32 bins per step
        const int seg = (((nb + 1) + PB_SIL_WARPS - 1) / PB_SIL_WARPS + 31) & ~31;
        {
            unsigned long long carry = 0;
            const int s_lo = warp * seg, s_hi = min(s_lo + seg, nb + 1);
            for (int k0 = s_lo; k0 < s_hi; k0 += 32) {
                const int k = k0 + lane;
                const unsigned long long v = k < s_hi ? bins[k] : 0;
                unsigned long long inc = v;
                PB_UNROLL for (int o = 1; o < 32; o <<= 1) { const unsigned long long up = __shfl_up_sync(PB_FULL_MASK, inc, o); if (lane >= o) inc += up; }
                if (k < s_hi) bins[k] = carry + inc - v;
                carry += __shfl_sync(PB_FULL_MASK, inc, 31);
            }
            if (lane == 0) warp_tot[warp] = carry;
        }
        __syncthreads();
        // ---- 3. one flag per window start j in [base .. wend): silent <=> S < limit * n  (or an empty slice: rms 0)
        for (int k = threadIdx.x; k < wend - base; k += blockDim.x) {
            const int j = base + k;
            unsigned long long off_a = 0, off_b = 0;
            const int sa = k / seg, sb = (k + win_ms) / seg;
            for (int w = 0; w < PB_SIL_WARPS; w++) { const unsigned long long wt = warp_tot[w]; if (w < sa) off_a += wt; if (w < sb) off_b += wt; }
            const unsigned long long S = (bins[k + win_ms] + off_b) - (bins[k] + off_a);
            const long long sf = pb_sil_frame(j, F.per_ms), ef = pb_sil_frame(j + win_ms, F.per_ms);
            const long long real = (ef < F.nx ? ef : F.nx) - (sf < F.nx ? sf : F.nx);
            const long long cnt = real > 0 ? ef - sf : 0;
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
