/*
 * prosody_b200.h — C ABI of libprosody_b200.so (hand-written sm_100a CUDA behind plain pointers and sizes).
 *
 * The reference (hi-paris/Prosody-Control-French-TTS) has no FFI: its hot path is four Python closures
 * inside AudioPipeline.measure_prosody_and_build_ssml that call parselmouth / pyloudnorm / pydub once per
 * (wav file, t0, t1).  Each entry point below is the BATCHED form of one of those closures; the units of
 * work are the reference's own call arguments (file, t0, t1[, meter rate]) in float64 seconds.
 *
 *   pb_median_pitch_batch   <- get_median_pitch(wav, t0, t1)      Code/audioPipeline.py:326-335
 *   pb_lufs_batch           <- get_lufs(wav, meter, t0, t1)       Code/audioPipeline.py:338-358
 *   pb_part_duration_batch  <- get_part_duration(wav, t0, t1)     Code/audioPipeline.py:314-323
 *                              get_duration(wav)                  Code/audioPipeline.py:360-361
 *   pb_extract_batch        <- the four above for one list of units, host PCM in, host records out
 *                              (what one pass of the step's loops :375-400 / :495-521 needs per unit)
 *   pb_intensity_batch      <- Sound.to_intensity()               Code/visualisation/Compare_speech_noenhanced.py:19-26
 *   pb_legacy_loudness_batch<- _calculate_loudness(path, s, e)    Code/Pipeline/compute_loudness_adjustments.py:8-25
 *   (has_t1 = 2 units)      <- calculate_pitch_segment            Code/Pipeline/compute_pitch_adjustments.py:167-208
 *   pb_split_on_silence_*   <- pydub.silence.split_on_silence     Code/Preprocessing/preprocess_audio.py:41-46
 *   pb_reduce_intervals     <- per-word aggregation of the frame tracks (BASELINE.json north_star)
 *   pb_segment_baselines, pb_syntagme_deltas, pb_ema_clamp <- the step's host arithmetic  Code/audioPipeline.py:401-424, 515-602
 *   pb_textgrid_*           <- textgrid.TextGrid.fromFile(p)[0]   Code/Preprocessing/gen_break_ssml.py:19-26
 *
 * Conventions
 *   - All functions return PB_OK (0) or a PB_E* code; pb_last_error(h) gives the message.
 *   - A "file" is mono 16-bit PCM (what the reference pipeline writes); all files of a call live in one
 *     concatenated int16 buffer, file f occupying [file_off[f], file_off[f]+file_nx[f]).
 *   - `pcm_on_device` selects whether `pcm` is a device pointer (HBM-resident input) or a host pointer
 *     (the library stages it through its own stream; pin it for full PCIe speed).  The library works on its own
 *     streams: device PCM must be complete when the call is made (synchronise the stream that produced it).
 *   - Unit descriptors and per-unit results are HOST arrays owned by the caller.  Optional per-frame outputs
 *     are host arrays sized with pb_pitch_plan().
 *   - The library owns only the handle: tables per analysis geometry and a scratch arena that grows on demand.
 *     A handle is bound to one device and one stream and is not thread-safe; use one per GPU.
 *   - There is no CPU fallback: without a CUDA device pb_create fails with PB_ENODEVICE.
 */
#ifndef PROSODY_B200_H
#define PROSODY_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PB_ABI_VERSION 2

enum {
    PB_OK = 0,
    PB_EINVAL = 1,      /* bad argument */
    PB_ENODEVICE = 2,   /* no usable CUDA device */
    PB_ECUDA = 3,       /* CUDA runtime error (message in pb_last_error) */
    PB_ENOMEM = 4,
    PB_EUNSUPPORTED = 5 /* geometry outside the kernels' compiled range */
};

/* per-unit status (what the reference's third-party call would have done) */
enum {
    PB_UNIT_OK = 0,
    PB_UNIT_TOO_SHORT = 1,     /* Praat throws: slice shorter than periods_per_window / pitch_floor */
    PB_UNIT_NO_SAMPLES = 2,    /* Praat throws: extracted Sound would contain no samples */
    PB_UNIT_WINDOW = 3,        /* Praat throws: analysis window too short */
    PB_UNIT_LUFS_FALLBACK = 16,/* flag: slice empty or < 0.4 s -> whole-file loudness (audioPipeline.py:345-358) */
    PB_UNIT_LUFS_ERROR = 32,   /* even the whole file is < 0.4 s: pyloudnorm ValueError escapes */
    PB_UNIT_SLICE_ERROR = 64   /* pydub TooManyMissingFrames */
};

typedef struct PbHandle PbHandle;

/* parselmouth Sound.to_pitch(time_step, pitch_floor, pitch_ceiling) == Praat Sound_to_Pitch_ac with the
 * remaining arguments at Praat's defaults; all of them are exposed so the legacy callers
 * (Code/Pipeline/compute_pitch_adjustments.py:191-199) can be served too. */
typedef struct PbPitchParams {
    double time_step;            /* 0 -> auto = periods_per_window / pitch_floor / 4 */
    double pitch_floor;          /* reference: 150 */
    double pitch_ceiling;        /* reference: 600 */
    double periods_per_window;   /* 3.0 */
    double silence_threshold;    /* 0.03 */
    double voicing_threshold;    /* 0.45 */
    double octave_cost;          /* 0.01 */
    double octave_jump_cost;     /* 0.35 */
    double voiced_unvoiced_cost; /* 0.14 */
    int32_t max_candidates;      /* 15 */
    int32_t reserved;
} PbPitchParams;

/* One unit of work = one call of a reference closure. SoA, host memory, n_units entries each. */
typedef struct PbUnits {
    int64_t n_units;
    const int64_t* file_off;   /* sample offset of the unit's file inside pcm */
    const int64_t* file_nx;    /* samples in that file */
    const double* rate;        /* file sample rate (Hz) */
    const int32_t* has_t1;     /* 0: whole file (t1 is None); 1: slice [t0, t1]; 2 (pitch only): slice with
                                  extract_part(preserve_times=False), the legacy callers' form */
    const double* t0;          /* seconds (the reference passes start_ms/1000) */
    const double* t1;
    const double* meter_rate;  /* pyln.Meter(rate) the reference built for this call; may be NULL for pitch-only */
} PbUnits;

/* kernel timings of the last batch call, measured with CUDA events on the handle's stream */
typedef struct PbTimings {
    float total_ms;        /* first enqueue .. last result on host */
    float h2d_ms;          /* PCM + descriptor upload */
    float unit_stats_ms;   /* K0: per-unit mean / global peak */
    float frames_ms;       /* K1+K2 together: framing, FFT autocorrelation, candidates (acf_ms + cand_ms + the pair-position kernel) */
    float path_ms;         /* K3: Viterbi path finder + median of voiced */
    float lufs_ms;         /* K4: K-weighting + gated loudness */
    float intensity_ms;    /* K4b */
    float d2h_ms;
    int64_t n_frames;      /* pitch frames analysed */
    int64_t n_lufs_samples;/* samples filtered */
    int32_t n_launches;    /* kernels launched */
    float host_plan_ms;    /* host wall time spent planning the units before the first enqueue */
    float acf_ms;          /* K1: windowing + FFT autocorrelation -> global scratch (dominant kernel) */
    float cand_ms;         /* K2: candidate search + sinc refinement */
} PbTimings;

int pb_abi_version(void);
int pb_create(int device, PbHandle** out);
void pb_destroy(PbHandle* h);
const char* pb_last_error(const PbHandle* h);
/* stream: a cudaStream_t to enqueue on (0 -> the handle's own non-blocking stream) */
int pb_set_stream(PbHandle* h, void* stream);
int pb_get_timings(const PbHandle* h, PbTimings* out);
int pb_device_info(const PbHandle* h, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* total_mem);

void pb_pitch_params_default(PbPitchParams* p);

/* Host-only planning: what Praat would do with each unit (status) and how many frames it yields.
 * frame_off (n_units+1 entries, may be NULL) receives the exclusive prefix sum of n_frames over OK units. */
int pb_pitch_plan(const PbPitchParams* p, const PbUnits* u, int32_t* status, int32_t* n_frames, int64_t* frame_off);

/* get_median_pitch, batched.  Outputs (host): median_f0 (0.0 when no voiced frame), n_voiced, n_frames, status.
 * Optional per-frame outputs (host, frame_off[n_units] entries, layout given by pb_pitch_plan): f0 / strength
 * of the selected candidate (parselmouth selected_array), intensity = Praat's relative frame intensity. */
int pb_median_pitch_batch(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int pcm_on_device,
                          const PbUnits* u, const PbPitchParams* p,
                          double* median_f0, int32_t* n_voiced, int32_t* n_frames, int32_t* status,
                          float* frame_f0, float* frame_strength, float* frame_intensity);

/* get_lufs, batched (pydub ms slicing, peak normalisation, pyloudnorm K-weighted gated loudness, and the
 * reference's two whole-file fallbacks).  lufs may be -inf exactly where pyloudnorm returns -inf. */
int pb_lufs_batch(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int pcm_on_device,
                  const PbUnits* u, double* lufs, int32_t* status);

/* get_part_duration / get_duration, batched (host arithmetic only; h may be NULL). */
int pb_part_duration_batch(const PbUnits* u, double* duration_s, int32_t* status);

/* Everything one pass of the step needs for its units. want_* select the work; outputs may be NULL when
 * not wanted.  This is the call the end-to-end benchmark times with host PCM. */
int pb_extract_batch(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int pcm_on_device,
                     const PbUnits* u, const PbPitchParams* p,
                     const uint8_t* want_pitch, const uint8_t* want_lufs,
                     double* median_f0, int32_t* n_voiced, int32_t* n_frames,
                     double* lufs, double* duration_s, int32_t* status);

/* The same call in two halves, so that the host can prepare the next batch while the GPU works on this one (the reference's own
 * concurrency is one process per voice, Code/audioPipeline.py:1141-1150; here it is batches in flight).  pb_extract_submit plans
 * the units, enqueues the uploads, the kernels and the result download on the handle's streams and returns; n_frames and
 * duration_s (host arithmetic) are already final.  pb_extract_wait blocks until the GPU is done and fills median_f0, n_voiced,
 * lufs and status.  The unit arrays may be reused after submit returns; pcm (host) and the output arrays must stay valid until
 * wait returns.  One batch per handle may be pending: use two handles to keep two batches in flight. */
int pb_extract_submit(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int pcm_on_device,
                      const PbUnits* u, const PbPitchParams* p,
                      const uint8_t* want_pitch, const uint8_t* want_lufs,
                      double* median_f0, int32_t* n_voiced, int32_t* n_frames,
                      double* lufs, double* duration_s, int32_t* status);
int pb_extract_wait(PbHandle* h);

/* Praat Sound_to_Intensity (parselmouth Sound.to_intensity(minimum_pitch=100, time_step=0, subtract_mean=True)),
 * whole files only.  n_frames[f] / frame_off from pb_intensity_plan; intensity_db has frame_off[n] entries. */
int pb_intensity_plan(const PbUnits* u, double minimum_pitch, double time_step, int32_t* status,
                      int32_t* n_frames, int64_t* frame_off, double* t_first, double* dt);
int pb_intensity_batch(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int pcm_on_device,
                       const PbUnits* u, double minimum_pitch, double time_step, int subtract_mean,
                       float* intensity_db, int32_t* status);

/* The legacy loudness of Code/Pipeline/compute_pitch_adjustments.py:157-164 (_calculate_loudness), batched:
 * 20*log10(sqrt(|mean(samples**2)|)) over audio[start*1000:end*1000] (pydub FLOAT-millisecond slice; t0/t1 carry
 * start/end in seconds), with numpy's int16 wrap-around of samples**2 reproduced bit for bit.  An empty slice
 * yields NaN and a silent one -inf, as numpy does.  sum_sq / count (optional, may be NULL) receive the exact
 * integer sum of the wrapped squares and the mean's denominator, for hosts that must finish the formula with
 * their own log10 (numpy's differs from libm's in the last bit). */
int pb_legacy_loudness_batch(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int pcm_on_device,
                             const PbUnits* u, double* loudness_db, int64_t* sum_sq, int64_t* count);

/* pydub.silence.split_on_silence(audio, min_silence_len, silence_thresh, keep_silence), seek_step 1, batched over
 * whole mono files (Code/Preprocessing/preprocess_audio.py:41-46; config.yaml silence: 1000 ms / -50 dBFS / 300 ms).
 * keep_silence_ms < 0 means keep_silence=True (keep everything).  File f yields segments
 * [seg_off[f], seg_off[f+1]): the [start_ms, end_ms) of the final audio[start:end] slices, and (optional outputs)
 * the sample range first/n_samples inside the file plus the n_pad zeros pydub appends when the last millisecond
 * overshoots the data.  capacity = entries in the seg_* arrays, at least pb_split_on_silence_bound(). */
int64_t pb_split_on_silence_bound(const PbUnits* files, int min_silence_len_ms);
int pb_split_on_silence_batch(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int pcm_on_device, const PbUnits* files,
                              int min_silence_len_ms, double silence_thresh_db, int keep_silence_ms, int64_t capacity,
                              int64_t* seg_off, int32_t* seg_start_ms, int32_t* seg_end_ms,
                              int64_t* seg_first_sample, int32_t* seg_n_samples, int32_t* seg_n_pad);

/* ---- host-side arithmetic of the step (float64, same libm calls and operation order as the reference's Python) */

/* prosody settings read by the step (Code/audioPipeline.py:127-139, config.yaml prosody_settings) */
typedef struct PbDeltaParams {
    double pitch_semitones;            /* P_ST */
    double pitch_lower_clip_factor;
    double volume_pct;
    double rate_percent;               /* R_PCT */
    double threshold_duration_before_slowing_down;
    double slow_floor_per_sec;
} PbDeltaParams;

/* Per-syntagme raw deltas (Code/audioPipeline.py:515-577): pitch % from 12*log2(p_nat/base_f0) clipped in semitones,
 * volume % from the loudness gap to the baseline, rate % from words per second nat vs synth with the asymmetric
 * length scaling, extra slow-down and clamps. All arrays have n entries. */
int pb_syntagme_deltas(int64_t n, const double* p_nat, const double* base_f0, const double* base_loud, const double* l_syn,
                       const int32_t* word_count, const double* nat_total_s, const double* syn_total_s, const int32_t* pause_ms,
                       const PbDeltaParams* prm, double* raw_pitch, double* raw_volume, double* raw_rate);

/* Per-segment baselines (Code/audioPipeline.py:401-424): np.median of the voiced p_nat (> 0), of l_nat and of rate_ratio
 * over all segments (window < 0 for None, or window >= n) or over the sliding window [i - window/2, i + window/2 + 1).
 * An empty voiced set gives NaN, as `float(np.median([])) or 1.0` does.  All arrays have n entries. */
int pb_segment_baselines(int64_t n, const double* p_nat, const double* l_nat, const double* rate_ratio, int32_t window,
                         double* f0, double* loud, double* rate);

/* EMA over all rows then the forward jump clamp (Code/audioPipeline.py:593-602). out may alias x. */
int pb_ema_clamp(const double* x, int64_t n, double alpha, double max_jump, double* out);

/* ---- aggregation of per-frame tracks over word / syntagme intervals (the "interval reduction" of the path) */

/* Frame times of the pitch analysis of every unit: frame k of unit i sits at t_first[i] + k * dt[i] (host only). */
int pb_pitch_frame_times(const PbPitchParams* p, const PbUnits* u, double* t_first, double* dt);

/* Segmented reduction on the GPU.  Series s (e.g. the per-frame f0 of unit s from pb_median_pitch_batch) occupies
 * [frame_off[s], frame_off[s+1]) of f0 / track2; interval j takes the frames of series iv_series[j] whose time lies in
 * [iv_tmin[j], iv_tmax[j]] (Praat's Sampled_getWindowSamples rule, float64).  Outputs per interval: n_frames, n_voiced
 * (f0 > 0), np.median and mean of the voiced f0 (0.0 when none, like get_median_pitch), mean of track2 (intensity; may be
 * NULL together with mean_track2).  tracks_on_device: f0 / track2 are device pointers. */
int pb_reduce_intervals(PbHandle* h, int64_t n_series, const int64_t* frame_off, const double* t_first, const double* dt,
                        const float* f0, const float* track2, int tracks_on_device,
                        int64_t n_intervals, const int64_t* iv_series, const double* iv_tmin, const double* iv_tmax,
                        int32_t* n_frames, int32_t* n_voiced, double* median_f0, double* mean_f0, double* mean_track2);

/* ---- TextGrid input at corpus scale (host only, no handle): one tier of many files, parsed on host threads.
 * Replaces per-file `textgrid.TextGrid.fromFile(path)[0]` (Code/Preprocessing/gen_break_ssml.py:19-26).
 * status per file: 0 ok, 1 unreadable, 2 not a TextGrid / truncated, 3 no such tier.  pb_textgrid_copy fills caller
 * arrays: status/xmin/xmax [n_files], iv_off [n_files+1], tmin/tmax [n_intervals], mark_off [n_intervals+1] into the
 * UTF-8 `marks` pool [mark_bytes].  Times are rounded to 5 decimals and intervals with min >= max dropped, as the
 * textgrid package does.  n_threads <= 0: all host cores. */
typedef struct PbTextGridBatch PbTextGridBatch;
int pb_textgrid_parse_files(const char* const* paths, int64_t n_files, int tier_index, int n_threads, PbTextGridBatch** out);
int pb_textgrid_sizes(const PbTextGridBatch* b, int64_t* n_files, int64_t* n_intervals, int64_t* mark_bytes);
int pb_textgrid_copy(const PbTextGridBatch* b, int32_t* status, double* xmin, double* xmax, int64_t* iv_off,
                     double* tmin, double* tmax, int64_t* mark_off, char* marks);
void pb_textgrid_free(PbTextGridBatch* b);

/* ---- SSML emitters + CSV tables at corpus scale (host only, no handle): Code/audioPipeline.py:604-711.
 * One row per syntagme: segment name and text as UTF-8 pools with n+1 offsets, pause in ms, the three percentages the emitters print
 * with '{x:+.2f}%'.  Produces the exact bytes pandas.DataFrame(rows).to_csv(path, index=False) writes for BDD_ssml.csv (one row per
 * segment, in order of first appearance), BDD_syntagme_ssml.csv and BDD_syntagme_for_synth.csv, as malloc'ed buffers the caller
 * releases with pb_ssml_free.  pause_factor = inter_syntagme_pause_factor; n_threads <= 0: up to 16 host threads. */
int pb_ssml_csv(int64_t n, const char* seg_names, const int64_t* seg_off, const char* texts, const int64_t* text_off,
                const int32_t* pause_ms, const double* pitch, const double* rate, const double* volume, double pause_factor,
                const char* voice, int n_threads, char** csv_final, int64_t* n_final, char** csv_syntagme, int64_t* n_syntagme,
                char** csv_synth, int64_t* n_synth);
void pb_ssml_free(char* p);

#ifdef __cplusplus
}
#endif
#endif /* PROSODY_B200_H */
