// pb_api.cu — the C ABI of libprosody_b200.so (see include/prosody_b200.h): handle, device buffers, per-geometry
// tables, host-side planning and kernel orchestration.  Built by nvcc for sm_100a; the same source is compiled by
// g++ against the SIMT emulator for CPU-side tests only (tests/simt_emu).
//
// One batch call = plan every unit on the host (float64 index arithmetic, pb_plan.h), then for each PCM SEGMENT:
// upload that slice of the PCM on the copy stream, and on the compute stream (after an event) upload the descriptors of
// the units whose samples are now resident and launch K0 / K1+K2 / K3 / K4 for them.  With device-resident PCM there
// is a single segment; with host PCM the upload of segment s+1 overlaps the kernels of segment s.
#include "../../include/prosody_b200.h"
#include "pb_rt.h"
#include "pb_plan.h"
#include "pb_pitch.cuh"
#include "pb_pitch_acf.cuh"
#include "pb_pitch_cand.cuh"
#include "pb_pitch_path.cuh"
#include "pb_lufs.cuh"
#include "pb_silence.cuh"
#include "pb_intervals.cuh"

// descriptors: pinned host staging -> device, read by the SMs over PCIe (does not queue behind the PCM in the copy engine)
__global__ void pb_copy16_kernel(const int4* __restrict__ src, int4* __restrict__ dst, long long n16) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return 0;
        if (p) pbrt_free(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 256;
        if (pbrt_malloc(&p, want)) { p = nullptr; return 1; }
        cap = want; return 0;
    }
    void release() { if (p) pbrt_free(p); p = nullptr; cap = 0; }
};
struct HostBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return 0;
        if (p) pbrt_free_host(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 256;
        if (pbrt_malloc_host(&p, want)) { p = nullptr; return 1; }
        cap = want; return 0;
    }
    void release() { if (p) pbrt_free_host(p); p = nullptr; cap = 0; }
};

struct PitchTables {        // per (nw, log2n): lives for the life of the handle
    DevBuf window, inv_wr, tw_a, tw_b, half_tab;
};

struct EvPair { pbEvent_t a, b; int kind; };   // kind: index into the PbTimings float fields

struct LufsKey {
    int64_t first, len, npad; double meter;
    bool operator==(const LufsKey& o) const { return first == o.first && len == o.len && npad == o.npad && meter == o.meter; }
};
struct LufsKeyHash {
    size_t operator()(const LufsKey& k) const {
        uint64_t hsh = (uint64_t)k.first * 0x9E3779B97F4A7C15ull;
        hsh ^= ((uint64_t)k.len + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
        hsh ^= (uint64_t)k.npad * 0x165667B19E3779F9ull;
        uint64_t mb; memcpy(&mb, &k.meter, 8);
        hsh ^= mb * 0x27D4EB2F165667C5ull;
        return (size_t)(hsh ^ (hsh >> 29));
    }
};

struct PitchClass { PbGeomHost g; int gstatus = 0; double rate = 0.0; };

// everything the host decides about a batch before any GPU work
struct LufsResolved { int64_t a, b, npad; int st; };

struct BatchPlan {
    // pitch
    std::vector<PbUnitPlan> pplan;            // per caller unit
    std::vector<int32_t> pclass;              // class index per caller unit, -1 = no pitch work
    std::vector<PitchClass> classes;
    int max_cand = 0;
    int64_t total_frames = 0;
    // loudness: compact list of the units actually measured (duplicates removed)
    std::vector<PbLufsUnitDev> lunits;
    std::vector<int64_t> lneed;               // pcm offset one past the last sample each compact unit reads
    std::vector<PbMeterDev> meters;
    std::vector<std::pair<int64_t, int64_t>> dups;     // (caller unit, caller unit it duplicates)
    int64_t lufs_samples = 0;
    std::vector<int32_t> seen;                // dedup hash table (open addressing)
    std::vector<LufsKey> seen_key;            // key of every compact unit
    // per-call scratch kept here so its capacity (and its pages) survive from call to call
    std::vector<int32_t> pstat, lflags;
    std::vector<LufsResolved> lres;           // slice arithmetic of every unit (filled in parallel)
    std::vector<int64_t> fbase;               // first frame of every staged pitch unit
    std::vector<std::vector<int64_t>> pids, lids, by_class;
    void reset() {
        classes.clear(); max_cand = 0; total_frames = 0;       // pplan / pclass keep their size: plan_pitch rewrites them
        lunits.clear(); lneed.clear(); meters.clear(); dups.clear(); lufs_samples = 0;
        seen_key.clear();
    }
};

}  // namespace

struct PbHandle {
    int device = 0;
    int sm_count = 1, cc_major = 0, cc_minor = 0;
    long long total_mem = 0;
    pbStream_t own_stream = 0, stream = 0, copy_stream = 0, lufs_stream = 0;   // lufs_stream: loudness kernels run beside the path finder
    std::string err;
    // device buffers (grow on demand)
    DevBuf pcm, units, pair_off, cand_f, cand_s, ncand, inten, psi, sel_f, sel_s, med, nvoiced;
    DevBuf lunits, meters, lstate, lenergy, lufs, pairpos;
    DevBuf path_longs, path_jobs, path_ctr, path_T, path_D, path_maps, path_exits;   // K3, long chains (blocked path finder)
    DevBuf stats_longs, stats_ctr;                                                    // K0, long units
    DevBuf lufs_longs, lufs_jobs, lufs_ctr, lufs_P, lufs_Z, lufs_S;                   // K4, long units
    DevBuf racf, slot_fr, work_ctr;                        // K1 -> K2: autocorrelations of one launch chunk, slot -> frame map, K2's chunk counter
    size_t cand_smem = 0; int cand_per_sm = 1;             // K2: footprint its function attribute was set for, resident CTAs per SM
    HostBuf stage_units, stage_pairs, stage_lunits, stage_meters, stage_out;
    size_t su_off = 0, sp_off = 0, sl_off = 0;           // running offsets (elements) into the pinned staging buffers
    std::map<std::pair<int64_t, int>, PitchTables*> tables;
    std::vector<EvPair> evs;
    std::vector<pbEvent_t> ev_pool;
    size_t ev_used = 0;
    PbTimings last;
    std::map<std::pair<int, size_t>, int> occ_cache;      // (LOG2N, dynamic smem bytes) -> resident CTAs per SM
    std::map<int, size_t> occ_last;                        // LOG2N -> footprint the function attributes were last set for
    BatchPlan plan;                                        // host plan of the call in progress (scratch reused across calls)
    int64_t cur_pcm_len = 0;                               // samples in the pcm buffer of the call in progress
    // a submitted batch whose results have not been collected yet (pb_extract_submit / pb_extract_wait)
    struct Pending { bool active = false; int64_t n = 0; bool do_pitch = false, do_lufs = false;
                     double* median_f0 = nullptr; int32_t* n_voiced = nullptr; double* lufs = nullptr; int32_t* status = nullptr; } pending;
    size_t sil_smem = 0; int sil_per_sm = 2, sil_nv = 0;               // K5: footprint its function attribute was set for, resident CTAs per SM
};

namespace {

int fail(PbHandle* h, int code, const char* fmt, const char* detail = "") {
    char buf[512];
    snprintf(buf, sizeof buf, fmt, detail);
    if (h) h->err = buf;
    return code;
}
#define PB_CK(call, what) do { if (call) return fail(h, PB_ECUDA, "%s", (std::string(what) + ": " + pbrt_error()).c_str()); } while (0)
#define PB_CKMEM(call, what) do { if (call) return fail(h, PB_ENOMEM, "out of memory: %s", what); } while (0)

pbEvent_t* next_event(PbHandle* h) {
    if (h->ev_used == h->ev_pool.size()) { pbEvent_t e; pbrt_event_create(&e); h->ev_pool.push_back(e); }
    return &h->ev_pool[h->ev_used++];
}
struct ScopedEv {     // records an event pair around a section of a stream
    PbHandle* h; size_t idx; pbStream_t s;
    ScopedEv(PbHandle* h_, int kind, pbStream_t s_ = 0) : h(h_), s(s_ ? s_ : h_->stream) {
        EvPair p; p.kind = kind; p.a = *next_event(h); p.b = *next_event(h);
        pbrt_event_record(&p.a, s);
        h->evs.push_back(p); idx = h->evs.size() - 1;
    }
    ~ScopedEv() { pbrt_event_record(&h->evs[idx].b, s); }
};
enum { EV_TOTAL = 0, EV_H2D, EV_STATS, EV_FRAMES, EV_PATH, EV_LUFS, EV_INTENSITY, EV_D2H, EV_ACF, EV_CAND };

void begin_call(PbHandle* h) {
    pbrt_stream_sync(h->lufs_stream);          // belt and braces: failed calls drain their streams themselves (DrainOnError)
    h->evs.clear(); h->ev_used = 0;
    h->su_off = h->sp_off = h->sl_off = 0;
    memset(&h->last, 0, sizeof h->last);
}
void end_call(PbHandle* h) {   // after the streams have been synchronised
    float* f = &h->last.total_ms;                       // kinds 0..7 are the first eight floats of PbTimings, in order
    for (auto& p : h->evs) {
        const float ms = pbrt_event_ms(p.a, p.b);
        if (p.kind == EV_ACF) h->last.acf_ms += ms;
        else if (p.kind == EV_CAND) h->last.cand_ms += ms;
        else f[p.kind] += ms;
    }
}

// ------------------------------------------------------------------------------------------------ tables
template <int LOG2N> void fill_twiddles(std::vector<float2>& a, std::vector<float2>& b) {
    typedef PbFftCfg<LOG2N> C;
    const double PI2 = 6.283185307179586476925286766559;
    a.resize((size_t)C::R * C::R);
    for (int t = 0; t < C::R; t++) for (int k = 0; k < C::R; k++) {
        const double ang = -PI2 * (double)((t * k) % (C::R * C::R)) / (double)(C::R * C::R);
        a[(size_t)t * C::R + k] = make_float2((float)cos(ang), (float)sin(ang));
    }
    const int F = C::F > 1 ? C::F : 1, RR = C::R * C::R;
    b.resize((size_t)F * RR);
    for (int t = 0; t < F; t++) for (int k = 0; k < RR; k++) {
        const double ang = -PI2 * (double)(((long long)t * k) % C::N) / (double)C::N;
        b[(size_t)t * RR + k] = make_float2((float)cos(ang), (float)sin(ang));
    }
}

int get_tables(PbHandle* h, const PbGeomHost& g, PitchTables** out) {
    auto key = std::make_pair(g.nw, (int)g.log2n);
    auto it = h->tables.find(key);
    if (it != h->tables.end()) { *out = it->second; return PB_OK; }
    const double PI2 = 6.283185307179586476925286766559;
    const int nw = (int)g.nw, B = (int)g.brent_ixmax;
    std::vector<double> w(nw);
    for (int i = 1; i <= nw; i++) w[i - 1] = 0.5 - 0.5 * cos((double)i * PI2 / (double)(nw + 1));
    std::vector<float> wf(nw), iw(B + 2);
    for (int i = 0; i < nw; i++) wf[i] = (float)w[i];
    double ac0 = 0.0;
    for (int i = 0; i < nw; i++) ac0 += w[i] * w[i];
    for (int lag = 0; lag <= B; lag++) {
        double s = 0.0;
        for (int i = 0; i + lag < nw; i++) s += w[i] * w[i + lag];
        iw[lag] = (float)(ac0 / s);
    }
    iw[B + 1] = 0.0f;
    std::vector<float2> ta, tb;
    switch (g.log2n) {
        case 8: fill_twiddles<8>(ta, tb); break;
        case 9: fill_twiddles<9>(ta, tb); break;
        case 10: fill_twiddles<10>(ta, tb); break;
        case 11: fill_twiddles<11>(ta, tb); break;
        case 12: fill_twiddles<12>(ta, tb); break;
        case 13: fill_twiddles<13>(ta, tb); break;
        default: return fail(h, PB_EUNSUPPORTED, "analysis window of %s samples needs an FFT beyond 8192 points", std::to_string(nw).c_str());
    }
    // windowed-sinc coefficients of Praat's NUM_interpolate_sinc at phi = 1/2, depth 70 (both sides are identical there)
    std::vector<float> ht(72, 0.0f);
    for (int m = 0; m < 70; m++) {
        const double d = 0.5 + m, c = (1.0 + cos(3.14159265358979323846 * d / 70.5)) / d / PI2;
        ht[m] = (float)((m & 1) ? -c : c);
    }
    PitchTables* t = new PitchTables();
    if (t->half_tab.ensure(ht.size() * 4) || t->window.ensure(wf.size() * 4) || t->inv_wr.ensure(iw.size() * 4) || t->tw_a.ensure(ta.size() * 8) || t->tw_b.ensure(tb.size() * 8)) {
        delete t; return fail(h, PB_ENOMEM, "out of memory: %s", "pitch tables");
    }
    // pageable -> device copies of a few KB; synchronous on purpose (tables are built once per geometry)
    pbrt_h2d(t->half_tab.p, ht.data(), ht.size() * 4, h->stream);
    pbrt_h2d(t->window.p, wf.data(), wf.size() * 4, h->stream);
    pbrt_h2d(t->inv_wr.p, iw.data(), iw.size() * 4, h->stream);
    pbrt_h2d(t->tw_a.p, ta.data(), ta.size() * 8, h->stream);
    pbrt_h2d(t->tw_b.p, tb.data(), tb.size() * 8, h->stream);
    pbrt_stream_sync(h->stream);
    h->tables[key] = t;
    *out = t;
    return PB_OK;
}

// ------------------------------------------------------------------------------------------------ planning (host only)
// fn(i0, i1) over [0, n) on a few host threads (the per-unit planning is independent float64 arithmetic)
template <class F> void pb_parallel_for(int64_t n, int64_t grain, F fn) {
    const int64_t want = grain > 0 ? n / grain : 1;
    // one process per GPU shares the host with its siblings (torchrun exports LOCAL_WORLD_SIZE): split the cores
    static const unsigned share = [] { const char* e = getenv("LOCAL_WORLD_SIZE"); const int v = e ? atoi(e) : 1; return (unsigned)(v > 1 ? v : 1); }();
    const unsigned cores = std::max(1u, std::thread::hardware_concurrency() / share);
    const unsigned nt = (unsigned)std::max<int64_t>(1, std::min<int64_t>(std::min(16u, cores), want));
    if (nt <= 1) { fn((int64_t)0, n); return; }
    std::vector<std::thread> pool;
    const int64_t per = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; t++) {
        const int64_t i0 = (int64_t)t * per, i1 = std::min<int64_t>(n, i0 + per);
        if (i0 < i1) pool.emplace_back(fn, i0, i1);
    }
    for (auto& th : pool) th.join();
}

int validate_units(PbHandle* h, const PbUnits* u, int64_t pcm_len) {
    if (!u || u->n_units < 0) return fail(h, PB_EINVAL, "%s", "units is null or n_units < 0");
    if (u->n_units && (!u->file_off || !u->file_nx || !u->rate || !u->has_t1 || !u->t0 || !u->t1))
        return fail(h, PB_EINVAL, "%s", "unit arrays must not be null");
    for (int64_t i = 0; i < u->n_units; i++) {
        if (u->file_off[i] < 0 || u->file_nx[i] < 0 || (pcm_len >= 0 && u->file_off[i] + u->file_nx[i] > pcm_len))
            return fail(h, PB_EINVAL, "unit %s: file range outside the pcm buffer", std::to_string(i).c_str());
        if (!(u->rate[i] > 0.0)) return fail(h, PB_EINVAL, "unit %s: rate must be positive", std::to_string(i).c_str());
        if (u->file_nx[i] > 0x7fffffffLL) return fail(h, PB_EUNSUPPORTED, "unit %s: file longer than 2^31 samples", std::to_string(i).c_str());
    }
    return PB_OK;
}

// Praat refuses these before any analysis (Sound_to_Pitch_any: minimumPitch / periodsPerWindow positive, at least two
// candidates); a non-finite or non-positive value would otherwise turn into an undefined int64 in pb_geom_for_rate.
const char* pitch_params_error(const PbPitchParams* p) {
    if (!p) return "pitch params is null";
    if (!(p->pitch_floor > 0.0) || !std::isfinite(p->pitch_floor)) return "pitch_floor must be positive and finite";
    if (!(p->pitch_ceiling > 0.0) || !std::isfinite(p->pitch_ceiling)) return "pitch_ceiling must be positive and finite";
    if (!(p->periods_per_window > 0.0) || !std::isfinite(p->periods_per_window)) return "periods_per_window must be positive and finite";
    if (!(p->time_step >= 0.0) || !std::isfinite(p->time_step)) return "time_step must be >= 0 (0 = automatic)";
    if (p->max_candidates < 2) return "max_candidates must be at least 2 (Praat: \"maximum number of candidates should be greater than 1\")";
    return nullptr;
}

// status / n_frames must be zero-initialised by the caller; only wanted units are touched
int plan_pitch(PbHandle* h, const PbUnits* u, const PbPitchParams* p, const uint8_t* want, int32_t* status, int32_t* n_frames, BatchPlan& bp) {
    const int64_t n = u->n_units;
    if (bp.pplan.size() != (size_t)n) bp.pplan.resize((size_t)n);     // every wanted entry is rewritten below: no re-zeroing per call
    bp.pclass.resize((size_t)n);
    // geometry class of every wanted unit (sequential: a handful of distinct rates), then the float64 planning of the
    // units themselves split over a few host threads, then the totals
    double last_rate = -1.0; int last_cls = -1;
    for (int64_t i = 0; i < n; i++) {
        if (want && !want[i]) { bp.pclass[(size_t)i] = -1; continue; }
        int ci = last_cls;
        if (u->rate[i] != last_rate) {
            ci = -1;
            for (size_t k = 0; k < bp.classes.size(); k++) if (bp.classes[k].rate == u->rate[i]) { ci = (int)k; break; }
            if (ci < 0) {
                PitchClass pc; pc.rate = u->rate[i]; pc.gstatus = pb_geom_for_rate(u->rate[i], *p, pc.g);
                bp.classes.push_back(pc); ci = (int)bp.classes.size() - 1;
            }
            last_rate = u->rate[i]; last_cls = ci;
        }
        bp.pclass[(size_t)i] = ci;
    }
    pb_parallel_for(n, 16384, [&](int64_t i0, int64_t i1) {
        for (int64_t i = i0; i < i1; i++) {
            const int ci = bp.pclass[(size_t)i];
            if (ci < 0) continue;
            const PitchClass& pc = bp.classes[(size_t)ci];
            PbUnitPlan& pl = bp.pplan[(size_t)i];
            pb_plan_pitch_unit(u->file_nx[i], u->rate[i], u->has_t1[i], u->t0[i], u->t1[i], *p, pc.g, pc.gstatus, pl);
            status[i] = pl.status; n_frames[i] = pl.n_frames;
            if (pl.status != PB_UNIT_OK) bp.pclass[(size_t)i] = -1;
        }
    });
    for (int64_t i = 0; i < n; i++) {
        const int ci = bp.pclass[(size_t)i];
        if (ci >= 0) { bp.total_frames += bp.pplan[(size_t)i].n_frames; bp.max_cand = bp.classes[(size_t)ci].g.max_cand; }
    }
    if (bp.max_cand > PB_MAXC) return fail(h, PB_EUNSUPPORTED, "%s", "pitch_ceiling / pitch_floor exceeds 32 candidates per frame");
    return PB_OK;
}

// flags must be zero-initialised by the caller; only wanted units are touched
int plan_lufs(PbHandle* h, const PbUnits* u, const uint8_t* want, int32_t* flags, BatchPlan& bp) {
    const int64_t n = u->n_units;
    std::map<double, int> meter_ix;
    // Units that resolve to the same samples and meter have the same loudness (the reference's < 0.4 s / empty-slice
    // fallbacks send every short syntagme of a file to that file's whole-file value): measure once, copy on the host.
    // flat open-addressing table (slot = 1 + index into bp.lunits, 0 = empty), kept in the plan scratch across calls
    bp.seen_key.clear();
    const LufsKeyHash hasher;
    double last_mr = -1.0; int last_meter = -1;
    bp.lunits.reserve((size_t)n); bp.lneed.reserve((size_t)n); bp.seen_key.reserve((size_t)n);
    // the slice arithmetic of every unit in parallel, the de-duplication below in order
    bp.lres.resize((size_t)n);
    pb_parallel_for(n, 16384, [&](int64_t i0, int64_t i1) {
        for (int64_t i = i0; i < i1; i++) {
            if (want && !want[i]) continue;
            const double mr = u->meter_rate ? u->meter_rate[i] : u->rate[i];
            LufsResolved& r = bp.lres[(size_t)i];
            r.st = pb_lufs_resolve(u->file_nx[i], u->rate[i], mr, u->has_t1[i], u->t0[i], u->t1[i], &r.a, &r.b, &r.npad);
        }
    });
    size_t n_whole = 0;                                          // units that resolve to a whole file: the table's load
    for (int64_t i = 0; i < n; i++) {
        if (want && !want[i]) continue;
        const LufsResolved& rs = bp.lres[(size_t)i];
        n_whole += (rs.a == 0 && rs.b == u->file_nx[i] && rs.npad == 0);
    }
    size_t cap = 1024;
    while (cap < n_whole * 2 + 16) cap <<= 1;
    bp.seen.assign(cap, 0);
    for (int64_t i = 0; i < n; i++) {
        if (want && !want[i]) continue;
        const double mr = u->meter_rate ? u->meter_rate[i] : u->rate[i];
        const LufsResolved& rs = bp.lres[(size_t)i];
        const int64_t a = rs.a, b = rs.b, npad = rs.npad;
        const int st = rs.st;
        flags[i] = st;
        if (st & (PB_UNIT_LUFS_ERROR | PB_UNIT_SLICE_ERROR)) continue;
        const LufsKey key{u->file_off[i] + a, b - a, npad, mr};
        // only units that resolve to a WHOLE file are looked up (that is where the reference's fallbacks pile up, and
        // it keeps the table small enough to stay in cache); two identical slices are simply measured twice
        if (a == 0 && b == u->file_nx[i] && npad == 0) {
            size_t slot = hasher(key) & (cap - 1);
            bool dup = false;
            while (bp.seen[slot]) {
                const size_t k = (size_t)bp.seen[slot] - 1;
                if (bp.seen_key[k] == key) { bp.dups.push_back(std::make_pair(i, (int64_t)bp.lunits[k].out_index)); dup = true; break; }
                slot = (slot + 1) & (cap - 1);
            }
            if (dup) continue;
            bp.seen[slot] = (int32_t)bp.lunits.size() + 1;
        }
        bp.seen_key.push_back(key);
        if (mr == last_mr) {                                    // nearly every unit of a voice uses the same meter
            PbLufsUnitDev d;
            d.pcm_off = u->file_off[i]; d.a = a; d.b = b; d.npad = npad; d.chunk_off = 0;
            const int64_t len = b - a + npad;
            d.n_blocks = (int32_t)pb_lufs_num_blocks(len, mr); d.n_chunks = d.n_blocks + 3;
            d.meter = last_meter; d.out_index = (int)i; d.inv_peak = 1.0;
            bp.lunits.push_back(d);
            bp.lneed.push_back(u->file_off[i] + b);
            bp.lufs_samples += len;
            continue;
        }
        auto it = meter_ix.find(mr);
        if (it == meter_ix.end()) {
            PbMeterDev md; memset(&md, 0, sizeof md);
            pb_kweight_coeffs(mr, md.b1, md.a1, md.b2, md.a2);
            md.rate = mr;
            int L0 = (int)std::floor(0.1 * mr) - 2; if (L0 < 1) L0 = 1;
            md.L0 = L0;
            // columns of A^L: run the homogeneous recurrence from each basis state
            for (int c = 0; c < 4; c++) {
                double s[4] = {0, 0, 0, 0}; s[c] = 1.0;
                for (int step = 1; step < L0 + PB_LUFS_NM; step++) {
                    const double y1 = s[0];
                    const double np0 = -md.a1[1] * y1 + s[1], np1 = -md.a1[2] * y1;
                    const double y2 = md.b2[0] * y1 + s[2];
                    const double nq0 = md.b2[1] * y1 - md.a2[1] * y2 + s[3], nq1 = md.b2[2] * y1 - md.a2[2] * y2;
                    s[0] = np0; s[1] = np1; s[2] = nq0; s[3] = nq1;
                    if (step >= L0) for (int r = 0; r < 4; r++) md.M[step - L0][r * 4 + c] = s[r];
                }
            }
            bp.meters.push_back(md);
            it = meter_ix.insert(std::make_pair(mr, (int)bp.meters.size() - 1)).first;
        }
        PbLufsUnitDev d;
        d.pcm_off = u->file_off[i]; d.a = a; d.b = b; d.npad = npad; d.chunk_off = 0;
        const int64_t len = b - a + npad;
        d.n_blocks = (int32_t)pb_lufs_num_blocks(len, mr); d.n_chunks = d.n_blocks + 3;
        d.meter = it->second; d.out_index = (int)i; d.inv_peak = 1.0;
        last_mr = mr; last_meter = it->second;
        bp.lunits.push_back(d);
        bp.lneed.push_back(u->file_off[i] + b);
        bp.lufs_samples += len;
    }
    return PB_OK;
}

// ------------------------------------------------------------------------------------------------ launches
// K1 (autocorrelation of every frame pair -> global scratch) and K2 (candidates from the scratch), in chunks of pairs so the
// scratch stays bounded (PB_RACF_BYTES, default 4 GiB): 2 * rstride_g floats per pair.
template <int LOG2N>
int launch_frames(PbHandle* h, const int16_t* d_pcm, const PbUnitDev* d_units, const int32_t* d_pair_off, const PbPitchGeomDev& gm,
                  float* cand_f, float* cand_s, uint8_t* ncand, float* inten) {
    typedef PbFftCfg<LOG2N> C;
    size_t smem = (size_t)C::GROUPS_PER_CTA * ((C::BUF + 8 * C::G) * sizeof(float2) + (size_t)gm.pre_cap * sizeof(int16_t)) + C::GROUPS_PER_CTA * sizeof(pbMbar);
    const int threads = C::WARPS_PER_CTA * 32;
    // N = 2048 with a window of at most 1024 samples and lags below 512 (24 kHz / 22.05 kHz at a 75 Hz floor, 44.1 kHz at 150 Hz):
    // two independent 1024-point pipelines per pair (pb_pitch_acf_split_kernel<2>).  PB_ACF_SPLIT=0 keeps the general kernel;
    // PB_ACF_SPLIT=4 also runs N = 4096 (44.1 kHz at 75 Hz) as FOUR pipelines — measured no faster than the general 32 x 32 x 4
    // kernel (8.05 against 8.08 ms for 599 k frames: two of the four pipelines read each other's spectra, five group barriers per
    // pair), so it is not the default.
    static const int acf_split = [] { const char* e = getenv("PB_ACF_SPLIT"); return e ? atoi(e) : 1; }();
    constexpr int NP = LOG2N == 11 ? 2 : (LOG2N == 12 ? 4 : 0);
    const bool split = NP != 0 && (NP == 2 ? acf_split != 0 : acf_split == 4) && gm.nw <= C::N / 2 && gm.brent_ixmax + 2 <= (NP == 2 ? 512 : 1024) && PB_WPC == 4;
    if (split) {
        const int groups = PB_WPC / NP > 0 ? PB_WPC / NP : 1;
        smem = (size_t)groups * ((NP * (1024 + 32 + 8) + 8 * NP) * sizeof(float2) + (size_t)gm.pre_cap * sizeof(int16_t)) + groups * sizeof(pbMbar);
    }
    static const int acf_ctas = [] { const char* e = getenv("PB_ACF_CTAS"); return e ? atoi(e) : 0; }();
    static const int acf_wsync = [] { const char* e = getenv("PB_ACF_WSYNC"); return e ? atoi(e) : 1; }();     // measured 3 % faster than named barriers
    auto kfn = pb_pitch_acf_kernel<LOG2N, C::MIN_CTAS, false>;
    if constexpr (LOG2N == 10) {
        if (acf_ctas == 5) kfn = acf_wsync ? pb_pitch_acf_kernel<LOG2N, 5, true> : pb_pitch_acf_kernel<LOG2N, 5, false>;
        else if (acf_wsync) kfn = pb_pitch_acf_kernel<LOG2N, C::MIN_CTAS, true>;
    }
    if constexpr (NP != 0) { if (split) kfn = pb_pitch_acf_split_kernel<NP, C::MIN_CTAS>; }
    static const int cand_ctas = [] { const char* e = getenv("PB_CAND_CTAS"); return e ? atoi(e) : 8; }();
    auto cfn = cand_ctas == 10 ? pb_pitch_cand_kernel<10> : cand_ctas == 12 ? pb_pitch_cand_kernel<12> : pb_pitch_cand_kernel<8>;
    int per_sm = 2;
    const int rstride_g = (gm.brent_ixmax + 2 + 3) & ~3;                 // floats per frame in the scratch (16-byte rows for the bulk copies)
    const size_t cand_smem = (size_t)PB_CAND_WARPS * ((size_t)(2 * (rstride_g + PB_MIR) + 3 * PB_MAXC) * sizeof(float) + 2 * sizeof(pbMbar)) + 72 * sizeof(float);
#ifndef PB_SIMT_EMU
    // function attributes persist, so they are (re)applied whenever the shared-memory footprint of this instantiation
    // changes (another analysis geometry with the same FFT size)
    const auto okey = std::make_pair((int)LOG2N + (split ? 100 : 0), smem);
    auto oc = h->occ_cache.find(okey);
    if (oc == h->occ_cache.end() || h->occ_last[LOG2N + (split ? 100 : 0)] != smem) {
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return fail(h, PB_ECUDA, "cudaFuncSetAttribute: %s", pbrt_error());
        cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        // ask for just the shared memory the resident CTAs use: the rest of the 228 KB stays L1 for the lookup tables
        // (window, 1/windowR, twiddles)
        int pct = (int)((100 * ((size_t)per_sm * (smem + 1024)) + 228 * 1024 - 1) / (228 * 1024));
        if (pct > 100) pct = 100;
        cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        h->occ_cache[okey] = per_sm;
        h->occ_last[LOG2N + (split ? 100 : 0)] = smem;
    } else per_sm = oc->second;
    if (h->cand_smem != cand_smem) {
        if (cudaFuncSetAttribute(cfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cand_smem) != cudaSuccess)
            return fail(h, PB_ECUDA, "cudaFuncSetAttribute: %s", pbrt_error());
        cudaFuncSetAttribute(cfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int cps = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cps, cfn, PB_CAND_WARPS * 32, cand_smem) != cudaSuccess || cps < 1) cps = 1;
        h->cand_per_sm = cps; h->cand_smem = cand_smem;
    }
#else
    h->cand_per_sm = 2;
#endif
    { const char* e = getenv("PB_FRAMES_CTAS"); if (e && atoi(e) > 0 && atoi(e) < per_sm) per_sm = atoi(e); }   // experiments: fewer resident CTAs
    // frame positions of every pair (float64 arithmetic, once)
    PB_CKMEM(h->pairpos.ensure((size_t)gm.n_pairs * sizeof(int4) + 16), "pair positions");
    {
        const int pgrid = (int)std::max(1LL, std::min(((long long)gm.n_pairs + 255) / 256, (long long)h->sm_count * 8));
        PB_LAUNCH(pb_pair_pos_kernel, dim3(pgrid), dim3(256), 0, h->stream, d_pcm, d_units, d_pair_off, gm, (int4*)h->pairpos.p);
        h->last.n_launches++;
    }
    static const size_t racf_budget = [] { const char* e = getenv("PB_RACF_BYTES"); const long long v = e ? atoll(e) : 0; return (size_t)(v > 0 ? v : (4LL << 30)); }();
    const size_t pair_bytes = 2 * (size_t)rstride_g * sizeof(float);
    const long long chunk_pairs = std::max<long long>(1024, (long long)(racf_budget / pair_bytes));
    const long long max_items = std::min<long long>(gm.n_pairs, chunk_pairs);
    PB_CKMEM(h->racf.ensure((size_t)max_items * pair_bytes + 64) || h->slot_fr.ensure((size_t)max_items * 2 * sizeof(long long) + 64) ||
             h->work_ctr.ensure(sizeof(unsigned) * 4096), "autocorrelation scratch");
    const long long n_chunks = ((long long)gm.n_pairs + chunk_pairs - 1) / chunk_pairs;
    if (n_chunks > 4096) return fail(h, PB_EUNSUPPORTED, "%s", "PB_RACF_BYTES too small for this batch");
    PB_CK(pbrt_memset(h->work_ctr.p, 0, sizeof(unsigned) * (size_t)n_chunks, h->stream), "memset");
    for (long long ck = 0; ck < n_chunks; ck++) {
        const int item0 = (int)(ck * chunk_pairs);
        const int n_items = (int)std::min<long long>(chunk_pairs, (long long)gm.n_pairs - item0);
        const long long need = ((long long)n_items + C::GROUPS_PER_CTA - 1) / C::GROUPS_PER_CTA;
        const int grid = (int)std::max(1LL, std::min(need, (long long)h->sm_count * per_sm));
        {
            ScopedEv ev(h, EV_ACF);
            PB_LAUNCH(kfn, dim3(grid), dim3(threads), smem, h->stream, d_pcm, d_units, d_pair_off, (const int4*)h->pairpos.p, gm, item0, n_items, rstride_g,
                      (float*)h->racf.p, (long long*)h->slot_fr.p, cand_f, cand_s, ncand, inten);
        }
        {
            ScopedEv ev(h, EV_CAND);
            const int n_slots = 2 * n_items;
            const int cgrid = (int)std::max(1LL, std::min<long long>(((long long)n_slots + 32 * PB_CAND_WARPS - 1) / (32 * PB_CAND_WARPS), (long long)h->sm_count * h->cand_per_sm));
            PB_LAUNCH(cfn, dim3(cgrid), dim3(PB_CAND_WARPS * 32), cand_smem, h->stream, (const float*)h->racf.p, (const long long*)h->slot_fr.p,
                      n_slots, rstride_g, gm, cand_f, cand_s, ncand, (unsigned*)h->work_ctr.p + ck);
        }
        h->last.n_launches += 2;
    }
    return PB_OK;
}

// Enqueue the F0 path for the caller units `ids` (all planned OK). Frame arrays are reused by every launch group:
// launches are stream-ordered, and per-frame values are only read back when there is a single segment.
// Results land in h->med / h->nvoiced (device, caller-indexed).
struct PitchLaunch { int cls; size_t uoff, poff, m; int64_t pairs; int64_t long_units = 0, long_blocks = 0, long_stats = 0; };
// K0: units with more samples than this are reduced piecewise by the whole grid (pb_pitch.cuh)
long long stats_long_nx() { const char* e = getenv("PB_STATS_LONG"); const long long x = e ? atoll(e) : 0; return x > 0 ? x : (1LL << 22); }
// K3: units with more frames than this are cut into blocks of path_block_len() frames (pb_pitch_path.cuh, "long chains")
// (read at every call: the tests switch them)
// (a launch group with few units cannot fill the GPU with one-warp walks: there the blocked form already pays at 2 048 frames)
int path_long_thresh(size_t units_in_group, int sm_count) {
    const char* e = getenv("PB_PATH_LONG"); const int x = e ? atoi(e) : 0;
    if (x > 0) return x;
    return units_in_group < (size_t)sm_count * 4 ? 2048 : 8192;
}
int path_block_len() { const char* e = getenv("PB_PATH_BLOCK"); const int x = e ? atoi(e) : 0; return x >= 2 && x <= PB_PATHL_BLOCK_MAX ? x : PB_PATHL_BLOCK_MAX; }
struct LufsLaunch { size_t off, m; int64_t chunks; int64_t long_units = 0, long_groups = 0; };
// K4: units with more 100 ms chunks than this go through the long-unit kernels, their chain cut into groups (pb_lufs.cuh, "long units")
int lufs_long_chunks() { const char* e = getenv("PB_LUFS_LONG"); const int x = e ? atoi(e) : 0; return x > 0 ? x : 4096; }
int lufs_group_len() { const char* e = getenv("PB_LUFS_GROUP"); const int x = e ? atoi(e) : 0; return x >= 1 && x <= 4096 ? x : 64; }

// Fill the pinned descriptor staging for the caller units `ids` (all planned OK), one launch group per geometry class.
int stage_pitch(PbHandle* h, const PbUnits* u, const BatchPlan& bp, const std::vector<int64_t>& ids,
                const std::vector<int64_t>* frame_off_by_unit, std::vector<PitchLaunch>& out) {
    if (ids.empty()) return PB_OK;
    const size_t ncls = bp.classes.size();
    std::vector<std::vector<int64_t>>& by_class = h->plan.by_class;       // scratch with reused capacity
    if (by_class.size() < ncls) by_class.resize(ncls);
    for (auto& v : by_class) v.clear();
    for (int64_t i : ids) by_class[(size_t)bp.pclass[(size_t)i]].push_back(i);
    int64_t frame_base = 0;
    for (size_t ci = 0; ci < ncls; ci++) {
        const std::vector<int64_t>& cid = by_class[ci];
        const size_t m = cid.size();
        if (!m) continue;
        PitchTables* tb_unused = nullptr;                      // build / upload the tables now, before any PCM copy is queued
        int rc = get_tables(h, bp.classes[ci].g, &tb_unused);
        if (rc != PB_OK) return rc;
        PbUnitDev* su = (PbUnitDev*)h->stage_units.p + h->su_off;
        int32_t* sp = (int32_t*)h->stage_pairs.p + h->sp_off;
        // running pair / frame offsets first (sequential, two adds per unit), then the 80-byte descriptors in parallel
        int64_t pairs = 0, n_long = 0, n_long_blocks = 0, n_long_stats = 0;
        const long long long_nx = stats_long_nx(); const int path_long = path_long_thresh(m, h->sm_count), path_block = path_block_len();   // (getenv: not per unit)
        std::vector<int64_t>& fbase = h->plan.fbase;
        fbase.resize(m);
        for (size_t k = 0; k < m; k++) {
            const PbUnitPlan& pl = bp.pplan[(size_t)cid[k]];
            sp[k] = (int32_t)pairs; fbase[k] = frame_base;
            pairs += (pl.n_frames + 1) / 2;
            frame_base += pl.n_frames;
            if (pl.nx > long_nx) n_long_stats++;
            if (pl.n_frames > path_long) { n_long++; n_long_blocks += (pl.n_frames + path_block - 1) / path_block; }
            if (pairs > 0x7ffffff0LL) return fail(h, PB_EUNSUPPORTED, "%s", "more than 2^31 frame pairs in one launch; split the batch");
        }
        sp[m] = (int32_t)pairs;
        pb_parallel_for((int64_t)m, 16384, [&](int64_t k0, int64_t k1) {
            for (int64_t k = k0; k < k1; k++) {
                const int64_t i = cid[(size_t)k];
                const PbUnitPlan& pl = bp.pplan[(size_t)i];
                PbUnitDev& d = su[k];
                d.pcm_off = u->file_off[i]; d.ix1 = pl.ix1; d.nx = pl.nx;
                d.frame_off = frame_off_by_unit ? (*frame_off_by_unit)[(size_t)i] : fbase[(size_t)k];
                d.x1 = pl.x1; d.t1 = pl.t1; d.mean = 0.0; d.global_peak = 0.0;
                d.file_nx = (int32_t)u->file_nx[i]; d.n_frames = pl.n_frames; d.pair_off = sp[k]; d.out_index = (int32_t)i;
            }
        });
        PitchLaunch L; L.cls = (int)ci; L.uoff = h->su_off; L.poff = h->sp_off; L.m = m; L.pairs = pairs;
        L.long_units = n_long; L.long_blocks = n_long_blocks; L.long_stats = n_long_stats;
        out.push_back(L);
        h->su_off += m; h->sp_off += m + 1;
    }
    return PB_OK;
}

// Launch K0, K1+K2, K3 for one staged group (its descriptors are already on the device). Frame arrays are reused by
// every group: launches are stream-ordered, and per-frame values are only read back when there is a single segment.
// Results land in h->med / h->nvoiced (device, caller-indexed).
int launch_pitch_group(PbHandle* h, const int16_t* d_pcm, const PbPitchParams* p, const BatchPlan& bp, const PitchLaunch& L) {
    PbRange r_("pb: launch K0 stats / K1 acf / K2 candidates / K3 path");
    {
        const PitchClass& pc = bp.classes[(size_t)L.cls];
        const size_t m = L.m;
        const int64_t pairs = L.pairs;
        PitchTables* tb = nullptr;
        int rc = get_tables(h, pc.g, &tb);
        if (rc != PB_OK) return rc;
        PbUnitDev* du = (PbUnitDev*)h->units.p + L.uoff;
        int32_t* dp = (int32_t*)h->pair_off.p + L.poff;
        PbPitchGeomDev gm;
        memset(&gm, 0, sizeof gm);
        gm.nw = (int)pc.g.nw; gm.half_nw = (int)pc.g.half_nw; gm.nsamp_period = (int)pc.g.nsamp_period; gm.half_period = (int)pc.g.half_period;
        gm.brent_ixmax = (int)pc.g.brent_ixmax;
        gm.scan_lim = (int)std::min(pc.g.max_lag, pc.g.brent_ixmax);
        gm.max_cand = pc.g.max_cand; gm.n_units = (int)m; gm.n_pairs = (int)pairs;
        // a maximum at lag i refines to a lag <= i+1: below this lag its frequency stays above the ceiling (never voiced)
        gm.min_refine_lag = (int)std::floor(1.0 / pc.g.dx / pc.g.ceiling) - 1;
        { const char* e = getenv("PB_PHASE_SYNC"); gm.phase_sync = e ? atoi(e) : 1; }
        {
            // staging buffer of the sample prefetch: the frame samples a pair reads (window and local-mean span of frame A,
            // the same shifted by one hop for frame B) plus up to 7 samples of 16-byte alignment slack on each side
            const int mean_n0 = gm.half_nw - gm.nsamp_period;
            const int span_lo = std::min(0, mean_n0), span_hi = std::max(gm.nw, mean_n0 + 2 * gm.nsamp_period);
            const int hop_max = (int)std::ceil(pc.g.dt / pc.g.dx) + 2;
            gm.pre_cap = ((span_hi - span_lo) + hop_max + 7 + 15) & ~7;
            gm.pcm_len = h->cur_pcm_len;
        }
        gm.sr = (float)(1.0 / pc.g.dx); gm.half_voicing = (float)(0.5 * p->voicing_threshold);
        gm.octave_cost = (float)p->octave_cost; gm.min_pitch = (float)p->pitch_floor;
        gm.dx = pc.g.dx; gm.dt = pc.g.dt; gm.ceiling = pc.g.ceiling; gm.silence_threshold = p->silence_threshold;
        gm.voicing_threshold = p->voicing_threshold; gm.octave_cost_d = p->octave_cost; gm.octave_jump_cost = p->octave_jump_cost;
        gm.voiced_unvoiced_cost = p->voiced_unvoiced_cost;
        gm.path_long = pc.g.max_cand <= 16 ? path_long_thresh(m, h->sm_count) : 0x7fffffff;      // the blocked path finder packs 16 candidates per word
        gm.path_block = path_block_len();
        gm.window = (const float*)tb->window.p; gm.inv_wr = (const float*)tb->inv_wr.p;
        gm.tw_a = (const float2*)tb->tw_a.p; gm.tw_b = (const float2*)tb->tw_b.p; gm.half_tab = (const float*)tb->half_tab.p;
        {
            ScopedEv ev(h, EV_STATS);
            int grid = (int)std::max<size_t>(1, std::min<size_t>((m + 7) / 8, (size_t)h->sm_count * 8));     // one warp per unit
            const long long long_nx = stats_long_nx();
            PB_LAUNCH(pb_unit_stats_kernel, dim3(grid), dim3(256), 0, h->stream, d_pcm, du, (int)m, long_nx);
            h->last.n_launches++;
            if (L.long_stats > 0) {
                PB_CKMEM(h->stats_longs.ensure((size_t)L.long_stats * sizeof(PbStatsLong)) || h->stats_ctr.ensure(16), "long unit statistics");
                PB_CK(pbrt_memset(h->stats_ctr.p, 0, 16, h->stream), "memset");
                const int g_idx = (int)std::max<size_t>(1, std::min<size_t>((m + 255) / 256, (size_t)h->sm_count));
                PB_LAUNCH(pb_unit_stats_long_index_kernel, dim3(g_idx), dim3(256), 0, h->stream, (const PbUnitDev*)du, (int)m, long_nx, (PbStatsLong*)h->stats_longs.p, (int*)h->stats_ctr.p);
                PB_LAUNCH(pb_unit_stats_long_kernel, dim3(h->sm_count * 8), dim3(256), 0, h->stream, d_pcm, (const PbUnitDev*)du, (PbStatsLong*)h->stats_longs.p, (const int*)h->stats_ctr.p);
                PB_LAUNCH(pb_unit_stats_long_fin_kernel, dim3(1), dim3(256), 0, h->stream, du, (const PbStatsLong*)h->stats_longs.p, (const int*)h->stats_ctr.p);
                h->last.n_launches += 3;
            }
        }
        {
            ScopedEv ev(h, EV_FRAMES);
            float* cf = (float*)h->cand_f.p; float* cs = (float*)h->cand_s.p;
            uint8_t* nc = (uint8_t*)h->ncand.p; float* it = (float*)h->inten.p;
            switch (pc.g.log2n) {
                case 8: rc = launch_frames<8>(h, d_pcm, du, dp, gm, cf, cs, nc, it); break;
                case 9: rc = launch_frames<9>(h, d_pcm, du, dp, gm, cf, cs, nc, it); break;
                case 10: rc = launch_frames<10>(h, d_pcm, du, dp, gm, cf, cs, nc, it); break;
                case 11: rc = launch_frames<11>(h, d_pcm, du, dp, gm, cf, cs, nc, it); break;
                case 12: rc = launch_frames<12>(h, d_pcm, du, dp, gm, cf, cs, nc, it); break;
                case 13: rc = launch_frames<13>(h, d_pcm, du, dp, gm, cf, cs, nc, it); break;
                default: rc = fail(h, PB_EUNSUPPORTED, "%s", "FFT size"); break;
            }
            if (rc != PB_OK) return rc;
        }
        {
            ScopedEv ev(h, EV_PATH);
            const int wpb = 4;
            int grid = (int)std::min<size_t>((m + wpb - 1) / wpb, (size_t)h->sm_count * 16);
            PB_LAUNCH(pb_pitch_path_kernel, dim3(grid), dim3(wpb * 32), 0, h->stream, (const PbUnitDev*)du, gm, (const float*)h->cand_f.p,
                      (const float*)h->cand_s.p, (const uint8_t*)h->ncand.p, (const float*)h->inten.p, (uint8_t*)h->psi.p,
                      (float*)h->sel_f.p, (float*)h->sel_s.p, (double*)h->med.p, (int32_t*)h->nvoiced.p);
            h->last.n_launches++;
            if (L.long_units > 0 && gm.path_long != 0x7fffffff) {
                // long chains: blocked (max, +) path finder, see pb_pitch_path.cuh
                const size_t nl = (size_t)L.long_units, nbk = (size_t)L.long_blocks;
                PB_CKMEM(h->path_longs.ensure(nl * sizeof(PbLongUnit)) || h->path_jobs.ensure(nbk * sizeof(int2)) || h->path_ctr.ensure(16) ||
                         h->path_T.ensure(nbk * 256 * sizeof(double)) || h->path_D.ensure(nbk * 16 * sizeof(double)) ||
                         h->path_maps.ensure(nbk * 8) || h->path_exits.ensure(nbk), "blocked path finder");
                PB_CK(pbrt_memset(h->path_ctr.p, 0, 16, h->stream), "memset");
                PbLongUnit* lg = (PbLongUnit*)h->path_longs.p; int2* jb = (int2*)h->path_jobs.p; int* ctr = (int*)h->path_ctr.p;
                double* T = (double*)h->path_T.p; double* D = (double*)h->path_D.p;
                unsigned long long* maps = (unsigned long long*)h->path_maps.p; uint8_t* exits = (uint8_t*)h->path_exits.p;
                const float* cf = (const float*)h->cand_f.p; const float* cs = (const float*)h->cand_s.p;
                const uint8_t* nc = (const uint8_t*)h->ncand.p; const float* it = (const float*)h->inten.p;
                const int cap = h->sm_count * 16;
                const int g_idx = (int)std::max<size_t>(1, std::min<size_t>((m + 255) / 256, (size_t)h->sm_count));
                const int g_entry = (int)std::max<size_t>(1, std::min<size_t>((8 * nbk + PB_PATHL_WARPS - 1) / PB_PATHL_WARPS, (size_t)cap));
                const int g_blk = (int)std::max<size_t>(1, std::min<size_t>((nbk + PB_PATHL_WARPS - 1) / PB_PATHL_WARPS, (size_t)cap));
                const int g_unit = (int)std::max<size_t>(1, std::min<size_t>(nl, (size_t)cap));
                PB_LAUNCH(pb_path_long_index_kernel, dim3(g_idx), dim3(256), 0, h->stream, (const PbUnitDev*)du, gm, lg, jb, ctr);
                PB_LAUNCH(pb_path_block_entry_kernel, dim3(g_entry), dim3(PB_PATHL_WARPS * 32), 0, h->stream, (const PbUnitDev*)du, gm, cf, cs, nc, it,
                          (const PbLongUnit*)lg, (const int2*)jb, (const int*)ctr, T);
                PB_LAUNCH(pb_path_block_scan_kernel, dim3(g_unit), dim3(32), 0, h->stream, (const PbUnitDev*)du, gm, nc, lg, (const int*)ctr, (const double*)T, D);
                PB_LAUNCH(pb_path_block_final_kernel, dim3(g_blk), dim3(PB_PATHL_WARPS * 32), 0, h->stream, (const PbUnitDev*)du, gm, cf, cs, nc, it,
                          (const PbLongUnit*)lg, (const int2*)jb, (const int*)ctr, (const double*)D, (uint8_t*)h->psi.p, maps);
                PB_LAUNCH(pb_path_block_link_kernel, dim3(g_unit), dim3(32), 0, h->stream, (const PbLongUnit*)lg, (const int*)ctr, (const unsigned long long*)maps, exits);
                PB_LAUNCH(pb_path_block_select_kernel, dim3(g_blk), dim3(PB_PATHL_WARPS * 32), 0, h->stream, (const PbUnitDev*)du, gm, cf, cs,
                          (const PbLongUnit*)lg, (const int2*)jb, (const int*)ctr, (const uint8_t*)h->psi.p, (const uint8_t*)exits,
                          (float*)h->sel_f.p, (float*)h->sel_s.p);
                PB_LAUNCH(pb_path_median_long_kernel, dim3(g_unit), dim3(1024), 0, h->stream, (const PbUnitDev*)du, (const PbLongUnit*)lg, (const int*)ctr,
                          (const float*)h->sel_f.p, (double*)h->med.p, (int32_t*)h->nvoiced.p);
                h->last.n_launches += 7;
            }
        }
        PB_CK(pbrt_last_error(), "pitch kernels");
    }
    return PB_OK;
}

// Enqueue the loudness path for the compact units `ids` (indices into bp.lunits). Results land in h->lufs.
void stage_lufs(PbHandle* h, const BatchPlan& bp, const std::vector<int64_t>& ids, LufsLaunch& L) {
    const size_t m = ids.size();
    PbLufsUnitDev* su = (PbLufsUnitDev*)h->stage_lunits.p + h->sl_off;
    int64_t chunks = 0;
    const int long_chunks = lufs_long_chunks(), group = lufs_group_len();
    L.long_units = 0; L.long_groups = 0;
    for (size_t k = 0; k < m; k++) {
        su[k] = bp.lunits[(size_t)ids[k]]; su[k].chunk_off = chunks; chunks += su[k].n_chunks;
        if (su[k].n_chunks > long_chunks) { L.long_units++; L.long_groups += (su[k].n_chunks + group - 1) / group; }
    }
    L.off = h->sl_off; L.m = m; L.chunks = chunks;
    h->sl_off += m;
}

int launch_lufs_group(PbHandle* h, const int16_t* d_pcm, const LufsLaunch& L, pbStream_t stream) {
    PbRange r_("pb: launch K4 loudness (peak, chunk, scan, chunk, gate)");
    const size_t m = L.m;
    if (!m) return PB_OK;
    PbLufsUnitDev* du = (PbLufsUnitDev*)h->lunits.p + L.off;
    const int64_t chunks = L.chunks;
    {
        ScopedEv ev(h, EV_LUFS, stream);
        const PbMeterDev* dm = (const PbMeterDev*)h->meters.p;
        double* st = (double*)h->lstate.p; double* en = (double*)h->lenergy.p;
        const int long_chunks = lufs_long_chunks(), group = lufs_group_len();
        const bool has_long = L.long_units > 0;
        PbLufsLong* lg = nullptr; int2* jb = nullptr; int* ctr = nullptr; double* P = nullptr; double* Z = nullptr; double* S = nullptr;
        int g_long = 1, g_grp = 1;
        if (has_long) {
            const size_t nl = (size_t)L.long_units, ng = (size_t)L.long_groups;
            PB_CKMEM(h->lufs_longs.ensure(nl * sizeof(PbLufsLong)) || h->lufs_jobs.ensure(ng * sizeof(int2)) || h->lufs_ctr.ensure(16) ||
                     h->lufs_P.ensure(ng * 16 * 8) || h->lufs_Z.ensure(ng * 4 * 8) || h->lufs_S.ensure(ng * 4 * 8), "long loudness units");
            lg = (PbLufsLong*)h->lufs_longs.p; jb = (int2*)h->lufs_jobs.p; ctr = (int*)h->lufs_ctr.p;
            P = (double*)h->lufs_P.p; Z = (double*)h->lufs_Z.p; S = (double*)h->lufs_S.p;
            g_long = (int)std::min<size_t>(nl, (size_t)h->sm_count * 16);
            g_grp = (int)std::max<size_t>(1, std::min<size_t>((ng + 31) / 32, (size_t)h->sm_count * 16));
            PB_CK(pbrt_memset(ctr, 0, 16, stream), "memset");
            const int g_idx = (int)std::max<size_t>(1, std::min<size_t>((m + 255) / 256, (size_t)h->sm_count));
            PB_LAUNCH(pb_lufs_long_index_kernel, dim3(g_idx), dim3(256), 0, stream, (const PbLufsUnitDev*)du, (int)m, long_chunks, group, lg, jb, ctr);
            PB_LAUNCH(pb_lufs_peak_long_kernel, dim3(h->sm_count * 8), dim3(256), 0, stream, d_pcm, (const PbLufsUnitDev*)du, lg, (const int*)ctr);
            PB_LAUNCH(pb_lufs_peak_fin_kernel, dim3(1), dim3(256), 0, stream, du, (const PbLufsLong*)lg, (const int*)ctr);
            h->last.n_launches += 3;
        }
        int g1 = (int)std::max<size_t>(1, std::min<size_t>((m + 7) / 8, (size_t)h->sm_count * 8));           // one warp per unit
        PB_LAUNCH(pb_lufs_peak_kernel, dim3(g1), dim3(256), 0, stream, d_pcm, du, (int)m, long_chunks);
        int gc = (int)std::max<int64_t>(1, std::min<int64_t>((chunks + 127) / 128, (int64_t)h->sm_count * 16));
        auto k_state = pb_lufs_chunk_kernel<false>; auto k_energy = pb_lufs_chunk_kernel<true>;
        PB_LAUNCH(k_state, dim3(gc), dim3(128), 0, stream, d_pcm, (const PbLufsUnitDev*)du, (int)m, dm, (long long)chunks, st, en);
        int gu = (int)std::max<size_t>(1, std::min<size_t>((m + 127) / 128, (size_t)h->sm_count * 16));
        int gs = (int)std::max<size_t>(1, std::min<size_t>((m + 31) / 32, (size_t)h->sm_count * 16));         // 8 units per warp
        PB_LAUNCH(pb_lufs_scan_kernel, dim3(gs), dim3(128), 0, stream, (const PbLufsUnitDev*)du, (int)m, dm, st, long_chunks);
        if (has_long) {
            auto k_p1 = pb_lufs_scan_group_kernel<1>; auto k_p3 = pb_lufs_scan_group_kernel<3>;
            PB_LAUNCH(k_p1, dim3(g_grp), dim3(128), 0, stream, (const PbLufsUnitDev*)du, dm, (const PbLufsLong*)lg, (const int2*)jb, (const int*)ctr, group, st, P, Z, (const double*)S);
            PB_LAUNCH(pb_lufs_scan_link_kernel, dim3(g_long), dim3(32), 0, stream, (const PbLufsLong*)lg, (const int*)ctr, (const double*)P, (const double*)Z, S);
            PB_LAUNCH(k_p3, dim3(g_grp), dim3(128), 0, stream, (const PbLufsUnitDev*)du, dm, (const PbLufsLong*)lg, (const int2*)jb, (const int*)ctr, group, st, P, Z, (const double*)S);
            h->last.n_launches += 3;
        }
        PB_LAUNCH(k_energy, dim3(gc), dim3(128), 0, stream, d_pcm, (const PbLufsUnitDev*)du, (int)m, dm, (long long)chunks, st, en);
        PB_LAUNCH(pb_lufs_gate_kernel, dim3(gu), dim3(128), 0, stream, (const PbLufsUnitDev*)du, (int)m, dm, (const double*)en, (double*)h->lufs.p, long_chunks);
        if (has_long) {
            PB_LAUNCH(pb_lufs_gate_long_kernel, dim3(g_long), dim3(256), 0, stream, (const PbLufsUnitDev*)du, (const PbLufsLong*)lg, (const int*)ctr, dm,
                      (const double*)en, (double*)h->lufs.p);
            h->last.n_launches++;
        }
        h->last.n_launches += 5;
    }
    PB_CK(pbrt_last_error(), "lufs kernels");
    return PB_OK;
}

// ------------------------------------------------------------------------------------------------ one batch call
struct BatchOut {
    double* median_f0 = nullptr; int32_t* n_voiced = nullptr; int32_t* n_frames = nullptr;   // pitch (all or none)
    double* lufs = nullptr; double* duration_s = nullptr; int32_t* status = nullptr;
    float* frame_f0 = nullptr; float* frame_strength = nullptr; float* frame_intensity = nullptr;   // force a single segment
};

int finish_batch(PbHandle* h);

// Plans the batch, enqueues every copy and kernel and the result download, and returns without waiting for the GPU.  The outputs are
// filled by finish_batch (which the blocking entry points call right away, and pb_extract_wait later).
int submit_batch(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int on_device, const PbUnits* u, const PbPitchParams* p,
                 const uint8_t* want_pitch, const uint8_t* want_lufs, const BatchOut& o) {
    if (h->pending.active) return fail(h, PB_EINVAL, "%s", "a submitted batch is still pending on this handle: call pb_extract_wait first");
    PbRange r_submit("pb: submit batch (plan + enqueue)");
    int rc = validate_units(h, u, pcm_len);
    if (rc != PB_OK) return rc;
    if (o.median_f0) { const char* pe = pitch_params_error(p); if (pe) return fail(h, PB_EINVAL, "%s", pe); }
    pbrt_set_device(h->device);
    begin_call(h);
    // Any early return below may leave copies / kernels in flight that read the caller's (pinned) PCM or the staging
    // buffers: drain all three streams before handing control back on failure.
    struct DrainOnError {
        PbHandle* h; int* rc;
        ~DrainOnError() { if (*rc != PB_OK) { pbrt_stream_sync(h->copy_stream); pbrt_stream_sync(h->stream); pbrt_stream_sync(h->lufs_stream); } }
    };
    int rc_final = PB_ECUDA;                                   // set to PB_OK on the one successful exit
    DrainOnError drain{h, &rc_final};
    h->cur_pcm_len = pcm_len;
    const int64_t n = u->n_units;
    const bool do_pitch = o.median_f0 && o.n_voiced && o.n_frames, do_lufs = o.lufs != nullptr;
    const bool want_frames = o.frame_f0 || o.frame_strength || o.frame_intensity;
    BatchPlan& bp = h->plan;
    bp.reset();
    std::vector<int32_t>& pstat = bp.pstat; std::vector<int32_t>& lflags = bp.lflags;
    pstat.assign((size_t)n, 0); lflags.assign((size_t)n, 0);
    const auto t_plan0 = std::chrono::steady_clock::now();
    static const bool plan_profile = getenv("PB_PLAN_PROFILE") != nullptr;      // development aid: host planning laps on stderr
    auto t_lap = t_plan0;
    auto lap = [&](const char* what) {
        if (!plan_profile) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[pb plan] %-14s %.3f ms\n", what, std::chrono::duration<float, std::milli>(now - t_lap).count());
        t_lap = now;
    };
    // Order of the call (everything the host computes overlaps something the GPU or the copy engine does):
    //   * host PCM: the upload of a first segment starts NOW on the copy stream; both plans are made while it is in flight,
    //     then the rest of the buffer is cut where the PLANNED work says (below) and queued behind it;
    //   * resident PCM: the pitch kernels are launched before the loudness units are even planned;
    //   * descriptors reach the device through a small copy kernel on the compute stream that reads the pinned staging
    //     buffer directly (a DMA copy would queue behind the PCM: the copy engine is FIFO).
    const size_t pcm_bytes = (size_t)pcm_len * 2;
    const bool segmented = !on_device && !want_frames && pcm_bytes >= ((size_t)64 << 20);
    PB_CKMEM(h->med.ensure((size_t)n * 8 + 8) || h->nvoiced.ensure((size_t)n * 4 + 4) || h->lufs.ensure((size_t)n * 8 + 8) ||
             h->stage_out.ensure((size_t)n * 20 + 64), "unit results");
    if (!on_device) PB_CKMEM(h->pcm.ensure(pcm_bytes + 64), "pcm staging");
    const int16_t* d_pcm = on_device ? pcm : (const int16_t*)h->pcm.p;
    std::unique_ptr<ScopedEv> evt(new ScopedEv(h, EV_TOTAL));       // closed once the result download is enqueued
    std::vector<int64_t> seg_end;                                    // exclusive end sample of every PCM segment
    std::vector<pbEvent_t> seg_done;
    constexpr int NB = 512;                                          // planning histogram: work per 1/512 of the buffer
    auto bin_edge = [&](int b) { return b >= NB ? pcm_len : std::min<int64_t>(pcm_len, (((pcm_len * b) / NB + 127) & ~(int64_t)127)); };
    auto enqueue_upload = [&](int64_t a, int64_t b) -> int {
        PbRange r_("pb: enqueue PCM upload segment");
        ScopedEv ev(h, EV_H2D, h->copy_stream);
        if (b > a) PB_CK(pbrt_h2d((char*)h->pcm.p + a * 2, pcm + a, (size_t)(b - a) * 2, h->copy_stream), "pcm upload");
        seg_end.push_back(b);
        seg_done.push_back(*next_event(h));
        pbrt_event_record(&seg_done.back(), h->copy_stream);
        return PB_OK;
    };
    const int first_bins = 2 * NB / 32;                              // ~ what uploads while the host plans
    if (on_device) seg_end.push_back(pcm_len);
    else { rc = enqueue_upload(0, segmented ? bin_edge(first_bins) : pcm_len); if (rc != PB_OK) return rc; }
    if (do_pitch) PB_CK(pbrt_memset(h->med.p, 0, (size_t)n * 8, h->stream) || pbrt_memset(h->nvoiced.p, 0, (size_t)n * 4, h->stream), "memset");
    if (do_lufs) PB_CK(pbrt_memset(h->lufs.p, 0xff, (size_t)n * 8, h->stream), "memset");      // all-ones = NaN
    // pinned staging -> device, by the SMs (16-byte words; the staging buffers are allocated with slack)
    auto upload = [&](void* dst, const void* src_pinned, size_t bytes, pbStream_t stream) {
        if (!bytes) return;
        const long long n16 = (long long)((bytes + 15) / 16);
        const int grid = (int)std::max<long long>(1, std::min<long long>((n16 + 255) / 256, (long long)h->sm_count * 4));
        PB_LAUNCH(pb_copy16_kernel, dim3(grid), dim3(256), 0, stream, (const int4*)pbrt_host_device_ptr(src_pinned), (int4*)dst, n16);
        h->last.n_launches++;
    };
    // the loudness stream starts behind the result memsets; its descriptors and kernels then depend on nothing the pitch
    // kernels do (same-stream order would park them behind the frames kernel)
    if (do_lufs) {
        pbEvent_t cleared = *next_event(h);
        pbrt_event_record(&cleared, h->stream);
        PB_CK(pbrt_stream_wait_event(h->lufs_stream, cleared), "stream wait");
    }
    auto seg_of = [&](int64_t need_end) { int s = 0; const int ns = (int)seg_end.size(); while (s + 1 < ns && need_end > seg_end[(size_t)s]) s++; return s; };

    std::vector<std::vector<int64_t>>& pids = bp.pids; std::vector<std::vector<int64_t>>& lids = bp.lids;
    std::vector<std::vector<PitchLaunch>> pl;
    std::vector<LufsLaunch> ll;
    std::vector<int64_t> frame_off;
    bool pitch_launched = false;

    auto plan_pitch_units = [&]() -> int {
        PbRange r_("pb: plan pitch units (host, float64)");
        memset(o.n_frames, 0, (size_t)n * 4);
        int r = plan_pitch(h, u, p, want_pitch, pstat.data(), o.n_frames, bp);
        lap("plan_pitch");
        return r;
    };
    auto plan_lufs_units = [&]() -> int {
        PbRange r_("pb: plan loudness units (host, float64)");
        int r = plan_lufs(h, u, want_lufs, lflags.data(), bp);
        lap("plan_lufs");
        return r;
    };
    // units -> segments, descriptors into pinned staging, descriptors to the device (the segments are final by now)
    // round 0: every unit (one segment or all segments known); round 1: only the units of the first segment (the others
    // are not cut yet); round 2: the rest.  Descriptors already on the device are not uploaded again.
    size_t su_sent = 0, sp_sent = 0;
    auto stage_upload_pitch = [&](int round) -> int {
        PbRange r_("pb: stage + upload pitch descriptors");
        const int n_seg = (int)seg_end.size();
        if (pids.size() < (size_t)n_seg) pids.resize((size_t)n_seg);
        if (pl.size() < (size_t)n_seg) pl.resize((size_t)n_seg);
        for (int s = (round == 2 ? 1 : 0); s < n_seg; s++) pids[(size_t)s].clear();
        std::vector<int64_t> seg_frames((size_t)n_seg, 0);
        size_t n_pok = 0;
        for (int64_t i = 0; i < n; i++) if (bp.pclass[(size_t)i] >= 0) {
            n_pok++;
            const int64_t need = u->file_off[i] + u->file_nx[i];
            const bool in_first = need <= seg_end[0];
            if (round == 1 && !in_first) continue;
            if (round == 2 && in_first) continue;
            const int s = seg_of(need);
            pids[(size_t)s].push_back(i); seg_frames[(size_t)s] += bp.pplan[(size_t)i].n_frames;
        }
        // per-frame outputs follow the caller's unit order (the layout pb_pitch_plan reports)
        if (want_frames) {
            frame_off.assign((size_t)n + 1, 0);
            int64_t acc = 0;
            for (int64_t i = 0; i < n; i++) { frame_off[(size_t)i] = acc; if (bp.pclass[(size_t)i] >= 0) acc += bp.pplan[(size_t)i].n_frames; }
            frame_off[(size_t)n] = acc;
        }
        // frame arrays are shared by the segments (stream order); when the cuts are not known yet, size them for everything
        const size_t T = round == 0 ? (size_t)*std::max_element(seg_frames.begin(), seg_frames.end()) : (size_t)bp.total_frames;
        const size_t mc = (size_t)(bp.max_cand > 0 ? bp.max_cand : 1);
        const size_t n_groups = 16 * bp.classes.size() + 4;           // launch groups (segments x classes) + alignment slack
        if (n_pok && round != 2)
            PB_CKMEM(h->cand_f.ensure(T * mc * 4) || h->cand_s.ensure(T * mc * 4) || h->ncand.ensure(T) || h->inten.ensure(T * 4) ||
                     h->psi.ensure(T * std::max<size_t>(mc, 8)) ||   /* K3 packs the back-pointers of a frame into one 64-bit word when max_cand <= 16 */
                     h->sel_f.ensure(T * 4) || h->sel_s.ensure(T * 4) ||
                     h->stage_units.ensure(n_pok * sizeof(PbUnitDev) + 16) || h->units.ensure(n_pok * sizeof(PbUnitDev) + 16) ||
                     h->stage_pairs.ensure((n_pok + n_groups) * 4 + 16) || h->pair_off.ensure((n_pok + n_groups) * 4 + 16), "pitch buffers");
        for (int s = (round == 2 ? 1 : 0); s < (round == 1 ? 1 : n_seg); s++) {
            int r = stage_pitch(h, u, bp, pids[(size_t)s], want_frames ? &frame_off : nullptr, pl[(size_t)s]);
            if (r != PB_OK) return r;
        }
        h->sp_off = (h->sp_off + 3) & ~(size_t)3;                     // keep the next round's upload 16-byte aligned
        lap("stage_pitch");
        upload((char*)h->units.p + su_sent * sizeof(PbUnitDev), (const char*)h->stage_units.p + su_sent * sizeof(PbUnitDev), (h->su_off - su_sent) * sizeof(PbUnitDev), h->stream);
        upload((char*)h->pair_off.p + sp_sent * 4, (const char*)h->stage_pairs.p + sp_sent * 4, (h->sp_off - sp_sent) * 4, h->stream);
        su_sent = h->su_off; sp_sent = h->sp_off;
        return PB_OK;
    };
    auto stage_upload_lufs = [&]() -> int {
        PbRange r_("pb: stage + upload loudness descriptors");
        const int n_seg = (int)seg_end.size();
        if (lids.size() < (size_t)n_seg) lids.resize((size_t)n_seg);
        for (auto& v : lids) v.clear();
        ll.assign((size_t)n_seg, LufsLaunch());
        std::vector<int64_t> seg_chunks((size_t)n_seg, 0);
        for (size_t k = 0; k < bp.lunits.size(); k++) {
            const int s = seg_of(bp.lneed[k]);
            lids[(size_t)s].push_back((int64_t)k); seg_chunks[(size_t)s] += bp.lunits[k].n_chunks;
        }
        const size_t CH = (size_t)*std::max_element(seg_chunks.begin(), seg_chunks.end());
        if (!bp.lunits.empty()) {
            PB_CKMEM(h->stage_lunits.ensure(bp.lunits.size() * sizeof(PbLufsUnitDev) + 16) || h->lunits.ensure(bp.lunits.size() * sizeof(PbLufsUnitDev) + 16) ||
                     h->meters.ensure(bp.meters.size() * sizeof(PbMeterDev) + 16) || h->stage_meters.ensure(bp.meters.size() * sizeof(PbMeterDev) + 16) ||
                     h->lstate.ensure(CH * 32) || h->lenergy.ensure(CH * 8), "loudness buffers");
            memcpy(h->stage_meters.p, bp.meters.data(), bp.meters.size() * sizeof(PbMeterDev));
        }
        for (int s = 0; s < n_seg; s++) {
            ll[(size_t)s].off = 0; ll[(size_t)s].m = 0; ll[(size_t)s].chunks = 0;
            if (!lids[(size_t)s].empty()) stage_lufs(h, bp, lids[(size_t)s], ll[(size_t)s]);
        }
        lap("stage_lufs");
        upload(h->lunits.p, h->stage_lunits.p, h->sl_off * sizeof(PbLufsUnitDev), h->lufs_stream);
        upload(h->meters.p, h->stage_meters.p, bp.meters.size() * sizeof(PbMeterDev), h->lufs_stream);
        return PB_OK;
    };

    if (!segmented) {
        // one segment: pitch descriptors up and pitch kernels running before the loudness units are planned
        if (do_pitch) {
            if ((rc = plan_pitch_units()) != PB_OK || (rc = stage_upload_pitch(0)) != PB_OK) return rc;
            if (!on_device) PB_CK(pbrt_stream_wait_event(h->stream, seg_done[0]), "stream wait");
            for (const PitchLaunch& L : pl[0]) { rc = launch_pitch_group(h, d_pcm, p, bp, L); if (rc != PB_OK) return rc; }
            pitch_launched = true;
        }
        if (do_lufs && ((rc = plan_lufs_units()) != PB_OK || (rc = stage_upload_lufs()) != PB_OK)) return rc;
    } else {
        // The pitch units of the first segment are planned, staged and launched as soon as that segment has landed; the
        // loudness plan and everything below happen while they run.  The rest of the buffer is then cut by PLANNED work:
        // a kernel can only start when its segment has landed, so the last segment should carry little work (it lands
        // when the upload ends) and the ones before it similar amounts (fewer, fuller launches than a fixed grid of cuts).
        double t_first_launch_ms = 0.0, work_first_ms = 0.0;
        if (do_pitch) {
            if ((rc = plan_pitch_units()) != PB_OK || (rc = stage_upload_pitch(1)) != PB_OK) return rc;
            PB_CK(pbrt_stream_wait_event(h->stream, seg_done[0]), "stream wait");
            for (const PitchLaunch& L : pl[0]) { rc = launch_pitch_group(h, d_pcm, p, bp, L); if (rc != PB_OK) return rc; }
            pl[0].clear();
            t_first_launch_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_plan0).count();
            for (int64_t i : pids[0]) work_first_ms += 5.2e-6 * (double)bp.pplan[(size_t)i].n_frames;
        }
        if (do_lufs && (rc = plan_lufs_units()) != PB_OK) return rc;
        std::vector<double> work((size_t)NB, 0.0);               // estimated GPU milliseconds per bin (measured rates)
        auto bin_of = [&](int64_t need_end) { int64_t b = need_end > 0 ? ((need_end - 1) * NB) / pcm_len : 0; return (size_t)(b >= NB ? NB - 1 : b); };
        if (do_pitch) for (int64_t i = 0; i < n; i++) if (bp.pclass[(size_t)i] >= 0 && u->file_off[i] + u->file_nx[i] > seg_end[0])
            work[bin_of(u->file_off[i] + u->file_nx[i])] += 5.2e-6 * (double)bp.pplan[(size_t)i].n_frames;
        for (size_t k = 0; k < bp.lunits.size(); k++)
            work[bin_of(bp.lneed[k])] += 2.6e-9 * (double)(bp.lunits[k].b - bp.lunits[k].a + bp.lunits[k].npad);
        double total = 0.0; for (double w : work) total += w;
        int tail = NB;                                           // the last segment: the longest suffix with little work
        { double sfx = 0.0; const double budget = std::max(3.0, 0.05 * total);
          while (tail > first_bins + 1 && sfx + work[(size_t)tail - 1] <= budget) { sfx += work[(size_t)tail - 1]; tail--; } }
        // Between the first and the last segment: every cut goes as far as the upload will have got by the time the GPU
        // runs out of the work it already has (a simulation with the measured rates), at least 1/16 of the buffer each.
        const double ms_per_bin = (double)pcm_bytes / NB / 55e6;     // ~55 GB/s pinned host -> device
        const int min_bins = NB / 16;
        double t_free = std::max(t_first_launch_ms, first_bins * ms_per_bin) + work_first_ms;   // when the GPU has drained segment 0
        for (int last = first_bins; last < tail;) {
            int e = std::min(tail, last + min_bins);
            while (e < tail && (e + 1) * ms_per_bin <= t_free) e++;
            if (tail - e < min_bins) e = tail;
            rc = enqueue_upload(bin_edge(last), bin_edge(e)); if (rc != PB_OK) return rc;
            double w = 0.0; for (int b = last; b < e; b++) w += work[(size_t)b];
            t_free = std::max(t_free, e * ms_per_bin) + w;
            last = e;
        }
        if (tail < NB) { rc = enqueue_upload(bin_edge(tail), pcm_len); if (rc != PB_OK) return rc; }
        lap("segments");
        if (do_pitch && (rc = stage_upload_pitch(2)) != PB_OK) return rc;
        if (do_lufs && (rc = stage_upload_lufs()) != PB_OK) return rc;
    }
    const int n_seg = (int)seg_end.size();
    if (pl.size() < (size_t)n_seg) pl.resize((size_t)n_seg);
    if (ll.size() < (size_t)n_seg) ll.resize((size_t)n_seg, LufsLaunch());
    h->last.n_frames = bp.total_frames; h->last.n_lufs_samples = bp.lufs_samples;
    h->last.host_plan_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_plan0).count();

    {
        // The loudness kernels go to their own stream: they depend only on their descriptors (uploaded on that same
        // stream) and on their PCM segment, and they share the SMs with the latency-bound path finder and
        // the tails of the frames kernel instead of queueing behind them.
        for (int s = 0; s < n_seg; s++) {
            if (!on_device && !(pitch_launched && s == 0)) PB_CK(pbrt_stream_wait_event(h->stream, seg_done[(size_t)s]), "stream wait");
            if (!pitch_launched) for (const PitchLaunch& L : pl[(size_t)s]) { rc = launch_pitch_group(h, d_pcm, p, bp, L); if (rc != PB_OK) return rc; }
            if (do_lufs) {
                if (!on_device) PB_CK(pbrt_stream_wait_event(h->lufs_stream, seg_done[(size_t)s]), "stream wait");
                rc = launch_lufs_group(h, d_pcm, ll[(size_t)s], h->lufs_stream); if (rc != PB_OK) return rc;
            }
        }
        if (do_lufs) {
            pbEvent_t lufs_done = *next_event(h);
            pbrt_event_record(&lufs_done, h->lufs_stream);
            PB_CK(pbrt_stream_wait_event(h->stream, lufs_done), "stream wait");
        }
        ScopedEv evd(h, EV_D2H);
        char* so = (char*)h->stage_out.p;
        if (do_pitch) PB_CK(pbrt_d2h(so, h->med.p, (size_t)n * 8, h->stream) || pbrt_d2h(so + (size_t)n * 8, h->nvoiced.p, (size_t)n * 4, h->stream), "result download");
        if (do_lufs) PB_CK(pbrt_d2h(so + (size_t)n * 12, h->lufs.p, (size_t)n * 8, h->stream), "result download");
        if (want_frames && bp.total_frames > 0) {
            const size_t fb = (size_t)bp.total_frames * 4;
            if (o.frame_f0) PB_CK(pbrt_d2h(o.frame_f0, h->sel_f.p, fb, h->stream), "frame download");
            if (o.frame_strength) PB_CK(pbrt_d2h(o.frame_strength, h->sel_s.p, fb, h->stream), "frame download");
            if (o.frame_intensity) PB_CK(pbrt_d2h(o.frame_intensity, h->inten.p, fb, h->stream), "frame download");
        }
    }
    evt.reset();
    // host arithmetic overlaps the GPU work
    if (o.duration_s) for (int64_t i = 0; i < n; i++) {
        int st; o.duration_s[i] = pb_part_duration(u->file_nx[i], u->rate[i], u->has_t1[i], u->t0[i], u->t1[i], &st);
    }
    h->pending.active = true; h->pending.n = n; h->pending.do_pitch = do_pitch; h->pending.do_lufs = do_lufs;
    h->pending.median_f0 = o.median_f0; h->pending.n_voiced = o.n_voiced; h->pending.lufs = o.lufs; h->pending.status = o.status;
    rc_final = PB_OK;
    return PB_OK;
}

// Waits for the submitted batch and hands its results to the caller's arrays.
int finish_batch(PbHandle* h) {
    if (!h->pending.active) return fail(h, PB_EINVAL, "%s", "no submitted batch to wait for");
    PbRange r_wait("pb: wait for batch + copy results out");
    h->pending.active = false;
    pbrt_set_device(h->device);
    int rc = PB_OK;
    if (pbrt_stream_sync(h->stream) | pbrt_stream_sync(h->copy_stream) | pbrt_stream_sync(h->lufs_stream))
        rc = fail(h, PB_ECUDA, "%s", (std::string("stream sync: ") + pbrt_error()).c_str());
    if (rc != PB_OK) return rc;
    end_call(h);
    const BatchPlan& bp = h->plan;
    const int64_t n = h->pending.n;
    const char* so = (const char*)h->stage_out.p;
    if (h->pending.do_pitch) { memcpy(h->pending.median_f0, so, (size_t)n * 8); memcpy(h->pending.n_voiced, so + (size_t)n * 8, (size_t)n * 4); }
    if (h->pending.do_lufs) {
        memcpy(h->pending.lufs, so + (size_t)n * 12, (size_t)n * 8);
        for (auto& d : bp.dups) h->pending.lufs[d.first] = h->pending.lufs[d.second];
    }
    if (h->pending.status) for (int64_t i = 0; i < n; i++) h->pending.status[i] = bp.pstat[(size_t)i] | bp.lflags[(size_t)i];
    return PB_OK;
}

int run_batch(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int on_device, const PbUnits* u, const PbPitchParams* p,
              const uint8_t* want_pitch, const uint8_t* want_lufs, const BatchOut& o) {
    const int rc = submit_batch(h, pcm, pcm_len, on_device, u, p, want_pitch, want_lufs, o);
    return rc != PB_OK ? rc : finish_batch(h);
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int pb_abi_version(void) { return PB_ABI_VERSION; }

int pb_create(int device, PbHandle** out) {
    if (!out) return PB_EINVAL;
    *out = nullptr;
    const int ndev = pbrt_device_count();
    if (ndev <= 0 || device < 0 || device >= ndev) return PB_ENODEVICE;
    if (pbrt_set_device(device)) return PB_ENODEVICE;
    PbHandle* h = new PbHandle();
    h->device = device;
    if (pbrt_props(device, &h->sm_count, &h->cc_major, &h->cc_minor, &h->total_mem)) { delete h; return PB_ENODEVICE; }
    if (pbrt_stream_create(&h->own_stream) || pbrt_stream_create(&h->copy_stream) || pbrt_stream_create(&h->lufs_stream)) { delete h; return PB_ECUDA; }
    h->stream = h->own_stream;
    memset(&h->last, 0, sizeof h->last);
    *out = h;
    return PB_OK;
}

void pb_destroy(PbHandle* h) {
    if (!h) return;
    pbrt_set_device(h->device);
    pbrt_stream_sync(h->stream);
    pbrt_stream_sync(h->copy_stream);
    pbrt_stream_sync(h->lufs_stream);
    DevBuf* dbs[] = {&h->pcm, &h->units, &h->pair_off, &h->cand_f, &h->cand_s, &h->ncand, &h->inten, &h->psi, &h->sel_f, &h->sel_s, &h->stats_longs, &h->stats_ctr, &h->lufs_longs, &h->lufs_jobs, &h->lufs_ctr, &h->lufs_P, &h->lufs_Z, &h->lufs_S, &h->path_longs, &h->path_jobs, &h->path_ctr, &h->path_T, &h->path_D, &h->path_maps, &h->path_exits,
                     &h->med, &h->nvoiced, &h->lunits, &h->meters, &h->lstate, &h->lenergy, &h->lufs, &h->pairpos, &h->racf, &h->slot_fr, &h->work_ctr};
    for (auto* b : dbs) b->release();
    HostBuf* hbs[] = {&h->stage_units, &h->stage_pairs, &h->stage_lunits, &h->stage_meters, &h->stage_out};
    for (auto* b : hbs) b->release();
    for (auto& kv : h->tables) { kv.second->window.release(); kv.second->inv_wr.release(); kv.second->tw_a.release(); kv.second->tw_b.release(); kv.second->half_tab.release(); delete kv.second; }
    for (auto& e : h->ev_pool) pbrt_event_destroy(e);
    pbrt_stream_destroy(h->own_stream);
    pbrt_stream_destroy(h->copy_stream);
    pbrt_stream_destroy(h->lufs_stream);
    delete h;
}

const char* pb_last_error(const PbHandle* h) { return h ? h->err.c_str() : "no handle"; }

int pb_set_stream(PbHandle* h, void* stream) {
    if (!h) return PB_EINVAL;
    h->stream = stream ? (pbStream_t)stream : h->own_stream;
    return PB_OK;
}

int pb_get_timings(const PbHandle* h, PbTimings* out) {
    if (!h || !out) return PB_EINVAL;
    *out = h->last;
    return PB_OK;
}

int pb_device_info(const PbHandle* h, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* total_mem) {
    if (!h) return PB_EINVAL;
    if (sm_count) *sm_count = h->sm_count;
    if (cc_major) *cc_major = h->cc_major;
    if (cc_minor) *cc_minor = h->cc_minor;
    if (total_mem) *total_mem = h->total_mem;
    return PB_OK;
}

int pb_median_pitch_batch(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int pcm_on_device, const PbUnits* u, const PbPitchParams* p,
                          double* median_f0, int32_t* n_voiced, int32_t* n_frames, int32_t* status,
                          float* frame_f0, float* frame_strength, float* frame_intensity) {
    if (!h) return PB_EINVAL;
    if (!pcm || !p || !median_f0 || !n_voiced || !n_frames || !status) return fail(h, PB_EINVAL, "%s", "null argument");
    BatchOut o;
    o.median_f0 = median_f0; o.n_voiced = n_voiced; o.n_frames = n_frames; o.status = status;
    o.frame_f0 = frame_f0; o.frame_strength = frame_strength; o.frame_intensity = frame_intensity;
    return run_batch(h, pcm, pcm_len, pcm_on_device, u, p, nullptr, nullptr, o);
}

int pb_lufs_batch(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int pcm_on_device, const PbUnits* u, double* lufs, int32_t* status) {
    if (!h) return PB_EINVAL;
    if (!pcm || !lufs || !status) return fail(h, PB_EINVAL, "%s", "null argument");
    BatchOut o;
    o.lufs = lufs; o.status = status;
    PbPitchParams p; pb_pitch_params_default(&p);
    return run_batch(h, pcm, pcm_len, pcm_on_device, u, &p, nullptr, nullptr, o);
}

int pb_extract_batch(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int pcm_on_device, const PbUnits* u, const PbPitchParams* p,
                     const uint8_t* want_pitch, const uint8_t* want_lufs,
                     double* median_f0, int32_t* n_voiced, int32_t* n_frames, double* lufs, double* duration_s, int32_t* status) {
    if (!h) return PB_EINVAL;
    if (!pcm || !p || !status) return fail(h, PB_EINVAL, "%s", "null argument");
    BatchOut o;
    o.median_f0 = median_f0; o.n_voiced = n_voiced; o.n_frames = n_frames; o.lufs = lufs; o.duration_s = duration_s; o.status = status;
    return run_batch(h, pcm, pcm_len, pcm_on_device, u, p, want_pitch, want_lufs, o);
}

int pb_extract_submit(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int pcm_on_device, const PbUnits* u, const PbPitchParams* p,
                      const uint8_t* want_pitch, const uint8_t* want_lufs,
                      double* median_f0, int32_t* n_voiced, int32_t* n_frames, double* lufs, double* duration_s, int32_t* status) {
    if (!h) return PB_EINVAL;
    if (!pcm || !p || !status) return fail(h, PB_EINVAL, "%s", "null argument");
    BatchOut o;
    o.median_f0 = median_f0; o.n_voiced = n_voiced; o.n_frames = n_frames; o.lufs = lufs; o.duration_s = duration_s; o.status = status;
    return submit_batch(h, pcm, pcm_len, pcm_on_device, u, p, want_pitch, want_lufs, o);
}

int pb_extract_wait(PbHandle* h) {
    if (!h) return PB_EINVAL;
    return finish_batch(h);
}

#include "pb_api_host.inc"
#include "pb_api_next.inc"
#include "pb_api_silence.inc"
#include "pb_textgrid.inc"
#include "pb_ssml.inc"

}  // extern "C"
