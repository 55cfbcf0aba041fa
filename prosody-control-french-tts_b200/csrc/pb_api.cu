// pb_api.cu — the C ABI of libprosody_b200.so (see include/prosody_b200.h): handle, device buffers, per-geometry
// tables, host-side planning and kernel orchestration.  Built by nvcc for sm_100a; the same source is compiled by
// g++ against the SIMT emulator for CPU-side tests only (tests/simt_emu).
#include "../../include/prosody_b200.h"
#include "pb_rt.h"
#include "pb_plan.h"
#include "pb_pitch.cuh"
#include "pb_pitch_frames.cuh"
#include "pb_pitch_path.cuh"
#include "pb_lufs.cuh"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
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
    DevBuf window, inv_wr, tw_a, tw_b;
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

}  // namespace

struct PbHandle {
    int device = 0;
    int sm_count = 1, cc_major = 0, cc_minor = 0;
    long long total_mem = 0;
    pbStream_t own_stream = 0, stream = 0;
    std::string err;
    // device buffers (grow on demand)
    DevBuf pcm, units, pair_off, cand_f, cand_s, ncand, inten, psi, sel_f, sel_s, med, nvoiced;
    DevBuf lunits, meters, lstate, lenergy, lufs;
    HostBuf stage_units, stage_pairs, stage_lunits, stage_meters, stage_out;
    std::map<std::pair<int64_t, int>, PitchTables*> tables;
    std::vector<EvPair> evs;
    std::vector<pbEvent_t> ev_pool;
    size_t ev_used = 0;
    PbTimings last;
    std::map<int, int> occ_cache;
    std::vector<std::pair<int64_t, int64_t>> lufs_dups;   // (unit, unit it duplicates): filled per call by enqueue_lufs
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
struct ScopedEv {     // records an event pair around a section of the stream
    PbHandle* h; size_t idx;
    ScopedEv(PbHandle* h_, int kind) : h(h_) {
        EvPair p; p.kind = kind; p.a = *next_event(h); p.b = *next_event(h);
        pbrt_event_record(&p.a, h->stream);
        h->evs.push_back(p); idx = h->evs.size() - 1;
    }
    ~ScopedEv() { pbrt_event_record(&h->evs[idx].b, h->stream); }
};
enum { EV_TOTAL = 0, EV_H2D, EV_STATS, EV_FRAMES, EV_PATH, EV_LUFS, EV_INTENSITY, EV_D2H };

void begin_call(PbHandle* h) {
    h->evs.clear(); h->ev_used = 0;
    memset(&h->last, 0, sizeof h->last);
}
void end_call(PbHandle* h) {   // after the stream has been synchronised
    float* f = &h->last.total_ms;
    for (auto& p : h->evs) f[p.kind] += pbrt_event_ms(p.a, p.b);
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
    PitchTables* t = new PitchTables();
    if (t->window.ensure(wf.size() * 4) || t->inv_wr.ensure(iw.size() * 4) || t->tw_a.ensure(ta.size() * 8) || t->tw_b.ensure(tb.size() * 8)) {
        delete t; return fail(h, PB_ENOMEM, "out of memory: %s", "pitch tables");
    }
    // pageable -> device copies of a few KB; synchronous on purpose (tables are built once per geometry)
    pbrt_h2d(t->window.p, wf.data(), wf.size() * 4, h->stream);
    pbrt_h2d(t->inv_wr.p, iw.data(), iw.size() * 4, h->stream);
    pbrt_h2d(t->tw_a.p, ta.data(), ta.size() * 8, h->stream);
    pbrt_h2d(t->tw_b.p, tb.data(), tb.size() * 8, h->stream);
    pbrt_stream_sync(h->stream);
    h->tables[key] = t;
    *out = t;
    return PB_OK;
}

// ------------------------------------------------------------------------------------------------ launches
template <int LOG2N>
int launch_frames(PbHandle* h, const int16_t* d_pcm, const PbUnitDev* d_units, const int32_t* d_pair_off, const PbPitchGeomDev& gm,
                  float* cand_f, float* cand_s, uint8_t* ncand, float* inten) {
    typedef PbFftCfg<LOG2N> C;
    const size_t smem = (size_t)C::GROUPS_PER_CTA * (C::BUF + 8 * C::G) * sizeof(float2);
    const int threads = C::WARPS_PER_CTA * 32;
    auto kfn = pb_pitch_frames_kernel<LOG2N>;
    int per_sm = 2;
#ifndef PB_SIMT_EMU
    auto oc = h->occ_cache.find(LOG2N);
    if (oc == h->occ_cache.end()) {
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return fail(h, PB_ECUDA, "cudaFuncSetAttribute: %s", pbrt_error());
        cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        h->occ_cache[LOG2N] = per_sm;
    } else per_sm = oc->second;
#endif
    long long need = ((long long)gm.n_pairs + C::GROUPS_PER_CTA - 1) / C::GROUPS_PER_CTA;
    long long cap = (long long)h->sm_count * per_sm;
    int grid = (int)std::max(1LL, std::min(need, cap));
    PB_LAUNCH(kfn, dim3(grid), dim3(threads), smem, h->stream, d_pcm, d_units, d_pair_off, gm, cand_f, cand_s, ncand, inten);
    h->last.n_launches++;
    return PB_OK;
}

struct PitchClass {
    PbGeomHost g; int gstatus = 0;
    std::vector<int> ids;       // caller indices of the OK units
};

// Stage host PCM on the device if needed. Returns the device pointer in *d_pcm.
int stage_pcm(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int on_device, const int16_t** d_pcm) {
    if (on_device) { *d_pcm = pcm; return PB_OK; }
    PB_CKMEM(h->pcm.ensure((size_t)pcm_len * 2 + 64), "pcm staging");
    ScopedEv ev(h, EV_H2D);
    PB_CK(pbrt_h2d(h->pcm.p, pcm, (size_t)pcm_len * 2, h->stream), "pcm upload");
    *d_pcm = (const int16_t*)h->pcm.p;
    return PB_OK;
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

// Enqueue the whole F0 path for the wanted units. Results land in h->med / h->nvoiced (device, caller-indexed)
// and per-frame selected values in h->sel_f / h->sel_s / h->inten.
int enqueue_pitch(PbHandle* h, const int16_t* d_pcm, const PbUnits* u, const PbPitchParams* p, const uint8_t* want,
                  int32_t* status, int32_t* n_frames, std::vector<int64_t>& frame_off, int64_t* total_frames_out) {
    const int64_t n = u->n_units;
    std::map<double, PitchClass> classes;
    std::vector<PbUnitPlan> plans((size_t)n);
    frame_off.assign((size_t)n + 1, 0);
    int64_t total = 0; int max_cand = 0;
    for (int64_t i = 0; i < n; i++) {
        frame_off[(size_t)i] = total;
        PbUnitPlan& pl = plans[(size_t)i];
        pl.status = PB_UNIT_OK; pl.n_frames = 0;
        if (want && !want[i]) { status[i] = PB_UNIT_OK; n_frames[i] = 0; continue; }
        auto it = classes.find(u->rate[i]);
        if (it == classes.end()) {
            PitchClass pc; pc.gstatus = pb_geom_for_rate(u->rate[i], *p, pc.g);
            it = classes.insert(std::make_pair(u->rate[i], pc)).first;
        }
        PitchClass& pc = it->second;
        pb_plan_pitch_unit(u->file_nx[i], u->rate[i], u->has_t1[i], u->t0[i], u->t1[i], *p, pc.g, pc.gstatus, pl);
        status[i] = pl.status; n_frames[i] = pl.n_frames;
        if (pl.status == PB_UNIT_OK) { pc.ids.push_back((int)i); total += pl.n_frames; max_cand = pc.g.max_cand; }
    }
    frame_off[(size_t)n] = total;
    *total_frames_out = total;
    h->last.n_frames += total;
    PB_CKMEM(h->med.ensure((size_t)n * 8 + 8) || h->nvoiced.ensure((size_t)n * 4 + 4), "unit results");
    PB_CK(pbrt_memset(h->med.p, 0, (size_t)n * 8, h->stream) || pbrt_memset(h->nvoiced.p, 0, (size_t)n * 4, h->stream), "memset");
    if (total == 0) return PB_OK;
    if (max_cand > PB_MAXC) return fail(h, PB_EUNSUPPORTED, "%s", "pitch_ceiling / pitch_floor exceeds 32 candidates per frame");
    const size_t T = (size_t)total;
    PB_CKMEM(h->cand_f.ensure(T * max_cand * 4) || h->cand_s.ensure(T * max_cand * 4) || h->ncand.ensure(T) ||
             h->inten.ensure(T * 4) || h->psi.ensure(T * max_cand) || h->sel_f.ensure(T * 4) || h->sel_s.ensure(T * 4), "frame arrays");
    // stage all classes' descriptors in one pinned buffer each
    size_t n_ok = 0; for (auto& kv : classes) n_ok += kv.second.ids.size();
    PB_CKMEM(h->stage_units.ensure(n_ok * sizeof(PbUnitDev)) || h->stage_pairs.ensure((n_ok + classes.size()) * 4) ||
             h->units.ensure(n_ok * sizeof(PbUnitDev)) || h->pair_off.ensure((n_ok + classes.size()) * 4), "unit descriptors");
    size_t uoff = 0, poff = 0;
    for (auto& kv : classes) {
        PitchClass& pc = kv.second;
        const size_t m = pc.ids.size();
        if (!m) continue;
        PitchTables* tb = nullptr;
        int rc = get_tables(h, pc.g, &tb);
        if (rc != PB_OK) return rc;
        PbUnitDev* su = (PbUnitDev*)h->stage_units.p + uoff;
        int32_t* sp = (int32_t*)h->stage_pairs.p + poff;
        int64_t pairs = 0;
        for (size_t k = 0; k < m; k++) {
            const int i = pc.ids[k];
            const PbUnitPlan& pl = plans[(size_t)i];
            PbUnitDev& d = su[k];
            d.pcm_off = u->file_off[i]; d.ix1 = pl.ix1; d.nx = pl.nx; d.frame_off = frame_off[(size_t)i];
            d.x1 = pl.x1; d.t1 = pl.t1; d.mean = 0.0; d.global_peak = 0.0;
            d.file_nx = (int32_t)u->file_nx[i]; d.n_frames = pl.n_frames; d.pair_off = (int32_t)pairs; d.out_index = i;
            sp[k] = (int32_t)pairs;
            pairs += (pl.n_frames + 1) / 2;
            if (pairs > 0x7ffffff0LL) return fail(h, PB_EUNSUPPORTED, "%s", "more than 2^31 frame pairs in one call; split the batch");
        }
        sp[m] = (int32_t)pairs;
        PbUnitDev* du = (PbUnitDev*)h->units.p + uoff;
        int32_t* dp = (int32_t*)h->pair_off.p + poff;
        {
            ScopedEv ev(h, EV_H2D);
            PB_CK(pbrt_h2d(du, su, m * sizeof(PbUnitDev), h->stream) || pbrt_h2d(dp, sp, (m + 1) * 4, h->stream), "descriptor upload");
        }
        PbPitchGeomDev gm;
        memset(&gm, 0, sizeof gm);
        gm.nw = (int)pc.g.nw; gm.half_nw = (int)pc.g.half_nw; gm.nsamp_period = (int)pc.g.nsamp_period; gm.half_period = (int)pc.g.half_period;
        gm.brent_ixmax = (int)pc.g.brent_ixmax;
        gm.scan_lim = (int)std::min(pc.g.max_lag, pc.g.brent_ixmax);
        gm.max_cand = pc.g.max_cand; gm.n_units = (int)m; gm.n_pairs = (int)pairs;
        // a maximum at lag i refines to a lag <= i+1: below this lag its frequency stays above the ceiling (never voiced)
        gm.min_refine_lag = (int)std::floor(1.0 / pc.g.dx / pc.g.ceiling) - 1;
        gm.sr = (float)(1.0 / pc.g.dx); gm.half_voicing = (float)(0.5 * p->voicing_threshold);
        gm.octave_cost = (float)p->octave_cost; gm.min_pitch = (float)p->pitch_floor;
        gm.dx = pc.g.dx; gm.dt = pc.g.dt; gm.ceiling = pc.g.ceiling; gm.silence_threshold = p->silence_threshold;
        gm.voicing_threshold = p->voicing_threshold; gm.octave_cost_d = p->octave_cost; gm.octave_jump_cost = p->octave_jump_cost;
        gm.voiced_unvoiced_cost = p->voiced_unvoiced_cost;
        gm.window = (const float*)tb->window.p; gm.inv_wr = (const float*)tb->inv_wr.p;
        gm.tw_a = (const float2*)tb->tw_a.p; gm.tw_b = (const float2*)tb->tw_b.p;
        {
            ScopedEv ev(h, EV_STATS);
            int grid = (int)std::min<size_t>(m, (size_t)h->sm_count * 8);
            PB_LAUNCH(pb_unit_stats_kernel, dim3(grid), dim3(256), 0, h->stream, d_pcm, du, (int)m);
            h->last.n_launches++;
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
            PB_LAUNCH(pb_pitch_path_kernel, dim3(grid), dim3(wpb * 32), 0, h->stream, du, gm, (const float*)h->cand_f.p,
                      (const float*)h->cand_s.p, (const uint8_t*)h->ncand.p, (const float*)h->inten.p, (uint8_t*)h->psi.p,
                      (float*)h->sel_f.p, (float*)h->sel_s.p, (double*)h->med.p, (int32_t*)h->nvoiced.p);
            h->last.n_launches++;
        }
        PB_CK(pbrt_last_error(), "pitch kernels");
        uoff += m; poff += m + 1;
    }
    return PB_OK;
}

// Enqueue the loudness path. Results land in h->lufs (device, caller-indexed; NaN where the reference would raise).
int enqueue_lufs(PbHandle* h, const int16_t* d_pcm, const PbUnits* u, const uint8_t* want, int32_t* status_flags) {
    const int64_t n = u->n_units;
    std::map<double, int> meter_ix;
    std::vector<PbMeterDev> meters;
    PB_CKMEM(h->stage_lunits.ensure((size_t)n * sizeof(PbLufsUnitDev) + 64), "lufs descriptors");
    PbLufsUnitDev* su = (PbLufsUnitDev*)h->stage_lunits.p;
    size_t m = 0; int64_t chunks = 0, samples = 0;
    // Units that resolve to the same samples and meter have the same loudness (the reference's < 0.4 s / empty-slice
    // fallbacks send every short syntagme of a file to that file's whole-file value): measure once, copy on the host.
    h->lufs_dups.clear();
    std::unordered_map<LufsKey, int64_t, LufsKeyHash> seen;
    seen.reserve((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        status_flags[i] = 0;
        if (want && !want[i]) continue;
        const double mr = u->meter_rate ? u->meter_rate[i] : u->rate[i];
        int64_t a, b, npad;
        const int st = pb_lufs_resolve(u->file_nx[i], u->rate[i], mr, u->has_t1[i], u->t0[i], u->t1[i], &a, &b, &npad);
        status_flags[i] = st;
        if (st & (PB_UNIT_LUFS_ERROR | PB_UNIT_SLICE_ERROR)) continue;
        {
            const LufsKey key{u->file_off[i] + a, b - a, npad, mr};
            auto ins = seen.emplace(key, i);
            if (!ins.second) { h->lufs_dups.push_back(std::make_pair(i, ins.first->second)); continue; }
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
            meters.push_back(md);
            it = meter_ix.insert(std::make_pair(mr, (int)meters.size() - 1)).first;
        }
        PbLufsUnitDev& d = su[m++];
        d.pcm_off = u->file_off[i]; d.a = a; d.b = b; d.npad = npad; d.chunk_off = chunks;
        const int64_t len = b - a + npad;
        d.n_blocks = (int32_t)pb_lufs_num_blocks(len, mr); d.n_chunks = d.n_blocks + 3;
        d.meter = it->second; d.out_index = (int)i; d.inv_peak = 1.0;
        chunks += d.n_chunks; samples += len;
    }
    h->last.n_lufs_samples += samples;
    PB_CKMEM(h->lufs.ensure((size_t)n * 8 + 8), "lufs results");
    PB_CK(pbrt_memset(h->lufs.p, 0xff, (size_t)n * 8, h->stream), "memset");      // all-ones = NaN
    if (!m) return PB_OK;
    PB_CKMEM(h->lunits.ensure(m * sizeof(PbLufsUnitDev)) || h->meters.ensure(meters.size() * sizeof(PbMeterDev)) ||
             h->stage_meters.ensure(meters.size() * sizeof(PbMeterDev)) || h->lstate.ensure((size_t)chunks * 32) ||
             h->lenergy.ensure((size_t)chunks * 8), "lufs buffers");
    memcpy(h->stage_meters.p, meters.data(), meters.size() * sizeof(PbMeterDev));
    {
        ScopedEv ev(h, EV_H2D);
        PB_CK(pbrt_h2d(h->lunits.p, su, m * sizeof(PbLufsUnitDev), h->stream) ||
              pbrt_h2d(h->meters.p, h->stage_meters.p, meters.size() * sizeof(PbMeterDev), h->stream), "lufs descriptor upload");
    }
    {
        ScopedEv ev(h, EV_LUFS);
        PbLufsUnitDev* du = (PbLufsUnitDev*)h->lunits.p;
        const PbMeterDev* dm = (const PbMeterDev*)h->meters.p;
        double* st = (double*)h->lstate.p; double* en = (double*)h->lenergy.p;
        int g1 = (int)std::min<size_t>(m, (size_t)h->sm_count * 8);
        PB_LAUNCH(pb_lufs_peak_kernel, dim3(g1), dim3(256), 0, h->stream, d_pcm, du, (int)m);
        int gc = (int)std::max<int64_t>(1, std::min<int64_t>((chunks + 127) / 128, (int64_t)h->sm_count * 16));
        auto k0 = pb_lufs_chunk_kernel<false>; auto k1 = pb_lufs_chunk_kernel<true>;
        PB_LAUNCH(k0, dim3(gc), dim3(128), 0, h->stream, d_pcm, (const PbLufsUnitDev*)du, (int)m, dm, (long long)chunks, st, en);
        int gu = (int)std::max<size_t>(1, std::min<size_t>((m + 127) / 128, (size_t)h->sm_count * 16));
        PB_LAUNCH(pb_lufs_scan_kernel, dim3(gu), dim3(128), 0, h->stream, (const PbLufsUnitDev*)du, (int)m, dm, st);
        PB_LAUNCH(k1, dim3(gc), dim3(128), 0, h->stream, d_pcm, (const PbLufsUnitDev*)du, (int)m, dm, (long long)chunks, st, en);
        PB_LAUNCH(pb_lufs_gate_kernel, dim3(gu), dim3(128), 0, h->stream, (const PbLufsUnitDev*)du, (int)m, dm, (const double*)en, (double*)h->lufs.p);
        h->last.n_launches += 5;
    }
    PB_CK(pbrt_last_error(), "lufs kernels");
    return PB_OK;
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
    if (pbrt_stream_create(&h->own_stream)) { delete h; return PB_ECUDA; }
    h->stream = h->own_stream;
    memset(&h->last, 0, sizeof h->last);
    *out = h;
    return PB_OK;
}

void pb_destroy(PbHandle* h) {
    if (!h) return;
    pbrt_set_device(h->device);
    pbrt_stream_sync(h->stream);
    DevBuf* dbs[] = {&h->pcm, &h->units, &h->pair_off, &h->cand_f, &h->cand_s, &h->ncand, &h->inten, &h->psi, &h->sel_f, &h->sel_s,
                     &h->med, &h->nvoiced, &h->lunits, &h->meters, &h->lstate, &h->lenergy, &h->lufs};
    for (auto* b : dbs) b->release();
    HostBuf* hbs[] = {&h->stage_units, &h->stage_pairs, &h->stage_lunits, &h->stage_meters, &h->stage_out};
    for (auto* b : hbs) b->release();
    for (auto& kv : h->tables) { kv.second->window.release(); kv.second->inv_wr.release(); kv.second->tw_a.release(); kv.second->tw_b.release(); delete kv.second; }
    for (auto& e : h->ev_pool) pbrt_event_destroy(e);
    pbrt_stream_destroy(h->own_stream);
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

void pb_pitch_params_default(PbPitchParams* p) {
    if (!p) return;
    p->time_step = 0.0; p->pitch_floor = 150.0; p->pitch_ceiling = 600.0;     // Code/audioPipeline.py:329,332
    p->periods_per_window = 3.0; p->silence_threshold = 0.03; p->voicing_threshold = 0.45;
    p->octave_cost = 0.01; p->octave_jump_cost = 0.35; p->voiced_unvoiced_cost = 0.14;
    p->max_candidates = 15; p->reserved = 0;
}

int pb_pitch_plan(const PbPitchParams* p, const PbUnits* u, int32_t* status, int32_t* n_frames, int64_t* frame_off) {
    if (!p || !u || !status || !n_frames) return PB_EINVAL;
    std::map<double, std::pair<PbGeomHost, int>> geoms;
    int64_t total = 0;
    for (int64_t i = 0; i < u->n_units; i++) {
        auto it = geoms.find(u->rate[i]);
        if (it == geoms.end()) {
            PbGeomHost g; int st = pb_geom_for_rate(u->rate[i], *p, g);
            it = geoms.insert(std::make_pair(u->rate[i], std::make_pair(g, st))).first;
        }
        PbUnitPlan pl;
        pb_plan_pitch_unit(u->file_nx[i], u->rate[i], u->has_t1[i], u->t0[i], u->t1[i], *p, it->second.first, it->second.second, pl);
        status[i] = pl.status; n_frames[i] = pl.n_frames;
        if (frame_off) frame_off[i] = total;
        if (pl.status == PB_UNIT_OK) total += pl.n_frames;
    }
    if (frame_off) frame_off[u->n_units] = total;
    return PB_OK;
}

int pb_part_duration_batch(const PbUnits* u, double* duration_s, int32_t* status) {
    if (!u || !duration_s) return PB_EINVAL;
    for (int64_t i = 0; i < u->n_units; i++) {
        int st;
        duration_s[i] = pb_part_duration(u->file_nx[i], u->rate[i], u->has_t1[i], u->t0[i], u->t1[i], &st);
        if (status) status[i] = st;
    }
    return PB_OK;
}

int pb_median_pitch_batch(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int pcm_on_device, const PbUnits* u, const PbPitchParams* p,
                          double* median_f0, int32_t* n_voiced, int32_t* n_frames, int32_t* status,
                          float* frame_f0, float* frame_strength, float* frame_intensity) {
    if (!h) return PB_EINVAL;
    if (!pcm || !p || !median_f0 || !n_voiced || !n_frames || !status) return fail(h, PB_EINVAL, "%s", "null argument");
    int rc = validate_units(h, u, pcm_len);
    if (rc != PB_OK) return rc;
    pbrt_set_device(h->device);
    begin_call(h);
    const int64_t n = u->n_units;
    std::vector<int64_t> frame_off; int64_t total = 0;
    {
        ScopedEv evt(h, EV_TOTAL);
        const int16_t* d_pcm = nullptr;
        rc = stage_pcm(h, pcm, pcm_len, pcm_on_device, &d_pcm);
        if (rc != PB_OK) return rc;
        rc = enqueue_pitch(h, d_pcm, u, p, nullptr, status, n_frames, frame_off, &total);
        if (rc != PB_OK) return rc;
        ScopedEv evd(h, EV_D2H);
        PB_CKMEM(h->stage_out.ensure((size_t)n * 12 + 64), "result staging");
        PB_CK(pbrt_d2h(h->stage_out.p, h->med.p, (size_t)n * 8, h->stream) ||
              pbrt_d2h((char*)h->stage_out.p + (size_t)n * 8, h->nvoiced.p, (size_t)n * 4, h->stream), "result download");
        if (total > 0) {
            if (frame_f0) PB_CK(pbrt_d2h(frame_f0, h->sel_f.p, (size_t)total * 4, h->stream), "frame download");
            if (frame_strength) PB_CK(pbrt_d2h(frame_strength, h->sel_s.p, (size_t)total * 4, h->stream), "frame download");
            if (frame_intensity) PB_CK(pbrt_d2h(frame_intensity, h->inten.p, (size_t)total * 4, h->stream), "frame download");
        }
    }
    PB_CK(pbrt_stream_sync(h->stream), "stream sync");
    end_call(h);
    memcpy(median_f0, h->stage_out.p, (size_t)n * 8);
    memcpy(n_voiced, (char*)h->stage_out.p + (size_t)n * 8, (size_t)n * 4);
    return PB_OK;
}

int pb_lufs_batch(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int pcm_on_device, const PbUnits* u, double* lufs, int32_t* status) {
    if (!h) return PB_EINVAL;
    if (!pcm || !lufs || !status) return fail(h, PB_EINVAL, "%s", "null argument");
    int rc = validate_units(h, u, pcm_len);
    if (rc != PB_OK) return rc;
    pbrt_set_device(h->device);
    begin_call(h);
    const int64_t n = u->n_units;
    {
        ScopedEv evt(h, EV_TOTAL);
        const int16_t* d_pcm = nullptr;
        rc = stage_pcm(h, pcm, pcm_len, pcm_on_device, &d_pcm);
        if (rc != PB_OK) return rc;
        rc = enqueue_lufs(h, d_pcm, u, nullptr, status);
        if (rc != PB_OK) return rc;
        ScopedEv evd(h, EV_D2H);
        PB_CKMEM(h->stage_out.ensure((size_t)n * 8 + 64), "result staging");
        PB_CK(pbrt_d2h(h->stage_out.p, h->lufs.p, (size_t)n * 8, h->stream), "result download");
    }
    PB_CK(pbrt_stream_sync(h->stream), "stream sync");
    end_call(h);
    memcpy(lufs, h->stage_out.p, (size_t)n * 8);
    for (auto& d : h->lufs_dups) lufs[d.first] = lufs[d.second];
    return PB_OK;
}

int pb_extract_batch(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int pcm_on_device, const PbUnits* u, const PbPitchParams* p,
                     const uint8_t* want_pitch, const uint8_t* want_lufs,
                     double* median_f0, int32_t* n_voiced, int32_t* n_frames, double* lufs, double* duration_s, int32_t* status) {
    if (!h) return PB_EINVAL;
    if (!pcm || !p || !status) return fail(h, PB_EINVAL, "%s", "null argument");
    int rc = validate_units(h, u, pcm_len);
    if (rc != PB_OK) return rc;
    pbrt_set_device(h->device);
    begin_call(h);
    const int64_t n = u->n_units;
    const bool do_pitch = median_f0 && n_voiced && n_frames, do_lufs = lufs != nullptr;
    std::vector<int64_t> frame_off; int64_t total = 0;
    std::vector<int32_t> lflags((size_t)n, 0), pstat((size_t)n, 0);
    {
        ScopedEv evt(h, EV_TOTAL);
        const int16_t* d_pcm = nullptr;
        rc = stage_pcm(h, pcm, pcm_len, pcm_on_device, &d_pcm);
        if (rc != PB_OK) return rc;
        if (do_pitch) { rc = enqueue_pitch(h, d_pcm, u, p, want_pitch, pstat.data(), n_frames, frame_off, &total); if (rc != PB_OK) return rc; }
        if (do_lufs) { rc = enqueue_lufs(h, d_pcm, u, want_lufs, lflags.data()); if (rc != PB_OK) return rc; }
        ScopedEv evd(h, EV_D2H);
        PB_CKMEM(h->stage_out.ensure((size_t)n * 20 + 64), "result staging");
        char* so = (char*)h->stage_out.p;
        if (do_pitch) PB_CK(pbrt_d2h(so, h->med.p, (size_t)n * 8, h->stream) || pbrt_d2h(so + (size_t)n * 8, h->nvoiced.p, (size_t)n * 4, h->stream), "result download");
        if (do_lufs) PB_CK(pbrt_d2h(so + (size_t)n * 12, h->lufs.p, (size_t)n * 8, h->stream), "result download");
    }
    // host arithmetic overlaps the GPU work
    if (duration_s) for (int64_t i = 0; i < n; i++) {
        int st; duration_s[i] = pb_part_duration(u->file_nx[i], u->rate[i], u->has_t1[i], u->t0[i], u->t1[i], &st);
    }
    PB_CK(pbrt_stream_sync(h->stream), "stream sync");
    end_call(h);
    const char* so = (const char*)h->stage_out.p;
    if (do_pitch) { memcpy(median_f0, so, (size_t)n * 8); memcpy(n_voiced, so + (size_t)n * 8, (size_t)n * 4); }
    if (do_lufs) {
        memcpy(lufs, so + (size_t)n * 12, (size_t)n * 8);
        for (auto& d : h->lufs_dups) lufs[d.first] = lufs[d.second];
    }
    for (int64_t i = 0; i < n; i++) status[i] = pstat[(size_t)i] | lflags[(size_t)i];
    return PB_OK;
}

int pb_intensity_plan(const PbUnits* u, double minimum_pitch, double time_step, int32_t* status, int32_t* n_frames,
                      int64_t* frame_off, double* t_first, double* dt_out) {
    (void)u; (void)minimum_pitch; (void)time_step; (void)status; (void)n_frames; (void)frame_off; (void)t_first; (void)dt_out;
    return PB_EUNSUPPORTED;
}
int pb_intensity_batch(PbHandle* h, const int16_t* pcm, int64_t pcm_len, int pcm_on_device, const PbUnits* u, double minimum_pitch,
                       double time_step, int subtract_mean, float* intensity_db, int32_t* status) {
    (void)pcm; (void)pcm_len; (void)pcm_on_device; (void)u; (void)minimum_pitch; (void)time_step; (void)subtract_mean; (void)intensity_db; (void)status;
    return fail(h, PB_EUNSUPPORTED, "%s", "pb_intensity_batch is not wired yet");
}

int pb_syntagme_deltas(int64_t n, const double* p_nat, const double* base_f0, const double* base_loud, const double* l_syn,
                       const int32_t* word_count, const double* nat_total_s, const double* syn_total_s, const int32_t* pause_ms,
                       const PbDeltaParams* prm, double* raw_pitch, double* raw_volume, double* raw_rate) {
    if (n < 0 || !prm || (n && (!p_nat || !base_f0 || !base_loud || !l_syn || !word_count || !nat_total_s || !syn_total_s ||
                                !pause_ms || !raw_pitch || !raw_volume || !raw_rate))) return PB_EINVAL;
    // np.clip(x, lo, hi) == minimum(maximum(x, lo), hi), NaN-propagating
    auto clip = [](double x, double lo, double hi) { if (x != x) return x; x = x < lo ? lo : x; return x > hi ? hi : x; };
    for (int64_t i = 0; i < n; i++) {
        const double pause_s = (double)pause_ms[i] / 1000.0;
        double d_nat = nat_total_s[i] - pause_s; if (!(d_nat > 1e-4)) d_nat = 1e-4;      // max(x, 1e-4)
        double d_syn = syn_total_s[i] - pause_s; if (!(d_syn > 1e-4)) d_syn = 1e-4;
        double p_pct = 0.0;
        if (p_nat[i] > 0.0) {
            double st = 12.0 * log2(p_nat[i] / base_f0[i]);
            st = clip(st, -prm->pitch_semitones * prm->pitch_lower_clip_factor, prm->pitch_semitones);
            p_pct = (pow(2.0, st / 12.0) - 1.0) * 100.0;
        }
        const double db_diff = base_loud[i] - l_syn[i];
        double v_pct = (pow(10.0, db_diff / 20.0) - 1.0) * 100.0;
        v_pct = clip(v_pct, -prm->volume_pct, prm->volume_pct);
        double rp = 0.0;
        if (word_count[i] > 0) {
            const double nat_r = (double)word_count[i] / d_nat, syn_r = (double)word_count[i] / d_syn;
            rp = (nat_r - syn_r) / syn_r * 100.0;
        }
        const double length_s = d_nat;
        double slow = 1.0, fast = 1.0;
        if (!(length_s <= 1.0)) { slow = pow(length_s, 1.5); fast = sqrt(length_s); }
        rp = rp < 0.0 ? rp * slow : rp / fast;
        double over = length_s - prm->threshold_duration_before_slowing_down; if (!(over > 0.0)) over = 0.0;
        rp = rp - over * prm->slow_floor_per_sec;
        const double lo = length_s > 5.0 ? prm->rate_percent * 1.5 : prm->rate_percent;
        const double hi = length_s > 5.0 ? prm->rate_percent * 0.5 : prm->rate_percent;
        rp = clip(rp, -lo, hi);
        raw_pitch[i] = p_pct; raw_volume[i] = v_pct; raw_rate[i] = rp;
    }
    return PB_OK;
}

int pb_ema_clamp(const double* x, int64_t n, double alpha, double max_jump, double* out) {
    if (n < 0 || (n && (!x || !out))) return PB_EINVAL;
    if (n == 0) return PB_OK;
    const double beta = 1.0 - alpha;
    double s = x[0];
    out[0] = s;
    for (int64_t i = 1; i < n; i++) { s = alpha * x[i] + beta * s; out[i] = s; }
    for (int64_t i = 1; i < n; i++) {
        const double d = out[i] - out[i - 1];
        if (fabs(d) > max_jump) out[i] = out[i - 1] + (d > 0.0 ? 1.0 : -1.0) * max_jump;
    }
    return PB_OK;
}

}  // extern "C"
