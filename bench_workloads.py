"""Synthetic workloads of bench.py and scripts/flip_listing.py, one per BASELINE.json config:

    c1  the ten repo clips (44.1 kHz real speech) + seed-0 TextGrids + 16 kHz x0.93 raw twins, reference pitch parameters
    c2  10 000 x 5 s @ 16 kHz, floor 75 / ceiling 600, word grids + paired 4.65 s raw-synth stream            (the bench default)
    c3  2 000 x 20 s @ 24 kHz, MFA-style word grids, paired 18.6 s raw-synth stream, full SSML-delta output
    c4  N x 1 h @ 22.05 kHz: silence segmentation on the GPU, per-segment analysis, and the unsegmented hour
    c5  100 h mixed corpus (60 % 16 kHz 3-12 s, 30 % 24 kHz 10-30 s, 10 % 44.1 kHz 20-40 s), length-balanced over the ranks

Everything is seeded; audio is generated on the GPU (prosody_b200.synth).  A workload is a list of step.Segment over ONE
int16 buffer plus the prosody / pitch settings; `shard(rank, world)` cuts it for strong scaling.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

FLOOR75 = dict(pitch_floor=75.0, pitch_ceiling=600.0)


@dataclass
class Workload:
    name: str
    description: str
    pcm: object                 # torch int16 tensor on the device (natural + synthetic audio, concatenated)
    segments: list              # prosody_b200.step.Segment
    pitch: dict
    prosody: dict
    audio_s: float              # natural audio seconds in `segments`
    total_audio_s: float = 0.0  # natural audio seconds of the whole corpus (strong scaling: all ranks)
    extra: dict = field(default_factory=dict)


def _segments_uniform(n_utt, dur, syn_dur, sr, seed, device, grid_seed=None):
    """n_utt natural utterances of `dur` s followed by their synthetic twins of `syn_dur` s, all at `sr`."""
    import torch
    from prosody_b200 import step as S
    from prosody_b200 import synth
    nat_n, syn_n = int(round(dur * sr)), int(round(syn_dur * sr))
    pcm = torch.empty(n_utt * (nat_n + syn_n), dtype=torch.int16, device=device)
    synth.make_corpus(n_utt, dur, sr, seed=seed, device=device, out=pcm[:n_utt * nat_n].view(n_utt, nat_n))
    synth.make_corpus(n_utt, syn_dur, sr, seed=seed + 7919, device=device, out=pcm[n_utt * nat_n:].view(n_utt, syn_n))
    grids = synth.make_word_grid(n_utt, dur, seed=seed if grid_seed is None else grid_seed)
    base = n_utt * nat_n
    segs = [S.Segment(f"segment_ph{i + 1}", i * nat_n, nat_n, sr, grids[i], base + i * syn_n, syn_n, sr) for i in range(n_utt)]
    return pcm, segs


def c2(device, n_utt=10000, seed=1234):
    from prosody_b200 import step as S
    pcm, segs = _segments_uniform(n_utt, 5.0, 4.65, 16000, seed, device)
    return Workload("c2", f"{n_utt} synthetic 5 s utterances per GPU, 16 kHz mono s16, F0 75-600 Hz, 10 ms hop, word grids + paired 4.65 s raw-synth stream",
                    pcm, segs, dict(FLOOR75), dict(S.DEFAULT_PROSODY), n_utt * 5.0, n_utt * 5.0)


def c3(device, n_utt=2000, seed=2345):
    from prosody_b200 import step as S
    pcm, segs = _segments_uniform(n_utt, 20.0, 18.6, 24000, seed, device)
    prosody = dict(S.DEFAULT_PROSODY); prosody.update(pitch_semitones=1.3, smoothing_alpha=0.2, end_punctuation_pause_ms=400, baseline_window=50)
    return Workload("c3", f"{n_utt} synthetic 20 s utterances per GPU, 24 kHz mono s16 (Azure-TTS-shaped), MFA-style word intervals, paired 18.6 s raw-synth "
                          f"stream, F0 75-600 Hz, sliding 50-segment baselines, full SSML-delta output (strings built every step)",
                    pcm, segs, dict(FLOOR75), prosody, n_utt * 20.0, n_utt * 20.0)


def c1(device):
    """The reference's own clips; file reading is part of the timed step (bench.py), so only paths are prepared here."""
    import sys
    import tempfile
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent / "tests" / "golden"))
    import make_c1_fixture as C1
    root = Path(tempfile.mkdtemp(prefix="c1_voice_"))
    v = C1.build_voice(root)
    audio_s = sum(len(s[1]) / s[2] for s in v["segments"])
    from prosody_b200 import step as S
    prosody = dict(S.DEFAULT_PROSODY); prosody.update(pitch_semitones=1.3, smoothing_alpha=0.2, end_punctuation_pause_ms=400)
    return Workload("c1", "the reference's ten example clips (Data/voice/records/audio, 44.1 kHz mono s16, 161.85 s of real speech) through the FILE-LEVEL "
                          "drop-in: WAV + TextGrid read from disk, three CSVs written, reference pitch parameters (floor 150 / ceiling 600)",
                    None, [], dict(S.REFERENCE_PITCH), prosody, audio_s, audio_s, extra=dict(voice=v, root=root))


# ---------------------------------------------------------------------------------------------------- c5: mixed corpus
C5_MIX = ((16000, 3, 12, 0.60), (24000, 10, 30, 0.30), (44100, 20, 40, 0.10))      # rate, min s, max s, share of the hours


def c5_catalogue(hours=100.0, seed=4567):
    """The corpus as a list of (rate, duration_s, bucket_index, row_in_bucket): integer-second lengths, uniform within each class.
    Buckets = (rate, duration): every bucket is one synth.make_corpus call with its own seed."""
    rng = np.random.default_rng(seed)
    utts, buckets = [], {}
    for rate, lo, hi, share in C5_MIX:
        target = hours * 3600.0 * share
        acc = 0.0
        while acc < target:
            d = int(rng.integers(lo, hi + 1))
            key = (rate, d)
            row = buckets.get(key, 0)
            buckets[key] = row + 1
            utts.append((rate, d, key, row))
            acc += d
    return utts, buckets


def c5(device, rank=0, world=1, hours=100.0, seed=4567):
    """This rank's share of the mixed corpus: utterances dealt to ranks by length-balanced bucketing on their frame counts
    (shard.partition_by_cost), PCM generated bucket by bucket (identical for every world size), only the owned rows kept."""
    import torch
    from prosody_b200 import shard
    from prosody_b200 import step as S
    from prosody_b200 import synth
    utts, buckets = c5_catalogue(hours, seed)
    # cost = pitch frames of the whole file + its syntagme slices ~ 2 x duration / hop (hop 10 ms at floor 75), times the FFT size class
    fft_cost = {16000: 1.0, 24000: 2.2, 44100: 4.6}
    costs = [fft_cost[r] * d for r, d, _, _ in utts]
    owners = shard.partition_by_cost(costs, world)
    mine = owners[rank]
    # rows of every bucket this rank owns
    need = {}
    for i in mine:
        need.setdefault(utts[i][2], []).append((utts[i][3], i))
    total = 0
    for (rate, d), rows in need.items():
        total += len(rows) * (int(round(d * rate)) + int(round(0.93 * d * rate)))
    pcm = torch.empty(total, dtype=torch.int16, device=device)
    segs, off, audio_s = [], 0, 0.0
    for bi, key in enumerate(sorted(need)):
        rate, d = key
        rows = sorted(need[key])
        n_b = buckets[key]
        bseed = seed + 101 * rate // 1000 + 7 * d
        nat = synth.make_corpus(n_b, float(d), rate, seed=bseed, device=device)
        syn = synth.make_corpus(n_b, 0.93 * d, rate, seed=bseed + 7919, device=device)
        grids = synth.make_word_grid(n_b, float(d), seed=bseed)
        sel = torch.tensor([r for r, _ in rows], device=device)
        nn, sn = nat.shape[1], syn.shape[1]
        m = len(rows)
        pcm[off:off + m * nn].view(m, nn).copy_(nat.index_select(0, sel))
        pcm[off + m * nn:off + m * (nn + sn)].view(m, sn).copy_(syn.index_select(0, sel))
        for k, (r, gi) in enumerate(rows):
            segs.append((gi, S.Segment(f"segment_ph{gi + 1}", off + k * nn, nn, rate, grids[r], off + m * nn + k * sn, sn, rate)))
            audio_s += d
        off += m * (nn + sn)
        del nat, syn
    segs.sort(key=lambda t: t[0])                       # global utterance order (the EMA runs over rows in segment order)
    total_audio = float(sum(d for _, d, _, _ in utts))
    load = [sum(costs[i] for i in o) for o in owners]
    return Workload("c5", f"{hours:g} h mixed synthetic corpus ({len(utts)} utterances: 60 % 16 kHz 3-12 s, 30 % 24 kHz 10-30 s, 10 % 44.1 kHz 20-40 s, paired "
                          f"x0.93 raw-synth stream), F0 75-600 Hz, utterances dealt to {world} rank(s) by length-balanced bucketing on estimated frame cost",
                    pcm, [s for _, s in segs], dict(FLOOR75), dict(S.DEFAULT_PROSODY), audio_s, total_audio,
                    extra=dict(global_ids=[g for g, _ in segs], n_utts_total=len(utts), planned_load=load,
                               planned_imbalance=max(load) / (sum(load) / len(load)) if load else 1.0))


# ---------------------------------------------------------------------------------------------------- c4: long-form
def c4_recordings(device, n_hours=4, seed=3456):
    """n_hours one-hour 22.05 kHz recordings (720 x 5 s of speech-shaped audio each) with a 1.1-2.3 s pause every 19 s."""
    import torch
    from prosody_b200 import synth
    sr = 22050
    per = 3600 * sr
    pcm = torch.empty(n_hours * per, dtype=torch.int16, device=device)
    for h in range(n_hours):
        x = pcm[h * per:(h + 1) * per]
        synth.make_corpus(720, 5.0, sr, seed=seed + h, device=device, out=x.view(720, 5 * sr))
        t = np.arange(0, 3600, 19.0)
        for k, a in enumerate(t[1:]):
            i0 = int(a * sr); i1 = i0 + int((1.1 + 0.1 * (k % 13)) * sr)
            x[i0:i1] = (x[i0:i1].float() * 0.004).to(torch.int16)
    return pcm, sr, per
