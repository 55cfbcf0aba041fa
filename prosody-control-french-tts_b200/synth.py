"""Synthetic French-speech-shaped audio and word grids.

Benchmark / test data only — there is no network for real corpora.  Everything is seeded and written with torch
ops so the same generator runs on the CPU (tests, small) and on the GPU (bench, 10^9 samples).

Signal: glottal-pulse-like harmonic source  sum_k a_k sin(k * phase), a_k ~ 1/k, with an f0 contour =
speaker median (uniform in log-frequency over [90, 260] Hz) x declination x 4-6 Hz vibrato; voiced stretches
alternate with unvoiced (shaped noise, -25 dB) and silent (-60 dB) ones; peak 0.7 full scale; int16 mono.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def _segment_labels(n_utt: int, n_samp: int, sr: float, gen: torch.Generator, device) -> torch.Tensor:
    """-> int8 [n_utt, n_samp]: 0 silence, 1 voiced, 2 unvoiced. Random alternation with speech-like durations."""
    max_seg = int(n_samp / sr / 0.08) + 4
    u = torch.rand(n_utt, max_seg, generator=gen, device=device)
    kind = torch.rand(n_utt, max_seg, generator=gen, device=device)
    # voiced runs U(0.08, 0.6) s; unvoiced U(0.04, 0.18) s; silences U(0.05, 0.5) s
    lab = torch.where(kind < 0.62, 1, torch.where(kind < 0.85, 2, 0)).to(torch.int8)
    dur = torch.where(lab == 1, 0.08 + 0.52 * u, torch.where(lab == 2, 0.04 + 0.14 * u, 0.05 + 0.45 * u))
    ends = torch.cumsum(dur, dim=1) * sr
    t = torch.arange(n_samp, device=device, dtype=torch.float32).expand(n_utt, n_samp).contiguous()
    idx = torch.searchsorted(ends.contiguous(), t).clamp_(max=max_seg - 1)
    return torch.gather(lab, 1, idx)


def make_corpus(n_utt: int, dur_s: float, sr: int, seed: int = 1234, device="cpu", chunk: int = 256,
                n_harm: int = 30, out: torch.Tensor | None = None) -> torch.Tensor:
    """-> int16 tensor [n_utt, n_samp] on `device` (or written into `out`)."""
    device = torch.device(device)
    n_samp = int(round(dur_s * sr))
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    pcm = out if out is not None else torch.empty(n_utt, n_samp, dtype=torch.int16, device=device)
    t = torch.arange(n_samp, device=device, dtype=torch.float32) / sr
    for s in range(0, n_utt, chunk):
        m = min(chunk, n_utt - s)
        med = 90.0 * torch.exp(torch.rand(m, 1, generator=gen, device=device) * math.log(260.0 / 90.0))
        decl = 1.0 - 0.12 * (t / max(dur_s, 1e-3)) * torch.rand(m, 1, generator=gen, device=device)
        vib_f = 4.0 + 2.0 * torch.rand(m, 1, generator=gen, device=device)
        vib_d = 0.01 + 0.03 * torch.rand(m, 1, generator=gen, device=device)
        slow = 0.08 * torch.sin(2 * math.pi * (0.3 + 0.5 * torch.rand(m, 1, generator=gen, device=device)) * t +
                                6.28 * torch.rand(m, 1, generator=gen, device=device))
        f0 = med * decl * (1.0 + vib_d * torch.sin(2 * math.pi * vib_f * t)) * (1.0 + slow)
        phase = 2 * math.pi * torch.cumsum(f0.double() / sr, dim=1)
        phase = torch.remainder(phase, 2 * math.pi).float()
        x = torch.zeros(m, n_samp, device=device)
        tilt = 0.8 + 0.6 * torch.rand(m, 1, generator=gen, device=device)
        for k in range(1, n_harm + 1):
            ok = (f0 * k) < 0.45 * sr
            x += torch.where(ok, torch.sin(k * phase + 0.7 * k) / (k ** tilt), torch.zeros((), device=device))
        lab = _segment_labels(m, n_samp, float(sr), gen, device)
        noise = torch.randn(m, n_samp, generator=gen, device=device)
        # crude spectral shaping of the noise: first difference (high-pass) mixed with the raw noise
        shaped = 0.6 * noise + 0.4 * torch.diff(noise, dim=1, prepend=noise[:, :1])
        amp = 0.4 + 0.6 * torch.rand(m, 1, generator=gen, device=device)
        env = 0.75 + 0.25 * torch.sin(2 * math.pi * 2.3 * t + 6.28 * torch.rand(m, 1, generator=gen, device=device))
        sig = torch.where(lab == 1, x * env, torch.where(lab == 2, shaped * 10 ** (-25 / 20) * 2.0, noise * 10 ** (-60 / 20)))
        sig = sig + 10 ** (-55 / 20) * noise
        peak = sig.abs().amax(dim=1, keepdim=True).clamp_min(1e-9)
        sig = sig / peak * (0.7 * amp)
        pcm[s:s + m] = torch.round(sig * 32767.0).clamp_(-32768, 32767).to(torch.int16)
    return pcm


def make_word_grid(n_utt: int, dur_s: float, seed: int = 0):
    """Deterministic word / pause grid per utterance, TextGrid-tier-0 style: list of lists of (tmin, tmax, mark).
    Words U(0.12, 0.55) s, a pause after a word with p = 0.18 lasting U(0.05, 0.9) s, a sentence end every 8-20 words
    (the mark then ends with '.').  Marks are French-looking tokens; empty mark = silence."""
    rng = np.random.default_rng(seed)
    lexicon = ["le", "chat", "mange", "une", "pomme", "dans", "jardin", "et", "puis", "il", "regarde", "les", "oiseaux",
               "qui", "chantent", "sur", "toit", "de", "maison", "bleue", "avec", "son", "ami", "fidèle", "pendant", "que"]
    grids = []
    for _ in range(n_utt):
        t, out, since = 0.0, [], 0
        next_end = int(rng.integers(8, 21))
        lead = float(rng.uniform(0.0, 0.3))
        if lead > 0.02:
            out.append((0.0, round(lead, 3), ""))
            t = round(lead, 3)
        while t < dur_s - 0.15:
            w = float(rng.uniform(0.12, 0.55))
            e = min(round(t + w, 3), round(dur_s, 3))
            word = lexicon[int(rng.integers(len(lexicon)))]
            since += 1
            if since >= next_end:
                word += "."
                since, next_end = 0, int(rng.integers(8, 21))
            out.append((t, e, word))
            t = e
            if rng.random() < 0.18 and t < dur_s - 0.2:
                p = float(rng.uniform(0.05, 0.9))
                e = min(round(t + p, 3), round(dur_s, 3))
                out.append((t, e, ""))
                t = e
        end = round(dur_s, 3)
        if end - t >= 0.06:
            out.append((t, end, ""))
        elif t < end and out:
            # a shorter tail would be a slice Praat refuses (< 3 / pitch_floor s aborts the reference step): extend instead
            a, _, mark = out[-1]
            out[-1] = (a, end, mark)
        grids.append(out)
    return grids
