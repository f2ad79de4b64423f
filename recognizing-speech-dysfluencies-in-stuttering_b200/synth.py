"""Deterministic synthetic UCLASS-style clips.

Shared workload generator for the parity tests, ``bench.py`` and ``smoke()``.  Clip
``i`` depends only on ``seed + i`` so every rank / process / oracle run can rebuild
exactly the same float32 samples without any file or network access.
"""
from __future__ import annotations

import numpy as np

SR = 16000
CLIP_SAMPLES = 3 * SR           # BASELINE.json: "synthetic 16 kHz 3-s clips"
BASE_SEED = 1234


def synth_clip(i: int, n: int = CLIP_SAMPLES, seed: int = BASE_SEED) -> np.ndarray:
    """Voiced harmonic source (f0 ~ U(90,300) Hz, 8 harmonics, 1/k roll-off, random
    phase) x raised-cosine syllable gate (U(2,6) Hz, duty 0.6 -> silent gaps exercise the
    spectral gate and the top-dB clamp) + white noise (sigma 0.01), peak 0.5, float32."""
    rng = np.random.default_rng(seed + i)
    t = np.arange(n, dtype=np.float64) / SR
    f0 = rng.uniform(90.0, 300.0)
    phases = rng.uniform(0.0, 2.0 * np.pi, size=8)
    voiced = np.zeros(n, dtype=np.float64)
    for k in range(1, 9):
        voiced += np.sin(2.0 * np.pi * f0 * k * t + phases[k - 1]) / k
    rate = rng.uniform(2.0, 6.0)
    gate_phase = rng.uniform(0.0, 1.0)
    u = np.mod(rate * t + gate_phase, 1.0)
    gate = np.where(u < 0.6, 0.5 - 0.5 * np.cos(2.0 * np.pi * u / 0.6), 0.0)
    y = voiced * gate + rng.normal(0.0, 0.01, size=n)
    peak = np.max(np.abs(y)) if n else 0.0
    if peak > 0:
        y = y * (0.5 / peak)
    return y.astype(np.float32)


def synth_batch(count: int, n: int = CLIP_SAMPLES, first: int = 0, seed: int = BASE_SEED) -> np.ndarray:
    """float32 [count, n]: clips first .. first+count-1."""
    out = np.empty((count, n), dtype=np.float32)
    for j in range(count):
        out[j] = synth_clip(first + j, n, seed)
    return out


def edge_clips(seed: int = BASE_SEED) -> list[tuple[str, np.ndarray]]:
    """The edge set of SURVEY.md 8d appended to config 1."""
    cases = []
    for n in (4095, 4096, 47999, 48001, 160000):
        cases.append((f"len{n}", synth_clip(100000 + n, n, seed)))
    cases.append(("all_zero", np.zeros(CLIP_SAMPLES, dtype=np.float32)))
    cases.append(("dc_0.25", np.full(CLIP_SAMPLES, 0.25, dtype=np.float32)))
    t = np.arange(CLIP_SAMPLES)
    cases.append(("square_200Hz", np.where((t // 40) % 2 == 0, 1.0, -1.0).astype(np.float32)))
    imp = np.zeros(CLIP_SAMPLES, dtype=np.float32)
    imp[CLIP_SAMPLES // 2] = 1.0
    cases.append(("impulse", imp))
    return cases
