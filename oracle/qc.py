"""Oracle: the reference's per-file quality-control scalars.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/pipeline1.py:151-186 (written to output_results/per_file_analysis.csv at :402-422):

    snr_db(y)                      :151-165  frames of 400 / hop 160 (librosa.util.frame: no padding), energy per
                                             frame, noise = frames below the 25th percentile, 10 log10(mean / (noise + 1e-10))
    spectral_flatness_mean(y)      :168-174  mean over frames of librosa.feature.spectral_flatness(S=|stft(y)|)
                                             (n_fft 2048, hop 512, amin 1e-10, power 2)
    high_freq_energy_ratio(y, sr)  :177-186  sum |rfft(y)|^2 above 4 kHz / (sum |rfft(y)|^2 + 1e-10), full-length FFT

All three are float32 computations in the reference (y is float32 and numpy keeps it); the restatement keeps the
dtypes.  Pinned against the *_after columns of per_file_analysis.csv for the committed clear_audio WAVs
(tests/golden/ref_qc_after.npz, tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np

from .features import power_spectrogram

SR = 16000
SNR_FRAME = int(0.025 * SR)      # 400
SNR_HOP = int(0.010 * SR)        # 160


def snr_db(y) -> float:
    if y is None or len(y) < SNR_FRAME:
        return 0.0
    y = np.asarray(y)
    n_frames = 1 + (len(y) - SNR_FRAME) // SNR_HOP
    idx = np.arange(SNR_FRAME)[:, None] + SNR_HOP * np.arange(n_frames)[None, :]
    frames = y[idx]                                            # [400, n_frames] like librosa.util.frame
    energy = np.sum(frames ** 2, axis=0)
    noise_mask = energy < np.percentile(energy, 25)
    if noise_mask.sum() == 0:
        return 0.0
    noise_power = np.mean(energy[noise_mask])
    signal_power = np.mean(energy)
    return float(10.0 * np.log10(signal_power / (noise_power + 1e-10)))


def spectral_flatness_mean(y) -> float:
    try:
        y = np.asarray(y, dtype=np.float32)
        if y.size == 0 or not np.isfinite(y).all():
            raise ValueError("invalid audio")                  # librosa.util.valid_audio
        S = np.sqrt(power_spectrogram(y))                      # np.abs(librosa.stft(y)), float32 [1025, T]
        S_thresh = np.maximum(np.float32(1e-10), S ** 2.0)
        gmean = np.exp(np.mean(np.log(S_thresh), axis=-2, keepdims=True))
        amean = np.mean(S_thresh, axis=-2, keepdims=True)
        return float(np.mean(gmean / amean))
    except Exception:
        return 0.0


def high_freq_energy_ratio(y, sr: int = SR) -> float:
    try:
        y = np.asarray(y)
        fft = np.fft.rfft(y)
        freqs = np.fft.rfftfreq(len(y), 1.0 / sr)
        high_mask = freqs > 4000
        total_energy = np.sum(np.abs(fft) ** 2)
        high_energy = np.sum(np.abs(fft[high_mask]) ** 2)
        return float(high_energy / (total_energy + 1e-10))
    except Exception:
        return 0.0


def qc_metrics(y) -> np.ndarray:
    """[snr_db, spectral_flatness_mean, high_freq_energy_ratio] as float64[3]."""
    return np.array([snr_db(y), spectral_flatness_mean(y), high_freq_energy_ratio(y, SR)], dtype=np.float64)
