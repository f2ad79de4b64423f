"""Oracle: the 149-dim feature function of the reference.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/pipeline1.py:206-265 (``extract_audio_features``,
``extract_text_features``, ``extract_features``; identical copy at
main1.py:665-715).  The reference passes no DSP parameters, so the arithmetic is
librosa's defaults (librosa >= 0.10 semantics: zero ``pad_mode``); each helper
below names the librosa routine it restates.  Only numpy + scipy are used.

Pinned against the reference's own artefacts by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import functools

import numpy as np
import scipy.fftpack

SR = 16000
N_FFT = 2048
HOP = 512
N_BINS = N_FFT // 2 + 1          # 1025
N_MELS = 128
N_MFCC = 20                      # pipeline1.py:79  MFCC_N
N_CHROMA = 12
AUDIO_FEATURE_LEN = 144          # pipeline1.py:84
TEXT_FEATURE_LEN = 5             # pipeline1.py:85
TOTAL_FEATURE_LEN = 149          # pipeline1.py:86
DELTA_WIDTH = 9


def n_frames(n: int) -> int:
    """librosa.stft(center=True): T = 1 + n // hop."""
    return 1 + n // HOP


# ----------------------------------------------------------------------------
# A.1  STFT -> power            (librosa.core.spectrum._spectrogram / stft)
# ----------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def hann_periodic(n: int) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True) in float64."""
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


def stft_frames(y: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    """Centre-pad n_fft//2 zeros each side and cut frames; returns [T, n_fft] (y's dtype)."""
    pad = n_fft // 2
    yp = np.zeros(len(y) + 2 * pad, dtype=y.dtype)
    yp[pad:pad + len(y)] = y
    T = 1 + len(y) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(T)[:, None]
    return yp[idx]


def power_spectrogram(y: np.ndarray) -> np.ndarray:
    """|stft|^2 as float32 [1025, T].

    librosa multiplies the float32 frames by the float64 window (-> float64), runs
    rfft in float64, stores complex64, then ``np.abs(D) ** 2`` in float32.
    """
    frames = stft_frames(np.asarray(y, dtype=np.float32), N_FFT, HOP)
    D = np.fft.rfft(hann_periodic(N_FFT)[None, :] * frames, axis=1).astype(np.complex64)
    S = np.abs(D) ** 2.0
    return np.ascontiguousarray(S.T.astype(np.float32))


# ----------------------------------------------------------------------------
# A.2  mel filterbank + dB      (librosa.filters.mel, librosa.power_to_db)
# ----------------------------------------------------------------------------
def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if f.ndim:
        m = f >= min_log_hz
        mels[m] = min_log_mel + np.log(f[m] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    log_t = m >= min_log_mel
    freqs[log_t] = min_log_hz * np.exp(logstep * (m[log_t] - min_log_mel))
    return freqs


@functools.lru_cache(maxsize=None)
def mel_filterbank() -> np.ndarray:
    """librosa.filters.mel(sr=16000, n_fft=2048, n_mels=128, fmin=0, fmax=8000,
    htk=False, norm='slaney', dtype=float32) -> float32 [128, 1025]."""
    weights = np.zeros((N_MELS, N_BINS), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(N_FFT, 1.0 / SR)
    mel_pts = np.linspace(_hz_to_mel(0.0), _hz_to_mel(SR / 2.0), N_MELS + 2)
    mel_f = _mel_to_hz(mel_pts)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(N_MELS):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:N_MELS + 2] - mel_f[:N_MELS])
    weights *= enorm[:, np.newaxis]          # float32 *= float64, rounded back to float32
    return weights


def mel_power(S: np.ndarray) -> np.ndarray:
    """einsum('...ft,mf->...mt') in float32 -> [128, T]."""
    return np.einsum("ft,mf->mt", S, mel_filterbank(), optimize=True)


def power_to_db(M: np.ndarray) -> np.ndarray:
    """librosa.power_to_db(ref=1.0, amin=1e-10, top_db=80.0); the max is over the WHOLE clip."""
    log_spec = 10.0 * np.log10(np.maximum(np.float32(1e-10), M))
    log_spec = log_spec.astype(np.float32)
    return np.maximum(log_spec, log_spec.max() - np.float32(80.0))


# ----------------------------------------------------------------------------
# A.3  DCT-II ortho, first 20   (scipy.fftpack.dct inside librosa.feature.mfcc)
# ----------------------------------------------------------------------------
def mfcc_from_logmel(L: np.ndarray) -> np.ndarray:
    return scipy.fftpack.dct(L, axis=0, type=2, norm="ortho")[:N_MFCC]


@functools.lru_cache(maxsize=None)
def dct_matrix() -> np.ndarray:
    """Explicit ortho DCT-II rows 0..19 (float64 [20,128]); used to cross-check and by table tests."""
    m = np.arange(N_MELS, dtype=np.float64)
    k = np.arange(N_MFCC, dtype=np.float64)[:, None]
    C = np.cos(np.pi * k * (2.0 * m + 1.0) / (2.0 * N_MELS)) * np.sqrt(2.0 / N_MELS)
    C[0] *= np.sqrt(0.5)
    return C


# ----------------------------------------------------------------------------
# A.4  delta / delta-delta      (librosa.feature.delta = savgol_filter, mode='interp')
# ----------------------------------------------------------------------------
DELTA1_TAPS = np.arange(-4, 5, dtype=np.float64) / 60.0
DELTA2_TAPS = np.array([28, 7, -8, -17, -20, -17, -8, 7, 28], dtype=np.float64) / 462.0


def delta(x: np.ndarray, order: int) -> np.ndarray:
    """Savitzky-Golay derivative, window 9, polyorder=order, edges = polynomial fit
    (for polyorder==deriv the fitted derivative is constant, i.e. the first/last
    interior value is replicated).  Raises like librosa when T < 9."""
    T = x.shape[-1]
    if DELTA_WIDTH > T:
        raise ValueError(
            f"when mode='interp', width={DELTA_WIDTH} cannot exceed data.shape[axis]={T}")
    taps = DELTA1_TAPS if order == 1 else DELTA2_TAPS
    xd = x.astype(np.float64)
    out = np.empty_like(xd)
    acc = np.zeros((x.shape[0], T - 8), dtype=np.float64)
    for j in range(9):
        acc += taps[j] * xd[:, j:j + T - 8]
    out[:, 4:T - 4] = acc
    out[:, :4] = acc[:, :1]
    out[:, T - 4:] = acc[:, -1:]
    return out.astype(np.float32)


# ----------------------------------------------------------------------------
# A.5  chroma_stft with estimated tuning
# ----------------------------------------------------------------------------
def piptrack_peaks(S: np.ndarray):
    """librosa.piptrack(S=S, fmin=150, fmax=4000, threshold=0.1) reduced to its peak
    list: returns (pitch float32[npk], mag float32[npk]) for every (bin, frame) that
    is a thresholded local maximum inside [150, 4000) Hz."""
    S = np.abs(S)
    ref = np.float32(0.1) * S.max(axis=0, keepdims=True)
    Q = S * (S > ref)
    Qp = np.pad(Q, ((1, 1), (0, 0)), mode="edge")
    lmax = (Q > Qp[:-2]) & (Q >= Qp[2:])                     # librosa.util.localmax, axis=-2
    freqs = np.fft.rfftfreq(N_FFT, 1.0 / SR)
    lmax &= ((freqs >= 150.0) & (freqs < 4000.0))[:, None]
    kk, tt = np.nonzero(lmax)
    # parabolic interpolation (librosa >= 0.10 _parabolic_interpolation, numba float64 math
    # on float32 inputs) and np.gradient (float32)
    xm = S[kk - 1, tt]
    x0 = S[kk, tt]
    xp = S[kk + 1, tt]
    a = xp.astype(np.float64) + xm.astype(np.float64) - 2.0 * x0.astype(np.float64)
    b = (xp.astype(np.float64) - xm.astype(np.float64)) / 2.0
    with np.errstate(divide="ignore", invalid="ignore"):
        shift = np.where(np.abs(b) >= np.abs(a), 0.0, -b / a).astype(np.float32)
    avg = ((xp - xm) / np.float32(2.0)).astype(np.float32)
    dskew = (np.float32(0.5) * avg * shift).astype(np.float32)
    pitch = ((kk + shift) * float(SR) / N_FFT).astype(np.float32)   # int + f32 -> f64 -> f32
    mag = (x0 + dskew).astype(np.float32)
    return pitch, mag


TUNING_EDGES = np.linspace(-0.5, 0.5, 101)      # librosa.pitch_tuning(resolution=0.01)


def estimate_tuning(S: np.ndarray) -> float:
    """librosa.estimate_tuning(S=S, sr=16000, bins_per_octave=12) -> one of 100 values."""
    pitch, mag = piptrack_peaks(S)
    keep = pitch > 0
    if not keep.any():
        return 0.0
    thr = np.median(mag[keep])
    f = pitch[(mag >= thr) & keep]
    f = f[f > 0]
    if f.size == 0:
        return 0.0
    octs = np.log2(f / np.float32(27.5))                    # float32 math (hz_to_octs)
    resid = np.mod(np.float32(12) * octs, np.float32(1.0)).astype(np.float32)
    resid[resid >= 0.5] -= 1.0
    counts, edges = np.histogram(resid, TUNING_EDGES)
    return float(edges[np.argmax(counts)])


def tuning_index(tuning: float) -> int:
    return int(np.argmin(np.abs(TUNING_EDGES[:100] - tuning)))


@functools.lru_cache(maxsize=128)
def chroma_filterbank(tuning: float) -> np.ndarray:
    """librosa.filters.chroma(sr=16000, n_fft=2048, tuning=tuning, n_chroma=12,
    ctroct=5, octwidth=2, norm=2, base_c=True, dtype=float32) -> float32 [12, 1025]."""
    freqs = np.linspace(0, SR, N_FFT, endpoint=False)[1:]
    a440 = 440.0 * 2.0 ** (tuning / N_CHROMA)
    frqbins = N_CHROMA * np.log2(freqs / (a440 / 16.0))
    frqbins = np.concatenate(([frqbins[0] - 1.5 * N_CHROMA], frqbins))
    binwidth = np.concatenate((np.maximum(frqbins[1:] - frqbins[:-1], 1.0), [1]))
    D = np.subtract.outer(frqbins, np.arange(0, N_CHROMA, dtype="d")).T
    half = np.round(float(N_CHROMA) / 2)
    D = np.remainder(D + half + 10 * N_CHROMA, N_CHROMA) - half
    wts = np.exp(-0.5 * (2 * D / np.tile(binwidth, (N_CHROMA, 1))) ** 2)
    wts = wts / np.sqrt(np.sum(wts ** 2, axis=0, keepdims=True))          # util.normalize(norm=2, axis=0)
    wts *= np.tile(np.exp(-0.5 * (((frqbins / N_CHROMA - 5.0) / 2.0) ** 2)), (N_CHROMA, 1))
    wts = np.roll(wts, -3, axis=0)
    return np.ascontiguousarray(wts[:, :N_BINS], dtype=np.float32)


def _normalize_inf(X: np.ndarray, axis: int) -> np.ndarray:
    """librosa.util.normalize(norm=inf): divide by max|.|; below tiny -> left unchanged."""
    if not np.isfinite(X).all():
        raise ValueError("Input must be finite")
    mag = np.abs(X).astype(float)
    length = np.max(mag, axis=axis, keepdims=True)
    length[length < np.finfo(X.dtype).tiny] = 1.0
    out = np.empty_like(X)
    out[:] = X / length
    return out


def chroma_stft(S: np.ndarray, tuning: float | None = None) -> np.ndarray:
    if tuning is None:
        tuning = estimate_tuning(S)
    raw = np.einsum("cf,ft->ct", chroma_filterbank(tuning), S, optimize=True)
    return _normalize_inf(raw, axis=0)


# ----------------------------------------------------------------------------
# A.8  statistics + assembly
# ----------------------------------------------------------------------------
def _stat_pair(mat: np.ndarray) -> np.ndarray:
    return np.hstack([np.mean(mat, axis=1), np.std(mat, axis=1)])


def intermediates(y: np.ndarray) -> dict:
    """Stage-by-stage values for one clip (used by the stage-wise parity tests)."""
    y = np.asarray(y)
    if y.ndim != 1 or y.size == 0:
        raise ValueError("audio must be a non-empty 1-D array")
    if not np.isfinite(y).all():
        raise ValueError("Audio buffer is not finite everywhere")      # librosa.util.valid_audio
    y = y.astype(np.float32)
    S = power_spectrogram(y)
    M = mel_power(S)
    L = power_to_db(M)
    mfcc = mfcc_from_logmel(L).astype(np.float32)
    d1 = delta(mfcc, 1)
    d2 = delta(mfcc, 2)
    tuning = estimate_tuning(S)
    chroma = chroma_stft(S, tuning)
    return dict(power=S, mel=M, logmel=L, mfcc=mfcc, delta=d1, delta2=d2, tuning=tuning, chroma=chroma)


def extract_audio_features(y, sr: int = SR) -> np.ndarray:
    """pipeline1.py:206-239 -- float32[144]; ``None`` or ANY exception (T < 9 frames,
    non-finite samples, empty input) -> zeros."""
    if y is None:
        return np.zeros(AUDIO_FEATURE_LEN, dtype=np.float32)
    if sr != SR:
        raise NotImplementedError("oracle is frozen at sr=16000 like the reference's TARGET_SR")
    try:
        it = intermediates(y)
        feats = np.hstack([_stat_pair(it["mfcc"]), _stat_pair(it["delta"]), _stat_pair(it["delta2"]),
                           np.hstack([np.mean(it["chroma"], axis=1), np.std(it["chroma"], axis=1)])])
        return feats.astype(np.float32)
    except Exception:
        return np.zeros(AUDIO_FEATURE_LEN, dtype=np.float32)


def extract_features(y, sr: int = SR, transcript: str = "") -> np.ndarray:
    """pipeline1.py:257-265 -- float32[149]; transcripts are always "" in the reference
    (pipeline1.py:366,399) so the 5 text statistics are zeros (pipeline1.py:243-244)."""
    if transcript:
        raise NotImplementedError("text statistics are outside the hot path (SURVEY 2.1)")
    out = np.zeros(TOTAL_FEATURE_LEN, dtype=np.float32)
    out[:AUDIO_FEATURE_LEN] = extract_audio_features(y, sr)
    return out
