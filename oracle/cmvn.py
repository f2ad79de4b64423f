"""Oracle: the reference's global feature standardisation.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/pipeline1.py:470-473 (``StandardScaler().fit(X)`` then
``.transform(X)``; main1.py:848-852 likewise): per-feature float64 mean and
population variance (ddof 0) over all clips, ``scale = sqrt(var)`` with zero
variance -> 1.0, output ``(x - mean) / scale`` in the input's dtype (float32 for the cache matrix).
Pinned against output_results/scaler_after.pkl (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np


def fit(X: np.ndarray):
    """-> (mean float64[D], var float64[D], scale float64[D], n)."""
    X = np.asarray(X)
    n = X.shape[0]
    Xd = X.astype(np.float64)
    mean = Xd.sum(axis=0) / n
    centered = Xd - mean
    var = (centered ** 2).sum(axis=0)
    var -= centered.sum(axis=0) ** 2 / n          # sklearn's corrected two-pass term
    var /= n
    scale = np.sqrt(var)
    scale[scale < 10 * np.finfo(np.float64).eps] = 1.0     # sklearn _handle_zeros_in_scale
    return mean, var, scale, n


def transform(X: np.ndarray, mean: np.ndarray, scale: np.ndarray) -> np.ndarray:
    """``StandardScaler.transform``: the output keeps X's dtype and the arithmetic runs in it
    (``X -= mean_.astype(X.dtype); X /= scale_.astype(X.dtype)``, sklearn >= 1.4 as installed here;
    checked bit-for-bit against sklearn in tests/test_oracle_golden.py).  The reference always passes
    the float32 matrix np.vstack'ed from cache_features/ (pipeline1.py:455-473)."""
    X = np.array(X, copy=True)
    if X.dtype not in (np.float32, np.float64):
        X = X.astype(np.float64)
    X -= np.asarray(mean).astype(X.dtype)
    X /= np.asarray(scale).astype(X.dtype)
    return X


def moments(X: np.ndarray) -> np.ndarray:
    """The all-reducible form used across GPUs: float64 [1 + 2D] = [n, sum x, sum x^2]."""
    Xd = np.asarray(X, dtype=np.float64)
    return np.concatenate(([Xd.shape[0]], Xd.sum(axis=0), (Xd ** 2).sum(axis=0)))


def from_moments(acc: np.ndarray):
    D = (len(acc) - 1) // 2
    n = acc[0]
    mean = acc[1:1 + D] / n
    var = np.maximum(acc[1 + D:] / n - mean ** 2, 0.0)
    scale = np.sqrt(var)
    scale[scale < 10 * np.finfo(np.float64).eps] = 1.0
    return mean, var, scale, n
