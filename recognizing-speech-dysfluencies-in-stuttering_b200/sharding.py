"""Static clip sharding across the GPUs of one node.

Every clip is an independent unit (own padding, own top-dB maximum, own tuning, own denoise
state), so the hot path needs no exchange: rank r of W owns the contiguous index range
[floor(r N / W), floor((r+1) N / W)), which keeps the concatenated output rows in the
reference's sorted-path order (pipeline1.py:97).  The only collective is the CMVN all-reduce
in scaler.py.
"""
from __future__ import annotations


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return (rank * n_items) // world, ((rank + 1) * n_items) // world


def sliding_windows(n_samples: int, win: int = 48000, hop: int = 24000):
    """Window starts of the long-form segmenter (BASELINE config 4: win 48 000 / hop 24 000 ->
    2399 windows per hour).  Each window is an independent clip for the feature function."""
    if n_samples < win:
        return []
    return list(range(0, n_samples - win + 1, hop))
