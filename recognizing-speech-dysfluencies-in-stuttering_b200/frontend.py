"""Host-side mirror of the reference's feature interface, backed by libdysb200.so.

Same names, argument meaning and error behaviour as the reference's functions on this path
(/root/reference/pipeline1.py; second copy in main1.py):

    load_audio                 pipeline1.py:100-106
    clean_audio_and_cache      pipeline1.py:126-146
    extract_audio_features     pipeline1.py:206-239
    extract_text_features      pipeline1.py:242-254
    extract_features           pipeline1.py:257-265
    cached_extract_features    pipeline1.py:429-440   (a closure there; module-level here)
    snr_db / spectral_flatness_mean / high_freq_energy_ratio   pipeline1.py:151-186   (QC scalars, reporting only)

plus the batched entry points the reference lacks (``extract_features_batch``,
``extract_features_host``, ``extract_features_longform``, ``build_feature_cache``, ``preprocess_corpus``).  PyTorch is only plumbing here: device
memory, streams, pinned host buffers.  All arithmetic runs in the CUDA library through its C
ABI; there is no CPU fallback -- without the built library or a CUDA device calls raise.
"""
from __future__ import annotations

import logging
import os
import re
from collections import Counter
from typing import Sequence

import numpy as np
import torch

from . import _lib, mp3io, wavio
from ._lib import AUDIO_FEATURE_LEN, FEATURE_LEN, SAMPLE_RATE, STATUS_CLEAN_FALLBACK, DysError

# same module-level knobs as the reference (pipeline1.py:29-32, 77-86)
CACHE_DIR = "cache_features"
CLEAR_DIR = "clear_audio"
TARGET_SR = SAMPLE_RATE
MFCC_N = 20
TEXT_FEATURE_LEN = 5
TOTAL_FEATURE_LEN = FEATURE_LEN
PROP_DECREASE = 1.0          # pipeline1.py:140 default; main1.py:605 / main.py:657 use 0.8

_log = logging.getLogger(__name__)


# ------------------------------------------------------------------------------------------
# device plumbing
# ------------------------------------------------------------------------------------------
def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise DysError("no CUDA device visible: the dysfluency front-end has no CPU path")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise DysError(f"device must be a CUDA device, got {device}")
    return torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())


class _Arena:
    """Grow-only uint8 scratch tensor per (device, slot, stream).  Work queued on different streams never shares a
    buffer (two asynchronous calls on two streams would otherwise race on the denoised clips and the scratch); a buffer
    that is replaced by a larger one is handed back to the caching allocator only after the stream that used it has
    passed this point (record_stream)."""

    def __init__(self):
        self._bufs: dict = {}

    @staticmethod
    def _key(device: torch.device, slot: int):
        return (device.index, slot, int(torch.cuda.current_stream(device).cuda_stream))

    def get(self, device: torch.device, nbytes: int, slot: int = 0) -> torch.Tensor:
        key = self._key(device, slot)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            if buf is not None:
                buf.record_stream(torch.cuda.current_stream(device))
            self._bufs[key] = None
            buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf

    def size(self, device: torch.device, slot: int = 0) -> int:
        buf = self._bufs.get(self._key(device, slot))
        return 0 if buf is None else int(buf.numel())

    def clear(self):
        self._bufs.clear()


_arena = _Arena()
abi_calls = {"features": 0}          # C-ABI feature calls issued by this module (tests: the run-in-pieces path)


def release_workspaces():
    """Frees the cached device scratch."""
    _arena.clear()


def _workspace_limit(dev: torch.device, full: int, slot: int) -> int | None:
    """Largest workspace one C-ABI call may ask for, or None for "no limit applies".  DYS_MAX_WORKSPACE_MB is always
    honoured when set.  Otherwise half of the free device memory -- but cudaMemGetInfo is a slow, occasionally blocking
    driver query, so it is only asked when a workspace above 1 GiB would have to be allocated (the 400-clip chunks of
    the host streaming path never are)."""
    env = os.environ.get("DYS_MAX_WORKSPACE_MB")
    if env:
        return max(1, int(env)) << 20
    if full <= (1 << 30) or _arena.size(dev, slot) >= full:
        return None
    free, _ = torch.cuda.mem_get_info(dev)
    return max(free // 2, 1 << 28)


def _run_device(d_audio: torch.Tensor, d_starts: torch.Tensor, d_lengths: torch.Tensor, max_len: int, denoise: bool,
                prop_decrease: float, want_pcm: bool, d_pcm_starts: torch.Tensor | None, total_pcm: int, slot: int = 0,
                workspace_bytes: int | None = None, pcm_out: torch.Tensor | None = None):
    """C-ABI call(s) on the current stream. Returns (raw, clean | None, status, pcm | None) device tensors.
    The per-clip workspace (denoised float32 + PCM-16 copies) grows with the batch: a batch whose workspace would not
    fit the limit is run as consecutive pieces over the same arena -- clips are independent, so the rows are the same."""
    lib = _lib.load()
    dev = d_audio.device
    n = int(d_starts.numel())
    flag = 1 if denoise else 0
    with torch.cuda.device(dev):
        _lib.check(lib.dys_init(), "dys_init")
        stream = torch.cuda.current_stream(dev).cuda_stream
        raw = torch.empty((n, FEATURE_LEN), dtype=torch.float32, device=dev)
        clean = torch.empty((n, FEATURE_LEN), dtype=torch.float32, device=dev) if denoise else None
        status = torch.empty(((2 if denoise else 1) * n,), dtype=torch.int32, device=dev)
        pcm = pcm_out if pcm_out is not None else (
            torch.empty((total_pcm,), dtype=torch.int16, device=dev) if (denoise and want_pcm) else None)
        piece = n
        full = lib.dys_workspace_bytes(n, max_len, flag)
        if workspace_bytes is None and n > 1:
            limit = _workspace_limit(dev, full, slot)
            if limit is not None and full > limit:
                lo, hi = 1, n                                   # largest piece whose workspace fits (monotone in the count)
                while lo < hi:
                    mid = (lo + hi + 1) // 2
                    if lib.dys_workspace_bytes(mid, max_len, flag) <= limit:
                        lo = mid
                    else:
                        hi = mid - 1
                piece = lo
        need = lib.dys_workspace_bytes(piece, max_len, flag) if workspace_bytes is None else workspace_bytes
        ws = _arena.get(dev, max(int(need), 256), slot)
        pcm_in = d_audio.dtype == torch.int16                   # 16-bit samples: value = q / 32768 (librosa.load on a PCM-16 WAV)
        f_raw = lib.dys_features_raw_pcm16 if pcm_in else lib.dys_features_raw
        f_both = lib.dys_features_raw_clean_pcm16 if pcm_in else lib.dys_features_raw_clean
        for i0 in range(0, n, piece):
            m = min(piece, n - i0)
            st = status if m == n else torch.empty(((2 if denoise else 1) * m,), dtype=torch.int32, device=dev)
            abi_calls["features"] += 1
            if not denoise:
                _lib.check(f_raw(d_audio.data_ptr(), d_starts[i0:].data_ptr(), d_lengths[i0:].data_ptr(), m,
                                                max_len, raw[i0:].data_ptr(), st.data_ptr(), ws.data_ptr(), int(need), stream),
                           "dys_features_raw")
            else:
                _lib.check(f_both(d_audio.data_ptr(), d_starts[i0:].data_ptr(), d_lengths[i0:].data_ptr(), m,
                                                      max_len, float(prop_decrease), raw[i0:].data_ptr(), clean[i0:].data_ptr(),
                                                      st.data_ptr(), pcm.data_ptr() if pcm is not None else None,
                                                      d_pcm_starts[i0:].data_ptr() if pcm is not None else None,
                                                      ws.data_ptr(), int(need), stream), "dys_features_raw_clean")
            if m != n:
                status[i0:i0 + m] = st[:m]
                if denoise:
                    status[n + i0:n + i0 + m] = st[m:]
        return raw, clean, status, pcm


_uniform_cache: dict = {}


def _uniform_index(dev: torch.device, B: int, n: int):
    """starts / lengths of a [B, n] batch, kept on the device (two pageable host-to-device copies per call otherwise:
    they would make the host wait for everything queued on the stream before)."""
    key = (dev.index, B, n)
    hit = _uniform_cache.get(key)
    if hit is None:
        if len(_uniform_cache) > 16:
            _uniform_cache.clear()
        hit = ((torch.arange(B, dtype=torch.int64, device=dev) * n), torch.full((B,), n, dtype=torch.int32, device=dev))
        torch.cuda.current_stream(dev).synchronize()             # later calls may run on other streams
        _uniform_cache[key] = hit
    return hit


def _run_length_binned(d_audio, d_starts, d_lens, h_lens: np.ndarray, denoise: bool, prop: float, want_pcm: bool,
                       d_pcm_starts, total_pcm: int, bins: int = 1):
    """Ragged batches in length-sorted order: clips of similar length then share launch groups and
    waves of CTAs, so no SM idles behind one long clip; rows are scattered back into the caller's order.  A clip's
    vector does not depend on its batch, so the result is bit-identical.  Measured on a UCLASS-like length mix
    (9 050 clips, 0.45 - 10.1 s): 7.6e5 audio-s/s as given, 8.2e5 sorted (uniform 3-s clips: 8.7e5).  ``bins`` > 1
    additionally cuts the sorted order into separate calls with their own ``max_len`` (less scratch per clip, but
    smaller launch groups: slower in that measurement, kept for memory-tight callers).  Uniform batches take the
    plain single-call path."""
    n = len(h_lens)
    max_len = int(h_lens.max()) if n else 0
    if n < 256 or max_len <= 1.25 * float(np.median(h_lens)):
        return _run_device(d_audio, d_starts, d_lens, max_len, denoise, prop, want_pcm, d_pcm_starts, total_pcm)
    dev = d_audio.device
    order = np.argsort(h_lens, kind="stable")
    raw = torch.empty((n, FEATURE_LEN), dtype=torch.float32, device=dev)
    clean = torch.empty((n, FEATURE_LEN), dtype=torch.float32, device=dev) if denoise else None
    status = torch.empty(((2 if denoise else 1) * n,), dtype=torch.int32, device=dev)
    pcm = torch.empty((total_pcm,), dtype=torch.int16, device=dev) if want_pcm else None
    for b in range(bins):
        idx_h = order[(b * n) // bins:((b + 1) * n) // bins]
        if len(idx_h) == 0:
            continue
        idx = torch.from_numpy(idx_h).to(dev)
        m = len(idx_h)
        r, c, st, _ = _run_device(d_audio, d_starts[idx], d_lens[idx], int(h_lens[idx_h].max()), denoise, prop, want_pcm,
                                  d_pcm_starts[idx] if want_pcm else None, total_pcm, pcm_out=pcm)
        raw[idx] = r
        status[idx] = st[:m]
        if denoise:
            clean[idx] = c
            status[idx + n] = st[m:]
    return raw, clean, status, pcm


def _all_int16(clips: Sequence) -> bool:
    seen = False
    for c in clips:
        if c is None:
            continue
        dt = c.dtype if isinstance(c, (np.ndarray, torch.Tensor)) else None
        if dt not in (np.dtype(np.int16), torch.int16):
            return False
        seen = True
    return seen


def _pack_host(clips: Sequence, pcm16: bool = False) -> tuple[torch.Tensor, np.ndarray, np.ndarray, int]:
    """Packs variable-length host clips into one pinned buffer (float32, or int16 for PCM-16 clips); clip starts are
    4-sample aligned."""
    lens = np.asarray([0 if c is None else int(np.asarray(c).shape[0]) for c in clips], dtype=np.int64)
    padded = (lens + 3) & ~3
    starts = np.zeros(len(clips), dtype=np.int64)
    if len(clips) > 1:
        starts[1:] = np.cumsum(padded)[:-1]
    total = int(padded.sum()) if len(clips) else 0
    buf = torch.zeros(max(total, 4), dtype=torch.int16 if pcm16 else torch.float32).pin_memory()
    view = buf.numpy()
    for c, s, n in zip(clips, starts, lens):
        if n:
            view[s:s + n] = np.asarray(c, dtype=np.int16 if pcm16 else np.float32).reshape(-1)
    return buf, starts, lens.astype(np.int32), int(lens.max()) if len(clips) else 0


# ------------------------------------------------------------------------------------------
# batched entry points
# ------------------------------------------------------------------------------------------
def extract_features_batch(audio, lengths=None, starts=None, sr: int = TARGET_SR, denoise: bool = False,
                           prop_decrease: float | None = None, return_status: bool = False, return_pcm: bool = False,
                           device=None):
    """149-dim feature vectors for a batch of 16 kHz clips (row i == reference ``extract_features(clip_i, sr)``).

    audio   * list of 1-D float arrays (numpy / CPU torch), variable length, or
            * 2-D [B, n] numpy / torch tensor (CPU or CUDA) of equal-length clips, or
            * 1-D CUDA/CPU tensor of packed samples with ``starts`` (int64[B]) and ``lengths`` (int32[B]);
              windows may overlap (long-form sliding windows need no copy).
            int16 input (every clip of a list, or the tensor) is taken as PCM-16, value = q / 32768 -- what
            ``librosa.load`` returns for a 16-bit WAV -- and stays 16-bit all the way into the kernels.
    denoise False -> raw[B,149];  True -> (raw[B,149], clean[B,149]) where clean goes through the
            reference's spectral gate, peak normalisation and PCM-16 round trip.
    Returns CUDA float32 tensors (plus int32 status [B] or [2B], plus a list of int16 PCM tensors).
    """
    if sr != TARGET_SR:
        raise ValueError(f"only sr={TARGET_SR} is supported (the reference always resamples to TARGET_SR)")
    prop = PROP_DECREASE if prop_decrease is None else float(prop_decrease)
    dev = _device(device if device is not None else (audio.device if isinstance(audio, torch.Tensor) and audio.is_cuda else None))
    with torch.cuda.device(dev):
        if isinstance(audio, (list, tuple)):
            host, h_starts, h_lens, max_len = _pack_host(audio, pcm16=_all_int16(audio))
            d_audio = host.to(dev, non_blocking=True)
        else:
            if isinstance(audio, torch.Tensor):
                t = audio
            else:
                a = np.asarray(audio)
                t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.int16 if a.dtype == np.int16 else np.float32))
            if t.dtype not in (torch.float32, torch.int16):
                t = t.float()
            if t.dim() == 2:
                B, n = t.shape
                t = t.contiguous()
                h_starts = np.arange(B, dtype=np.int64) * n
                h_lens = np.full(B, n, dtype=np.int32) if lengths is None else np.asarray(lengths, dtype=np.int32)
                if len(h_lens) != B or (B and (h_lens.min() < 0 or h_lens.max() > n)):
                    raise ValueError("lengths must hold one value in [0, n] per row of the [B, n] batch")
                max_len = n
                d_audio = t.reshape(-1).to(dev, non_blocking=True)
            elif t.dim() == 1 and starts is not None and lengths is not None:
                h_starts = np.asarray(starts.cpu() if isinstance(starts, torch.Tensor) else starts, dtype=np.int64)
                h_lens = np.asarray(lengths.cpu() if isinstance(lengths, torch.Tensor) else lengths, dtype=np.int32)
                if len(h_starts) and (h_starts.min() < 0 or int((h_starts + h_lens).max()) > t.numel()):
                    raise ValueError("starts/lengths reach outside the sample buffer")
                max_len = int(h_lens.max()) if len(h_lens) else 0
                d_audio = t.contiguous().to(dev, non_blocking=True)
            else:
                raise ValueError("audio must be a list of clips, a [B, n] array, or packed samples with starts+lengths")
        B = len(h_starts)
        if B == 0:
            empty = torch.zeros((0, FEATURE_LEN), dtype=torch.float32, device=dev)
            res = (empty, empty.clone()) if denoise else empty
            return res
        uniform = (not isinstance(audio, (list, tuple))) and t.dim() == 2 and lengths is None
        if uniform:                                               # [B, n] batches: the index tensors are cached on the device
            d_starts, d_lens = _uniform_index(dev, B, max_len)
        else:
            d_starts = torch.from_numpy(h_starts).to(dev, non_blocking=True)
            d_lens = torch.from_numpy(np.ascontiguousarray(h_lens)).to(dev, non_blocking=True)
        pcm_starts = None
        total_pcm = 0
        if denoise and return_pcm:
            pl = np.maximum(h_lens.astype(np.int64), 0)
            hp = np.zeros(B, dtype=np.int64)
            hp[1:] = np.cumsum(pl)[:-1]
            total_pcm = int(pl.sum())
            pcm_starts = torch.from_numpy(hp).to(dev, non_blocking=True)
        raw, clean, status, pcm = _run_length_binned(d_audio, d_starts, d_lens, h_lens, denoise, prop,
                                                     bool(denoise and return_pcm), pcm_starts, max(total_pcm, 1))
        # keep inputs alive until the stream has consumed them
        for tns in (d_audio, d_starts, d_lens, pcm_starts):
            if tns is not None:
                tns.record_stream(torch.cuda.current_stream(dev))
        out = [raw, clean] if denoise else [raw]
        if return_status:
            out.append(status)
        if denoise and return_pcm:
            hp_list = hp.tolist()
            out.append([pcm[s:s + int(n)] for s, n in zip(hp_list, np.maximum(h_lens, 0))])
        return out[0] if len(out) == 1 else tuple(out)


def extract_features_host(audio: torch.Tensor, denoise: bool = True, prop_decrease: float | None = None,
                          chunk_clips: int | None = None, out_raw: torch.Tensor | None = None,
                          out_clean: torch.Tensor | None = None, device=None, compute_streams: int = 3):
    """End-to-end host path: equal-length clips [B, n] in (preferably pinned) HOST memory ->
    host float32 [B,149] raw (and clean).

    One copy stream pushes the whole batch to the device back to back, chunk by chunk, and records an event per
    chunk, so the PCIe link is never idle.  ``compute_streams`` compute streams take the chunks in turn: each waits
    for its chunk's event, runs the kernels on it (own scratch arena) and sends the 2 x 149 floats per clip back into
    the (pinned) result tensors.  Compute trails the copy by about one chunk, so chunks are small; two streams let a
    chunk's kernels fill the idle tail of its predecessor's, which is what makes small chunks efficient (measured on
    B200, 10 000 3-s clips: 42.6 ms with one compute stream and 1250-clip chunks, 39.9 ms with two streams and 625,
    37.8 ms with three and 400; the copy alone takes 34.5 ms, the kernels alone 34.4 ms)."""
    if audio.is_cuda or audio.dim() != 2 or audio.dtype not in (torch.float32, torch.int16):
        raise ValueError("audio must be a 2-D float32 or int16 (PCM-16) CPU tensor [B, n]")
    prop = PROP_DECREASE if prop_decrease is None else float(prop_decrease)
    dev = _device(device)
    B, n = audio.shape
    pcm_in = audio.dtype == torch.int16           # half the bytes over PCIe; bit-identical to feeding int16 / 32768.0f
    esz = 2 if pcm_in else 4
    if out_raw is None:
        out_raw = torch.empty((B, FEATURE_LEN), dtype=torch.float32).pin_memory()
    if denoise and out_clean is None:
        out_clean = torch.empty((B, FEATURE_LEN), dtype=torch.float32).pin_memory()
    if B == 0:
        return (out_raw, out_clean) if denoise else out_raw
    if chunk_clips is None:
        # float32 samples: the copy is the bottleneck, small chunks keep the kernels close behind it; PCM-16 halves the copy
        # and the kernels become the bottleneck, which larger launch groups serve better (measured, 10 000 3-s clips,
        # tools/sweep_e2e_pcm.py / sweep_e2e_f32.py: PCM-16 33.9 ms at 400 clips per chunk, 31.7 at 592, 32.0 at 800, 32.3 at
        # 1184; float32 36.2 ms at 296, 36.8 at 400, 37.6 at 888)
        chunk_clips = max(1, (592 * 48000) // max(n, 1)) if pcm_in else max(1, (296 * 48000) // max(n, 1))
    chunk = max(1, min(int(chunk_clips), B))
    head = [max(1, chunk // 4), max(1, chunk // 2)] if B >= 3 * chunk else []      # short first copies: kernels start early
    sizes, left = list(head), B - sum(head)
    while left > 0:
        sizes.append(min(chunk, left))
        left -= sizes[-1]
    n_comp = max(1, min(int(compute_streams), 4))
    lib = _lib.load()
    flag = 1 if denoise else 0
    with torch.cuda.device(dev), _host_lock:
        _lib.check(lib.dys_init(), "dys_init")
        cur = torch.cuda.current_stream(dev)
        streams = _host_streams(dev, 1 + n_comp)
        copy_s, comp = streams[0], streams[1:]
        staging = _arena.get(dev, B * n * esz, slot=10)[:B * n * esz].view(audio.dtype)
        f_both = lib.dys_features_raw_clean_pcm16 if pcm_in else lib.dys_features_raw_clean
        f_raw = lib.dys_features_raw_pcm16 if pcm_in else lib.dys_features_raw
        biggest = max(sizes)
        base_starts = (torch.arange(biggest, dtype=torch.int64) * n).to(dev)
        base_lens = torch.full((biggest,), n, dtype=torch.int32, device=dev)
        # per compute stream: its own scratch arena and result rows, allocated once per call -- the per-chunk host work is
        # then one library call and two result copies (the host thread is the part of this path that a busy box delays)
        need = int(lib.dys_workspace_bytes(biggest, n, flag))
        slots = []
        for k in range(n_comp):
            ws = _arena.get(dev, max(need, 256), slot=1 + k)
            rows = _arena.get(dev, biggest * (2 * FEATURE_LEN * 4 + 8), slot=20 + k)
            raw_k = rows[:biggest * FEATURE_LEN * 4].view(torch.float32).view(biggest, FEATURE_LEN)
            clean_k = rows[biggest * FEATURE_LEN * 4:2 * biggest * FEATURE_LEN * 4].view(torch.float32).view(biggest, FEATURE_LEN)
            st_k = rows[2 * biggest * FEATURE_LEN * 4:].view(torch.int32)
            slots.append((ws, raw_k, clean_k, st_k))
        for s in streams:
            s.wait_stream(cur)
        landed, c0 = [], 0
        with torch.cuda.stream(copy_s):
            for cnt in sizes:
                staging[c0 * n:(c0 + cnt) * n].copy_(audio[c0:c0 + cnt].reshape(-1), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_s)
                landed.append(ev)
                c0 += cnt
        c0 = 0
        in_ptr, st_ptr, ln_ptr = staging.data_ptr(), base_starts.data_ptr(), base_lens.data_ptr()
        for i, (cnt, ev) in enumerate(zip(sizes, landed)):
            s = comp[i % n_comp]
            ws, raw_k, clean_k, st_k = slots[i % n_comp]
            with torch.cuda.stream(s):
                s.wait_event(ev)
                if denoise:
                    rc = f_both(in_ptr + c0 * n * esz, st_ptr, ln_ptr, cnt, n, prop, raw_k.data_ptr(),
                                clean_k.data_ptr(), st_k.data_ptr(), None, None, ws.data_ptr(), need, s.cuda_stream)
                else:
                    rc = f_raw(in_ptr + c0 * n * esz, st_ptr, ln_ptr, cnt, n, raw_k.data_ptr(), st_k.data_ptr(),
                               ws.data_ptr(), need, s.cuda_stream)
                _lib.check(rc, "dys_features_raw_clean" if denoise else "dys_features_raw")
                out_raw[c0:c0 + cnt].copy_(raw_k[:cnt], non_blocking=True)
                if denoise:
                    out_clean[c0:c0 + cnt].copy_(clean_k[:cnt], non_blocking=True)
            c0 += cnt
        for s in comp:
            cur.wait_stream(s)
        cur.synchronize()
    return (out_raw, out_clean) if denoise else out_raw


_pinned_cache: dict = {}


def _pinned_rows(which: str, rows: int) -> torch.Tensor:
    """Grow-only pinned float32 [rows, 149] staging for results (pinning a fresh buffer per call costs a millisecond or
    two of cudaHostRegister); only used under _host_lock."""
    buf = _pinned_cache.get(which)
    if buf is None or buf.shape[0] < rows:
        buf = torch.empty((max(rows, 1), FEATURE_LEN), dtype=torch.float32).pin_memory()
        _pinned_cache[which] = buf
    return buf[:rows]


class PackedClips:
    """Ragged host clips packed once into ONE pinned buffer, in length-sorted order (clips of similar length share
    launch groups and waves of CTAs), ready to be streamed to the GPU chunk by chunk.  ``order[j]`` is the caller's index
    of the j-th packed clip."""

    def __init__(self, clips: Sequence, pcm16: bool | None = None, sort: bool = True):
        lens = np.asarray([0 if c is None else int(np.asarray(c).shape[0]) for c in clips], dtype=np.int64)
        self.order = np.argsort(lens, kind="stable") if sort else np.arange(len(clips))
        self.pcm16 = _all_int16(clips) if pcm16 is None else bool(pcm16)
        ordered = [clips[i] for i in self.order]
        self.buf, self.starts, self.lengths, self.max_len = _pack_host(ordered, pcm16=self.pcm16)
        self.n_clips = len(clips)

    @property
    def total_samples(self) -> int:
        return int(self.lengths.astype(np.int64).sum())


def extract_features_host_packed(packed: PackedClips, denoise: bool = True, prop_decrease: float | None = None,
                                 chunk_samples: int | None = None, out_raw: torch.Tensor | None = None,
                                 out_clean: torch.Tensor | None = None, device=None, compute_streams: int = 3):
    """End-to-end host path for RAGGED clips (the real corpus: 0.45 - 10.1 s): the packed pinned buffer is pushed to the
    device back to back in chunks of about ``chunk_samples`` samples cut at clip boundaries; compute streams take the
    chunks in turn (one C-ABI call each over that chunk's clip range -- clips are addressed by starts + lengths, so the
    device copy keeps the host layout) and send the rows back.  Returns host float32 [B,149] raw (and clean) in the
    CALLER's clip order."""
    prop = PROP_DECREASE if prop_decrease is None else float(prop_decrease)
    dev = _device(device)
    B = packed.n_clips
    if out_raw is None:
        out_raw = torch.empty((B, FEATURE_LEN), dtype=torch.float32).pin_memory()
    if denoise and out_clean is None:
        out_clean = torch.empty((B, FEATURE_LEN), dtype=torch.float32).pin_memory()
    if B == 0:
        return (out_raw, out_clean) if denoise else out_raw
    esz = 2 if packed.pcm16 else 4
    if chunk_samples is None:
        chunk_samples = (592 if packed.pcm16 else 296) * 48000                   # see extract_features_host
    ends = packed.starts + ((packed.lengths.astype(np.int64) + 3) & ~3)          # padded extent of every clip in the buffer
    cuts, c0 = [], 0                                                              # (first clip, end clip, first sample, end sample)
    while c0 < B:
        limit = packed.starts[c0] + max(int(chunk_samples), int(ends[c0] - packed.starts[c0]))
        c1 = int(np.searchsorted(ends, limit, side="right"))
        c1 = max(c1, c0 + 1)
        cuts.append((c0, c1, int(packed.starts[c0]), int(ends[c1 - 1])))
        c0 = c1
    lib = _lib.load()
    flag = 1 if denoise else 0
    n_comp = max(1, min(int(compute_streams), 4))
    with torch.cuda.device(dev), _host_lock:
        srt_raw = _pinned_rows("raw", B)                                          # rows in packed (sorted) order
        srt_clean = _pinned_rows("clean", B) if denoise else None
        _lib.check(lib.dys_init(), "dys_init")
        cur = torch.cuda.current_stream(dev)
        streams = _host_streams(dev, 1 + n_comp)
        copy_s, comp = streams[0], streams[1:]
        total = int(ends[-1])
        staging = _arena.get(dev, max(total * esz, 256), slot=11)[:total * esz].view(packed.buf.dtype)
        d_starts = torch.from_numpy(packed.starts).to(dev)
        d_lens = torch.from_numpy(np.ascontiguousarray(packed.lengths)).to(dev)
        biggest = max(c1 - c0 for c0, c1, _, _ in cuts)
        need = max(int(lib.dys_workspace_bytes(c1 - c0, int(packed.lengths[c0:c1].max()), flag)) for c0, c1, _, _ in cuts)
        slots = []
        for k in range(n_comp):
            ws = _arena.get(dev, max(need, 256), slot=31 + k)
            rows = _arena.get(dev, biggest * (2 * FEATURE_LEN * 4 + 8), slot=41 + k)
            raw_k = rows[:biggest * FEATURE_LEN * 4].view(torch.float32).view(biggest, FEATURE_LEN)
            clean_k = rows[biggest * FEATURE_LEN * 4:2 * biggest * FEATURE_LEN * 4].view(torch.float32).view(biggest, FEATURE_LEN)
            st_k = rows[2 * biggest * FEATURE_LEN * 4:].view(torch.int32)
            slots.append((ws, raw_k, clean_k, st_k))
        for s_ in streams:
            s_.wait_stream(cur)
        landed = []
        with torch.cuda.stream(copy_s):
            for _, _, s0, s1 in cuts:
                staging[s0:s1].copy_(packed.buf[s0:s1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_s)
                landed.append(ev)
        f_both = lib.dys_features_raw_clean_pcm16 if packed.pcm16 else lib.dys_features_raw_clean
        f_raw = lib.dys_features_raw_pcm16 if packed.pcm16 else lib.dys_features_raw
        in_ptr, st_ptr, ln_ptr = staging.data_ptr(), d_starts.data_ptr(), d_lens.data_ptr()
        for i, ((c0, c1, _, _), ev) in enumerate(zip(cuts, landed)):
            s_ = comp[i % n_comp]
            ws, raw_k, clean_k, st_k = slots[i % n_comp]
            cnt, mx = c1 - c0, int(packed.lengths[c0:c1].max())
            with torch.cuda.stream(s_):
                s_.wait_event(ev)
                if denoise:
                    rc = f_both(in_ptr, st_ptr + 8 * c0, ln_ptr + 4 * c0, cnt, mx, prop, raw_k.data_ptr(), clean_k.data_ptr(),
                                st_k.data_ptr(), None, None, ws.data_ptr(), need, s_.cuda_stream)
                else:
                    rc = f_raw(in_ptr, st_ptr + 8 * c0, ln_ptr + 4 * c0, cnt, mx, raw_k.data_ptr(), st_k.data_ptr(),
                               ws.data_ptr(), need, s_.cuda_stream)
                _lib.check(rc, "dys_features_raw_clean" if denoise else "dys_features_raw")
                srt_raw[c0:c1].copy_(raw_k[:cnt], non_blocking=True)
                if denoise:
                    srt_clean[c0:c1].copy_(clean_k[:cnt], non_blocking=True)
        for s_ in comp:
            cur.wait_stream(s_)
        cur.synchronize()
        idx = torch.from_numpy(packed.order)
        out_raw[idx] = srt_raw                                                    # back to the caller's order (host, 149 floats per clip)
        if denoise:
            out_clean[idx] = srt_clean
    return (out_raw, out_clean) if denoise else out_raw


def extract_features_longform(recording, win: int = 48000, hop: int = 24000, denoise: bool = True,
                              prop_decrease: float | None = None, rank: int = 0, world: int = 1, device=None):
    """Long-form recordings (BASELINE config 4): sliding windows of ``win`` samples every ``hop`` samples, each an
    independent clip for the reference's feature function (own centre padding, own top-dB maximum, own tuning,
    own 30 000-sample denoise padding) -- the reference has no segmenter, this is the batched form of calling
    ``extract_features`` / ``clean_audio_and_cache`` once per window.

    The windows are addressed in place (starts + lengths into one sample buffer), so every sample of the recording
    crosses PCIe and is stored in HBM once although windows overlap.  With ``world`` > 1 rank ``rank`` takes the
    contiguous window range ``sharding.shard_range`` gives it and uploads only the samples that range covers.

    Returns (starts int64 numpy [n_local], raw[n_local,149], clean[n_local,149] | None) -- CUDA float32 tensors."""
    from . import sharding
    if isinstance(recording, torch.Tensor):
        rec = recording.reshape(-1)
        if rec.dtype != torch.float32:
            rec = rec.float()
    else:
        rec = torch.from_numpy(np.ascontiguousarray(recording, dtype=np.float32).reshape(-1))
    all_starts = np.asarray(sharding.sliding_windows(int(rec.numel()), win, hop), dtype=np.int64)
    lo, hi = sharding.shard_range(len(all_starts), rank, world)
    starts = all_starts[lo:hi]
    dev = _device(device if device is not None else (rec.device if rec.is_cuda else None))
    if len(starts) == 0:
        empty = torch.zeros((0, FEATURE_LEN), dtype=torch.float32, device=dev)
        return starts, empty, (empty.clone() if denoise else None)
    s0, s1 = int(starts[0]), int(starts[-1]) + win
    with torch.cuda.device(dev):
        span = rec[s0:s1].to(dev, non_blocking=True)
        lens = np.full(len(starts), win, dtype=np.int32)
        out = extract_features_batch(span, lengths=lens, starts=starts - s0, denoise=denoise, prop_decrease=prop_decrease,
                                     device=dev)
    if denoise:
        return starts, out[0], out[1]
    return starts, out, None


_streams: dict = {}
_host_lock = __import__("threading").RLock()      # the host streaming paths own fixed streams and arena slots: one call at a time


def _host_streams(dev: torch.device, count: int = 2):
    have = _streams.setdefault(dev.index, [])
    while len(have) < count:
        have.append(torch.cuda.Stream(device=dev))
    return have[:count]


# ------------------------------------------------------------------------------------------
# the reference's single-clip interface
# ------------------------------------------------------------------------------------------
def resample_to_16k(clips: Sequence, sr_in: int, device=None, return_device: bool = False):
    """The rate-conversion half of ``librosa.load(path, sr=16000)`` (pipeline1.py:102) for a batch of mono clips at
    ``sr_in`` Hz (float32, or int16 PCM): one H2D copy, one kernel (soxr "HQ" restated, see dys_resample.cu), clip i
    comes back with ``ceil(len_i * 16000 / sr_in)`` float32 samples.  ``return_device`` -> (packed CUDA float32 buffer,
    int64 starts, int32 lengths) ready for ``extract_features_batch(buf, starts=..., lengths=...)`` without a host trip."""
    lib = _lib.load()
    dev = _device(device)
    if sr_in == TARGET_SR:
        raise ValueError("clips are already at 16 kHz")
    pcm16 = _all_int16(clips)
    with torch.cuda.device(dev):
        _lib.check(lib.dys_init(), "dys_init")
        host, h_starts, h_lens, max_len = _pack_host(clips, pcm16=pcm16)
        out_lens = np.asarray([lib.dys_resampled_length(int(n), int(sr_in)) for n in h_lens], dtype=np.int64)
        if len(out_lens) and out_lens.min() < 0:
            raise ValueError(f"unsupported sample rate {sr_in}")
        padded = (out_lens + 3) & ~3
        o_starts = np.zeros(len(clips), dtype=np.int64)
        if len(clips) > 1:
            o_starts[1:] = np.cumsum(padded)[:-1]
        total = int(padded.sum()) if len(clips) else 0
        d_in = host.to(dev, non_blocking=True)
        d_is = torch.from_numpy(h_starts).to(dev, non_blocking=True)
        d_il = torch.from_numpy(np.ascontiguousarray(h_lens)).to(dev, non_blocking=True)
        d_os = torch.from_numpy(o_starts).to(dev, non_blocking=True)
        d_out = torch.zeros(max(total, 4), dtype=torch.float32, device=dev)
        if len(clips) and max_len > 0:
            _lib.check(lib.dys_resample_to_16k(d_in.data_ptr(), 1 if pcm16 else 0, int(sr_in), d_is.data_ptr(), d_il.data_ptr(),
                                               len(clips), int(max_len), d_out.data_ptr(), d_os.data_ptr(),
                                               torch.cuda.current_stream(dev).cuda_stream), "dys_resample_to_16k")
        for t in (d_in, d_is, d_il, d_os):
            t.record_stream(torch.cuda.current_stream(dev))
        if return_device:
            return d_out, o_starts, out_lens.astype(np.int32)
        h = d_out.cpu().numpy()
        return [h[s:s + n].copy() for s, n in zip(o_starts, out_lens)]


def _decode_file(path: str):
    """-> (mono samples at the file's own rate: float32, or int16 for PCM-16 WAV; sample rate)."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".mp3":
        return mp3io.read_mp3(path)
    if ext == ".wav":
        try:
            return wavio.read_wav_pcm16(path)           # mono PCM-16 stays 16-bit
        except ValueError:
            return wavio.read_wav(path)                 # other widths / float / multi-channel: float32 mono like librosa
    raise ValueError(f"unsupported audio format {ext!r} (WAV/PCM-16 and MP3 are decoded here)")


def load_audio(path: str, sr: int = TARGET_SR):
    """pipeline1.py:100-106: ``librosa.load(path, sr=16000, mono=True)``.  Returns (float32[n], sr) or (None, None)
    after logging.  Mono PCM-16 WAV and MPEG Layer III are decoded on the host (file I/O); a file whose own rate is not
    ``sr`` -- the reference's whole MP3 corpus is 22 050 Hz -- is converted on the GPU (``resample_to_16k``)."""
    if sr != TARGET_SR:
        raise ValueError(f"only sr={TARGET_SR} is supported (the reference always passes TARGET_SR)")
    try:
        y, s = _decode_file(path)
        if y.dtype == np.int16 and s == sr:
            return (y.astype(np.float32) / np.float32(32768.0)), s
        if s != sr:
            y = resample_to_16k([y], s)[0]
        return np.asarray(y, dtype=np.float32), sr
    except DysError:
        raise
    except Exception as e:  # noqa: BLE001 - mirrors the reference's blanket handler
        logging.error(f"load_audio fail {path}: {e}")
        return None, None


def load_audio_batch(paths: Sequence[str], io_threads: int = 8):
    """``load_audio`` for many files: decoding runs on a thread pool, rate conversion as one GPU batch per source
    rate.  -> list of float32 arrays (``None`` where the reference's ``load_audio`` would return ``(None, None)``)."""
    from concurrent.futures import ThreadPoolExecutor

    def dec(p):
        try:
            return _decode_file(p)
        except Exception as e:  # noqa: BLE001
            logging.error(f"load_audio fail {p}: {e}")
            return None
    with ThreadPoolExecutor(max_workers=max(1, io_threads)) as pool:
        decoded = list(pool.map(dec, paths))
    out = [None] * len(paths)
    by_rate: dict = {}
    for i, d in enumerate(decoded):
        if d is None:
            continue
        y, s = d
        if s == TARGET_SR:
            out[i] = (y.astype(np.float32) / np.float32(32768.0)) if y.dtype == np.int16 else np.asarray(y, np.float32)
        else:
            by_rate.setdefault((s, y.dtype == np.int16), []).append(i)
    for (s, _), idx in by_rate.items():
        try:
            for i, y in zip(idx, resample_to_16k([decoded[i][0] for i in idx], s)):
                out[i] = y
        except DysError:
            raise
        except Exception as e:  # noqa: BLE001
            logging.error(f"load_audio fail ({len(idx)} files at {s} Hz): {e}")
    return out


def repetition_stats_from_text(text: str):
    """pipeline1.py:191-201."""
    if not text:
        return {"repetition_count": 0, "repetition_ratio": 0.0, "unique_ratio": 0.0}
    words = re.findall(r"\b\w+\b", text.lower())
    if len(words) < 1:
        return {"repetition_count": 0, "repetition_ratio": 0.0, "unique_ratio": 0.0}
    counts = Counter(words)
    repeats = sum(c - 1 for c in counts.values() if c > 1)
    return {"repetition_count": float(repeats), "repetition_ratio": float(repeats / len(words)),
            "unique_ratio": float(len(set(words)) / len(words))}


def extract_text_features(text: str) -> np.ndarray:
    """pipeline1.py:242-254 (host-side; the reference always passes "" -> zeros(5))."""
    if not text:
        return np.zeros(TEXT_FEATURE_LEN, dtype=np.float32)
    rep = repetition_stats_from_text(text)
    words = re.findall(r"\b\w+\b", text.lower())
    return np.array([float(len(text)), float(len(words)), rep["repetition_count"], rep["repetition_ratio"],
                     rep["unique_ratio"]], dtype=np.float32)


def extract_audio_features(y, sr: int = TARGET_SR) -> np.ndarray:
    """pipeline1.py:206-239 -> float32[144]; ``None`` and every failure mode of the reference's
    try-block (fewer than 9 frames, non-finite samples, empty input) give zeros."""
    if y is None:
        return np.zeros(AUDIO_FEATURE_LEN, dtype=np.float32)
    try:
        y = np.asarray(y, dtype=np.float32).reshape(-1)
        if y.size == 0:
            raise ValueError("empty clip")
        feats = extract_features_batch([y], sr=sr)
        return feats[0, :AUDIO_FEATURE_LEN].cpu().numpy()
    except DysError:
        raise                                   # a missing GPU/library is not a data error: fail loudly
    except Exception as e:  # noqa: BLE001
        logging.error(f"extract_audio_features error: {e}")
        return np.zeros(AUDIO_FEATURE_LEN, dtype=np.float32)


def extract_features(y, sr: int = TARGET_SR, transcript: str = "") -> np.ndarray:
    """pipeline1.py:257-265 -> float32[149]."""
    feats = np.hstack([extract_audio_features(y, sr), extract_text_features(transcript)]).astype(np.float32)
    if feats.size != TOTAL_FEATURE_LEN:
        out = np.zeros(TOTAL_FEATURE_LEN, dtype=np.float32)
        out[:min(feats.size, TOTAL_FEATURE_LEN)] = feats[:TOTAL_FEATURE_LEN]
        return out
    return feats


def clean_audio(y, prop_decrease: float | None = None):
    """In-memory body of clean_audio_and_cache (pipeline1.py:140-142): int16 PCM as the reference
    writes it, or ``None`` where the reference's except-branch fires (NaN from an all-zero clip, ...)."""
    if y is None:
        return None
    y = np.asarray(y, dtype=np.float32).reshape(-1)
    if y.size == 0:
        return None
    _, _, status, pcm = extract_features_batch([y], denoise=True, prop_decrease=prop_decrease, return_status=True,
                                               return_pcm=True)
    if int(status[1].item()) & STATUS_CLEAN_FALLBACK:
        return None
    return pcm[0].cpu().numpy()


def clean_audio_and_cache(in_path: str):
    """pipeline1.py:126-146: writes CLEAR_DIR/<stem>.wav (PCM-16) and returns its path; an existing
    file is reused; any failure is logged and gives ``None``."""
    base = os.path.basename(in_path).rsplit(".", 1)[0]
    out_path = os.path.normpath(os.path.join(CLEAR_DIR, f"{base}.wav"))
    if os.path.exists(out_path):
        return out_path
    y, sr = load_audio(in_path, sr=TARGET_SR)
    if y is None:
        return None
    try:
        pcm = clean_audio(y)
        if pcm is None:
            raise ValueError("Input must be finite")      # librosa.util.normalize's complaint
        os.makedirs(CLEAR_DIR, exist_ok=True)
        wavio.write_wav_pcm16(out_path, pcm, sr)
        return out_path
    except DysError:
        raise
    except Exception as e:  # noqa: BLE001
        logging.error(f"clean_audio fail {in_path}: {e}")
        return None


def cached_extract_features(path: str, transcript: str, suffix: str) -> np.ndarray:
    """pipeline1.py:429-440: cache key is CACHE_DIR/<basename-stem>_<suffix>_feats.npy."""
    base = os.path.basename(path).rsplit(".", 1)[0]
    cache_file = os.path.normpath(os.path.join(CACHE_DIR, f"{base}_{suffix}_feats.npy"))
    try:
        return np.array(np.load(cache_file, allow_pickle=False))      # same files; no unpickling of a shared cache directory
    except Exception:  # noqa: BLE001
        pass
    y, sr = load_audio(path, sr=TARGET_SR)
    feats = extract_features(y, sr if sr is not None else TARGET_SR, transcript)
    os.makedirs(CACHE_DIR, exist_ok=True)
    np.save(cache_file, feats)
    return feats


def _stem(path: str) -> str:
    return os.path.basename(path).rsplit(".", 1)[0]


def build_feature_cache(paths: Sequence[str], overwrite: bool = False, io_threads: int = 8):
    """Batched replacement for the reference's two per-file loops (pipeline1.py:371-417 and :447-453): loads every
    readable file (WAV or MP3, any rate: ``load_audio_batch``), runs batched GPU passes and writes the artefacts the
    reference writes -- CLEAR_DIR/<stem>.wav, CACHE_DIR/<stem>_raw_feats.npy, CACHE_DIR/<stem>_clean_feats.npy
    (byte-identical .npy headers: np.save of float32 (149,)); file writes run on a thread pool.

    Cache semantics are the reference's, artefact by artefact (unless ``overwrite``):
      * an existing CLEAR_DIR/<stem>.wav is reused as it is (pipeline1.py:134-135) -- and a missing clean vector is then
        computed FROM THAT FILE (pipeline1.py:437 loads the WAV), not from a fresh denoise, so WAV and vector agree;
      * an existing ``_raw_feats.npy`` / ``_clean_feats.npy`` is loaded, never rewritten (pipeline1.py:434-436);
      * caches are keyed by the basename stem only (pipeline1.py:132, :432), so two inputs with the same stem alias:
        the second one reuses the first one's artefacts (16 stems of the corpus do).
    Returns (X_before, X_after, kept_paths)."""
    from concurrent.futures import ThreadPoolExecutor

    def files_of(stem):
        return (os.path.normpath(os.path.join(CLEAR_DIR, f"{stem}.wav")),
                os.path.normpath(os.path.join(CACHE_DIR, f"{stem}_raw_feats.npy")),
                os.path.normpath(os.path.join(CACHE_DIR, f"{stem}_clean_feats.npy")))

    loaded = load_audio_batch(paths, io_threads)
    kept, stems, first = [], [], {}                    # first: stem -> its first clip (the one whose artefacts everybody shares)
    for p, y in zip(paths, loaded):
        if y is None:
            continue                                   # reference: skipped += 1
        kept.append(p)
        stems.append(_stem(p))
        first.setdefault(stems[-1], y)
    n = len(kept)
    Xb = np.empty((n, FEATURE_LEN), np.float32)
    Xa = np.empty((n, FEATURE_LEN), np.float32)
    if n == 0:
        return Xb, Xa, []
    os.makedirs(CLEAR_DIR, exist_ok=True)
    os.makedirs(CACHE_DIR, exist_ok=True)
    gate, raw_only, from_wav = [], [], []              # stems per GPU pass
    for stem in first:
        wav, f_raw, f_clean = files_of(stem)
        have_wav = os.path.exists(wav) and not overwrite
        need_raw = overwrite or not os.path.exists(f_raw)
        need_clean = overwrite or not os.path.exists(f_clean)
        if not have_wav:
            gate.append(stem)                          # the reference's loop A would denoise and write the WAV
        else:
            if need_raw:
                raw_only.append(stem)
            if need_clean:
                from_wav.append(stem)
    vec_raw, vec_clean, jobs = {}, {}, []
    with ThreadPoolExecutor(max_workers=max(1, io_threads)) as pool:
        if gate:
            raw, clean, status, pcm = extract_features_batch([first[s] for s in gate], denoise=True, return_status=True,
                                                             return_pcm=True)
            raw, clean, status = raw.cpu().numpy(), clean.cpu().numpy(), status.cpu().numpy()
            for i, stem in enumerate(gate):
                wav, f_raw, f_clean = files_of(stem)
                if status[len(gate) + i] & STATUS_CLEAN_FALLBACK:
                    logging.error(f"clean_audio fail {stem}: Input must be finite")
                else:
                    jobs.append(pool.submit(wavio.write_wav_pcm16, wav, pcm[i].cpu().numpy(), TARGET_SR))
                if overwrite or not os.path.exists(f_raw):
                    vec_raw[stem] = raw[i]
                    jobs.append(pool.submit(np.save, f_raw, raw[i]))
                if overwrite or not os.path.exists(f_clean):
                    vec_clean[stem] = clean[i]
                    jobs.append(pool.submit(np.save, f_clean, clean[i]))
        if raw_only:
            raw = extract_features_batch([first[s] for s in raw_only]).cpu().numpy()
            for i, stem in enumerate(raw_only):
                vec_raw[stem] = raw[i]
                jobs.append(pool.submit(np.save, files_of(stem)[1], raw[i]))
        if from_wav:                                   # the clean vector of an existing WAV is the feature vector OF THAT WAV
            wavs = []
            for s_ in from_wav:
                try:
                    wavs.append(wavio.read_wav_pcm16(files_of(s_)[0])[0])     # the reference's own WAVs: mono PCM-16, kept 16-bit
                except Exception:  # noqa: BLE001 - a foreign file: read it like load_audio would; unreadable -> the raw clip
                    y_, _ = load_audio(files_of(s_)[0])                       # (pipeline1.py:385-387 falls back to the raw file)
                    wavs.append(first[s_] if y_ is None else y_)
            if _all_int16(wavs):
                cl = extract_features_batch(wavs).cpu().numpy()
            else:                                                             # mixed dtypes: float32 of q / 32768 gives the same bits
                cl = extract_features_batch([w.astype(np.float32) / np.float32(32768.0) if w.dtype == np.int16 else w
                                             for w in wavs]).cpu().numpy()
            for i, stem in enumerate(from_wav):
                vec_clean[stem] = cl[i]
                jobs.append(pool.submit(np.save, files_of(stem)[2], cl[i]))
        for i, stem in enumerate(stems):
            _, f_raw, f_clean = files_of(stem)
            Xb[i] = vec_raw[stem] if stem in vec_raw else np.load(f_raw, allow_pickle=False)
            Xa[i] = vec_clean[stem] if stem in vec_clean else np.load(f_clean, allow_pickle=False)
        for j in jobs:
            j.result()
    return Xb, Xa, kept


# ------------------------------------------------------------------------------------------
# per-file QC scalars (pipeline1.py:151-186; reporting only)
# ------------------------------------------------------------------------------------------
def qc_metrics_batch(clips: Sequence, device=None) -> np.ndarray:
    """float32 [B, 3] = (snr_db, spectral_flatness_mean, high_freq_energy_ratio) of every clip -- the three numbers
    the reference logs per raw and per cleaned file into output_results/per_file_analysis.csv
    (pipeline1.py:379-381, 394-396).  ``None`` / empty clips give the reference's fall-back values (0.0)."""
    lib = _lib.load()
    dev = _device(device)
    B = len(clips)
    if B == 0:
        return np.zeros((0, 3), dtype=np.float32)
    with torch.cuda.device(dev):
        _lib.check(lib.dys_init(), "dys_init")
        host, h_starts, h_lens, max_len = _pack_host(clips)
        d_audio = host.to(dev, non_blocking=True)
        d_starts = torch.from_numpy(h_starts).to(dev, non_blocking=True)
        d_lens = torch.from_numpy(np.ascontiguousarray(h_lens)).to(dev, non_blocking=True)
        need = int(lib.dys_qc_workspace_bytes(B, max_len))
        ws = _arena.get(dev, max(need, 256), slot=5)
        out = torch.zeros((B, 3), dtype=torch.float32, device=dev)
        _lib.check(lib.dys_qc_metrics(d_audio.data_ptr(), d_starts.data_ptr(), d_lens.data_ptr(), B, max_len, out.data_ptr(),
                                      ws.data_ptr(), need, torch.cuda.current_stream(dev).cuda_stream), "dys_qc_metrics")
        return out.cpu().numpy()


def _qc_one(y, column: int) -> float:
    if y is None:
        return 0.0
    y = np.asarray(y, dtype=np.float32).reshape(-1)
    if y.size == 0:
        return 0.0
    return float(qc_metrics_batch([y])[0, column])


def snr_db(y) -> float:
    """pipeline1.py:151-165."""
    return _qc_one(y, 0)


def spectral_flatness_mean(y) -> float:
    """pipeline1.py:168-174."""
    return _qc_one(y, 1)


def high_freq_energy_ratio(y, sr: int = TARGET_SR) -> float:
    """pipeline1.py:177-186 (sr must be 16000, like everywhere in this package)."""
    if sr != TARGET_SR:
        raise ValueError(f"only sr={TARGET_SR} is supported")
    return _qc_one(y, 2)


# ------------------------------------------------------------------------------------------
# the reference's preprocessing + feature loops as one batched pass
# ------------------------------------------------------------------------------------------
OUTPUT_DIR = "output_results"            # pipeline1.py:32


def preprocess_corpus(paths: Sequence[str], output_dir: str | None = None):
    """Everything ``run_pipeline`` does up to the feature matrices (pipeline1.py:356-456) in batched GPU passes:
    per file the QC scalars before and after cleaning, the cleaned WAV in CLEAR_DIR, both cached vectors in
    CACHE_DIR, ``per_file_analysis.csv`` with the reference's columns, and the label of every kept file (its
    directory name, pipeline1.py:372).  Unreadable files are skipped like in the reference; a file whose cleaning
    failed is analysed and featurised from its raw samples (pipeline1.py:385-387).
    Returns (rows, X_before, X_after, labels, kept_paths); the matrices are what pipeline1.py:455-456 vstacks."""
    out_dir = OUTPUT_DIR if output_dir is None else output_dir
    X_before, X_after, kept = build_feature_cache(paths)
    raw_clips, clean_clips = [], []
    for p in kept:
        y, _ = load_audio(p, sr=TARGET_SR)
        raw_clips.append(y)
        wav = os.path.normpath(os.path.join(CLEAR_DIR, f"{_stem(p)}.wav"))
        yc, _ = load_audio(wav, sr=TARGET_SR) if os.path.exists(wav) else (None, None)
        clean_clips.append(y if yc is None else yc)
    qb = qc_metrics_batch(raw_clips) if kept else np.zeros((0, 3), np.float32)
    qa = qc_metrics_batch(clean_clips) if kept else np.zeros((0, 3), np.float32)
    rows, labels = [], []
    for i, p in enumerate(kept):
        label = os.path.basename(os.path.dirname(p)) or "unknown"
        labels.append(label)
        rows.append({"file": os.path.basename(p), "label": label, "duration_sec": float(len(raw_clips[i]) / TARGET_SR),
                     "snr_before_db": float(qb[i, 0]), "snr_after_db": float(qa[i, 0]),
                     "spectral_flatness_before": float(qb[i, 1]), "spectral_flatness_after": float(qa[i, 1]),
                     "hf_energy_ratio_before": float(qb[i, 2]), "hf_energy_ratio_after": float(qa[i, 2]),
                     "transcript": ""})
    if rows:
        import pandas as pd
        os.makedirs(out_dir, exist_ok=True)
        pd.DataFrame(rows).to_csv(os.path.join(out_dir, "per_file_analysis.csv"), index=False)
    return rows, X_before, X_after, labels, kept


# ------------------------------------------------------------------------------------------
# introspection (parity tests)
# ------------------------------------------------------------------------------------------
def debug_feature_stages(y, device=None) -> dict:
    """Stage-wise intermediates of one clip, as numpy arrays laid out like the oracle's."""
    lib = _lib.load()
    dev = _device(device)
    y = np.ascontiguousarray(y, dtype=np.float32)
    n = int(y.size)
    T = 1 + n // 512
    with torch.cuda.device(dev):
        _lib.check(lib.dys_init(), "dys_init")
        d = torch.from_numpy(y).to(dev) if n else torch.zeros(4, device=dev)
        power = torch.zeros((T, 1032), dtype=torch.float32, device=dev)
        logmel = torch.zeros((T, 128), dtype=torch.float32, device=dev)
        mfcc = torch.zeros((T, 20), dtype=torch.float32, device=dev)
        chroma = torch.zeros((T, 12), dtype=torch.float32, device=dev)
        scal = torch.zeros(4, dtype=torch.int32, device=dev)
        out = torch.zeros(FEATURE_LEN, dtype=torch.float32, device=dev)
        _lib.check(lib.dys_debug_feature_stages(d.data_ptr(), n, power.data_ptr(), logmel.data_ptr(), mfcc.data_ptr(),
                                                chroma.data_ptr(), scal.data_ptr(), out.data_ptr(),
                                                torch.cuda.current_stream(dev).cuda_stream), "dys_debug_feature_stages")
        torch.cuda.synchronize(dev)
        s = scal.cpu().numpy()
    return dict(power=power[:, :1025].cpu().numpy().T, logmel_unclamped=logmel.cpu().numpy().T, mfcc=mfcc.cpu().numpy().T,
                chroma=chroma.cpu().numpy().T, frames=int(s[0]), tuning_index=int(s[1]), peak_count=int(s[2]),
                status=int(s[3]), features=out.cpu().numpy())


def debug_denoise(y, prop_decrease: float = 1.0, device=None):
    """-> (float32[n] reduce_noise output before normalisation, peak, fallback flag)."""
    lib = _lib.load()
    dev = _device(device)
    y = np.ascontiguousarray(y, dtype=np.float32)
    with torch.cuda.device(dev):
        _lib.check(lib.dys_init(), "dys_init")
        d = torch.from_numpy(y).to(dev)
        clean = torch.zeros_like(d)
        info = torch.zeros(2, dtype=torch.float32, device=dev)
        _lib.check(lib.dys_debug_denoise(d.data_ptr(), int(y.size), float(prop_decrease), clean.data_ptr(), info.data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream), "dys_debug_denoise")
        torch.cuda.synchronize(dev)
        i = info.cpu().numpy()
    return clean.cpu().numpy(), float(i[0]), int(i[1])


def get_table(which: int, arg: int = 0) -> np.ndarray:
    """Host copy of a lookup table of the library (see dys_get_table in the header)."""
    lib = _lib.load()
    shapes = {0: ((128, 1025), np.float32), 1: ((20, 128), np.float32), 2: ((1025, 12), np.float32),
              3: ((2048,), np.float32), 4: ((101,), np.float64), 5: ((40,), np.float64), 6: ((256,), np.float64),
              7: ((1,), np.float64), 8: ((260,), np.int32), 9: ((32 * 88,), np.float32)}
    shape, dt = shapes[which]
    out = np.empty(shape, dtype=dt)
    got = lib.dys_get_table(which, arg, out.ctypes.data, out.size)
    if got != out.size:
        raise DysError(f"dys_get_table({which}) returned {got}")
    return out
