"""The front-end as registered PyTorch operators (``torch.ops.dysb200.*``).

BASELINE.json's north star asks for "a drop-in Python/PyTorch op ... calling CUDA through a thin C-ABI extension":
these operators are that op.  They take the packed device layout of the C ABI (one float32 sample buffer, int64 clip
starts, int32 clip lengths) and return new CUDA tensors; shapes are registered for ``torch.compile`` / fake-tensor
tracing.  CPU tensors are rejected -- there is no CPU implementation.

    raw, status              = torch.ops.dysb200.features_raw(audio, starts, lengths, max_len)
    raw, clean, status       = torch.ops.dysb200.features_raw_clean(audio, starts, lengths, max_len, prop_decrease)
    qc                       = torch.ops.dysb200.qc_metrics(audio, starts, lengths, max_len)
    audio16k                 = torch.ops.dysb200.resample_to_16k(audio, starts, lengths, max_len, sr_in, out_starts, total_out)

``audio`` may be float32 or int16 (PCM-16, value = q / 32768: the library's *_pcm16 entry points) for the feature
operators and the rate converter.

Row i of ``raw`` is the reference's ``extract_features(clip_i, 16000)`` (pipeline1.py:257-265); ``clean`` is the same
function applied to the clip after ``clean_audio_and_cache`` (pipeline1.py:126-146).
"""
from __future__ import annotations

import torch

from . import _lib, frontend
from ._lib import FEATURE_LEN, DysError


def _need_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise DysError("dysb200 operators need CUDA tensors: the front-end has no CPU path")


@torch.library.custom_op("dysb200::features_raw", mutates_args=(), device_types="cuda")
def features_raw(audio: torch.Tensor, starts: torch.Tensor, lengths: torch.Tensor, max_len: int) -> tuple[torch.Tensor, torch.Tensor]:
    _need_cuda(audio, starts, lengths)
    raw, _, status, _ = frontend._run_device(audio.contiguous(), starts.contiguous(), lengths.contiguous(), int(max_len),
                                             False, 1.0, False, None, 1)
    return raw, status


@features_raw.register_fake
def _(audio, starts, lengths, max_len):
    n = starts.shape[0]
    return audio.new_empty((n, FEATURE_LEN), dtype=torch.float32), lengths.new_empty((n,), dtype=torch.int32)


@torch.library.custom_op("dysb200::features_raw_clean", mutates_args=(), device_types="cuda")
def features_raw_clean(audio: torch.Tensor, starts: torch.Tensor, lengths: torch.Tensor, max_len: int,
                       prop_decrease: float) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    _need_cuda(audio, starts, lengths)
    raw, clean, status, _ = frontend._run_device(audio.contiguous(), starts.contiguous(), lengths.contiguous(), int(max_len),
                                                 True, float(prop_decrease), False, None, 1)
    return raw, clean, status


@features_raw_clean.register_fake
def _(audio, starts, lengths, max_len, prop_decrease):
    n = starts.shape[0]
    return (audio.new_empty((n, FEATURE_LEN), dtype=torch.float32), audio.new_empty((n, FEATURE_LEN), dtype=torch.float32),
            lengths.new_empty((2 * n,), dtype=torch.int32))


@torch.library.custom_op("dysb200::qc_metrics", mutates_args=(), device_types="cuda")
def qc_metrics(audio: torch.Tensor, starts: torch.Tensor, lengths: torch.Tensor, max_len: int) -> torch.Tensor:
    _need_cuda(audio, starts, lengths)
    lib = _lib.load()
    dev = audio.device
    n = int(starts.shape[0])
    out = torch.zeros((n, 3), dtype=torch.float32, device=dev)
    if n == 0:
        return out
    with torch.cuda.device(dev):
        _lib.check(lib.dys_init(), "dys_init")
        need = int(lib.dys_qc_workspace_bytes(n, int(max_len)))
        ws = frontend._arena.get(dev, max(need, 256), slot=5)
        a, s, ln = audio.contiguous(), starts.contiguous(), lengths.contiguous()
        _lib.check(lib.dys_qc_metrics(a.data_ptr(), s.data_ptr(), ln.data_ptr(), n, int(max_len), out.data_ptr(), ws.data_ptr(),
                                      need, torch.cuda.current_stream(dev).cuda_stream), "dys_qc_metrics")
    return out


@qc_metrics.register_fake
def _(audio, starts, lengths, max_len):
    return audio.new_empty((starts.shape[0], 3))


@torch.library.custom_op("dysb200::resample_to_16k", mutates_args=(), device_types="cuda")
def resample_to_16k(audio: torch.Tensor, starts: torch.Tensor, lengths: torch.Tensor, max_len: int, sr_in: int,
                    out_starts: torch.Tensor, total_out: int) -> torch.Tensor:
    """The rate-conversion half of librosa.load(path, sr=16000) (pipeline1.py:102): clip i (``lengths[i]`` samples at
    ``starts[i]``, ``sr_in`` Hz) -> ceil(lengths[i] * 16000 / sr_in) float32 samples at ``out_starts[i]`` of the result."""
    _need_cuda(audio, starts, lengths, out_starts)
    lib = _lib.load()
    dev = audio.device
    out = torch.zeros((int(total_out),), dtype=torch.float32, device=dev)
    n = int(starts.shape[0])
    if n == 0:
        return out
    if audio.dtype not in (torch.float32, torch.int16):
        raise DysError("audio must be float32 or int16")
    with torch.cuda.device(dev):
        _lib.check(lib.dys_init(), "dys_init")
        a, s, ln, os_ = audio.contiguous(), starts.contiguous(), lengths.contiguous(), out_starts.contiguous()
        _lib.check(lib.dys_resample_to_16k(a.data_ptr(), 1 if a.dtype == torch.int16 else 0, int(sr_in), s.data_ptr(), ln.data_ptr(),
                                           n, int(max_len), out.data_ptr(), os_.data_ptr(),
                                           torch.cuda.current_stream(dev).cuda_stream), "dys_resample_to_16k")
    return out


@resample_to_16k.register_fake
def _(audio, starts, lengths, max_len, sr_in, out_starts, total_out):
    return audio.new_empty((total_out,), dtype=torch.float32)
