"""B200-native audio front-end of kishormb/Recognizing-Speech-Dysfluencies-in-Stuttering.

Drop-in for the reference's feature path only (16 kHz clip -> ``*_raw_feats.npy`` /
``*_clean_feats.npy``); the classifier, UI and reporting stay the reference's own code.
Importing the package does not need a GPU; calling into it does (no CPU fallback).
"""
from . import _lib, sharding, synth, wavio  # noqa: F401
from ._lib import (AUDIO_FEATURE_LEN, FEATURE_LEN, SAMPLE_RATE, STATUS_BAD_LENGTH,  # noqa: F401
                   STATUS_CLEAN_FALLBACK, STATUS_NONFINITE, STATUS_SHORT, DysError)

__all__ = ["frontend", "scaler", "torch_ops", "sharding", "synth", "wavio", "DysError", "FEATURE_LEN"]


def __getattr__(name):
    # torch-dependent modules load lazily so that table/ABI checks stay light
    if name in ("frontend", "scaler", "torch_ops"):            # torch_ops registers torch.ops.dysb200.*
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
