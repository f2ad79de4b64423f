#!/usr/bin/env python
"""Chunk-size / stream-count sweep of the PCM-16 host streaming path (10 000 3-s clips, pinned int16)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dysb200 as pkg
fe = pkg.frontend
torch.cuda.set_device(0)
base = torch.from_numpy(pkg.synth.synth_batch(100))
host = (base.repeat(100, 1) * 32768.0 * 0.9).round().clamp(-32768, 32767).to(torch.int16).pin_memory()
raw = torch.empty((10000, 149)).pin_memory(); clean = torch.empty((10000, 149)).pin_memory()
for streams in (3,):
    for chunk in (592, 740, 800, 888, 1036, 1184, 1480):
        for _ in range(3):
            fe.extract_features_host(host, chunk_clips=chunk, out_raw=raw, out_clean=clean, compute_streams=streams)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(6):
            fe.extract_features_host(host, chunk_clips=chunk, out_raw=raw, out_clean=clean, compute_streams=streams)
        torch.cuda.synchronize()
        print(f"streams {streams} chunk {chunk}: {(time.perf_counter() - t0) / 6 * 1e3:.2f} ms", flush=True)
