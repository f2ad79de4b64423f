#!/usr/bin/env python
"""Step time of mid-size batches (the strong-scaling shards and the host path's chunks) with the gate's two CTA geometries."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, time
sys.path.insert(0, %r)
import torch, dysb200 as pkg
fe = pkg.frontend
torch.cuda.set_device(0)
base = torch.from_numpy(pkg.synth.synth_batch(50)).cuda()
for n in (600, 800, 888, 1250, 2000, 2500, 5000, 10000):
    X = base.repeat((n + 49) // 50, 1)[:n].contiguous()
    for _ in range(3): fe.extract_features_batch(X, denoise=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fe.extract_features_batch(X, denoise=True)
    e1.record(); torch.cuda.synchronize()
    print(n, round(e0.elapsed_time(e1) / 10, 3), "ms", round(e0.elapsed_time(e1) / 10 / n * 1e3, 3), "us/clip", flush=True)
''' % ROOT
for mode in ("1", "0", ""):
    env = dict(os.environ)
    if mode: env["DYS_GATE_BIG"] = mode
    else: env.pop("DYS_GATE_BIG", None)
    print("DYS_GATE_BIG =", mode or "(auto)", flush=True)
    subprocess.run([sys.executable, "-c", CODE], env=env, check=True)
