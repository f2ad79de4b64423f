import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
import dysb200 as pkg
fe = pkg.frontend
torch.cuda.set_device(0)
base = torch.from_numpy(pkg.synth.synth_batch(100))
host = (base.repeat(100, 1) * 0.9).pin_memory()
raw = torch.empty((10000, 149)).pin_memory(); clean = torch.empty((10000, 149)).pin_memory()
for chunk in (296, 394, 400, 444, 592, 888):
    for _ in range(3):
        fe.extract_features_host(host, chunk_clips=chunk, out_raw=raw, out_clean=clean, compute_streams=3)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(6):
        fe.extract_features_host(host, chunk_clips=chunk, out_raw=raw, out_clean=clean, compute_streams=3)
    torch.cuda.synchronize()
    print(f"f32 chunk {chunk}: {(time.perf_counter() - t0) / 6 * 1e3:.2f} ms", flush=True)
