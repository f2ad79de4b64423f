"""Does a polling nvidia-smi slow the host-side CUDA calls of the streaming path?  60 calls without and with it."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dysb200 as pkg
fe = pkg.frontend
N, L = 10000, 48000
host = torch.from_numpy(pkg.synth.synth_batch(100)).repeat(N // 100, 1).contiguous().pin_memory()
out_raw = torch.empty((N, 149)).pin_memory(); out_clean = torch.empty((N, 149)).pin_memory()
def series(n=60, **kw):
    ts = []
    for i in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fe.extract_features_host(host, out_raw=out_raw, out_clean=out_clean, **kw)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return ts
def report(name, ts):
    s = sorted(ts)
    print("%-34s median %.1f  p90 %.1f  max %.1f  mean %.1f  slow(>55ms) %d/%d" % (name, s[len(s) // 2], s[int(len(s) * 0.9)], s[-1], sum(ts) / len(ts), sum(t > 55 for t in ts), len(ts)), flush=True)
series(5)
ref = fe.extract_features_batch(host.cuda(), denoise=True)
assert torch.equal(out_raw, ref[0].cpu()) and torch.equal(out_clean, ref[1].cpu())
for rnd in range(2):
    for chunk, ns in ((400, 3), (800, 3), (1200, 3), (800, 2)):
        report("chunk %d, %d compute streams" % (chunk, ns), series(40, chunk_clips=chunk, compute_streams=ns))
