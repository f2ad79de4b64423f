"""Where the host streaming path's time above the pure-copy floor goes: the public call with and without the denoise
branch (raw only needs a fifth of the kernel time), the scaler fit alone, and the per-chunk timeline at the default
configuration."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dysb200 as pkg
fe, sc = pkg.frontend, pkg.scaler
N, L = 10000, 48000
host = torch.from_numpy(pkg.synth.synth_batch(100)).repeat(N // 100, 1).contiguous().pin_memory()
dev = torch.device("cuda", 0)
out_raw = torch.empty((N, 149)).pin_memory(); out_clean = torch.empty((N, 149)).pin_memory()
d_all = torch.empty((N, L), dtype=torch.float32, device=dev)


def dev_ms(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


print("copy only            %.2f ms" % dev_ms(lambda: d_all.copy_(host, non_blocking=True)))
for chunk, ns in ((400, 3), (200, 3)):
    print("host path raw only   %.2f ms  (chunk %d, %d streams)" % (dev_ms(lambda: fe.extract_features_host(
        host, denoise=False, chunk_clips=chunk, out_raw=out_raw, compute_streams=ns)), chunk, ns))
    print("host path raw+clean  %.2f ms  (chunk %d, %d streams)" % (dev_ms(lambda: fe.extract_features_host(
        host, denoise=True, chunk_clips=chunk, out_raw=out_raw, out_clean=out_clean, compute_streams=ns)), chunk, ns))
s = sc.GlobalScaler()
print("scaler fit from pinned rows  %.3f ms" % dev_ms(lambda: s.fit(out_clean.to(dev, non_blocking=True))))
print("device-resident raw+clean    %.2f ms" % dev_ms(lambda: fe.extract_features_batch(d_all, denoise=True)))
