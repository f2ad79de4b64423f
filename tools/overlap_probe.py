"""Probe: does running two half-batches concurrently on two streams (their gate and feature phases interleave on
the SMs) beat one full batch?  Device-resident inputs, CUDA events.  Informs the sub-batch scheduling only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dysb200 as pkg
fe = pkg.frontend
N, L = 10000, 48000
base = torch.from_numpy(pkg.synth.synth_batch(100)).cuda()
X = base.repeat(N // 100, 1).contiguous()
streams = [torch.cuda.Stream() for _ in range(4)]

def run_split(k):
    cur = torch.cuda.current_stream()
    outs = []
    per = N // k
    for i in range(k):
        s = streams[i]
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            xi = X[i * per:(i + 1) * per]
            starts = torch.arange(per, dtype=torch.int64, device="cuda") * L
            lens = torch.full((per,), L, dtype=torch.int32, device="cuda")
            outs.append(fe._run_device(xi.reshape(-1), starts, lens, L, True, 1.0, False, None, 1, slot=20 + i))
    for i in range(k):
        cur.wait_stream(streams[i])
    return outs

def timeit(fn, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

print("one batch      : %.2f ms" % timeit(lambda: fe.extract_features_batch(X, denoise=True)))
for k in (1, 2, 3, 4):
    print("%d streams      : %.2f ms" % (k, timeit(lambda: run_split(k))))
a = fe.extract_features_batch(X, denoise=True)
b = run_split(2)
torch.cuda.synchronize()
print("identical:", torch.equal(a[0], torch.cat([o[0] for o in b])), torch.equal(a[1], torch.cat([o[1] for o in b])))

# raw-only features of N clips on a side stream while the full raw+clean pass runs on the main stream
def both():
    cur = torch.cuda.current_stream()
    s = streams[0]
    s.wait_stream(cur)
    starts = torch.arange(N, dtype=torch.int64, device="cuda") * L
    lens = torch.full((N,), L, dtype=torch.int32, device="cuda")
    with torch.cuda.stream(s):
        r = fe._run_device(X.reshape(-1), starts, lens, L, False, 1.0, False, None, 1, slot=30)
    o = fe._run_device(X.reshape(-1), starts, lens, L, True, 1.0, False, None, 1, slot=31)
    cur.wait_stream(s)
    return r, o
t_raw = timeit(lambda: fe.extract_features_batch(X, denoise=False))
t_full = timeit(lambda: fe.extract_features_batch(X, denoise=True))
t_both = timeit(both)
print("raw-only %.2f ms, raw+clean %.2f ms, sum %.2f ms, concurrent on two streams %.2f ms" % (t_raw, t_full, t_raw + t_full, t_both))
