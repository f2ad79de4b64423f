"""Throughput on a corpus-like ragged batch (UCLASS clips: 0.45 .. 10.1 s, median 2.07 s, mean 2.39 s) against the
uniform 3-s benchmark batch: does the grid sized for the longest clip cost anything?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dysb200 as pkg
fe = pkg.frontend
rng = np.random.default_rng(0)
n = 9050
dur = np.clip(rng.lognormal(np.log(2.07), 0.55, n), 0.45, 10.1)
lens = (dur * 16000).astype(np.int64)
base = torch.from_numpy(pkg.synth.synth_clip(0, 170000)).cuda()
starts = np.zeros(n, np.int64)
total = int(((lens + 3) & ~3).sum())
audio = base.repeat(total // base.numel() + 1)[:total].contiguous()
starts[1:] = np.cumsum((lens + 3) & ~3)[:-1]
d_st, d_ln = torch.from_numpy(starts).cuda(), torch.from_numpy(lens.astype(np.int32)).cuda()

def run(order=None):
    st, ln = (d_st, d_ln) if order is None else (d_st[order], d_ln[order])
    return fe._run_device(audio, st.contiguous(), ln.contiguous(), int(lens.max()), True, 1.0, False, None, 1, slot=50)

def timeit(fn, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
secs = lens.sum() / 16000
ms = timeit(run)
print("ragged  : %d clips, %.0f audio-s, mean %.2f s, max %.2f s: %.2f ms -> %.3e audio-s/s" % (n, secs, secs / n, lens.max() / 16000, ms, secs / (ms * 1e-3)))
order = torch.from_numpy(np.argsort(lens)).cuda()
ms2 = timeit(lambda: run(order))
print("sorted  : %.2f ms -> %.3e audio-s/s" % (ms2, secs / (ms2 * 1e-3)))
ms4 = timeit(lambda: fe.extract_features_batch(audio, lengths=lens.astype(np.int32), starts=starts, denoise=True))
print("public  : %.2f ms -> %.3e audio-s/s  (extract_features_batch: length-sorted order, rows scattered back)" % (ms4, secs / (ms4 * 1e-3)))
X = base[:48000].repeat(7200, 1)
ms3 = timeit(lambda: fe.extract_features_batch(X, denoise=True))
print("uniform : 7200 x 3 s = %.0f audio-s: %.2f ms -> %.3e audio-s/s" % (7200 * 3, ms3, 7200 * 3 / (ms3 * 1e-3)))
