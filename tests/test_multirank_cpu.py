"""N > 1 host logic on CPU: two gloo ranks shard the reference's 905-row feature matrix like the
GPUs shard clips (contiguous ranges), all-reduce the 299 float64 moments -- the one
collective of the path -- and must land on the reference's committed StandardScaler
(output_results/scaler_after.pkl, fixture tests/golden/ref_scaler_after.npz).

The per-rank moment kernel itself (dys_cmvn_accumulate) needs a GPU and is covered by the -m gpu
tests; here the checker's numpy moments stand in for it so that sharding, the collective, the
two-pass schedule and the finalisation are exercised without a device.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG_NAME, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _moments_about(X, shift):
    Xd = X.astype(np.float64) - shift
    return np.concatenate(([X.shape[0]], Xd.sum(axis=0), (Xd ** 2).sum(axis=0)))


def _worker(rank, world, port, out_dir):
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module(PKG_NAME)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", "ref_scaler_after.npz"))
        lo, hi = pkg.sharding.shard_range(g["X"].shape[0], rank, world)
        X = g["X"][lo:hi]
        sc = pkg.scaler
        # the default schedule: ONE all-reduce of the moments about zero (GlobalScaler.fit) ...
        acc0 = sc.allreduce_moments(torch.from_numpy(_moments_about(X, 0.0)))
        mean_1, var_1, scale_1, n_1 = sc.finalize_moments(acc0, None)
        # ... and the two-pass one (GlobalScaler(two_pass=True)): global mean, then moments about it
        mean0 = (acc0[1:150] / acc0[0]).numpy()
        acc1 = sc.allreduce_moments(torch.from_numpy(_moments_about(X, mean0)))
        mean, var, scale, n = sc.finalize_moments(acc1, mean0)
        np.testing.assert_allclose(mean_1, mean, rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(var_1, var, rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(scale_1, scale, rtol=1e-11, atol=1e-13)
        mean, var, scale, n = mean_1, var_1, scale_1, n_1        # the reference's pickle is compared with the one-pass result
        # every rank holds the same statistics afterwards
        gathered = [torch.zeros(149, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(mean))
        assert all(torch.equal(gathered[0], t) for t in gathered)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), mean=mean, var=var, scale=scale, n=n, lo=lo, hi=hi)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_cmvn_allreduce_matches_reference_scaler(tmp_path, golden_dir):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = np.load(os.path.join(golden_dir, "ref_scaler_after.npz"))
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert int(r0["lo"]) == 0 and int(r0["hi"]) == int(r1["lo"]) and int(r1["hi"]) == 905
    for r in (r0, r1):
        assert int(r["n"]) == 905
        np.testing.assert_allclose(r["mean"], g["mean"], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(r["var"], g["var"], rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(r["scale"], g["scale"], rtol=1e-11, atol=1e-13)
        assert np.all(r["scale"][144:] == 1.0)


def test_bench_shard_plan_covers_every_clip_once(pkg):
    """bench.py gives rank r the contiguous clip range shard_range(total, r, world)."""
    for world in (1, 2, 4, 8):
        total = 10000 * world
        seen = np.zeros(total, dtype=np.int32)
        for r in range(world):
            lo, hi = pkg.sharding.shard_range(total, r, world)
            assert hi - lo == 10000
            seen[lo:hi] += 1
        assert (seen == 1).all()
