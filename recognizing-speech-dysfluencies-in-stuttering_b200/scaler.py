"""Global feature standardisation ("CMVN") across clips -- and across GPUs.

Mirror of ``StandardScaler().fit(X)`` / ``.transform(X)`` in the reference
(/root/reference/pipeline1.py:470-473; main1.py:848-852): float64 per-feature mean and
population variance over ALL clips, ``scale_ = sqrt(var_)`` with constant features -> 1.0.

Each rank reduces its own [N_rank, 149] feature block on the GPU (``dys_cmvn_accumulate``)
to 299 doubles; the ONE collective of the whole path is ONE sum all-reduce of that vector
(NCCL over NVLink when ranks are GPUs; gloo in the CPU tests).  The moments are taken about zero in
float64 with a fixed summation order: for this feature set (|mean| / std <= ~10) the variance keeps
more than 13 digits, inside the 1e-11 the tests hold it to against scikit-learn's own result
(output_results/scaler_after.pkl).  ``GlobalScaler(two_pass=True)`` re-centres on the all-reduced
mean with a second pass and a second all-reduce (sklearn's corrected two-pass form) for feature
sets whose mean dwarfs their spread.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import CMVN_ACC_LEN, CMVN_PARTIALS, FEATURE_LEN


def _allreduce_sum(vec: torch.Tensor, group) -> torch.Tensor:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        backend = dist.get_backend(group)
        if backend == "gloo" and vec.is_cuda:
            host = vec.cpu()
            dist.all_reduce(host, op=dist.ReduceOp.SUM, group=group)
            vec.copy_(host)
        else:
            dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec


def allreduce_moments(acc: torch.Tensor, group=None) -> torch.Tensor:
    """The path's one collective: in-place sum all-reduce of a rank's float64 [299] moment vector
    ([n, sum_f (x - s), sum_f (x - s)^2]) over the process group (NCCL on GPUs, gloo on CPU)."""
    if acc.dtype != torch.float64 or acc.numel() != CMVN_ACC_LEN:
        raise ValueError("acc must be a float64 vector of 299 elements")
    return _allreduce_sum(acc, group)


def finalize_moments(acc, shift=None):
    """Host float64 mirror of dys_cmvn_finalize: (mean, var, scale, n) from all-reduced moments about
    ``shift`` (None = 0); constant features get scale 1.0 like sklearn's StandardScaler."""
    acc = np.asarray(acc.cpu() if isinstance(acc, torch.Tensor) else acc, dtype=np.float64)
    n = acc[0]
    m1 = acc[1:1 + FEATURE_LEN] / n
    mean = m1 + (0.0 if shift is None else np.asarray(shift, dtype=np.float64))
    var = np.maximum(acc[1 + FEATURE_LEN:] / n - m1 * m1, 0.0)
    eps = np.finfo(np.float64).eps
    scale = np.sqrt(var)
    const = (var <= n * eps * var + (n * mean * eps) ** 2) | (scale < 10 * eps)
    scale[const] = 1.0
    return mean, var, scale, int(n)


class GlobalScaler:
    """``fit`` / ``transform`` with sklearn's attribute names (mean_, var_, scale_, n_samples_seen_)."""

    def __init__(self, group=None, two_pass: bool = False):
        self.group = group
        self.two_pass = bool(two_pass)
        self.mean_ = self.var_ = self.scale_ = None
        self._n = None

    def _moments(self, X: torch.Tensor, shift: torch.Tensor | None) -> torch.Tensor:
        lib = _lib.load()
        acc = torch.empty(CMVN_ACC_LEN, dtype=torch.float64, device=X.device)
        partials = torch.empty(CMVN_PARTIALS, dtype=torch.float64, device=X.device)
        with torch.cuda.device(X.device):
            _lib.check(lib.dys_cmvn_accumulate(X.data_ptr() if X.numel() else None, X.shape[0],
                                               shift.data_ptr() if shift is not None else None, acc.data_ptr(),
                                               partials.data_ptr(), torch.cuda.current_stream(X.device).cuda_stream),
                       "dys_cmvn_accumulate")
        return acc

    def fit(self, X: torch.Tensor):
        """X: this rank's float32 [N_rank, 149] CUDA tensor. Collective when torch.distributed is initialised."""
        if not (isinstance(X, torch.Tensor) and X.is_cuda and X.dtype == torch.float32 and X.dim() == 2
                and X.shape[1] == FEATURE_LEN):
            raise ValueError("X must be a float32 CUDA tensor of shape [N, 149]")
        X = X.contiguous()
        lib = _lib.load()
        acc1 = _allreduce_sum(self._moments(X, None), self.group)           # the one collective: 299 float64
        n = acc1[0]
        mean0 = None
        if self.two_pass:
            mean0 = (acc1[1:1 + FEATURE_LEN] / n).contiguous()
            acc1 = _allreduce_sum(self._moments(X, mean0), self.group)
        mean = torch.empty(FEATURE_LEN, dtype=torch.float64, device=X.device)
        scale = torch.empty(FEATURE_LEN, dtype=torch.float64, device=X.device)
        with torch.cuda.device(X.device):
            _lib.check(lib.dys_cmvn_finalize(acc1.data_ptr(), mean0.data_ptr() if mean0 is not None else None, mean.data_ptr(),
                                             scale.data_ptr(), torch.cuda.current_stream(X.device).cuda_stream),
                       "dys_cmvn_finalize")
        m1 = acc1[1:1 + FEATURE_LEN] / n
        self.var_ = torch.clamp(acc1[1 + FEATURE_LEN:] / n - m1 * m1, min=0.0)
        self.mean_, self.scale_ = mean, scale
        self._n = n                      # device scalar: reading it would force a host sync inside fit()
        return self

    @property
    def n_samples_seen_(self) -> int:
        return 0 if self._n is None else int(self._n.item())

    def transform(self, X: torch.Tensor) -> torch.Tensor:
        if self.mean_ is None:
            raise RuntimeError("GlobalScaler.transform called before fit")
        lib = _lib.load()
        X = X.contiguous()
        out = torch.empty_like(X)
        with torch.cuda.device(X.device):
            _lib.check(lib.dys_cmvn_apply(X.data_ptr() if X.numel() else None, X.shape[0], self.mean_.data_ptr(),
                                          self.scale_.data_ptr(), out.data_ptr() if X.numel() else None,
                                          torch.cuda.current_stream(X.device).cuda_stream), "dys_cmvn_apply")
        return out

    def fit_transform(self, X: torch.Tensor) -> torch.Tensor:
        return self.fit(X).transform(X)

    def to_sklearn(self):
        """A fitted ``sklearn.preprocessing.StandardScaler`` carrying this fit, i.e. the object the reference
        pickles as ``output_results/scaler_after.pkl`` (main1.py:851) and its sidebar predictor reloads
        (main1.py:958-987): the reference scripts keep working on statistics computed on the GPUs."""
        if self.mean_ is None:
            raise RuntimeError("GlobalScaler.to_sklearn called before fit")
        from sklearn.preprocessing import StandardScaler
        sk = StandardScaler()
        sk.mean_ = self.mean_.cpu().numpy().copy()
        sk.var_ = self.var_.cpu().numpy().copy()
        sk.scale_ = self.scale_.cpu().numpy().copy()
        sk.n_samples_seen_ = np.int64(self.n_samples_seen_)
        sk.n_features_in_ = FEATURE_LEN
        return sk


def classifier_inputs(X: torch.Tensor, scaler: "GlobalScaler | None" = None, group=None, gather: bool = True):
    """The reference classifier's input loader (pipeline1.py:455-473; main1.py:847-852,987) for features that
    are already on the GPUs: fit the global scaler over every rank's rows (one all-reduce), standardise on the
    device, and hand sklearn a host float32 matrix in the reference's row order.

    X      this rank's float32 [N_rank, 149] CUDA tensor (rows in shard order, see sharding.shard_range)
    scaler a fitted GlobalScaler to apply (inference: main1.py:987) or None to fit one here (training)
    gather True: every rank receives all ranks' standardised rows (all-gather of the row blocks)
    Returns (Z float32 numpy [N or N_rank, 149], scaler)."""
    import torch.distributed as dist
    if scaler is None:
        scaler = GlobalScaler(group).fit(X)
    Z = scaler.transform(X)
    if gather and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        counts = torch.zeros(world, dtype=torch.int64, device=Z.device)
        counts[dist.get_rank(group)] = Z.shape[0]
        dist.all_reduce(counts, group=group)
        blocks = [torch.empty((int(c), FEATURE_LEN), dtype=Z.dtype, device=Z.device) for c in counts.tolist()]
        dist.all_gather(blocks, Z, group=group)
        Z = torch.cat(blocks, dim=0)
    host = torch.empty(Z.shape, dtype=Z.dtype).pin_memory()
    host.copy_(Z, non_blocking=True)
    torch.cuda.current_stream(Z.device).synchronize()
    return host.numpy(), scaler
