"""Static clip sharding across the GPUs of one node.

Every clip is an independent unit (own padding, own top-dB maximum, own tuning, own denoise
state), so the hot path needs no exchange: rank r of W owns the contiguous index range
[floor(r N / W), floor((r+1) N / W)), which keeps the concatenated output rows in the
reference's sorted-path order (pipeline1.py:97).  The only collective is the CMVN all-reduce
in scaler.py.
"""
from __future__ import annotations

import os


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return (rank * n_items) // world, ((rank + 1) * n_items) // world


def spread_device_index(local_rank: int, world: int, n_visible: int) -> int:
    """GPU for a rank when fewer ranks than GPUs run on one box: alternate between the two halves of the box (rank 0 ->
    GPU 0, rank 1 -> GPU G/2, rank 2 -> GPU 1, ...).  On 8-GPU HGX-style hosts the halves hang off different host
    bridges, and host-to-device streaming is bound by what one bridge delivers (measured on this pool: 4 ranks on GPUs
    0-3 get 28.8 GB/s each, on GPUs 0, 4, 1, 5 they get 52.8 GB/s).  Identity when all GPUs are used."""
    if not (1 < world < n_visible) or n_visible % 2:
        return local_rank
    return (local_rank % 2) * (n_visible // 2) + local_rank // 2


def sliding_windows(n_samples: int, win: int = 48000, hop: int = 24000):
    """Window starts of the long-form segmenter (BASELINE config 4: win 48 000 / hop 24 000 ->
    2399 windows per hour).  Each window is an independent clip for the feature function."""
    if n_samples < win:
        return []
    return list(range(0, n_samples - win + 1, hop))


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Best effort: pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that pinned host
    staging buffers allocated afterwards are first-touched in memory local to that GPU's PCIe root.  With one
    process per GPU on a two-socket box this keeps every rank's host-to-device stream on its own socket instead of
    all eight reading one socket's DRAM (measured: 23 GB/s per GPU unbound at 8 GPUs vs 55 GB/s for one GPU alone).
    Returns what it did ({"node": n, "cpus": k}) or why it did nothing; never raises."""
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return {"skipped": "no NUMA affinity reported for " + bus}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return {"skipped": f"node {node} has no CPU this process may use"}
        os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed), "bus": bus}
    except Exception as e:  # noqa: BLE001 - purely an optimisation
        return {"skipped": f"{type(e).__name__}: {e}"}
