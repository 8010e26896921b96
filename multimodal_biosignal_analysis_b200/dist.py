"""Sharding of surrogates / permutations across ranks (one process per GPU, torch.distributed).

The path shards by index range with no data-path collective: every rank holds the (small)
replicated inputs, computes a contiguous slice of the surrogate / permutation indices and only the
per-index max-statistic vectors (all_gather) and the exceedance histograms (all_reduce) are
exchanged - NCCL over NVLink on GPUs, gloo in the CPU tests.  Indices are global, so results do
not depend on the number of ranks.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int | None = None, world_size: int | None = None) -> tuple[int, int]:
    """Contiguous slice [begin, end) of range(n) owned by ``rank`` (sizes differ by at most 1)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(n, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def all_gather_ranges(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """Concatenate per-rank slices produced with ``shard_range`` into the full length-n vector."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_range(n_total, r, ws) for r in range(ws)]
    width = max(e - b for b, e in sizes)
    if n_total == width * ws:                                    # equal slices: one collective, no repacking
        out = torch.empty(n_total, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    padded = torch.zeros(width, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    out = torch.empty(width * ws, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * width: r * width + (e - b)] for r, (b, e) in enumerate(sizes)])


_GATHER_INDEX: dict = {}


def all_gather_frequency_slices(exceed: torch.Tensor, max_local: torch.Tensor, f_range) -> tuple:
    """Combine a null whose FREQUENCY axis is split with ``shard_range(F)``: rank r owns ``exceed[f_b:f_e]`` and the
    maxima of all surrogates over those bins.  One all-gather carries [slice of the counts | float bits of the
    maxima] of every rank; returns (full counts, element-wise max over the ranks' maxima).  The host side is a
    handful of launches (pack, all-gather, one row gather, one max): at 8 ranks a 10,000-surrogate null is a
    sub-millisecond job and every extra launch shows."""
    rank, ws = world()
    if ws == 1:
        return exceed, max_local
    F = exceed.shape[0]
    per_bin = exceed[0].numel()
    sizes = [shard_range(F, r, ws) for r in range(ws)]
    width = max(e - b for b, e in sizes)
    n_s = max_local.numel()
    fb, fe = (int(f_range[0]), int(f_range[1])) if f_range is not None else sizes[rank]
    n_cnt = width * per_bin
    mine = torch.empty(n_cnt + n_s, dtype=torch.int32, device=exceed.device)
    mine[: (fe - fb) * per_bin] = exceed[fb:fe].reshape(-1).view(torch.int32)
    if fe - fb < width:
        mine[(fe - fb) * per_bin: n_cnt] = 0
    mine[n_cnt:] = max_local.contiguous().view(torch.int32)
    out = torch.empty(ws * mine.numel(), dtype=torch.int32, device=exceed.device)
    dist.all_gather_into_tensor(out, mine)
    out = out.view(ws, mine.numel())
    rows = out[:, :n_cnt].reshape(ws * width, per_bin)           # row r * width + k = bin sizes[r][0] + k
    if F == ws * width:
        full = rows
    else:
        key = (F, ws, str(exceed.device))
        idx = _GATHER_INDEX.get(key)
        if idx is None:
            idx = torch.tensor([r * width + k for r, (b, e) in enumerate(sizes) for k in range(e - b)],
                               dtype=torch.int64, device=exceed.device)
            _GATHER_INDEX[key] = idx
        full = rows.index_select(0, idx)
    max_stat = out[:, n_cnt:].view(torch.float32).amax(dim=0)
    return full.view(exceed.dtype).view(exceed.shape), max_stat


def round_robin(n: int, rank: int | None = None, world_size: int | None = None) -> range:
    """Indices of range(n) owned by ``rank`` when independent units (subject-conditions) are dealt out in turn."""
    if rank is None or world_size is None:
        rank, world_size = world()
    return range(rank, n, world_size)


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    _, ws = world()
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def all_reduce_max_(t: torch.Tensor) -> torch.Tensor:
    _, ws = world()
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


def bind_to_gpu_numa_node(device_index: int | None = None) -> int | None:
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off (one process per GPU): pinned host
    buffers allocated afterwards are node-local, so several ranks streaming recordings over PCIe do not all pull
    from one socket's memory.  Returns the node, or None when the topology cannot be read (then nothing changes)."""
    import os
    try:
        if device_index is None:
            device_index = torch.cuda.current_device()
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:                                             # no sysfs, no such attribute, not permitted, ...
        return None

