"""Multi-GPU plumbing: independent inputs shard over ranks (one process per GPU).

The reference has no distributed mode (running N calculator processes is its only way to use N
cores).  Here the batch dimension is partitioned; the circuit is replicated.  The only
collective on the path is the one-time broadcast of the circuit description from rank 0 (NCCL on
GPUs, gloo in the CPU tests), after which every rank lowers the tape itself; timing and status are
reduced with MAX / gathered.  No data-path collective: witnesses stay on (or are written from) the rank
that computed them.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous block [lo, hi) of `total` independent inputs owned by `rank`"""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_blob(blob: bytes | None, rank: int, world: int, device: str = "cpu") -> bytes:
    """circuit description from rank 0 to every rank (one collective at load time)"""
    if world == 1:
        assert blob is not None
        return blob
    import torch
    import torch.distributed as dist
    n = torch.tensor([len(blob) if rank == 0 else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, 0)
    if rank == 0:
        buf = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
    else:
        buf = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


def all_reduce_max(values: List[float], world: int, device: str = "cpu") -> List[float]:
    if world == 1:
        return list(values)
    import torch
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def gather_int32(local: np.ndarray, sizes: List[int], rank: int, world: int, device: str = "cpu") -> np.ndarray | None:
    """per-instance status words of all shards on rank 0 (tiny)"""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    m = max(sizes)
    pad = np.zeros(m, dtype=np.int32)
    pad[: local.shape[0]] = local
    t = torch.from_numpy(pad).to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    if rank != 0:
        return None
    return np.concatenate([o.cpu().numpy()[: sizes[r]] for r, o in enumerate(out)])
