"""Multi-GPU plumbing: independent inputs shard over ranks (one process per GPU).

The reference has no distributed mode (running N calculator processes is its only way to use N
cores).  Here the batch dimension is partitioned; the circuit is replicated.  Collectives:
the one-time broadcast of the circuit from rank 0 - either its description (`broadcast_blob`, any backend: every
rank lowers) or, through the library's own NCCL communicator (`Comm`, include/circom_b200.h cw_comm_*), the LOWERED
tape + CSR + dictionaries (`Comm.broadcast_circuit`: only rank 0 lowers) -, the gather of packed witness records on a
root (`Comm.gather_witness_packed`, grouped ncclSend/ncclRecv over NVLink) and the all-reduce of the per-instance
status counters.  Timing is reduced with MAX.
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


class Comm:
    """The library's NCCL communicator (cw_comm_*): created from a unique id that rank 0 generates and the host
    program ships to the other ranks - here through torch.distributed, whatever its backend."""

    def __init__(self, rank: int, world: int, device: int):
        import ctypes
        import torch
        import torch.distributed as dist
        from .native import lib, check
        self.rank, self.world, self.device = rank, world, device
        ident = (ctypes.c_uint8 * 128)()
        if rank == 0:
            check(lib.cw_comm_unique_id(ident))
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor(list(ident), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0)
        ident = (ctypes.c_uint8 * 128)(*t.cpu().tolist())
        self._h = ctypes.c_void_p()
        check(lib.cw_comm_init(ident, rank, world, device, ctypes.byref(self._h)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            from .native import lib
            lib.cw_comm_destroy(h)

    def broadcast_circuit(self, circuit, root: int = 0):
        """the root's lowered circuit on every rank (one NCCL broadcast of the blob; non-root ranks do not lower)"""
        import ctypes
        from .native import lib, check
        from .witness_calculator import Circuit
        h = ctypes.c_void_p(circuit._h.value if circuit is not None else None)
        check(lib.cw_circuit_broadcast(self._h, ctypes.byref(h), root))
        return circuit if self.rank == root else Circuit.from_handle(h)

    def stats(self):
        import ctypes
        from .native import lib, check
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        check(lib.cw_comm_stats(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def gather_witness_packed(self, batch, first: int, count: int, root: int = 0):
        """packed records of instances [first, first+count) of every rank's batch on the root's GPU.
        Returns (torch uint32-as-int32 tensor [world][count][words] on the root else None, device ms)"""
        import ctypes
        import torch
        from .native import lib, check
        info, _ = batch.circuit.pack_info(entries=False)
        words = info[0]
        ms = ctypes.c_float()
        if self.rank == root:
            recv = torch.empty((self.world, count, words), dtype=torch.int32, device="cuda")
            check(lib.cw_batch_gather_witness_packed(self._h, batch._h, first, count, root, ctypes.c_void_p(recv.data_ptr()),
                                                     None, ctypes.byref(ms)))
            return recv, ms.value
        send = torch.empty((count, words), dtype=torch.int32, device="cuda")
        check(lib.cw_batch_gather_witness_packed(self._h, batch._h, first, count, root, None,
                                                 ctypes.c_void_p(send.data_ptr()), ctypes.byref(ms)))
        return None, ms.value

    def status_allreduce(self, batch):
        import ctypes
        from .native import lib, check
        out = (ctypes.c_uint64 * 2)()
        check(lib.cw_status_allreduce(self._h, batch._h, out))
        return int(out[0]), int(out[1])


def gather_witness_packed(batch, circuit, rank: int, world: int, count: int, reps: int = 3, comm: "Comm | None" = None):
    """bench leg: gather `count` packed witnesses per rank on rank 0, `reps` times; returns the figures on rank 0"""
    import torch
    import torch.distributed as dist
    own = comm is None
    if own:
        comm = Comm(rank, world, torch.cuda.current_device())
    info, _ = circuit.pack_info(entries=False)
    best = None
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        dist.barrier()
        recv, ms = comm.gather_witness_packed(batch, 0, count, 0)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        best = ms if best is None else min(best, ms)
        del recv
    sent, received = comm.stats()
    res = {"witnesses_per_rank": count, "bytes_per_witness_packed": info[0] * 4, "ms": best,
           "root_receive_GBps": (world - 1) * count * info[0] * 4 / (best / 1e3) / 1e9,
           "dense_equivalent_GBps": (world - 1) * count * circuit.n_witness * 32 / (best / 1e3) / 1e9,
           "comm_bytes_received_on_root": received,
           "collective": "grouped ncclSend/ncclRecv of packed records (cw_batch_gather_witness_packed), max over ranks of the device time of pack + transfer"}
    if own:
        del comm
    return res
