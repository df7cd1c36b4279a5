"""World-size-2 test of the N>1 host logic on CPU (gloo): circuit broadcast, identical lowering on
every rank, batch sharding, status gather.  The tape is executed by tests/hostsim (no GPU here)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from circom_b200.distributed import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_batch():
    for total in (1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            got = [shard_range(total, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == total
            assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from circom_b200.circuit import CircuitDesc
    from circom_b200 import circuits as C
    from circom_b200.distributed import broadcast_blob, gather_int32, all_reduce_max
    from circom_b200.witness_calculator import Circuit
    from tests.util import hostsim_run
    import random
    blob = None
    if rank == 0:
        d0 = CircuitDesc("bn128")
        d0.set_main(C.less_than(d0, 16))
        blob = d0.to_bytes()
    blob = broadcast_blob(blob, rank, world)
    c = Circuit(blob, host_only=True)          # every rank lowers the same description
    sig = np.array([c.stats["n_tape_ops"], c.stats["n_levels"], c.n_witness, int(c.witness2signal().sum())], dtype=np.int64)
    t = torch.from_numpy(sig.copy())
    dist.broadcast(t, 0)
    assert (t.numpy() == sig).all(), "ranks lowered different tapes"
    # the LOWERED circuit travels too (what cw_circuit_broadcast sends over NCCL on GPUs): only rank 0 lowered it
    low = broadcast_blob(c.serialize() if rank == 0 else None, rank, world)
    c2 = Circuit.deserialize(low)
    assert c2.stats == c.stats and c2.serialize() == low
    for x, y in zip(c.tape(), c2.tape()):
        assert (x == y).all()
    assert (c2.witness2signal() == c.witness2signal()).all() and c2.input_signal_size("in") == 2
    # the global batch, deterministic on every rank; each rank computes only its shard
    total = 11
    rng = random.Random(42)
    ins = [{"in": [rng.randrange(65536), rng.randrange(65536)]} for _ in range(total)]
    lo, hi = shard_range(total, rank, world)
    d = CircuitDesc("bn128")
    d.set_main(C.less_than(d, 16))
    wit, st, _, _ = hostsim_run(d, ins[lo:hi])
    outs = [int(w[1, 0]) for w in wit]         # LessThan output signal
    sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    allst = gather_int32(np.array(outs, dtype=np.int32), sizes, rank, world)
    tmax = all_reduce_max([float(rank + 1)], world)
    assert tmax == [float(world)]
    if rank == 0:
        exp = [int(i["in"][0] < i["in"][1]) for i in ins]
        assert allst.tolist() == exp
        open(out_path, "w").write("ok")
    dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "ok.txt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
