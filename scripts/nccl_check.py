#!/usr/bin/env python
"""Multi-GPU plumbing check (run under torchrun, one rank per GPU): the library's NCCL communicator, ONE broadcast of
the LOWERED circuit from rank 0 (the other ranks never lower), shards of independent inputs, the gather of packed witness
records on rank 0 (decoded there and compared with the oracle for EVERY rank's instances) and the status all-reduce.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/nccl_check.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from circom_b200.circuit import CircuitDesc
from circom_b200 import circuits as C
from circom_b200.distributed import Comm, shard_range
from circom_b200.witness_calculator import Circuit, Batch
from tests.util import flat_inputs


def main():
    rank, world, dev = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    comm = Comm(rank, world, dev)
    d = CircuitDesc("bn128")
    d.set_main(C.ecdsa_scale(d, 2, 5))
    t0 = time.time()
    c0 = Circuit(d) if rank == 0 else None            # only rank 0 lowers
    circuit = comm.broadcast_circuit(c0, 0)
    t_bc = time.time() - t0
    # the global batch is deterministic; every rank computes its shard
    total = 64 * world + 5
    rng = np.random.default_rng(3)
    ins = [{"a": [int(x) for x in rng.integers(0, 2**63, 8)], "b": [int(x) for x in rng.integers(0, 2**63, 8)]} for _ in range(total)]
    lo, hi = shard_range(total, rank, world)
    n = 64                                             # gather the first 64 instances of every shard
    b = Batch(circuit, hi - lo, dev)
    b.set_inputs(flat_inputs(d, ins[lo:hi]))
    b.run()
    bad = comm.status_allreduce(b)
    assert bad == (0, 0), bad
    recv, ms = comm.gather_witness_packed(b, 0, n, 0)
    sent, received = comm.stats()
    if rank == 0:
        from oracle.ir_eval import evaluate
        info, ent = circuit.pack_info()
        off = [0, info[1], info[1] + info[2], info[1] + info[2] + 2 * info[3]]
        rec = recv.cpu().numpy().view(np.uint32)
        w2s = circuit.witness2signal().astype(np.int64)
        ent = ent.tolist()
        for r in range(world):
            rlo, _ = shard_range(total, r, world)
            for i in (0, n - 1):
                exp = evaluate(d, ins[rlo + i])
                got = []
                for e in ent:
                    cls, idx = e >> 30, e & 0x3FFFFFFF
                    if cls <= 1:
                        got.append((int(rec[r, i, off[cls] + (idx >> 5)]) >> (idx & 31)) & 1)
                    elif cls == 2:
                        got.append(int(rec[r, i, off[2] + 2 * idx]) | (int(rec[r, i, off[2] + 2 * idx + 1]) << 32))
                    else:
                        got.append(sum(int(rec[r, i, off[3] + 8 * idx + k]) << (32 * k) for k in range(8)))
                assert got == [exp[k] for k in w2s], (r, i)
        assert received >= (world - 1) * n * info[0] * 4
        print("nccl_check ok: world %d, lowered-circuit broadcast %.2f s, gather of %d packed witnesses per rank %.3f ms, "
              "rank 0 received %d bytes" % (world, t_bc, n, ms, received))
    else:
        assert sent >= n * circuit.pack_info(entries=False)[0][0] * 4 and circuit.stats["n_tape_ops"] > 0
    dist.barrier()
    del comm
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
