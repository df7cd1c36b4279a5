#!/usr/bin/env python
"""Where the end-to-end time of a single-chunk config goes (development probe, not a bench line):
host inputs -> set_inputs -> run -> witness (packed D2H + host expansion), each timed with a synchronisation."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="sha256compression")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--chunks", type=int, default=1)
    a = ap.parse_args()
    import bench, torch
    from circom_b200.witness_calculator import Circuit, Batch, aligned_empty
    desc, label, _ = bench.make_workload(argparse.Namespace(workload=a.workload, batch_per_gpu=0, lanes=8, chain=132))
    c = Circuit(desc)
    W = c.n_witness
    inputs = bench.synth_inputs(desc, a.workload, a.batch, 7)
    pin = torch.from_numpy(inputs.view(np.int64)).pin_memory().numpy().view(np.uint64)
    chunk = a.batch // a.chunks
    pair = [Batch(c, chunk, 0), Batch(c, chunk, 0)]
    outs = [aligned_empty((chunk, W, 4)) for _ in range(2)]
    for rep in range(4):
        t = [time.perf_counter()]
        b = pair[0]
        b.set_inputs(pin[:chunk]); torch.cuda.synchronize(); t.append(time.perf_counter())
        b.run(sync=True); t.append(time.perf_counter())
        b.witness_async(outs[0]); b.witness_wait(); t.append(time.perf_counter())
        print("serial  chunk %d: set_inputs %.2f ms, run %.2f ms (kernel %.2f), witness %.2f ms (d2h %.1f MB)" % (
            chunk, 1e3 * (t[1] - t[0]), 1e3 * (t[2] - t[1]), b.last_ms()[0], 1e3 * (t[3] - t[2]), b.last_d2h_bytes() / 1e6), flush=True)
    for rep in range(4):
        t0 = time.perf_counter()
        infl = [False, False]
        for k in range(a.chunks):
            B = pair[k & 1]
            if infl[k & 1]:
                B.witness_wait()
            B.set_inputs(pin[k * chunk:(k + 1) * chunk]); B.run(sync=False); B.witness_async(outs[k & 1]); infl[k & 1] = True
        for j in range(2):
            if infl[j]:
                pair[j].witness_wait()
        dt = time.perf_counter() - t0
        print("pipelined %d x %d: %.2f ms -> %.0f witnesses/s" % (a.chunks, chunk, 1e3 * dt, a.batch / dt), flush=True)


if __name__ == "__main__":
    main()
