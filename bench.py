#!/usr/bin/env python
"""Benchmark of the hot path: batched witness generation (+ R1CS check) on B200.

  python bench.py --gpus N --steps K --warmup W          our arm (one process per GPU under torchrun)
  python bench.py --impl reference ...                   the reference's CPU path on the host cores

Metric (BASELINE.json): witnesses/s on the ~1M-constraint BN254 circuit; the R1CS check is reported
beside it as Mconstraints/s.  A step = one pass of the hot path over one batch of synthetic inputs:
stage inputs -> execute the instruction tape -> gather the canonical witness vectors.
`value` times steps with the inputs already in HBM; `e2e` times the same through the
reference-facing API with host buffers (pinned H2D of the inputs and D2H of the witnesses inside the
timed region).  Weak scaling: every GPU processes its own batch of independent inputs; the only
collective is the one-time NCCL broadcast of the circuit description.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ecdsa_scale", choices=["ecdsa_scale", "sha256compression", "poseidon2", "sha256_512_bls"])
    ap.add_argument("--batch-per-gpu", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=8)
    ap.add_argument("--chain", type=int, default=132)
    ap.add_argument("--no-r1cs", action="store_true")
    ap.add_argument("--e2e-pinned-gb", type=float, default=12.0, help="cap of the pinned witness buffer of the e2e leg")
    ap.add_argument("--e2e-steps", type=int, default=-1, help="timed end-to-end steps (default: min(steps, 3); 0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
def make_workload(args):
    from circom_b200.circuit import CircuitDesc
    from circom_b200 import circuits as C
    d = CircuitDesc("bls12381" if args.workload == "sha256_512_bls" else "bn128")
    if args.workload == "sha256_512_bls":   # BASELINE.json configs[4]
        d.set_main(C.sha256(d, 512), "sha256_512_bls")
        return d, "Sha256(512) full preimage, BLS12-381 Fr", args.batch_per_gpu or 1024
    if args.workload == "ecdsa_scale":
        d.set_main(C.ecdsa_scale(d, args.lanes, args.chain), "ecdsa_scale_%dx%d" % (args.lanes, args.chain))
        label = "ecdsa-scale synthetic (secp256k1 BigMultModP chains %dx%d, 4x64-bit limbs), BN254" % (args.lanes, args.chain)
        batch = args.batch_per_gpu or 2048   # 2048 x 50 MB of value slots = 103 GB of the 180 GB
    elif args.workload == "sha256compression":
        d.set_main(C.sha256_compression(d), "sha256compression")
        label = "Sha256compression, BN254"
        batch = args.batch_per_gpu or 1024
    else:
        d.set_main(C.poseidon(d, 2), "poseidon2")
        label = "Poseidon(2), BN254"
        batch = args.batch_per_gpu or 65536
    return d, label, batch


def synth_inputs(desc, workload: str, batch: int, seed: int) -> np.ndarray:
    """uint64 [batch][n_inputs][4] canonical synthetic inputs."""
    rng = np.random.default_rng(seed)
    n_in = desc.main.n_in
    a = np.zeros((batch, n_in, 4), dtype=np.uint64)
    if workload == "ecdsa_scale":      # 64-bit limbs
        a[:, :, 0] = rng.integers(0, 2**64, size=(batch, n_in), dtype=np.uint64)
    elif workload in ("sha256compression", "sha256_512_bls"):  # bits
        a[:, :, 0] = rng.integers(0, 2, size=(batch, n_in), dtype=np.uint64)
    else:                                # field elements (top limb kept below q's)
        a[:, :, :] = rng.integers(0, 2**64, size=(batch, n_in, 4), dtype=np.uint64)
        a[:, :, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    return a


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        import tempfile
        self.path = tempfile.mktemp(prefix="cwclk", suffix=".csv")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50", "-f", self.path],
                                         stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            time.sleep(0.5)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        try:
            self.lines = [ln.strip() for ln in open(self.path)]
            os.remove(self.path)
        except OSError:
            self.lines = []
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(workload_name: str, batch: int, kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu capture, if it was taken on this configuration"""
    for name in ("r01c_traffic.json", "r01b_traffic.json"):   # newest capture first
        try:
            j = json.load(open(os.path.join(ROOT, "profiles", name)))
            if j["workload"] == workload_name and j["batch_per_gpu"] == batch:
                return int(j[kernel]["dram_bytes_read"] + j[kernel]["dram_bytes_write"])
        except Exception:
            pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
def cpu_reference_run(desc, args, inputs: np.ndarray, seconds: float):
    """The reference's own CPU path on this box's host cores, on a bounded sample of the workload.
    kind "reference": the reference runtime (common/main.cpp + calcwit.cpp + generic fr.cpp, built by
    oracle/build_ref.py) linked with the hand-lowered <circuit>.cpp, one process per input as the
    reference works (`<bin> input.json out.wtns`), `cores` processes at a time.
    kind "port": the C restatement oracle/cw_oracle.c on `cores` threads."""
    from oracle import c_oracle
    cores = os.cpu_count() or 1
    calc = os.path.join(ROOT, "oracle", "_ref", "calc", desc.name)
    if os.path.exists(calc) and os.path.exists(calc + ".dat"):
        import tempfile
        from concurrent.futures import ThreadPoolExecutor
        names = desc.main_inputs()
        td = tempfile.mkdtemp(prefix="cwref")

        def write_json(i):
            row = inputs[i % inputs.shape[0]]
            obj, k = {}, 0
            for name, _gid, n in names:
                vals = [str(int.from_bytes(row[k + j].tobytes(), "little")) for j in range(n)]
                obj[name] = vals if n > 1 else vals[0]
                k += n
            p = os.path.join(td, "in%d.json" % i)
            json.dump(obj, open(p, "w"))
            return p

        def one(i):
            subprocess.run([calc, os.path.join(td, "in%d.json" % i), os.path.join(td, "o%d.wtns" % i)], check=True,
                           stdout=subprocess.DEVNULL)
            try:
                os.remove(os.path.join(td, "o%d.wtns" % i))
            except OSError:
                pass
        write_json(0)
        t0 = time.time()
        one(0)
        t1 = time.time() - t0
        # the calculator scales poorly on many-core hosts (every process allocates and writes its own
        # multi-MB signal array and .wtns); try several degrees of parallelism and keep the best
        levels = sorted({max(1, cores // 8), max(1, cores // 4), max(1, cores // 2), cores})
        budget = max(2.0, seconds / len(levels))
        best = None
        made = 1
        for par in levels:
            n = int(max(par, min(par * 4, par * budget / max(t1, 1e-3))))
            for i in range(made, n):
                write_json(i)
            made = max(made, n)
            t0 = time.time()
            with ThreadPoolExecutor(par) as ex:
                list(ex.map(one, range(n)))
            dtl = time.time() - t0
            if best is None or n / dtl > best[0]:
                best = (n / dtl, par, n, dtl)
        rate, par_best, n, dt = best
        import shutil
        shutil.rmtree(td, ignore_errors=True)
        # the same work without process start / JSON / file output: C restatement on all cores
        orc = c_oracle.COracle(desc.to_bytes())
        port = 0.0
        for par in levels:
            n2 = par * 2
            t0 = time.time()
            orc.run_many(inputs[np.arange(n2) % inputs.shape[0]], par)
            port = max(port, n2 / (time.time() - t0))
        return {"value": rate, "unit": "witnesses/s", "cores": par_best, "host_cores": cores, "kind": "reference",
                "in_memory_port_witnesses_per_s": port,
                "sample": "%d inputs, one reference-calculator process per input (json in, .wtns out), %d at a time "
                          "(best of %s), --no_asm arithmetic, %.1f s; single process %.3f s/witness"
                          % (n, par_best, levels, dt, t1)}
    orc = c_oracle.COracle(desc.to_bytes())
    t0 = time.time()
    orc.run_many(inputs[:1], 1)
    t1 = time.time() - t0
    n = int(max(cores, min(cores * 8, cores * seconds / max(t1, 1e-3))))
    idx = np.arange(n) % inputs.shape[0]
    t0 = time.time()
    orc.run_many(inputs[idx], cores)
    dt = time.time() - t0
    return {"value": n / dt, "unit": "witnesses/s", "cores": cores, "kind": "port",
            "sample": "%d inputs through oracle/cw_oracle.c on %d threads, %.1f s; single thread %.3f s/witness"
                      % (n, cores, dt, t1)}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    desc, label, batch = make_workload(args)

    if args.impl == "reference":
        if rank != 0:
            return
        inputs = synth_inputs(desc, args.workload, 256, 1234)
        vals, walls = [], []
        for _ in range(max(1, args.steps)):   # a step = one bounded sample of the workload (no warm-up needed on the CPU)
            t0 = time.time()
            vals.append(cpu_reference_run(desc, args, inputs, max(4.0, args.cpu_seconds / max(1, args.steps))))
            walls.append(time.time() - t0)
        best = max(vals, key=lambda v: v["value"])
        v = float(np.mean([x["value"] for x in vals]))
        # (nothing of the CUDA back end is loaded in this arm: the size figure comes from the circuit description)
        out = {"impl": "reference", "metric": "witnesses/s", "value": v, "unit": "witnesses/s", "n_gpus": args.gpus,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(walls)),
               "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "u256 (4x u64 limbs, GMP mpn)", "data": "synthetic",
               "config": {"workload": label, "n_signals": desc.total_signals},
               "cpu_baseline": dict(best, value=v),
               "e2e": {"value": v, "unit": "witnesses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(out))
        return

    import torch
    import torch.distributed as dist
    from circom_b200.witness_calculator import Circuit, Batch, R1cs
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # one-time collective: rank 0's circuit description is broadcast over NCCL, every rank lowers it
    from circom_b200.distributed import broadcast_blob
    blob = broadcast_blob(desc.to_bytes() if rank == 0 else None, rank, world, device="cuda")
    circuit = Circuit(blob)
    st = circuit.stats
    b = Batch(circuit, batch, local_rank)
    n_in, W = circuit.n_inputs, circuit.n_witness
    inputs = synth_inputs(desc, args.workload, batch, 1000 + rank)
    pin_in = torch.empty((batch, n_in, 4), dtype=torch.int64, pin_memory=True)
    pin_in.numpy().view(np.uint64)[:] = inputs
    dev_in = pin_in.cuda()
    torch.cuda.synchronize()
    # end-to-end leg: its own batch, capped so that the pinned host buffer for the witnesses stays bounded
    # (e2e is PCIe-bound and independent of the batch size; 8 ranks x 38.8 GB of pinned memory is not)
    e2e_batch = batch
    pinned_cap = min(args.e2e_pinned_gb, 48.0 / world) * 1e9   # all ranks pin memory of the same host
    while e2e_batch > 64 and e2e_batch * W * 32 > pinned_cap:
        e2e_batch //= 2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        b.set_inputs(None, device_ptr=dev_in.data_ptr())
        b.run(sync=False)

    e2e_state = {}

    def step_e2e():
        if not e2e_state:
            e2e_state["b"] = b if e2e_batch == batch else Batch(circuit, e2e_batch, local_rank)
            e2e_state["out"] = torch.empty((e2e_batch, W, 4), dtype=torch.int64, pin_memory=True)
        eb = e2e_state["b"]
        eb.set_inputs(pin_in.numpy().view(np.uint64)[:e2e_batch])
        eb.run(sync=False)
        eb.witness(out=e2e_state["out"].numpy().view(np.uint64))

    # ---- device-resident timing ------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()            # nvidia-smi needs a moment to attach: it samples warm-up + timed steps (all under load)
    for _ in range(args.warmup):
        step_resident()
    b.sync()
    barrier()
    t0 = time.perf_counter()
    exec_ms = gather_ms = 0.0
    for _ in range(args.steps):
        step_resident()
        b.sync()
        e, g = b.last_ms()
        exec_ms += e
        gather_ms += g
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_ms = exec_ms + gather_ms  # CUDA events on the batch stream: stage+exec, gather
    t = torch.tensor([dev_ms, wall * 1e3, exec_ms, gather_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, exec_ms, gather_ms = [float(x) for x in t.tolist()]
    status = b.status()
    # CW_BENCH_NOCHECK: only for the diagnostic kernel builds of scripts/sweep_wrap.sh (garbage results)
    assert os.environ.get("CW_BENCH_NOCHECK") or not status.any(), "witness generation reported failing asserts: %r" % status[:8]

    # ---- end to end through the API with host buffers --------------------------------------------
    e2e_steps = min(args.steps, 3) if args.e2e_steps < 0 else args.e2e_steps
    e2e_s = None
    if e2e_steps > 0:
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        barrier()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    # ---- R1CS check on the device-resident witnesses ---------------------------------------------
    r1cs_ms = None
    if not args.no_r1cs:
        r = R1cs(circuit)
        fb, _ = r.check_batch(b, device=local_rank)
        assert (fb == -1).all(), "R1CS check failed on generated witnesses"
        ms = []
        for _ in range(max(2, args.steps)):
            fb, m = r.check_batch(b, device=local_rank)
            ms.append(m)
        t = torch.tensor([float(np.mean(ms))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        r1cs_ms = float(t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_batch = batch * world
    wit_s = total_batch * args.steps / (dev_ms / 1e3)
    peak, peak_src = measured_peaks()
    # algorithmic bytes per instance: every written slot once + witness + inputs
    s_w = st["n_slots"] - 1 - n_in
    b_wit = 32 * s_w + 32 * n_in   # witness entries are slots: written once by the tape, no gather pass
    exec_per_launch_ms = exec_ms / args.steps
    achieved = batch * (32 * s_w + 32 * n_in) / (exec_per_launch_ms / 1e3) / 1e9
    out = {
        "metric": "witnesses/s", "value": wit_s, "unit": "witnesses/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u256 (8x u32 limbs, Montgomery)", "data": "synthetic",
        "config": {"workload": label, "batch_per_gpu": batch, "global_batch": total_batch,
                   "n_constraints": st["n_constraints"], "n_signals": st["n_signals"], "n_tape_ops": st["n_tape_ops"],
                   "n_levels": st["n_levels"], "parallelism": "batch-sharded x%d" % world,
                   "l2": "working set %.1f GB per step >> L2, rewritten every step" % (batch * st["n_slots"] * 32 / 1e9)},
        "wall_ms_per_step": wall_ms / args.steps,
        "kernel_ms": {"tape_exec+stage": exec_ms / args.steps},
        "e2e": {"value": (e2e_batch * world * e2e_steps / e2e_s) if e2e_s else None, "unit": "witnesses/s",
                "steps": e2e_steps, "batch_per_gpu": e2e_batch,
                "h2d_bytes_per_step": int(e2e_batch * n_in * 32),
                "d2h_bytes_per_step": (e2e_state["b"].last_d2h_bytes() if e2e_state else int(e2e_batch * W * 32)),
                "host_witness_bytes_per_step": int(e2e_batch * W * 32)},
        "gpu_launches": 2 * args.steps,   # stage_inputs_kernel + tape_exec_kernel per step
        "clocks": clocks,
        "roofline": {"kernel": "tape_exec_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic(desc.name, batch, "tape_exec_kernel"),
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_witness": b_wit,
                     "operand_traffic_upper_bound_GBps": batch * st["n_tape_ops"] * 96 / (exec_per_launch_ms / 1e3) / 1e9},
    }
    if r1cs_ms is not None:
        nnz, m = st["n_nnz"], st["n_constraints"]
        b_r1cs = nnz * 8 + 3 * (m + 1) * 8 + 32 * st["n_constants"] + batch * (32 * W + 8)
        out["r1cs"] = {"mconstraints_per_s": total_batch * m / (r1cs_ms / 1e3) / 1e6, "ms": r1cs_ms,
                       "roofline": {"kernel": "r1cs_check_kernel", "bound": "hbm",
                                    "achieved": b_r1cs / (r1cs_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                    "frac": b_r1cs / (r1cs_ms / 1e3) / 1e9 / peak, "traffic": None}}
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_reference_run(desc, args, inputs, args.cpu_seconds)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
