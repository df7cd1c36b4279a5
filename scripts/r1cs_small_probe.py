"""A/B of the integer-row pass of the R1CS check (csrc/r1cs_small.h): per config the check of one batch on its value store
with CW_R1CS_SMALL=0 / 1, medians of CUDA-event times, results compared.  Prints one JSON line per config.
usage: python scripts/r1cs_small_probe.py [out.jsonl]"""
import json
import os
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from circom_b200.circuit import CircuitDesc  # noqa: E402
from circom_b200 import circuits as C  # noqa: E402
from circom_b200.witness_calculator import Circuit, Batch, R1cs  # noqa: E402
from tests.util import flat_inputs  # noqa: E402

out = open(sys.argv[1], "a") if len(sys.argv) > 1 else None
CONFIGS = [
    ("C2 Sha256compression bn128", "bn128", lambda d: C.sha256_compression(d), 1024, None),
    ("C4 Sha256(512) bls12381", "bls12381", lambda d: C.sha256(d, 512), 1024, None),
    ("C4 Sha256(512) bls12381, 32 per tile", "bls12381", lambda d: C.sha256(d, 512), 1024, "5"),
    ("Poseidon(2) bn128", "bn128", lambda d: C.poseidon(d, 2), 4096, None),
]
only = os.environ.get("PROBE_ONLY")
for name, prime, mk, batch, bt in CONFIGS:
    if only and only not in name:
        continue
    if bt is None:
        os.environ.pop("CW_BT_LOG2", None)
    else:
        os.environ["CW_BT_LOG2"] = bt
    d = CircuitDesc(prime)
    d.set_main(mk(d))
    rng = np.random.default_rng(1)
    ins = [{n: ([int(x) for x in rng.integers(0, 2, sz)] if sz > 1 else int(rng.integers(0, 2))) for n, _g, sz in d.main_inputs()}
           for _ in range(batch)]
    t0 = time.time()
    c = Circuit(d)
    b = Batch(c, batch)
    b.set_inputs(flat_inputs(d, ins))
    b.run()
    res = {"config": name, "batch": batch, "bt_log2": bt, "constraints": c.stats["n_constraints"], "setup_s": round(time.time() - t0, 2)}
    fbs = {}
    for small in ("0", "1"):
        os.environ["CW_R1CS_SMALL"] = small
        r = R1cs(c)
        info = r.compiled_info(b)
        ms = []
        for it in range(8):
            fb, t = r.check_batch(b)
            ms.append(t)
        fbs[small] = fb
        med = statistics.median(ms[2:])
        res["small_" + small] = {"ms": round(med, 4), "mconstraints_per_s": round(batch * c.stats["n_constraints"] / med / 1e3, 1),
                                 "integer_rows": info["integer_rows"], "general_rows": info["general_rows"]}
    res["same_result"] = bool((fbs["0"] == fbs["1"]).all())
    res["all_satisfied"] = bool((fbs["1"] == -1).all())
    res["speedup"] = round(res["small_0"]["ms"] / res["small_1"]["ms"], 3)
    line = json.dumps(res)
    print(line, flush=True)
    if out:
        out.write(line + "\n")
        out.flush()
