"""Conformance of the .cb2c format specification (docs/CB2C.md): a file written by a stand-alone C program that follows
only the specification must equal, byte for byte, what the Python DSL writes for the same circuit; it must load, lower
and compute the expected witness - in the library's lowering (run by the CPU build of the device code) and in the
oracle, which parses the format independently."""
import os
import subprocess

import numpy as np

from circom_b200.circuit import CircuitDesc
from circom_b200 import circuits as C
from circom_b200.witness_calculator import Circuit
from tests.util import ROOT, hostsim, limbs_to_ints


def dsl_conf():
    d = CircuitDesc("bn128")
    m2 = C.multiplier2(d)

    def build(t):
        x, y = t.input("x"), t.input("y")
        out = t.output("bits", 2)
        p = t.output("p")
        lc = t.const(0)
        for k in range(2):
            t.assign(out[k], (x >> k) & 1)
            t.constrain(out[k] * (out[k] - 1), 0)
            lc = lc + out[k] * (1 << k)
        t.constrain(lc, x)
        c = t.component("m", m2)
        t.assign_constrained(c["a"], x)
        t.assign_constrained(c["b"], y)
        t.assign_constrained(p, c["c"] + 1)
    d.set_main(d.template("Conf", (), build), "conf")
    return d


def test_c_writer_from_the_spec_matches_the_dsl_and_runs(tmp_path):
    import ctypes
    exe, out = str(tmp_path / "cb2c_conf"), str(tmp_path / "conf.cb2c")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cb2c_writer", "cb2c_conf.c")])
    subprocess.check_call([exe, out])
    blob = open(out, "rb").read()
    d = dsl_conf()
    assert blob == d.to_bytes()
    # loads and lowers like the DSL's description (same object code, same witness list)
    c = Circuit(blob, host_only=True)
    assert c.stats == Circuit(d, host_only=True).stats and c.n_witness == 7   # 9 signals; m.a = x and m.b = y merged away
    # runs: the library's lowering on the CPU build of the device code, and the oracle's own parser + evaluator
    from oracle.c_oracle import COracle
    hs = hostsim()
    ins = np.zeros((3, 2, 4), dtype=np.uint64)
    ins[:, 0, 0] = [3, 2, 1]
    ins[:, 1, 0] = [11, 5, 2**63]
    W = c.n_witness
    wit = np.zeros((3, W, 4), dtype=np.uint64)
    st = np.zeros(3, dtype=np.int32)
    rc = hs.hs_run(blob, ctypes.c_size_t(len(blob)), 0, ins.ctypes.data_as(ctypes.c_void_p), 3,
                   wit.ctypes.data_as(ctypes.c_void_p), st.ctypes.data_as(ctypes.c_void_p), None)
    assert rc == 0 and not st.any()
    w2s = c.witness2signal().astype(np.int64)
    ow, ost = COracle(blob).run(ins)
    assert not ost.any() and (ow[:, w2s] == wit).all()
    for i, (x, y) in enumerate(((3, 11), (2, 5), (1, 2**63))):
        assert limbs_to_ints(wit[i]) == [1, x & 1, (x >> 1) & 1, x * y + 1, x, y, x * y]
