"""Golden fixtures: `.wtns` files written by the REFERENCE witness calculators (tests/golden/make_golden.py ran the
binaries built from the reference sources) for fixed inputs, incl. the field's edge values.  They pin
  * the C oracle (oracle/cw_oracle.c) and the CPU build of the device code (tests/hostsim)   [-m "not gpu"]
  * the CUDA path through the public API (`calculateWTNSBin`)                                  [-m gpu]
byte for byte, on machines where neither /root/reference nor oracle/_ref exists."""
from __future__ import annotations

import glob
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

from circom_b200.witness_calculator import parse_value
from oracle import build_calcs, c_oracle
from tests.test_oracle_c import wtns_frame
from tests.util import flat_inputs, hostsim_run

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLDEN, "*.json")))


def load(name):
    meta = json.load(open(os.path.join(GOLDEN, name + ".json")))
    raws = []
    for i, sha in enumerate(meta["sha256"]):
        raw = zlib.decompress(open(os.path.join(GOLDEN, "%s_%d.wtns.z" % (name, i)), "rb").read())
        assert hashlib.sha256(raw).hexdigest() == sha
        raws.append(raw)
    return meta, raws


def int_inputs(d, inputs):
    out = []
    for inp in inputs:
        out.append({k: ([parse_value(x, d.q) for x in v] if isinstance(v, list) else parse_value(v, d.q))
                    for k, v in inp.items()})
    return out


def test_fixtures_present():
    assert len(NAMES) >= 9


def test_sha256_fixtures_are_sha256():
    """the reference calculator's outputs for the SHA-256 compression circuit are the digests hashlib computes"""
    import hashlib
    meta, raws = load("sha256compression")
    for inp, raw in zip(meta["inputs"], raws):
        block = int("".join(inp["inp"]), 2).to_bytes(64, "big")
        n = int.from_bytes(block[56:], "big") // 8
        body = raw[76:]
        bits = "".join(str(int.from_bytes(body[32 * (1 + i):32 * (2 + i)], "little")) for i in range(256))
        assert int(bits, 2).to_bytes(32, "big") == hashlib.sha256(block[:n]).digest()


@pytest.mark.parametrize("name", NAMES)
def test_oracle_and_host_build_reproduce_reference_wtns(name):
    meta, raws = load(name)
    d = build_calcs.make_desc(name)
    assert d.prime == meta["prime"]
    ins = int_inputs(d, meta["inputs"])
    arr = flat_inputs(d, ins)
    wit, st = c_oracle.COracle(d.to_bytes()).run(arr)
    assert not st.any()
    for i, raw in enumerate(raws):
        assert wtns_frame(d.q, wit[i]) == raw, (name, i, "C oracle")
    for flags in (4, 4 | 48):   # CW_FLAG_O0: every signal is a witness entry, as in the .dat; + the compact value store
        hw, hst, _, w2s = hostsim_run(d, ins, flags=flags)
        assert not hst.any() and w2s.tolist() == list(range(d.total_signals))
        for i, raw in enumerate(raws):
            assert wtns_frame(d.q, hw[i]) == raw, (name, i, "device code built for the CPU", flags)


# every fixture, the SHA-256 circuits and the 1.2 M-constraint one included; the goldilocks fixtures stay with the oracle and the
# CPU build of the device code (goldilocks circuits run on the GPU in tests/test_gpu_parity.py::test_other_primes_run_circuits)
GPU_NAMES = [n for n in NAMES if not n.endswith("_gl")]


@pytest.mark.gpu
@pytest.mark.parametrize("name", GPU_NAMES)
def test_gpu_wtns_equals_reference_fixture(name):
    from circom_b200.witness_calculator import Circuit, WitnessCalculator
    meta, raws = load(name)
    d = build_calcs.make_desc(name)
    wc = WitnessCalculator(Circuit(d, o0=True))
    for inp, raw in zip(meta["inputs"], raws):
        assert bytes(wc.calculateWTNSBin(inp)) == raw
    # and as one batch
    w = wc.calculate_witness_batch(meta["inputs"])
    for i, raw in enumerate(raws):
        assert wtns_frame(d.q, np.ascontiguousarray(w[i])) == raw
