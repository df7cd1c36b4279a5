"""GPU parity tests: the sm_100a kernels, called through the C ABI, against the oracle.
Bit-exact (integer arithmetic)."""
import ctypes
import random
import zlib

import numpy as np
import pytest

from circom_b200 import native
from circom_b200.circuit import CircuitDesc, OPS
from circom_b200 import circuits as C
from circom_b200.witness_calculator import Circuit, Batch, R1cs, WitnessCalculator, builder
from oracle.field_model import Field, OP_NAMES
from oracle.ir_eval import evaluate, check_r1cs
from tests.util import ints_to_limbs, limbs_to_ints, edge_values, rand_operand, flat_inputs
from tests.test_lowering_cpu import CIRCUITS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prime", [0, 1])
def test_device_field_ops(prime):
    """device Fr_* equivalents (fr.hpp:28-70) over all operators, random + edge operands"""
    F = Field(["bn128", "bls12381"][prime])
    q = F.q
    rng = random.Random(991 + prime)
    edges = edge_values(q)
    n = 20000
    A = [rand_operand(rng, q, edges) for _ in range(n)]
    B = [rand_operand(rng, q, edges) if rng.random() > 0.25 else rng.randrange(300) for _ in range(n)]
    Cc = [rng.choice([0, 1, rng.randrange(q)]) for _ in range(n)]
    a, b, c = ints_to_limbs(A), ints_to_limbs(B), ints_to_limbs(Cc)
    r = np.zeros((n, 4), dtype=np.uint64)
    for op in list(range(1, 24)) + [OPS["SELECT"], 28]:
        m = n if op not in (OPS["POW"], OPS["DIV"], 28) else 2000
        bb = b
        if op in (OPS["IDIV"], OPS["MOD"]):
            Bz = [x if x else 1 for x in B]
            bb = ints_to_limbs(Bz)
        else:
            Bz = B
        native.check(native.lib.cw_fr_batch_op(prime, op, a.ctypes.data, bb.ctypes.data, c.ctypes.data,
                                               r.ctypes.data, m, 0))
        got = limbs_to_ints(r[:m])
        for i in range(m):
            exp = F.inv(A[i]) if op == 28 else F.apply(op, A[i], Bz[i], Cc[i])
            assert got[i] == exp, (OP_NAMES.get(op, op), hex(A[i]), hex(Bz[i]), hex(got[i]), hex(exp))


@pytest.mark.parametrize("prime", ["bn128", "bls12381"])
@pytest.mark.parametrize("name", sorted(CIRCUITS))
@pytest.mark.parametrize("batch", [1, 37])
def test_circuit_witness_matches_oracle(prime, name, batch):
    mk, gen = CIRCUITS[name]
    d = CircuitDesc(prime)
    d.set_main(mk(d))
    rng = random.Random(zlib.crc32((prime + name).encode()) + batch)
    ins = [gen(rng, d.q) for _ in range(batch)]
    wc = builder(d)
    wit = wc.calculate_witness_batch(ins)
    w2s = wc.circuit.witness2signal().astype(np.int64)
    for i, inp in enumerate(ins):
        exp = evaluate(d, inp)
        assert limbs_to_ints(wit[i]) == [exp[k] for k in w2s], (prime, name, i)
    # algebraic self-check on the GPU: A.w o B.w == C.w
    fb, _ = R1cs(wc.circuit).check(wit)
    assert (fb == -1).all()


@pytest.mark.parametrize("bt", ["0", "2", "5"])
def test_tile_layouts_agree(bt, monkeypatch):
    """every instance-tile width of the slot layout gives the same witnesses"""
    monkeypatch.setenv("CW_BT_LOG2", bt)
    d = CircuitDesc("bn128")
    d.set_main(C.all_ops(d))
    rng = random.Random(5)
    ins = [CIRCUITS["all_ops"][1](rng, d.q) for _ in range(45)]
    wc = builder(d)
    wit = wc.calculate_witness_batch(ins)
    w2s = wc.circuit.witness2signal().astype(np.int64)
    for i, inp in enumerate(ins):
        exp = evaluate(d, inp)
        assert limbs_to_ints(wit[i]) == [exp[k] for k in w2s]


def test_reference_surface_single_input():
    """calculateWitness / calculateBinWitness / calculateWTNSBin (witness_calculator.js:176-276) and the
    docs' worked example (computing-the-witness.md:16-24)."""
    d = CircuitDesc("bn128")
    d.set_main(C.multiplier2(d))
    wc = builder(d)
    assert wc.calculateWitness({"a": "3", "b": "11"}) == [1, 33, 3, 11]
    assert wc.calculateBinWitness({"a": 3, "b": 11}) == b"".join(int(x).to_bytes(32, "little") for x in (1, 33, 3, 11))
    wtns = wc.calculateWTNSBin({"a": "0x3", "b": "0b1011"})
    q = d.q
    exp = (b"wtns" + (2).to_bytes(4, "little") + (2).to_bytes(4, "little") + (1).to_bytes(4, "little") +
           (40).to_bytes(8, "little") + (32).to_bytes(4, "little") + q.to_bytes(32, "little") +
           (4).to_bytes(4, "little") + (2).to_bytes(4, "little") + (128).to_bytes(8, "little") +
           b"".join(int(x).to_bytes(32, "little") for x in (1, 33, 3, 11)))
    assert wtns == exp and len(wtns) == 204


def test_input_errors_follow_reference():
    d = CircuitDesc("bn128")
    d.set_main(C.multiplier_n(d, 4))
    wc = builder(d)
    with pytest.raises(ValueError, match="Not enough values"):
        wc.calculateWitness({"in": [1, 2, 3]})
    with pytest.raises(ValueError, match="Too many values"):
        wc.calculateWitness({"in": [1, 2, 3, 4, 5]})
    with pytest.raises(ValueError, match="not found"):
        wc.calculateWitness({"in": [1, 2, 3, 4], "zz": 1})
    with pytest.raises(ValueError, match="Not all inputs"):
        wc.calculateWitness({})
    b = Batch(wc.circuit, 2)
    b.set_input(0, "in", 0, 5)
    with pytest.raises(native.CwError, match="assigned twice"):
        b.set_input(0, "in", 0, 5)
    with pytest.raises(native.CwError, match="Not all inputs"):
        b.run()


def test_assert_and_r1cs_violation_detected():
    d = CircuitDesc("bn128")

    def build(t):
        a = t.input("a")
        b = t.input("b")
        o = t.output("o")
        t.assign(o, a + b)
        t.constrain(a * b, o)
    d.set_main(d.template("Bad", (), build))
    c = Circuit(d)
    b = Batch(c, 3)
    b.set_inputs(flat_inputs(d, [{"a": 2, "b": 2}, {"a": 2, "b": 3}, {"a": 0, "b": 0}]))
    b.run()
    assert b.status().tolist() == [0, 1, 0]
    fb, _ = R1cs(c).check(b.witness())
    assert fb.tolist() == [-1, 0, -1]
    # the same check straight from device memory
    fb2, _ = R1cs(c).check(None, batch=3, device_ptr=b.witness_device_ptr())
    assert fb2.tolist() == [-1, 0, -1]


def test_r1cs_check_reads_witness_in_place():
    """the tape writes witness entries into the first slots of each instance: the R1CS check and the
    host copy read them there (strided), the dense device copy is made only on request"""
    d = CircuitDesc("bn128")
    d.set_main(C.less_than(d, 16))
    c = Circuit(d)
    rng = random.Random(8)
    ins = [{"in": [rng.randrange(65536), rng.randrange(65536)]} for _ in range(50)]
    b = Batch(c, len(ins))
    b.set_inputs(flat_inputs(d, ins))
    b.run()
    r = R1cs(c)
    fb1, _ = r.check_batch(b)
    ptr, stride = b.witness_strided()
    assert stride == c.stats["n_slots"] and stride >= c.n_witness
    wit = b.witness()
    fb2, _ = r.check(wit)
    fb3, _ = r.check(None, batch=len(ins), device_ptr=b.witness_device_ptr())
    assert (fb1 == -1).all() and (fb2 == -1).all() and (fb3 == -1).all()
    w2s = c.witness2signal().astype(np.int64)
    for i, inp in enumerate(ins):
        exp = evaluate(d, inp)
        assert limbs_to_ints(wit[i]) == [exp[k] for k in w2s]


def test_r1cs_first_violated_row_matches_oracle():
    """corrupt single witness entries (bits set to 2, the packed value changed): the smallest violated row
    reported by the GPU check equals the oracle's, also for boolean rows that are checked inside the
    recomposition-sum thread"""
    from oracle.c_oracle import COracle
    d = CircuitDesc("bn128")
    d.set_main(C.num2bits(d, 16))     # no signal=signal rows: row numbering equals the oracle's
    c = Circuit(d)
    b = Batch(c, 1)
    b.set_inputs(flat_inputs(d, [{"in": 0xBEEF}]))
    b.run()
    good = b.witness()
    orc = COracle(d.to_bytes())
    r = R1cs(c)
    cases = []
    for wire in (1, 5, 16, 17):       # out[0], out[4], out[15], in
        bad = good.copy()
        bad[0, wire, 0] = 2 if wire != 17 else 0xBEEE
        cases.append(bad)
    batch = np.concatenate(cases, axis=0)
    fb, _ = r.check(batch)
    exp = orc.r1cs_check(batch)
    assert (exp >= 0).all() and fb.tolist() == exp.tolist()


def test_r1cs_lane_group_kernel_matches(monkeypatch):
    """the opt-in kernel that checks long rows with 8 lanes per (row, instance) (CW_R1CS_SPLIT=1) reports the
    same first violated row as the default kernel and the oracle"""
    from oracle.c_oracle import COracle
    d = CircuitDesc("bn128")
    d.set_main(C.num2bits(d, 64))     # one 65-term recomposition row, 64 boolean rows
    c = Circuit(d)
    b = Batch(c, 1)
    b.set_inputs(flat_inputs(d, [{"in": 0xDEADBEEFCAFEF00D}]))
    b.run()
    good = b.witness()
    cases = [good.copy()]
    for wire in (1, 9, 40, 64, 65):   # out[0], out[8], out[39], out[63], in
        bad = good.copy()
        bad[0, wire, 0] = 2 if wire != 65 else 0xDEADBEEFCAFEF00C
        cases.append(bad)
    batch = np.concatenate(cases, axis=0)
    exp = COracle(d.to_bytes()).r1cs_check(batch)
    r = R1cs(c)
    monkeypatch.setenv("CW_R1CS_SPLIT", "0")
    fb0, _ = r.check(batch)
    monkeypatch.setenv("CW_R1CS_SPLIT", "1")
    fb1, _ = r.check(batch)
    assert exp[0] == -1 and (exp[1:] >= 0).all()
    assert fb0.tolist() == exp.tolist() and fb1.tolist() == exp.tolist()


def test_packed_device_to_host_transfer_equals_plain_copy(monkeypatch):
    """witness entries proven to be bits / 64-bit values cross PCIe packed and are zero-extended on the
    host: the host array must equal the plain pitched copy, with fewer bytes transferred"""
    d = CircuitDesc("bn128")
    d.set_main(C.ecdsa_scale(d, 2, 3))
    c = Circuit(d)
    rng = random.Random(2)
    ins = [{"a": [rng.randrange(2**64) for _ in range(8)], "b": [rng.randrange(2**64) for _ in range(8)]} for _ in range(37)]
    b = Batch(c, len(ins))
    b.set_inputs(flat_inputs(d, ins))
    b.run()
    monkeypatch.setenv("CW_PACKED_D2H", "0")
    plain = b.witness().copy()
    plain_bytes = b.last_d2h_bytes()
    monkeypatch.setenv("CW_PACKED_D2H", "1")
    packed = b.witness().copy()
    packed_bytes = b.last_d2h_bytes()
    assert (plain == packed).all()
    assert plain_bytes == len(ins) * c.n_witness * 32 and packed_bytes * 4 < plain_bytes
    w2s = c.witness2signal().astype(np.int64)
    exp = evaluate(d, ins[5])
    assert limbs_to_ints(packed[5]) == [exp[k] for k in w2s]
