"""CPU checks of the lowering (flatten.cpp) and of the host+device field source (fr_device.cuh):
the tape is executed by tests/hostsim (which compiles the same source the kernels use) and
compared bit-for-bit with the oracle.  No GPU needed; the GPU parity tests are in test_gpu_*.py."""
import ctypes
import random

import numpy as np
import pytest

from circom_b200.circuit import CircuitDesc, OPS
from circom_b200 import circuits as C
from oracle.field_model import Field, OP_NAMES, DivisionByZero
from oracle.ir_eval import evaluate, check_r1cs
from tests.util import hostsim, hostsim_run, hostsim_run_r1cs, ints_to_limbs, limbs_to_ints, edge_values, rand_operand, PRIME_NAMES

CIRCUITS = {
    "multiplier2": (lambda d: C.multiplier2(d), lambda r, q: {"a": r.randrange(q), "b": r.randrange(q)}),
    "all_ops": (lambda d: C.all_ops(d),
                lambda r, q: {"a": r.choice([0, 1, q - 1, r.randrange(q), r.randrange(2**64)]),
                              "b": r.choice([0, 1, 5, 255, q - 3, r.randrange(q), r.randrange(300)])}),
    "less_than8": (lambda d: C.less_than(d, 8), lambda r, q: {"in": [r.randrange(256), r.randrange(256)]}),
    "num2bits64": (lambda d: C.num2bits(d, 64), lambda r, q: {"in": r.randrange(2**64)}),
    "multiplier_n6": (lambda d: C.multiplier_n(d, 6), lambda r, q: {"in": [r.randrange(q) for _ in range(6)]}),
    "is_zero": (lambda d: C.is_zero(d), lambda r, q: {"in": r.choice([0, r.randrange(q)])}),
    # hints computed by circom functions with run-time loops, branches and array indexing
    "int_div32": (lambda d: C.int_div(d, 32),
                  lambda r, q: {"a": r.choice([0, 1, 2**32 - 1, r.randrange(2**32)]),
                                "b": r.choice([1, 2**32 - 1, r.randrange(1, 2**r.randrange(1, 33))])}),
    # one call, several results (`var qr[3] = f(a, b);`): all used / first result dead / only the first used
    "int_div_arr_all": (lambda d: C.int_div_array(d, 32, "all"),
                        lambda r, q: {"a": r.choice([0, 1, 2**32 - 1, r.randrange(2**32)]),
                                      "b": r.choice([1, 2**32 - 1, r.randrange(1, 2**r.randrange(1, 33))])}),
    "int_div_arr_tail": (lambda d: C.int_div_array(d, 32, "tail"),
                         lambda r, q: {"a": r.randrange(2**32), "b": r.choice([0, 1, r.randrange(1, 2**32)])}),
    "int_div_arr_head": (lambda d: C.int_div_array(d, 32, "head"),
                         lambda r, q: {"a": r.randrange(2**32), "b": r.choice([0, 1, r.randrange(1, 2**32)])}),
    # a function that calls two other functions inside a data-dependent loop (nested frames in the interpreter)
    "gcd32": (lambda d: C.gcd_circuit(d, 32),
              lambda r, q: {"a": r.choice([0, 1, 2**32 - 1, r.randrange(2**32), 2 * 3 * 5 * 7 * 11 * 13 * r.randrange(1, 1000)]),
                            "b": r.choice([0, 1, r.randrange(1, 2**32), 2 * 3 * 5 * 7 * r.randrange(1, 100000)])}),
    # a component array of mixed templates (io map in the description) and a load at a run-time address (expanded by the producer)
    "mixed_array": (lambda d: C.mixed_array(d), lambda r, q: {"a": [r.randrange(q) for _ in range(3)], "b": r.choice([0, 1, r.randrange(q)])}),
    "logging": (lambda d: C.logging(d), lambda r, q: {"a": r.choice([0, 1, q - 1, r.randrange(q)]), "b": r.randrange(q)}),   # log() calls: no tape ops
    "table_lookup8": (lambda d: C.table_lookup(d, 8), lambda r, q: {"table": [r.choice([0, 1, q - 1, r.randrange(q)]) for _ in range(8)],
                                                                     "sel": r.randrange(8)}),
    # the bench circuit's BigMultModP with its quotient / remainder hints computed by a long_div-style function
    "ecdsa_calls_1x2": (lambda d: C.ecdsa_scale(d, 1, 2, hints="functions"),
                        lambda r, q: {"a": [r.choice([2**64 - 1, r.getrandbits(64)]) for _ in range(4)],
                                      "b": [r.choice([2**64 - 1, 0, r.getrandbits(64)]) for _ in range(4)]}),
}


@pytest.mark.parametrize("prime", ["bn128", "bls12381", "grumpkin", "pallas", "vesta", "secq256r1", "bls12377", "goldilocks"])
@pytest.mark.parametrize("name", sorted(CIRCUITS))
def test_tape_matches_oracle(prime, name):
    if prime == "goldilocks" and name == "ecdsa_calls_1x2":
        pytest.skip("products of 64-bit limbs need a field above 2^130")
    mk, gen = CIRCUITS[name]
    d = CircuitDesc(prime)
    d.set_main(mk(d))
    import zlib
    rng = random.Random(zlib.crc32((prime + name).encode()))
    ins = [gen(rng, d.q) for _ in range(24)]
    for flags in ((0, 4, 16, 32, 48, 52, 64, 112) if prime in ("bn128", "bls12381") else (0, 48, 112)):  # default, O0, BITPLANE, REUSE, both (COMPACT), COMPACT + O0, FUSE, COMPACT + FUSE
        wit, st, stats, w2s = hostsim_run(d, ins, flags=flags)
        if flags == 4:
            assert w2s.tolist() == list(range(d.total_signals))
        assert w2s[0] == 0 and (np.diff(w2s) > 0).all()
        assert set(range(1, 1 + d.main.n_out + d.main.n_in)) <= set(w2s.tolist())
        for i, inp in enumerate(ins):
            exp = evaluate(d, inp)
            assert check_r1cs(d, exp) == 0
            assert limbs_to_ints(wit[i]) == [exp[k] for k in w2s], (prime, name, i)
        assert not st.any()


def test_wide_tapes_and_compact_layouts():
    import hashlib
    """Tapes with wide levels under the compact value store: bit runs in the bit plane (CW_FLAG_BITPLANE) and
    temporaries sharing slots (CW_FLAG_REUSE).  The simulator checks that no level reads what it writes (the device
    runs a level's work items in any order) and compares every value.  Known answers: python-int secp256k1 arithmetic."""
    from circom_b200.circuits.bigint import ecdsa_scale_expected
    from circom_b200.witness_calculator import Circuit
    rng = random.Random(77)
    d = CircuitDesc("bn128")
    d.set_main(C.ecdsa_scale(d, 2, 5))
    a = [rng.getrandbits(64) for _ in range(8)]
    b = [rng.getrandbits(64) for _ in range(8)]
    wit, st, stats, w2s = hostsim_run(d, [{"a": a, "b": b}])
    assert not st.any()
    assert limbs_to_ints(wit[0])[1:9] == ecdsa_scale_expected(a, b, 2, 5)
    # bit-plane layout (CW_FLAG_BITPLANE): the range-check bits leave the 32-byte slot store, same witness
    wit_bp, st_bp, stats_bp, w2s_bp = hostsim_run(d, [{"a": a, "b": b}], flags=16)
    assert not st_bp.any() and (wit_bp == wit).all() and (w2s_bp == w2s).all()
    assert int(stats_bp[3]) < int(stats[3]) // 2
    # + slot reuse: the value store shrinks by another large factor, same witness
    wit_c, st_c, stats_c, w2s_c = hostsim_run(d, [{"a": a, "b": b}], flags=48)
    assert not st_c.any() and (wit_c == wit).all() and (w2s_c == w2s).all()
    assert int(stats_c[3]) < int(stats_bp[3]) // 2 and int(stats_c[7]) > 0
    stc = Circuit(d, host_only=True, compact=False).stats
    assert stc["n_bitwords"] == 0 and stc["n_resident_slots"] == stc["n_witness"]
    stcc = Circuit(d, host_only=True).stats
    assert stcc["n_bitwords"] == int(stats_c[7]) and stcc["n_resident_slots"] < stcc["n_slots"] < stc["n_slots"] // 4
    # census of the slots by proven width: every slot is counted once, the range-checked bits dominate
    from circom_b200 import native
    cc = Circuit(d, host_only=True, compact=False)
    census = (ctypes.c_uint64 * 4)()
    assert native.lib.cw_circuit_slot_census(cc._h, census) == 0
    assert sum(census) == stc["n_slots"] and census[0] > stc["n_slots"] // 2

    d = CircuitDesc("bn128")
    d.set_main(C.sha256(d, 64))
    msg = bytes(rng.getrandbits(8) for _ in range(8))
    bits = [(byte >> (7 - k)) & 1 for byte in msg for k in range(8)]
    wit, st, stats, w2s = hostsim_run(d, [{"in": bits}])
    assert not st.any()
    out = limbs_to_ints(wit[0])[1:257]
    digest = hashlib.sha256(msg).digest()
    assert out == [(byte >> (7 - k)) & 1 for byte in digest for k in range(8)]


def test_assert_failure_is_reported():
    """`===` violated -> status k+1 of the first failing assert (reference: assert(Fr_isTrue(..)) aborts,
    assert_bucket.rs:70-88)."""
    d = CircuitDesc("bn128")

    def build(t):
        a = t.input("a")
        b = t.input("b")
        o = t.output("o")
        t.assign(o, a + b)
        t.constrain(a * b, o)       # only true for special inputs
        t.constrain(a, b)           # second assert
    d.set_main(d.template("Bad", (), build))
    wit, st, _, _ = hostsim_run(d, [{"a": 2, "b": 2}, {"a": 2, "b": 3}, {"a": 0, "b": 5}])
    assert st.tolist() == [0, 1, 1]
    wit, st, _, _ = hostsim_run(d, [{"a": 2, "b": 3}], flags=1)  # CW_FLAG_NO_ASSERTS
    assert st.tolist() == [0]


@pytest.mark.parametrize("prime", range(8))
def test_device_field_source_vs_model(prime):
    """every operator of fr_device.cuh (compiled for the host) against the python model, for all seven 256-bit primes
    (secq256r1 is a full 256-bit modulus: the ninth limb of the Montgomery product and of the division matter)"""
    F = Field(PRIME_NAMES[prime])
    q = F.q
    rng = random.Random(77 + prime)
    edges = edge_values(q)
    hs = hostsim()
    n = 3000
    A = [rand_operand(rng, q, edges) for _ in range(n)]
    B = [rand_operand(rng, q, edges) if rng.random() > 0.25 else rng.randrange(300) for _ in range(n)]
    Cc = [rng.choice([0, 1, rng.randrange(q)]) for _ in range(n)]
    a, b, c = ints_to_limbs(A), ints_to_limbs(B), ints_to_limbs(Cc)
    r = np.zeros((n, 4), dtype=np.uint64)
    for op in list(range(1, 24)) + [OPS["SELECT"], 28]:
        m = n if op not in (OPS["POW"], OPS["DIV"], 28) else 150
        err = hs.hs_fr_op(prime, op, a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                          c.ctypes.data_as(ctypes.c_void_p), r.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(m))
        got = limbs_to_ints(r[:m])
        for i in range(m):
            if op == 28:
                exp = F.inv(A[i])
            elif op in (OPS["IDIV"], OPS["MOD"]) and B[i] == 0:
                assert err == 1
                continue
            else:
                exp = F.apply(op, A[i], B[i], Cc[i])
            assert got[i] == exp, (OP_NAMES.get(op, op), hex(A[i]), hex(B[i]), hex(got[i]), hex(exp))


def test_r1cs_check_arithmetic_and_violation_detection():
    hs = hostsim()
    d = CircuitDesc("bn128")
    d.set_main(C.less_than(d, 16))
    blob = d.to_bytes()
    rng = random.Random(3)
    ins = [{"in": [rng.randrange(65536), rng.randrange(65536)]} for _ in range(6)]
    wit, st, stats, w2s = hostsim_run(d, ins)
    fb = np.zeros(len(ins), dtype=np.int64)
    assert hs.hs_r1cs_check(blob, ctypes.c_size_t(len(blob)), wit.ctypes.data_as(ctypes.c_void_p), len(ins),
                            fb.ctypes.data_as(ctypes.c_void_p)) == 0
    assert (fb == -1).all()
    bad = wit.copy()
    bad[2, 3, 0] ^= 1  # flip one bit of one wire of instance 2
    assert hs.hs_r1cs_check(blob, ctypes.c_size_t(len(blob)), bad.ctypes.data_as(ctypes.c_void_p), len(ins),
                            fb.ctypes.data_as(ctypes.c_void_p)) == 0
    assert fb[2] >= 0 and (np.delete(fb, 2) == -1).all()



def test_narrow_register_machine_ops():
    """fr_device.cuh vmn_apply (the 128-bit machine hint functions run on first): whenever it produces a result it is the
    field result of the python model; it gives up exactly when the result leaves 128 bits / the operator is not an
    integer one, and it must not give up on the common cases (that is its point)"""
    hs = hostsim()
    rng = random.Random(5)
    M = 2**128

    def small():
        return rng.choice([0, 1, 2, 2**64 - 1, 2**64, 2**127, M - 1, rng.getrandbits(rng.randrange(1, 129)),
                           rng.getrandbits(rng.randrange(1, 65)), rng.randrange(140)])
    for prime in ("bn128", "secq256r1"):
        F = Field(prime)
        for op in list(range(1, 26)) + [28]:
            n = 4000
            A = [small() for _ in range(n)]
            B = [small() for _ in range(n)]
            Cc = [rng.choice([0, 1, small()]) for _ in range(n)]
            v = np.array([[a & (2**64 - 1), a >> 64, b & (2**64 - 1), b >> 64, c & (2**64 - 1), c >> 64]
                          for a, b, c in zip(A, B, Cc)], dtype=np.uint64)
            out = np.zeros((n, 2), dtype=np.uint64)
            ok = np.zeros(n, dtype=np.uint8)
            hs.hs_vmn_apply(op, v.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p),
                            ok.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n))
            n_ok = 0
            for i in range(n):
                try:
                    exp = F.apply(op, A[i], B[i], Cc[i]) if op != 28 else F.inv(A[i])
                except DivisionByZero:
                    assert not ok[i]          # the full-width machine reports it
                    continue
                if ok[i]:
                    n_ok += 1
                    assert int(out[i, 0]) | (int(out[i, 1]) << 64) == exp, (OP_NAMES.get(op, op), hex(A[i]), hex(B[i]))
                else:
                    # giving up is only allowed when the narrow machine has no answer
                    fits = exp < M
                    name = OP_NAMES.get(op, str(op))
                    if name in ("ADD", "SUB", "LT", "GT", "LEQ", "GEQ", "EQ", "NEQ", "LOR", "LAND", "LNOT", "BOR", "BAND",
                                "BXOR", "COPY", "SELECT"):
                        assert not fits or name == "SUB" and A[i] < B[i], (name, hex(A[i]), hex(B[i]))
                    if name == "MUL" and A[i] < 2**64 and B[i] < 2**64:
                        raise AssertionError("64 x 64 product refused")
                    if name in ("SHR", "SHL") and B[i] < 128 and fits and name == "SHR":
                        raise AssertionError("shift refused")
            name = OP_NAMES.get(op, str(op))
            if name in ("DIV", "POW", "BNOT", "INV") or op == 28:
                assert n_ok == 0
            elif name not in ("NEG", "IDIV", "MOD"):
                assert n_ok > n // 4, name


def test_calls_run_on_the_narrow_machine_and_fall_back():
    """calls of limb-arithmetic functions finish on the 128-bit machine; a call whose values leave 128 bits (a negative
    difference, a field division, a wide argument) is repeated at full width - same witness either way"""
    hs = hostsim()
    na, wi = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    rng = random.Random(9)
    d = CircuitDesc("bn128")
    d.set_main(C.ecdsa_scale(d, 1, 2, hints="functions"))
    ins = [{"a": [rng.getrandbits(64) for _ in range(4)], "b": [rng.getrandbits(64) for _ in range(4)]} for _ in range(6)]
    hs.hs_vm_counters(ctypes.byref(na), ctypes.byref(wi))
    wit, st, _, w2s = hostsim_run(d, ins)
    hs.hs_vm_counters(ctypes.byref(na), ctypes.byref(wi))
    assert not st.any() and na.value == 6 * 2 and wi.value == 0
    for i, inp in enumerate(ins):
        exp = evaluate(d, inp)
        assert limbs_to_ints(wit[i]) == [exp[k] for k in w2s]

    # a function that leaves the integers depending on its arguments
    d = CircuitDesc("bn128")

    def body(f):
        a, b = f.param(0), f.param(1)
        x = f.var(a - b)                  # negative when a < b
        f.if_begin(b.gt(1000))
        f.set(x, x + a / b)               # field division
        f.if_end()
        f.ret(x * 3)
    fn = d.function("mix", 2, body)

    def build(t):
        a, b = t.input("a"), t.input("b")
        o = t.output("o")
        t.assign(o, t.call(fn, [a, b]))
    d.set_main(d.template("Mix", (), build))
    q = d.q
    ins = [{"a": 7, "b": 5}, {"a": 5, "b": 7}, {"a": 2**70, "b": 2**69}, {"a": 9, "b": 2000}, {"a": q - 1, "b": 1},
           {"a": 2**128, "b": 1}, {"a": 2**127, "b": 3}, {"a": 2**70, "b": 999}, {"a": 2**100, "b": 1}]
    # narrow: (7, 5), (2^70, 999), (2^100, 1); wide: a < b, the two field divisions (b > 1000), the two wide arguments,
    # (2^127 - 3) * 3 >= 2^128
    wit, st, _, w2s = hostsim_run(d, ins)
    hs.hs_vm_counters(ctypes.byref(na), ctypes.byref(wi))
    assert not st.any() and na.value == 3 and wi.value == 6
    for i, inp in enumerate(ins):
        exp = evaluate(d, inp)
        assert limbs_to_ints(wit[i]) == [exp[k] for k in w2s], i


@pytest.mark.parametrize("prime", range(8))
def test_modular_inverse_by_division_steps(prime):
    """fr_device.cuh fr_modinv / fr_inv_mont (safegcd, 600 division steps on 30-bit limbs) against python's pow(x, -1, q) for
    every prime: edge values (0 -> 0 as the reference, 1, q-1, powers of two, values around the limb boundaries) and
    random ones; field division through the same path"""
    name = PRIME_NAMES[prime]
    F = Field(name)
    q = F.q
    rng = random.Random(400 + prime)
    hs = hostsim()
    vals = [0, 1, 2, 3, q - 1, q - 2, (q - 1) // 2, (q + 1) // 2] + [2**k for k in range(1, 253, 7)] + \
           [2**(30 * k) - 1 for k in range(1, 9)] + [2**(30 * k) for k in range(1, 9)] + [2**(30 * k) + 1 for k in range(1, 9)] + \
           [q - 2**k for k in range(1, 250, 11)]
    vals = [v % q for v in vals]
    vals += [rng.randrange(q) for _ in range(3000 - len(vals))]
    A = vals
    B = [rng.choice([1, rng.randrange(q)]) for _ in A]
    a, b = ints_to_limbs(A), ints_to_limbs(B)
    c = np.zeros_like(a)
    r = np.zeros((len(A), 4), dtype=np.uint64)
    hs.hs_fr_op(prime, 28, a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), c.ctypes.data_as(ctypes.c_void_p),
                r.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(A)))
    got = limbs_to_ints(r)
    for x, g in zip(A, got):
        assert g == F.inv(x), hex(x)
        assert x == 0 or g * x % q == 1
    # b / a (DIV = 2) goes through the same inverse
    hs.hs_fr_op(prime, OPS["DIV"], b.ctypes.data_as(ctypes.c_void_p), a.ctypes.data_as(ctypes.c_void_p),
                c.ctypes.data_as(ctypes.c_void_p), r.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(A)))
    got = limbs_to_ints(r)
    for x, y, g in zip(A, B, got):
        assert g == F.div(y, x), (hex(y), hex(x))


@pytest.mark.parametrize("flags", [0, 48])
def test_integer_rows_of_the_r1cs_check(flags):
    """The compiled R1CS (r1cs_compile.cpp) with its integer rows (r1cs_small.h: the functions r1cs_small_kernel runs) on the
    value store the tape leaves, against the definition A.w * B.w = C.w evaluated in the field: a SHA-256 compression
    (almost every row is small by shape; with bit inputs no value is wide), the same witness with one entry overwritten
    (0 <-> 1, a byte, a 17-bit value, a field-sized value), and inputs that are not bits (rows marked wide go through
    the general path and must give the same answer)."""
    d = CircuitDesc("bn128")
    d.set_main(C.sha256_compression(d))
    rng = random.Random(50 + flags)
    names = [(n, sz) for n, _g, sz in d.main_inputs()]
    ins = [{n: [rng.randrange(2) for _ in range(sz)] for n, sz in names} for _ in range(2)]
    fc, fp, cnt = hostsim_run_r1cs(d, ins, flags)
    assert fc.tolist() == fp.tolist() == [-1, -1]
    assert cnt[0] > 20000 and cnt[0] > 10 * cnt[2] and cnt[1] == 0
    W = len(hostsim_run(d, ins[:1], flags)[3])
    hits = 0
    for trial in range(24):
        wire = rng.randrange(1, W)
        val = [0, 1, 1, 0, 2, 255, 1 << 16, (1 << 16) - 1, 1 << 40, d.q - 1, d.q >> 1][trial % 11]
        fc, fp, cnt = hostsim_run_r1cs(d, ins[:1], flags, tamper=(wire, val))
        assert fc.tolist() == fp.tolist(), (wire, val)
        hits += fp[0] >= 0
    assert hits >= 8
    # inputs that are not bits: the integer rows that meet them are handed over, the answer stays the definition's
    wild = {n: [rng.choice([0, 1, 1, 1 << 16, 70000, d.q - 1, rng.randrange(d.q)]) for _ in range(sz)] for n, sz in names}
    fc, fp, cnt = hostsim_run_r1cs(d, [wild, ins[0]], flags)
    assert fc.tolist() == fp.tolist() and fc[1] == -1 and cnt[1] > 0


def test_integer_rows_are_left_alone_where_they_do_not_pay():
    """rows of limbs (the big-integer circuit) and small fields (goldilocks: |a*b - c| is not below q) keep the general path"""
    d = CircuitDesc("bn128")
    d.set_main(C.ecdsa_scale(d, 2, 5))
    rng = random.Random(9)
    ins = [{n: [rng.getrandbits(64) for _ in range(sz)] for n, _g, sz in d.main_inputs()}]
    fc, fp, cnt = hostsim_run_r1cs(d, ins, 48)
    assert fc.tolist() == fp.tolist() == [-1] and cnt[0] == 0 and cnt[2] > 0
    g = CircuitDesc("goldilocks")
    g.set_main(C.less_than(g, 12))
    fc, fp, cnt = hostsim_run_r1cs(g, [{"in": [77, 3000]}], 48)
    assert fc.tolist() == fp.tolist() == [-1] and cnt[0] == 0


def test_indexed_load_expansion_and_its_range_assert():
    """`out <-- table[sel]` (Template.load_indexed): the description carries the expansion (EQ / SELECT per element, adder
    trees, an assert that the index hit), the evaluator indexes directly; an index outside the array - where the reference
    reads whatever lies beside it - gives a non-zero status instead of a witness"""
    d = CircuitDesc("bn128")
    d.set_main(C.table_lookup(d, 5))
    t = d.main
    ops, n_tmp = t.expanded_ops()
    assert any(op == OPS["LOADSIG"] for op, *_ in t.ops) and not any(op == OPS["LOADSIG"] for op, *_ in ops)
    assert len(ops) == len(t.ops) - 1 + 5 + 5 + 4 + 4 + 1 and n_tmp > t.n_tmp
    ins = [{"table": [10, 20, 30, 40, 50], "sel": k} for k in range(5)]
    for flags in (0, 4, 48, 112):
        wit, st, _, w2s = hostsim_run(d, ins, flags=flags)
        assert not st.any()
        for k in range(5):
            assert limbs_to_ints(wit[k])[1] == (10 * (k + 1)) ** 2
    bad = [{"table": [1, 2, 3, 4, 5], "sel": 5}, {"table": [1, 2, 3, 4, 5], "sel": d.q - 1}, {"table": [1, 2, 3, 4, 5], "sel": 2**40}]
    _, st, _, _ = hostsim_run(d, bad, flags=48)
    assert (st != 0).all()
    for b in bad:
        with pytest.raises(Exception):
            evaluate(d, b)
