"""File formats checked with an independent python reader: the `.r1cs` the back end writes is parsed here
from the format definition (constraint_writers/src/r1cs_writer.rs:6-14,49-72,246-269,328-341) and its
constraints are evaluated with python integers on oracle witnesses; plus lowering edge cases."""
import os
import struct

import numpy as np
import pytest

from circom_b200.circuit import CircuitDesc
from circom_b200 import circuits as C
from circom_b200.witness_calculator import Circuit, R1cs
from oracle.ir_eval import evaluate
from tests.util import hostsim_run, limbs_to_ints


def parse_r1cs(raw: bytes):
    assert raw[:4] == b"r1cs"
    version, nsec = struct.unpack_from("<II", raw, 4)
    assert version == 1
    pos, secs = 12, {}
    for _ in range(nsec):
        ty, ln = struct.unpack_from("<IQ", raw, pos)
        secs[ty] = (pos + 12, ln)
        pos += 12 + ln
    assert pos == len(raw)
    h, _ = secs[1]
    fs = struct.unpack_from("<I", raw, h)[0]
    q = int.from_bytes(raw[h + 4:h + 4 + fs], "little")
    n_wires, n_out, n_pub, n_prv, n_labels, m = struct.unpack_from("<IIIIQI", raw, h + 4 + fs)
    c, _ = secs[2]
    cons = []
    for _ in range(m):
        row = []
        for _k in range(3):
            n = struct.unpack_from("<I", raw, c)[0]
            c += 4
            lc = {}
            prev = None
            for _j in range(n):
                w = struct.unpack_from("<I", raw, c)[0]
                key = w.to_bytes(4, "little").rstrip(b"\0") or b"\0"
                assert prev is None or prev < key, "wire ids must be sorted as little-endian byte strings"
                prev = key
                lc[w] = int.from_bytes(raw[c + 4:c + 4 + fs], "little")
                c += 4 + fs
            row.append(lc)
        cons.append(row)
    l, ln = secs[3]
    labels = struct.unpack_from("<%dQ" % n_wires, raw, l)
    return dict(q=q, n_wires=n_wires, n_out=n_out, n_pub=n_pub, n_prv=n_prv, n_labels=n_labels, cons=cons, labels=labels)


@pytest.mark.parametrize("name,mk,inp", [
    ("less_than", lambda d: C.less_than(d, 12), {"in": [77, 3000]}),
    ("poseidon", lambda d: C.poseidon(d, 2), {"inputs": [1, 2]}),
    ("ecdsa_small", lambda d: C.ecdsa_scale(d, 1, 2), {"a": [2**64 - 1, 5, 7, 11], "b": [13, 17, 19, 2**63]}),
])
@pytest.mark.parametrize("o0", [False, True])
def test_written_r1cs_is_satisfied_by_oracle_witness(name, mk, inp, o0, tmp_path):
    d = CircuitDesc("bn128")
    d.set_main(mk(d))
    c = Circuit(d, host_only=True, o0=o0)
    p = str(tmp_path / "c.r1cs")
    R1cs(c).write(p, d.main.n_out, 0, d.main.n_in)
    r = parse_r1cs(open(p, "rb").read())
    assert r["q"] == d.q and r["n_wires"] == c.n_witness and len(r["cons"]) == c.stats["n_constraints"]
    assert (r["n_out"], r["n_prv"]) == (d.main.n_out, d.main.n_in)
    sig = evaluate(d, inp)
    w2s = c.witness2signal().astype(np.int64)
    w = [sig[k] for k in w2s]
    if o0:
        assert w2s.tolist() == list(range(d.total_signals))
    for A, B, Cc in r["cons"]:
        a = sum(v * w[k] for k, v in A.items()) % d.q
        b = sum(v * w[k] for k, v in B.items()) % d.q
        cc = sum(v * w[k] for k, v in Cc.items()) % d.q
        assert (a * b - cc) % d.q == 0
    if not o0:   # no `signal = signal` row survives
        for A, B, Cc in r["cons"]:
            trivial = not A and not B and len(Cc) == 2 and 0 not in Cc and sum(Cc.values()) % d.q == 0
            assert not trivial


def _basic_circom(d):
    """`basic.circom` of the reference's documentation (formats/constraints-json.md:34-50)."""
    def internal(t):
        i = t.input("in", 2)
        out = t.output("out")
        t.assign_constrained(out, i[0] * i[1])
    internal_t = d.template("Internal", (), internal)

    def main(t):
        i = t.input("in", 2)
        out = t.output("out")
        c = t.component("c", internal_t)
        t.assign_constrained(c.sig("in", 0), i[0])
        t.assign_constrained(c.sig("in", 1), i[1] + 2 * i[0] + 1)
        t.assign_constrained(out, c.sig("out"))
    return d.template("Main", (), main)


def test_r1cs_of_the_documented_basic_circuit(tmp_path):
    """The only constraint data pinned anywhere in the reference tree: the R1CS of `basic.circom` printed in
    formats/constraints-json.md:57-59 (default --O1) and :77-81 (--O0), with the signal -> witness maps of
    formats/sym.md:50-55 / :69-74.  Wire numbering and every coefficient (q-1 for -1) must agree; the order
    of the rows is the compiler's DAG order and is compared as a set."""
    q = 21888242871839275222246405745257275088548364400416034343698204186575808495616 + 1
    m1 = q - 1
    doc_o1 = [({2: m1}, {4: 1}, {1: m1}), ({}, {}, {0: 1, 2: 2, 3: 1, 4: m1})]
    doc_o0 = [({}, {}, {2: 1, 5: m1}), ({}, {}, {0: 1, 2: 2, 3: 1, 6: m1}), ({}, {}, {1: m1, 4: 1}),
              ({5: m1}, {6: 1}, {4: m1})]
    sym_o1 = [0, 1, 2, 3, 6]            # witness -> signal: sym.md lines "6,4,0,main.c.in[1]"; 4 and 5 eliminated
    # the `--sym` files printed in formats/sym.md:50-55 (--O1) and :69-74 (--O0)
    symfile_o1 = ["1,1,1,main.out", "2,2,1,main.in[0]", "3,3,1,main.in[1]", "4,-1,0,main.c.out", "5,-1,0,main.c.in[0]",
                  "6,4,0,main.c.in[1]"]
    symfile_o0 = ["1,1,1,main.out", "2,2,1,main.in[0]", "3,3,1,main.in[1]", "4,4,0,main.c.out", "5,5,0,main.c.in[0]",
                  "6,6,0,main.c.in[1]"]
    for o0, symfile in ((False, symfile_o1), (True, symfile_o0)):
        d = CircuitDesc("bn128")
        d.set_main(_basic_circom(d))
        c = Circuit(d, host_only=True, o0=o0, symbols=True)
        assert d.sym_lines(c.witness2signal()) == symfile
        p = d.write_sym(str(tmp_path / "basic.sym"), c.witness2signal())
        assert open(p).read() == "".join(x + "\n" for x in symfile)
        # the C library's writer (cw_circuit_write_sym, names from the symbols section of the description)
        c.write_sym(str(tmp_path / "basic_c.sym"))
        assert open(str(tmp_path / "basic_c.sym")).read() == "".join(x + "\n" for x in symfile)
    for o0, doc, sym in ((False, doc_o1, sym_o1), (True, doc_o0, list(range(7)))):
        d = CircuitDesc("bn128")
        assert d.q == q
        d.set_main(_basic_circom(d))
        c = Circuit(d, host_only=True, o0=o0)
        assert c.witness2signal().tolist() == sym
        p = str(tmp_path / ("basic_o%d.r1cs" % (0 if o0 else 1)))
        R1cs(c).write(p, 1, 0, 2)
        r = parse_r1cs(open(p, "rb").read())
        key = lambda row: tuple(tuple(sorted(lc.items())) for lc in row)  # noqa: E731
        assert sorted(map(key, r["cons"])) == sorted(map(key, doc)), (o0, r["cons"])
        # and the witness of the documented shape [1, x*(y+2x+1), x, y, y+2x+1] satisfies it
        x, y = 3, 11
        sig = evaluate(d, {"in": [x, y]})
        w = [sig[k] for k in sym]
        assert w[:4] == [1, x * (y + 2 * x + 1), x, y]


def test_constant_and_inputless_circuits():
    """signals that are compile-time constants, a sub-component without inputs (runs at creation,
    template.rs:274-278), two outputs with the same value, an output equal to an input"""
    d = CircuitDesc("bn128")

    def k(t):
        o = t.output("k", 2)
        t.assign_constrained(o[0], 41)
        t.assign_constrained(o[1], t.const(41) + 1)
    kt = d.template("Konst", (), k)

    def main(t):
        x = t.input("x")
        o = t.output("o", 4)
        c = t.component("k", kt)
        t.assign(o[0], c["k", 0] + x)
        t.assign(o[1], c["k", 1])
        t.assign(o[2], x)          # output aliases an input without a constraint
        t.assign(o[3], x * 1)      # and a second alias of the same value
    d.set_main(d.template("Main", (), main))
    wit, st, stats, w2s = hostsim_run(d, [{"x": 5}, {"x": 0}])
    for i, x in enumerate((5, 0)):
        exp = evaluate(d, {"x": x})
        assert exp[1:5] == [41 + x, 42, x, x]
        assert limbs_to_ints(wit[i]) == [exp[k] for k in w2s]


def test_batch_of_one_and_status_of_bad_input():
    d = CircuitDesc("bn128")
    d.set_main(C.num2bits(d, 8))
    wit, st, _, w2s = hostsim_run(d, [{"in": 255}])
    assert st.tolist() == [0]
    wit, st, _, _ = hostsim_run(d, [{"in": 256}, {"in": 3}])   # 256 does not fit 8 bits: recomposition assert fails
    assert st[0] > 0 and st[1] == 0


def test_failing_assert_number_matches_oracle():
    """status = 1 + number of the first failing `===` in the reference's execution order, also after the
    lowering fused / dropped asserts"""
    from oracle.c_oracle import COracle
    from tests.util import flat_inputs
    d = CircuitDesc("bn128")
    d.set_main(C.num2bits(d, 8))
    ins = [{"in": 256}, {"in": 255}, {"in": 2**200}]
    wit, st, _, _ = hostsim_run(d, ins)
    ow, ost = COracle(d.to_bytes()).run(flat_inputs(d, ins))
    assert st.tolist() == ost.tolist() == [9, 0, 9]


def test_r1cs_with_custom_gate_sections_loads(tmp_path):
    """circom writes two more sections (types 4 and 5: custom gates list / applications,
    constraint_writers/src/r1cs_writer.rs:356-454) when a circuit uses custom templates; a reader must skip
    sections it does not consume (r1cs_reader.rs walks sections by type and size)."""
    import ctypes
    from circom_b200 import native
    d = CircuitDesc("bn128")
    d.set_main(C.less_than(d, 8))
    c = Circuit(d, host_only=True)
    p = str(tmp_path / "a.r1cs")
    R1cs(c).write(p, 1, 0, 2)
    raw = bytearray(open(p, "rb").read())
    nsec = struct.unpack_from("<I", raw, 8)[0]
    raw[8:12] = struct.pack("<I", nsec + 2)
    raw += struct.pack("<IQ", 4, 4) + struct.pack("<I", 0)
    raw += struct.pack("<IQ", 5, 4) + struct.pack("<I", 0)
    p2 = str(tmp_path / "b.r1cs")
    open(p2, "wb").write(raw)
    r = R1cs(p2)
    assert r.n_constraints == c.stats["n_constraints"] and r.n_wires == c.n_witness


def test_custom_gate_sections_round_trip(tmp_path):
    """sections 4 / 5 (custom gates used: NUL-terminated template name + field-element parameters; custom gates applied:
    gate index + wire list - constraint_writers/src/r1cs_writer.rs:356-454, read back by r1cs_reader.rs:343-419) survive
    load -> write byte for byte, in the reference's section order (constraints, header, wire map, gates used, applied)"""
    d = CircuitDesc("bn128")
    d.set_main(C.less_than(d, 8))
    c = Circuit(d, host_only=True)
    p = str(tmp_path / "a.r1cs")
    R1cs(c).write(p)
    raw = bytearray(open(p, "rb").read())
    raw[8:12] = struct.pack("<I", 5)
    used = struct.pack("<I", 2)
    used += b"Poseidon\0" + struct.pack("<I", 2) + (3).to_bytes(32, "little") + (d.q - 1).to_bytes(32, "little")
    used += b"EmptyGate\0" + struct.pack("<I", 0)
    applied = struct.pack("<I", 3)
    applied += struct.pack("<II", 0, 3) + struct.pack("<3Q", 1, 2, 5)
    applied += struct.pack("<II", 1, 0)
    applied += struct.pack("<II", 0, 1) + struct.pack("<Q", 7)
    raw += struct.pack("<IQ", 4, len(used)) + used + struct.pack("<IQ", 5, len(applied)) + applied
    p2 = str(tmp_path / "b.r1cs")
    open(p2, "wb").write(raw)
    r = R1cs(p2)
    assert r.n_constraints == c.stats["n_constraints"]
    p3 = str(tmp_path / "c.r1cs")
    r.write(p3)
    assert open(p3, "rb").read() == bytes(raw)
    # an application that names a gate which is not in the list is refused
    bad = bytearray(raw)
    off = len(raw) - len(applied) + 4
    bad[off:off + 4] = struct.pack("<I", 9)
    open(p2, "wb").write(bad)
    with pytest.raises(Exception):
        R1cs(p2)


def test_wtns_reader_round_trip(tmp_path):
    """cw_wtns_read parses what writeBinWitness writes (main.cpp:288-334; golden fixture bytes from the reference
    calculator) and refuses damaged files"""
    import ctypes
    import zlib
    from circom_b200 import native
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    raw = zlib.decompress(open(os.path.join(here, "poseidon2_0.wtns.z"), "rb").read())
    p = str(tmp_path / "p.wtns")
    open(p, "wb").write(raw)
    pid, n = ctypes.c_int(), ctypes.c_uint64()
    assert native.lib.cw_wtns_read(p.encode(), ctypes.byref(pid), ctypes.byref(n), None, 0) == 0
    assert pid.value == 0 and n.value * 32 + 76 == len(raw)
    out = np.zeros((n.value, 4), dtype=np.uint64)
    assert native.lib.cw_wtns_read(p.encode(), ctypes.byref(pid), ctypes.byref(n), out.ctypes.data, n.value) == 0
    assert out.tobytes() == raw[76:]
    for damaged in (raw[:100], raw[:4] + b"\x03" + raw[5:], b"wtnz" + raw[4:], raw[:76] + b"\xff" * 32 + raw[108:]):
        open(p, "wb").write(damaged)
        assert native.lib.cw_wtns_read(p.encode(), ctypes.byref(pid), ctypes.byref(n), None, 0) == native.CW_EFORMAT


def test_goldilocks_files_carry_8_byte_elements(tmp_path):
    """goldilocks: field size 8 in `.r1cs` (constraint_list/src/r1cs_porting.rs:6-10: whole 64-bit words of the prime)
    and n8 = 8 in `.wtns` (c_elements/common64/main.cpp:312-353); in memory a value stays 4 x u64 with zero upper words"""
    import ctypes
    from circom_b200 import native
    d = CircuitDesc("goldilocks")
    d.set_main(C.less_than(d, 12))
    c = Circuit(d, host_only=True)
    p = str(tmp_path / "g.r1cs")
    R1cs(c).write(p, d.main.n_out, 0, d.main.n_in)
    raw = open(p, "rb").read()
    r = parse_r1cs(raw)
    h = raw.index(struct.pack("<IQ", 1, 4 + 8 + 28)) + 12      # header section: 4 + field size + 28 bytes
    assert struct.unpack_from("<I", raw, h)[0] == 8 and r["q"] == d.q == 2**64 - 2**32 + 1
    sig = evaluate(d, {"in": [77, 3000]})
    w = [sig[k] for k in c.witness2signal().astype(np.int64)]
    assert len(r["cons"]) == c.stats["n_constraints"] > 0
    for A, B, Cc in r["cons"]:
        a, b, cc = (sum(v * w[k] for k, v in X.items()) % d.q for X in (A, B, Cc))
        assert (a * b - cc) % d.q == 0
    # read -> write reproduces the file
    r2 = R1cs(p)
    p2 = str(tmp_path / "g2.r1cs")
    r2.write(p2)
    assert open(p2, "rb").read() == raw
    # .wtns as the reference's goldilocks runtime writes it
    n = len(w)
    wt = b"wtns" + struct.pack("<II", 2, 2) + struct.pack("<IQ", 1, 8 + 8) + struct.pack("<IQI", 8, d.q, n) + \
        struct.pack("<IQ", 2, 8 * n) + b"".join(struct.pack("<Q", v) for v in w)
    wp = str(tmp_path / "g.wtns")
    open(wp, "wb").write(wt)
    pid, cnt = ctypes.c_int(), ctypes.c_uint64()
    out = np.zeros((n, 4), dtype=np.uint64)
    assert native.lib.cw_wtns_read(wp.encode(), ctypes.byref(pid), ctypes.byref(cnt), out.ctypes.data, n) == 0
    assert pid.value == 7 and cnt.value == n and limbs_of(out) == w
    # a 32-byte-element file that names the goldilocks prime is malformed; so is a value that is not reduced
    bad = b"wtns" + struct.pack("<II", 2, 2) + struct.pack("<IQ", 1, 8 + 32) + struct.pack("<I", 32) + d.q.to_bytes(32, "little") + \
        struct.pack("<I", 1) + struct.pack("<IQ", 2, 32) + (1).to_bytes(32, "little")
    open(wp, "wb").write(bad)
    assert native.lib.cw_wtns_read(wp.encode(), ctypes.byref(pid), ctypes.byref(cnt), None, 0) == native.CW_EFORMAT
    open(wp, "wb").write(wt[:-8] + struct.pack("<Q", d.q))
    assert native.lib.cw_wtns_read(wp.encode(), ctypes.byref(pid), ctypes.byref(cnt), None, 0) == native.CW_EFORMAT


def limbs_of(a):
    return [int.from_bytes(r.tobytes(), "little") for r in np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)]


def test_failed_assert_message_names_template_and_component_trace():
    """cw_circuit_assert_info: the reference's "Failed assert in template/function <T> ... Followed trace of components: <path>"
    (c_code_generator.rs:461-468) for the assert number the status reports - checked against where the evaluator raises,
    with asserts inside nested sub-components (LessThan -> Num2Bits), with and without the symbols section"""
    from tests.util import hostsim_run
    from oracle.ir_eval import AssertFailed
    d = CircuitDesc("bn128")
    lt = C.less_than(d, 4)

    def build(t):
        a = t.input("a", 2)
        out = t.output("out")
        c0 = t.component("lo", lt)
        c1 = t.component("hi[1]", lt)
        t.assign_constrained(c0["in", 0], a[0])
        t.assign_constrained(c0["in", 1], 5)
        t.assign_constrained(c1["in", 0], a[1])
        t.assign_constrained(c1["in", 1], 9)
        t.assign_constrained(out, c0["out"] + c1["out"])
    d.set_main(d.template("Two", (), build))
    c = Circuit(d, host_only=True, symbols=True)
    plain = Circuit(d, host_only=True)
    # a[1] = 100 overflows the 5-bit range check inside hi[1]'s Num2Bits: (100 + 16 - 9) needs 7 bits
    _, st, _, _ = hostsim_run(d, [{"a": [3, 100]}, {"a": [100, 3]}, {"a": [3, 4]}])
    assert st[2] == 0 and st[0] > 0 and st[1] > 0
    m0, m1 = c.assert_info(int(st[0]) - 1), c.assert_info(int(st[1]) - 1)
    assert m0 == "Failed assert in template/function Num2Bits_5. Followed trace of components: main.hi[1].n2b"
    assert m1 == "Failed assert in template/function Num2Bits_5. Followed trace of components: main.lo.n2b"
    assert plain.assert_info(int(st[0]) - 1) == "Failed assert in template/function Num2Bits_5"
    for inp, name in (({"a": [3, 100]}, "Num2Bits_5"), ({"a": [100, 3]}, "Num2Bits_5")):
        with pytest.raises(AssertFailed, match=name):
            evaluate(d, inp)
    with pytest.raises(Exception):
        c.assert_info(10**6)
