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


def dsl_conf(log=False):
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
        if log:
            t.log("p =", p)
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


def test_symbols_section_and_sym_file(tmp_path):
    """The optional symbols section, written from the spec by the C writer, equals the DSL's; `cw_circuit_write_sym` prints the
    reference's `.sym` lines (sym_writer.rs:10-14: `signal,witness,node,name`; order of dag/src/sym_porting.rs:16-33: a
    component's signals, then its sub-components; witness = -1 for a signal that the signal merging removed)."""
    exe, out = str(tmp_path / "cb2c_conf"), str(tmp_path / "conf_sym.cb2c")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cb2c_writer", "cb2c_conf.c")])
    subprocess.check_call([exe, out, "sym"])
    blob = open(out, "rb").read()
    d = dsl_conf()
    assert blob == d.to_bytes(symbols=True) and blob.startswith(d.to_bytes()) and blob[len(d.to_bytes()):][:4] == b"SYMS"
    c = Circuit(blob, host_only=True)
    assert c.stats == Circuit(d, host_only=True).stats      # names change nothing of the lowering
    sym = str(tmp_path / "conf.sym")
    c.write_sym(sym)
    assert open(sym).read() == ("1,1,1,main.bits[0]\n2,2,1,main.bits[1]\n3,3,1,main.p\n4,4,1,main.x\n5,5,1,main.y\n"
                                "6,6,0,main.m.c\n7,-1,0,main.m.a\n8,-1,0,main.m.b\n")
    # --O0 keeps every signal in the witness
    Circuit(blob, host_only=True, o0=True).write_sym(sym)
    assert [ln.split(",")[1] for ln in open(sym).read().split()] == [str(i) for i in range(1, 9)]
    # a description without the section: refused, not guessed
    import pytest
    with pytest.raises(Exception, match="no symbols"):
        Circuit(d, host_only=True).write_sym(sym)
    # a bigger tree: one line per signal but the constant one, ids ascending, the witness column is the inverse of
    # witness2signal, paths follow the component tree
    d2 = CircuitDesc("bn128")
    d2.set_main(C.ecdsa_scale(d2, 1, 2))
    c2 = Circuit(d2, host_only=True, symbols=True)
    c2.write_sym(sym)
    lines = [ln.split(",") for ln in open(sym).read().split()]
    assert [int(x[0]) for x in lines] == list(range(1, d2.total_signals))
    w2s = c2.witness2signal().tolist()
    inv = {s: i for i, s in enumerate(w2s)}
    assert all(int(x[1]) == inv.get(int(x[0]), -1) for x in lines)
    assert lines[0][3].startswith("main.") and any(x[3].count(".") >= 3 for x in lines)
    assert len({x[3] for x in lines}) == len(lines)         # qualified names are unique
    assert open(sym).read() == "".join(x + "\n" for x in d2.sym_lines(w2s))   # = the Python writer, node ids included
    for mk in (lambda dd: C.int_div(dd, 8), lambda dd: C.sha256(dd, 8), lambda dd: C.poseidon(dd, 2)):
        d3 = CircuitDesc("bn128")
        d3.set_main(mk(d3))
        c3 = Circuit(d3, host_only=True, symbols=True)
        c3.write_sym(sym)
        assert open(sym).read() == "".join(x + "\n" for x in d3.sym_lines(c3.witness2signal()))


def test_io_map_section_from_the_spec(tmp_path):
    """The optional io-map section, written from the spec by the C writer, equals the DSL's (alone and followed by the
    symbols); the library copies it into the `.dat` in the layout the reference's loadCircuit reads (main.cpp:57-93)"""
    from oracle.emit_ref_cpp import dat_bytes, io_map_bytes
    exe = str(tmp_path / "cb2c_conf")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cb2c_writer", "cb2c_conf.c")])
    d = dsl_conf()
    d.mark_mixed_array(*d.templates)
    for mode, sym in (("iomap", False), ("iomap+sym", True)):
        out = str(tmp_path / ("conf_%s.cb2c" % mode))
        subprocess.check_call([exe, out, mode])
        blob = open(out, "rb").read()
        assert blob == d.to_bytes(symbols=sym) and b"IOMP" in blob
        c = Circuit(blob, host_only=True, o0=True)
        p = str(tmp_path / "conf.dat")
        c.write_dat(p)
        assert open(p, "rb").read() == dat_bytes(d) and dat_bytes(d).endswith(io_map_bytes(d))
    # template ids, then per template: #signals, {offset, #dims - 1, dims but the first, size, bus}
    assert io_map_bytes(d) == np.array([0, 1, 3, 0, 0, 1, 0, 1, 0, 1, 0, 2, 0, 1, 0,
                                        4, 0, 0, 1, 0, 2, 0, 1, 0, 3, 0, 1, 0, 4, 0, 1, 0], dtype="<u4").tobytes()


def test_log_ops_and_string_table_from_the_spec(tmp_path):
    """LOG ops and the LOGS section written from the spec by the C writer equal the DSL's (alone and with the two other
    optional sections behind them); the text comes out as the reference prints it"""
    exe = str(tmp_path / "cb2c_conf")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cb2c_writer", "cb2c_conf.c")])
    d = dsl_conf(log=True)
    out = str(tmp_path / "conf_log.cb2c")
    subprocess.check_call([exe, out, "log"])
    blob = open(out, "rb").read()
    assert blob == d.to_bytes() and b"LOGS" in blob
    d.mark_mixed_array(*d.templates)
    subprocess.check_call([exe, out, "log+iomap+sym"])
    assert open(out, "rb").read() == d.to_bytes(symbols=True)
    c = Circuit(blob, host_only=True)
    from oracle.ir_eval import evaluate
    sig = evaluate(d, {"x": 3, "y": 11})
    from tests.util import ints_to_limbs
    assert c.format_log(ints_to_limbs([sig[k] for k in c.witness2signal().astype(np.int64)])) == "p = 34\n"


def test_rust_producer_numbers_match_the_format():
    """integration/cuda_elements is not compiled in this image; what can be pinned is that the numbers it writes are the
    format's: the opcode enum against circuit.py's OPS (= what flatten.cpp::parse accepts), the reference kinds, the prime
    ids, the section tags and the order of the header words"""
    import os
    import re
    from circom_b200.circuit import OPS, PRIME_IDS
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "integration", "cuda_elements")
    src = open(os.path.join(root, "code_producers", "src", "cuda_elements", "mod.rs")).read()
    body = re.search(r"pub enum Op \{(.*?)\}", src, re.S).group(1)
    val, ops = 0, {}
    for item in [x.strip() for x in body.replace("\n", " ").split(",") if x.strip()]:
        m = re.match(r"(\w+)(?:\s*=\s*(\d+))?$", item)
        val = int(m.group(2)) if m.group(2) else val + 1
        ops[m.group(1)] = val
    assert ops and all(OPS[k] == v for k, v in ops.items() if k != "INV") and ops["INV"] == 28
    assert {"MUL", "SELECT", "ASSERT_EQ", "CALL", "ARG", "LOADX", "STOREX", "RET", "JZ", "JMP"} <= set(ops)
    for name, pid in PRIME_IDS.items():
        assert re.search(r'"%s" => Ok\(%d\)' % (name, pid), src), name
    kinds = re.search(r"let \(k, s, i\).*?match self \{(.*?)\};", src, re.S).group(1)
    for variant, k in (("None", 0), ("Own", 1), ("Sub", 2), ("Const", 3), ("Tmp", 4), ("One", 5)):
        assert re.search(r"Ref::%s\b[^=]*=> \(%d," % (variant, k), kinds), variant
    assert src.index('b"CB2C"') < src.index('b"LOGS"') < src.index('b"IOMP"') < src.index('b"SYMS"') and ops["LOG"] == 29
    head = re.search(r"for v in \[1u32, self\.prime, (.*?)\]", src, re.S).group(1)
    assert [w.strip().split(".")[1].split(" ")[0] for w in head.split(",")][:3] == ["consts", "templates", "main"]
    low = open(os.path.join(root, "compiler", "src", "cuda_lowering.rs")).read()
    for fn in ("mapped_address", "dynamic_address", "indexed_signal_load", "produce_cb2c"):
        assert "fn %s" % fn in low
