"""Circuit descriptions (.cb2c) and .r1cs files are untrusted inputs of the library: truncated, corrupted or
hostile files must be rejected with an error code (or load, if the damage was harmless) - never crash or
allocate without bound.  (The same corpus was run under AddressSanitizer/UBSan while hardening the parser.)"""
import ctypes
import random
import struct

import pytest

from circom_b200 import circuits as C
from circom_b200 import native
from circom_b200.circuit import CircuitDesc
from circom_b200.witness_calculator import Circuit, R1cs

lib = native.lib
CW_FLAG_HOST_ONLY = 2


def try_load(buf: bytes) -> int:
    h = ctypes.c_void_p()
    rc = lib.cw_circuit_load_mem(buf, len(buf), CW_FLAG_HOST_ONLY, ctypes.byref(h))
    if rc == 0:
        lib.cw_circuit_destroy(h)
    return rc


def mutate(rng, src: bytes) -> bytes:
    b = bytearray(src)
    mode = rng.randrange(5)
    if mode == 0:
        b = b[:rng.randrange(len(b))]
    elif mode == 1:
        for _ in range(rng.randrange(1, 6)):
            b[rng.randrange(len(b))] = rng.randrange(256)
    elif mode == 2:
        for _ in range(rng.randrange(1, 4)):
            p = rng.randrange(0, len(b) - 4)
            b[p:p + 4] = rng.choice([0, 1, 2, 3, 0xFFFFFFFF, 0x7FFFFFFF, 0x80000000, rng.randrange(2**32),
                                     rng.randrange(64)]).to_bytes(4, "little")
    elif mode == 3:
        p = rng.randrange(len(b))
        b = b[:p] + bytes(rng.randrange(256) for _ in range(rng.randrange(1, 40))) + b[p:]
    else:
        p = rng.randrange(0, len(b) - 8)
        v = int.from_bytes(b[p:p + 8], "little") ^ (1 << rng.randrange(64))
        b[p:p + 8] = v.to_bytes(8, "little")
    return bytes(b)


def descriptions():
    out = []
    for mk in (lambda d: C.multiplier2(d), lambda d: C.less_than(d, 8), lambda d: C.int_div(d, 32),
               lambda d: C.num2bits(d, 16), lambda d: C.is_zero(d), lambda d: C.all_ops(d)):
        d = CircuitDesc("bn128")
        d.set_main(mk(d))
        out.append(d.to_bytes())
    return out


def test_mutated_descriptions_are_rejected_or_load():
    rng = random.Random(2024)
    srcs = descriptions()
    codes = {0: 0, native.CW_EFORMAT: 0}
    for _ in range(1500):
        rc = try_load(mutate(rng, rng.choice(srcs)))
        assert rc in codes, rc
        codes[rc] += 1
    assert codes[native.CW_EFORMAT] > 500 and codes[0] > 50   # both outcomes occur


def test_hostile_descriptions():
    d = CircuitDesc("bn128")
    d.set_main(C.multiplier2(d))
    good = d.to_bytes()
    assert try_load(good) == 0
    # header: magic, version, prime, n_consts, n_templates, main, n_names, n_funcs
    head = struct.unpack_from("<4s7I", good)
    for field, value in ((3, 0xFFFFFFFF), (4, 0xFFFFFFFF), (4, 0), (5, 7), (6, 0x10000000), (7, 0xFFFFFF), (2, 9), (1, 2)):
        h = list(head)
        h[field] = value
        assert try_load(struct.pack("<4s7I", *h) + good[32:]) == native.CW_EFORMAT, field
    # a name whose length field is 2^32 - 1
    assert try_load(good[:32 + 32 * head[3]] + struct.pack("<I", 0xFFFFFFFF) + good[32 + 32 * head[3] + 4:]) == native.CW_EFORMAT
    # 40 nested templates with two sub-components each describe 2^40 components in a few kilobytes
    d = CircuitDesc("bn128")
    t = C.multiplier2(d)
    for k in range(40):
        def build(tt, prev=t, k=k):
            a = tt.input("a")
            b = tt.input("b")
            o = tt.output("o")
            x, y = tt.component("x", prev), tt.component("y", prev)
            names = [n for n, _ in prev.sigs["in"]]
            for comp in (x, y):
                tt.assign_constrained(comp[names[0]], a)
                tt.assign_constrained(comp[names[1]], b)
            out = prev.sigs["out"][0][0]
            tt.assign_constrained(o, x[out] + y[out])
        t = d.template("Nest%d" % k, (), build)
    d.set_main(t)
    assert try_load(d.to_bytes()) == native.CW_EFORMAT
    assert b"too large" in lib.cw_last_error()


def test_mutated_r1cs_files(tmp_path):
    rng = random.Random(7)
    srcs = []
    for k, mk in enumerate((lambda d: C.multiplier2(d), lambda d: C.less_than(d, 8), lambda d: C.num2bits(d, 16))):
        d = CircuitDesc("bn128")
        d.set_main(mk(d))
        p = str(tmp_path / ("src%d.r1cs" % k))
        R1cs(Circuit(d, host_only=True)).write(p, 1, 0, 2)
        srcs.append(open(p, "rb").read())
    ok = bad = 0
    for i in range(600):
        p = str(tmp_path / "m.r1cs")
        open(p, "wb").write(mutate(rng, rng.choice(srcs)))
        h = ctypes.c_void_p()
        rc = lib.cw_r1cs_load(p.encode(), ctypes.byref(h))
        assert rc in (0, native.CW_EFORMAT, native.CW_EIO), rc
        if rc == 0:
            ok += 1
            lib.cw_r1cs_destroy(h)
        else:
            bad += 1
    assert ok > 20 and bad > 200


def test_cli_rejects_pathological_json(tmp_path):
    """input.json is untrusted too: the CLI's reader bounds its recursion (a 2-million-deep array used to
    overflow the stack) and reports malformed files as errors"""
    import os
    import subprocess
    from circom_b200 import build
    cli = os.path.join(os.path.dirname(build.LIB), "circom_cuda_witness")
    if not os.path.exists(cli):
        pytest.skip("CLI not built")
    d = CircuitDesc("bn128")
    d.set_main(C.less_than(d, 8))
    cb = d.save(str(tmp_path / "lt.cb2c"))
    for text, needle in (('{"in":[' + "[" * 2000000 + "1" + "]" * 2000000 + ",2]}", b"nesting too deep"),
                         ('{"in":[1,2', b"JSON"), ("", b"JSON"), ('{"in":["12', b"JSON")):
        p = str(tmp_path / "in.json")
        open(p, "w").write(text)
        r = subprocess.run([cli, cb, p, str(tmp_path / "o.wtns")], capture_output=True)
        assert r.returncode == 1 and needle in r.stderr, (text[:20], r.returncode, r.stderr[-200:])
