"""The C-ABI library builds, loads and exports every symbol include/circom_b200.h declares.
No compute calls here (no GPU); lowering-only entry points are exercised with CW_FLAG_HOST_ONLY."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest

from circom_b200 import native
from circom_b200.circuit import CircuitDesc
from circom_b200 import circuits as C
from circom_b200.witness_calculator import Circuit, R1cs, fnv_hash, qualify_input, parse_value

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "circom_b200.h")).read()
    names = set(re.findall(r"\b(cw_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) > 35
    for n in names:
        assert hasattr(native.lib, n), "missing export " + n
    assert native.lib.cw_version() >= 100


def test_no_device_is_an_error_not_a_fallback():
    if native.lib.cw_device_count() > 0:
        pytest.skip("GPU present")
    d = CircuitDesc("bn128")
    d.set_main(C.multiplier2(d))
    c = Circuit(d)
    h = ctypes.c_void_p()
    rc = native.lib.cw_batch_create(c._h, 4, 0, ctypes.byref(h))
    assert rc == native.CW_ENODEV
    assert b"no CPU execution path" in native.lib.cw_last_error()


def test_metadata_matches_reference_getters():
    d = CircuitDesc("bn128")
    d.set_main(C.multiplier2(d))
    c = Circuit(d, host_only=True)
    L = native.lib
    assert L.cw_get_main_input_signal_start(c._h) == 2      # get_main_input_signal_start = outputs + 1
    assert L.cw_get_main_input_signal_no(c._h) == 2
    assert L.cw_get_total_signal_no(c._h) == 4
    assert L.cw_get_size_of_witness(c._h) == 4
    assert L.cw_get_number_of_components(c._h) == 1
    assert L.cw_get_size_of_input_hashmap(c._h) == 256      # max(256, 2^ceil(log2 n)) c_elements/mod.rs:167-169
    assert c.input_signal_size("a") == 1 and c.input_signal_size("b") == 1
    assert c.input_signal_id("a") == 2 and c.input_signal_id("b") == 3
    assert c.input_signal_size("nope") == -1
    assert fnv_hash("a") == L.cw_fnv1a(b"a")
    assert c.prime == d.q


def test_dat_layout(tmp_path):
    """hash map (24-byte entries, linear probing on hash % size) + witness2signal list,
    c_code_generator.rs:575-603,605-614."""
    d = CircuitDesc("bn128")
    d.set_main(C.multiplier2(d))
    c = Circuit(d, host_only=True)
    p = str(tmp_path / "m.dat")
    c.write_dat(p)
    raw = open(p, "rb").read()
    assert len(raw) == 256 * 24 + 4 * 8 + 40 * len(d.consts)
    ent = {}
    for i in range(256):
        h, sid, sz = struct.unpack_from("<QQQ", raw, i * 24)
        if sid:
            ent[h] = (i, sid, sz)
    assert ent[fnv_hash("a")][1:] == (2, 1) and ent[fnv_hash("b")][1:] == (3, 1)
    assert ent[fnv_hash("a")][0] == fnv_hash("a") % 256
    assert struct.unpack_from("<4Q", raw, 256 * 24) == (0, 1, 2, 3)


def test_input_flattening_rules():
    out = {}
    qualify_input("", {"a": 1, "b": {"c": [1, 2], "d": [{"x": 1}, {"x": 2}]}}, out)
    assert out == {"a": 1, "b.c": [1, 2], "b.d[0].x": 1, "b.d[1].x": 2}
    q = 101
    assert parse_value("0x10", q) == 16 and parse_value("0b11", q) == 3 and parse_value("0o17", q) == 15
    assert parse_value("205", q) == 3 and parse_value(7, q) == 7
    with pytest.raises(ValueError):
        parse_value("-3", q)      # the reference accepts no sign (main.cpp:126-142)


def test_r1cs_write_read_roundtrip(tmp_path):
    d = CircuitDesc("bls12381")
    d.set_main(C.less_than(d, 8))
    c = Circuit(d, host_only=True)
    r = R1cs(c)
    assert r.n_constraints == c.stats["n_constraints"] and r.prime_id == 1
    p1, p2 = str(tmp_path / "a.r1cs"), str(tmp_path / "b.r1cs")
    r.write(p1, 1, 0, 2)
    r2 = R1cs(p1)
    assert (r2.n_wires, r2.n_constraints, r2.nnz) == (r.n_wires, r.n_constraints, r.nnz)
    r2.write(p2, 1, 0, 2)
    assert open(p1, "rb").read() == open(p2, "rb").read()
    raw = open(p1, "rb").read()
    assert raw[:4] == b"r1cs" and struct.unpack_from("<II", raw, 4) == (1, 3)
    assert struct.unpack_from("<I", raw, 12)[0] == 2   # constraints section comes first (r1cs_porting.rs:19-53)


def test_packed_record_layout_and_host_expansion():
    """cw_circuit_pack_info describes the packed device->host record; cw_circuit_expand_record (the host half of
    cw_batch_get_witness) turns a record into the reference's 32-byte rows.  A record assembled in Python from an
    oracle witness must expand to that witness with every store width the CPU has (SSE2 / AVX2 / AVX-512 paths)."""
    import random
    from circom_b200.circuits.bigint import ecdsa_scale
    from oracle.ir_eval import evaluate
    rng = random.Random(4)
    for compact in (False, True):
        d = CircuitDesc("bn128")
        d.set_main(ecdsa_scale(d, 1, 2))
        c = Circuit(d, host_only=True, compact=compact)
        info, ent = c.pack_info()
        words, n_plane, n_xbits, n_u64, n_full = info
        assert words * 4 < c.n_witness * 32 // 8 and (words % 4) == 0
        if compact:
            assert n_plane == c.stats["n_bitwords"] > 0
        inp = {"a": [rng.getrandbits(64) for _ in range(4)], "b": [rng.getrandbits(64) for _ in range(4)]}
        exp = evaluate(d, inp)
        w2s = c.witness2signal().astype(np.int64)
        wit = [exp[k] for k in w2s]
        rec = np.zeros(words, dtype=np.uint32)
        off = [0, n_plane, n_plane + n_xbits, n_plane + n_xbits + 2 * n_u64]
        for v, e in zip(wit, ent.tolist()):
            cls, idx = e >> 30, e & 0x3FFFFFFF
            if cls <= 1:
                assert v in (0, 1)
                rec[off[cls] + (idx >> 5)] |= np.uint32(v << (idx & 31))
            elif cls == 2:
                assert v < 2**64
                rec[off[2] + 2 * idx] = v & 0xFFFFFFFF
                rec[off[2] + 2 * idx + 1] = v >> 32
            else:
                for k in range(8):
                    rec[off[3] + 8 * idx + k] = (v >> (32 * k)) & 0xFFFFFFFF
        for bits in (128, 256, 512, 0):
            for misalign in (0, 1):     # 32-byte aligned rows take the streaming stores, others plain ones
                buf = np.full(c.n_witness * 4 + 8, 0xDEADBEEFDEADBEEF, dtype=np.uint64)
                base = (-buf.ctypes.data // 8) % 4 + misalign
                rows = buf[base:base + c.n_witness * 4]
                assert native.lib.cw_circuit_expand_record(c._h, rec.ctypes.data, rows.ctypes.data, bits) == 0
                got = [int.from_bytes(rows[4 * i:4 * i + 4].tobytes(), "little") for i in range(c.n_witness)]
                assert got == wit, (compact, bits, misalign)
                assert int(buf[base + c.n_witness * 4]) == 0xDEADBEEFDEADBEEF and (base == 0 or int(buf[base - 1]) == 0xDEADBEEFDEADBEEF)
    assert native.lib.cw_host_expand_isa() in (b"avx512", b"avx2", b"sse2")


@pytest.mark.parametrize("prime", ["bn128", "bls12381", "goldilocks"])
def test_dat_equals_the_file_the_reference_runtime_loads(prime, tmp_path):
    """cw_circuit_write_dat (--O0 witness list) == the .dat the reference calculators of oracle/_ref are linked with
    (oracle/emit_ref_cpp.dat_bytes, consumed by the reference's loadCircuit, main.cpp:22-124): hash map, witness list
    and the constants in their 40-byte tagged Montgomery form (short constants incl. negative ones, long ones); for a
    component array of mixed templates also templateInsId2IOSignalInfo (c_code_generator.rs:681-735) - the reference
    calculator of that circuit (oracle/_ref/calc/mixed_array) reads its sub-component signals through it"""
    from oracle.emit_ref_cpp import dat_bytes, io_map_bytes
    for mk in (lambda d: C.all_ops(d), lambda d: C.poseidon(d, 2), lambda d: C.int_div(d, 32), lambda d: C.less_than(d, 8),
               lambda d: C.mixed_array(d)):
        d = CircuitDesc(prime)
        d.set_main(mk(d))
        d.const_id(-5)
        d.const_id(2**31 - 1)
        d.const_id(-2**31)
        d.const_id(2**31)
        d.const_id(-2**31 - 1)
        c = Circuit(d, host_only=True, o0=True)
        p = str(tmp_path / "c.dat")
        c.write_dat(p)
        assert open(p, "rb").read() == dat_bytes(d)
    assert len(io_map_bytes(d)) == 4 * (3 + 3 + 3 * 3 * 4) and open(p, "rb").read().endswith(io_map_bytes(d))
