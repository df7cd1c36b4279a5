"""Builds the REAL reference witness calculators (reference runtime + hand-lowered <circuit>.cpp,
see emit_ref_cpp.py) for the circuits the parity tests and the CPU baseline use, into
oracle/_ref/calc/ (TEST INFRASTRUCTURE).  Needs /root/reference; on the GPU box the prebuilt
binaries are used.  A calculator is rebuilt only when missing (the big ones take minutes of g++)."""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
CALC_DIR = os.path.join(HERE, "_ref", "calc")


def circuits():
    """name -> (prime, builder(desc) -> main template)"""
    from circom_b200 import circuits as C
    return {
        "multiplier2": ("bn128", lambda d: C.multiplier2(d)),
        "all_ops": ("bn128", lambda d: C.all_ops(d)),
        "all_ops_bls": ("bls12381", lambda d: C.all_ops(d)),
        "less_than8": ("bn128", lambda d: C.less_than(d, 8)),
        # goldilocks: the reference's other runtime (c_elements/common64 + goldilocks/fr.hpp), u64 values, operators as expressions
        "all_ops_gl": ("goldilocks", lambda d: C.all_ops(d)),
        "less_than8_gl": ("goldilocks", lambda d: C.less_than(d, 8)),
        "mixed_array_gl": ("goldilocks", lambda d: C.mixed_array(d)),
        # a component array of mixed templates: the calculator reads sub-component signals through the io map of its .dat
        "mixed_array": ("bn128", lambda d: C.mixed_array(d)),
        # `out <-- table[sel]`: the calculator loads at a run-time address, the description carries the expansion
        "table_lookup8": ("bn128", lambda d: C.table_lookup(d, 8)),
        # log() calls: the calculator prints them (stdout is part of the comparison)
        "logging": ("bn128", lambda d: C.logging(d)),
        "poseidon2": ("bn128", lambda d: C.poseidon(d, 2)),
        "int_div32": ("bn128", lambda d: C.int_div(d, 32)),
        "int_div_arr32": ("bn128", lambda d: C.int_div_array(d, 32, "all")),     # `var qr[3] = f(a, b);`: one call, three results
        "ecdsa_calls_2x5": ("bn128", lambda d: C.ecdsa_scale(d, 2, 5, hints="functions")),
        "gcd32": ("bn128", lambda d: C.gcd_circuit(d, 32)),                        # functions calling functions
        "ecdsa_scale_2x5": ("bn128", lambda d: C.ecdsa_scale(d, 2, 5)),
        "sha256compression": ("bn128", lambda d: C.sha256_compression(d)),
        "sha256_64_bls": ("bls12381", lambda d: C.sha256(d, 64)),
        "ecdsa_scale_8x132": ("bn128", lambda d: C.ecdsa_scale(d, 8, 132)),
        "sha256_512_bls": ("bls12381", lambda d: C.sha256(d, 512)),   # BASELINE.json configs[4]
        # the bench circuit with function-computed hints (circom-ecdsa style)
        "ecdsa_scale_calls_8x132": ("bn128", lambda d: C.ecdsa_scale(d, 8, 132, hints="functions")),
    }


# calculators whose generated C++ takes g++ -O3 more than ten minutes each: built only when asked for by name
# (tests/golden/make_golden.py), never by the default build
SLOW = ("sha256compression", "sha256_64_bls", "sha256_512_bls")


def make_desc(name: str):
    from circom_b200.circuit import CircuitDesc
    prime, mk = circuits()[name]
    d = CircuitDesc(prime)
    d.set_main(mk(d), name)
    return d


def calc_path(name: str) -> str:
    return os.path.join(CALC_DIR, name)


def build(names=None, force: bool = False):
    from . import build_ref
    from .emit_ref_cpp import build_reference_calculator
    if not build_ref.have_reference():
        return []
    built = []
    for name in (names or [n for n in circuits() if n not in SLOW]):
        p = calc_path(name)
        if not force and os.path.exists(p) and os.path.exists(p + ".dat"):
            continue
        build_reference_calculator(make_desc(name), CALC_DIR, name)
        built.append(name)
    return built


if __name__ == "__main__":
    print("built:", build([a for a in sys.argv[1:] if a != "--force"] or None, force="--force" in sys.argv))
