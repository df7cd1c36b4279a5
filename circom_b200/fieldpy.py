"""Compile-time field arithmetic for constant folding in the circuit DSL.

Plays the role circom_algebra/src/modular_arithmetic.rs:26-215 plays inside the
reference compiler (folding operators whose operands are known at compile
time).  Plain Python integers; not a throughput path.
"""
from __future__ import annotations

from .circuit import OPS


def _val(q, x):
    return x - q if x > (q >> 1) else x


def _wrap(q, x):
    x &= (1 << q.bit_length()) - 1
    return x - q if x >= q else x


def _shl(q, a, k):
    n = ((q.bit_length() + 63) // 64) * 64
    return _wrap(q, (a << k) & ((1 << n) - 1))


def apply(q: int, op: int, a: int, b: int = 0) -> int:
    qb = q.bit_length()
    n = ((qb + 63) // 64) * 64
    if op == OPS["MUL"]: return a * b % q
    if op == OPS["ADD"]: return (a + b) % q
    if op == OPS["SUB"]: return (a - b) % q
    if op == OPS["NEG"]: return (-a) % q
    if op == OPS["DIV"]: return a * (pow(b, -1, q) if b else 0) % q
    if op == OPS["POW"]: return pow(a, b, q)
    if op == OPS["IDIV"]: return a // b
    if op == OPS["MOD"]: return a % b
    if op == OPS["SHL"]:
        if b < qb: return _shl(q, a, b)
        if b > q - qb: return a >> (q - b)
        return 0
    if op == OPS["SHR"]:
        if b < qb: return a >> b
        if b > q - qb: return _shl(q, a, q - b)
        return 0
    if op == OPS["LT"]: return int(_val(q, a) < _val(q, b))
    if op == OPS["GT"]: return int(_val(q, a) > _val(q, b))
    if op == OPS["LEQ"]: return int(_val(q, a) <= _val(q, b))
    if op == OPS["GEQ"]: return int(_val(q, a) >= _val(q, b))
    if op == OPS["EQ"]: return int(a == b)
    if op == OPS["NEQ"]: return int(a != b)
    if op == OPS["LOR"]: return int(bool(a) or bool(b))
    if op == OPS["LAND"]: return int(bool(a) and bool(b))
    if op == OPS["LNOT"]: return int(not a)
    if op == OPS["BAND"]: return _wrap(q, a & b)
    if op == OPS["BOR"]: return _wrap(q, a | b)
    if op == OPS["BXOR"]: return _wrap(q, a ^ b)
    if op == OPS["BNOT"]: return _wrap(q, ~a & ((1 << n) - 1))
    if op == OPS["COPY"]: return a
    raise ValueError("cannot fold op %d" % op)
