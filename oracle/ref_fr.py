"""ctypes access to the compiled REFERENCE field library oracle/_ref/libfr_<prime>.so
(TEST INFRASTRUCTURE).  The library is the reference's own generic/fr.cpp
(rendered by oracle/build_ref.py); symbols are C++-mangled because
generic/fr.hpp declares them without extern "C".
"""
from __future__ import annotations

import ctypes
import os

from .field_model import PRIMES, OPS

HERE = os.path.dirname(os.path.abspath(__file__))

Fr_SHORT = 0x00000000
Fr_SHORTMONTGOMERY = 0x40000000
Fr_LONG = 0x80000000
Fr_LONGMONTGOMERY = 0xC0000000


class FrElement(ctypes.Structure):
    _pack_ = 1
    _fields_ = [("shortVal", ctypes.c_int32), ("type", ctypes.c_uint32), ("longVal", ctypes.c_uint64 * 4)]


_BIN = {
    "MUL": "mul", "DIV": "div", "ADD": "add", "SUB": "sub", "POW": "pow", "IDIV": "idiv", "MOD": "mod",
    "SHL": "shl", "SHR": "shr", "LEQ": "leq", "GEQ": "geq", "LT": "lt", "GT": "gt", "EQ": "eq", "NEQ": "neq",
    "LOR": "lor", "LAND": "land", "BOR": "bor", "BAND": "band", "BXOR": "bxor",
}
_UN = {"LNOT": "lnot", "BNOT": "bnot", "NEG": "neg"}


def _mangle(name: str, nargs: int) -> str:
    full = "Fr_" + name
    return "_Z%d%sP9FrElement%s" % (len(full), full, "S0_" * (nargs - 1))


class RefFr:
    """Thin handle on the reference library for one prime."""

    def __init__(self, prime: str):
        path = os.path.join(HERE, "_ref", "libfr_%s.so" % prime)
        if not os.path.exists(path):
            from . import build_ref
            build_ref.build_prime(prime)
        self.lib = ctypes.CDLL(path)
        self.q = PRIMES[prime]
        self.R = 1 << 256
        self._fn = {}
        for op, n in _BIN.items():
            self._fn[OPS[op]] = (getattr(self.lib, _mangle(n, 3)), 3)
        for op, n in _UN.items():
            self._fn[OPS[op]] = (getattr(self.lib, _mangle(n, 2)), 2)
        self._toLongNormal = getattr(self.lib, _mangle("toLongNormal", 2))
        self._toMontgomery = getattr(self.lib, _mangle("toMontgomery", 2))
        self._isTrue = getattr(self.lib, "_Z9Fr_isTrueP9FrElement")
        self._isTrue.restype = ctypes.c_int
        self._toInt = getattr(self.lib, "_Z8Fr_toIntP9FrElement")
        self._toInt.restype = ctypes.c_int
        self._str2element = getattr(self.lib, "_Z14Fr_str2elementP9FrElementPKcj")

    # representations ---------------------------------------------------------
    def make(self, v: int, rep: str = "long") -> FrElement:
        """rep: 'long' (canonical), 'mont' (long Montgomery), 'short' (int32, v given signed)"""
        e = FrElement()
        if rep == "short":
            assert -(1 << 31) <= v < (1 << 31)
            e.shortVal = v
            e.type = Fr_SHORT
            return e
        v %= self.q
        if rep == "mont":
            v = (v * self.R) % self.q
            e.type = Fr_LONGMONTGOMERY
        else:
            e.type = Fr_LONG
        for i in range(4):
            e.longVal[i] = (v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF
        return e

    def canon(self, e: FrElement) -> int:
        out = FrElement()
        self._toLongNormal(ctypes.byref(out), ctypes.byref(e))
        return sum(int(out.longVal[i]) << (64 * i) for i in range(4))

    def apply(self, op: int, a: FrElement, b: FrElement | None = None) -> int:
        fn, n = self._fn[op]
        r = FrElement()
        if n == 3:
            fn(ctypes.byref(r), ctypes.byref(a), ctypes.byref(b))
        else:
            fn(ctypes.byref(r), ctypes.byref(a))
        return self.canon(r)

    def is_true(self, a: FrElement) -> int:
        return int(self._isTrue(ctypes.byref(a)))

    def str2element(self, s: str, base: int = 10) -> int:
        e = FrElement()
        self._str2element(ctypes.byref(e), s.encode(), ctypes.c_uint(base))
        return self.canon(e)
