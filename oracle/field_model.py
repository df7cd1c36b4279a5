"""Python big-int model of circom's runtime field operators (TEST INFRASTRUCTURE).

CPU restatement of the *value-level* semantics of the reference's `Fr_*`
runtime (code_producers/src/c_elements/generic/fr.cpp), cross-checked
against the operator spec in circom_algebra/src/modular_arithmetic.rs:26-215
and mkdocs/docs/circom-language/basic-operators.md:40-118.  Every witness value
the reference writes goes through Fr_toLongNormal (common/main.cpp:330), so
only the canonical integer in [0,q) matters; this model works on canonical
integers.  tests/test_oracle_ref.py pins it against the compiled reference
(oracle/_ref/libfr_<prime>.so) over all operand representations.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import
this module.
"""
from __future__ import annotations

# program_structure/src/utils/constants.rs:3-6
PRIMES = {
    "bn128": 21888242871839275222246405745257275088548364400416034343698204186575808495617,
    "bls12381": 52435875175126190479447740508185965837690552500527637822603658699938581184513,
    "grumpkin": 21888242871839275222246405745257275088696311157297823662689037894645226208583,
    "pallas": 28948022309329048855892746252171976963363056481941560715954676764349967630337,
    "vesta": 28948022309329048855892746252171976963363056481941647379679742748393362948097,
    "secq256r1": 115792089210356248762697446949407573530086143415290314195533631308867097853951,
    "bls12377": 8444461749428370424248824938781546531375899335154063827935233455917409239041,
    "goldilocks": 18446744069414584321,
}

# IR opcodes: OperatorType of compiler/src/intermediate_representation/compute_bucket.rs:7-34
# plus the moves / control the other buckets express.
OPS = {
    "NOP": 0, "MUL": 1, "DIV": 2, "ADD": 3, "SUB": 4, "POW": 5, "IDIV": 6, "MOD": 7,
    "SHL": 8, "SHR": 9, "LEQ": 10, "GEQ": 11, "LT": 12, "GT": 13, "EQ": 14, "NEQ": 15,
    "LOR": 16, "LAND": 17, "LNOT": 18, "BOR": 19, "BAND": 20, "BXOR": 21, "BNOT": 22,
    "NEG": 23, "COPY": 24, "SELECT": 25, "ASSERT": 26, "ASSERT_EQ": 27,
    "JMP": 40, "JZ": 41, "RET": 42, "LOADX": 43, "STOREX": 44, "CALL": 45, "ARG": 46,
}
OP_NAMES = {v: k for k, v in OPS.items()}


class DivisionByZero(Exception):
    """Fr_idiv / Fr_mod by zero: GMP aborts the reference process (generic/fr.cpp:2835-2875)."""


class Field:
    def __init__(self, prime: str | int):
        self.q = PRIMES[prime] if isinstance(prime, str) else int(prime)
        self.name = prime if isinstance(prime, str) else "custom"
        self.qbits = self.q.bit_length()
        self.half = self.q >> 1          # generic/fr.cpp:9  (q-1)/2 == q>>1 for odd q
        self.mask = (1 << self.qbits) - 1  # lboMask on the top limb (generic/fr.cpp:16)
        self.n64 = (self.qbits + 63) // 64
        self.R = 1 << (64 * self.n64)

    # --- helpers -----------------------------------------------------------
    def val(self, x: int) -> int:
        """signed view used by comparisons (generic/fr.cpp:1184-1218; modular_arithmetic.rs:154-164)"""
        return x - self.q if x > self.half else x

    def _wrap_bits(self, x: int) -> int:
        """top-limb mask then one conditional subtraction (generic/fr.cpp:293-327)"""
        x &= self.mask
        return x - self.q if x >= self.q else x

    # --- arithmetic (generic/fr.cpp:19-86,110-164) ---------------------------
    def add(self, a, b): return (a + b) % self.q
    def sub(self, a, b): return (a - b) % self.q
    def mul(self, a, b): return (a * b) % self.q
    def neg(self, a): return (-a) % self.q

    def inv(self, a):
        # mpz_invert's failure on 0 is ignored by the reference (generic/fr.cpp:2895-2906);
        # with GMP 6.3.0 the result element is 0 (SURVEY.md Appendix D probe; pinned in tests).
        return pow(a, -1, self.q) if a % self.q else 0

    def div(self, a, b): return (a * self.inv(b)) % self.q   # generic/fr.cpp:2908-2912

    def pow(self, a, b): return pow(a, b, self.q)             # generic/fr.cpp:2877-2893

    def idiv(self, a, b):                                     # generic/fr.cpp:2835-2857
        if b == 0:
            raise DivisionByZero()
        return a // b

    def mod(self, a, b):                                      # generic/fr.cpp:2859-2875
        if b == 0:
            raise DivisionByZero()
        return a % b

    # --- shifts (generic/fr.cpp:329-364,1995-2027,2157-2307) -----------------
    def _shl_raw(self, a, k): return self._wrap_bits((a << k) & (self.R - 1))
    def _shr_raw(self, a, k): return a >> k

    def shl(self, a, b):
        if b < self.qbits:
            return self._shl_raw(a, b)
        if b > self.q - self.qbits:            # "negative" shift amount -j, j < qbits
            return self._shr_raw(a, self.q - b)
        return 0

    def shr(self, a, b):
        if b < self.qbits:
            return self._shr_raw(a, b)
        if b > self.q - self.qbits:
            return self._shl_raw(a, self.q - b)
        return 0

    # --- comparisons (generic/fr.cpp:1294-1363,1469-1537,1595-1765) ----------
    def lt(self, a, b): return int(self.val(a) < self.val(b))
    def gt(self, a, b): return int(self.val(a) > self.val(b))
    def leq(self, a, b): return int(self.val(a) <= self.val(b))
    def geq(self, a, b): return int(self.val(a) >= self.val(b))
    def eq(self, a, b): return int(a == b)
    def neq(self, a, b): return int(a != b)

    # --- boolean (generic/fr.cpp:1540-1580,1771-1797) ------------------------
    def lor(self, a, b): return int(a != 0 or b != 0)
    def land(self, a, b): return int(a != 0 and b != 0)
    def lnot(self, a): return int(a == 0)

    # --- bitwise (generic/fr.cpp:293-327,366-376,1938-2755) ------------------
    def band(self, a, b): return self._wrap_bits(a & b)
    def bor(self, a, b): return self._wrap_bits(a | b)
    def bxor(self, a, b): return self._wrap_bits(a ^ b)
    def bnot(self, a): return self._wrap_bits(~a & (self.R - 1))

    # --- dispatch by IR opcode ----------------------------------------------
    def apply(self, op: int, a: int = 0, b: int = 0, c: int = 0) -> int:
        f = _DISPATCH[op]
        return f(self, a, b, c)


_DISPATCH = {
    OPS["NOP"]: lambda F, a, b, c: 0,
    OPS["MUL"]: lambda F, a, b, c: F.mul(a, b),
    OPS["DIV"]: lambda F, a, b, c: F.div(a, b),
    OPS["ADD"]: lambda F, a, b, c: F.add(a, b),
    OPS["SUB"]: lambda F, a, b, c: F.sub(a, b),
    OPS["POW"]: lambda F, a, b, c: F.pow(a, b),
    OPS["IDIV"]: lambda F, a, b, c: F.idiv(a, b),
    OPS["MOD"]: lambda F, a, b, c: F.mod(a, b),
    OPS["SHL"]: lambda F, a, b, c: F.shl(a, b),
    OPS["SHR"]: lambda F, a, b, c: F.shr(a, b),
    OPS["LEQ"]: lambda F, a, b, c: F.leq(a, b),
    OPS["GEQ"]: lambda F, a, b, c: F.geq(a, b),
    OPS["LT"]: lambda F, a, b, c: F.lt(a, b),
    OPS["GT"]: lambda F, a, b, c: F.gt(a, b),
    OPS["EQ"]: lambda F, a, b, c: F.eq(a, b),
    OPS["NEQ"]: lambda F, a, b, c: F.neq(a, b),
    OPS["LOR"]: lambda F, a, b, c: F.lor(a, b),
    OPS["LAND"]: lambda F, a, b, c: F.land(a, b),
    OPS["LNOT"]: lambda F, a, b, c: F.lnot(a),
    OPS["BOR"]: lambda F, a, b, c: F.bor(a, b),
    OPS["BAND"]: lambda F, a, b, c: F.band(a, b),
    OPS["BXOR"]: lambda F, a, b, c: F.bxor(a, b),
    OPS["BNOT"]: lambda F, a, b, c: F.bnot(a),
    OPS["NEG"]: lambda F, a, b, c: F.neg(a),
    OPS["COPY"]: lambda F, a, b, c: a,
    # `c ? a : b` is a BranchBucket on Fr_isTrue(c) (branch_bucket.rs:100-122)
    OPS["SELECT"]: lambda F, a, b, c: a if c != 0 else b,
}
