"""Circuit description: the information circom's compiler hands to a code producer.

The Rust compiler cannot run in this environment, so circuits are authored in a
small Python DSL that plays the role of `compiler/` up to the point where the
reference calls a code producer (compiler/src/circuit_design/circuit.rs:596
`Circuit::produce_c`): a list of *template instances* (one per distinct
template + parameters, like `TemplateCodeInfo`, circuit_design/template.rs:12-31),
each with

  * its own signals numbered outputs, inputs, intermediates
    (constraint_generation/src/execution_data/executed_template.rs:442-552),
  * its sub-components in creation order (CreateCmpBucket),
  * a straight-line body of field operations over own signals, sub-component
    signals, constants and temporaries (Compute / Load / Store buckets with all
    compile-time-bounded loops unrolled), and
  * its R1CS constraints  A*B - C = 0  (circom_algebra/src/algebra.rs:113-145).

`CircuitDesc.save()` writes the binary ".cb2c" description that the CUDA
back end (circom_b200/csrc/flatten.cpp) lowers to the flat, levelised
instruction tape, and that oracle/ consumes independently.
"""
from __future__ import annotations

import struct
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

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
# the primes of program_structure/src/utils/constants.rs:3-13 (goldilocks: 64-bit values in the same 32-byte elements)
PRIME_IDS = {"bn128": 0, "bls12381": 1, "grumpkin": 2, "pallas": 3, "vesta": 4, "secq256r1": 5, "bls12377": 6, "goldilocks": 7}

# OperatorType (compiler/src/intermediate_representation/compute_bucket.rs:7-34) + moves
OPS = {
    "NOP": 0, "MUL": 1, "DIV": 2, "ADD": 3, "SUB": 4, "POW": 5, "IDIV": 6, "MOD": 7,
    "SHL": 8, "SHR": 9, "LEQ": 10, "GEQ": 11, "LT": 12, "GT": 13, "EQ": 14, "NEQ": 15,
    "LOR": 16, "LAND": 17, "LNOT": 18, "BOR": 19, "BAND": 20, "BXOR": 21, "BNOT": 22,
    "NEG": 23, "COPY": 24, "SELECT": 25, "ASSERT": 26, "ASSERT_EQ": 27,
    # one argument of a log() call (LogBucket, log_bucket.rs:104-162): a = the value (a signal, a constant) or NONE for a
    # string, b = (NONE, string id), c = (NONE, 1) on the last argument of the call
    "LOG": 29,
    # producer-level only (never in a .cb2c): d <- own signal array element a[toInt(b)], c = (NONE, extent) - Template.load_indexed
    "LOADSIG": 48,
    # function bodies (FunctionCodeInfo, compiler/src/circuit_design/function.rs) and their call sites
    "JMP": 40, "JZ": 41, "RET": 42, "LOADX": 43, "STOREX": 44, "CALL": 45, "ARG": 46,
}
OP_NAMES = {v: k for k, v in OPS.items()}

# reference kinds
K_NONE, K_OWN, K_SUB, K_CONST, K_TMP, K_ONE = 0, 1, 2, 3, 4, 5
Ref = Tuple[int, int, int]  # (kind, sub, idx)
NONE_REF: Ref = (K_NONE, 0, 0)
ONE_REF: Ref = (K_ONE, 0, 0)


def pack_ref(r: Ref) -> int:
    return (r[0] << 56) | (r[1] << 32) | r[2]


class Expr:
    """A value inside a template body: where it lives (`ref`) plus, when it is at
    most quadratic in the signals, its symbolic form (the role of
    circom_algebra::ArithmeticExpression, algebra.rs:9)."""

    __slots__ = ("t", "_ref", "terms", "lin", "quad", "const")

    def __init__(self, t: "Template", ref: Optional[Ref], lin=None, quad=None, const=None, terms=None):
        self.t = t
        self._ref = ref
        self.terms = terms  # pending n-ary sum: list of (sign, Expr); materialised on first use
        self.lin = lin      # dict key->coeff ; key = Ref of a signal or ONE_REF
        self.quad = quad    # (A, B, C) dicts
        self.const = const  # int if compile-time constant

    @property
    def ref(self) -> Ref:
        """Where the value lives.  A pending sum is emitted here as a balanced tree of ADD/SUB:
        field addition is associative and commutative, so re-associating the accumulation chains
        circom programs write (`lc += x*2**k`) changes no value but shortens the dependency
        depth of the tape from n to log2(n)."""
        if self._ref is None:
            self._ref = self.t._materialise(self.terms)
            self.terms = None
        return self._ref

    # -- coercion -------------------------------------------------------------
    def _co(self, o) -> "Expr":
        return o if isinstance(o, Expr) else self.t.const(o)

    # -- arithmetic -----------------------------------------------------------
    def __add__(self, o): return self.t._add(self, self._co(o))
    def __radd__(self, o): return self.t._add(self._co(o), self)
    def __sub__(self, o): return self.t._sub(self, self._co(o))
    def __rsub__(self, o): return self.t._sub(self._co(o), self)
    def __mul__(self, o): return self.t._mul(self, self._co(o))
    def __rmul__(self, o): return self.t._mul(self._co(o), self)
    def __neg__(self): return self.t._neg(self)
    def __truediv__(self, o): return self.t._bin("DIV", self, self._co(o))
    def __floordiv__(self, o): return self.t._bin("IDIV", self, self._co(o))
    def __mod__(self, o): return self.t._bin("MOD", self, self._co(o))
    def __pow__(self, o): return self.t._bin("POW", self, self._co(o))
    def __lshift__(self, o): return self.t._bin("SHL", self, self._co(o))
    def __rshift__(self, o): return self.t._bin("SHR", self, self._co(o))
    def __and__(self, o): return self.t._bin("BAND", self, self._co(o))
    def __or__(self, o): return self.t._bin("BOR", self, self._co(o))
    def __xor__(self, o): return self.t._bin("BXOR", self, self._co(o))
    def __invert__(self): return self.t._un("BNOT", self)
    # comparisons are explicit methods (python's == must stay identity)
    def lt(self, o): return self.t._bin("LT", self, self._co(o))
    def gt(self, o): return self.t._bin("GT", self, self._co(o))
    def leq(self, o): return self.t._bin("LEQ", self, self._co(o))
    def geq(self, o): return self.t._bin("GEQ", self, self._co(o))
    def eq(self, o): return self.t._bin("EQ", self, self._co(o))
    def neq(self, o): return self.t._bin("NEQ", self, self._co(o))
    def land(self, o): return self.t._bin("LAND", self, self._co(o))
    def lor(self, o): return self.t._bin("LOR", self, self._co(o))
    def lnot(self): return self.t._un("LNOT", self)


def _lin_add(a: dict, b: dict, q: int, sign: int = 1) -> dict:
    r = dict(a)
    for k, v in b.items():
        nv = (r.get(k, 0) + sign * v) % q
        if nv:
            r[k] = nv
        else:
            r.pop(k, None)
    return r


def _lin_scale(a: dict, c: int, q: int) -> dict:
    c %= q
    if c == 0:
        return {}
    return {k: (v * c) % q for k, v in a.items()}


class Sub:
    """Handle on a sub-component inside a template body."""

    def __init__(self, parent: "Template", index: int, tmpl: "Template", name: str):
        self.parent, self.index, self.tmpl, self.name = parent, index, tmpl, name

    def sig(self, name: str, i: Optional[int] = None):
        off, size, cat = self.tmpl.sig_info[name]
        if i is None and size == 1 and not self.tmpl.sig_is_array[name]:
            return self._mk(name, 0)
        if i is None:
            return [self._mk(name, j) for j in range(size)]
        return self._mk(name, i)

    def _mk(self, name: str, j: int) -> Expr:
        ref = (K_SUB, self.index, ("name", name, j))  # resolved at finalize
        return Expr(self.parent, ref, lin={ref: 1})

    def __getitem__(self, key):
        if isinstance(key, tuple):
            return self.sig(key[0], key[1])
        return self.sig(key)


class Template:
    """One template instance (template name + concrete parameters)."""

    def __init__(self, desc: "CircuitDesc", name: str):
        self.desc = desc
        self.q = desc.q
        self.name = name
        self.sigs: Dict[str, list] = {"out": [], "in": [], "inter": []}  # (name, size)
        self.sig_info: Dict[str, Tuple[int, int, str]] = {}
        self.sig_is_array: Dict[str, bool] = {}
        self.subs: List[Sub] = []
        self.n_tmp = 0
        self.ops: List[Tuple[int, Ref, Ref, Ref, Ref]] = []
        self.constraints: List[Tuple[dict, dict, dict]] = []
        self.finalized = False
        self.id = -1

    # -- declarations -----------------------------------------------------------
    def _decl(self, cat: str, name: str, size: Optional[int]):
        assert not self.finalized and name not in self.sig_info
        n = 1 if size is None else size
        self.sigs[cat].append((name, n))
        self.sig_info[name] = (-1, n, cat)
        self.sig_is_array[name] = size is not None
        mk = lambda j: Expr(self, (K_OWN, 0, ("name", name, j)), lin={(K_OWN, 0, ("name", name, j)): 1})
        return mk(0) if size is None else [mk(j) for j in range(n)]

    def output(self, name, size=None): return self._decl("out", name, size)
    def input(self, name, size=None): return self._decl("in", name, size)
    def signal(self, name, size=None): return self._decl("inter", name, size)

    def component(self, name: str, tmpl: "Template") -> Sub:
        assert tmpl.finalized, "instantiate finished templates only"
        s = Sub(self, len(self.subs), tmpl, name)
        self.subs.append(s)
        return s

    # -- values -------------------------------------------------------------------
    def const(self, v: int) -> Expr:
        v = int(v) % self.q
        cid = self.desc.const_id(v)
        return Expr(self, (K_CONST, 0, cid), lin=({ONE_REF: v} if v else {}), const=v)

    def _tmp(self) -> Ref:
        self.n_tmp += 1
        return (K_TMP, 0, self.n_tmp - 1)

    def _emit(self, op: str, a: Expr, b: Optional[Expr] = None, c: Optional[Expr] = None) -> Ref:
        dst = self._tmp()
        self.ops.append((OPS[op], dst, a.ref, b.ref if b is not None else NONE_REF,
                         c.ref if c is not None else NONE_REF))
        return dst

    def _fold(self, op: str, a: Expr, b: Optional[Expr]) -> Optional[Expr]:
        if a.const is not None and (b is None or b.const is not None):
            from . import fieldpy
            v = fieldpy.apply(self.q, OPS[op], a.const, b.const if b is not None else 0)
            return self.const(v)
        return None

    @staticmethod
    def _terms_of(e: Expr, sign: int):
        if e._ref is None and e.terms is not None:
            return [(sg * sign, x) for sg, x in e.terms]
        return [(sign, e)]

    def _materialise(self, terms) -> Ref:
        pos = [x for sg, x in terms if sg > 0]
        neg = [x for sg, x in terms if sg < 0]

        def tree(xs):
            xs = list(xs)
            while len(xs) > 1:
                nxt = []
                for i in range(0, len(xs) - 1, 2):
                    nxt.append(Expr(self, self._emit("ADD", xs[i], xs[i + 1])))
                if len(xs) & 1:
                    nxt.append(xs[-1])
                xs = nxt
            return xs[0]
        if pos and neg:
            return self._emit("SUB", tree(pos), tree(neg))
        if pos:
            p = tree(pos)
            return p.ref
        return self._emit("NEG", tree(neg))

    def _addsub(self, a: Expr, b: Expr, sign: int) -> Expr:
        f = self._fold("ADD" if sign > 0 else "SUB", a, b)
        if f: return f
        if a.const == 0 and sign > 0:
            return b
        if b.const == 0:
            return a
        r = Expr(self, None, terms=self._terms_of(a, 1) + self._terms_of(b, sign))
        q = self.q
        if a.lin is not None and b.lin is not None:
            r.lin = _lin_add(a.lin, b.lin, q, sign)
        elif a.quad is not None and b.lin is not None:
            r.quad = (a.quad[0], a.quad[1], _lin_add(a.quad[2], b.lin, q, sign))
        elif b.quad is not None and a.lin is not None:
            if sign > 0:
                r.quad = (b.quad[0], b.quad[1], _lin_add(b.quad[2], a.lin, q))
            else:
                r.quad = (_lin_scale(b.quad[0], q - 1, q), b.quad[1], _lin_add(a.lin, b.quad[2], q, -1))
        return r

    def _add(self, a: Expr, b: Expr) -> Expr:
        return self._addsub(a, b, 1)

    def _sub(self, a: Expr, b: Expr) -> Expr:
        return self._addsub(a, b, -1)

    def _mul(self, a: Expr, b: Expr) -> Expr:
        f = self._fold("MUL", a, b)
        if f: return f
        r = Expr(self, self._emit("MUL", a, b))
        if a.const is not None or b.const is not None:
            c, x = (a, b) if a.const is not None else (b, a)
            if x.lin is not None:
                r.lin = _lin_scale(x.lin, c.const, self.q)
            elif x.quad is not None:
                r.quad = (_lin_scale(x.quad[0], c.const, self.q), x.quad[1],
                          _lin_scale(x.quad[2], c.const, self.q))
        elif a.lin is not None and b.lin is not None:
            r.quad = (a.lin, b.lin, {})
        return r

    def _neg(self, a: Expr) -> Expr:
        f = self._fold("NEG", a, None)
        if f: return f
        r = Expr(self, self._emit("NEG", a))
        if a.lin is not None:
            r.lin = _lin_scale(a.lin, self.q - 1, self.q)
        elif a.quad is not None:
            r.quad = (_lin_scale(a.quad[0], self.q - 1, self.q), a.quad[1],
                      _lin_scale(a.quad[2], self.q - 1, self.q))
        return r

    def _bin(self, op: str, a: Expr, b: Expr) -> Expr:
        f = self._fold(op, a, b)
        if f: return f
        return Expr(self, self._emit(op, a, b))

    def _un(self, op: str, a: Expr) -> Expr:
        f = self._fold(op, a, None)
        if f: return f
        return Expr(self, self._emit(op, a))

    def select(self, cond: Expr, a, b) -> Expr:
        """`cond ? a : b` (BranchBucket on Fr_isTrue, branch_bucket.rs:100-122); both arms
        are evaluated, so they must be total (inv(0)=0 is)."""
        a = a if isinstance(a, Expr) else self.const(a)
        b = b if isinstance(b, Expr) else self.const(b)
        return Expr(self, self._emit("SELECT", a, b, cond))

    def log(self, *args) -> None:
        """`log(a, "text", b, ...)`: the reference prints the arguments separated by blanks, values as canonical decimals, and a
        newline (log_bucket.rs:104-162).  Arguments: strings, constants, signals (own or of a sub-component); the value of a
        logged signal is read from the witness after the run (cw_batch_log), so expressions have to be logged through the
        signal that holds them."""
        assert args, "log() needs an argument"
        for k, a in enumerate(args):
            last = (K_NONE, 0, 1 if k == len(args) - 1 else 0)
            if isinstance(a, str):
                assert a and all(0x20 <= ord(ch) < 0x7F and ch not in '%\\"' for ch in a), "log string: printable ASCII without % \\ \""
                self.ops.append((OPS["LOG"], NONE_REF, NONE_REF, (K_NONE, 0, self.desc.string_id(a)), last))
            else:
                e = a if isinstance(a, Expr) else self.const(a)
                assert e._ref is not None and e._ref[0] in (K_OWN, K_SUB, K_CONST, K_ONE), "log: a signal or a constant"
                self.ops.append((OPS["LOG"], NONE_REF, e.ref, NONE_REF, last))

    def load_indexed(self, arr: Sequence[Expr], idx: Expr) -> Expr:
        """`arr[idx]` in a `<--` expression with a SIGNAL-dependent index - the reference compiler prints a load whose address
        is computed at run time, `&signalValues[mySignalStart + base + Fr_toInt(idx)]` (LoadBucket with an Indexed location
        over ToAddress, load_bucket.rs:325-447, compute_bucket.rs:361-363).  Tape operands are static, so a producer for this
        back end expands the access over the array's extent: n comparisons, n selections, two adder trees, and an assert that
        the index hit an element (the reference reads whatever lies beside the array there).  `arr` is a declared signal array
        of this template.  The expansion is written at serialisation (Template.expanded_ops); evaluators of the description
        and the reference-calculator emitter see the access itself."""
        assert arr and all(e._ref is not None and e._ref[0] == K_OWN for e in arr), "load_indexed: an own signal array"
        name = arr[0]._ref[2][1]
        assert [e._ref[2] for e in arr] == [("name", name, j) for j in range(len(arr))] and self.sig_info[name][1] == len(arr)
        dst = self._tmp()
        self.ops.append((OPS["LOADSIG"], dst, arr[0].ref, idx.ref, (K_NONE, 0, len(arr))))
        return Expr(self, dst)

    def expanded_ops(self):
        """(ops as a `.cb2c` carries them, n_tmp): LOADSIG replaced by primitive operations on fresh temporaries"""
        if not any(op == OPS["LOADSIG"] for op, *_ in self.ops):
            return self.ops, self.n_tmp
        out, n_tmp = [], self.n_tmp
        cref = lambda v: (K_CONST, 0, self.desc.const_id(v))

        def tmp():
            nonlocal n_tmp
            n_tmp += 1
            return (K_TMP, 0, n_tmp - 1)

        def tree(refs, last=None):
            refs = list(refs)
            while len(refs) > 1:
                nxt = []
                for i in range(0, len(refs) - 1, 2):
                    d = last if (last is not None and len(refs) == 2) else tmp()
                    out.append((OPS["ADD"], d, refs[i], refs[i + 1], NONE_REF))
                    nxt.append(d)
                if len(refs) & 1:
                    nxt.append(refs[-1])
                refs = nxt
            return refs[0]
        for op, d, a, b, c in self.ops:
            if op != OPS["LOADSIG"]:
                out.append((op, d, a, b, c))
                continue
            n, base = c[2], a[2]
            eqs, sels = [], []
            for i in range(n):
                e = tmp()
                out.append((OPS["EQ"], e, b, cref(i), NONE_REF))
                eqs.append(e)
            for i in range(n):
                sl = d if n == 1 else tmp()
                out.append((OPS["SELECT"], sl, (K_OWN, 0, base + i), cref(0), eqs[i]))
                sels.append(sl)
            if n > 1:
                tree(sels, last=d)
            out.append((OPS["ASSERT"], NONE_REF, tree(eqs), NONE_REF, NONE_REF))
        return out, n_tmp

    def call(self, fn: "Function", args: Sequence) -> Expr:
        """`f(args...)` inside an expression (CallBucket, call_bucket.rs:466-533): the arguments are copied to
        the callee's variables, the body runs with its own data-dependent loops and branches, one field
        element comes back.  Functions returning arrays are called once per element (extra index argument)."""
        assert len(args) == fn.n_params, "wrong number of arguments for %s" % fn.name
        exprs = [a if isinstance(a, Expr) else self.const(a) for a in args]
        for e in exprs:
            self.ops.append((OPS["ARG"], NONE_REF, e.ref, NONE_REF, NONE_REF))
        dst = self._tmp()
        self.ops.append((OPS["CALL"], dst, (K_NONE, 0, fn.id), (K_NONE, 0, len(exprs)), NONE_REF))
        return Expr(self, dst)

    def call_array(self, fn: "Function", args: Sequence, n: int) -> List[Expr]:
        """`var r[n] = f(args...)`: ONE call whose n results land in n consecutive temporaries (the reference's
        CallBucket with a destination of size n, call_bucket.rs:466-533)"""
        assert len(args) == fn.n_params and 1 <= n <= fn.n_results, "bad array call of %s" % fn.name
        exprs = [a if isinstance(a, Expr) else self.const(a) for a in args]
        for e in exprs:
            self.ops.append((OPS["ARG"], NONE_REF, e.ref, NONE_REF, NONE_REF))
        dsts = [self._tmp() for _ in range(n)]
        self.ops.append((OPS["CALL"], dsts[0], (K_NONE, 0, fn.id), (K_NONE, 0, len(exprs)), (K_NONE, 0, n)))
        return [Expr(self, d) for d in dsts]

    # -- statements -------------------------------------------------------------------
    def assign(self, dst: Expr, src) -> None:
        """`dst <-- src` (StoreBucket, store_bucket.rs:607-646)."""
        src = src if isinstance(src, Expr) else self.const(src)
        assert dst.ref[0] in (K_OWN, K_SUB)
        self.ops.append((OPS["COPY"], dst.ref, src.ref, NONE_REF, NONE_REF))

    def constrain(self, lhs, rhs, emit_assert: bool = True) -> None:
        """`lhs === rhs`: records the R1CS constraint and (sanity_check >= 1,
        assert_bucket.rs:70-88) the run-time equality assert."""
        lhs = lhs if isinstance(lhs, Expr) else self.const(lhs)
        rhs = rhs if isinstance(rhs, Expr) else self.const(rhs)
        self._record_constraint(lhs, rhs)
        if emit_assert:
            self.ops.append((OPS["ASSERT_EQ"], NONE_REF, lhs.ref, rhs.ref, NONE_REF))

    def _record_constraint(self, lhs: Expr, rhs: Expr) -> None:
        q = self.q
        # The reference forms E = lhs - rhs (sub = add(lhs, -1 * rhs); a constant times a quadratic a*b + c scales
        # a and c, algebra.rs:387-395,441-450) and stores A = E.a, B = E.b, C = -E.c
        # (transform_expression_to_constraint_form, algebra.rs:113-145).  Same normal form here, so that the
        # coefficients of the written .r1cs are the reference's (`out <== x*y` is (-x) * y = -out).
        if lhs.quad is not None and rhs.lin is not None:
            A, B, C = lhs.quad[0], lhs.quad[1], _lin_add(rhs.lin, lhs.quad[2], q, -1)
        elif rhs.quad is not None and lhs.lin is not None:
            A, B, C = _lin_scale(rhs.quad[0], q - 1, q), rhs.quad[1], _lin_add(rhs.quad[2], lhs.lin, q, -1)
        elif lhs.lin is not None and rhs.lin is not None:
            A, B, C = {}, {}, _lin_add(rhs.lin, lhs.lin, q, -1)
        else:
            raise ValueError("non quadratic constraint in template %s" % self.name)
        self.constraints.append((A, B, C))

    def assign_constrained(self, dst: Expr, src) -> None:
        """`dst <== src`  =  `dst <-- src; dst === src`.  The assert of a freshly stored
        value is vacuous and is not emitted as a tape op."""
        src = src if isinstance(src, Expr) else self.const(src)
        self.assign(dst, src)
        self._record_constraint(dst, src)

    # -- finalisation -------------------------------------------------------------------
    def finalize(self) -> "Template":
        assert not self.finalized
        off = 0
        for cat in ("out", "in", "inter"):
            for name, n in self.sigs[cat]:
                self.sig_info[name] = (off, n, cat)
                off += n
        self.n_out = sum(n for _, n in self.sigs["out"])
        self.n_in = sum(n for _, n in self.sigs["in"])
        self.n_inter = sum(n for _, n in self.sigs["inter"])
        self.n_own = off

        def res(r: Ref) -> Ref:
            if r[0] == K_OWN and isinstance(r[2], tuple):
                _, name, j = r[2]
                o, n, _c = self.sig_info[name]
                assert 0 <= j < n
                return (K_OWN, 0, o + j)
            if r[0] == K_SUB and isinstance(r[2], tuple):
                _, name, j = r[2]
                o, n, _c = self.subs[r[1]].tmpl.sig_info[name]
                assert 0 <= j < n
                return (K_SUB, r[1], o + j)
            return r

        self.ops = [(op, res(d), res(a), res(b), res(c)) for op, d, a, b, c in self.ops]
        self.constraints = [tuple({res(k): v for k, v in lc.items()} for lc in con)
                            for con in self.constraints]
        self.total_signals = self.n_own + sum(s.tmpl.total_signals for s in self.subs)
        self.total_components = 1 + sum(s.tmpl.total_components for s in self.subs)
        self.finalized = True
        self.id = len(self.desc.templates)
        self.desc.templates.append(self)
        return self


class FReg:
    """A variable (register) of a function body.  Arithmetic on registers emits instructions into the
    function and yields fresh registers; `fb.set(var, value)` assigns (circom `var` semantics: mutable)."""

    __slots__ = ("fb", "idx")

    def __init__(self, fb: "Function", idx: int):
        self.fb, self.idx = fb, idx

    def _b(self, op, o): return self.fb._emit_bin(op, self, o)
    def _rb(self, op, o): return self.fb._emit_bin(op, o, self)
    def __add__(self, o): return self._b("ADD", o)
    def __radd__(self, o): return self._rb("ADD", o)
    def __sub__(self, o): return self._b("SUB", o)
    def __rsub__(self, o): return self._rb("SUB", o)
    def __mul__(self, o): return self._b("MUL", o)
    def __rmul__(self, o): return self._rb("MUL", o)
    def __truediv__(self, o): return self._b("DIV", o)
    def __floordiv__(self, o): return self._b("IDIV", o)
    def __mod__(self, o): return self._b("MOD", o)
    def __pow__(self, o): return self._b("POW", o)
    def __lshift__(self, o): return self._b("SHL", o)
    def __rshift__(self, o): return self._b("SHR", o)
    def __and__(self, o): return self._b("BAND", o)
    def __or__(self, o): return self._b("BOR", o)
    def __xor__(self, o): return self._b("BXOR", o)
    def __neg__(self): return self.fb._emit_bin("NEG", self, None)
    def lt(self, o): return self._b("LT", o)
    def gt(self, o): return self._b("GT", o)
    def leq(self, o): return self._b("LEQ", o)
    def geq(self, o): return self._b("GEQ", o)
    def eq(self, o): return self._b("EQ", o)
    def neq(self, o): return self._b("NEQ", o)
    def land(self, o): return self._b("LAND", o)
    def lor(self, o): return self._b("LOR", o)
    def lnot(self): return self.fb._emit_bin("LNOT", self, None)


class Function:
    """A circom `function`: parameters and `var`s are registers, the body is a list of register
    instructions with jumps, so loops and branches may depend on run-time values (LoopBucket /
    BranchBucket with Fr_isTrue conditions, loop_bucket.rs:76-91, branch_bucket.rs:100-122; array
    variables indexed through Fr_toInt, compute_bucket.rs:361-363)."""

    def __init__(self, desc: "CircuitDesc", name: str, n_params: int):
        self.desc, self.name, self.n_params = desc, name, n_params
        self.n_regs = n_params
        self.n_results = 1
        self._arrays: List[Tuple[int, int]] = []          # (first register, length) of every `var` / parameter array
        self.code: List[Tuple[int, Ref, Ref, Ref, Ref]] = []
        self._loops: List[Tuple[int, List[int]]] = []
        self._ifs: List[List[int]] = []
        self.id = -1

    # -- registers -----------------------------------------------------------------------
    def param(self, i: int) -> FReg:
        assert 0 <= i < self.n_params
        return FReg(self, i)

    def var(self, init=0) -> FReg:
        r = FReg(self, self.n_regs)
        self.n_regs += 1
        self.set(r, init)
        return r

    def array(self, n: int, init=0) -> int:
        """n consecutive registers; returns the index of the first (for load/store with a run-time index)"""
        base = self.n_regs
        self.n_regs += n
        self._arrays.append((base, n))
        for k in range(n):
            self.set(FReg(self, base + k), init)
        return base

    def param_array(self, first: int, n: int) -> int:
        """parameters first .. first+n-1 as one array (`function f(x[n])`: array parameters are consecutive variables and
        are indexed in place); returns the base for load / store"""
        assert 0 <= first and first + n <= self.n_params
        self._arrays.append((first, n))
        return first

    def _extent(self, base: int) -> int:
        """registers from `base` to the end of the array that contains it (0: not inside a declared array)"""
        for start, n in self._arrays:
            if start <= base < start + n:
                return start + n - base
        return 0

    def _operand(self, o) -> Ref:
        if isinstance(o, FReg):
            return (K_TMP, 0, o.idx)
        return (K_CONST, 0, self.desc.const_id(int(o)))

    def _emit_bin(self, op: str, a, b) -> FReg:
        r = FReg(self, self.n_regs)
        self.n_regs += 1
        self.code.append((OPS[op], (K_TMP, 0, r.idx), self._operand(a), self._operand(b) if b is not None else NONE_REF,
                          NONE_REF))
        return r

    # -- statements ------------------------------------------------------------------------
    def set(self, dst: FReg, src) -> None:
        self.code.append((OPS["COPY"], (K_TMP, 0, dst.idx), self._operand(src), NONE_REF, NONE_REF))

    def load(self, base: int, idx: FReg) -> FReg:
        r = FReg(self, self.n_regs)
        self.n_regs += 1
        # operand c = (NONE, extent): the index must stay below it (0 = unknown: up to the last register)
        self.code.append((OPS["LOADX"], (K_TMP, 0, r.idx), (K_NONE, 0, base), (K_TMP, 0, idx.idx), (K_NONE, 0, self._extent(base))))
        return r

    def store(self, base: int, idx: FReg, src) -> None:
        self.code.append((OPS["STOREX"], (K_NONE, 0, self._extent(base)), (K_NONE, 0, base), (K_TMP, 0, idx.idx),
                          self._operand(src)))

    def loop_begin(self) -> None:
        self._loops.append((len(self.code), []))

    def loop_break_if_zero(self, cond: FReg) -> None:
        """`while (cond)`: leaves the innermost loop when cond is zero"""
        self._loops[-1][1].append(len(self.code))
        self.code.append((OPS["JZ"], NONE_REF, (K_TMP, 0, cond.idx), (K_NONE, 0, 0), NONE_REF))

    def loop_end(self) -> None:
        head, breaks = self._loops.pop()
        self.code.append((OPS["JMP"], NONE_REF, (K_NONE, 0, head), NONE_REF, NONE_REF))
        for pc in breaks:
            op, d, a, _b, c = self.code[pc]
            self.code[pc] = (op, d, a, (K_NONE, 0, len(self.code)), c)

    def if_begin(self, cond: FReg) -> None:
        self._ifs.append([len(self.code)])
        self.code.append((OPS["JZ"], NONE_REF, (K_TMP, 0, cond.idx), (K_NONE, 0, 0), NONE_REF))

    def if_else(self) -> None:
        pcs = self._ifs[-1]
        pcs.append(len(self.code))
        self.code.append((OPS["JMP"], NONE_REF, (K_NONE, 0, 0), NONE_REF, NONE_REF))
        op, d, a, _b, c = self.code[pcs[0]]
        self.code[pcs[0]] = (op, d, a, (K_NONE, 0, len(self.code)), c)

    def if_end(self) -> None:
        pcs = self._ifs.pop()
        if len(pcs) == 1:
            op, d, a, _b, c = self.code[pcs[0]]
            self.code[pcs[0]] = (op, d, a, (K_NONE, 0, len(self.code)), c)
        else:
            op, d, _a, b, c = self.code[pcs[1]]
            self.code[pcs[1]] = (op, d, (K_NONE, 0, len(self.code)), b, c)

    def call(self, fn: "Function", args: Sequence) -> FReg:
        """`x = g(args...)` inside a function body: g must be an earlier (finalized) function"""
        return self.call_array(fn, args, 1)[0]

    def call_array(self, fn: "Function", args: Sequence, n: int) -> List[FReg]:
        """`var r[n] = g(args...)`: the arguments are copied into consecutive fresh registers (the callee's parameters, as the
        C++ producer fills `lvarcall`), the n results land in n consecutive registers"""
        assert fn.id >= 0 and len(args) == fn.n_params and 1 <= n <= fn.n_results, "bad call of %s" % fn.name
        base = self.n_regs
        self.n_regs += fn.n_params
        for k, a in enumerate(args):
            self.set(FReg(self, base + k), a)
        d = self.n_regs
        self.n_regs += n
        self.code.append((OPS["CALL"], (K_TMP, 0, d), (K_NONE, 0, fn.id), (K_TMP, 0, base) if fn.n_params else NONE_REF,
                          (K_NONE, 0, n if n > 1 else 0)))
        return [FReg(self, d + k) for k in range(n)]

    def ret(self, value) -> None:
        self.code.append((OPS["RET"], NONE_REF, self._operand(value), NONE_REF, NONE_REF))

    def ret_array(self, base: int, n: int) -> None:
        """`return arr;` for a `var arr[n]` (ReturnBucket with_size = n, return_bucket.rs:70-120: the reference copies n
        elements into the caller's destination): registers base .. base+n-1 are the results of a `call_array`"""
        assert 1 <= n and base + n <= self.n_regs
        self.n_results = max(self.n_results, n)
        self.code.append((OPS["RET"], NONE_REF, (K_TMP, 0, base), (K_NONE, 0, n), NONE_REF))

    def finalize(self) -> "Function":
        assert not self._loops and not self._ifs and self.code and self.code[-1][0] == OPS["RET"]
        self.id = len(self.desc.functions)
        self.desc.functions.append(self)
        return self


class CircuitDesc:
    def __init__(self, prime: str = "bn128"):
        self.prime = prime
        self.q = PRIMES[prime]
        self.templates: List[Template] = []
        self.functions: List[Function] = []
        self.consts: List[int] = []
        self._cid: Dict[int, int] = {}
        self._cache: Dict[tuple, Template] = {}
        self.main: Optional[Template] = None
        self.name = "circuit"
        self.io_map_templates: set = set()   # template ids that sit in component arrays of mixed templates (mark_mixed_array)
        self.strings: List[str] = []          # the string table of log() (CProducer::get_string_table)

    def string_id(self, text: str) -> int:
        if text not in self.strings:
            self.strings.append(text)
        return self.strings.index(text)

    def const_id(self, v: int) -> int:
        v %= self.q
        i = self._cid.get(v)
        if i is None:
            i = len(self.consts)
            self.consts.append(v)
            self._cid[v] = i
        return i

    def template(self, name: str, params: tuple, build: Callable[["Template"], None]) -> Template:
        """Get-or-build the instance of template `name` for concrete `params`."""
        key = (name, params)
        t = self._cache.get(key)
        if t is None:
            label = name if not params else "%s_%s" % (name, "_".join(str(p) for p in params))
            t = Template(self, "".join(ch if ch.isalnum() else "_" for ch in label))
            build(t)
            t.finalize()
            self._cache[key] = t
        return t

    def function(self, name: str, n_params: int, build: Callable[[Function], None]) -> Function:
        key = ("fn", name, n_params)
        f = self._cache.get(key)
        if f is None:
            f = Function(self, name, n_params)
            build(f)
            f.finalize()
            self._cache[key] = f
        return f

    def set_main(self, t: Template, name: Optional[str] = None) -> "CircuitDesc":
        self.main = t
        self.name = name or t.name
        return self

    # -- sizes ---------------------------------------------------------------------------
    @property
    def total_signals(self) -> int:
        return 1 + self.main.total_signals

    def main_inputs(self) -> List[Tuple[str, int, int]]:
        """(name, global signal id of first element, size) for the main inputs, in signal order
        (the content of InputHashMap, c_code_generator.rs:575-603)."""
        out = []
        for name, n in self.main.sigs["in"]:
            off, _, _ = self.main.sig_info[name]
            out.append((name, 1 + off, n))
        return out

    def main_outputs(self) -> List[Tuple[str, int, int]]:
        out = []
        for name, n in self.main.sigs["out"]:
            off, _, _ = self.main.sig_info[name]
            out.append((name, 1 + off, n))
        return out

    # -- .sym -------------------------------------------------------------------------------
    def sym_lines(self, witness2signal: Sequence[int]) -> List[str]:
        """The lines of the reference's `--sym` file (`#s,#w,#c,name`, mkdocs/docs/circom-language/formats/sym.md;
        writer constraint_writers/src/sym_writer.rs:4-14, traversal constraint_list/src/sym_porting.rs:14-38):
        pre-order over the component tree, `s` the signal number, `w` its witness position or -1 when the signal
        was merged away, `c` the DAG node of the component's template instance (nodes are numbered when their
        first instance finishes executing: children before parents), `name` the qualified name."""
        w_of = {int(s): i for i, s in enumerate(witness2signal)}
        node_id: Dict[int, int] = {}

        def number(t: "Template"):
            if t.id in node_id:
                return
            for s in t.subs:
                number(s.tmpl)
            node_id[t.id] = len(node_id)
        number(self.main)
        lines: List[str] = []

        def visit(t: "Template", start: int, path: str):
            off = start
            for cat in ("out", "in", "inter"):
                for name, n in t.sigs[cat]:
                    for j in range(n):
                        sig = off + j
                        label = "%s.%s%s" % (path, name, "[%d]" % j if t.sig_is_array[name] else "")
                        lines.append("%d,%d,%d,%s" % (sig, w_of.get(sig, -1), node_id[t.id], label))
                    off += n
            for s in t.subs:
                visit(s.tmpl, off, path + "." + s.name)
                off += s.tmpl.total_signals
        visit(self.main, 1, "main")
        return lines

    def write_sym(self, path: str, witness2signal: Sequence[int]) -> str:
        with open(path, "w") as f:
            f.write("".join(line + "\n" for line in self.sym_lines(witness2signal)))
        return path

    # -- serialisation -------------------------------------------------------------------
    def symbol_names(self, t: "Template") -> Tuple[List[str], List[str]]:
        """(names of the own signals in numbering order, array elements spelled out; names of the sub-components):
        the content of the symbols section, what `--sym` prints per component (dag/src/sym_porting.rs:16-33)"""
        own = []
        for cat in ("out", "in", "inter"):
            for name, n in t.sigs[cat]:
                own += ["%s[%d]" % (name, j) for j in range(n)] if t.sig_is_array[name] else [name]
        return own, [s.name for s in t.subs]

    def to_bytes(self, symbols: bool = False) -> bytes:
        """the `.cb2c` file (docs/CB2C.md); symbols=True appends the optional symbols section (`cw_circuit_write_sym`)"""
        import numpy as np
        assert self.main is not None
        # constraint coefficients go through the constant table as well
        blobs = []
        expanded = {t.id: t.expanded_ops() for t in self.templates}   # (first: the expansions add constants)
        for t in self.templates:
            nb = t.name.encode()
            nb += b"\0" * ((-len(nb)) % 4)
            cons_words = []
            nterms = 0
            for con in t.constraints:
                for lc in con:
                    cons_words.append(len(lc))
                    for k in sorted(lc.keys(), key=pack_ref):
                        cons_words.append(pack_ref(k))
                        cons_words.append(self.const_id(lc[k]))
                        nterms += 1
            hdr = struct.pack("<I", len(t.name.encode())) + nb
            t_ops, t_n_tmp = expanded[t.id]
            hdr += struct.pack("<8I", t.n_out, t.n_in, t.n_inter, len(t.subs), t_n_tmp, len(t_ops),
                               len(t.constraints), nterms)
            subs = np.array([s.tmpl.id for s in t.subs], dtype="<u4").tobytes()
            ops = np.array([[op, pack_ref(d), pack_ref(a), pack_ref(b), pack_ref(c)]
                            for op, d, a, b, c in t_ops], dtype="<u8").reshape(-1, 5).tobytes()
            cons = np.array(cons_words, dtype="<u8").tobytes()
            blobs.append(hdr + subs + ops + cons)
        names = b""
        ins = self.main_inputs()
        for name, gid, n in ins:
            nb = name.encode()
            names += struct.pack("<I", len(nb)) + nb + b"\0" * ((-len(nb)) % 4)
            names += struct.pack("<II", gid, n)
        funcs = b""
        for f in self.functions:
            nb = f.name.encode()
            funcs += struct.pack("<I", len(nb)) + nb + b"\0" * ((-len(nb)) % 4)
            funcs += struct.pack("<3I", f.n_params, f.n_regs, len(f.code))
            funcs += np.array([[op, pack_ref(d), pack_ref(a), pack_ref(b), pack_ref(c)] for op, d, a, b, c in f.code],
                              dtype="<u8").reshape(-1, 5).tobytes()
        head = b"CB2C" + struct.pack("<7I", 1, PRIME_IDS[self.prime], len(self.consts), len(self.templates),
                                     self.main.id, len(ins), len(self.functions))
        consts = b"".join(int(c).to_bytes(32, "little") for c in self.consts)
        syms = b""
        if symbols:
            def pstr(x: str) -> bytes:
                nb = x.encode()
                return struct.pack("<I", len(nb)) + nb + b"\0" * ((-len(nb)) % 4)
            syms = b"SYMS"
            for t in self.templates:
                own, subs = self.symbol_names(t)
                syms += b"".join(pstr(x) for x in own) + b"".join(pstr(x) for x in subs)
        logs = b""
        if self.strings:      # optional string table of log(): "LOGS", u32 count, count x str
            logs = b"LOGS" + struct.pack("<I", len(self.strings))
            for x in self.strings:
                nb = x.encode()
                logs += struct.pack("<I", len(nb)) + nb + b"\0" * ((-len(nb)) % 4)
        iomp = b""
        if self.io_map_templates:
            # the compiler's TemplateInstanceIOMap for the template instances that sit in component arrays of mixed
            # templates (build_input_output_list, compiler/src/circuit_design/build.rs:498-520): docs/CB2C.md, IOMP
            iomp = b"IOMP" + struct.pack("<I", len(self.io_map_templates))
            for tid in sorted(self.io_map_templates):
                defs = self.io_defs(self.templates[tid])
                iomp += struct.pack("<II", tid, len(defs))
                for off, lengths, size, bus in defs:
                    iomp += struct.pack("<II", off, len(lengths)) + b"".join(struct.pack("<I", x) for x in lengths)
                    iomp += struct.pack("<II", size, bus)
        return head + consts + b"".join(blobs) + names + funcs + logs + iomp + syms

    @staticmethod
    def io_defs(t: "Template"):
        """IODef list of a template instance: (offset = local id of element 0, dimensions, element size, bus id) per output
        and input signal, in declaration order = signal code order (build.rs:498-520)"""
        defs, off = [], 0
        for cat in ("out", "in"):
            for name, n in t.sigs[cat]:
                defs.append((off, [n] if t.sig_is_array[name] else [], 1, 0))
                off += n
        return defs

    def mark_mixed_array(self, *templates: "Template") -> None:
        """declare that these template instances occur together in one component array (`component ops[2]; ops[0] = A();
        ops[1] = B();`): the reference compiler then addresses their signals through the io map (`LocationRule::Mapped`,
        store_bucket.rs:498-566) and ships the map in the `.dat`"""
        for t in templates:
            self.io_map_templates.add(t.id)

    def save(self, path: str) -> str:
        with open(path, "wb") as f:
            f.write(self.to_bytes())
        return path
