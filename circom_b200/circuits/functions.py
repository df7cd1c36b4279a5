"""Circuits whose `<--` hints are computed by circom *functions* with data-dependent control flow
(the style of circom-ecdsa's bigint_func.circom: `long_div`, `log_ceil`, ...): loops whose trip count
depends on run-time values, branches on run-time conditions, `var` arrays indexed by loop counters."""
from __future__ import annotations

from ..circuit import CircuitDesc, Function, Template
from .basic import less_than, num2bits


def fn_bit_length(d: CircuitDesc) -> Function:
    """function bit_length(x) { var n = 0; while (x > 0) { x = x >> 1; n++; } return n; }"""
    def build(f: Function):
        x = f.var(f.param(0))
        n = f.var(0)
        f.loop_begin()
        f.loop_break_if_zero(x.neq(0))
        f.set(x, x >> 1)
        f.set(n, n + 1)
        f.loop_end()
        f.ret(n)
    return d.function("bit_length", 1, build)


def fn_divmod(d: CircuitDesc, nbits: int = 64) -> Function:
    """function divmod(a, b, sel): schoolbook restoring division over the bits of `a` (kept in a `var`
    array filled by a loop and read back with a run-time index), returns the quotient (sel == 0) or the
    remainder (sel != 0).  b == 0 returns 0."""
    def build(f: Function):
        a, b, sel = f.param(0), f.param(1), f.param(2)
        bits = f.array(nbits)
        i = f.var(0)
        t = f.var(a)
        top = f.var(0)
        # bits[i] = (a >> i) & 1, remember the highest set bit
        f.loop_begin()
        f.loop_break_if_zero(t.neq(0))
        f.store(bits, i, t & 1)
        f.set(t, t >> 1)
        f.set(i, i + 1)
        f.loop_end()
        f.set(top, i)
        q = f.var(0)
        r = f.var(0)
        f.if_begin(b.neq(0))
        f.loop_begin()
        f.loop_break_if_zero(top.neq(0))
        f.set(top, top - 1)
        f.set(r, r * 2 + f.load(bits, top))
        f.set(q, q * 2)
        f.if_begin(r.geq(b))
        f.set(r, r - b)
        f.set(q, q + 1)
        f.if_end()
        f.loop_end()
        f.if_end()
        res = f.var(q)
        f.if_begin(sel.neq(0))
        f.set(res, r)
        f.if_end()
        f.ret(res)
    return d.function("divmod%d" % nbits, 3, build)


def int_div(d: CircuitDesc, nbits: int = 32) -> Template:
    """q, r with a = q*b + r, 0 <= r < b (b != 0), hints from the `divmod` function, constrained with
    range checks on q, r and LessThan(r, b)."""
    fdiv = fn_divmod(d, 64)
    fbl = fn_bit_length(d)
    n2b = num2bits(d, nbits)
    lt = less_than(d, nbits)

    def build(t: Template):
        a = t.input("a")
        b = t.input("b")
        q = t.output("q")
        r = t.output("r")
        nb = t.output("nbits")
        t.assign(q, t.call(fdiv, [a, b, 0]))
        t.assign(r, t.call(fdiv, [a, b, 1]))
        t.assign(nb, t.call(fbl, [a]))
        t.constrain(q * b + r, a)
        cq = t.component("rq", n2b)
        t.assign_constrained(cq["in"], q)
        cr = t.component("rr", n2b)
        t.assign_constrained(cr["in"], r)
        c = t.component("lt", lt)
        t.assign_constrained(c["in", 0], r)
        t.assign_constrained(c["in", 1], b)
        t.constrain(c["out"], 1)
    return d.template("IntDiv", (nbits,), build)


def fn_divmod_array(d: CircuitDesc, nbits: int = 64) -> Function:
    """function divmod_arr(a, b) { var out[3]; ... return out; }: quotient, remainder and bit length of `a` from ONE
    call (`var qr[3] = divmod_arr(a, b);`).  An early `return out;` for b == 0 makes it a function with two RETs."""
    def build(f: Function):
        a, b = f.param(0), f.param(1)
        out = f.array(3)
        bits = f.array(nbits)
        i = f.var(0)
        t = f.var(a)
        f.loop_begin()
        f.loop_break_if_zero(t.neq(0))
        f.store(bits, i, t & 1)
        f.set(t, t >> 1)
        f.set(i, i + 1)
        f.loop_end()
        two = f.var(2)
        f.store(out, two, i)
        f.if_begin(b.eq(0))
        f.ret_array(out, 3)
        f.if_end()
        q = f.var(0)
        r = f.var(0)
        f.loop_begin()
        f.loop_break_if_zero(i.neq(0))
        f.set(i, i - 1)
        f.set(r, r * 2 + f.load(bits, i))
        f.set(q, q * 2)
        f.if_begin(r.geq(b))
        f.set(r, r - b)
        f.set(q, q + 1)
        f.if_end()
        f.loop_end()
        zero = f.var(0)
        one = f.var(1)
        f.store(out, zero, q)
        f.store(out, one, r)
        f.ret_array(out, 3)
    return d.function("divmod_arr%d" % nbits, 2, build)


def int_div_array(d: CircuitDesc, nbits: int = 32, use: str = "all") -> Template:
    """IntDiv with its three hints from one array-returning call.  use = "all": q, r, nbits are outputs;
    "tail": only r and nbits are read (the first result of the call is dead); "head": only q (the others are dead)."""
    fdiv = fn_divmod_array(d, 64)
    n2b = num2bits(d, nbits)

    def build(t: Template):
        a = t.input("a")
        b = t.input("b")
        if use == "head":
            q = t.output("q")
            res = t.call_array(fdiv, [a, b], 3)
            t.assign(q, res[0])
            cq = t.component("rq", n2b)
            t.assign_constrained(cq["in"], q)
            return
        if use == "tail":
            r = t.output("r")
            nb = t.output("nbits")
            res = t.call_array(fdiv, [a, b], 3)
            t.assign(r, res[1])
            t.assign(nb, res[2])
            cr = t.component("rr", n2b)
            t.assign_constrained(cr["in"], r)
            return
        q = t.output("q")
        r = t.output("r")
        nb = t.output("nbits")
        res = t.call_array(fdiv, [a, b], 2)            # fewer results than the function returns: the first two
        t.assign(q, res[0])
        t.assign(r, res[1] + 0 * res[0])
        t.assign(nb, t.call_array(fdiv, [a, b], 3)[2])
        t.constrain(q * b + r, a)
        cq = t.component("rq", n2b)
        t.assign_constrained(cq["in"], q)
        cr = t.component("rr", n2b)
        t.assign_constrained(cr["in"], r)
    return d.template("IntDivArr_" + use, (nbits,), build)


def fn_gcd(d: CircuitDesc) -> Function:
    """function gcd(a, b) { while (b != 0) { var qr[3] = divmod_arr(a, b); a = b; b = qr[1]; } return a; } - a function
    that calls another one (with an array result) inside a loop whose trip count depends on the data, the way
    circom-ecdsa's bigint functions are built from `long_div` / `short_div` / `long_sub`"""
    fdiv = fn_divmod_array(d, 64)
    fbl = fn_bit_length(d)

    def build(f: Function):
        a = f.var(f.param(0))
        b = f.var(f.param(1))
        steps = f.var(0)
        f.loop_begin()
        f.loop_break_if_zero(b.neq(0))
        qr = f.call_array(fdiv, [a, b], 3)
        f.set(a, b)
        f.set(b, qr[1])
        f.set(steps, steps + f.call(fbl, [qr[0]]))      # a second callee, scalar result
        f.loop_end()
        out = f.array(2)
        f.store(out, f.var(0), a)
        f.store(out, f.var(1), steps)
        f.ret_array(out, 2)
    return d.function("gcd", 2, build)


def gcd_circuit(d: CircuitDesc, nbits: int = 32) -> Template:
    """g = gcd(a, b) as a hint (two levels of function calls), constrained only by range checks: a test of the call path"""
    fg = fn_gcd(d)
    n2b = num2bits(d, nbits)

    def build(t: Template):
        a = t.input("a")
        b = t.input("b")
        g = t.output("g")
        s = t.output("steps")
        res = t.call_array(fg, [a, b], 2)
        t.assign(g, res[0])
        t.assign(s, res[1])
        c = t.component("rg", n2b)
        t.assign_constrained(c["in"], g)
    return d.template("Gcd", (nbits,), build)
