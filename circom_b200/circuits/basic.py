"""Small circuits: the only ones whose circom source is in the reference tree.

  Multiplier2   mkdocs/docs/getting-started/writing-circuits.md:17-28
  IsZero        mkdocs/docs/circom-language/basic-operators.md:133-147
  Num2Bits      mkdocs/docs/circom-language/basic-operators.md:151-170
  MultiplierN / AndN   mkdocs/docs/more-circuits/more-basic-circuits.md
plus `AllOps`, a synthetic template that exercises every OperatorType once
(compiler/src/intermediate_representation/compute_bucket.rs:7-34).
"""
from __future__ import annotations

from ..circuit import CircuitDesc, Template


def multiplier2(d: CircuitDesc) -> Template:
    def build(t: Template):
        a = t.input("a")
        b = t.input("b")
        c = t.output("c")
        t.assign_constrained(c, a * b)
    return d.template("Multiplier2", (), build)


def is_zero(d: CircuitDesc) -> Template:
    def build(t: Template):
        i = t.input("in")
        out = t.output("out")
        inv = t.signal("inv")
        # inv <-- in!=0 ? 1/in : 0   (both arms evaluated; 1/0 == 0 in the runtime)
        t.assign(inv, t.select(i.neq(0), t.const(1) / i, 0))
        t.assign_constrained(out, -i * inv + 1)
        t.constrain(i * out, 0)
    return d.template("IsZero", (), build)


def num2bits(d: CircuitDesc, n: int) -> Template:
    def build(t: Template):
        i = t.input("in")
        out = t.output("out", n)
        lc1 = t.const(0)
        e2 = 1
        for k in range(n):
            t.assign(out[k], (i >> k) & 1)
            t.constrain(out[k] * (out[k] - 1), 0)
            lc1 = lc1 + out[k] * e2
            e2 = e2 + e2
        t.constrain(lc1, i)
    return d.template("Num2Bits", (n,), build)


def bits2num(d: CircuitDesc, n: int) -> Template:
    def build(t: Template):
        i = t.input("in", n)
        out = t.output("out")
        lc1 = t.const(0)
        e2 = 1
        for k in range(n):
            lc1 = lc1 + i[k] * e2
            e2 = e2 + e2
        t.assign_constrained(out, lc1)
    return d.template("Bits2Num", (n,), build)


def less_than(d: CircuitDesc, n: int) -> Template:
    """circomlib-style LessThan(n): Num2Bits(n+1) of in[0] + 2^n - in[1]."""
    n2b = num2bits(d, n + 1)

    def build(t: Template):
        i = t.input("in", 2)
        out = t.output("out")
        c = t.component("n2b", n2b)
        t.assign_constrained(c["in"], i[0] + (1 << n) - i[1])
        t.assign_constrained(out, 1 - c["out", n])
    return d.template("LessThan", (n,), build)


def multiplier_n(d: CircuitDesc, n: int) -> Template:
    m2 = multiplier2(d)

    def build(t: Template):
        i = t.input("in", n)
        out = t.output("out")
        comps = [t.component("comp[%d]" % k, m2) for k in range(n - 1)]
        t.assign_constrained(comps[0]["a"], i[0])
        t.assign_constrained(comps[0]["b"], i[1])
        for k in range(n - 2):
            t.assign_constrained(comps[k + 1]["a"], comps[k]["c"])
            t.assign_constrained(comps[k + 1]["b"], i[k + 2])
        t.assign_constrained(out, comps[n - 2]["c"])
    return d.template("MultiplierN", (n,), build)


def all_ops(d: CircuitDesc) -> Template:
    """One use of every runtime operator; outputs are `<--` hints so that no constraint
    restricts the input domain (division by zero in `\\` and `%` is guarded)."""
    iz = is_zero(d)

    def build(t: Template):
        a = t.input("a")
        b = t.input("b")
        outs = t.output("o", 27)
        vals = [
            a * b, a / b, a + b, a - b, a ** (b & 15),
            a // t.select(b.eq(0), 1, b), a % t.select(b.eq(0), 1, b),
            a << (b & 255), a >> (b & 255), a << b, a >> b,
            a.leq(b), a.geq(b), a.lt(b), a.gt(b), a.eq(b), a.neq(b),
            a.lor(b), a.land(b), a.lnot(), a | b, a & b, a ^ b, ~a, -a,
        ]
        for k, v in enumerate(vals):
            t.assign(outs[k], v)
        z = t.component("iz", iz)
        t.assign_constrained(z["in"], a - b)
        t.assign_constrained(outs[25], z["out"])
        t.assign_constrained(outs[26], a * b + a)
    return d.template("AllOps", (), build)


def mixed_array(d: CircuitDesc) -> Template:
    """A component array whose elements are instances of DIFFERENT templates -

        template Acc(n) { signal input in[n]; signal input k; signal output out; ... }
        component ops[3];  ops[0] = Acc(2);  ops[1] = Acc(3);  ops[2] = Scale();

    - the case in which the reference compiler cannot know a sub-component signal's offset statically (the inputs of
    Acc(2) and Acc(3) start at the same place but `k` does not) and emits `LocationRule::Mapped` accesses that look the
    offset up in templateInsId2IOSignalInfo at run time (store_bucket.rs:498-566, load_bucket.rs:264-322).  A producer that
    unrolls the loops knows every element's template: the description below is what it writes, and the io map travels
    to the `.dat` (CircuitDesc.mark_mixed_array)."""
    def acc(n):
        def build(t: Template):
            xs = t.input("in", n)
            k = t.input("k")
            out = t.output("out")
            s = xs[0]
            for x in xs[1:]:
                s = s + x
            t.assign_constrained(out, s * k)
        return d.template("Acc", (n,), build)

    def scale(t: Template):
        out = t.output("out", 2)
        x = t.input("x")
        y = t.input("y")
        t.assign_constrained(out[0], x * y)
        t.assign_constrained(out[1], x * x)
    a2, a3, sc = acc(2), acc(3), d.template("Scale", (), scale)
    d.mark_mixed_array(a2, a3, sc)

    def build(t: Template):
        a = t.input("a", 3)
        b = t.input("b")
        out = t.output("out", 4)
        o0 = t.component("ops[0]", a2)
        o1 = t.component("ops[1]", a3)
        o2 = t.component("ops[2]", sc)
        t.assign_constrained(o0.sig("in", 0), a[0])
        t.assign_constrained(o0.sig("in", 1), a[1])
        t.assign_constrained(o0.sig("k"), b)
        for j in range(3):
            t.assign_constrained(o1.sig("in", j), a[j])
        t.assign_constrained(o1.sig("k"), o0.sig("out"))
        t.assign_constrained(o2.sig("x"), o1.sig("out"))
        t.assign_constrained(o2.sig("y"), b)
        t.assign_constrained(out[0], o0.sig("out"))
        t.assign_constrained(out[1], o1.sig("out"))
        t.assign_constrained(out[2], o2.sig("out", 0))
        t.assign_constrained(out[3], o2.sig("out", 1))
    return d.template("MixedArray", (), build)


def table_lookup(d: CircuitDesc, n: int = 8) -> Template:
    """`out <-- table[sel]` with a signal index - a witness hint the reference compiles to a load at a run-time address
    (Template.load_indexed) - constrained the circomlib way: a one-hot decomposition of sel (IsEqual per position via the
    inverse trick is overkill here: eq[i] * (sel - i) === 0, sum eq[i] === 1) and out === sum eq[i] * table[i]."""
    def build(t: Template):
        table = t.input("table", n)
        sel = t.input("sel")
        out = t.output("out")
        picked = t.signal("picked")
        t.assign(picked, t.load_indexed(table, sel))          # <-- table[sel]
        eq = t.signal("eq", n)
        prod = t.signal("prod", n)
        acc_e, acc_p = None, None
        for i in range(n):
            t.assign(eq[i], sel.eq(i))
            t.constrain(eq[i] * (sel - i), t.const(0))
            t.assign_constrained(prod[i], eq[i] * table[i])
            acc_e = eq[i] if acc_e is None else acc_e + eq[i]
            acc_p = prod[i] if acc_p is None else acc_p + prod[i]
        t.constrain(acc_e, t.const(1))
        t.constrain(picked, acc_p)
        t.assign_constrained(out, picked * picked)
    return d.template("TableLookup", (n,), build)


def logging(d: CircuitDesc) -> Template:
    """log() calls in a component tree: strings, constants, signals of the component and of a sub-component, a signal that
    the signal merging eliminates (its value is printed through the entry it was merged into), calls in execution order
    (the sub-component logs when its last input arrives)."""
    def inner(t: Template):
        x = t.input("x")
        y = t.input("y")
        out = t.output("out")
        t.assign_constrained(out, x * y)
        t.log("inner", x, y, "->", out)
    it = d.template("Inner", (), inner)

    def build(t: Template):
        a = t.input("a")
        b = t.input("b")
        out = t.output("out")
        t.log("start", a, 42)
        c = t.component("c", it)
        t.assign_constrained(c.sig("x"), a)
        t.log("between the inputs of c")
        t.assign_constrained(c.sig("y"), b + 1)
        m = t.signal("m")
        t.assign_constrained(m, c.sig("out"))          # m = c.out: merged away by the signal elimination
        t.assign_constrained(out, m + a)
        t.log(m, c.sig("out"), out, -1)
    return d.template("Logging", (), build)
