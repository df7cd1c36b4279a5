"""Randomised circuits against the Python evaluator: the lowering's value-preserving rewrites (range analysis,
bit-field fusion, shifts for x * 2^k, products without reduction, x * 0, representation inference, bit runs, the
forwarding-ring flags) are exercised on expression DAGs they were not hand-written for.  The tape runs on the CPU
build of the device code (tests/hostsim) with all rewrites, with CW_FLAG_NO_PEEPHOLE and with CW_FLAG_O0; every
witness entry must equal the evaluator's, which follows the reference's operator semantics
(circom_algebra/src/modular_arithmetic.rs:26-215, generic/fr.cpp)."""
import random

import pytest

from circom_b200.circuit import CircuitDesc
from circom_b200 import circuits as C
from oracle.ir_eval import evaluate, check_r1cs
from oracle.field_model import DivisionByZero
from tests.util import hostsim_run, limbs_to_ints, edge_values

CW_FLAG_O0, CW_FLAG_NO_PEEPHOLE, CW_FLAG_BITPLANE, CW_FLAG_REUSE, CW_FLAG_COMPACT = 4, 8, 16, 32, 48


def random_template(d, rng, n_in, n_vals, with_components):
    """values are `<--` hints (no constraint restricts the inputs); sub-components get masked values so that
    their own constraints (Num2Bits recomposition, LessThan) hold for every input"""
    n2b = {n: C.num2bits(d, n) for n in (5, 33, 64)} if with_components else {}
    q = d.q

    def build(t):
        ins = t.input("x", n_in)
        outs = t.output("o", n_vals)
        pool = [ins[i] for i in range(n_in)]
        small = []    # values known (by construction) to be < 2^64

        def pick():
            return rng.choice(pool)

        def const():
            return rng.choice([0, 1, 2, 3, 5, 255, 2**16, 2**31 - 1, 2**32, 2**63, 2**64 - 1, q - 1, q - 2,
                               rng.randrange(q), rng.randrange(2**20), 1 << rng.randrange(0, 253)])

        for k in range(n_vals):
            kind = rng.randrange(16)
            a, b = pick(), pick()
            if kind == 0:      # bit-field of a value
                v = (a >> rng.randrange(0, 254)) & ((1 << rng.randrange(1, 65)) - 1)
                small.append(v)
            elif kind == 1:    # low mask
                v = a & ((1 << rng.randrange(1, 200)) - 1)
            elif kind == 2:    # product with a constant (0, 1, powers of two, small, large)
                v = (rng.choice(small) if small and rng.random() < 0.6 else a) * const()
            elif kind == 3:    # product of two values, often both small
                v = (rng.choice(small) * rng.choice(small)) if len(small) >= 2 and rng.random() < 0.7 else a * b
                if len(small) >= 2 and rng.random() < 0.3:
                    small.append(v & ((1 << 64) - 1))
            elif kind == 4:
                v = a + b if rng.random() < 0.5 else a - b
            elif kind == 5:    # recomposition of a few bits of one source, in and out of order
                src = pick()
                lo = rng.randrange(0, 200)
                terms = [((src >> (lo + i)) & 1) * (1 << (lo + i)) for i in range(rng.randrange(2, 12))]
                rng.shuffle(terms) if rng.random() < 0.3 else None
                v = terms[0]
                for x in terms[1:]:
                    v = v + x
            elif kind == 6 and rng.random() < 0.4:
                # shift by a SIGNAL: amounts >= q - qbits reverse the direction (generic/fr.cpp:2157-2173), so the
                # operand's width bounds nothing; the result then feeds a rewrite that trusts the width
                sh = (a & 0xFF) >> b if rng.random() < 0.5 else (a & 0xFF) << b
                v = sh * (1 << rng.choice([1, 100, 200])) if rng.random() < 0.7 else sh + (b & 0xFFFF) * 3
            elif kind == 6:
                v = a << rng.choice([0, 1, 7, 64, 200, 253, 254, 300]) if rng.random() < 0.5 else a >> rng.choice([0, 1, 31, 32, 33, 64, 253, 254, 300])
            elif kind == 7:
                v = rng.choice([a.lt(b), a.gt(b), a.leq(b), a.geq(b), a.eq(b), a.neq(b), a.lor(b), a.land(b), a.lnot()])
                small.append(v)
            elif kind == 8:
                v = t.select(a.lt(b), a, b) if rng.random() < 0.5 else t.select(a & 1, b, const())
            elif kind == 9:    # integer division / remainder by a non-zero value
                dv = (b & ((1 << rng.randrange(1, 64)) - 1)) + 1
                v = a // dv if rng.random() < 0.5 else a % dv
            elif kind == 10:
                v = rng.choice([a | b, a ^ b, a & b, ~a, -a])
            elif kind == 11:   # field division and small powers
                v = a / (b + 1) if rng.random() < 0.5 else a ** (b & 7)
            elif kind == 12 and small:   # sums and scaled sums of small values (no reduction needed)
                v = rng.choice(small) * rng.choice([3, 10, 2**20, 2**64]) + rng.choice(small)
            elif kind == 13:
                v = (a * b + a) * const()
            elif kind == 14 and n2b:     # a range-checked value: Num2Bits on a masked value
                n = rng.choice(sorted(n2b))
                comp = t.component("n2b_%d" % k, n2b[n])
                masked = t.signal("m_%d" % k)
                t.assign(masked, a & ((1 << n) - 1))
                t.assign_constrained(comp["in"], masked)
                v = comp["out"][rng.randrange(n)] + comp["out"][rng.randrange(n)] * 2
                small.append(masked)
            else:
                v = a * const() + b
            t.assign(outs[k], v)
            pool.append(outs[k])
    return d.template("Fuzz", (), build)


def rand_input(rng, q, edges):
    r = rng.random()
    if r < 0.35:
        return rng.choice(edges)
    if r < 0.6:
        return rng.randrange(2 ** rng.choice([1, 8, 31, 32, 64, 128]))
    return rng.randrange(q)


@pytest.mark.parametrize("prime", ["bn128", "bls12381", "goldilocks"])
@pytest.mark.parametrize("seed", range(40))
def test_random_circuits_match_the_evaluator(prime, seed):
    rng = random.Random(1000 * seed + (7 if prime == "bn128" else 13))
    d = CircuitDesc(prime)
    n_in = rng.randrange(2, 5)
    d.set_main(random_template(d, rng, n_in, n_vals=rng.randrange(20, 70), with_components=seed % 2 == 0))
    edges = edge_values(d.q)
    ins = [{"x": [rand_input(rng, d.q, edges) for _ in range(n_in)]} for _ in range(10)]
    expected = [evaluate(d, inp) for inp in ins]
    for e in expected:
        assert check_r1cs(d, e) == 0
    for flags in (0, CW_FLAG_NO_PEEPHOLE, CW_FLAG_O0, CW_FLAG_BITPLANE, CW_FLAG_BITPLANE | CW_FLAG_O0, CW_FLAG_REUSE, CW_FLAG_COMPACT,
                  CW_FLAG_COMPACT | CW_FLAG_O0, CW_FLAG_REUSE | CW_FLAG_NO_PEEPHOLE, 64, 64 | CW_FLAG_COMPACT):
        wit, st, stats, w2s = hostsim_run(d, ins, flags=flags)
        assert not st.any(), (prime, seed, flags, st)
        for i, e in enumerate(expected):
            got = limbs_to_ints(wit[i])
            want = [e[k] for k in w2s]
            if got != want:
                bad = [j for j in range(len(got)) if got[j] != want[j]][:5]
                raise AssertionError("prime %s seed %d flags %d input %d: witness entries %s differ" % (prime, seed, flags, i, bad))


@pytest.mark.parametrize("prime", ["bn128", "bls12381", "goldilocks"])
def test_shift_by_negative_signal_amount_is_not_treated_as_narrow(prime):
    """r = (x & 0xFF) >> y with y = q - 200 is a LEFT shift by 200 (Fr_shr, generic/fr.cpp:2189-2263): r is ~208 bits
    wide, and r * 2^100 must be reduced modulo q - the range analysis once took r for an 8-bit value"""
    d = CircuitDesc(prime)

    def build(t):
        x, y = t.input("x"), t.input("y")
        o = t.output("o", 3)
        r = (x & 0xFF) >> y
        t.assign(o[0], r * (1 << 100))
        l = (x & 0xFF) << y
        t.assign(o[1], l * (1 << 100))
        t.assign(o[2], ((x & 0xFFFF) >> (y & 0xF)) * (1 << 230))   # a provably plain amount keeps the fast path
    d.set_main(d.template("ShiftNeg", (), build))
    q = d.q
    ins = [{"x": xv, "y": yv} for xv in (0xAB, q - 1, 0x1FF) for yv in (q - 200, q - 1, q - 253, 3, 0, 254, 255, q - 254, 2**64)]
    expected = [evaluate(d, inp) for inp in ins]
    for flags in (0, CW_FLAG_NO_PEEPHOLE, CW_FLAG_BITPLANE, CW_FLAG_COMPACT):
        wit, st, stats, w2s = hostsim_run(d, ins, flags=flags)
        assert not st.any()
        for i, e in enumerate(expected):
            assert limbs_to_ints(wit[i]) == [e[k] for k in w2s], (prime, flags, ins[i])


def random_function(d, rng, n_params, callees=(), name="fz"):
    """a random structured function body: scalars, `var` arrays with run-time indices (kept in range by masking), nested
    if / else and counted loops, values that are reused long after their definition, registers read before they are
    written (the frame is zero-initialised), scalar or array return.  Exercises the lowering's register allocation
    (liveness over loops and branches) and both register machines (values leave 128 bits now and then)."""
    from circom_b200.circuit import FReg

    def build(f):
        scalars = [f.param(i) for i in range(n_params)]
        arrays = []
        if n_params >= 4 and rng.random() < 0.5:
            arrays.append((f.param_array(0, 4), 4))
            scalars = scalars[4:] or [f.var(3)]
        for _ in range(rng.randrange(0, 3)):
            n = rng.choice([2, 4, 8])
            arrays.append((f.array(n, rng.randrange(5)), n))
        for _ in range(rng.randrange(1, 4)):
            scalars.append(f.var(rng.choice([0, 1, 7, rng.getrandbits(64)])))
        for _ in range(rng.randrange(0, 2)):           # never initialised: reads the zero of the fresh frame
            r = FReg(f, f.n_regs)
            f.n_regs += 1
            scalars.append(r)

        def operand():
            return rng.choice(scalars) if rng.random() < 0.8 else rng.choice([0, 1, 2, 3, 255, 2**32, 2**64 - 1, d.q - 1])

        def expr(depth=0):
            a = rng.choice(scalars)
            if callees and rng.random() < 0.12:       # a nested call (the callee may itself call an earlier function)
                g = rng.choice(callees)
                n = rng.randrange(1, g.n_results + 1)
                res = f.call_array(g, [operand() for _ in range(g.n_params)], n)
                return res[rng.randrange(n)] + 0
            k = rng.random()
            if k < 0.45:
                op = rng.choice(["__add__", "__mul__", "__and__", "__or__", "__xor__", "lt", "gt", "eq", "neq", "leq", "geq"])
                return getattr(a, op)(operand() if depth or rng.random() < 0.6 else expr(1))
            if k < 0.6:
                return a - operand()
            if k < 0.75:
                return a >> rng.randrange(0, 70)
            if k < 0.85:
                return a << rng.randrange(0, 40)
            if k < 0.92 and arrays:
                base, n = rng.choice(arrays)
                return f.load(base, a & (n - 1))
            if k < 0.96:
                return a // (operand() if rng.random() < 0.5 else 3)
            return a / 7 if rng.random() < 0.5 else -a

        def block(depth, budget):
            for _ in range(rng.randrange(1, budget)):
                k = rng.random()
                if k < 0.5 or depth >= 3:
                    tgt = rng.choice(scalars[n_params if not arrays or arrays[0][0] else 0:] or scalars)
                    if tgt.idx < n_params and arrays and arrays[0][0] == 0:
                        continue
                    f.set(tgt, expr())
                elif k < 0.65 and arrays:
                    base, n = rng.choice(arrays)
                    f.store(base, rng.choice(scalars) & (n - 1), expr())
                elif k < 0.85:
                    f.if_begin(expr())
                    block(depth + 1, 4)
                    if rng.random() < 0.5:
                        f.if_else()
                        block(depth + 1, 4)
                    f.if_end()
                else:
                    i = f.var(0)
                    n = rng.randrange(1, 5)
                    f.loop_begin()
                    f.loop_break_if_zero(i.lt(n))
                    block(depth + 1, 4)
                    f.set(i, i + 1)
                    f.loop_end()
                    scalars.append(i)
        block(0, 8)
        outs = [a for a in arrays if a[0] >= n_params]
        if outs and rng.random() < 0.5:
            base, n = rng.choice(outs)
            f.ret_array(base, n)
        else:
            f.ret(expr())
    return d.function(name, n_params, build)


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("CW_FUZZ_FUNCS", "150"))))
def test_random_function_bodies_match_the_evaluator(seed):
    import ctypes
    from tests.util import hostsim
    rng = random.Random(1000 + seed)
    prime = rng.choice(["bn128", "bls12381", "secq256r1"]) if seed % 5 else "goldilocks"   # (goldilocks: the full-width machine only)
    d = CircuitDesc(prime)
    n_params = rng.randrange(1, 6)
    callees = []
    if seed % 3 == 0:        # every third circuit: one or two helper functions, the second may call the first
        for k in range(rng.randrange(1, 3)):
            callees.append(random_function(d, rng, rng.randrange(1, 4), tuple(callees), "helper%d" % k))
    fn = random_function(d, rng, n_params, tuple(callees))
    n_res = fn.n_results

    def build(t):
        ins = t.input("x", n_params)
        outs = t.output("o", n_res)
        res = t.call_array(fn, ins, n_res) if n_res > 1 else [t.call(fn, ins)]
        for k in range(n_res):
            t.assign(outs[k], res[k])
    d.set_main(d.template("Fz", (), build))
    q = d.q
    ins = [{"x": [rng.choice([0, 1, 2, rng.getrandbits(64), rng.getrandbits(32), rng.getrandbits(120), q - 1, rng.randrange(q)])
                  for _ in range(n_params)]} for _ in range(12)]
    ok_ins, exps = [], []
    for inp in ins:
        try:
            exps.append(evaluate(d, inp))
            ok_ins.append(inp)
        except (DivisionByZero, AssertionError, RuntimeError):
            pass          # division by zero / a runaway loop: error statuses are covered elsewhere
    if not ok_ins:
        return
    hs = hostsim()
    na, wi = ctypes.c_ulonglong(), ctypes.c_ulonglong()
    hs.hs_vm_counters(ctypes.byref(na), ctypes.byref(wi))
    for flags in (0, 48, 112):      # plain, compact store, compact + fused work items (calls stay items of their own)
        wit, st, _, w2s = hostsim_run(d, ok_ins, flags=flags)
        assert not st.any()
        for i, exp in enumerate(exps):
            assert limbs_to_ints(wit[i]) == [exp[k] for k in w2s], (seed, i)
    # the packed frame never exceeds the declared one
    from circom_b200.witness_calculator import Circuit
    assert Circuit(d, host_only=True).functions()[-1]["n_regs"] <= fn.n_regs


def bit_logic_template(d, rng, n_in):
    """random circuits of the kind the integer rows are for: range checks (boolean rows, recomposition sums that become runs
    of plane bits), and / xor / majority-style products of bits, linear forms with +-2^k coefficients up to 2^44, a few
    field-sized coefficients and long sums in between"""
    widths = [rng.choice([3, 8, 16, 33]) for _ in range(n_in)]

    def build(t):
        xs = t.input("x", n_in)
        out = t.output("out")
        vals = []
        for i, x in enumerate(xs):
            b = t.signal("b%d" % i, widths[i])
            acc = t.const(0)
            for k in range(widths[i]):
                t.assign(b[k], (x >> k) & 1)
                t.constrain(b[k] * (b[k] - 1), 0)
                acc = acc + b[k] * (1 << k)
            if rng.random() < 0.8:
                t.constrain(acc, x)
            vals += b
        for j in range(rng.randrange(12, 40)):
            a, b2, c = rng.choice(vals), rng.choice(vals), rng.choice(vals)
            s = t.signal("v%d" % j)
            kind = rng.choice(["and", "xor", "lin", "lin", "big", "sum"])
            if kind == "and":
                t.assign_constrained(s, a * b2)
            elif kind == "xor":
                t.assign_constrained(s, a + b2 - 2 * a * b2)
            elif kind == "lin":
                k1, k2 = rng.randrange(0, 31), rng.randrange(0, 46)
                t.assign_constrained(s, (a * (1 << k1) - b2 + c * (1 << k2)) * rng.choice(vals))
            elif kind == "big":
                t.assign_constrained(s, (a * rng.randrange(d.q) + b2) * c)
            else:
                e = t.const(0)
                for _ in range(rng.randrange(30, 70)):
                    e = e + rng.choice(vals) * (1 << rng.randrange(0, 20)) * rng.choice([1, -1])
                t.assign_constrained(s, e * a)
            vals.append(s)
        t.assign_constrained(out, vals[-1] * vals[-2])
    return d.template("BitLogic%d" % rng.randrange(1 << 30), (), build)


@pytest.mark.parametrize("seed", range(30))
def test_compiled_r1cs_with_integer_rows_on_random_circuits(seed, monkeypatch):
    """the compiled R1CS (general rows, integer rows with their grouped records, absorbed and free-standing boolean rows,
    runs of plane bits) on the value store of random circuits, with the integer rows forced on however few they are:
    equal to the definition on valid witnesses and on witnesses with one entry overwritten"""
    from tests.util import hostsim_run_r1cs
    monkeypatch.setenv("CW_R1CS_SMALL_ALWAYS", "1")
    rng = random.Random(31337 + seed)
    d = CircuitDesc("bn128" if seed % 3 else "bls12381")
    n_in = rng.randrange(2, 5)
    if seed % 2:
        d.set_main(random_template(d, rng, n_in, n_vals=rng.randrange(20, 70), with_components=seed % 4 == 1))
    else:
        d.set_main(bit_logic_template(d, rng, n_in))
    edges = edge_values(d.q)
    small = lambda: rng.choice([0, 1, 1, 2, 3, 255, 65535, 65536, rng.randrange(1 << 16), rng.randrange(1 << 33)])
    ins = [{"x": [small() if rng.random() < 0.7 else rand_input(rng, d.q, edges) for _ in range(n_in)]} for _ in range(6)]
    n_small = n_wide = 0
    for flags in (0, CW_FLAG_COMPACT):
        fc, fp, cnt = hostsim_run_r1cs(d, ins, flags)
        assert fc.tolist() == fp.tolist(), (seed, flags)
        if seed % 2:
            assert fc.tolist() == [-1] * len(ins)
        n_small += cnt[0]
        n_wide += cnt[1]
        W = len(hostsim_run(d, ins[:1], flags)[3])
        for trial in range(8):
            wire = rng.randrange(1, W)
            val = rng.choice([0, 1, 2, 65535, 65536, 1 << 40, d.q - 1, rng.randrange(d.q)])
            fc, fp, _ = hostsim_run_r1cs(d, ins[:2], flags, tamper=(wire, val))
            assert fc.tolist() == fp.tolist(), (seed, flags, wire, val)
    if seed % 2 == 0:
        assert n_small >= 8
