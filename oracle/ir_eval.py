"""Pure-Python big-int execution of a circuit description (TEST INFRASTRUCTURE).

Restates the reference runtime's execution model for generated code
(compiler/src/circuit_design/template.rs:177-472): components are created with
`inputCounter = number_of_inputs` (template.rs:204-209), a store into a
sub-component input decrements the counter and runs the sub-component when it
reaches zero (intermediate_representation/store_bucket.rs:660-734), templates
without inputs run at creation (template.rs:274-278).  Signal numbering: global
signal 0 is the constant 1 (common/calcwit.cpp:34), main starts at 1, each
component owns [signalStart, signalStart + n_own) followed by its
sub-components' blocks (create_component_bucket.rs:219-233).

Small circuits only (python loops).  Only tests may import this.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

from .field_model import Field, OPS

K_NONE, K_OWN, K_SUB, K_CONST, K_TMP, K_ONE = 0, 1, 2, 3, 4, 5


LOG_SINK: list = []   # what the log() calls of the last evaluate() printed (cleared by the caller)


class AssertFailed(Exception):
    pass


def run_function(desc, F, fn, args):
    """A function body: registers are mutable variables, loops / branches test Fr_isTrue of run-time values
    (loop_bucket.rs:76-91, branch_bucket.rs:100-122), array variables are indexed through Fr_toInt
    (compute_bucket.rs:361-363; signed view, must be a valid index)."""
    regs = [0] * fn.n_regs
    regs[:fn.n_params] = list(args)
    consts = desc.consts

    def val(r):
        if r[0] == K_TMP: return regs[r[2]]
        if r[0] == K_CONST: return consts[r[2]]
        return 0

    def to_int(v):
        v = F.val(v)
        assert -(1 << 31) <= v < (1 << 31)
        return v

    pc, steps = 0, 0
    while True:
        steps += 1
        if steps > 10_000_000:
            raise RuntimeError("function %s does not terminate" % fn.name)
        op, d, a, b, c = fn.code[pc]
        pc += 1
        if op == OPS["JMP"]:
            pc = a[2]
        elif op == OPS["JZ"]:
            if val(a) == 0:
                pc = b[2]
        elif op == OPS["RET"]:
            n = b[2] if b[0] == 0 else 0          # array return: registers a .. a+n-1 (return_bucket.rs with_size)
            if n > 1:
                return [regs[a[2] + k] for k in range(n)]
            return val(a)
        elif op == OPS["CALL"]:                  # nested call: the callee's parameters are the registers b .. b+n_params-1
            callee = desc.functions[a[2]]
            assert callee.id < fn.id, "a function may only call earlier functions"
            res = run_function(desc, F, callee, [regs[b[2] + k] for k in range(callee.n_params)])
            n_res = c[2] if c[0] == 0 else 0
            if n_res > 1:
                assert isinstance(res, list) and len(res) >= n_res
                for k in range(n_res):
                    regs[d[2] + k] = res[k]
            else:
                regs[d[2]] = res[0] if isinstance(res, list) else res
        elif op == OPS["LOADX"]:
            i = a[2] + to_int(val(b))
            assert 0 <= i < (a[2] + c[2] if c[0] == 0 and c[2] else fn.n_regs)   # operand c: extent of the array
            regs[d[2]] = regs[i]
        elif op == OPS["STOREX"]:
            i = a[2] + to_int(val(b))
            assert 0 <= i < (a[2] + d[2] if d[0] == 0 and d[2] else fn.n_regs)      # operand d: extent of the array
            regs[i] = val(c)
        else:
            regs[d[2]] = F.apply(op, val(a), val(b), val(c))


class _Comp:
    __slots__ = ("tmpl", "start", "counter", "subs", "ran")


def evaluate(desc, inputs: Dict[str, Sequence[int]], check_asserts: bool = True) -> List[int]:
    """Returns the value of every signal, indexed by global signal id."""
    F = Field(desc.prime)
    S = desc.total_signals
    sig = [None] * S
    sig[0] = 1
    consts = desc.consts

    def create(tmpl, start):
        c = _Comp()
        c.tmpl, c.start, c.counter, c.ran = tmpl, start, tmpl.n_in, False
        c.subs = [None] * len(tmpl.subs)
        return c

    def run(c):
        assert not c.ran
        c.ran = True
        t = c.tmpl
        tmp = [None] * t.n_tmp
        off = c.start + t.n_own
        for i, s in enumerate(t.subs):     # CreateCmp buckets
            c.subs[i] = create(s.tmpl, off)
            off += s.tmpl.total_signals
            if s.tmpl.n_in == 0:
                run(c.subs[i])

        def load(r):
            k = r[0]
            if k == K_OWN: v = sig[c.start + r[2]]
            elif k == K_SUB: v = sig[c.subs[r[1]].start + r[2]]
            elif k == K_CONST: v = consts[r[2]]
            elif k == K_TMP: v = tmp[r[2]]
            elif k == K_ONE: v = 1
            else: return 0
            if v is None:
                raise RuntimeError("read of unassigned value %r in %s" % (r, t.name))
            return v

        argstack = []
        for op, d, a, b, cc in t.ops:
            if op == OPS["ARG"]:
                argstack.append(load(a))
                continue
            if op == OPS["CALL"]:
                n = b[2]
                args = argstack[len(argstack) - n:]
                del argstack[len(argstack) - n:]
                res = run_function(desc, F, desc.functions[a[2]], args)
                n_res = cc[2] if cc[0] == 0 else 0
                if n_res > 1:                      # one call, n_res results in consecutive temporaries
                    assert isinstance(res, list) and len(res) >= n_res
                    for k in range(n_res):
                        tmp[d[2] + k] = res[k]
                else:
                    tmp[d[2]] = res[0] if isinstance(res, list) else res
                continue
            if op == OPS["ASSERT_EQ"]:
                if check_asserts and load(a) != load(b):
                    raise AssertFailed(t.name)
                continue
            if op == OPS["ASSERT"]:
                if check_asserts and load(a) == 0:
                    raise AssertFailed(t.name)
                continue
            if op == 29:   # LOG: one argument of a log() call (log_bucket.rs:104-162)
                if a[0] == 0:
                    text = desc.strings[b[2]]
                else:
                    text = str(load(a))
                LOG_SINK.append(text + ("\n" if cc[2] else " "))
                continue
            if op == 48:   # LOADSIG (producer level, circuit.py: Template.load_indexed): own signal array a[toInt(b)], extent cc
                iv = load(b)
                iv = iv - F.q if iv > F.half else iv           # Fr_toInt on the signed view (generic/fr.cpp:2766-2803)
                if not 0 <= iv < cc[2]:
                    raise AssertFailed(t.name)                 # the reference reads beside the array here
                tmp[d[2]] = sig[c.start + a[2] + iv]
                assert tmp[d[2]] is not None, "indexed load of a signal that is not assigned yet"
                continue
            v = F.apply(op, load(a), load(b), load(cc))
            if d[0] == K_TMP:
                tmp[d[2]] = v
            elif d[0] == K_OWN:
                g = c.start + d[2]
                assert sig[g] is None, "signal assigned twice"
                sig[g] = v
            elif d[0] == K_SUB:
                sc = c.subs[d[1]]
                g = sc.start + d[2]
                assert sig[g] is None, "signal assigned twice"
                sig[g] = v
                if sc.tmpl.n_out <= d[2] < sc.tmpl.n_out + sc.tmpl.n_in:
                    sc.counter -= 1
                    if sc.counter == 0:
                        run(sc)
        for sc in c.subs:
            assert sc.ran, "sub-component never triggered"

    main = create(desc.main, 1)
    for name, gid, n in desc.main_inputs():
        vals = inputs[name]
        if n == 1 and not isinstance(vals, (list, tuple)):
            vals = [vals]
        assert len(vals) == n, name
        for j, v in enumerate(vals):
            sig[gid + j] = int(v) % F.q
    run(main)
    return sig


def check_r1cs(desc, sig: List[int]) -> int:
    """Number of violated constraints (A.w * B.w - C.w != 0), evaluated straight from the
    description with python ints."""
    q = desc.q
    bad = 0

    def walk(t, start):
        nonlocal bad
        offs = []
        off = start + t.n_own
        for s in t.subs:
            offs.append(off)
            off += s.tmpl.total_signals

        def val(k):
            if k[0] == K_OWN: return sig[start + k[2]]
            if k[0] == K_SUB: return sig[offs[k[1]] + k[2]]
            if k[0] == K_ONE: return 1
            raise ValueError(k)

        for A, B, C in t.constraints:
            a = sum(v * val(k) for k, v in A.items()) % q
            b = sum(v * val(k) for k, v in B.items()) % q
            c = sum(v * val(k) for k, v in C.items()) % q
            if (a * b - c) % q:
                bad += 1
        for s, o in zip(t.subs, offs):
            walk(s.tmpl, o)

    walk(desc.main, 1)
    return bad
