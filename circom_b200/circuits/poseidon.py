"""Poseidon(nInputs) over BN254, structured like circomlib's (pre-optimisation) poseidon.circom:
Ark / Sigma (x^5) / Mix components, R_F = 8 full rounds, R_P from circomlib's table.

circomlib is not in the reference tree, so the round constants and the MDS matrix are derived
here with the Poseidon paper's Grain LFSR procedure (generate_parameters_grain.sage); the result
reproduces circomlibjs' published test value poseidon([1,2]) =
0x115cc0f5e7d690413df64c6b9662e9cf2a3617f2743245519e19607a4417189a (checked in tests).
"""
from __future__ import annotations

from functools import lru_cache

from ..circuit import CircuitDesc, Template, PRIMES

N_ROUNDS_P = [56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65, 70, 60, 64, 68]
N_ROUNDS_F = 8


def _grain(n, t, rf, rp):
    def bits(v, w):
        return [(v >> (w - 1 - i)) & 1 for i in range(w)]
    s = bits(1, 2) + bits(0, 4) + bits(n, 12) + bits(t, 12) + bits(rf, 10) + bits(rp, 10) + [1] * 30

    def step():
        nb = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(nb)
        return nb
    for _ in range(160):
        step()
    while True:
        nb = step()
        while nb == 0:
            step()
            nb = step()
        yield step()


@lru_cache(maxsize=None)
def poseidon_params(t: int, q: int = PRIMES["bn128"]):
    """(round constants [(R_F+R_P)*t], MDS matrix [t][t]) for width t."""
    n = q.bit_length()
    rf, rp = N_ROUNDS_F, N_ROUNDS_P[t - 2]
    g = _grain(n, t, rf, rp)

    def rb(k):
        v = 0
        for _ in range(k):
            v = (v << 1) | next(g)
        return v
    rc = []
    for _ in range((rf + rp) * t):
        v = rb(n)
        while v >= q:
            v = rb(n)
        rc.append(v)
    while True:
        rl = [rb(n) % q for _ in range(2 * t)]
        if len(set(rl)) != 2 * t:
            continue
        xs, ys = rl[:t], rl[t:]
        if any((x + y) % q == 0 for x in xs for y in ys):
            continue
        m = [[pow((xs[i] + ys[j]) % q, -1, q) for j in range(t)] for i in range(t)]
        return rc, m


def poseidon_hash(inputs, q: int = PRIMES["bn128"]) -> int:
    """plain-python Poseidon (known-answer side of the tests)"""
    t = len(inputs) + 1
    rc, m = poseidon_params(t, q)
    rf, rp = N_ROUNDS_F, N_ROUNDS_P[t - 2]
    st = [0] + [int(x) % q for x in inputs]
    for r in range(rf + rp):
        st = [(st[i] + rc[r * t + i]) % q for i in range(t)]
        if r < rf // 2 or r >= rf // 2 + rp:
            st = [pow(x, 5, q) for x in st]
        else:
            st[0] = pow(st[0], 5, q)
        st = [sum(m[i][j] * st[j] for j in range(t)) % q for i in range(t)]
    return st[0]


def sigma(d: CircuitDesc) -> Template:
    def build(t: Template):
        i = t.input("in")
        out = t.output("out")
        in2 = t.signal("in2")
        in4 = t.signal("in4")
        t.assign_constrained(in2, i * i)
        t.assign_constrained(in4, in2 * in2)
        t.assign_constrained(out, in4 * i)
    return d.template("Sigma", (), build)


def ark(d: CircuitDesc, tw: int, r: int) -> Template:
    rc, _ = poseidon_params(tw, d.q)

    def build(t: Template):
        i = t.input("in", tw)
        out = t.output("out", tw)
        for k in range(tw):
            t.assign_constrained(out[k], i[k] + rc[r + k])
    return d.template("Ark", (tw, r), build)


def mix(d: CircuitDesc, tw: int) -> Template:
    _, m = poseidon_params(tw, d.q)

    def build(t: Template):
        i = t.input("in", tw)
        out = t.output("out", tw)
        for a in range(tw):
            lc = t.const(0)
            for b in range(tw):
                lc = lc + i[b] * m[a][b]
            t.assign_constrained(out[a], lc)
    return d.template("Mix", (tw,), build)


def poseidon(d: CircuitDesc, n_inputs: int) -> Template:
    tw = n_inputs + 1
    rf, rp = N_ROUNDS_F, N_ROUNDS_P[tw - 2]
    sg, mx = sigma(d), mix(d, tw)

    def build(t: Template):
        inputs = t.input("inputs", n_inputs)
        out = t.output("out")
        state = [t.const(0)] + list(inputs)
        for r in range(rf + rp):
            a = t.component("ark[%d]" % r, ark(d, tw, r * tw))
            for k in range(tw):
                t.assign_constrained(a["in", k], state[k])
            cur = [a["out", k] for k in range(tw)]
            if r < rf // 2 or r >= rf // 2 + rp:
                for k in range(tw):
                    s = t.component("sigmaF[%d][%d]" % (r, k), sg)
                    t.assign_constrained(s["in"], cur[k])
                    cur[k] = s["out"]
            else:
                s = t.component("sigmaP[%d]" % r, sg)
                t.assign_constrained(s["in"], cur[0])
                cur[0] = s["out"]
            m = t.component("mix[%d]" % r, mx)
            for k in range(tw):
                t.assign_constrained(m["in", k], cur[k])
            state = [m["out", k] for k in range(tw)]
        t.assign_constrained(out, state[0])
    return d.template("Poseidon", (n_inputs,), build)
