"""Non-native big-integer arithmetic circuits (k limbs of n bits), in the style of 0xPARC
circom-ecdsa's bigint.circom: BigMultNoCarry (polynomial identity at 2k-1 points), limb range
checks through Num2Bits, CheckCarryToZero with range-checked carries, BigMultModP over the
secp256k1 base field — and `ecdsa_scale`, the ~1M-constraint BN254 workload of BASELINE.json.

circom-ecdsa is NOT in the reference tree and cannot be reproduced here (no compiler, no
sources); `ecdsa_scale` is a documented synthetic of the same size and operation mix (64-bit
limb products, `<--` hints computed with shifts / masks / comparisons, thousands of Num2Bits
range checks).  circom-ecdsa computes the quotient / remainder hints with data-dependent
functions (`long_div`); here they are straight-line, using the special form of the secp256k1
prime p = 2^256 - 2^32 - 977:  X = Xhi*2^256 + Xlo = Xhi*p + (Xhi*delta + Xlo).
Known-answer side of the tests: python integers.
"""
from __future__ import annotations

from ..circuit import CircuitDesc, Template
from .basic import num2bits

SECP256K1_P = 2**256 - 2**32 - 977
N, KL = 64, 4            # limb bits, limbs
M64 = (1 << 64) - 1


def _limbs(x: int, k: int = KL):
    return [(x >> (N * i)) & M64 for i in range(k)]


def _split(t: Template, x):
    """x (canonical integer < 2^190) -> (x mod 2^64, x >> 64) as `<--` style values"""
    return x & M64, x >> N


def fn_modp_limb(d: CircuitDesc):
    """function long_div_p(X[8]) -> var out[9]: quotient (5 limbs) and remainder (4 limbs) of the 512-bit X (eight
    64-bit limbs) by the secp256k1 prime, returned as ONE array - the job of circom-ecdsa's `long_div` hint
    (bigint_func.circom), written like it: `var` arrays indexed by run-time loop counters, a folding loop whose trip
    count depends on the data, a limb-wise comparison with early exit, a conditional subtraction."""
    from ..circuit import Function
    delta = (1 << 32) + 977
    p_l = _limbs(SECP256K1_P)

    def build(f: Function):
        X = f.param_array(0, 2 * KL)     # `function long_div_p(X[8])`: the parameter array is indexed in place
        Q = f.array(2 * KL + 1)          # out[0..4] = quotient, out[5..8] = remainder: one contiguous `var out[9]`
        R = Q + KL + 1
        Pp = f.array(KL)
        for i in range(KL):
            f.store(Pp, f.var(i), p_l[i])
        four = KL
        # Q = X_hi ; R = X_lo + X_hi * delta (with carries), h = carry out
        i = f.var(0)
        c = f.var(0)
        f.loop_begin()
        f.loop_break_if_zero(i.lt(four))
        hi = f.load(X, i + four)
        f.store(Q, i, hi)
        s = f.load(X, i) + hi * delta + c
        f.store(R, i, s & M64)
        f.set(c, s >> N)
        f.set(i, i + 1)
        f.loop_end()
        f.store(Q, f.var(four), 0)
        h = f.var(c)
        # while (h != 0): R += h * delta, Q += h   (X = Q*p + R is kept; at most three rounds)
        f.loop_begin()
        f.loop_break_if_zero(h.neq(0))
        f.set(c, h * delta)
        qc = f.var(h)
        f.set(i, 0)
        f.loop_begin()
        f.loop_break_if_zero(i.lt(four))
        s2 = f.load(R, i) + c
        f.store(R, i, s2 & M64)
        f.set(c, s2 >> N)
        f.set(i, i + 1)
        f.loop_end()
        f.set(h, c)
        f.set(i, 0)
        f.loop_begin()                                  # Q += qc with carries
        f.loop_break_if_zero(i.leq(four))
        s3 = f.load(Q, i) + qc
        f.store(Q, i, s3 & M64)
        f.set(qc, s3 >> N)
        f.set(i, i + 1)
        f.loop_end()
        f.loop_end()
        # R >= p ?  compare from the top limb down, leave at the first difference
        ge = f.var(1)
        j = f.var(four)
        f.loop_begin()
        f.loop_break_if_zero(j.neq(0))
        f.set(j, j - 1)
        rj = f.load(R, j)
        pj = f.load(Pp, j)
        f.if_begin(rj.neq(pj))
        f.set(ge, rj.gt(pj))
        f.set(j, 0)
        f.if_end()
        f.loop_end()
        f.if_begin(ge)
        b = f.var(0)                                    # R -= p
        f.set(i, 0)
        f.loop_begin()
        f.loop_break_if_zero(i.lt(four))
        s4 = f.load(R, i) + (1 << N) - f.load(Pp, i) - b
        f.store(R, i, s4 & M64)
        f.set(b, 1 - (s4 >> N))
        f.set(i, i + 1)
        f.loop_end()
        one = f.var(1)                                  # Q += 1
        f.set(i, 0)
        f.loop_begin()
        f.loop_break_if_zero(i.leq(four))
        s5 = f.load(Q, i) + one
        f.store(Q, i, s5 & M64)
        f.set(one, s5 >> N)
        f.set(i, i + 1)
        f.loop_end()
        f.if_end()
        f.ret_array(Q, 2 * KL + 1)
    return d.function("long_div_p", 2 * KL, build)


def big_mult_mod_p(d: CircuitDesc, hints: str = "inline") -> Template:
    """out = a*b mod p for 4x64-bit limb operands (limbs of a, b must be < 2^64).  hints = "inline": quotient and
    remainder hints as straight-line shifts and masks; "functions": `var qr[9] = long_div_p(P)`, one array-valued
    call (the circom-ecdsa style: hints computed by functions with data-dependent control flow)."""
    fmod = fn_modp_limb(d) if hints == "functions" else None
    n2b64 = num2bits(d, N)
    n2b_carry = num2bits(d, 72)
    p_l = _limbs(SECP256K1_P)
    delta = (1 << 32) + 977

    def build(t: Template):
        a = t.input("a", KL)
        b = t.input("b", KL)
        out = t.output("out", KL)
        quo = t.signal("q", KL + 1)
        prod = t.signal("prod", 2 * KL - 1)      # no-carry product coefficients of a*b
        qp = t.signal("qp", 2 * KL)             # no-carry product coefficients of q*p
        carry = t.signal("carry", 2 * KL - 1)

        # ---- a*b as polynomial coefficients (BigMultNoCarry: hints + identity at 2k-1 points)
        coef = []
        for m in range(2 * KL - 1):
            acc = t.const(0)
            for i in range(KL):
                j = m - i
                if 0 <= j < KL:
                    acc = acc + a[i] * b[j]
            coef.append(acc)
            t.assign(prod[m], acc)
        for x in range(2 * KL - 1):
            pa = t.const(0)
            pb = t.const(0)
            pc = t.const(0)
            for i in range(KL):
                pa = pa + a[i] * (x ** i)
                pb = pb + b[i] * (x ** i)
            for m in range(2 * KL - 1):
                pc = pc + prod[m] * (x ** m)
            t.constrain(pa * pb, pc)

        # ---- hint: (quotient, remainder) of the 512-bit product by p, straight-line -------------
        # normalise to 8 proper limbs
        P = []
        c = t.const(0)
        for m in range(2 * KL - 1):
            s = prod[m] + c
            lo, c = _split(t, s)
            P.append(lo)
        P.append(c)                                # limb 7 (< 2^64)
        if fmod is not None:
            qr = t.call_array(fmod, P, 2 * KL + 1)     # var qr[9] = long_div_p(P): ONE call
            for i in range(KL):
                t.assign(out[i], qr[KL + 1 + i])
            for i in range(KL + 1):
                t.assign(quo[i], qr[i])
        # first fold: R1 = Xlo + Xhi*delta (5 limbs), Q1 = Xhi
        R = []
        c = t.const(0)
        for i in range(KL):
            s = P[i] + P[KL + i] * delta + c
            lo, c = _split(t, s)
            R.append(lo)
        h2 = c                                     # < 2^34
        # second fold: R2 = R1lo + h2*delta, Q2 = h2
        c = h2 * delta
        R2 = []
        for i in range(KL):
            s = R[i] + c
            lo, c = _split(t, s)
            R2.append(lo)
        h3 = c                                     # 0 or 1; if 1 the low part is tiny: fold once more
        c = h3 * delta
        R3 = []
        for i in range(KL):
            s = R2[i] + c
            lo, c = _split(t, s)
            R3.append(lo)
        # quotient so far: Xhi + h2 + h3
        Q = []
        c = h2 + h3
        for i in range(KL):
            s = P[KL + i] + c
            lo, c = _split(t, s)
            Q.append(lo)
        Q.append(c)
        # final conditional subtraction of p (R3 < 2^256 < 2p)
        borrow = t.const(0)
        D = []
        for i in range(KL):
            s = R3[i] + (1 << N) - p_l[i] - borrow
            lo, hi = _split(t, s)
            D.append(lo)
            borrow = 1 - hi
        ge = 1 - borrow                            # 1 iff R3 >= p
        rem = [t.select(ge, D[i], R3[i]) for i in range(KL)]
        c = ge
        Qf = []
        for i in range(KL + 1):
            s = Q[i] + c
            lo, c = _split(t, s)
            Qf.append(lo)
        if fmod is None:
            for i in range(KL):
                t.assign(out[i], rem[i])
            for i in range(KL + 1):
                t.assign(quo[i], Qf[i])

        # ---- range checks: limbs of out and q are 64-bit (Num2Bits) --------------------------
        for i in range(KL):
            c1 = t.component("rc_out[%d]" % i, n2b64)
            t.assign_constrained(c1["in"], out[i])
        for i in range(KL + 1):
            c2 = t.component("rc_q[%d]" % i, n2b64)
            t.assign_constrained(c2["in"], quo[i])

        # ---- q*p as polynomial coefficients (p constant: linear constraints) -------------------
        for m in range(2 * KL):
            acc = t.const(0)
            for i in range(KL + 1):
                j = m - i
                if 0 <= j < KL:
                    acc = acc + quo[i] * p_l[j]
            t.assign_constrained(qp[m], acc)

        # ---- a*b - q*p - out == 0 limb-wise with carries (CheckCarryToZero) ---------------------
        # in[m] = prod[m] - qp[m] - out[m]; in[m] + carry[m-1] = carry[m] * 2^64 ; carries are signed
        # and bounded by 2^70: carry + 2^71 is range-checked to 72 bits.
        prev = t.const(0)
        OFF = 1 << 71
        for m in range(2 * KL):
            lhs = (prod[m] if m < 2 * KL - 1 else t.const(0)) - qp[m] - (out[m] if m < KL else t.const(0)) + prev
            if m < 2 * KL - 1:
                # hint: exact signed division by 2^64, computed on the shifted non-negative value
                cv = ((lhs + (OFF << N)) >> N) - OFF
                t.assign(carry[m], cv)
                t.constrain(lhs, carry[m] * (1 << N))
                rc = t.component("rc_carry[%d]" % m, n2b_carry)
                t.assign_constrained(rc["in"], carry[m] + OFF)
                prev = carry[m]
            else:
                t.constrain(lhs, 0)
    return d.template("BigMultModP" if fmod is None else "BigMultModPfn", (N, KL), build)


def ecdsa_scale(d: CircuitDesc, lanes: int = 8, steps: int = 132, hints: str = "inline") -> Template:
    """`lanes` independent chains of `steps` modular multiplications over the secp256k1 base field
    (alternating squarings and multiplications, like the field operations of a double-and-add
    ladder).  8 x 132 gives ~1.0M R1CS constraints."""
    mm = big_mult_mod_p(d, hints)
    n2b64 = num2bits(d, N)

    def build(t: Template):
        a = t.input("a", lanes * KL)
        b = t.input("b", lanes * KL)
        out = t.output("out", lanes * KL)
        for l in range(lanes):
            for i in range(KL):                      # inputs are range checked
                r1 = t.component("rca[%d][%d]" % (l, i), n2b64)
                t.assign_constrained(r1["in"], a[l * KL + i])
                r2 = t.component("rcb[%d][%d]" % (l, i), n2b64)
                t.assign_constrained(r2["in"], b[l * KL + i])
            x = [a[l * KL + i] for i in range(KL)]
            y = [b[l * KL + i] for i in range(KL)]
            for s in range(steps):
                c = t.component("mm[%d][%d]" % (l, s), mm)
                other = x if s % 2 == 0 else y
                for i in range(KL):
                    t.assign_constrained(c["a", i], x[i])
                    t.assign_constrained(c["b", i], other[i])
                x = [c["out", i] for i in range(KL)]
            for i in range(KL):
                t.assign_constrained(out[l * KL + i], x[i])
    return d.template("EcdsaScale" if hints == "inline" else "EcdsaScaleFn", (lanes, steps), build)


def ecdsa_scale_expected(a_vals, b_vals, lanes: int = 8, steps: int = 132):
    """python-int model of `ecdsa_scale` outputs (limbs), for known-answer tests"""
    out = []
    for l in range(lanes):
        x = sum(a_vals[l * KL + i] << (N * i) for i in range(KL))
        y = sum(b_vals[l * KL + i] << (N * i) for i in range(KL))
        for s in range(steps):
            x = (x * (x if s % 2 == 0 else y)) % SECP256K1_P
        out += _limbs(x)
    return out
