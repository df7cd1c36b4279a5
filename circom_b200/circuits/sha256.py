"""SHA-256 circuits in the style of circomlib's sha256 (bit-decomposed words, BinSum adders with
`<--` bit hints, Xor3 / Ch / Maj gadgets): Sha256compression and Sha256(nBits).

circomlib is not in the reference tree: these templates are authored from FIPS 180-4 and follow
circomlib's gadget structure; rotations / shifts are pure re-indexing (what the reference's
constraint simplifier leaves of RotR / ShR).  Known-answer side of the tests: hashlib.sha256.
Bit conventions: state words `hin` are LSB-first per 32-bit word, message bits `inp` and digest
bits `out` are MSB-first per word (as in circomlib's sha256compression.circom).
"""
from __future__ import annotations

from ..circuit import CircuitDesc, Template

K = [
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2,
]
H0 = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]


def _rotr(x, r):  # x: list of 32 bit-exprs, LSB first
    return [x[(k + r) % 32] for k in range(32)]


def _shr(t, x, r):
    return [x[k + r] if k + r < 32 else t.const(0) for k in range(32)]


def _xor3_bits(t: Template, name: str, a, b, c):
    """out = a ^ b ^ c per bit, circomlib Xor3: mid = b*c; out = a*(1-2b-2c+4mid) + b + c - 2mid"""
    out = []
    for k in range(32):
        if c[k].const is not None and c[k].const == 0:
            o = t.signal("%s_o[%d]" % (name, k))
            t.assign_constrained(o, a[k] * (1 - 2 * b[k]) + b[k])
        else:
            mid = t.signal("%s_m[%d]" % (name, k))
            t.assign_constrained(mid, b[k] * c[k])
            o = t.signal("%s_o[%d]" % (name, k))
            t.assign_constrained(o, a[k] * (1 - 2 * b[k] - 2 * c[k] + 4 * mid) + b[k] + c[k] - 2 * mid)
        out.append(o)
    return out


def bin_sum(d: CircuitDesc, n: int, ops: int) -> Template:
    """circomlib BinSum(n, ops): bits of the sum of `ops` n-bit numbers"""
    nout = ((2 ** n - 1) * ops).bit_length()

    def build(t: Template):
        ins = [t.input("in%d" % j, n) for j in range(ops)]
        out = t.output("out", nout)
        lin = t.const(0)
        for j in range(ops):
            e2 = 1
            for k in range(n):
                lin = lin + ins[j][k] * e2
                e2 *= 2
        lout = t.const(0)
        e2 = 1
        for k in range(nout):
            t.assign(out[k], (lin >> k) & 1)
            t.constrain(out[k] * (out[k] - 1), 0)
            lout = lout + out[k] * e2
            e2 *= 2
        t.constrain(lin, lout)
    return d.template("BinSum", (n, ops), build)


def small_sigma(d: CircuitDesc, ra: int, rb: int, rc: int) -> Template:
    def build(t: Template):
        i = t.input("in", 32)
        out = t.output("out", 32)
        x = _xor3_bits(t, "x", _rotr(i, ra), _rotr(i, rb), _shr(t, i, rc))
        for k in range(32):
            t.assign_constrained(out[k], x[k])
    return d.template("SmallSigma", (ra, rb, rc), build)


def big_sigma(d: CircuitDesc, ra: int, rb: int, rc: int) -> Template:
    def build(t: Template):
        i = t.input("in", 32)
        out = t.output("out", 32)
        x = _xor3_bits(t, "x", _rotr(i, ra), _rotr(i, rb), _rotr(i, rc))
        for k in range(32):
            t.assign_constrained(out[k], x[k])
    return d.template("BigSigma", (ra, rb, rc), build)


def ch(d: CircuitDesc) -> Template:
    def build(t: Template):
        a, b, c = t.input("a", 32), t.input("b", 32), t.input("c", 32)
        out = t.output("out", 32)
        for k in range(32):
            t.assign_constrained(out[k], a[k] * (b[k] - c[k]) + c[k])
    return d.template("Ch_t", (32,), build)


def maj(d: CircuitDesc) -> Template:
    def build(t: Template):
        a, b, c = t.input("a", 32), t.input("b", 32), t.input("c", 32)
        out = t.output("out", 32)
        mid = t.signal("mid", 32)
        for k in range(32):
            t.assign_constrained(mid[k], b[k] * c[k])
            t.assign_constrained(out[k], a[k] * (b[k] + c[k] - 2 * mid[k]) + mid[k])
    return d.template("Maj_t", (32,), build)


def t1(d: CircuitDesc) -> Template:
    bs, chh, s5 = big_sigma(d, 6, 11, 25), ch(d), bin_sum(d, 32, 5)

    def build(t: Template):
        h, e, f, g = t.input("h", 32), t.input("e", 32), t.input("f", 32), t.input("g", 32)
        k, w = t.input("k", 32), t.input("w", 32)
        out = t.output("out", 32)
        c_ch = t.component("ch", chh)
        c_bs = t.component("bigsigma1", bs)
        c_sum = t.component("sum", s5)
        for i in range(32):
            t.assign_constrained(c_bs["in", i], e[i])
            t.assign_constrained(c_ch["a", i], e[i])
            t.assign_constrained(c_ch["b", i], f[i])
            t.assign_constrained(c_ch["c", i], g[i])
        for i in range(32):
            t.assign_constrained(c_sum["in0", i], h[i])
            t.assign_constrained(c_sum["in1", i], c_bs["out", i])
            t.assign_constrained(c_sum["in2", i], c_ch["out", i])
            t.assign_constrained(c_sum["in3", i], k[i])
            t.assign_constrained(c_sum["in4", i], w[i])
        for i in range(32):
            t.assign_constrained(out[i], c_sum["out", i])
    return d.template("T1", (), build)


def t2(d: CircuitDesc) -> Template:
    bs, mj, s2 = big_sigma(d, 2, 13, 22), maj(d), bin_sum(d, 32, 2)

    def build(t: Template):
        a, b, c = t.input("a", 32), t.input("b", 32), t.input("c", 32)
        out = t.output("out", 32)
        c_bs = t.component("bigsigma0", bs)
        c_mj = t.component("maj", mj)
        c_sum = t.component("sum", s2)
        for i in range(32):
            t.assign_constrained(c_bs["in", i], a[i])
            t.assign_constrained(c_mj["a", i], a[i])
            t.assign_constrained(c_mj["b", i], b[i])
            t.assign_constrained(c_mj["c", i], c[i])
        for i in range(32):
            t.assign_constrained(c_sum["in0", i], c_bs["out", i])
            t.assign_constrained(c_sum["in1", i], c_mj["out", i])
        for i in range(32):
            t.assign_constrained(out[i], c_sum["out", i])
    return d.template("T2", (), build)


def sigma_plus(d: CircuitDesc) -> Template:
    s0, s1, s4 = small_sigma(d, 7, 18, 3), small_sigma(d, 17, 19, 10), bin_sum(d, 32, 4)

    def build(t: Template):
        in2, in7, in15, in16 = t.input("in2", 32), t.input("in7", 32), t.input("in15", 32), t.input("in16", 32)
        out = t.output("out", 32)
        c1 = t.component("sigma1", s1)
        c0 = t.component("sigma0", s0)
        cs = t.component("sum", s4)
        for i in range(32):
            t.assign_constrained(c1["in", i], in2[i])
            t.assign_constrained(c0["in", i], in15[i])
        for i in range(32):
            t.assign_constrained(cs["in0", i], c1["out", i])
            t.assign_constrained(cs["in1", i], in7[i])
            t.assign_constrained(cs["in2", i], c0["out", i])
            t.assign_constrained(cs["in3", i], in16[i])
        for i in range(32):
            t.assign_constrained(out[i], cs["out", i])
    return d.template("SigmaPlus", (), build)


def _compression_body(t: Template, d: CircuitDesc, hin, inp, tag: str = ""):
    """64 rounds over state bits `hin` (8 words, LSB first) and message bits `inp` (MSB first per word);
    returns the 8 output words (LSB first) as lists of bit expressions."""
    sp, tt1, tt2, s2 = sigma_plus(d), t1(d), t2(d), bin_sum(d, 32, 2)
    w = []
    for r in range(64):
        if r < 16:
            w.append([inp[r * 32 + 31 - k] for k in range(32)])
        else:
            c = t.component("%ssigmaPlus[%d]" % (tag, r - 16), sp)
            for k in range(32):
                t.assign_constrained(c["in2", k], w[r - 2][k])
                t.assign_constrained(c["in7", k], w[r - 7][k])
                t.assign_constrained(c["in15", k], w[r - 15][k])
                t.assign_constrained(c["in16", k], w[r - 16][k])
            w.append([c["out", k] for k in range(32)])
    st = [[hin[i * 32 + k] for k in range(32)] for i in range(8)]
    a, b, c_, dd, e, f, g, h = st
    for r in range(64):
        c1 = t.component("%st1[%d]" % (tag, r), tt1)
        c2 = t.component("%st2[%d]" % (tag, r), tt2)
        for k in range(32):
            t.assign_constrained(c1["h", k], h[k])
            t.assign_constrained(c1["e", k], e[k])
            t.assign_constrained(c1["f", k], f[k])
            t.assign_constrained(c1["g", k], g[k])
            t.assign_constrained(c1["k", k], (K[r] >> k) & 1)
            t.assign_constrained(c1["w", k], w[r][k])
            t.assign_constrained(c2["a", k], a[k])
            t.assign_constrained(c2["b", k], b[k])
            t.assign_constrained(c2["c", k], c_[k])
        se = t.component("%ssume[%d]" % (tag, r), s2)
        sa = t.component("%ssuma[%d]" % (tag, r), s2)
        for k in range(32):
            t.assign_constrained(se["in0", k], dd[k])
            t.assign_constrained(se["in1", k], c1["out", k])
            t.assign_constrained(sa["in0", k], c1["out", k])
            t.assign_constrained(sa["in1", k], c2["out", k])
        h, g, f = g, f, e
        e = [se["out", k] for k in range(32)]
        dd, c_, b = c_, b, a
        a = [sa["out", k] for k in range(32)]
    fin = []
    for i, wd in enumerate([a, b, c_, dd, e, f, g, h]):
        fs = t.component("%sfsum[%d]" % (tag, i), s2)
        for k in range(32):
            t.assign_constrained(fs["in0", k], st[i][k])
            t.assign_constrained(fs["in1", k], wd[k])
        fin.append([fs["out", k] for k in range(32)])
    return fin


def sha256_compression(d: CircuitDesc) -> Template:
    def build(t: Template):
        hin = t.input("hin", 256)
        inp = t.input("inp", 512)
        out = t.output("out", 256)
        fin = _compression_body(t, d, hin, inp)
        for i in range(8):
            for k in range(32):
                t.assign_constrained(out[i * 32 + 31 - k], fin[i][k])
    return d.template("Sha256compression", (), build)


def sha256(d: CircuitDesc, n_bits: int) -> Template:
    """Sha256(nBits): padding + chained compressions (circomlib sha256.circom)"""
    n_blocks = (n_bits + 64) // 512 + 1

    def build(t: Template):
        i = t.input("in", n_bits)
        out = t.output("out", 256)
        padded = list(i) + [t.const(1)] + [t.const(0)] * (n_blocks * 512 - n_bits - 1 - 64)
        padded += [t.const((n_bits >> (63 - k)) & 1) for k in range(64)]
        state = [t.const((H0[j] >> k) & 1) for j in range(8) for k in range(32)]
        for blk in range(n_blocks):
            fin = _compression_body(t, d, state, padded[blk * 512:(blk + 1) * 512], tag="b%d_" % blk)
            state = [fin[j][k] for j in range(8) for k in range(32)]
        for j in range(8):
            for k in range(32):
                t.assign_constrained(out[j * 32 + 31 - k], state[j * 32 + k])
    return d.template("Sha256", (n_bits,), build)
