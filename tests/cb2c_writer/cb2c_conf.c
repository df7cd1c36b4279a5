/* TEST INFRASTRUCTURE: a stand-alone writer of one .cb2c file, written from docs/CB2C.md alone (no code shared with the
 * Python DSL or the library).  It describes the circuit
 *
 *     template Multiplier2() { signal input a, b; signal output c; c <== a*b; }
 *     template Conf() { signal input x, y; signal output bits[2], p;
 *         lc = 0; for k in 0..2 { bits[k] <-- (x >> k) & 1; bits[k]*(bits[k]-1) === 0; lc += bits[k]*2^k; } lc === x;
 *         component m = Multiplier2(); m.a <== x; m.b <== y; p <== m.c + 1; }
 *
 * tests/test_cb2c_spec_cpu.py checks that the bytes equal what the DSL writes for the same circuit and that the file
 * loads, lowers and computes the expected witness.   usage: cb2c_conf out.cb2c [sym|iomap|iomap+sym|log|log+iomap+sym]   (sym: with the symbols
 * section; iomap: with the io-map section, as if both templates sat in a component array of mixed templates) */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static FILE *f;
static void u32(uint32_t v) { fwrite(&v, 4, 1, f); }
static void u64(uint64_t v) { fwrite(&v, 8, 1, f); }
static void str(const char *s) {
    uint32_t n = (uint32_t)strlen(s), pad = (4 - n % 4) % 4;
    u32(n);
    fwrite(s, 1, n, f);
    fwrite("\0\0\0", 1, pad, f);
}
enum { NONE = 0, OWN = 1, SUB = 2, CONST = 3, TMP = 4, ONE = 5 };
static uint64_t ref(int kind, int sub, uint32_t idx) { return ((uint64_t)kind << 56) | ((uint64_t)sub << 32) | idx; }
static void op(int code, uint64_t d, uint64_t a, uint64_t b, uint64_t c) { u64(code); u64(d); u64(a); u64(b); u64(c); }
enum { MUL = 1, ADD = 3, SUB_ = 4, SHR = 9, BAND = 20, COPY = 24, ASSERT_EQ = 27 };
/* bn128 prime, 4 x u64 little-endian limbs, and q - 1, q - 2 */
static const uint64_t Q[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static void constant_small(uint64_t v) { u64(v); u64(0); u64(0); u64(0); }
static void constant_q_minus(uint64_t k) { u64(Q[0] - k); u64(Q[1]); u64(Q[2]); u64(Q[3]); }
static void term(uint64_t r, uint32_t cid) { u64(r); u64(cid); }

int main(int argc, char **argv) {
    if (argc < 2 || !(f = fopen(argv[1], "wb"))) return 1;
    /* constant ids: 0, 1, 2, q-1, q-2 */
    enum { C0 = 0, C1 = 1, C2 = 2, CM1 = 3, CM2 = 4 };
    fwrite("CB2C", 1, 4, f);
    u32(1); u32(0); u32(5); u32(2); u32(1); u32(2); u32(0); /* version, prime, n_consts, n_templates, main, n_names, n_funcs */
    constant_small(0); constant_small(1); constant_small(2); constant_q_minus(1); constant_q_minus(2);

    /* ---- template 0: Multiplier2 (signals: c=0 | a=1, b=2) ---- */
    str("Multiplier2");
    u32(1); u32(2); u32(0); u32(0); u32(1); u32(2); u32(1); u32(3); /* n_out n_in n_inter n_sub n_tmp n_ops n_cons n_terms */
    op(MUL, ref(TMP, 0, 0), ref(OWN, 0, 1), ref(OWN, 0, 2), 0);
    op(COPY, ref(OWN, 0, 0), ref(TMP, 0, 0), 0, 0);
    /* c <== a*b in the reference's normal form: (-a) * b - (-c) = 0 */
    u64(1); term(ref(OWN, 0, 1), CM1);
    u64(1); term(ref(OWN, 0, 2), C1);
    u64(1); term(ref(OWN, 0, 0), CM1);

    /* ---- template 1: Conf (signals: bits[0]=0 bits[1]=1 p=2 | x=3 y=4; sub 0 = Multiplier2) ---- */
    str("Conf");
    const int with_log = argc > 2 && strstr(argv[2], "log") != NULL;   /* ... and `log("p =", p);` at the end of Conf */
    u32(3); u32(2); u32(0); u32(1); u32(12); u32(with_log ? 22 : 20); u32(6); u32(16);
    u32(0); /* subs */
    for (int k = 0; k < 2; ++k) {
        uint32_t t0 = k ? 5 : 0, t1 = k ? 6 : 1, t2 = k ? 7 : 2, t3 = k ? 8 : 3, t4 = k ? 9 : 4;
        op(SHR, ref(TMP, 0, t0), ref(OWN, 0, 3), ref(CONST, 0, k ? C1 : C0), 0);
        op(BAND, ref(TMP, 0, t1), ref(TMP, 0, t0), ref(CONST, 0, C1), 0);
        op(COPY, ref(OWN, 0, k), ref(TMP, 0, t1), 0, 0);
        op(SUB_, ref(TMP, 0, t3), ref(OWN, 0, k), ref(CONST, 0, C1), 0);
        op(MUL, ref(TMP, 0, t2), ref(OWN, 0, k), ref(TMP, 0, t3), 0);
        op(ASSERT_EQ, 0, ref(TMP, 0, t2), ref(CONST, 0, C0), 0);
        op(MUL, ref(TMP, 0, t4), ref(OWN, 0, k), ref(CONST, 0, k ? C2 : C1), 0);
    }
    op(ADD, ref(TMP, 0, 10), ref(TMP, 0, 4), ref(TMP, 0, 9), 0);
    op(ASSERT_EQ, 0, ref(TMP, 0, 10), ref(OWN, 0, 3), 0);
    op(COPY, ref(SUB, 0, 1), ref(OWN, 0, 3), 0, 0);   /* m.a <== x */
    op(COPY, ref(SUB, 0, 2), ref(OWN, 0, 4), 0, 0);   /* m.b <== y : the last input, Multiplier2 runs here */
    op(ADD, ref(TMP, 0, 11), ref(SUB, 0, 0), ref(CONST, 0, C1), 0);
    op(COPY, ref(OWN, 0, 2), ref(TMP, 0, 11), 0, 0);
    if (with_log) {   /* one LOG op per argument: a string (b = string id), then the signal p with c = 1: last argument */
        op(29, 0, 0, ref(NONE, 0, 0), ref(NONE, 0, 0));
        op(29, 0, ref(OWN, 0, 2), 0, ref(NONE, 0, 1));
    }
    /* constraints; terms in ascending reference order */
    for (int k = 0; k < 2; ++k) { /* bits[k] * (bits[k] - 1) = 0 */
        u64(1); term(ref(OWN, 0, k), C1);
        u64(2); term(ref(OWN, 0, k), C1); term(ref(ONE, 0, 0), CM1);
        u64(0);
    }
    u64(0); u64(0); u64(3); term(ref(OWN, 0, 0), CM1); term(ref(OWN, 0, 1), CM2); term(ref(OWN, 0, 3), C1);  /* x - bits0 - 2 bits1 */
    u64(0); u64(0); u64(2); term(ref(OWN, 0, 3), C1); term(ref(SUB, 0, 1), CM1);                              /* x - m.a */
    u64(0); u64(0); u64(2); term(ref(OWN, 0, 4), C1); term(ref(SUB, 0, 2), CM1);                              /* y - m.b */
    u64(0); u64(0); u64(3); term(ref(OWN, 0, 2), CM1); term(ref(SUB, 0, 0), C1); term(ref(ONE, 0, 0), C1);    /* m.c + 1 - p */

    /* ---- main-input names: global ids (one = 0, outputs 1..3, inputs 4, 5) ---- */
    str("x"); u32(4); u32(1);
    str("y"); u32(5); u32(1);
    /* ---- no functions; optional symbols section: per template the own signals in numbering order, then the sub-components ---- */
    if (with_log) {   /* optional string table of log(): before the io map and the symbols */
        fwrite("LOGS", 1, 4, f);
        u32(1);
        str("p =");
    }
    if (argc > 2 && strstr(argv[2], "iomap")) {
        /* optional io-map section: entries in ascending template order; per signal {offset, #dims, dims.., element size, bus id} */
        fwrite("IOMP", 1, 4, f);
        u32(2);
        u32(0); u32(3);                       /* Multiplier2: c, a, b */
        u32(0); u32(0); u32(1); u32(0);
        u32(1); u32(0); u32(1); u32(0);
        u32(2); u32(0); u32(1); u32(0);
        u32(1); u32(4);                       /* Conf: bits[2], p, x, y */
        u32(0); u32(1); u32(2); u32(1); u32(0);
        u32(2); u32(0); u32(1); u32(0);
        u32(3); u32(0); u32(1); u32(0);
        u32(4); u32(0); u32(1); u32(0);
    }
    if (argc > 2 && strstr(argv[2], "sym")) {
        fwrite("SYMS", 1, 4, f);
        str("c"); str("a"); str("b");                                       /* Multiplier2 */
        str("bits[0]"); str("bits[1]"); str("p"); str("x"); str("y"); str("m");  /* Conf: 5 signals, 1 sub-component */
    }
    fclose(f);
    return 0;
}
