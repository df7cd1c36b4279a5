// TEST INFRASTRUCTURE: C entry point over the REFERENCE's goldilocks field library, which is header-only
// (code_producers/src/c_elements/goldilocks/fr.hpp, included where it lies: -I <reference>/.../goldilocks).
// Built by oracle/build_ref.py into oracle/_ref/libfr_goldilocks.so; tests/test_oracle_ref.py pins the python
// model (oracle/field_model.py, Field("goldilocks")) against it.  Opcodes: OperatorType order
// (compiler/src/intermediate_representation/compute_bucket.rs:7-34), as everywhere in this repository.
#include <stdint.h>
typedef unsigned int uint;
#include "fr.hpp"

extern "C" int gl_apply(int op, uint64_t a, uint64_t b, uint64_t *r) {
    switch (op) {
        case 1: *r = Fr_mul(a, b); return 0;
        case 2: *r = Fr_div(a, b); return 0;
        case 3: *r = Fr_add(a, b); return 0;
        case 4: *r = Fr_sub(a, b); return 0;
        case 5: *r = Fr_pow(a, b); return 0;
        case 6: if (b == 0) return 1; *r = Fr_idiv(a, b); return 0;   // the reference divides by zero (SIGFPE)
        case 7: if (b == 0) return 1; *r = Fr_mod(a, b); return 0;
        case 8: *r = Fr_shl(a, b); return 0;
        case 9: *r = Fr_shr(a, b); return 0;
        case 10: *r = Fr_leq(a, b); return 0;
        case 11: *r = Fr_geq(a, b); return 0;
        case 12: *r = Fr_lt(a, b); return 0;
        case 13: *r = Fr_gt(a, b); return 0;
        case 14: *r = Fr_eq(a, b); return 0;
        case 15: *r = Fr_neq(a, b); return 0;
        case 16: *r = Fr_lor(a, b); return 0;
        case 17: *r = Fr_land(a, b); return 0;
        case 18: *r = Fr_lnot(a); return 0;
        case 19: *r = Fr_bor(a, b); return 0;
        case 20: *r = Fr_band(a, b); return 0;
        case 21: *r = Fr_bxor(a, b); return 0;
        case 22: *r = Fr_bnot(a); return 0;
        case 23: *r = Fr_neg(a); return 0;
        case 28: *r = Fr_inv(a); return 0;
    }
    return -1;
}
extern "C" int gl_to_int(uint64_t a) { return Fr_toInt(a); }
extern "C" int gl_is_true(uint64_t a) { return Fr_isTrue(a); }
