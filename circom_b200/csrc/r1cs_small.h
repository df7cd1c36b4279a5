// Integer rows of the R1CS check.
//
// The constraints of boolean logic, adders and range checks (the whole of a SHA-256 circuit) are rows whose three linear
// combinations are short sums of +-2^k times wires that hold bits or other small numbers.  For such a row nothing has to
// happen in the field: with |A.w|, |B.w|, |C.w| < 2^61 the product fits 122 bits and  A.w * B.w = C.w  holds modulo a
// 254-bit prime exactly when it holds over the integers (|A.w * B.w - C.w| < 2^123 < q).  The host lists the rows that
// are small BY SHAPE (r1cs_compile.cpp: every coefficient +-2^k with k <= R1CS_SMALL_MAX_SHIFT, at most
// R1CS_SMALL_MAX_TERMS terms per linear combination, runs of plane bits below 2^R1CS_SMALL_MAX_BITS); whether the VALUES
// are small (below 2^16: bits, bytes, carries) only the run shows - hash circuits do not constrain their inputs to be bits,
// so no range analysis proves it.  r1cs_small_kernel therefore tests every value it reads and, when one
// is wider, marks the row in a bitmap; the general kernel then decides the marked rows (for all instances) in the field.
//
// Bound: a term is below 2^16 * 2^40 = 2^56 (a run: below 2^56), 32 terms per linear combination: |sum| < 2^61.
//
// `__host__ __device__`: tests/hostsim runs the same functions on the CPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CW_SM_HD __host__ __device__ __forceinline__
#else
#define CW_SM_HD inline
#endif

namespace cw {

constexpr uint32_t R1CS_SMALL_MAX_SHIFT = 40, R1CS_SMALL_MAX_BITS = 56, R1CS_SMALL_MAX_TERMS = 32;

// one term on a stored value: x0 = its low 32 bits, upper = the OR of its other seven limbs; kw = the term's kind word
// (kinds 1 / 2: +-1, 3 / 4: +-2^k with k in bits 8-15)
CW_SM_HD void r1cs_small_term(long long &acc, uint32_t &wide, uint32_t kw, uint32_t x0, uint32_t upper) {
    const uint32_t kd = kw & 0xFFu, sh = (kw >> 8) & 0xFFu;
    wide |= upper | (x0 >> 16);
    const long long t = (long long)((unsigned long long)(x0 & 0xFFFFu) << sh);
    acc += (kd == 2u || kd == 4u) ? -t : t;
}
// a run of plane bits (kinds 5 / 6): `count` bits of one plane word from bit `first`, times +-2^k
CW_SM_HD void r1cs_small_run(long long &acc, uint32_t kw, uint32_t plane_word) {
    const uint32_t kd = kw & 0xFFu, sh = (kw >> 8) & 0xFFu, first = (kw >> 16) & 31u, cnt = ((kw >> 21) & 31u) + 1u;
    const uint32_t word = (plane_word >> first) & (cnt >= 32u ? 0xFFFFFFFFu : ((1u << cnt) - 1u));
    const long long t = (long long)((unsigned long long)word << sh);
    acc += kd == 6u ? -t : t;
}
// a * b == c over the integers (|a|, |b|, |c| < 2^61)
CW_SM_HD bool r1cs_small_holds(long long a, long long b, long long c) {
#if defined(__CUDA_ARCH__)
    return a * b == c && __mul64hi(a, b) == (c >> 63);
#else
    return (__int128)a * (__int128)b == (__int128)c;
#endif
}

}  // namespace cw
