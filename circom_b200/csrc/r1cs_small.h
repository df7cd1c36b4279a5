// Integer rows of the R1CS check.
//
// The constraints of boolean logic, adders and range checks (the whole of a SHA-256 circuit) are rows whose three linear
// combinations are short sums of +-2^k times wires that hold bits or other small numbers.  For such a row nothing has to
// happen in the field: with |A.w|, |B.w|, |C.w| < 2^61 the product fits 122 bits and  A.w * B.w = C.w  holds modulo a
// 254-bit prime exactly when it holds over the integers (|A.w * B.w - C.w| < 2^123 < q).  The host lists the rows that
// are small BY SHAPE (r1cs_compile.cpp: every coefficient +-2^k, at most R1CS_SMALL_MAX_TERMS terms per linear
// combination, and the sum of the terms' bounds - 2^(16 + k) per term, 2^(k + count) per run of plane bits - below 2^61); whether the VALUES
// are small (below 2^16: bits, bytes, carries) only the run shows - hash circuits do not constrain their inputs to be bits,
// so no range analysis proves it.  r1cs_small_kernel therefore tests every value it reads and, when one
// is wider, marks the row in a bitmap; the general kernel then decides the marked rows (for all instances) in the field.
//
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

constexpr uint32_t R1CS_SMALL_MAX_SHIFT = 44, R1CS_SMALL_MAX_BITS = 60, R1CS_SMALL_MAX_TERMS = 255;
constexpr int R1CS_SMALL_SUM_BITS = 61;   // the host admits a row when the bound of each of its three sums stays below 2^61

// The integer rows have a term list of their own (r1cs_compile.cpp: build_small_groups): 8-byte records, the rows in groups
// of 32 consecutive rows of the sorted order, the records of a group interleaved - record t of row r of the group at
// base + t * 32 + r - and every row of a group padded to the group's term counts with records of coefficient 0.  With
// one-instance tiles a warp is such a group: its 32 record loads are one 256-byte line, its loop bounds are uniform; with
// 32-instance tiles a warp is one row and the record a broadcast.
struct R1csSmallRec {
    uint32_t loc;   // SM_RUN: plane word index; SM_BIT (= OPERAND_BIT): plane bit position; else the slot id.  Flags below.
    uint32_t mag;   // |coefficient| = 2^k: k in bits 0-5, SM_PAD: padding (coefficient 0);  runs: first | (count - 1) << 5 | k << 10
};
constexpr uint32_t SM_NEG = 0x80000000u, SM_RUN = 0x40000000u, SM_BIT = 0x20000000u, SM_BROW = 0x10000000u,
                   SM_PAD = 0x100u, SM_LOC = 0x0FFFFFFFu, SM_BITPOS = 0x1FFFFFFFu;   // SM_BROW (slot terms only - a plane bit position owns bit 28):
                                                                    // the wire's boolean row rides on this term (sbrow[])

// one term: x0 = low 32 bits of the value, upper = OR of its other limbs
CW_SM_HD void r1cs_small_acc(unsigned long long &pos, unsigned long long &neg, uint32_t &wide, uint32_t loc, uint32_t mag,
                             uint32_t x0, uint32_t upper) {
    wide |= upper | (x0 >> 16);
    const unsigned long long t = (unsigned long long)(x0 & ((mag & SM_PAD) ? 0u : 0xFFFFu)) << (mag & 63u);
    if (loc & SM_NEG) neg += t;
    else pos += t;
}
// a run of plane bits: `count` bits of one plane word from bit `first`, times +-2^k (k + count <= R1CS_SMALL_MAX_BITS)
CW_SM_HD void r1cs_small_acc_run(unsigned long long &pos, unsigned long long &neg, uint32_t loc, uint32_t mag, uint32_t plane_word) {
    const uint32_t first = mag & 31u, cnt = ((mag >> 5) & 31u) + 1u, sh = mag >> 10;
    const uint32_t word = (plane_word >> first) & (cnt >= 32u ? 0xFFFFFFFFu : ((1u << cnt) - 1u));
    const unsigned long long t = (unsigned long long)word << sh;
    if (loc & SM_NEG) neg += t;
    else pos += t;
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
