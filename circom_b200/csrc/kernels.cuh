// sm_100a kernels: tape execution, input staging, witness gather, R1CS check, field batch ops.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fr_device.cuh"

namespace cw {

// Per-prime parameters live in constant memory so that modulus limbs are read as c[bank][imm]
// instruction operands (no registers, no loads).
__constant__ FrParams c_fr[2];

// ---- value-slot storage ---------------------------------------------------------------------
// One instance tile holds BT = 1 << bt_log2 instances.  A slot (256-bit value) of a tile is two
// 16-byte halves, each stored for the BT instances contiguously:
//     uint4 index = (tile * n_slots + slot) * 2 * BT + half * BT + instance_in_tile
// so a (warp of) thread(s) working on BT instances of one op issues 128-bit loads over
// BT*16 contiguous bytes per half; with BT = 1 this is the plain 32-byte element (one DRAM sector).
// sm_100 moves a whole 32-byte element with one instruction (LDG/STG.E.ENL2.256): half the memory
// instructions and half the L1/L2 requests of a pair of 128-bit accesses.  32-byte alignment required.
__device__ __forceinline__ void ldg256(u32 *v, const void *p) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p)
                 : "memory");
}
__device__ __forceinline__ void ldg256_nc(u32 *v, const void *p) {  // data that no thread of the kernel writes
    asm("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
        : "l"(p));
}
__device__ __forceinline__ void stg256(void *p, const u32 *v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

__device__ __forceinline__ void load_slot(u32 *v, const uint4 *__restrict__ tile_base, u32 slot, u32 bt_log2,
                                          u32 inst) {
    if (bt_log2 == 0) {  // one instance per tile: the element is contiguous
        ldg256(v, tile_base + ((size_t)slot << 1));
        return;
    }
    size_t i = ((size_t)slot << (bt_log2 + 1)) + inst;
    uint4 lo = tile_base[i];
    uint4 hi = tile_base[i + ((size_t)1 << bt_log2)];
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
    v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
}
__device__ __forceinline__ void store_slot(const u32 *v, uint4 *__restrict__ tile_base, u32 slot, u32 bt_log2,
                                           u32 inst) {
    if (bt_log2 == 0) {
        stg256(tile_base + ((size_t)slot << 1), v);
        return;
    }
    size_t i = ((size_t)slot << (bt_log2 + 1)) + inst;
    tile_base[i] = make_uint4(v[0], v[1], v[2], v[3]);
    tile_base[i + ((size_t)1 << bt_log2)] = make_uint4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void load_const(u32 *v, const uint4 *__restrict__ consts, u32 idx) {
    uint4 lo = __ldg(&consts[2 * (size_t)idx]);
    uint4 hi = __ldg(&consts[2 * (size_t)idx + 1]);
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
    v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
}
__device__ __forceinline__ void load_operand(u32 *v, u32 operand, const uint4 *__restrict__ tile_base,
                                             const uint4 *__restrict__ consts, u32 bt_log2, u32 inst) {
    if (operand & 0x80000000u) load_const(v, consts, operand & 0x7FFFFFFFu);
    else load_slot(v, tile_base, operand, bt_log2, inst);
}

__device__ __forceinline__ u32 u256_bitlen_dev(const u32 *a) {
    u32 n = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (a[i]) n = 32u * i + (32u - __clz(a[i]));
    return n;
}

struct TapeDev {
    const uint4 *ops;          // {opcode, a, b, c}
    const u32 *level_start;    // n_levels + 1
    const uint4 *consts;       // 2 per constant
    const u32 *input_slot;     // slot of main input k
    const u32 *fn_code;        // register-machine code of the circuit's functions (5 words per instruction)
    const u32 *fn_info;        // per function {code offset, n_instr, n_regs, n_params}
    const u32 *call_tab;       // per call {function, n_args, arg operands...}
    u32 n_levels;
    u32 n_slots;
    u32 n_inputs;
};

// ---- inputs: inputs[batch][n_inputs][8 u32] canonical -> slots 1..n_inputs, slot 0 = 1 ----------
__global__ void stage_inputs_kernel(TapeDev tp, const uint4 *__restrict__ inputs, uint4 *__restrict__ slots,
                                    u32 batch, u32 batch_padded, u32 bt_log2) {
    size_t total = (size_t)batch_padded * (tp.n_inputs + 1);
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        u32 inst = (u32)(w % batch_padded);
        u32 k = (u32)(w / batch_padded);  // 0 = constant one, 1.. = input k-1
        u32 tile = inst >> bt_log2, li = inst & ((1u << bt_log2) - 1);
        uint4 *base = slots + (((size_t)tile * tp.n_slots) << (bt_log2 + 1));
        u32 v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (k == 0) v[0] = 1;
        else if (inst < batch) {
            const uint4 *src = inputs + ((size_t)inst * tp.n_inputs + (k - 1)) * 2;
            uint4 lo = src[0], hi = src[1];
            v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
            v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
        }
        store_slot(v, base, k == 0 ? 0u : __ldg(&tp.input_slot[k - 1]), bt_log2, li);
    }
}

// ---- the tape interpreter ---------------------------------------------------------------------
// One CTA owns one tile of BT instances and walks the levels of the tape; within a level the work
// items (op, instance) are spread over the CTA's threads, instance fastest.  Values produced in
// level l are consumed in later levels by other threads of the same CTA only, so a CTA barrier
// per level is the only synchronisation (no grid-wide sync, tiles are independent).
// A function call (circom `function` with run-time loops / branches): the thread copies the arguments into
// the callee's registers (local memory: they are indexed dynamically) and interprets the body.
template <int PRIME>
__device__ __noinline__ void exec_call(const TapeDev &tp, u32 call_off, const uint4 *base, u32 bt_log2, u32 li, u32 *r,
                                       int *err) {
    const FrParams &P = c_fr[PRIME];
    const u32 *ct = tp.call_tab + call_off;
    const u32 f = __ldg(&ct[0]), n_args = __ldg(&ct[1]);
    FnInfo fi;
    fi.code_off = __ldg(&tp.fn_info[4 * f]);
    fi.n_instr = __ldg(&tp.fn_info[4 * f + 1]);
    fi.n_regs = __ldg(&tp.fn_info[4 * f + 2]);
    fi.n_params = __ldg(&tp.fn_info[4 * f + 3]);
    u32 regs[VM_MAX_REGS * 8];
    for (u32 k = 0; k < fi.n_regs * 8; ++k) regs[k] = 0;
    for (u32 k = 0; k < n_args; ++k) {
        u32 v[8];
        load_operand(v, __ldg(&ct[2 + k]), base, tp.consts, bt_log2, li);  // call-table operands carry no ring flag
        for (int j = 0; j < 8; ++j) regs[8 * k + j] = v[j];
    }
    int e = 0;
    vm_run(tp.fn_code, fi, regs, reinterpret_cast<const u32 *>(tp.consts), r, P, e);
    *err = e;
}

// Shared-memory forwarding ring (BT = 1 layouts): every single-value result is also deposited at
// ring[dst % RING_N] (two 16-byte halves in separate arrays: consecutive entries are conflict-free), and an
// operand the lowering flagged with bit 30 is read from there instead of from L2 - most operands of a level
// were produced a few levels earlier by the same CTA.  RING_N must equal CW_RING_SIZE of tape.h: the flags
// are computed for exactly this size.
constexpr u32 RING_N = 512;
constexpr u32 OPD_CONST = 0x80000000u, OPD_RING = 0x40000000u, OPD_SLOT = 0x00FFFFFFu;

template <bool RING>
__device__ __forceinline__ void load_operand_t(u32 *v, u32 operand, const uint4 *__restrict__ tile_base,
                                               const uint4 *__restrict__ consts, u32 bt_log2, u32 li,
                                               const uint4 *ring) {
    if (operand & OPD_CONST) {
        load_const(v, consts, operand & 0x7FFFFFFFu);
    } else if (RING && (operand & OPD_RING)) {
        const u32 i = operand & (RING_N - 1u);
        const uint4 lo = ring[i], hi = ring[RING_N + i];
        v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
        v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    } else {
        load_slot(v, tile_base, operand & OPD_SLOT, bt_log2, li);
    }
}

// HAS_CALLS selects the build that contains the function interpreter (more registers, a local-memory
// frame); tapes without calls - all circuits whose hints are straight-line - use the lean build.
#ifndef CW_TAPE_LB
#define CW_TAPE_LB 1024
#endif
#ifndef CW_TAPE_MINB
#define CW_TAPE_MINB 1
#endif
// BT >= 0 fixes the tile size at compile time (BT = 0, one instance per CTA, is the common layout: the slot
// address arithmetic then folds to `base + slot * 32`); BT < 0 takes it from the launch argument.
template <int PRIME, bool HAS_CALLS, bool RING, int BT>
__global__ void __launch_bounds__(CW_TAPE_LB, CW_TAPE_MINB)
    tape_exec_kernel(TapeDev tp, uint4 *__restrict__ slots, u32 bt_log2_arg, u32 *__restrict__ first_assert,
                     int *__restrict__ err, u32 batch) {
    extern __shared__ uint4 ring[];  // RING: 2 * RING_N entries (16 KB)
    const FrParams &P = c_fr[PRIME];
    const u32 bt_log2 = BT >= 0 ? (u32)BT : bt_log2_arg;
    constexpr bool COOP = BT == 0 && !HAS_CALLS;  // warp-cooperative bit-run stores (needs blockDim % 32 == 0)
    const u32 tile = blockIdx.x;
    const u32 bt_mask = (1u << bt_log2) - 1;
    uint4 *base = slots + (((size_t)tile * tp.n_slots) << (bt_log2 + 1));
    u32 lb = tp.level_start[0];
    u32 le = tp.n_levels ? tp.level_start[1] : lb;
    // the tape word of a thread's first work item of the next level is fetched before the barrier of the
    // current one, taking one memory round trip off the per-level critical path
    uint4 pre = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < ((le - lb) << bt_log2)) pre = __ldg(&tp.ops[lb + (threadIdx.x >> bt_log2)]);
    for (u32 l = 0; l < tp.n_levels; ++l) {
        const u32 n = (le - lb) << bt_log2;
        const u32 le_next = (l + 1 < tp.n_levels) ? tp.level_start[l + 2] : le;
        // COOP (one instance per CTA): the warp walks the level together - lanes beyond the level's end idle in
        // the body - so that the bit runs of its lanes can be stored cooperatively afterwards
        for (u32 w0 = COOP ? (threadIdx.x & ~31u) : threadIdx.x; w0 < n; w0 += blockDim.x) {
            const u32 w = COOP ? w0 + (threadIdx.x & 31u) : w0;
            u32 run_dst = 0, run_n = 0, run_bits = 0;
            if (!COOP || w < n) do {
            const u32 oi = lb + (w >> bt_log2);
            const u32 li = w & bt_mask;
            const uint4 opw = (w == threadIdx.x) ? pre : __ldg(&tp.ops[oi]);
            const u32 opcode = opw.x & 0xFFu, dst = opw.x >> 8;
            const u32 inst = (tile << bt_log2) + li;
            u32 r[8];
            if (HAS_CALLS && opcode == OP_CALL) {
                int e = 0;
                exec_call<PRIME>(tp, opw.y, base, bt_log2, li, r, &e);
                if (e && inst < batch) err[inst] = 1;
            } else if (opcode == OP_BITS && ((opw.w >> 16) & 0xFFu) <= 32u && !(opw.y & OPD_CONST)) {
                // narrow bit-field of a slot value: fetch only the one or two 32-bit words that hold it
                const u32 k = opw.w & 0xFFFFu, m = (opw.w >> 16) & 0xFFu, run = (opw.w >> 24) + 1u;
                const u32 wd = k >> 5, sh = k & 31u;
                const bool two = sh + m + run - 1u > 32u && wd < 7u;
                u32 lo, hi = 0;
                if (RING && (opw.y & OPD_RING)) {
                    const u32 *rw = reinterpret_cast<const u32 *>(ring);
                    const u32 i = opw.y & (RING_N - 1u), w1 = wd + 1u;
                    lo = rw[(((wd >> 2) * RING_N + i) << 2) + (wd & 3u)];
                    if (two) hi = rw[(((w1 >> 2) * RING_N + i) << 2) + (w1 & 3u)];
                } else {
                    const u32 *words = reinterpret_cast<const u32 *>(base);
                    const size_t src = (size_t)(opw.y & OPD_SLOT) << (bt_log2 + 1);
                    const u32 w1 = wd + 1u;
                    lo = words[((src + ((size_t)(wd >> 2) << bt_log2) + li) << 2) + (wd & 3u)];
                    if (two) hi = words[((src + ((size_t)(w1 >> 2) << bt_log2) + li) << 2) + (w1 & 3u)];
                }
                const unsigned long long window = (((unsigned long long)hi << 32) | lo) >> sh;
#pragma unroll
                for (int i = 1; i < 8; ++i) r[i] = 0;
                if (run > 1u) {
                    // a run writes `run` (<= 32) consecutive slots, one bit each; runs bypass the ring
                    if (COOP) {  // stored by the whole warp after the body
                        run_dst = dst;
                        run_n = run;
                        run_bits = (u32)window;
                    } else {
                        for (u32 j = 0; j < run; ++j) {
                            r[0] = (u32)(window >> j) & 1u;
                            store_slot(r, base, dst + j, bt_log2, li);
                        }
                    }
                    continue;  // (leaves the do { } while (0) body)
                }
                r[0] = (u32)window & (m >= 32u ? 0xFFFFFFFFu : ((1u << m) - 1u));
            } else {
                u32 a[8], b[8];
                load_operand_t<RING>(a, opw.y, base, tp.consts, bt_log2, li, ring);
                load_operand_t<RING>(b, opw.z, base, tp.consts, bt_log2, li, ring);
                if (opcode == OP_SELECT) {
                    u32 c[8];
                    load_operand_t<RING>(c, opw.w, base, tp.consts, bt_log2, li, ring);
                    bool t = !u256_is_zero(c);
#pragma unroll
                    for (int i = 0; i < 8; ++i) r[i] = t ? a[i] : b[i];
                } else if (opcode == OP_ASSERT_EQ || opcode == OP_ASSERT || opcode == OP_ASSERT_BOOL ||
                           opcode == OP_ASSERT_FITS) {
                    bool ok = opcode == OP_ASSERT_EQ     ? u256_eq(a, b)
                              : opcode == OP_ASSERT      ? !u256_is_zero(a)
                              : opcode == OP_ASSERT_BOOL ? (u256_is_zero(a) || u256_eq(a, b))
                                                         : (u256_bitlen_dev(a) <= b[0]);
                    if (!ok && inst < batch) atomicMin(&first_assert[inst], opw.w);
                    continue;  // asserts have no destination value
                } else {
                    int e = 0;
                    fr_exec(opcode, r, a, b, opw.w, P, e);
                    if (e && inst < batch) err[inst] = 1;
                }
            }
            store_slot(r, base, dst, bt_log2, li);
            if (RING) {
                ring[dst & (RING_N - 1u)] = make_uint4(r[0], r[1], r[2], r[3]);
                ring[RING_N + (dst & (RING_N - 1u))] = make_uint4(r[4], r[5], r[6], r[7]);
            }
            } while (0);
            if (COOP) {
                // Bit runs, warp-cooperatively: the slots of a run are consecutive, so lane j stores bit j and one
                // store instruction covers run * 32 contiguous bytes (whole 128-byte lines) - a lane streaming its
                // own run would touch one line per instruction and lane, 32 different lines per instruction.
                unsigned pending = __ballot_sync(0xFFFFFFFFu, run_n != 0u);
                const u32 lane = threadIdx.x & 31u;
                while (pending) {
                    const int src = __ffs(pending) - 1;
                    pending &= pending - 1u;
                    const u32 d = __shfl_sync(0xFFFFFFFFu, run_dst, src);
                    const u32 cnt = __shfl_sync(0xFFFFFFFFu, run_n, src);
                    const u32 bits = __shfl_sync(0xFFFFFFFFu, run_bits, src);
                    if (lane < cnt) {
                        u32 r[8] = {(bits >> lane) & 1u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                        stg256(base + ((size_t)(d + lane) << 1), r);
                    }
                }
            }
        }
        if (threadIdx.x < ((le_next - le) << bt_log2)) pre = __ldg(&tp.ops[le + (threadIdx.x >> bt_log2)]);
        lb = le;
        le = le_next;
        __syncthreads();
    }
}

// ---- witness compaction for tile layouts with BT > 1: out[inst][w] = slot w (witness entries are the
// first n_witness slots, already canonical).  With BT = 1 the witness rows are contiguous inside the slot
// store and are read in place (strided) or copied with cudaMemcpy2D.
__global__ void witness_compact_kernel(const uint4 *__restrict__ slots, uint4 *__restrict__ out, u32 n_slots,
                                       u32 n_witness, u32 batch, u32 bt_log2) {
    const u32 bt = 1u << bt_log2;
    const u32 tiles = (batch + bt - 1) >> bt_log2;
    size_t total = (size_t)tiles * n_witness * bt;
    for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
        u32 li = (u32)(w & (bt - 1));
        size_t rest = w >> bt_log2;
        u32 wi = (u32)(rest % n_witness);
        u32 tile = (u32)(rest / n_witness);
        u32 inst = (tile << bt_log2) + li;
        if (inst >= batch) continue;
        const uint4 *base = slots + (((size_t)tile * n_slots) << (bt_log2 + 1));
        u32 v[8];
        load_slot(v, base, wi, bt_log2, li);
        uint4 *dst = out + ((size_t)inst * n_witness + wi) * 2;
        dst[0] = make_uint4(v[0], v[1], v[2], v[3]);
        dst[1] = make_uint4(v[4], v[5], v[6], v[7]);
    }
}

// ---- packed witness for the device->host transfer -------------------------------------------------------
// Most witness entries of real circuits are bits or 64-bit limbs stored as 32-byte field elements.  The
// lowering knows an upper bound of every entry's bit length (range analysis); entries proven to be one bit
// travel as one bit, entries proven <= 64 bits as 8 bytes, the rest as 32 bytes, and the host expands them
// back to the canonical 32-byte rows.  Each value is re-checked here against its class: a violation raises
// `flag` and the caller falls back to the plain copy.  Per-instance packed layout (32-bit words):
// [bit words][u64 entries][full entries].  BT = 1 layout only.
__global__ void witness_pack_kernel(const uint4 *__restrict__ slots, u32 n_slots, const u32 *__restrict__ bit_wire,
                                    u32 n_bits, const u32 *__restrict__ u64_wire, u32 n_u64,
                                    const u32 *__restrict__ full_wire, u32 n_full, u32 *__restrict__ packed,
                                    size_t words_per_inst, u32 batch, int *__restrict__ flag) {
    const u32 n_bit_words = (n_bits + 31u) >> 5;
    const size_t items = (size_t)n_bit_words + n_u64 + n_full;
    for (u32 inst = blockIdx.y; inst < batch; inst += gridDim.y) {
        const uint4 *base = slots + (size_t)inst * n_slots * 2;
        u32 *out = packed + (size_t)inst * words_per_inst;
        for (size_t it = blockIdx.x * (size_t)blockDim.x + threadIdx.x; it < items; it += (size_t)gridDim.x * blockDim.x) {
            if (it < n_bit_words) {
                u32 word = 0, bad = 0;
                const u32 j0 = (u32)it << 5;
#pragma unroll 4
                for (u32 j = 0; j < 32u; ++j) {
                    if (j0 + j < n_bits) {
                        const u32 w = __ldg(&bit_wire[j0 + j]);
                        u32 x[8];
                        ldg256_nc(x, base + 2 * (size_t)w);
                        bad |= x[1] | x[2] | x[3] | x[4] | x[5] | x[6] | x[7] | (x[0] & ~1u);
                        word |= (x[0] & 1u) << j;
                    }
                }
                out[it] = word;
                if (bad) *flag = 1;
            } else if (it < (size_t)n_bit_words + n_u64) {
                const u32 k = (u32)(it - n_bit_words);
                const u32 w = __ldg(&u64_wire[k]);
                u32 x[8];
                ldg256_nc(x, base + 2 * (size_t)w);
                if (x[2] | x[3] | x[4] | x[5] | x[6] | x[7]) *flag = 1;
                out[n_bit_words + 2 * (size_t)k] = x[0];
                out[n_bit_words + 2 * (size_t)k + 1] = x[1];
            } else {
                const u32 k = (u32)(it - n_bit_words - n_u64);
                const u32 w = __ldg(&full_wire[k]);
                u32 x[8];
                ldg256_nc(x, base + 2 * (size_t)w);
                u32 *o = out + n_bit_words + 2 * (size_t)n_u64 + 8 * (size_t)k;
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = x[i];
            }
        }
    }
}

// ---- R1CS check: A.w * B.w == C.w for every row and instance ------------------------------------
// CSR: row_ptr[3m+1] (A, B, C blocks per row), col[nnz] (wire), coef[nnz] (dictionary index).
// Per dictionary entry: dictM = coefficient * R mod q and a kind word
//     bits 0-7  : 0 general, 1 = +1, 2 = -1, 3 = +2^k, 4 = -2^k        bits 8-15 : k
// witness[inst][n_wires][8 u32] canonical.
//
// Work decomposition: thread = (row, instance).  Rows are visited through `perm`, a host-side stable
// sort of the rows by structure (term counts and coefficient kinds), so that the 32 rows of a warp have
// the same length and take the same branches: circuits mix 3-term boolean rows with 65-term
// bit-sum rows, and with the natural order every warp would run at the speed of its longest row.
// blockIdx.y walks groups of instances; a thread re-walks its row for each instance of the group,
// so the row's col/coef words come from L1 after the first instance.
//
// Arithmetic: +-1 coefficients are modular add/sub; +-2^k coefficients shift the witness value when
// the shifted value provably stays below q (runtime check of the value's bit length; a Montgomery
// product otherwise); the row product a*b is skipped when a or b is 0 or 1 (boolean-constraint rows).
struct R1csDev {
    const unsigned long long *row_ptr;
    // per term {wire, coefficient dictionary index, kind word of the coefficient, row id of the boolean constraint
    // of the wire that is checked alongside or ~0}: one 128-bit load per term
    const uint4 *terms;
    const uint4 *dictM;
    const u32 *perm;
    u32 n_constraints;
    u32 n_wires;
    u32 inst_per_block;
    unsigned long long w_stride;  // distance between two instances' witness rows, in 32-byte elements
};

template <int PRIME>
__device__ __forceinline__ void r1cs_lc(u32 *acc, const R1csDev &R, unsigned long long b, unsigned long long e,
                                        const uint4 *__restrict__ w, const FrParams &P,
                                        unsigned long long *__restrict__ first_bad_inst, u32 step = 1) {
    u256_set_u32(acc, 0);
    for (unsigned long long k = b; k < e; k += step) {
        // one 16-byte record per term: {wire, dictionary index, kind word, absorbed boolean row}
        const uint4 term = __ldg(&R.terms[k]);
        const u32 c = term.x, ci = term.y, kw = term.z, brow = term.w;
        u32 x[8];
        ldg256_nc(x, w + 2 * (size_t)c);
        u32 t[8];
        u32 kd = kw & 0xFF, sh = kw >> 8;
        bool neg = (kd == 2) || (kd == 4);
        const u32 upper = x[1] | x[2] | x[3] | x[4] | x[5] | x[6] | x[7];
        // the boolean constraint x*(x-1) = 0 of this wire is checked here, while its value is in registers
        if (brow != 0xFFFFFFFFu && (upper || x[0] > 1u)) atomicMin(first_bad_inst, (unsigned long long)brow);
        if (kd >= 3) {
            if (!upper && sh + 32u < P.qbits) {  // x < 2^32 and x * 2^sh < 2^(qbits-1) < q
                // one-limb value times 2^sh: place the (at most 64-bit) shifted value, no reduction needed
                const u32 wd = sh >> 5, s = sh & 31u;
                const u32 l = x[0] << s, h = s ? (x[0] >> (32u - s)) : 0u;
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = ((u32)i == wd) ? l : (((u32)i == wd + 1u) ? h : 0u);
            } else if (u256_bitlen_dev(x) + sh < P.qbits) {   // x * 2^sh < 2^(qbits-1) < q : plain shift
                u32 y[8];
                u256_shl(y, x, sh);
                u256_set(x, y);
            } else {
                u32 cm[8], p[8];
                load_const(cm, R.dictM, ci);
                fr_mont_mul(p, cm, x, P);           // (cR) * x / R = c*x, sign included
                u256_set(x, p);
                neg = false;
            }
        } else if (kd == 0) {
            u32 cm[8], p[8];
            load_const(cm, R.dictM, ci);
            fr_mont_mul(p, cm, x, P);
            u256_set(x, p);
        }
        if (neg) fr_sub(t, acc, x, P);
        else fr_add(t, acc, x, P);
        u256_set(acc, t);
    }
}

// a * b == c for canonical a, b, c; the product is skipped for a or b in {0, 1, -1}
__device__ __forceinline__ bool r1cs_row_holds(const u32 *a, const u32 *b, const u32 *c, const FrParams &P) {
    bool ok;
    u32 ha = a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7];
    u32 hb = b[1] | b[2] | b[3] | b[4] | b[5] | b[6] | b[7];
    if ((!ha && a[0] == 0) || (!hb && b[0] == 0)) ok = u256_is_zero(c);
    else if (!ha && a[0] == 1) ok = u256_eq(b, c);
    else if (!hb && b[0] == 1) ok = u256_eq(a, c);
    else if (a[0] + 1u == P.q[0] && a[1] == P.q[1] && a[2] == P.q[2] && a[3] == P.q[3] && a[4] == P.q[4] &&
             a[5] == P.q[5] && a[6] == P.q[6] && a[7] == P.q[7]) {
        // a = -1: rows `out <== x*y` are stored as (-x) * y = -out (the reference's normal form), so a
        // bit-valued x = 1 lands here: -b == c
        u32 s[8];
        fr_add(s, b, c, P);
        ok = u256_is_zero(s);
    } else {
        u32 ab[8], c1[8];
        fr_mont_mul(ab, a, b, P);  // a*b/R
        fr_from_mont(c1, c, P);    // c/R
        ok = u256_eq(ab, c1);
    }
    return ok;
}

// MINB = CTAs per SM the register budget is cut for.  Measured on the bench circuit (long rows, bound by memory
// latency), batch 1024: 3 (78 registers, no spills) 16.4 ms, 4 (64) 13.7 ms, 5 (48, 240 B spilled) 13.0 ms, 6 (40)
// 17.0 ms - resident warps buy more than the spills cost.  Circuits of short rows (SHA-256: ~5 terms and a
// Montgomery product per row) are arithmetic-bound and prefer the unspilled build; the host picks by the mean
// row length.
template <int PRIME, int MINB>
__global__ void __launch_bounds__(256, MINB) r1cs_check_kernel(R1csDev R, const uint4 *__restrict__ witness, u32 batch,
                                                         unsigned long long *__restrict__ first_bad) {
    const FrParams &P = c_fr[PRIME];
    const u32 i0 = blockIdx.y * R.inst_per_block;
    const u32 i1 = min(batch, i0 + R.inst_per_block);
    for (u32 rs = blockIdx.x * blockDim.x + threadIdx.x; rs < R.n_constraints; rs += gridDim.x * blockDim.x) {
        const u32 row = __ldg(&R.perm[rs]);
        const unsigned long long p0 = R.row_ptr[3 * (size_t)row], p1 = R.row_ptr[3 * (size_t)row + 1],
                                 p2 = R.row_ptr[3 * (size_t)row + 2], p3 = R.row_ptr[3 * (size_t)row + 3];
        for (u32 inst = i0; inst < i1; ++inst) {
            const uint4 *w = witness + (size_t)inst * R.w_stride * 2;
            u32 a[8], b[8], c[8];
            r1cs_lc<PRIME>(a, R, p0, p1, w, P, &first_bad[inst]);
            r1cs_lc<PRIME>(b, R, p1, p2, w, P, &first_bad[inst]);
            r1cs_lc<PRIME>(c, R, p2, p3, w, P, &first_bad[inst]);
            const bool ok = r1cs_row_holds(a, b, c, P);
            if (!ok) atomicMin(&first_bad[inst], (unsigned long long)row);
        }
    }
}

// Long rows (the 65-term recomposition sums of range checks, polynomial identities): G lanes share one
// (row, instance); lane g takes terms g, g + G, ... of each linear combination and the partial sums are
// combined with a butterfly of modular additions.  The terms of such a row reference consecutive witness
// entries, so the G lanes read G adjacent 32-byte elements (whole 128-byte lines) where one thread walking
// the row alone touches one sector per load; and the dependent chain per thread is G times shorter.
template <int PRIME, int G>
__global__ void __launch_bounds__(256) r1cs_check_split_kernel(R1csDev R, const uint4 *__restrict__ witness, u32 batch,
                                                               unsigned long long *__restrict__ first_bad) {
    const FrParams &P = c_fr[PRIME];
    const u32 i0 = blockIdx.y * R.inst_per_block;
    const u32 i1 = min(batch, i0 + R.inst_per_block);
    const u32 g = threadIdx.x & (G - 1);
    const u32 gmask = ((G == 32) ? 0xFFFFFFFFu : ((1u << G) - 1u)) << ((threadIdx.x & 31u) & ~(u32)(G - 1));
    const u32 per_block = blockDim.x / G;
    for (u32 rs = blockIdx.x * per_block + threadIdx.x / G; rs < R.n_constraints; rs += gridDim.x * per_block) {
        const u32 row = __ldg(&R.perm[rs]);
        unsigned long long p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = R.row_ptr[3 * (size_t)row + j];
        for (u32 inst = i0; inst < i1; ++inst) {
            const uint4 *w = witness + (size_t)inst * R.w_stride * 2;
            u32 lc[3][8];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                r1cs_lc<PRIME>(lc[j], R, p[j] + g, p[j + 1], w, P, &first_bad[inst], G);
                if (p[j + 1] == p[j]) continue;  // empty combination: every lane holds 0
#pragma unroll
                for (int off = G / 2; off > 0; off >>= 1) {
                    u32 o[8], t[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) o[i] = __shfl_xor_sync(gmask, lc[j][i], off);
                    fr_add(t, lc[j], o, P);
                    u256_set(lc[j], t);
                }
            }
            if (g == 0 && !r1cs_row_holds(lc[0], lc[1], lc[2], P)) atomicMin(&first_bad[inst], (unsigned long long)row);
        }
    }
}

// boolean rows x*(x-1) = 0: the witness value must be 0 or 1.  Thread = (boolean row, instance); the
// wires of consecutive boolean rows are consecutive witness entries (the bits of one decomposition), so
// a warp reads a contiguous 1 KB run of the instance's witness row.  Purely memory-bound (32 B per row).
__global__ void __launch_bounds__(256) r1cs_bool_kernel(const u32 *__restrict__ wire, const u32 *__restrict__ rows,
                                                        u32 n_bool, const uint4 *__restrict__ witness,
                                                        unsigned long long w_stride, u32 batch,
                                                        unsigned long long *__restrict__ first_bad) {
    for (u32 inst = blockIdx.y; inst < batch; inst += gridDim.y) {
        const uint4 *w = witness + (size_t)inst * w_stride * 2;
        for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < n_bool; k += gridDim.x * blockDim.x) {
            const u32 c = __ldg(&wire[k]);
            u32 x[8];
            ldg256_nc(x, w + 2 * (size_t)c);
            const u32 rest = x[1] | x[2] | x[3] | x[4] | x[5] | x[6] | x[7];
            if (rest || x[0] > 1u) atomicMin(&first_bad[inst], (unsigned long long)__ldg(&rows[k]));
        }
    }
}

// ---- batched single field op (parity tests of the device Fr_* equivalents) ---------------------
// canonical in / canonical out; the kernel applies the same representation rules as the lowering
template <int PRIME>
__global__ void fr_batch_op_kernel(int op, const uint4 *__restrict__ A, const uint4 *__restrict__ B,
                                   const uint4 *__restrict__ C, uint4 *__restrict__ Rr, size_t n,
                                   int *__restrict__ err) {
    const FrParams &P = c_fr[PRIME];
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        u32 a[8], b[8] = {0, 0, 0, 0, 0, 0, 0, 0}, c[8] = {0, 0, 0, 0, 0, 0, 0, 0}, r[8];
        load_const(a, A, (u32)i);
        if (B) load_const(b, B, (u32)i);
        if (C) load_const(c, C, (u32)i);
        int e = 0;
        if (op == OP_MUL) {  // canonical x canonical: convert one side
            u32 am[8];
            fr_to_mont(am, a, P);
            fr_mont_mul(r, am, b, P);
        } else if (op == 2 /* DIV */) {
            u32 bm[8], im[8];
            fr_to_mont(bm, b, P);
            fr_inv_mont(im, bm, P);
            fr_mont_mul(r, im, a, P);
        } else if (op == OP_POW) {
            u32 am[8], rm[8];
            fr_to_mont(am, a, P);
            fr_pow_mont(rm, am, b, P);
            fr_from_mont(r, rm, P);
        } else if (op == OP_INV) {
            u32 am[8], rm[8];
            fr_to_mont(am, a, P);
            fr_inv_mont(rm, am, P);
            fr_from_mont(r, rm, P);
        } else if (op == OP_SELECT) {
            bool t = !u256_is_zero(c);
            for (int k = 0; k < 8; ++k) r[k] = t ? a[k] : b[k];
        } else {
            fr_exec((u32)op, r, a, b, 0, P, e);
        }
        if (e) err[0] = 1;
        Rr[2 * i] = make_uint4(r[0], r[1], r[2], r[3]);
        Rr[2 * i + 1] = make_uint4(r[4], r[5], r[6], r[7]);
    }
}

// ---- Montgomery-multiplication throughput probe ------------------------------------------------
template <int PRIME>
__global__ void fr_mul_bench_kernel(uint4 *__restrict__ data, size_t n, int iters) {
    const FrParams &P = c_fr[PRIME];
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 x[8], y[8];
    load_const(x, data, (u32)i);
    u256_set(y, x);
    y[0] ^= 0x9E3779B9u & 0x0FFFFFFFu;
#pragma unroll 1
    for (int k = 0; k < iters; ++k) {
        u32 t[8];
        fr_mont_mul(t, x, y, P);
        u256_set(x, t);
    }
    data[2 * i] = make_uint4(x[0], x[1], x[2], x[3]);
    data[2 * i + 1] = make_uint4(x[4], x[5], x[6], x[7]);
}

}  // namespace cw
