// C ABI (include/circom_b200.h) over the lowering (flatten.cpp), the formats (formats.cpp) and the
// sm_100a kernels (kernels.cuh).  There is no CPU execution path: every compute entry point
// returns CW_ENODEV when no CUDA device is present.
#include <cuda_runtime.h>
#include <emmintrin.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/circom_b200.h"
#include "kernels.cuh"
#include "tape.h"

using namespace cw;

static_assert(cw::RING_N == cw::CW_RING_SIZE, "kernel ring size must match the lowering's");
namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(CW_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                 \
    } while (0)

FrParams make_dev_params(const FieldParams &F) {
    FrParams p;
    memset(&p, 0, sizeof(p));
    auto split = [](u32 *dst, const U256 &v) {
        for (int i = 0; i < 4; ++i) {
            dst[2 * i] = (u32)v.v[i];
            dst[2 * i + 1] = (u32)(v.v[i] >> 32);
        }
    };
    split(p.q, F.q);
    split(p.half, F.half);
    split(p.r1, F.r1);
    split(p.r2, F.r2);
    U256 two = u256_from_u64(2), qm2;
    u256_sub(qm2, F.q, two);
    split(p.qm2, qm2);
    p.np32 = F.np32;
    p.qbits = F.qbits;
    p.top_mask = (F.qbits - 224 >= 32) ? 0xFFFFFFFFu : ((1u << (F.qbits - 224)) - 1u);
    return p;
}

std::mutex g_dev_mutex;
std::map<int, bool> g_dev_ready;

int ensure_device(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(CW_ENODEV, "no CUDA device available (circom_b200 has no CPU execution path)");
    }
    if (device < 0 || device >= n) return fail(CW_EINVAL, "bad device index");
    CU(cudaSetDevice(device));
    std::lock_guard<std::mutex> lk(g_dev_mutex);
    if (!g_dev_ready[device]) {
        FrParams h[2] = {make_dev_params(make_field(0)), make_dev_params(make_field(1))};
        CU(cudaMemcpyToSymbol(c_fr, h, sizeof(h)));
        g_dev_ready[device] = true;
    }
    return CW_OK;
}

struct DevTape {
    uint4 *ops = nullptr;
    u32 *level_start = nullptr;
    uint4 *consts = nullptr;
    u32 *input_slot = nullptr, *fn_code = nullptr, *fn_info = nullptr, *call_tab = nullptr;
    u32 *pk_bit = nullptr, *pk_u64 = nullptr, *pk_full = nullptr;
};
struct DevR1cs {
    unsigned long long *row_ptr = nullptr;
    uint4 *terms = nullptr;  // per term {wire, dictionary index, kind word, absorbed boolean row}
    uint4 *dictM = nullptr;
    u32 *perm = nullptr, *bool_wire = nullptr, *bool_row = nullptr;
    u32 n_general = 0, n_bool = 0;
    u32 n_long = 0;  // perm[0, n_long): rows with >= R1CS_SPLIT_MIN terms, checked by lane groups
    u32 mean_row_terms = 0;  // terms per general row
};
constexpr uint64_t R1CS_SPLIT_MIN = 16;
constexpr int R1CS_SPLIT_G = 8;

template <class T>
int upload(T **dst, const void *src, size_t bytes) {
    CU(cudaMalloc((void **)dst, bytes ? bytes : 16));
    if (bytes) CU(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    return CW_OK;
}

int env_int(const char *name, int dflt) {
    const char *s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}

}  // namespace

struct cw_circuit {
    Tape tape;
    mutable std::mutex mu;
    mutable std::map<int, DevTape> dev;
};

struct cw_r1cs {
    R1csData data;
    FieldParams F;
    std::mutex mu;
    std::map<int, DevR1cs> dev;
};

struct cw_batch {
    const cw_circuit *c = nullptr;
    int device = 0;
    u32 batch = 0, batch_padded = 0, bt_log2 = 0, threads = 256;
    cudaStream_t stream = nullptr;
    uint4 *slots = nullptr, *inputs_d = nullptr, *witness_d = nullptr;
    u32 *first_assert_d = nullptr;
    int *err_d = nullptr;
    DevTape dt;
    std::vector<uint64_t> host_inputs;  // [batch][n_inputs][4]
    std::vector<uint8_t> assigned;      // [batch][n_inputs]
    std::vector<u32> remaining;         // [batch]
    bool host_inputs_dirty = false;
    bool inputs_on_device = false;
    bool ran = false;
    bool compact_valid = false;  // witness_d holds the contiguous copy of the current run
    u32 *packed_d = nullptr, *packed_h = nullptr;  // packed witness staging (device / pinned host)
    int *pack_flag_d = nullptr;
    uint64_t last_d2h_bytes = 0;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
};

static int get_dev_tape(const cw_circuit *c, int device, DevTape &out) {
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->dev.find(device);
    if (it != c->dev.end()) {
        out = it->second;
        return CW_OK;
    }
    const Tape &t = c->tape;
    DevTape d;
    int rc;
    if ((rc = upload(&d.ops, t.ops.data(), t.ops.size() * 4))) return rc;
    if ((rc = upload(&d.level_start, t.level_start.data(), t.level_start.size() * 4))) return rc;
    if ((rc = upload(&d.consts, t.consts.data(), t.consts.size() * 32))) return rc;
    if ((rc = upload(&d.input_slot, t.input_slot.data(), t.input_slot.size() * 4))) return rc;
    if ((rc = upload(&d.fn_code, t.fn_code.data(), t.fn_code.size() * 4))) return rc;
    if ((rc = upload(&d.fn_info, t.fn_info.data(), t.fn_info.size() * 4))) return rc;
    if ((rc = upload(&d.call_tab, t.call_tab.data(), t.call_tab.size() * 4))) return rc;
    if ((rc = upload(&d.pk_bit, t.pk_bit_wire.data(), t.pk_bit_wire.size() * 4))) return rc;
    if ((rc = upload(&d.pk_u64, t.pk_u64_wire.data(), t.pk_u64_wire.size() * 4))) return rc;
    if ((rc = upload(&d.pk_full, t.pk_full_wire.data(), t.pk_full_wire.size() * 4))) return rc;
    c->dev[device] = d;
    out = d;
    return CW_OK;
}

extern "C" {

int cw_version(void) { return 100; }
const char *cw_last_error(void) { return g_err.c_str(); }
int cw_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int cw_circuit_load_mem(const void *data, size_t len, uint32_t flags, cw_circuit **out) {
    if (!data || !out) return fail(CW_EINVAL, "null argument");
    cw_circuit *c = new cw_circuit();
    try {
        lower_circuit((const uint8_t *)data, len, flags, c->tape);
    } catch (const std::exception &e) {
        delete c;
        return fail(CW_EFORMAT, e.what());
    }
    *out = c;
    return CW_OK;
}

int cw_circuit_load(const char *path, uint32_t flags, cw_circuit **out) {
    if (!path || !out) return fail(CW_EINVAL, "null argument");
    FILE *f = fopen(path, "rb");
    if (!f) return fail(CW_EIO, std::string("cannot open ") + path);
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> buf(sz);
    size_t rd = sz ? fread(buf.data(), 1, sz, f) : 0;
    fclose(f);
    if ((long)rd != sz) return fail(CW_EIO, "short read");
    return cw_circuit_load_mem(buf.data(), buf.size(), flags, out);
}

void cw_circuit_destroy(cw_circuit *c) {
    if (!c) return;
    for (auto &kv : c->dev) {
        cudaSetDevice(kv.first);
        cudaFree(kv.second.ops);
        cudaFree(kv.second.level_start);
        cudaFree(kv.second.consts);
        cudaFree(kv.second.input_slot);
        cudaFree(kv.second.fn_code);
        cudaFree(kv.second.fn_info);
        cudaFree(kv.second.call_tab);
        cudaFree(kv.second.pk_bit);
        cudaFree(kv.second.pk_u64);
        cudaFree(kv.second.pk_full);
    }
    delete c;
}

int cw_circuit_stats(const cw_circuit *c, cw_stats *o) {
    if (!c || !o) return fail(CW_EINVAL, "null argument");
    const Tape &t = c->tape;
    memset(o, 0, sizeof(*o));
    o->n_signals = t.n_signals;
    o->n_witness = t.n_witness;
    o->n_inputs = t.n_inputs;
    o->n_outputs = t.n_outputs;
    o->n_components = t.n_components;
    o->n_constants = t.consts.size();
    o->n_ir_ops = t.n_ir_ops;
    o->n_tape_ops = t.n_tape_ops();
    o->n_slots = t.n_slots;
    o->n_levels = t.n_levels();
    o->n_constraints = t.r1cs.n_constraints;
    o->n_nnz = t.r1cs.col.size();
    o->n_mul_ops = t.n_mul_ops;
    o->n_conv_ops = t.n_conv_ops;
    o->max_level_width = t.max_level_width;
    o->n_slot_operands = t.n_slot_operands;
    o->n_ring_operands = t.n_ring_operands;
    return CW_OK;
}

int cw_circuit_prime(const cw_circuit *c, int *prime_id, uint64_t q[4]) {
    if (!c) return fail(CW_EINVAL, "null argument");
    if (prime_id) *prime_id = c->tape.F.prime_id;
    if (q) memcpy(q, c->tape.F.q.v, 32);
    return CW_OK;
}

uint32_t cw_get_main_input_signal_start(const cw_circuit *c) { return (uint32_t)c->tape.n_outputs + 1; }
uint32_t cw_get_main_input_signal_no(const cw_circuit *c) { return (uint32_t)c->tape.n_inputs; }
uint32_t cw_get_total_signal_no(const cw_circuit *c) { return (uint32_t)c->tape.n_signals; }
uint32_t cw_get_number_of_components(const cw_circuit *c) { return (uint32_t)c->tape.n_components; }
uint32_t cw_get_size_of_input_hashmap(const cw_circuit *c) { return (uint32_t)c->tape.hashmap.size(); }
uint32_t cw_get_size_of_witness(const cw_circuit *c) { return (uint32_t)c->tape.n_witness; }
uint32_t cw_get_size_of_constants(const cw_circuit *c) { return (uint32_t)c->tape.consts.size(); }

uint64_t cw_fnv1a(const char *name) { return fnv1a(name, strlen(name)); }

// getInputSignalHashPosition (calcwit.cpp:51-69)
static int hash_pos(const Tape &t, uint64_t h, size_t *pos) {
    size_t n = t.hashmap.size();
    size_t p = (size_t)(h % n);
    if (t.hashmap[p].hash != h || t.hashmap[p].signalid == 0) {
        size_t ini = p;
        p = (p + 1) % n;
        while (p != ini) {
            if (t.hashmap[p].hash == h && t.hashmap[p].signalid != 0) {
                *pos = p;
                return CW_OK;
            }
            if (t.hashmap[p].signalid == 0) return fail(CW_ENOTFOUND, "Signal not found");
            p = (p + 1) % n;
        }
        return fail(CW_ENOTFOUND, "Signals not found");
    }
    *pos = p;
    return CW_OK;
}

int cw_get_input_signal_size(const cw_circuit *c, uint64_t h, uint64_t *size) {
    size_t p;
    int rc = hash_pos(c->tape, h, &p);
    if (rc) return rc;
    *size = c->tape.hashmap[p].signalsize;
    return CW_OK;
}
int cw_get_input_signal_id(const cw_circuit *c, uint64_t h, uint64_t *id) {
    size_t p;
    int rc = hash_pos(c->tape, h, &p);
    if (rc) return rc;
    *id = c->tape.hashmap[p].signalid;
    return CW_OK;
}

int cw_circuit_tape(const cw_circuit *c, uint32_t *ops, uint32_t *level_start, uint32_t *witness_slot) {
    const Tape &t = c->tape;
    if (ops) memcpy(ops, t.ops.data(), t.ops.size() * 4);
    if (level_start) memcpy(level_start, t.level_start.data(), t.level_start.size() * 4);
    if (witness_slot) memcpy(witness_slot, t.witness_slot.data(), t.witness_slot.size() * 4);
    return CW_OK;
}

int cw_circuit_slot_census(const cw_circuit *c, uint64_t out[4]) {
    if (!c || !out) return fail(CW_EINVAL, "null argument");
    memcpy(out, c->tape.slot_census, sizeof(c->tape.slot_census));
    return CW_OK;
}

int cw_circuit_witness2signal(const cw_circuit *c, uint64_t *out) {
    if (!c || !out) return fail(CW_EINVAL, "null argument");
    memcpy(out, c->tape.witness2signal.data(), c->tape.witness2signal.size() * 8);
    return CW_OK;
}

int cw_circuit_write_dat(const cw_circuit *c, const char *path) {
    try {
        write_dat(c->tape, path);
    } catch (const std::exception &e) {
        return fail(CW_EIO, e.what());
    }
    return CW_OK;
}

// ---- batch ------------------------------------------------------------------------------------
int cw_batch_create(const cw_circuit *c, uint32_t batch, int device, cw_batch **out) {
    if (!c || !out || batch == 0) return fail(CW_EINVAL, "bad argument");
    if (c->tape.flags & CW_FLAG_HOST_ONLY) return fail(CW_ESTATE, "circuit was loaded with CW_FLAG_HOST_ONLY");
    if (c->tape.n_bitwords)  // the lowering and its CPU verification exist (tests), the kernels do not read the layout yet
        return fail(CW_ESTATE, "CW_FLAG_BITPLANE tapes cannot be executed by this build (lowering-only preview)");
    int rc = ensure_device(device);
    if (rc) return rc;
    const Tape &t = c->tape;
    cw_batch *b = new cw_batch();
    b->c = c;
    b->device = device;
    b->batch = batch;
    // tile size: keep at least ~4 CTAs per SM in flight before widening tiles for coalescing
    int bt = env_int("CW_BT_LOG2", -1);
    if (bt < 0) {
        // BT = 1 keeps witness rows contiguous in the slot store (no compaction) and is as fast as wider tiles
        // whenever a level has enough ops to fill warps; very narrow tapes (Poseidon: ~3 ops per level)
        // need instances side by side in a warp instead
        uint64_t avg_w = t.n_levels() ? t.n_tape_ops() / t.n_levels() + 1 : 1;
        bt = 0;
        while (bt < 5 && (avg_w << bt) < 64 && (batch >> (bt + 1)) >= 296u) ++bt;
    }
    if (bt > 5) bt = 5;
    b->bt_log2 = (u32)bt;
    u32 btn = 1u << bt;
    b->batch_padded = (batch + btn - 1) / btn * btn;
    int th = env_int("CW_THREADS", 0);
    if (th <= 0) {
        // enough threads for a typical level: average width x tile, clamped to [64, 512]
        uint64_t avg = t.n_levels() ? (t.n_tape_ops() / t.n_levels() + 1) * btn : 64;
        th = 64;
        while (th < 512 && (uint64_t)th < avg) th <<= 1;
        // many tiles per SM hide latency better than wide CTAs: keep <= ~1024 resident threads per SM
        // (measured on B200: batch 256 -> 512 threads, 512 -> 256, 1024 -> 128)
        u32 tiles = b->batch_padded >> bt;
        u32 per_sm = (tiles + 147) / 148;
        while (th > 64 && (u32)th * per_sm > 1024) th >>= 1;
    }
    th = (th + 31) / 32 * 32;
    if (th > 1024) th = 1024;
    b->threads = (u32)th;
    size_t slot_bytes = (size_t)b->batch_padded * t.n_slots * 32;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    size_t need = slot_bytes + (size_t)batch * t.n_inputs * 32 + (64u << 20);
    if (need > free_b) {
        delete b;
        return fail(CW_ECUDA, "batch needs " + std::to_string(need >> 20) + " MiB of device memory, " +
                                  std::to_string(free_b >> 20) + " MiB free");
    }
    if ((rc = get_dev_tape(c, device, b->dt))) { delete b; return rc; }
    CU(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    CU(cudaMalloc((void **)&b->slots, slot_bytes));
    CU(cudaMalloc((void **)&b->inputs_d, std::max<size_t>((size_t)batch * t.n_inputs * 32, 32)));
    CU(cudaMalloc((void **)&b->first_assert_d, (size_t)batch * 4));
    CU(cudaMalloc((void **)&b->err_d, (size_t)batch * 4));
    for (auto &e : b->ev) CU(cudaEventCreate(&e));
    b->host_inputs.assign((size_t)batch * t.n_inputs * 4, 0);
    b->assigned.assign((size_t)batch * t.n_inputs, 0);
    b->remaining.assign(batch, (u32)t.n_inputs);
    *out = b;
    return CW_OK;
}

void cw_batch_destroy(cw_batch *b) {
    if (!b) return;
    cudaSetDevice(b->device);
    cudaFree(b->slots);
    cudaFree(b->inputs_d);
    cudaFree(b->witness_d);
    cudaFree(b->packed_d);
    cudaFree(b->pack_flag_d);
    if (b->packed_h) cudaFreeHost(b->packed_h);
    cudaFree(b->first_assert_d);
    cudaFree(b->err_d);
    for (auto &e : b->ev)
        if (e) cudaEventDestroy(e);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

int cw_batch_set_input(cw_batch *b, uint32_t inst, uint64_t h, uint32_t idx, const uint64_t limbs[4]) {
    if (!b || inst >= b->batch) return fail(CW_EINVAL, "bad instance");
    const Tape &t = b->c->tape;
    if (b->remaining[inst] == 0) return fail(CW_ESTATE, "No more signals to be assigned");
    size_t p;
    int rc = hash_pos(t, h, &p);
    if (rc) return rc;
    if (idx >= t.hashmap[p].signalsize) return fail(CW_EINVAL, "Input signal array access exceeds the size");
    uint64_t si = t.hashmap[p].signalid + idx;
    uint64_t k = si - (t.n_outputs + 1);
    if (si < t.n_outputs + 1 || k >= t.n_inputs) return fail(CW_EINVAL, "input signal outside the main inputs");
    if (b->assigned[(size_t)inst * t.n_inputs + k]) return fail(CW_ESTATE, "Signal assigned twice: " + std::to_string(si));
    U256 v;
    memcpy(v.v, limbs, 32);
    if (!(v < t.F.q)) return fail(CW_EINVAL, "input value not reduced modulo the field prime");
    memcpy(&b->host_inputs[((size_t)inst * t.n_inputs + k) * 4], limbs, 32);
    b->assigned[(size_t)inst * t.n_inputs + k] = 1;
    b->remaining[inst]--;
    b->host_inputs_dirty = true;
    return CW_OK;
}

int cw_batch_remaining_inputs(const cw_batch *b, uint32_t inst, uint32_t *rem) {
    if (!b || inst >= b->batch) return fail(CW_EINVAL, "bad instance");
    *rem = b->remaining[inst];
    return CW_OK;
}

int cw_batch_set_inputs(cw_batch *b, const uint64_t *inputs, int is_device_ptr) {
    if (!b || !inputs) return fail(CW_EINVAL, "null argument");
    const Tape &t = b->c->tape;
    CU(cudaSetDevice(b->device));
    size_t bytes = (size_t)b->batch * t.n_inputs * 32;
    CU(cudaMemcpyAsync(b->inputs_d, inputs, bytes, is_device_ptr ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                       b->stream));
    std::fill(b->remaining.begin(), b->remaining.end(), 0u);
    b->host_inputs_dirty = false;
    b->inputs_on_device = true;
    return CW_OK;
}

int cw_batch_run(cw_batch *b) {
    if (!b) return fail(CW_EINVAL, "null argument");
    const Tape &t = b->c->tape;
    CU(cudaSetDevice(b->device));
    if (b->host_inputs_dirty || !b->inputs_on_device) {
        for (u32 i = 0; i < b->batch; ++i)
            if (b->remaining[i])
                return fail(CW_ESTATE, "Not all inputs have been set. Only " +
                                           std::to_string(t.n_inputs - b->remaining[i]) + " out of " +
                                           std::to_string(t.n_inputs) + " (instance " + std::to_string(i) + ")");
        CU(cudaMemcpyAsync(b->inputs_d, b->host_inputs.data(), b->host_inputs.size() * 8, cudaMemcpyHostToDevice,
                           b->stream));
        b->host_inputs_dirty = false;
        b->inputs_on_device = true;
    }
    TapeDev tp;
    tp.ops = b->dt.ops;
    tp.level_start = b->dt.level_start;
    tp.consts = b->dt.consts;
    tp.n_levels = (u32)t.n_levels();
    tp.n_slots = t.n_slots;
    tp.input_slot = b->dt.input_slot;
    tp.fn_code = b->dt.fn_code;
    tp.fn_info = b->dt.fn_info;
    tp.call_tab = b->dt.call_tab;
    tp.n_inputs = (u32)t.n_inputs;
    CU(cudaMemsetAsync(b->first_assert_d, 0xFF, (size_t)b->batch * 4, b->stream));
    CU(cudaMemsetAsync(b->err_d, 0, (size_t)b->batch * 4, b->stream));
    CU(cudaEventRecord(b->ev[0], b->stream));
    {
        size_t total = (size_t)b->batch_padded * (t.n_inputs + 1);
        u32 grid = (u32)std::min<size_t>((total + 255) / 256, 148 * 8);
        stage_inputs_kernel<<<grid, 256, 0, b->stream>>>(tp, b->inputs_d, b->slots, b->batch, b->batch_padded, b->bt_log2);
    }
    u32 tiles = b->batch_padded >> b->bt_log2;
    if (tp.n_levels) {
        const bool calls = !t.call_tab.empty();
        const u32 th = calls ? std::min<u32>(b->threads, 256u) : b->threads;  // the interpreter build has a large frame
        // The shared-memory forwarding ring (kernels.cuh) is opt-in (env CW_RING=1; BT = 1 layouts): measured on
        // B200 it does not shorten the step (21.99 ms with, 21.51 ms without, batch 1024 of the bench circuit -
        // operand latency is not what bounds the interpreter) and its 16 KB per CTA come out of L1.
        const bool ring = !calls && b->bt_log2 == 0 && env_int("CW_RING", 0) != 0;
        const size_t smem = ring ? (size_t)2 * RING_N * sizeof(uint4) : 0;
#define CW_LAUNCH_TAPE(PR, CALLS, RING_, BT_)                                                                  \
    tape_exec_kernel<PR, CALLS, RING_, BT_><<<tiles, th, smem, b->stream>>>(tp, b->slots, b->bt_log2,            \
                                                                              b->first_assert_d, b->err_d, b->batch)
        // builds: calls (runtime tile size), ring (BT = 1 instance), plain with BT = 1 instance fixed at compile
        // time, plain with the tile size as an argument
        const bool bt0 = b->bt_log2 == 0;
        if (t.F.prime_id == 0) {
            if (calls) CW_LAUNCH_TAPE(0, true, false, -1);
            else if (ring) CW_LAUNCH_TAPE(0, false, true, 0);
            else if (bt0) CW_LAUNCH_TAPE(0, false, false, 0);
            else CW_LAUNCH_TAPE(0, false, false, -1);
        } else {
            if (calls) CW_LAUNCH_TAPE(1, true, false, -1);
            else if (ring) CW_LAUNCH_TAPE(1, false, true, 0);
            else if (bt0) CW_LAUNCH_TAPE(1, false, false, 0);
            else CW_LAUNCH_TAPE(1, false, false, -1);
        }
#undef CW_LAUNCH_TAPE
    }
    CU(cudaEventRecord(b->ev[1], b->stream));
    b->compact_valid = false;  // witness rows are slots [0, n_witness) of each instance: nothing to gather
    CU(cudaEventRecord(b->ev[2], b->stream));
    CU(cudaGetLastError());
    b->ran = true;
    return CW_OK;
}

int cw_batch_sync(cw_batch *b) {
    if (!b) return fail(CW_EINVAL, "null argument");
    CU(cudaSetDevice(b->device));
    CU(cudaStreamSynchronize(b->stream));
    return CW_OK;
}

int cw_batch_status(cw_batch *b, int32_t *status) {
    if (!b || !status) return fail(CW_EINVAL, "null argument");
    if (!b->ran) return fail(CW_ESTATE, "batch has not been run");
    CU(cudaSetDevice(b->device));
    std::vector<u32> fa(b->batch);
    std::vector<int> er(b->batch);
    CU(cudaMemcpyAsync(fa.data(), b->first_assert_d, (size_t)b->batch * 4, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaMemcpyAsync(er.data(), b->err_d, (size_t)b->batch * 4, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    for (u32 i = 0; i < b->batch; ++i) {
        if (er[i]) status[i] = -1;  // division by zero: the reference process aborts inside GMP
        else status[i] = fa[i] == 0xFFFFFFFFu ? 0 : (int32_t)(fa[i] + 1);
    }
    return CW_OK;
}

// contiguous [batch][n_witness] copy in device memory (only needed for tile layouts with BT > 1 or when
// a caller insists on a dense device array)
static int compact_witness(cw_batch *b) {
    if (b->compact_valid) return CW_OK;
    const Tape &t = b->c->tape;
    if (!b->witness_d) CU(cudaMalloc((void **)&b->witness_d, (size_t)b->batch * t.n_witness * 32));
    if (b->bt_log2 == 0) {
        CU(cudaMemcpy2DAsync(b->witness_d, (size_t)t.n_witness * 32, b->slots, (size_t)t.n_slots * 32,
                             (size_t)t.n_witness * 32, b->batch, cudaMemcpyDeviceToDevice, b->stream));
    } else {
        u32 tiles = b->batch_padded >> b->bt_log2;
        size_t total = (size_t)tiles * t.n_witness << b->bt_log2;
        u32 grid = (u32)std::min<size_t>((total + 255) / 256, 148 * 16);
        witness_compact_kernel<<<grid, 256, 0, b->stream>>>(b->slots, b->witness_d, t.n_slots, (u32)t.n_witness, b->batch, b->bt_log2);
        CU(cudaGetLastError());
    }
    b->compact_valid = true;
    return CW_OK;
}

// Packed transfer: entries the lowering proved to be one bit / <= 64 bits cross PCIe as that, the host
// expands them to the canonical 32-byte rows (zero-extension only - no field arithmetic happens on the CPU).
// For circuits made of bit decompositions this cuts the device->host bytes by an order of magnitude; the
// expansion runs on a few host threads at memory speed.  CW_PACKED_D2H=0 forces the plain pitched copy.
static int get_witness_packed(cw_batch *b, uint64_t *out, bool *done) {
    const Tape &t = b->c->tape;
    *done = false;
    const size_t n0 = t.pk_bit_wire.size(), n1 = t.pk_u64_wire.size(), n2 = t.pk_full_wire.size();
    const size_t bit_words = (n0 + 31) / 32;
    size_t words = bit_words + 2 * n1 + 8 * n2;
    words = (words + 3) & ~(size_t)3;
    if (b->bt_log2 != 0 || env_int("CW_PACKED_D2H", 1) == 0 || words * 4 * 2 > (size_t)t.n_witness * 32) return CW_OK;
    const size_t bytes = words * 4 * b->batch;
    if (!b->packed_d) {
        CU(cudaMalloc((void **)&b->packed_d, bytes));
        CU(cudaMallocHost((void **)&b->packed_h, bytes));
        CU(cudaMalloc((void **)&b->pack_flag_d, 4));
    }
    CU(cudaMemsetAsync(b->pack_flag_d, 0, 4, b->stream));
    dim3 grid((u32)std::min<size_t>((bit_words + n1 + n2 + 255) / 256, 148 * 4), std::min<u32>(b->batch, 65535u));
    if (grid.x == 0) grid.x = 1;
    witness_pack_kernel<<<grid, 256, 0, b->stream>>>(b->slots, t.n_slots, b->dt.pk_bit, (u32)n0, b->dt.pk_u64, (u32)n1,
                                                     b->dt.pk_full, (u32)n2, b->packed_d, words, b->batch, b->pack_flag_d);
    CU(cudaGetLastError());
    int flag = 0;
    CU(cudaMemcpyAsync(b->packed_h, b->packed_d, bytes, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaMemcpyAsync(&flag, b->pack_flag_d, 4, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    if (flag) return CW_OK;  // a value exceeded its static class (never expected): caller does the plain copy
    b->last_d2h_bytes = bytes;
    // host-side expansion: one sequential pass per instance over the witness entries, each 32-byte row
    // written exactly once with streaming stores (the three packed streams are in witness order).  The pass
    // is bound by host memory write bandwidth; measured on the B200 host: 8 threads 3.4 k witnesses/s,
    // 16 threads 3.2 k, 32 threads 2.6 k (37.9 MB rows), against 1.27 k for the plain PCIe copy.
    const uint8_t *cls = t.wit_class.data();
    const size_t W = t.n_witness;
    unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)env_int("CW_UNPACK_THREADS", 8)));
    nt = std::min<unsigned>(nt, b->batch);
    const bool aligned = (((uintptr_t)out) & 15u) == 0;
    std::vector<std::thread> th;
    for (unsigned tid = 0; tid < nt; ++tid)
        th.emplace_back([=]() {
            for (uint32_t inst = tid; inst < b->batch; inst += nt) {
                const uint32_t *p = b->packed_h + (size_t)inst * words;
                const uint32_t *pu = p + bit_words, *pf = pu + 2 * n1;
                uint64_t *row = out + (size_t)inst * W * 4;
                size_t bi = 0, ui = 0, fi = 0;
                const __m128i zero = _mm_setzero_si128();
                for (size_t i = 0; i < W; ++i) {
                    __m128i lo, hi = zero;
                    const uint8_t cl = cls[i];
                    if (cl == 0) {
                        lo = _mm_cvtsi64_si128((long long)((p[bi >> 5] >> (bi & 31)) & 1u));
                        ++bi;
                    } else if (cl == 1) {
                        lo = _mm_cvtsi64_si128((long long)((uint64_t)pu[2 * ui] | ((uint64_t)pu[2 * ui + 1] << 32)));
                        ++ui;
                    } else {
                        lo = _mm_loadu_si128((const __m128i *)(pf + 8 * fi));
                        hi = _mm_loadu_si128((const __m128i *)(pf + 8 * fi + 4));
                        ++fi;
                    }
                    if (aligned) {
                        _mm_stream_si128((__m128i *)(row + 4 * i), lo);
                        _mm_stream_si128((__m128i *)(row + 4 * i + 2), hi);
                    } else {
                        _mm_storeu_si128((__m128i *)(row + 4 * i), lo);
                        _mm_storeu_si128((__m128i *)(row + 4 * i + 2), hi);
                    }
                }
            }
            _mm_sfence();
        });
    for (auto &x : th) x.join();
    *done = true;
    return CW_OK;
}

int cw_batch_get_witness(cw_batch *b, uint64_t *out) {
    if (!b || !out) return fail(CW_EINVAL, "null argument");
    if (!b->ran) return fail(CW_ESTATE, "batch has not been run");
    const Tape &t = b->c->tape;
    CU(cudaSetDevice(b->device));
    bool done = false;
    int rc = get_witness_packed(b, out, &done);
    if (rc) return rc;
    if (done) return CW_OK;
    b->last_d2h_bytes = (uint64_t)b->batch * t.n_witness * 32;
    if (b->bt_log2 == 0) {  // rows are read in place: pitched device-to-host copy
        CU(cudaMemcpy2DAsync(out, (size_t)t.n_witness * 32, b->slots, (size_t)t.n_slots * 32, (size_t)t.n_witness * 32,
                             b->batch, cudaMemcpyDeviceToHost, b->stream));
    } else {
        rc = compact_witness(b);
        if (rc) return rc;
        CU(cudaMemcpyAsync(out, b->witness_d, (size_t)b->batch * t.n_witness * 32, cudaMemcpyDeviceToHost, b->stream));
    }
    CU(cudaStreamSynchronize(b->stream));
    return CW_OK;
}

uint64_t cw_batch_last_d2h_bytes(const cw_batch *b) { return b ? b->last_d2h_bytes : 0; }

int cw_batch_witness_device(cw_batch *b, const uint64_t **dptr) {
    if (!b || !dptr) return fail(CW_EINVAL, "null argument");
    if (!b->ran) return fail(CW_ESTATE, "batch has not been run");
    CU(cudaSetDevice(b->device));
    int rc = compact_witness(b);
    if (rc) return rc;
    *dptr = (const uint64_t *)b->witness_d;
    return CW_OK;
}

int cw_batch_witness_strided(cw_batch *b, const uint64_t **dptr, uint64_t *stride_elems) {
    if (!b || !dptr || !stride_elems) return fail(CW_EINVAL, "null argument");
    if (!b->ran) return fail(CW_ESTATE, "batch has not been run");
    if (b->bt_log2 == 0) {
        *dptr = (const uint64_t *)b->slots;
        *stride_elems = b->c->tape.n_slots;
        return CW_OK;
    }
    int rc = cw_batch_witness_device(b, dptr);
    *stride_elems = b->c->tape.n_witness;
    return rc;
}

void *cw_batch_stream(cw_batch *b) { return b ? (void *)b->stream : nullptr; }

int cw_batch_last_ms(cw_batch *b, float *exec_ms, float *gather_ms) {
    if (!b || !b->ran) return fail(CW_ESTATE, "batch has not been run");
    CU(cudaSetDevice(b->device));
    CU(cudaEventSynchronize(b->ev[2]));
    if (exec_ms) CU(cudaEventElapsedTime(exec_ms, b->ev[0], b->ev[1]));
    if (gather_ms) *gather_ms = 0.f;  // no gather pass: witness rows are written in place by the tape
    return CW_OK;
}

int cw_batch_wtns_bytes(cw_batch *b, uint32_t inst, uint8_t *out, size_t cap, size_t *len) {
    if (!b || inst >= b->batch) return fail(CW_EINVAL, "bad instance");
    if (!b->ran) return fail(CW_ESTATE, "batch has not been run");
    const Tape &t = b->c->tape;
    size_t need = 76 + 32 * (size_t)t.n_witness;
    if (len) *len = need;
    if (!out) return CW_OK;
    if (cap < need) return fail(CW_EINVAL, "buffer too small");
    CU(cudaSetDevice(b->device));
    std::vector<uint64_t> w((size_t)t.n_witness * 4);
    const uint4 *row;
    if (b->bt_log2 == 0) row = b->slots + (size_t)inst * t.n_slots * 2;
    else {
        int rc = compact_witness(b);
        if (rc) return rc;
        row = b->witness_d + (size_t)inst * t.n_witness * 2;
    }
    CU(cudaMemcpyAsync(w.data(), row, (size_t)t.n_witness * 32, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    std::vector<uint8_t> bytes = wtns_bytes(t.F, w.data(), t.n_witness);
    memcpy(out, bytes.data(), need);
    return CW_OK;
}

int cw_batch_write_wtns(cw_batch *b, uint32_t inst, const char *path) {
    size_t need = 0;
    int rc = cw_batch_wtns_bytes(b, inst, nullptr, 0, &need);
    if (rc) return rc;
    std::vector<uint8_t> buf(need);
    if ((rc = cw_batch_wtns_bytes(b, inst, buf.data(), need, &need))) return rc;
    FILE *f = fopen(path, "wb");
    if (!f) return fail(CW_EIO, std::string("cannot open ") + path);
    size_t wr = fwrite(buf.data(), 1, need, f);
    fclose(f);
    return wr == need ? CW_OK : fail(CW_EIO, "short write");
}

// ---- R1CS -------------------------------------------------------------------------------------
int cw_r1cs_from_circuit(const cw_circuit *c, cw_r1cs **out) {
    if (!c || !out) return fail(CW_EINVAL, "null argument");
    cw_r1cs *r = new cw_r1cs();
    r->data = c->tape.r1cs;
    r->F = c->tape.F;
    *out = r;
    return CW_OK;
}
int cw_r1cs_load(const char *path, cw_r1cs **out) {
    if (!path || !out) return fail(CW_EINVAL, "null argument");
    cw_r1cs *r = new cw_r1cs();
    try {
        read_r1cs(path, r->data);
    } catch (const std::exception &e) {
        delete r;
        return fail(CW_EFORMAT, e.what());
    }
    r->F = make_field(r->data.prime_id);
    *out = r;
    return CW_OK;
}
int cw_r1cs_write(const cw_r1cs *r, const char *path, uint32_t n_pub_out, uint32_t n_pub_in, uint32_t n_prv_in) {
    if (!r || !path) return fail(CW_EINVAL, "null argument");
    try {
        R1csData d = r->data;  // CW_KEEP: the count the circuit / the loaded file carries
        if (n_pub_out != CW_KEEP) d.n_pub_out = n_pub_out;
        if (n_pub_in != CW_KEEP) d.n_pub_in = n_pub_in;
        if (n_prv_in != CW_KEEP) d.n_prv_in = n_prv_in;
        write_r1cs(d, r->F, path);
    } catch (const std::exception &e) {
        return fail(CW_EIO, e.what());
    }
    return CW_OK;
}
int cw_r1cs_info(const cw_r1cs *r, uint64_t *n_wires, uint64_t *n_constraints, uint64_t *nnz, int *prime_id) {
    if (!r) return fail(CW_EINVAL, "null argument");
    if (n_wires) *n_wires = r->data.n_wires;
    if (n_constraints) *n_constraints = r->data.n_constraints;
    if (nnz) *nnz = r->data.col.size();
    if (prime_id) *prime_id = r->data.prime_id;
    return CW_OK;
}
void cw_r1cs_destroy(cw_r1cs *r) {
    if (!r) return;
    for (auto &kv : r->dev) {
        cudaSetDevice(kv.first);
        cudaFree(kv.second.row_ptr);
        cudaFree(kv.second.terms);
        cudaFree(kv.second.dictM);
        cudaFree(kv.second.perm);
        cudaFree(kv.second.bool_wire);
        cudaFree(kv.second.bool_row);
    }
    delete r;
}

int cw_r1cs_check(cw_r1cs *r, const uint64_t *witness, int is_device_ptr, uint32_t batch, int device,
                  int64_t *first_bad, float *kernel_ms) {
    if (!r) return fail(CW_EINVAL, "bad argument");
    return cw_r1cs_check_strided(r, witness, r->data.n_wires, is_device_ptr, batch, device, first_bad, kernel_ms);
}

int cw_r1cs_check_strided(cw_r1cs *r, const uint64_t *witness, uint64_t stride_elems, int is_device_ptr, uint32_t batch,
                          int device, int64_t *first_bad, float *kernel_ms) {
    if (!r || !witness || !first_bad || batch == 0 || stride_elems < r->data.n_wires) return fail(CW_EINVAL, "bad argument");
    if (is_device_ptr && ((uintptr_t)witness & 31u))
        return fail(CW_EINVAL, "device witness pointer must be 32-byte aligned (elements are read with 256-bit loads)");
    int rc = ensure_device(device);
    if (rc) return rc;
    DevR1cs d;
    {
        std::lock_guard<std::mutex> lk(r->mu);
        auto it = r->dev.find(device);
        if (it == r->dev.end()) {
            const R1csData &R = r->data;
            std::vector<U256> dm(R.dict.size());
            std::vector<unsigned short> kind(R.dict.size());
            auto pow2_exp = [](const U256 &v) -> int {  // k if v == 2^k, else -1
                int k = -1;
                for (int i = 0; i < 256; ++i)
                    if ((v.v[i >> 6] >> (i & 63)) & 1) {
                        if (k >= 0) return -1;
                        k = i;
                    }
                return k;
            };
            for (size_t i = 0; i < R.dict.size(); ++i) {
                dm[i] = r->F.to_mont(R.dict[i]);
                U256 negv;
                u256_sub(negv, r->F.q, R.dict[i]);
                int kp = pow2_exp(R.dict[i]), kn = R.dict[i].is_zero() ? -1 : pow2_exp(negv);
                if (kp == 0) kind[i] = 1;
                else if (kn == 0) kind[i] = 2;
                else if (kp > 0 && kp < 250) kind[i] = (unsigned short)(3 | (kp << 8));
                else if (kn > 0 && kn < 250) kind[i] = (unsigned short)(4 | (kn << 8));
                else kind[i] = 0;
            }
            // rows sorted by structure so that the rows of a warp have equal length and branch alike
            size_t m = R.n_constraints;
            std::vector<uint64_t> sig(m);
            for (size_t row = 0; row < m; ++row) {
                uint64_t na = R.row_ptr[3 * row + 1] - R.row_ptr[3 * row], nb = R.row_ptr[3 * row + 2] - R.row_ptr[3 * row + 1],
                         nc = R.row_ptr[3 * row + 3] - R.row_ptr[3 * row + 2];
                uint64_t h = 1469598103934665603ull;
                for (uint64_t k = R.row_ptr[3 * row]; k < R.row_ptr[3 * row + 3]; ++k) h = (h ^ (kind[R.coef[k]] & 0xFF)) * 1099511628211ull;
                uint64_t total = std::min<uint64_t>(na + nb + nc, 0xFFFF);
                sig[row] = (total << 48) | ((std::min<uint64_t>(na, 255)) << 40) | ((std::min<uint64_t>(nb, 255)) << 32) | (h & 0xFFFFFFFFull);
            }
            // boolean rows  x * (x - 1) = 0  (A = {x:1}, B = {x:1, one:-1}, C = {} or A/B swapped) are the bulk
            // of circom circuits (every Num2Bits / range-check bit): they only need `w[x] in {0,1}` and
            // get their own memory-bound kernel; the remaining rows go through the general kernel.
            std::vector<u32> perm, bool_wire, bool_row;
            auto is_unit = [&](uint64_t k, int want) { return (kind[R.coef[k]] & 0xFF) == want && (kind[R.coef[k]] >> 8) == 0; };
            for (size_t row = 0; row < m; ++row) {
                uint64_t p0 = R.row_ptr[3 * row], p1 = R.row_ptr[3 * row + 1], p2 = R.row_ptr[3 * row + 2], p3 = R.row_ptr[3 * row + 3];
                bool is_bool = false;
                u32 wire = 0;
                if (p3 == p2 && (p1 - p0) + (p2 - p1) == 3) {
                    uint64_t s0 = (p1 - p0 == 1) ? p0 : p1, l0 = (p1 - p0 == 1) ? p1 : p0;  // single-term block / two-term block
                    // two-term block is sorted by wire: {one: -1, x: +1}
                    if (is_unit(s0, 1) && R.col[s0] != 0 && R.col[l0] == 0 && is_unit(l0, 2) && R.col[l0 + 1] == R.col[s0] && is_unit(l0 + 1, 1)) {
                        is_bool = true;
                        wire = R.col[s0];
                    }
                }
                if (is_bool) { bool_wire.push_back(wire); bool_row.push_back((u32)row); }
                else perm.push_back((u32)row);
            }
            std::stable_sort(perm.begin(), perm.end(), [&](u32 x, u32 y) { return sig[x] > sig[y]; });
            // a boolean row whose wire is a term of a general row (the bit of a decomposition inside its
            // recomposition sum) is checked by that term's thread while the value is in registers: the witness
            // is then read once instead of twice
            std::vector<u32> term_bool(R.col.size(), 0xFFFFFFFFu);
            {
                std::vector<u32> wire2bool(R.n_wires, 0xFFFFFFFFu);
                for (size_t i = 0; i < bool_wire.size(); ++i)
                    if (wire2bool[bool_wire[i]] == 0xFFFFFFFFu) wire2bool[bool_wire[i]] = (u32)i;
                std::vector<uint8_t> absorbed(bool_wire.size(), 0);
                for (u32 row : perm)
                    for (uint64_t k = R.row_ptr[3 * (size_t)row]; k < R.row_ptr[3 * (size_t)row + 3]; ++k) {
                        u32 bi = wire2bool[R.col[k]];
                        if (bi != 0xFFFFFFFFu && !absorbed[bi]) {
                            absorbed[bi] = 1;
                            term_bool[k] = bool_row[bi];
                        }
                    }
                size_t o = 0;
                for (size_t i = 0; i < bool_wire.size(); ++i)
                    if (!absorbed[i]) { bool_wire[o] = bool_wire[i]; bool_row[o] = bool_row[i]; ++o; }
                bool_wire.resize(o);
                bool_row.resize(o);
            }
            {
                std::vector<uint4> terms(R.col.size());
                for (size_t k = 0; k < R.col.size(); ++k)
                    terms[k] = make_uint4(R.col[k], R.coef[k], kind[R.coef[k]], term_bool[k]);
                if ((rc = upload(&d.terms, terms.data(), terms.size() * sizeof(uint4)))) return rc;
            }
            d.n_general = (u32)perm.size();
            d.n_bool = (u32)bool_wire.size();
            {
                uint64_t terms_general = 0;
                for (u32 row : perm) terms_general += R.row_ptr[3 * (size_t)row + 3] - R.row_ptr[3 * (size_t)row];
                d.mean_row_terms = perm.empty() ? 0 : (u32)(terms_general / perm.size());
            }
            d.n_long = 0;  // perm is sorted by decreasing term count
            while (d.n_long < d.n_general && (sig[perm[d.n_long]] >> 48) >= R1CS_SPLIT_MIN) ++d.n_long;
            if ((rc = upload(&d.bool_wire, bool_wire.data(), bool_wire.size() * 4))) return rc;
            if ((rc = upload(&d.bool_row, bool_row.data(), bool_row.size() * 4))) return rc;
            if ((rc = upload(&d.row_ptr, R.row_ptr.data(), R.row_ptr.size() * 8))) return rc;
            if ((rc = upload(&d.dictM, dm.data(), dm.size() * 32))) return rc;
            if ((rc = upload(&d.perm, perm.data(), perm.size() * 4))) return rc;
            r->dev[device] = d;
        } else d = it->second;
    }
    const R1csData &R = r->data;
    const uint4 *w_d = (const uint4 *)witness;
    uint4 *tmp = nullptr;
    if (!is_device_ptr) {
        CU(cudaMalloc((void **)&tmp, (size_t)batch * R.n_wires * 32));
        CU(cudaMemcpy2D(tmp, (size_t)R.n_wires * 32, witness, (size_t)stride_elems * 32, (size_t)R.n_wires * 32, batch,
                        cudaMemcpyHostToDevice));
        w_d = tmp;
        stride_elems = R.n_wires;
    }
    unsigned long long *fb_d = nullptr;
    CU(cudaMalloc((void **)&fb_d, (size_t)batch * 8));
    CU(cudaMemset(fb_d, 0xFF, (size_t)batch * 8));
    R1csDev rd;
    rd.row_ptr = d.row_ptr;
    rd.terms = d.terms;
    rd.dictM = d.dictM;
    rd.perm = d.perm;
    rd.n_wires = (u32)R.n_wires;
    rd.w_stride = stride_elems;
    // general rows: the long ones (perm[0, n_long)) by lane groups, the rest one thread per (row, instance)
    // (lane groups are opt-in, env CW_R1CS_SPLIT=1: measured 20.7 ms against 18.4 ms for the one-thread-per-row
    // kernel on the bench circuit, batch 1024 - the butterfly costs more than the coalescing gains)
    const u32 n_long = env_int("CW_R1CS_SPLIT", 0) ? d.n_long : 0u, n_short = d.n_general - n_long;
    // instance groups: enough blocks to fill the GPU, as many instances per block as that allows
    auto plan = [&](u32 rows_per_block, u32 n_rows, u32 *row_blocks) {
        *row_blocks = (u32)std::min<uint64_t>(((uint64_t)n_rows + rows_per_block - 1) / rows_per_block, 148 * 8);
        if (*row_blocks == 0) *row_blocks = 1;
        u32 ipb = 1;
        while (ipb < 8 && (uint64_t)*row_blocks * ((batch + 2 * ipb - 1) / (2 * ipb)) >= 148ull * 16) ipb *= 2;
        ipb = (u32)std::max(1, env_int("CW_R1CS_IPB", (int)ipb));
        while ((batch + ipb - 1) / ipb > 65535u) ipb *= 2;  // grid.y limit: every instance must have a block
        return ipb;
    };
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0));
    if (n_long) {
        u32 row_blocks;
        rd.inst_per_block = plan(256 / R1CS_SPLIT_G, n_long, &row_blocks);
        rd.perm = d.perm;
        rd.n_constraints = n_long;
        dim3 grid(row_blocks, std::min<u32>((batch + rd.inst_per_block - 1) / rd.inst_per_block, 65535u));
        if (R.prime_id == 0) r1cs_check_split_kernel<0, R1CS_SPLIT_G><<<grid, 256>>>(rd, w_d, batch, fb_d);
        else r1cs_check_split_kernel<1, R1CS_SPLIT_G><<<grid, 256>>>(rd, w_d, batch, fb_d);
    }
    if (n_short) {
        u32 row_blocks;
        rd.inst_per_block = plan(256, n_short, &row_blocks);
        rd.perm = d.perm + n_long;
        rd.n_constraints = n_short;  // rows visited through perm
        dim3 grid(row_blocks, std::min<u32>((batch + rd.inst_per_block - 1) / rd.inst_per_block, 65535u));
        // long rows: many resident warps (48 registers); short rows: the unspilled build (78 registers)
        const bool lean = env_int("CW_R1CS_LEAN", d.mean_row_terms >= 12 ? 1 : 0) != 0;
        if (R.prime_id == 0) {
            if (lean) r1cs_check_kernel<0, 5><<<grid, 256>>>(rd, w_d, batch, fb_d);
            else r1cs_check_kernel<0, 3><<<grid, 256>>>(rd, w_d, batch, fb_d);
        } else {
            if (lean) r1cs_check_kernel<1, 5><<<grid, 256>>>(rd, w_d, batch, fb_d);
            else r1cs_check_kernel<1, 3><<<grid, 256>>>(rd, w_d, batch, fb_d);
        }
    }
    if (d.n_bool) {
        dim3 bgrid((u32)std::min<uint64_t>(((uint64_t)d.n_bool + 255) / 256, 148 * 8), std::min<u32>(batch, 65535u));
        r1cs_bool_kernel<<<bgrid, 256>>>(d.bool_wire, d.bool_row, d.n_bool, w_d, stride_elems, batch, fb_d);
    }
    CU(cudaEventRecord(e1));
    CU(cudaGetLastError());
    std::vector<unsigned long long> fb(batch);
    CU(cudaMemcpy(fb.data(), fb_d, (size_t)batch * 8, cudaMemcpyDeviceToHost));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    if (kernel_ms) *kernel_ms = ms;
    for (u32 i = 0; i < batch; ++i) first_bad[i] = fb[i] == ~0ull ? -1 : (int64_t)fb[i];
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(fb_d);
    if (tmp) cudaFree(tmp);
    return CW_OK;
}

// ---- field batch ops ---------------------------------------------------------------------------
int cw_fr_batch_op(int prime_id, int op, const uint64_t *a, const uint64_t *b, const uint64_t *c, uint64_t *r,
                   size_t n, int device) {
    if (!a || !r || prime_id < 0 || prime_id > 1) return fail(CW_EINVAL, "bad argument");
    int rc = ensure_device(device);
    if (rc) return rc;
    uint4 *A = nullptr, *B = nullptr, *C = nullptr, *Rr = nullptr;
    int *err = nullptr;
    if ((rc = upload(&A, a, n * 32))) return rc;
    if (b && (rc = upload(&B, b, n * 32))) return rc;
    if (c && (rc = upload(&C, c, n * 32))) return rc;
    CU(cudaMalloc((void **)&Rr, n * 32 + 32));
    CU(cudaMalloc((void **)&err, 4));
    CU(cudaMemset(err, 0, 4));
    u32 grid = (u32)std::min<size_t>((n + 127) / 128, 148 * 16);
    if (!grid) grid = 1;
    if (prime_id == 0) fr_batch_op_kernel<0><<<grid, 128>>>(op, A, B, C, Rr, n, err);
    else fr_batch_op_kernel<1><<<grid, 128>>>(op, A, B, C, Rr, n, err);
    CU(cudaGetLastError());
    CU(cudaMemcpy(r, Rr, n * 32, cudaMemcpyDeviceToHost));
    int herr = 0;
    CU(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
    cudaFree(A);
    cudaFree(B);
    cudaFree(C);
    cudaFree(Rr);
    cudaFree(err);
    return herr ? fail(CW_EINVAL, "division by zero in batch op") : CW_OK;
}

int cw_fr_mul_bench(int prime_id, size_t n, int iters, int device, float *ms) {
    int rc = ensure_device(device);
    if (rc) return rc;
    std::vector<uint64_t> h(n * 4);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (auto &x : h) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        x = s;
    }
    for (size_t i = 0; i < n; ++i) h[4 * i + 3] &= 0x0FFFFFFFFFFFFFFFull;
    uint4 *d = nullptr;
    if ((rc = upload(&d, h.data(), n * 32))) return rc;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    u32 grid = (u32)((n + 255) / 256);
    for (int rep = 0; rep < 2; ++rep) {
        CU(cudaEventRecord(e0));
        if (prime_id == 0) fr_mul_bench_kernel<0><<<grid, 256>>>(d, n, iters);
        else fr_mul_bench_kernel<1><<<grid, 256>>>(d, n, iters);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
    }
    CU(cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return CW_OK;
}

}  // extern "C"
