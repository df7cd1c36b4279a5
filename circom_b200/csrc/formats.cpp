// Byte-compatible file formats: .r1cs (read / write), .wtns (write), .dat (write).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <unordered_map>

#include "tape.h"

namespace cw {

namespace {
struct File {
    FILE *f;
    File(const std::string &p, const char *mode) : f(fopen(p.c_str(), mode)) {
        if (!f) throw std::runtime_error("cannot open " + p);
    }
    ~File() {
        if (f) fclose(f);
    }
    void w(const void *p, size_t n) {
        if (n && fwrite(p, 1, n, f) != n) throw std::runtime_error("write failed");
    }
    template <class T>
    void put(T v) {
        w(&v, sizeof(T));
    }
};

// minimal little-endian byte string of a wire id, as BigInt::to_bytes_le gives it
// (constraint_writers/src/r1cs_writer.rs:262-266); [0] for zero.
inline int le_len(uint32_t x) { return x >> 24 ? 4 : x >> 16 ? 3 : x >> 8 ? 2 : 1; }
inline bool wire_less(uint32_t a, uint32_t b) {
    int la = le_len(a), lb = le_len(b);
    for (int i = 0; i < std::min(la, lb); ++i) {
        uint8_t ba = (a >> (8 * i)) & 0xFF, bb = (b >> (8 * i)) & 0xFF;
        if (ba != bb) return ba < bb;
    }
    return la < lb;
}
}  // namespace

// Section order and header layout follow constraint_list/src/r1cs_porting.rs:19-53 and
// constraint_writers/src/r1cs_writer.rs:6-14,49-72,93-101,246-269,328-341.
void write_r1cs(const R1csData &r, const FieldParams &F, const std::string &path) {
    File f(path, "wb");
    f.w("r1cs", 4);
    f.put<uint32_t>(1);
    f.put<uint32_t>(3);
    // constraints section first
    uint64_t nnz = r.col.size();
    uint64_t m = r.n_constraints;
    uint64_t csize = 3 * m * 4 + nnz * (4 + 32);
    f.put<uint32_t>(2);
    f.put<uint64_t>(csize);
    std::vector<uint32_t> order;
    for (uint64_t k = 0; k < 3 * m; ++k) {
        uint64_t b = r.row_ptr[k], e = r.row_ptr[k + 1];
        f.put<uint32_t>((uint32_t)(e - b));
        order.resize(e - b);
        for (uint64_t i = b; i < e; ++i) order[i - b] = (uint32_t)i;
        std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return wire_less(r.col[x], r.col[y]); });
        for (uint32_t i : order) {
            f.put<uint32_t>(r.col[i]);
            f.w(r.dict[r.coef[i]].v, 32);
        }
    }
    // header
    f.put<uint32_t>(1);
    f.put<uint64_t>(4 + 32 + 4 * 4 + 8 + 4);
    f.put<uint32_t>(32);
    f.w(F.q.v, 32);
    f.put<uint32_t>((uint32_t)r.n_wires);
    f.put<uint32_t>(r.n_pub_out);
    f.put<uint32_t>(r.n_pub_in);
    f.put<uint32_t>(r.n_prv_in);
    f.put<uint64_t>(r.n_wires);  // number of labels
    f.put<uint32_t>((uint32_t)m);
    // wire -> label map
    f.put<uint32_t>(3);
    f.put<uint64_t>(8 * r.n_wires);
    for (uint64_t i = 0; i < r.n_wires; ++i) f.put<uint64_t>(i);
}

void read_r1cs(const std::string &path, R1csData &out) {
    File f(path, "rb");
    fseek(f.f, 0, SEEK_END);
    long sz = ftell(f.f);
    fseek(f.f, 0, SEEK_SET);
    std::vector<uint8_t> buf(sz);
    if (sz && fread(buf.data(), 1, sz, f.f) != (size_t)sz) throw std::runtime_error("r1cs: read failed");
    auto need = [&](size_t off, size_t n) {
        if (n > buf.size() || off > buf.size() - n) throw std::runtime_error("r1cs: truncated file");  // (no wrap for n near 2^64)
    };
    auto u32 = [&](size_t off) { need(off, 4); uint32_t v; memcpy(&v, &buf[off], 4); return v; };
    auto u64 = [&](size_t off) { need(off, 8); uint64_t v; memcpy(&v, &buf[off], 8); return v; };
    need(0, 12);
    if (memcmp(buf.data(), "r1cs", 4)) throw std::runtime_error("r1cs: bad magic");
    if (u32(4) != 1) throw std::runtime_error("r1cs: unsupported version");
    uint32_t nsec = u32(8);
    size_t pos = 12, hdr = 0, cons = 0, hdr_len = 0, cons_len = 0;
    for (uint32_t s = 0; s < nsec; ++s) {
        uint32_t ty = u32(pos);
        uint64_t len = u64(pos + 4);
        pos += 12;
        need(pos, len);
        if (ty == 1) { hdr = pos; hdr_len = len; }
        if (ty == 2) { cons = pos; cons_len = len; }
        pos += len;
    }
    if (!hdr || !cons) throw std::runtime_error("r1cs: missing header or constraint section");
    need(hdr, 4 + 32 + 28);
    uint32_t fs = u32(hdr);
    if (fs != 32 || hdr_len < 4 + 32 + 28) throw std::runtime_error("r1cs: only 32-byte fields are supported");
    U256 q;
    memcpy(q.v, &buf[hdr + 4], 32);
    FieldParams f0 = make_field(0), f1 = make_field(1);
    if (q == f0.q) out.prime_id = 0;
    else if (q == f1.q) out.prime_id = 1;
    else throw std::runtime_error("r1cs: unsupported prime");
    size_t h = hdr + 4 + 32;
    out.n_wires = u32(h);
    out.n_pub_out = u32(h + 4);
    out.n_pub_in = u32(h + 8);
    out.n_prv_in = u32(h + 12);
    out.n_constraints = u32(h + 24);
    out.row_ptr.assign(1, 0);
    out.col.clear();
    out.coef.clear();
    out.dict.clear();
    std::unordered_map<std::string, uint32_t> idx;
    size_t p = cons, end = cons + cons_len;
    for (uint64_t k = 0; k < 3 * out.n_constraints; ++k) {
        if (p + 4 > end) throw std::runtime_error("r1cs: constraint section too short");
        uint32_t n = u32(p);
        p += 4;
        if (p + (size_t)n * 36 > end) throw std::runtime_error("r1cs: constraint section too short");
        size_t first = out.col.size();
        for (uint32_t j = 0; j < n; ++j) {
            uint32_t w = u32(p);
            if (w >= out.n_wires) throw std::runtime_error("r1cs: wire id out of range");
            std::string key((const char *)&buf[p + 4], 32);
            auto it = idx.find(key);
            uint32_t id;
            if (it == idx.end()) {
                U256 c;
                memcpy(c.v, key.data(), 32);
                id = (uint32_t)out.dict.size();
                out.dict.push_back(c);
                idx.emplace(std::move(key), id);
            } else id = it->second;
            out.col.push_back(w);
            out.coef.push_back(id);
            p += 36;
        }
        // numeric order inside the CSR row (files carry the byte-string order)
        std::vector<std::pair<uint32_t, uint32_t>> tmp;
        for (size_t i = first; i < out.col.size(); ++i) tmp.emplace_back(out.col[i], out.coef[i]);
        std::sort(tmp.begin(), tmp.end());
        for (size_t i = 0; i < tmp.size(); ++i) {
            out.col[first + i] = tmp[i].first;
            out.coef[first + i] = tmp[i].second;
        }
        out.row_ptr.push_back(out.col.size());
    }
    if (out.dict.empty()) out.dict.push_back(u256_from_u64(0));
}

// writeBinWitness (c_elements/common/main.cpp:288-334)
std::vector<uint8_t> wtns_bytes(const FieldParams &F, const uint64_t *witness, uint64_t n_witness) {
    std::vector<uint8_t> o(76 + 32 * n_witness);
    uint8_t *p = o.data();
    auto put32 = [&](uint32_t v) { memcpy(p, &v, 4); p += 4; };
    auto put64 = [&](uint64_t v) { memcpy(p, &v, 8); p += 8; };
    memcpy(p, "wtns", 4);
    p += 4;
    put32(2);
    put32(2);
    put32(1);
    put64(8 + 32);
    put32(32);
    memcpy(p, F.q.v, 32);
    p += 32;
    put32((uint32_t)n_witness);
    put32(2);
    put64(32ull * n_witness);
    memcpy(p, witness, 32 * n_witness);
    return o;
}

// generate_dat_file (c_code_generator.rs:818-865) for a circuit without run-time constants / io maps:
// the input hash map followed by the witness -> signal list.
void write_dat(const Tape &t, const std::string &path) {
    File f(path, "wb");
    for (const HashEntry &e : t.hashmap) {
        f.put<uint64_t>(e.hash);
        f.put<uint64_t>(e.signalid);
        f.put<uint64_t>(e.signalsize);
    }
    for (uint64_t i = 0; i < t.n_witness; ++i) f.put<uint64_t>(t.witness2signal[i]);
}

}  // namespace cw
