// Host side of the packed device->host transfer: the layout of the packed witness records, their expansion to the
// reference's 32-byte rows (zero-extension only - no field arithmetic happens on the CPU) and the worker threads
// that run it.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <vector>

#include "tape.h"

namespace cw {

// The witness of one instance arrives as the packed record written by witness_pack_kernel.  The expansion is
// described once per circuit by segments of consecutive witness entries that come from the same section.
struct PackSeg {
    uint32_t kind;   // 0 plane run, 1 bits outside the plane, 2 u64 entries, 3 full entries
    uint32_t start;  // first witness entry
    uint32_t count;
    uint32_t src;    // plane run: word * 32 + first bit; otherwise the index inside the section
};
struct PackLayout {
    std::vector<PackSeg> segs;
    std::vector<uint32_t> bit_loc, u64_loc, full_loc;  // slot ids of the entries outside the plane, in witness order
    size_t n_plane_words = 0, n_bit_words = 0, words = 0;  // words: per instance, rounded to 16 bytes
};
// cls: per witness entry 0 bit / 1 <= 64 bits / 2 full; nullptr: the classes the lowering proved (t.wit_class)
void build_pack_layout(const Tape &t, PackLayout &L, const uint8_t *cls = nullptr);

// one instance: packed record -> W rows of 32 bytes, each written exactly once, front to back, with streaming stores
// (force_bits: 0 = the widest stores the CPU has, else at most 128 / 256 / 512-bit stores - for tests)
void expand_record(const PackLayout &L, const uint32_t *rec, uint64_t *row_out, int force_bits = 0);
const char *expand_isa();  // "avx512" / "avx2" / "sse2": the store width expand_record uses on this CPU

// Persistent worker threads (created on first use, shared by all batches of the process).  The expansion writes
// 32 bytes per witness entry - tens of MB per instance - so it runs at the speed of the host's memory system and
// placement decides that speed on a multi-socket host: the workers are pinned to NUMA nodes (round robin over the
// nodes in use) and work items are assigned STATICALLY (item key -> worker), so that the rows of instance i of a
// caller's buffer are always written by the same worker; the first pass places those pages on that worker's node
// (first touch), every later pass over the same buffer writes node-local memory.
class Pool {
  public:
    // node_hint >= 0 with several ranks on the host (LOCAL_WORLD_SIZE > 1): all workers on that node (the GPU's)
    static Pool &get(int node_hint = -1);
    unsigned size() const;
    // runs fn(i) for i in [0, n) on the workers, item i on worker (key0 + i) % size(); returns when all are done
    void parallel_for(size_t n, size_t key0, const std::function<void(size_t)> &fn);
    const char *describe() const;  // e.g. "32 threads over 2 NUMA nodes, static"

  private:
    explicit Pool(int node_hint);
    ~Pool();
    struct Impl;
    Impl *p_;
};

}  // namespace cw
