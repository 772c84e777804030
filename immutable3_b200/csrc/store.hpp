// store.hpp — host-side model of an opened data directory: the SegmentManager replacement
// (SegmentManager.scala:20-112) restricted to this handle's canonical segment slice.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "common.hpp"

namespace imm3 {

struct ColumnStore {
    ColumnMeta meta;
    std::vector<SegmentFile> segs;  // owned slice, canonical (file-name-sorted) order
    int64_t encoded_bytes = 0;      // sum of the slice's block bytes == bytes resident in HBM
    // device side (engine.cu)
    uint8_t* d_arena = nullptr;
    size_t arena_bytes = 0;
    uint8_t* h_mirror = nullptr;     // pinned copy of the slice (IMM3_OPEN_KEEP_HOST)
    std::vector<uint32_t> word_off;  // PFOR_INT: nblocks+1 offsets in 32-bit words into the arena
    uint32_t* d_word_off = nullptr;
    int64_t max_block_words = 0;     // PFOR_INT: largest encoded block, in 32-bit words
    int64_t max_tile_bytes = 0;      // PFOR_INT: largest 8-block tile as the filter kernel stages it (16-byte aligned start and size)
    void* d_stats = nullptr;         // PFOR_INT: BlockStat per block (exact min / max), computed on the GPU at open
    int64_t max_tile32_bytes = 0;    // PFOR_INT: largest 32-block tile (the quad filter kernel's CTA tile)
};

struct TableStore {
    TableMeta meta;
    std::string dir;
    int nsegments = 0;              // whole table (segments of the first column, SegmentManager.scala:94-99)
    int seg_begin = 0, seg_end = 0; // owned canonical slice
    std::vector<int> file_ids;      // canonical position -> numeric id, whole table
    std::vector<ColumnStore> cols;
    int64_t nrows = 0;
    int64_t nblocks = 0;
    std::vector<uint64_t> row_start;  // nblocks+1 canonical row ordinals of the slice's blocks
    int max_block_rows = 0;
    uint64_t* d_row_start = nullptr;
};

// Canonical slice of `n` segments owned by `rank` of `world`.
inline void shard_range(int n, int rank, int world, int* begin, int* end) {
    *begin = (int)((int64_t)rank * n / world);
    *end = (int)((int64_t)(rank + 1) * n / world);
}

// Discover + validate every table of data_dir; fills everything except device pointers.
// `visit(col, seg_index_in_slice, bytes, nbytes)` is not used here: block contents are only
// inspected for PFOR_INT headers.
int load_tables(const std::string& data_dir, int rank, int world, std::vector<TableStore>* out);

// Validate one PFOR_INT block (big-endian words + 8 pad bytes) and return its value count.
int pfor_validate_block(const uint8_t* bytes, int64_t nbytes, int32_t* n_out);

// Read-only mapping of a file prefix.
struct FileMap {
    const uint8_t* data = nullptr;
    size_t len = 0;
    int open(const std::string& path, size_t need);
    void close();
    ~FileMap() { close(); }
};

}  // namespace imm3
