// engine.cu — the C ABI of include/imm3.h: SegmentManager upload path + query execution.
//
//  imm3_open        = new SegmentManager(dataDir)            SegmentManager.scala:20-112
//                     + staging of every owned segment into HBM through pinned async copies
//  imm3_query*      = Engine.execute, Project branch         Engine.scala:158-198
//                     = ScanOp -> SelectOp* -> ProjectOp      Scan.scala, Select.scala, Project.scala
//  result accessors = Iterator[Row] / Row                    Project.scala:17-81, Record.scala:3-14
//
// HBM layout: one arena per (table, column) holding the owned segments' block bytes back to back in
// canonical order.  For DENSE_* columns that is a flat array of values indexed by the canonical row
// ordinal (blocks and segments need no per-block metadata on the device); the arena is padded to a
// whole tile so the last TMA bulk copy stays in bounds.  For PFOR_INT columns it is the stream of
// big-endian words plus a word offset per block.
//
// There is no CPU execution path in this file: every query runs the CUDA kernels of kernels.cu.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "kernels.hpp"
#include "plan.hpp"
#include "store.hpp"

using namespace imm3;

#define CUDA_TRY(expr)                                                                                     \
    do {                                                                                                   \
        cudaError_t e_ = (expr);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return fail(e_ == cudaErrorMemoryAllocation ? IMM3_ERR_OOM : IMM3_ERR_CUDA, "%s: %s (%s:%d)", #expr, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                       \
    } while (0)

namespace {

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
};

inline double now_us() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec * 1e6 + (double)ts.tv_nsec * 1e-3;
}
thread_local double g_t_launched = 0, g_t_synced = 0;  // stamps taken inside run_scan_once (last launch queued / GPU done)
thread_local bool g_eager_times = false;               // two-phase LIMIT queries add their phases' device times up: no lazy reading

// Grow-only cache of device / pinned-host buffers: result columns are handed back on
// imm3_result_free and reused by the next query (cudaMalloc / cudaMallocHost cost milliseconds).
struct BufPool {
    bool pinned_host = false;
    std::vector<Buf> free_list;
    int acquire(size_t bytes, Buf* out) {
        if (bytes < 256) bytes = 256;
        int best = -1;
        for (size_t i = 0; i < free_list.size(); i++)
            if (free_list[i].cap >= bytes && (best < 0 || free_list[i].cap < free_list[(size_t)best].cap)) best = (int)i;
        if (best >= 0) {
            *out = free_list[(size_t)best];
            free_list.erase(free_list.begin() + best);
            return 0;
        }
        // drop the largest cached buffer that is too small, so the pool does not accumulate
        if (!free_list.empty()) {
            size_t big = 0;
            for (size_t i = 1; i < free_list.size(); i++)
                if (free_list[i].cap > free_list[big].cap) big = i;
            if (pinned_host) cudaFreeHost(free_list[big].p); else cudaFree(free_list[big].p);
            free_list.erase(free_list.begin() + (long)big);
        }
        size_t cap = bytes + bytes / 8;
        cap = (cap + 255) & ~(size_t)255;
        void* p = nullptr;
        cudaError_t e = pinned_host ? cudaMallocHost(&p, cap) : cudaMalloc(&p, cap);
        if (e != cudaSuccess) {
            cudaGetLastError();
            release_all();
            e = pinned_host ? cudaMallocHost(&p, cap) : cudaMalloc(&p, cap);
        }
        if (e != cudaSuccess)
            return fail(IMM3_ERR_OOM, "%s of %zu bytes failed: %s", pinned_host ? "cudaMallocHost" : "cudaMalloc", cap,
                        cudaGetErrorString(e));
        out->p = p;
        out->cap = cap;
        return 0;
    }
    void release(Buf b) {
        if (b.p) free_list.push_back(b);
    }
    void release_all() {
        for (auto& b : free_list) {
            if (pinned_host) cudaFreeHost(b.p); else cudaFree(b.p);
        }
        free_list.clear();
    }
};

}  // namespace

struct imm3_db {
    std::string dir;
    int device = 0, rank = 0, world = 1;
    uint32_t flags = 0;
    bool host_only = false;
    std::vector<TableStore> tables;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // imm3_result_fetch_async: device->host copies overlap the next query's staging
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_mid = nullptr;
    ScanCtrl* d_ctrl = nullptr;   // first member of a device CtrlBlock (ScanCtrl + CommOut)
    CtrlBlock* h_ctrl = nullptr;  // pinned copy, refreshed once per launch sequence
    unsigned long long pub_seq = 0;  // publish (plan.hpp): sequence number of the last query whose last kernel writes h_ctrl itself
    unsigned long long ev_seq = 0, query_seq = 0;  // the query ev0 / ev1 were recorded for (device times are read lazily)
    int live_results = 0;         // imm3_result objects that still point at this db (imm3_close refuses while > 0)
    // Count exchange over NVLink peer memory (imm3_comm_*): the local mailbox, every rank's mailbox as mapped here.
    unsigned long long* d_mailbox = nullptr;
    unsigned long long* comm_peer[kMaxWorld] = {};
    bool comm_on = false;
    uint32_t comm_epoch = 0;
    unsigned long long comm_timeout_ns = 30000000000ull;
    unsigned long long* d_status = nullptr;
    size_t status_cap = 0;
    uint32_t epoch = 0;
    int num_sms = 0;
    BufPool dev_pool, host_pool;
    Buf d_bitmap, h_bitmap;
    Buf d_span_cnt, d_tile_cnt, d_tile_off;  // multi-pass pipeline scratch (grow-only)
    Buf d_agg_table, d_agg_out, d_agg_cnt;   // aggregation: global hash table, compacted groups, [groups, overflow] counters
    Buf h_agg_out;                           // pinned read-back of the compacted groups
    Buf d_work;                              // blocks_prune_kernel: work list of tiles for the filter kernel ([0] = count)
    Buf d_scan_part;                         // offset_scan_kernel: epoch-tagged chunk sums (zeroed when (re)allocated)
    Buf d_tile_list;                         // offset_scan_kernel (block pipeline): the non-empty tiles, for the emit kernel
    Buf d_grp_sum;                           // blocks_group_emit_kernel: match counts per group of 1024 blocks (zero between queries)
    uint32_t scan_epoch = 0;
    Buf d_trace;                             // IMM3_TRACE debugging buffer
    // Emit-kernel feedback: result density class (1 dense, 0 sparse) last seen for a query shape (table, filter
    // columns and kinds, select list).  A known class launches only the matching emit kernel; both kernels are correct
    // for any result, so a stale hint costs time, never rows.
    std::unordered_map<std::string, int> emit_hint;
    std::string explain_buf;
};

struct imm3_result {
    imm3_db* db = nullptr;
    int ncols = 0;
    std::vector<std::string> names;
    std::vector<int> types, widths;
    std::vector<Buf> d_cols, h_cols;
    int64_t local_count = 0;
    int64_t g_offset = 0, g_take = 0, g_total = 0;  // after the count exchange (single handle: 0, local_count, local_count)
    int world = 1;
    int64_t rank_counts[kMaxWorld] = {};
    int agg_first_col = -1;  // aggregate result: index of the first aggregate column (rows live in host buffers only)
    int64_t fetched = 0;
    int64_t pending = -1;        // rows of an imm3_result_fetch_async still in flight (-1 = none)
    cudaEvent_t copied = nullptr;  // recorded on the copy stream after the last device->host copy
    double device_ms = 0;            // < 0: not read yet (timing_seq says which query's events hold it)
    unsigned long long timing_seq = 0;
    double stage_ms[2] = {0, 0};
    double host_us[5] = {0, 0, 0, 0, 0};  // wall clock inside imm3_query_begin: plan, buffers + device plan, launches, wait for the GPU, epilogue
    int launches = 0;
    int64_t alg_bytes = 0;
};

namespace {

TableStore* find_table(imm3_db* db, const char* name) {
    if (!name) { set_error("table name is NULL"); return nullptr; }
    for (auto& t : db->tables)
        if (t.meta.name == name) return &t;
    // SegmentManager.getTable, SegmentManager.scala:89-92
    fail(IMM3_ERR_NOT_FOUND, "Table %s does not exist in SegmentManager", name);
    return nullptr;
}

int use_device(imm3_db* db) {
    if (db->host_only) return fail(IMM3_ERR_STATE, "handle was opened with IMM3_OPEN_HOST_ONLY: no device work (there is no CPU fallback)");
    CUDA_TRY(cudaSetDevice(db->device));
    return 0;
}

// ---- SegmentManager upload path: files -> pinned staging -> HBM --------------------------------
// The owned slice of every column is one contiguous byte range in canonical order (its segments' block bytes back to
// back).  It is cut into pieces of at most kStageBytes; a pool of I/O threads takes pieces off a shared counter, pread()s
// each piece from the segment files (page cache / tmpfs) STRAIGHT into one of its two pinned staging buffers - no mmap,
// no intermediate copy - and queues the host->device copy on its own stream, so file reads, PCIe copies of different
// threads and the refill of the other buffer all overlap.  (Round 1: one thread memcpy-ing mmap pages into a 2 x 32 MiB
// ring - 1.4 GB/s.)  The same piece reader fills the pinned host mirror that imm3_reupload re-stages from.
constexpr size_t kStageBytes = 4u << 20;
constexpr int kMaxStageThreads = 12;

struct Piece {
    ColumnStore* col;
    size_t off;    // byte offset inside the column's payload
    size_t bytes;
    size_t seg;    // first segment file touched, and the offset inside it
    size_t seg_off;
};

void cut_pieces(ColumnStore& col, std::vector<Piece>* out) {
    size_t seg = 0, seg_off = 0, off = 0;
    const size_t total = (size_t)col.encoded_bytes;
    while (off < total) {
        while (seg < col.segs.size() && seg_off >= (size_t)col.segs[seg].nbytes) { seg++; seg_off = 0; }
        const size_t take = std::min(kStageBytes, total - off);
        out->push_back(Piece{&col, off, take, seg, seg_off});
        size_t left = take;  // advance (seg, seg_off) by `take` bytes
        while (left) {
            const size_t in_seg = (size_t)col.segs[seg].nbytes - seg_off;
            if (left < in_seg) { seg_off += left; left = 0; }
            else { left -= in_seg; seg++; seg_off = 0; }
        }
        off += take;
    }
}

// pread the bytes of one piece into dst (host memory).
int read_piece(const Piece& pc, uint8_t* dst) {
    size_t seg = pc.seg, seg_off = pc.seg_off, done = 0;
    while (done < pc.bytes) {
        const SegmentFile& sf = pc.col->segs[seg];
        const size_t take = std::min(pc.bytes - done, (size_t)sf.nbytes - seg_off);
        if (take) {
            int fd = ::open(sf.path.c_str(), O_RDONLY);
            if (fd < 0) return fail(IMM3_ERR_IO, "open %s: %s", sf.path.c_str(), strerror(errno));
            size_t got = 0;
            while (got < take) {
                const ssize_t r = ::pread(fd, dst + done + got, take - got, (off_t)(seg_off + got));
                if (r <= 0) {
                    ::close(fd);
                    return fail(IMM3_ERR_IO, "read %s: %s", sf.path.c_str(), r < 0 ? strerror(errno) : "file is shorter than its block offsets");
                }
                got += (size_t)r;
            }
            ::close(fd);
        }
        done += take;
        seg++;
        seg_off = 0;
    }
    return 0;
}

// Run `pieces` on the I/O pool.  to_device: through pinned staging buffers into the columns' arenas; otherwise straight into
// the columns' pinned host mirrors.
int run_pieces(imm3_db* db, const std::vector<Piece>& pieces, bool to_device) {
    if (pieces.empty()) return 0;
    const int nthreads = std::max(1, std::min<int>(std::min(io_threads(), kMaxStageThreads), (int)pieces.size()));
    // Pinning memory is the slow part of a staged upload (tens of ms per 16 MiB, serialised inside the driver): ONE pinned
    // block for all staging threads, allocated on first use and kept for the life of the process.
    uint8_t* pool_base = nullptr;
    if (to_device) {
        static std::mutex pin_mu;
        static uint8_t* pin_block = nullptr;
        std::lock_guard<std::mutex> lock(pin_mu);
        if (!pin_block) CUDA_TRY(cudaHostAlloc(&pin_block, (size_t)kMaxStageThreads * 2 * kStageBytes, cudaHostAllocPortable));
        pool_base = pin_block;
    }
    static std::mutex run_mu;  // (one staged upload at a time per process: the pinned block is shared)
    std::unique_lock<std::mutex> run_lock(run_mu, std::defer_lock);
    if (to_device) run_lock.lock();
    std::atomic<size_t> next(0);
    std::atomic<int> first_rc(0);
    std::atomic<long long> us_alloc(0), us_read(0), us_wait(0);  // summed over the threads (IMM3_OPEN_TRACE)
    std::mutex mu;
    std::string why;
    auto report = [&](int rc) {
        std::lock_guard<std::mutex> lock(mu);
        if (!first_rc.load()) {
            why = last_error();
            first_rc.store(rc);
        }
    };
    auto work = [&](int tid) {
        uint8_t* stage[2] = {nullptr, nullptr};
        cudaEvent_t ev[2] = {nullptr, nullptr};
        cudaStream_t st = nullptr;
        auto body = [&]() -> int {
            if (to_device) {
                const double ta = now_us();
                CUDA_TRY(cudaSetDevice(db->device));
                CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
                for (int i = 0; i < 2; i++) {
                    stage[i] = pool_base + ((size_t)tid * 2 + (size_t)i) * kStageBytes;  // (carved out of the process-wide pinned block)
                    CUDA_TRY(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
                }
                us_alloc += (long long)(now_us() - ta);
            }
            int cur = 0;
            bool used[2] = {false, false};
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= pieces.size() || first_rc.load()) break;
                const Piece& pc = pieces[i];
                if (!to_device) {
                    int rc = read_piece(pc, pc.col->h_mirror + pc.off);
                    if (rc) return rc;
                    continue;
                }
                const double tw = now_us();
                if (used[cur]) CUDA_TRY(cudaEventSynchronize(ev[cur]));  // the buffer's previous copy has retired
                const double tr = now_us();
                int rc = read_piece(pc, stage[cur]);
                if (rc) return rc;
                us_wait += (long long)(tr - tw);
                us_read += (long long)(now_us() - tr);
                CUDA_TRY(cudaMemcpyAsync(pc.col->d_arena + pc.off, stage[cur], pc.bytes, cudaMemcpyHostToDevice, st));
                CUDA_TRY(cudaEventRecord(ev[cur], st));
                used[cur] = true;
                cur ^= 1;
            }
            if (st) CUDA_TRY(cudaStreamSynchronize(st));
            return 0;
        };
        const int rc = body();
        if (rc) report(rc);
        if (st) cudaStreamSynchronize(st);
        for (int i = 0; i < 2; i++)
            if (ev[i]) cudaEventDestroy(ev[i]);
        if (st) cudaStreamDestroy(st);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    if (int rc = first_rc.load()) return fail(rc, "%s", why.c_str());
    if (getenv("IMM3_OPEN_TRACE") && to_device)
        fprintf(stderr, "[imm3_open] %d staging threads; per thread: pinned buffers %.1f ms, file reads %.1f ms, waiting for copies %.1f ms\n", nthreads,
                us_alloc.load() * 1e-3 / nthreads, us_read.load() * 1e-3 / nthreads, us_wait.load() * 1e-3 / nthreads);
    return 0;
}

int upload_all(imm3_db* db) {
    const bool otrace = getenv("IMM3_OPEN_TRACE") != nullptr;
    const double t0 = now_us();
    std::vector<Piece> pieces;
    for (auto& t : db->tables) {
        for (auto& col : t.cols) {
            const bool dense = col.meta.codec != IMM3_CODEC_PFOR_INT;
            const size_t payload = (size_t)col.encoded_bytes;
            size_t arena = dense ? (size_t)((t.nrows + kDenseMaxTileRows - 1) / kDenseMaxTileRows) * kDenseMaxTileRows * (size_t)col.meta.width : payload;
            arena += 256;
            CUDA_TRY(cudaMalloc(&col.d_arena, arena));
            col.arena_bytes = arena;
            CUDA_TRY(cudaMemsetAsync(col.d_arena + payload, 0, arena - payload, db->stream));  // zero padding up to a whole tile
            if (!dense) {
                CUDA_TRY(cudaMalloc(&col.d_word_off, col.word_off.size() * sizeof(uint32_t) + 256));  // (+256: the filter kernels' TMA reads up to 36 entries per tile)
                CUDA_TRY(cudaMemcpyAsync(col.d_word_off, col.word_off.data(), col.word_off.size() * sizeof(uint32_t),
                                         cudaMemcpyHostToDevice, db->stream));
            }
            cut_pieces(col, &pieces);
        }
        CUDA_TRY(cudaMalloc(&t.d_row_start, t.row_start.size() * sizeof(uint64_t) + 512));  // (+512: ... and up to 34 row ordinals)
        CUDA_TRY(cudaMemcpyAsync(t.d_row_start, t.row_start.data(), t.row_start.size() * sizeof(uint64_t),
                                 cudaMemcpyHostToDevice, db->stream));
    }
    const double t1 = now_us();
    int rc = run_pieces(db, pieces, true);
    cudaError_t e = cudaStreamSynchronize(db->stream);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(IMM3_ERR_CUDA, "upload: %s", cudaGetErrorString(e));
    if (otrace) {
        size_t bytes = 0;
        for (auto& pc : pieces) bytes += pc.bytes;
        fprintf(stderr, "[imm3_open] arenas allocated in %.1f ms; %zu pieces, %.1f MB staged in %.1f ms = %.1f GB/s\n", (t1 - t0) * 1e-3, pieces.size(),
                bytes * 1e-6, (now_us() - t1) * 1e-3, bytes * 1e-3 / (now_us() - t1));
    }
    const double t2 = now_us();
    // Block statistics of the encoded INT columns (exact min / max per block, one decode pass on the GPU): what the stubs
    // SegmentStats / check() of the reference (Segment.scala:18-30) were meant to hold.  Range queries prune with them.
    if (!(db->flags & IMM3_OPEN_NO_STATS)) {
        for (auto& t : db->tables) {
            if (t.max_block_rows > 1024 || t.nblocks == 0) continue;  // (the warp-per-block decoder takes blocks of <= 1024 rows)
            for (auto& col : t.cols) {
                if (col.meta.codec != IMM3_CODEC_PFOR_INT) continue;
                CUDA_TRY(cudaMalloc(&col.d_stats, (size_t)t.nblocks * sizeof(BlockStat)));
                PforCol pc;
                pc.words = reinterpret_cast<const uint32_t*>(col.d_arena);
                pc.word_off = col.d_word_off;
                const int cap = (int)std::min<int64_t>(1120, ((col.max_block_words + 4 + 31) / 32) * 32);
                CUDA_TRY(launch_block_stats(pc, t.d_row_start, t.nblocks, cap, db->num_sms, (BlockStat*)col.d_stats, db->stream));
            }
        }
        CUDA_TRY(cudaStreamSynchronize(db->stream));
        if (otrace) fprintf(stderr, "[imm3_open] block statistics in %.1f ms\n", (now_us() - t2) * 1e-3);
    }
    return 0;
}

// Pinned host mirror of a column's payload, built on first use (imm3_reupload): the open path never pins a whole column.
int ensure_mirror(imm3_db* db, ColumnStore& col) {
    if (col.h_mirror || !col.encoded_bytes) return 0;
    CUDA_TRY(cudaMallocHost(&col.h_mirror, (size_t)col.encoded_bytes));
    std::vector<Piece> pieces;
    cut_pieces(col, &pieces);
    return run_pieces(db, pieces, false);
}

void free_device_side(imm3_db* db) {
    if (db->host_only) return;
    cudaSetDevice(db->device);
    if (db->stream) cudaStreamSynchronize(db->stream);
    for (auto& t : db->tables) {
        for (auto& c : t.cols) {
            if (c.d_arena) cudaFree(c.d_arena);
            if (c.d_word_off) cudaFree(c.d_word_off);
            if (c.d_stats) cudaFree(c.d_stats);
            if (c.h_mirror) cudaFreeHost(c.h_mirror);
        }
        if (t.d_row_start) cudaFree(t.d_row_start);
    }
    db->dev_pool.release_all();
    db->host_pool.release_all();
    if (db->d_bitmap.p) cudaFree(db->d_bitmap.p);
    if (db->d_span_cnt.p) cudaFree(db->d_span_cnt.p);
    if (db->d_tile_cnt.p) cudaFree(db->d_tile_cnt.p);
    if (db->d_tile_off.p) cudaFree(db->d_tile_off.p);
    if (db->d_trace.p) cudaFree(db->d_trace.p);
    if (db->d_scan_part.p) cudaFree(db->d_scan_part.p);
    if (db->d_tile_list.p) cudaFree(db->d_tile_list.p);
    if (db->d_grp_sum.p) cudaFree(db->d_grp_sum.p);
    if (db->d_work.p) cudaFree(db->d_work.p);
    if (db->d_agg_table.p) cudaFree(db->d_agg_table.p);
    if (db->d_agg_out.p) cudaFree(db->d_agg_out.p);
    if (db->d_agg_cnt.p) cudaFree(db->d_agg_cnt.p);
    if (db->h_agg_out.p) cudaFreeHost(db->h_agg_out.p);
    if (db->h_bitmap.p) cudaFreeHost(db->h_bitmap.p);
    if (db->d_status) cudaFree(db->d_status);
    if (db->d_ctrl) cudaFree(db->d_ctrl);
    if (db->h_ctrl) cudaFreeHost(db->h_ctrl);
    for (int i = 0; i < kMaxWorld; i++)
        if (db->comm_peer[i] && db->comm_peer[i] != db->d_mailbox) cudaIpcCloseMemHandle(db->comm_peer[i]);
    if (db->d_mailbox) cudaFree(db->d_mailbox);
    if (db->ev0) cudaEventDestroy(db->ev0);
    if (db->ev1) cudaEventDestroy(db->ev1);
    if (db->ev_mid) cudaEventDestroy(db->ev_mid);
    if (db->copy_stream) cudaStreamDestroy(db->copy_stream);
    if (db->own_stream) cudaStreamDestroy(db->own_stream);
    cudaGetLastError();
}

// ---- query planning on top of the logical plan --------------------------------------------------
struct Prepared {
    TableStore* table = nullptr;
    LogicalPlan lp;
    bool block_mode = false;
    bool multipass = false;   // dense tables: filter -> scan -> emit kernels instead of the fused single pass
    bool blocks_multi = false;  // block mode: warp-per-block filter kernel -> offset scan -> emit kernel (no look-back chain)
    bool for_bitmap = false;    // imm3_filter_bitmap: the canonical-row bitmap comes from the single-pass kernels
    bool quad = false;          // block mode: the single-range-predicate filter kernel (lane = block x super-block)
    bool lane = false;          // block mode: ... its lane-per-block successor (every warp its own TMA ring)
    int lane_warps = 8;         // ... warps per CTA of that kernel
    bool prune = false;         // block mode: every predicate is a range on an encoded column with block statistics: blocks_prune_kernel first
    PrunePlan pp;
    bool hybrid = false;        // block mode, no predicate on an encoded column: DENSE filter kernel (row space) -> block emit kernel
    size_t blocks_emit_smem = 0;
    int64_t prefix_blocks = 0;  // small LIMIT on a block table: the pipeline first runs over this many leading blocks (0 = no prefix)
    int64_t prefix_rows = 0;    // small LIMIT on a dense table: ... over this many leading rows (a multiple of the tile size)
    int grid_blocks_emit = 0;
    int grid_emit = 0;
    bool emit_general = false;  // select list needs the general gather kernel (> 4 columns or a width other than 1/2/4)
    int grid_emit_stream = 0, emit_stage_bytes = 0, emit_ring = 0;  // streaming emit kernel (dense results); 0 = not usable
    std::string shape_key;    // key of imm3_db::emit_hint
    size_t emit_smem = 0;
    ScanPlan sp;
    size_t dyn_smem = 0;
    int grid = 0;
    // real OR (imm3_query_begin_dnf): the second and later conjunctions of the disjunction - their filter kernels run behind this
    // plan's and OR their rows into its bitmap; counts, offsets and the emit kernels then see the union
    std::vector<std::unique_ptr<Prepared>> or_terms;
};

const char* kernel_name(const imm3_db* db, const TableStore& t, const LogicalPlan& lp, bool* block_mode) {
    bool blocks = lp.uses_pfor || (db->flags & IMM3_OPEN_FORCE_BLOCKS);  // tests: cross-check the two kernels
    (void)t;
    *block_mode = blocks;
    if (lp.always_empty) return "none(always_empty)";
    return blocks ? "blocks_filter -> blocks_emit" : ((db->flags & IMM3_OPEN_NO_TMA) ? "filter(direct) -> emit" : "filter(tma) -> emit");
}

// Dense tables run the three-kernel pipeline (filter -> offset scan -> emit: no cross-CTA dependency); a small LIMIT makes it
// run over a prefix of the table first (Prepared::prefix_rows).  (Round 1 also carried a fused single-pass kernel behind
// IMM3_PATH=fused - 0.7x the pipeline's rate, never the default, and the home of that round's one wrong result; deleted.)
bool choose_multipass(const LogicalPlan& lp, bool block_mode) {
    (void)lp;
    return !block_mode;
}

int prepare(imm3_db* db, const char* table, const imm3_pred* preds, int npreds, const char* const* proj, int nproj,
            int64_t limit, Prepared* pr) {
    pr->table = find_table(db, table);
    if (!pr->table) return IMM3_ERR_NOT_FOUND;
    int rc = build_logical_plan(pr->table->meta, preds, npreds, proj, nproj, limit, &pr->lp);
    if (rc) return rc;
    kernel_name(db, *pr->table, pr->lp, &pr->block_mode);
    pr->multipass = choose_multipass(pr->lp, pr->block_mode);
    pr->shape_key = pr->table->meta.name + "|";
    for (auto& f : pr->lp.filters) pr->shape_key += std::to_string(f.col_idx) + ":" + std::to_string(f.kind) + ",";
    pr->shape_key += "|";
    for (int c : pr->lp.proj) pr->shape_key += std::to_string(c) + ",";
    return 0;
}

// Fill the device plan (everything but result pointers and the bitmap).
int fill_scan_plan(imm3_db* db, Prepared* pr) {
    TableStore& t = *pr->table;
    const LogicalPlan& lp = pr->lp;
    ScanPlan& sp = pr->sp;
    std::memset(&sp, 0, sizeof sp);
    sp.nrows = t.nrows;
    sp.limit = lp.limit > 0 ? lp.limit : INT64_MAX;
    sp.row_start = t.d_row_start;
    sp.max_block_rows = t.max_block_rows;
    sp.nfilter = (int)lp.filters.size();
    sp.nproj = (int)lp.proj.size();
    std::vector<int> pfor_cols;
    auto pfor_slot = [&](int ci) -> int {
        if (t.cols[(size_t)ci].meta.codec != IMM3_CODEC_PFOR_INT) return -1;
        for (size_t i = 0; i < pfor_cols.size(); i++)
            if (pfor_cols[i] == ci) return (int)i;
        pfor_cols.push_back(ci);
        return (int)pfor_cols.size() - 1;
    };
    int lit_at = 0;
    for (int i = 0; i < sp.nfilter; i++) {
        const LogicalFilter& lf = lp.filters[(size_t)i];
        const ColumnStore& c = t.cols[(size_t)lf.col_idx];
        FilterCol& f = sp.filter[i];
        f.width = c.meta.width;
        f.kind = lf.kind;
        f.pfor_slot = pfor_slot(lf.col_idx);
        f.base = f.pfor_slot >= 0 ? nullptr : c.d_arena;
        f.smem_off = -1;
        if (lf.kind == kFilterStrMatch) {
            f.nlit = (int)lf.lits.size();
            f.lit_off = lit_at;
            for (auto& s : lf.lits) {
                std::memcpy(sp.lits + lit_at, s.data(), (size_t)f.width);
                lit_at += f.width;
            }
        } else {
            f.lo = (int32_t)lf.lo;
            f.span = (uint32_t)(lf.hi - lf.lo);
        }
    }
    sp.lit_bytes = lit_at;
    for (int i = 0; i < sp.nproj; i++) {
        const int ci = lp.proj[(size_t)i];
        const ColumnStore& c = t.cols[(size_t)ci];
        ProjCol& p = sp.proj[i];
        p.width = c.meta.width;
        p.pfor_slot = pfor_slot(ci);
        p.base = p.pfor_slot >= 0 ? nullptr : c.d_arena;
        p.filter_idx = -1;
        for (int k = 0; k < sp.nfilter; k++)
            if (lp.filters[(size_t)k].col_idx == ci) p.filter_idx = k;
    }
    // A filter column that is also projected is read again by the emit kernel: ask L2 to keep it if all such
    // columns together fit in ~2/3 of the 126 MB L2.
    {
        int64_t again = 0;
        for (int i = 0; i < sp.nproj; i++)
            if (sp.proj[i].filter_idx >= 0) again += t.nrows * sp.filter[sp.proj[i].filter_idx].width;
        for (int i = 0; i < sp.nproj; i++)
            if (sp.proj[i].filter_idx >= 0) sp.filter[sp.proj[i].filter_idx].keep_l2 = (again > 0 && again <= (int64_t)(getenv("IMM3_KEEP_L2_MB") ? atoi(getenv("IMM3_KEEP_L2_MB")) : 84) << 20) ? 1 : 0;
    }
    sp.npfor = (int)pfor_cols.size();
    for (int i = 0; i < sp.npfor; i++) {
        const ColumnStore& c = t.cols[(size_t)pfor_cols[(size_t)i]];
        sp.pfor[i].words = reinterpret_cast<const uint32_t*>(c.d_arena);
        sp.pfor[i].word_off = c.d_word_off;
    }
    if (++db->epoch >= 0x3FFFFFu) {  // 22-bit tag wrapped: clear the status words once
        if (db->d_status) CUDA_TRY(cudaMemsetAsync(db->d_status, 0, db->status_cap * sizeof(unsigned long long), db->stream));
        db->epoch = 1;
    }
    sp.epoch = db->epoch;
    if (const char* e = getenv("IMM3_DEBUG")) sp.debug = (uint32_t)atoi(e);

    int occ = 0;
    // one 8192-row sub-tile of every (dense) filter column, and where each column sits inside a TMA stage
    auto dense_sub_bytes = [&]() {
        int row_bytes = 0;
        for (int i = 0; i < sp.nfilter; i++) row_bytes += sp.filter[i].width;
        return kDenseTileRowsPerWord * row_bytes;
    };
    auto dense_stage_offsets = [&](int sub_bytes) {
        const bool can_stage = !(db->flags & IMM3_OPEN_NO_TMA) && sub_bytes > 0 && sub_bytes <= 56 * 1024;
        int off = 0;
        for (int i = 0; i < sp.nfilter; i++) {
            sp.filter[i].smem_off = can_stage ? off : -1;
            off += kDenseTileRowsPerWord * sp.filter[i].width;
        }
        return can_stage;
    };
    // Multi-pass filter kernel (K1): tile = 8192 rows; a TMA ring of ~32 KiB, so that four CTAs share an SM.
    auto config_dense_filter = [&]() -> cudaError_t {
        const int stage_bytes = dense_sub_bytes();
        const bool stage_ok = dense_stage_offsets(stage_bytes);
        sp.words_per_lane = 1;
        sp.ntiles = (t.nrows + kDenseTileRowsPerWord - 1) / kDenseTileRowsPerWord;
        int stages = 0;
        if (stage_ok) stages = std::max(2, std::min(kMaxFilterStages, (32 * 1024) / stage_bytes));
        if (const char* e = getenv("IMM3_FILTER_STAGES")) {
            int v = atoi(e);
            if (stage_ok && v >= 2 && v <= kMaxFilterStages && (size_t)v * stage_bytes <= 200 * 1024) stages = v;
        }
        sp.stages = stages;
        sp.stage_bytes = stage_bytes;
        pr->dyn_smem = (size_t)stages * (size_t)stage_bytes;
        return filter_kernel_occupancy(pr->dyn_smem, &occ);
    };
    if (pr->block_mode) {
        if (t.max_block_rows > kMaxBlockRows)
            return fail(IMM3_ERR_UNSUPPORTED, "block-mode kernel stages blocks of at most %d rows, table %s has a block of %d",
                        kMaxBlockRows, t.meta.name.c_str(), t.max_block_rows);
        const char* path = getenv("IMM3_PATH");
        pr->blocks_multi = !pr->for_bitmap && t.max_block_rows <= 1024 && !(path && !strcmp(path, "fused"));
        // Small LIMIT: the multi-pass pipeline has no early exit, so it first runs over a prefix of the table (64 rows per
        // requested row, at least 4 M rows); only if that does not fill the LIMIT is the whole table scanned.  (The
        // single-pass kernel does stop early, but its look-back chain costs 13 ns per block when the rows come late.)
        if (pr->blocks_multi && lp.limit > 0 && lp.limit <= (1 << 20) && !getenv("IMM3_NO_PREFIX")) {
            const int64_t min_rows = getenv("IMM3_PREFIX_ROWS") ? std::max(1024, atoi(getenv("IMM3_PREFIX_ROWS"))) : (4 << 20);  // (tests shrink it)
            const int64_t want_rows = std::max<int64_t>(min_rows, lp.limit * 64);
            const int64_t nb = (want_rows + 1023) / 1024;
            if (nb * 2 <= t.nblocks) pr->prefix_blocks = nb;
        }
        if (pr->blocks_multi) {
            bool filters_dense = true;
            for (int i = 0; i < sp.nfilter; i++) filters_dense = filters_dense && sp.filter[i].pfor_slot < 0;
            pr->hybrid = filters_dense && !getenv("IMM3_NO_HYBRID");
            sp.ntiles = (t.nblocks + 7) / 8;  // the offset scan works on tiles of 8 blocks
            int64_t cap = 0;  // per-warp scratch for the byte-swapped words of one encoded block
            for (int ci : pfor_cols) cap = std::max<int64_t>(cap, t.cols[(size_t)ci].max_block_words);
            sp.blk_words_cap = (int)std::min<int64_t>(1120, ((cap + 4 + 31) / 32) * 32);
            // Filter kernel: a TMA ring of raw tiles (8 blocks of every encoded column that carries a predicate + metadata).
            int64_t tile_cap = 0;
            int nstaged = 0;
            sp.pfor_filter_mask = 0;
            for (int i = 0; i < sp.nfilter; i++)
                if (sp.filter[i].pfor_slot >= 0) sp.pfor_filter_mask |= 1u << sp.filter[i].pfor_slot;
            for (int s = 0; s < sp.npfor; s++)
                if ((sp.pfor_filter_mask >> s) & 1u) {
                    tile_cap = std::max<int64_t>(tile_cap, t.cols[(size_t)pfor_cols[(size_t)s]].max_tile_bytes);
                    nstaged++;
                }
            sp.blk_tile_bytes = (int)((tile_cap + 16 + 15) & ~15ll);
            const int slot_bytes = blocks_filter_slot_bytes(nstaged, sp.blk_tile_bytes);
            int ring = std::max(2, std::min(kMaxFilterStages, (48 * 1024) / slot_bytes));
            if (const char* e = getenv("IMM3_BLOCKS_STAGES")) ring = std::max(2, std::min(kMaxFilterStages, atoi(e)));
            while (ring > 2 && (size_t)ring * slot_bytes > 200 * 1024) ring--;
            sp.stages = ring;  // (the row-space variant below re-plans these two for the dense filter kernel)
            sp.stage_bytes = slot_bytes;
            pr->dyn_smem = blocks_filter_smem_bytes(nstaged, sp.blk_tile_bytes, ring);
            // The only predicate is a range on one encoded column (C4): the quad kernel - lane = (block, super-block), CTA tile
            // of 32 blocks - if such a tile of this column fits a ring slot (i.e. the column actually compresses).
            pr->quad = false;
            pr->lane = false;
            if (sp.nfilter == 1 && sp.filter[0].pfor_slot >= 0 && sp.filter[0].kind == kFilterI32Range && !getenv("IMM3_NO_QUAD")) {
                const int64_t cap32 = (t.cols[(size_t)pfor_cols[(size_t)sp.filter[0].pfor_slot]].max_tile32_bytes + 16 + 15) & ~15ll;
                const int qslot = blocks_filter_quad_slot_bytes((int)cap32);
                if (qslot <= 40 * 1024) {
                    pr->quad = true;
                    int qring = std::max(2, std::min(4, (40 * 1024) / qslot));
                    if (const char* e = getenv("IMM3_BLOCKS_STAGES")) qring = std::max(2, std::min(kMaxFilterStages, atoi(e)));
                    sp.blk_tile_bytes = (int)cap32;
                    sp.stages = qring;
                    sp.stage_bytes = qslot;
                    pr->dyn_smem = (size_t)qring * (size_t)qslot;
                }
                // lane = block: every warp of the CTA runs its own ring of 2 .. 4 such slots; one CTA per SM with as many warps
                // (8 .. 16) as its shared memory holds two-slot rings for
                const size_t lane_smem_cap = 222 * 1024;
                if (pr->quad && 2 * 8 * (size_t)qslot <= lane_smem_cap && !getenv("IMM3_NO_LANE")) {
                    pr->lane = true;
                    int lring = 2;
                    if (const char* e = getenv("IMM3_LANE_STAGES")) lring = std::max(2, std::min(4, atoi(e)));
                    while (lring > 2 && (size_t)lring * 4 * (size_t)qslot > lane_smem_cap) lring--;
                    int lw = (int)std::min<size_t>(16, lane_smem_cap / ((size_t)lring * (size_t)qslot));
                    if (const char* e = getenv("IMM3_LANE_WARPS")) lw = std::max(4, std::min(lw, atoi(e)));
                    pr->lane_warps = lw;
                    sp.stages = lring;
                    pr->dyn_smem = (size_t)lring * (size_t)lw * (size_t)qslot;
                }
            }
            // Pruning: every predicate is a range on an encoded column that has block statistics.
            pr->prune = false;
            if (sp.nfilter > 0 && !pr->hybrid && !getenv("IMM3_NO_PRUNE")) {
                bool ok = true;
                std::memset(&pr->pp, 0, sizeof pr->pp);
                for (int i = 0; i < sp.nfilter && ok; i++) {
                    const FilterCol& f = sp.filter[i];
                    ok = f.pfor_slot >= 0 && f.kind == kFilterI32Range && t.cols[(size_t)pfor_cols[(size_t)f.pfor_slot]].d_stats != nullptr;
                    if (ok) {
                        pr->pp.stats[i] = (const BlockStat*)t.cols[(size_t)pfor_cols[(size_t)f.pfor_slot]].d_stats;
                        pr->pp.lo[i] = f.lo;
                        pr->pp.hi[i] = (int32_t)((int64_t)f.lo + (int64_t)f.span);
                    }
                }
                pr->pp.nfilter = sp.nfilter;
                pr->pp.group_shift = pr->quad ? 5 : 3;
                pr->prune = ok;
                // a pruned scan decides whole blocks from 8 bytes each: the whole slice costs what a prefix would - no prefix phase (with a
                // communicator phase A then is this rank's whole slice and phase B the count exchange alone: the protocol is unchanged)
                if (ok) pr->prefix_blocks = 0;
            }
            pr->blocks_emit_smem = blocks_emit_smem_bytes(sp.npfor, sp.blk_words_cap);
            int occ_e = 0;
            CUDA_TRY(blocks_multi_occupancy(pr->dyn_smem, pr->blocks_emit_smem, pr->hybrid ? nullptr : &occ, &occ_e, pr->lane ? (2 | (pr->lane_warps << 8)) : (pr->quad ? 1 : 0), pr->hybrid));
            if (occ_e < 1) return fail(IMM3_ERR_CUDA, "block emit kernel does not fit on an SM (dynamic shared memory %zu bytes)", pr->blocks_emit_smem);
            pr->grid_blocks_emit = (int)std::max<int64_t>(1, std::min<int64_t>((t.nblocks + 7) / 8, (int64_t)db->num_sms * std::max(1, occ_e)));
            if (pr->hybrid) {
                // No predicate touches an encoded column: the dense filter kernel runs over the table's row space (dense
                // columns are contiguous in HBM whatever the block framing) and only the emit kernel works per block.
                CUDA_TRY(config_dense_filter());
            }
        } else {
            sp.ntiles = t.nblocks;
            pr->dyn_smem = blocks_kernel_smem_bytes(sp.npfor, t.max_block_rows);
            CUDA_TRY(blocks_kernel_occupancy(pr->dyn_smem, &occ));
        }
    } else {
        if (pr->multipass) {
            if (lp.limit > 0 && lp.limit <= (1 << 20) && !getenv("IMM3_NO_PREFIX")) {  // (same policy as the block tables above)
                const int64_t min_rows = getenv("IMM3_PREFIX_ROWS") ? std::max(1024, atoi(getenv("IMM3_PREFIX_ROWS"))) : (4 << 20);
                const int64_t want = (std::max<int64_t>(min_rows, lp.limit * 64) + kDenseTileRowsPerWord - 1) / kDenseTileRowsPerWord * kDenseTileRowsPerWord;
                if (want * 2 <= t.nrows) pr->prefix_rows = want;
            }
            CUDA_TRY(config_dense_filter());
            const int W = 1;
            const int tile_rows = kDenseTileRowsPerWord;
            int occ_emit = 0;
            pr->emit_general = sp.nproj > 4 || getenv("IMM3_EMIT_GENERAL");
            for (int i = 0; i < sp.nproj; i++) pr->emit_general = pr->emit_general || !(sp.proj[i].width == 1 || sp.proj[i].width == 2 || sp.proj[i].width == 4);
            CUDA_TRY(emit_kernel_occupancy(pr->emit_general, &occ_emit));
            const int64_t nspans = sp.ntiles * (tile_rows / 1024);
            pr->grid_emit = (int)std::max<int64_t>(1, std::min<int64_t>((nspans + 7) / 8, (int64_t)db->num_sms * std::max(1, occ_emit)));
            // Streaming emit kernel (dense results): a stage = bitmap words + span counts + one 8192-row tile of every
            // projected column, if that fits.
            int pstage = 0;
            for (int i = 0; i < sp.nproj; i++) {
                sp.proj[i].stage_off = pstage;
                pstage += 1024 * sp.proj[i].width;
            }
            pr->emit_stage_bytes = 0;
            if (sp.nproj > 0 && 8 * pstage <= 64 * 1024 && !(db->flags & IMM3_OPEN_NO_TMA) && !getenv("IMM3_NO_EMIT_STREAM")) {
                pr->emit_stage_bytes = emit_stream_header_bytes() + 8 * pstage;
                pr->emit_ring = std::max(2, std::min(4, (80 * 1024) / pr->emit_stage_bytes));
                if (const char* e = getenv("IMM3_EMIT_STAGES")) {
                    int v = atoi(e);
                    if (v >= 2 && v <= 4) pr->emit_ring = v;
                }
                pr->emit_smem = emit_stream_smem_bytes(pr->emit_stage_bytes, pr->emit_ring);
                int occ_s = 0;
                CUDA_TRY(emit_stream_occupancy(pr->emit_smem, &occ_s));
                if (occ_s < 1) pr->emit_stage_bytes = 0;
                pr->grid_emit_stream = (int)std::max<int64_t>(1, std::min<int64_t>(sp.ntiles * W, (int64_t)db->num_sms * std::max(1, occ_s)));
            }
        }
    }
    if (sp.ntiles >= (int64_t)0x7FFFFFFF) return fail(IMM3_ERR_UNSUPPORTED, "too many tiles (%lld)", (long long)sp.ntiles);
    if (occ < 1) return fail(IMM3_ERR_CUDA, "kernel does not fit on an SM (dynamic shared memory %zu bytes)", pr->dyn_smem);
    const int64_t persistent = (int64_t)db->num_sms * occ;
    pr->grid = (int)std::max<int64_t>(1, std::min<int64_t>(sp.ntiles, persistent));
    const size_t status_need = 2 * (size_t)status_round_up(sp.ntiles);  // tile counts + (dense kernel) tile offsets
    if (status_need > db->status_cap) {
        if (db->d_status) CUDA_TRY(cudaFree(db->d_status));
        db->d_status = nullptr;
        size_t cap = status_need + status_need / 4 + 1024;
        CUDA_TRY(cudaMalloc(&db->d_status, cap * sizeof(unsigned long long)));
        CUDA_TRY(cudaMemsetAsync(db->d_status, 0, cap * sizeof(unsigned long long), db->stream));
        db->status_cap = cap;
    }
    return 0;
}

int ensure_buf(Buf* b, size_t bytes) {
    if (b->cap >= bytes) return 0;
    if (b->p) cudaFree(b->p);
    *b = Buf();
    size_t cap = bytes + bytes / 8 + 4096;
    CUDA_TRY(cudaMalloc(&b->p, cap));
    b->cap = cap;
    return 0;
}

// Launch the kernels of one query and wait for the match count.
// Launch the count-exchange kernel of one round (every rank of the communicator must launch the same rounds in the same
// order).  has_count = 0: this rank ran no kernel in this query (empty slice) and contributes 0.
int launch_exchange(imm3_db* db, int64_t limit, int has_count, unsigned long long pub_seq = 0) {
    CommPlan cp;
    std::memset(&cp, 0, sizeof cp);
    cp.pub = pub_seq ? db->h_ctrl : nullptr;
    cp.pub_seq = pub_seq;
    for (int i = 0; i < db->world; i++) cp.peer[i] = db->comm_peer[i];
    cp.rank = db->rank;
    cp.world = db->world;
    if (((++db->comm_epoch) & 0xFFFFFFu) == 0) ++db->comm_epoch;  // tag 0 is what an untouched mailbox holds
    cp.epoch = db->comm_epoch;
    cp.has_count = has_count;
    cp.limit = limit;
    cp.timeout_ns = db->comm_timeout_ns;
    CUDA_TRY(launch_count_exchange(cp, db->d_ctrl, &reinterpret_cast<CtrlBlock*>(db->d_ctrl)->x, db->stream));
    return 0;
}

// A round of the exchange without a scan of this rank's own: its slice is empty (has_count = 0) or its previous phase
// already covered the whole slice (has_count = 1: ctrl->total still holds that count).
int exchange_only(imm3_db* db, int64_t limit, int has_count) {
    int rc = launch_exchange(db, limit, has_count);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(db->h_ctrl, db->d_ctrl, sizeof(CtrlBlock), cudaMemcpyDeviceToHost, db->stream));
    CUDA_TRY(cudaStreamSynchronize(db->stream));
    if (db->h_ctrl->x.error)
        return fail(IMM3_ERR_COMM, "count exchange: a peer's count did not arrive within %llu ms (ranks must issue the same queries in the same order)",
                    db->comm_timeout_ns / 1000000ull);
    return 0;
}

// Who turns the tile counts into offsets: the filter kernel's last CTA (small tables), or offset_scan_kernel - one CTA per
// 4096 counts, launched behind the filter kernel (large tables: 122 K counts at 1 B rows take one SM ~50 us).
bool scan_inline_for(int64_t ntiles) {
    if (const char* e = getenv("IMM3_SCAN")) {
        if (!strcmp(e, "inline")) return true;
        if (!strcmp(e, "kernel")) return false;
    }
    return ntiles <= scan_inline_max_tiles();
}
int prepare_scan_buffers(imm3_db* db, int64_t ntiles, bool want_list) {
    if (want_list) {
        int rc = ensure_buf(&db->d_tile_list, (size_t)(ntiles + 16) * 4);
        if (rc) return rc;
    }
    const size_t nchunks = (size_t)((ntiles + 4095) / 4096);
    if (db->d_scan_part.cap < nchunks * 16 + 512) {  // (+ the scan-emit kernel's chunks-done counter, 128 bytes behind the sums)
        if (db->d_scan_part.p) cudaFree(db->d_scan_part.p);
        db->d_scan_part = Buf();
        const size_t cap = nchunks * 16 + 4096;
        CUDA_TRY(cudaMalloc(&db->d_scan_part.p, cap));
        CUDA_TRY(cudaMemsetAsync(db->d_scan_part.p, 0, cap, db->stream));
        db->d_scan_part.cap = cap;
    }
    if (((++db->scan_epoch) & 0xFFFFFFu) == 0) ++db->scan_epoch;  // (tag 0 = never written)
    return 0;
}
int launch_scan_kernel(imm3_db* db, const ScanPlan& sp, int64_t ntiles, int* launches, bool want_list = false) {
    {
        int rc = prepare_scan_buffers(db, ntiles, want_list);
        if (rc) return rc;
    }
    const size_t nchunks = (size_t)((ntiles + 4095) / 4096);
    (void)nchunks;
    CUDA_TRY(launch_offset_scan((const uint32_t*)db->d_tile_cnt.p, (unsigned long long*)db->d_tile_off.p, ntiles, sp.limit, db->scan_epoch,
                                (unsigned long long*)db->d_scan_part.p, db->d_ctrl, want_list ? (unsigned int*)db->d_tile_list.p : nullptr, db->stream));
    (*launches)++;
    return 0;
}

// Real OR: the filter kernels of the later terms, behind the first term's (same bitmap, same span / tile counts: every pass
// rewrites the counts of the union so far; only the last pass may run the inline offset scan).
int launch_or_terms(imm3_db* db, Prepared* pr, int scan_inline, int* launches) {
    for (size_t i = 0; i < pr->or_terms.size(); i++) {
        Prepared* tm = pr->or_terms[i].get();
        tm->sp.bitmap = pr->sp.bitmap;
        tm->sp.or_accumulate = 1;
        tm->sp.nrows = pr->sp.nrows;
        tm->sp.ntiles = pr->sp.ntiles;
        tm->sp.limit = pr->sp.limit;
        tm->sp.debug = pr->sp.debug;
        tm->sp.scan_inline = (i + 1 == pr->or_terms.size()) ? scan_inline : 0;
        CUDA_TRY(launch_filter(tm->sp, tm->sp.bitmap, (uint32_t*)db->d_span_cnt.p, (uint32_t*)db->d_tile_cnt.p, (unsigned long long*)db->d_tile_off.p,
                               db->d_ctrl, (int)std::max<int64_t>(1, std::min<int64_t>(tm->grid, tm->sp.ntiles)), tm->dyn_smem, db->stream));
        (*launches)++;
    }
    return 0;
}

int run_scan_once(imm3_db* db, Prepared* pr, double* ms, int64_t* total, int* launches, double* stage_ms, int64_t nblocks_use, bool exchange) {
    bool have_mid = false;
    pr->sp.scan_inline = 1;
    *launches = 0;
    // publish (plan.hpp): the query's last kernel writes the pinned control block itself and the host polls it
    const bool publish_ok = !getenv("IMM3_NO_PUBLISH");
    const unsigned long long pub_seq = ++db->pub_seq;
    bool published = false;
    if (pr->block_mode && pr->hybrid) {
        // dense filter kernel over the row space -> block emit kernel (decodes only blocks with surviving rows)
        TableStore& t = *pr->table;
        const int64_t ntiles = pr->sp.ntiles, nspans = ntiles * 8;
        int rc;
        // (+64 words: the row-space emit kernel reads, for every lane of a block's warp, the word holding bit R0 + 32*lane
        // and the one after it - up to 33 words past the last row's word for a 1-row tail block at the end of the slice)
        if ((rc = ensure_buf(&db->d_bitmap, (size_t)(ntiles * kDenseTileRowsPerWord / 32 + 64) * 4))) return rc;
        pr->sp.bitmap = (uint32_t*)db->d_bitmap.p;
        if ((rc = ensure_buf(&db->d_span_cnt, (size_t)(nspans + 8) * 4))) return rc;
        const size_t ntiles_pad = ((size_t)ntiles + 4095) / 4096 * 4096 + 16;  // whole rounds of the offset scan
        if ((rc = ensure_buf(&db->d_tile_cnt, ntiles_pad * 4))) return rc;
        if ((rc = ensure_buf(&db->d_tile_off, ntiles_pad * 8))) return rc;
        const int scan_inline_h = (pr->sp.nfilter == 0 || scan_inline_for(ntiles)) ? 1 : 0;
        pr->sp.scan_inline = pr->or_terms.empty() ? scan_inline_h : 0;
        CUDA_TRY(cudaEventRecord(db->ev0, db->stream));
        CUDA_TRY(launch_filter(pr->sp, pr->sp.bitmap, (uint32_t*)db->d_span_cnt.p, (uint32_t*)db->d_tile_cnt.p,
                               (unsigned long long*)db->d_tile_off.p, db->d_ctrl, pr->grid, pr->dyn_smem, db->stream));
        *launches = 1;
        if ((rc = launch_or_terms(db, pr, scan_inline_h, launches))) return rc;
        pr->sp.scan_inline = scan_inline_h;
        if (!pr->sp.scan_inline && (rc = launch_scan_kernel(db, pr->sp, ntiles, launches))) return rc;
        const bool pdl = pr->sp.nproj > 0 && !getenv("IMM3_NO_PDL");  // (no event may sit between a kernel and its programmatic dependent)
        if (!pdl) {
            CUDA_TRY(cudaEventRecord(db->ev_mid, db->stream));
            have_mid = true;
        }
        if (pr->sp.nproj > 0) {
            CUDA_TRY(launch_blocks_emit(pr->sp, pr->sp.bitmap, (const uint32_t*)db->d_span_cnt.p, nullptr, (const unsigned long long*)db->d_tile_off.p,
                                        nblocks_use, db->d_ctrl, true, pdl, (int)std::min<int64_t>(pr->grid_blocks_emit, (nblocks_use + 7) / 8),
                                        pr->blocks_emit_smem, db->stream));
            (*launches)++;
        }
        CUDA_TRY(cudaEventRecord(db->ev1, db->stream));
    } else if (pr->block_mode && pr->blocks_multi) {
        TableStore& t = *pr->table;
        const int64_t nblocks = nblocks_use, ntiles = pr->sp.ntiles;
        int rc;
        if ((rc = ensure_buf(&db->d_bitmap, (size_t)(nblocks * 32 + 2) * 4))) return rc;
        if ((rc = ensure_buf(&db->d_span_cnt, (size_t)(nblocks + 8 + 1024) * 4))) return rc;  // (+ whole 16-byte loads of a group's counts)
        const size_t ntiles_pad = ((size_t)ntiles + 4095) / 4096 * 4096 + 16;  // whole rounds of the offset scan
        if ((rc = ensure_buf(&db->d_tile_cnt, ntiles_pad * 4))) return rc;
        if ((rc = ensure_buf(&db->d_tile_off, ntiles_pad * 8))) return rc;
        pr->sp.scan_inline = (scan_inline_for(ntiles) && !(pr->lane && pr->lane_warps < 8)) ? 1 : 0;  // (the inline scan is written for 8 warps)
        const bool publish_here = publish_ok && !(exchange && db->comm_on);  // (a sharded table: the exchange kernel publishes)
        int fused_grid = 0;  // > 0: offset scan + emit as ONE kernel behind the filter kernel (one encoded column projected)
        int group_grid = 0, ngroups = 0;  // > 0: no offset scan at all - blocks_group_emit_kernel behind the lane kernel
        if (pr->sp.nproj == 1 && pr->sp.proj[0].pfor_slot >= 0 && !getenv("IMM3_NO_SCANEMIT")) {
            if (pr->lane && !getenv("IMM3_NO_GROUPEMIT")) CUDA_TRY(blocks_group_emit_grid(db->num_sms, nblocks, &group_grid, &ngroups));
            if (group_grid > 0) {
                if (!db->d_grp_sum.p) {
                    CUDA_TRY(cudaMalloc(&db->d_grp_sum.p, blocks_group_sum_bytes()));
                    db->d_grp_sum.cap = blocks_group_sum_bytes();
                    CUDA_TRY(cudaMemsetAsync(db->d_grp_sum.p, 0, db->d_grp_sum.cap, db->stream));
                }
                pr->sp.scan_inline = 0;
            } else {
                CUDA_TRY(blocks_scan_emit_grid(db->num_sms, ntiles, &fused_grid));
                if (fused_grid > 0) {
                    if ((rc = prepare_scan_buffers(db, ntiles, true))) return rc;
                    pr->sp.scan_inline = 0;
                }
            }
        }
        uint32_t* const grp_sum = group_grid > 0 ? (uint32_t*)db->d_grp_sum.p : nullptr;
        const unsigned int* work = nullptr;
        if (pr->prune) {
            if ((rc = ensure_buf(&db->d_work, (size_t)((nblocks + 7) / 8 + 2) * 4))) return rc;
            work = (const unsigned int*)db->d_work.p;
        }
        if ((pr->sp.debug & 16u) && getenv("IMM3_TRACE")) {  // debugging: phase stamps (min / max over warps)
            int rc0 = ensure_buf(&db->d_trace, (64 + 1024) * 8);
            if (rc0) return rc0;
            std::vector<unsigned long long> init(64 + 1024);
            for (int i = 0; i < 64; i++) init[(size_t)i] = (i & 1) ? 0ull : ~0ull;
            CUDA_TRY(cudaMemcpyAsync(db->d_trace.p, init.data(), init.size() * 8, cudaMemcpyHostToDevice, db->stream));
            CUDA_TRY(cudaStreamSynchronize(db->stream));
            pr->sp.trace = (unsigned long long*)db->d_trace.p;
        }
        CUDA_TRY(cudaEventRecord(db->ev0, db->stream));
        if (pr->prune) {
            // whole blocks decided from their min / max; only the tiles a window edge cuts through reach the filter kernel
            CUDA_TRY(cudaMemsetAsync(db->d_work.p, 0, 4, db->stream));
            CUDA_TRY(launch_blocks_prune(pr->pp, t.d_row_start, nblocks, ntiles, (uint32_t*)db->d_span_cnt.p, (uint32_t*)db->d_tile_cnt.p,
                                         (unsigned int*)db->d_work.p, db->num_sms, db->stream, grp_sum));
            (*launches)++;
        }
        CUDA_TRY(launch_blocks_filter(pr->sp, (uint32_t*)db->d_bitmap.p, (uint32_t*)db->d_span_cnt.p, (uint32_t*)db->d_tile_cnt.p,
                                      (unsigned long long*)db->d_tile_off.p, db->d_ctrl, nblocks,
                                      pr->lane ? (int)std::max<int64_t>(1, std::min<int64_t>(pr->grid, (nblocks + 32 * pr->lane_warps - 1) / (32 * pr->lane_warps)))
                                               : (pr->quad ? (int)std::max<int64_t>(1, std::min<int64_t>(pr->grid, (nblocks + 31) / 32)) : pr->grid),
                                      pr->dyn_smem, pr->lane ? (2 | (pr->lane_warps << 8)) : (pr->quad ? 1 : 0), work, db->stream, grp_sum));
        (*launches)++;
        if (group_grid > 0) {
            // one encoded column projected, lane kernel in front: every emit warp finds its rows from the group sums - no scan
            const bool pdl = !getenv("IMM3_NO_PDL");
            if (!pdl) {
                CUDA_TRY(cudaEventRecord(db->ev_mid, db->stream));
                have_mid = true;
            }
            CUDA_TRY(launch_blocks_group_emit(pr->sp, (const uint32_t*)db->d_bitmap.p, (const uint32_t*)db->d_span_cnt.p, grp_sum, nblocks, ngroups,
                                              db->d_ctrl, pdl, group_grid, db->stream, publish_here ? db->h_ctrl : nullptr, publish_here ? pub_seq : 0));
            published = publish_here;
            (*launches)++;
            CUDA_TRY(cudaEventRecord(db->ev1, db->stream));
        } else if (fused_grid > 0) {
            // one encoded column projected: offset scan + emit in one small-footprint kernel, resident while the filter kernel runs
            const bool pdl = !getenv("IMM3_NO_PDL");
            if (!pdl) {
                CUDA_TRY(cudaEventRecord(db->ev_mid, db->stream));
                have_mid = true;
            }
            CUDA_TRY(launch_blocks_scan_emit(pr->sp, (const uint32_t*)db->d_bitmap.p, (const uint32_t*)db->d_span_cnt.p, (const uint32_t*)db->d_tile_cnt.p,
                                             (unsigned long long*)db->d_tile_off.p, nblocks, db->scan_epoch, (unsigned long long*)db->d_scan_part.p,
                                             db->d_ctrl, (unsigned int*)db->d_tile_list.p, pdl, fused_grid, db->stream,
                                             publish_here ? db->h_ctrl : nullptr, publish_here ? pub_seq : 0));
            published = publish_here;
            (*launches)++;
            CUDA_TRY(cudaEventRecord(db->ev1, db->stream));
        } else {
        if (!pr->sp.scan_inline && (rc = launch_scan_kernel(db, pr->sp, ntiles, launches, true))) return rc;
        const bool pdl = pr->sp.nproj > 0 && !getenv("IMM3_NO_PDL");
        if (!pdl) {
            CUDA_TRY(cudaEventRecord(db->ev_mid, db->stream));
            have_mid = true;
        }
        if (pr->sp.nproj > 0) {
            CUDA_TRY(launch_blocks_emit(pr->sp, (const uint32_t*)db->d_bitmap.p, (const uint32_t*)db->d_span_cnt.p,
                                        pr->sp.scan_inline ? nullptr : (const unsigned int*)db->d_tile_list.p,
                                        (const unsigned long long*)db->d_tile_off.p, nblocks, db->d_ctrl, false, pdl,
                                        (int)std::min<int64_t>(pr->grid_blocks_emit, (nblocks + 7) / 8), pr->blocks_emit_smem, db->stream));
            (*launches)++;
        }
        CUDA_TRY(cudaEventRecord(db->ev1, db->stream));
        }
    } else if (pr->multipass) {
        if ((pr->sp.debug & 16u) && getenv("IMM3_TRACE")) {  // debugging: phase stamps (min / max over CTAs)
            int rc0 = ensure_buf(&db->d_trace, 64 * 8);
            if (rc0) return rc0;
            std::vector<unsigned long long> init(64);
            for (int i = 0; i < 64; i++) init[(size_t)i] = (i & 1) ? 0ull : ~0ull;
            CUDA_TRY(cudaMemcpyAsync(db->d_trace.p, init.data(), 64 * 8, cudaMemcpyHostToDevice, db->stream));
            CUDA_TRY(cudaStreamSynchronize(db->stream));
            pr->sp.trace = (unsigned long long*)db->d_trace.p;
        }
        const int64_t tile_rows = (int64_t)kDenseTileRowsPerWord * pr->sp.words_per_lane;
        const int64_t ntiles = pr->sp.ntiles, nspans = ntiles * (tile_rows / 1024);
        const int64_t nsub = ntiles * pr->sp.words_per_lane;  // 8192-row sub-tiles: the unit of the offset scan
        int rc;
        if (!pr->sp.bitmap) {
            if ((rc = ensure_buf(&db->d_bitmap, (size_t)(ntiles * tile_rows / 32 + 2) * 4))) return rc;
            pr->sp.bitmap = (uint32_t*)db->d_bitmap.p;
        }
        if ((rc = ensure_buf(&db->d_span_cnt, (size_t)nspans * 4))) return rc;
        const size_t nsub_pad = ((size_t)nsub + 4095) / 4096 * 4096 + 16;  // whole rounds of the offset scan
        if ((rc = ensure_buf(&db->d_tile_cnt, nsub_pad * 4))) return rc;
        if ((rc = ensure_buf(&db->d_tile_off, nsub_pad * 8))) return rc;
        // Emit-kernel choice known from an earlier query of this shape (see emit_hint).
        const int stream_ok = pr->emit_stage_bytes > 0 && pr->sp.nproj > 0;
        const char* force = getenv("IMM3_EMIT");  // experiment: "stream" / "gather" = that kernel takes every result, "both" = no feedback
        int hint = -1;
        if (stream_ok && !(force && !strcmp(force, "both"))) {
            auto it = db->emit_hint.find(pr->shape_key);
            if (it != db->emit_hint.end()) hint = it->second;
            if (force && !strcmp(force, "stream")) hint = 1;
            if (force && !strcmp(force, "gather")) hint = 0;
        }
        const bool only_stream = stream_ok && hint == 1, only_gather = stream_ok && hint == 0;
        const bool pdl = pr->sp.nproj > 0 && !getenv("IMM3_NO_PDL");  // the first emit kernel launched is a programmatic dependent
        const int scan_inline_d = (pr->sp.nfilter == 0 || scan_inline_for(nsub)) ? 1 : 0;
        pr->sp.scan_inline = pr->or_terms.empty() ? scan_inline_d : 0;
        CUDA_TRY(cudaEventRecord(db->ev0, db->stream));
        CUDA_TRY(launch_filter(pr->sp, pr->sp.bitmap, (uint32_t*)db->d_span_cnt.p, (uint32_t*)db->d_tile_cnt.p,
                               (unsigned long long*)db->d_tile_off.p, db->d_ctrl, pr->grid, pr->dyn_smem, db->stream));
        *launches = 1;
        if ((rc = launch_or_terms(db, pr, scan_inline_d, launches))) return rc;
        pr->sp.scan_inline = scan_inline_d;
        if (!pr->sp.scan_inline && (rc = launch_scan_kernel(db, pr->sp, nsub, launches))) return rc;
        // The streaming emit kernel is launched as a programmatic dependent of the filter kernel (its prologue overlaps the
        // filter kernel's tail), so no event may sit between the two; IMM3_NO_PDL=1 restores per-stage timing.
        if (!pdl) {
            CUDA_TRY(cudaEventRecord(db->ev_mid, db->stream));
            have_mid = true;
        }
        if (pr->sp.nproj > 0) {
            // Two emit kernels, one of which does the work: which one is decided on the device from the match count.
            if (stream_ok && !only_gather) {
                CUDA_TRY(launch_emit_stream(pr->sp, pr->sp.bitmap, (const uint32_t*)db->d_span_cnt.p, (const uint32_t*)db->d_tile_cnt.p,
                                            (const unsigned long long*)db->d_tile_off.p, nsub, pr->emit_ring, pr->emit_stage_bytes,
                                            only_stream ? -1 : 1, pr->grid_emit_stream, pr->emit_smem, db->d_ctrl, pdl, db->stream));
                (*launches)++;
            }
            if (!only_stream) {
                CUDA_TRY(launch_emit(pr->sp, pr->sp.bitmap, (const uint32_t*)db->d_span_cnt.p, (const unsigned long long*)db->d_tile_off.p, 8,
                                     nspans, pr->grid_emit, stream_ok && !only_gather, db->d_ctrl, pr->emit_general, pdl && (only_gather || !stream_ok),
                                     db->stream));
                (*launches)++;
            }
        }
        CUDA_TRY(cudaEventRecord(db->ev1, db->stream));
    } else {
        if (!pr->block_mode && getenv("IMM3_TRACE")) {  // debugging: per-tile globaltimer stamps
            int rc = ensure_buf(&db->d_trace, (size_t)pr->sp.ntiles * 64);
            if (rc) return rc;
            CUDA_TRY(cudaMemsetAsync(db->d_trace.p, 0, (size_t)pr->sp.ntiles * 64, db->stream));
            pr->sp.trace = (unsigned long long*)db->d_trace.p;
        }
        CUDA_TRY(cudaEventRecord(db->ev0, db->stream));
        CUDA_TRY(launch_scan_blocks(pr->sp, db->d_ctrl, db->d_status, pr->grid, pr->dyn_smem, db->stream));  // (single-pass block kernel: blocks > 1024 rows, bitmaps)
        CUDA_TRY(cudaEventRecord(db->ev1, db->stream));
        *launches = 1;
    }
    if (exchange && db->comm_on) {
        // The per-rank counts are exchanged by the GPUs themselves (k_comm.cuh), behind the query's last kernel: no host
        // round trip between the kernels and the exchange, one synchronisation per query.
        int rc = launch_exchange(db, pr->sp.limit, 1, publish_ok ? pub_seq : 0);
        if (rc) return rc;
        (*launches)++;
        CUDA_TRY(cudaEventRecord(db->ev1, db->stream));  // (re-recorded: the timed span now ends behind the exchange)
        published = publish_ok;
    }
    db->ev_seq = ++db->query_seq;
    if (published) {
        g_t_launched = now_us();
        const volatile unsigned long long* ps = &db->h_ctrl->pub_seq;
        bool seen = false;
        for (unsigned spins = 0;; spins++) {
            const unsigned long long word = __atomic_load_n(ps, __ATOMIC_ACQUIRE);
            if ((word >> 41) == (pub_seq & 0x7FFFFFull)) {
                if (!(exchange && db->comm_on)) {  // (blocks_scan_emit_kernel packs what the host needs into the word itself)
                    db->h_ctrl->c.total = word & ((1ull << 40) - 1ull);
                    db->h_ctrl->c.error = (word >> 40) & 1u ? 1u : 0u;
                    db->h_ctrl->c.dense_rows = 0;
                }
                seen = true;
                break;
            }
            if ((spins & 0xFFFu) == 0xFFFu && now_us() - g_t_launched > 2e5) break;  // 0.2 s: something is wrong - ask the runtime
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        if (!seen) {  // (a kernel fault, or a peer that never came: the ordinary path reports it)
            CUDA_TRY(cudaMemcpyAsync(db->h_ctrl, db->d_ctrl, sizeof(CtrlBlock) - sizeof(unsigned long long), cudaMemcpyDeviceToHost, db->stream));
            CUDA_TRY(cudaStreamSynchronize(db->stream));
        }
        g_t_synced = now_us();
    } else {
        CUDA_TRY(cudaMemcpyAsync(db->h_ctrl, db->d_ctrl, sizeof(CtrlBlock) - sizeof(unsigned long long), cudaMemcpyDeviceToHost, db->stream));
        g_t_launched = now_us();
        CUDA_TRY(cudaStreamSynchronize(db->stream));
        g_t_synced = now_us();
    }
    if (db->h_ctrl->c.error) return fail(IMM3_ERR_CUDA, "kernel watchdog fired (code %u)", db->h_ctrl->c.error);
    if (exchange && db->comm_on && db->h_ctrl->x.error)
        return fail(IMM3_ERR_COMM, "count exchange: a peer's count did not arrive within %llu ms (ranks must issue the same queries in the same order)",
                    db->comm_timeout_ns / 1000000ull);
    float f = 0;
    if (published && !have_mid && !g_eager_times && !getenv("IMM3_EAGER_TIMES")) {
        *ms = -1.0;  // read lazily (imm3_result_device_ms): the events complete a moment after the publish, waiting for them here costs 2-3 us
        if (stage_ms) stage_ms[0] = -1.0, stage_ms[1] = 0;
    } else {
    if (published) CUDA_TRY(cudaEventSynchronize(db->ev1));
    CUDA_TRY(cudaEventElapsedTime(&f, db->ev0, db->ev1));
    *ms = f;
    if (stage_ms) {
        stage_ms[0] = f;
        stage_ms[1] = 0;
        if (have_mid) {
            float a = 0, b = 0;
            CUDA_TRY(cudaEventElapsedTime(&a, db->ev0, db->ev_mid));
            CUDA_TRY(cudaEventElapsedTime(&b, db->ev_mid, db->ev1));
            stage_ms[0] = a;
            stage_ms[1] = b;
        }
    }
    }
    *total = (int64_t)db->h_ctrl->c.total;
    if (pr->multipass && pr->sp.nproj > 0 && pr->emit_stage_bytes > 0)  // feedback for the next query of this shape
        db->emit_hint[pr->shape_key] = (db->h_ctrl->c.total > 0 && db->h_ctrl->c.dense_rows * 2 >= db->h_ctrl->c.total) ? 1 : 0;
    if (pr->sp.trace && (pr->sp.debug & 16u)) {
        unsigned long long h[64];
        CUDA_TRY(cudaMemcpy(h, pr->sp.trace, sizeof h, cudaMemcpyDeviceToHost));
        if (FILE* f = fopen(getenv("IMM3_TRACE"), "w")) {
            const char* names[] = {"K1 entry", "K1 cta done", "K1 scan start", "K1 scan end", "K3s entry", "K3s first tile", "K3s cta done", "-",
                                   "K3b entry", "K3b dep ready", "K3b tiles seen", "K3b tile meta", "K3b block start", "K3b block done", "K3b warp done", "K3b words here", "K3b store start", "K3b store done"};
            for (int i = 0; i < 18; i++)
                if (h[2 * i] != ~0ull)
                    fprintf(f, "%-16s min %9.2f us  max %9.2f us\n", names[i], (double)(h[2 * i] - h[0]) / 1e3, (double)(h[2 * i + 1] - h[0]) / 1e3);
            for (int i = 0; i < 8; i++)
                if (h[32 + i] != ~0ull && h[32 + i] != 0) fprintf(f, "scan round %d after block scan: %9.2f us\n", i, (double)(h[32 + i] - h[0]) / 1e3);
            if (pr->lane) {  // per CTA of the lane kernel: done time, SM, tiles decided
                std::vector<unsigned long long> pc(1024);
                if (cudaMemcpy(pc.data(), pr->sp.trace + 64, 1024 * 8, cudaMemcpyDeviceToHost) == cudaSuccess)
                    for (int i = 0; i < 512 && pc[2 * i]; i++)
                        fprintf(f, "cta %3d done %9.2f us  sm %3llu  tiles %llu\n", i, (double)(pc[2 * i] - h[0]) / 1e3, pc[2 * i + 1] >> 32, pc[2 * i + 1] & 0xFFFFFFFFull);
            }
            fclose(f);
        }
    } else if (pr->sp.trace) {
        std::vector<unsigned long long> h((size_t)pr->sp.ntiles * 8);
        CUDA_TRY(cudaMemcpy(h.data(), pr->sp.trace, h.size() * 8, cudaMemcpyDeviceToHost));
        if (FILE* f = fopen(getenv("IMM3_TRACE"), "wb")) {
            fwrite(h.data(), 8, h.size(), f);
            fclose(f);
        }
    }
    return 0;
}

// A small LIMIT is served "prefix first": phase A scans a leading part of the slice, phase B the whole slice if that did
// not fill the LIMIT.  With a communicator every rank must run the same rounds, so the policy depends only on the query:
// phase A everywhere (a rank whose slice is too small for a prefix scans all of it), then - unless RANK 0's phase A alone
// fills the LIMIT, in which case the global result is its first `limit` rows - phase B everywhere.
bool small_limit_policy(const LogicalPlan& lp) { return lp.limit > 0 && lp.limit <= (1 << 20) && !getenv("IMM3_NO_PREFIX"); }

// Launch the kernels of one query and wait for the match count (see Prepared::prefix_blocks for the two-phase LIMIT).
int run_scan(imm3_db* db, Prepared* pr, double* ms, int64_t* total, int* launches, double* stage_ms = nullptr) {
    TableStore& t = *pr->table;
    const bool comm = db->comm_on && !pr->for_bitmap;
    const bool block_prefix = pr->block_mode && pr->blocks_multi && pr->prefix_blocks > 0;
    const bool dense_prefix = !pr->block_mode && pr->multipass && pr->prefix_rows > 0;
    const bool two_phase = comm ? small_limit_policy(pr->lp) : (block_prefix || dense_prefix);
    if (!two_phase) return run_scan_once(db, pr, ms, total, launches, stage_ms, t.nblocks, comm);
    struct EagerGuard {
        EagerGuard() { g_eager_times = true; }
        ~EagerGuard() { g_eager_times = false; }
    } eager_guard;
    // phase A: the leading blocks / rows only
    const ScanPlan full = pr->sp;
    const int grid_full = pr->grid;
    const bool local_prefix = block_prefix || dense_prefix;
    int64_t nb = t.nblocks;
    if (local_prefix) {
        nb = pr->prefix_blocks;
        if (dense_prefix) {
            pr->sp.nrows = pr->prefix_rows;
            pr->sp.ntiles = pr->prefix_rows / kDenseTileRowsPerWord;
        } else if (pr->hybrid) {
            pr->sp.nrows = (int64_t)t.row_start[(size_t)nb];
            pr->sp.ntiles = (pr->sp.nrows + kDenseTileRowsPerWord - 1) / kDenseTileRowsPerWord;
        } else {
            pr->sp.ntiles = (nb + 7) / 8;
        }
        pr->grid = (int)std::max<int64_t>(1, std::min<int64_t>(pr->sp.ntiles, grid_full));
    }
    double ms_a = 0, st_a[2] = {0, 0};
    int launches_a = 0;
    int rc = run_scan_once(db, pr, &ms_a, total, &launches_a, st_a, nb, comm);
    pr->sp = full;
    pr->grid = grid_full;
    if (rc) return rc;
    *ms = ms_a;
    *launches = launches_a;
    if (stage_ms) stage_ms[0] = st_a[0], stage_ms[1] = st_a[1];
    if (comm ? (int64_t)db->h_ctrl->x.counts[0] >= pr->lp.limit : *total >= pr->lp.limit) return 0;
    // phase B: the prefix did not fill the LIMIT - the whole table
    if (!local_prefix) return exchange_only(db, pr->sp.limit, 1);  // (comm only: phase A was this rank's whole slice already)
    double ms_b = 0, st_b[2] = {0, 0};
    int launches_b = 0;
    if ((rc = run_scan_once(db, pr, &ms_b, total, &launches_b, st_b, t.nblocks, comm))) return rc;
    *ms += ms_b;
    *launches += launches_b;
    if (stage_ms) stage_ms[0] += st_b[0], stage_ms[1] += st_b[1];
    return 0;
}

// java.lang.Double.toString for an integral value of int32 magnitude (all Min/MaxDoubleAggr ever hold here): "18.0",
// "-128.0", and computerized scientific notation from 10^7 on ("1.0E7", "2.147483647E9").
std::string java_double_to_string_integral(double v) {
    if (v == 0) return "0.0";
    const bool neg = v < 0;
    long long a = (long long)(neg ? -v : v);
    std::string digits = std::to_string(a);
    std::string out = neg ? "-" : "";
    if (a < 10000000ll) return out + digits + ".0";
    const int exp10 = (int)digits.size() - 1;
    std::string frac = digits.substr(1);
    while (frac.size() > 1 && frac.back() == '0') frac.pop_back();
    return out + digits.substr(0, 1) + "." + frac + "E" + std::to_string(exp10);
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* imm3_last_error(void) { return last_error(); }
int imm3_abi_version(void) { return IMM3_ABI_VERSION; }

int imm3_open(const char* data_dir, const imm3_open_opts* opts, imm3_db** out) {
    if (!data_dir || !out) return fail(IMM3_ERR_INVALID_ARG, "imm3_open: NULL argument");
    imm3_open_opts o = {0, 0, 1, 0};
    if (opts) o = *opts;
    if (o.world < 1 || o.rank < 0 || o.rank >= o.world) return fail(IMM3_ERR_INVALID_ARG, "imm3_open: rank %d of world %d", o.rank, o.world);
    std::unique_ptr<imm3_db> db(new imm3_db());
    db->dir = data_dir;
    db->device = o.device;
    db->rank = o.rank;
    db->world = o.world;
    db->flags = o.flags;
    db->host_only = (o.flags & IMM3_OPEN_HOST_ONLY) != 0;
    db->host_pool.pinned_host = true;
    const bool otrace = getenv("IMM3_OPEN_TRACE") != nullptr;  // phase times of the open path on stderr
    const double t_open0 = now_us();
    int rc = load_tables(db->dir, o.rank, o.world, &db->tables);
    if (rc) return rc;
    if (otrace) fprintf(stderr, "[imm3_open] metadata + validation %.1f ms (%d I/O threads)\n", (now_us() - t_open0) * 1e-3, io_threads());
    if (!db->host_only) {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            return fail(IMM3_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                        e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        }
        if (o.device < 0 || o.device >= ndev) return fail(IMM3_ERR_INVALID_ARG, "imm3_open: device %d of %d", o.device, ndev);
        auto body = [&]() -> int {
            CUDA_TRY(cudaSetDevice(db->device));
            cudaDeviceProp prop;
            CUDA_TRY(cudaGetDeviceProperties(&prop, db->device));
            if (prop.major < 10) return fail(IMM3_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", db->device, prop.major, prop.minor);
            db->num_sms = prop.multiProcessorCount;
            if (const char* g = getenv("IMM3_L2_FETCH")) {  // experiment: DRAM->L2 fetch granularity hint (32 / 64 / 128 bytes)
                CUDA_TRY(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g)));
            }
            CUDA_TRY(cudaStreamCreateWithFlags(&db->own_stream, cudaStreamNonBlocking));
            db->stream = db->own_stream;
            CUDA_TRY(cudaEventCreate(&db->ev0));
            CUDA_TRY(cudaEventCreate(&db->ev1));
            CUDA_TRY(cudaEventCreate(&db->ev_mid));
            CUDA_TRY(cudaMalloc(&db->d_ctrl, sizeof(CtrlBlock)));
            CUDA_TRY(cudaMemsetAsync(db->d_ctrl, 0, sizeof(CtrlBlock), db->stream));
            CUDA_TRY(cudaMallocHost(&db->h_ctrl, sizeof(CtrlBlock)));
            std::memset(db->h_ctrl, 0, sizeof(CtrlBlock));
            if (const char* e = getenv("IMM3_COMM_TIMEOUT_MS")) db->comm_timeout_ns = (unsigned long long)std::max(1, atoi(e)) * 1000000ull;
            if (otrace) fprintf(stderr, "[imm3_open] device + stream setup at %.1f ms\n", (now_us() - t_open0) * 1e-3);
            return upload_all(db.get());
        };
        rc = body();
        if (rc) {
            std::string why = last_error();
            free_device_side(db.get());
            return fail(rc, "%s", why.c_str());
        }
    }
    *out = db.release();
    return 0;
}

int imm3_close(imm3_db* db) {
    if (!db) return 0;
    // Results borrow buffers from the db's pools and dereference it in fetch / wait / free: closing under them would be a
    // use-after-free.  The caller frees its results first (the Python mirror does so in SegmentManager.close).
    if (db->live_results > 0)
        return fail(IMM3_ERR_STATE, "imm3_close: %d result(s) of this handle are still open (imm3_result_free them first)", db->live_results);
    free_device_side(db);
    delete db;
    return 0;
}

// ---- count exchange over NVLink peer memory -----------------------------------------------------
int imm3_comm_local_handle(imm3_db* db, void* handle_out) {
    if (!db || !handle_out) return fail(IMM3_ERR_INVALID_ARG, "imm3_comm_local_handle: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == IMM3_COMM_HANDLE_BYTES, "IMM3_COMM_HANDLE_BYTES must be sizeof(cudaIpcMemHandle_t)");
    int rc = use_device(db);
    if (rc) return rc;
    if (db->world > kMaxWorld) return fail(IMM3_ERR_UNSUPPORTED, "count exchange supports at most %d ranks (world = %d)", kMaxWorld, db->world);
    if (!db->d_mailbox) {
        CUDA_TRY(cudaMalloc(&db->d_mailbox, kMailboxBytes));
        CUDA_TRY(cudaMemset(db->d_mailbox, 0, kMailboxBytes));
    }
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, db->d_mailbox));
    std::memcpy(handle_out, &h, sizeof h);
    return 0;
}

int imm3_comm_connect(imm3_db* db, const void* handles, int nhandles) {
    if (!db || !handles) return fail(IMM3_ERR_INVALID_ARG, "imm3_comm_connect: NULL argument");
    int rc = use_device(db);
    if (rc) return rc;
    if (nhandles != db->world) return fail(IMM3_ERR_INVALID_ARG, "imm3_comm_connect: %d handles for a world of %d", nhandles, db->world);
    if (!db->d_mailbox) return fail(IMM3_ERR_STATE, "imm3_comm_connect: call imm3_comm_local_handle first");
    if (db->comm_on) return fail(IMM3_ERR_STATE, "imm3_comm_connect: already connected");
    for (int i = 0; i < db->world; i++) {
        if (i == db->rank) {
            db->comm_peer[i] = db->d_mailbox;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const uint8_t*)handles + (size_t)i * sizeof h, sizeof h);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int k = 0; k < i; k++) {
                if (db->comm_peer[k] && db->comm_peer[k] != db->d_mailbox) cudaIpcCloseMemHandle(db->comm_peer[k]);
                db->comm_peer[k] = nullptr;
            }
            return fail(IMM3_ERR_COMM, "imm3_comm_connect: cannot map the mailbox of rank %d (%s); the ranks must be processes on one NVLink-connected node", i,
                        cudaGetErrorString(e));
        }
        db->comm_peer[i] = (unsigned long long*)p;
    }
    db->comm_on = db->world > 1;
    db->comm_epoch = 0;
    return 0;
}

int imm3_table_count(imm3_db* db) { return db ? (int)db->tables.size() : IMM3_ERR_INVALID_ARG; }

const char* imm3_table_name(imm3_db* db, int idx) {
    if (!db || idx < 0 || idx >= (int)db->tables.size()) return nullptr;
    return db->tables[(size_t)idx].meta.name.c_str();
}

int imm3_table_info(imm3_db* db, const char* table, imm3_table_desc* out) {
    if (!db || !out) return fail(IMM3_ERR_INVALID_ARG, "imm3_table_info: NULL argument");
    TableStore* t = find_table(db, table);
    if (!t) return IMM3_ERR_NOT_FOUND;
    out->ncols = (int)t->cols.size();
    out->block_size = t->meta.block_size;
    out->nsegments = t->nsegments;
    out->seg_begin = t->seg_begin;
    out->seg_end = t->seg_end;
    out->nrows = t->nrows;
    out->nblocks = t->nblocks;
    out->resident_bytes = 0;
    for (auto& c : t->cols) out->resident_bytes += c.encoded_bytes;
    return 0;
}

int imm3_column_info(imm3_db* db, const char* table, int col_idx, imm3_column_desc* out) {
    if (!db || !out) return fail(IMM3_ERR_INVALID_ARG, "imm3_column_info: NULL argument");
    TableStore* t = find_table(db, table);
    if (!t) return IMM3_ERR_NOT_FOUND;
    if (col_idx < 0 || col_idx >= (int)t->cols.size()) return fail(IMM3_ERR_NOT_FOUND, "column index %d out of range", col_idx);
    const ColumnStore& c = t->cols[(size_t)col_idx];
    std::memset(out, 0, sizeof *out);
    snprintf(out->name, sizeof out->name, "%s", c.meta.name.c_str());
    out->column_type = c.meta.ctype;
    out->codec = c.meta.codec;
    out->width = c.meta.width;
    out->encoded_bytes = c.encoded_bytes;
    return 0;
}

int imm3_segment_file_id(imm3_db* db, const char* table, int canonical_idx, int32_t* out_id) {
    if (!db || !out_id) return fail(IMM3_ERR_INVALID_ARG, "imm3_segment_file_id: NULL argument");
    TableStore* t = find_table(db, table);
    if (!t) return IMM3_ERR_NOT_FOUND;
    if (canonical_idx < 0 || canonical_idx >= t->nsegments) return fail(IMM3_ERR_NOT_FOUND, "segment index %d out of range", canonical_idx);
    *out_id = t->file_ids[(size_t)canonical_idx];
    return 0;
}

int imm3_set_stream(imm3_db* db, void* cuda_stream) {
    if (!db) return fail(IMM3_ERR_INVALID_ARG, "imm3_set_stream: db is NULL");
    int rc = use_device(db);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(db->stream));
    db->stream = cuda_stream ? (cudaStream_t)cuda_stream : db->own_stream;
    return 0;
}

int imm3_sync(imm3_db* db) {
    if (!db) return fail(IMM3_ERR_INVALID_ARG, "imm3_sync: db is NULL");
    int rc = use_device(db);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(db->stream));
    return 0;
}

int imm3_reupload(imm3_db* db, const char* table, const char* const* cols, int ncols, int64_t* out_bytes) {
    if (!db) return fail(IMM3_ERR_INVALID_ARG, "imm3_reupload: db is NULL");
    int rc = use_device(db);
    if (rc) return rc;
    if (!(db->flags & IMM3_OPEN_KEEP_HOST)) return fail(IMM3_ERR_STATE, "imm3_reupload needs IMM3_OPEN_KEEP_HOST");
    TableStore* t = find_table(db, table);
    if (!t) return IMM3_ERR_NOT_FOUND;
    int64_t bytes = 0;
    for (auto& c : t->cols) {
        bool want = ncols <= 0 || !cols;
        for (int i = 0; !want && i < ncols; i++) want = cols[i] && c.meta.name == cols[i];
        if (!want || !c.encoded_bytes) continue;
        if ((rc = ensure_mirror(db, c))) return rc;  // first use: pin + read the column's files once
        CUDA_TRY(cudaMemcpyAsync(c.d_arena, c.h_mirror, (size_t)c.encoded_bytes, cudaMemcpyHostToDevice, db->stream));
        bytes += c.encoded_bytes;
    }
    if (ncols > 0 && cols)
        for (int i = 0; i < ncols; i++) {
            bool found = false;
            for (auto& c : t->cols) found = found || (cols[i] && c.meta.name == cols[i]);
            if (!found) return fail(IMM3_ERR_NOT_FOUND, "Column %s does not exist in table %s", cols[i] ? cols[i] : "(null)", t->meta.name.c_str());
        }
    if (out_bytes) *out_bytes = bytes;
    return 0;
}

int imm3_explain(imm3_db* db, const char* table, const imm3_pred* preds, int npreds, const char* const* proj_cols,
                 int nproj, int64_t limit, const char** json) {
    if (!db || !json) return fail(IMM3_ERR_INVALID_ARG, "imm3_explain: NULL argument");
    Prepared pr;
    int rc = prepare(db, table, preds, npreds, proj_cols, nproj, limit, &pr);
    if (rc) return rc;
    bool bm;
    db->explain_buf = explain_json(pr.lp, kernel_name(db, *pr.table, pr.lp, &bm));
    *json = db->explain_buf.c_str();
    return 0;
}

// One conjunction (imm3_query_begin) or a disjunction of conjunctions (imm3_query_begin_dnf): terms[i] = (predicates, count).
static int query_begin_terms(imm3_db* db, const char* table, const std::vector<std::pair<const imm3_pred*, int>>& terms,
                             const char* const* proj_cols, int nproj, int64_t limit, imm3_result** out) {
    if (!db || !out) return fail(IMM3_ERR_INVALID_ARG, "imm3_query_begin: NULL argument");
    const double t_in = now_us();
    // validation before any device work; a term that can never hold (a wrong-length literal, an empty range) drops out of a
    // disjunction; the first term that can hold carries the query, the others only add their filter kernels
    std::vector<std::unique_ptr<Prepared>> prepared;
    int rc = 0;
    for (auto& tm : terms) {
        std::unique_ptr<Prepared> p(new Prepared());
        if ((rc = prepare(db, table, tm.first, tm.second, proj_cols, nproj, limit, p.get()))) return rc;
        prepared.push_back(std::move(p));
    }
    size_t main_i = 0;
    while (main_i + 1 < prepared.size() && prepared[main_i]->lp.always_empty) main_i++;
    Prepared& pr = *prepared[main_i];
    std::vector<std::unique_ptr<Prepared>> extra;
    for (size_t i = main_i + 1; i < prepared.size(); i++)
        if (!prepared[i]->lp.always_empty) extra.push_back(std::move(prepared[i]));
    if ((rc = use_device(db))) return rc;
    const double t_plan = now_us();
    double t_dev = t_plan;
    g_t_launched = g_t_synced = 0;
    TableStore& t = *pr.table;
    std::unique_ptr<imm3_result> r(new imm3_result());
    r->db = db;
    r->ncols = (int)pr.lp.proj.size();
    const int64_t capacity = limit > 0 ? std::min<int64_t>(limit, t.nrows) : t.nrows;
    auto give_back = [&]() {
        for (auto& b : r->d_cols) db->dev_pool.release(b);
        r->d_cols.clear();
    };
    for (int i = 0; i < r->ncols; i++) {
        const ColumnMeta& c = t.meta.cols[(size_t)pr.lp.proj[(size_t)i]];
        r->names.push_back(c.name);
        r->types.push_back(c.ctype);
        r->widths.push_back(c.width);
        Buf b;
        if ((rc = db->dev_pool.acquire((size_t)capacity * (size_t)c.width, &b))) { give_back(); return rc; }
        r->d_cols.push_back(b);
        r->h_cols.emplace_back();
    }
    if (!pr.lp.always_empty && t.nrows > 0) {
        if ((rc = fill_scan_plan(db, &pr))) { give_back(); return rc; }
        for (int i = 0; i < r->ncols; i++) pr.sp.proj[i].out = (uint8_t*)r->d_cols[(size_t)i].p;
        pr.sp.bitmap = nullptr;
        if (!extra.empty()) {
            // real OR: every term's predicates are decided by the row-space filter kernel (dense columns); a disjunction over
            // an encoded column would need the block filter kernels to accumulate as well - not built
            auto row_space = [](const Prepared& p) { return p.block_mode ? (p.blocks_multi && p.hybrid) : p.multipass; };
            bool ok = row_space(pr);
            for (auto& tm : extra) {
                if ((rc = fill_scan_plan(db, tm.get()))) { give_back(); return rc; }
                ok = ok && row_space(*tm);
            }
            if (!ok) {
                give_back();
                return fail(IMM3_ERR_UNSUPPORTED, "imm3_query_begin_dnf: a disjunction needs every predicate on a dense (DENSE_INT / DENSE_TINYINT / DENSE_STRING) column");
            }
            pr.or_terms = std::move(extra);
        }
        t_dev = now_us();
        if ((rc = run_scan(db, &pr, &r->device_ms, &r->local_count, &r->launches, r->stage_ms))) { give_back(); return rc; }
        r->timing_seq = db->ev_seq;
    } else if (db->comm_on && !pr.lp.always_empty) {
        // Empty slice (more ranks than segments): this rank still takes part in every round of the exchange.  (A predicate
        // that can never hold is a property of the query, known to every rank: nobody exchanges anything.)
        const int64_t lim = limit > 0 ? limit : INT64_MAX;
        if ((rc = exchange_only(db, lim, 0))) { give_back(); return rc; }
        if (small_limit_policy(pr.lp) && (int64_t)db->h_ctrl->x.counts[0] < pr.lp.limit && (rc = exchange_only(db, lim, 0))) { give_back(); return rc; }
        r->launches = 1;
    }
    // Global placement of this rank's rows: from the on-device exchange, or trivially for a single handle.
    r->world = 1;
    r->g_offset = 0;
    r->g_take = r->g_total = r->local_count;
    r->rank_counts[0] = r->local_count;
    if (db->comm_on && !pr.lp.always_empty) {
        const CommOut& x = db->h_ctrl->x;
        r->world = db->world;
        r->g_offset = (int64_t)x.g_offset;
        r->g_take = (int64_t)x.g_take;
        r->g_total = (int64_t)x.g_total;
        for (int i = 0; i < db->world; i++) r->rank_counts[i] = (int64_t)x.counts[i];
    } else if (db->comm_on) {
        r->world = db->world;
    }
    // Algorithmic bytes: filter columns' encoded bytes once + per surviving row the
    // project-only widths read and every projected width written (+ 4 B/block for PFOR offsets).
    {
        int64_t a = 0, per_row = 0;
        std::vector<int> seen;
        auto count_filters = [&](const LogicalPlan& lp) {
            for (auto& f : lp.filters) {
                if (std::find(seen.begin(), seen.end(), f.col_idx) != seen.end()) continue;  // (a column several terms of a disjunction test: once)
                const ColumnStore& c = t.cols[(size_t)f.col_idx];
                a += c.encoded_bytes + (c.meta.codec == IMM3_CODEC_PFOR_INT ? 4 * t.nblocks : 0);
                seen.push_back(f.col_idx);
            }
        };
        count_filters(pr.lp);
        for (auto& tm : pr.or_terms) count_filters(tm->lp);
        for (int ci : pr.lp.proj) {
            per_row += t.cols[(size_t)ci].meta.width;
            if (std::find(seen.begin(), seen.end(), ci) == seen.end()) {
                per_row += t.cols[(size_t)ci].meta.width;
                seen.push_back(ci);
            }
        }
        r->alg_bytes = a + per_row * r->local_count;
    }
    db->live_results++;
    {
        const double t_out = now_us();
        r->host_us[0] = t_plan - t_in;
        r->host_us[1] = t_dev - t_plan;
        if (g_t_launched > 0) {
            r->host_us[2] = g_t_launched - t_dev;     // (two-phase LIMIT queries: the last phase's stamps)
            r->host_us[3] = g_t_synced - g_t_launched;
            r->host_us[4] = t_out - g_t_synced;
        }
    }
    *out = r.release();
    return 0;
}

int imm3_query_begin(imm3_db* db, const char* table, const imm3_pred* preds, int npreds, const char* const* proj_cols,
                     int nproj, int64_t limit, imm3_result** out) {
    if (npreds < 0 || (npreds > 0 && !preds)) return fail(IMM3_ERR_INVALID_ARG, "imm3_query_begin: predicates");
    return query_begin_terms(db, table, {{preds, npreds}}, proj_cols, nproj, limit, out);
}

int imm3_query_begin_dnf(imm3_db* db, const char* table, const imm3_pred* preds, const int32_t* term_sizes, int nterms,
                         const char* const* proj_cols, int nproj, int64_t limit, imm3_result** out) {
    if (!term_sizes || nterms < 1 || nterms > IMM3_MAX_OR_TERMS) return fail(IMM3_ERR_INVALID_ARG, "imm3_query_begin_dnf: 1 .. %d terms", IMM3_MAX_OR_TERMS);
    std::vector<std::pair<const imm3_pred*, int>> terms;
    int at = 0;
    for (int i = 0; i < nterms; i++) {
        if (term_sizes[i] < 0 || (term_sizes[i] > 0 && !preds)) return fail(IMM3_ERR_INVALID_ARG, "imm3_query_begin_dnf: term %d", i);
        if (term_sizes[i] == 0) return query_begin_terms(db, table, {{nullptr, 0}}, proj_cols, nproj, limit, out);  // `... or true`: every row
        terms.push_back({preds + at, term_sizes[i]});
        at += term_sizes[i];
    }
    return query_begin_terms(db, table, terms, proj_cols, nproj, limit, out);
}

int64_t imm3_result_local_count(const imm3_result* r) { return r ? r->local_count : IMM3_ERR_INVALID_ARG; }
int64_t imm3_result_global_offset(const imm3_result* r) { return r ? r->g_offset : IMM3_ERR_INVALID_ARG; }
int64_t imm3_result_take(const imm3_result* r) { return r ? r->g_take : IMM3_ERR_INVALID_ARG; }
int64_t imm3_result_global_count(const imm3_result* r) { return r ? r->g_total : IMM3_ERR_INVALID_ARG; }
int imm3_result_rank_counts(const imm3_result* r, int64_t* counts, int cap) {
    if (!r || !counts || cap < r->world) return fail(IMM3_ERR_INVALID_ARG, "imm3_result_rank_counts: need room for %d counts", r ? r->world : 0);
    for (int i = 0; i < r->world; i++) counts[i] = r->rank_counts[i];
    return r->world;
}

int imm3_result_fetch(imm3_result* r, int64_t nrows) {
    if (!r) return fail(IMM3_ERR_INVALID_ARG, "imm3_result_fetch: result is NULL");
    if (nrows < 0 || nrows > r->local_count) return fail(IMM3_ERR_INVALID_ARG, "imm3_result_fetch: %lld rows of %lld", (long long)nrows, (long long)r->local_count);
    if (r->agg_first_col >= 0) return 0;  // (aggregate results are materialised on the host by imm3_query_agg)
    imm3_db* db = r->db;
    int rc = use_device(db);
    if (rc) return rc;
    for (int i = 0; i < r->ncols; i++) {
        const size_t bytes = (size_t)nrows * (size_t)r->widths[(size_t)i];
        if (r->h_cols[(size_t)i].cap < bytes || !r->h_cols[(size_t)i].p) {
            db->host_pool.release(r->h_cols[(size_t)i]);
            r->h_cols[(size_t)i] = Buf();
            if ((rc = db->host_pool.acquire(bytes, &r->h_cols[(size_t)i]))) return rc;
        }
        if (bytes) CUDA_TRY(cudaMemcpyAsync(r->h_cols[(size_t)i].p, r->d_cols[(size_t)i].p, bytes, cudaMemcpyDeviceToHost, db->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(db->stream));
    r->fetched = nrows;
    return 0;
}

int imm3_result_fetch_async(imm3_result* r, int64_t nrows) {
    if (!r) return fail(IMM3_ERR_INVALID_ARG, "imm3_result_fetch_async: result is NULL");
    if (nrows < 0 || nrows > r->local_count) return fail(IMM3_ERR_INVALID_ARG, "imm3_result_fetch_async: %lld rows of %lld", (long long)nrows, (long long)r->local_count);
    if (r->agg_first_col >= 0) return 0;
    imm3_db* db = r->db;
    int rc = use_device(db);
    if (rc) return rc;
    if (!db->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&db->copy_stream, cudaStreamNonBlocking));
    // (imm3_query_begin returned after the kernels had finished: the rows are final, the copy stream needs no event)
    for (int i = 0; i < r->ncols; i++) {
        const size_t bytes = (size_t)nrows * (size_t)r->widths[(size_t)i];
        if (r->h_cols[(size_t)i].cap < bytes || !r->h_cols[(size_t)i].p) {
            db->host_pool.release(r->h_cols[(size_t)i]);
            r->h_cols[(size_t)i] = Buf();
            if ((rc = db->host_pool.acquire(bytes, &r->h_cols[(size_t)i]))) return rc;
        }
        if (bytes) CUDA_TRY(cudaMemcpyAsync(r->h_cols[(size_t)i].p, r->d_cols[(size_t)i].p, bytes, cudaMemcpyDeviceToHost, db->copy_stream));
    }
    if (!r->copied) CUDA_TRY(cudaEventCreateWithFlags(&r->copied, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(r->copied, db->copy_stream));
    r->pending = nrows;
    return 0;
}

int imm3_result_wait(imm3_result* r) {
    if (!r) return fail(IMM3_ERR_INVALID_ARG, "imm3_result_wait: result is NULL");
    if (r->pending < 0) return 0;  // nothing in flight
    int rc = use_device(r->db);
    if (rc) return rc;
    CUDA_TRY(cudaEventSynchronize(r->copied));
    r->fetched = r->pending;
    r->pending = -1;
    return 0;
}

int imm3_query(imm3_db* db, const char* table, const imm3_pred* preds, int npreds, const char* const* proj_cols,
               int nproj, int64_t limit, imm3_result** out) {
    imm3_result* r = nullptr;
    int rc = imm3_query_begin(db, table, preds, npreds, proj_cols, nproj, limit, &r);
    if (rc) return rc;
    if ((rc = imm3_result_fetch(r, r->g_take))) {  // (a single handle: all local rows; with a communicator: this rank's share)
        std::string why = last_error();
        imm3_result_free(r);
        return fail(rc, "%s", why.c_str());
    }
    *out = r;
    return 0;
}

int imm3_query_agg(imm3_db* db, const char* table, const imm3_pred* preds, int npreds, const imm3_agg* aggs, int naggs,
                   const char* const* group_cols, int ngroup, imm3_result** out) {
    if (!db || !out || (naggs > 0 && !aggs) || (ngroup > 0 && !group_cols)) return fail(IMM3_ERR_INVALID_ARG, "imm3_query_agg: NULL argument");
    if (naggs < 1) return fail(IMM3_ERR_INVALID_ARG, "imm3_query_agg: no aggregates");
    if (naggs > kMaxAggs) return fail(IMM3_ERR_UNSUPPORTED, "imm3_query_agg: at most %d aggregates", kMaxAggs);
    if (ngroup > kMaxGroupCols) return fail(IMM3_ERR_UNSUPPORTED, "imm3_query_agg: at most %d group-by columns", kMaxGroupCols);
    Prepared pr;
    int rc = prepare(db, table, preds, npreds, nullptr, 0, 0, &pr);  // validation before any device work
    if (rc) return rc;
    TableStore& t = *pr.table;
    auto find_col = [&](const char* name) -> int {
        if (!name) return -1;
        for (size_t i = 0; i < t.cols.size(); i++)
            if (t.cols[i].meta.name == name) return (int)i;
        return -1;
    };
    AggPlan ap;
    std::memset(&ap, 0, sizeof ap);
    std::unique_ptr<imm3_result> r(new imm3_result());
    r->db = db;
    std::vector<int> gidx, aidx;
    int key_bits = 0;
    for (int g = 0; g < ngroup; g++) {
        const int ci = find_col(group_cols[g]);
        if (ci < 0) return fail(IMM3_ERR_NOT_FOUND, "Column %s does not exist in table %s", group_cols[g] ? group_cols[g] : "(null)", t.meta.name.c_str());
        const ColumnStore& c = t.cols[(size_t)ci];
        if (c.meta.codec == IMM3_CODEC_PFOR_INT) return fail(IMM3_ERR_UNSUPPORTED, "group by a sorted-int-codec column (%s) is not supported", c.meta.name.c_str());
        if (key_bits + 8 * c.meta.width > 56) return fail(IMM3_ERR_UNSUPPORTED, "group-by cells must pack into 7 bytes");
        ap.group[g].width = c.meta.width;
        ap.group[g].key_shift = key_bits;
        key_bits += 8 * c.meta.width;
        gidx.push_back(ci);
        r->names.push_back(c.meta.name);
        r->types.push_back(c.meta.ctype);
        r->widths.push_back(c.meta.width);
    }
    for (int a = 0; a < naggs; a++) {
        const int ci = find_col(aggs[a].col);
        if (ci < 0) return fail(IMM3_ERR_NOT_FOUND, "Column %s does not exist in table %s", aggs[a].col ? aggs[a].col : "(null)", t.meta.name.c_str());
        const ColumnStore& c = t.cols[(size_t)ci];
        const char* suffix = "_count";
        if (aggs[a].op == IMM3_AGG_COUNT) {
            ap.agg[a].op = kAggCount;
        } else if (aggs[a].op == IMM3_AGG_MIN || aggs[a].op == IMM3_AGG_MAX) {
            if (c.meta.ctype == IMM3_COL_STRING)
                return fail(IMM3_ERR_UNSUPPORTED, "min / max on a STRING column (the reference maps both to MaxStringAggr, Engine.scala:137,147)");
            if (c.meta.codec == IMM3_CODEC_PFOR_INT) return fail(IMM3_ERR_UNSUPPORTED, "min / max on a sorted-int-codec column (%s) is not supported", c.meta.name.c_str());
            ap.agg[a].op = aggs[a].op == IMM3_AGG_MIN ? kAggMin : kAggMax;
            suffix = aggs[a].op == IMM3_AGG_MIN ? "_min" : "_max";
        } else {
            return fail(IMM3_ERR_UNSUPPORTED, "Unknown Aggregate type (Engine.scala:153: only Min, Max and Count are resolved)");
        }
        ap.agg[a].width = c.meta.width;
        aidx.push_back(ci);
        r->names.push_back(c.meta.name + suffix);  // alias.getOrElse(col + "_max"), Engine.scala:136-152
        r->types.push_back(ap.agg[a].op == kAggCount ? IMM3_COL_COUNT : IMM3_COL_DOUBLE);
        r->widths.push_back(8);
    }
    r->ncols = ngroup + naggs;
    r->agg_first_col = ngroup;
    r->h_cols.resize((size_t)r->ncols);
    ap.naggs = naggs;
    ap.ngroup = ngroup;
    if (pr.block_mode && !pr.lp.filters.empty()) {
        for (auto& f : pr.lp.filters)
            if (t.cols[(size_t)f.col_idx].meta.codec == IMM3_CODEC_PFOR_INT)
                return fail(IMM3_ERR_UNSUPPORTED, "aggregation with a predicate on a sorted-int-codec column (%s) is not supported yet", t.cols[(size_t)f.col_idx].meta.name.c_str());
    }
    if ((rc = use_device(db))) return rc;
    int64_t ngroups_out = 0;
    std::vector<AggEntry> groups;
    if (!pr.lp.always_empty && t.nrows > 0) {
        // The filter runs in row space exactly as for a Project query without a select list (dense filter kernel: bitmap,
        // span counts, total); the aggregation kernel is queued behind it, one synchronisation for the whole query.
        const bool was_block = pr.block_mode;
        pr.block_mode = false;
        pr.multipass = true;
        pr.for_bitmap = false;
        if ((rc = fill_scan_plan(db, &pr))) return rc;
        pr.block_mode = was_block;
        ScanPlan& sp = pr.sp;
        sp.bitmap = nullptr;
        const int64_t ntiles = sp.ntiles, nspans = ntiles * 8;
        if ((rc = ensure_buf(&db->d_bitmap, (size_t)(ntiles * kDenseTileRowsPerWord / 32 + 64) * 4))) return rc;
        if ((rc = ensure_buf(&db->d_span_cnt, (size_t)(nspans + 8) * 4))) return rc;
        const size_t ntiles_pad = ((size_t)ntiles + 4095) / 4096 * 4096 + 16;
        if ((rc = ensure_buf(&db->d_tile_cnt, ntiles_pad * 4))) return rc;
        if ((rc = ensure_buf(&db->d_tile_off, ntiles_pad * 8))) return rc;
        ap.table_slots = 1u << 16;
        if (const char* e = getenv("IMM3_AGG_SLOTS")) {  // (tests shrink it to exercise the overflow report)
            uint32_t v = (uint32_t)atoi(e);
            if (v >= 64 && (v & (v - 1)) == 0) ap.table_slots = v;
        }
        if ((rc = ensure_buf(&db->d_agg_table, (size_t)ap.table_slots * sizeof(AggEntry)))) return rc;
        if ((rc = ensure_buf(&db->d_agg_out, (size_t)ap.table_slots * sizeof(AggEntry)))) return rc;
        if ((rc = ensure_buf(&db->d_agg_cnt, 64))) return rc;
        ap.nrows = t.nrows;
        ap.ntiles = ntiles;
        for (int g = 0; g < ngroup; g++) ap.group[g].base = t.cols[(size_t)gidx[(size_t)g]].d_arena;
        for (int a = 0; a < naggs; a++) ap.agg[a].base = t.cols[(size_t)aidx[(size_t)a]].d_arena;
        sp.scan_inline = 1;  // (only the total matters here; the last CTA's scan provides it)
        CUDA_TRY(cudaEventRecord(db->ev0, db->stream));
        CUDA_TRY(launch_agg_init((AggEntry*)db->d_agg_table.p, ap.table_slots, ap, (unsigned int*)db->d_agg_cnt.p, db->stream));
        CUDA_TRY(launch_filter(sp, (uint32_t*)db->d_bitmap.p, (uint32_t*)db->d_span_cnt.p, (uint32_t*)db->d_tile_cnt.p,
                               (unsigned long long*)db->d_tile_off.p, db->d_ctrl, pr.grid, pr.dyn_smem, db->stream));
        CUDA_TRY(launch_agg(ap, (const uint32_t*)db->d_bitmap.p, (const uint32_t*)db->d_span_cnt.p, (AggEntry*)db->d_agg_table.p,
                            (AggEntry*)db->d_agg_out.p, (unsigned int*)db->d_agg_cnt.p, db->d_ctrl, db->num_sms, db->stream));
        CUDA_TRY(cudaEventRecord(db->ev1, db->stream));
        r->launches = 4;
        unsigned int counters[2] = {0, 0};
        CUDA_TRY(cudaMemcpyAsync(db->h_ctrl, db->d_ctrl, sizeof(CtrlBlock), cudaMemcpyDeviceToHost, db->stream));
        CUDA_TRY(cudaMemcpyAsync(counters, db->d_agg_cnt.p, sizeof counters, cudaMemcpyDeviceToHost, db->stream));
        CUDA_TRY(cudaStreamSynchronize(db->stream));
        if (db->h_ctrl->c.error) return fail(IMM3_ERR_CUDA, "kernel watchdog fired (code %u)", db->h_ctrl->c.error);
        if (counters[1]) return fail(IMM3_ERR_UNSUPPORTED, "aggregation: more than %u groups", ap.table_slots);
        float f = 0;
        CUDA_TRY(cudaEventElapsedTime(&f, db->ev0, db->ev1));
        r->device_ms = f;
        r->stage_ms[0] = f;
        ngroups_out = counters[0];
        groups.resize((size_t)ngroups_out);
        if (ngroups_out) CUDA_TRY(cudaMemcpy(groups.data(), db->d_agg_out.p, (size_t)ngroups_out * sizeof(AggEntry), cudaMemcpyDeviceToHost));
        std::sort(groups.begin(), groups.end(), [](const AggEntry& x, const AggEntry& y) { return x.first_row < y.first_row; });
        // algorithmic bytes: filter columns once + per selected row the cells of the group and aggregate columns
        int64_t bytes = 0, per_row = 0;
        for (auto& f2 : pr.lp.filters) bytes += t.cols[(size_t)f2.col_idx].encoded_bytes;
        for (int ci : gidx) per_row += t.cols[(size_t)ci].meta.width;
        for (int a = 0; a < naggs; a++)
            if (ap.agg[a].op != kAggCount) per_row += ap.agg[a].width;
        r->alg_bytes = bytes + per_row * (int64_t)db->h_ctrl->c.total;
    }
    // rows of the result: group cells, then the aggregates (COUNT: int64, MIN / MAX: double)
    for (int c = 0; c < r->ncols; c++) {
        const size_t w = (size_t)r->widths[(size_t)c];
        if ((rc = db->host_pool.acquire(std::max<size_t>(1, (size_t)ngroups_out * w), &r->h_cols[(size_t)c]))) {
            for (auto& b : r->h_cols) db->host_pool.release(b);
            return rc;
        }
        uint8_t* dst = (uint8_t*)r->h_cols[(size_t)c].p;
        for (int64_t i = 0; i < ngroups_out; i++) {
            const AggEntry& e = groups[(size_t)i];
            if (c < ngroup) {
                const unsigned long long cell = e.key >> ap.group[c].key_shift;
                for (size_t b = 0; b < w; b++) dst[(size_t)i * w + b] = (uint8_t)(cell >> (8 * b));
            } else {
                const int a = c - ngroup;
                if (ap.agg[a].op == kAggCount) {
                    const int64_t v = e.val[a];
                    std::memcpy(dst + (size_t)i * 8, &v, 8);
                } else {
                    const double v = (double)e.val[a];  // value.toDouble, ProjectAggregate.scala:176-178
                    std::memcpy(dst + (size_t)i * 8, &v, 8);
                }
            }
        }
    }
    r->local_count = r->fetched = ngroups_out;
    r->g_offset = 0;
    r->g_take = r->g_total = ngroups_out;
    r->world = 1;
    db->live_results++;
    *out = r.release();
    return 0;
}

int imm3_query_sql(imm3_db* db, const char* sql, imm3_result** out) {
    if (!db || !out) return fail(IMM3_ERR_INVALID_ARG, "imm3_query_sql: NULL argument");
    ParsedQuery q;
    int rc = parse_sql(sql, &q);
    if (rc) return rc;
    std::vector<const char*> proj;
    for (auto& s : q.proj) proj.push_back(s.c_str());
    return imm3_query(db, q.table.c_str(), q.preds.data(), (int)q.preds.size(), proj.data(), (int)proj.size(), q.limit, out);
}

int64_t imm3_result_nrows(const imm3_result* r) { return r ? r->fetched : IMM3_ERR_INVALID_ARG; }
int imm3_result_ncols(const imm3_result* r) { return r ? r->ncols : IMM3_ERR_INVALID_ARG; }
int imm3_result_col_type(const imm3_result* r, int c) { return (r && c >= 0 && c < r->ncols) ? r->types[(size_t)c] : IMM3_ERR_INVALID_ARG; }
int imm3_result_col_width(const imm3_result* r, int c) { return (r && c >= 0 && c < r->ncols) ? r->widths[(size_t)c] : IMM3_ERR_INVALID_ARG; }
const char* imm3_result_col_name(const imm3_result* r, int c) { return (r && c >= 0 && c < r->ncols) ? r->names[(size_t)c].c_str() : nullptr; }
const void* imm3_result_col_data(const imm3_result* r, int c) { return (r && c >= 0 && c < r->ncols) ? r->h_cols[(size_t)c].p : nullptr; }
const void* imm3_result_col_device(const imm3_result* r, int c) { return (r && c >= 0 && c < r->ncols && (size_t)c < r->d_cols.size()) ? r->d_cols[(size_t)c].p : nullptr; }
// Device times of a published query (plan.hpp) are read from the events on first use - as long as no later query has
// recorded them again (then the figure is gone: -1).
static void resolve_times(imm3_result* r) {
    if (r->device_ms >= 0 || !r->db || r->timing_seq == 0 || r->db->ev_seq != r->timing_seq) return;
    imm3_db* db = r->db;
    if (use_device(db)) return;
    float f = 0;
    if (cudaEventSynchronize(db->ev1) != cudaSuccess || cudaEventElapsedTime(&f, db->ev0, db->ev1) != cudaSuccess) return;
    r->device_ms = f;
    r->stage_ms[0] = f;
}
double imm3_result_device_ms(const imm3_result* r) {
    if (!r) return -1.0;
    resolve_times(const_cast<imm3_result*>(r));
    return r->device_ms;
}
int imm3_result_kernel_launches(const imm3_result* r) { return r ? r->launches : IMM3_ERR_INVALID_ARG; }
double imm3_result_host_us(const imm3_result* r, int phase) { return (r && phase >= 0 && phase < 5) ? r->host_us[phase] : -1.0; }
double imm3_result_stage_ms(const imm3_result* r, int stage) {
    if (!r || stage < 0 || stage >= 2) return -1.0;
    resolve_times(const_cast<imm3_result*>(r));
    return r->stage_ms[stage];
}
int64_t imm3_result_algorithmic_bytes(const imm3_result* r) { return r ? r->alg_bytes : IMM3_ERR_INVALID_ARG; }

// Row.toString = xs.mkString("Row(", ",", ")")  (Record.scala:13)
int imm3_result_format_row(const imm3_result* r, int64_t row, char* buf, size_t buflen) {
    if (!r || !buf || row < 0 || row >= r->fetched) return fail(IMM3_ERR_INVALID_ARG, "imm3_result_format_row: bad arguments");
    std::string s = "Row(";
    const int c0 = r->agg_first_col >= 0 ? r->agg_first_col : 0;  // aggregate rows print the aggregators' repr only (ProjectAggregateQueue.scala:48-50)
    for (int c = c0; c < r->ncols; c++) {
        if (c > c0) s += ",";
        const uint8_t* p = (const uint8_t*)r->h_cols[(size_t)c].p + (size_t)row * (size_t)r->widths[(size_t)c];
        if (r->types[(size_t)c] == IMM3_COL_COUNT) {
            int64_t v;
            std::memcpy(&v, p, 8);
            s += std::to_string(v);  // Long.toString
        } else if (r->types[(size_t)c] == IMM3_COL_DOUBLE) {
            double v;
            std::memcpy(&v, p, 8);
            s += java_double_to_string_integral(v);
        } else if (r->types[(size_t)c] == IMM3_COL_INT) {
            int32_t v;
            std::memcpy(&v, p, 4);
            s += std::to_string(v);
        } else if (r->types[(size_t)c] == IMM3_COL_TINYINT) {
            s += std::to_string((int)(int8_t)p[0]);
        } else {
            s.append((const char*)p, (size_t)r->widths[(size_t)c]);
        }
    }
    s += ")";
    if (s.size() + 1 > buflen) return fail(IMM3_ERR_INVALID_ARG, "imm3_result_format_row: buffer too small (%zu needed)", s.size() + 1);
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}

int imm3_result_free(imm3_result* r) {
    if (!r) return 0;
    if (r->pending >= 0) cudaEventSynchronize(r->copied);  // never hand buffers back while a copy is reading them
    if (r->copied) cudaEventDestroy(r->copied);
    for (auto& b : r->d_cols) r->db->dev_pool.release(b);
    for (auto& b : r->h_cols) r->db->host_pool.release(b);
    r->db->live_results--;
    delete r;
    return 0;
}

int imm3_filter_bitmap(imm3_db* db, const char* table, const imm3_pred* preds, int npreds, const uint32_t** words,
                       int64_t* nwords, int64_t* nselected) {
    if (!db || !words || !nwords || !nselected) return fail(IMM3_ERR_INVALID_ARG, "imm3_filter_bitmap: NULL argument");
    Prepared pr;
    int rc = prepare(db, table, preds, npreds, nullptr, 0, 0, &pr);
    if (rc) return rc;
    if ((rc = use_device(db))) return rc;
    TableStore& t = *pr.table;
    const size_t need_words = (size_t)((t.nrows + kDenseMaxTileRows - 1) / kDenseMaxTileRows) * (kDenseMaxTileRows / 32) + 2;
    if ((rc = ensure_buf(&db->d_bitmap, need_words * 4))) return rc;
    if (db->h_bitmap.cap < need_words * 4) {
        if (db->h_bitmap.p) cudaFreeHost(db->h_bitmap.p);
        db->h_bitmap = Buf();
        CUDA_TRY(cudaMallocHost(&db->h_bitmap.p, need_words * 4));
        db->h_bitmap.cap = need_words * 4;
    }
    CUDA_TRY(cudaMemsetAsync(db->d_bitmap.p, 0, need_words * 4, db->stream));
    int64_t total = 0;
    if (!pr.lp.always_empty && t.nrows > 0) {
        pr.for_bitmap = true;
        if ((rc = fill_scan_plan(db, &pr))) return rc;
        pr.sp.bitmap = (uint32_t*)db->d_bitmap.p;
        pr.sp.limit = INT64_MAX;
        double ms;
        int launches;
        if ((rc = run_scan(db, &pr, &ms, &total, &launches))) return rc;
    }
    const size_t out_words = (size_t)((t.nrows + 31) / 32);
    if (out_words) CUDA_TRY(cudaMemcpyAsync(db->h_bitmap.p, db->d_bitmap.p, out_words * 4, cudaMemcpyDeviceToHost, db->stream));
    CUDA_TRY(cudaStreamSynchronize(db->stream));
    *words = (const uint32_t*)db->h_bitmap.p;
    *nwords = (int64_t)out_words;
    *nselected = total;
    return 0;
}

}  // extern "C"
