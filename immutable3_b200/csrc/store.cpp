// store.cpp — table discovery and validation (host only, no CUDA):
//   SegmentManager.getTables / getSegmentFiles / getSegmentMetaFiles   SegmentManager.scala:27-79
//   Segment.BlockIterator framing                                       Segment.scala:154-181
// What the reference leaves undefined is rejected up front with IMM3_ERR_BAD_FORMAT instead of
// throwing (or silently misaligning) in the middle of a scan:
//   * a column with a different number of .dat and .meta files, or columns with different segment
//     counts (the reference indexes every column's list with the first column's count);
//   * block offsets that do not start at 0, decrease, or run past the file;
//   * a dense block whose byte length is not a multiple of the value width (the reference's decode
//     loop would fabricate one extra value from stale bytes, DenseCodec.scala:41-44);
//   * columns whose blocks hold different numbers of rows (batch size is taken from the first used
//     column, Scan.scala:55, and the other vectors are then indexed out of bounds).
#include "store.hpp"

#include <dirent.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstring>

namespace imm3 {

int FileMap::open(const std::string& path, size_t need) {
    close();
    if (need == 0) return 0;
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return fail(IMM3_ERR_IO, "open %s: %s", path.c_str(), strerror(errno));
    struct stat st;
    if (fstat(fd, &st)) { ::close(fd); return fail(IMM3_ERR_IO, "stat %s: %s", path.c_str(), strerror(errno)); }
    if ((size_t)st.st_size < need) { ::close(fd); return fail(IMM3_ERR_BAD_FORMAT, "%s: file has %lld bytes, block offsets need %zu", path.c_str(), (long long)st.st_size, need); }
    void* p = mmap(nullptr, need, PROT_READ, MAP_PRIVATE, fd, 0);  // FileChannel.map READ_ONLY, SegmentManager.scala:81-87
    ::close(fd);
    if (p == MAP_FAILED) return fail(IMM3_ERR_IO, "mmap %s: %s", path.c_str(), strerror(errno));
    data = (const uint8_t*)p;
    len = need;
    return 0;
}
void FileMap::close() {
    if (data) munmap((void*)data, len);
    data = nullptr;
    len = 0;
}

static inline uint32_t be32(const uint8_t* p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

int pfor_validate_block(const uint8_t* bytes, int64_t nbytes, int32_t* n_out) {
    if (nbytes < 12 || (nbytes & 3)) return fail(IMM3_ERR_BAD_FORMAT, "PFOR_INT block of %lld bytes", (long long)nbytes);
    const int64_t nw = nbytes / 4 - 2;  // the encoder appends 8 zero bytes (PFORCodec.scala:20)
    const int64_t n64 = (int64_t)be32(bytes);
    if (n64 > (1 << 28)) return fail(IMM3_ERR_BAD_FORMAT, "PFOR_INT block claims %lld values", (long long)n64);
    const int32_t n = (int32_t)n64;
    const int32_t packed = n - n % 32;
    int64_t ip = 1;
    int32_t s = 0;
    for (; s + 128 <= packed; s += 128) {
        if (ip >= nw) return fail(IMM3_ERR_BAD_FORMAT, "PFOR_INT block truncated (header)");
        uint32_t h = be32(bytes + 4 * ip++);
        for (int q = 0; q < 4; q++) {
            uint32_t b = (h >> (24 - 8 * q)) & 0xFF;
            if (b > 32) return fail(IMM3_ERR_BAD_FORMAT, "PFOR_INT bit width %u", b);
            ip += b;
        }
    }
    for (; s < packed; s += 32) {
        if (ip >= nw) return fail(IMM3_ERR_BAD_FORMAT, "PFOR_INT block truncated (header)");
        uint32_t b = be32(bytes + 4 * ip++);
        if (b > 32) return fail(IMM3_ERR_BAD_FORMAT, "PFOR_INT bit width %u", b);
        ip += b;
    }
    if (ip > nw) return fail(IMM3_ERR_BAD_FORMAT, "PFOR_INT block truncated (payload)");
    // var-byte remainder: words hold the byte stream little-endian
    int32_t left = n - packed;
    int64_t byte_pos = 0;
    const int64_t vb_bytes = (nw - ip) * 4;
    while (left > 0) {
        int len = 0;
        for (;;) {
            if (byte_pos >= vb_bytes) return fail(IMM3_ERR_BAD_FORMAT, "PFOR_INT block truncated (var-byte)");
            int64_t w = ip + byte_pos / 4;
            uint32_t c = (be32(bytes + 4 * w) >> (8 * (byte_pos & 3))) & 0xFF;
            byte_pos++;
            if (++len > 5) return fail(IMM3_ERR_BAD_FORMAT, "PFOR_INT var-byte value longer than 5 bytes");
            if (c & 0x80) break;
        }
        left--;
    }
    if (ip + (byte_pos + 3) / 4 != nw)
        return fail(IMM3_ERR_BAD_FORMAT, "PFOR_INT block has %lld words, its contents need %lld", (long long)nw,
                    (long long)(ip + (byte_pos + 3) / 4));
    *n_out = n;
    return 0;
}

static int parse_file_id(const std::string& file, const std::string& col) {
    return atoi(file.c_str() + col.size() + 1);
}

static int load_table(const std::string& data_dir, const std::string& dir_name, int rank, int world, TableStore* t) {
    // TableIO.load(dataDir, parent.getName)  (SegmentManager.scala:32-34, Table.scala:45-48)
    std::string meta_path = data_dir + "/" + dir_name + "/_table.meta", txt;
    int rc = read_text_file(meta_path, &txt);
    if (rc) return rc;
    if ((rc = parse_table_meta(txt, meta_path, &t->meta))) return rc;
    if (t->meta.cols.empty()) return fail(IMM3_ERR_BAD_FORMAT, "%s: table has no columns", meta_path.c_str());
    for (size_t i = 0; i < t->meta.cols.size(); i++)
        for (size_t j = i + 1; j < t->meta.cols.size(); j++)
            if (t->meta.cols[i].name == t->meta.cols[j].name)
                return fail(IMM3_ERR_BAD_FORMAT, "%s: duplicate column %s", meta_path.c_str(), t->meta.cols[i].name.c_str());
    // Segment files are looked up under <dataDir>/<table.name> (SegmentManager.scala:39,62)
    t->dir = data_dir + "/" + t->meta.name;

    t->cols.resize(t->meta.cols.size());
    std::vector<std::vector<std::string>> dats(t->cols.size()), metas(t->cols.size());
    for (size_t c = 0; c < t->cols.size(); c++) {
        t->cols[c].meta = t->meta.cols[c];
        if ((rc = list_segment_files(t->dir, t->cols[c].meta.name, ".dat", &dats[c]))) return rc;
        if ((rc = list_segment_files(t->dir, t->cols[c].meta.name, ".meta", &metas[c]))) return rc;
        if (dats[c].size() != metas[c].size())
            return fail(IMM3_ERR_BAD_FORMAT, "table %s column %s: %zu .dat files but %zu .meta files", t->meta.name.c_str(),
                        t->cols[c].meta.name.c_str(), dats[c].size(), metas[c].size());
        if (dats[c].size() != dats[0].size())
            return fail(IMM3_ERR_BAD_FORMAT, "table %s: column %s has %zu segments, column %s has %zu", t->meta.name.c_str(),
                        t->cols[c].meta.name.c_str(), dats[c].size(), t->cols[0].meta.name.c_str(), dats[0].size());
    }
    t->nsegments = (int)dats[0].size();
    t->file_ids.clear();
    for (auto& f : dats[0]) t->file_ids.push_back(parse_file_id(f, t->cols[0].meta.name));
    shard_range(t->nsegments, rank, world, &t->seg_begin, &t->seg_end);

    // Per column: offsets of the owned segments, block row counts.  One task per (column, segment) file pair on the I/O
    // thread pool (a 1 B-row table is 2931 .meta files and 277 MB of sorted-int blocks to validate); merged in order below.
    const int nown = t->seg_end - t->seg_begin;
    const size_t ncols = t->cols.size();
    struct SegTask {
        SegmentFile sf;
        std::vector<int32_t> rows;
    };
    std::vector<SegTask> tasks(ncols * (size_t)nown);
    rc = parallel_for((int64_t)tasks.size(), io_threads(), [&](int64_t ti) -> int {
        const size_t c = (size_t)ti / (size_t)nown;
        const int s = t->seg_begin + (int)((size_t)ti % (size_t)nown);
        const ColumnStore& col = t->cols[c];
        SegTask& task = tasks[(size_t)ti];
        SegmentFile& sf = task.sf;
        int rc = 0;
        sf.path = t->dir + "/" + dats[c][(size_t)s];
        sf.file_id = parse_file_id(dats[c][(size_t)s], col.meta.name);
        std::string mpath = t->dir + "/" + metas[c][(size_t)s], mtxt;  // i-th sorted .meta pairs with i-th sorted .dat
        if ((rc = read_text_file(mpath, &mtxt))) return rc;
        if ((rc = parse_segment_meta(mtxt, mpath, &sf.offsets))) return rc;
        if (sf.offsets.empty()) return fail(IMM3_ERR_BAD_FORMAT, "%s: empty blockOffset array", mpath.c_str());
        if (sf.offsets[0] != 0) return fail(IMM3_ERR_BAD_FORMAT, "%s: blockOffset must start at 0", mpath.c_str());
        for (size_t b = 1; b < sf.offsets.size(); b++)
            if (sf.offsets[b] < sf.offsets[b - 1]) return fail(IMM3_ERR_BAD_FORMAT, "%s: blockOffset decreases", mpath.c_str());
        sf.nbytes = sf.offsets.back();
        struct stat st;
        if (stat(sf.path.c_str(), &st)) return fail(IMM3_ERR_IO, "stat %s: %s", sf.path.c_str(), strerror(errno));
        if (st.st_size < sf.nbytes)
            return fail(IMM3_ERR_BAD_FORMAT, "%s: %lld bytes on disk, block offsets need %lld", sf.path.c_str(),
                        (long long)st.st_size, (long long)sf.nbytes);
        const size_t nb = sf.offsets.size() - 1;
        task.rows.reserve(nb);
        if (col.meta.codec == IMM3_CODEC_PFOR_INT) {
            FileMap fm;
            if ((rc = fm.open(sf.path, (size_t)sf.nbytes))) return rc;
            for (size_t b = 0; b < nb; b++) {
                int32_t n = 0;
                if ((rc = pfor_validate_block(fm.data + sf.offsets[b], sf.offsets[b + 1] - sf.offsets[b], &n))) {
                    std::string why = last_error();
                    return fail(rc, "%s block %zu: %s", sf.path.c_str(), b, why.c_str());
                }
                task.rows.push_back(n);
            }
        } else {
            for (size_t b = 0; b < nb; b++) {
                int64_t len = (int64_t)sf.offsets[b + 1] - sf.offsets[b];
                if (len % col.meta.width)
                    return fail(IMM3_ERR_BAD_FORMAT, "%s block %zu: %lld bytes is not a multiple of the value width %d",
                                sf.path.c_str(), b, (long long)len, col.meta.width);
                task.rows.push_back((int32_t)(len / col.meta.width));
            }
        }
        return 0;
    });
    if (rc) return rc;
    std::vector<std::vector<int32_t>> block_rows(ncols);
    for (size_t c = 0; c < ncols; c++) {
        ColumnStore& col = t->cols[c];
        for (int s = 0; s < nown; s++) {
            SegTask& task = tasks[c * (size_t)nown + (size_t)s];
            block_rows[c].insert(block_rows[c].end(), task.rows.begin(), task.rows.end());
            col.encoded_bytes += task.sf.nbytes;
            col.segs.push_back(std::move(task.sf));
        }
        if (block_rows[c] != block_rows[0])
            return fail(IMM3_ERR_BAD_FORMAT, "table %s: columns %s and %s do not hold the same rows per block",
                        t->meta.name.c_str(), t->cols[c].meta.name.c_str(), t->cols[0].meta.name.c_str());
    }
    t->nblocks = (int64_t)block_rows[0].size();
    t->row_start.assign(1, 0);
    t->max_block_rows = 0;
    for (int32_t n : block_rows[0]) {
        t->row_start.push_back(t->row_start.back() + (uint64_t)n);
        if (n > t->max_block_rows) t->max_block_rows = n;
    }
    t->nrows = (int64_t)t->row_start.back();

    // PFOR columns: word offsets of every block inside the packed arena.
    for (auto& col : t->cols) {
        if (col.meta.codec != IMM3_CODEC_PFOR_INT) continue;
        col.word_off.clear();
        int64_t base = 0;
        for (auto& sf : col.segs) {
            for (size_t b = 0; b + 1 < sf.offsets.size(); b++) col.word_off.push_back((uint32_t)((base + sf.offsets[b]) / 4));
            base += sf.nbytes;
        }
        col.word_off.push_back((uint32_t)(base / 4));
        col.max_block_words = 0;
        for (size_t b = 0; b + 1 < col.word_off.size(); b++)
            col.max_block_words = std::max<int64_t>(col.max_block_words, (int64_t)col.word_off[b + 1] - (int64_t)col.word_off[b]);
        col.max_tile_bytes = 0;
        const size_t nb = col.word_off.size() - 1;
        for (size_t b = 0; b < nb; b += 8) {
            const uint64_t w0 = col.word_off[b] & ~3u, w8 = col.word_off[std::min(b + 8, nb)];
            col.max_tile_bytes = std::max<int64_t>(col.max_tile_bytes, (int64_t)(((w8 - w0) * 4 + 15) & ~15ull));
        }
        col.max_tile32_bytes = 0;
        for (size_t b = 0; b < nb; b += 32) {
            const uint64_t w0 = col.word_off[b] & ~3u, w1 = col.word_off[std::min(b + 32, nb)];
            col.max_tile32_bytes = std::max<int64_t>(col.max_tile32_bytes, (int64_t)(((w1 - w0) * 4 + 15) & ~15ull));
        }
        if (base / 4 > 0xFFFFFFFFll) return fail(IMM3_ERR_UNSUPPORTED, "PFOR_INT column %s exceeds 16 GiB per GPU", col.meta.name.c_str());
    }
    return 0;
}

int load_tables(const std::string& data_dir, int rank, int world, std::vector<TableStore>* out) {
    DIR* d = opendir(data_dir.c_str());
    if (!d) return fail(IMM3_ERR_IO, "cannot open data dir %s: %s", data_dir.c_str(), strerror(errno));
    std::vector<std::string> dirs;
    while (struct dirent* e = readdir(d)) {
        std::string n = e->d_name;
        if (n == "." || n == "..") continue;
        struct stat st;
        if (!stat((data_dir + "/" + n).c_str(), &st) && S_ISDIR(st.st_mode)) dirs.push_back(n);  // listFiles().filter(_.isDirectory)
    }
    closedir(d);
    out->clear();
    for (auto& n : dirs) {
        out->emplace_back();
        int rc = load_table(data_dir, n, rank, world, &out->back());
        if (rc) return rc;
    }
    return 0;
}

}  // namespace imm3
