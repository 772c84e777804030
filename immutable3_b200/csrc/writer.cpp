// writer.cpp — writer side of the table format:
//   SegmentWriter            Segment.scala:70-152   (lazy block flush, blockOffset bookkeeping)
//   LoaderCli main loop      LoaderCli.scala:113-154 (roll to a new segment when remaining == 0)
//   TableIO.clear / store    Table.scala:50-66
//   PFORCodecInt.encode      PFORCodec.scala:17-28  (JavaFastPFOR IntegratedIntCompressor 0.1.10;
//                            the library is not in the reference tree: restated from its published
//                            algorithm, byte compatibility with the real jar is UNPINNED)
// plus the deterministic synthetic tables of BASELINE.md.  Host-only code, no CUDA.
#include <dirent.h>
#include <sys/stat.h>

#include <cerrno>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "common.hpp"

namespace imm3 {

// ---------------------------------------------------------------------------------------------
// Sorted-integer codec, encoder.  Bit-stream formulation: each 32-value mini-block of width b is
// the little-endian bit string d0 | d1<<b | d2<<2b ... cut into 32-bit words.
// ---------------------------------------------------------------------------------------------
namespace {

inline int bit_length(uint32_t x) { return x ? 32 - __builtin_clz(x) : 0; }

// Width of one mini-block = bit length of the OR of its 32 deltas (Util.maxdiffbits).
inline int miniblock_width(uint32_t base, const int32_t* v) {
    uint32_t acc = 0, prev = base;
    for (int i = 0; i < 32; i++) {
        acc |= (uint32_t)v[i] - prev;
        prev = (uint32_t)v[i];
    }
    return bit_length(acc);
}

// Appends the packed mini-block to `out` (IntegratedBitPacking.integratedpack<b>).
inline void pack_miniblock(uint32_t base, const int32_t* v, int b, std::vector<uint32_t>& out) {
    if (b == 0) return;  // all 32 values equal the running base
    if (b == 32) {       // integratedpack32 copies the VALUES, not the deltas
        for (int i = 0; i < 32; i++) out.push_back((uint32_t)v[i]);
        return;
    }
    uint64_t acc = 0;
    int have = 0;
    uint32_t prev = base;
    for (int i = 0; i < 32; i++) {
        uint64_t d = (uint32_t)v[i] - prev;
        prev = (uint32_t)v[i];
        acc |= d << have;
        have += b;
        if (have >= 32) {
            out.push_back((uint32_t)acc);
            acc >>= 32;
            have -= 32;
        }
    }
    // 32*b bits is a whole number of words, nothing is left over.
}

// IntegratedIntCompressor.compress: [n] ++ binary-packed prefix ++ var-byte remainder.
void iic_compress(const int32_t* in, int32_t n, std::vector<uint32_t>& out) {
    out.clear();
    out.push_back((uint32_t)n);
    uint32_t base = 0;  // initvalue = 0 for every block, so blocks decode independently
    const int32_t packed = n - n % 32;
    int32_t s = 0;
    while (s + 128 <= packed) {  // super-block: one header word, four mini-blocks
        int b[4];
        uint32_t bases[4];
        uint32_t run = base;
        for (int m = 0; m < 4; m++) {
            bases[m] = run;
            b[m] = miniblock_width(run, in + s + 32 * m);
            run = (uint32_t)in[s + 32 * m + 31];
        }
        out.push_back(((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | (uint32_t)b[3]);
        for (int m = 0; m < 4; m++) pack_miniblock(bases[m], in + s + 32 * m, b[m], out);
        base = run;
        s += 128;
    }
    while (s < packed) {  // left-over mini-blocks carry their own header word
        int b = miniblock_width(base, in + s);
        out.push_back((uint32_t)b);
        pack_miniblock(base, in + s, b, out);
        base = (uint32_t)in[s + 31];
        s += 32;
    }
    if (n > packed) {  // IntegratedVariableByte: 7-bit groups, low first, LAST byte has bit 7 set
        std::vector<uint8_t> bytes;
        for (int32_t k = packed; k < n; k++) {
            uint32_t d = (uint32_t)in[k] - base;
            base = (uint32_t)in[k];
            while (d >= 0x80u) {
                bytes.push_back((uint8_t)(d & 0x7Fu));
                d >>= 7;
            }
            bytes.push_back((uint8_t)(d | 0x80u));
        }
        while (bytes.size() % 4) bytes.push_back(0);
        for (size_t i = 0; i < bytes.size(); i += 4)  // ByteOrder.LITTLE_ENDIAN byte buffer viewed as ints
            out.push_back((uint32_t)bytes[i] | ((uint32_t)bytes[i + 1] << 8) | ((uint32_t)bytes[i + 2] << 16) |
                          ((uint32_t)bytes[i + 3] << 24));
    }
}

// PFORCodecInt.encode: putInt is big-endian; the whole (4*words + 8)-byte backing array is emitted.
void pfor_encode_bytes(const int32_t* in, int32_t n, std::vector<uint8_t>& out, std::vector<uint32_t>& scratch) {
    iic_compress(in, n, scratch);
    out.resize(scratch.size() * 4 + 8);
    for (size_t i = 0; i < scratch.size(); i++) {
        uint32_t w = scratch[i];
        out[4 * i] = (uint8_t)(w >> 24);
        out[4 * i + 1] = (uint8_t)(w >> 16);
        out[4 * i + 2] = (uint8_t)(w >> 8);
        out[4 * i + 3] = (uint8_t)w;
    }
    std::memset(out.data() + scratch.size() * 4, 0, 8);
}

int mkdirs(const std::string& path) {
    std::string cur;
    for (size_t i = 0; i <= path.size(); i++) {
        if (i == path.size() || path[i] == '/') {
            if (!cur.empty() && mkdir(cur.c_str(), 0777) && errno != EEXIST)
                return fail(IMM3_ERR_IO, "mkdir %s: %s", cur.c_str(), strerror(errno));
        }
        if (i < path.size()) cur.push_back(path[i]);
    }
    return 0;
}

// One SegmentWriter (Segment.scala:70-152) that also owns the LoaderCli roll to `newSegment()`.
class ColumnWriter {
  public:
    ColumnWriter(std::string dir, ColumnMeta col, int block_size, int segment_size, int first_id)
        : dir_(std::move(dir)), col_(std::move(col)), B_(block_size), S_(segment_size), id_(first_id) {}
    ~ColumnWriter() {
        if (f_) fclose(f_);
    }

    int open_segment() {
        std::string p = dir_ + "/" + col_.name + "_" + std::to_string(id_) + ".dat";
        f_ = fopen(p.c_str(), "wb");  // RandomAccessFile(..., "rw") + setLength(0)
        if (!f_) return fail(IMM3_ERR_IO, "open %s: %s", p.c_str(), strerror(errno));
        offsets_.assign(1, 0);  // blockBufferOffsets += 0
        buf_.clear();
        buf_.reserve((size_t)B_ * col_.width);
        records_ = 0;
        return 0;
    }

    // LoaderCli.scala:142-148 + SegmentWriter.write (Segment.scala:99-112), for `n` values at once.
    int append(const uint8_t* cells, int64_t n) {
        const size_t w = (size_t)col_.width;
        while (n > 0) {
            if (remaining() <= 0) {  // seg.close(); segs(segName) = seg.newSegment()
                int rc = close_segment();
                if (rc) return rc;
                id_++;
                if ((rc = open_segment())) return rc;
            }
            if (records_ == B_) {  // the (B+1)-th value flushes the full block, then is buffered
                int rc = flush();
                if (rc) return rc;
                buf_.insert(buf_.end(), cells, cells + w);
                records_ = 1;
                cells += w;
                n--;
                continue;  // re-check remaining: the flush may have filled the segment
            }
            int64_t take = B_ - records_ < n ? B_ - records_ : n;
            buf_.insert(buf_.end(), cells, cells + (size_t)take * w);
            records_ += (int)take;
            cells += (size_t)take * w;
            n -= take;
        }
        return 0;
    }

    // SegmentWriter.flush, Segment.scala:114-128
    int flush() {
        const uint8_t* data = buf_.data();
        size_t len = buf_.size();
        if (col_.codec == IMM3_CODEC_PFOR_INT) {
            // bytes -> Int via bytesToValue (little-endian) -> IntegratedIntCompressor
            ints_.resize(len / 4);
            std::memcpy(ints_.data(), data, ints_.size() * 4);  // host is little-endian
            pfor_encode_bytes(ints_.data(), (int32_t)ints_.size(), enc_, words_);
            data = enc_.data();
            len = enc_.size();
        }
        if (len && fwrite(data, 1, len, f_) != len) return fail(IMM3_ERR_IO, "short write on %s_%d.dat", col_.name.c_str(), id_);
        if ((int64_t)offsets_.back() + (int64_t)len > INT32_MAX)
            return fail(IMM3_ERR_UNSUPPORTED, "segment %s_%d exceeds 2 GiB (block offsets are Int, Segment.scala:33)", col_.name.c_str(), id_);
        offsets_.push_back(offsets_.back() + (int32_t)len);
        buf_.clear();
        records_ = 0;
        return 0;
    }

    // SegmentWriter.close, Segment.scala:144-151
    int close_segment() {
        if (!f_) return 0;
        if (!buf_.empty()) {
            int rc = flush();
            if (rc) return rc;
        }
        fclose(f_);
        f_ = nullptr;
        std::string p = dir_ + "/" + col_.name + "_" + std::to_string(id_) + ".meta";
        FILE* m = fopen(p.c_str(), "wb");
        if (!m) return fail(IMM3_ERR_IO, "open %s: %s", p.c_str(), strerror(errno));
        std::string s = "{\"blockOffset\":[";  // SegmentMeta.toJsonValue, Segment.scala:41-45
        for (size_t i = 0; i < offsets_.size(); i++) {
            if (i) s += ",";
            s += std::to_string(offsets_[i]);
        }
        s += "]}";
        bool ok = fwrite(s.data(), 1, s.size(), m) == s.size();
        fclose(m);
        return ok ? 0 : fail(IMM3_ERR_IO, "short write on %s", p.c_str());
    }

    int remaining() const { return S_ - ((int)offsets_.size() - 1); }  // Segment.scala:139-142
    const ColumnMeta& col() const { return col_; }

  private:
    std::string dir_;
    ColumnMeta col_;
    int B_, S_, id_;
    FILE* f_ = nullptr;
    std::vector<int32_t> offsets_;
    std::vector<uint8_t> buf_, enc_;
    std::vector<int32_t> ints_;
    std::vector<uint32_t> words_;
    int records_ = 0;
};

}  // namespace

struct Writer {
    std::string table_dir;
    std::vector<std::unique_ptr<ColumnWriter>> cols;
};

static int writer_open(const char* data_dir, const char* table, const char* const* specs, int ncols, int B, int S,
                       int first_id, int write_meta, Writer** out) {
    if (!data_dir || !table || !specs || ncols <= 0 || B <= 0 || S <= 0 || first_id < 0)
        return fail(IMM3_ERR_INVALID_ARG, "imm3_writer_open: bad arguments");
    TableMeta tm;
    tm.name = table;
    tm.block_size = B;
    for (int i = 0; i < ncols; i++) {
        ColumnMeta c;
        int rc = parse_col_spec(specs[i], &c);
        if (rc) return rc;
        tm.cols.push_back(c);
    }
    std::unique_ptr<Writer> w(new Writer());
    w->table_dir = std::string(data_dir) + "/" + table;
    int rc = mkdirs(w->table_dir);  // Files.createDirectories, Table.scala:52
    if (rc) return rc;
    if (write_meta) {
        // TableIO.clear: delete the regular files of the table directory (Table.scala:61-66)
        if (DIR* d = opendir(w->table_dir.c_str())) {
            while (struct dirent* e = readdir(d)) {
                std::string p = w->table_dir + "/" + e->d_name;
                struct stat st;
                if (!stat(p.c_str(), &st) && S_ISREG(st.st_mode)) remove(p.c_str());
            }
            closedir(d);
        }
        std::string p = w->table_dir + "/_table.meta";
        FILE* f = fopen(p.c_str(), "wb");
        if (!f) return fail(IMM3_ERR_IO, "open %s: %s", p.c_str(), strerror(errno));
        std::string s = table_meta_json(tm);
        bool ok = fwrite(s.data(), 1, s.size(), f) == s.size();
        fclose(f);
        if (!ok) return fail(IMM3_ERR_IO, "short write on %s", p.c_str());
    }
    for (auto& c : tm.cols) {
        w->cols.emplace_back(new ColumnWriter(w->table_dir, c, B, S, first_id));
        if ((rc = w->cols.back()->open_segment())) return rc;
    }
    *out = w.release();
    return 0;
}

// stringToValue of the three DataTypes (DataType.scala:39,59,68), stricter on STRING length.
static int parse_cell(const ColumnMeta& c, const std::string& s, std::vector<uint8_t>& cell) {
    cell.resize((size_t)c.width);
    if (c.ctype == IMM3_COL_STRING) {
        if ((int)s.size() != c.width)
            return fail(IMM3_ERR_INVALID_ARG, "column %s: value '%s' is %zu bytes, DENSE_STRING size is %d "
                        "(the reference would silently misalign the block)", c.name.c_str(), s.c_str(), s.size(), c.width);
        std::memcpy(cell.data(), s.data(), (size_t)c.width);
        return 0;
    }
    // Java Integer.parseInt / Byte.parseByte: optional sign, decimal digits only, range checked.
    if (s.empty()) return fail(IMM3_ERR_INVALID_ARG, "column %s: empty numeric value", c.name.c_str());
    size_t i = (s[0] == '-' || s[0] == '+') ? 1 : 0;
    if (i == s.size()) return fail(IMM3_ERR_INVALID_ARG, "column %s: bad number '%s'", c.name.c_str(), s.c_str());
    int64_t v = 0;
    for (; i < s.size(); i++) {
        if (s[i] < '0' || s[i] > '9') return fail(IMM3_ERR_INVALID_ARG, "column %s: bad number '%s'", c.name.c_str(), s.c_str());
        v = v * 10 + (s[i] - '0');
        if (v > (int64_t)INT32_MAX + 1) return fail(IMM3_ERR_INVALID_ARG, "column %s: '%s' out of range", c.name.c_str(), s.c_str());
    }
    if (s[0] == '-') v = -v;
    if (c.ctype == IMM3_COL_TINYINT) {
        if (v < -128 || v > 127) return fail(IMM3_ERR_INVALID_ARG, "column %s: '%s' out of TINYINT range", c.name.c_str(), s.c_str());
        cell[0] = (uint8_t)(int8_t)v;
    } else {
        if (v < INT32_MIN || v > INT32_MAX) return fail(IMM3_ERR_INVALID_ARG, "column %s: '%s' out of INT range", c.name.c_str(), s.c_str());
        uint32_t u = (uint32_t)(int32_t)v;  // IntType.valueToBytes: low byte first (DataType.scala:40-47)
        cell[0] = (uint8_t)u; cell[1] = (uint8_t)(u >> 8); cell[2] = (uint8_t)(u >> 16); cell[3] = (uint8_t)(u >> 24);
    }
    return 0;
}

static std::string trim(const std::string& s) {  // java String.trim: strip chars <= ' '
    size_t a = 0, b = s.size();
    while (a < b && (unsigned char)s[a] <= ' ') a++;
    while (b > a && (unsigned char)s[b - 1] <= ' ') b--;
    return s.substr(a, b - a);
}

static int writer_append_csv_line(Writer* w, const char* line) {
    // line.split(",").map(_.trim); Java's split drops trailing empty strings (LoaderCli.scala:136)
    std::vector<std::string> vals;
    {
        std::string cur;
        for (const char* p = line; *p && *p != '\n' && *p != '\r'; p++) {
            if (*p == ',') { vals.push_back(cur); cur.clear(); } else cur.push_back(*p);
        }
        vals.push_back(cur);
        while (!vals.empty() && vals.back().empty()) vals.pop_back();
    }
    if (vals.size() > w->cols.size())
        return fail(IMM3_ERR_INVALID_ARG, "CSV line has %zu fields, table has %zu columns", vals.size(), w->cols.size());
    std::vector<uint8_t> cell;
    for (size_t i = 0; i < vals.size(); i++) {  // for (idx <- 0 until vals.size): column idx of --cols
        int rc = parse_cell(w->cols[i]->col(), trim(vals[i]), cell);
        if (rc) return rc;
        if ((rc = w->cols[i]->append(cell.data(), 1))) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Synthetic tables (BASELINE.md "Configurations to measure").
// ---------------------------------------------------------------------------------------------
static const char kStates[51][3] = {
    "AL", "AK", "AZ", "AR", "CA", "CO", "CT", "DE", "FL", "GA", "HI", "ID", "IL", "IN", "IA", "KS", "KY",
    "LA", "ME", "MD", "MA", "MI", "MN", "MS", "MO", "MT", "NE", "NV", "NH", "NJ", "NM", "NY", "NC", "ND",
    "OH", "OK", "OR", "PA", "RI", "SC", "SD", "TN", "TX", "UT", "VT", "VA", "WA", "WV", "WI", "WY", "DC"};

static inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static const uint64_t kSeed = 42;

static inline void synth_row(int64_t row, int32_t* id, int8_t* age, char* state) {
    *id = (int32_t)row;  // sorted, so it suits the sorted-integer codec
    *age = (int8_t)((splitmix64(((kSeed ^ 1) << 32) ^ (uint64_t)row) >> 32) % 100u);
    const char* s = kStates[(splitmix64(((kSeed ^ 2) << 32) ^ (uint64_t)row) >> 32) % 51u];
    state[0] = s[0];
    state[1] = s[1];
}

}  // namespace imm3

using namespace imm3;

extern "C" {

struct imm3_writer {
    Writer* w;
};

int imm3_writer_open(const char* data_dir, const char* table, const char* const* col_specs, int ncols,
                     int32_t block_size, int32_t segment_size, int32_t first_segment_id, int write_table_meta,
                     imm3_writer** out) {
    if (!out) return fail(IMM3_ERR_INVALID_ARG, "imm3_writer_open: out is NULL");
    Writer* w = nullptr;
    int rc = writer_open(data_dir, table, col_specs, ncols, block_size, segment_size, first_segment_id,
                         write_table_meta, &w);
    if (rc) return rc;
    *out = new imm3_writer{w};
    return 0;
}

int imm3_writer_append(imm3_writer* h, const void* const* col_data, int64_t nrows) {
    if (!h || !h->w || !col_data || nrows < 0) return fail(IMM3_ERR_INVALID_ARG, "imm3_writer_append: bad arguments");
    for (size_t i = 0; i < h->w->cols.size(); i++) {
        if (!col_data[i]) return fail(IMM3_ERR_INVALID_ARG, "imm3_writer_append: column %zu data is NULL", i);
        int rc = h->w->cols[i]->append((const uint8_t*)col_data[i], nrows);  // typed arrays are already LE cells
        if (rc) return rc;
    }
    return 0;
}

int imm3_writer_append_csv_line(imm3_writer* h, const char* line) {
    if (!h || !h->w || !line) return fail(IMM3_ERR_INVALID_ARG, "imm3_writer_append_csv_line: bad arguments");
    return writer_append_csv_line(h->w, line);
}

int imm3_writer_close(imm3_writer* h) {
    if (!h) return 0;
    int rc = 0;
    if (h->w) {
        for (auto& c : h->w->cols) {  // for ((_, seg) <- segs) seg.close()  (LoaderCli.scala:152-154)
            int r = c->close_segment();
            if (r && !rc) rc = r;
        }
        delete h->w;
    }
    delete h;
    return rc;
}

int imm3_load_csv(const char* data_dir, const char* table, const char* const* col_specs, int ncols,
                  int32_t block_size, int32_t segment_size, const char* csv_path) {
    if (!csv_path) return fail(IMM3_ERR_INVALID_ARG, "imm3_load_csv: csv_path is NULL");
    FILE* f = fopen(csv_path, "rb");
    if (!f) return fail(IMM3_ERR_IO, "open %s: %s", csv_path, strerror(errno));
    imm3_writer* w = nullptr;
    int rc = imm3_writer_open(data_dir, table, col_specs, ncols, block_size, segment_size, 0, 1, &w);
    if (rc) { fclose(f); return rc; }
    char* line = nullptr;
    size_t cap = 0;
    bool first = true;
    while (getline(&line, &cap, f) >= 0) {
        if (first) { first = false; continue; }  // val first = lines.next(): header discarded
        if ((rc = imm3_writer_append_csv_line(w, line))) break;
    }
    free(line);
    fclose(f);
    int rc2 = imm3_writer_close(w);
    return rc ? rc : rc2;
}

int64_t imm3_pfor_encode(const int32_t* values, int32_t n, uint8_t* out, int64_t out_cap) {
    if (n < 0 || (n > 0 && !values)) return fail(IMM3_ERR_INVALID_ARG, "imm3_pfor_encode: bad arguments");
    std::vector<uint8_t> enc;
    std::vector<uint32_t> scratch;
    pfor_encode_bytes(values, n, enc, scratch);
    if (!out) return (int64_t)enc.size();
    if (out_cap < (int64_t)enc.size()) return fail(IMM3_ERR_INVALID_ARG, "imm3_pfor_encode: need %zu bytes", enc.size());
    std::memcpy(out, enc.data(), enc.size());
    return (int64_t)enc.size();
}

void imm3_synth_row(int64_t row, int32_t* id, int8_t* age, char state[2]) { synth_row(row, id, age, state); }

int imm3_synth_write(const char* data_dir, const char* table, int64_t nrows, int32_t block_size, int32_t segment_size,
                     int32_t id_codec, int32_t seg_id_begin, int32_t seg_id_end, int write_table_meta) {
    if (nrows < 0 || block_size <= 0 || segment_size <= 0) return fail(IMM3_ERR_INVALID_ARG, "imm3_synth_write: bad sizes");
    if (id_codec != IMM3_CODEC_DENSE_INT && id_codec != IMM3_CODEC_PFOR_INT)
        return fail(IMM3_ERR_INVALID_ARG, "imm3_synth_write: id_codec must be DENSE_INT or PFOR_INT");
    const int64_t rows_per_seg = (int64_t)block_size * segment_size + 1;
    const int64_t nseg = (nrows + rows_per_seg - 1) / rows_per_seg;
    if (seg_id_end < 0 || seg_id_end > nseg) seg_id_end = (int32_t)nseg;
    if (seg_id_begin < 0) seg_id_begin = 0;
    const char* specs[3] = {id_codec == IMM3_CODEC_PFOR_INT ? "id:PFOR_INT" : "id:DENSE_INT", "state:DENSE_STRING:size=2",
                            "age:DENSE_TINYINT"};
    if (write_table_meta) {  // clear + _table.meta only (rank 0 of a cooperative write)
        imm3_writer* w = nullptr;
        int rc = imm3_writer_open(data_dir, table, specs, 3, block_size, segment_size, 0, 1, &w);
        if (rc) return rc;
        // drop the empty segment 0 this created; real segments are written below
        imm3_writer_close(w);
        std::string base = std::string(data_dir) + "/" + table + "/";
        for (const char* c : {"id", "state", "age"}) {
            remove((base + c + "_0.dat").c_str());
            remove((base + c + "_0.meta").c_str());
        }
    }
    // Segments are independent files: written by a small pool of threads (a 1 B-row table is 977 segments).
    const int nsegs = seg_id_end - seg_id_begin;
    if (nsegs <= 0) return 0;
    int nthreads = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("IMM3_WRITER_THREADS")) nthreads = atoi(e);
    nthreads = std::max(1, std::min(nthreads, nsegs));
    std::atomic<int> next(seg_id_begin);
    std::atomic<int> first_rc(0);
    std::vector<std::string> errors((size_t)nthreads);
    auto work = [&](int tid) {
        const int64_t chunk = 1 << 18;
        std::vector<int32_t> ids((size_t)chunk);
        std::vector<int8_t> ages((size_t)chunk);
        std::vector<char> states((size_t)chunk * 2);
        for (;;) {
            const int seg = next.fetch_add(1);
            if (seg >= seg_id_end || first_rc.load()) return;
            imm3_writer* w = nullptr;
            int rc = imm3_writer_open(data_dir, table, specs, 3, block_size, segment_size, seg, 0, &w);
            int64_t r0 = seg * rows_per_seg, r1 = r0 + rows_per_seg < nrows ? r0 + rows_per_seg : nrows;
            for (int64_t r = r0; !rc && r < r1; r += chunk) {
                int64_t n = r1 - r < chunk ? r1 - r : chunk;
                for (int64_t i = 0; i < n; i++) synth_row(r + i, &ids[(size_t)i], &ages[(size_t)i], &states[(size_t)i * 2]);
                const void* cols[3] = {ids.data(), states.data(), ages.data()};
                rc = imm3_writer_append(w, cols, n);
            }
            if (w) {
                int rc2 = imm3_writer_close(w);
                if (!rc) rc = rc2;
            }
            if (rc) {
                errors[(size_t)tid] = last_error();  // (the message is thread-local)
                int zero = 0;
                first_rc.compare_exchange_strong(zero, rc);
                return;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < nthreads; i++) pool.emplace_back(work, i);
    work(0);
    for (auto& th : pool) th.join();
    if (int rc = first_rc.load()) {
        for (auto& e : errors)
            if (!e.empty()) return fail(rc, "%s", e.c_str());
        return fail(rc, "imm3_synth_write failed");
    }
    return 0;
}

}  // extern "C"
