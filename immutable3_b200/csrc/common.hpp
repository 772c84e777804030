// common.hpp — shared host-side declarations of libimm3gpu (error channel, metadata model).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include "../../include/imm3.h"

namespace imm3 {

// Thread-local message behind imm3_last_error().
void set_error(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
const char* last_error();

// Host threads for the embarrassingly parallel parts of imm3_open (one .meta / .dat file per task): hardware
// concurrency capped at 16, IMM3_IO_THREADS overrides.
int io_threads();
// fn(i) for i in [0, n) on up to `nthreads` threads (dynamic assignment); returns the first non-zero status, with that
// thread's error message re-published on the calling thread.
int parallel_for(int64_t n, int nthreads, const std::function<int(int64_t)>& fn);

// JVM narrowing used by the predicate constants (Select.scala:65,73,103,111,141,149).
int32_t d2i(double d);  // Scala Double.toInt
int8_t d2b(double d);   // Scala Double.toByte

// ---------------------------------------------------------------------------------------------
// Metadata model: Table / Column (Table.scala:9, Column.scala:18) + segment files
// (SegmentManager.scala:38-79) + SegmentMeta.blockOffsets (Segment.scala:33).
// ---------------------------------------------------------------------------------------------
struct ColumnMeta {
    std::string name;
    int ctype = 0;  // imm3_column_type
    int codec = 0;  // imm3_codec
    int width = 0;  // decoded bytes per value
    std::vector<std::pair<std::string, std::string>> attrs;  // dtypeAttrs, insertion order
};

struct TableMeta {
    std::string name;
    int block_size = 0;
    std::vector<ColumnMeta> cols;
};

struct SegmentFile {
    std::string path;       // <dir>/<table>/<col>_<id>.dat
    int file_id = 0;        // numeric id in the file name
    int64_t nbytes = 0;     // bytes covered by the block offsets (what the reference can ever read)
    std::vector<int32_t> offsets;  // blockOffset array
};

// "name:CODEC[:k=v;k=v]" -> ColumnMeta (LoaderCliParser.parseCol, LoaderCli.scala:66-80 + Column.make)
int parse_col_spec(const char* spec, ColumnMeta* out);
std::string table_meta_json(const TableMeta& t);  // TableIO.toJsonValue, Table.scala:27-35
int parse_table_meta(const std::string& json, const std::string& origin, TableMeta* out);
int parse_segment_meta(const std::string& json, const std::string& origin, std::vector<int32_t>* offsets);
int read_text_file(const std::string& path, std::string* out);
// listFiles().filter(startsWith(col_) && endsWith(suffix)).sortBy(getName)  (SegmentManager.scala:38-42)
int list_segment_files(const std::string& table_dir, const std::string& col, const char* suffix,
                       std::vector<std::string>* names);

}  // namespace imm3
