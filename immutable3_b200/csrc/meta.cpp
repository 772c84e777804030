// meta.cpp — error channel, JVM narrowing, and the metadata files of a table directory:
//   <dataDir>/<table>/_table.meta   TableIO (Table.scala:26-59), Column JSON (Column.scala:21-38)
//   <dataDir>/<table>/<col>_<id>.meta  SegmentMeta (Segment.scala:33-58)
#include <dirent.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <sstream>
#include <thread>

#include "common.hpp"
#include "json_min.hpp"

namespace imm3 {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
const char* last_error() { return g_err; }

int io_threads() {
    int n = (int)std::thread::hardware_concurrency();
    if (n < 1) n = 1;
    if (n > 16) n = 16;
    if (const char* e = getenv("IMM3_IO_THREADS")) n = std::max(1, atoi(e));
    return n;
}

int parallel_for(int64_t n, int nthreads, const std::function<int(int64_t)>& fn) {
    if (n <= 0) return 0;
    if (nthreads > n) nthreads = (int)n;
    if (nthreads <= 1) {
        for (int64_t i = 0; i < n; i++)
            if (int rc = fn(i)) return rc;
        return 0;
    }
    std::atomic<int64_t> next(0);
    std::atomic<int> first(0);
    std::mutex mu;
    std::string why;
    auto work = [&]() {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n || first.load()) return;
            if (int rc = fn(i)) {
                std::lock_guard<std::mutex> lock(mu);
                if (!first.load()) {
                    why = last_error();  // (the message is thread-local: carry it over to the caller)
                    first.store(rc);
                }
                return;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    if (int rc = first.load()) return fail(rc, "%s", why.c_str());
    return 0;
}

// Scala Double.toInt == JVM d2i: NaN -> 0, saturating, else truncation toward zero.
int32_t d2i(double d) {
    if (d != d) return 0;
    if (d >= 2147483647.0) return INT32_MAX;
    if (d <= -2147483648.0) return INT32_MIN;
    return (int32_t)d;
}
// Scala Double.toByte == (byte)(int)d.
int8_t d2b(double d) { return (int8_t)(uint8_t)((uint32_t)d2i(d) & 0xFFu); }

static const char* kColumnTypeNames[] = {"INT", "TINYINT", "STRING"};
static const char* kCodecNames[] = {"PFOR_INT", "DENSE_INT", "DENSE_TINYINT", "DENSE_STRING"};

static int codec_from_name(const std::string& s) {
    for (int i = 0; i < 4; i++)
        if (s == kCodecNames[i]) return i;
    return -1;
}

// Width of a decoded value is a property of the codec (Column.getCodec, Column.scala:57-63).
static int finish_column(ColumnMeta* c) {
    switch (c->codec) {
        case IMM3_CODEC_PFOR_INT:
        case IMM3_CODEC_DENSE_INT: c->width = 4; break;
        case IMM3_CODEC_DENSE_TINYINT: c->width = 1; break;
        case IMM3_CODEC_DENSE_STRING: {
            const std::string* sz = nullptr;
            for (auto& kv : c->attrs)
                if (kv.first == "size") sz = &kv.second;
            if (!sz) return fail(IMM3_ERR_BAD_FORMAT, "column %s: DENSE_STRING needs dtypeAttrs size", c->name.c_str());
            char* e = nullptr;
            long v = std::strtol(sz->c_str(), &e, 10);
            if (e == sz->c_str() || *e || v <= 0 || v > 256)
                return fail(IMM3_ERR_BAD_FORMAT, "column %s: unsupported string size '%s' (1..256)", c->name.c_str(), sz->c_str());
            c->width = (int)v;
            break;
        }
        default: return fail(IMM3_ERR_BAD_FORMAT, "column %s: unknown codec", c->name.c_str());
    }
    return 0;
}

int parse_col_spec(const char* spec, ColumnMeta* out) {
    // parts = colArg.split(":"); options only when there are exactly 3 parts (LoaderCli.scala:70-80)
    std::vector<std::string> parts;
    {
        std::string s(spec), cur;
        for (char ch : s) {
            if (ch == ':') { parts.push_back(cur); cur.clear(); } else cur.push_back(ch);
        }
        parts.push_back(cur);
    }
    if (parts.size() < 2) return fail(IMM3_ERR_INVALID_ARG, "column spec '%s' is not NAME:CODEC[:OPTIONS]", spec);
    *out = ColumnMeta();
    out->name = parts[0];
    out->codec = codec_from_name(parts[1]);
    if (out->codec < 0) return fail(IMM3_ERR_INVALID_ARG, "column spec '%s': unknown codec %s", spec, parts[1].c_str());
    if (parts.size() == 3) {
        // parseColOptions: split(";") then split("=") -> (head, last)  (LoaderCli.scala:66-68)
        std::stringstream ss(parts[2]);
        std::string kv;
        while (std::getline(ss, kv, ';')) {
            if (kv.empty()) continue;
            size_t eq = kv.find('=');
            size_t leq = kv.rfind('=');
            std::string k = eq == std::string::npos ? kv : kv.substr(0, eq);
            std::string v = leq == std::string::npos ? kv : kv.substr(leq + 1);
            out->attrs.emplace_back(k, v);
        }
    }
    // Column.make: the column type follows from the codec (Column.scala:45-55)
    switch (out->codec) {
        case IMM3_CODEC_DENSE_INT:
        case IMM3_CODEC_PFOR_INT: out->ctype = IMM3_COL_INT; break;
        case IMM3_CODEC_DENSE_TINYINT: out->ctype = IMM3_COL_TINYINT; break;
        default: out->ctype = IMM3_COL_STRING;
    }
    return finish_column(out);
}

std::string table_meta_json(const TableMeta& t) {
    // ujson compact rendering; whole Doubles print without a fraction.
    std::string s = "{\"name\":\"" + json_escape(t.name) + "\",\"columns\":[";
    for (size_t i = 0; i < t.cols.size(); i++) {
        const ColumnMeta& c = t.cols[i];
        if (i) s += ",";
        s += "{\"name\":\"" + json_escape(c.name) + "\",\"columnType\":\"" + kColumnTypeNames[c.ctype] +
             "\",\"codec\":\"" + kCodecNames[c.codec] + "\",\"dtypeAttrs\":{";
        for (size_t k = 0; k < c.attrs.size(); k++) {
            if (k) s += ",";
            s += "\"" + json_escape(c.attrs[k].first) + "\":\"" + json_escape(c.attrs[k].second) + "\"";
        }
        s += "}}";
    }
    s += "],\"blockSize\":" + std::to_string(t.block_size) + "}";
    return s;
}

int parse_table_meta(const std::string& json, const std::string& origin, TableMeta* out) {
    JsonValue j;
    if (!JsonParser(json).parse(&j)) return fail(IMM3_ERR_BAD_FORMAT, "%s: invalid JSON", origin.c_str());
    const JsonValue *jn = j.get("name"), *jc = j.get("columns"), *jb = j.get("blockSize");
    if (!jn || jn->type != JsonValue::Str || !jc || jc->type != JsonValue::Arr || !jb || jb->type != JsonValue::Num)
        return fail(IMM3_ERR_BAD_FORMAT, "%s: expected {name, columns[], blockSize}", origin.c_str());
    *out = TableMeta();
    out->name = jn->str;
    out->block_size = d2i(jb->num);  // jsonValue.obj("blockSize").num.toInt (Table.scala:41)
    if (out->block_size <= 0) return fail(IMM3_ERR_BAD_FORMAT, "%s: blockSize %d", origin.c_str(), out->block_size);
    for (auto& c : jc->arr) {
        const JsonValue *cn = c.get("name"), *ct = c.get("columnType"), *cc = c.get("codec"), *ca = c.get("dtypeAttrs");
        if (!cn || !ct || !cc || !ca || cn->type != JsonValue::Str || ct->type != JsonValue::Str ||
            cc->type != JsonValue::Str || ca->type != JsonValue::Obj)
            return fail(IMM3_ERR_BAD_FORMAT, "%s: malformed column entry", origin.c_str());
        ColumnMeta m;
        m.name = cn->str;
        m.ctype = -1;
        for (int i = 0; i < 3; i++)
            if (ct->str == kColumnTypeNames[i]) m.ctype = i;
        m.codec = codec_from_name(cc->str);
        if (m.ctype < 0) return fail(IMM3_ERR_BAD_FORMAT, "%s: column %s: unknown columnType %s", origin.c_str(), m.name.c_str(), ct->str.c_str());
        if (m.codec < 0) return fail(IMM3_ERR_BAD_FORMAT, "%s: column %s: unknown codec %s", origin.c_str(), m.name.c_str(), cc->str.c_str());
        for (auto& kv : ca->obj) {
            if (kv.second.type != JsonValue::Str)
                return fail(IMM3_ERR_BAD_FORMAT, "%s: column %s: dtypeAttrs values must be strings", origin.c_str(), m.name.c_str());
            m.attrs.emplace_back(kv.first, kv.second.str);
        }
        // The decode type is the codec's, the operators dispatch on the decoded vector type
        // (Scan.scala:36-49); a columnType that disagrees with the codec cannot be produced by
        // Column.make and is rejected.
        int want = (m.codec == IMM3_CODEC_DENSE_TINYINT) ? IMM3_COL_TINYINT
                   : (m.codec == IMM3_CODEC_DENSE_STRING) ? IMM3_COL_STRING : IMM3_COL_INT;
        if (m.ctype != want)
            return fail(IMM3_ERR_BAD_FORMAT, "%s: column %s: columnType %s does not match codec %s", origin.c_str(),
                        m.name.c_str(), ct->str.c_str(), cc->str.c_str());
        int rc = finish_column(&m);
        if (rc) return rc;
        out->cols.push_back(std::move(m));
    }
    return 0;
}

int parse_segment_meta(const std::string& json, const std::string& origin, std::vector<int32_t>* offsets) {
    JsonValue j;
    if (!JsonParser(json).parse(&j)) return fail(IMM3_ERR_BAD_FORMAT, "%s: invalid JSON", origin.c_str());
    const JsonValue* a = j.get("blockOffset");  // the key is singular (Segment.scala:37,43)
    if (!a || a->type != JsonValue::Arr) return fail(IMM3_ERR_BAD_FORMAT, "%s: no blockOffset array", origin.c_str());
    offsets->clear();
    offsets->reserve(a->arr.size());
    for (auto& v : a->arr) {
        if (v.type != JsonValue::Num) return fail(IMM3_ERR_BAD_FORMAT, "%s: non-numeric block offset", origin.c_str());
        offsets->push_back(d2i(v.num));  // x.num.toInt
    }
    return 0;
}

int read_text_file(const std::string& path, std::string* out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return fail(IMM3_ERR_IO, "cannot read %s", path.c_str());
    std::stringstream ss;
    ss << f.rdbuf();
    *out = ss.str();
    return 0;
}

int list_segment_files(const std::string& table_dir, const std::string& col, const char* suffix,
                       std::vector<std::string>* names) {
    DIR* d = opendir(table_dir.c_str());
    if (!d) return fail(IMM3_ERR_IO, "cannot list %s", table_dir.c_str());
    const std::string prefix = col + "_";
    const size_t sl = std::strlen(suffix);
    names->clear();
    while (struct dirent* e = readdir(d)) {
        std::string n = e->d_name;
        if (n.size() < prefix.size() + sl) continue;
        if (n.compare(0, prefix.size(), prefix) != 0) continue;
        if (n.compare(n.size() - sl, sl, suffix) != 0) continue;
        names->push_back(n);
    }
    closedir(d);
    // sortBy(f => f.getName): String.compareTo == byte order for ASCII names, so
    // "id_1.dat" < "id_10.dat" < "id_2.dat" (SURVEY.md §3.4-8).
    std::sort(names->begin(), names->end());
    return 0;
}

}  // namespace imm3
