// sql.cpp — restatement of the reference's parser-combinator grammar for the Project form
// (SQLParser.scala:8-129), so the C ABI can take the same text the reference CLI takes
// (SqlCli.scala:60).  JavaTokenParsers semantics: whitespace is skipped before every literal and
// regex; a literal matches as a PREFIX (no word boundary); `|` backtracks on failure.
//
//   query        := queryProjectAgg | queryProjectAggNoGroup | queryProject        (:13)
//   queryProject := "select" repsep(fieldIdent, ",") "from" fieldIdent where limit (:27-31, :98-99)
//   where        := opt("where" filter)                                            (:51-56)
//   filter       := "(" repsep(filter,"and") ")" | "(" repsep(filter,"or") ")"
//                 | fieldIdent "=" value | fieldIdent "=" "'" value "'"
//                 | fieldIdent ">" value | fieldIdent "<" value                    (:58-96)
//   limit        := opt("limit" [\d]+)                                             (:37)
//   fieldIdent   := [\w\#]+      value := [\w0-9\#]+                               (:121-123)
// Aggregates (sum/min/max/count, group by) parse in the reference but are outside this path
//  and return IMM3_ERR_UNSUPPORTED.
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "plan.hpp"

namespace imm3 {

namespace {

struct Cursor {
    const char* s;
    size_t pos;
};

inline bool is_word(char c) {  // Java \w = [a-zA-Z_0-9]; plus '#'
    return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' || c == '#';
}
inline void skip_ws(Cursor& c) {  // JavaTokenParsers whiteSpace = \s+
    while (c.s[c.pos] == ' ' || c.s[c.pos] == '\t' || c.s[c.pos] == '\n' || c.s[c.pos] == '\r' || c.s[c.pos] == '\f' ||
           c.s[c.pos] == '\v')
        c.pos++;
}
inline bool lit(Cursor& c, const char* word) {
    Cursor t = c;
    skip_ws(t);
    size_t n = std::strlen(word);
    if (std::strncmp(t.s + t.pos, word, n)) return false;
    t.pos += n;
    c = t;
    return true;
}
inline bool ident(Cursor& c, std::string* out) {
    Cursor t = c;
    skip_ws(t);
    size_t b = t.pos;
    while (is_word(t.s[t.pos])) t.pos++;
    if (t.pos == b) return false;
    out->assign(t.s + b, t.pos - b);
    c = t;
    return true;
}

// java.lang.Double.parseDouble restricted to the characters `value` can contain.
bool java_parse_double(const std::string& s, double* out) {
    if (s == "NaN") { *out = std::nan(""); return true; }
    if (s == "Infinity") { *out = INFINITY; return true; }
    size_t n = s.size(), i = 0;
    std::string body = s;
    if (n > 2 && s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) {
        i = 2;
        size_t d0 = i;
        while (i < n && std::isxdigit((unsigned char)s[i])) i++;
        if (i == d0 || i >= n || (s[i] != 'p' && s[i] != 'P')) return false;
        i++;
        size_t e0 = i;
        while (i < n && s[i] >= '0' && s[i] <= '9') i++;
        if (i == e0) return false;
    } else {
        while (i < n && s[i] >= '0' && s[i] <= '9') i++;
        if (i == 0) return false;
        if (i < n && (s[i] == 'e' || s[i] == 'E')) {
            i++;
            size_t e0 = i;
            while (i < n && s[i] >= '0' && s[i] <= '9') i++;
            if (i == e0) return false;
        }
    }
    if (i < n && (s[i] == 'f' || s[i] == 'F' || s[i] == 'd' || s[i] == 'D')) { body = s.substr(0, i); i++; }
    if (i != n) return false;
    char* e = nullptr;
    *out = std::strtod(body.c_str(), &e);
    return e && *e == 0;
}

enum { kOk = 0, kNoMatch = 1, kError = 2 };

struct Ctx {
    ParsedQuery* q;
    std::string err;
};

int parse_filter(Cursor& c, Ctx& ctx);

// "(" ~> repsep(filter, sep) <~ ")"
int parse_group(Cursor& c, Ctx& ctx, const char* sep) {
    Cursor t = c;
    size_t mark = ctx.q->preds.size();
    auto rollback = [&] {
        ctx.q->preds.resize(mark);
        ctx.q->pred_cols.resize(mark);
        ctx.q->pred_strs.resize(mark);
    };
    if (!lit(t, "(")) return kNoMatch;
    int count = 0;
    int rc = parse_filter(t, ctx);
    if (rc == kError) return rc;
    if (rc == kOk) {
        count = 1;
        for (;;) {
            Cursor u = t;
            if (!lit(u, sep)) break;
            rc = parse_filter(u, ctx);
            if (rc == kError) return rc;
            if (rc != kOk) break;  // repsep stops before a separator that is not followed by an element
            t = u;
            count++;
        }
    }
    if (!lit(t, ")")) { rollback(); return kNoMatch; }
    if (count == 0) {  // xs.head on an empty list
        ctx.err = "empty filter list: head of empty list";
        return kError;
    }
    c = t;
    return kOk;
}

int parse_cmp(Cursor& c, Ctx& ctx, const char* opstr, int op) {
    Cursor t = c;
    std::string f, v;
    if (!ident(t, &f) || !lit(t, opstr) || !ident(t, &v)) return kNoMatch;
    double d;
    if (!java_parse_double(v, &d)) {  // v.toDouble throws inside the semantic action
        ctx.err = "java.lang.NumberFormatException: For input string: \"" + v + "\"";
        return kError;
    }
    imm3_pred p{};
    p.op = op;
    p.num = d;
    ctx.q->pred_cols.push_back(f);
    ctx.q->pred_strs.emplace_back();
    ctx.q->preds.push_back(p);
    c = t;
    return kOk;
}

int parse_eq_string(Cursor& c, Ctx& ctx) {
    Cursor t = c;
    std::string f, v;
    if (!ident(t, &f) || !lit(t, "=") || !lit(t, "'") || !ident(t, &v) || !lit(t, "'")) return kNoMatch;
    imm3_pred p{};
    p.op = IMM3_OP_MATCH;  // Select(f, Match(List(v)))  (SQLParser.scala:80-84)
    ctx.q->pred_cols.push_back(f);
    ctx.q->pred_strs.push_back({v});
    ctx.q->preds.push_back(p);
    c = t;
    return kOk;
}

int parse_filter(Cursor& c, Ctx& ctx) {
    int rc;
    if ((rc = parse_group(c, ctx, "and")) != kNoMatch) return rc;
    if ((rc = parse_group(c, ctx, "or")) != kNoMatch) return rc;  // Or is evaluated as And (Engine.scala:240)
    if ((rc = parse_cmp(c, ctx, "=", IMM3_OP_EQ)) != kNoMatch) return rc;
    if ((rc = parse_eq_string(c, ctx)) != kNoMatch) return rc;
    if ((rc = parse_cmp(c, ctx, ">", IMM3_OP_GT)) != kNoMatch) return rc;
    if ((rc = parse_cmp(c, ctx, "<", IMM3_OP_LT)) != kNoMatch) return rc;
    return kNoMatch;
}

}  // namespace

void ParsedQuery::fix_pointers() {
    pred_str_ptrs.assign(preds.size(), {});
    for (size_t i = 0; i < preds.size(); i++) {
        preds[i].col = pred_cols[i].c_str();
        for (auto& s : pred_strs[i]) pred_str_ptrs[i].push_back(s.c_str());
        preds[i].strs = pred_str_ptrs[i].empty() ? nullptr : pred_str_ptrs[i].data();
        preds[i].nstrs = (int32_t)pred_str_ptrs[i].size();
    }
}

int parse_sql(const char* sql, ParsedQuery* out) {
    if (!sql) return fail(IMM3_ERR_INVALID_ARG, "sql is NULL");
    *out = ParsedQuery();
    Ctx ctx{out, ""};
    Cursor c{sql, 0};
    if (!lit(c, "select")) return fail(IMM3_ERR_INVALID_ARG, "SQL parse failure: 'select' expected");
    {  // selectProjectAgg is tried first (SQLParser.scala:13): an aggregate call in the select list
        for (const char* agg : {"sum", "min", "max", "count"}) {
            Cursor t = c;
            if (lit(t, agg) && lit(t, "("))
                return fail(IMM3_ERR_UNSUPPORTED, "aggregate queries (ProjectAgg) are outside the scan/filter/project path");
        }
    }
    std::string f;
    if (ident(c, &f)) {  // repsep(fieldIdent, ",")
        out->proj.push_back(f);
        for (;;) {
            Cursor t = c;
            if (!lit(t, ",") || !ident(t, &f)) break;
            out->proj.push_back(f);
            c = t;
        }
    }
    if (!lit(c, "from")) return fail(IMM3_ERR_INVALID_ARG, "SQL parse failure: 'from' expected at offset %zu", c.pos);
    if (!ident(c, &out->table)) return fail(IMM3_ERR_INVALID_ARG, "SQL parse failure: table name expected at offset %zu", c.pos);
    {
        Cursor t = c;
        if (lit(t, "where")) {
            int rc = parse_filter(t, ctx);
            if (rc == kError) return fail(IMM3_ERR_INVALID_ARG, "SQL parse failure: %s", ctx.err.c_str());
            if (rc == kOk) c = t;  // opt(...): on failure nothing is consumed and parseAll fails below
        }
    }
    {
        Cursor t = c;
        if (lit(t, "limit")) {
            skip_ws(t);
            size_t b = t.pos;
            while (t.s[t.pos] >= '0' && t.s[t.pos] <= '9') t.pos++;
            if (t.pos > b) {
                std::string digits(t.s + b, t.pos - b);
                errno = 0;
                unsigned long long v = std::strtoull(digits.c_str(), nullptr, 10);
                if (errno || v > 2147483647ull)  // x.toInt
                    return fail(IMM3_ERR_INVALID_ARG, "SQL parse failure: java.lang.NumberFormatException: For input string: \"%s\"", digits.c_str());
                out->limit = (int64_t)v;
                c = t;
            }
        }
    }
    skip_ws(c);
    if (c.s[c.pos] != 0) return fail(IMM3_ERR_INVALID_ARG, "SQL parse failure at offset %zu: '%s'", c.pos, c.s + c.pos);
    out->fix_pointers();
    return 0;
}

}  // namespace imm3
