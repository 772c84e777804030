/*
 * oracle.c — CPU restatement of immutable3's Scanner -> RangeFilter/MatchFilter -> Project/LIMIT.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Plain C11 + pthreads, no dependencies.
 * Every function cites the reference file:line it follows (relative to the reference checkout).
 * PFOR_INT section: PARITY UNPINNED (JavaFastPFOR 0.1.10 is not under /root/reference).
 */
#define _GNU_SOURCE
#include "oracle.h"

#include <dirent.h>
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

static __thread char g_err[512];
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
const char* orc_last_error(void) { return g_err; }
void orc_free(void* p) { free(p); }

/* ===================================================================================== */
/* Value-level restatements                                                               */
/* ===================================================================================== */

/* Conversions.bytesToInt, Conversions.scala:17-24.  In Scala `+` binds tighter than `<<`, so
 *   result = result + (b3&0xFF) << 8   parses as   (result + (b3&0xFF)) << 8
 * i.e. Horner's rule over b3,b2,b1 then + b0: little-endian two's complement. */
int32_t orc_bytes_to_int(const uint8_t* b) {
    uint32_t r = 0;
    r = (r + (uint32_t)b[3]) << 8;
    r = (r + (uint32_t)b[2]) << 8;
    r = (r + (uint32_t)b[1]) << 8;
    r = r + (uint32_t)b[0];
    return (int32_t)r;
}

/* IntType.valueToBytes, DataType.scala:40-47 */
void orc_int_to_bytes(int32_t v, uint8_t* out) {
    out[3] = (uint8_t)((v >> 24) & 0xFF);
    out[2] = (uint8_t)((v >> 16) & 0xFF);
    out[1] = (uint8_t)((v >> 8) & 0xFF);
    out[0] = (uint8_t)(v & 0xFF);
}

/* Scala `Double.toInt` = JVM d2i (JLS 5.1.3): NaN -> 0, saturate, else round toward zero.
 * Used by Select.scala:65,103,141 on INT columns. */
int32_t orc_d2i(double d) {
    if (d != d) return 0;
    if (d >= 2147483647.0) return INT32_MAX;
    if (d <= -2147483648.0) return INT32_MIN;
    return (int32_t)d; /* C truncates toward zero */
}

/* Scala `Double.toByte` = (byte)(int)d: d2i then keep the low 8 bits, signed.
 * Used by Select.scala:73,111,149 on TINYINT columns. */
int8_t orc_d2b(double d) { return (int8_t)(uint8_t)((uint32_t)orc_d2i(d) & 0xFFu); }

/* DenseCodecInt/TinyInt/String.decode, DenseCodec.scala:34-74:
 *     val chunk = new Array[Byte](dtype.size); while (data.read(chunk) != -1) segment += bytesToValue(chunk)
 * ByteArrayInputStream.read(chunk) copies min(remaining, size) bytes and returns that count (never -1
 * while bytes remain), so a ragged tail yields one extra value whose trailing bytes are STALE bytes of
 * the previous chunk (zeros if it is the first).  The cells are written raw (width bytes each); callers
 * interpret them with orc_bytes_to_int / int8 / k-byte string. */
int64_t orc_dense_decode(const uint8_t* bytes, int64_t nbytes, int width, uint8_t* out_cells) {
    uint8_t chunk[256];
    if (width <= 0 || width > (int)sizeof chunk) return -1;
    if (nbytes % width == 0) { /* whole values only: every read fills the chunk, the cells ARE the bytes */
        memcpy(out_cells, bytes, (size_t)nbytes);
        return nbytes / width;
    }
    memset(chunk, 0, (size_t)width);
    int64_t pos = 0, n = 0;
    while (pos < nbytes) {
        int64_t take = nbytes - pos < width ? nbytes - pos : width;
        memcpy(chunk, bytes + pos, (size_t)take);
        pos += take;
        memcpy(out_cells + n * width, chunk, (size_t)width);
        n++;
    }
    return n;
}

/* ===================================================================================== */
/* Sorted-integer codec: JavaFastPFOR 0.1.10 IntegratedIntCompressor  (PARITY UNPINNED)    */
/*   compress():  out[0] = n; SkippableIntegratedComposition(IntegratedBinaryPacking,      */
/*                IntegratedVariableByte).headlessCompress(..., initvalue = 0)             */
/* Call sites in the reference: PFORCodec.scala:15,18,31,49.                               */
/* ===================================================================================== */

static int bits32(uint32_t x) { return x ? 32 - __builtin_clz(x) : 0; }

/* Util.maxdiffbits(initoffset, in, pos, 32) */
static int maxdiffbits(int32_t init, const int32_t* in) {
    uint32_t mask = (uint32_t)in[0] - (uint32_t)init;
    for (int k = 1; k < 32; k++) mask |= (uint32_t)in[k] - (uint32_t)in[k - 1];
    return bits32(mask);
}

/* IntegratedBitPacking.integratedpack(initoffset, in, inpos, out, outpos, bit):
 * bit==0: nothing; bit==32: raw copy of the 32 VALUES; else the 32 deltas packed LSB-first,
 * value j at bit offset j*bit of the little-endian word stream. */
static void integratedpack(int32_t init, const int32_t* in, uint32_t* out, int bit) {
    if (bit == 0) return;
    if (bit == 32) {
        for (int k = 0; k < 32; k++) out[k] = (uint32_t)in[k];
        return;
    }
    for (int k = 0; k < bit; k++) out[k] = 0;
    uint32_t prev = (uint32_t)init;
    for (int k = 0; k < 32; k++) {
        uint32_t d = (uint32_t)in[k] - prev;
        prev = (uint32_t)in[k];
        int off = k * bit, w = off >> 5, s = off & 31;
        out[w] |= d << s;
        if (s + bit > 32) out[w + 1] |= d >> (32 - s);
    }
}

static void integratedunpack(int32_t init, const uint32_t* in, int32_t* out, int bit) {
    if (bit == 0) { /* integratedunpack0: Arrays.fill(out, initoffset) */
        for (int k = 0; k < 32; k++) out[k] = init;
        return;
    }
    if (bit == 32) {
        for (int k = 0; k < 32; k++) out[k] = (int32_t)in[k];
        return;
    }
    uint32_t mask = (1u << bit) - 1u, prev = (uint32_t)init;
    for (int k = 0; k < 32; k++) {
        int off = k * bit, w = off >> 5, s = off & 31;
        uint32_t d = in[w] >> s;
        if (s + bit > 32) d |= in[w + 1] << (32 - s);
        prev += d & mask;
        out[k] = (int32_t)prev;
    }
}

int64_t orc_iic_compress(const int32_t* in, int32_t n, int32_t* out_words, int64_t cap) {
    /* worst case: 1 + n/32 headers + n words + var-byte (<= 5 bytes per value, < 32 values) */
    int64_t need = 1 + (int64_t)n + n / 32 + 48;
    if (cap < need) return -need;
    uint32_t* out = (uint32_t*)out_words;
    int64_t op = 0;
    out[op++] = (uint32_t)n;
    int32_t init = 0;
    int32_t packed = n - n % 32; /* Util.greatestMultiple(inlength, 32) */
    int32_t s = 0;
    /* IntegratedBinaryPacking.headlessCompress: groups of 4 mini-blocks share one header word */
    for (; s + 128 - 1 < packed; s += 128) {
        int b1 = maxdiffbits(init, in + s);
        int b2 = maxdiffbits(in[s + 31], in + s + 32);
        int b3 = maxdiffbits(in[s + 63], in + s + 64);
        int b4 = maxdiffbits(in[s + 95], in + s + 96);
        out[op++] = ((uint32_t)b1 << 24) | ((uint32_t)b2 << 16) | ((uint32_t)b3 << 8) | (uint32_t)b4;
        integratedpack(init, in + s, out + op, b1); op += b1;
        integratedpack(in[s + 31], in + s + 32, out + op, b2); op += b2;
        integratedpack(in[s + 63], in + s + 64, out + op, b3); op += b3;
        integratedpack(in[s + 95], in + s + 96, out + op, b4); op += b4;
        init = in[s + 127];
    }
    for (; s < packed; s += 32) {
        int b = maxdiffbits(init, in + s);
        out[op++] = (uint32_t)b;
        integratedpack(init, in + s, out + op, b); op += b;
        init = in[s + 31];
    }
    /* IntegratedVariableByte.headlessCompress on the remaining n%32 values: 7 data bits per byte,
     * low group first, the LAST byte of a value has bit 7 set; bytes packed little-endian into
     * words, zero padded to a word boundary. */
    int32_t rem = n - packed;
    if (rem > 0) {
        uint8_t buf[32 * 5 + 4];
        int bp = 0;
        for (int k = packed; k < n; k++) {
            uint32_t val = (uint32_t)in[k] - (uint32_t)init;
            init = in[k];
            if (val < (1u << 7)) {
                buf[bp++] = (uint8_t)(val | 0x80);
            } else if (val < (1u << 14)) {
                buf[bp++] = (uint8_t)(val & 0x7F);
                buf[bp++] = (uint8_t)((val >> 7) | 0x80);
            } else if (val < (1u << 21)) {
                buf[bp++] = (uint8_t)(val & 0x7F);
                buf[bp++] = (uint8_t)((val >> 7) & 0x7F);
                buf[bp++] = (uint8_t)((val >> 14) | 0x80);
            } else if (val < (1u << 28)) {
                buf[bp++] = (uint8_t)(val & 0x7F);
                buf[bp++] = (uint8_t)((val >> 7) & 0x7F);
                buf[bp++] = (uint8_t)((val >> 14) & 0x7F);
                buf[bp++] = (uint8_t)((val >> 21) | 0x80);
            } else {
                buf[bp++] = (uint8_t)(val & 0x7F);
                buf[bp++] = (uint8_t)((val >> 7) & 0x7F);
                buf[bp++] = (uint8_t)((val >> 14) & 0x7F);
                buf[bp++] = (uint8_t)((val >> 21) & 0x7F);
                buf[bp++] = (uint8_t)((val >> 28) | 0x80);
            }
        }
        while (bp % 4) buf[bp++] = 0;
        for (int i = 0; i < bp; i += 4)
            out[op++] = (uint32_t)buf[i] | ((uint32_t)buf[i + 1] << 8) | ((uint32_t)buf[i + 2] << 16) |
                        ((uint32_t)buf[i + 3] << 24);
    }
    return op;
}

int32_t orc_iic_uncompress(const int32_t* words, int64_t nwords, int32_t* out, int32_t cap) {
    const uint32_t* in = (const uint32_t*)words;
    if (nwords < 1) return -1;
    int32_t n = (int32_t)in[0];
    if (n < 0 || n > cap) return -1;
    int64_t ip = 1;
    int32_t init = 0;
    int32_t packed = n - n % 32;
    int32_t s = 0;
    for (; s + 128 - 1 < packed; s += 128) {
        if (ip >= nwords) return -1;
        uint32_t h = in[ip++];
        int b[4] = {(int)(h >> 24), (int)((h >> 16) & 0xFF), (int)((h >> 8) & 0xFF), (int)(h & 0xFF)};
        for (int m = 0; m < 4; m++) {
            if (b[m] > 32 || ip + b[m] > nwords) return -1;
            integratedunpack(init, in + ip, out + s + 32 * m, b[m]);
            ip += b[m];
            init = out[s + 32 * m + 31];
        }
    }
    for (; s < packed; s += 32) {
        if (ip >= nwords) return -1;
        int b = (int)in[ip++];
        if (b < 0 || b > 32 || ip + b > nwords) return -1;
        integratedunpack(init, in + ip, out + s, b);
        ip += b;
        init = out[s + 31];
    }
    /* IntegratedVariableByte.headlessUncompress */
    int sh = 0;
    uint32_t v = 0;
    int shift = 0;
    for (int32_t k = packed; k < n;) {
        if (ip >= nwords) return -1;
        uint32_t c = in[ip] >> sh;
        sh += 8;
        ip += sh >> 5;
        sh &= 31;
        v += (c & 127u) << shift;
        if (c & 128u) {
            out[k] = (int32_t)(v + (uint32_t)init);
            init = out[k];
            k++;
            v = 0;
            shift = 0;
        } else {
            shift += 7;
        }
    }
    return n;
}

/* PFORCodecInt.encode, PFORCodec.scala:17-28: ByteBuffer.allocate(words*4 + 8), putInt (big-endian,
 * the JVM default order) per word, then the WHOLE backing array is written: 8 zero bytes trail. */
int64_t orc_pfor_encode_block(const int32_t* in, int32_t n, uint8_t* out, int64_t cap) {
    int64_t wcap = 1 + (int64_t)n + n / 32 + 48;
    int32_t* w = (int32_t*)malloc((size_t)wcap * 4);
    if (!w) return -1;
    int64_t nw = orc_iic_compress(in, n, w, wcap);
    if (nw < 0) { free(w); return -1; }
    int64_t nbytes = nw * 4 + 8;
    if (!out) { free(w); return nbytes; }
    if (cap < nbytes) { free(w); return -nbytes; }
    for (int64_t i = 0; i < nw; i++) {
        uint32_t x = (uint32_t)w[i];
        out[4 * i + 0] = (uint8_t)(x >> 24);
        out[4 * i + 1] = (uint8_t)(x >> 16);
        out[4 * i + 2] = (uint8_t)(x >> 8);
        out[4 * i + 3] = (uint8_t)x;
    }
    memset(out + nw * 4, 0, 8);
    free(w);
    return nbytes;
}

int32_t orc_pfor_decode_block(const uint8_t* bytes, int64_t nbytes, int32_t* out, int32_t cap) {
    if (nbytes < 12 || nbytes % 4) return -1;
    int64_t nw = nbytes / 4 - 2; /* strip the 8 pad bytes */
    int32_t* w = (int32_t*)malloc((size_t)(nw + 2) * 4);
    if (!w) return -1;
    for (int64_t i = 0; i < nw; i++)
        w[i] = (int32_t)(((uint32_t)bytes[4 * i] << 24) | ((uint32_t)bytes[4 * i + 1] << 16) |
                         ((uint32_t)bytes[4 * i + 2] << 8) | (uint32_t)bytes[4 * i + 3]);
    int32_t n = orc_iic_uncompress(w, nw, out, cap);
    free(w);
    return n;
}

/* ===================================================================================== */
/* Minimal JSON reader (ujson-written files: Table.scala:27-35, Column.scala:21-29,        */
/* Segment.scala:41-45)                                                                   */
/* ===================================================================================== */

typedef enum { J_NULL, J_BOOL, J_NUM, J_STR, J_ARR, J_OBJ } jtype;
typedef struct jval {
    jtype t;
    double num;
    char* str;
    struct jval* items; /* array items or object values */
    char** keys;        /* object keys */
    int n;
} jval;

static void jfree(jval* v) {
    if (!v) return;
    free(v->str);
    for (int i = 0; i < v->n; i++) {
        jfree(&v->items[i]);
        if (v->keys) free(v->keys[i]);
    }
    free(v->items);
    free(v->keys);
    v->str = NULL; v->items = NULL; v->keys = NULL; v->n = 0;
}
static void jskip(const char** p) {
    while (**p == ' ' || **p == '\n' || **p == '\t' || **p == '\r') (*p)++;
}
static int jparse(const char** p, jval* out);
static int jstring(const char** p, char** out) {
    if (**p != '"') return -1;
    (*p)++;
    size_t cap = 32, n = 0;
    char* s = (char*)malloc(cap);
    while (**p && **p != '"') {
        char c = **p;
        if (c == '\\') {
            (*p)++;
            c = **p;
            switch (c) {
                case 'n': c = '\n'; break;
                case 't': c = '\t'; break;
                case 'r': c = '\r'; break;
                case 'b': c = '\b'; break;
                case 'f': c = '\f'; break;
                case 'u': { /* \uXXXX: ASCII subset is enough for identifiers */
                    unsigned x = 0;
                    for (int i = 1; i <= 4; i++) {
                        char h = (*p)[i];
                        if (!h) { free(s); return -1; }
                        x = x * 16 + (unsigned)(h <= '9' ? h - '0' : (h | 32) - 'a' + 10);
                    }
                    (*p) += 4;
                    c = (char)x;
                    break;
                }
                default: break; /* \" \\ \/ */
            }
        }
        if (n + 2 > cap) { cap *= 2; s = (char*)realloc(s, cap); }
        s[n++] = c;
        (*p)++;
    }
    if (**p != '"') { free(s); return -1; }
    (*p)++;
    s[n] = 0;
    *out = s;
    return 0;
}
static int jparse(const char** p, jval* out) {
    memset(out, 0, sizeof *out);
    jskip(p);
    char c = **p;
    if (c == '{' || c == '[') {
        int obj = c == '{';
        out->t = obj ? J_OBJ : J_ARR;
        (*p)++;
        int cap = 8;
        out->items = (jval*)calloc((size_t)cap, sizeof(jval));
        if (obj) out->keys = (char**)calloc((size_t)cap, sizeof(char*));
        jskip(p);
        if (**p == (obj ? '}' : ']')) { (*p)++; return 0; }
        for (;;) {
            if (out->n == cap) {
                cap *= 2;
                out->items = (jval*)realloc(out->items, (size_t)cap * sizeof(jval));
                if (obj) out->keys = (char**)realloc(out->keys, (size_t)cap * sizeof(char*));
            }
            jskip(p);
            if (obj) {
                char* k = NULL;
                if (jstring(p, &k)) return -1;
                out->keys[out->n] = k;
                jskip(p);
                if (**p != ':') { out->n++; memset(&out->items[out->n - 1], 0, sizeof(jval)); return -1; }
                (*p)++;
            }
            int rc = jparse(p, &out->items[out->n]);
            out->n++;
            if (rc) return -1;
            jskip(p);
            if (**p == ',') { (*p)++; continue; }
            if (**p == (obj ? '}' : ']')) { (*p)++; return 0; }
            return -1;
        }
    }
    if (c == '"') { out->t = J_STR; return jstring(p, &out->str); }
    if (!strncmp(*p, "true", 4)) { out->t = J_BOOL; out->num = 1; *p += 4; return 0; }
    if (!strncmp(*p, "false", 5)) { out->t = J_BOOL; *p += 5; return 0; }
    if (!strncmp(*p, "null", 4)) { out->t = J_NULL; *p += 4; return 0; }
    char* end = NULL;
    out->num = strtod(*p, &end);
    if (end == *p) return -1;
    out->t = J_NUM;
    *p = end;
    return 0;
}
static const jval* jget(const jval* o, const char* key) {
    if (!o || o->t != J_OBJ) return NULL;
    for (int i = 0; i < o->n; i++)
        if (!strcmp(o->keys[i], key)) return &o->items[i];
    return NULL;
}
static char* read_file(const char* path, size_t* len) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    char* s = (char*)malloc((size_t)n + 1);
    if (fread(s, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(s); return NULL; }
    fclose(f);
    s[n] = 0;
    if (len) *len = (size_t)n;
    return s;
}

/* ===================================================================================== */
/* SegmentManager restatement                                                             */
/* ===================================================================================== */

typedef struct {
    char file[300];   /* file name, the sort key of SegmentManager.scala:41 */
    int file_id;
    uint8_t* data;    /* mmap, FileChannel.map READ_ONLY (SegmentManager.scala:81-87) */
    int64_t nbytes;
    int32_t* offsets; /* SegmentMeta.blockOffsets (Segment.scala:33) */
    int noffsets;
} seg_t;

typedef struct {
    char name[128];
    int ctype, codec, size, width;
    seg_t* segs;
    int nsegs;
} col_t;

typedef struct {
    char name[128];
    int block_size;
    col_t* cols;
    int ncols;
} table_t;

struct orc_db {
    char dir[1024];
    table_t* tables;
    int ntables;
};

static int ends_with(const char* s, const char* suf) {
    size_t a = strlen(s), b = strlen(suf);
    return a >= b && !strcmp(s + a - b, suf);
}
typedef struct { char file[300]; } fname_t;
/* sortBy(f => f.getName): String.compareTo, byte order on ASCII names */
static int cmp_fname(const void* a, const void* b) { return strcmp(((const fname_t*)a)->file, ((const fname_t*)b)->file); }

/* getSegmentFiles / getSegmentMetaFiles: listFiles().filter(startsWith(col+"_") && endsWith(suffix))
 * .sortBy(_.getName), SegmentManager.scala:38-42, 61-65. */
static int list_files(const char* dir, const char* col, const char* suffix, fname_t** out, int* n) {
    DIR* d = opendir(dir);
    if (!d) return fail(ORC_ERR_IO, "cannot list %s", dir);
    char prefix[160];
    snprintf(prefix, sizeof prefix, "%s_", col);
    size_t pl = strlen(prefix);
    int cap = 64, cnt = 0;
    fname_t* v = (fname_t*)malloc((size_t)cap * sizeof *v);
    struct dirent* e;
    while ((e = readdir(d))) {
        if (strncmp(e->d_name, prefix, pl) || !ends_with(e->d_name, suffix)) continue;
        if (cnt == cap) { cap *= 2; v = (fname_t*)realloc(v, (size_t)cap * sizeof *v); }
        snprintf(v[cnt].file, sizeof v[cnt].file, "%s", e->d_name);
        cnt++;
    }
    closedir(d);
    qsort(v, (size_t)cnt, sizeof *v, cmp_fname);
    *out = v;
    *n = cnt;
    return 0;
}

static int load_column(const char* tdir, col_t* c) {
    fname_t *dats = NULL, *metas = NULL;
    int nd = 0, nm = 0, rc;
    if ((rc = list_files(tdir, c->name, ".dat", &dats, &nd))) return rc;
    if ((rc = list_files(tdir, c->name, ".meta", &metas, &nm))) { free(dats); return rc; }
    if (nd != nm) { free(dats); free(metas); return fail(ORC_ERR_BAD_FORMAT, "%s/%s: %d .dat vs %d .meta", tdir, c->name, nd, nm); }
    c->segs = (seg_t*)calloc((size_t)(nd ? nd : 1), sizeof(seg_t));
    c->nsegs = nd;
    for (int i = 0; i < nd; i++) {
        seg_t* s = &c->segs[i];
        snprintf(s->file, sizeof s->file, "%s", dats[i].file);
        s->file_id = atoi(dats[i].file + strlen(c->name) + 1);
        char path[1400];
        snprintf(path, sizeof path, "%s/%s", tdir, dats[i].file);
        int fd = open(path, O_RDONLY);
        if (fd < 0) { rc = fail(ORC_ERR_IO, "open %s: %s", path, strerror(errno)); goto done; }
        struct stat st;
        fstat(fd, &st);
        s->nbytes = st.st_size;
        if (st.st_size > 0) {
            s->data = (uint8_t*)mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (s->data == MAP_FAILED) { s->data = NULL; close(fd); rc = fail(ORC_ERR_IO, "mmap %s", path); goto done; }
        }
        close(fd);
        /* SegmentMeta.load: json.obj("blockOffset").arr.map(_.num.toInt), Segment.scala:36-50.
         * The i-th .meta in sorted order pairs with the i-th .dat in sorted order. */
        snprintf(path, sizeof path, "%s/%s", tdir, metas[i].file);
        char* txt = read_file(path, NULL);
        if (!txt) { rc = fail(ORC_ERR_IO, "read %s", path); goto done; }
        const char* p = txt;
        jval j;
        int prc = jparse(&p, &j);
        const jval* arr = prc ? NULL : jget(&j, "blockOffset");
        if (!arr || arr->t != J_ARR) { jfree(&j); free(txt); rc = fail(ORC_ERR_BAD_FORMAT, "%s: no blockOffset array", path); goto done; }
        s->noffsets = arr->n;
        s->offsets = (int32_t*)malloc((size_t)(arr->n ? arr->n : 1) * 4);
        for (int k = 0; k < arr->n; k++) s->offsets[k] = orc_d2i(arr->items[k].num);
        jfree(&j);
        free(txt);
    }
    rc = 0;
done:
    free(dats);
    free(metas);
    return rc;
}

/* TableIO.load + Column.fromJsonValue, Table.scala:37-48, Column.scala:31-38 */
static int load_table(const char* data_dir, const char* name, table_t* t) {
    char path[1400], tdir[1200];
    snprintf(tdir, sizeof tdir, "%s/%s", data_dir, name);
    snprintf(path, sizeof path, "%s/_table.meta", tdir);
    char* txt = read_file(path, NULL);
    if (!txt) return fail(ORC_ERR_IO, "read %s", path);
    const char* p = txt;
    jval j;
    if (jparse(&p, &j)) { jfree(&j); free(txt); return fail(ORC_ERR_BAD_FORMAT, "%s: bad JSON", path); }
    const jval *jn = jget(&j, "name"), *jc = jget(&j, "columns"), *jb = jget(&j, "blockSize");
    if (!jn || jn->t != J_STR || !jc || jc->t != J_ARR || !jb || jb->t != J_NUM) {
        jfree(&j); free(txt);
        return fail(ORC_ERR_BAD_FORMAT, "%s: missing name/columns/blockSize", path);
    }
    snprintf(t->name, sizeof t->name, "%s", jn->str);
    t->block_size = orc_d2i(jb->num);
    t->ncols = jc->n;
    t->cols = (col_t*)calloc((size_t)(jc->n ? jc->n : 1), sizeof(col_t));
    int rc = 0;
    for (int i = 0; i < jc->n && !rc; i++) {
        const jval* c = &jc->items[i];
        const jval *cn = jget(c, "name"), *ct = jget(c, "columnType"), *cc = jget(c, "codec"), *ca = jget(c, "dtypeAttrs");
        if (!cn || !ct || !cc || !ca || cn->t != J_STR || ct->t != J_STR || cc->t != J_STR || ca->t != J_OBJ) {
            rc = fail(ORC_ERR_BAD_FORMAT, "%s: bad column %d", path, i);
            break;
        }
        col_t* col = &t->cols[i];
        snprintf(col->name, sizeof col->name, "%s", cn->str);
        if (!strcmp(ct->str, "INT")) col->ctype = ORC_COL_INT;
        else if (!strcmp(ct->str, "TINYINT")) col->ctype = ORC_COL_TINYINT;
        else if (!strcmp(ct->str, "STRING")) col->ctype = ORC_COL_STRING;
        else { rc = fail(ORC_ERR_BAD_FORMAT, "unknown columnType %s", ct->str); break; }
        if (!strcmp(cc->str, "PFOR_INT")) col->codec = ORC_CODEC_PFOR_INT;
        else if (!strcmp(cc->str, "DENSE_INT")) col->codec = ORC_CODEC_DENSE_INT;
        else if (!strcmp(cc->str, "DENSE_TINYINT")) col->codec = ORC_CODEC_DENSE_TINYINT;
        else if (!strcmp(cc->str, "DENSE_STRING")) col->codec = ORC_CODEC_DENSE_STRING;
        else { rc = fail(ORC_ERR_BAD_FORMAT, "unknown codec %s", cc->str); break; }
        /* Column.getCodec, Column.scala:57-63: the decode width comes from the CODEC */
        switch (col->codec) {
            case ORC_CODEC_PFOR_INT:
            case ORC_CODEC_DENSE_INT: col->width = 4; break;
            case ORC_CODEC_DENSE_TINYINT: col->width = 1; break;
            default: {
                const jval* sz = jget(ca, "size");
                if (!sz || sz->t != J_STR) { rc = fail(ORC_ERR_BAD_FORMAT, "column %s: DENSE_STRING needs dtypeAttrs.size", col->name); break; }
                col->size = atoi(sz->str);
                col->width = col->size;
                if (col->size <= 0 || col->size > 256) rc = fail(ORC_ERR_BAD_FORMAT, "column %s: size %d", col->name, col->size);
            }
        }
        if (!rc) rc = load_column(tdir, col);
    }
    jfree(&j);
    free(txt);
    return rc;
}

int orc_open(const char* data_dir, orc_db** out) {
    /* SegmentManager.getTables: every sub-directory is a table, SegmentManager.scala:27-35 */
    DIR* d = opendir(data_dir);
    if (!d) return fail(ORC_ERR_IO, "cannot open data dir %s", data_dir);
    orc_db* db = (orc_db*)calloc(1, sizeof *db);
    snprintf(db->dir, sizeof db->dir, "%s", data_dir);
    int cap = 8;
    db->tables = (table_t*)calloc((size_t)cap, sizeof(table_t));
    struct dirent* e;
    int rc = 0;
    while ((e = readdir(d))) {
        if (!strcmp(e->d_name, ".") || !strcmp(e->d_name, "..")) continue;
        char p[1400];
        snprintf(p, sizeof p, "%s/%s", data_dir, e->d_name);
        struct stat st;
        if (stat(p, &st) || !S_ISDIR(st.st_mode)) continue;
        if (db->ntables == cap) {
            cap *= 2;
            db->tables = (table_t*)realloc(db->tables, (size_t)cap * sizeof(table_t));
            memset(db->tables + db->ntables, 0, (size_t)(cap - db->ntables) * sizeof(table_t));
        }
        rc = load_table(data_dir, e->d_name, &db->tables[db->ntables]);
        db->ntables++;
        if (rc) break;
    }
    closedir(d);
    if (rc) { orc_close(db); return rc; }
    *out = db;
    return 0;
}

void orc_close(orc_db* db) {
    if (!db) return;
    for (int t = 0; t < db->ntables; t++) {
        table_t* tb = &db->tables[t];
        for (int c = 0; c < tb->ncols; c++) {
            col_t* col = &tb->cols[c];
            for (int s = 0; s < col->nsegs; s++) {
                if (col->segs[s].data) munmap(col->segs[s].data, (size_t)col->segs[s].nbytes);
                free(col->segs[s].offsets);
            }
            free(col->segs);
        }
        free(tb->cols);
    }
    free(db->tables);
    free(db);
}

/* SegmentManager.getTable, SegmentManager.scala:89-92 */
static table_t* get_table(orc_db* db, const char* name) {
    for (int i = 0; i < db->ntables; i++)
        if (!strcmp(db->tables[i].name, name)) return &db->tables[i];
    fail(ORC_ERR_NOT_FOUND, "Table %s does not exist in SegmentManager", name);
    return NULL;
}
/* Table.getColumn, Table.scala:10-13 */
static col_t* get_column(table_t* t, const char* name) {
    for (int i = 0; i < t->ncols; i++)
        if (!strcmp(t->cols[i].name, name)) return &t->cols[i];
    fail(ORC_ERR_NOT_FOUND, "Column %s does not exist in table %s", name, t->name);
    return NULL;
}

/* getTableSegmentCount: segments of the FIRST column, SegmentManager.scala:94-99 */
int orc_table_nsegments(orc_db* db, const char* table) {
    table_t* t = get_table(db, table);
    if (!t) return ORC_ERR_NOT_FOUND;
    return t->ncols ? t->cols[0].nsegs : 0;
}
int orc_table_block_size(orc_db* db, const char* table) {
    table_t* t = get_table(db, table);
    return t ? t->block_size : ORC_ERR_NOT_FOUND;
}
int orc_table_ncols(orc_db* db, const char* table) {
    table_t* t = get_table(db, table);
    return t ? t->ncols : ORC_ERR_NOT_FOUND;
}
int orc_segment_file_id(orc_db* db, const char* table, int idx) {
    table_t* t = get_table(db, table);
    if (!t || !t->ncols || idx < 0 || idx >= t->cols[0].nsegs) return -1;
    return t->cols[0].segs[idx].file_id;
}

/* ===================================================================================== */
/* The hot path: ScanOp -> SelectOp* -> ProjectOp                                          */
/* ===================================================================================== */

typedef struct {
    col_t* col;
    int op;
    int32_t ival;     /* gt.toInt  (Select.scala:65,103,141)  */
    int8_t bval;      /* gt.toByte (Select.scala:73,111,149) */
    const char* const* strs;
    int nstrs;
    int used_idx;     /* index into the used-column list */
} pred_t;

typedef struct {
    table_t* table;
    col_t* used[64];      /* Engine.getColumns, Engine.scala:85-106 (select columns, then project columns, de-duplicated) */
    int nused;
    pred_t preds[64];
    int npreds;
    int proj_used[64];    /* project column -> used index, in select-list order (Project.scala:55-57) */
    int nproj;
    int64_t limit;
} plan_t;

typedef struct {
    uint8_t** cols;  /* per project column: growing byte buffer */
    int64_t nrows, cap;
    int64_t nmatched;
    int32_t* batch_sel; /* selected count per batch (for the reference-throw simulation) */
    int nbatches;
    uint32_t* bitmap;   /* optional: per-row selection bits in segment order */
    int64_t bitmap_rows;
    int err;
    char errmsg[256];
} seg_out_t;

static int plan_build(orc_db* db, const char* table, const orc_pred* preds, int npreds,
                      const char* const* proj, int nproj, int64_t limit, plan_t* pl) {
    memset(pl, 0, sizeof *pl);
    if (npreds > 64 || nproj > 64 || npreds < 0 || nproj < 0) return fail(ORC_ERR_INVALID_ARG, "too many predicates/columns");
    pl->table = get_table(db, table);
    if (!pl->table) return ORC_ERR_NOT_FOUND;
    pl->limit = limit;
    /* rec(query.select): columns of the Select leaves, Engine.scala:86-93 */
    for (int i = 0; i < npreds; i++) {
        col_t* c = get_column(pl->table, preds[i].col);
        if (!c) return ORC_ERR_NOT_FOUND;
        int u = -1;
        for (int k = 0; k < pl->nused; k++) if (pl->used[k] == c) u = k;
        if (u < 0) { u = pl->nused; pl->used[pl->nused++] = c; }
        pred_t* p = &pl->preds[pl->npreds++];
        p->col = c;
        p->op = preds[i].op;
        p->used_idx = u;
        p->strs = preds[i].strs;
        p->nstrs = preds[i].nstrs;
        switch (p->op) {
            case ORC_OP_GT: case ORC_OP_LT: case ORC_OP_EQ:
                /* `case _ => throw new Exception("Unsupported column vector")`, Select.scala:80,118,156 */
                if (c->ctype == ORC_COL_STRING) return fail(ORC_ERR_UNSUPPORTED, "Unsupported column vector");
                p->ival = orc_d2i(preds[i].num);
                p->bval = orc_d2b(preds[i].num);
                break;
            case ORC_OP_MATCH:
                if (c->ctype != ORC_COL_STRING) return fail(ORC_ERR_UNSUPPORTED, "Unsupported column vector"); /* Select.scala:41 */
                break;
            default: /* Select.scala:22 */
                return fail(ORC_ERR_UNSUPPORTED, "Unsupported condition");
        }
    }
    for (int i = 0; i < nproj; i++) {
        col_t* c = get_column(pl->table, proj[i]);
        if (!c) return ORC_ERR_NOT_FOUND;
        int u = -1;
        for (int k = 0; k < pl->nused; k++) if (pl->used[k] == c) u = k;
        if (u < 0) { u = pl->nused; pl->used[pl->nused++] = c; }
        pl->proj_used[pl->nproj++] = u;
    }
    return 0;
}

/* One PipelineThread.run (Engine.scala:247-262) over segment `seg`, producing the rows the
 * intended ProjectIterator would emit from that segment's batches. */
static void run_segment(const plan_t* pl, int seg, int want_bitmap, seg_out_t* o) {
    const int bs_hint = pl->table->block_size > 0 ? pl->table->block_size : 1024;
    int nused = pl->nused;
    /* ScanOp: one BlockIterator per used column, Scan.scala:25-26; hasNext follows the FIRST used
     * column (Scan.scala:72).  With no used column at all (no predicates, empty projection) the
     * reference fails on segmentIters.head; we scan the table's first column to count rows. */
    col_t* fallback_first = (nused == 0 && pl->table->ncols) ? &pl->table->cols[0] : NULL;
    col_t* firstc = nused ? pl->used[0] : fallback_first;
    if (!firstc) return;
    if (seg >= firstc->nsegs) { o->err = ORC_ERR_BAD_FORMAT; snprintf(o->errmsg, sizeof o->errmsg, "segment %d missing", seg); return; }
    int nblocks = firstc->segs[seg].noffsets - 1;
    if (nblocks < 0) nblocks = 0;
    o->batch_sel = (int32_t*)calloc((size_t)(nblocks ? nblocks : 1), sizeof(int32_t));
    o->nbatches = 0;

    int64_t cellcap = (int64_t)bs_hint + 8;
    uint8_t* cells[64];
    int64_t cellcaps[64];
    int32_t* ints[64];
    for (int u = 0; u < 64; u++) { cells[u] = NULL; ints[u] = NULL; cellcaps[u] = 0; }
    uint64_t* bits = NULL;
    int64_t bitcap = 0;
    int nloc = nused ? nused : 1;

    for (int blk = 0; blk < nblocks; blk++) {
        int64_t sizes[64];
        /* ---- ScanOp.DataVectorIterator.next, Scan.scala:28-70 ---- */
        for (int u = 0; u < nloc; u++) {
            col_t* c = nused ? pl->used[u] : fallback_first;
            if (seg >= c->nsegs || blk + 1 >= c->segs[seg].noffsets) {
                o->err = ORC_ERR_BAD_FORMAT;
                snprintf(o->errmsg, sizeof o->errmsg, "column %s: block %d of segment %d missing", c->name, blk, seg);
                goto out;
            }
            const seg_t* s = &c->segs[seg];
            /* Segment.BlockIterator.next, Segment.scala:162-170 */
            int64_t b0 = s->offsets[blk], b1 = s->offsets[blk + 1];
            if (b0 < 0 || b1 < b0 || b1 > s->nbytes) {
                o->err = ORC_ERR_BAD_FORMAT;
                snprintf(o->errmsg, sizeof o->errmsg, "column %s: bad block offsets", c->name);
                goto out;
            }
            const uint8_t* bytes = s->data + b0;
            int64_t nb = b1 - b0;
            if (c->codec == ORC_CODEC_PFOR_INT) {
                int32_t cap = (int32_t)(nb >= 4 ? ((uint32_t)bytes[0] << 24 | (uint32_t)bytes[1] << 16 | (uint32_t)bytes[2] << 8 | bytes[3]) : 0);
                if (cap < 0 || cap > (1 << 26)) { o->err = ORC_ERR_BAD_FORMAT; snprintf(o->errmsg, sizeof o->errmsg, "PFOR block count %d", cap); goto out; }
                if (cap + 32 > cellcaps[u]) {
                    cellcaps[u] = cap + 32;
                    ints[u] = (int32_t*)realloc(ints[u], (size_t)cellcaps[u] * 4);
                }
                int32_t n = orc_pfor_decode_block(bytes, nb, ints[u], (int32_t)cellcaps[u]);
                if (n < 0) { o->err = ORC_ERR_BAD_FORMAT; snprintf(o->errmsg, sizeof o->errmsg, "column %s: bad PFOR block", c->name); goto out; }
                sizes[u] = n;
            } else {
                int64_t n = (nb + c->width - 1) / c->width;
                if (n + 1 > cellcaps[u]) {
                    cellcaps[u] = n + cellcap;
                    cells[u] = (uint8_t*)realloc(cells[u], (size_t)cellcaps[u] * (size_t)c->width);
                    if (c->ctype == ORC_COL_INT) ints[u] = (int32_t*)realloc(ints[u], (size_t)cellcaps[u] * 4);
                }
                sizes[u] = orc_dense_decode(bytes, nb, c->width, cells[u]);
                if (c->ctype == ORC_COL_INT) /* IntType.bytesToValue, DataType.scala:48 */
                    for (int64_t i = 0; i < sizes[u]; i++) ints[u][i] = orc_bytes_to_int(cells[u] + 4 * i);
            }
        }
        int64_t vec_size = sizes[0]; /* Scan.scala:55: based on first column */
        /* bitSet = all ones, Scan.scala:56-57 */
        int64_t nw = (vec_size + 63) / 64;
        if (nw > bitcap) { bitcap = nw + 16; bits = (uint64_t*)realloc(bits, (size_t)bitcap * 8); }
        for (int64_t w = 0; w < nw; w++) bits[w] = ~0ull;
        if (vec_size & 63) bits[nw - 1] = (1ull << (vec_size & 63)) - 1;

        /* ---- SelectOp chain, left to right; each iterator scans ALL positions and only removes
         *      bits (Select.scala:36-39, 67-78, 105-116, 143-154) => conjunction, Engine.scala:237-245 ---- */
        for (int pi = 0; pi < pl->npreds; pi++) {
            const pred_t* p = &pl->preds[pi];
            int u = p->used_idx;
            if (sizes[u] < vec_size) { /* data(x) would throw ArrayIndexOutOfBounds */
                o->err = ORC_ERR_BAD_FORMAT;
                snprintf(o->errmsg, sizeof o->errmsg, "column %s shorter than batch", p->col->name);
                goto out;
            }
            /* `for (x <- 0 until vec.size) if (!(data(x) OP c)) vec.selected.remove(x)`: a bit survives iff the
             * predicate holds.  Evaluated one 64-row bitmap word at a time (same result, no per-row branch). */
#define ORC_PASS(TYPE, DATA, EXPR)                                                          \
    do {                                                                                    \
        const TYPE* d = (const TYPE*)(DATA);                                                \
        for (int64_t w = 0; w < nw; w++) {                                                  \
            const int64_t x0 = w * 64, xe = x0 + 64 < vec_size ? x0 + 64 : vec_size;        \
            uint64_t m = 0;                                                                 \
            for (int64_t x = x0; x < xe; x++) m |= (uint64_t)(EXPR) << (x - x0);            \
            bits[w] &= m;                                                                   \
        }                                                                                   \
    } while (0)
            if (p->col->ctype == ORC_COL_INT) {
                const int32_t c = p->ival;
                if (p->op == ORC_OP_GT) ORC_PASS(int32_t, ints[u], d[x] > c);
                else if (p->op == ORC_OP_LT) ORC_PASS(int32_t, ints[u], d[x] < c);
                else ORC_PASS(int32_t, ints[u], d[x] == c);
            } else if (p->col->ctype == ORC_COL_TINYINT) {
                const int8_t c = p->bval;
                if (p->op == ORC_OP_GT) ORC_PASS(int8_t, cells[u], d[x] > c);
                else if (p->op == ORC_OP_LT) ORC_PASS(int8_t, cells[u], d[x] < c);
                else ORC_PASS(int8_t, cells[u], d[x] == c);
#undef ORC_PASS
            } else {
                /* matchValues.contains(data(x)), Select.scala:37 — String equality; with ASCII cells and
                 * literals this is "same length k and same bytes". */
                int k = p->col->width;
                const uint8_t* d = cells[u];
                for (int64_t x = 0; x < vec_size; x++) {
                    int hit = 0;
                    for (int s = 0; s < p->nstrs && !hit; s++)
                        hit = (int)strlen(p->strs[s]) == k && !memcmp(d + x * k, p->strs[s], (size_t)k);
                    if (!hit) bits[x >> 6] &= ~(1ull << (x & 63));
                }
            }
        }
        int32_t selcount = 0;
        for (int64_t w = 0; w < nw; w++) selcount += __builtin_popcountll(bits[w]);
        o->batch_sel[o->nbatches++] = selcount;
        o->nmatched += selcount;

        if (want_bitmap) {
            int64_t need = o->bitmap_rows + vec_size;
            o->bitmap = (uint32_t*)realloc(o->bitmap, (size_t)((need + 31) / 32 + 1) * 4);
            for (int64_t x = 0; x < vec_size; x++) {
                int64_t r = o->bitmap_rows + x;
                if ((r & 31) == 0) o->bitmap[r >> 5] = 0;
                if (bits[x >> 6] >> (x & 63) & 1) o->bitmap[r >> 5] |= 1u << (r & 31);
            }
            o->bitmap_rows = need;
            continue;
        }

        /* ---- ProjectOp.ProjectIterator, Project.scala:37-64: ascending selected positions, cells in
         *      select-list order.  Intended semantics: batches with no selected row are skipped (B2). ---- */
        if (selcount == 0) continue;
        for (int pc = 0; pc < pl->nproj; pc++)
            if (sizes[pl->proj_used[pc]] < vec_size) {
                o->err = ORC_ERR_BAD_FORMAT;
                snprintf(o->errmsg, sizeof o->errmsg, "projected column shorter than batch");
                goto out;
            }
        if (o->nrows + selcount > o->cap) {
            o->cap = (o->nrows + selcount) * 2 + 1024;
            for (int pc = 0; pc < pl->nproj; pc++)
                o->cols[pc] = (uint8_t*)realloc(o->cols[pc], (size_t)o->cap * (size_t)pl->used[pl->proj_used[pc]]->width);
        }
        for (int64_t w = 0; w < nw; w++) {
            uint64_t m = bits[w];
            while (m) {
                int64_t pos = w * 64 + __builtin_ctzll(m);
                m &= m - 1;
                for (int pc = 0; pc < pl->nproj; pc++) {
                    int u = pl->proj_used[pc];
                    col_t* c = pl->used[u];
                    if (c->ctype == ORC_COL_INT) ((int32_t*)o->cols[pc])[o->nrows] = ints[u][pos];
                    else if (c->ctype == ORC_COL_TINYINT) o->cols[pc][o->nrows] = cells[u][pos];
                    else memcpy(o->cols[pc] + o->nrows * c->width, cells[u] + pos * c->width, (size_t)c->width);
                }
                o->nrows++;
            }
        }
        if (pl->limit > 0 && o->nrows >= pl->limit) break; /* consumer stops asking, Project.scala:73-77 */
    }
out:
    for (int u = 0; u < 64; u++) { free(cells[u]); free(ints[u]); }
    free(bits);
}

struct orc_result {
    int ncols;
    int ctype[64], width[64];
    uint8_t* cols[64];
    int64_t nrows, nmatched;
    int ref_throw;
    int64_t ref_rows;
};

typedef struct {
    const plan_t* pl;
    seg_out_t* outs;
    int seg_begin, seg_end;
    int want_bitmap;
    int next; /* atomic */
} pool_t;

static void* worker(void* arg) {
    pool_t* p = (pool_t*)arg;
    for (;;) {
        int i = __atomic_fetch_add(&p->next, 1, __ATOMIC_RELAXED);
        if (p->seg_begin + i >= p->seg_end) break;
        run_segment(p->pl, p->seg_begin + i, p->want_bitmap, &p->outs[i]);
    }
    return NULL;
}

static int run_all(const plan_t* pl, int nthreads, int seg_begin, int seg_end, int want_bitmap, seg_out_t** outs_p, int* nseg_p) {
    table_t* t = pl->table;
    int nseg_total = t->ncols ? t->cols[0].nsegs : 0; /* Engine.scala:161 */
    if (seg_end < 0 || seg_end > nseg_total) seg_end = nseg_total;
    if (seg_begin < 0) seg_begin = 0;
    if (seg_begin > seg_end) seg_begin = seg_end;
    int nseg = seg_end - seg_begin;
    seg_out_t* outs = (seg_out_t*)calloc((size_t)(nseg ? nseg : 1), sizeof(seg_out_t));
    for (int i = 0; i < nseg; i++) outs[i].cols = (uint8_t**)calloc(64, sizeof(uint8_t*));
    pool_t pool = {pl, outs, seg_begin, seg_end, want_bitmap, 0};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nseg) nthreads = nseg > 0 ? nseg : 1;
    if (nthreads == 1) {
        /* --cpu-count 1: Futures run in index order; with a LIMIT the consumer stops once satisfied
         * (the remaining producers would block on the bounded queue). */
        int64_t have = 0;
        for (int i = 0; i < nseg; i++) {
            if (!want_bitmap && pl->limit > 0 && have >= pl->limit) break;
            run_segment(pl, seg_begin + i, want_bitmap, &outs[i]);
            have += outs[i].nrows;
        }
    } else {
        pthread_t* th = (pthread_t*)malloc((size_t)nthreads * sizeof(pthread_t));
        for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, worker, &pool);
        for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
        free(th);
    }
    *outs_p = outs;
    *nseg_p = nseg;
    return 0;
}

static void free_outs(seg_out_t* outs, int nseg) {
    for (int i = 0; i < nseg; i++) {
        for (int c = 0; c < 64; c++) free(outs[i].cols[c]);
        free(outs[i].cols);
        free(outs[i].batch_sel);
        free(outs[i].bitmap);
    }
    free(outs);
}

int orc_query(orc_db* db, const char* table, const orc_pred* preds, int npreds,
              const char* const* proj_cols, int nproj, int64_t limit, int nthreads,
              int seg_begin, int seg_end, orc_result** out) {
    plan_t* pl = (plan_t*)malloc(sizeof *pl);
    int rc = plan_build(db, table, preds, npreds, proj_cols, nproj, limit, pl);
    if (rc) { free(pl); return rc; }
    seg_out_t* outs;
    int nseg;
    run_all(pl, nthreads, seg_begin, seg_end, 0, &outs, &nseg);
    for (int i = 0; i < nseg; i++)
        if (outs[i].err) {
            rc = fail(outs[i].err, "%s", outs[i].errmsg);
            free_outs(outs, nseg);
            free(pl);
            return rc;
        }
    orc_result* r = (orc_result*)calloc(1, sizeof *r);
    r->ncols = pl->nproj;
    int64_t total = 0;
    for (int i = 0; i < nseg; i++) { total += outs[i].nrows; r->nmatched += outs[i].nmatched; }
    if (limit > 0 && total > limit) total = limit; /* Project.scala:73-77 */
    r->nrows = total;
    for (int pc = 0; pc < pl->nproj; pc++) {
        col_t* c = pl->used[pl->proj_used[pc]];
        r->ctype[pc] = c->ctype;
        r->width[pc] = c->width;
        r->cols[pc] = (uint8_t*)malloc((size_t)(total ? total : 1) * (size_t)c->width);
        /* fan-in in canonical segment order */
        int64_t at = 0;
        for (int i = 0; i < nseg && at < total; i++) {
            int64_t take = outs[i].nrows;
            if (at + take > total) take = total - at;
            if (take > 0) memcpy(r->cols[pc] + at * c->width, outs[i].cols[pc], (size_t)take * (size_t)c->width);
            at += take;
        }
    }
    /* Where would the unmodified reference have thrown with --cpu-count 1?
     * Only segment 0 matters: its end-of-segment marker is the first None in the queue. */
    {
        int64_t emitted = 0;
        r->ref_throw = 1;
        if (nseg > 0 && seg_begin == 0) {
            for (int b = 0; b < outs[0].nbatches; b++) {
                int32_t sel = outs[0].batch_sel[b];
                if (sel == 0) { r->ref_throw = 2; break; }
                int64_t take = sel;
                if (limit > 0 && emitted + take > limit) take = limit - emitted;
                emitted += take;
                if (limit > 0 && emitted >= limit) { r->ref_throw = 0; break; }
            }
        }
        r->ref_rows = emitted;
    }
    free_outs(outs, nseg);
    free(pl);
    *out = r;
    return 0;
}

int64_t orc_result_nrows(const orc_result* r) { return r->nrows; }
int64_t orc_result_nmatched(const orc_result* r) { return r->nmatched; }
int orc_result_ncols(const orc_result* r) { return r->ncols; }
int orc_result_col_type(const orc_result* r, int c) { return r->ctype[c]; }
int orc_result_col_width(const orc_result* r, int c) { return r->width[c]; }
const void* orc_result_col_data(const orc_result* r, int c) { return r->cols[c]; }
int orc_result_ref_throw(const orc_result* r, int64_t* rows_before) {
    if (rows_before) *rows_before = r->ref_rows;
    return r->ref_throw;
}

/* Row.toString = xs.mkString("Row(", ",", ")"), Record.scala:13; Int/Byte print in decimal. */
int orc_result_format_row(const orc_result* r, int64_t row, char* buf, size_t buflen) {
    if (row < 0 || row >= r->nrows) return -1;
    size_t n = 0;
#define PUT(...) do { int k_ = snprintf(buf + n, n < buflen ? buflen - n : 0, __VA_ARGS__); if (k_ < 0) return -1; n += (size_t)k_; } while (0)
    PUT("Row(");
    for (int c = 0; c < r->ncols; c++) {
        if (c) PUT(",");
        if (r->ctype[c] == ORC_COL_INT) PUT("%d", ((const int32_t*)r->cols[c])[row]);
        else if (r->ctype[c] == ORC_COL_TINYINT) PUT("%d", (int)((const int8_t*)r->cols[c])[row]);
        else PUT("%.*s", r->width[c], (const char*)r->cols[c] + row * r->width[c]);
    }
    PUT(")");
#undef PUT
    return (int)n;
}

void orc_result_free(orc_result* r) {
    if (!r) return;
    for (int c = 0; c < 64; c++) free(r->cols[c]);
    free(r);
}

int orc_filter_bitmap(orc_db* db, const char* table, const orc_pred* preds, int npreds,
                      int seg_begin, int seg_end, uint32_t** words, int64_t* nwords, int64_t* nselected) {
    plan_t* pl = (plan_t*)malloc(sizeof *pl);
    int rc = plan_build(db, table, preds, npreds, NULL, 0, 0, pl);
    if (rc) { free(pl); return rc; }
    seg_out_t* outs;
    int nseg;
    run_all(pl, 1, seg_begin, seg_end, 1, &outs, &nseg);
    int64_t rows = 0, sel = 0;
    for (int i = 0; i < nseg; i++) {
        if (outs[i].err) { rc = fail(outs[i].err, "%s", outs[i].errmsg); free_outs(outs, nseg); free(pl); return rc; }
        rows += outs[i].bitmap_rows;
        sel += outs[i].nmatched;
    }
    uint32_t* w = (uint32_t*)calloc((size_t)((rows + 31) / 32 + 1), 4);
    int64_t at = 0;
    for (int i = 0; i < nseg; i++)
        for (int64_t x = 0; x < outs[i].bitmap_rows; x++, at++)
            if (outs[i].bitmap[x >> 5] >> (x & 31) & 1) w[at >> 5] |= 1u << (at & 31);
    free_outs(outs, nseg);
    free(pl);
    *words = w;
    *nwords = (rows + 31) / 32;
    *nselected = sel;
    return 0;
}

/* Rows (decoded values of the first column) in canonical segments [seg_begin, seg_end). */
int64_t orc_table_nrows(orc_db* db, const char* table, int seg_begin, int seg_end) {
    table_t* t = get_table(db, table);
    if (!t) return ORC_ERR_NOT_FOUND;
    if (!t->ncols) return 0;
    col_t* c = &t->cols[0];
    if (seg_end < 0 || seg_end > c->nsegs) seg_end = c->nsegs;
    int64_t rows = 0;
    for (int s = seg_begin < 0 ? 0 : seg_begin; s < seg_end; s++) {
        const seg_t* sg = &c->segs[s];
        for (int b = 0; b + 1 < sg->noffsets; b++) {
            int64_t nb = sg->offsets[b + 1] - sg->offsets[b];
            if (c->codec == ORC_CODEC_PFOR_INT) {
                const uint8_t* p = sg->data + sg->offsets[b];
                rows += nb >= 4 ? (int64_t)((uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]) : 0;
            } else {
                rows += (nb + c->width - 1) / c->width;
            }
        }
    }
    return rows;
}

/* ===================================================================================== */
/* Writer side (test fixtures and the reference arm's tables need no product code)         */
/*   SegmentWriter      Segment.scala:70-152   lazy block flush: a block is written when the */
/*                      (blockSize+1)-th value arrives, close() writes what is buffered      */
/*   LoaderCli loop     LoaderCli.scala:142-148 roll to a new segment when remaining == 0    */
/*                      BEFORE a write => a full segment holds S*B + 1 rows = S full blocks  */
/*                      + one 1-row tail block                               */
/*   TableIO.store      Table.scala:27-35,50-59 compact JSON                                 */
/* ===================================================================================== */
static int mkdir_p(const char* path) {
    char tmp[4096];
    snprintf(tmp, sizeof tmp, "%s", path);
    for (char* q = tmp + 1; *q; q++)
        if (*q == '/') {
            *q = 0;
            if (mkdir(tmp, 0777) && errno != EEXIST) return fail(ORC_ERR_IO, "mkdir %s: %s", tmp, strerror(errno));
            *q = '/';
        }
    if (mkdir(tmp, 0777) && errno != EEXIST) return fail(ORC_ERR_IO, "mkdir %s: %s", tmp, strerror(errno));
    return 0;
}

/* One segment of one column: `nrows` cells (little-endian, `width` bytes each) cut into blocks of block_size rows (the
 * last one shorter), each block encoded by the column's codec; <col>_<id>.dat + <col>_<id>.meta (Segment.scala:75-76,
 * 41-45, 120-121, 144-151). */
static int write_segment(const char* tdir, const char* col, int codec, int width, int seg_id, const uint8_t* cells,
                         int64_t nrows, int block_size) {
    char path[4096];
    snprintf(path, sizeof path, "%s/%s_%d.dat", tdir, col, seg_id);
    FILE* f = fopen(path, "wb");
    if (!f) return fail(ORC_ERR_IO, "open %s: %s", path, strerror(errno));
    const int64_t nblocks = (nrows + block_size - 1) / block_size;
    int32_t* offs = (int32_t*)malloc((size_t)(nblocks + 1) * 4);
    uint8_t* enc = (uint8_t*)malloc((size_t)block_size * 4 + (size_t)block_size / 8 + 512);
    int32_t* vals = (int32_t*)malloc((size_t)block_size * 4);
    int rc = 0;
    int64_t at = 0;
    if (!offs || !enc || !vals) rc = fail(ORC_ERR_OOM, "out of memory");
    if (!rc) offs[0] = 0;
    for (int64_t b = 0; !rc && b < nblocks; b++) {
        const int64_t r0 = b * block_size, n = nrows - r0 < block_size ? nrows - r0 : block_size;
        const uint8_t* src = cells + r0 * width;
        int64_t len = n * width;
        if (codec == ORC_CODEC_PFOR_INT) { /* SegmentWriter.flush: codec.encode(buffer), Segment.scala:120 */
            for (int64_t i = 0; i < n; i++) vals[i] = orc_bytes_to_int(src + 4 * i);
            len = orc_pfor_encode_block(vals, (int32_t)n, enc, (int64_t)block_size * 4 + block_size / 8 + 512);
            if (len < 0) { rc = fail(ORC_ERR_OOM, "sorted-int encode failed"); break; }
            src = enc;
        }
        if (fwrite(src, 1, (size_t)len, f) != (size_t)len) rc = fail(ORC_ERR_IO, "short write on %s", path);
        at += len;
        if (at > INT32_MAX) rc = fail(ORC_ERR_UNSUPPORTED, "segment %s exceeds 2 GiB (block offsets are Int, Segment.scala:33)", path);
        offs[b + 1] = (int32_t)at;
    }
    fclose(f);
    if (!rc) {
        snprintf(path, sizeof path, "%s/%s_%d.meta", tdir, col, seg_id);
        FILE* m = fopen(path, "wb");
        if (!m) rc = fail(ORC_ERR_IO, "open %s: %s", path, strerror(errno));
        else {
            fputs("{\"blockOffset\":[", m);
            for (int64_t b = 0; b <= nblocks; b++) fprintf(m, b ? ",%d" : "%d", offs[b]);
            fputs("]}", m);
            fclose(m);
        }
    }
    free(offs);
    free(enc);
    free(vals);
    return rc;
}

/* TableIO.clear + store: remove the regular files of the table directory, write _table.meta. */
int orc_write_table_meta(const char* data_dir, const char* table, const char* meta_json) {
    char tdir[4096], path[4400];
    snprintf(tdir, sizeof tdir, "%s/%s", data_dir, table);
    int rc = mkdir_p(tdir);
    if (rc) return rc;
    DIR* d = opendir(tdir);
    if (d) {
        struct dirent* e;
        while ((e = readdir(d))) {
            if (e->d_name[0] == '.' && (!e->d_name[1] || (e->d_name[1] == '.' && !e->d_name[2]))) continue;
            snprintf(path, sizeof path, "%s/%s", tdir, e->d_name);
            struct stat st;
            if (!stat(path, &st) && S_ISREG(st.st_mode)) remove(path);
        }
        closedir(d);
    }
    snprintf(path, sizeof path, "%s/_table.meta", tdir);
    FILE* f = fopen(path, "wb");
    if (!f) return fail(ORC_ERR_IO, "open %s: %s", path, strerror(errno));
    fputs(meta_json, f);
    fclose(f);
    return 0;
}

/* A whole column given as one array of cells: segments 0, 1, ... in the loader's layout. */
int orc_write_column(const char* data_dir, const char* table, const char* col, int codec, int width, const void* cells,
                     int64_t nrows, int block_size, int segment_size) {
    if (block_size <= 0 || segment_size <= 0 || width <= 0 || nrows < 0) return fail(ORC_ERR_INVALID_ARG, "orc_write_column: bad sizes");
    char tdir[4096];
    snprintf(tdir, sizeof tdir, "%s/%s", data_dir, table);
    const int64_t per_seg = (int64_t)block_size * segment_size + 1;
    int64_t nseg = (nrows + per_seg - 1) / per_seg;
    if (nseg == 0) nseg = 1; /* the writer of segment 0 exists before the first row arrives (LoaderCli.scala:118-127) */
    for (int64_t s = 0; s < nseg; s++) {
        const int64_t r0 = s * per_seg, n = nrows - r0 < per_seg ? nrows - r0 : per_seg;
        int rc = write_segment(tdir, col, codec, width, (int)s, (const uint8_t*)cells + r0 * width, n > 0 ? n : 0, block_size);
        if (rc) return rc;
    }
    return 0;
}

/* The synthetic tables of BASELINE.md: id = row index, age uniform [0,100), state uniform over 51 two-letter codes;
 * counter-based PRNG (splitmix64 of (seed ^ column) << 32 ^ row), seed 42.  Numeric segment ids [seg_begin, seg_end). */
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
void orc_synth_row(int64_t row, int32_t* id, int8_t* age, char state[2]) {
    *id = (int32_t)row;
    *age = (int8_t)((splitmix64(((42ull ^ 1) << 32) ^ (uint64_t)row) >> 32) % 100u);
    const char* s = kStates[(splitmix64(((42ull ^ 2) << 32) ^ (uint64_t)row) >> 32) % 51u];
    state[0] = s[0];
    state[1] = s[1];
}

typedef struct {
    const char* tdir;
    int64_t nrows, per_seg;
    int block_size, id_codec, seg_end;
    int* next;
    int* rc;
    pthread_mutex_t* mu;
} synth_job_t;

static void* synth_worker(void* arg) {
    synth_job_t* j = (synth_job_t*)arg;
    uint8_t* ids = (uint8_t*)malloc((size_t)j->per_seg * 4);
    uint8_t* ages = (uint8_t*)malloc((size_t)j->per_seg);
    uint8_t* states = (uint8_t*)malloc((size_t)j->per_seg * 2);
    for (;;) {
        pthread_mutex_lock(j->mu);
        const int seg = (*j->next)++;
        const int stop = *j->rc != 0;
        pthread_mutex_unlock(j->mu);
        if (seg >= j->seg_end || stop) break;
        int rc = (!ids || !ages || !states) ? ORC_ERR_OOM : 0;
        const int64_t r0 = (int64_t)seg * j->per_seg, n = j->nrows - r0 < j->per_seg ? j->nrows - r0 : j->per_seg;
        for (int64_t i = 0; !rc && i < n; i++) {
            int32_t id;
            orc_synth_row(r0 + i, &id, (int8_t*)&ages[i], (char*)&states[2 * i]);
            orc_int_to_bytes(id, ids + 4 * i);
        }
        if (!rc) rc = write_segment(j->tdir, "id", j->id_codec, 4, seg, ids, n, j->block_size);
        if (!rc) rc = write_segment(j->tdir, "state", ORC_CODEC_DENSE_STRING, 2, seg, states, n, j->block_size);
        if (!rc) rc = write_segment(j->tdir, "age", ORC_CODEC_DENSE_TINYINT, 1, seg, ages, n, j->block_size);
        if (rc) {
            pthread_mutex_lock(j->mu);
            if (!*j->rc) *j->rc = rc;
            pthread_mutex_unlock(j->mu);
        }
    }
    free(ids);
    free(ages);
    free(states);
    return NULL;
}

int orc_synth_write(const char* data_dir, const char* table, int64_t nrows, int32_t block_size, int32_t segment_size,
                    int32_t id_codec, int32_t seg_begin, int32_t seg_end, int write_table_meta, int nthreads) {
    if (nrows < 0 || block_size <= 0 || segment_size <= 0) return fail(ORC_ERR_INVALID_ARG, "orc_synth_write: bad sizes");
    char tdir[4096];
    snprintf(tdir, sizeof tdir, "%s/%s", data_dir, table);
    if (write_table_meta) {
        char meta[1024];
        snprintf(meta, sizeof meta,
                 "{\"name\":\"%s\",\"columns\":[{\"name\":\"id\",\"columnType\":\"INT\",\"codec\":\"%s\",\"dtypeAttrs\":{}},"
                 "{\"name\":\"state\",\"columnType\":\"STRING\",\"codec\":\"DENSE_STRING\",\"dtypeAttrs\":{\"size\":\"2\"}},"
                 "{\"name\":\"age\",\"columnType\":\"TINYINT\",\"codec\":\"DENSE_TINYINT\",\"dtypeAttrs\":{}}],\"blockSize\":%d}",
                 table, id_codec == ORC_CODEC_PFOR_INT ? "PFOR_INT" : "DENSE_INT", block_size);
        int rc = orc_write_table_meta(data_dir, table, meta);
        if (rc) return rc;
    }
    const int64_t per_seg = (int64_t)block_size * segment_size + 1;
    const int64_t nseg = (nrows + per_seg - 1) / per_seg;
    if (seg_end < 0 || seg_end > nseg) seg_end = (int32_t)nseg;
    if (seg_begin < 0) seg_begin = 0;
    if (seg_end <= seg_begin) return 0;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > seg_end - seg_begin) nthreads = seg_end - seg_begin;
    if (nthreads > 256) nthreads = 256;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    int next = seg_begin, rc = 0;
    synth_job_t job = {tdir, nrows, per_seg, block_size, id_codec, seg_end, &next, &rc, &mu};
    pthread_t th[256];
    for (int i = 1; i < nthreads; i++) pthread_create(&th[i], NULL, synth_worker, &job);
    synth_worker(&job);
    for (int i = 1; i < nthreads; i++) pthread_join(th[i], NULL);
    return rc ? fail(rc, "orc_synth_write failed (status %d)", rc) : 0;
}

/* ===================================================================================== */
/* Engine.execute, ProjectAgg branch (Engine.scala:130-156, 200-232)                        */
/*   ProjectAggOp.ProjectAggIterator.runAggs   ProjectAggregate.scala:160-222                */
/*     per batch, per SELECTED row: key = group values joined by "_" (:152,166-170);        */
/*     resultMap.getOrElseUpdate(key, fresh aggregators) - a LinkedHashMap, so groups keep  */
/*     the order of their first row (:128,178); CountAggr.add counts the row (:21-24),      */
/*     Min/MaxDoubleAggr.add(value.toDouble) (:37-59,176-178,190-192)                       */
/*   ProjectAggregateQueueOp.init              ProjectAggregateQueue.scala:21-45            */
/*     merges the per-segment maps in arrival order (one worker: segment order) with        */
/*     combine(): counts add, min/max of min/max                                            */
/* Net effect with one worker: one group per distinct key in order of first appearance in   */
/* canonical row order.  Output row of the reference = the aggregators' repr only           */
/* (ProjectAggregateQueue.scala:48-50) - in the iteration order of a mutable.HashMap keyed  */
/* by alias; INTENDED order (used here): the select list's.  Min on a STRING column is      */
/* resolved to MaxStringAggr (Engine.scala:147) and Min/MaxStringAggr are not what this     */
/* restatement covers: ORC_ERR_UNSUPPORTED, as are Sum / Avg (Engine.scala:153).            */
/* Result columns: the group columns, then per aggregate COUNT -> int64, MIN/MAX -> double. */
/* ===================================================================================== */
enum { ORC_COL_COUNT = 3, ORC_COL_DOUBLE = 4 };

int orc_query_agg(orc_db* db, const char* table, const orc_pred* preds, int npreds, const orc_agg* aggs, int naggs,
                  const char* const* group_cols, int ngroup, int nthreads, int seg_begin, int seg_end, orc_result** out) {
    if (naggs < 1 || naggs > 16 || ngroup < 0 || ngroup > 8) return fail(ORC_ERR_INVALID_ARG, "orc_query_agg: bad counts");
    table_t* t = get_table(db, table);
    if (!t) return fail(ORC_ERR_NOT_FOUND, "Table %s does not exist in SegmentManager", table);
    /* usedColumns = aggs' columns ++ group columns (Engine.scala:97-101): project them all, unlimited */
    const char* proj[24];
    for (int g = 0; g < ngroup; g++) proj[g] = group_cols[g];
    for (int a = 0; a < naggs; a++) {
        proj[ngroup + a] = aggs[a].col;
        if (aggs[a].op != ORC_AGG_COUNT && aggs[a].op != ORC_AGG_MIN && aggs[a].op != ORC_AGG_MAX)
            return fail(ORC_ERR_UNSUPPORTED, "Unknown Aggregate type");
        col_t* c = get_column(t, aggs[a].col);
        if (c && c->ctype == ORC_COL_STRING && aggs[a].op != ORC_AGG_COUNT) return fail(ORC_ERR_UNSUPPORTED, "min / max on a STRING column");
    }
    orc_result* rows = NULL;
    int rc = orc_query(db, table, preds, npreds, proj, ngroup + naggs, 0, nthreads, seg_begin, seg_end, &rows);
    if (rc) return rc;
    const int64_t n = rows->nrows;
    int keyw = 0;
    for (int g = 0; g < ngroup; g++) keyw += rows->width[g];
    /* open-addressing map: key bytes -> group index, groups appended in first-appearance order */
    int64_t cap = 1024, ngroups = 0;
    int64_t* slot = (int64_t*)malloc((size_t)cap * 8);
    for (int64_t i = 0; i < cap; i++) slot[i] = -1;
    uint8_t* gkeys = NULL;
    int64_t* cnt = NULL;
    double* val = NULL; /* [group][agg] */
    int64_t gcap = 0;
    uint8_t key[256];
    for (int64_t r = 0; r < n; r++) {
        int at = 0;
        for (int g = 0; g < ngroup; g++) {
            memcpy(key + at, rows->cols[g] + r * rows->width[g], (size_t)rows->width[g]);
            at += rows->width[g];
        }
        uint64_t h = 1469598103934665603ull;
        for (int i = 0; i < keyw; i++) h = (h ^ key[i]) * 1099511628211ull;
        int64_t gi = -1;
        for (int64_t p = (int64_t)(h & (uint64_t)(cap - 1));; p = (p + 1) & (cap - 1)) {
            if (slot[p] < 0) {
                if (ngroups == gcap) {
                    gcap = gcap ? gcap * 2 : 256;
                    gkeys = (uint8_t*)realloc(gkeys, (size_t)gcap * (size_t)(keyw ? keyw : 1));
                    cnt = (int64_t*)realloc(cnt, (size_t)gcap * (size_t)naggs * 8);
                    val = (double*)realloc(val, (size_t)gcap * (size_t)naggs * 8);
                }
                gi = ngroups++;
                slot[p] = gi;
                memcpy(gkeys + gi * keyw, key, (size_t)keyw);
                for (int a = 0; a < naggs; a++) { /* aggregator.make: fresh state (ProjectAggregate.scala:21,38,50) */
                    cnt[gi * naggs + a] = 0;
                    val[gi * naggs + a] = aggs[a].op == ORC_AGG_MAX ? -1.7976931348623157e308 : 1.7976931348623157e308; /* Double.MinValue / MaxValue */
                }
                break;
            }
            if (!memcmp(gkeys + slot[p] * keyw, key, (size_t)keyw)) { gi = slot[p]; break; }
        }
        for (int a = 0; a < naggs; a++) {
            const int c = ngroup + a;
            if (aggs[a].op == ORC_AGG_COUNT) { cnt[gi * naggs + a]++; continue; }
            const double v = rows->ctype[c] == ORC_COL_INT ? (double)orc_bytes_to_int(rows->cols[c] + r * 4) : (double)(int8_t)rows->cols[c][r];
            double* m = &val[gi * naggs + a];
            if (aggs[a].op == ORC_AGG_MAX ? v > *m : v < *m) *m = v;
        }
        if (ngroups * 2 > cap) { /* grow the index */
            cap *= 2;
            slot = (int64_t*)realloc(slot, (size_t)cap * 8);
            for (int64_t i = 0; i < cap; i++) slot[i] = -1;
            for (int64_t g2 = 0; g2 < ngroups; g2++) {
                uint64_t h2 = 1469598103934665603ull;
                for (int i = 0; i < keyw; i++) h2 = (h2 ^ gkeys[g2 * keyw + i]) * 1099511628211ull;
                int64_t p = (int64_t)(h2 & (uint64_t)(cap - 1));
                while (slot[p] >= 0) p = (p + 1) & (cap - 1);
                slot[p] = g2;
            }
        }
    }
    orc_result* res = (orc_result*)calloc(1, sizeof *res);
    res->ncols = ngroup + naggs;
    res->nrows = ngroups;
    res->nmatched = n;
    int at = 0;
    for (int g = 0; g < ngroup; g++) {
        res->ctype[g] = rows->ctype[g];
        res->width[g] = rows->width[g];
        res->cols[g] = (uint8_t*)malloc((size_t)(ngroups ? ngroups : 1) * (size_t)rows->width[g]);
        for (int64_t i = 0; i < ngroups; i++) memcpy(res->cols[g] + i * rows->width[g], gkeys + i * keyw + at, (size_t)rows->width[g]);
        at += rows->width[g];
    }
    for (int a = 0; a < naggs; a++) {
        const int c = ngroup + a;
        res->ctype[c] = aggs[a].op == ORC_AGG_COUNT ? ORC_COL_COUNT : ORC_COL_DOUBLE;
        res->width[c] = 8;
        res->cols[c] = (uint8_t*)malloc((size_t)(ngroups ? ngroups : 1) * 8);
        for (int64_t i = 0; i < ngroups; i++) {
            if (aggs[a].op == ORC_AGG_COUNT) memcpy(res->cols[c] + i * 8, &cnt[i * naggs + a], 8);
            else memcpy(res->cols[c] + i * 8, &val[i * naggs + a], 8);
        }
    }
    free(slot);
    free(gkeys);
    free(cnt);
    free(val);
    orc_result_free(rows);
    *out = res;
    return 0;
}
