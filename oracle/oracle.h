/*
 * oracle.h — CPU restatement of immutable3's scan / filter / project path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under immutable3_b200/ may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use
 * it, and only as the checker or the CPU baseline — never as the product path.
 *
 * Parity status: the reference ships no tests, fixtures or golden vectors and cannot be run here
 * (no JVM), so DENSE_* decode, predicates, Project/LIMIT and the segment layout are pinned by
 * restatement of the cited Scala plus the hand-derived known-answer vectors of SURVEY.md §8c
 * (tests/test_oracle_kat.py).  The sorted-integer codec delegates to JavaFastPFOR 0.1.10
 * (me.lemire.integercompression:JavaFastPFOR, project/Dependencies.scala:4), whose source is NOT
 * under /root/reference: its published algorithm (IntegratedIntCompressor =
 * SkippableIntegratedComposition(IntegratedBinaryPacking, IntegratedVariableByte)) is restated
 * from the upstream library's documentation — byte compatibility with real JavaFastPFOR output is
 * "PARITY UNPINNED".
 *
 * All citations are relative to the reference checkout.
 */
#ifndef IMM3_ORACLE_H
#define IMM3_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_COL_INT = 0, ORC_COL_TINYINT = 1, ORC_COL_STRING = 2 };                  /* Column.scala:13-16 */
enum { ORC_CODEC_PFOR_INT = 0, ORC_CODEC_DENSE_INT = 1, ORC_CODEC_DENSE_TINYINT = 2,
       ORC_CODEC_DENSE_STRING = 3 };                                                /* Codec.scala:20-23  */
enum { ORC_OP_GT = 1, ORC_OP_LT = 2, ORC_OP_EQ = 3, ORC_OP_MATCH = 4, ORC_OP_NOTMATCH = 5, ORC_OP_NOOP = 6 };
enum { ORC_OK = 0, ORC_ERR_NOT_FOUND = -1, ORC_ERR_UNSUPPORTED = -2, ORC_ERR_BAD_FORMAT = -3,
       ORC_ERR_OOM = -5, ORC_ERR_INVALID_ARG = -6, ORC_ERR_IO = -7 };

typedef struct orc_db orc_db;
typedef struct orc_result orc_result;

typedef struct orc_pred {
    const char* col;
    int32_t op;
    double num;
    const char* const* strs;
    int32_t nstrs;
} orc_pred;

/* ---- value-level restatements (unit-testable) ---- */
int32_t orc_bytes_to_int(const uint8_t* b);          /* Conversions.scala:17-24 */
void orc_int_to_bytes(int32_t v, uint8_t* out);      /* DataType.scala:40-47    */
int32_t orc_d2i(double d);                           /* Scala Double.toInt  (JLS 5.1.3) */
int8_t orc_d2b(double d);                            /* Scala Double.toByte = (byte)(int)d */
/* DenseCodec*.decode loops (DenseCodec.scala:34-74) incl. the stale-chunk behaviour of a ragged tail:
 * returns the number of values written (ceil(nbytes/width)). */
int64_t orc_dense_decode(const uint8_t* bytes, int64_t nbytes, int width, uint8_t* out_cells);

/* JavaFastPFOR IntegratedIntCompressor (PARITY UNPINNED, see header). */
int64_t orc_iic_compress(const int32_t* in, int32_t n, int32_t* out_words, int64_t cap);
int32_t orc_iic_uncompress(const int32_t* words, int64_t nwords, int32_t* out, int32_t cap);
/* PFORCodecInt.encode (PFORCodec.scala:17-28): big-endian words + 8 trailing zero bytes. */
int64_t orc_pfor_encode_block(const int32_t* in, int32_t n, uint8_t* out, int64_t cap);
/* Inverse of the above (the reference's own decode is broken). */
int32_t orc_pfor_decode_block(const uint8_t* bytes, int64_t nbytes, int32_t* out, int32_t cap);

/* ---- SegmentManager (SegmentManager.scala:20-112) ---- */
int orc_open(const char* data_dir, orc_db** out);
void orc_close(orc_db* db);
int orc_table_nsegments(orc_db* db, const char* table);       /* getTableSegmentCount */
int orc_table_block_size(orc_db* db, const char* table);
int orc_table_ncols(orc_db* db, const char* table);
int orc_segment_file_id(orc_db* db, const char* table, int canonical_idx); /* id in `<col>_<id>.dat` */
int64_t orc_table_nrows(orc_db* db, const char* table, int seg_begin, int seg_end);

/* ---- Engine.execute, Project branch (Engine.scala:158-198) with the INTENDED semantics of
 *      SURVEY.md §3.4 (B1: drain all segments; B2: skip empty batches), canonical order.
 *      Segments [seg_begin, seg_end) of the canonical list; seg_end < 0 = all.
 *      nthreads mirrors --cpu-count (one task per segment, Engine.scala:176-180). ---- */
int orc_query(orc_db* db, const char* table, const orc_pred* preds, int npreds,
              const char* const* proj_cols, int nproj, int64_t limit, int nthreads,
              int seg_begin, int seg_end, orc_result** out);
int64_t orc_result_nrows(const orc_result* r);
int64_t orc_result_nmatched(const orc_result* r);   /* matches before LIMIT (only exact when limit<=0) */
int orc_result_ncols(const orc_result* r);
int orc_result_col_type(const orc_result* r, int c);
int orc_result_col_width(const orc_result* r, int c);
const void* orc_result_col_data(const orc_result* r, int c);
/* Where the real reference would have thrown when run with --cpu-count 1: 0 = it completes
 * (LIMIT reached first), 1 = B1 (None.get at the first end-of-segment marker,
 * ResultQueue.scala:22), 2 = B2 (empty batch, Project.scala:50-57).  rows = rows it prints first. */
int orc_result_ref_throw(const orc_result* r, int64_t* rows_before);
int orc_result_format_row(const orc_result* r, int64_t row, char* buf, size_t buflen); /* Record.scala:13 */
void orc_result_free(orc_result* r);

/* ---- Engine.execute, ProjectAgg branch: count / min / max ... group by (ProjectAggregate.scala:115-226,
 *      ProjectAggregateQueue.scala:9-54).  One row per group in first-appearance canonical order; columns = the group
 *      columns, then per aggregate COUNT -> int64 (column type 3), MIN / MAX -> double (column type 4). ---- */
enum { ORC_AGG_COUNT = 0, ORC_AGG_MIN = 1, ORC_AGG_MAX = 2 };
typedef struct orc_agg {
    const char* col;
    int32_t op;
} orc_agg;
int orc_query_agg(orc_db* db, const char* table, const orc_pred* preds, int npreds, const orc_agg* aggs, int naggs,
                  const char* const* group_cols, int ngroup, int nthreads, int seg_begin, int seg_end, orc_result** out);

/* Selection bitmap of the conjunction alone, canonical row order, bit i of word w = row 32w+i. */
int orc_filter_bitmap(orc_db* db, const char* table, const orc_pred* preds, int npreds,
                      int seg_begin, int seg_end, uint32_t** words, int64_t* nwords, int64_t* nselected);
void orc_free(void* p);

/* ---- Writer side: SegmentWriter (Segment.scala:70-152) + the LoaderCli roll (LoaderCli.scala:142-148) +
 *      TableIO.store (Table.scala:27-35,50-59), so that fixtures and the reference arm's tables are made without any
 *      product code.  Cells are little-endian, `width` bytes each (DataType.scala:40-47, 60, 69). ---- */
int orc_write_table_meta(const char* data_dir, const char* table, const char* meta_json); /* clears the table dir */
int orc_write_column(const char* data_dir, const char* table, const char* col, int codec, int width, const void* cells,
                     int64_t nrows, int block_size, int segment_size);
/* BASELINE.md's synthetic tables (id:DENSE_INT|PFOR_INT, state:DENSE_STRING size=2, age:DENSE_TINYINT). */
void orc_synth_row(int64_t row, int32_t* id, int8_t* age, char state[2]);
int orc_synth_write(const char* data_dir, const char* table, int64_t nrows, int32_t block_size, int32_t segment_size,
                    int32_t id_codec, int32_t seg_begin, int32_t seg_end, int write_table_meta, int nthreads);

const char* orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
