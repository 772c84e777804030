// Reference-side binding for libimm3gpu.so (include/imm3.h); see INTEGRATION.md.  Not compiled in this repository's build image (no JVM).
class Imm3Jni { @native def open(dataDir: String, device: Int, rank: Int, world: Int): Long
                @native def querySql(db: Long, sql: String): Long            // imm3_query_sql: the CLI's own grammar (SQLParser.scala:8-129)
                @native def nrows(res: Long): Long;  @native def ncols(res: Long): Int
                @native def column(res: Long, col: Int): java.nio.ByteBuffer // NewDirectByteBuffer(imm3_result_col_data, nrows*width)
                @native def free(res: Long): Unit;   @native def close(db: Long): Unit }
