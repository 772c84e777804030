// Reference-side binding for libimm3gpu.so (include/imm3.h); see INTEGRATION.md.  Not compiled in this repository's build image (no JVM).
// engine/src/main/scala/immutabledb/engine/gpu/Imm3.scala
package immutabledb.engine.gpu
import java.lang.foreign._
import java.lang.foreign.ValueLayout._
import java.lang.invoke.MethodHandle

object Imm3 {
  private val linker = Linker.nativeLinker()
  private val lib    = SymbolLookup.libraryLookup("libimm3gpu.so", Arena.global())
  private def h(name: String, fd: FunctionDescriptor): MethodHandle = linker.downcallHandle(lib.find(name).get, fd)

  val open       = h("imm3_open",        FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS))
  val close      = h("imm3_close",       FunctionDescriptor.of(JAVA_INT, ADDRESS))
  val query      = h("imm3_query",       FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_INT, JAVA_LONG, ADDRESS))
  // imm3_query_agg(db, table, preds, npreds, aggs, naggs, group_cols, ngroup, &out): ProjectAggOp + ProjectAggregateQueueOp
  val queryAgg   = h("imm3_query_agg",   FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS))
  val nrows      = h("imm3_result_nrows",     FunctionDescriptor.of(JAVA_LONG, ADDRESS))
  val ncols      = h("imm3_result_ncols",     FunctionDescriptor.of(JAVA_INT, ADDRESS))
  val colType    = h("imm3_result_col_type",  FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT))
  val colWidth   = h("imm3_result_col_width", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT))
  val colData    = h("imm3_result_col_data",  FunctionDescriptor.of(ADDRESS, ADDRESS, JAVA_INT))
  val resultFree = h("imm3_result_free",      FunctionDescriptor.of(JAVA_INT, ADDRESS))
  val lastError  = h("imm3_last_error",       FunctionDescriptor.of(ADDRESS))

  // struct imm3_pred { const char* col; int32 op; double num; const char* const* strs; int32 nstrs; }  (imm3.h)
  val PRED: MemoryLayout = MemoryLayout.structLayout(
    ADDRESS.withName("col"), JAVA_INT.withName("op"), MemoryLayout.paddingLayout(4),
    JAVA_DOUBLE.withName("num"), ADDRESS.withName("strs"), JAVA_INT.withName("nstrs"), MemoryLayout.paddingLayout(4))
  // struct imm3_agg { const char* col; int32 op; }  op: 0 count, 1 min, 2 max (imm3_agg_op)
  val AGG: MemoryLayout = MemoryLayout.structLayout(ADDRESS.withName("col"), JAVA_INT.withName("op"), MemoryLayout.paddingLayout(4))
  // struct imm3_open_opts { int32 device, rank, world; uint32 flags; }
  val OPTS: MemoryLayout = MemoryLayout.structLayout(JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT)
}
