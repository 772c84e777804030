// Reference-side binding for libimm3gpu.so (include/imm3.h); see INTEGRATION.md.  Not compiled in this repository's build image (no JVM).
// engine/src/main/scala/immutabledb/engine/gpu/GpuEngine.scala — same signature as Engine.execute (Engine.scala:158)
class GpuEngine(dataDir: String, device: Int = 0) extends AutoCloseable {
  private val arena = Arena.ofShared()
  private val db: MemorySegment = {                       // replaces `new SegmentManager(dataDir)` (SqlCli.scala:65)
    val out  = arena.allocate(ADDRESS)
    val opts = arena.allocate(Imm3.OPTS); opts.set(JAVA_INT, 0, device); opts.set(JAVA_INT, 8, 1)
    check(Imm3.open.invoke(arena.allocateFrom(dataDir), opts, out).asInstanceOf[Int])
    out.get(ADDRESS, 0)
  }

  /** Flatten the And/Or tree left to right exactly like PipelineThread.runOps (Engine.scala:237-245). */
  private def leaves(s: Select): List[Select] = s match {
    case And(a, b) => leaves(a) ++ leaves(b)
    case Or(a, b)  => leaves(a) ++ leaves(b)               // the reference drops the tag: `or` behaves as `and`
    case NoSelect  => Nil
    case leaf      => List(leaf)
  }

  /** The conjunction as an array of imm3_pred (include/imm3.h). */
  private def packPreds(a: Arena, ls: List[Select]): MemorySegment = {
      val preds = a.allocate(Imm3.PRED, math.max(ls.size, 1))
      ls.zipWithIndex.foreach { case (Select(col, cond), i) =>
        val p = preds.asSlice(i * Imm3.PRED.byteSize, Imm3.PRED.byteSize)
        p.set(ADDRESS, 0, a.allocateFrom(col))
        cond match {
          case GT(v)    => p.set(JAVA_INT, 8, 1); p.set(JAVA_DOUBLE, 16, v)
          case LT(v)    => p.set(JAVA_INT, 8, 2); p.set(JAVA_DOUBLE, 16, v)
          case EQ(v)    => p.set(JAVA_INT, 8, 3); p.set(JAVA_DOUBLE, 16, v)
          case Match(vs) =>
            val arr = a.allocate(ADDRESS, vs.size)
            vs.zipWithIndex.foreach { case (s, k) => arr.setAtIndex(ADDRESS, k, a.allocateFrom(s)) }
            p.set(JAVA_INT, 8, 4); p.set(ADDRESS, 24, arr); p.set(JAVA_INT, 32, vs.size)
          case _        => p.set(JAVA_INT, 8, 6)           // NotMatch / NoOp: the library answers IMM3_ERR_UNSUPPORTED (Select.scala:22)
        }
      }
      preds
  }

  def execute(q: Query): Either[Throwable, Iterator[Row]] = q match {
    case Query(table, select, Project(cols, limit)) =>
      val a     = Arena.ofConfined()
      val ls    = leaves(select)
      val preds = packPreds(a, ls)
      val proj = a.allocate(ADDRESS, math.max(cols.size, 1))
      cols.zipWithIndex.foreach { case (c, k) => proj.setAtIndex(ADDRESS, k, a.allocateFrom(c)) }
      val out = a.allocate(ADDRESS)
      val rc  = Imm3.query.invoke(db, a.allocateFrom(table), preds, ls.size, proj, cols.size, limit.toLong, out).asInstanceOf[Int]
      if (rc != 0) { a.close(); Left(new Exception(err())) }          // Left(Throwable), never a hang (Engine.scala:182-188 hangs)
      else Right(rows(out.get(ADDRESS, 0), cols.size, a))
    case Query(table, select, ProjectAgg(aggs, groupBy)) =>                 // Engine.scala:199-229 + resolveProjectOp (130-156)
      val a     = Arena.ofConfined()
      val ls    = leaves(select)
      val preds = packPreds(a, ls)
      val as    = a.allocate(Imm3.AGG, math.max(aggs.size, 1))
      aggs.zipWithIndex.foreach { case (agg, i) =>
        val (col, op) = agg match {
          case Count(c, _) => (c, 0)
          case Min(c, _)   => (c, 1)
          case Max(c, _)   => (c, 2)
          case other       => (other.col, 4)              // Avg: the reference throws "Unknown Aggregate type" (Engine.scala:153); so does the library
        }
        val e = as.asSlice(i * Imm3.AGG.byteSize, Imm3.AGG.byteSize)
        e.set(ADDRESS, 0, a.allocateFrom(col)); e.set(JAVA_INT, 8, op)
      }
      val gb = a.allocate(ADDRESS, math.max(groupBy.size, 1))
      groupBy.zipWithIndex.foreach { case (c, k) => gb.setAtIndex(ADDRESS, k, a.allocateFrom(c)) }
      val out = a.allocate(ADDRESS)
      val rc  = Imm3.queryAgg.invoke(db, a.allocateFrom(table), preds, ls.size, as, aggs.size, gb, groupBy.size, out).asInstanceOf[Int]
      if (rc != 0) { a.close(); Left(new Exception(err())) }
      else Right(aggRows(out.get(ADDRESS, 0), groupBy.size, aggs, a))
  }

  /** One Row per group, the aggregators' repr only and in select-list order (ProjectAggregateQueue.scala:48-50): count -> Long,
    * min / max of INT / TINYINT -> Double (MinDoubleAggr / MaxDoubleAggr, ProjectAggregate.scala:36-58).  The result's first
    * `ngroup` columns are the group cells; the library returns the groups in order of first appearance. */
  private def aggRows(r: MemorySegment, ngroup: Int, aggs: List[Aggregate], a: Arena): Iterator[Row] = {
    val n    = Imm3.nrows.invoke(r).asInstanceOf[Long]
    val data = aggs.indices.map(k => Imm3.colData.invoke(r, ngroup + k).asInstanceOf[MemorySegment].reinterpret(n * 8))
    new Iterator[Row] {
      private var i = 0L
      def hasNext: Boolean = { val more = i < n; if (!more && i == n) { Imm3.resultFree.invoke(r); a.close(); i += 1 }; more }
      def next(): Row = {
        val cells = aggs.zipWithIndex.map {
          case (_: Count, k) => data(k).get(JAVA_LONG_UNALIGNED, i * 8)           // IMM3_COL_COUNT: int64
          case (_, k)        => data(k).get(JAVA_DOUBLE_UNALIGNED, i * 8)         // IMM3_COL_DOUBLE
        }
        i += 1; Row.fromSeq(cells)
      }
    }
  }

  /** Column-major host buffers -> lazy Iterator[Row] (Record.scala:7-14); Int / Byte / String cells as in the reference. */
  private def rows(r: MemorySegment, ncols: Int, a: Arena): Iterator[Row] = {
    val n     = Imm3.nrows.invoke(r).asInstanceOf[Long]
    val types = (0 until ncols).map(c => Imm3.colType.invoke(r, c).asInstanceOf[Int])
    val width = (0 until ncols).map(c => Imm3.colWidth.invoke(r, c).asInstanceOf[Int])
    val data  = (0 until ncols).map(c => Imm3.colData.invoke(r, c).asInstanceOf[MemorySegment].reinterpret(n * width(c)))
    new Iterator[Row] {
      private var i = 0L
      def hasNext: Boolean = { val more = i < n; if (!more && i == n) { Imm3.resultFree.invoke(r); a.close(); i += 1 }; more }
      def next(): Row = {
        val cells = (0 until ncols).map { c => types(c) match {
          case 0 => data(c).get(JAVA_INT_UNALIGNED, i * 4)                         // IMM3_COL_INT: little-endian int32
          case 1 => data(c).get(JAVA_BYTE, i)                                      // IMM3_COL_TINYINT
          case _ => new String(data(c).asSlice(i * width(c), width(c)).toArray(JAVA_BYTE)) // STRING(k), DataType.scala:70
        }}
        i += 1; Row.fromSeq(cells)
      }
    }
  }
  private def err(): String = Imm3.lastError.invoke().asInstanceOf[MemorySegment].reinterpret(512).getString(0)
  private def check(rc: Int): Unit = if (rc != 0) throw new Exception(err())
  def close(): Unit = { Imm3.close.invoke(db); arena.close() }
}
