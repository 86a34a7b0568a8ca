/*
 * JNA binding of liblakeside_b200.so -- the B200 drop-in for what Commons.toGlobResultSet obtains from DuckDB
 * (core/src/main/scala/com/cardinal/utils/Commons.scala:200-254).  Same pattern as the reference's own native binding,
 * NLPUtils.RegexpInterface (core/src/main/scala/com/cardinal/utils/ast/queries/NLPUtils.scala:30-52): a trait extending
 * com.sun.jna.Library, loaded once from /app/libs.  Every parameter is a C string, a pointer or an int / long / double;
 * no struct crosses the boundary (include/lakeside_b200.h).
 *
 * NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK / Scala / Gradle there): the Python ctypes binding
 * (lakeside_b200/_lib.py) drives the same entry points with the same calling convention and is what the tests exercise.
 * Build: add this directory to core's source sets; JNA 5.12.1 is already a dependency (core/build.gradle:68).
 */
package com.cardinal.gpu

import com.sun.jna.{Library, Native, Pointer}
import com.sun.jna.ptr.{IntByReference, LongByReference, PointerByReference}

trait LakesideB200 extends Library {
  // ---- library ----
  def lk_init(optionsJson: String): Int
  def lk_shutdown(): Unit
  def lk_last_error(): String
  def lk_version(): String
  def lk_device_count(): Int

  // ---- HBM-resident segment cache (device-side continuation of the worker's segment file cache, WorkerApi.scala:53-64) ----
  def lk_cache_stats(stats: Array[Long]): Int   // [capacity, resident bytes, segments, column hits, column misses, evicted segments]
  def lk_cache_configure(capacityBytes: Long): Int
  def lk_cache_clear(): Unit

  // ---- one-shot glob evaluation: Commons.toGlobResultSet ----
  def lk_eval(pushDownRequestJson: String, parquetPaths: Array[String], nPaths: Int, out: PointerByReference): Int

  // ---- staged evaluation: resident segments, fused multi-aggregate pass (AVG = SUM + COUNT in one scan) ----
  def lk_query_create(pushDownRequestJson: String, optionsJson: String, out: PointerByReference): Int
  def lk_query_add_segment_file(q: Pointer, path: String): Int
  def lk_query_prepare(q: Pointer): Int
  def lk_query_execute(q: Pointer): Int
  def lk_query_finalize(q: Pointer, out: PointerByReference): Int
  def lk_query_eval(q: Pointer, aggregation: String, chartType: String, metricType: String, out: Array[Double], cap: Long): Long
  def lk_query_destroy(q: Pointer): Unit
  // Formula.eval (Formula.scala:32-69) over the reduced rows of two finalized queries; a constant side passes null
  def lk_formula_eval(e1: Pointer, e2: Pointer, specJson: String, cap: Long, outTs: Array[Long], outValue: Array[Double], outSide: Array[Int],
                      outRow: Array[Long], nOut: LongByReference): Int

  // ---- result: what Commons.toDataPoint reads from the JDBC ResultSet (Commons.scala:399-462) ----
  def lk_result_num_rows(r: Pointer): Long
  def lk_result_num_values(r: Pointer): Int
  def lk_result_num_cols(r: Pointer): Int
  def lk_result_col_name(r: Pointer, col: Int): String
  def lk_result_ts(r: Pointer): Pointer                 // int64[num_rows]
  def lk_result_value(r: Pointer, a: Int): Pointer       // double[num_rows]
  def lk_result_tag_codes(r: Pointer, t: Int): Pointer   // int32[num_rows], -1 = SQL NULL
  def lk_result_tag_dict(r: Pointer, t: Int, n: IntByReference, strings: PointerByReference): Int
  def lk_result_get_long(r: Pointer, row: Long, col: Int): Long      // 1-based column, like JDBC
  def lk_result_get_double(r: Pointer, row: Long, col: Int): Double  // SQL NULL -> 0.0, like ResultSet.getDouble
  def lk_result_get_string(r: Pointer, row: Long, col: Int): String  // SQL NULL -> null
  def lk_result_to_sse(r: Pointer, row0: Long, row1: Long, sketchKeys: Array[String], nKeys: Int, fallbackTags: Array[String],
                       nFallback: Int, buf: Array[Byte], cap: Long): Long
  def lk_result_free(r: Pointer): Unit

  // ---- K-way merge of per-segment streams (QueryEngineV2.mergeSortedSource, QueryEngineV2.scala:76-97) ----
  def lk_merge_streams(k: Int, ts: Array[Pointer], gid: Array[Pointer], value: Array[Pointer], lens: Array[Long], reverse: Int,
                       outTs: Pointer, outGid: Pointer, outValue: Pointer, outSrc: Pointer): Int
}

object LakesideB200 {
  val LK_OK = 0
  val LK_ERR_INVALID = 1
  val LK_ERR_UNSUPPORTED = 2 // shape outside the GPU path: the adapter falls back to DuckDB
  val LK_ERR_IO = 3
  val LK_ERR_CUDA = 4
  val LK_ERR_QUERY = 5       // the reference's SQL would not bind: stream nothing (Commons.scala:249-253)
  val LK_ERR_NOMEM = 6

  /** Set LAKESIDE_GPU=1 on the worker to route aggregate push-downs through the library. */
  lazy val enabled: Boolean = sys.env.get("LAKESIDE_GPU").contains("1")

  lazy val lib: LakesideB200 = {
    val path = sys.env.getOrElse("LAKESIDE_B200_LIB", "/app/libs/liblakeside_b200.so") // next to lib-trigram.so (query-worker/Dockerfile:27-28)
    val l = Native.load(path, classOf[LakesideB200])
    if (l.lk_init(sys.env.getOrElse("LAKESIDE_B200_OPTIONS", null)) != LK_OK)
      throw new RuntimeException(s"lakeside_b200: ${l.lk_last_error()}") // as NLPUtils does on a load failure (NLPUtils.scala:48-52)
    l
  }
}
