/*
 * The switch a maintainer adds at the top of Commons.toGlobResultSet (core/src/main/scala/com/cardinal/utils/Commons.scala:200):
 *
 *   GpuGlob.tryEvaluate(queryId, pushDownRequest, localParquet, sr => toParquetFilePath(localParquet, sr)) match {
 *     case GpuGlob.Rows(statement, resultSet) => return (statement, resultSet, null)   // connection = null is guarded at :325
 *     case GpuGlob.Nothing                    => return (null, null, null)             // log-and-stream-nothing, as :249-253
 *     case GpuGlob.UseDuckDb                  => // fall through to the JDBC path below
 *   }
 *
 * Everything downstream (resultSetToSource, toDataPoint, PushDownAggregatorStage, dataPointResponseToSSE, the HTTP route) is
 * untouched.  NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK); see LakesideB200.scala.
 */
package com.cardinal.gpu

import com.cardinal.model.{PushDownRequest, SegmentRequest}
import com.sun.jna.ptr.PointerByReference
import org.slf4j.LoggerFactory

import java.sql.{ResultSet, Statement}

object GpuGlob {
  sealed trait Outcome
  final case class Rows(statement: Statement, resultSet: ResultSet) extends Outcome
  case object Nothing extends Outcome
  case object UseDuckDb extends Outcome

  private val logger = LoggerFactory.getLogger(getClass)

  def tryEvaluate(queryId: String, request: PushDownRequest, localParquet: Boolean, pathOf: SegmentRequest => String): Outcome = {
    // the GPU path covers the aggregate push-down over local, sealed segments; the library itself answers LK_ERR_UNSUPPORTED for
    // percentile / ces rollups, extract / compute sub-queries, compressed pages ... (SURVEY §8b "Error convention")
    // aggregate push-downs and tag queries with a tagDataType (two JDBC columns: tag, count); exemplar queries and `SELECT *`
    // tag queries select whole rows and stay with DuckDB
    val tagCount = request.isTagQuery && request.tagDataType.isDefined
    if (!LakesideB200.enabled || !localParquet || (request.isTagQuery && !tagCount) || (!tagCount && request.baseExpr.chartOpts.isEmpty)) return UseDuckDb
    val lib = LakesideB200.lib
    val paths = request.segmentRequests.map(pathOf).toArray
    val out = new PointerByReference()
    val start = System.currentTimeMillis()
    lib.lk_eval(PushDownRequest.toJson(request), paths, paths.length, out) match {
      case LakesideB200.LK_OK =>
        logger.info(s"[$queryId][gpu-glob/${paths.length}] toResultSet took ${System.currentTimeMillis() - start}ms")
        Rows(LkResultSet.noopStatement, LkResultSet(out.getValue))
      case LakesideB200.LK_ERR_UNSUPPORTED =>
        logger.debug(s"[$queryId] lakeside_b200 declines: ${lib.lk_last_error()}")
        UseDuckDb
      case rc =>
        logger.error(s"[$queryId] lakeside_b200 error $rc: ${lib.lk_last_error()}")
        Nothing
    }
  }
}
