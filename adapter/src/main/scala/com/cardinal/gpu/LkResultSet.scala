/*
 * java.sql.ResultSet / Statement views over an lk_result, for Commons.resultSetToSource and Commons.toDataPoint
 * (core/src/main/scala/com/cardinal/utils/Commons.scala:280-341, 399-462).  Those consumers call only
 *   resultSet.next()                                  :297
 *   resultSet.getMetaData.getColumnCount / getColumnName(i)   :403-410, 431
 *   resultSet.getLong(1), getDouble(2), getString(i)  :425-432
 *   resultSet.close(), statement.close(), connection.close()  :319-326   (connection may be null: guarded there)
 * so the views are dynamic proxies that answer exactly these and raise SQLFeatureNotSupportedException for the other
 * ~190 methods of the interfaces.  NOT COMPILED IN THIS REPOSITORY'S IMAGE (no JDK); see LakesideB200.scala.
 */
package com.cardinal.gpu

import com.sun.jna.Pointer

import java.lang.reflect.{InvocationHandler, Method, Proxy}
import java.sql.{ResultSet, ResultSetMetaData, SQLFeatureNotSupportedException, Statement}

object LkResultSet {
  private def proxy[T](cls: Class[T])(f: PartialFunction[(String, Array[AnyRef]), Any]): T =
    Proxy
      .newProxyInstance(cls.getClassLoader, Array[Class[_]](cls), new InvocationHandler {
        override def invoke(p: Any, m: Method, args: Array[AnyRef]): AnyRef = {
          val a = if (args == null) Array.empty[AnyRef] else args
          f.applyOrElse((m.getName, a), (_: (String, Array[AnyRef])) => throw new SQLFeatureNotSupportedException(m.getName)).asInstanceOf[AnyRef]
        }
      })
      .asInstanceOf[T]

  /** Rows of `r` in timestamp order; owns `r` (freed by close()). */
  def apply(r: Pointer): ResultSet = {
    val lib = LakesideB200.lib
    val rows = lib.lk_result_num_rows(r)
    val cols = lib.lk_result_num_cols(r)
    val names = (0 until cols).map(i => lib.lk_result_col_name(r, i)).toArray
    var row = -1L
    var closed = false
    var lastNull = false
    val meta = proxy(classOf[ResultSetMetaData]) {
      case ("getColumnCount", _)   => Int.box(cols)
      case ("getColumnName", a)    => names(a(0).asInstanceOf[Integer] - 1)
      case ("getColumnLabel", a)   => names(a(0).asInstanceOf[Integer] - 1)
    }
    proxy(classOf[ResultSet]) {
      case ("next", _)        => row += 1; Boolean.box(row < rows)
      case ("getMetaData", _) => meta
      case ("getLong", a) if a(0).isInstanceOf[Integer]   => Long.box(lib.lk_result_get_long(r, row, a(0).asInstanceOf[Integer]))
      case ("getDouble", a) if a(0).isInstanceOf[Integer] => Double.box(lib.lk_result_get_double(r, row, a(0).asInstanceOf[Integer])) // SQL NULL -> 0.0
      case ("getString", a) if a(0).isInstanceOf[Integer] =>
        val s = lib.lk_result_get_string(r, row, a(0).asInstanceOf[Integer]) // SQL NULL -> null (toDataPoint drops null / "" / "null" tags, :433)
        lastNull = s == null
        s
      case ("wasNull", _)  => Boolean.box(lastNull)
      case ("isClosed", _) => Boolean.box(closed)
      case ("close", _)    => if (!closed) { closed = true; lib.lk_result_free(r) }; null
    }
  }

  /** resultSetToSource closes the statement it was handed (Commons.scala:321): nothing to release here. */
  val noopStatement: Statement = proxy(classOf[Statement]) {
    case ("close", _)    => null
    case ("isClosed", _) => Boolean.box(true)
  }
}
