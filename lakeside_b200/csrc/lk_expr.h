// DataExpr (the reference's BaseExpr) and PushDownRequest: types + JSON decode + leaf predicates.
// Mirrors core/src/main/scala/com/cardinal/utils/ast/ASTUtils.scala:124-137, 222-229, 276-417 and
// core/src/main/scala/com/cardinal/model/SegmentRequest.scala:45-98.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "lk_json.h"
#include "lk_regex.h"

namespace lk {

enum Truth : uint8_t { F = 0, T = 1, N = 2 };  // Kleene truth values (SQL three-valued logic)

struct Clause {
  enum Kind { Leaf, And, Or, Not } kind = Leaf;
  // Leaf
  std::string k;
  std::vector<std::string> v;
  std::string op;
  bool extracted = false, computed = false;
  std::string data_type = "string";
  // And / Or / Not
  std::unique_ptr<Clause> a, b;
};

struct ChartOptions {
  std::string aggregation = "sum";
  std::vector<std::string> group_bys;
  std::string type = "count";
  bool has_rollup = false;
  std::string rollup;
  bool has_field_name = false, has_field_type = false;
  std::string field_name, field_type;
};

struct BaseExpr {
  std::string id = "_";
  std::string dataset = "metrics";
  std::unique_ptr<Clause> filter;
  bool has_extract = false, has_compute = false, has_chart = false;
  ChartOptions chart;
  std::string metric_type = "gauge";
};

struct SegmentRequest {
  std::string dataset, segment_id;
  int64_t step_ms = 10000, start_ts = 0, end_ts = 0;
};

struct PushDownRequest {
  BaseExpr expr;
  std::vector<SegmentRequest> segments;
  bool reverse_sort = false, is_tag_query = false, has_tag_data_type = false;
  std::string tag_name, tag_type;  // tagDataType {tagName, dataType} (model/query/common/TagDataType.scala:19)
};

PushDownRequest parse_push_down_request(const std::string& json);
BaseExpr parse_base_expr(const Json& node);

// One compiled filter leaf (BaseExpr.filterSqlAndAccumulateFields, BaseExpr.scala:433-513).
struct LeafPredicate {
  enum Op { Exists, Eq, Ne, In, NotIn, RegexMatch, Gt, Ge, Lt, Le } op;
  std::vector<std::string> values;
  std::unique_ptr<Regex> re;
  double number = 0;  // gt/ge/lt/le constant (dataType "number")
  bool is_numeric_op() const { return op >= Gt; }
  // VARCHAR cell; s == nullptr is SQL NULL
  Truth eval_string(const std::string* s) const;
  // numeric cell with DuckDB's total order (NaN greatest)
  Truth eval_number(bool is_null, double x) const;
};
LeafPredicate compile_leaf(const Clause& leaf);

// Collects leaves in evaluation order; fills `cols` with the distinct referenced column names.
void collect_leaves(const Clause& c, std::vector<const Clause*>& leaves);
// fieldSet(): filter fields NOT under a NotClause + groupBys (BaseExpr.scala:648-663)
void field_set(const BaseExpr& e, std::vector<std::string>& out);
void all_filter_columns(const Clause& c, std::vector<std::string>& out);

}  // namespace lk
