#include "lk_expr.h"

#include <algorithm>
#include <cmath>

namespace lk {

static std::unique_ptr<Clause> handle_filter(const Json& node);

static std::unique_ptr<Clause> to_basic_filter(const Json& node) {
  // ASTUtils.scala:276-288
  auto c = std::make_unique<Clause>();
  c->kind = Clause::Leaf;
  const Json* k = node.get("k");
  LK_CHECK(k && k->text(), LK_ERR_INVALID, "No `k` provided in filter!");
  c->k = k->str;
  const Json* op = node.get("op");
  LK_CHECK(op && op->text(), LK_ERR_INVALID, "No op provided for filter!");
  c->op = op->str;
  if (const Json* v = node.get("v"); v && v->is_arr())
    for (auto& e : v->arr) {
      LK_CHECK(e.text(), LK_ERR_INVALID, "filter values must be strings (JsonNode.textValue)");
      c->v.push_back(e.str);
    }
  LK_CHECK(!c->v.empty() || c->op == "exists", LK_ERR_INVALID, "No value for key = " + c->k + " provided in filter!");
  if (const Json* e = node.get("extracted")) c->extracted = e->as_bool();
  if (const Json* e = node.get("computed")) c->computed = e->as_bool();
  if (const Json* e = node.get("dataType"); e && e->text()) c->data_type = e->str;
  return c;
}

static std::unique_ptr<Clause> to_binary_clause(const Json& node) {
  // ASTUtils.scala:379-404: every non-textual member is a sub-clause; folded left with `op`
  const Json* op = node.get("op");
  LK_CHECK(op != nullptr, LK_ERR_INVALID, "No `op` provided in binary query clause!");
  LK_CHECK(op->text(), LK_ERR_INVALID, "binary clause op must be a string");
  LK_CHECK(op->str == "and" || op->str == "or", LK_ERR_INVALID, "unknown binary op " + op->str);  // BaseExpr.scala:506-508
  std::vector<std::unique_ptr<Clause>> clauses;
  for (auto& kv : node.obj)
    if (!kv.second.is_str()) {
      LK_CHECK(kv.second.is_obj(), LK_ERR_INVALID, "binary clause members must be objects");
      clauses.push_back(handle_filter(kv.second));
    }
  LK_CHECK(clauses.size() >= 2, LK_ERR_INVALID, "Atleast two clauses required in a binary clause!");
  std::unique_ptr<Clause> acc;
  for (auto& c : clauses) {
    if (!acc) { acc = std::move(c); continue; }
    auto b = std::make_unique<Clause>();
    b->kind = op->str == "and" ? Clause::And : Clause::Or;
    b->op = op->str;
    b->a = std::move(acc);
    b->b = std::move(c);
    acc = std::move(b);
  }
  return acc;
}

static std::unique_ptr<Clause> handle_filter(const Json& node) {
  // ASTUtils.scala:406-417
  LK_CHECK(node.is_obj(), LK_ERR_INVALID, "filter clause must be an object");
  if (const Json* n = node.get("not"); n && !n->is_null()) {
    auto c = std::make_unique<Clause>();
    c->kind = Clause::Not;
    c->a = handle_filter(*n);
    return c;
  }
  if (const Json* k = node.get("k"); k && !k->is_null()) return to_basic_filter(node);
  return to_binary_clause(node);
}

BaseExpr parse_base_expr(const Json& node) {
  // ASTUtils.toBaseExpr, ASTUtils.scala:296-377
  LK_CHECK(node.is_obj(), LK_ERR_INVALID, "baseExpr must be an object");
  BaseExpr e;
  if (const Json* j = node.get("id"); j && j->text()) e.id = j->str;
  if (const Json* j = node.get("dataset")) {
    LK_CHECK(j->text(), LK_ERR_INVALID, "dataset must be a string");
    e.dataset = j->str;
  }
  LK_CHECK(e.dataset == "metrics" || e.dataset == "logs" || e.dataset == "traces", LK_ERR_INVALID,
           "Invalid dataset: " + e.dataset);  // BaseExpr.scala:213
  if (const Json* j = node.get("metricType"); j && j->text()) e.metric_type = j->str;
  if (const Json* j = node.get("extract"); j && !j->is_null()) e.has_extract = true;
  if (const Json* j = node.get("compute"); j && !j->is_null()) e.has_compute = true;
  if (const Json* c = node.get("chart"); c && c->is_obj()) {
    e.has_chart = true;
    if (const Json* g = c->get("groupBys"); g && g->is_arr())
      for (auto& x : g->arr) {
        LK_CHECK(x.text(), LK_ERR_INVALID, "groupBys must be strings");
        e.chart.group_bys.push_back(x.str);
      }
    if (const Json* a = c->get("aggregation"); a && a->text()) e.chart.aggregation = a->str;
    if (const Json* r = c->get("rollup"); r && r->text()) { e.chart.has_rollup = true; e.chart.rollup = r->str; }
    if (const Json* t = c->get("type"); t && t->text()) e.chart.type = t->str;
    if (const Json* f = c->get("fieldName"); f && f->text()) { e.chart.has_field_name = true; e.chart.field_name = f->str; }
    if (const Json* f = c->get("fieldType"); f && f->text()) { e.chart.has_field_type = true; e.chart.field_type = f->str; }
  }
  const Json* f = node.get("filter");
  LK_CHECK(f && !f->is_null(), LK_ERR_INVALID, "No filter provided!");
  e.filter = handle_filter(*f);
  return e;
}

PushDownRequest parse_push_down_request(const std::string& json) {
  // PushDownRequest.fromJson, SegmentRequest.scala:45-60
  Json root = parse_json(json);
  LK_CHECK(root.is_obj(), LK_ERR_INVALID, "PushDownRequest must be a JSON object");
  PushDownRequest r;
  const Json* b = root.get("baseExpr");
  LK_CHECK(b != nullptr, LK_ERR_INVALID, "PushDownRequest.baseExpr missing");
  r.expr = parse_base_expr(*b);
  const Json* s = root.get("segmentRequests");
  LK_CHECK(s && s->is_arr(), LK_ERR_INVALID, "PushDownRequest.segmentRequests missing");
  for (auto& sr : s->arr) {
    LK_CHECK(sr.is_obj(), LK_ERR_INVALID, "segmentRequest must be an object");
    SegmentRequest q;
    if (const Json* j = sr.get("dataset"); j && j->text()) q.dataset = j->str;
    if (const Json* j = sr.get("segmentId"); j && j->text()) q.segment_id = j->str;
    const Json* st = sr.get("stepInMillis");
    const Json* a = sr.get("startTs");
    const Json* z = sr.get("endTs");
    LK_CHECK(st && a && z, LK_ERR_INVALID, "segmentRequest needs stepInMillis, startTs, endTs");
    q.step_ms = st->as_i64();
    q.start_ts = a->as_i64();
    q.end_ts = z->as_i64();
    r.segments.push_back(q);
  }
  if (const Json* j = root.get("reverseSort")) r.reverse_sort = j->as_bool();
  if (const Json* j = root.get("isTagQuery")) r.is_tag_query = j->as_bool();
  if (const Json* j = root.get("tagDataType"); j && !j->is_null()) {
    LK_CHECK(j->is_obj(), LK_ERR_INVALID, "tagDataType must be an object");
    r.has_tag_data_type = true;
    if (const Json* n = j->get("tagName"); n && n->text()) r.tag_name = n->str;
    if (const Json* t = j->get("dataType"); t && t->text()) r.tag_type = t->str;
    LK_CHECK(!r.tag_name.empty(), LK_ERR_INVALID, "tagDataType needs a tagName");
  }
  return r;
}

LeafPredicate compile_leaf(const Clause& c) {
  LeafPredicate p;
  const std::string& op = c.op;
  auto need_number = [&]() {
    // normalizedValue(), BaseExpr.scala:454-459.  duration/datasize go through QuantityParser (outside the hot path).
    LK_CHECK(c.data_type == "number", LK_ERR_UNSUPPORTED,
             "comparison operators are supported for dataType=number only (got " + c.data_type + ")");
    LK_CHECK(c.v.size() == 1, LK_ERR_INVALID, "filter value is a list of values for dataType: " + c.data_type);
    char* end = nullptr;
    p.number = strtod(c.v[0].c_str(), &end);
    LK_CHECK(end && *end == 0 && !c.v[0].empty(), LK_ERR_INVALID, "NumberFormatException: " + c.v[0]);
  };
  if (c.data_type == "number" || c.data_type == "duration" || c.data_type == "datasize")
    LK_CHECK(c.v.size() == 1, LK_ERR_INVALID, "filter value is a list of values for dataType: " + c.data_type);  // :450-452
  if (op == "has" || op == "exists") p.op = LeafPredicate::Exists;
  else if (op == "eq") { p.op = LeafPredicate::Eq; p.values = {c.v.at(0)}; }
  else if (op == "!=") { p.op = LeafPredicate::Ne; p.values = {c.v.at(0)}; }
  else if (op == "in") { p.op = LeafPredicate::In; p.values = c.v; }
  else if (op == "not_in") { p.op = LeafPredicate::NotIn; p.values = c.v; }
  else if (op == "regex") { p.op = LeafPredicate::RegexMatch; p.re = std::make_unique<Regex>(c.v.at(0), true); }
  else if (op == "contains") { p.op = LeafPredicate::RegexMatch; p.re = std::make_unique<Regex>(".*" + c.v.at(0) + ".*", true); }
  else if (op == "gt") { p.op = LeafPredicate::Gt; need_number(); }
  else if (op == "ge") { p.op = LeafPredicate::Ge; need_number(); }
  else if (op == "lt") { p.op = LeafPredicate::Lt; need_number(); }
  else if (op == "le") { p.op = LeafPredicate::Le; need_number(); }
  else fail(LK_ERR_INVALID, "Invalid operator " + op);  // BaseExpr.scala:503
  return p;
}

Truth LeafPredicate::eval_string(const std::string* s) const {
  if (op == Exists) return s ? T : F;
  if (!s) return N;
  switch (op) {
    case Eq: return *s == values[0] ? T : F;
    case Ne: return *s != values[0] ? T : F;
    case In: return std::find(values.begin(), values.end(), *s) != values.end() ? T : F;
    case NotIn: return std::find(values.begin(), values.end(), *s) == values.end() ? T : F;
    case RegexMatch: return re->search(*s) ? T : F;
    default: fail(LK_ERR_UNSUPPORTED, "numeric comparison on a string column");
  }
}

static int cmp_total(double a, double b) {  // DuckDB: NaN is the greatest value, NaN == NaN
  bool an = std::isnan(a), bn = std::isnan(b);
  if (an || bn) return (int)an - (int)bn;
  return (a > b) - (a < b);
}

Truth LeafPredicate::eval_number(bool is_null, double x) const {
  if (op == Exists) return is_null ? F : T;
  if (is_null) return N;
  int r = cmp_total(x, number);
  switch (op) {
    case Gt: return r > 0 ? T : F;
    case Ge: return r >= 0 ? T : F;
    case Lt: return r < 0 ? T : F;
    case Le: return r <= 0 ? T : F;
    default: fail(LK_ERR_UNSUPPORTED, "string operator on a numeric column");
  }
}

void collect_leaves(const Clause& c, std::vector<const Clause*>& leaves) {
  if (c.kind == Clause::Leaf) { leaves.push_back(&c); return; }
  collect_leaves(*c.a, leaves);
  if (c.b) collect_leaves(*c.b, leaves);
}

static void filter_field_set(const Clause& c, std::vector<std::string>& out) {
  // BaseExpr.scala:652-663: Filter and BinaryClause only; NotClause contributes nothing (`case _ =>`)
  if (c.kind == Clause::Leaf) out.push_back(c.k);
  else if (c.kind == Clause::And || c.kind == Clause::Or) { filter_field_set(*c.a, out); filter_field_set(*c.b, out); }
}

void field_set(const BaseExpr& e, std::vector<std::string>& out) {
  filter_field_set(*e.filter, out);
  if (e.has_chart)
    for (auto& g : e.chart.group_bys) out.push_back(g);
  std::sort(out.begin(), out.end());
  out.erase(std::unique(out.begin(), out.end()), out.end());
}

void all_filter_columns(const Clause& c, std::vector<std::string>& out) {
  std::vector<const Clause*> leaves;
  collect_leaves(c, leaves);
  for (auto* l : leaves) out.push_back(l->k);
  std::sort(out.begin(), out.end());
  out.erase(std::unique(out.begin(), out.end()), out.end());
}

}  // namespace lk
