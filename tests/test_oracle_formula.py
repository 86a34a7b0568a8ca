"""The oracle's restatement of Formula.eval / ConstantExpr (Formula.scala:32-69, ASTUtils.scala:42-89) against the one test the
reference holds for it -- FormulaListenerTest.testAddZero (core/src/test/.../FormulaListenerTest.scala:78-103): the formula
``((a + 0) / b) * 100`` over a SketchGroup that carries a single map-sketch input for ``b`` yields exactly ONE result -- and
against hand-computed cases of every rule the restatement encodes."""
import math

import lakeside_oracle as lo


def _b_map(ts=1, value=1.0, tags=None):
    # BaseExpr b of the reference's test: chart {aggregation: sum}, no group-bys -> key "default" (BaseExpr.scala:665-695)
    return {"default": (ts, value, tags if tags is not None else {"name": "foo"})}


def test_reference_test_add_zero_has_one_result():
    ts = 1
    a = {}  # BaseExpr a has no sketch input in the group (and no chart options): BaseExpr.eval returns an empty map
    zero = lo.constant_expr_eval(0.0, [], ts, [])        # no group-bys anywhere: {"default": (ts, 0.0, {})}  (ASTUtils.scala:51-55)
    assert zero == {"default": (ts, 0.0, {})}
    a_plus_0 = lo.formula_eval("add", a, zero)           # add fills the missing side with 0 and the other side's tags
    assert a_plus_0 == {"default": (ts, 0.0, {})}
    ratio = lo.formula_eval("div", a_plus_0, _b_map())   # b = 1.0 != 0: 0.0 / 1.0
    assert ratio == {"default": (ts, 0.0, {})}
    result = lo.formula_eval("mul", ratio, lo.constant_expr_eval(100.0, [], ts, []))
    assert len(result) == 1                              # FormulaListenerTest.scala:102: assert(result.size == 1)
    assert result["default"][1] == 0.0


def test_missing_side_only_add_fills():
    e1 = {"x": (5, 2.0, {"g": "x"})}
    e2 = {"y": (5, 3.0, {"g": "y"})}
    got = lo.formula_eval("add", e1, e2)
    assert got == {"x": (5, 2.0, {"g": "x"}), "y": (5, 3.0, {"g": "y"})}  # the filled side takes the other's timestamp and tags
    for op in ("sub", "mul", "div"):
        assert lo.formula_eval(op, e1, e2) == {}


def test_zero_divisor_is_missing_data_and_nan_passes():
    e1 = {"k": (1, 4.0, {"t": "1"})}
    assert lo.formula_eval("div", e1, {"k": (1, 0.0, {})}) == {}
    assert lo.formula_eval("div", e1, {"k": (1, -0.0, {})}) == {}
    got = lo.formula_eval("div", e1, {"k": (1, float("nan"), {})})  # NaN != 0 is true on the JVM: the division happens
    assert math.isnan(got["k"][1]) and got["k"][2] == {"t": "1"}
    assert lo.formula_eval("div", e1, {"k": (1, 8.0, {})})["k"][1] == 0.5


def test_constant_takes_the_keys_of_the_group_inputs():
    ins = [lo.SketchInput(7, {"a": "1", "b": "x"}, {}), lo.SketchInput(7, {"a": "2"}, {}), lo.SketchInput(7, {"a": "1", "b": "x", "n": "later"}, {})]
    got = lo.constant_expr_eval(3.0, ["b", "a"], 7, ins)  # key = sorted group-by names, missing -> "" (ASTUtils.scala:87-89)
    assert set(got) == {"1:x", "2:"}
    assert got["1:x"] == (7, 3.0, {"a": "1", "b": "x", "n": "later"})  # a later input of the same key replaces an earlier one
    assert got["2:"] == (7, 3.0, {"a": "2"})
