"""Boundary B2: the post-optimisation callback must take exactly the plan shapes the library
covers and leave everything else to Polars.  Polars itself is not installable here, so the IR is mocked with
classes named like the reference's node/expression views (crates/polars-python/src/lazyframe/visitor/nodes.rs,
visitor/expr_nodes.rs); only the matcher runs — no kernel is launched."""
import types

import pytest

from polars_b200 import engine


def make(_kind, **fields):
    obj = type(_kind, (), {})()
    for k, v in fields.items():
        setattr(obj, k, v)
    return obj


class FakeTraverser:
    """node ids -> IR nodes, expression ids -> expression nodes; records set_udf."""

    def __init__(self, nodes, exprs, root):
        self.nodes, self.exprs, self.cur, self.udf = nodes, exprs, root, None

    def view_current_node(self):
        return self.nodes[self.cur]

    def set_node(self, n):
        self.cur = n

    def get_node(self):
        return self.cur

    def view_expression(self, e):
        return self.exprs[e]

    def set_udf(self, fn):
        self.udf = fn


def expr_ir(node, name):
    return types.SimpleNamespace(node=node, output_name=name)


def group_by_plan(agg_name="sum", n_keys=1, with_filter=True, rhs_literal=True, scan_selection=None, agg_options=None, gb_options=None):
    exprs = {0: make("Column", name="key"), 1: make("Column", name="x"), 2: make("Agg", name=agg_name, arguments=[1], options=agg_options), 3: make("Len"),
             4: make("BinaryExpr", left=1, op="Operator.Gt", right=5), 5: make("Literal", value=0) if rhs_literal else make("Column", name="y"),
             6: make("Column", name="key2")}
    scan = make("DataFrameScan", df=object(), projection=None, selection=scan_selection)
    nodes = {10: scan}
    child = 10
    if with_filter:
        nodes[11] = make("Filter", input=10, predicate=expr_ir(4, "p"))
        child = 11
    keys = [expr_ir(0, "key")] + ([expr_ir(6, "key2")] if n_keys == 2 else [])
    nodes[12] = make("GroupBy", input=child, keys=keys, aggs=[expr_ir(2, "x"), expr_ir(3, "len")], maintain_order=True,
                     options=gb_options if gb_options is not None else make("GroupbyOptions", slice=None, dynamic=None, rolling=None))
    return FakeTraverser(nodes, exprs, 12)


def test_filter_group_by_is_taken():
    for kwargs in (dict(with_filter=True), dict(with_filter=False), dict(n_keys=2)):
        nt = group_by_plan(**kwargs)
        engine.execute_with_b200(nt)
        assert callable(nt.udf) and nt.cur == 12          # the traverser is left on the replaced root


@pytest.mark.parametrize("kwargs", [dict(agg_name="median"), dict(rhs_literal=False), dict(scan_selection=object()),
                                    dict(agg_name="min", agg_options=True), dict(agg_name="max", agg_options=True),       # propagate_nans
                                    dict(gb_options=make("GroupbyOptions", slice=(0, 10), dynamic=None, rolling=None)),
                                    dict(gb_options=make("GroupbyOptions", slice=None, dynamic=object(), rolling=None)),
                                    dict(gb_options=make("GroupbyOptions", slice=None, dynamic=None, rolling=object()))])
def test_unsupported_group_by_shapes_are_left_to_polars(kwargs):
    nt = group_by_plan(**kwargs)
    engine.execute_with_b200(nt)
    assert nt.udf is None
    with pytest.raises(Exception):
        engine.execute_with_b200(group_by_plan(**kwargs), raise_on_fail=True)


def test_count_with_include_nulls_is_len():
    """pl.col(x).len() lowers to Agg count(include_nulls=True) (dsl/mod.rs:923-929): group length, not the non-null count."""
    nt = group_by_plan(agg_name="count", agg_options=True)
    assert engine._parse_agg(nt, expr_ir(2, "x")) == ("len", None, "x")
    nt = group_by_plan(agg_name="count", agg_options=False)
    assert engine._parse_agg(nt, expr_ir(2, "x")) == ("count", "x", "x")
    nt = group_by_plan(agg_name="min", agg_options=False)
    assert engine._parse_agg(nt, expr_ir(2, "x")) == ("min", "x", "x")


def test_var_std_ddof_first_last_n_unique():
    """Agg.options of var / std is ddof (visitor/expr_nodes.rs:1027-1040) and travels in the library's aggregation kind; first / last /
    n_unique carry no option; n_unique is a single-key aggregation of the library, so a two-key plan with it is left to Polars."""
    nt = group_by_plan(agg_name="var", agg_options=1)
    assert engine._parse_agg(nt, expr_ir(2, "x")) == ("var:1", "x", "x")
    nt = group_by_plan(agg_name="std", agg_options=0)
    assert engine._parse_agg(nt, expr_ir(2, "x")) == ("std:0", "x", "x")
    for name in ("first", "last", "n_unique"):
        nt = group_by_plan(agg_name=name)
        assert engine._parse_agg(nt, expr_ir(2, "x")) == (name, "x", "x")
        engine.execute_with_b200(nt)
        assert callable(nt.udf)
    for bad in (dict(agg_name="var", agg_options=None), dict(agg_name="std", agg_options=True), dict(agg_name="first", agg_options=1),
                dict(agg_name="n_unique", n_keys=2)):
        nt = group_by_plan(**bad)
        engine.execute_with_b200(nt)
        assert nt.udf is None, bad
    import polars_b200 as plb
    assert plb._agg_kind("var:0") == 8 and plb._agg_kind("std:2") == (9 | (2 << 16)) and plb._agg_kind("n_unique") == 10


def join_plan(how="Inner", n_keys=1, nulls_equal=False, slc=None, suffix="_right", coalesce=True, order="none", options=None):
    exprs = {0: make("Column", name="k"), 1: make("Column", name="k"), 2: make("Column", name="k2")}
    nodes = {20: make("DataFrameScan", df=object(), projection=None, selection=None), 21: make("DataFrameScan", df=object(), projection=["k", "r"], selection=None)}
    on = [expr_ir(0, "k")] + ([expr_ir(2, "k2")] if n_keys == 2 else [])
    nodes[22] = make("Join", input_left=20, input_right=21, left_on=on, right_on=[expr_ir(1, "k")] + on[1:],
                     options=options if options is not None else (how, nulls_equal, slc, suffix, coalesce, order))
    return FakeTraverser(nodes, exprs, 22)


def test_single_key_joins_are_taken_and_others_left():
    for how in ("Inner", "Left", "Semi", "Anti"):
        nt = join_plan(how)
        engine.execute_with_b200(nt)
        assert callable(nt.udf) and nt.cur == 22
    for nt in (join_plan("Full"), join_plan("Cross"), join_plan("Right"), join_plan("Inner", n_keys=2)):
        engine.execute_with_b200(nt)
        assert nt.udf is None


def test_join_options_are_read_not_guessed():
    """Every option of the Join node (visitor/nodes.rs:590-651) is either implemented by the UDF or the node is left to Polars."""
    taken = [join_plan("Inner", nulls_equal=True), join_plan("Inner", suffix="_r"), join_plan("Inner", order="left_right"),
             join_plan("Left", order="right"), join_plan("Semi", coalesce=False)]
    for nt in taken:
        engine.execute_with_b200(nt)
        assert callable(nt.udf)
    assert engine._join_options(("Inner", True, None, "_r", True, "right_left")) == ("inner", True, "_r", "right_left")
    left = [join_plan("Inner", slc=(0, 5)), join_plan("Inner", coalesce=False), join_plan("Left", coalesce=False), join_plan("Semi", order="left"),
            join_plan(options=("Inner", False)),                                                   # wrong arity
            # an asof join whose `left_by` column happens to be called "left": how is a tuple, never a substring match
            join_plan(options=(("AsOf", "backward", None, None, ["left"], ["inner"], True, True), False, None, "_right", True, "none"))]
    for nt in left:
        engine.execute_with_b200(nt)
        assert nt.udf is None
        with pytest.raises(Exception):
            engine.execute_with_b200(nt, raise_on_fail=True)


def test_other_roots_are_untouched():
    nt = FakeTraverser({1: make("Sort", input=0)}, {}, 1)
    engine.execute_with_b200(nt)
    assert nt.udf is None
