#!/usr/bin/env python
"""Writes tests/golden/reference_cases.json: the known-answer cases the REFERENCE's own test-suite holds for the hot
path, transcribed by hand — the Go tests cannot run here (no Go toolchain; fetchAndBuildHashTable /
runJoinWorker / shuffleIntermData / consumeIntermData are course stubs), so nothing in this file is computed by our
code: every `expect` is the literal the reference test asserts, with its file:line.  The CPU suite checks the oracle
against these; the GPU suite checks the CUDA path (through the C-ABI) against the same file.

Run:  python tests/golden/make_golden.py      (rewrites reference_cases.json deterministically)
"""
import json
import os

I, U, F, B = "int64", "uint64", "float64", "bytes"
N = None
MAXU = (1 << 64) - 1

# ---- executor/join_test.go TestJoin (:36-182).  Tables as lists of rows; `lhs` / `rhs` in SQL order (output row =
# lhs columns ++ rhs columns, joiner.go:145-150); lkey / rkey = join key column indices; `where` = the statement's WHERE
# clause (or inner-join other-condition) as a Python predicate over the joined row r (None = NULL), applied by the
# test AFTER the operator, the way the plan puts a Selection above the join; `outer_selected` = an ON-clause condition
# on the outer side only, which reaches the executor as outerSideFilter / selected[] (join.go:328); `select` =
# projected output columns; ordered = the test compares without .Sort() / ORDER BY.
t, t1 = [[1, 1], [2, 2]], [[2, 3], [4, 4]]
ta, tb = [[1, 1], [2, 2], [3, 3]], [[1, 2], [1, 3], [1, 4], [3, 4], [4, 5]]
seven = [[i] for i in range(1, 8)]
JOIN = [
    dict(name="left_outer_where", cite="executor/join_test.go:74-75", lhs=t, rhs=t1, lkey=[0], rkey=[0], type="left",
         where="r[0] == 1 or (r[3] is not None and r[3] > 20)", expect=[[1, 1, N, N]]),
    dict(name="right_outer_t1_t_where", cite="executor/join_test.go:76-77", lhs=t1, rhs=t, lkey=[0], rkey=[0], type="right",
         where="r[2] == 1 or (r[1] is not None and r[1] > 20)", expect=[[N, N, 1, 1]]),
    dict(name="right_outer_t_t1_where", cite="executor/join_test.go:78-79", lhs=t, rhs=t1, lkey=[0], rkey=[0], type="right",
         where="(r[0] is not None and r[0] == 1) or r[3] > 20", expect=[]),
    dict(name="left_outer_where_false", cite="executor/join_test.go:80-81", lhs=t, rhs=t1, lkey=[0], rkey=[0], type="left",
         where="r[2] is not None and r[2] == 3", expect=[]),
    dict(name="left_outer_on_clause_outer_condition", cite="executor/join_test.go:82-83", lhs=t, rhs=t1, lkey=[0], rkey=[0], type="left",
         outer_selected=[0, 1], expect=[[1, 1, N, N], [2, 2, 2, 3]]),
    dict(name="self_join_3x3_duplicates", cite="executor/join_test.go:100-104", lhs=[[1], [1], [1]], rhs=[[1], [1], [1]], lkey=[0], rkey=[0],
         type="inner", expect=[[1, 1]] * 9),
    dict(name="self_join_1_to_7", cite="executor/join_test.go:111-113", lhs=seven, rhs=seven, lkey=[0], rkey=[0], type="inner", select=[0],
         expect=[[i] for i in range(1, 8)]),
    dict(name="self_join_sum_gt_5", cite="executor/join_test.go:115-116", lhs=seven, rhs=seven, lkey=[0], rkey=[0], type="inner",
         where="r[0] + r[1] > 5", select=[0], expect=[[3], [4], [5], [6], [7]]),
    dict(name="multi_match_probe_then_insertion_order", cite="executor/join_test.go:134-136", lhs=ta, rhs=tb, lkey=[0], rkey=[0],
         type="inner", build="rhs", expect=[[1, 1, 1, 2], [1, 1, 1, 3], [1, 1, 1, 4], [3, 3, 3, 4]], ordered=True),
    dict(name="inner_other_condition", cite="executor/join_test.go:137-139", lhs=tb, rhs=ta, lkey=[0], rkey=[0], type="inner",
         where="r[2] < r[1]", expect=[[1, 2, 1, 1], [1, 3, 1, 1], [1, 4, 1, 1], [3, 4, 3, 3]]),
    dict(name="right_outer_unmatched_row", cite="executor/join_test.go:144-146", lhs=ta, rhs=tb, lkey=[0], rkey=[0], type="right",
         expect=[[1, 1, 1, 2], [1, 1, 1, 3], [1, 1, 1, 4], [3, 3, 3, 4], [N, N, 4, 5]]),
    dict(name="issue5255_varchar_and_float_payload", cite="executor/join_test.go:337-346 (t1(a int, b varchar(64), c float) join t2(a))",
         lhs=[[1, "2017-11-29", 2.2]], rhs=[[1]], ltypes=[I, B, "float32"], rtypes=[I], lkey=[0], rkey=[0], type="inner",
         expect=[[1, "2017-11-29", 2.2, 1]]),
    dict(name="two_column_key_self_join", cite="executor/join_test.go:381-388 (t t1 join t t2 on t1.b = t2.b and t1.a = t2.a)",
         lhs=[[1, 1], [1, 2], [2, 1], [2, 2]], rhs=[[1, 1], [1, 2], [2, 1], [2, 2]], lkey=[1, 0], rkey=[1, 0], type="inner",
         expect=[[1, 1, 1, 1], [1, 2, 1, 2], [2, 1, 2, 1], [2, 2, 2, 2]]),
    dict(name="left_outer_no_key_match_with_outer_condition", cite="executor/join_test.go:372-379 (t1 left outer join t2 on t1.a=t2.a and t1.a!=3)",
         lhs=[[i, 100] for i in range(1, 6)], rhs=[[i * 100, 10000] for i in range(1, 6)], lkey=[0], rkey=[0], type="left",
         outer_selected=[1, 1, 0, 1, 1], expect=[[i, 100, N, N] for i in range(1, 6)]),
    dict(name="issue5278_second_left_join", cite="executor/join_test.go:348-356 ((t left join tt) left join t ttt on t.a=ttt.a; tt is empty)",
         lhs=[[1, 1, N, N]], rhs=[[1, 1]], lkey=[0], rkey=[0], type="left", expect=[[1, 1, N, N, 1, 1]]),
    dict(name="inner_100x100_limit_1_then_close", cite="executor/join_test.go:175-182", lhs=[[1]] * 100, rhs=[[1]] * 100, lkey=[0], rkey=[0],
         type="inner", limit=1, expect=[[1, 1]], total_rows=10000),
]

# ---- executor/aggfuncs/*_test.go + executor/aggregate_test.go + executor/executor_test.go.
# input columns: list of (type, values); funcs: (name, arg column or -1 for a constant non-NULL argument)
five_i, five_f = [0, 1, 2, 3, 4], [0.0, 1.0, 2.0, 3.0, 4.0]
AGG = [
    dict(name="sum_avg_count_int", cite="aggfuncs/func_sum_test.go, func_avg_test.go, func_count_test.go (harness aggfunc_test.go:69-205)",
         cols=[[I, five_i]], group_by=[], funcs=[["sum", 0], ["avg", 0], ["count", 0]], expect=[[10, 2, 5]]),
    dict(name="sum_avg_count_double", cite="aggfuncs/func_sum_test.go, func_avg_test.go", cols=[[F, five_f]], group_by=[],
         funcs=[["sum", 0], ["avg", 0], ["count", 0]], expect=[[10.0, 2.0, 5]]),
    dict(name="empty_input_scalar", cite="aggfuncs/aggfunc_test.go:176-190; executor/aggregate_test.go:58", cols=[[I, []]], group_by=[],
         funcs=[["sum", 0], ["avg", 0], ["count", 0], ["max", 0], ["min", 0]], expect=[[N, N, 0, N, N]]),
    dict(name="empty_input_group_by", cite="executor/aggregate_test.go:57", cols=[[I, []]], group_by=[0], funcs=[["count", 0]], expect=[]),
    dict(name="merge_partials_int", cite="aggfuncs/func_sum_test.go TestMergePartialResult4Sum, func_avg_test.go TestMergePartialResult4Avg",
         cols=[[I, five_i + [2, 3, 4]]], group_by=[], funcs=[["sum", 0], ["avg", 0]], expect=[[19, 2]], partial_workers=2),
    dict(name="merge_partials_double", cite="aggfuncs/func_avg_test.go TestMergePartialResult4Avg", cols=[[F, five_f + [2.0, 3.0, 4.0]]],
         group_by=[], funcs=[["sum", 0], ["avg", 0]], expect=[[19.0, 2.375]], partial_workers=2),
    dict(name="max_min_first_row", cite="aggfuncs/func_max_min_test.go, func_first_row_test.go", cols=[[I, five_i]], group_by=[],
         funcs=[["max", 0], ["min", 0], ["firstrow", 0]], expect=[[4, 0, 0]]),
    dict(name="min_max_with_null", cite="executor/aggregate_test.go:74-81 TestAggEliminator", cols=[[I, [1, 2, 3, 4]], [I, [-1, -2, 1, N]]],
         group_by=[], funcs=[["max", 0], ["min", 1]], expect=[[4, -2]]),
    dict(name="min_max_b_times_b", cite="executor/aggregate_test.go:79-80 (b*b pre-projected)", cols=[[I, [1, 4, 1, N]]], group_by=[],
         funcs=[["max", 0], ["min", 0]], expect=[[4, 1]]),
    dict(name="count_one_row_group_by", cite="executor/aggregate_test.go:60-62", cols=[[I, [0]], [I, [0]]], group_by=[0], funcs=[["count", 1]],
         expect=[[1]]),
    dict(name="count_group_by_a_b", cite="executor/aggregate_test.go:64-65 (where b>0 pre-filtered)",
         cols=[[I, [1, 3, 3, 2, 1, 1]], [I, [1, 3, 2, 1, 1, 1]]], group_by=[0, 1], funcs=[["count", 0]], expect=[[1], [1], [1], [3]]),
    dict(name="count_star_group_by_c", cite="executor/executor_test.go:975 (rows (1,1,1),(2,1,1),(3,1,2),(4,2,3))",
         cols=[[I, [1, 2, 3, 4]], [I, [1, 1, 1, 2]], [I, [1, 1, 2, 3]]], group_by=[2], funcs=[["count", -1], ["firstrow", 2]],
         expect=[[2, 1], [1, 2], [1, 3]]),
    dict(name="sum_c_group_by_b", cite="executor/executor_test.go:976", cols=[[I, [1, 2, 3, 4]], [I, [1, 1, 1, 2]], [I, [1, 1, 2, 3]]],
         group_by=[1], funcs=[["sum", 2]], expect=[[3], [4]]),
]

# ---- expression/builtin_*_test.go known answers for the vectorized builtins (one-row columns)
EXPR = [
    dict(op="plus", cite="expression/builtin_arithmetic_test.go:116-128", args=[[I, 12], [I, 1]], expect=[I, 13]),
    dict(op="plus", cite="expression/builtin_arithmetic_test.go:131-143", args=[[F, 1.01001], [F, -0.01]], expect=[F, 1.00001]),
    dict(op="plus", cite="expression/builtin_arithmetic_test.go:146-158", args=[[F, N], [F, -0.11101]], expect=[F, N]),
    dict(op="plus", cite="expression/builtin_arithmetic_test.go:161-173", args=[[F, N], [F, N]], expect=[F, N]),
    dict(op="minus", cite="expression/builtin_arithmetic_test.go:178-190", args=[[I, 12], [I, 1]], expect=[I, 11]),
    dict(op="minus", cite="expression/builtin_arithmetic_test.go:193-205", args=[[F, 1.01001], [F, -0.01]], expect=[F, 1.02001]),
    dict(op="minus", cite="expression/builtin_arithmetic_test.go:208-220", args=[[F, N], [F, -0.11101]], expect=[F, N]),
    dict(op="minus", cite="expression/builtin_arithmetic_test.go:223-235", args=[[F, 1.01], [F, N]], expect=[F, N]),
    dict(op="mul", cite="expression/builtin_arithmetic_test.go:259-262", args=[[I, 11], [I, 11]], expect=[I, 121]),
    dict(op="mul", cite="expression/builtin_arithmetic_test.go:263-266", args=[[U, 11], [U, 11]], expect=[U, 121]),
    dict(op="mul", cite="expression/builtin_arithmetic_test.go:267-270", args=[[F, 11.0], [F, 11.0]], expect=[F, 121.0]),
    dict(op="mul", cite="expression/builtin_arithmetic_test.go:271-274", args=[[F, N], [F, -0.11101]], expect=[F, N]),
    dict(op="lt", cite="expression/builtin_compare_test.go:34", args=[[I, 1], [I, 1]], expect=[I, 0]),
    dict(op="lt", cite="expression/builtin_compare_test.go:36", args=[[F, 1.1], [F, 1.1]], expect=[I, 0]),
    dict(op="eq", cite="expression/builtin_compare_test.go:37", args=[[U, 1], [U, 1]], expect=[I, 1]),
    dict(op="in", cite="expression/builtin_other_test.go:32", args=[[I, 1], [I, 1], [I, 2], [I, 3]], expect=[I, 1]),
    dict(op="in", cite="expression/builtin_other_test.go:33", args=[[I, 1], [I, 0], [I, 2], [I, 3]], expect=[I, 0]),
    dict(op="in", cite="expression/builtin_other_test.go:34", args=[[I, 1], [I, N], [I, 2], [I, 3]], expect=[I, N]),
    dict(op="in", cite="expression/builtin_other_test.go:35", args=[[I, N], [I, N], [I, 2], [I, 3]], expect=[I, N]),
    dict(op="in", cite="expression/builtin_other_test.go:36", args=[[U, 0], [I, 0], [I, 2], [I, 3]], expect=[I, 1]),
    dict(op="in", cite="expression/builtin_other_test.go:37", args=[[U, MAXU], [U, MAXU], [I, 2], [I, 3]], expect=[I, 1]),
    dict(op="in", cite="expression/builtin_other_test.go:38", args=[[I, -1], [U, MAXU], [I, 2], [I, 3]], expect=[I, 0]),
    dict(op="in", cite="expression/builtin_other_test.go:39", args=[[U, MAXU], [I, -1], [I, 2], [I, 3]], expect=[I, 0]),
]

# ---- expression/builtin_string_test.go (string-typed argument cases; the int / float argument cases go through a CAST
# that is planner work, not part of the vectorized signature)
EXPR += [
    dict(op="length", cite="expression/builtin_string_test.go:31", args=[[B, "abc"]], expect=[I, 3]),
    dict(op="length", cite="expression/builtin_string_test.go:32", args=[[B, "\u4f60\u597d"]], expect=[I, 6]),
    dict(op="length", cite="expression/builtin_string_test.go:35", args=[[B, N]], expect=[I, N]),
    dict(op="strcmp", cite="expression/builtin_string_test.go:69", args=[[B, "123"], [B, "123"]], expect=[I, 0]),
    dict(op="strcmp", cite="expression/builtin_string_test.go:70", args=[[B, "123"], [B, "1"]], expect=[I, 1]),
    dict(op="strcmp", cite="expression/builtin_string_test.go:71", args=[[B, "1"], [B, "123"]], expect=[I, -1]),
    dict(op="strcmp", cite="expression/builtin_string_test.go:72", args=[[B, "123"], [B, "45"]], expect=[I, -1]),
    dict(op="strcmp", cite="expression/builtin_string_test.go:75", args=[[B, N], [B, "123"]], expect=[I, N]),
    dict(op="strcmp", cite="expression/builtin_string_test.go:76", args=[[B, "123"], [B, N]], expect=[I, N]),
    dict(op="strcmp", cite="expression/builtin_string_test.go:77", args=[[B, ""], [B, "123"]], expect=[I, -1]),
    dict(op="strcmp", cite="expression/builtin_string_test.go:78", args=[[B, "123"], [B, ""]], expect=[I, 1]),
    dict(op="strcmp", cite="expression/builtin_string_test.go:79", args=[[B, ""], [B, ""]], expect=[I, 0]),
    dict(op="strcmp", cite="expression/builtin_string_test.go:80", args=[[B, ""], [B, N]], expect=[I, N]),
    dict(op="strcmp", cite="expression/builtin_string_test.go:82", args=[[B, N], [B, N]], expect=[I, N]),
    dict(op="lt", cite="expression/builtin_compare_test.go:35", args=[[B, "123"], [B, "123"]], expect=[I, 0]),
]

# ---- expression/evaluator_test.go TestUnaryOp (:117-139), expression/builtin_test.go TestIsNullFunc (:54-68)
EXPR += [
    dict(op="neg", cite="expression/evaluator_test.go:124", args=[[F, N]], expect=[F, N]),
    dict(op="neg", cite="expression/evaluator_test.go:125", args=[[F, 1.0]], expect=[F, -1.0]),
    dict(op="neg", cite="expression/evaluator_test.go:126", args=[[I, 1]], expect=[I, -1]),
    dict(op="neg", cite="expression/evaluator_test.go:128", args=[[U, 1]], expect=[I, -1]),
    dict(op="isnull", cite="expression/builtin_test.go:56-60", args=[[I, 1]], expect=[I, 0]),
    dict(op="isnull", cite="expression/builtin_test.go:62-66", args=[[I, N]], expect=[I, 1]),
]

# ---- util/codec/codec_test.go TestHashChunkRow (:735-769) as join-key equalities
KEYEQ = [
    dict(cite="util/codec/codec_test.go:741-747", a=[U, 1], b=[I, 1], equal=True),
    dict(cite="util/codec/codec_test.go:749-755", a=[U, MAXU], b=[I, -1], equal=False),
]

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_cases.json")
    with open(out, "w") as f:
        json.dump(dict(join=JOIN, agg=AGG, expr=EXPR, key_equality=KEYEQ), f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote", out, len(JOIN), "join", len(AGG), "agg", len(EXPR), "expr cases")
