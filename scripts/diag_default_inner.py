import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_py as O
from tinysql_b200 import _lib as L
from tinysql_b200.chunk import FLOAT64, INT64, Column
from tinysql_b200.executor import LEFT_OUTER_JOIN, HashJoinExec, MockDataSource
L.check(L.load().tq_init(0))
b = [Column(INT64, [1, 2, 2]), Column(INT64, [10, 20, 21]), Column(FLOAT64, [0.5, 1.5, 2.5])]
p = [Column(INT64, [2, 3, 1, 9], [True, True, True, True]), Column(INT64, [100, 101, 102, 103])]
bt, pt = [INT64, INT64, FLOAT64], [INT64, INT64]
for conds in ((), [(0, 3, 1)]):
    e = HashJoinExec(MockDataSource(pt, p), MockDataSource(bt, b), [0], [0], LEFT_OUTER_JOIN, False, None, 0, other_conditions=conds, default_inner=[None, 0, 2.5])
    e.Open(); got = e.drain(); e.Close()
    want = O.hash_join(LEFT_OUTER_JOIN, False, bt, b, pt, p, [0], [0], None, conds, default_inner=[None, 0, 2.5])
    print("conds", conds); print(" got ", got.rows()); print(" want", want.rows())
