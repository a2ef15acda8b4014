"""GPU parity for the string (var-len column) builtins — builtin{LT..NE}StringSig, builtinStrcmpSig, builtinLengthSig,
builtinStringIsNullSig — against the oracle's restatement (types.CompareString byte order, MergeNulls)."""
import numpy as np
import pytest

import oracle_py as O
from tinysql_b200 import expression as E
from tinysql_b200.chunk import BYTES, Column
from util import assert_col_equal

pytestmark = pytest.mark.gpu


def rand_strings(rng, n, max_len, null_frac, alphabet=4):
    """short alphabet + shared prefixes so that ties, prefixes and long common runs actually occur"""
    out = []
    for i in range(n):
        if rng.random() < null_frac:
            out.append(None)
            continue
        l = int(rng.integers(0, max_len + 1))
        out.append(rng.integers(97, 97 + alphabet, l, dtype=np.uint8).tobytes())
    return out


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 1000, 20011])
def test_string_compare_and_unary(lib, n):
    rng = np.random.default_rng(n + 5)
    a_cells, b_cells = rand_strings(rng, n, 70, 0.1), rand_strings(rng, n, 70, 0.1)
    for i in range(0, n, 3):  # equal strings and proper prefixes
        if a_cells[i] is not None:
            b_cells[i] = a_cells[i] if i % 2 else a_cells[i][: len(a_cells[i]) // 2]
    a, b = Column(BYTES, a_cells), Column(BYTES, b_cells)
    for op in range(7):
        rc, want = O.vec_compare_string(op, a, b)
        assert rc == 0
        assert_col_equal(E.vec_compare_string(op, a, b), want, check_null_slots=True)
    for op in (E.STR_LENGTH, E.STR_ISNULL):
        rc, want = O.vec_string_unary(op, a)
        assert rc == 0
        assert_col_equal(E.vec_string_unary(op, a), want, check_null_slots=True)


def test_string_compare_long_cells_and_high_bytes(lib):
    """5 KiB cells differing only in the last byte / by length; bytes >= 0x80 compare as UNSIGNED (Go string order)"""
    base = bytes(range(256)) * 20
    a = Column(BYTES, [base, base, base + b"x", base, b"\xff", b"\x7f", "你好".encode(), b""])
    b = Column(BYTES, [base, base[:-1] + b"\x00", base, base + b"x", b"\x7f", b"\xff", "你".encode(), b""])
    for op in range(7):
        rc, want = O.vec_compare_string(op, a, b)
        assert_col_equal(E.vec_compare_string(op, a, b), want, check_null_slots=True)
    assert E.vec_compare_string(E.STRCMP, a, b).tolist() == [0, 1, 1, -1, 1, -1, 1, 0]
    assert E.vec_string_unary(E.STR_LENGTH, a).tolist() == [5120, 5120, 5121, 5120, 1, 1, 6, 0]


@pytest.mark.parametrize("n", [0, 1, 33, 1000, 20011])
def test_string_if_ifnull_in(lib, n):
    """builtinIfStringSig / builtinIfNullStringSig / builtinInStringSig (builtin_control_vec_generated.go:81-112,209-262,
    builtin_other_vec_generated.go:97-149) against the oracle"""
    from tinysql_b200.chunk import INT64
    rng = np.random.default_rng(n + 9)
    a, b = Column(BYTES, rand_strings(rng, n, 40, 0.2, alphabet=2)), Column(BYTES, rand_strings(rng, n, 40, 0.2, alphabet=2))
    cond = Column(INT64, rng.integers(-1, 2, n), rng.random(n) > 0.2)
    rc, want = O.vec_pick_string(0, cond, a, b)
    assert rc == 0 and E.vec_if_string(cond, a, b).tolist() == want.tolist()
    rc, want = O.vec_pick_string(1, None, a, b)
    assert rc == 0 and E.vec_ifnull_string(a, b).tolist() == want.tolist()
    short = [Column(BYTES, rand_strings(rng, n, 3, 0.15, alphabet=2)) for _ in range(4)]
    x = Column(BYTES, rand_strings(rng, n, 3, 0.15, alphabet=2))
    for k in (1, 4):
        rc, want = O.vec_in_string(x, short[:k])
        assert rc == 0
        assert_col_equal(E.vec_in_string(x, short[:k]), want, check_null_slots=True)


@pytest.mark.parametrize("n", [0, 5, 64, 10007])
def test_in_real_and_real_to_bool(lib, n):
    """builtinInRealSig (builtin_other_vec_generated.go:151-204); toBool for ETReal (expression.go:296-307, RoundFloat)"""
    from tinysql_b200.chunk import FLOAT64
    rng = np.random.default_rng(n)
    vals = np.array([0.0, -0.0, 0.25, 0.5, -0.5, 0.49999999999999994, 1.5, np.nan, np.inf, -3.0])
    a = Column(FLOAT64, rng.choice(vals, n), rng.random(n) > 0.15)
    lst = [Column(FLOAT64, rng.choice(vals, n), rng.random(n) > 0.15) for _ in range(3)]
    for k in (1, 3):
        rc, want = O.vec_in_real(a, lst[:k])
        assert rc == 0
        assert_col_equal(E.vec_in_real(a, lst[:k]), want, check_null_slots=True)
    assert np.array_equal(E.vectorized_filter_real(a), O.vec_filter_real(a))
