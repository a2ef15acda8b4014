"""GPU parity for toBool of an ETString filter expression (tq_vec_filter_string, csrc/strbool.cu + strnum.cuh) against the
oracle's literal restatement of types.StrToInt: the reference's own vectors, then a differential fuzz over the alphabet the
prefix scanner distinguishes."""
import numpy as np
import pytest

import oracle_py as O
from strnum_cases import ALL, fuzz_strings
from tinysql_b200 import _lib as L
from tinysql_b200 import expression as E
from tinysql_b200.chunk import BYTES, Column

pytestmark = pytest.mark.gpu


def check(cells):
    a = Column(BYTES, cells)
    want, err = O.vec_filter_string(a)
    if err:
        with pytest.raises(L.TQError) as ei:
            E.vectorized_filter_string(a)
        assert ei.value.status == L.TQ_ERR_OVERFLOW_BIGINT
    else:
        got = E.vectorized_filter_string(a)
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, [(cells[i], int(got[i]), int(want[i])) for i in bad[:5]]


def test_reference_vectors(lib):
    cells = [s for s, _ in ALL]
    got = E.vectorized_filter_string(Column(BYTES, cells))
    assert list(got) == [1 if v != 0 else 0 for _, v in ALL]
    check(cells + [None, b"  12  ", b"0.49999", b"0.5"])


def test_error_of_the_last_non_null_row(lib):
    check([b"1", None, b"99999999999999999999", b"0", None])      # an overflow in the middle is overwritten: no error
    check([b"1", b"99999999999999999999", None])                   # ... at the last non-NULL row it is reported
    check([b"1.5e30"])
    check([])


@pytest.mark.parametrize("seed", range(4))
def test_differential_fuzz(lib, seed, n=20000):
    rng = np.random.default_rng(seed)
    cells = fuzz_strings(rng, n)
    cells = [None if rng.random() < 0.05 else c for c in cells]
    # row by row first (the failing rows of a column-wide run would hide behind the last-row error rule) ...
    vals = [O.str_to_int(c) for c in cells if c is not None]
    ok_cells = [c for c in cells if c is None or not O.str_to_int(c)[1]]
    check(ok_cells)                                   # ... a column without failing rows: every selected[] value is compared
    assert any(e for _, e in vals)
    for c in [c for c in cells if c is not None and O.str_to_int(c)[1]][:200]:
        check([b"1", c])                              # each failing row as the last row: the error must surface
        check([c, b"0"])                              # ... and as a non-last row: no error, selected = (value != 0)
