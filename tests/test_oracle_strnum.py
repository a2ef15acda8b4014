"""CPU suite: the oracle's literal restatement of types.StrToInt (types/convert.go:224-405, the SELECT statement context) against
the reference's own vectors, plus a few properties the Go code implies."""
import oracle_py as O
from strnum_cases import ALL
from tinysql_b200.chunk import BYTES, Column


def test_reference_vectors():
    for s, want in ALL:
        assert O.str_to_int(s) == (want, False), s


def test_limits_and_quirks():
    assert O.str_to_int(b"9223372036854775807") == (9223372036854775807, False)
    assert O.str_to_int(b"9223372036854775808") == (9223372036854775807, True)       # ParseInt range error -> ErrOverflow
    assert O.str_to_int(b"-9223372036854775808") == (-9223372036854775808, False)
    assert O.str_to_int(b"-9223372036854775809") == (-9223372036854775808, True)
    assert O.str_to_int(b"9223372036854775807.5") == (9223372036854775807, True)      # rounds up to 2^63
    assert O.str_to_int(b"1.5e30") == (0, True)          # intCnt > 21: "1.5" goes to ParseInt (floatStrToIntStr :369-376)
    assert O.str_to_int(b"125e342") == (125, False)      # ... and "125" parses (the reference's own vector)
    assert O.str_to_int(b"1-5") == (0, True)             # the prefix scan lets a sign through at index 1 (eIdx starts at 0)
    assert O.str_to_int(b"  12  ") == (12, False)        # strings.TrimSpace
    assert O.str_to_int(b"0.49999") == (0, False) and O.str_to_int(b"0.5") == (1, False)   # roundIntStr looks at ONE digit


def test_filter_takes_the_error_of_the_last_non_null_row():
    a = Column(BYTES, [b"1", None, b"99999999999999999999", b"0", None])
    sel, err = O.vec_filter_string(a)
    assert list(sel) == [1, 0, 1, 0, 0] and err is False           # the overflow row is not the last one: err = err1 overwrote it
    a = Column(BYTES, [b"1", b"99999999999999999999", None])
    sel, err = O.vec_filter_string(a)
    assert list(sel) == [1, 1, 0] and err is True
