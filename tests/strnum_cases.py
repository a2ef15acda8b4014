"""Known answers of types.StrToInt transcribed from the reference's own tests (types/convert_test.go), as (input, value) with no
error expected; used by the oracle test (CPU) and the device test (GPU / emulation)."""
# TestStrToNum :199-212 (the rows with truncateAsErr == false, i.e. truncation is not an error — a SELECT's statement context)
STR_TO_INT = [(b"0", 0), (b"-1", -1), (b"100", 100), (b"65.0", 65), (b"", 0), (b"xx", 0), (b"11xx", 11), (b"xx11", 0)]
# TestGetValidFloat tests2 :558-577 — floatStrToIntStr(origin) == expected; StrToInt = ParseInt(expected)
FLOAT_STR_TO_INT = [(b"1e9223372036854775807", 1), (b"125e342", 125), (b"1e21", 1), (b"1e5", 100000), (b"-123.45678e5", -12345678), (b"+0.5", 1),
                    (b"-0.5", -1), (b".5e0", 1), (b"+.5e0", 1), (b"-.5e0", -1), (b".5", 1), (b"123.456789e5", 12345679), (b"123.456784e5", 12345678),
                    (b"+999.9999e2", 100000)]
# TestGetValidFloat :530-556 — the valid float prefix, through floatStrToIntStr and ParseInt
VALID_PREFIX = [(b"-100", -100), (b"1abc", 1), (b"-1-1", -1), (b"+1+1", 1), (b"123..34", 123), (b"123.23E-10", 0), (b"1.1e1.3", 11), (b"11e1.3", 110),
                (b"1.1e-13a", 0), (b"1.", 1), (b".1", 0), (b"123e+", 123), (b"123.e", 123)]
ALL = STR_TO_INT + FLOAT_STR_TO_INT + VALID_PREFIX


def fuzz_strings(rng, n):
    """strings over the alphabet the prefix scanner distinguishes, plus structured numbers around the int64 limits"""
    alphabet = [b"0", b"1", b"4", b"5", b"9", b"9", b".", b"e", b"E", b"+", b"-", b" ", b"x", b"\t"]
    out = []
    for _ in range(n):
        kind = rng.integers(0, 6)
        if kind == 0:
            out.append(b"".join(alphabet[i] for i in rng.integers(0, len(alphabet), int(rng.integers(0, 12)))))
        elif kind == 1:   # sign, digits, optional fraction, optional exponent
            s = [b"", b"+", b"-"][int(rng.integers(0, 3))] + b"".join(b"%d" % d for d in rng.integers(0, 10, int(rng.integers(0, 24))))
            if rng.random() < 0.6:
                s += b"." + b"".join(b"%d" % d for d in rng.integers(0, 10, int(rng.integers(0, 6))))
            if rng.random() < 0.5:
                s += [b"e", b"E"][int(rng.integers(0, 2))] + [b"", b"+", b"-"][int(rng.integers(0, 3))] + b"%d" % int(rng.integers(0, 30))
            out.append(s)
        elif kind == 2:   # around the BIGINT limits
            base = [9223372036854775807, 9223372036854775808, 18446744073709551615, 922337203685477580][int(rng.integers(0, 4))] + int(rng.integers(-2, 3))
            s = [b"", b"-", b"+"][int(rng.integers(0, 3))] + b"%d" % base
            s += [b"", b".4", b".5", b".49", b"e0", b"e1", b"e-1", b"e-19"][int(rng.integers(0, 8))]
            out.append(s)
        elif kind == 3:   # white space and trailing garbage
            out.append(b" \t" + b"%d" % int(rng.integers(-50, 50)) + [b"", b" ", b"abc", b" 7", b"\n"][int(rng.integers(0, 5))])
        elif kind == 4:   # tiny magnitudes: rounding to 0 or 1
            out.append([b"0.4", b"0.5", b"-0.4", b"-0.5", b"0.49999", b"4e-1", b"5e-1", b"-5e-1", b"+5e-1", b"49e-2", b"5e-2", b".5e1", b"0", b"-0", b"00", b"-00.4",
                        b"0e5", b"-0e5", b"0.0e-3", b"1e-30", b"1e30", b"1.5e30", b"-1e25", b"1e99999999999999999999", b"1e-99999999999999999999"][int(rng.integers(0, 25))])
        else:
            out.append(bytes(rng.integers(0, 256, int(rng.integers(0, 8)), dtype="uint8")))
    return out
