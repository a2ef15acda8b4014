// strnum.cuh — types.StrToInt (types/convert.go:224-232) for one var-len cell, in streaming form: the value test toBool needs
// for an ETString filter expression (expression/expression.go:308-322).
//
// The reference builds three intermediate strings — TrimSpace(s), the valid float prefix (getValidFloatPrefix :430-475), the
// integer string (floatStrToIntStr :318-405, roundIntStr :283-311) — and hands the last one to strconv.ParseInt.  A thread
// cannot allocate strings, and does not need to: the integer string is always
//     [an optional '-'] ++ the first `cnt` bytes of (prefix[a0, a1) ++ prefix[b0, b1)) ++ `zeros` x '0',  magnitude + 1 if `round`
// so ParseInt is evaluated over that description directly, one pass, with a saturating 64-bit accumulator.  ParseInt failing
// (a non-digit after the sign — the prefix scan lets a sign through at index 1, "1-5" — or a value outside int64) is what
// StrToInt reports as ErrOverflow("BIGINT"); the value is then 0 (syntax) or MaxInt64 / MinInt64 (range).
// Statement context = a SELECT (InSelectStmt: "" is "0"; truncation is a warning, not an error; CastStrToIntStrict == false).
// strings.TrimSpace is restated for the ASCII white space characters.
#pragma once
#include <cstdint>

namespace tqd {

struct IntStrView {
  const uint8_t *p;     // the valid float prefix
  int64_t a0, a1, b0, b1;
  int64_t cnt;          // bytes taken from a ++ b
  int64_t zeros;        // '0' bytes appended
  bool neg_prefix;      // an explicit '-' in front
  bool round;           // roundIntStr fired: magnitude + 1
};

// byte j of a ++ b
__host__ __device__ inline uint8_t view_byte(const IntStrView &v, int64_t j) {
  const int64_t la = v.a1 - v.a0;
  return j < la ? v.p[v.a0 + j] : v.p[v.b0 + (j - la)];
}

// strconv.ParseInt(view, 10, 64).  Returns the value; *fail = 1 syntax error (value 0) or 2 range error (value = max / min).
__host__ __device__ inline int64_t parse_int_view(const IntStrView &v, int *fail) {
  bool neg = v.neg_prefix, have_sign = v.neg_prefix, syntax = false, ovf = false;
  uint64_t acc = 0;
  int64_t ndig = 0;
  for (int64_t j = 0; j < v.cnt; j++) {
    const uint8_t c = view_byte(v, j);
    if (j == 0 && !have_sign && (c == '+' || c == '-')) { neg = c == '-'; continue; }
    if (c < '0' || c > '9') { syntax = true; break; }
    const uint64_t d = (uint64_t)(c - '0');
    if (acc > (0xFFFFFFFFFFFFFFFFull - d) / 10) ovf = true; else if (!ovf) acc = acc * 10 + d;
    ndig++;
  }
  if (!syntax) {
    for (int64_t z = 0; z < v.zeros; z++) { if (acc > 0xFFFFFFFFFFFFFFFFull / 10) ovf = true; else if (!ovf) acc *= 10; ndig++; }
    if (ndig == 0) syntax = true;   // "" or a lone sign
  }
  if (syntax) { *fail = 1; return 0; }
  if (v.round) { if (acc == 0xFFFFFFFFFFFFFFFFull) ovf = true; else if (!ovf) acc++; }
  if (!neg && (ovf || acc > 0x7FFFFFFFFFFFFFFFull)) { *fail = 2; return 0x7FFFFFFFFFFFFFFFll; }
  if (neg && (ovf || acc > 0x8000000000000000ull)) { *fail = 2; return (int64_t)0x8000000000000000ull; }
  *fail = 0;
  return neg ? (int64_t)(0 - acc) : (int64_t)acc;
}

// strconv.Atoi of prefix[lo, hi) (the exponent): *fail != 0 on syntax / range error
__host__ __device__ inline int64_t atoi_range(const uint8_t *p, int64_t lo, int64_t hi, int *fail) {
  IntStrView v{p, lo, hi, 0, 0, hi - lo, 0, false, false};
  return parse_int_view(v, fail);
}

// types.StrToInt(sc, s): the value; *overflow_err = ParseInt failed
__host__ __device__ inline int64_t str_to_int(const uint8_t *s, int64_t n, int *overflow_err) {
  while (n > 0 && (s[0] == ' ' || (s[0] >= '\t' && s[0] <= '\r'))) { s++; n--; }   // strings.TrimSpace (ASCII)
  while (n > 0 && (s[n - 1] == ' ' || (s[n - 1] >= '\t' && s[n - 1] <= '\r'))) n--;
  // getValidFloatPrefix (:436-474).  eIdx starts at 0, so "i != eIdx + 1" lets a sign through at index 1 — kept as written.
  bool saw_dot = false, saw_digit = false;
  int64_t vl = 0, e_scan = 0;
  for (int64_t i = 0; i < n; i++) {
    const uint8_t c = s[i];
    if (c == '+' || c == '-') { if (i != 0 && i != e_scan + 1) break; }
    else if (c == '.') { if (saw_dot || e_scan > 0) break; saw_dot = true; if (saw_digit) vl = i + 1; }
    else if (c == 'e' || c == 'E') { if (!saw_digit) break; if (e_scan != 0) break; e_scan = i; }
    else if (c < '0' || c > '9') break;
    else { saw_digit = true; vl = i + 1; }
  }
  *overflow_err = 0;
  if (vl == 0) return 0;   // valid = "0" (also the empty string of a SELECT)
  // floatStrToIntStr (:319-328): the LAST '.' and 'e' / 'E' inside the prefix
  int64_t dot = -1, eidx = -1;
  for (int64_t i = 0; i < vl; i++) { const uint8_t c = s[i]; if (c == '.') dot = i; else if (c == 'e' || c == 'E') eidx = i; }
  IntStrView v{s, 0, 0, 0, 0, 0, 0, false, false};
  if (eidx == -1) {
    if (dot == -1) { v.a1 = vl; v.cnt = vl; }                                  // :330-332 the prefix itself
    else {                                                                      // :333-351
      const int64_t off = (s[0] == '-' || s[0] == '+') ? 1 : 0;
      const int64_t d = dot - off, dl = vl - off;
      if (d == 0) v.zeros = 1;                                                  // intStr = "0"
      else { v.a0 = off; v.a1 = off + d; v.cnt = d; }
      if (dl > d + 1) v.round = s[off + d + 1] >= '5';
      // "-" is put back unless intStr is the single character '0'; on a zero magnitude the sign changes nothing
      v.neg_prefix = s[0] == '-' && !(v.cnt <= 1 && !v.round && (v.cnt == 0 || s[v.a0] == '0'));
    }
  } else {
    // digits = prefix[:dot] ++ prefix[dot+1:eidx]  (or prefix[:eidx]); intCnt = len(integer part incl. sign) + exponent
    int64_t int_cnt, dlen;
    if (dot == -1) { v.a0 = 0; v.a1 = eidx; int_cnt = eidx; dlen = eidx; }
    else { v.a0 = 0; v.a1 = dot; v.b0 = dot + 1; v.b1 = eidx; int_cnt = dot; dlen = dot + (eidx - dot - 1); }
    int afail = 0;
    const int64_t exp = atoi_range(s, eidx + 1, vl, &afail);
    if (afail) { *overflow_err = 1; return 0; }                                 // :364-367 returns the prefix ("…e…"): ParseInt syntax error
    int_cnt = (int64_t)((uint64_t)int_cnt + (uint64_t)exp);                     // Go's int addition wraps
    if (exp >= 0 && (int_cnt > 21 || int_cnt < 0)) { v.a0 = 0; v.a1 = eidx; v.b0 = v.b1 = 0; v.cnt = eidx; }   // :369-376 prefix[:eidx] (a '.' in it fails ParseInt)
    else if (int_cnt <= 0) {                                                    // :377-383
      v.zeros = 1;
      if (int_cnt == 0 && dlen > 0) { const uint8_t c = view_byte(v, 0); v.round = c >= '0' && c <= '9' && c >= '5'; }
    } else if (int_cnt == 1 && (view_byte(v, 0) == '-' || view_byte(v, 0) == '+')) {   // :384-393 (digits[0]: prefix[0], or prefix[1] for ".+5e1")
      const bool minus = view_byte(v, 0) == '-';
      v.zeros = 1;
      if (dlen > 1) v.round = view_byte(v, 1) >= '5';
      v.neg_prefix = v.round && minus;                                          // the sign is put back only in front of "1"
    } else if (int_cnt <= dlen) {                                               // :394-398
      v.cnt = int_cnt;
      if (int_cnt < dlen) v.round = view_byte(v, int_cnt) >= '5';
    } else { v.cnt = dlen; v.zeros = int_cnt - dlen; }                          // :399-403 scientific notation with extra zeros
  }
  int fail = 0;
  const int64_t val = parse_int_view(v, &fail);
  *overflow_err = fail != 0;
  return val;
}

}  // namespace tqd
