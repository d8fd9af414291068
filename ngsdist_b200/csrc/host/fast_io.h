// Host-side I/O helpers of the drop-in command line (SURVEY §8f rows N1 reader / N2 writer).
//
//   fmt_fixed10   the writer's "%.10f" (join(), shared/gen_func.cpp:479-496; matrix print ngsDist.cpp:282-287), exact:
//                 the decimal expansion of the binary double rounded half-to-even on the exact value, which is what
//                 glibc's printf does -- byte-identical output, ~20x faster than snprintf and thread-safe, so a
//                 101 x 2000^2 bootstrap output is no longer slower than the GPU work that produced it.
//   parse_number  the reader's strtod for the common plain-decimal tokens (<= 19 significant digits after stripping,
//                 no exponent): mantissa / 10^k with k <= 22 and mantissa < 2^53 is correctly rounded in one IEEE
//                 division (Clinger's fast path) and therefore equals strtod bit for bit; everything else goes to strtod.
#pragma once

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace fastio {

// writes v as printf("%.10f") would; returns the number of characters (no terminator)
static inline int fmt_fixed10(char *dst, double v) {
  uint64_t bits;
  memcpy(&bits, &v, 8);
  char *p = dst;
  const bool neg = bits >> 63;
  const int bexp = (int) ((bits >> 52) & 0x7FF);
  const uint64_t frac = bits & ((1ull << 52) - 1);
  if (bexp == 0x7FF) {                       // glibc: "inf" / "-inf" / "nan" / "-nan"
    if (neg) *p++ = '-';
    memcpy(p, frac ? "nan" : "inf", 3);
    return (int) (p + 3 - dst);
  }
  const uint64_t mant = bexp ? (frac | (1ull << 52)) : frac;
  const int e2 = (bexp ? bexp : 1) - 1075;  // value = mant * 2^e2
  if (e2 >= 0) return snprintf(dst, 400, "%.10f", v);          // >= 2^52: never a distance; leave it to libc
  const int sh = -e2;                        // 1 .. 1074
  unsigned __int128 q;
  if (sh >= 120) {
    q = 0;                                   // mant * 1e10 < 2^87: far below half a unit of the last place
  } else {
    const unsigned __int128 P = (unsigned __int128) mant * 10000000000ull;
    q = P >> sh;
    const unsigned __int128 rem = P & ((((unsigned __int128) 1) << sh) - 1), half = ((unsigned __int128) 1) << (sh - 1);
    if (rem > half || (rem == half && (q & 1))) q++;
  }
  const uint64_t ip = (uint64_t) (q / 10000000000ull);
  uint64_t fp = (uint64_t) (q % 10000000000ull);
  if (neg) *p++ = '-';
  char tmp[24];
  int n = 0;
  uint64_t x = ip;
  do { tmp[n++] = (char) ('0' + x % 10); x /= 10; } while (x);
  while (n) *p++ = tmp[--n];
  *p++ = '.';
  for (int k = 9; k >= 0; k--) { p[k] = (char) ('0' + fp % 10); fp /= 10; }
  p += 10;
  return (int) (p - dst);
}

static const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                  1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// Parses the token [s, e) completely as a double (the reference only accepts tokens that strtod consumes entirely,
// shared/gen_func.cpp:404-407).  Returns false when the token is not a number.
static inline bool parse_number(const char *s, const char *e, double *out) {
  const char *p = s;
  bool neg = false;
  if (p < e && (*p == '-' || *p == '+')) { neg = *p == '-'; p++; }
  uint64_t m = 0;
  int digits = 0, frac_digits = 0;
  bool any = false, dot = false, simple = true;
  for (; p < e; p++) {
    const char ch = *p;
    if (ch >= '0' && ch <= '9') {
      any = true;
      if (digits < 19) {
        m = m * 10 + (uint64_t) (ch - '0');
        if (m) digits++;
        if (dot) frac_digits++;
      } else {
        simple = false;                      // too many significant digits for the exact fast path
      }
    } else if (ch == '.' && !dot) {
      dot = true;
    } else {
      simple = false;                        // exponent, inf, nan, hex, garbage: let strtod decide
      break;
    }
  }
  if (simple && any && m < (1ull << 53) && frac_digits <= 22) {
    const double v = (double) m / kPow10[frac_digits];
    *out = neg ? -v : v;
    return true;
  }
  // general case: strtod on a NUL-terminated copy (tokens are short)
  char buf[128];
  const size_t len = (size_t) (e - s);
  if (len == 0) return false;
  char *endp = nullptr;
  if (len < sizeof(buf)) {
    memcpy(buf, s, len);
    buf[len] = 0;
    const double v = strtod(buf, &endp);
    if (endp != buf + len || endp == buf) return false;
    *out = v;
    return true;
  }
  char *big = (char *) malloc(len + 1);
  memcpy(big, s, len);
  big[len] = 0;
  const double v = strtod(big, &endp);
  const bool ok = endp == big + len && endp != big;
  free(big);
  if (ok) *out = v;
  return ok;
}

}  // namespace fastio
