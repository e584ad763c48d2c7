// text_io.cuh -- per-line / per-id text rules of the reference's input and output formats, written once for
// the device kernels (text_io.cu) and for the host (the serial parser in srw_host.cpp uses the same tokeniser
// rules; tests/emu compiles this header with g++ and checks it against the oracle without a GPU).
//
//   input  (A1): UniformRandomWalk.loadGraph URW:23-34 / VCutRandomWalk.loadGraph VRW:19-34 -- a line is
//                `triplet.split("\\s+")`, ids are `toInt` (java.lang.Integer.parseInt), the optional weight is
//                `toFloat` (java.lang.Float.parseFloat) with 1.0f on failure, the optional partition id `toInt`.
//   output (A11): RandomWalk.save RW:234-241 -- `path.mkString("\t")`, one line per path.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SRW_TXT_HD __host__ __device__ __forceinline__
#else
#define SRW_TXT_HD inline
#endif

// ---- output: decimal text of an int32 (java.lang.Integer.toString) ----
SRW_TXT_HD int srw_dec_len(int32_t v) {
  uint32_t u = v < 0 ? 0u - (uint32_t)v : (uint32_t)v;
  int n = v < 0 ? 2 : 1;
  if (u >= 1000000000u) return n + 9;
  if (u >= 100000000u) return n + 8;
  if (u >= 10000000u) return n + 7;
  if (u >= 1000000u) return n + 6;
  if (u >= 100000u) return n + 5;
  if (u >= 10000u) return n + 4;
  if (u >= 1000u) return n + 3;
  if (u >= 100u) return n + 2;
  if (u >= 10u) return n + 1;
  return n;
}
// writes exactly srw_dec_len(v) characters at out
SRW_TXT_HD int srw_dec_write(int32_t v, char *out) {
  const int n = srw_dec_len(v);
  uint32_t u = v < 0 ? 0u - (uint32_t)v : (uint32_t)v;
  if (v < 0) out[0] = '-';
  for (int i = n - 1; i >= (v < 0 ? 1 : 0); --i) { out[i] = (char)('0' + (u % 10u)); u /= 10u; }
  return n;
}

// ---- input ----
// java.util.regex \s = [ \t\n\x0B\f\r]
SRW_TXT_HD bool srw_java_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\x0B' || c == '\f' || c == '\r'; }
SRW_TXT_HD bool srw_line_end(char c) { return c == '\n' || c == '\r'; }   // Hadoop LineRecordReader: \n, \r or \r\n

// java.lang.Integer.parseInt on [s, s+n): optional sign, ASCII digits, no overflow.
SRW_TXT_HD bool srw_java_int(const char *s, int64_t n, int32_t *out) {
  if (n <= 0) return false;
  int64_t i = 0;
  bool neg = false;
  if (s[0] == '-' || s[0] == '+') { neg = s[0] == '-'; i = 1; }
  if (i == n) return false;
  int64_t v = 0;
  for (; i < n; ++i) {
    if (s[i] < '0' || s[i] > '9') return false;
    v = v * 10 + (s[i] - '0');
    if (v > 2147483648LL) return false;
  }
  if (neg) v = -v;
  if (v > 2147483647LL || v < -2147483648LL) return false;
  *out = (int32_t)v;
  return true;
}

// Float.parseFloat, the part that can be done exactly in float arithmetic: [sign] digits [. digits] [e|E [sign] digits]
// whose decimal significand (trailing zeros stripped) is below 2^24 with a decimal exponent within +-10 -- then
// value = (float)m * or / 10^|k| is ONE correctly rounded operation on two exactly representable operands, i.e.
// the correctly rounded result parseFloat returns.  Returns 1 = parsed; -1 = outside this fast path (hex, NaN,
// Infinity, f/d suffix, long significand, big exponent, or not a float at all): the caller hands the token to the
// host parser, which applies Java's full grammar and the reference's 1.0f fallback (URW:31).
SRW_TXT_HD int srw_float_fast(const char *s, int64_t n, float *out) {
  if (n <= 0) return -1;
  int64_t i = 0;
  bool neg = false;
  if (s[0] == '-' || s[0] == '+') { neg = s[0] == '-'; i = 1; }
  uint64_t m = 0;
  int nd = 0, frac = 0;
  bool seen_dot = false, any_digit = false;
  for (; i < n; ++i) {
    const char c = s[i];
    if (c >= '0' && c <= '9') {
      any_digit = true;
      if (seen_dot) frac++;
      if (m == 0 && c == '0') continue;                                     // leading zeros carry no significance
      if (nd == 18) return -1;
      m = m * 10u + (uint64_t)(c - '0');
      nd++;
    } else if (c == '.') {
      if (seen_dot) return -1;
      seen_dot = true;
    } else break;
  }
  if (!any_digit) return -1;
  int e10 = 0;
  if (i < n) {
    if (s[i] != 'e' && s[i] != 'E') return -1;                              // suffix, hex, NaN, Infinity, junk: host decides
    ++i;
    bool eneg = false;
    if (i < n && (s[i] == '-' || s[i] == '+')) { eneg = s[i] == '-'; ++i; }
    if (i == n) return -1;
    int ev = 0;
    for (; i < n; ++i) {
      if (s[i] < '0' || s[i] > '9') return -1;
      if (ev < 100000) ev = ev * 10 + (s[i] - '0');
    }
    e10 = eneg ? -ev : ev;
  }
  int k = e10 - frac;
  while (m != 0 && m % 10u == 0) { m /= 10u; k++; }
  float v;
  if (m == 0) v = 0.0f;
  else {
    if (m >= 16777216u || k > 10 || k < -10) return -1;
    const float p10[11] = {1.f, 10.f, 100.f, 1000.f, 10000.f, 100000.f, 1000000.f, 10000000.f, 100000000.f, 1000000000.f, 10000000000.f};
    const int ka = k < 0 ? -k : k;
#ifdef __CUDA_ARCH__
    v = k >= 0 ? __fmul_rn((float)m, p10[ka]) : __fdiv_rn((float)m, p10[ka]);
#else
    v = k >= 0 ? (float)m * p10[ka] : (float)m / p10[ka];
#endif
  }
  *out = neg ? -v : v;
  return 1;
}

enum : int { SRW_LINE_OK = 0, SRW_LINE_HOST_FLOAT = 1, SRW_LINE_ERROR = 2 };

// One line [b, e) of the edge list (terminator excluded).  Mirrors srw_edges_parse_buffer in srw_host.cpp.
SRW_TXT_HD int srw_parse_line(const char *buf, int64_t b, int64_t e, int weighted, int partitioned, int32_t *src, int32_t *dst,
                              int32_t *pid, float *w) {
  // triplet.split("\\s+") (URW:26): leading whitespace -> empty first token; empty line -> [""]
  int64_t tb[4] = {0, 0, 0, 0}, tn[4] = {0, 0, 0, 0};      // tokens 0, 1, 2 and the LAST token
  int ntok = 0;
  int64_t i = b;
  if (i == e || srw_java_ws(buf[i])) { tb[0] = i; tn[0] = 0; ntok = 1; }
  while (i < e) {
    while (i < e && srw_java_ws(buf[i])) i++;
    if (i == e) break;
    const int64_t s0 = i;
    while (i < e && !srw_java_ws(buf[i])) i++;
    if (ntok < 3) { tb[ntok] = s0; tn[ntok] = i - s0; }
    tb[3] = s0; tn[3] = i - s0;
    ntok++;
  }
  if (ntok <= 3 && ntok >= 1) { tb[3] = tb[ntok - 1]; tn[3] = tn[ntok - 1]; }
  *pid = 0;
  *w = 1.0f;
  if (!srw_java_int(buf + tb[0], tn[0], src)) return SRW_LINE_ERROR;        // parts(0).toInt
  if (ntok < 2) return SRW_LINE_ERROR;                                      // parts(1): ArrayIndexOutOfBounds
  if (!srw_java_int(buf + tb[1], tn[1], dst)) return SRW_LINE_ERROR;
  bool want_w;
  if (!partitioned) {
    want_w = weighted && ntok > 2;                                          // URW:29-32
  } else {
    if (ntok > 2 && !srw_java_int(buf + tb[2], tn[2], pid)) *pid = 0;       // VRW:23-26 (random in the reference)
    want_w = weighted && ntok > 3;                                          // VRW:29-32
  }
  if (want_w) {
    const int r = srw_float_fast(buf + tb[3], tn[3], w);
    if (r < 0) { *w = 1.0f; return SRW_LINE_HOST_FLOAT; }
  }
  return SRW_LINE_OK;
}
