// emu_text.cpp -- TEST INFRASTRUCTURE: the per-line parse rules and the per-id formatting rules of
// stellar-random-walk_b200/csrc/text_io.cuh (the functions the CUDA kernels in text_io.cu call), compiled for the
// host and driven the way the kernels drive them: line starts by the IsLineStart predicate, one srw_parse_line per
// line; one srw_dec_len / srw_dec_write per id.  The tests compare the results with the host parser of libsrw
// (srw_edges_parse_buffer) and with the oracle's formatter.
#include <stdint.h>
#include <string.h>

#include "../../stellar-random-walk_b200/csrc/text_io.cuh"

// returns the number of lines (even when cap is too small); flag[l] = SRW_LINE_*
extern "C" int64_t emu_parse_buffer(const char *buf, int64_t len, int weighted, int partitioned, int32_t *src, int32_t *dst,
                                    float *w, int32_t *pid, uint8_t *flag, int64_t cap) {
  int64_t n = 0;
  for (int64_t i = 0; i < len; ++i) {
    const bool start = i == 0 || buf[i - 1] == '\n' || (buf[i - 1] == '\r' && buf[i] != '\n');   // text_io.cu IsLineStart
    if (!start) continue;
    int64_t e = i;
    while (e < len && !srw_line_end(buf[e])) e++;
    if (n < cap) {
      int32_t s = 0, d = 0, p = 0;
      float wt = 1.0f;
      flag[n] = (uint8_t)srw_parse_line(buf, i, e, weighted, partitioned, &s, &d, &p, &wt);
      src[n] = s; dst[n] = d; w[n] = wt; pid[n] = p;
    }
    n++;
  }
  return n;
}

extern "C" int emu_float_fast(const char *s, int64_t n, float *out) { return srw_float_fast(s, n, out); }

// RW:234-241 with the kernel's arithmetic: returns bytes needed; writes when cap suffices
extern "C" int64_t emu_format_paths(const int32_t *paths, const int32_t *lens, int64_t n, int32_t stride, char *out, int64_t cap) {
  int64_t need = 0;
  for (int64_t i = 0; i < n; ++i) {
    int64_t b = 0;
    for (int j = 0; j < lens[i]; ++j) b += srw_dec_len(paths[i * stride + j]) + 1;
    need += lens[i] > 0 ? b : 1;
  }
  if (!out || cap < need) return need;
  char *d = out;
  for (int64_t i = 0; i < n; ++i) {
    if (lens[i] <= 0) { *d++ = '\n'; continue; }
    for (int j = 0; j < lens[i]; ++j) {
      d += srw_dec_write(paths[i * stride + j], d);
      *d++ = (j == lens[i] - 1) ? '\n' : '\t';
    }
  }
  return need;
}
