// text_io.cu -- the text on either side of the walk, on the device:
//
//   A1  edge-list text -> (src, dst, weight, partition id) arrays in HBM.  Replaces the per-line closure of
//       UniformRandomWalk.loadGraph (URW:23-34) / VCutRandomWalk.loadGraph (VRW:19-34): the file is pushed
//       through the device in line-aligned chunks; line starts are compacted with one select pass, then ONE
//       THREAD PARSES ONE LINE with the JVM rules of text_io.cuh.  Weight tokens outside the exact float fast
//       path (hex floats, NaN/Infinity, f/d suffixes, >7-digit significands) are re-parsed on the host.
//   A11 paths -> `<output>/path/part-NNNNN` (RandomWalk.save RW:234-241): ONE WARP FORMATS ONE PATH -- lane j
//       sizes id j, a warp scan places it, digits are staged in shared memory and leave as 16-byte stores.
//       srw_walk_save streams walk -> format -> D2H -> write() so that neither the paths nor their text ever
//       has to fit in host memory (RMAT-26: 107 GB of ids, ~240 GB of text per 10 rounds).
//
// Both are HBM-streaming integer/byte work (no tensor cores): bounds and measurements in DESIGN.md section 4.
#include <cuda_runtime.h>
#include <dirent.h>
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cub/cub.cuh>
#include <string>
#include <vector>

#include "srw_internal.h"
#include "text_io.cuh"

namespace {

struct DBuf {
  void *p = nullptr;
  size_t bytes = 0;
  ~DBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) {
    if (p) { cudaFree(p); p = nullptr; }
    bytes = n;
    return cudaMalloc(&p, n ? n : 1);
  }
  template <class T> T *as() const { return reinterpret_cast<T *>(p); }
  void *release() { void *q = p; p = nullptr; bytes = 0; return q; }
};
struct PinBuf {
  void *p = nullptr;
  ~PinBuf() { if (p) cudaFreeHost(p); }
  cudaError_t alloc(size_t n) { return cudaHostAlloc(&p, n ? n : 1, cudaHostAllocDefault); }
  template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// ------------------------------------------------------------------------------------------
// A11: formatter
// ------------------------------------------------------------------------------------------
constexpr int kFmtThreads = 128;   // 4 warps, one path each per pass

// bytes of line i: sum(len(id)) + one separator per id (tabs, then the newline)
__global__ void k_line_bytes(int64_t n, int32_t stride, const int32_t *__restrict__ paths, const int32_t *__restrict__ lens,
                             int64_t *__restrict__ line_bytes) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += n_warps) {
    const int32_t len = lens[i];
    const int32_t *row = paths + i * stride;
    int b = 0;
    for (int j = lane; j < len; j += 32) b += srw_dec_len(row[j]) + 1;
#pragma unroll
    for (int o = 16; o; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
    if (lane == 0) line_bytes[i] = len > 0 ? b : 1;    // an empty path would still be one (empty) line
  }
}

// line_off: exclusive prefix sum of line_bytes (relative to `out`).  Dynamic shared memory: warps * (line_cap + 32).
__global__ void k_format_lines(int64_t n, int32_t stride, const int32_t *__restrict__ paths, const int32_t *__restrict__ lens,
                               const int64_t *__restrict__ line_off, char *__restrict__ out, int line_cap) {
  extern __shared__ __align__(16) char sm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  char *mine = sm + (size_t)wib * (size_t)(line_cap + 32);
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += n_warps) {
    const int32_t len = lens[i];
    const int32_t *row = paths + i * stride;
    char *g = out + line_off[i];
    const int mis = (int)(reinterpret_cast<uintptr_t>(g) & 15);    // stage at the same phase as the destination
    char *s = mine + mis;
    int pos = 0;
    for (int j0 = 0; j0 < len; j0 += 32) {
      const int j = j0 + lane;
      const int32_t v = j < len ? row[j] : 0;
      const int l = j < len ? srw_dec_len(v) + 1 : 0;
      int incl = l;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (j < len) {
        char *d = s + pos + incl - l;
        srw_dec_write(v, d);
        d[l - 1] = (j == len - 1) ? '\n' : '\t';
      }
      pos += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (len <= 0) { if (lane == 0) s[0] = '\n'; pos = 1; }
    __syncwarp();
    // head bytes up to the first 16-byte boundary, whole int4 chunks, tail bytes
    const int head = mis ? min(16 - mis, pos) : 0;
    if (lane < head) g[lane] = s[lane];
    const int body = (pos - head) >> 4;
    const int4 *s4 = reinterpret_cast<const int4 *>(s + head);
    int4 *g4 = reinterpret_cast<int4 *>(g + head);
    for (int c = lane; c < body; c += 32) g4[c] = s4[c];
    const int done = head + (body << 4);
    if (done + lane < pos) g[done + lane] = s[done + lane];
    __syncwarp();
  }
}

struct Formatter {
  DBuf off;            // [cap_paths + 1] int64
  DBuf scan_tmp;
  int64_t cap_paths = 0;
  int warps = 4, line_cap = 0;
  size_t smem = 0;

  srw_status init(int64_t max_paths, int32_t stride) {
    cap_paths = max_paths;
    SRW_CUDA(off.alloc((size_t)(max_paths + 1) * 8));
    size_t tb = 0;
    SRW_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, (int64_t *)nullptr, (int64_t *)nullptr, max_paths + 1));
    SRW_CUDA(scan_tmp.alloc(tb));
    line_cap = (int)(((int64_t)stride * 12 + 15) & ~15LL);
    warps = 4;
    while (warps > 1 && (size_t)warps * (size_t)(line_cap + 32) > 48 * 1024) warps >>= 1;
    smem = (size_t)warps * (size_t)(line_cap + 32);
    if (smem > 48 * 1024) {
      if (smem > 227 * 1024) { srw_set_error("walkLength %d: a line of %d bytes does not fit the formatter's staging buffer", stride - 2, line_cap); return SRW_ERR_UNSUPPORTED; }
      SRW_CUDA(cudaFuncSetAttribute(k_format_lines, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    return SRW_OK;
  }
  // line offsets of paths [0, n) into off[0..n]; *total = bytes of their text
  srw_status measure(const int32_t *d_paths, const int32_t *d_lens, int64_t n, int32_t stride, int64_t *total, cudaStream_t st) {
    if (n > cap_paths) { srw_set_error("formatter: %lld paths > capacity %lld", (long long)n, (long long)cap_paths); return SRW_ERR_ARG; }
    int64_t *o = off.as<int64_t>();
    SRW_CUDA(cudaMemsetAsync(o + n, 0, 8, st));
    if (n > 0) {
      const int64_t blocks = std::min<int64_t>((n * 32 + kFmtThreads - 1) / kFmtThreads, 148 * 16);
      k_line_bytes<<<(unsigned)blocks, kFmtThreads, 0, st>>>(n, stride, d_paths, d_lens, o);
    }
    size_t tb = scan_tmp.bytes;
    SRW_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp.p, tb, o, o, n + 1, st));
    SRW_CUDA(cudaMemcpyAsync(total, o + n, 8, cudaMemcpyDeviceToHost, st));
    SRW_CUDA(cudaStreamSynchronize(st));
    return SRW_OK;
  }
  srw_status emit(const int32_t *d_paths, const int32_t *d_lens, int64_t n, int32_t stride, char *d_text, cudaStream_t st) {
    if (n <= 0) return SRW_OK;
    const int threads = warps * 32;
    const int64_t blocks = std::min<int64_t>((n + warps - 1) / warps, 148 * 16);
    k_format_lines<<<(unsigned)blocks, threads, smem, st>>>(n, stride, d_paths, d_lens, off.as<int64_t>(), d_text, line_cap);
    SRW_CUDA(cudaGetLastError());
    return SRW_OK;
  }
};

int mkdir_p(const std::string &dir) {
  std::string cur;
  for (size_t i = 0; i <= dir.size(); ++i) {
    if (i == dir.size() || dir[i] == '/') {
      if (!cur.empty() && mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST) return -1;
    }
    if (i < dir.size()) cur.push_back(dir[i]);
  }
  return 0;
}

bool write_all(int fd, const char *p, size_t n) {
  while (n) {
    const ssize_t w = write(fd, p, n);
    if (w < 0) { if (errno == EINTR) continue; return false; }
    p += w; n -= (size_t)w;
  }
  return true;
}

}  // namespace

extern "C" srw_status srw_paths_format_device(const int32_t *d_paths, const int32_t *d_lens, int64_t n_paths, int32_t stride,
                                              char *d_text, int64_t cap, int64_t *needed, void *stream) {
  SRW_TRY(srw_require_device());
  if (n_paths < 0 || stride < 1 || (n_paths > 0 && (!d_paths || !d_lens))) { srw_set_error("srw_paths_format_device: bad argument"); return SRW_ERR_ARG; }
  Formatter f;
  SRW_TRY(f.init(n_paths, stride));
  int64_t total = 0;
  cudaStream_t st = (cudaStream_t)stream;
  SRW_TRY(f.measure(d_paths, d_lens, n_paths, stride, &total, st));
  if (needed) *needed = total;
  if (!d_text) return SRW_OK;
  if (cap < total) { srw_set_error("srw_paths_format_device: buffer of %lld bytes < %lld needed", (long long)cap, (long long)total); return SRW_ERR_ARG; }
  SRW_TRY(f.emit(d_paths, d_lens, n_paths, stride, d_text, st));
  SRW_CUDA(cudaStreamSynchronize(st));
  return SRW_OK;
}

namespace {
__global__ void k_count_short(int64_t n, int32_t stride, const int32_t *__restrict__ lens, unsigned long long *out) {
  unsigned long long c = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) c += lens[i] < stride ? 1u : 0u;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
thread_local int64_t t_short_paths = 0;
}  // namespace
int64_t srw_last_short_paths() { return t_short_paths; }

// RW:75-176 + RW:234-241 streamed: numWalks rounds in walker order, formatted on the device, written as
// `parts` contiguous blocks of lines (Spark's repartition spreads lines arbitrarily, RW:240).
extern "C" srw_status srw_walk_save(const srw_graph *g, const srw_params *params) {
  SRW_TRY(srw_require_device());
  if (!g || !params || !params->output[0]) { srw_set_error("srw_walk_save: no graph / output path"); return SRW_ERR_ARG; }
  if (params->num_walks < 0) { srw_set_error("numWalks must be >= 0"); return SRW_ERR_ARG; }
  const bool multi = !g->shards.empty();
  if (params->num_gpus > 1 && !multi) { srw_set_error("--gpus %d: the graph was loaded on one GPU (load it with the same --gpus)", params->num_gpus); return SRW_ERR_ARG; }
  const std::string dir = std::string(params->output) + "/path";          // Property.pathSuffix
  struct stat stt;
  if (stat(dir.c_str(), &stt) == 0) {   // Hadoop saveAsTextFile refuses an existing directory
    srw_set_error("FileAlreadyExistsException: Output directory %s already exists", dir.c_str());
    return SRW_ERR_IO;
  }
  if (mkdir_p(dir) != 0) { srw_set_error("cannot create %s: %s", dir.c_str(), strerror(errno)); return SRW_ERR_IO; }
  SRW_CUDA(cudaSetDevice(g->device));
  const int32_t stride = params->walk_length + 2;
  const int64_t total = (int64_t)params->num_walks * g->nv;
  int parts = params->single_output ? 1 : params->rdd_partitions;          // Main:64-69
  if (parts < 1) parts = 1;

  // walk batch: up to 2^24 walkers (enough to fill the machine), bounded by a third of the free memory
  size_t free_b = 0, total_b = 0;
  SRW_CUDA(cudaMemGetInfo(&free_b, &total_b));
  int64_t batch = std::min<int64_t>({total, (int64_t)1 << 24, (int64_t)(free_b / 3) / ((int64_t)stride * 4 + 4)});
  if (batch < 1) batch = 1;
  if (multi && g->nv > 0)   // the sharded walk runs whole rounds (device 0 also holds its shard and ~420 bytes of exchange block per walker)
    batch = g->nv * std::max<int64_t>(1, std::min<int64_t>({(int64_t)params->num_walks, ((int64_t)1 << 25) / g->nv, (int64_t)(free_b / 3) / (g->nv * ((int64_t)stride * 4 + 420))}));
  // text chunk: worst case 12 bytes per id; two device and two pinned host buffers
  const int64_t chunk_bytes = getenv("SRW_SAVE_CHUNK_BYTES") ? atoll(getenv("SRW_SAVE_CHUNK_BYTES")) : ((int64_t)256 << 20);
  int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(batch, chunk_bytes / ((int64_t)stride * 12)));
  const int64_t text_cap = chunk * stride * 12;
  DBuf d_paths, d_lens, d_text[2], d_short;
  t_short_paths = 0;
  SRW_CUDA(d_short.alloc(8));
  PinBuf h_text[2], h_off[2];
  SRW_CUDA(d_paths.alloc((size_t)batch * stride * 4));
  SRW_CUDA(d_lens.alloc((size_t)batch * 4));
  for (int b = 0; b < 2; ++b) {
    SRW_CUDA(d_text[b].alloc((size_t)text_cap));
    SRW_CUDA(h_text[b].alloc((size_t)text_cap));
    SRW_CUDA(h_off[b].alloc((size_t)(chunk + 1) * 8));
  }
  Formatter fmt;
  SRW_TRY(fmt.init(chunk, stride));
  // stream, events and the open part file are released on EVERY exit (the SRW_CUDA early returns below included)
  struct SaveGuard {
    cudaStream_t st = nullptr;
    cudaEvent_t ready[2] = {nullptr, nullptr};
    int fd = -1;
    ~SaveGuard() {
      for (int b = 0; b < 2; ++b) if (ready[b]) cudaEventDestroy(ready[b]);
      if (st) cudaStreamDestroy(st);
      if (fd >= 0) close(fd);
    }
  } guard;
  SRW_CUDA(cudaStreamCreate(&guard.st));
  for (int b = 0; b < 2; ++b) SRW_CUDA(cudaEventCreateWithFlags(&guard.ready[b], cudaEventDisableTiming));
  cudaStream_t st = guard.st;
  cudaEvent_t *ready = guard.ready;

  // output files: part k holds paths [total*k/parts, total*(k+1)/parts)
  int file_k = -1;
  int &fd = guard.fd;
  bool io_ok = true;
  std::string io_err;
  auto open_part = [&](int k) {
    if (fd >= 0) { if (close(fd) != 0) io_ok = false; fd = -1; }
    char name[32];
    snprintf(name, sizeof(name), "/part-%05d", k);
    fd = open((dir + name).c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (fd < 0) { io_ok = false; io_err = "cannot write " + dir + name + ": " + strerror(errno); }
    file_k = k;
  };
  auto part_end = [&](int k) { return total * (int64_t)(k + 1) / parts; };
  // writes the text of paths [p0, p0 + n) (host copy: text + n+1 line offsets)
  auto write_chunk = [&](int64_t p0, int64_t n, const char *text, const int64_t *off) {
    int64_t done = 0;
    while (done < n && io_ok) {
      if (file_k < 0) open_part(0);
      while (p0 + done >= part_end(file_k) && file_k + 1 < parts) open_part(file_k + 1);
      const int64_t upto = std::min(n, part_end(file_k) - p0);
      const int64_t take = file_k + 1 < parts ? upto : n;
      if (take > done) {
        if (fd >= 0 && !write_all(fd, text + off[done], (size_t)(off[take] - off[done]))) { io_ok = false; io_err = std::string("short write: ") + strerror(errno); }
        done = take;
      }
    }
  };

  double kernel_ms = 0;
  int64_t launches = 0, steps = 0, props = 0, mem = 0, logs = 0, text_bytes = 0;
  struct Pending { int64_t p0 = 0, n = 0; bool live = false; } pend[2];
  int slot = 0;
  srw_status rc = SRW_OK;
  for (int64_t first = 0; first < total && rc == SRW_OK && io_ok; first += batch) {
    const int64_t nb = std::min(batch, total - first);
    rc = srw_walk_device(g, params, (uint64_t)first, nb, d_paths.as<int32_t>(), d_lens.as<int32_t>(), st);
    if (rc != SRW_OK) break;
    if (multi) SRW_CUDA(cudaSetDevice(g->device));
    {
      // walkers that stopped at a vertex without out-neighbours (RW:115-119; the reference's `Zero Neighbors` accumulator)
      SRW_CUDA(cudaMemsetAsync(d_short.p, 0, 8, st));
      k_count_short<<<148 * 4, 256, 0, st>>>(nb, stride, d_lens.as<int32_t>(), (unsigned long long *)d_short.p);
      unsigned long long h = 0;
      SRW_CUDA(cudaMemcpyAsync(&h, d_short.p, 8, cudaMemcpyDeviceToHost, st));
      SRW_CUDA(cudaStreamSynchronize(st));
      t_short_paths += (int64_t)h;
    }
    srw_walk_info wi;
    srw_last_walk_info(&wi);
    kernel_ms += wi.kernel_ms; launches += wi.kernel_launches; steps += wi.steps;
    props += wi.proposals; mem += wi.member_tests; logs += wi.probes_log2;
    for (int64_t c0 = 0; c0 < nb && rc == SRW_OK && io_ok; c0 += chunk) {
      const int64_t n = std::min(chunk, nb - c0);
      const int32_t *pp = d_paths.as<int32_t>() + c0 * stride, *ll = d_lens.as<int32_t>() + c0;
      int64_t bytes = 0;
      rc = fmt.measure(pp, ll, n, stride, &bytes, st);
      if (rc != SRW_OK) break;
      rc = fmt.emit(pp, ll, n, stride, d_text[slot].as<char>(), st);
      if (rc != SRW_OK) break;
      if (cudaMemcpyAsync(h_text[slot].p, d_text[slot].p, (size_t)bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaMemcpyAsync(h_off[slot].p, fmt.off.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
          cudaEventRecord(ready[slot], st) != cudaSuccess) {
        srw_set_error("srw_walk_save: device-to-host copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        rc = SRW_ERR_CUDA;
        break;
      }
      pend[slot] = {first + c0, n, true};
      text_bytes += bytes;
      launches += 3;
      // while that chunk is formatted and copied, write the previous one
      const int other = slot ^ 1;
      if (pend[other].live) {
        SRW_CUDA(cudaEventSynchronize(ready[other]));
        write_chunk(pend[other].p0, pend[other].n, h_text[other].as<char>(), h_off[other].as<int64_t>());
        pend[other].live = false;
      }
      // the formatter's offset array is reused by the next measure(): this chunk's copy must have left the device
      SRW_CUDA(cudaEventSynchronize(ready[slot]));
      slot = other;
    }
  }
  for (int b = 0; b < 2 && rc == SRW_OK && io_ok; ++b) {
    const int s2 = slot ^ b ^ 1;     // oldest first
    if (pend[s2].live) {
      SRW_CUDA(cudaEventSynchronize(ready[s2]));
      write_chunk(pend[s2].p0, pend[s2].n, h_text[s2].as<char>(), h_off[s2].as<int64_t>());
      pend[s2].live = false;
    }
  }
  if (rc == SRW_OK && io_ok) {
    if (file_k < 0) open_part(0);
    while (file_k + 1 < parts && io_ok) open_part(file_k + 1);    // trailing (possibly empty) part files, as Spark writes them
  }
  if (fd >= 0) { if (close(fd) != 0) io_ok = false; fd = -1; }
  if (rc != SRW_OK) return rc;
  if (!io_ok) { srw_set_error("%s", io_err.empty() ? "I/O error while writing the path files" : io_err.c_str()); return SRW_ERR_IO; }
  FILE *f = fopen((dir + "/_SUCCESS").c_str(), "wb");
  if (f) fclose(f);
  srw_set_walk_info(kernel_ms, launches, steps, props, mem, logs);
  (void)text_bytes;
  return SRW_OK;
}

// ------------------------------------------------------------------------------------------
// A1: edge-list parser
// ------------------------------------------------------------------------------------------
namespace {

struct IsLineStart {
  const char *buf;
  __host__ __device__ unsigned long long operator()(uint32_t i) const {   // 0 / 1: a select flag and a countable value
    if (i == 0) return true;
    const char p = buf[i - 1];
    return p == '\n' || (p == '\r' && buf[i] != '\n');
  }
};

struct ParseStatus {
  unsigned long long first_error;   // line index inside the chunk, ~0 = none
  unsigned long long n_host_float;
};

__global__ void k_parse_lines(const char *__restrict__ buf, uint32_t len, const uint32_t *__restrict__ starts, int64_t n_lines,
                              int weighted, int partitioned, int32_t *__restrict__ src, int32_t *__restrict__ dst,
                              float *__restrict__ w, int32_t *__restrict__ pid, uint8_t *__restrict__ flag, ParseStatus *status) {
  for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < n_lines; l += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = starts[l];
    int64_t e = b;
    while (e < (int64_t)len && !srw_line_end(buf[e])) e++;
    int32_t s = 0, d = 0, p = 0;
    float wt = 1.0f;
    const int r = srw_parse_line(buf, b, e, weighted, partitioned, &s, &d, &p, &wt);
    src[l] = s; dst[l] = d;
    if (w) w[l] = wt;
    if (pid) pid[l] = p;
    if (flag) flag[l] = (uint8_t)r;
    if (r == SRW_LINE_ERROR) atomicMin(&status->first_error, (unsigned long long)l);
    else if (r == SRW_LINE_HOST_FLOAT) atomicAdd(&status->n_host_float, 1ULL);
  }
}

__global__ void k_scatter_f32(int64_t n, const int64_t *__restrict__ idx, const float *__restrict__ val, float *out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[idx[i]] = val[i];
}

// end of the chunk that starts at c0: the last line terminator inside the window (a "\r\n" is never split)
size_t chunk_end(const char *t, size_t len, size_t c0, size_t want) {
  size_t end = std::min(len, c0 + want);
  if (end == len) return len;
  size_t e = end;
  while (e > c0 && !srw_line_end(t[e - 1])) e--;
  if (e == c0) {                               // one line longer than the window: run forward to its end
    e = end;
    while (e < len && !srw_line_end(t[e])) e++;
    if (e < len) e++;
  }
  if (e < len && t[e - 1] == '\r' && t[e] == '\n') e++;
  return e;
}

struct ChunkOut {
  int64_t n = 0;
  int32_t *src = nullptr, *dst = nullptr, *pid = nullptr;
  float *w = nullptr;
};

}  // namespace

// Host text -> device edge arrays (caller frees with cudaFree).  *d_w is NULL unless `weighted`, *d_pid unless `partitioned`.
srw_status srw_parse_text_device(const char *h_text, size_t len, int weighted, int partitioned, int64_t *n_out, int32_t **d_src,
                                 int32_t **d_dst, float **d_w, int32_t **d_pid) {
  SRW_TRY(srw_require_device());
  const size_t chunk_want = getenv("SRW_PARSE_CHUNK_BYTES") ? (size_t)atoll(getenv("SRW_PARSE_CHUNK_BYTES")) : ((size_t)1 << 30);
  const size_t want = std::min<size_t>(std::max<size_t>(chunk_want, 16), (size_t)1 << 30);   // chunk positions are u32, CUB counts are int
  std::vector<ChunkOut> outs;
  auto free_outs = [&]() { for (auto &c : outs) { cudaFree(c.src); cudaFree(c.dst); cudaFree(c.w); cudaFree(c.pid); } outs.clear(); };
  DBuf text, starts, nsel, sel_tmp, flag, status;
  const bool timing = getenv("SRW_IO_TIMING") != nullptr;     // stderr: where the parse time goes (profiles/run_io.py)
  double t_h2d = 0, t_split = 0, t_parse = 0;
  auto now = []() { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  SRW_CUDA(nsel.alloc(8));
  SRW_CUDA(status.alloc(sizeof(ParseStatus)));
  int64_t lines_before = 0;
  size_t c0 = 0;
  while (c0 < len) {
    const size_t c1 = chunk_end(h_text, len, c0, want);
    const size_t clen = c1 - c0;
    if (clen >= ((size_t)1 << 32) - 1) { free_outs(); srw_set_error("a single line of %zu bytes cannot be parsed", clen); return SRW_ERR_PARSE; }
    if (text.bytes < clen + 1) SRW_CUDA(text.alloc(clen + 1));
    auto t0 = now();
    SRW_CUDA(cudaMemcpy(text.p, h_text + c0, clen, cudaMemcpyHostToDevice));
    auto t1 = now();
    // line starts: every position whose predecessor ends a line
    cub::CountingInputIterator<uint32_t> pos(0);
    cub::TransformInputIterator<unsigned long long, IsLineStart, cub::CountingInputIterator<uint32_t>> is_start(pos, IsLineStart{text.as<char>()});
    // empty lines are 1 byte each, so the number of starts can reach clen: count first, then size the array
    size_t tb = 0;
    SRW_CUDA(cub::DeviceReduce::Sum(nullptr, tb, is_start, nsel.as<unsigned long long>(), (int64_t)clen));
    size_t tb2 = 0;
    SRW_CUDA(cub::DeviceSelect::Flagged(nullptr, tb2, pos, is_start, starts.as<uint32_t>(), nsel.as<unsigned long long>(), (int64_t)clen));
    if (sel_tmp.bytes < std::max(tb, tb2)) SRW_CUDA(sel_tmp.alloc(std::max(tb, tb2)));
    tb = sel_tmp.bytes;
    SRW_CUDA(cub::DeviceReduce::Sum(sel_tmp.p, tb, is_start, nsel.as<unsigned long long>(), (int64_t)clen));
    unsigned long long n_lines_u = 0;
    SRW_CUDA(cudaMemcpy(&n_lines_u, nsel.p, 8, cudaMemcpyDeviceToHost));
    const int64_t n_lines = (int64_t)n_lines_u;
    if (starts.bytes < (size_t)(n_lines + 1) * 4) SRW_CUDA(starts.alloc((size_t)(n_lines + 1) * 4));
    tb2 = sel_tmp.bytes;
    SRW_CUDA(cub::DeviceSelect::Flagged(sel_tmp.p, tb2, pos, is_start, starts.as<uint32_t>(), nsel.as<unsigned long long>(), (int64_t)clen));
    if (timing) SRW_CUDA(cudaDeviceSynchronize());
    auto t2 = now();
    ChunkOut co;
    co.n = n_lines;
    outs.push_back(co);
    ChunkOut &o = outs.back();
    SRW_CUDA(cudaMalloc(&o.src, (size_t)std::max<int64_t>(n_lines, 1) * 4));
    SRW_CUDA(cudaMalloc(&o.dst, (size_t)std::max<int64_t>(n_lines, 1) * 4));
    if (weighted) SRW_CUDA(cudaMalloc(&o.w, (size_t)std::max<int64_t>(n_lines, 1) * 4));
    if (partitioned) SRW_CUDA(cudaMalloc(&o.pid, (size_t)std::max<int64_t>(n_lines, 1) * 4));
    if (flag.bytes < (size_t)n_lines + 1) SRW_CUDA(flag.alloc((size_t)n_lines + 1));
    ParseStatus hs{~0ULL, 0ULL};
    SRW_CUDA(cudaMemcpy(status.p, &hs, sizeof(hs), cudaMemcpyHostToDevice));
    if (n_lines > 0) {
      const int64_t blocks = std::min<int64_t>((n_lines + 255) / 256, 148 * 32);
      k_parse_lines<<<(unsigned)blocks, 256>>>(text.as<char>(), (uint32_t)clen, starts.as<uint32_t>(), n_lines, weighted, partitioned,
                                                o.src, o.dst, o.w, o.pid, flag.as<uint8_t>(), status.as<ParseStatus>());
      SRW_CUDA(cudaGetLastError());
    }
    SRW_CUDA(cudaMemcpy(&hs, status.p, sizeof(hs), cudaMemcpyDeviceToHost));
    auto t3 = now();
    t_h2d += secs(t0, t1); t_split += secs(t1, t2); t_parse += secs(t2, t3);
    if (hs.first_error != ~0ULL) {
      // the reference throws from the executor that meets the bad line; here the FIRST bad line is reported, with
      // the message of the host parser (same rules) and its line number in the file
      uint32_t sb = 0;
      SRW_CUDA(cudaMemcpy(&sb, starts.as<uint32_t>() + hs.first_error, 4, cudaMemcpyDeviceToHost));
      size_t e = c0 + sb;
      while (e < len && !srw_line_end(h_text[e])) e++;
      srw_edges *tmp = nullptr;
      std::string msg = "malformed line";
      const bool empty = e == c0 + sb;           // the host parser sees an empty line only through its terminator
      if (srw_edges_parse_buffer(empty ? "\n" : h_text + c0 + sb, empty ? 1 : e - (c0 + sb), weighted, partitioned, &tmp) != SRW_OK) {
        msg = srw_last_error();
        const size_t colon = msg.find(": ");
        if (msg.compare(0, 5, "line ") == 0 && colon != std::string::npos) msg = msg.substr(colon + 2);
      } else srw_edges_free(tmp);
      free_outs();
      srw_set_error("line %lld: %s", (long long)(lines_before + (int64_t)hs.first_error + 1), msg.c_str());
      return SRW_ERR_PARSE;
    }
    if (hs.n_host_float > 0) {
      // weight tokens outside the device fast path: Java's full Float.parseFloat grammar lives on the host
      std::vector<uint8_t> hf((size_t)n_lines);
      std::vector<uint32_t> hst((size_t)n_lines);
      SRW_CUDA(cudaMemcpy(hf.data(), flag.p, (size_t)n_lines, cudaMemcpyDeviceToHost));
      SRW_CUDA(cudaMemcpy(hst.data(), starts.p, (size_t)n_lines * 4, cudaMemcpyDeviceToHost));
      std::vector<int64_t> idx;
      std::vector<float> val;
      for (int64_t l = 0; l < n_lines; ++l) {
        if (hf[(size_t)l] != SRW_LINE_HOST_FLOAT) continue;
        size_t b = c0 + hst[(size_t)l], e = b;
        while (e < len && !srw_line_end(h_text[e])) e++;
        srw_edges *tmp = nullptr;
        if (srw_edges_parse_buffer(h_text + b, e - b, weighted, partitioned, &tmp) != SRW_OK) { free_outs(); return SRW_ERR_PARSE; }
        idx.push_back(l);
        val.push_back(tmp->w.empty() ? 1.0f : tmp->w[0]);
        srw_edges_free(tmp);
      }
      DBuf di, dv;
      SRW_CUDA(di.alloc(idx.size() * 8));
      SRW_CUDA(dv.alloc(val.size() * 4));
      SRW_CUDA(cudaMemcpy(di.p, idx.data(), idx.size() * 8, cudaMemcpyHostToDevice));
      SRW_CUDA(cudaMemcpy(dv.p, val.data(), val.size() * 4, cudaMemcpyHostToDevice));
      k_scatter_f32<<<(unsigned)((idx.size() + 255) / 256), 256>>>((int64_t)idx.size(), di.as<int64_t>(), dv.as<float>(), o.w);
      SRW_CUDA(cudaDeviceSynchronize());
    }
    lines_before += n_lines;
    c0 = c1;
  }
  // concatenate the chunks (a single chunk is handed over as it is)
  int64_t n = 0;
  for (auto &c : outs) n += c.n;
  if (timing)
    fprintf(stderr, "[srw io] parse: %zu bytes, %lld lines, %zu chunk(s): H2D %.3f s, line split %.3f s, alloc+parse kernel %.3f s\n", len,
            (long long)n, outs.size(), t_h2d, t_split, t_parse);
  *n_out = n;
  *d_src = *d_dst = nullptr;
  if (d_w) *d_w = nullptr;
  if (d_pid) *d_pid = nullptr;
  if (outs.size() == 1) {
    *d_src = outs[0].src; *d_dst = outs[0].dst;
    if (d_w) *d_w = outs[0].w; else cudaFree(outs[0].w);
    if (d_pid) *d_pid = outs[0].pid; else cudaFree(outs[0].pid);
    outs.clear();
    return SRW_OK;
  }
  const size_t nb = (size_t)std::max<int64_t>(n, 1) * 4;
  int32_t *s = nullptr, *d = nullptr, *p = nullptr;
  float *w = nullptr;
  cudaError_t ce = cudaMalloc(&s, nb);
  if (ce == cudaSuccess) ce = cudaMalloc(&d, nb);
  if (ce == cudaSuccess && weighted && d_w) ce = cudaMalloc(&w, nb);
  if (ce == cudaSuccess && partitioned && d_pid) ce = cudaMalloc(&p, nb);
  int64_t at = 0;
  for (auto &c : outs) {
    if (ce != cudaSuccess) break;
    if (c.n > 0) {
      ce = cudaMemcpy(s + at, c.src, (size_t)c.n * 4, cudaMemcpyDeviceToDevice);
      if (ce == cudaSuccess) ce = cudaMemcpy(d + at, c.dst, (size_t)c.n * 4, cudaMemcpyDeviceToDevice);
      if (ce == cudaSuccess && w) ce = cudaMemcpy(w + at, c.w, (size_t)c.n * 4, cudaMemcpyDeviceToDevice);
      if (ce == cudaSuccess && p) ce = cudaMemcpy(p + at, c.pid, (size_t)c.n * 4, cudaMemcpyDeviceToDevice);
    }
    at += c.n;
  }
  free_outs();
  if (ce != cudaSuccess) {
    cudaFree(s); cudaFree(d); cudaFree(w); cudaFree(p);
    srw_set_error("edge-list parse: %s", cudaGetErrorString(ce));
    return SRW_ERR_CUDA;
  }
  *d_src = s; *d_dst = d;
  if (d_w) *d_w = w;
  if (d_pid) *d_pid = p;
  return SRW_OK;
}

// The device parser with a host result: the same srw_edges a caller gets from srw_edges_parse_buffer.
extern "C" srw_status srw_edges_parse_buffer_device(const char *buf, size_t len, int weighted, int partitioned, srw_edges **out) {
  if (!out || (len && !buf)) return SRW_ERR_ARG;
  int64_t n = 0;
  int32_t *s = nullptr, *d = nullptr, *p = nullptr;
  float *w = nullptr;
  SRW_TRY(srw_parse_text_device(buf, len, weighted, partitioned, &n, &s, &d, &w, &p));
  srw_edges *E = new srw_edges();
  E->has_pid = partitioned != 0;
  E->src.resize((size_t)n); E->dst.resize((size_t)n); E->w.assign((size_t)n, 1.0f);
  if (partitioned) E->pid.resize((size_t)n);
  cudaError_t ce = cudaSuccess;
  if (n > 0) {
    ce = cudaMemcpy(E->src.data(), s, (size_t)n * 4, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess) ce = cudaMemcpy(E->dst.data(), d, (size_t)n * 4, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && w) ce = cudaMemcpy(E->w.data(), w, (size_t)n * 4, cudaMemcpyDeviceToHost);
    if (ce == cudaSuccess && p) ce = cudaMemcpy(E->pid.data(), p, (size_t)n * 4, cudaMemcpyDeviceToHost);
  }
  cudaFree(s); cudaFree(d); cudaFree(w); cudaFree(p);
  if (ce != cudaSuccess) { delete E; srw_set_error("edge-list parse: %s", cudaGetErrorString(ce)); return SRW_ERR_CUDA; }
  *out = E;
  return SRW_OK;
}

// A1 + A2 from a file: mmap -> device parse -> CSR build; the edge arrays never exist on the host.
// SparkContext.textFile accepts a directory (URW:23): every regular file in it, in name order, files whose name starts
// with '_' or '.' skipped (Hadoop's hidden-file filter: _SUCCESS, .crc).  A file's last line needs no terminator.
// Main:54-57: Params decide which walker runs -- one GPU, or (--gpus N) one vertex-range shard per GPU of this process
static srw_status build_for(const srw_params *params, int64_t n, const int32_t *s, const int32_t *d, const float *w, const int32_t *p,
                            unsigned flags, srw_graph **out) {
  // (--partitioned true: VRW -- the partition-id column is the shard map, owner(v) = getPartition(v) mod num_gpus)
  if (params->num_gpus > 1) return srw_build_graph_device_multi(n, s, d, w, params->directed, params->num_gpus, out, params->partitioned ? p : nullptr, -1.0);
  return srw_build_graph_device(n, s, d, w, p, params->directed, flags, out);
}

static srw_status load_directory(const srw_params *params, unsigned flags, srw_graph **out) {
  std::vector<std::string> names;
  DIR *dp = opendir(params->input);
  if (!dp) { srw_set_error("Input path does not exist: %s", params->input); return SRW_ERR_IO; }
  while (struct dirent *de = readdir(dp)) {
    if (de->d_name[0] == '_' || de->d_name[0] == '.') continue;
    const std::string path = std::string(params->input) + "/" + de->d_name;
    struct stat st;
    if (stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode)) names.push_back(path);
  }
  closedir(dp);
  std::sort(names.begin(), names.end());
  std::string text;
  for (const std::string &path : names) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { srw_set_error("cannot read %s: %s", path.c_str(), strerror(errno)); return SRW_ERR_IO; }
    char chunk[1 << 16];
    size_t r;
    while ((r = fread(chunk, 1, sizeof(chunk), f)) > 0) text.append(chunk, r);
    fclose(f);
    if (!text.empty() && !srw_line_end(text.back())) text.push_back('\n');
  }
  int64_t n = 0;
  int32_t *s = nullptr, *d = nullptr, *p = nullptr;
  float *w = nullptr;
  SRW_TRY(srw_parse_text_device(text.data(), text.size(), params->weighted, params->partitioned, &n, &s, &d, &w, &p));
  srw_status rc = build_for(params, n, s, d, w, p, flags, out);
  cudaFree(s); cudaFree(d); cudaFree(w); cudaFree(p);
  return rc;
}

srw_status srw_graph_load_device(const srw_params *params, unsigned flags, srw_graph **out) {
  {
    struct stat sd;
    if (stat(params->input, &sd) == 0 && S_ISDIR(sd.st_mode)) return load_directory(params, flags, out);
  }
  const int fd = open(params->input, O_RDONLY);
  if (fd < 0) { srw_set_error("Input path does not exist: %s", params->input); return SRW_ERR_IO; }
  struct stat stt;
  if (fstat(fd, &stt) != 0) { close(fd); srw_set_error("cannot stat %s: %s", params->input, strerror(errno)); return SRW_ERR_IO; }
  const size_t len = (size_t)stt.st_size;
  const char *text = nullptr;
  void *map = nullptr;
  std::string small;
  if (len > 0) {
    map = mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
    if (map == MAP_FAILED) {                     // not mappable (pipe, odd filesystem): read it
      map = nullptr;
      small.resize(len);
      size_t got = 0;
      while (got < len) { const ssize_t r = read(fd, &small[got], len - got); if (r <= 0) break; got += (size_t)r; }
      small.resize(got);
      text = small.data();
    } else {
      madvise(map, len, MADV_SEQUENTIAL);
      text = (const char *)map;
    }
  }
  close(fd);
  int64_t n = 0;
  int32_t *s = nullptr, *d = nullptr, *p = nullptr;
  float *w = nullptr;
  srw_status rc = srw_parse_text_device(text, map ? len : small.size(), params->weighted, params->partitioned, &n, &s, &d, &w, &p);
  if (map) munmap(map, len);
  if (rc != SRW_OK) return rc;
  rc = build_for(params, n, s, d, w, p, flags, out);
  cudaFree(s); cudaFree(d); cudaFree(w); cudaFree(p);
  return rc;
}
