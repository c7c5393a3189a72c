// sb_internal.h -- shared internals of libsedumi_b200 (context, error handling, buffers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <utility>
#include <vector>
#include "sedumi_b200.h"

namespace sb {

struct Context {
  bool inited = false;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  bool profiling = false;
  // bump arena for per-call scratch (reset at the start of each public entry point)
  std::vector<std::pair<char *, size_t>> arena_chunks;
  size_t arena_cur = 0, arena_off = 0;
};
// Scratch that lives until the next arena_reset(); never freed in between, so kernels
// enqueued on the stream may keep using it.  Returns nullptr on allocation failure.
void *arena_alloc(size_t bytes);
void arena_reset();
template <typename T> inline T *arena(size_t count) { return (T *)arena_alloc(count * sizeof(T)); }
Context &ctx();
void set_error(const char *fmt, ...);
int  ensure_init();

#define SB_CUDA(call)                                                                 \
  do {                                                                                \
    cudaError_t _e = (call);                                                          \
    if (_e != cudaSuccess) {                                                          \
      sb::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,              \
                    cudaGetErrorString(_e));                                          \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

#define SB_CHECK(cond, ...)                                                           \
  do {                                                                                \
    if (!(cond)) { sb::set_error(__VA_ARGS__); return 1; }                            \
  } while (0)

#define SB_TRY(expr)                                                                  \
  do { int _rc = (expr); if (_rc) return _rc; } while (0)

// Count + check a kernel launch on the library stream (and time it when profiling is on).
void prof_mark(const char *name);
#define SB_LAUNCH_CHECK_N(name)                                                       \
  do {                                                                                \
    sb::ctx().launches++;                                                             \
    if (sb::ctx().profiling) sb::prof_mark(name);                                     \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      sb::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,          \
                    cudaGetErrorString(_e));                                          \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

// RAII device buffer (freed on scope exit).
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  int alloc(size_t count) {
    release();
    n = count;
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu B) failed: %s", count * sizeof(T), cudaGetErrorString(e)); p = nullptr; return 1; }
    return 0;
  }
  int upload(const T *h, size_t count) {
    if (alloc(count)) return 1;
    if (count) {
      cudaError_t e = cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, ctx().stream);
      if (e != cudaSuccess) { set_error("H2D failed: %s", cudaGetErrorString(e)); return 1; }
    }
    return 0;
  }
  int upload(const std::vector<T> &v) { return upload(v.data(), v.size()); }
  int download(T *h, size_t count) const {
    if (!count) return 0;
    cudaError_t e = cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, ctx().stream);
    if (e != cudaSuccess) { set_error("D2H failed: %s", cudaGetErrorString(e)); return 1; }
    return 0;
  }
  int zero() {
    if (!n) return 0;
    cudaError_t e = cudaMemsetAsync(p, 0, n * sizeof(T), ctx().stream);
    if (e != cudaSuccess) { set_error("memset failed: %s", cudaGetErrorString(e)); return 1; }
    return 0;
  }
};

// Content hash of a host buffer.  Buffers above 512 KB are hashed in 256 KB chunks by a small pool of host threads
// (the plugin boundary hashes ~65 MB per IPM iteration to recognise plans and device mirrors; one core does
// 15 GB/s); the result is a function of the bytes only, independent of the number of threads.
uint64_t hash64(const void *data, size_t bytes, uint64_t seed = 0x9E3779B97F4A7C15ull);

// Fast 64-bit content hash: four independent multiply-xorshift lanes over 32-byte blocks
// (the dependent multiply chain of a single-lane hash limits it to ~2 GB/s; this runs at memory speed).
inline uint64_t hash64_st(const void *data, size_t bytes, uint64_t seed = 0x9E3779B97F4A7C15ull) {
  const unsigned char *p = (const unsigned char *)data;
  const uint64_t M = 0xFF51AFD7ED558CCDull;
  uint64_t a = seed ^ bytes, b = seed * 3, c = seed * 5, d = seed * 7;
  size_t nblk = bytes / 32;
  const uint64_t *w = (const uint64_t *)p;
  for (size_t i = 0; i < nblk; i++, w += 4) {
    a = (a ^ w[0]) * M; a ^= a >> 32;
    b = (b ^ w[1]) * M; b ^= b >> 32;
    c = (c ^ w[2]) * M; c ^= c >> 32;
    d = (d ^ w[3]) * M; d ^= d >> 32;
  }
  uint64_t h = a ^ (b * 0xC4CEB9FE1A85EC53ull) ^ (c << 1) ^ (d * M);
  for (size_t i = nblk * 32; i < bytes; i++) { h ^= p[i]; h *= 1099511628211ull; }
  h ^= h >> 33; h *= M; h ^= h >> 33;
  return h;
}

// ---- device mirrors of host arrays, keyed by content: a host entry asks for the
// device copy of an input; if an array with the same size and hash is already resident (typically
// the output of the previous plugin call: ADA travelling getada1 -> 2 -> 3 -> blkchol, or L.L used by
// eight solves) the upload is skipped.  Semantically invisible: same bytes in, same result out.
struct MirrorSlot { uint64_t hash = 0; size_t bytes = 0, cap = 0; void *dev = nullptr; uint64_t stamp = 0; int tag = 0; };
// Returns the device copy of (host, bytes); *hit tells whether an upload was avoided.
void *mirror_input(const void *host, size_t bytes, uint64_t *hash_out = nullptr, bool *hit = nullptr);
// A device buffer (persistent slot) to write an output into; after the D2H copy call
// mirror_publish(slot_dev, host, bytes) so that the next call can find it.
void *mirror_output_slot(size_t bytes);
void mirror_publish(void *slot_dev, const void *host, size_t bytes);

inline uint64_t fnv1a(const void *data, size_t bytes, uint64_t h = 1469598103934665603ull) {
  return hash64(data, bytes, h);
}
inline uint64_t fnv1a_old(const void *data, size_t bytes, uint64_t h = 1469598103934665603ull) {
  const unsigned char *p = (const unsigned char *)data;
  // word-at-a-time variant (not the canonical byte FNV, just a fast content hash)
  size_t nw = bytes / 8;
  const uint64_t *w = (const uint64_t *)p;
  for (size_t i = 0; i < nw; i++) { h ^= w[i]; h *= 1099511628211ull; h ^= h >> 29; }
  for (size_t i = nw * 8; i < bytes; i++) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

inline int to_i32(const sb_idx *src, size_t n, std::vector<int> &dst, const char *what) {
  dst.resize(n);
  for (size_t i = 0; i < n; i++) {
    if (src[i] < 0 || src[i] > 2147483647LL) { set_error("%s[%zu]=%lld out of int32 range", what, i, (long long)src[i]); return 1; }
    dst[i] = (int)src[i];
  }
  return 0;
}

}  // namespace sb
