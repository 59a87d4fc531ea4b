// Shared host-side helpers: error reporting, launch counting, checked CUDA calls.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/ovo_b200.h"

namespace ovo {

int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);

#define OVO_CUDA(expr)                                                                                    \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess)                                                                                \
      return ::ovo::set_error(OVO_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define OVO_CHECK_LAUNCH()                                                                               \
  do {                                                                                                   \
    ::ovo::count_launch();                                                                               \
    cudaError_t _e = cudaGetLastError();                                                                 \
    if (_e != cudaSuccess)                                                                               \
      return ::ovo::set_error(OVO_E_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
  } while (0)

#define OVO_REQUIRE(cond, ...)                                    \
  do {                                                            \
    if (!(cond)) return ::ovo::set_error(OVO_E_INVALID, __VA_ARGS__); \
  } while (0)

#define OVO_TRY(expr)         \
  do {                        \
    int _r = (expr);          \
    if (_r != OVO_OK) return _r; \
  } while (0)

// The stream-ordered temporaries of the mask post-processing come from the device's default memory pool; by default that
// pool returns memory to the OS at every synchronisation, which makes the next cudaMallocAsync a real allocation
// (milliseconds).  Keep it cached.
void keep_default_mempool_cached();

// Stream-ordered temporaries of one call: every pointer handed out is released with cudaFreeAsync on the same stream when the
// guard goes out of scope, on the success path and on every early error return alike.
struct AsyncTemps {
  cudaStream_t s;
  void* ptrs[32];
  int n = 0;
  explicit AsyncTemps(cudaStream_t stream) : s(stream) {}
  AsyncTemps(const AsyncTemps&) = delete;
  AsyncTemps& operator=(const AsyncTemps&) = delete;
  ~AsyncTemps() { for (int i = 0; i < n; ++i) cudaFreeAsync(ptrs[i], s); }
  template <typename T>
  cudaError_t alloc(T** p, size_t bytes) {
    *p = nullptr;
    if (n >= 32) return cudaErrorMemoryAllocation;
    const cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(p), bytes ? bytes : 1, s);
    if (e == cudaSuccess) ptrs[n++] = *p;
    return e;
  }
  void release(void* p) {   // early release of one buffer (e.g. a table that is re-allocated at another size)
    for (int i = 0; i < n; ++i)
      if (ptrs[i] == p) { cudaFreeAsync(p, s); ptrs[i] = ptrs[--n]; return; }
  }
};

inline int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

// Optional per-launch timing (ovo_profile_begin / ovo_profile_report): CUDA events recorded on the launching
// stream around each kernel, aggregated per kernel class.  Off by default; when on, the encoder runs eagerly.
enum ProfClass : int { PROF_GEMM = 0, PROF_ATTN, PROF_LN, PROF_PRE, PROF_POOL, PROF_ASSOC, PROF_FUSE, PROF_QUERY, PROF_OTHER, PROF_NCLASS };
bool profiling();
struct ProfScope {
  cudaStream_t s;
  int idx;
  ProfScope(cudaStream_t stream, int cls, double flops, double bytes);
  ~ProfScope();
};

}  // namespace ovo
