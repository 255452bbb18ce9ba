// Shared host/device helpers for libpixelsynth_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pixelsynth_b200.h"

namespace ps {

extern thread_local char g_err_detail[512];
extern thread_local long long g_launches;
// SMs the stream's kernels may occupy: the device's count, or the partition's when the stream belongs to a CUDA green
// context (cuGreenCtxStreamCreate) -- persistent kernels size their grid with this.
int stream_sms(cudaStream_t stream);

inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_err_detail, sizeof(g_err_detail), fmt, a, b);
  return code;
}

#define PS_CHECK_ARG(cond)                                                         \
  do {                                                                             \
    if (!(cond)) return ps::fail(PS_EINVAL, "%s: argument check failed: %s", __func__, #cond); \
  } while (0)

#define PS_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) return ps::fail(PS_ECUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

// after every kernel launch
#define PS_LAUNCHED()                                                                              \
  do {                                                                                             \
    ++ps::g_launches;                                                                              \
    cudaError_t e__ = cudaGetLastError();                                                          \
    if (e__ != cudaSuccess) return ps::fail(PS_ECUDA, "%s: launch failed: %s", __func__, cudaGetErrorString(e__)); \
  } while (0)

// Optional per-kernel device timing (ps_timing_enable / ps_timing_collect): CUDA events recorded on the
// launching stream around selected kernels; used by bench.py for the roofline of the dominant kernel.
extern thread_local int g_timing_on;
void timing_begin(const char* name, cudaStream_t stream);
void timing_end(cudaStream_t stream);
#define PS_TIME_BEGIN(name, stream) \
  do {                              \
    if (ps::g_timing_on) ps::timing_begin(name, stream); \
  } while (0)
#define PS_TIME_END(stream) \
  do {                      \
    if (ps::g_timing_on) ps::timing_end(stream); \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Wedge watchdog of the tensor-core kernels (tc05.cuh): 8 pinned, mapped host words shared by every device; word 0 != 0
// once any barrier wait of any launch timed out (sticky until ps_wedge_reset).  wedge_check() is the first thing every
// C-ABI entry point that launches such a kernel does.
constexpr int PS_WEDGE_WORDS = 8 + 256 * 8 + 64 * 4;  // the record + snapshots of up to 32 progress waiters and 32 mbarrier waiters
unsigned int* wedge_host_words();
int wedge_check(const char* func);
// called by a translation unit before it launches: points its copy of g_wedge_host at the host words and clears its
// device-side flag, once per (host thread, device)
#define PS_WEDGE_ARM()                                                                                \
  do {                                                                                                \
    static thread_local int armed_dev__ = -1;                                                         \
    int dev__ = 0;                                                                                    \
    PS_CUDA(cudaGetDevice(&dev__));                                                                   \
    if (armed_dev__ != dev__) {                                                                       \
      unsigned int* h__ = ps::wedge_host_words();                                                     \
      if (!h__) return ps::fail(PS_ECUDA, "%s: cannot allocate the pinned watchdog words%s", __func__); \
      unsigned int* d__ = nullptr;                                                                    \
      PS_CUDA(cudaHostGetDevicePointer((void**)&d__, h__, 0));                                        \
      unsigned int zero__[8] = {0};                                                                   \
      PS_CUDA(cudaMemcpyToSymbol(g_wedge_host, &d__, sizeof(d__)));                                   \
      PS_CUDA(cudaMemcpyToSymbol(g_wedge, zero__, sizeof(zero__)));                                   \
      armed_dev__ = dev__;                                                                            \
    }                                                                                                 \
  } while (0)

// bump allocator over the caller's workspace
struct Workspace {
  char* base;
  size_t cap, off;
  Workspace(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
  template <class T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* r = (T*)(base + off);
    off += count * sizeof(T);
    return r;
  }
  bool ok() const { return off <= cap; }
};

}  // namespace ps
