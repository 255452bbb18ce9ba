// Library-wide C-ABI bookkeeping: version, error strings, launch counter.
#include <string.h>

#include <vector>

#include "common.cuh"

namespace ps {
thread_local char g_err_detail[512] = {0};
thread_local long long g_launches = 0;
thread_local int g_timing_on = 0;

struct TimedLaunch {
  const char* name;
  cudaEvent_t a, b;
};
static thread_local std::vector<TimedLaunch> g_timed;

void timing_begin(const char* name, cudaStream_t stream) {
  TimedLaunch t{name, nullptr, nullptr};
  if (cudaEventCreate(&t.a) != cudaSuccess || cudaEventCreate(&t.b) != cudaSuccess) return;
  cudaEventRecord(t.a, stream);
  g_timed.push_back(t);
}

void timing_end(cudaStream_t stream) {
  if (!g_timed.empty()) cudaEventRecord(g_timed.back().b, stream);
}
}  // namespace ps

extern "C" {

int ps_abi_version(void) { return PS_ABI_VERSION; }

const char* ps_error_string(int code) {
  switch (code) {
    case PS_OK: return "ok";
    case PS_EINVAL: return "invalid argument";
    case PS_ECUDA: return "CUDA error";
    case PS_EWORKSPACE: return "workspace too small";
    case PS_EUNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
  }
}

const char* ps_last_error_detail(void) { return ps::g_err_detail; }

long long ps_launch_count(void) { return ps::g_launches; }
void ps_launch_count_reset(void) { ps::g_launches = 0; }

void ps_timing_enable(int on) { ps::g_timing_on = on; }

int ps_timing_collect(const char* kernel, double* total_ms, int* launches) {
  double tot = 0.0;
  int n = 0;
  for (auto& t : ps::g_timed) {
    if (cudaEventSynchronize(t.b) == cudaSuccess && (!kernel || strcmp(kernel, t.name) == 0)) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
        tot += ms;
        ++n;
      }
    }
  }
  if (!kernel) {  // a NULL name sums every timed launch and releases the events
    for (auto& t : ps::g_timed) {
      cudaEventDestroy(t.a);
      cudaEventDestroy(t.b);
    }
    ps::g_timed.clear();
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = n;
  return PS_OK;
}

}  // extern "C"
