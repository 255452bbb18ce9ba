// Library-wide C-ABI bookkeeping: version, error strings, launch counter.
#include <string.h>

#include <mutex>
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

unsigned int* wedge_host_words() {
  static unsigned int* words = nullptr;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (!words) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, ps::PS_WEDGE_WORDS * sizeof(unsigned int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return nullptr;
    memset(p, 0, ps::PS_WEDGE_WORDS * sizeof(unsigned int));
    words = (unsigned int*)p;
  }
  return words;
}

int wedge_check(const char* func) {
  unsigned int* w = wedge_host_words();
  if (w && *(volatile unsigned int*)w) {
    char buf[200];
    snprintf(buf, sizeof(buf), "block %u thread %u %s 0x%x value %u", w[1], w[2],
             w[3] == 0xffffffffu ? "progress wait, needed" : "mbarrier at shared", w[3] == 0xffffffffu ? w[4] : w[3], w[4]);
    return fail(PS_ECUDA, "%s: an earlier launch wedged its barrier protocol and its results are garbage (%s); ps_wedge_reset() clears this",
                func, buf);
  }
  return PS_OK;
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

int ps_wedge_poll(unsigned int* info8) {
  unsigned int* w = ps::wedge_host_words();
  if (!w) return 0;
  if (info8) memcpy(info8, w, 8 * sizeof(unsigned int));
  return *(volatile unsigned int*)w != 0;
}

/* developer aid: the snapshot of progress waiters taken when the watchdog tripped (8 words per waiter, see tc05.cuh) */
int ps_wedge_log(unsigned int* out, int max_words) {
  unsigned int* w = ps::wedge_host_words();
  if (!w || !out) return 0;
  const int n = max_words < ps::PS_WEDGE_WORDS ? max_words : ps::PS_WEDGE_WORDS;
  memcpy(out, w, (size_t)n * sizeof(unsigned int));
  return n;
}

void ps_wedge_reset(void) {
  unsigned int* w = ps::wedge_host_words();
  if (w) memset(w, 0, ps::PS_WEDGE_WORDS * sizeof(unsigned int));
}

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
