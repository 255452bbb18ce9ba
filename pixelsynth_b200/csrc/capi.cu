// Library-wide C-ABI bookkeeping: version, error strings, launch counter.
#include <cuda.h>
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

int stream_sms(cudaStream_t stream) {
  static thread_local int dev_cached = -1, dev_sms = 148;
  static thread_local cudaStream_t last_stream = nullptr;
  static thread_local int last_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return dev_sms;
  if (dev != dev_cached) {
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, dev);
    dev_cached = dev;
    last_sms = 0;
  }
  if (!stream || stream == cudaStreamLegacy || stream == cudaStreamPerThread) return dev_sms;
  if (last_sms && stream == last_stream) return last_sms;
  typedef CUresult (*GetGreenFn)(CUstream, CUgreenCtx*);
  typedef CUresult (*GetResFn)(CUgreenCtx, CUdevResource*, CUdevResourceType);
  static GetGreenFn get_green = nullptr;
  static GetResFn get_res = nullptr;
  static bool looked = false;
  if (!looked) {
    void *a = nullptr, *b = nullptr;
    cudaDriverEntryPointQueryResult qa, qb;
    if (cudaGetDriverEntryPoint("cuStreamGetGreenCtx", &a, cudaEnableDefault, &qa) == cudaSuccess &&
        qa == cudaDriverEntryPointSuccess &&
        cudaGetDriverEntryPoint("cuGreenCtxGetDevResource", &b, cudaEnableDefault, &qb) == cudaSuccess &&
        qb == cudaDriverEntryPointSuccess) {
      get_green = (GetGreenFn)a;
      get_res = (GetResFn)b;
    }
    looked = true;
  }
  int n = dev_sms;
  if (get_green) {
    CUgreenCtx g = nullptr;
    CUdevResource r;
    memset(&r, 0, sizeof(r));
    if (get_green((CUstream)stream, &g) == CUDA_SUCCESS && g && get_res(g, &r, CU_DEV_RESOURCE_TYPE_SM) == CUDA_SUCCESS &&
        r.sm.smCount > 0)
      n = (int)r.sm.smCount < dev_sms ? (int)r.sm.smCount : dev_sms;
  }
  last_stream = stream;
  last_sms = n;
  return n;
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
// ---- SM partitions (CUDA green contexts) --------------------------------------------------------------------------
struct SmPartition {
  CUgreenCtx ctx[2];
  std::vector<cudaStream_t> streams;
};

template <class F>
static bool driver_fn(const char* name, F* out) {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return false;
  *out = (F)sym;
  return true;
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

int ps_sm_partition_create(int device, int small_sms, int n_small_streams, void** small_streams, int n_big_streams,
                           void** big_streams, int* small_count, int* big_count, void** handle) {
  PS_CHECK_ARG(device >= 0 && small_sms > 0 && n_small_streams >= 0 && n_big_streams >= 0 && handle);
  PS_CHECK_ARG((n_small_streams == 0 || small_streams) && (n_big_streams == 0 || big_streams));
  CUresult (*dev_get)(CUdevice*, int) = nullptr;
  CUresult (*dev_res)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
  CUresult (*split)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int) = nullptr;
  CUresult (*gen_desc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
  CUresult (*ctx_create)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
  CUresult (*stream_create)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
  CUresult (*ctx_destroy)(CUgreenCtx) = nullptr;
  if (!ps::driver_fn("cuDeviceGet", &dev_get) || !ps::driver_fn("cuDeviceGetDevResource", &dev_res) ||
      !ps::driver_fn("cuDevSmResourceSplitByCount", &split) || !ps::driver_fn("cuDevResourceGenerateDesc", &gen_desc) ||
      !ps::driver_fn("cuGreenCtxCreate", &ctx_create) || !ps::driver_fn("cuGreenCtxStreamCreate", &stream_create) ||
      !ps::driver_fn("cuGreenCtxDestroy", &ctx_destroy))
    return ps::fail(PS_EUNSUPPORTED, "%s: this driver has no green-context entry points%s", __func__, "");
  PS_CUDA(cudaSetDevice(device));
  PS_CUDA(cudaFree(0));  // the primary context exists
  CUdevice dev;
  CUdevResource all, part, rest;
  unsigned int groups = 1;
  memset(&all, 0, sizeof(all));
  memset(&part, 0, sizeof(part));
  memset(&rest, 0, sizeof(rest));
  if (dev_get(&dev, device) != CUDA_SUCCESS || dev_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS)
    return ps::fail(PS_ECUDA, "%s: cuDeviceGetDevResource failed%s", __func__, "");
  if ((unsigned)small_sms >= all.sm.smCount)
    return ps::fail(PS_EINVAL, "%s: small_sms leaves nothing of the device's SMs%s", __func__, "");
  if (split(&part, &groups, &all, &rest, 0, (unsigned)small_sms) != CUDA_SUCCESS || groups != 1 || rest.sm.smCount == 0)
    return ps::fail(PS_ECUDA, "%s: cuDevSmResourceSplitByCount failed%s", __func__, "");
  ps::SmPartition* sp = new ps::SmPartition();
  sp->ctx[0] = sp->ctx[1] = nullptr;
  CUdevResource* res[2] = {&part, &rest};
  void** outs[2] = {small_streams, big_streams};
  const int counts[2] = {n_small_streams, n_big_streams};
  bool ok = true;
  for (int i = 0; i < 2 && ok; ++i) {
    CUdevResourceDesc desc;
    ok = gen_desc(&desc, res[i], 1) == CUDA_SUCCESS && ctx_create(&sp->ctx[i], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS;
    for (int k = 0; k < counts[i] && ok; ++k) {
      CUstream st = nullptr;
      ok = stream_create(&st, sp->ctx[i], CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS;
      if (ok) {
        sp->streams.push_back((cudaStream_t)st);
        outs[i][k] = (void*)st;
      }
    }
  }
  if (!ok) {
    for (cudaStream_t st : sp->streams) cudaStreamDestroy(st);
    for (int i = 0; i < 2; ++i)
      if (sp->ctx[i]) ctx_destroy(sp->ctx[i]);
    delete sp;
    return ps::fail(PS_ECUDA, "%s: creating the green contexts / streams failed%s", __func__, "");
  }
  if (small_count) *small_count = (int)part.sm.smCount;
  if (big_count) *big_count = (int)rest.sm.smCount;
  *handle = sp;
  return PS_OK;
}

int ps_sm_partition_stream(void* handle, int big, int high_priority, void** stream) {
  PS_CHECK_ARG(handle && stream);
  ps::SmPartition* sp = (ps::SmPartition*)handle;
  CUresult (*stream_create)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
  if (!ps::driver_fn("cuGreenCtxStreamCreate", &stream_create))
    return ps::fail(PS_EUNSUPPORTED, "%s: this driver has no green-context entry points%s", __func__, "");
  int least = 0, greatest = 0;
  PS_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));  // numerically lower = scheduled first
  CUstream st = nullptr;
  if (stream_create(&st, sp->ctx[big ? 1 : 0], CU_STREAM_NON_BLOCKING, high_priority ? greatest : 0) != CUDA_SUCCESS)
    return ps::fail(PS_ECUDA, "%s: cuGreenCtxStreamCreate failed%s", __func__, "");
  sp->streams.push_back((cudaStream_t)st);
  *stream = (void*)st;
  return PS_OK;
}

int ps_sm_partition_destroy(void* handle) {
  if (!handle) return PS_OK;
  ps::SmPartition* sp = (ps::SmPartition*)handle;
  CUresult (*ctx_destroy)(CUgreenCtx) = nullptr;
  for (cudaStream_t st : sp->streams) {
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
  }
  if (ps::driver_fn("cuGreenCtxDestroy", &ctx_destroy))
    for (int i = 0; i < 2; ++i)
      if (sp->ctx[i]) ctx_destroy(sp->ctx[i]);
  delete sp;
  return PS_OK;
}

int ps_stream_sm_count(void* stream) { return ps::stream_sms((cudaStream_t)stream); }

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
