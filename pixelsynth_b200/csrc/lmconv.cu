// Activation-cached sampler for the locally-masked-convolution PixelCNN over VQ-VAE-2 codes.
//
// Replaces models/lmconv/sample.py:8-73 (sample) driving models/lmconv/model.py:110-155 (OurPixelCNN.forward):
// the reference re-runs the whole 32x32 network for every sampled token (11.43 GFLOP/token).  The masks make the
// network causal in generation order, so the activations of a cell depend only on cells generated earlier and can
// be computed ONCE, when the cell's turn comes (11.16 MFLOP/token).  One persistent CTA per
// image walks the cells in generation order without returning to the host; per cell it evaluates the 42-layer
// column -- u_init gather over the one-hot codes, 14 gated resnets (conv_input, nin_skip, conv_out, PONO, gate),
// 4 dilated convs, nin_out -- reading the cached activations of the (masked-in) neighbours, then draws the code
// from softmax(logits / T) with the caller's uniform number and writes it where later cells will read it.
// Masked-out taps are skipped outright: their weights are never fetched.
//
// Layouts: weights bf16 [tap][cin][cout] (cout contiguous: a k-row is one coalesced segment), fp32 accumulation;
// activation cache fp32 (B, 33 tensors, 1024 cells, 80 channels) in global memory (L2-resident neighbourhood).
// Images are independent, so there is no inter-CTA communication at all.
#include <cuda_bf16.h>

#include "common.cuh"

namespace ps {

constexpr int LM_THREADS = 1024;  // many k-groups: the column GEMVs are L2-latency bound, so rows per thread must be few
constexpr int LM_G = 32;          // grid side
constexpr int LM_CELLS = LM_G * LM_G;
constexpr int LM_F = 80;          // nr_filters
constexpr int LM_CLASSES = 512;
constexpr int LM_TENSORS = 33;
constexpr int LM_NOPS = 18;

struct LmOp {
  int kind;  // 0 gated resnet, 1 dilated conv + PONO
  int og, a, mid, out;
  int w_in, b_in, w_skip, b_skip, w_out, b_out;
};

struct LmParams {
  const __nv_bfloat16* W;
  const float* bias;
  int w_uinit, b_uinit, w_nin, b_nin;
  LmOp ops[LM_NOPS];
  const int* order;            // (B, 1024) cell index per step
  const uint16_t* words;       // (B, 3, 1024)
  const uint8_t* sample_mask;  // (B, 1024)
  long long* codes;            // (B, 1024) in/out
  const float* uniforms;       // (B, ustride)
  int ustride;
  float inv_temperature;
  float* cache;       // (B, 33, 1024, 80)
  float* logits_out;  // (B, 1024, 512) or null
  const int* nsteps;  // (B)
  int sample;
};

struct LmSmem {
  float xin[9 * 2 * LM_F];
  float partial[8192];
  float res[LM_CLASSES];
  float og[LM_F], mid[LM_F];
  int taps[3][9], nbr[3][9], nact[3];
  float red[32];
  float bcast[2];
  int token;
};

__device__ __forceinline__ float elu1(float v) { return v > 0.0f ? v : expm1f(v); }

// out[co] = bias[co] + sum_k xin[k] * W[row(k)][co], k over (active tap, cin); result in sm.res[0..Cout)
template <int COUT>
__device__ __forceinline__ void gemv(const LmParams& p, LmSmem& sm, int w_off, int b_off, int cin, const int* taps,
                                     int nact) {
  constexpr int TPR = COUT / 8;              // threads per k-row, 8 output channels each
  constexpr int NG = LM_THREADS / TPR;       // k groups
  const int tid = threadIdx.x;
  const int g = tid / TPR, cq = tid - g * TPR;
  const int K = nact * cin;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (g < NG) {
    const int k0 = (int)((long long)K * g / NG), k1 = (int)((long long)K * (g + 1) / NG);
    int ti = k0 / cin, ci = k0 - ti * cin;
    const __nv_bfloat16* wbase = p.W + w_off + cq * 8;
    int k = k0;
    while (k < k1) {
      const int run = min(k1 - k, cin - ci);  // stay inside one tap block: rows are contiguous there
      const __nv_bfloat16* wr = wbase + (size_t)(taps[ti] * cin + ci) * COUT;
      const float* xr = sm.xin + k;
      int j = 0;
      for (; j + 8 <= run; j += 8) {
        uint4 w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = __ldg(reinterpret_cast<const uint4*>(wr + (size_t)(j + u) * COUT));
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float x = xr[j + u];
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&w[u]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(h[e]);
            acc[2 * e] = fmaf(x, f.x, acc[2 * e]);
            acc[2 * e + 1] = fmaf(x, f.y, acc[2 * e + 1]);
          }
        }
      }
      for (; j < run; ++j) {
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(wr + (size_t)j * COUT));
        const float x = xr[j];
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(h[e]);
          acc[2 * e] = fmaf(x, f.x, acc[2 * e]);
          acc[2 * e + 1] = fmaf(x, f.y, acc[2 * e + 1]);
        }
      }
      k += run;
      ci += run;
      if (ci == cin) {
        ci = 0;
        ++ti;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) sm.partial[g * COUT + cq * 8 + e] = acc[e];
  }
  __syncthreads();
  {  // R threads per output channel sum the k-group partials, then a shuffle tree
    constexpr int R = COUT <= 128 ? 8 : (COUT <= 256 ? 4 : 2);
    const int co = tid / R, part = tid - co * R;
    float s = 0.f;
    if (co < COUT)
      for (int gg = part; gg < NG; gg += R) s += sm.partial[gg * COUT + co];
#pragma unroll
    for (int off = R / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (co < COUT && part == 0) sm.res[co] = s + p.bias[b_off + co];
  }
  __syncthreads();
}

// positional normalisation of v[0..80): (x - mean) / sqrt(var_unbiased + 1e-5)   (layers.py:224-236)
__device__ __forceinline__ void pono(LmSmem& sm, float* v) {
  const int tid = threadIdx.x;
  if (tid < 32) {
    float s = 0.f;
    for (int i = tid; i < LM_F; i += 32) s += v[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const float mean = s / LM_F;
    float q = 0.f;
    for (int i = tid; i < LM_F; i += 32) {
      const float d = v[i] - mean;
      q += d * d;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
    const float inv = 1.0f / sqrtf(q / (LM_F - 1) + 1e-5f);
    for (int i = tid; i < LM_F; i += 32) v[i] = (v[i] - mean) * inv;
  }
  __syncthreads();
}

// xin[j*2F + c] = elu(t[nbr_j][c]), xin[j*2F + F + c] = elu(-t[nbr_j][c]) for the active taps of mask m
__device__ __forceinline__ void gather_celu(LmSmem& sm, const float* tensor, int m) {
  const int n = sm.nact[m] * LM_F;
  for (int i = threadIdx.x; i < n; i += LM_THREADS) {
    const int j = i / LM_F, c = i - j * LM_F;
    const float v = tensor[(size_t)sm.nbr[m][j] * LM_F + c];
    sm.xin[j * 2 * LM_F + c] = elu1(v);
    sm.xin[j * 2 * LM_F + LM_F + c] = elu1(-v);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(LM_THREADS) lmconv_sample_kernel(const LmParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LmSmem& sm = *reinterpret_cast<LmSmem*>(smem_raw);
  const int b = blockIdx.x, tid = threadIdx.x;
  const int* order = p.order + (size_t)b * LM_CELLS;
  const uint16_t* words = p.words + (size_t)b * 3 * LM_CELLS;
  long long* codes = p.codes + (size_t)b * LM_CELLS;
  float* cache = p.cache + (size_t)b * LM_TENSORS * LM_CELLS * LM_F;
  auto tensor = [&](int id) { return cache + (size_t)id * LM_CELLS * LM_F; };
  const int nsteps = p.nsteps[b];
  int drawn = 0;

  for (int t = 0; t < nsteps; ++t) {
    const int pos = order[t];
    const int r = pos / LM_G, c = pos - r * LM_G;
    if (tid < 3) {  // active taps + neighbour cells of the three masks (A dil 1, B dil 1, B dil 2)
      const unsigned w = words[tid * LM_CELLS + pos];
      const int dil = tid == 2 ? 2 : 1;
      int n = 0;
      for (int tp = 0; tp < 9; ++tp)
        if (w >> tp & 1u) {
          sm.taps[tid][n] = tp;
          sm.nbr[tid][n] = (r + (tp / 3 - 1) * dil) * LM_G + c + (tp % 3 - 1) * dil;
          ++n;
        }
      sm.nact[tid] = n;
    }
    __syncthreads();

    // ---- u_init over [one-hot(code) | ones]: a gather of weight columns (mask A) ----
    if (tid < LM_F) {
      float s = p.bias[p.b_uinit + tid];
      for (int j = 0; j < sm.nact[0]; ++j) {
        const int tp = sm.taps[0][j];
        const int code = (int)codes[sm.nbr[0][j]];
        const __nv_bfloat16* wt = p.W + p.w_uinit + (size_t)tp * (LM_CLASSES + 1) * LM_F + tid;
        s += __bfloat162float(wt[(size_t)code * LM_F]) + __bfloat162float(wt[(size_t)LM_CLASSES * LM_F]);
      }
      sm.og[tid] = s;
    }
    __syncthreads();
    pono(sm, sm.og);
    if (tid < LM_F) tensor(0)[(size_t)pos * LM_F + tid] = sm.og[tid];
    __syncthreads();

    for (int oi = 0; oi < LM_NOPS; ++oi) {
      const LmOp& op = p.ops[oi];
      if (op.kind == 0) {
        // x = PONO(conv_input(concat_elu(og))) [+ nin_skip(concat_elu(a))]
        gather_celu(sm, tensor(op.og), 1);
        gemv<LM_F>(p, sm, op.w_in, op.b_in, 2 * LM_F, sm.taps[1], sm.nact[1]);
        if (tid < LM_F) sm.mid[tid] = sm.res[tid];
        __syncthreads();
        pono(sm, sm.mid);
        if (op.a >= 0) {
          if (tid < LM_F) {
            const float v = tensor(op.a)[(size_t)pos * LM_F + tid];
            sm.xin[tid] = elu1(v);
            sm.xin[LM_F + tid] = elu1(-v);
          }
          __syncthreads();
          const int one_tap[1] = {0};
          gemv<LM_F>(p, sm, op.w_skip, op.b_skip, 2 * LM_F, one_tap, 1);
          if (tid < LM_F) sm.mid[tid] += sm.res[tid];
        }
        if (tid < LM_F) tensor(op.mid)[(size_t)pos * LM_F + tid] = sm.mid[tid];
        __syncthreads();
        // y = conv_out(concat_elu(x)); out = og + PONO(y[:80]) * sigmoid(y[80:])
        gather_celu(sm, tensor(op.mid), 1);
        gemv<2 * LM_F>(p, sm, op.w_out, op.b_out, 2 * LM_F, sm.taps[1], sm.nact[1]);
        pono(sm, sm.res);
        if (tid < LM_F) {
          const float o = sm.og[tid] + sm.res[tid] / (1.0f + __expf(-sm.res[LM_F + tid]));
          sm.og[tid] = o;
          tensor(op.out)[(size_t)pos * LM_F + tid] = o;
        }
        __syncthreads();
      } else {
        // dilated masked conv on the raw stream + PONO
        const int n = sm.nact[2] * LM_F;
        for (int i = tid; i < n; i += LM_THREADS) {
          const int j = i / LM_F, ch = i - j * LM_F;
          sm.xin[i] = tensor(op.og)[(size_t)sm.nbr[2][j] * LM_F + ch];
        }
        __syncthreads();
        gemv<LM_F>(p, sm, op.w_in, op.b_in, LM_F, sm.taps[2], sm.nact[2]);
        if (tid < LM_F) sm.og[tid] = sm.res[tid];
        __syncthreads();
        pono(sm, sm.og);
        if (tid < LM_F) tensor(op.out)[(size_t)pos * LM_F + tid] = sm.og[tid];
        __syncthreads();
      }
    }

    const bool do_sample = p.sample && p.sample_mask[(size_t)b * LM_CELLS + pos];
    if (!do_sample && !p.logits_out) continue;  // known cell: only its activations were needed

    // ---- logits = nin_out(elu(u)) ----
    if (tid < LM_F) sm.xin[tid] = elu1(sm.og[tid]);
    __syncthreads();
    {
      const int one_tap[1] = {0};
      gemv<LM_CLASSES>(p, sm, p.w_nin, p.b_nin, LM_F, one_tap, 1);
    }
    if (p.logits_out) {
      float* lo = p.logits_out + ((size_t)b * LM_CELLS + pos) * LM_CLASSES;
      for (int i = tid; i < LM_CLASSES; i += LM_THREADS) lo[i] = sm.res[i];
    }
    if (do_sample) {
      // token = first j with cumsum(softmax(logits / T))_j > u
      const int lane = tid & 31, warp = tid >> 5;
      const bool act = tid < LM_CLASSES / 2;  // 256 threads own two logits each
      const float l0 = act ? sm.res[2 * tid] * p.inv_temperature : -INFINITY;
      const float l1 = act ? sm.res[2 * tid + 1] * p.inv_temperature : -INFINITY;
      float mx = fmaxf(l0, l1);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      if (lane == 0 && warp < 8) sm.red[warp] = mx;
      __syncthreads();
      if (tid == 0) {
        float m = sm.red[0];
        for (int w = 1; w < 8; ++w) m = fmaxf(m, sm.red[w]);
        sm.bcast[0] = m;
        sm.token = LM_CLASSES - 1;
      }
      __syncthreads();
      mx = sm.bcast[0];
      const float e0 = act ? __expf(l0 - mx) : 0.f, e1 = act ? __expf(l1 - mx) : 0.f;
      float incl = e0 + e1;  // inclusive scan of the per-thread pair sums
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const float v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
      }
      if (lane == 31 && warp < 8) sm.red[warp] = incl;
      __syncthreads();
      float base = 0.f, total = 0.f;
      for (int w = 0; w < 8; ++w) {
        if (w < warp) base += sm.red[w];
        total += sm.red[w];
      }
      const float thr = p.uniforms[(size_t)b * p.ustride + drawn] * total;
      const float c0 = base + incl - e1, c1 = base + incl;
      if (act) {
        if (c0 > thr)
          atomicMin(&sm.token, 2 * tid);
        else if (c1 > thr)
          atomicMin(&sm.token, 2 * tid + 1);
      }
      __syncthreads();
      if (tid == 0) codes[pos] = sm.token;
      ++drawn;
      __syncthreads();
    }
  }
}

}  // namespace ps

using namespace ps;

extern "C" {

size_t ps_lmconv_cache_bytes(int B) { return (size_t)(B > 0 ? B : 0) * LM_TENSORS * LM_CELLS * LM_F * sizeof(float); }

int ps_lmconv_sample(const ps_lmconv_weights* w, int B, const int* order, const uint16_t* words,
                     const uint8_t* sample_mask, long long* codes, const float* uniforms, int uniforms_stride,
                     float temperature, const int* nsteps, int sample, float* logits_out, void* cache,
                     size_t cache_bytes, void* stream) {
  PS_CHECK_ARG(w && w->weights && w->bias && order && words && sample_mask && codes && nsteps && cache);
  PS_CHECK_ARG(B >= 0 && temperature > 0.0f);
  PS_CHECK_ARG(!sample || uniforms);
  if (B == 0) return PS_OK;
  if (cache_bytes < ps_lmconv_cache_bytes(B)) return fail(PS_EWORKSPACE, "%s: activation cache too small%s", __func__);
  LmParams p;
  memset(&p, 0, sizeof(p));
  p.W = (const __nv_bfloat16*)w->weights;
  p.bias = w->bias;
  p.w_uinit = w->w_uinit;
  p.b_uinit = w->b_uinit;
  p.w_nin = w->w_nin;
  p.b_nin = w->b_nin;
  for (int i = 0; i < LM_NOPS; ++i) {
    const ps_lmconv_op& s = w->ops[i];
    LmOp& d = p.ops[i];
    d.kind = s.kind;
    d.og = s.og;
    d.a = s.a;
    d.mid = s.mid;
    d.out = s.out;
    d.w_in = s.w_in;
    d.b_in = s.b_in;
    d.w_skip = s.w_skip;
    d.b_skip = s.b_skip;
    d.w_out = s.w_out;
    d.b_out = s.b_out;
    PS_CHECK_ARG(d.og >= 0 && d.og < LM_TENSORS && d.out >= 0 && d.out < LM_TENSORS && d.a < LM_TENSORS);
  }
  p.order = order;
  p.words = words;
  p.sample_mask = sample_mask;
  p.codes = codes;
  p.uniforms = uniforms;
  p.ustride = uniforms_stride;
  p.inv_temperature = 1.0f / temperature;
  p.cache = (float*)cache;
  p.logits_out = logits_out;
  p.nsteps = nsteps;
  p.sample = sample;
  PS_TIME_BEGIN("lmconv_sample_kernel", (cudaStream_t)stream);
  lmconv_sample_kernel<<<B, LM_THREADS, sizeof(LmSmem), (cudaStream_t)stream>>>(p);
  PS_TIME_END((cudaStream_t)stream);
  PS_LAUNCHED();
  return PS_OK;
}

}  // extern "C"
