// Memory-bound glue kernels between the tensor-core convolutions (all NHWC bf16 unless stated):
//   layout conversion (+ the decoder's mask channel, architectures.py:154), AvgPool2d(3,2,1) / bilinear x2
//   resampling with the next layer's noise-conditioned batch-norm + ReLU folded in (blocks.py:45-63,
//   normalization.py:39-47,159-171), LinearNoiseLayer gain/bias (normalization.py:39-47), VQ nearest-code search and
//   code embedding (vqvae.py:41-48,76-77), get_combined (z_buffermodel.py:703-708), tanh(residual) head
//   (architectures.py:157-160).
#include <cuda_bf16.h>

#include "common.cuh"

namespace ps {

__device__ __forceinline__ float ew_act(float v, int act) {
  switch (act) {
    case PS_ACT_RELU: return fmaxf(v, 0.0f);
    case PS_ACT_LEAKY02: return v > 0.0f ? v : 0.2f * v;
    case PS_ACT_TANH: return tanhf(v);
    case PS_ACT_ELU: return v > 0.0f ? v : expm1f(v);
    default: return v;
  }
}

// (N,C,H,W) f32 -> (N,H,W,cstride) bf16; channel C optionally receives mask ? 0 : 1 (float(~background_mask));
// remaining channels are zeroed.
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, int N, int C, int H, int W,
                                                           const uint8_t* __restrict__ mask, __nv_bfloat16* __restrict__ out,
                                                           int cstride) {
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)N * H * W;
  if (pix >= total) return;
  const size_t hw = (size_t)H * W;
  const size_t n = pix / hw, r = pix - n * hw;
  __nv_bfloat16* o = out + pix * cstride;
  auto value = [&](int c) {
    if (c < C) return x[(n * C + c) * hw + r];
    return (c == C && mask) ? (mask[pix] ? 0.0f : 1.0f) : 0.0f;
  };
  if ((cstride & 7) == 0 && (((uintptr_t)out) & 15) == 0) {  // one 16-byte store per 8 channels
    for (int c0 = 0; c0 < cstride; c0 += 8) {
      __nv_bfloat162 h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(value(c0 + 2 * j), value(c0 + 2 * j + 1));
      *reinterpret_cast<uint4*>(o + c0) = *reinterpret_cast<const uint4*>(h);
    }
    return;
  }
  for (int c = 0; c < cstride; ++c) o[c] = __float2bfloat16(value(c));
}

struct ResampleOut {
  __nv_bfloat16* ptr;
  const float* scale;
  const float* shift;
  int per_sample, act, cstride, coffset;
};

// y = act(v * scale + shift) of one (pixel, 8-channel group) into up to two NHWC bf16 outputs.  The activation is
// selected once per group (a switch around the 8-channel loop, not inside it): these kernels are instruction-bound.
template <int ACT>
__device__ __forceinline__ void store_group(const float (&t)[8], __nv_bfloat16* dst) {
  float y[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) y[j] = ew_act(t[j], ACT);
  __nv_bfloat162 h0 = __floats2bfloat162_rn(y[0], y[1]), h1 = __floats2bfloat162_rn(y[2], y[3]);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(y[4], y[5]), h3 = __floats2bfloat162_rn(y[6], y[7]);
  uint4 w;
  w.x = *reinterpret_cast<uint32_t*>(&h0);
  w.y = *reinterpret_cast<uint32_t*>(&h1);
  w.z = *reinterpret_cast<uint32_t*>(&h2);
  w.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(dst) = w;
}

__device__ __forceinline__ void resample_store(const float (&v)[8], int n, int g, int oy, int ox, int Ho, int Wo, int C,
                                               const ResampleOut& o0, const ResampleOut& o1) {
  const ResampleOut* outs[2] = {&o0, &o1};
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const ResampleOut& o = *outs[k];
    if (!o.ptr) continue;
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = v[j];
    if (o.scale) {
      const float4* sc = reinterpret_cast<const float4*>(o.scale + (o.per_sample ? (size_t)n * C : 0) + g * 8);
      const float4 a = __ldg(sc), c = __ldg(sc + 1);
      t[0] *= a.x, t[1] *= a.y, t[2] *= a.z, t[3] *= a.w, t[4] *= c.x, t[5] *= c.y, t[6] *= c.z, t[7] *= c.w;
    }
    if (o.shift) {
      const float4* sh = reinterpret_cast<const float4*>(o.shift + (o.per_sample ? (size_t)n * C : 0) + g * 8);
      const float4 a = __ldg(sh), c = __ldg(sh + 1);
      t[0] += a.x, t[1] += a.y, t[2] += a.z, t[3] += a.w, t[4] += c.x, t[5] += c.y, t[6] += c.z, t[7] += c.w;
    }
    __nv_bfloat16* dst = o.ptr + (((size_t)n * Ho + oy) * Wo + ox) * o.cstride + o.coffset + g * 8;
    switch (o.act) {
      case PS_ACT_RELU: store_group<PS_ACT_RELU>(t, dst); break;
      case PS_ACT_LEAKY02: store_group<PS_ACT_LEAKY02>(t, dst); break;
      case PS_ACT_TANH: store_group<PS_ACT_TANH>(t, dst); break;
      case PS_ACT_ELU: store_group<PS_ACT_ELU>(t, dst); break;
      default: store_group<PS_ACT_NONE>(t, dst); break;
    }
  }
}

// mode 0: identity, 1: AvgPool2d(3, stride 2, pad 1, count_include_pad), 2: bilinear x2 (align_corners=False),
// 3: avg_pool2d(3, 2, 1, count_include_pad=False) (the discriminator's downsample, discriminators.py:170-177),
// 4: MaxPool2d(3, 2, 1) (torchvision resnet18).
// One thread per (output pixel, 8-channel group); up to two outputs y = act(v * scale + shift).
__global__ void __launch_bounds__(256) resample_kernel(const __nv_bfloat16* __restrict__ in, int N, int H, int W, int C,
                                                       int in_cstride, int mode, int Ho, int Wo, ResampleOut o0,
                                                       ResampleOut o1) {
  // grid = (ceil(Wo * groups / 256), Ho, N): one 32-bit division per thread instead of four 64-bit ones on a flat index
  const unsigned groups = (unsigned)C >> 3;
  const unsigned ix = blockIdx.x * blockDim.x + threadIdx.x;
  if (ix >= (unsigned)Wo * groups) return;
  const int ox = (int)(ix / groups);
  const int g = (int)(ix - (unsigned)ox * groups);
  const int oy = blockIdx.y;
  const int n = blockIdx.z;
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto accum = [&](int y, int x, float w) {
    const uint4 raw = *reinterpret_cast<const uint4*>(in + (((size_t)n * H + y) * W + x) * in_cstride + g * 8);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h[j]);
      v[2 * j] += w * f.x;
      v[2 * j + 1] += w * f.y;
    }
  };
  if (mode == 0) {
    accum(oy, ox, 1.0f);
  } else if (mode == 1) {
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int y = 2 * oy + dy, x = 2 * ox + dx;
        if (y >= 0 && y < H && x >= 0 && x < W) accum(y, x, 1.0f / 9.0f);
      }
  } else if (mode == 3) {
    int cnt = 0;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int y = 2 * oy + dy, x = 2 * ox + dx;
        if (y >= 0 && y < H && x >= 0 && x < W) {
          accum(y, x, 1.0f);
          ++cnt;
        }
      }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] /= (float)cnt;
  } else if (mode == 4) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = -INFINITY;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int y = 2 * oy + dy, x = 2 * ox + dx;
        if (y < 0 || y >= H || x < 0 || x >= W) continue;
        const uint4 raw = *reinterpret_cast<const uint4*>(in + (((size_t)n * H + y) * W + x) * in_cstride + g * 8);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h[j]);
          v[2 * j] = fmaxf(v[2 * j], f.x);
          v[2 * j + 1] = fmaxf(v[2 * j + 1], f.y);
        }
      }
  } else {
    // PyTorch upsample_bilinear2d, align_corners=False: src = (dst + 0.5) / 2 - 0.5 clamped at 0
    const float sy = fmaxf((oy + 0.5f) * 0.5f - 0.5f, 0.0f), sx = fmaxf((ox + 0.5f) * 0.5f - 0.5f, 0.0f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = sy - y0, lx = sx - x0;
    accum(y0, x0, (1.0f - ly) * (1.0f - lx));
    accum(y0, x1, (1.0f - ly) * lx);
    accum(y1, x0, ly * (1.0f - lx));
    accum(y1, x1, ly * lx);
  }
  resample_store(v, n, g, oy, ox, Ho, Wo, C, o0, o1);
}

// Bilinear x2 (PyTorch upsample_bilinear2d, align_corners=False) with one thread per INPUT cell (i, j) in
// [-1, H-1] x [-1, W-1] and 8-channel group: the four inputs (i, j), (i, j+1), (i+1, j), (i+1, j+1) (clamped) are loaded
// once and give the up to four outputs (2i+1 | 2i+2, 2j+1 | 2j+2) -- a quarter of the loads and address arithmetic of
// the one-thread-per-output kernel.  Per output the weights and the order of the four terms are exactly those of
// resample_kernel's mode 2 (src = (dst + 0.5) / 2 - 0.5 clamped at 0), so the results are bit-identical.
__global__ void __launch_bounds__(256, 4) upsample2x_kernel(const __nv_bfloat16* __restrict__ in, int N, int H, int W, int C,
                                                         int in_cstride, ResampleOut o0, ResampleOut o1) {
  const unsigned groups = (unsigned)C >> 3;
  const unsigned ix = blockIdx.x * blockDim.x + threadIdx.x;
  if (ix >= (unsigned)(W + 1) * groups) return;
  const int jj = (int)(ix / groups);
  const int g = (int)(ix - (unsigned)jj * groups);
  const int j = jj - 1, i = (int)blockIdx.y - 1;
  const int n = blockIdx.z;
  const int r0 = max(i, 0), r1 = min(i + 1, H - 1), c0 = max(j, 0), c1 = min(j + 1, W - 1);
  float f[4][8];
  {
    const int rr[4] = {r0, r0, r1, r1}, cc[4] = {c0, c1, c0, c1};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const uint4 raw = *reinterpret_cast<const uint4*>(in + (((size_t)n * H + rr[t]) * W + cc[t]) * in_cstride + g * 8);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 v2 = __bfloat1622float2(h[q]);
        f[t][2 * q] = v2.x;
        f[t][2 * q + 1] = v2.y;
      }
    }
  }
  const int Ho = 2 * H, Wo = 2 * W;
#pragma unroll
  for (int dy = 1; dy <= 2; ++dy) {
    const int oy = 2 * i + dy;
    if (oy < 0 || oy >= Ho) continue;
    const float sy = fmaxf((oy + 0.5f) * 0.5f - 0.5f, 0.0f);
    const float ly = sy - (float)(int)sy;
#pragma unroll
    for (int dx = 1; dx <= 2; ++dx) {
      const int ox = 2 * j + dx;
      if (ox < 0 || ox >= Wo) continue;
      const float sx = fmaxf((ox + 0.5f) * 0.5f - 0.5f, 0.0f);
      const float lx = sx - (float)(int)sx;
      const float w[4] = {(1.0f - ly) * (1.0f - lx), (1.0f - ly) * lx, ly * (1.0f - lx), ly * lx};
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] += w[t] * f[t][q];
      resample_store(v, n, g, oy, ox, Ho, Wo, C, o0, o1);
    }
  }
}

// LinearNoiseLayer + bn (eval): gain = 1 + Wg z, bias = Wb z; scale = rsqrt(var + eps) * gain;
// shift = bias - mean * scale   (fused_bn: x * scale - (mean * scale - bias)).  Wg, Wb: (C, Z) already divided
// by their spectral norm.  Channels c >= C of the (N, cpad) outputs get scale 0 / shift 0.
__global__ void noise_affine_kernel(const float* __restrict__ z, int N, int Z, const float* __restrict__ Wg,
                                    const float* __restrict__ Wb, const float* __restrict__ mean,
                                    const float* __restrict__ var, float eps, int C, int cpad, float* __restrict__ scale,
                                    float* __restrict__ shift) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * cpad) return;
  const int n = idx / cpad, c = idx - n * cpad;
  float s = 0.0f, t = 0.0f;
  if (c < C) {
    float g = 0.0f, b = 0.0f;
    for (int k = 0; k < Z; ++k) {
      const float zk = z[n * Z + k];
      g += Wg[c * Z + k] * zk;
      b += Wb[c * Z + k] * zk;
    }
    s = rsqrtf(var[c] + eps) * (1.0f + g);
    t = b - mean[c] * s;
  }
  scale[idx] = s;
  shift[idx] = t;
}

// Nearest codebook entry: argmax_j -(|x|^2 - 2 x.E_j + |E_j|^2)  (vqvae.py:42-48).  x: (N, D, HW) fp32 (NCHW),
// embed: (D, J).  Ties resolve to the smallest j (torch.max returns the first maximum).
// One CTA per 32 pixels: their D-vectors sit in shared memory, a thread owns codes tid, tid + 256, ... and walks the
// codebook once for all 32 pixels (one coalesced load per (d, code) instead of one per (d, code, pixel)); per pixel the
// sums run over d in the same order as a plain loop would, so the scores do not depend on the tiling.
constexpr int VQ_PX = 32;
constexpr int VQ_MAXD = 64;
__global__ void __launch_bounds__(256) vq_argmin_kernel(const float* __restrict__ x, int N, int D, int HW,
                                                        const float* __restrict__ embed, int J,
                                                        long long* __restrict__ ids) {
  __shared__ __align__(16) float xs[VQ_MAXD][VQ_PX];
  __shared__ float x2s[VQ_PX];
  __shared__ float wbest[8][VQ_PX];
  __shared__ int wj[8][VQ_PX];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long total = (long long)N * HW;
  const long long base = (long long)blockIdx.x * VQ_PX;
  for (int i = tid; i < D * VQ_PX; i += 256) {
    const int d = i / VQ_PX, p = i - d * VQ_PX;
    const long long gp = base + p;
    float v = 0.0f;
    if (gp < total) {
      const long long n = gp / HW, r = gp - n * HW;
      v = x[(size_t)n * D * HW + (size_t)d * HW + r];
    }
    xs[d][p] = v;
  }
  __syncthreads();
  if (tid < VQ_PX) {
    float x2 = 0.0f;
    for (int d = 0; d < D; ++d) x2 += xs[d][tid] * xs[d][tid];
    x2s[tid] = x2;
  }
  __syncthreads();
  float best[VQ_PX];
  int bj[VQ_PX];
#pragma unroll
  for (int p = 0; p < VQ_PX; ++p) {
    best[p] = -INFINITY;
    bj[p] = 0x7fffffff;
  }
  for (int j = tid; j < J; j += 256) {
    float dot[VQ_PX];
#pragma unroll
    for (int p = 0; p < VQ_PX; ++p) dot[p] = 0.0f;
    float e2 = 0.0f;
    for (int d = 0; d < D; ++d) {
      const float e = embed[(size_t)d * J + j];
      e2 += e * e;
#pragma unroll
      for (int p4 = 0; p4 < VQ_PX / 4; ++p4) {
        const float4 v = *reinterpret_cast<const float4*>(&xs[d][4 * p4]);
        dot[4 * p4 + 0] += v.x * e;
        dot[4 * p4 + 1] += v.y * e;
        dot[4 * p4 + 2] += v.z * e;
        dot[4 * p4 + 3] += v.w * e;
      }
    }
#pragma unroll
    for (int p = 0; p < VQ_PX; ++p) {
      const float score = -(x2s[p] - 2.0f * dot[p] + e2);
      if (score > best[p]) {  // j ascends within a thread: the first maximum stays
        best[p] = score;
        bj[p] = j;
      }
    }
  }
  // per pixel: best over the warp's lanes, lane p keeps pixel p's; then over the 8 warps
  float mine = -INFINITY;
  int minej = 0x7fffffff;
#pragma unroll
  for (int p = 0; p < VQ_PX; ++p) {
    float b = best[p];
    int j = bj[p];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, b, off);
      const int oj = __shfl_xor_sync(0xffffffffu, j, off);
      if (ob > b || (ob == b && oj < j)) {
        b = ob;
        j = oj;
      }
    }
    if (lane == p) {
      mine = b;
      minej = j;
    }
  }
  wbest[warp][lane] = mine;
  wj[warp][lane] = minej;
  __syncthreads();
  if (tid < VQ_PX && base + tid < total) {
    float b = wbest[0][tid];
    int j = wj[0][tid];
    for (int w = 1; w < 8; ++w) {
      const float ob = wbest[w][tid];
      const int oj = wj[w][tid];
      if (ob > b || (ob == b && oj < j)) {
        b = ob;
        j = oj;
      }
    }
    ids[base + tid] = j;
  }
}

// embed_code: ids (N*HW) -> NHWC bf16 (N*HW, D) from embed (D, J)
__global__ void embed_kernel(const long long* __restrict__ ids, int total, int D, const float* __restrict__ embed, int J,
                             __nv_bfloat16* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total * D) return;
  const int p = idx / D, d = idx - p * D;
  // an id outside the codebook (a caller's bug) must not index outside `embed`: it reads code 0 .. J-1 by clamping
  const long long id = ids[p];
  const int j = id < 0 ? 0 : (id >= J ? J - 1 : (int)id);
  out[idx] = __float2bfloat16(embed[(size_t)d * J + j]);
}

// get_combined: a * (1 - bg) + b * bg per pixel, NCHW f32, bg (N,H,W) u8
__global__ void combine_kernel(const float* __restrict__ a, const float* __restrict__ b, const uint8_t* __restrict__ bg,
                               int N, int C, int HW, float* __restrict__ out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * C * HW) return;
  const size_t n = idx / ((size_t)C * HW), r = idx % HW;
  const float m = bg[n * HW + r] ? 1.0f : 0.0f;
  out[idx] = a[idx] * (1.0f - m) + b[idx] * m;
}

// out = tanh(v + x) (predict_residual) or tanh(v) + x (normalize_before_residual), elementwise f32
__global__ void tanh_residual_kernel(const float* __restrict__ v, const float* __restrict__ x, size_t n, int before,
                                     float* __restrict__ out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  out[idx] = before ? tanhf(v[idx]) + x[idx] : tanhf(v[idx] + x[idx]);
}


// InstanceNorm2d(affine=False) statistics of an NHWC bf16 tensor (the discriminator's norm layer,
// normalization.py:78-79): per (sample, channel) over the H*W pixels, biased variance;
// scale = rsqrt(var + eps), shift = -mean * scale, so that y = x * scale + shift.  grid (C/8, N), 256 threads.
__global__ void __launch_bounds__(256) instnorm_stats_kernel(const __nv_bfloat16* __restrict__ x, int HW, int C, int cstride,
                                                             float eps, float* __restrict__ scale, float* __restrict__ shift) {
  const int g = blockIdx.x, n = blockIdx.y;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int p = threadIdx.x; p < HW; p += 256) {
    const uint4 raw = *reinterpret_cast<const uint4*>(x + ((size_t)n * HW + p) * cstride + g * 8);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h[j]);
      s[2 * j] += f.x;
      s[2 * j + 1] += f.y;
      q[2 * j] = fmaf(f.x, f.x, q[2 * j]);
      q[2 * j + 1] = fmaf(f.y, f.y, q[2 * j + 1]);
    }
  }
  __shared__ float red[2][8][8];  // [sum | sumsq][warp][channel]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
      q[j] += __shfl_xor_sync(0xffffffffu, q[j], o);
    }
    if (lane == 0) {
      red[0][warp][j] = s[j];
      red[1][warp][j] = q[j];
    }
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int j = threadIdx.x;
    float ts = 0.f, tq = 0.f;
    for (int w = 0; w < 8; ++w) {
      ts += red[0][w][j];
      tq += red[1][w][j];
    }
    const float mean = ts / (float)HW;
    const float var = fmaxf(tq / (float)HW - mean * mean, 0.f);
    const float r = rsqrtf(var + eps);
    const int c = g * 8 + j;
    if (c < C) {
      scale[(size_t)n * C + c] = r;
      shift[(size_t)n * C + c] = -mean * r;
    }
  }
}

// The classifier's input as the reference builds it (z_buffermodel.py:105-110,256-257): the candidate's image 0, a
// (3,256,256) f32 buffer, is REINTERPRETED as (256,256,3) (reshape, not permute), scaled to uint8 by truncation, resized
// to 224x224 with PIL's antialiased bilinear filter, divided by 255 and normalised with the ImageNet statistics.
// The filter is Pillow's 8-bit path bit for bit: coefficients in 22-bit fixed point (tap0[224] first tap, kk[224][4]
// integer weights, built on the host like precompute_coeffs / normalize_coeffs_8bpc), horizontal pass then vertical
// pass, each (2^21 + sum kk * p) >> 22 clipped to uint8.  -> NHWC bf16 (M,224,224,8), channels 3..7 zero.
__global__ void __launch_bounds__(256) classifier_input_kernel(const float* __restrict__ img, long long img_stride, int M,
                                                               const int* __restrict__ tap0, const int* __restrict__ kk,
                                                               __nv_bfloat16* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * 224 * 224) return;
  const int ox = idx % 224, oy = (idx / 224) % 224, m = idx / (224 * 224);
  const float* src = img + (size_t)m * img_stride;
  const int x0 = tap0[ox], y0 = tap0[oy];
  int acc[3] = {1 << 21, 1 << 21, 1 << 21};
  for (int ky = 0; ky < 4; ++ky) {
    const int wy = kk[oy * 4 + ky];
    if (wy == 0) continue;
    const int y = y0 + ky;
    int h[3] = {1 << 21, 1 << 21, 1 << 21};
    for (int kx = 0; kx < 4; ++kx) {
      const int wx = kk[ox * 4 + kx];
      if (wx == 0) continue;
      const float* px = src + ((size_t)y * 256 + (x0 + kx)) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int u8 = (int)fminf(fmaxf((px[c] * .5f + .5f) * 255.f, 0.f), 255.f);  // astype(np.uint8): truncation
        h[c] += wx * u8;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += wy * min(max(h[c] >> 22, 0), 255);
  }
  const float mean[3] = {0.485f, 0.456f, 0.406f}, sd[3] = {0.229f, 0.224f, 0.225f};
  __nv_bfloat16 o[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) o[c] = __float2bfloat16(0.f);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float u8 = (float)min(max(acc[c] >> 22, 0), 255);
    o[c] = __float2bfloat16((u8 / 255.f - mean[c]) / sd[c]);
  }
  *reinterpret_cast<uint4*>(out + (size_t)idx * 8) = *reinterpret_cast<const uint4*>(o);
}

}  // namespace ps

using namespace ps;

extern "C" {

int ps_nchw_to_nhwc_bf16(const float* x, int N, int C, int H, int W, const uint8_t* mask, void* out, int cstride,
                         void* stream) {
  PS_CHECK_ARG(x && out && N >= 0 && C >= 1 && H >= 1 && W >= 1 && cstride >= C + (mask ? 1 : 0));
  const size_t total = (size_t)N * H * W;
  if (total == 0) return PS_OK;
  nchw_to_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, N, C, H, W, mask,
                                                                                        (__nv_bfloat16*)out, cstride);
  PS_LAUNCHED();
  return PS_OK;
}

int ps_resample(const void* in, int N, int H, int W, int C, int in_cstride, int mode, const ps_conv_output* out0,
                const ps_conv_output* out1, void* stream) {
  PS_CHECK_ARG(in && out0 && N >= 0 && H >= 1 && W >= 1 && C >= 8 && C % 8 == 0 && in_cstride % 8 == 0);
  PS_CHECK_ARG(mode >= 0 && mode <= 4 && N <= 65535 && H <= 32767);
  const bool half = mode == 1 || mode == 3 || mode == 4;
  const int Ho = half ? (H + 1) / 2 : (mode == 2 ? 2 * H : H);
  const int Wo = half ? (W + 1) / 2 : (mode == 2 ? 2 * W : W);
  ResampleOut o[2];
  memset(o, 0, sizeof(o));
  const ps_conv_output* src[2] = {out0, out1};
  for (int k = 0; k < 2; ++k)
    if (src[k] && src[k]->ptr) {
      PS_CHECK_ARG(src[k]->cstride % 8 == 0 && src[k]->coffset % 8 == 0);
      PS_CHECK_ARG(((uintptr_t)src[k]->scale & 15) == 0 && ((uintptr_t)src[k]->shift & 15) == 0);  // read as float4
      o[k].ptr = (__nv_bfloat16*)src[k]->ptr;
      o[k].scale = src[k]->scale;
      o[k].shift = src[k]->shift;
      o[k].per_sample = src[k]->per_sample;
      o[k].act = src[k]->act;
      o[k].cstride = src[k]->cstride;
      o[k].coffset = src[k]->coffset;
    }
  const size_t total = (size_t)N * Ho * Wo * (C / 8);
  if (total == 0) return PS_OK;
  dim3 grid((unsigned)(((size_t)Wo * (C / 8) + 255) / 256), (unsigned)Ho, (unsigned)N);
  static const bool generic_up = getenv("PS_UPSAMPLE_GENERIC") != nullptr;  // developer A/B switch
  if (mode == 2 && !generic_up) {
    dim3 grid2((unsigned)(((size_t)(W + 1) * (C / 8) + 255) / 256), (unsigned)(H + 1), (unsigned)N);
    upsample2x_kernel<<<grid2, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, N, H, W, C, in_cstride, o[0], o[1]);
  } else {
    resample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)in, N, H, W, C, in_cstride, mode, Ho, Wo, o[0], o[1]);
  }
  PS_LAUNCHED();
  return PS_OK;
}

int ps_instance_norm_stats(const void* x, int N, int HW, int C, int cstride, float eps, float* scale, float* shift,
                           void* stream) {
  PS_CHECK_ARG(x && scale && shift && N >= 0 && HW >= 1 && C >= 1 && cstride >= C && cstride % 8 == 0 && N <= 65535);
  if (N == 0) return PS_OK;
  instnorm_stats_kernel<<<dim3((unsigned)((C + 7) / 8), (unsigned)N), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, HW, C, cstride, eps, scale, shift);
  PS_LAUNCHED();
  return PS_OK;
}

int ps_classifier_input(const float* img, long long img_stride, int M, const int* tap0, const int* kk, void* out,
                        void* stream) {
  PS_CHECK_ARG(img && tap0 && kk && out && M >= 0 && img_stride >= 3 * 256 * 256);
  if (M == 0) return PS_OK;
  const int total = M * 224 * 224;
  classifier_input_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(img, img_stride, M, tap0, kk,
                                                                               (__nv_bfloat16*)out);
  PS_LAUNCHED();
  return PS_OK;
}

int ps_noise_affine(const float* z, int N, int Z, const float* Wg, const float* Wb, const float* mean, const float* var,
                    float eps, int C, int cpad, float* scale, float* shift, void* stream) {
  PS_CHECK_ARG(z && Wg && Wb && mean && var && scale && shift && N >= 1 && Z >= 1 && C >= 1 && cpad >= C);
  const int total = N * cpad;
  noise_affine_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(z, N, Z, Wg, Wb, mean, var, eps, C, cpad,
                                                                             scale, shift);
  PS_LAUNCHED();
  return PS_OK;
}

int ps_vq_argmin(const float* x, int N, int D, int HW, const float* embed, int J, long long* ids, void* stream) {
  PS_CHECK_ARG(x && embed && ids && N >= 0 && D >= 1 && HW >= 1 && J >= 1);
  PS_CHECK_ARG(D <= VQ_MAXD);
  const size_t pixels = (size_t)N * HW;
  if (pixels == 0) return PS_OK;
  vq_argmin_kernel<<<(unsigned)((pixels + VQ_PX - 1) / VQ_PX), 256, 0, (cudaStream_t)stream>>>(x, N, D, HW, embed, J, ids);
  PS_LAUNCHED();
  return PS_OK;
}

int ps_embed_codes(const long long* ids, int total, int D, const float* embed, int J, void* out, void* stream) {
  PS_CHECK_ARG(ids && embed && out && total >= 0 && D >= 1 && J >= 1);
  if (total == 0) return PS_OK;
  embed_kernel<<<(total * D + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ids, total, D, embed, J, (__nv_bfloat16*)out);
  PS_LAUNCHED();
  return PS_OK;
}

int ps_combine(const float* a, const float* b, const uint8_t* bg, int N, int C, int HW, float* out, void* stream) {
  PS_CHECK_ARG(a && b && bg && out && N >= 0 && C >= 1 && HW >= 1);
  const size_t total = (size_t)N * C * HW;
  if (total == 0) return PS_OK;
  combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, b, bg, N, C, HW, out);
  PS_LAUNCHED();
  return PS_OK;
}

int ps_tanh_residual(const float* v, const float* x, long long n, int normalize_before_residual, float* out,
                     void* stream) {
  PS_CHECK_ARG(v && x && out && n >= 0);
  if (n == 0) return PS_OK;
  tanh_residual_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(v, x, (size_t)n,
                                                                                      normalize_before_residual, out);
  PS_LAUNCHED();
  return PS_OK;
}

}  // extern "C"
