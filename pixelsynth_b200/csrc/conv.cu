// Implicit-GEMM convolution for sm_100a: TMA-staged NHWC tiles -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM)
// -> fused epilogue.  One kernel serves every dense layer of the PixelSynth inference path:
//   reference models/networks/architectures.py:174-279 (Unet: 4x4 stride-2 and 3x3 convs),
//   reference models/layers/blocks.py:33-74 (ResNet_Block of the refinement decoder: 3x3 + fused 1x1 skip),
//   reference models/vqvae2/vqvae.py:80-161 (VQ-VAE-2 encoder/decoder: 4x4 s2, 3x3, 1x1, 4x4 transposed s2).
//
// Formulation.  Activations are NHWC bf16 (C padded to a multiple of 8).  A CTA owns one tile of 128 output
// pixels (TN images x TH rows x TW columns) x BN output channels.  The GEMM K dimension runs over
// (tap, 64-channel chunk): for tap (dy,dx) the A operand is the input window shifted by the tap, fetched by ONE
// 4-D TMA box {64 ch, TW, TH, TN} at element strides {1,s,s,1}; out-of-image coordinates are zero-filled by the
// TMA unit, which is exactly the convolution's zero padding.  Both operands land in shared memory in the
// 128-byte-swizzled K-major layout tcgen05 consumes directly.  A second input tensor with its own taps can be
// accumulated into the same tile (the decoder's 1x1 skip convolution, blocks.py:43,65-66).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2..5 = epilogue (each reads its 32-lane TMEM quadrant with tcgen05.ld, applies bias / residual /
// per-sample scale+shift / activation and writes up to two NHWC outputs).  smem ring of STAGES stages guarded by
// full/empty mbarriers; MMA completion is signalled with tcgen05.commit.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc05.cuh"

namespace ps {

constexpr int CONV_THREADS = 192;
constexpr int BM = 128;  // output pixels per tile (UMMA M)
constexpr int BK = 64;   // channels per K step (128 bytes of bf16 = one swizzle row)
constexpr int A_STAGE_BYTES = BM * BK * 2;

struct ConvKernelParams {
  // tile geometry: TW * TH * TN == 128
  int TW, TH, TN;
  int tiles_x, tiles_y;
  int Hout, Wout, N;  // output grid this launch computes (per phase for transposed convs)
  int stride;         // input pixel = out * stride + tap offset
  int ntaps[2], kchunks[2];
  int dy[2][16], dx[2][16];
  int wrow[2][16];  // first weight row (of the packed [rows][Cin_pad] matrix) of each tap
  int BN;           // output channels per CTA (multiple of 16, <= 256)
  int stages;
  // epilogue
  int Cout;  // real output channels (columns >= Cout are dropped)
  const float* bias;
  const __nv_bfloat16* residual;  // NHWC on the full output grid, or null
  int res_cstride;
  // output o (o = 0, 1): y = act(v * scale + shift); scale/shift per channel ([Cout]) or per sample ([N][Cout])
  __nv_bfloat16* out[2];
  const float* scale[2];
  const float* shift[2];
  int per_sample[2];
  int act[2];
  int out_cstride[2], out_coffset[2];
  float* out_f32_nchw;  // optional fp32 NCHW copy of output 0's values
  float act_param[2];   // sigmoid-depth: y = sigmoid(v) * act_param[0] + act_param[1]
  // full output geometry: tile pixel (oy, ox) -> (oy * out_sy + out_py, ox * out_sx + out_px)
  int out_H, out_W, out_sy, out_sx, out_py, out_px;
};

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_TANH = 3, ACT_SIGMOID_AFFINE = 4, ACT_ELU = 5 };

__device__ __forceinline__ float apply_act(float v, int act, const float* ap) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.0f);
    case ACT_LEAKY: return v > 0.0f ? v : 0.2f * v;
    case ACT_TANH: return tanhf(v);
    case ACT_SIGMOID_AFFINE: return ap[0] / (1.0f + __expf(-v)) + ap[1];
    case ACT_ELU: return v > 0.0f ? v : expm1f(v);
    default: return v;
  }
}

__global__ void __launch_bounds__(CONV_THREADS, 1)
    conv_igemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                      const __grid_constant__ CUtensorMap mapW, const ConvKernelParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = p.BN;
  const int stage_bytes = A_STAGE_BYTES + BN * BK * 2;
  // carve: [stages x (A | B)] then barriers
  unsigned char* tiles = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(tiles + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* accum_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);

  // tile coordinates
  int tile = blockIdx.x;
  const int tx = tile % p.tiles_x;
  tile /= p.tiles_x;
  const int ty = tile % p.tiles_y;
  const int tn = tile / p.tiles_y;
  const int ox0 = tx * p.TW, oy0 = ty * p.TH, n0 = tn * p.TN;
  const int ncol0 = blockIdx.y * BN;
  const int kiters = p.ntaps[0] * p.kchunks[0] + p.ntaps[1] * p.kchunks[1];

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapW);
    if (p.ntaps[1]) tma_prefetch_desc(&mapA1);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(accum_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // TMEM: power-of-two column count >= 32 covering BN fp32 columns
    const uint32_t ncols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int src = 0; src < 2; ++src) {
        const CUtensorMap* mA = src ? &mapA1 : &mapA0;
        for (int t = 0; t < p.ntaps[src]; ++t) {
          const int ix0 = ox0 * p.stride + p.dx[src][t];
          const int iy0 = oy0 * p.stride + p.dy[src][t];
          for (int kc = 0; kc < p.kchunks[src]; ++kc, ++it) {
            const int s = it % p.stages;
            const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
            mbar_wait(&empty_bar[s], ph ^ 1u);
            unsigned char* a = tiles + (size_t)s * stage_bytes;
            mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
            tma_load_4d(mA, &full_bar[s], a, kc * BK, ix0, iy0, n0);
            tma_load_2d(&mapW, &full_bar[s], a + A_STAGE_BYTES, kc * BK, p.wrow[src][t] + ncol0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = umma_idesc_bf16(BN);
    for (int it = 0; it < kiters; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
      mbar_wait(&full_bar[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a = smem_u32(tiles + (size_t)s * stage_bytes);
        const uint32_t b = a + A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_bf16(tmem_base, umma_desc_sw128(a + k * 32), umma_desc_sw128(b + k * 32), idesc, (it | k) ? 1u : 0u);
        umma_commit(&empty_bar[s]);                      // frees the smem stage once these MMAs have read it
        if (it == kiters - 1) umma_commit(accum_bar);    // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 =====
    const int q = warp & 3;
    const int m = q * 32 + lane;  // tile row = output pixel
    const int nl = m / (p.TH * p.TW);
    const int rem = m - nl * (p.TH * p.TW);
    const int oy = oy0 + rem / p.TW, ox = ox0 + rem % p.TW;
    const int n = n0 + nl;
    const bool valid = n < p.N && oy < p.Hout && ox < p.Wout;
    const int fy = oy * p.out_sy + p.out_py, fx = ox * p.out_sx + p.out_px;
    const size_t pixel = ((size_t)n * p.out_H + fy) * p.out_W + fx;
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (!valid) continue;
      const int cbase = ncol0 + c0;
      if (cbase >= p.Cout) continue;
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        f[j] = __uint_as_float(v[j]);
        const int c = cbase + j;
        if (p.bias && c < p.Cout) f[j] += __ldg(p.bias + c);
      }
      if (p.residual) {
        const __nv_bfloat16* r = p.residual + pixel * p.res_cstride + cbase;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (cbase + j < p.Cout) f[j] += __bfloat162float(r[j]);
      }
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        if (!p.out[o]) continue;
        const float* sc = p.scale[o] ? p.scale[o] + (p.per_sample[o] ? (size_t)n * p.Cout : 0) : nullptr;
        const float* sh = p.shift[o] ? p.shift[o] + (p.per_sample[o] ? (size_t)n * p.Cout : 0) : nullptr;
        __nv_bfloat16* dst = p.out[o] + pixel * p.out_cstride[o] + p.out_coffset[o] + cbase;
        float y[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = cbase + j;
          float t = f[j];
          if (c < p.Cout) {
            if (sc) t *= __ldg(sc + c);
            if (sh) t += __ldg(sh + c);
          }
          y[j] = apply_act(t, p.act[o], p.act_param);
        }
        const bool full = cbase + 32 <= p.Cout && ((((uintptr_t)dst) & 15) == 0);
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(y[j], y[j + 1]), h1 = __floats2bfloat162_rn(y[j + 2], y[j + 3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(y[j + 4], y[j + 5]), h3 = __floats2bfloat162_rn(y[j + 6], y[j + 7]);
            uint4 w;
            w.x = *reinterpret_cast<uint32_t*>(&h0);
            w.y = *reinterpret_cast<uint32_t*>(&h1);
            w.z = *reinterpret_cast<uint32_t*>(&h2);
            w.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(dst + j) = w;
          }
        } else {
          for (int j = 0; j < 32; ++j)
            if (cbase + j < p.Cout) dst[j] = __float2bfloat16(y[j]);
        }
        if (o == 0 && p.out_f32_nchw) {
          for (int j = 0; j < 32; ++j)
            if (cbase + j < p.Cout)
              p.out_f32_nchw[(((size_t)n * p.Cout + cbase + j) * p.out_H + fy) * p.out_W + fx] = y[j];
        }
      }
      if (!p.out[0] && p.out_f32_nchw) {  // fp32-only output
        const float* sc = p.scale[0] ? p.scale[0] + (p.per_sample[0] ? (size_t)n * p.Cout : 0) : nullptr;
        const float* sh = p.shift[0] ? p.shift[0] + (p.per_sample[0] ? (size_t)n * p.Cout : 0) : nullptr;
        for (int j = 0; j < 32; ++j) {
          const int c = cbase + j;
          if (c < p.Cout) {
            float t = f[j];
            if (sc) t *= __ldg(sc + c);
            if (sh) t += __ldg(sh + c);
            p.out_f32_nchw[(((size_t)n * p.Cout + c) * p.out_H + fy) * p.out_W + fx] = apply_act(t, p.act[0], p.act_param);
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    const uint32_t ncols = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  }
  return fn;
}

// NHWC bf16 activation: dims (C, W, H, N), box (64, TW*s, TH*s, TN) traversed at element strides (1, s, s, 1)
static int make_act_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int cstride, int TW, int TH, int TN,
                        int s) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(PS_ECUDA, "%s: cuTensorMapEncodeTiled unavailable%s", __func__);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)cstride * 2, (cuuint64_t)W * cstride * 2, (cuuint64_t)H * W * cstride * 2};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(TW * s), (cuuint32_t)(TH * s), (cuuint32_t)TN};
  cuuint32_t estr[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%d", (int)r);
    return fail(PS_ECUDA, "%s: cuTensorMapEncodeTiled(activation) failed: CUresult %s", __func__, buf);
  }
  return PS_OK;
}

// packed weights: [rows][Cin_pad] bf16, K-major; box (64, BN)
static int make_w_map(CUtensorMap* m, const void* ptr, int rows, int cin_pad, int BN) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(PS_ECUDA, "%s: cuTensorMapEncodeTiled unavailable%s", __func__);
  cuuint64_t dims[2] = {(cuuint64_t)cin_pad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cin_pad * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%d", (int)r);
    return fail(PS_ECUDA, "%s: cuTensorMapEncodeTiled(weights) failed: CUresult %s", __func__, buf);
  }
  return PS_OK;
}

}  // namespace ps

using namespace ps;

extern "C" int ps_conv_igemm(const ps_conv_desc* d, void* stream) {
  PS_CHECK_ARG(d != nullptr);
  PS_CHECK_ARG(d->in[0].ptr && d->weights);
  PS_CHECK_ARG(d->N >= 1 && d->Hout >= 1 && d->Wout >= 1 && d->Cout >= 1);
  PS_CHECK_ARG(d->stride == 1 || d->stride == 2);
  PS_CHECK_ARG(d->in[0].ntaps >= 1 && d->in[0].ntaps <= 16 && d->in[1].ntaps >= 0 && d->in[1].ntaps <= 16);
  PS_CHECK_ARG(d->cout_pad % 16 == 0 && d->cout_pad >= d->Cout);
  PS_CHECK_ARG(d->out[0].ptr || d->out[1].ptr || d->out_f32_nchw);
  ConvKernelParams p;
  memset(&p, 0, sizeof(p));
  // tile shape: widest power-of-two column count <= 16 that the output width fills, then rows, then images
  int TW = 16;
  while (TW > 1 && TW / 2 >= d->Wout) TW /= 2;
  int TH = BM / TW;
  while (TH > 1 && TH / 2 >= d->Hout) TH /= 2;
  if (TH > 8 && d->Hout >= 8 && TW == 16) TH = 8;
  int TN = BM / (TW * TH);
  p.TW = TW;
  p.TH = TH;
  p.TN = TN;
  p.tiles_x = (d->Wout + TW - 1) / TW;
  p.tiles_y = (d->Hout + TH - 1) / TH;
  const int tiles_n = (d->N + TN - 1) / TN;
  p.Hout = d->Hout;
  p.Wout = d->Wout;
  p.N = d->N;
  p.stride = d->stride;
  // BN: whole padded Cout when it fits one UMMA (<= 256), else 128-wide column blocks
  int BN = d->cout_pad <= 256 ? d->cout_pad : 128;
  PS_CHECK_ARG(d->cout_pad % BN == 0);
  p.BN = BN;
  const int stage_bytes = A_STAGE_BYTES + BN * BK * 2;
  // <= 128 columns: 3 stages (97 KB) so two CTAs share an SM and one's epilogue overlaps the other's main loop
  const int stages = BN <= 128 ? 3 : 4;
  p.stages = stages;
  CUtensorMap mapA[2], mapW;
  memset(mapA, 0, sizeof(mapA));
  int wrows = 0;
  for (int s = 0; s < 2; ++s) {
    const ps_conv_input& in = d->in[s];
    p.ntaps[s] = in.ntaps;
    if (in.ntaps == 0) continue;
    PS_CHECK_ARG(in.ptr && in.C % 8 == 0 && in.cstride % 8 == 0 && in.cstride >= in.C);
    p.kchunks[s] = (in.C + BK - 1) / BK;
    for (int t = 0; t < in.ntaps; ++t) {
      p.dy[s][t] = in.dy[t];
      p.dx[s][t] = in.dx[t];
      p.wrow[s][t] = in.wrow[t];
      if (in.wrow[t] + d->cout_pad > wrows) wrows = in.wrow[t] + d->cout_pad;
    }
    int rc = make_act_map(&mapA[s], in.ptr, d->N, in.H, in.W, in.C, in.cstride, TW, TH, TN, d->stride);
    if (rc != PS_OK) return rc;
  }
  PS_CHECK_ARG(d->w_rows >= wrows && d->w_cin_pad % BK == 0);
  PS_CHECK_ARG(d->w_cin_pad >= p.kchunks[0] * BK && (p.ntaps[1] == 0 || d->w_cin_pad >= p.kchunks[1] * BK));
  if (p.ntaps[1] == 0) mapA[1] = mapA[0];
  int rc = make_w_map(&mapW, d->weights, d->w_rows, d->w_cin_pad, BN);
  if (rc != PS_OK) return rc;
  p.Cout = d->Cout;
  p.bias = d->bias;
  p.residual = (const __nv_bfloat16*)d->residual;
  p.res_cstride = d->res_cstride;
  for (int o = 0; o < 2; ++o) {
    p.out[o] = (__nv_bfloat16*)d->out[o].ptr;
    p.scale[o] = d->out[o].scale;
    p.shift[o] = d->out[o].shift;
    p.per_sample[o] = d->out[o].per_sample;
    p.act[o] = d->out[o].act;
    p.out_cstride[o] = d->out[o].cstride;
    p.out_coffset[o] = d->out[o].coffset;
  }
  p.out_f32_nchw = d->out_f32_nchw;
  p.act_param[0] = d->act_param[0];
  p.act_param[1] = d->act_param[1];
  p.out_H = d->out_H ? d->out_H : d->Hout;
  p.out_W = d->out_W ? d->out_W : d->Wout;
  p.out_sy = d->out_sy ? d->out_sy : 1;
  p.out_sx = d->out_sx ? d->out_sx : 1;
  p.out_py = d->out_py;
  p.out_px = d->out_px;

  const size_t smem_bytes = 1024 + (size_t)stages * stage_bytes + (2 * stages + 1) * sizeof(uint64_t) + 16;
  static thread_local int attr_dev = -1;
  int dev = 0;
  PS_CUDA(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    PS_CUDA(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_dev = dev;
  }
  dim3 grid(p.tiles_x * p.tiles_y * tiles_n, d->cout_pad / BN);
  PS_TIME_BEGIN("conv_igemm_kernel", (cudaStream_t)stream);
  conv_igemm_kernel<<<grid, CONV_THREADS, smem_bytes, (cudaStream_t)stream>>>(mapA[0], mapA[1], mapW, p);
  PS_TIME_END((cudaStream_t)stream);
  PS_LAUNCHED();
  return PS_OK;
}
