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
// Persistent kernel, one CTA per SM, each walking tiles blockIdx.x, blockIdx.x + gridDim.x, ...  Warp roles (192
// threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue (two warps per
// 32-lane TMEM quadrant, each taking every other 32-column block: tcgen05.ld, bias / residual / per-sample
// scale+shift / activation, up to two NHWC outputs; the epilogue is a chain of dependent loads, so it wants warps).  The smem ring of `stages` stages (full/empty mbarriers) runs on across tiles, and the
// accumulator is double buffered in TMEM (acc_full / acc_empty mbarriers), so the epilogue of tile i overlaps the
// main loop of tile i + 1.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc05.cuh"

namespace ps {

constexpr int CONV_THREADS = 320;  // producer, MMA issuer, 8 epilogue warps
constexpr int BM = 128;  // output pixels per tile (UMMA M)
constexpr int BK = 64;   // channels per K step (128 bytes of bf16 = one swizzle row)
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int CONV_MAX_STAGES = 8;
// halo mode: tile = 8 x 16 output pixels; the halo tile is 18 rows of 10 pixels x 64 channels.  The 128-byte swizzle
// is a function of the shared-memory address for TMA and MMA alike, so the 8-pixel row groups of a shifted window may
// start on any 128-byte row.
constexpr int HALO_W = 10, HALO_H = 18;
constexpr int HALO_LOAD_BYTES = HALO_W * HALO_H * BK * 2;         // 23040 bytes per TMA box
constexpr int HALO_BYTES = (HALO_LOAD_BYTES + 1023) / 1024 * 1024;  // buffer stride

struct ConvKernelParams {
  // tile geometry: TW * TH * TN == 128
  int TW, TH, TN;
  int tiles_x, tiles_y;
  int tiles_spatial, total_tiles;  // tile index = column block * tiles_spatial + spatial tile
  int Hout, Wout, N;  // output grid this launch computes (per phase for transposed convs)
  int stride;         // input pixel = out * stride + tap offset
  int ntaps[2], kchunks[2];
  int dy[2][16], dx[2][16];
  int wrow[2][16];  // first weight row (of the packed [rows][Cin_pad] matrix) of each tap
  int BN;           // output channels per CTA (multiple of 16, <= 256)
  int stages;
  int halo;         // 3x3 stride-1 mode: every tap is read in place from one (TH+2) x (TW+2)-pixel halo tile per 64-channel chunk
  int a_stage_bytes;  // A part of a ring stage (0 in halo mode: no k-iteration stages an A tile there)
  // A CTA works on GROUPS of tpg (1 or 2) spatially adjacent tiles of one column block: a weight tile that has been
  // brought into shared memory is multiplied against every tile of the group before it is released, so with tpg = 2 the
  // weights -- 9 x 16 KB per 64 input channels, re-read for every tile, the kernel's largest L2 stream -- are read half as often.
  int tpg, groups_per_cb, total_groups;
  int nbuf;         // TMEM accumulator buffers (2 * tpg): the epilogue of one group overlaps the main loop of the next
  int nhalo;        // halo buffers (2 * tpg)
  // Narrow layers (Cout <= 16, e.g. the 128 -> 3 and 3 -> 3 convolutions at 256 x 256): a tap's weight tile is 2 KB and
  // a tile's main loop is 18 of them, each a TMA round trip -- the layer was bound by TMA latency, not by any roofline.
  // wgroup = taps of input 0 whose weight tiles one TMA brings into one ring stage (ntaps[0] there, 1 elsewhere).
  int wgroup;
  int debug;
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

// y[0..32) = act(f * scale + shift) -> 64 bytes of bf16 at dst (16-byte aligned); the activation is a compile-time
// constant here so the 32-wide loop carries no switch
// Warp-private staging tile: 32 rows x 64 bytes (one 32-channel bf16 block per output pixel).  A lane owns a row when
// it computes, and a 16-byte chunk of every eighth row when it talks to global memory (four lanes cover a row's
// 64 contiguous bytes, so a warp instruction touches 8 lines instead of 32).  Chunk c of row r lives at chunk
// c ^ ((r >> 1) & 3): both access patterns are then bank-conflict free.
__device__ __forceinline__ uint32_t stage_addr(uint32_t base, int r, int c) { return base + r * 64 + ((c ^ ((r >> 1) & 3)) << 4); }
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// y[0..32) = act(f * scale + shift) as bf16 into row `lane` of the staging tile; the activation is a compile-time
// constant here so the 32-wide loop carries no switch
template <int ACT>
__device__ __forceinline__ void stage_block32(const float* f, const float4* sc, const float4* sh, const float* ap,
                                              uint32_t stage, int lane) {
  float y[32];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 a = sc[j], b = sh[j];
    y[4 * j] = apply_act(fmaf(f[4 * j], a.x, b.x), ACT, ap);
    y[4 * j + 1] = apply_act(fmaf(f[4 * j + 1], a.y, b.y), ACT, ap);
    y[4 * j + 2] = apply_act(fmaf(f[4 * j + 2], a.z, b.z), ACT, ap);
    y[4 * j + 3] = apply_act(fmaf(f[4 * j + 3], a.w, b.w), ACT, ap);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(y[8 * j], y[8 * j + 1]), h1 = __floats2bfloat162_rn(y[8 * j + 2], y[8 * j + 3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(y[8 * j + 4], y[8 * j + 5]), h3 = __floats2bfloat162_rn(y[8 * j + 6], y[8 * j + 7]);
    uint4 w;
    w.x = *reinterpret_cast<uint32_t*>(&h0);
    w.y = *reinterpret_cast<uint32_t*>(&h1);
    w.z = *reinterpret_cast<uint32_t*>(&h2);
    w.w = *reinterpret_cast<uint32_t*>(&h3);
    sts128(stage_addr(stage, lane, j), w);
  }
}

__global__ void __launch_bounds__(CONV_THREADS, 1)
    conv_igemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                      const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapH,
                      const __grid_constant__ CUtensorMap mapH1, const __grid_constant__ CUtensorMap mapWG,
                      const ConvKernelParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform to the compiler as well
  const int BN = p.BN;
  const int BNp = (BN + 31) & ~31;  // TMEM columns per accumulator buffer
  const int stage_bytes = p.a_stage_bytes + p.wgroup * BN * BK * 2;
  // carve: [nhalo halo tiles (halo mode)] [stages x (A | B)] then barriers
  unsigned char* halo_tiles = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  unsigned char* tiles = halo_tiles + (p.halo ? p.nhalo * HALO_BYTES : 0);
  uint64_t* full_bar = (uint64_t*)(tiles + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* acc_full = empty_bar + p.stages;
  uint64_t* acc_empty = acc_full + 4;
  uint64_t* halo_full = acc_empty + 4;
  uint64_t* halo_empty = halo_full + 4;
  uint32_t* tmem_slot = (uint32_t*)(halo_empty + 4);
  unsigned char* param_smem = (unsigned char*)(tmem_slot + 4);  // 16-byte aligned: 8 warps x 2 buffers x 40 float4
  unsigned char* stage_smem = param_smem + 8 * 2 * 40 * 16;      // 8 warps x 2 staging tiles of 2 KB
  const int kiters = p.ntaps[0] * p.kchunks[0] + p.ntaps[1] * p.kchunks[1];
  const int acc_cols = p.nbuf * BNp;
  const uint32_t ncols = acc_cols <= 32 ? 32u : (acc_cols <= 64 ? 64u : (acc_cols <= 128 ? 128u : (acc_cols <= 256 ? 256u : 512u)));
  // nbuf = nhalo = 2 * tpg is 2 or 4: masks and shifts (a division by a run-time value leaves the uniform datapath,
  // and everything derived from it would reach the MMA instructions through per-use R2UR moves)
  const int tpg = p.tpg, bshift = tpg, bmask = 2 * tpg - 1;
  // group gi -> column block, first spatial tile and number of tiles (the last group of a column block may hold one)
  auto group_cb = [&](int gi) { return gi / p.groups_per_cb; };
  auto group_sp0 = [&](int gi) { return (gi % p.groups_per_cb) * tpg; };
  auto group_nt = [&](int gi) { return min(tpg, p.tiles_spatial - group_sp0(gi)); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapW);
    if (p.ntaps[1]) tma_prefetch_desc(&mapA1);
    if (p.halo) tma_prefetch_desc(&mapH);
    if (p.halo && p.ntaps[1]) tma_prefetch_desc(&mapH1);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < 4; ++b) {
        mbar_init(&acc_full[b], 1);
        mbar_init(&acc_empty[b], 8);  // one arrival per epilogue warp
        mbar_init(&halo_full[b], 1);
        mbar_init(&halo_empty[b], 1);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, ncols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: the whole warp walks the loop (uniform control flow, incremental ring state); one elected
    // lane arms the barrier and issues the copies.  Lane t keeps tap t's parameters; a shuffle fetches them. =====
    const int tl = lane & 15;
    const int my_dx0 = p.dx[0][tl], my_dy0 = p.dy[0][tl], my_w0 = p.wrow[0][tl];
    const int my_dx1 = p.dx[1][tl], my_dy1 = p.dy[1][tl], my_w1 = p.wrow[1][tl];
    const uint32_t w_bytes = (uint32_t)(BN * BK * 2);
    int s = 0, hit = 0;
    uint32_t ph = 1;  // parity of the `empty` phase to wait for (passes on a fresh barrier)
    for (int gi = blockIdx.x; gi < p.total_groups; gi += gridDim.x) {
      const int cb = group_cb(gi), sp0 = group_sp0(gi), nt = group_nt(gi), ncol0 = cb * BN;
      for (int src = 0; src < 2; ++src) {
        if (p.ntaps[src] == 0) continue;
        if (p.halo) {
          // per 64-channel chunk: one halo tile per tile of the group, then the weight tiles that are multiplied against them
          const CUtensorMap* mH = src ? &mapH1 : &mapH;
          for (int kc = 0; kc < p.kchunks[src]; ++kc) {
            for (int j = 0; j < nt; ++j, ++hit) {
              const int sp = sp0 + j;
              const int tx = sp % p.tiles_x, ty = (sp / p.tiles_x) % p.tiles_y, tn = sp / (p.tiles_x * p.tiles_y);
              const int hb = hit & bmask;
              mbar_wait(&halo_empty[hb], (((uint32_t)(hit >> bshift)) & 1u) ^ 1u);
              if (elect_one()) {
                if (p.debug & 8) {
                  mbar_arrive(&halo_full[hb]);
                } else {
                  mbar_expect_tx(&halo_full[hb], (uint32_t)HALO_LOAD_BYTES);
                  tma_load_4d(mH, &halo_full[hb], halo_tiles + (size_t)hb * HALO_BYTES, kc * BK, tx * p.TW - 1, ty * p.TH - 1,
                              tn * p.TN);
                }
              }
              __syncwarp();
            }
            const int wg = src ? 1 : p.wgroup;  // taps per weight stage
            for (int t = 0; t < p.ntaps[src]; t += wg) {
              const int wr = __shfl_sync(0xffffffffu, src ? my_w1 : my_w0, t);
              mbar_wait(&empty_bar[s], ph);
              if (elect_one()) {
                if (p.debug & 8) {
                  mbar_arrive(&full_bar[s]);
                } else {
                  mbar_expect_tx(&full_bar[s], w_bytes * (uint32_t)wg);
                  tma_load_2d(wg > 1 ? &mapWG : &mapW, &full_bar[s], tiles + (size_t)s * stage_bytes + p.a_stage_bytes, kc * BK,
                              wr + ncol0);
                }
              }
              __syncwarp();
              if (++s == p.stages) {
                s = 0;
                ph ^= 1u;
              }
            }
          }
          continue;
        }
        // tap mode (one tile per group): the tap's shifted window is fetched into the stage next to its weights
        const CUtensorMap* mA = src ? &mapA1 : &mapA0;
        const int tx = sp0 % p.tiles_x, ty = (sp0 / p.tiles_x) % p.tiles_y, tn = sp0 / (p.tiles_x * p.tiles_y);
        const int ox0 = tx * p.TW, oy0 = ty * p.TH, n0 = tn * p.TN;
        for (int t = 0; t < p.ntaps[src]; ++t) {
          const int dx = __shfl_sync(0xffffffffu, src ? my_dx1 : my_dx0, t);
          const int dy = __shfl_sync(0xffffffffu, src ? my_dy1 : my_dy0, t);
          const int wr = __shfl_sync(0xffffffffu, src ? my_w1 : my_w0, t);
          const int ix0 = ox0 * p.stride + dx;
          const int iy0 = oy0 * p.stride + dy;
          for (int kc = 0; kc < p.kchunks[src]; ++kc) {
            mbar_wait(&empty_bar[s], ph);
            if (elect_one()) {
              unsigned char* a = tiles + (size_t)s * stage_bytes;
              if (p.debug & 8) {
                mbar_arrive(&full_bar[s]);
              } else {
                mbar_expect_tx(&full_bar[s], (uint32_t)A_STAGE_BYTES + w_bytes);
                tma_load_4d(mA, &full_bar[s], a, kc * BK, ix0, iy0, n0);
                tma_load_2d(&mapW, &full_bar[s], a + p.a_stage_bytes, kc * BK, wr + ncol0);
              }
            }
            __syncwarp();
            if (++s == p.stages) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: whole warp, warp-uniform operands, one elected lane issues (tc05.cuh: umma_f16_kblock) =====
    const uint32_t idesc = umma_idesc_bf16(BN);
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(tiles));
    // lane t: tap t's start row in the halo tile, for either input
    const int my_r0 = (1 + p.dy[0][lane & 15]) * HALO_W + 1 + p.dx[0][lane & 15];
    const int my_r1 = (1 + p.dy[1][lane & 15]) * HALO_W + 1 + p.dx[1][lane & 15];
    int s = 0, lt = 0, hit = 0;
    uint32_t ph = 0;
    for (int gi = blockIdx.x; gi < p.total_groups; gi += gridDim.x) {
      const int nt = group_nt(gi);
      // scalars, not arrays: the MMA operands must stay warp-uniform registers (tc05.cuh, umma_f16_kblock)
      const int buf0 = lt & bmask, buf1 = (lt + 1) & bmask;
      mbar_wait(&acc_empty[buf0], (((uint32_t)(lt >> bshift)) & 1u) ^ 1u);  // the epilogue has drained this buffer
      if (nt == 2) mbar_wait(&acc_empty[buf1], (((uint32_t)((lt + 1) >> bshift)) & 1u) ^ 1u);
      // (the shuffles tell the compiler these are warp-uniform: they then live in uniform registers)
      const uint32_t d0 = __shfl_sync(0xffffffffu, tmem_base + (uint32_t)(buf0 * BNp), 0);
      const uint32_t d1 = __shfl_sync(0xffffffffu, tmem_base + (uint32_t)(buf1 * BNp), 0);
      const uint32_t d_last = nt == 2 ? d1 : d0;
      tc_fence_after();
      int it = 0;
      if (p.halo) {
        for (int src = 0; src < 2; ++src) {
          for (int kc = 0; kc < (p.ntaps[src] ? p.kchunks[src] : 0); ++kc) {
            const int hb0 = hit & bmask, hb1 = (hit + 1) & bmask;
            mbar_wait(&halo_full[hb0], ((uint32_t)(hit >> bshift)) & 1u);
            if (nt == 2) mbar_wait(&halo_full[hb1], ((uint32_t)((hit + 1) >> bshift)) & 1u);
            hit += nt;
            const uint32_t h0 = __shfl_sync(0xffffffffu, smem_u32(halo_tiles) + (uint32_t)(hb0 * HALO_BYTES), 0);
            const uint32_t h1 = __shfl_sync(0xffffffffu, smem_u32(halo_tiles) + (uint32_t)(hb1 * HALO_BYTES), 0);
            const uint32_t h_last = nt == 2 ? h1 : h0;
            tc_fence_after();
            const int wg = src ? 1 : p.wgroup;
            for (int t = 0; t < p.ntaps[src]; ++t, ++it) {
              const bool opens = wg == 1 || t == 0, releases = wg == 1 || t == p.ntaps[src] - 1;
              if (opens) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
              }
              // window shifted by the tap: starts (1+dy) halo rows and (1+dx) pixels in; 8-pixel row groups are one
              // halo row (10 pixels = 1280 bytes) apart.  The 128-byte swizzle is a function of the shared-memory
              // address bits, for TMA's writes and the MMA's reads alike, so a start row that is not a multiple of 8
              // needs nothing else (descriptor base offset 0; measured: a non-zero base offset reads the wrong chunks).
              const uint32_t r0 = (uint32_t)__shfl_sync(0xffffffffu, src ? my_r1 : my_r0, t);
              const uint32_t a_hi = (uint32_t)((HALO_W * 128) >> 4) | (1u << 14) | (2u << 29);
              const uint32_t b_lo = a_lo0 + (uint32_t)s * (uint32_t)(stage_bytes >> 4) + (uint32_t)(p.a_stage_bytes >> 4) +
                                    (wg > 1 ? (uint32_t)(t * BN * 8) : 0u);  // tap t's tile inside a grouped stage
              if (p.debug & 4) {
                if (releases) umma_commit_elect(&empty_bar[s]);
              } else {
                // the same weight tile against every tile of the group; the stage is released after its last use
                if (nt == 2) umma_f16_kblock_ahi_nc(d0, umma_desc_lo(h0 + r0 * 128u), a_hi, b_lo, idesc, it ? 1u : 0u);
                if (releases) umma_f16_kblock_ahi(d_last, umma_desc_lo(h_last + r0 * 128u), a_hi, b_lo, idesc, it ? 1u : 0u, &empty_bar[s]);
                else umma_f16_kblock_ahi_nc(d_last, umma_desc_lo(h_last + r0 * 128u), a_hi, b_lo, idesc, it ? 1u : 0u);
              }
              if (releases && ++s == p.stages) {
                s = 0;
                ph ^= 1u;
              }
            }
            // the halo tiles are free once their taps have been multiplied
            umma_commit_elect(&halo_empty[hb0]);
            if (nt == 2) umma_commit_elect(&halo_empty[hb1]);
          }
        }
      } else {
        for (; it < kiters; ++it) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)s * (uint32_t)(stage_bytes >> 4);
          if (p.debug & 4) umma_commit_elect(&empty_bar[s]);
          else umma_f16_kblock(d0, a_lo, a_lo + (uint32_t)(p.a_stage_bytes >> 4), idesc, it ? 1u : 0u, &empty_bar[s]);
          if (++s == p.stages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
      umma_commit_elect(&acc_full[buf0]);  // accumulators complete
      if (nt == 2) umma_commit_elect(&acc_full[buf1]);
      lt += nt;
    }
  } else {
    // ===== epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 and the 32-column blocks of parity (w-2)/4 =====
    // Work items = (tile, 32-column block).  The per-channel parameters of an item (bias, scale/shift of both outputs)
    // are the same for every row, so the warp fetches them once -- lane l loads one float4 -- a whole item ahead,
    // parks them in a warp-private shared-memory buffer and reads them back as broadcasts; the residual rows of the
    // next item are requested a whole item ahead too.  Nothing on the item's critical path waits for L2.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int m = q * 32 + lane;  // tile row = output pixel
    const int nl = m / (p.TH * p.TW);
    const int rem = m - nl * (p.TH * p.TW);
    const int my = rem / p.TW, mx = rem % p.TW;
    float4* pbuf = reinterpret_cast<float4*>(param_smem) + (warp - 2) * 2 * 40;
    const bool shared_params = p.TN == 1 || !(p.per_sample[0] || p.per_sample[1]);

    struct Geom {
      bool valid;
      int n, fy, fx, ncol0;
      size_t pixel;
    };
    auto geom = [&](int tile) {
      Geom g;
      const int sp = tile % p.tiles_spatial, cb = tile / p.tiles_spatial;
      const int tx = sp % p.tiles_x, ty = (sp / p.tiles_x) % p.tiles_y, tn = sp / (p.tiles_x * p.tiles_y);
      const int oy = ty * p.TH + my, ox = tx * p.TW + mx;
      g.n = tn * p.TN + nl;
      g.ncol0 = cb * BN;
      g.valid = g.n < p.N && oy < p.Hout && ox < p.Wout;
      g.fy = oy * p.out_sy + p.out_py;
      g.fx = ox * p.out_sx + p.out_px;
      g.pixel = ((size_t)g.n * p.out_H + g.fy) * p.out_W + g.fx;
      return g;
    };
    // float4 number idx (0..39) of an item's parameters: [bias | scale0 | shift0 | scale1 | shift1] x 8
    auto load_param = [&](int idx, int n_tile, int cbase) {
      const int k = idx >> 3, j = (idx & 7) * 4;
      const float* src = nullptr;
      float dflt = 0.f;
      if (k == 0) {
        src = p.bias;
      } else {
        const int o = (k - 1) >> 1;
        src = ((k - 1) & 1) ? p.shift[o] : p.scale[o];
        dflt = ((k - 1) & 1) ? 0.f : 1.f;
        if (src && p.per_sample[o]) src += (size_t)n_tile * p.Cout;
      }
      if (!src) return make_float4(dflt, dflt, dflt, dflt);
      return __ldg(reinterpret_cast<const float4*>(src + cbase + j));
    };
    auto block_full = [&](int cbase) { return cbase + 32 <= p.Cout && (p.Cout & 3) == 0; };
    auto prefetch_params = [&](int tile, int c0, float4& a, float4& b) {
      const int sp = tile % p.tiles_spatial, cb = tile / p.tiles_spatial;
      const int n_tile = (sp / (p.tiles_x * p.tiles_y)) * p.TN;
      const int cbase = cb * BN + c0;
      a = b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (shared_params && block_full(cbase)) {
        a = load_param(lane, n_tile, cbase);
        if (lane < 8) b = load_param(lane + 32, n_tile, cbase);
      }
    };
    // lane l talks to global memory for chunk l & 3 of rows 8i + (l >> 2), i = 0..3; the rows' pixel offsets and
    // validity come from the lanes that own them
    const int trow = lane >> 2, tchunk = lane & 3;
    auto prefetch_residual = [&](const Geom& g, int c0, uint4* rv) {
      const int cbase = g.ncol0 + c0;
      if (!p.residual || !block_full(cbase) || ((p.res_cstride | cbase) & 7)) return;  // warp-uniform
      const unsigned vmask = __ballot_sync(0xffffffffu, g.valid);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 8 * i + trow;
        const unsigned long long px = __shfl_sync(0xffffffffu, (unsigned long long)g.pixel, r);
        rv[i] = make_uint4(0, 0, 0, 0);
        if ((vmask >> r) & 1u)
          rv[i] = __ldg(reinterpret_cast<const uint4*>(p.residual + px * p.res_cstride + cbase) + tchunk);
      }
    };
    // the staging tile -> global rows, coalesced
    auto write_out = [&](uint32_t stage, __nv_bfloat16* base, int cstride, const Geom& g) {
      const unsigned vmask = __ballot_sync(0xffffffffu, g.valid);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 8 * i + trow;
        const unsigned long long px = __shfl_sync(0xffffffffu, (unsigned long long)g.pixel, r);
        const uint4 w = lds128(stage_addr(stage, r, tchunk));
        if ((vmask >> r) & 1u) *(reinterpret_cast<uint4*>(base + px * cstride) + tchunk) = w;
      }
    };
    const uint32_t stage0 = smem_u32(stage_smem) + (uint32_t)(warp - 2) * 4096u, stage1 = stage0 + 2048u;

    // this CTA's tiles in the order the MMA warp fills the accumulator buffers: group by group, tile by tile
    auto tile_of = [&](int gi, int gj) {
      return gi < p.total_groups ? group_cb(gi) * p.tiles_spatial + group_sp0(gi) + gj : p.total_tiles;
    };
    auto advance = [&](int& gi, int& gj) {
      if (gj + 1 < group_nt(gi)) {
        ++gj;
      } else {
        gi += (int)gridDim.x;
        gj = 0;
      }
    };
    int gi = blockIdx.x, gj = 0;
    int tile = tile_of(gi, gj), c0 = half * 32, lt = 0, item = 0;
    const bool any = tile < p.total_tiles && c0 < BN;
    float4 pf0, pf1;
    uint4 rv[4];
    Geom g;
    if (any) {
      g = geom(tile);
      prefetch_params(tile, c0, pf0, pf1);
      prefetch_residual(g, c0, rv);
    }
    while (any && tile < p.total_tiles) {
      // park this item's parameters, then put the next item's requests in flight
      float4* pb = pbuf + (item & 1) * 40;
      pb[lane] = pf0;
      if (lane < 8) pb[32 + lane] = pf1;
      __syncwarp();
      uint4 rcur[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) rcur[j] = rv[j];
      const Geom gc = g;
      const int c0c = c0, buf = lt & bmask;
      const bool first = c0 == half * 32;
      int ntile = tile, nc0 = c0 + 64, ngi = gi, ngj = gj;
      if (nc0 >= BN) {
        advance(ngi, ngj);
        ntile = tile_of(ngi, ngj);
        nc0 = half * 32;
      }
      const bool last = ntile != tile;
      if (ntile < p.total_tiles) {
        if (last) g = geom(ntile);
        prefetch_params(ntile, nc0, pf0, pf1);
        prefetch_residual(g, nc0, rv);
      }
      if (first) {
        mbar_wait(&acc_full[buf], ((uint32_t)(lt >> bshift)) & 1u);
        tc_fence_after();
      }
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BNp);
      const int cbase = gc.ncol0 + c0c;
      const bool full = block_full(cbase);
      const bool block_on = cbase < p.Cout && !(p.debug & 16);                 // warp-uniform
      // coalesced path: whole 32-channel block, warp-shared parameters, 16-byte aligned rows in every tensor
      bool fast = block_on && full && shared_params && (!p.residual || ((p.res_cstride | cbase) & 7) == 0);
#pragma unroll
      for (int o = 0; o < 2; ++o)
        if (p.out[o]) fast = fast && ((p.out_cstride[o] | (p.out_coffset[o] + cbase)) & 7) == 0 && (((uintptr_t)p.out[o]) & 15) == 0;
      const bool live = gc.valid && block_on;
      uint32_t v[32];
      if (p.debug & 64) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0;
      } else {
        tmem_ld32(tbase + (uint32_t)c0c, v);
      }
      if ((fast || live) && !(p.debug & 128)) {
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        const int n = gc.n;
        const size_t pixel = gc.pixel;
        if (fast) {
          // ---- fast path (all lanes, rows outside the image are masked at the stores) ----
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = pb[j];
            f[4 * j] += b4.x;
            f[4 * j + 1] += b4.y;
            f[4 * j + 2] += b4.z;
            f[4 * j + 3] += b4.w;
          }
          if (p.residual) {
            // the residual block arrived in transposed ownership: through the staging tile back to row ownership
#pragma unroll
            for (int i = 0; i < 4; ++i) sts128(stage_addr(stage0, 8 * i + trow, tchunk), rcur[i]);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 rr = lds128(stage_addr(stage0, lane, j));
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 t2 = __bfloat1622float2(h[k]);
                f[8 * j + 2 * k] += t2.x;
                f[8 * j + 2 * k + 1] += t2.y;
              }
            }
            __syncwarp();
          }
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            if (!p.out[o]) continue;
            const float4* sc4 = pb + 8 + 16 * o;
            const float4* sh4 = pb + 16 + 16 * o;
            const uint32_t st = o ? stage1 : stage0;
            switch (p.act[o]) {
              case ACT_RELU: stage_block32<ACT_RELU>(f, sc4, sh4, p.act_param, st, lane); break;
              case ACT_LEAKY: stage_block32<ACT_LEAKY>(f, sc4, sh4, p.act_param, st, lane); break;
              case ACT_NONE: stage_block32<ACT_NONE>(f, sc4, sh4, p.act_param, st, lane); break;
              case ACT_TANH: stage_block32<ACT_TANH>(f, sc4, sh4, p.act_param, st, lane); break;
              case ACT_ELU: stage_block32<ACT_ELU>(f, sc4, sh4, p.act_param, st, lane); break;
              default: stage_block32<ACT_SIGMOID_AFFINE>(f, sc4, sh4, p.act_param, st, lane); break;
            }
          }
          __syncwarp();
#pragma unroll
          for (int o = 0; o < 2; ++o)
            if (p.out[o] && !(p.debug & 32)) write_out(o ? stage1 : stage0, p.out[o] + p.out_coffset[o] + cbase, p.out_cstride[o], gc);
          __syncwarp();  // the staging tiles are rewritten by the next item
          if (p.out_f32_nchw && gc.valid) {
            const float* sc = reinterpret_cast<const float*>(pb + 8);
            const float* sh = reinterpret_cast<const float*>(pb + 16);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              p.out_f32_nchw[(((size_t)n * p.Cout + cbase + j) * p.out_H + gc.fy) * p.out_W + gc.fx] =
                  apply_act(fmaf(f[j], sc[j], sh[j]), p.act[0], p.act_param);
          }
        } else {
          // ---- generic path: ragged channel counts, per-sample parameters in a tile that spans images.  Every loop
          // is fully unrolled with constant indices: one dynamic index would move f[] to local memory for the
          // fast path as well. ----
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = cbase + j;
            if (c >= p.Cout) break;  // unrolled: the indices stay compile-time constants, the tail is skipped
            if (p.bias) f[j] += __ldg(p.bias + c);
            if (p.residual) f[j] += __bfloat162float(p.residual[pixel * p.res_cstride + c]);
          }
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            const bool f32_only = o == 0 && !p.out[0] && p.out_f32_nchw;
            if (!p.out[o] && !f32_only) continue;
            const float* sc = p.scale[o] ? p.scale[o] + (p.per_sample[o] ? (size_t)n * p.Cout : 0) : nullptr;
            const float* sh = p.shift[o] ? p.shift[o] + (p.per_sample[o] ? (size_t)n * p.Cout : 0) : nullptr;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int c = cbase + j;
              if (c >= p.Cout) break;
              float t = f[j];
              if (sc) t *= __ldg(sc + c);
              if (sh) t += __ldg(sh + c);
              const float y = apply_act(t, p.act[o], p.act_param);
              if (p.out[o]) p.out[o][pixel * p.out_cstride[o] + p.out_coffset[o] + c] = __float2bfloat16(y);
              if (o == 0 && p.out_f32_nchw)
                p.out_f32_nchw[(((size_t)n * p.Cout + c) * p.out_H + gc.fy) * p.out_W + gc.fx] = y;
            }
          }
        }
      }
      if (last) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        ++lt;
      }
      tile = ntile;
      gi = ngi;
      gj = ngj;
      c0 = nc0;
      ++item;
    }
    // a warp whose column parity has no block in this launch (BN <= 32) still releases every accumulator
    if (!any || half * 32 >= BN) {
      int lt2 = 0;
      for (int g2 = blockIdx.x; g2 < p.total_groups; g2 += gridDim.x)
        for (int j2 = 0; j2 < group_nt(g2); ++j2, ++lt2) {
          mbar_wait(&acc_full[lt2 & bmask], ((uint32_t)(lt2 >> bshift)) & 1u);
          if (lane == 0) mbar_arrive(&acc_empty[lt2 & bmask]);
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
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
  if (int rcw = wedge_check(__func__)) return rcw;
  PS_WEDGE_ARM();
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
  // halo mode: stride-1 convolution whose first input's taps all lie in the 3x3 window, on images that fill the
  // 8 x 16 tile.  The window of every tap is then read in place from one halo tile per 64-channel chunk instead of
  // being fetched once per tap (input traffic / 9 * 2.25).
  const int dbg = getenv("PS_CONV_DEBUG") ? atoi(getenv("PS_CONV_DEBUG")) : 0;
  bool halo = d->stride == 1 && d->Wout >= 8 && d->Hout >= 16 && !(dbg & 2);
  for (int t = 0; t < d->in[0].ntaps; ++t)
    halo = halo && d->in[0].dy[t] >= -1 && d->in[0].dy[t] <= 1 && d->in[0].dx[t] >= -1 && d->in[0].dx[t] <= 1;
  halo = halo && d->in[0].ntaps > 1;
  // the second input (the decoder's fused 1x1 skip) goes through halo tiles too: its taps must lie in the window as well
  for (int t = 0; t < d->in[1].ntaps; ++t)
    halo = halo && d->in[1].dy[t] >= -1 && d->in[1].dy[t] <= 1 && d->in[1].dx[t] >= -1 && d->in[1].dx[t] <= 1;
  if (halo) {
    TW = 8;
    TH = 16;
    TN = 1;
  }
  p.halo = halo ? 1 : 0;
  p.debug = dbg;
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
  const int sms = stream_sms((cudaStream_t)stream);  // the whole device, or the stream's green-context partition
  // few pixels, many channels (the U-Net's bottleneck: 1x1 .. 4x4 images): a handful of CTAs would each stream all the
  // weights of a wide column block through one SM; narrower column blocks spread that read over more SMs
  if (!(dbg & 512))
    while (BN >= 64 && BN % 64 == 0 && d->cout_pad % (BN / 2) == 0 &&
           p.tiles_x * p.tiles_y * tiles_n * (d->cout_pad / BN) * 2 <= sms)
      BN /= 2;
  PS_CHECK_ARG(d->cout_pad % BN == 0);
  p.BN = BN;
  // a ring stage carries an A tile only in tap mode (in halo mode every tap of both inputs is read from a halo tile)
  p.a_stage_bytes = halo ? 0 : A_STAGE_BYTES;
  // narrow layers: all taps' weight tiles of a 64-channel chunk in one TMA / one stage (rows must be tap-contiguous)
  p.wgroup = 1;
  if (halo && BN == d->cout_pad && d->in[0].ntaps * BN <= 256 && !(dbg & 256)) {
    bool contiguous = true;
    for (int t = 0; t < d->in[0].ntaps; ++t) contiguous = contiguous && d->in[0].wrow[t] == d->in[0].wrow[0] + t * BN;
    if (contiguous) p.wgroup = d->in[0].ntaps;
  }
  const int stage_bytes = p.a_stage_bytes + p.wgroup * BN * BK * 2;
  // groups of two tiles share every weight tile: needs four accumulator buffers (4 x BN <= 512 TMEM columns); it
  // only pays when there are more tiles than SMs (otherwise it would just idle half of them)
  static thread_local int attr_dev = -1;
  int dev = 0;
  PS_CUDA(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    PS_CUDA(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_dev = dev;
  }
  p.tiles_spatial = p.tiles_x * p.tiles_y * tiles_n;
  p.total_tiles = p.tiles_spatial * (d->cout_pad / BN);
  p.tpg = (halo && BN <= 128 && !(dbg & 1) && p.total_tiles >= 2 * sms) ? 2 : 1;
  p.nbuf = 2 * p.tpg;
  p.nhalo = 2 * p.tpg;
  // one persistent CTA per SM: as many ring stages as ~200 KB of shared memory hold
  int stages = (182 * 1024 - (halo ? p.nhalo * HALO_BYTES : 0)) / stage_bytes;  // 227 KB - barriers, parameter and staging tiles
  stages = stages > CONV_MAX_STAGES ? CONV_MAX_STAGES : (stages < 2 ? 2 : stages);
  p.stages = stages;
  CUtensorMap mapA[2], mapW, mapH, mapH1, mapWG;
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
  mapWG = mapW;
  if (p.wgroup > 1) {
    rc = make_w_map(&mapWG, d->weights, d->w_rows, d->w_cin_pad, p.wgroup * BN);
    if (rc != PS_OK) return rc;
  }
  mapH = mapH1 = mapA[0];
  if (halo) {
    const ps_conv_input& in = d->in[0];
    rc = make_act_map(&mapH, in.ptr, d->N, in.H, in.W, in.C, in.cstride, HALO_W, HALO_H, 1, 1);
    if (rc != PS_OK) return rc;
    if (d->in[1].ntaps > 0) {
      const ps_conv_input& in1 = d->in[1];
      rc = make_act_map(&mapH1, in1.ptr, d->N, in1.H, in1.W, in1.C, in1.cstride, HALO_W, HALO_H, 1, 1);
      if (rc != PS_OK) return rc;
    }
  }
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

  const size_t smem_bytes = 1024 + (halo ? p.nhalo * HALO_BYTES : 0) + (size_t)stages * stage_bytes +
                            (2 * stages + 16) * sizeof(uint64_t) + 16 + 8 * 2 * 40 * 16 + 8 * 4096;
  p.groups_per_cb = (p.tiles_spatial + p.tpg - 1) / p.tpg;
  p.total_groups = p.groups_per_cb * (d->cout_pad / BN);
  const int grid = p.total_groups < sms ? p.total_groups : sms;
  PS_TIME_BEGIN("conv_igemm_kernel", (cudaStream_t)stream);
  conv_igemm_kernel<<<grid, CONV_THREADS, smem_bytes, (cudaStream_t)stream>>>(mapA[0], mapA[1], mapW, mapH, mapH1, mapWG, p);
  PS_TIME_END((cudaStream_t)stream);
  PS_LAUNCHED();
  return PS_OK;
}
