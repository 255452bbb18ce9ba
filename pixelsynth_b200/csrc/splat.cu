// Z-buffer point splat for sm_100a: project -> bin -> per-tile sort -> per-pixel K-nearest -> composite.
//
// Replaces, behind ps_project_pts / ps_splat_points / ps_splat_fwd (include/pixelsynth_b200.h):
//   reference models/projection/z_buffer_manipulator.py:38-83,221-266  (project_pts[_cumulative])
//   reference models/layers/z_buffer_layers.py:55-131                   (RasterizePointsXYsBlending.forward)
//   PyTorch3D rasterize_points + compositing (not vendored; semantics in SURVEY.md Appendix A)
//
// Design (HBM-bound: 69.0 MB of mandatory output per 256x256 view when the idx/z maps are emitted):
//   1. bin_count / scan / bin_fill : every point is appended to the candidate list of each 8x8-pixel
//      tile its disc can touch (conservative box, +1 px margin).  Lists hold point ids only.
//   2. fine_kernel, one 64-thread CTA per tile: candidates are sorted ONCE per tile by the canonical
//      key (z, point id) in shared memory; each thread then walks the sorted list for its pixel, so a
//      pixel's hits come out already in output order and no per-pixel sort or heap is needed.  A second
//      phase gives one warp lane to each output slot: alpha, transmittance prefix product, weighted
//      feature sum, and 512-byte coalesced row stores of idx / zbuf / dist2.
//   3. tiles whose list exceeds the shared-memory capacity are queued and handled by fine_big_kernel,
//      which streams the list in sorted chunks and merges into a per-pixel running top-K, so no point
//      is ever dropped (PyTorch3D's binned path silently drops on bin overflow).
//   4. bgmask_kernel: separable k x k box dilation of the "pixel received no point" map.
//
// Bit-exact surfaces (idx, zbuf, dist2, pts) use __f*_rn intrinsics so ptxas never contracts to FMA;
// this matches oracle/splat_oracle.c built with -ffp-contract=off.
#include "common.cuh"

namespace ps {

constexpr int TILE = 8;
constexpr int FINE_THREADS = TILE * TILE;
constexpr int CAP = 512;    // candidates per tile handled by the shared-memory fast path
constexpr int CAPB = 1024;  // chunk size of the overflow path
constexpr int MAXK = PS_MAX_POINTS_PER_PIXEL;
constexpr int MAXC_SMEM = 4;  // feature channels cached in shared memory by the fast path
constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// canonical arithmetic helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dot4(const float* __restrict__ m, float x0, float x1, float x2, float x3) {
  float s = __fmul_rn(m[0], x0);
  s = __fadd_rn(s, __fmul_rn(m[1], x1));
  s = __fadd_rn(s, __fmul_rn(m[2], x2));
  s = __fadd_rn(s, __fmul_rn(m[3], x3));
  return s;
}

// xs[i] = linspace(0, W-1, W)[i] / (W-1) * 2 - 1      (z_buffer_manipulator.py:38)
__device__ __forceinline__ float grid_coord(int i, int W) {
  return __fsub_rn(__fmul_rn(__fdiv_rn((float)i, (float)(W - 1)), 2.0f), 1.0f);
}

// PyTorch3D PixToNdc
__device__ __forceinline__ float pix_to_ndc(int i, int S) {
  return __fadd_rn(-1.0f, __fdiv_rn(__fadd_rn(__fmul_rn(2.0f, (float)i), 1.0f), (float)S));
}

__device__ __forceinline__ float dist2_rn(float dx, float dy) {
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// z_buffer_manipulator.py:69-81
__device__ __forceinline__ void finish_point(float qx, float qy, float qz, float eps, float* __restrict__ o) {
  const bool masked = fabsf(qz) < eps;
  if (masked) qz = eps;
  const float nz = -qz;
  float sx = __fdiv_rn(qx, nz), sy = __fdiv_rn(qy, nz), sz = qz;
  if (masked) { sx = -10.0f; sy = -10.0f; sz = -10.0f; }
  o[0] = sx;
  o[1] = -sy;
  o[2] = -sz;
}

__device__ __forceinline__ void matmul4_entry(const float* a, const float* b, float* c, int e) {
  const int i = e >> 2, j = e & 3;
  float s = __fmul_rn(a[i * 4 + 0], b[0 * 4 + j]);
  s = __fadd_rn(s, __fmul_rn(a[i * 4 + 1], b[1 * 4 + j]));
  s = __fadd_rn(s, __fmul_rn(a[i * 4 + 2], b[2 * 4 + j]));
  s = __fadd_rn(s, __fmul_rn(a[i * 4 + 3], b[3 * 4 + j]));
  c[e] = s;
}

// ------------------------------------------------------------------------------------------------
// stage 1-2: projection
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) project_kernel(const float* __restrict__ depth, const float* __restrict__ mats,
                                                      int W, float eps, float* __restrict__ pts,
                                                      float* __restrict__ xyproj) {
  __shared__ float sK[16], sKinv[16], sRT[16];
  const int b = blockIdx.y;
  const int P = W * W;
  const float* m = mats + (size_t)b * 96;
  if (threadIdx.x < 16) {
    sK[threadIdx.x] = m[threadIdx.x];
    sKinv[threadIdx.x] = m[16 + threadIdx.x];
    matmul4_entry(m + 64, m + 48, sRT, threadIdx.x);  // RT = RT2 * RT1inv
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int sy = p / W, sx = p - sy * W;
  const float d = depth[(size_t)b * P + p];
  const float X0 = __fmul_rn(grid_coord(sx, W), d);
  const float X1 = __fmul_rn(-grid_coord(sy, W), d);
  const float X2 = -d;
  float c[4], w[4], q[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) c[r] = dot4(sKinv + 4 * r, X0, X1, X2, 1.0f);
#pragma unroll
  for (int r = 0; r < 4; ++r) w[r] = dot4(sRT + 4 * r, c[0], c[1], c[2], c[3]);
#pragma unroll
  for (int r = 0; r < 4; ++r) q[r] = dot4(sK + 4 * r, w[0], w[1], w[2], w[3]);
  if (xyproj) {
    float* o = xyproj + (size_t)b * 4 * P;
    o[p] = q[0];
    o[(size_t)P + p] = q[1];
    o[2 * (size_t)P + p] = (fabsf(q[2]) < eps) ? eps : q[2];
    o[3 * (size_t)P + p] = q[3];
  }
  finish_point(q[0], q[1], q[2], eps, pts + ((size_t)b * P + p) * 3);
}

__global__ void __launch_bounds__(256) project_cloud_kernel(const float* __restrict__ cloud,
                                                            const float* __restrict__ mats3, int P, float eps,
                                                            float* __restrict__ pts, float* __restrict__ xyproj) {
  __shared__ float sK[16], sRT[16];
  const int b = blockIdx.y;
  const float* m = mats3 + (size_t)b * 48;
  if (threadIdx.x < 16) {
    sK[threadIdx.x] = m[threadIdx.x];
    matmul4_entry(m + 16, m + 32, sRT, threadIdx.x);  // RT_last = RT2 * RT3inv
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const float* cb = cloud + (size_t)b * 4 * P;
  const float X0 = cb[p], X1 = cb[(size_t)P + p], X2 = cb[2 * (size_t)P + p], X3 = cb[3 * (size_t)P + p];
  float w[4], q[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) w[r] = dot4(sRT + 4 * r, X0, X1, X2, X3);
#pragma unroll
  for (int r = 0; r < 4; ++r) q[r] = dot4(sK + 4 * r, w[0], w[1], w[2], w[3]);
  if (xyproj) {
    float* o = xyproj + (size_t)b * 4 * P;
    o[p] = q[0];
    o[(size_t)P + p] = q[1];
    o[2 * (size_t)P + p] = (fabsf(q[2]) < eps) ? eps : q[2];
    o[3 * (size_t)P + p] = q[3];
  }
  finish_point(q[0], q[1], q[2], eps, pts + ((size_t)b * P + p) * 3);
}

// ------------------------------------------------------------------------------------------------
// stage 3a: binning
// ------------------------------------------------------------------------------------------------
struct BinGeom {
  int S, nt;     // image side, tiles per side
  float half_S;  // S/2
  float rp;      // radius in pixels + 1 px safety margin
  float lim;     // |x| beyond this can never touch the image
};

// Conservative set of tiles a point's disc can touch.  Every pixel that passes the exact test
// (z >= 0, dx*dx+dy*dy < r*r) lies inside this box; the box may contain pixels that fail it.
__device__ __forceinline__ bool tile_range(const float* __restrict__ pt, const BinGeom& g, int& tx0, int& tx1, int& ty0,
                                           int& ty1) {
  const float x = pt[0], y = pt[1], z = pt[2];
  if (!(z >= 0.0f)) return false;
  if (!(x > -g.lim && x < g.lim && y > -g.lim && y < g.lim)) return false;  // also rejects NaN
  const float fx = (x + 1.0f) * g.half_S - 0.5f;
  const float fy = (y + 1.0f) * g.half_S - 0.5f;
  int x0 = (int)floorf(fx - g.rp), x1 = (int)ceilf(fx + g.rp);
  int y0 = (int)floorf(fy - g.rp), y1 = (int)ceilf(fy + g.rp);
  x0 = max(x0, 0);
  y0 = max(y0, 0);
  x1 = min(x1, g.S - 1);
  y1 = min(y1, g.S - 1);
  if (x0 > x1 || y0 > y1) return false;
  tx0 = x0 / TILE;
  tx1 = x1 / TILE;
  ty0 = y0 / TILE;
  ty1 = y1 / TILE;
  return true;
}

__global__ void __launch_bounds__(256) bin_count_kernel(const float* __restrict__ pts, int P, BinGeom g,
                                                        int* __restrict__ tile_count) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  int tx0, tx1, ty0, ty1;
  if (!tile_range(pts + ((size_t)b * P + p) * 3, g, tx0, tx1, ty0, ty1)) return;
  int* tc = tile_count + (size_t)b * g.nt * g.nt;
  for (int ty = ty0; ty <= ty1; ++ty)
    for (int tx = tx0; tx <= tx1; ++tx) atomicAdd(tc + ty * g.nt + tx, 1);
}

// one CTA per view: exclusive scan of the view's tile counts; counts are zeroed for reuse as cursors
__global__ void __launch_bounds__(1024) bin_scan_kernel(int* __restrict__ tile_count, int* __restrict__ tile_start,
                                                        int nt2) {
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  const int b = blockIdx.x;
  int* tc = tile_count + (size_t)b * nt2;
  int* ts = tile_start + (size_t)b * nt2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nt2; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = (i < nt2) ? tc[i] : 0;
    int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(FULL, incl, off);
      if (lane >= off) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int ws = warp_sums[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(FULL, ws, off);
        if (lane >= off) ws += t;
      }
      warp_sums[lane] = ws;  // inclusive
    }
    __syncthreads();
    const int carry = carry_s;
    const int woff = (warp == 0) ? 0 : warp_sums[warp - 1];
    if (i < nt2) {
      ts[i] = carry + woff + incl - v;
      tc[i] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + warp_sums[31];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) bin_fill_kernel(const float* __restrict__ pts, int P, BinGeom g,
                                                       int* __restrict__ tile_count, const int* __restrict__ tile_start,
                                                       int* __restrict__ list, long long cap_per_view) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  int tx0, tx1, ty0, ty1;
  if (!tile_range(pts + ((size_t)b * P + p) * 3, g, tx0, tx1, ty0, ty1)) return;
  const int nt2 = g.nt * g.nt;
  int* tc = tile_count + (size_t)b * nt2;
  const int* ts = tile_start + (size_t)b * nt2;
  int* lst = list + (size_t)b * cap_per_view;
  for (int ty = ty0; ty <= ty1; ++ty)
    for (int tx = tx0; tx <= tx1; ++tx) {
      const int t = ty * g.nt + tx;
      const int slot = atomicAdd(tc + t, 1);
      lst[ts[t] + slot] = p;
    }
}

// ------------------------------------------------------------------------------------------------
// stage 3b-4: per-tile fine rasterisation + compositing
// ------------------------------------------------------------------------------------------------
struct FineParams {
  const float* pts;   // (B,P,3)
  const float* feat;  // (B,C,P)
  const int* tile_count;
  const int* tile_start;
  const int* list;
  long long cap_per_view;
  int P, C, S, K, nt;
  float r2;     // (float)radius * (float)radius
  float denom;  // (float)pow(radius, rad_pow)
  float tau;
  int accumulation;
  float* out;        // (B,C,S,S)
  uint8_t* empty;    // (B,S,S) 1 where the pixel received no point
  int32_t* idx;      // (B,S,S,K) or null
  float* zbuf;       // (B,S,S,K) or null
  float* dist2;      // (B,S,S,K) or null
  int* ovf_count;    // tiles that exceeded CAP
  int* ovf_list;
};

__device__ __forceinline__ float alpha_of(float d2, float denom, float tau) {
  float d = __fdiv_rn(d2, denom);
  d = fminf(fmaxf(d, 1e-3f), 1.0f);
  float a = 1.0f - sqrtf(d);
  if (tau != 1.0f) a = powf(a, tau);
  return a;
}

__device__ __forceinline__ void bitonic_sort_u64(unsigned long long* __restrict__ key, int N, int tid, int nthreads) {
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (N >> 1); t += nthreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int ixj = i | j;
        const unsigned long long a = key[i], c = key[ixj];
        const bool up = (i & k) == 0;
        if ((a > c) == up) {
          key[i] = c;
          key[ixj] = a;
        }
      }
      __syncthreads();
    }
  }
}

// Per-warp staging of one pixel's output rows so the global stores are 16-byte vectors.
struct RowStage {
  int idx[MAXK];
  float z[MAXK];
  float d2[MAXK];
  float w[MAXK];  // compositing weight per slot (generic-C path)
};

__device__ __forceinline__ void store_rows(const FineParams& q, const RowStage& st, size_t pixoff, int lane) {
  const int K = q.K;
  if ((K & 3) == 0) {
    if (4 * lane < K) {
      if (q.idx) reinterpret_cast<int4*>(q.idx + pixoff * K)[lane] = reinterpret_cast<const int4*>(st.idx)[lane];
      if (q.zbuf) reinterpret_cast<float4*>(q.zbuf + pixoff * K)[lane] = reinterpret_cast<const float4*>(st.z)[lane];
      if (q.dist2) reinterpret_cast<float4*>(q.dist2 + pixoff * K)[lane] = reinterpret_cast<const float4*>(st.d2)[lane];
    }
  } else {
    for (int k = lane; k < K; k += 32) {
      if (q.idx) q.idx[pixoff * K + k] = st.idx[k];
      if (q.zbuf) q.zbuf[pixoff * K + k] = st.z[k];
      if (q.dist2) q.dist2[pixoff * K + k] = st.d2[k];
    }
  }
}

template <int CAP_>
struct FineSmem {
  unsigned long long key[CAP_];
  float2 xy[CAP_];  // negated coordinates (z_buffer_layers.py:71-72)
  float z[CAP_];
  int p[CAP_];
  float f[MAXC_SMEM][CAP_];
  unsigned short lists[MAXK][FINE_THREADS];
  RowStage stage[FINE_THREADS / 32];
  float col[MAXC_SMEM][FINE_THREADS];
};

__global__ void __launch_bounds__(FINE_THREADS) fine_kernel(FineParams q) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FineSmem<CAP>& sm = *reinterpret_cast<FineSmem<CAP>*>(smem_raw);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int t = blockIdx.x;
  const int nt2 = q.nt * q.nt;
  const int ty = t / q.nt, tx = t - ty * q.nt;
  const int n = q.tile_count[(size_t)b * nt2 + t];
  if (n > CAP) {  // queue for fine_big_kernel
    if (tid == 0) q.ovf_list[atomicAdd(q.ovf_count, 1)] = b * nt2 + t;
    return;
  }
  const int S = q.S, K = q.K, P = q.P, C = q.C;
  const int xi = tx * TILE + (tid & 7), yi = ty * TILE + (tid >> 3);
  const bool inimg = xi < S && yi < S;
  const float xf = pix_to_ndc(S - 1 - xi, S), yf = pix_to_ndc(S - 1 - yi, S);
  const float* ptsb = q.pts + (size_t)b * P * 3;
  const bool cached = C <= MAXC_SMEM;

  // ---- load candidate keys, sort by (z, point id) ----
  int N = 2;
  while (N < n) N <<= 1;
  const int* lst = q.list + (size_t)b * q.cap_per_view + q.tile_start[(size_t)b * nt2 + t];
  for (int i = tid; i < N; i += FINE_THREADS) {
    unsigned long long key = ~0ull;
    if (i < n) {
      const int p = lst[i];
      const float z = ptsb[(size_t)p * 3 + 2] + 0.0f;  // -0 -> +0 so the bit pattern orders like the float
      key = ((unsigned long long)__float_as_uint(z) << 32) | (unsigned)p;
    }
    sm.key[i] = key;
  }
  __syncthreads();
  if (n > 1) bitonic_sort_u64(sm.key, N, tid, FINE_THREADS);
  for (int i = tid; i < n; i += FINE_THREADS) {
    const int p = (int)(unsigned)(sm.key[i] & 0xffffffffull);
    const float* pt = ptsb + (size_t)p * 3;
    sm.xy[i] = make_float2(-pt[0], -pt[1]);
    sm.z[i] = pt[2];
    sm.p[i] = p;
    if (cached)
      for (int c = 0; c < C; ++c) sm.f[c][i] = q.feat[((size_t)b * C + c) * P + p];
  }
  __syncthreads();

  // ---- phase 1: one thread per pixel walks the sorted candidates; hits come out in output order ----
  int cnt = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    if (__all_sync(FULL, cnt >= K)) break;
    const int i1 = min(i0 + 32, n);
    for (int i = i0; i < i1; ++i) {
      const float2 c = sm.xy[i];
      const float d2 = dist2_rn(c.x - xf, c.y - yf);
      if (d2 < q.r2 && cnt < K) {
        sm.lists[cnt][tid] = (unsigned short)i;
        ++cnt;
      }
    }
  }
  if (inimg) q.empty[((size_t)b * S + yi) * S + xi] = (cnt == 0);
  __syncwarp();

  // ---- phase 2: one lane per output slot ----
  RowStage& st = sm.stage[warp];
  const int rounds = (K + 31) >> 5;
  const int32_t base = (int32_t)((size_t)b * P);
  for (int j = 0; j < 32; ++j) {
    if (!__shfl_sync(FULL, (int)inimg, j)) continue;
    const int pix = warp * 32 + j;
    const int nh = __shfl_sync(FULL, cnt, j);
    const float xfj = __shfl_sync(FULL, xf, j), yfj = __shfl_sync(FULL, yf, j);
    const int pxi = __shfl_sync(FULL, xi, j), pyi = __shfl_sync(FULL, yi, j);
    float acc[MAXC_SMEM] = {0.f, 0.f, 0.f, 0.f};
    float tcarry = 1.0f, wsum = 0.0f;
    for (int r = 0; r < rounds; ++r) {
      const int k = r * 32 + lane;
      const bool valid = k < nh;
      float a = 0.0f, z = -1.0f, d2 = -1.0f;
      int pid = -1, ci = 0;
      if (valid) {
        ci = sm.lists[k][pix];
        const float2 c = sm.xy[ci];
        d2 = dist2_rn(c.x - xfj, c.y - yfj);
        z = sm.z[ci];
        pid = base + sm.p[ci];
        a = alpha_of(d2, q.denom, q.tau);
      }
      if (k < K) {
        st.idx[k] = pid;
        st.z[k] = z;
        st.d2[k] = d2;
      }
      if (r * 32 >= nh) continue;  // warp-uniform: nothing to composite in this round
      float wgt;
      if (q.accumulation == PS_ACCUM_ALPHACOMPOSITE) {
        float incl = 1.0f - a;  // a == 0 for invalid lanes
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const float v = __shfl_up_sync(FULL, incl, off);
          if (lane >= off) incl *= v;
        }
        float excl = __shfl_up_sync(FULL, incl, 1);
        if (lane == 0) excl = 1.0f;
        wgt = tcarry * excl * a;
        tcarry *= __shfl_sync(FULL, incl, 31);
      } else {
        wgt = a;
        wsum += a;
      }
      if (cached) {
        if (valid) {
#pragma unroll
          for (int c = 0; c < MAXC_SMEM; ++c)
            if (c < C) acc[c] += wgt * sm.f[c][ci];
        }
      } else if (k < K) {
        st.w[k] = valid ? wgt : 0.0f;
      }
    }
    __syncwarp();
    const size_t pixoff = ((size_t)b * S + pyi) * S + pxi;
    store_rows(q, st, pixoff, lane);
    if (q.accumulation != PS_ACCUM_ALPHACOMPOSITE) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) wsum += __shfl_xor_sync(FULL, wsum, off);
    }
    const float norm = (q.accumulation == PS_ACCUM_WSUMNORM) ? fmaxf(wsum, 1e-4f) : 1.0f;
    if (cached) {
#pragma unroll
      for (int c = 0; c < MAXC_SMEM; ++c) {
        if (c < C) {
          float v = acc[c];
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
          if (lane == 0) sm.col[c][pix] = (q.accumulation == PS_ACCUM_WSUMNORM) ? v / norm : v;
        }
      }
    } else {
      const int nhk = min(nh, K);
      for (int c = lane; c < C; c += 32) {
        const float* fc = q.feat + ((size_t)b * C + c) * P - base;  // st.idx holds packed ids
        float v = 0.0f;
        for (int k = 0; k < nhk; ++k) v += st.w[k] * fc[st.idx[k]];
        if (q.accumulation == PS_ACCUM_WSUMNORM) v = v / norm;
        q.out[(((size_t)b * C + c) * S + pyi) * S + pxi] = v;
      }
    }
    __syncwarp();
  }
  if (cached) {
    __syncthreads();
    if (inimg)
      for (int c = 0; c < C; ++c) q.out[(((size_t)b * C + c) * S + yi) * S + xi] = sm.col[c][tid];
  }
}

// Overflow path: tiles with more than CAP candidates.  Streams the candidate list in chunks of CAPB,
// sorts each chunk, and merges the chunk's hits (already ordered) into each pixel's running top-K.
struct BigSmem {
  unsigned long long key[CAPB];
  float2 xy[CAPB];
  unsigned long long topA[MAXK][FINE_THREADS];
  unsigned long long topB[MAXK][FINE_THREADS];
  RowStage stage[FINE_THREADS / 32];
};

__global__ void __launch_bounds__(FINE_THREADS) fine_big_kernel(FineParams q) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BigSmem& sm = *reinterpret_cast<BigSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nt2 = q.nt * q.nt;
  const int S = q.S, K = q.K, P = q.P, C = q.C;
  const int novf = *q.ovf_count;
  for (int o = blockIdx.x; o < novf; o += gridDim.x) {
    const int bt = q.ovf_list[o];
    const int b = bt / nt2, t = bt - b * nt2;
    const int ty = t / q.nt, tx = t - ty * q.nt;
    const int n = q.tile_count[bt];
    const int xi = tx * TILE + (tid & 7), yi = ty * TILE + (tid >> 3);
    const bool inimg = xi < S && yi < S;
    const float xf = pix_to_ndc(S - 1 - xi, S), yf = pix_to_ndc(S - 1 - yi, S);
    const float* ptsb = q.pts + (size_t)b * P * 3;
    const int* lst = q.list + (size_t)b * q.cap_per_view + q.tile_start[bt];
    unsigned long long(*A)[FINE_THREADS] = sm.topA;
    unsigned long long(*Bv)[FINE_THREADS] = sm.topB;
    int cntA = 0;
    for (int c0 = 0; c0 < n; c0 += CAPB) {
      const int m = min(CAPB, n - c0);
      int N = 2;
      while (N < m) N <<= 1;
      __syncthreads();
      for (int i = tid; i < N; i += FINE_THREADS) {
        unsigned long long key = ~0ull;
        if (i < m) {
          const int p = lst[c0 + i];
          const float z = ptsb[(size_t)p * 3 + 2] + 0.0f;
          key = ((unsigned long long)__float_as_uint(z) << 32) | (unsigned)p;
        }
        sm.key[i] = key;
      }
      __syncthreads();
      bitonic_sort_u64(sm.key, N, tid, FINE_THREADS);
      for (int i = tid; i < m; i += FINE_THREADS) {
        const int p = (int)(unsigned)(sm.key[i] & 0xffffffffull);
        sm.xy[i] = make_float2(-ptsb[(size_t)p * 3], -ptsb[(size_t)p * 3 + 1]);
      }
      __syncthreads();
      // merge: A (sorted, cntA) with this chunk's hits (sorted) -> Bv, keeping the K smallest
      int ia = 0, nb = 0;
      for (int i = 0; i < m; ++i) {
        const float2 c = sm.xy[i];
        const float d2 = dist2_rn(c.x - xf, c.y - yf);
        if (d2 < q.r2 && nb < K) {
          const unsigned long long key = sm.key[i];
          while (ia < cntA && nb < K && A[ia][tid] < key) Bv[nb++][tid] = A[ia++][tid];
          if (nb < K) Bv[nb++][tid] = key;
        }
      }
      while (ia < cntA && nb < K) Bv[nb++][tid] = A[ia++][tid];
      cntA = nb;
      unsigned long long(*tmp)[FINE_THREADS] = A;
      A = Bv;
      Bv = tmp;
    }
    if (inimg) q.empty[((size_t)b * S + yi) * S + xi] = (cntA == 0);
    __syncwarp();
    // phase 2 from keys; coordinates and features are gathered from global memory
    RowStage& st = sm.stage[warp];
    const int rounds = (K + 31) >> 5;
    const int32_t base = (int32_t)((size_t)b * P);
    for (int j = 0; j < 32; ++j) {
      if (!__shfl_sync(FULL, (int)inimg, j)) continue;
      const int pix = warp * 32 + j;
      const int nh = __shfl_sync(FULL, cntA, j);
      const float xfj = __shfl_sync(FULL, xf, j), yfj = __shfl_sync(FULL, yf, j);
      const int pxi = __shfl_sync(FULL, xi, j), pyi = __shfl_sync(FULL, yi, j);
      float tcarry = 1.0f, wsum = 0.0f;
      for (int r = 0; r < rounds; ++r) {
        const int k = r * 32 + lane;
        const bool valid = k < nh;
        float a = 0.0f, z = -1.0f, d2 = -1.0f;
        int pid = -1;
        if (valid) {
          const int p = (int)(unsigned)(A[k][pix] & 0xffffffffull);
          const float* pt = ptsb + (size_t)p * 3;
          d2 = dist2_rn(-pt[0] - xfj, -pt[1] - yfj);
          z = pt[2];
          pid = base + p;
          a = alpha_of(d2, q.denom, q.tau);
        }
        float wgt = 0.0f;
        if (r * 32 < nh) {
          if (q.accumulation == PS_ACCUM_ALPHACOMPOSITE) {
            float incl = 1.0f - a;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
              const float v = __shfl_up_sync(FULL, incl, off);
              if (lane >= off) incl *= v;
            }
            float excl = __shfl_up_sync(FULL, incl, 1);
            if (lane == 0) excl = 1.0f;
            wgt = tcarry * excl * a;
            tcarry *= __shfl_sync(FULL, incl, 31);
          } else {
            wgt = a;
            wsum += a;
          }
        }
        if (k < K) {
          st.idx[k] = pid;
          st.z[k] = z;
          st.d2[k] = d2;
          st.w[k] = valid ? wgt : 0.0f;
        }
      }
      __syncwarp();
      const size_t pixoff = ((size_t)b * S + pyi) * S + pxi;
      store_rows(q, st, pixoff, lane);
      if (q.accumulation != PS_ACCUM_ALPHACOMPOSITE) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) wsum += __shfl_xor_sync(FULL, wsum, off);
      }
      const float norm = fmaxf(wsum, 1e-4f);
      const int nhk = min(nh, K);
      for (int c = lane; c < C; c += 32) {
        const float* fc = q.feat + ((size_t)b * C + c) * P - base;
        float v = 0.0f;
        for (int k = 0; k < nhk; ++k) v += st.w[k] * fc[st.idx[k]];
        if (q.accumulation == PS_ACCUM_WSUMNORM) v = v / norm;
        q.out[(((size_t)b * C + c) * S + pyi) * S + pxi] = v;
      }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// stage 5: background mask = box_dilate_k(empty)       (z_buffer_layers.py:100-110)
// ------------------------------------------------------------------------------------------------
constexpr int BG_TILE = 32;
constexpr int BG_MAXH = 15;  // supports odd ksize <= 31

__global__ void __launch_bounds__(BG_TILE* BG_TILE) bgmask_kernel(const uint8_t* __restrict__ empty, int S, int h,
                                                                  uint8_t* __restrict__ bg) {
  __shared__ uint8_t s_in[BG_TILE + 2 * BG_MAXH][BG_TILE + 2 * BG_MAXH];
  __shared__ uint8_t s_row[BG_TILE + 2 * BG_MAXH][BG_TILE];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * BG_TILE, y0 = blockIdx.y * BG_TILE;
  const int ext = BG_TILE + 2 * h;
  const uint8_t* e = empty + (size_t)b * S * S;
  for (int i = threadIdx.x; i < ext * ext; i += blockDim.x) {
    const int ry = i / ext, rx = i - ry * ext;
    const int y = y0 + ry - h, x = x0 + rx - h;
    s_in[ry][rx] = (y >= 0 && y < S && x >= 0 && x < S) ? e[(size_t)y * S + x] : 0;  // zero padding
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ext * BG_TILE; i += blockDim.x) {
    const int ry = i / BG_TILE, cx = i - ry * BG_TILE;
    uint8_t v = 0;
    for (int d = 0; d <= 2 * h; ++d) v |= s_in[ry][cx + d];
    s_row[ry][cx] = v;
  }
  __syncthreads();
  const int cy = threadIdx.x / BG_TILE, cx = threadIdx.x - cy * BG_TILE;
  const int y = y0 + cy, x = x0 + cx;
  if (y < S && x < S) {
    uint8_t v = 0;
    for (int d = 0; d <= 2 * h; ++d) v |= s_row[cy + d][cx];
    bg[(size_t)b * S * S + (size_t)y * S + x] = v ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------
// host drivers
// ------------------------------------------------------------------------------------------------
static int tile_span(double radius_px) { return (int)((2.0 * (radius_px + 1.0) + 1.0) / TILE) + 2; }

struct SplatLayout {
  int nt, nt2;
  long long cap_per_view;
  size_t bytes;
  size_t off_count, off_start, off_list, off_empty, off_ovf;
};

static SplatLayout splat_layout(int B, int P, int S, double radius_px) {
  SplatLayout L;
  L.nt = (S + TILE - 1) / TILE;
  L.nt2 = L.nt * L.nt;
  const long long span = tile_span(radius_px);
  L.cap_per_view = (long long)P * span * span;
  if (L.cap_per_view > (long long)P * L.nt2) L.cap_per_view = (long long)P * L.nt2;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    off = align_up(off, 256);
    size_t r = off;
    off += bytes;
    return r;
  };
  L.off_count = take(sizeof(int) * (size_t)B * L.nt2);
  L.off_ovf = take(sizeof(int) * (1 + (size_t)B * L.nt2));  // adjacent to counts: one memset clears both
  L.off_start = take(sizeof(int) * (size_t)B * L.nt2);
  L.off_list = take(sizeof(int) * (size_t)B * (size_t)L.cap_per_view);
  L.off_empty = take((size_t)B * S * S);
  L.bytes = align_up(off, 256);
  return L;
}

static int splat_points_impl(const float* pts, const float* feat, int B, int P, int C, int S, int K, double radius_px,
                             double tau, int rad_pow, int accumulation, int bg_ksize, float* out, uint8_t* bg_mask,
                             int32_t* idx, float* zbuf, float* dist2, void* workspace, size_t workspace_bytes,
                             cudaStream_t stream) {
  PS_CHECK_ARG(pts && feat && out && bg_mask && workspace);
  PS_CHECK_ARG(B >= 0 && P >= 0 && C >= 1 && S >= 1);
  PS_CHECK_ARG(K >= 1);
  if (K > MAXK) return fail(PS_EUNSUPPORTED, "%s: points_per_pixel exceeds PS_MAX_POINTS_PER_PIXEL%s", __func__);
  PS_CHECK_ARG(accumulation >= 0 && accumulation <= 2);
  PS_CHECK_ARG(bg_ksize >= 1 && (bg_ksize & 1) == 1 && bg_ksize / 2 <= BG_MAXH);
  PS_CHECK_ARG(radius_px > 0.0 && radius_px <= 64.0);
  PS_CHECK_ARG((size_t)B * P < 0x7fffffffull);  // packed indices are int32, as in PyTorch3D
  if (B == 0) return PS_OK;
  const SplatLayout L = splat_layout(B, P, S, radius_px);
  if (workspace_bytes < L.bytes) return fail(PS_EWORKSPACE, "%s: workspace too small%s", __func__);
  char* ws = (char*)workspace;
  int* tile_count = (int*)(ws + L.off_count);
  int* ovf = (int*)(ws + L.off_ovf);
  int* tile_start = (int*)(ws + L.off_start);
  int* list = (int*)(ws + L.off_list);
  uint8_t* empty = (uint8_t*)(ws + L.off_empty);

  const double radius = radius_px / (double)S * 2.0;  // z_buffer_layers.py:77
  const float rf = (float)radius;
  BinGeom g;
  g.S = S;
  g.nt = L.nt;
  g.half_S = 0.5f * (float)S;
  g.rp = (float)(radius_px + 1.0);
  g.lim = 1.0f + rf + 4.0f / (float)S;

  PS_CUDA(cudaMemsetAsync(tile_count, 0, (L.off_ovf - L.off_count) + sizeof(int), stream));
  if (P > 0) {
    dim3 grid((P + 255) / 256, B);
    bin_count_kernel<<<grid, 256, 0, stream>>>(pts, P, g, tile_count);
    PS_LAUNCHED();
    bin_scan_kernel<<<B, 1024, 0, stream>>>(tile_count, tile_start, L.nt2);
    PS_LAUNCHED();
    bin_fill_kernel<<<grid, 256, 0, stream>>>(pts, P, g, tile_count, tile_start, list, L.cap_per_view);
    PS_LAUNCHED();
  }
  FineParams q;
  q.pts = pts;
  q.feat = feat;
  q.tile_count = tile_count;
  q.tile_start = tile_start;
  q.list = list;
  q.cap_per_view = L.cap_per_view;
  q.P = P;
  q.C = C;
  q.S = S;
  q.K = K;
  q.nt = L.nt;
  q.r2 = rf * rf;
  q.denom = (float)pow(radius, (double)rad_pow);
  q.tau = (float)tau;
  q.accumulation = accumulation;
  q.out = out;
  q.empty = empty;
  q.idx = idx;
  q.zbuf = zbuf;
  q.dist2 = dist2;
  q.ovf_count = ovf;
  q.ovf_list = ovf + 1;

  static thread_local int smem_set_for_device = -1;
  int dev = 0;
  PS_CUDA(cudaGetDevice(&dev));
  if (smem_set_for_device != dev) {
    PS_CUDA(cudaFuncSetAttribute(fine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FineSmem<CAP>)));
    PS_CUDA(cudaFuncSetAttribute(fine_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BigSmem)));
    smem_set_for_device = dev;
  }
  {
    dim3 grid(L.nt2, B);
    PS_TIME_BEGIN("fine_kernel", stream);
    fine_kernel<<<grid, FINE_THREADS, sizeof(FineSmem<CAP>), stream>>>(q);
    PS_TIME_END(stream);
    PS_LAUNCHED();
    fine_big_kernel<<<296, FINE_THREADS, sizeof(BigSmem), stream>>>(q);
    PS_LAUNCHED();
  }
  {
    dim3 grid((S + BG_TILE - 1) / BG_TILE, (S + BG_TILE - 1) / BG_TILE, B);
    bgmask_kernel<<<grid, BG_TILE * BG_TILE, 0, stream>>>(empty, S, bg_ksize / 2, bg_mask);
    PS_LAUNCHED();
  }
  return PS_OK;
}

}  // namespace ps

using namespace ps;

extern "C" {

int ps_project_pts(const float* depth, const float* mats, int B, int W, float eps, float* pts, float* xyproj,
                   void* stream) {
  PS_CHECK_ARG(depth && mats && pts);
  PS_CHECK_ARG(B >= 0 && W >= 2 && W <= 4096);
  if (B == 0) return PS_OK;
  const int P = W * W;
  dim3 grid((P + 255) / 256, B);
  project_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(depth, mats, W, eps, pts, xyproj);
  PS_LAUNCHED();
  return PS_OK;
}

int ps_project_cloud(const float* cloud, const float* mats3, int B, int P, float eps, float* pts, float* xyproj,
                     void* stream) {
  PS_CHECK_ARG(cloud && mats3 && pts);
  PS_CHECK_ARG(B >= 0 && P >= 0);
  if (B == 0 || P == 0) return PS_OK;
  dim3 grid((P + 255) / 256, B);
  project_cloud_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cloud, mats3, P, eps, pts, xyproj);
  PS_LAUNCHED();
  return PS_OK;
}

size_t ps_splat_workspace_bytes(int B, int P, int S, double radius_px) {
  if (B <= 0 || P < 0 || S < 1 || !(radius_px > 0.0)) return 256;
  return splat_layout(B, P, S, radius_px).bytes;
}

int ps_splat_points(const float* pts, const float* feat, int B, int P, int C, int S, int K, double radius_px,
                    double tau, int rad_pow, int accumulation, int bg_ksize, float* out, uint8_t* bg_mask,
                    int32_t* idx, float* zbuf, float* dist2, void* workspace, size_t workspace_bytes, void* stream) {
  return splat_points_impl(pts, feat, B, P, C, S, K, radius_px, tau, rad_pow, accumulation, bg_ksize, out, bg_mask, idx,
                           zbuf, dist2, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t ps_splat_fwd_workspace_bytes(int B, int W, int S, double radius_px) {
  if (B <= 0 || W < 2) return 256;
  const size_t P = (size_t)W * W;
  return align_up(sizeof(float) * 3 * P * B, 256) + ps_splat_workspace_bytes(B, (int)P, S, radius_px);
}

int ps_splat_fwd(const float* depth, const float* feat, const float* mats, int B, int W, int C, int S, int K,
                 double radius_px, double tau, int rad_pow, int accumulation, int bg_ksize, float eps, float* out,
                 uint8_t* bg_mask, int32_t* idx, float* zbuf, float* dist2, void* workspace, size_t workspace_bytes,
                 void* stream) {
  PS_CHECK_ARG(workspace);
  PS_CHECK_ARG(B >= 0 && W >= 2 && W <= 4096);
  if (B == 0) return PS_OK;
  const size_t P = (size_t)W * W;
  const size_t pts_bytes = align_up(sizeof(float) * 3 * P * B, 256);
  if (workspace_bytes < pts_bytes) return fail(PS_EWORKSPACE, "%s: workspace too small%s", __func__);
  float* pts = (float*)workspace;
  int rc = ps_project_pts(depth, mats, B, W, eps, pts, nullptr, stream);
  if (rc != PS_OK) return rc;
  return splat_points_impl(pts, feat, B, (int)P, C, S, K, radius_px, tau, rad_pow, accumulation, bg_ksize, out, bg_mask,
                           idx, zbuf, dist2, (char*)workspace + pts_bytes, workspace_bytes - pts_bytes,
                           (cudaStream_t)stream);
}

}  // extern "C"
