// Z-buffer point splat for sm_100a: project -> bin -> per-tile sort -> per-pixel K-nearest -> composite.
//
// Replaces, behind ps_project_pts / ps_splat_points / ps_splat_fwd (include/pixelsynth_b200.h):
//   reference models/projection/z_buffer_manipulator.py:38-83,221-266  (project_pts[_cumulative])
//   reference models/layers/z_buffer_layers.py:55-131                   (RasterizePointsXYsBlending.forward)
//   PyTorch3D rasterize_points + compositing (not vendored; semantics in SURVEY.md Appendix A)
//
// Design (HBM-bound: 69.0 MB of mandatory output per 256x256 view when the idx/z maps are emitted):
//   1. bin_kernel (projection fused in): every point is appended, with warp-aggregated atomics, to the
//      fixed-capacity candidate list of each 8x8-pixel tile its disc can reach (tight, conservative box).
//   2. fine_kernel, one 128-thread CTA per tile: candidates are sorted ONCE per tile by the canonical key
//      (z, point id) with a shared-memory counting sort; warps then ballot the exact membership test of 32
//      sorted candidates against the tile's 64 pixels, so every pixel's hits come out already in output
//      order as bit masks; two threads per pixel expand the masks and composite front to back; finally
//      each warp streams a pixel's K slots of idx / zbuf / dist2 with 16-byte coalesced stores.
//   3. tiles with more candidates than the shared-memory capacity are queued for fine_big_kernel, which
//      streams the list (or, if even the list overflowed, the whole cloud) in sorted chunks and merges into a
//      per-pixel running top-K, so no point is ever dropped (PyTorch3D's binned path drops on bin overflow).
//   4. bgmask_kernel: separable k x k box dilation of the "pixel received no point" map.
//
// Bit-exact surfaces (idx, zbuf, dist2, pts) use __f*_rn intrinsics so ptxas never contracts to FMA;
// this matches oracle/splat_oracle.c built with -ffp-contract=off.
#include "common.cuh"
#include "tc05.cuh"

namespace ps {

constexpr int TILE = 8;
constexpr int NPIX = TILE * TILE;  // pixels per tile
constexpr int FTPB = 128;          // threads of the tile kernel: two per pixel
constexpr int TPB = NPIX;          // threads of the overflow kernel: one per pixel
constexpr int CAP = 512;          // candidates per tile handled by the shared-memory fast path
constexpr int CAPG = 1024;        // capacity of a tile's candidate list in global memory
constexpr int CAPB = 1024;        // chunk size of the overflow path
constexpr int NB = 1024;          // buckets of the per-tile counting sort
constexpr int MAXBK = 32;         // largest bucket ranked by comparison; beyond that the tile is bitonic-sorted
constexpr int MAXK = PS_MAX_POINTS_PER_PIXEL;
constexpr int LSTRIDE = MAXK + 4;    // u16 per pixel row of the slot lists (8-byte aligned rows)
constexpr int BSTRIDE = CAP / 32 + 1;  // words per pixel row of the hit masks (odd: conflict-free columns)
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
// stage 3a: binning.  One pass: every point is appended (warp-aggregated atomics) to the fixed-capacity
// candidate list of each 8x8-pixel tile whose pixels its disc can reach.  tile_count keeps the true
// number of candidates even when it exceeds the list capacity, so nothing is dropped silently: such
// tiles are re-derived from the whole cloud by fine_big_kernel.
// ------------------------------------------------------------------------------------------------
struct BinGeom {
  int S, nt;     // image side, tiles per side
  float half_S;  // S/2
  float rp;      // radius in pixels + 1/64 px (covers the fp32 error of the pixel-coordinate estimate)
  float lim;     // |x| beyond this can never touch the image
};

// Tile range of the pixel columns/rows a point can reach.  A pixel passes the exact test only if
// |dx| < r(1+2^-22) and |dy| < r(1+2^-22) (dx*dx <= dx*dx+dy*dy under round-to-nearest), i.e. its index lies
// within radius_px(1+2^-21) + S*2^-22 of the point's pixel coordinate; rp adds 1/64 px on top of radius_px.
__device__ __forceinline__ bool tile_range(float x, float y, float z, const BinGeom& g, int& tx0, int& tx1, int& ty0,
                                           int& ty1) {
  if (!(z >= 0.0f)) return false;                                           // behind the camera, or NaN
  if (!(x > -g.lim && x < g.lim && y > -g.lim && y < g.lim)) return false;  // also rejects NaN
  const float fx = (x + 1.0f) * g.half_S - 0.5f;
  const float fy = (y + 1.0f) * g.half_S - 0.5f;
  int x0 = (int)ceilf(fx - g.rp), x1 = (int)floorf(fx + g.rp);
  int y0 = (int)ceilf(fy - g.rp), y1 = (int)floorf(fy + g.rp);
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

struct ProjectArgs {
  const float* depth;  // (B,P) when projecting in-kernel
  const float* mats;   // (B,6,16)
  int W;
  float eps;
};

// PROJECT: points are produced here from depth (project_pts fused in); otherwise read from pts (B,P,3).
// Either way a 16-byte-aligned copy (x, y, z, 0) goes to pts4 for the tile kernels.
template <bool PROJECT>
__global__ void __launch_bounds__(256) bin_kernel(ProjectArgs pa, const float* __restrict__ pts, int P, BinGeom g,
                                                  float4* __restrict__ pts4, int* __restrict__ tile_count,
                                                  int* __restrict__ list) {
  __shared__ float sK[16], sKinv[16], sRT[16];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  if (PROJECT) {
    const float* m = pa.mats + (size_t)b * 96;
    if (threadIdx.x < 16) {
      sK[threadIdx.x] = m[threadIdx.x];
      sKinv[threadIdx.x] = m[16 + threadIdx.x];
      matmul4_entry(m + 64, m + 48, sRT, threadIdx.x);  // RT = RT2 * RT1inv
    }
    __syncthreads();
  }
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  float o[3] = {0.f, 0.f, -1.f};
  if (p < P) {
    if (PROJECT) {
      const int W = pa.W;
      const int sy = p / W, sx = p - sy * W;
      const float d = pa.depth[(size_t)b * P + p];
      const float X0 = __fmul_rn(grid_coord(sx, W), d);
      const float X1 = __fmul_rn(-grid_coord(sy, W), d);
      const float X2 = -d;
      float c[4], w[4], q[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) c[r] = dot4(sKinv + 4 * r, X0, X1, X2, 1.0f);
#pragma unroll
      for (int r = 0; r < 4; ++r) w[r] = dot4(sRT + 4 * r, c[0], c[1], c[2], c[3]);
#pragma unroll
      for (int r = 0; r < 3; ++r) q[r] = dot4(sK + 4 * r, w[0], w[1], w[2], w[3]);
      finish_point(q[0], q[1], q[2], pa.eps, o);
    } else {
      const float* pt = pts + ((size_t)b * P + p) * 3;
      o[0] = pt[0];
      o[1] = pt[1];
      o[2] = pt[2];
    }
    pts4[(size_t)b * P + p] = make_float4(o[0], o[1], o[2], 0.0f);
  }
  int tx0 = 0, tx1 = -1, ty0 = 0, ty1 = -1;
  const bool ok = (p < P) && tile_range(o[0], o[1], o[2], g, tx0, tx1, ty0, ty1);
  const int ntx = ok ? tx1 - tx0 + 1 : 0;
  const int ntiles = ok ? ntx * (ty1 - ty0 + 1) : 0;
  const int maxt = __reduce_max_sync(FULL, ntiles);
  const int nt2 = g.nt * g.nt;
  int* tc = tile_count + (size_t)b * nt2;
  for (int k = 0; k < maxt; ++k) {
    const bool has = k < ntiles;
    int t = -1 - lane;  // unique per lane: forms a singleton group
    if (has) {
      const int dy = k / ntx;
      t = (ty0 + dy) * g.nt + tx0 + (k - dy * ntx);
    }
    const unsigned grp = __match_any_sync(FULL, t);
    const int leader = __ffs(grp) - 1;
    int base = 0;
    if (has && lane == leader) base = atomicAdd(tc + t, __popc(grp));
    base = __shfl_sync(FULL, base, leader);
    if (has) {
      const int slot = base + __popc(grp & ((1u << lane) - 1u));
      if (slot < CAPG) list[((size_t)b * nt2 + t) * CAPG + slot] = p;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// stage 3b-4: per-tile fine rasterisation + compositing
// ------------------------------------------------------------------------------------------------
struct FineParams {
  const float4* pts4;  // (B,P) x, y, z, 0 in project_pts' frame
  const float* feat;   // (B,C,P)
  const int* tile_count;
  const int* list;  // (B, nt2, CAPG)
  int P, C, S, K, nt;
  float r2;         // (float)radius * (float)radius
  float denom;      // (float)pow(radius, rad_pow)
  float inv_denom;  // 1/denom when that is exact (denom a power of two), else 0
  float tau;
  int accumulation;
  BinGeom g;       // for the whole-cloud rescan of tiles whose list overflowed
  float* out;      // (B,C,S,S)
  uint8_t* empty;  // (B,S,S) 1 where the pixel received no point
  int32_t* idx;    // (B,S,S,K) or null
  float* zbuf;     // (B,S,S,K) or null
  float* dist2;    // (B,S,S,K) or null
  int* ovf_count;  // tiles that exceeded CAP
  int* ovf_list;
};

// alpha = (1 - clamp(dist2 / r^rad_pow, 1e-3, 1)^0.5)^tau   (z_buffer_layers.py:89-98).  Feeds only the
// composited image (tolerance 2e-6), so the square root may be the approximate one.
__device__ __forceinline__ float alpha_of(float d2, const FineParams& q) {
  float d = (q.inv_denom != 0.0f) ? d2 * q.inv_denom : __fdiv_rn(d2, q.denom);
  d = fminf(fmaxf(d, 1e-3f), 1.0f);
  float s;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(d));
  float a = 1.0f - s;
  if (q.tau != 1.0f) a = powf(a, q.tau);
  return a;
}

// tau == 1 and an exact reciprocal of r^rad_pow
__device__ __forceinline__ float alpha_fast(float d2, float inv_denom) {
  const float d = fminf(fmaxf(d2 * inv_denom, 1e-3f), 1.0f);
  float s;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(d));
  return 1.0f - s;
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

struct SortScratch {
  unsigned hist[NB + 4];        // bucket counts -> exclusive starts; hist[NB] = n
  unsigned long long key[CAP];  // (z bits << 32) | point id, in bucket order
};
struct RasterScratch {
  unsigned bits[NPIX][BSTRIDE];         // per pixel: hit mask over the sorted candidates, 32 per word
  unsigned short lists[NPIX][LSTRIDE];  // per pixel: BYTE offset (candidate index * 16) of output slot k
};
// The sorted candidates as three arrays, each holding exactly what one stage reads, because the kernel is bound by
// shared-memory wavefronts (ncu: L1 data pipe 87 % busy, half of it bank conflicts of 16-byte-stride gathers):
//   xyrg  x, y (already negated, z_buffer_layers.py:71-72) + feature channels 0, 1   -> ballots (x, y), compositing
//   zid   z bits + PACKED index b*P+p (what the idx map stores); [CAP] = sentinel     -> map writer only
//   bw    feature channel 2 (NC = 3: 4-byte stride) or channels 2, 3 (NC = 4)         -> compositing
// A slot-list entry is the candidate's byte offset in xyrg; zid is at half, bw at a quarter / half of it.
struct FineSmem {
  float4 xyrg[CAP];
  uint2 zid[CAP + 2];  // [CAP]: z = -1, id = -1, what the padded list entries point at
  float bw[2 * CAP];
  union {
    SortScratch s;
    RasterScratch r;
  } u;
  float ndcx[TILE], ndcy[TILE];  // NDC centres of the tile's pixel columns / rows
  unsigned zmin, zmax, maxcount;
  unsigned warp_tot[FTPB / 32];
  unsigned long long list_bar;  // mbarrier of the bulk copy that stages the tile's candidate list (into bw, free until the scatter)
};

__device__ __forceinline__ float4 gather_feat(const float* __restrict__ featb, int C, int P, int pid) {
  float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
  f.x = __ldg(featb + pid);
  if (C > 1) f.y = __ldg(featb + (size_t)P + pid);
  if (C > 2) f.z = __ldg(featb + 2 * (size_t)P + pid);
  if (C > 3) f.w = __ldg(featb + 3 * (size_t)P + pid);
  return f;
}

__device__ __forceinline__ unsigned short list_entry(int ci) { return (unsigned short)(ci << 4); }
__device__ __forceinline__ float4 ld_xyrg(const FineSmem& sm, unsigned e) {
  return *reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(sm.xyrg) + e);
}
__device__ __forceinline__ float2 ld_xy(const FineSmem& sm, unsigned e) {
  return *reinterpret_cast<const float2*>(reinterpret_cast<const unsigned char*>(sm.xyrg) + e);
}
__device__ __forceinline__ uint2 ld_zid(const FineSmem& sm, unsigned e) {
  return *reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned char*>(sm.zid) + (e >> 1));
}
template <int NC>
__device__ __forceinline__ float2 ld_bw(const FineSmem& sm, unsigned e) {
  if (NC == 3) return make_float2(*reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(sm.bw) + (e >> 2)), 0.f);
  return *reinterpret_cast<const float2*>(reinterpret_cast<const unsigned char*>(sm.bw) + (e >> 1));
}
template <int NC>
__device__ __forceinline__ void st_cand(FineSmem& sm, int i, float x, float y, float z, int packed_id, float4 f) {
  sm.xyrg[i] = make_float4(x, y, f.x, f.y);
  sm.zid[i] = make_uint2(__float_as_uint(z), (unsigned)packed_id);
  if (NC == 3)
    sm.bw[i] = f.z;
  else
    *reinterpret_cast<float2*>(&sm.bw[2 * i]) = make_float2(f.z, f.w);
}

// Stage D when K % 4 == 0 and idx, zbuf are both requested.  Lane l owns slots 4l..4l+3 of every pixel; the warp owns
// pixels warp*16 .. warp*16+15 (two rows of the tile).  Fully unrolled over the 16 pixels; offsets in 16-byte units from
// the warp's first pixel; no dist2 arithmetic unless the dist2 map was asked for.
template <bool D2>
__device__ __forceinline__ void emit_maps(const FineSmem& sm, const FineParams& q, int b, int tx, int ty, int warp,
                                          int lane, int nh, bool inimg) {
  const int K = q.K, S = q.S;
  const int k0 = 4 * lane;
  const int kq = K >> 2;
  const unsigned inmask = __ballot_sync(FULL, inimg);  // bit 2j: pixel j of this warp lies inside the image
  const bool act = k0 < K;
  const size_t o0 = ((((size_t)b * S + (ty * TILE + warp * 2)) * S + tx * TILE) * K + k0) >> 2;
  int4* const pi = reinterpret_cast<int4*>(q.idx) + o0;
  float4* const pz = reinterpret_cast<float4*>(q.zbuf) + o0;
  float4* const pd = D2 ? reinterpret_cast<float4*>(q.dist2) + o0 : nullptr;
  const int rowq = S * kq;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (!((inmask >> (2 * j)) & 1u)) continue;  // warp-uniform
    const int nhj = __shfl_sync(FULL, nh, 2 * j);
    const int pj = warp * 16 + j;
    int4 iv = make_int4(-1, -1, -1, -1);
    float4 zv = make_float4(-1.f, -1.f, -1.f, -1.f), dv = make_float4(-1.f, -1.f, -1.f, -1.f);
    if (k0 < nhj) {
      // entries between nhj and the next multiple of 4 point at the sentinel (id -1, z -1)
      const uint2 L = *reinterpret_cast<const uint2*>(&sm.u.r.lists[pj][k0]);
      const unsigned e0 = L.x & 0xffffu, e1 = L.x >> 16, e2 = L.y & 0xffffu, e3 = L.y >> 16;
      const uint2 c0 = ld_zid(sm, e0), c1 = ld_zid(sm, e1), c2 = ld_zid(sm, e2), c3 = ld_zid(sm, e3);
      iv = make_int4((int)c0.y, (int)c1.y, (int)c2.y, (int)c3.y);
      zv = make_float4(__uint_as_float(c0.x), __uint_as_float(c1.x), __uint_as_float(c2.x), __uint_as_float(c3.x));
      if (D2) {
        const float xfj = sm.ndcx[j & 7], yfj = sm.ndcy[warp * 2 + (j >> 3)];
        const float2 p0 = ld_xy(sm, e0), p1 = ld_xy(sm, e1), p2 = ld_xy(sm, e2), p3 = ld_xy(sm, e3);
        dv.x = dist2_rn(__fsub_rn(p0.x, xfj), __fsub_rn(p0.y, yfj));
        const float d1 = dist2_rn(__fsub_rn(p1.x, xfj), __fsub_rn(p1.y, yfj));
        const float d2 = dist2_rn(__fsub_rn(p2.x, xfj), __fsub_rn(p2.y, yfj));
        const float d3 = dist2_rn(__fsub_rn(p3.x, xfj), __fsub_rn(p3.y, yfj));
        dv.y = (k0 + 1 < nhj) ? d1 : -1.0f;
        dv.z = (k0 + 2 < nhj) ? d2 : -1.0f;
        dv.w = (k0 + 3 < nhj) ? d3 : -1.0f;
      }
    }
    if (act) {
      const int du = (j >> 3) * rowq + (j & 7) * kq;
      __stcs(pi + du, iv);
      __stcs(pz + du, zv);
      if (D2) __stcs(pd + du, dv);
    }
  }
}

// One 128-thread CTA per 8x8-pixel tile.
//   A. load the tile's candidates; counting sort by (z, point id): 1024 buckets over the tile's z range,
//      exact rank inside the (small) buckets; a bitonic sort is the fallback for degenerate z distributions.
//   B. each warp takes 32 sorted candidates at a time and ballots the exact membership test against the 64
//      pixels: pixel p gets, per candidate block, a 32-bit hit word whose set bits are already in output order.
//      Every lane sees all 64 ballots; a select tree over the lane's own bits leaves lane l holding the words of
//      pixels l and l + 32, so a block costs two conflict-free stores instead of 64 single-lane ones.
//   C. two adjacent lanes per pixel: each expands half of the pixel's hit words into the slot list, then
//      composites half of the slots front to back; the halves combine as acc0 + T0 * acc1.
//   D. one warp per pixel row of K slots: gather (z, id) of the lane's four slots, 16-byte streaming stores.
// FAST: alpha compositing with tau == 1, an exact reciprocal of r^rad_pow and at most NC <= 4 feature channels
// (the reference's shipped configuration is NC = 3); otherwise the same code with the general formulas (NC = 4, further
// channels re-gathered from global memory).
template <bool FAST, int NC>
__global__ void __launch_bounds__(FTPB, 6) fine_kernel(FineParams q) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FineSmem& sm = *reinterpret_cast<FineSmem*>(smem_raw);
  constexpr int PER = CAP / FTPB;
  constexpr int NWARP = FTPB / 32;

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

  // ---- A. load + sort ----
  // The tile's candidate list (a contiguous run of point ids, at most 2 KB) is staged with ONE bulk asynchronous copy
  // (cp.async.bulk, the descriptor-less TMA path), issued first so that it flies while the CTA clears its histogram and
  // builds its NDC tables; the point records themselves are gathers by id, which no bulk copy expresses.
  int* ids_s = reinterpret_cast<int*>(sm.bw);
  uint64_t* lbar = reinterpret_cast<uint64_t*>(&sm.list_bar);
  if (tid == 0) {
    mbar_init(lbar, 1);
    mbar_fence_init();
    if (n > 0) {
      const uint32_t bytes = (uint32_t)(n * 4 + 15) & ~15u;
      mbar_expect_tx(lbar, bytes);
      bulk_load(ids_s, q.list + ((size_t)b * nt2 + t) * CAPG, bytes, lbar);
    }
  }
  for (int i = tid; i < (NB + 4) / 4; i += FTPB) reinterpret_cast<uint4*>(sm.u.s.hist)[i] = make_uint4(0, 0, 0, 0);
  if (tid < TILE) sm.ndcx[tid] = pix_to_ndc(S - 1 - (tx * TILE + tid), S);
  if (tid >= 32 && tid < 32 + TILE) sm.ndcy[tid - 32] = pix_to_ndc(S - 1 - (ty * TILE + tid - 32), S);
  if (tid == 65) sm.zid[CAP] = make_uint2(__float_as_uint(-1.0f), 0xffffffffu);
  if (tid == 64) {
    sm.zmin = 0xffffffffu;
    sm.zmax = 0u;
    sm.maxcount = 0u;
  }
  const int32_t idbase = (int32_t)((size_t)b * P);  // stored candidate ids are packed: b*P + p
  const float4* p4 = q.pts4 + (size_t)b * P;
  __syncthreads();  // the list barrier is initialised (and armed) before anybody waits on it
  if (n > 0) mbar_wait(lbar, 0);
  int id[PER];
  float cx[PER], cy[PER], cz[PER];
  unsigned zb[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = tid + j * FTPB;
    id[j] = (i < n) ? ids_s[i] : 0;
  }
  unsigned lmin = 0xffffffffu, lmax = 0u;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = tid + j * FTPB;
    if (i < n) {
      const float4 v = __ldg(p4 + id[j]);
      cx[j] = v.x;
      cy[j] = v.y;
      cz[j] = v.z;
      zb[j] = __float_as_uint(v.z + 0.0f);  // -0 -> +0: the bit pattern of z >= 0 orders like the float
      lmin = min(lmin, zb[j]);
      lmax = max(lmax, zb[j]);
    }
  }
  lmin = __reduce_min_sync(FULL, lmin);
  lmax = __reduce_max_sync(FULL, lmax);
  __syncthreads();
  if (lane == 0 && n > 0) {
    atomicMin(&sm.zmin, lmin);
    atomicMax(&sm.zmax, lmax);
  }
  __syncthreads();
  const unsigned zmin = sm.zmin;
  const unsigned range = sm.zmax - zmin;
  const int sh = (n > 0 && range >= (unsigned)NB) ? (32 - __clz(range) - 10) : 0;
  unsigned bo[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = tid + j * FTPB;
    if (i < n) {
      const unsigned bk = (zb[j] - zmin) >> sh;
      bo[j] = (bk << 16) | atomicAdd(&sm.u.s.hist[bk], 1u);
    }
  }
  __syncthreads();
  {  // exclusive scan of the NB bucket counts; each thread owns NB/FTPB consecutive buckets
    constexpr int BPT = NB / FTPB;
    unsigned v[BPT];
    unsigned tot = 0, mx = 0;
#pragma unroll
    for (int e = 0; e < BPT; e += 4) {
      const uint4 w = *reinterpret_cast<const uint4*>(&sm.u.s.hist[tid * BPT + e]);
      v[e] = w.x;
      v[e + 1] = w.y;
      v[e + 2] = w.z;
      v[e + 3] = w.w;
    }
#pragma unroll
    for (int e = 0; e < BPT; ++e) {
      const unsigned c = v[e];
      mx = max(mx, c);
      v[e] = tot;
      tot += c;
    }
    unsigned incl = tot;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned u = __shfl_up_sync(FULL, incl, off);
      if (lane >= off) incl += u;
    }
    mx = __reduce_max_sync(FULL, mx);
    if (lane == 31) sm.warp_tot[warp] = incl;
    if (lane == 0) atomicMax(&sm.maxcount, mx);
    __syncthreads();
    unsigned base = incl - tot;
#pragma unroll
    for (int w = 0; w < NWARP - 1; ++w)
      if (w < warp) base += sm.warp_tot[w];
#pragma unroll
    for (int e = 0; e < BPT; e += 4)
      *reinterpret_cast<uint4*>(&sm.u.s.hist[tid * BPT + e]) =
          make_uint4(base + v[e], base + v[e + 1], base + v[e + 2], base + v[e + 3]);
    if (tid == 0) sm.u.s.hist[NB] = (unsigned)n;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = tid + j * FTPB;
    if (i < n)
      sm.u.s.key[sm.u.s.hist[bo[j] >> 16] + (bo[j] & 0xffffu)] =
          ((unsigned long long)zb[j] << 32) | (unsigned)id[j];
  }
  __syncthreads();
  const float* featb = q.feat + (size_t)b * C * P;
  if (sm.maxcount <= (unsigned)MAXBK) {
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = tid + j * FTPB;
      if (i < n) {
        const unsigned bk = bo[j] >> 16;
        const unsigned s0 = sm.u.s.hist[bk], c = sm.u.s.hist[bk + 1] - s0;
        unsigned rank = 0;
        if (c > 1) {
          const unsigned long long mine = ((unsigned long long)zb[j] << 32) | (unsigned)id[j];
          for (unsigned m = 0; m < c; ++m) rank += (sm.u.s.key[s0 + m] < mine) ? 1u : 0u;
        }
        st_cand<NC>(sm, (int)(s0 + rank), -cx[j], -cy[j], cz[j], idbase + id[j], gather_feat(featb, C, P, id[j]));
      }
    }
  } else {  // degenerate z distribution (e.g. constant depth): full sort of the 64-bit keys
    int N = 2;
    while (N < n) N <<= 1;
    for (int i = n + tid; i < N; i += FTPB) sm.u.s.key[i] = ~0ull;
    __syncthreads();
    bitonic_sort_u64(sm.u.s.key, N, tid, FTPB);
    for (int i = tid; i < n; i += FTPB) {
      const int pid = (int)(unsigned)(sm.u.s.key[i] & 0xffffffffull);
      const float4 v = __ldg(p4 + pid);
      st_cand<NC>(sm, i, -v.x, -v.y, v.z, idbase + pid, gather_feat(featb, C, P, pid));
    }
  }
  __syncthreads();  // candidates complete; sort scratch is dead from here (aliased by bits/lists)

  // ---- B. membership ballots, 32 sorted candidates x 64 pixels per step ----
  const int nblk = (n + 31) >> 5;
  {
    float xfc[TILE], yfr[TILE];
#pragma unroll
    for (int c = 0; c < TILE; ++c) {
      xfc[c] = sm.ndcx[c];
      yfr[c] = sm.ndcy[c];
    }
    const bool l0 = lane & 1, l1 = lane & 2, l2 = lane & 4, l3 = lane & 8, l4 = lane & 16;
    for (int blk = warp; blk < nblk; blk += NWARP) {
      const int ci = blk * 32 + lane;
      float2 pxy = make_float2(1e30f, 1e30f);  // lanes past the end never hit
      if (ci < n) pxy = *reinterpret_cast<const float2*>(&sm.xyrg[ci]);
      float dx2[TILE], dy2[TILE];
#pragma unroll
      for (int c = 0; c < TILE; ++c) {
        const float dx = __fsub_rn(pxy.x, xfc[c]), dy = __fsub_rn(pxy.y, yfr[c]);
        dx2[c] = __fmul_rn(dx, dx);
        dy2[c] = __fmul_rn(dy, dy);
      }
      // rowsel[r] = the ballot of pixel (r, lane & 7); then rows lane >> 3 (and + 4) are picked the same way
      unsigned rowsel[TILE];
#pragma unroll
      for (int r = 0; r < TILE; ++r) {
        unsigned v[TILE];
#pragma unroll
        for (int c = 0; c < TILE; ++c) v[c] = __ballot_sync(FULL, __fadd_rn(dx2[c], dy2[r]) < q.r2);
        const unsigned a0 = l0 ? v[1] : v[0], a1 = l0 ? v[3] : v[2], a2 = l0 ? v[5] : v[4], a3 = l0 ? v[7] : v[6];
        const unsigned b0 = l1 ? a1 : a0, b1 = l1 ? a3 : a2;
        rowsel[r] = l2 ? b1 : b0;
      }
      const unsigned h0 = l3 ? rowsel[1] : rowsel[0], h1 = l3 ? rowsel[3] : rowsel[2];
      const unsigned h2 = l3 ? rowsel[5] : rowsel[4], h3 = l3 ? rowsel[7] : rowsel[6];
      sm.u.r.bits[lane][blk] = l4 ? h1 : h0;       // pixel index lane      = row lane >> 3,       column lane & 7
      sm.u.r.bits[lane + 32][blk] = l4 ? h3 : h2;  // pixel index lane + 32 = row (lane >> 3) + 4, column lane & 7
    }
  }
  __syncthreads();

  // ---- C. two lanes per pixel: slot list, then front-to-back compositing ----
  const int pix = tid >> 1, half = tid & 1;
  const int xi = tx * TILE + (pix & 7), yi = ty * TILE + (pix >> 3);
  const bool inimg = xi < S && yi < S;
  const float xf = sm.ndcx[pix & 7], yf = sm.ndcy[pix >> 3];
  int nh;
  {
    // lane 0 expands words [0, mid), lane 1 words [mid, nblk) starting at the hit count of the first half
    const int mid = (nblk + 1) >> 1;
    int first = 0;
    for (int blk = 0; blk < mid; ++blk) first += __popc(sm.u.r.bits[pix][blk]);
    int pos = half ? first : 0;
    const int b0 = half ? mid : 0, b1 = half ? nblk : mid;
    for (int blk = b0; blk < b1 && pos < K; ++blk) {
      unsigned w = sm.u.r.bits[pix][blk];
      while (w && pos < K) {
        const int bpos = __ffs(w) - 1;
        w &= w - 1;
        sm.u.r.lists[pix][pos++] = list_entry(blk * 32 + bpos);
      }
    }
    nh = min(__shfl_sync(FULL, pos, lane | 1), K);  // lane 1 ends at the pixel's total hit count
    if (half)
      for (int e = nh; e < ((nh + 3) & ~3); ++e) sm.u.r.lists[pix][e] = list_entry(CAP);
  }
  __syncwarp();
  {
    const int m2 = (nh + 1) >> 1;
    const int k0 = half ? m2 : 0, k1 = half ? nh : m2;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f, T = 1.0f, wsum = 0.0f;
    const bool ac = FAST || q.accumulation == PS_ACCUM_ALPHACOMPOSITE;
    const unsigned short* lp = sm.u.r.lists[pix];
    int k = k0;
    for (; k + 1 < k1; k += 2) {  // two hits per trip: independent loads and alphas, serial transmittance
      const unsigned e0 = lp[k], e1 = lp[k + 1];
      const float4 c0 = ld_xyrg(sm, e0), c1 = ld_xyrg(sm, e1);
      const float2 g0 = ld_bw<NC>(sm, e0), g1 = ld_bw<NC>(sm, e1);
      const float d20 = dist2_rn(__fsub_rn(c0.x, xf), __fsub_rn(c0.y, yf));
      const float d21 = dist2_rn(__fsub_rn(c1.x, xf), __fsub_rn(c1.y, yf));
      const float a0 = FAST ? alpha_fast(d20, q.inv_denom) : alpha_of(d20, q);
      const float a1 = FAST ? alpha_fast(d21, q.inv_denom) : alpha_of(d21, q);
      float w0 = a0, w1 = a1;
      if (ac) {
        w0 = T * a0;
        T *= 1.0f - a0;
        w1 = T * a1;
        T *= 1.0f - a1;
      } else {
        wsum += a0 + a1;
      }
      acc0 = fmaf(w1, c1.z, fmaf(w0, c0.z, acc0));
      acc1 = fmaf(w1, c1.w, fmaf(w0, c0.w, acc1));
      acc2 = fmaf(w1, g1.x, fmaf(w0, g0.x, acc2));
      if (NC > 3) acc3 = fmaf(w1, g1.y, fmaf(w0, g0.y, acc3));
    }
    if (k < k1) {
      const unsigned e = lp[k];
      const float4 c = ld_xyrg(sm, e);
      const float2 g = ld_bw<NC>(sm, e);
      const float d2 = dist2_rn(__fsub_rn(c.x, xf), __fsub_rn(c.y, yf));
      const float a = FAST ? alpha_fast(d2, q.inv_denom) : alpha_of(d2, q);
      float wgt = a;
      if (ac) {
        wgt = T * a;
        T *= 1.0f - a;
      } else {
        wsum += a;
      }
      acc0 = fmaf(wgt, c.z, acc0);
      acc1 = fmaf(wgt, c.w, acc1);
      acc2 = fmaf(wgt, g.x, acc2);
      if (NC > 3) acc3 = fmaf(wgt, g.y, acc3);
    }
    // combine the halves: second-half terms are attenuated by the first half's transmittance
    const float T0 = __shfl_sync(FULL, T, lane & ~1);
    const float sc = (half && ac) ? T0 : 1.0f;
    acc0 *= sc;
    acc1 *= sc;
    acc2 *= sc;
    acc3 *= sc;
    acc0 += __shfl_xor_sync(FULL, acc0, 1);
    acc1 += __shfl_xor_sync(FULL, acc1, 1);
    acc2 += __shfl_xor_sync(FULL, acc2, 1);
    acc3 += __shfl_xor_sync(FULL, acc3, 1);
    wsum += __shfl_xor_sync(FULL, wsum, 1);
    if (inimg && half == 0) {
      const bool dv = !FAST && q.accumulation == PS_ACCUM_WSUMNORM;
      const float norm = dv ? fmaxf(wsum, 1e-4f) : 1.0f;
      float* o = q.out + ((size_t)b * C * S + yi) * S + xi;
      const size_t cs = (size_t)S * S;
      o[0] = dv ? acc0 / norm : acc0;
      if (C > 1) o[cs] = dv ? acc1 / norm : acc1;
      if (C > 2) o[2 * cs] = dv ? acc2 / norm : acc2;
      if (NC > 3 && C > 3) o[3 * cs] = dv ? acc3 / norm : acc3;
      q.empty[((size_t)b * S + yi) * S + xi] = (nh == 0);
      if (!FAST) {
        // feature widths beyond 4 (non-RGB feature splats): channels re-gathered from global memory
        for (int c0 = 4; c0 < C; ++c0) {
          const float* fc = featb + (size_t)c0 * P;
          float acc = 0.f, Tc = 1.0f;
          for (int k = 0; k < nh; ++k) {
            const unsigned e = sm.u.r.lists[pix][k];
            const float2 pxy = ld_xy(sm, e);
            const float a = alpha_of(dist2_rn(__fsub_rn(pxy.x, xf), __fsub_rn(pxy.y, yf)), q);
            float wgt = a;
            if (ac) {
              wgt = Tc * a;
              Tc *= 1.0f - a;
            }
            acc += wgt * __ldg(fc + ((int)ld_zid(sm, e).y - idbase));
          }
          o[c0 * cs] = dv ? acc / norm : acc;
        }
      }
    }
  }

  // ---- D. maps: lane l owns output slots 4l..4l+3 of the current pixel; a warp streams its 16 pixels.
  // List entries between a pixel's hit count and the next multiple of 4 point at the sentinel (id -1, z -1), so a
  // lane either gathers four entries unconditionally or stores the -1 padding.
  if (q.idx && q.zbuf && (K & 3) == 0) {
    if (q.dist2)
      emit_maps<true>(sm, q, b, tx, ty, warp, lane, nh, inimg);
    else
      emit_maps<false>(sm, q, b, tx, ty, warp, lane, nh, inimg);
  } else if (q.idx || q.zbuf || q.dist2) {  // any subset of the maps, any K
    const int k0 = 4 * lane;
    for (int j = 0; j < 16; ++j) {
      if (!__shfl_sync(FULL, (int)inimg, 2 * j)) continue;
      const int pj = warp * 16 + j;
      const int nhj = __shfl_sync(FULL, nh, 2 * j);
      const float xfj = sm.ndcx[pj & 7], yfj = sm.ndcy[pj >> 3];
      const size_t o = (((size_t)b * S + (ty * TILE + (pj >> 3))) * S + tx * TILE + (pj & 7)) * K;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + e;
        if (k >= K) break;
        int idv = -1;
        float zv = -1.0f, dv = -1.0f;
        if (k < nhj) {
          const unsigned en = sm.u.r.lists[pj][k];
          const uint2 c = ld_zid(sm, en);
          const float2 pxy = ld_xy(sm, en);
          idv = (int)c.y;
          zv = __uint_as_float(c.x);
          dv = dist2_rn(__fsub_rn(pxy.x, xfj), __fsub_rn(pxy.y, yfj));
        }
        if (q.idx) q.idx[o + k] = idv;
        if (q.zbuf) q.zbuf[o + k] = zv;
        if (q.dist2) q.dist2[o + k] = dv;
      }
    }
  }
}

// Overflow path: tiles with more than CAP candidates.  Streams the candidates in chunks of CAPB (from the
// tile's list, or by rescanning the whole cloud when even the list overflowed), sorts each chunk, and merges
// the chunk's hits (already ordered) into each pixel's running top-K.  Slow, exact, never drops a point.
constexpr int BTH = 256;  // threads of the overflow kernel: all sort / fill / write, the first NPIX each own a pixel's top-K
struct BigSmem {
  unsigned long long key[CAPB];
  float2 xy[CAPB];
  unsigned long long topA[MAXK][TPB];
  unsigned long long topB[MAXK][TPB];
  int id4[BTH / 32][MAXK];
  float z4[BTH / 32][MAXK];
  float d4[BTH / 32][MAXK];
  float w4[BTH / 32][MAXK];
  int cnt[TPB];   // hits kept per pixel
  int fill;
};

__global__ void __launch_bounds__(BTH) fine_big_kernel(FineParams q) {
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
    const bool rescan = n > CAPG;  // the list is incomplete: derive the candidates from the whole cloud
    const int total = rescan ? P : n;
    const bool owner = tid < TPB;  // this thread merges pixel `tid` of the tile
    const int xi = tx * TILE + (tid & 7), yi = ty * TILE + ((tid >> 3) & 7);
    const float xf = pix_to_ndc(S - 1 - xi, S), yf = pix_to_ndc(S - 1 - yi, S);
    const float4* p4 = q.pts4 + (size_t)b * P;
    const int* lst = q.list + (size_t)bt * CAPG;
    unsigned long long(*A)[TPB] = sm.topA;
    unsigned long long(*Bv)[TPB] = sm.topB;
    int cntA = 0;
    for (int c0 = 0; c0 < total;) {
      // fill the chunk: the next CAPB list entries, or (rescan) the next points of the cloud whose tile range covers
      // this tile, compacted -- the order inside a chunk is irrelevant (it is sorted below), and so is the chunking
      // (the merge keeps the K smallest keys)
      int m;
      __syncthreads();
      if (!rescan) {
        m = min(CAPB, total - c0);
        for (int i = tid; i < m; i += BTH) {
          const int p = lst[c0 + i];
          sm.key[i] = ((unsigned long long)__float_as_uint(p4[p].z + 0.0f) << 32) | (unsigned)p;
        }
        c0 += m;
      } else {
        if (tid == 0) sm.fill = 0;
        __syncthreads();
        m = 0;
        while (c0 < total && m <= CAPB - 2 * BTH) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int p = c0 + e * BTH + tid;
            bool take = false;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < total) {
              v = p4[p];
              int a0, a1, b0, b1;
              take = tile_range(v.x, v.y, v.z, q.g, a0, a1, b0, b1) && tx >= a0 && tx <= a1 && ty >= b0 && ty <= b1;
            }
            const unsigned bal = __ballot_sync(FULL, take);
            int base = 0;
            if (lane == 0 && bal) base = atomicAdd(&sm.fill, __popc(bal));
            base = __shfl_sync(FULL, base, 0);
            if (take)
              sm.key[base + __popc(bal & ((1u << lane) - 1))] =
                  ((unsigned long long)__float_as_uint(v.z + 0.0f) << 32) | (unsigned)p;
          }
          c0 += 2 * BTH;
          __syncthreads();
          m = sm.fill;
        }
        if (m == 0) continue;  // (only at the end of the cloud)
      }
      int N = 2;
      while (N < m) N <<= 1;
      for (int i = m + tid; i < N; i += BTH) sm.key[i] = ~0ull;
      __syncthreads();
      bitonic_sort_u64(sm.key, N, tid, BTH);
      for (int i = tid; i < m; i += BTH) {
        const float4 v = p4[(int)(unsigned)(sm.key[i] & 0xffffffffull)];
        sm.xy[i] = make_float2(-v.x, -v.y);
      }
      __syncthreads();
      if (owner) {
        // merge: A (sorted, cntA) with this chunk's hits (sorted) -> Bv, keeping the K smallest
        int ia = 0, nb = 0;
        for (int i = 0; i < m; ++i) {
          const float2 c = sm.xy[i];
          const float d2 = dist2_rn(__fsub_rn(c.x, xf), __fsub_rn(c.y, yf));
          if (d2 < q.r2 && nb < K) {
            const unsigned long long key = sm.key[i];
            while (ia < cntA && nb < K && A[ia][tid] < key) Bv[nb++][tid] = A[ia++][tid];
            if (nb < K) Bv[nb++][tid] = key;
          }
        }
        while (ia < cntA && nb < K) Bv[nb++][tid] = A[ia++][tid];
        cntA = nb;
      }
      unsigned long long(*tmp)[TPB] = A;
      A = Bv;
      Bv = tmp;
    }
    if (owner) {
      sm.cnt[tid] = cntA;
      if (xi < S && yi < S) q.empty[((size_t)b * S + yi) * S + xi] = (cntA == 0);
    }
    __syncthreads();
    // every warp writes NPIX / 8 pixels: one lane per output slot; coordinates and features from global memory
    const int rounds = (K + 31) >> 5;
    const int32_t base = (int32_t)((size_t)b * P);
    for (int pix = warp; pix < TPB; pix += BTH / 32) {
      const int pxi = tx * TILE + (pix & 7), pyi = ty * TILE + (pix >> 3);
      if (pxi >= S || pyi >= S) continue;
      const int nh = sm.cnt[pix];
      const float xfj = pix_to_ndc(S - 1 - pxi, S), yfj = pix_to_ndc(S - 1 - pyi, S);
      float tcarry = 1.0f, wsum = 0.0f;
      // the lane's (up to MAXK / 32) candidates first: independent loads, one memory latency instead of one per round
      int pp[MAXK / 32];
      float4 pv[MAXK / 32];
#pragma unroll
      for (int r = 0; r < MAXK / 32; ++r) {
        const int k = r * 32 + lane;
        pp[r] = (r < rounds && k < nh) ? (int)(unsigned)(A[k][pix] & 0xffffffffull) : -1;
      }
#pragma unroll
      for (int r = 0; r < MAXK / 32; ++r) pv[r] = pp[r] >= 0 ? p4[pp[r]] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < MAXK / 32; ++r) {
        if (r >= rounds) break;
        const int k = r * 32 + lane;
        const bool valid = k < nh;
        float a = 0.0f, z = -1.0f, d2 = -1.0f;
        int pid = -1;
        if (valid) {
          const int p = pp[r];
          const float4 v = pv[r];
          d2 = dist2_rn(__fsub_rn(-v.x, xfj), __fsub_rn(-v.y, yfj));
          z = v.z;
          pid = base + p;
          a = alpha_of(d2, q);
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
          sm.id4[warp][k] = pid;
          sm.z4[warp][k] = z;
          sm.d4[warp][k] = d2;
          sm.w4[warp][k] = valid ? wgt : 0.0f;
        }
      }
      __syncwarp();
      const size_t pixoff = ((size_t)b * S + pyi) * S + pxi;
      for (int k = lane; k < K; k += 32) {
        if (q.idx) q.idx[pixoff * K + k] = sm.id4[warp][k];
        if (q.zbuf) q.zbuf[pixoff * K + k] = sm.z4[warp][k];
        if (q.dist2) q.dist2[pixoff * K + k] = sm.d4[warp][k];
      }
      if (q.accumulation != PS_ACCUM_ALPHACOMPOSITE) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) wsum += __shfl_xor_sync(FULL, wsum, off);
      }
      const float norm = fmaxf(wsum, 1e-4f);
      const int nhk = min(nh, K);
      // lanes share the K slots of one channel (independent loads in flight), then the warp adds the partial sums
      for (int c = 0; c < C; ++c) {
        const float* fc = q.feat + ((size_t)b * C + c) * P - base;
        float v = 0.0f;
        for (int k = lane; k < nhk; k += 32) v += sm.w4[warp][k] * fc[sm.id4[warp][k]];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
        if (q.accumulation == PS_ACCUM_WSUMNORM) v = v / norm;
        if (lane == 0) q.out[(((size_t)b * C + c) * S + pyi) * S + pxi] = v;
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
struct SplatLayout {
  int nt, nt2;
  size_t bytes;
  size_t off_count, off_ovf, off_list, off_empty, off_pts4;
};

static SplatLayout splat_layout(int B, int P, int S) {
  SplatLayout L;
  L.nt = (S + TILE - 1) / TILE;
  L.nt2 = L.nt * L.nt;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    off = align_up(off, 256);
    size_t r = off;
    off += bytes;
    return r;
  };
  L.off_count = take(sizeof(int) * (size_t)B * L.nt2);
  L.off_ovf = take(sizeof(int) * (1 + (size_t)B * L.nt2));  // adjacent to counts: one memset clears both
  L.off_list = take(sizeof(int) * (size_t)B * L.nt2 * CAPG);
  L.off_empty = take((size_t)B * S * S);
  L.off_pts4 = take(sizeof(float4) * (size_t)B * P);
  L.bytes = align_up(off, 256);
  return L;
}

// depth != null: project from depth (forward_justpts); else pts (B,P,3) is the cloud.
static int splat_impl(const float* depth, const float* mats, int W, float eps, const float* pts, const float* feat,
                      int B, int P, int C, int S, int K, double radius_px, double tau, int rad_pow, int accumulation,
                      int bg_ksize, float* out, uint8_t* bg_mask, int32_t* idx, float* zbuf, float* dist2,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  PS_CHECK_ARG((depth && mats) || pts);
  PS_CHECK_ARG((feat || P == 0) && out && bg_mask && workspace);
  PS_CHECK_ARG(B >= 0 && P >= 0 && C >= 1 && S >= 1 && S <= 8192);
  PS_CHECK_ARG(K >= 1);
  if (K > MAXK) return fail(PS_EUNSUPPORTED, "%s: points_per_pixel exceeds PS_MAX_POINTS_PER_PIXEL%s", __func__);
  PS_CHECK_ARG(accumulation >= 0 && accumulation <= 2);
  PS_CHECK_ARG(bg_ksize >= 1 && (bg_ksize & 1) == 1 && bg_ksize / 2 <= BG_MAXH);
  PS_CHECK_ARG(radius_px > 0.0 && radius_px <= 64.0);
  PS_CHECK_ARG((size_t)B * P < 0x7fffffffull);  // packed indices are int32, as in PyTorch3D
  if (B == 0) return PS_OK;
  const SplatLayout L = splat_layout(B, P, S);
  if (workspace_bytes < L.bytes) return fail(PS_EWORKSPACE, "%s: workspace too small%s", __func__);
  char* ws = (char*)workspace;
  int* tile_count = (int*)(ws + L.off_count);
  int* ovf = (int*)(ws + L.off_ovf);
  int* list = (int*)(ws + L.off_list);
  uint8_t* empty = (uint8_t*)(ws + L.off_empty);
  float4* pts4 = (float4*)(ws + L.off_pts4);

  const double radius = radius_px / (double)S * 2.0;  // z_buffer_layers.py:77
  const float rf = (float)radius;
  BinGeom g;
  g.S = S;
  g.nt = L.nt;
  g.half_S = 0.5f * (float)S;
  g.rp = (float)(radius_px + 1.0 / 64.0);
  g.lim = 1.0f + rf + 4.0f / (float)S;

  PS_CUDA(cudaMemsetAsync(tile_count, 0, (L.off_ovf - L.off_count) + sizeof(int), stream));
  if (P > 0) {
    dim3 grid((P + 255) / 256, B);
    ProjectArgs pa{depth, mats, W, eps};
    if (depth)
      bin_kernel<true><<<grid, 256, 0, stream>>>(pa, nullptr, P, g, pts4, tile_count, list);
    else
      bin_kernel<false><<<grid, 256, 0, stream>>>(pa, pts, P, g, pts4, tile_count, list);
    PS_LAUNCHED();
  }
  FineParams q;
  q.pts4 = pts4;
  q.feat = feat;
  q.tile_count = tile_count;
  q.list = list;
  q.P = P;
  q.C = C;
  q.S = S;
  q.K = K;
  q.nt = L.nt;
  q.r2 = rf * rf;
  q.denom = (float)pow(radius, (double)rad_pow);
  {
    int e = 0;
    q.inv_denom = (frexpf(q.denom, &e) == 0.5f) ? 1.0f / q.denom : 0.0f;  // exact reciprocal only
  }
  q.tau = (float)tau;
  q.accumulation = accumulation;
  q.g = g;
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
    PS_CUDA(cudaFuncSetAttribute(fine_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FineSmem)));
    PS_CUDA(cudaFuncSetAttribute(fine_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FineSmem)));
    PS_CUDA(cudaFuncSetAttribute(fine_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FineSmem)));
    PS_CUDA(cudaFuncSetAttribute(fine_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BigSmem)));
    smem_set_for_device = dev;
  }
  {
    dim3 grid(L.nt2, B);
    PS_TIME_BEGIN("fine_kernel", stream);
    const bool fast = accumulation == PS_ACCUM_ALPHACOMPOSITE && q.tau == 1.0f && q.inv_denom != 0.0f && C <= 4;
    if (fast && C <= 3)
      fine_kernel<true, 3><<<grid, FTPB, sizeof(FineSmem), stream>>>(q);
    else if (fast)
      fine_kernel<true, 4><<<grid, FTPB, sizeof(FineSmem), stream>>>(q);
    else
      fine_kernel<false, 4><<<grid, FTPB, sizeof(FineSmem), stream>>>(q);
    PS_TIME_END(stream);
    PS_LAUNCHED();
    PS_TIME_BEGIN("fine_big_kernel", stream);
    fine_big_kernel<<<148, BTH, sizeof(BigSmem), stream>>>(q);
    PS_TIME_END(stream);
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
  (void)radius_px;
  if (B <= 0 || P < 0 || S < 1) return 256;
  return splat_layout(B, P, S).bytes;
}

int ps_splat_points(const float* pts, const float* feat, int B, int P, int C, int S, int K, double radius_px,
                    double tau, int rad_pow, int accumulation, int bg_ksize, float* out, uint8_t* bg_mask,
                    int32_t* idx, float* zbuf, float* dist2, void* workspace, size_t workspace_bytes, void* stream) {
  PS_CHECK_ARG(pts || P == 0 || B == 0);
  static const float dummy = 0.f;
  return splat_impl(nullptr, nullptr, 0, 0.f, pts ? pts : &dummy, feat, B, P, C, S, K, radius_px, tau, rad_pow,
                    accumulation, bg_ksize, out, bg_mask, idx, zbuf, dist2, workspace, workspace_bytes,
                    (cudaStream_t)stream);
}

size_t ps_splat_fwd_workspace_bytes(int B, int W, int S, double radius_px) {
  if (B <= 0 || W < 2) return 256;
  return ps_splat_workspace_bytes(B, W * W, S, radius_px);
}

int ps_splat_fwd(const float* depth, const float* feat, const float* mats, int B, int W, int C, int S, int K,
                 double radius_px, double tau, int rad_pow, int accumulation, int bg_ksize, float eps, float* out,
                 uint8_t* bg_mask, int32_t* idx, float* zbuf, float* dist2, void* workspace, size_t workspace_bytes,
                 void* stream) {
  PS_CHECK_ARG(depth && mats);
  PS_CHECK_ARG(B >= 0 && W >= 2 && W <= 4096);
  return splat_impl(depth, mats, W, eps, nullptr, feat, B, W * W, C, S, K, radius_px, tau, rad_pow, accumulation,
                    bg_ksize, out, bg_mask, idx, zbuf, dist2, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
