/*
 * oracle/splat_oracle.c -- CPU restatement of the reference's z-buffer point splat.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pixelsynth_b200/ may import, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED by the reference: crockwell/pixelsynth ships no tests or golden vectors, and
 * the rasteriser arithmetic lives in PyTorch3D (pinned 0.4.0, docs/INSTALL.md:11; or 0.2.0 @
 * e3819a49, docs/INSTALL.md:59,74), which is not vendored and not installable offline.  What we
 * *can* pin is the reference's own glue: tests/golden/make_splat_golden.py runs the reference's
 * unmodified PtsManipulator.project_pts and RasterizePointsXYsBlending.forward (imported from
 * /root/reference, with `pytorch3d` stubbed by a brute-force numpy restatement of the published
 * algorithm) and this file is checked against those fixtures.
 *
 * What is restated, with the reference lines it follows:
 *   pso_project          models/projection/z_buffer_manipulator.py:38-48 (xyzs grid), :50-83 (project_pts)
 *   pso_rasterize*       models/layers/z_buffer_layers.py:71-72 (negate x,y), :77 (radius), :81-84
 *                        -> pytorch3d.renderer.points.rasterize_points, CPU "naive" semantics:
 *                        keep the K smallest-z points with z >= 0 and dx*dx+dy*dy < r*r, ascending
 *                        (z, packed index); unused slots = -1 (SURVEY.md Appendix A rules 1-4)
 *   pso_composite        z_buffer_layers.py:89-98 (alpha), :112-129 -> pytorch3d compositing
 *                        alpha_composite / weighted_sum / weighted_sum_norm (Appendix A rules 5-6)
 *   pso_bgmask           z_buffer_layers.py:100-110 (13x13 box dilation of "pixel has no point")
 *
 * Canonical arithmetic (shared, bit for bit, with the CUDA kernels): IEEE fp32, round-to-nearest,
 * no FMA contraction (build with -ffp-contract=off), 4-term dot products summed left to right
 * ((a0*b0 + a1*b1) + a2*b2) + a3*b3.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PSO_OK 0
#define PSO_EINVAL -1
#define PSO_ENOMEM -2

static inline float dot4(const float* m, const float* v) {
  float s = m[0] * v[0];
  s = s + m[1] * v[1];
  s = s + m[2] * v[2];
  s = s + m[3] * v[3];
  return s;
}

/* 4x4 * 4x4, row-major, same left-to-right dot rule. */
static void matmul4(const float* a, const float* b, float* c) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = a[i * 4 + 0] * b[0 * 4 + j];
      s = s + a[i * 4 + 1] * b[1 * 4 + j];
      s = s + a[i * 4 + 2] * b[2 * 4 + j];
      s = s + a[i * 4 + 3] * b[3 * 4 + j];
      c[i * 4 + j] = s;
    }
}

/* xs[i] = linspace(0, W-1, W)[i] / (W-1) * 2 - 1   (z_buffer_manipulator.py:38) */
static inline float grid_coord(int i, int W) {
  float t = (float)i / (float)(W - 1);
  t = t * 2.0f;
  return t - 1.0f;
}

/* Transform one homogeneous point already in "xy_proj" form into the sampler triple
 * (z_buffer_manipulator.py:69-81): mask |z|<EPS, z:=EPS, x/-z, y/-z, masked -> -10, flip (1,-1,-1). */
static inline void finish_point(float qx, float qy, float qz, float eps, float* out3) {
  int masked = fabsf(qz) < eps;
  if (masked) qz = eps;
  float nz = -qz;
  float sx = qx / nz, sy = qy / nz, sz = qz;
  if (masked) { sx = -10.0f; sy = -10.0f; sz = -10.0f; }
  out3[0] = sx * 1.0f;
  out3[1] = sy * -1.0f;
  out3[2] = sz * -1.0f;
}

/*
 * depth (B,P) with P = W*W, row-major pixel order p = y*W + x.
 * mats (B,6,16) row-major 4x4: [K, Kinv, RT1, RT1inv, RT2, RT2inv] -- the argument order of
 * PtsManipulator.forward_justpts (z_buffer_manipulator.py:85-87); only K, Kinv, RT1inv, RT2 are read.
 * pts (B,P,3) : the `sampler` returned by project_pts, permuted to point-major as at :103.
 * xyproj (B,4,P) or NULL: the pre-division homogeneous coords (what project_pts_cumulative returns, :266;
 *   note z already has EPS written into masked entries because the reference writes through a view, :73-74).
 */
int pso_project(const float* depth, const float* mats, int B, int W, float eps, float* pts, float* xyproj) {
  if (!depth || !mats || !pts || B < 0 || W < 2) return PSO_EINVAL;
  const int P = W * W;
  for (int b = 0; b < B; ++b) {
    const float* K = mats + (size_t)b * 96 + 0;
    const float* Kinv = mats + (size_t)b * 96 + 16;
    const float* RT1inv = mats + (size_t)b * 96 + 48;
    const float* RT2 = mats + (size_t)b * 96 + 64;
    float RT[16];
    matmul4(RT2, RT1inv, RT);
    for (int p = 0; p < P; ++p) {
      const int sy = p / W, sx = p % W;
      const float d = depth[(size_t)b * P + p];
      float X[4];
      X[0] = grid_coord(sx, W) * d;
      X[1] = (-grid_coord(sy, W)) * d;
      X[2] = -1.0f * d;
      X[3] = 1.0f; /* projected_coors[:, -1, :] = 1 */
      float c[4], w[4], q[4];
      for (int r = 0; r < 4; ++r) c[r] = dot4(Kinv + 4 * r, X);
      for (int r = 0; r < 4; ++r) w[r] = dot4(RT + 4 * r, c);
      for (int r = 0; r < 4; ++r) q[r] = dot4(K + 4 * r, w);
      if (xyproj) {
        float* o = xyproj + (size_t)b * 4 * P;
        o[0 * (size_t)P + p] = q[0];
        o[1 * (size_t)P + p] = q[1];
        o[2 * (size_t)P + p] = (fabsf(q[2]) < eps) ? eps : q[2];
        o[3 * (size_t)P + p] = q[3];
      }
      finish_point(q[0], q[1], q[2], eps, pts + ((size_t)b * P + p) * 3);
    }
  }
  return PSO_OK;
}

/*
 * Prior-cloud branch of project_pts_cumulative (z_buffer_manipulator.py:244-248, 253-264):
 * cloud (B,4,P) homogeneous points in the previous target camera's frame; mats2 (B,3,16) =
 * [K, RT2 (new target), RT3inv (previous target inverse)]; outputs as pso_project.
 */
int pso_project_cloud(const float* cloud, const float* mats3, int B, int P, float eps, float* pts, float* xyproj) {
  if (!cloud || !mats3 || !pts || B < 0 || P < 0) return PSO_EINVAL;
  for (int b = 0; b < B; ++b) {
    const float* K = mats3 + (size_t)b * 48 + 0;
    const float* RT2 = mats3 + (size_t)b * 48 + 16;
    const float* RT3inv = mats3 + (size_t)b * 48 + 32;
    float RT[16];
    matmul4(RT2, RT3inv, RT);
    const float* cb = cloud + (size_t)b * 4 * P;
    for (int p = 0; p < P; ++p) {
      float X[4] = {cb[p], cb[(size_t)P + p], cb[2 * (size_t)P + p], cb[3 * (size_t)P + p]};
      float w[4], q[4];
      for (int r = 0; r < 4; ++r) w[r] = dot4(RT + 4 * r, X);
      for (int r = 0; r < 4; ++r) q[r] = dot4(K + 4 * r, w);
      if (xyproj) {
        float* o = xyproj + (size_t)b * 4 * P;
        o[0 * (size_t)P + p] = q[0];
        o[1 * (size_t)P + p] = q[1];
        o[2 * (size_t)P + p] = (fabsf(q[2]) < eps) ? eps : q[2];
        o[3 * (size_t)P + p] = q[3];
      }
      finish_point(q[0], q[1], q[2], eps, pts + ((size_t)b * P + p) * 3);
    }
  }
  return PSO_OK;
}

/* PyTorch3D PixToNdc: -1 + (2*i + 1) / S, evaluated in fp32. */
static inline float pix_to_ndc(int i, int S) { return -1.0f + (2.0f * (float)i + 1.0f) / (float)S; }

typedef struct { float z; int32_t idx; float d2; } hit_t;

static int hit_cmp(const void* a, const void* b) {
  const hit_t* x = (const hit_t*)a;
  const hit_t* y = (const hit_t*)b;
  if (x->z < y->z) return -1;
  if (x->z > y->z) return 1;
  return (x->idx > y->idx) - (x->idx < y->idx);
}

/* membership test exactly as PyTorch3D's naive CPU rasteriser, in the frame after the layer's
 * negation of x,y (z_buffer_layers.py:71-72) and PyTorch3D's reversed pixel index. */
static inline int test_point(const float* pt, float xf, float yf, float r2, float* d2out) {
  const float px = -pt[0], py = -pt[1], pz = pt[2];
  if (!(pz >= 0.0f)) return 0; /* pz < 0 -> skip; NaN never contributes */
  const float dx = px - xf, dy = py - yf;
  const float d2 = dx * dx + dy * dy;
  if (d2 < r2) { *d2out = d2; return 1; }
  return 0;
}

static void emit_pixel(hit_t* hits, int n, int K, int32_t base, int32_t* idx, float* zbuf, float* dist2) {
  qsort(hits, (size_t)n, sizeof(hit_t), hit_cmp);
  for (int k = 0; k < K; ++k) {
    if (k < n) {
      idx[k] = base + hits[k].idx;
      zbuf[k] = hits[k].z;
      if (dist2) dist2[k] = hits[k].d2;
    } else {
      idx[k] = -1;
      zbuf[k] = -1.0f;
      if (dist2) dist2[k] = -1.0f;
    }
  }
}

/*
 * Brute force: every pixel tests every point.  O(S*S*P); use for S <= 64.
 * pts (B,P,3) in project_pts' frame; radius = float(opt.radius)/float(S)*2.0 (z_buffer_layers.py:77)
 * passed as the Python double it is there.  Outputs (B,S,S,K); idx holds packed indices b*P + p.
 */
int pso_rasterize_naive(const float* pts, int B, int P, int S, int K, double radius, int32_t* idx, float* zbuf,
                        float* dist2) {
  if (!pts || !idx || !zbuf || B < 0 || P < 0 || S < 1 || K < 1) return PSO_EINVAL;
  const float rf = (float)radius;
  const float r2 = rf * rf;
  hit_t* hits = (hit_t*)malloc(sizeof(hit_t) * (size_t)(P > 0 ? P : 1));
  if (!hits) return PSO_ENOMEM;
  for (int b = 0; b < B; ++b)
    for (int yi = 0; yi < S; ++yi) {
      const float yf = pix_to_ndc(S - 1 - yi, S);
      for (int xi = 0; xi < S; ++xi) {
        const float xf = pix_to_ndc(S - 1 - xi, S);
        int n = 0;
        for (int p = 0; p < P; ++p) {
          float d2;
          const float* pt = pts + ((size_t)b * P + p) * 3;
          if (test_point(pt, xf, yf, r2, &d2)) { hits[n].z = pt[2]; hits[n].idx = p; hits[n].d2 = d2; ++n; }
        }
        const size_t o = (((size_t)b * S + yi) * S + xi) * K;
        emit_pixel(hits, n, K, (int32_t)((size_t)b * P), idx + o, zbuf + o, dist2 ? dist2 + o : NULL);
      }
    }
  free(hits);
  return PSO_OK;
}

/*
 * Same result as pso_rasterize_naive, but points are first bucketed into T x T pixel tiles by a
 * conservative bounding box so 256x256 finishes in well under a second.  The per-pixel test and
 * ordering are the same code; the bucketing only prunes points that cannot pass the test.
 */
int pso_rasterize(const float* pts, int B, int P, int S, int K, double radius, int32_t* idx, float* zbuf,
                  float* dist2) {
  if (!pts || !idx || !zbuf || B < 0 || P < 0 || S < 1 || K < 1) return PSO_EINVAL;
  const int T = 8;
  const int nt = (S + T - 1) / T;
  const float rf = (float)radius;
  const float r2 = rf * rf;
  const double rp = radius * 0.5 * (double)S + 1.0; /* radius in pixels, +1 px safety margin */
  int* count = (int*)malloc(sizeof(int) * (size_t)(nt * nt + 1));
  int* start = (int*)malloc(sizeof(int) * (size_t)(nt * nt + 1));
  int* list = NULL;
  size_t list_cap = 0;
  hit_t* hits = (hit_t*)malloc(sizeof(hit_t) * (size_t)(P > 0 ? P : 1));
  if (!count || !start || !hits) { free(count); free(start); free(hits); return PSO_ENOMEM; }
  for (int b = 0; b < B; ++b) {
    const float* pb = pts + (size_t)b * P * 3;
    for (int pass = 0; pass < 2; ++pass) {
      memset(count, 0, sizeof(int) * (size_t)(nt * nt + 1));
      for (int p = 0; p < P; ++p) {
        const float* pt = pb + (size_t)p * 3;
        if (!(pt[2] >= 0.0f)) continue;
        /* pixel (xi, yi) has centre cx = -1 + (2*xi+1)/S in the un-negated frame (Appendix A rule 1) */
        const double fx = ((double)pt[0] + 1.0) * 0.5 * (double)S - 0.5;
        const double fy = ((double)pt[1] + 1.0) * 0.5 * (double)S - 0.5;
        if (!(fx == fx) || !(fy == fy)) continue;
        if (fx + rp < 0 || fy + rp < 0 || fx - rp > S - 1 || fy - rp > S - 1) continue;
        int x0 = (int)floor(fx - rp), x1 = (int)ceil(fx + rp), y0 = (int)floor(fy - rp), y1 = (int)ceil(fy + rp);
        if (x0 < 0) x0 = 0;
        if (y0 < 0) y0 = 0;
        if (x1 > S - 1) x1 = S - 1;
        if (y1 > S - 1) y1 = S - 1;
        for (int ty = y0 / T; ty <= y1 / T; ++ty)
          for (int tx = x0 / T; tx <= x1 / T; ++tx) {
            const int t = ty * nt + tx;
            if (pass == 1) list[start[t] + count[t]] = p;
            ++count[t];
          }
      }
      if (pass == 0) {
        size_t tot = 0;
        for (int t = 0; t < nt * nt; ++t) { start[t] = (int)tot; tot += (size_t)count[t]; }
        if (tot > list_cap) {
          free(list);
          list = (int*)malloc(sizeof(int) * (tot ? tot : 1));
          list_cap = tot;
          if (!list) { free(count); free(start); free(hits); return PSO_ENOMEM; }
        }
      }
    }
    for (int yi = 0; yi < S; ++yi) {
      const float yf = pix_to_ndc(S - 1 - yi, S);
      for (int xi = 0; xi < S; ++xi) {
        const float xf = pix_to_ndc(S - 1 - xi, S);
        const int t = (yi / T) * nt + (xi / T);
        int n = 0;
        for (int j = 0; j < count[t]; ++j) {
          const int p = list[start[t] + j];
          const float* pt = pb + (size_t)p * 3;
          float d2;
          if (test_point(pt, xf, yf, r2, &d2)) { hits[n].z = pt[2]; hits[n].idx = p; hits[n].d2 = d2; ++n; }
        }
        const size_t o = (((size_t)b * S + yi) * S + xi) * K;
        emit_pixel(hits, n, K, (int32_t)((size_t)b * P), idx + o, zbuf + o, dist2 ? dist2 + o : NULL);
      }
    }
  }
  free(count); free(start); free(list); free(hits);
  return PSO_OK;
}

/*
 * idx,dist2 (B,S,S,K) as produced above; feat (B,C,P) (the layer's `src`, z_buffer_layers.py:55);
 * out (B,C,S,S).  accumulation: 0 alphacomposite, 1 wsum, 2 wsumnorm (z_buffer_layers.py:112-129).
 * alpha = (1 - clamp(dist2 / radius^rad_pow, 1e-3, 1)^0.5)^tau  (z_buffer_layers.py:89-98).
 */
int pso_composite(const int32_t* idx, const float* dist2, const float* feat, int B, int P, int C, int S, int K,
                  double radius, int rad_pow, double tau, int accumulation, float* out) {
  if (!idx || !dist2 || !feat || !out) return PSO_EINVAL;
  const float denom = (float)pow(radius, (double)rad_pow);
  const float tauf = (float)tau;
  for (int b = 0; b < B; ++b)
    for (int yi = 0; yi < S; ++yi)
      for (int xi = 0; xi < S; ++xi) {
        const size_t o = (((size_t)b * S + yi) * S + xi) * K;
        for (int c = 0; c < C; ++c) {
          float acc = 0.0f, cum = 1.0f, wsum = 0.0f;
          for (int k = 0; k < K; ++k) {
            const int32_t n = idx[o + k];
            if (n < 0) continue;
            float d = dist2[o + k] / denom;
            if (d < 1e-3f) d = 1e-3f;
            if (d > 1.0f) d = 1.0f;
            float a = 1.0f - sqrtf(d);
            if (tauf != 1.0f) a = powf(a, tauf);
            const int32_t p = n - (int32_t)((size_t)b * P);
            const float f = feat[((size_t)b * C + c) * P + p];
            if (accumulation == 0) {
              acc = acc + f * cum * a;
              cum = cum * (1.0f - a);
            } else {
              acc = acc + a * f;
              wsum = wsum + a;
            }
          }
          if (accumulation == 2) acc = acc / (wsum > 1e-4f ? wsum : 1e-4f);
          out[(((size_t)b * C + c) * S + yi) * S + xi] = acc;
        }
      }
  return PSO_OK;
}

/* bg = box_dilate_ksize(idx[..., 0] < 0), zero padding ksize/2 (z_buffer_layers.py:100-110). */
int pso_bgmask(const int32_t* idx, int B, int S, int K, int ksize, uint8_t* bg) {
  if (!idx || !bg || ksize < 1) return PSO_EINVAL;
  const int h = ksize / 2;
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < S; ++y)
      for (int x = 0; x < S; ++x) {
        int any = 0;
        /* nn.Conv2d(k, padding=k//2): taps dy in [-h, ksize-1-h] */
        for (int dy = -h; dy <= ksize - 1 - h && !any; ++dy)
          for (int dx = -h; dx <= ksize - 1 - h; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if (yy < 0 || yy >= S || xx < 0 || xx >= S) continue;
            if (idx[(((size_t)b * S + yy) * S + xx) * K] < 0) { any = 1; break; }
          }
        bg[((size_t)b * S + y) * S + x] = (uint8_t)any;
      }
  return PSO_OK;
}
