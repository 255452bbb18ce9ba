// Wavefront sampler for the locally-masked-convolution PixelCNN over VQ-VAE-2 codes, on tcgen05 tensor cores.
//
// Replaces models/lmconv/sample.py:8-73 (sample) driving models/lmconv/model.py:110-155 (OurPixelCNN.forward):
// the reference re-runs the whole 32x32 network for every sampled token (11.43 GFLOP/token).
//
// Two facts make the loop tensor-core shaped:
//   1. The masks make the network causal in generation order, so a cell's 42-layer activation column depends only
//      on cells generated earlier and is computed ONCE (11.16 MFLOP/token).
//   2. A cell reads, at every layer, only its own column and the columns of its masked-in 3x3 (dilation 1 or 2)
//      neighbours.  Cells that are not neighbours are therefore independent regardless of their rank in the
//      order: the dependency DAG is levelled on the host (ps_lmconv_levels_host) and every level -- all the
//      known prefix cells of all images at level 0, then wavefronts of mutually independent sampled cells -- is
//      one launch in which each CTA drives a tile of 128 (image, cell) rows through the whole column.
//
// Per CTA: rows are the UMMA M dimension (one TMEM lane = one cell), output channels the N dimension, and
// (tap, input channel) the K dimension, walked in 64-wide chunks through a 6-stage shared-memory ring:
//   warp 0      streams the pre-swizzled fp16 weight tile of each chunk with cp.async.bulk (static schedule),
//   warps 6-9   gather the neighbours' cached activations of each chunk with zero-filling cp.async (a masked-out
//               tap is a zero row, so the mask costs no bandwidth),
//   warp 1      issues tcgen05.mma (M=128, N=80/160/128, K=16) into two ping-pong TMEM accumulators,
//   warps 2-5   epilogue, one thread per row: bias, PONO (a thread-local reduction over the 80 channels of its own
//               TMEM lane), gate / residual (the residual stream lives in spare TMEM columns), concat_elu, the
//               cache write, and the centre-tap operand chunks of the NEXT layer written straight into the ring
//               -- the only data a layer needs from the previous one -- so the non-centre chunks of layer l+1
//               are multiplied while the epilogue of layer l runs.
// After the last layer nin_out puts the 512 logits of each row in TMEM and the row's thread draws the token
// (softmax / temperature, inverse CDF with the caller's uniform).
//
// Layouts: activation cache fp16 (B, 33 tensors, 1024 cells, 240) = [elu(x) | elu(-x) | x]; weights fp16, one
// 128-byte-swizzled K-major [cout][64] tile per chunk in schedule order (pixelsynth_b200/lmconv.py packs them).
#include <cuda_fp16.h>

#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tc05.cuh"

namespace ps {

constexpr int TC_THREADS = 320;
constexpr int TC_STAGES = PS_LMCONV_STAGES;
constexpr int TC_A_BYTES = 128 * 128;  // 128 rows x 64 fp16
constexpr int TC_W_BYTES = 160 * 128;  // up to 160 output channels x 64 fp16
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_W_BYTES;
constexpr int LMT_F = 80;
constexpr int LMT_CELLS = 1024;
constexpr int LMT_TENSORS = 33;
constexpr int LMT_ACT = 240;     // fp16 per (tensor, cell): elu(x) | elu(-x) | x
constexpr int LMT_CLASSES = 512;
constexpr int COL_OG = 400;      // TMEM columns [400, 480): the row's residual stream (fp32)
constexpr int TC_NOPS = 18;
constexpr int TC_MAX_CHUNKS = 728;

enum { A_GATHER = 0, A_CENTRE = 1, A_EPILOGUE = 2 };

// ps_lmconv_chunk in 8 bytes.  x: w_off16 (24) | w_rows / 8 (5) | a_kind (2);  y: a_tensor (6) | mask (2) | cin == 160 (1)
// | kc (5) | reads the raw third of the cache row (1) | d_col / 16 (5) | flags (6)
struct Chunk {
  uint32_t w_off16;
  int w_rows, a_kind, a_tensor, mask, cin8, kc, ch_off8, d_col, flags;
};
__device__ __forceinline__ uint2 pack_chunk(const ps_lmconv_chunk& c) {
  uint2 r;
  r.x = (c.w_off16 & 0xffffffu) | ((uint32_t)(c.w_rows >> 3) << 24) | ((uint32_t)c.a_kind << 29);
  r.y = (uint32_t)c.a_tensor | ((uint32_t)c.mask << 6) | ((uint32_t)(c.cin8 == 20) << 8) | ((uint32_t)c.kc << 9) |
        ((uint32_t)(c.ch_off8 != 0) << 14) | ((uint32_t)(c.d_col >> 4) << 15) | ((uint32_t)c.flags << 20);
  return r;
}
__device__ __forceinline__ Chunk unpack_chunk(uint2 r) {
  Chunk c;
  c.w_off16 = r.x & 0xffffffu;
  c.w_rows = (int)((r.x >> 24) & 31u) << 3;
  c.a_kind = (int)(r.x >> 29) & 3;
  c.a_tensor = (int)(r.y & 63u);
  c.mask = (int)(r.y >> 6) & 3;
  c.cin8 = ((r.y >> 8) & 1u) ? 20 : 10;
  c.kc = (int)(r.y >> 9) & 31;
  c.ch_off8 = ((r.y >> 14) & 1u) ? 20 : 0;
  c.d_col = (int)((r.y >> 15) & 31u) << 4;
  c.flags = (int)(r.y >> 20) & 63;
  return c;
}
enum { FORM_NONE = 0, FORM_PAIR = 1, FORM_RAW = 2 };
enum { ROW_SAMPLED = 1u << 16, ROW_LOGITS = 1u << 17, ROW_VALID = 1u << 18 };

struct TcParams {
  const unsigned char* wblob;
  const ps_lmconv_chunk* chunks;
  int n_body, n_total;
  int epi_first[PS_LMCONV_MAX_GEMMS];
  const __half* w_uinit;  // [9][513][80]
  const float* bias;
  int b_uinit, b_nin;
  ps_lmconv_op ops[TC_NOPS];
  __half* act;
  const ps_lmconv_row* rows;
  int row_begin, row_end;
  int rows_per_cta;  // 16, 32, 64 or 128 rows of the 128-row UMMA tile carry work (small levels spread over more SMs)
  long long* codes;
  const float* uniforms;
  int ustride;
  float inv_temperature;
  float* logits_out;
  int debug;         // developer aid (PS_TC_DEBUG): bit0 skip the weight copies, bit1 skip the gather copies (timing only)
  long long* trace;  // developer aid: clock64 timestamps of CTA 0 (ps_lmconv_tc_set_trace), or null
};

#define TC_TRACE(slot, idx)                                                   \
  do {                                                                        \
    if (p.trace && blockIdx.x == 0) p.trace[(slot) * 1024 + (idx)] = clock64(); \
  } while (0)

struct TcSmem {
  uint64_t full[TC_STAGES], empty[TC_STAGES], acc_full[3], ctr[2];
  uint32_t tmem_slot, pad_;
  ps_lmconv_row rows[128];
  uint4 rowtab[128];  // per tile row, for the gather warps: cache base address (x, y), packed mask words (z)
  uint2 sched[TC_MAX_CHUNKS];  // the chunk schedule, packed (a dependent global load per chunk would pace every role)
};

// Operands are fp16, not bf16: every activation that reaches a multiply is O(1) (PONO outputs, their ELUs, the
// residual stream) and the weights are O(0.1), far inside fp16's range, and the 11-bit significand keeps the
// 33-layer column within 0.2% of the fp32 oracle where bf16 operands drift to 1%.  Same tensor-core rate.
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// concat_elu of one value: (elu(x), elu(-x)) with a single exponential
__device__ __forceinline__ void celu(float x, float& p, float& n) {
  const float ax = fabsf(x);
  const float e = ax < 0.03125f ? -ax * (1.0f - 0.5f * ax * (1.0f - 0.33333333f * ax)) : __expf(-ax) - 1.0f;  // expm1(-|x|)
  p = x > 0.0f ? x : e;
  n = x > 0.0f ? e : -x;
}

// v[0..16) += b[0..16) with four 16-byte loads (every bias block starts on a 16-byte boundary)
__device__ __forceinline__ void add_bias16(float* v, const float* b) {
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(b + i));
    v[i] += f.x;
    v[i + 1] += f.y;
    v[i + 2] += f.z;
    v[i + 3] += f.w;
  }
}
__device__ __forceinline__ void add_bias80(float* v, const float* b) {
#pragma unroll
  for (int j = 0; j < 5; ++j) add_bias16(v + 16 * j, b + 16 * j);
}

// positional normalisation over the 80 channels held by this thread (layers.py:224-236, unbiased variance)
__device__ __forceinline__ void pono80(float* v) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LMT_F; ++i) s += v[i];
  const float mean = s * (1.0f / LMT_F);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LMT_F; ++i) {
    const float d = v[i] - mean;
    q = fmaf(d, d, q);
  }
  const float inv = 1.0f / sqrtf(q * (1.0f / (LMT_F - 1)) + 1e-5f);
#pragma unroll
  for (int i = 0; i < LMT_F; ++i) v[i] = (v[i] - mean) * inv;
}

struct Epi {
  unsigned char* tiles;
  TcSmem* sm;
  int r;            // tile row of this thread
  uint32_t tlane;   // TMEM address of this thread's lane, column 0
  bool valid;
  __half* actrow;  // act + (b * 33 * 1024 + cell) * 240; tensor t adds t * 1024 * 240

  __device__ __forceinline__ uint32_t a_addr(int chunk, int kg) const {  // 16-byte group kg (0..7) of this row
    const int s = chunk % TC_STAGES;
    return smem_u32(tiles + (size_t)s * TC_STAGE_BYTES) + r * 128 + ((kg ^ (r & 7)) << 4);
  }
  // The stages of GEMM g's centre chunks are free once the chunk TC_STAGES before the last of them has been
  // multiplied; the MMA warp signals exactly that on ctr[g & 1] (chunk flag bit 4).  The ring's own `empty`
  // barriers cannot be used here: this role touches a stage only every few phases, and an mbarrier wait can tell
  // the current phase from the previous one only.
  __device__ __forceinline__ void acquire(int g) const { mbar_wait(&sm->ctr[g & 1], (uint32_t)(g >> 1) & 1u); }
  __device__ __forceinline__ void publish(int first, int count) const {
    fence_proxy_async();
    __syncwarp();
    if ((threadIdx.x & 31) < 8)
      for (int c = first; c < first + count; ++c) mbar_arrive(&sm->full[c % TC_STAGES]);
  }
  __device__ __forceinline__ void sts16(int first, int kgg, uint4 v) const {  // kgg = group index over the GEMM's centre K
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_addr(first + (kgg >> 3), kgg & 7)), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
  }
  // One 16-channel piece j (channels 16j..16j+15) of a finished tensor: cache write + centre operand of the next GEMM.
  __device__ __forceinline__ void emit16(int form, int first, int j, const float* x, int tensor, bool raw) const {
    uint4 p0, p1, n0, n1, r0, r1;
    {
      float p[16], n[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) celu(x[i], p[i], n[i]);
      p0 = make_uint4(pack_h2(p[0], p[1]), pack_h2(p[2], p[3]), pack_h2(p[4], p[5]), pack_h2(p[6], p[7]));
      p1 = make_uint4(pack_h2(p[8], p[9]), pack_h2(p[10], p[11]), pack_h2(p[12], p[13]), pack_h2(p[14], p[15]));
      n0 = make_uint4(pack_h2(n[0], n[1]), pack_h2(n[2], n[3]), pack_h2(n[4], n[5]), pack_h2(n[6], n[7]));
      n1 = make_uint4(pack_h2(n[8], n[9]), pack_h2(n[10], n[11]), pack_h2(n[12], n[13]), pack_h2(n[14], n[15]));
      r0 = make_uint4(pack_h2(x[0], x[1]), pack_h2(x[2], x[3]), pack_h2(x[4], x[5]), pack_h2(x[6], x[7]));
      r1 = make_uint4(pack_h2(x[8], x[9]), pack_h2(x[10], x[11]), pack_h2(x[12], x[13]), pack_h2(x[14], x[15]));
    }
    if (valid) {
      uint4* g = reinterpret_cast<uint4*>(actrow + (size_t)tensor * LMT_CELLS * LMT_ACT);
      g[2 * j] = p0;
      g[2 * j + 1] = p1;
      g[10 + 2 * j] = n0;
      g[10 + 2 * j + 1] = n1;
      if (raw) {
        g[20 + 2 * j] = r0;
        g[20 + 2 * j + 1] = r1;
      }
    }
    const uint4 z = make_uint4(0, 0, 0, 0);
    if (form == FORM_PAIR) {  // K = [elu(x) 0..79 | elu(-x) 80..159 | 0 .. 191]
      sts16(first, 2 * j, p0);
      sts16(first, 2 * j + 1, p1);
      sts16(first, 10 + 2 * j, n0);
      sts16(first, 10 + 2 * j + 1, n1);
      if (j < 4) sts16(first, 20 + j, z);
    } else if (form == FORM_RAW) {  // K = [x 0..79 | 0 .. 127]
      sts16(first, 2 * j, r0);
      sts16(first, 2 * j + 1, r1);
      sts16(first, 10 + j, z);
      if (j == 0) sts16(first, 15, z);
    }
  }
};

__global__ void __launch_bounds__(TC_THREADS, 1) lmconv_tc_kernel(const TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  TcSmem& sm = *reinterpret_cast<TcSmem*>(tiles + (size_t)TC_STAGES * TC_STAGE_BYTES);
  const int tid = threadIdx.x, lane = tid & 31;
  // the shuffle tells the compiler the warp index is warp-uniform: role branches become uniform branches and the
  // issuing roles' descriptor arithmetic can stay in uniform registers
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int row0 = p.row_begin + blockIdx.x * p.rows_per_cta;

  if (tid < 128) {
    ps_lmconv_row ri;
    ri.bc = 0;
    ri.w01 = 0;
    ri.w2_flags = 0;
    ri.uidx = 0;
    if (tid < p.rows_per_cta && row0 + tid < p.row_end) ri = p.rows[row0 + tid];
    sm.rows[tid] = ri;
  }
  for (int i = tid; i < p.n_total; i += TC_THREADS) sm.sched[i] = pack_chunk(p.chunks[i]);
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < TC_STAGES; ++s) {
        mbar_init(&sm.full[s], 33);  // 32 arrivals of the row writers (one gather warp, or 8 lanes of each epilogue warp) + the weight producer
        mbar_init(&sm.empty[s], 1);
      }
      for (int i = 0; i < 3; ++i) mbar_init(&sm.acc_full[i], 1);
      for (int i = 0; i < 2; ++i) mbar_init(&sm.ctr[i], 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(&sm.tmem_slot, 512);
  }
  tc_fence_before();
  const int need_logits =
      __syncthreads_or(tid < p.rows_per_cta && row0 + tid < p.row_end &&
                       (p.rows[row0 + tid].w2_flags & (ROW_SAMPLED | ROW_LOGITS)));
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, sm.tmem_slot, 0);
  if (tid == 0) TC_TRACE(7, 0);
#ifdef PS_TC_DEBUG
  if (tid == 0 && blockIdx.x == 0) printf("lmconv_tc: full[0] at smem 0x%x, rows %d..%d need_logits %d\n", smem_u32(&sm.full[0]), p.row_begin, p.row_end, need_logits);
#endif
  const int nchunks = need_logits ? p.n_total : p.n_body;

  if (warp == 0) {
    // ===== weight producer: whole warp, warp-uniform values, one elected lane issues (see umma_f16_kblock) =====
    {
      int s = 0;
      uint32_t ph = 1;
      for (int i = 0; i < nchunks; ++i) {
        uint2 raw = sm.sched[i];
        raw.x = __shfl_sync(0xffffffffu, raw.x, 0);  // warp-uniform by construction; now also to the compiler
        mbar_wait(&sm.empty[s], ph);
        if (lane == 0) TC_TRACE(3, i);
        const uint32_t bytes = ((raw.x >> 24) & 31u) << 10;  // w_rows * 128
        if (p.debug & 1) {
          if (lane == 0) mbar_arrive(&sm.full[s]);
        } else {
          bulk_load_elect(tiles + (size_t)s * TC_STAGE_BYTES + TC_A_BYTES, p.wblob + (size_t)(raw.x & 0xffffffu) * 16, bytes,
                          &sm.full[s]);
        }
        if (++s == TC_STAGES) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop with warp-uniform values; one elected lane issues =====
    {
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(tiles));
      const uint32_t idesc0 = umma_idesc_f16(0);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nchunks; ++i) {
        uint2 raw = sm.sched[i];
        raw.x = __shfl_sync(0xffffffffu, raw.x, 0);  // warp-uniform by construction; now also to the compiler
        raw.y = __shfl_sync(0xffffffffu, raw.y, 0);
        mbar_wait(&sm.full[s], ph);
        if (!(p.debug & 8)) fence_proxy_async();
        tc_fence_after();
        if (lane == 0) TC_TRACE(0, i);
        const uint32_t flags = raw.y >> 20;
        const uint32_t idesc = idesc0 | (((raw.x >> 24) & 31u) << 17);       // N >> 3 = w_rows / 8
        const uint32_t d = tmem_base + (((raw.y >> 15) & 31u) << 4);          // d_col
        const uint32_t a_lo = a_lo0 + (uint32_t)s * (TC_STAGE_BYTES >> 4);
        umma_f16_kblock(d, a_lo, a_lo + (TC_A_BYTES >> 4), idesc, flags & 1u, &sm.empty[s]);
        if (flags & 2u) umma_commit_elect(&sm.acc_full[(flags >> 2) & 3u]);
        if (flags & 16u) umma_commit_elect(&sm.ctr[(flags >> 5) & 1u]);
        if (++s == TC_STAGES) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp >= 6) {
    // ===== gather producers: warp w fills every 4th gathered chunk on its own =====
    // A chunk is 128 rows x 8 groups of 16 bytes.  Lane -> group g = lane & 7 of rows rs, rs + 4, ... (rs = lane >> 3),
    // so eight lanes read one contiguous 128-byte segment.  Whatever depends on the row only (cache base address,
    // the three mask words packed into one register: bit m*9+tap, bit 27 = row valid) sits in a shared-memory
    // table; per chunk a lane derives ONE (tap, channel group) from the descriptor, per row it tests one bit, adds
    // one offset and issues one zero-filling cp.async.  One warp per chunk keeps the per-chunk fixed cost (descriptor,
    // barrier wait, arrival) off the other three warps, which are busy with the next chunks.
    const int t = tid - 192;
    {
      const ps_lmconv_row ri = sm.rows[t];
      const bool v = (ri.w2_flags & ROW_VALID) != 0;
      const unsigned long long base =
          (unsigned long long)p.act +
          (v ? ((size_t)(ri.bc >> 10) * LMT_TENSORS * LMT_CELLS + (ri.bc & 1023)) * (LMT_ACT * 2) : 0);
      const uint32_t w27 =
          v ? ((ri.w01 & 0x1ffu) | (((ri.w01 >> 16) & 0x1ffu) << 9) | ((ri.w2_flags & 0x1ffu) << 18) | (1u << 27)) : 0u;
      sm.rowtab[t] = make_uint4((uint32_t)base, (uint32_t)(base >> 32), w27, 0u);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");  // the four gather warps
    const int gw = warp - 6, g = lane & 7, rs = lane >> 3;
    const uint32_t dst_even = smem_u32(tiles) + rs * 128 + ((g ^ rs) << 4);        // rows rs + 8m
    const uint32_t dst_odd = smem_u32(tiles) + (rs + 4) * 128 + ((g ^ (rs + 4)) << 4);  // rows rs + 4 + 8m
    const int npair = p.rows_per_cta >> 3;  // row pairs (rs + 8m, rs + 4 + 8m) of this lane
    int seen = 0;                            // gathered chunks so far: this warp takes those with seen % 4 == gw
    for (int i = 0; i < nchunks; ++i) {
      const Chunk ch = unpack_chunk(sm.sched[i]);
      if (ch.a_kind == A_EPILOGUE) continue;
      if ((seen++ & 3) != gw) continue;
      const int s = i % TC_STAGES;
      const int kg = ch.kc * 8 + g;
      int bitpos, off;  // mask bit to test, byte offset from the row's base
      if (ch.a_kind == A_GATHER) {
        const int slot = ch.cin8 == 20 ? kg / 20 : kg / 10;
        const int c8 = kg - slot * ch.cin8;
        const int tap = slot + (slot >= 4 ? 1 : 0);
        const int tr = (tap * 11) >> 5;  // tap / 3 for tap < 9
        const int dil = ch.mask == 2 ? 2 : 1;
        bitpos = ch.mask * 9 + tap;
        off = (((tr - 1) * 32 + (tap - 3 * tr - 1)) * dil + ch.a_tensor * LMT_CELLS) * (LMT_ACT * 2) + (ch.ch_off8 + c8) * 16;
      } else {
        bitpos = kg < ch.cin8 ? 27 : 31;
        off = ch.a_tensor * LMT_CELLS * (LMT_ACT * 2) + (ch.ch_off8 + kg) * 16;
      }
      mbar_wait(&sm.empty[s], ((uint32_t)(i / TC_STAGES) & 1u) ^ 1u);
      if (lane == 0) TC_TRACE(1, i);
      const uint32_t soff = s * TC_STAGE_BYTES;
      if (!(p.debug & 2)) {
#pragma unroll 4
        for (int m = 0; m < npair; ++m) {
          const uint4 r0 = sm.rowtab[rs + 8 * m], r1 = sm.rowtab[rs + 4 + 8 * m];
          const uint32_t ok0 = (r0.z >> bitpos) & 1u, ok1 = (r1.z >> bitpos) & 1u;
          const unsigned long long b0 = ((unsigned long long)r0.y << 32 | r0.x) + (ok0 ? off : 0);
          const unsigned long long b1 = ((unsigned long long)r1.y << 32 | r1.x) + (ok1 ? off : 0);
          cp_async16_zfill(dst_even + soff + m * 1024, (const void*)b0, ok0 << 4);
          cp_async16_zfill(dst_odd + soff + m * 1024, (const void*)b1, ok1 << 4);
        }
      }
      // asynchronous completion: the stage's full barrier gets this lane's arrival when its copies have landed, so
      // the warp never blocks on data and every free stage of the ring is in flight (the MMA warp orders the
      // landed generic-proxy writes before its async-proxy reads with fence.proxy.async)
      if (p.debug & 32)
        mbar_arrive(&sm.full[s]);  // timing experiment (with bit 1): how long does the asynchronous arrival itself take?
      else
        cp_async_arrive_noinc(&sm.full[s]);
      if (lane == 0) TC_TRACE(2, i);
    }
  } else {
    // ===== epilogue: one thread per row =====
    Epi e;
    e.tiles = tiles;
    e.sm = &sm;
    const int quad = warp & 3;
    e.r = quad * 32 + lane;
    e.tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const ps_lmconv_row ri = sm.rows[e.r];
    e.valid = (ri.w2_flags & ROW_VALID) != 0;
    const int b = ri.bc >> 10, cell = ri.bc & 1023;
    e.actrow = p.act + ((size_t)b * LMT_TENSORS * LMT_CELLS + cell) * LMT_ACT;
    const float* bias = p.bias;
    int g = 0;  // GEMM counter

    // ---- u_init over [one-hot(code) | ones]: a gather of weight rows (mask A), then PONO ----
    {
      float v[LMT_F];
#pragma unroll
      for (int i = 0; i < LMT_F; ++i) v[i] = 0.f;
      add_bias80(v, bias + p.b_uinit);
      const uint32_t w0 = ri.w01 & 0x1ffu;
      for (int tap = 0; tap < 9; ++tap) {
        if (!((w0 >> tap) & 1u)) continue;
        const int nbr = cell + (tap / 3 - 1) * 32 + (tap % 3 - 1);
        const int code = (int)p.codes[(size_t)b * LMT_CELLS + nbr];
        const uint4* wc = reinterpret_cast<const uint4*>(p.w_uinit + ((size_t)tap * (LMT_CLASSES + 1) + code) * LMT_F);
        const uint4* w1 = reinterpret_cast<const uint4*>(p.w_uinit + ((size_t)tap * (LMT_CLASSES + 1) + LMT_CLASSES) * LMT_F);
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          const uint4 a = __ldg(wc + j), c = __ldg(w1 + j);
          const __half2* ha = reinterpret_cast<const __half2*>(&a);
          const __half2* hc = reinterpret_cast<const __half2*>(&c);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 fa = __half22float2(ha[q]), fc = __half22float2(hc[q]);
            v[8 * j + 2 * q] += fa.x + fc.x;
            v[8 * j + 2 * q + 1] += fa.y + fc.y;
          }
        }
      }
      pono80(v);
      const int first = p.epi_first[0];
      e.acquire(0);
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        tmem_st16_nowait(e.tlane + COL_OG + 16 * j, v + 16 * j);
        e.emit16(FORM_PAIR, first, j, v + 16 * j, 0, true);
      }
      tmem_st_wait();
      e.publish(first, 3);
    }

    for (int oi = 0; oi < TC_NOPS; ++oi) {
      const ps_lmconv_op op = p.ops[oi];
      const int next_form = oi + 1 < TC_NOPS ? (p.ops[oi + 1].kind == 0 ? FORM_PAIR : FORM_RAW) : FORM_NONE;
      const int next_count = next_form == FORM_PAIR ? 3 : (next_form == FORM_RAW ? 2 : 0);
      if (op.kind == 0) {
        {  // x = PONO(conv_input(concat_elu(og))) [+ nin_skip(concat_elu(a))]
          const uint32_t col0 = (uint32_t)(g & 1) * 160u;
          mbar_wait(&sm.acc_full[g & 1], (uint32_t)(g >> 1) & 1u);
          if (e.r == 0) TC_TRACE(4, g);
          tc_fence_after();
          float x[LMT_F];
#pragma unroll
          for (int j = 0; j < 5; ++j) tmem_ld16_nowait(e.tlane + col0 + 16 * j, x + 16 * j);
          tmem_ld_wait();
          add_bias80(x, bias + op.b_in);
          pono80(x);
          if (op.a >= 0) {
#pragma unroll
            for (int j = 0; j < 5; ++j) {
              float sk[16];
              tmem_ld16_nowait(e.tlane + col0 + 80 + 16 * j, sk);
              add_bias16(x + 16 * j, bias + op.b_skip + 16 * j);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) x[16 * j + i] += sk[i];
            }
          }
          tc_fence_before();
          const int first = p.epi_first[g + 1];
          e.acquire(g + 1);
          if (e.r == 0) TC_TRACE(6, g);
#pragma unroll
          for (int j = 0; j < 5; ++j) e.emit16(FORM_PAIR, first, j, x + 16 * j, op.mid, false);
          e.publish(first, 3);
          if (e.r == 0) TC_TRACE(5, g);
          ++g;
        }
        {  // y = conv_out(concat_elu(x)); og += PONO(y[:80]) * sigmoid(y[80:])
          const uint32_t col0 = (uint32_t)(g & 1) * 160u;
          mbar_wait(&sm.acc_full[g & 1], (uint32_t)(g >> 1) & 1u);
          if (e.r == 0) TC_TRACE(4, g);
          tc_fence_after();
          float a[LMT_F];
#pragma unroll
          for (int j = 0; j < 5; ++j) tmem_ld16_nowait(e.tlane + col0 + 16 * j, a + 16 * j);
          tmem_ld_wait();
          add_bias80(a, bias + op.b_out);
          pono80(a);
          const int first = next_count ? p.epi_first[g + 1] : 0;
          if (next_count) e.acquire(g + 1);
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            float gt[16], o[16];
            tmem_ld16_nowait(e.tlane + col0 + 80 + 16 * j, gt);
            tmem_ld16_nowait(e.tlane + COL_OG + 16 * j, o);
            tmem_ld_wait();
            add_bias16(gt, bias + op.b_out + 80 + 16 * j);
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = fmaf(a[16 * j + i], 1.0f / (1.0f + __expf(-gt[i])), o[i]);
            tmem_st16_nowait(e.tlane + COL_OG + 16 * j, o);
            e.emit16(next_form, first, j, o, op.out, true);
          }
          tmem_st_wait();
          tc_fence_before();
          e.publish(first, next_count);
          if (e.r == 0) TC_TRACE(5, g);
          ++g;
        }
      } else {  // dilated masked conv on the raw stream + PONO becomes the new stream
        const uint32_t col0 = (uint32_t)(g & 1) * 160u;
        mbar_wait(&sm.acc_full[g & 1], (uint32_t)(g >> 1) & 1u);
        if (e.r == 0) TC_TRACE(4, g);
        tc_fence_after();
        float x[LMT_F];
#pragma unroll
        for (int j = 0; j < 5; ++j) tmem_ld16_nowait(e.tlane + col0 + 16 * j, x + 16 * j);
        tmem_ld_wait();
        add_bias80(x, bias + op.b_in);
        pono80(x);
        const int first = next_count ? p.epi_first[g + 1] : 0;
        if (next_count) e.acquire(g + 1);
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          tmem_st16_nowait(e.tlane + COL_OG + 16 * j, x + 16 * j);
          e.emit16(next_form, first, j, x + 16 * j, op.out, true);
        }
        tmem_st_wait();
        tc_fence_before();
        e.publish(first, next_count);
        if (e.r == 0) TC_TRACE(5, g);
        ++g;
      }
    }

    if (need_logits) {
      // ---- logits = nin_out(elu(u)): four 128-class quarters, A = elu(u) rewritten per quarter ----
      for (int q = 0; q < 4; ++q) {
        const int first = p.epi_first[g + q];
        // quarters 0-2 reuse stages whose last users (body chunks) completed before acc_full of the last GEMM fired;
        // quarter 3 reuses quarter 0's stages: the schedule signals that on ctr as if it were GEMM g + 1
        if (q == 3) e.acquire(g + 1);
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          float o[16], pp[16], nn;
          tmem_ld16_nowait(e.tlane + COL_OG + 16 * j, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) celu(o[i], pp[i], nn);
          e.sts16(first, 2 * j,
                  make_uint4(pack_h2(pp[0], pp[1]), pack_h2(pp[2], pp[3]), pack_h2(pp[4], pp[5]), pack_h2(pp[6], pp[7])));
          e.sts16(first, 2 * j + 1,
                  make_uint4(pack_h2(pp[8], pp[9]), pack_h2(pp[10], pp[11]), pack_h2(pp[12], pp[13]),
                             pack_h2(pp[14], pp[15])));
          e.sts16(first, 10 + j, make_uint4(0, 0, 0, 0));
          if (j == 0) e.sts16(first, 15, make_uint4(0, 0, 0, 0));
        }
        tc_fence_before();
        e.publish(first, 2);
      }
      mbar_wait(&sm.acc_full[2], 0);
      tc_fence_after();
      const bool sampled = e.valid && (ri.w2_flags & ROW_SAMPLED) && p.uniforms;
      const bool want = e.valid && (ri.w2_flags & ROW_LOGITS) && p.logits_out;
      float* lo = want ? p.logits_out + ((size_t)b * LMT_CELLS + cell) * LMT_CLASSES : nullptr;
      // pass 1: (optional) logits out, running max and sum of exp((l - max) / T)
      float mx = -INFINITY, sum = 0.f;
      for (int c0 = 0; c0 < LMT_CLASSES; c0 += 16) {
        float l[16];
        tmem_ld16_nowait(e.tlane + c0, l);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) l[i] += __ldg(bias + p.b_nin + c0 + i);
        if (lo) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(lo + c0 + i) = make_float4(l[i], l[i + 1], l[i + 2], l[i + 3]);
        }
        float m2 = mx;
#pragma unroll
        for (int i = 0; i < 16; ++i) m2 = fmaxf(m2, l[i] * p.inv_temperature);
        sum *= __expf(mx - m2);
#pragma unroll
        for (int i = 0; i < 16; ++i) sum += __expf(l[i] * p.inv_temperature - m2);
        mx = m2;
      }
      if (__any_sync(0xffffffffu, sampled)) {
        // pass 2: token = first j with cumsum(softmax(l / T))_j > u.  tcgen05.ld is warp-collective, so the loop
        // is kept warp-uniform: it runs until every sampled row of the warp has found its token.
        const float thr = sampled ? p.uniforms[(size_t)b * p.ustride + ri.uidx] * sum : 0.f;
        float cum = 0.f;
        int token = LMT_CLASSES - 1;
        bool found = !sampled;
        for (int c0 = 0; c0 < LMT_CLASSES; c0 += 16) {
          if (__all_sync(0xffffffffu, found)) break;
          float l[16];
          tmem_ld16_nowait(e.tlane + c0, l);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            cum += __expf((l[i] + __ldg(bias + p.b_nin + c0 + i)) * p.inv_temperature - mx);
            if (!found && cum > thr) {
              token = c0 + i;
              found = true;
            }
          }
        }
        if (sampled) p.codes[(size_t)b * LMT_CELLS + cell] = token;
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (tid == 0) TC_TRACE(7, 1);
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace ps

using namespace ps;

static thread_local long long* g_tc_trace = nullptr;

extern "C" {

// developer aid (tools/trace_lmconv.py): device buffer of 8 x 1024 int64 receiving CTA 0's timestamps, or NULL
void ps_lmconv_tc_set_trace(void* dev_buffer) { g_tc_trace = (long long*)dev_buffer; }

size_t ps_lmconv_tc_cache_bytes(int B) {
  return (size_t)(B > 0 ? B : 0) * LMT_TENSORS * LMT_CELLS * LMT_ACT * sizeof(__half);
}

// Levels of the dependency DAG (host): a cell sits one level above the highest of its masked-in neighbours (the first
// cell of the order, which reads nothing, is level 0).  mode 0: sampling -- cells ranked after the image's last
// sampled cell are dropped, images with nothing to sample produce no rows.  mode 1: teacher-forced logits of every
// cell.
int ps_lmconv_levels_host(const int* order, const uint16_t* words, const uint8_t* sample_mask, int B, int mode,
                          ps_lmconv_row* rows_out, int* level_offsets, int max_levels, int* n_levels) {
  PS_CHECK_ARG(order && words && rows_out && level_offsets && n_levels && B >= 0 && max_levels >= 2);
  PS_CHECK_ARG(mode == 1 || sample_mask);
  PS_CHECK_ARG(B < (1 << 20));
  std::vector<int> level((size_t)B * LMT_CELLS, -1), uidx((size_t)B * LMT_CELLS, 0), rank(LMT_CELLS);
  int top = 0;
  for (int b = 0; b < B; ++b) {
    const int* ord = order + (size_t)b * LMT_CELLS;
    const uint16_t* w = words + (size_t)b * 3 * LMT_CELLS;
    int* lv = level.data() + (size_t)b * LMT_CELLS;
    for (int i = 0; i < LMT_CELLS; ++i) {
      PS_CHECK_ARG(ord[i] >= 0 && ord[i] < LMT_CELLS);
      rank[ord[i]] = i;
    }
    int last = LMT_CELLS - 1;
    if (mode == 0) {
      const uint8_t* smk = sample_mask + (size_t)b * LMT_CELLS;
      int drawn = 0;
      last = -1;
      for (int i = 0; i < LMT_CELLS; ++i)
        if (smk[ord[i]]) {
          last = i;
          uidx[(size_t)b * LMT_CELLS + ord[i]] = drawn++;
        }
      if (last < 0) continue;  // nothing to sample: the image needs no work
    }
    for (int i = 0; i <= last; ++i) {
      const int cell = ord[i];
      const int r = cell / 32, c = cell % 32;
      int l = -1;
      for (int m = 0; m < 3; ++m) {
        const int dil = m == 2 ? 2 : 1;
        for (int t = 0; t < 9; ++t) {
          if (t == 4 || !((w[m * LMT_CELLS + cell] >> t) & 1)) continue;
          const int rr = r + (t / 3 - 1) * dil, cc = c + (t % 3 - 1) * dil;
          PS_CHECK_ARG(rr >= 0 && rr < 32 && cc >= 0 && cc < 32);       // masks never reach outside the grid
          PS_CHECK_ARG(rank[rr * 32 + cc] < i);                          // ... nor forward in the order
          l = std::max(l, lv[rr * 32 + cc]);
        }
      }
      lv[cell] = l + 1;
      top = std::max(top, l + 1);
    }
  }
  if (top + 1 > max_levels) return fail(PS_EWORKSPACE, "%s: more dependency levels than level_offsets holds%s", __func__);
  std::vector<int> count(top + 2, 0);
  for (size_t i = 0; i < level.size(); ++i)
    if (level[i] >= 0) ++count[level[i] + 1];
  for (int l = 0; l <= top; ++l) count[l + 1] += count[l];
  for (int l = 0; l <= top + 1; ++l) level_offsets[l] = count[l];
  std::vector<int> cursor(count.begin(), count.end() - 1);
  for (int b = 0; b < B; ++b)
    for (int cell = 0; cell < LMT_CELLS; ++cell) {
      const int l = level[(size_t)b * LMT_CELLS + cell];
      if (l < 0) continue;
      const uint16_t* w = words + (size_t)b * 3 * LMT_CELLS;
      ps_lmconv_row ri;
      ri.bc = (b << 10) | cell;
      ri.w01 = (uint32_t)w[cell] | ((uint32_t)w[LMT_CELLS + cell] << 16);
      ri.w2_flags = (uint32_t)w[2 * LMT_CELLS + cell] | ROW_VALID;
      if (mode == 1)
        ri.w2_flags |= ROW_LOGITS;
      else if (sample_mask[(size_t)b * LMT_CELLS + cell])
        ri.w2_flags |= ROW_SAMPLED;
      ri.uidx = uidx[(size_t)b * LMT_CELLS + cell];
      rows_out[cursor[l]++] = ri;
    }
  *n_levels = count[top + 1] == 0 ? 0 : top + 1;
  return PS_OK;
}

int ps_lmconv_tc_run(const ps_lmconv_plan* plan, int B, const ps_lmconv_row* rows_dev, const int* level_offsets_host,
                     int n_levels, long long* codes, const float* uniforms, int uniforms_stride, float temperature,
                     float* logits_out, void* cache, size_t cache_bytes, void* stream) {
  PS_CHECK_ARG(plan && plan->wblob && plan->chunks && plan->w_uinit && plan->bias && codes && cache);
  PS_CHECK_ARG(B >= 0 && n_levels >= 0 && temperature > 0.0f);
  PS_CHECK_ARG(plan->n_chunks_body > 0 && plan->n_chunks_total >= plan->n_chunks_body);
  PS_CHECK_ARG(plan->n_chunks_total <= TC_MAX_CHUNKS);
  if (B == 0 || n_levels == 0) return PS_OK;
  PS_CHECK_ARG(rows_dev && level_offsets_host);
  PS_CHECK_ARG(uniforms || logits_out);  // sampling needs the uniform numbers
  if (cache_bytes < ps_lmconv_tc_cache_bytes(B)) return fail(PS_EWORKSPACE, "%s: activation cache too small%s", __func__);
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.wblob = (const unsigned char*)plan->wblob;
  p.chunks = plan->chunks;
  p.n_body = plan->n_chunks_body;
  p.n_total = plan->n_chunks_total;
  memcpy(p.epi_first, plan->epi_first, sizeof(p.epi_first));
  p.w_uinit = (const __half*)plan->w_uinit;
  p.bias = plan->bias;
  p.b_uinit = plan->b_uinit;
  p.b_nin = plan->b_nin;
  memcpy(p.ops, plan->ops, sizeof(p.ops));
  for (int i = 0; i < TC_NOPS; ++i)
    PS_CHECK_ARG(p.ops[i].out >= 0 && p.ops[i].out < LMT_TENSORS && p.ops[i].mid < LMT_TENSORS);
  p.act = (__half*)cache;
  p.rows = rows_dev;
  p.codes = codes;
  p.uniforms = uniforms;
  p.ustride = uniforms_stride;
  p.inv_temperature = 1.0f / temperature;
  p.logits_out = logits_out;
  p.trace = g_tc_trace;
  p.debug = getenv("PS_TC_DEBUG") ? atoi(getenv("PS_TC_DEBUG")) : 0;
  const size_t smem_bytes = 1024 + (size_t)TC_STAGES * TC_STAGE_BYTES + sizeof(TcSmem);
  static thread_local int attr_dev = -1;
  int dev = 0;
  PS_CUDA(cudaGetDevice(&dev));
  if (attr_dev != dev) {
    PS_CUDA(cudaFuncSetAttribute(lmconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    attr_dev = dev;
  }
  int sms = 148;
  PS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PS_TIME_BEGIN("lmconv_tc_kernel", (cudaStream_t)stream);
  for (int l = 0; l < n_levels; ++l) {
    const int r0 = level_offsets_host[l], r1 = level_offsets_host[l + 1];
    if (r1 <= r0) continue;
    // a level is latency bound per CTA (a 36-GEMM dependent chain), so small levels use fewer rows of each
    // 128-row tile and more SMs; full tiles once the level fills the GPU
    int rpc = 32;
    while (rpc < 128 && (r1 - r0 + rpc - 1) / rpc > sms) rpc *= 2;
    p.row_begin = r0;
    p.row_end = r1;
    p.rows_per_cta = rpc;
    lmconv_tc_kernel<<<(r1 - r0 + rpc - 1) / rpc, TC_THREADS, smem_bytes, (cudaStream_t)stream>>>(p);
    PS_LAUNCHED();
  }
  PS_TIME_END((cudaStream_t)stream);
  if (getenv("PS_CHECK_WEDGE")) {  // developer aid: synchronise and report a wedged barrier protocol
    unsigned int w[8] = {0};
    PS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    PS_CUDA(cudaMemcpyFromSymbol(w, g_wedge, sizeof(w)));
    if (w[0]) {
      char buf[160];
      snprintf(buf, sizeof(buf), "block %u thread %u barrier smem 0x%x parity %u", w[1], w[2], w[3], w[4]);
      return fail(PS_ECUDA, "%s: barrier protocol wedged: %s", __func__, buf);
    }
  }
  return PS_OK;
}

}  // extern "C"
