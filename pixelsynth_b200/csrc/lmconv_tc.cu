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
// Per CTA (one tile of up to 128 rows): rows are the UMMA M dimension (one TMEM lane = one cell), output channels
// the N dimension, and (tap, input channel) the K dimension, walked in 64-wide chunks through a 6-stage
// shared-memory ring:
//   warp 0      streams the pre-swizzled fp16 weight tile of each chunk with cp.async.bulk (static schedule),
//   warps 6-9   gather the neighbours' cached activations of each chunk with zero-filling cp.async (a masked-out
//               tap is a zero row, so the mask costs no bandwidth),
//   warp 1      issues tcgen05.mma (M=128, N=80/160/128, K=16) into two ping-pong TMEM accumulators,
//   warps 2-5   epilogue, one thread per row: bias, PONO (a thread-local reduction over the 80 channels of its own
//               TMEM lane), gate / residual (the residual stream lives in spare TMEM columns), concat_elu, the
//               cache write, and the centre-tap operand of the NEXT layer -- the only data a layer needs from the
//               previous one -- written with tcgen05.st into TMEM columns the next layer's centre MMAs read as
//               their A operand (A-from-TMEM form), so the epilogue never waits for the ring and the non-centre
//               chunks of layer l+1 are multiplied while the epilogue of layer l runs,
//   warp 10     publishes the tile's progress (cache tensors complete) to global memory.
// After the last layer nin_out puts the 512 logits of each row in TMEM and the row's thread draws the token
// (softmax / temperature, inverse CDF with the caller's uniform).
//
// ONE launch runs every level: CTAs take tiles (sorted by level) from an atomic ticket, so a tile only ever waits
// for tiles that are already running or finished.  Layer j of a level needs the neighbours' layer j-1 columns, so
// the levels of the known prefix run as a software pipeline two layers apart (a gather warp polls the previous
// level's progress words before it reads the cache); a level that holds sampled cells additionally waits for the
// previous level's tokens, because its first layer reads them.
//
// Layouts: activation cache fp16 (B, 33 tensors, 1024 cells, 240) = [elu(x) | elu(-x) | x]; weights fp16, one
// 128-byte-swizzled K-major [cout][64] tile per chunk in schedule order (pixelsynth_b200/lmconv.py packs them).
#include <cuda_fp16.h>

#include <stdlib.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "common.cuh"
#include "hostpool.cuh"
#include "tc05.cuh"

namespace ps {

constexpr int TC_THREADS = 352;
constexpr int TC_STAGES = PS_LMCONV_STAGES;  // ring stages of a full 128-row tile
constexpr int TC_MAX_STAGES = 12;            // a tile of fewer rows keeps a smaller A tile per stage and gets more stages
constexpr int TC_A_BYTES = 128 * 128;  // 128 rows x 64 fp16
constexpr int TC_W_BYTES = 160 * 128;  // up to 160 output channels x 64 fp16
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_W_BYTES;
constexpr int LMT_F = 80;
constexpr int LMT_CELLS = 1024;
constexpr int LMT_TENSORS = 33;
constexpr int LMT_ACT = 240;     // fp16 per (tensor, cell): elu(x) | elu(-x) | x
constexpr int LMT_CLASSES = 512;
constexpr int COL_OG = 320;      // TMEM columns [320, 400): the row's residual stream (fp32)
constexpr int COL_A = 400;       // TMEM columns [400, 496): centre-tap A operand of the next GEMM (192 fp16 per row)
constexpr int PROG_DONE = 34;    // progress word of a finished tile: 33 cache tensors + the sampled tokens
constexpr int TC_NOPS = 18;
constexpr int TC_MAX_CHUNKS = 724;
constexpr int LMT_PART_COLS = 3680;    // fp32 partial-sum columns per row: 14 x 80 + 14 x 160 + 4 x 80 (one block per GEMM)
constexpr int HALO_GROUP = 4;          // GEMMs per halo tile
constexpr int CHAIN_ROWS = 32;         // rows of a chain tile
constexpr int PART_ROW_BYTES = 640;    // one row's partial sums of one GEMM in shared memory (160 fp32)
constexpr int PART_BUF_BYTES = CHAIN_ROWS * PART_ROW_BYTES;
constexpr int XBUF_ROW_BYTES = 384;    // split epilogue: one row's next centre operand (192 fp16) in the exchange buffer
constexpr int XBUF_BYTES = CHAIN_ROWS * XBUF_ROW_BYTES;
constexpr int STAT_BYTES = 2 * 4 * CHAIN_ROWS * 8;  // split epilogue: per (parity, warp, row) a float2 of partial statistics

enum { A_GATHER = 0, A_CENTRE = 1, A_EPILOGUE = 2, A_TMEM = 3, A_REUSE = 4 };

// ps_lmconv_chunk in 8 bytes.  x: w_off16 (24) | w_rows / 8 (5) | a_kind (3);  y: a_tensor (6) | mask (2) | cin == 160 (1)
// | kc (5) | reads the raw third of the cache row (1) | d_col / 16 (5) | flags (6) | GEMM index (6)
struct Chunk {
  uint32_t w_off16;
  int w_rows, a_kind, a_tensor, mask, cin8, kc, ch_off8, d_col, flags, gemm;
};
__device__ __forceinline__ uint2 pack_chunk(const ps_lmconv_chunk& c) {
  uint2 r;
  r.x = (c.w_off16 & 0xffffffu) | ((uint32_t)(c.w_rows >> 3) << 24) | ((uint32_t)c.a_kind << 29);
  r.y = (uint32_t)c.a_tensor | ((uint32_t)c.mask << 6) | ((uint32_t)(c.cin8 == 20) << 8) | ((uint32_t)c.kc << 9) |
        ((uint32_t)(c.ch_off8 != 0) << 14) | ((uint32_t)(c.d_col >> 4) << 15) | ((uint32_t)(c.flags & 63) << 20) |
        ((uint32_t)(c.gemm & 63) << 26);
  return r;
}
__device__ __forceinline__ Chunk unpack_chunk(uint2 r) {
  Chunk c;
  c.w_off16 = r.x & 0xffffffu;
  c.w_rows = (int)((r.x >> 24) & 31u) << 3;
  c.a_kind = (int)(r.x >> 29) & 7;
  c.a_tensor = (int)(r.y & 63u);
  c.mask = (int)(r.y >> 6) & 3;
  c.cin8 = ((r.y >> 8) & 1u) ? 20 : 10;
  c.kc = (int)(r.y >> 9) & 31;
  c.ch_off8 = ((r.y >> 14) & 1u) ? 20 : 0;
  c.d_col = (int)((r.y >> 15) & 31u) << 4;
  c.flags = (int)(r.y >> 20) & 63;
  c.gemm = (int)(r.y >> 26) & 63;
  return c;
}
enum { FORM_NONE = 0, FORM_PAIR = 1, FORM_RAW = 2 };
enum { ROW_SAMPLED = 1u << 16, ROW_LOGITS = 1u << 17, ROW_VALID = 1u << 18 };
// One tile = up to 128 consecutive rows of one level.  prev_first / prev_count: the tiles of the previous level;
// wait_start: progress the previous level must have published before this tile does anything (0, all 33 cache
// tensors, or PROG_DONE = its tokens as well).
// kind_g: tile kind | first GEMM << 8 | end GEMM << 16 (halo tiles).  FULL: the whole column of its rows (known prefix,
// teacher-forced logits).  Sampled levels are split: a HALO tile multiplies the gathered neighbour taps of a group of
// GEMMs for up to 128 rows and stores per-row fp32 partial sums (no dependence on the rows' own chain, so it runs ahead,
// a step behind the previous level); a CHAIN tile runs the dependent part of up to 32 rows -- nin_skip, centre tap,
// epilogue, token -- and adds the partial sums.  part_row0: the tile's first row in the partial-sum buffer; h_first: a
// chain tile's halo tiles (one per GEMM group, consecutive).
enum { TILE_FULL = 0, TILE_CHAIN = 1, TILE_HALO = 2 };
struct Tile {
  int row_begin, nrows, prev_first, prev_count, wait_start, kind_g, part_row0, h_first;
};

struct TcParams {
  const unsigned char* wblob;
  const ps_lmconv_chunk* chunks;
  int n_body, n_total;
  int logit_first;   // schedule index of the first nin_out chunk (quarter 0, kc 0)
  const __half* w_uinit;  // [9][513][80]
  const float* bias;
  int b_uinit, b_nin;
  ps_lmconv_op ops[TC_NOPS];
  const ps_lmconv_chunk* chunks_chain;
  int n_chain_body, n_chain_total;
  const ps_lmconv_chunk* chunks_halo;
  short halo_first[33], part_col[33];
  int halo_group;   // GEMMs per halo tile (host: 4, or 8 on a small SM partition)
  float* part;  // [2 x part_cap rows][LMT_PART_COLS] fp32 partial sums of the halo tiles
  unsigned long long raw_mask;  // bit t: cached tensor t is read through its raw third (by a dilated convolution)
  __half* act;
  const ps_lmconv_row* rows;
  const Tile* tiles;
  int n_tiles;
  unsigned int* sync;  // [0] ticket counter, [16 + t] progress of tile t
  long long* codes;
  const float* uniforms;
  int ustride;
  float inv_temperature;
  float* logits_out;
  int debug;         // developer aid (PS_TC_DEBUG): bit0 skip the weight copies, bit1 skip the gather copies (timing only)
  long long* trace;  // developer aid: clock64 timestamps of the LAST tile (ps_lmconv_tc_set_trace), or null
};

#define TC_TRACE(slot, idx)                                              \
  do {                                                                   \
    if (p.trace && traced) p.trace[(slot) * 1024 + (idx)] = clock64();   \
  } while (0)

struct TcSmem {
  uint64_t full[TC_MAX_STAGES], empty[TC_MAX_STAGES], acc_full[3], cfull;
  uint64_t acc_empty[2];        // halo tiles: the epilogue has drained the accumulator (no centre-operand hand-off there)
  uint64_t pfull[2], pempty[2];  // chain tiles: the partial sums of a GEMM have landed in / been read from their buffer
  uint32_t tmem_slot, pad_steps;
  uint32_t warp_steps[4];  // finished cache tensors per epilogue warp (the tile's progress is their minimum)
  Tile tile;
  int tile_index, pad_;
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

__device__ __forceinline__ float exp2f_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// concat_elu of one value: (elu(x), elu(-x)) with a single exponential
__device__ __forceinline__ void celu(float x, float& p, float& n) {
  // expm1(-|x|) as exp - 1: the absolute error (~2e-7) is far below the fp16 rounding of the operand it becomes
  const float e = exp2f_fast(-1.4426950408889634f * fabsf(x)) - 1.0f;
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
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < LMT_F; i += 4) {
    s0 += v[i];
    s1 += v[i + 1];
    s2 += v[i + 2];
    s3 += v[i + 3];
  }
  const float mean = ((s0 + s1) + (s2 + s3)) * (1.0f / LMT_F);
  float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
  for (int i = 0; i < LMT_F; i += 4) {
    const float d0 = v[i] - mean, d1 = v[i + 1] - mean, d2 = v[i + 2] - mean, d3 = v[i + 3] - mean;
    q0 = fmaf(d0, d0, q0);
    q1 = fmaf(d1, d1, q1);
    q2 = fmaf(d2, d2, q2);
    q3 = fmaf(d3, d3, q3);
  }
  const float inv = rsqrtf(((q0 + q1) + (q2 + q3)) * (1.0f / (LMT_F - 1)) + 1e-5f);
#pragma unroll
  for (int i = 0; i < LMT_F; ++i) v[i] = (v[i] - mean) * inv;
}

// progress words: written by one thread per tile (release), polled by the tiles of the next level (acquire)
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_cta_shared(const uint32_t* p) {
  unsigned int v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_cta_shared_inc(uint32_t* p) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(smem_u32(p)) : "memory");
}

// Whole warp: returns once every tile of the previous level has published progress >= need; returns the smallest
// progress seen (so the caller can skip later polls).  A wait of ~1 s trips the same watchdog as the mbarriers.
// developer aid: when the watchdog trips, every warp still inside wait_progress logs what it waits for (pinned host words)
static __device__ __noinline__ void wedge_snapshot(int first, int count, unsigned int need, unsigned int seen, int tag) {
  volatile unsigned int* h = g_wedge_host;
  if (!h) return;
  const unsigned int slot = atomicAdd(&g_wedge[7], 1u);
  if (slot >= 256u) return;
  volatile unsigned int* e = h + 8 + slot * 8;
  e[0] = blockIdx.x;
  e[1] = threadIdx.x;
  e[2] = (unsigned int)first;
  e[3] = (unsigned int)count;
  e[4] = need;
  e[5] = seen;
  e[6] = (unsigned int)tag;
  e[7] = 0xabcd0000u | slot;
  __threadfence_system();
}

__device__ __noinline__ unsigned int wait_progress(const unsigned int* prog, int first, int count, unsigned int need, int tag = 0) {
  const int lane = threadIdx.x & 31;
  const long long t0 = clock64();
  for (unsigned int spins = 0;; ++spins) {
    unsigned int v = 0xffffffffu;
    for (int t = first + lane; t < first + count; t += 32) v = min(v, ld_acquire_gpu(prog + t));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (v >= need) return v;
    if ((spins & 63u) == 63u) {
      if (*(volatile unsigned int*)&g_wedge[0]) {
        if (lane == 0) wedge_snapshot(first, count, need, v, tag);
        return 0xffffffffu;
      }
      if (clock64() - t0 > 2000000000ll) {
        if (lane == 0) {
          wedge_report(0xffffffffu, need);
          wedge_snapshot(first, count, need, v, tag);
        }
        return 0xffffffffu;
      }
    }
    __nanosleep(64);
  }
}

// 16-byte group kg (0..7) of row r of the A tile in the ring stage of schedule chunk `chunk` (128-byte swizzle)
__device__ __forceinline__ void sts_a(unsigned char* tiles, int slot_offset, int kg, int r, uint4 v) {
  const uint32_t addr = smem_u32(tiles + slot_offset) + r * 128 + ((kg ^ (r & 7)) << 4);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

struct Epi {
  TcSmem* sm;
  uint32_t tlane;   // TMEM address of this thread's lane, column 0
  bool valid;
  int dbg;
  __half* actrow;  // act + (b * 33 * 1024 + cell) * 240; tensor t adds t * 1024 * 240

  // the centre operand of the next GEMM is complete in TMEM: one arrival per epilogue warp
  __device__ __forceinline__ void publish() const {
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&sm->cfull);
  }
  // the cache rows of one more tensor have been written by this warp
  __device__ __forceinline__ void step_done() const {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) red_release_cta_shared_inc(&sm->warp_steps[(threadIdx.x >> 5) & 3]);
  }
  // One 16-channel piece j (channels 16j..16j+15) of a finished tensor: cache write + centre operand of the next GEMM.
  __device__ __forceinline__ void emit16(int form, int j, const float* x, int tensor, bool raw) const {
    uint32_t pp[8], nn[8], rr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float p0, n0, p1, n1;
      if (dbg & 8) {
        p0 = n0 = x[2 * i];
        p1 = n1 = x[2 * i + 1];
      } else {
        celu(x[2 * i], p0, n0);
        celu(x[2 * i + 1], p1, n1);
      }
      pp[i] = pack_h2(p0, p1);
      nn[i] = pack_h2(n0, n1);
      rr[i] = pack_h2(x[2 * i], x[2 * i + 1]);
    }
    if (form == FORM_PAIR) {  // K = [elu(x) 0..79 | elu(-x) 80..159 | 0 .. 191]: two fp16 per TMEM column
      tmem_st8_nowait(tlane + COL_A + 8 * j, pp);
      tmem_st8_nowait(tlane + COL_A + 40 + 8 * j, nn);
    } else if (form == FORM_RAW) {  // K = [x 0..79 | 0 .. 127]
      const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      tmem_st8_nowait(tlane + COL_A + 8 * j, rr);
      if (j < 3) tmem_st8_nowait(tlane + COL_A + 40 + 8 * j, z);
    }
    if (valid && !(dbg & 4)) {
      uint4* g = reinterpret_cast<uint4*>(actrow + (size_t)tensor * LMT_CELLS * LMT_ACT);
      g[2 * j] = make_uint4(pp[0], pp[1], pp[2], pp[3]);
      g[2 * j + 1] = make_uint4(pp[4], pp[5], pp[6], pp[7]);
      g[10 + 2 * j] = make_uint4(nn[0], nn[1], nn[2], nn[3]);
      g[10 + 2 * j + 1] = make_uint4(nn[4], nn[5], nn[6], nn[7]);
      if (raw) {
        g[20 + 2 * j] = make_uint4(rr[0], rr[1], rr[2], rr[3]);
        g[20 + 2 * j + 1] = make_uint4(rr[4], rr[5], rr[6], rr[7]);
      }
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

  // tiles are handed out in level order: whatever this tile waits for is already running or finished
  if (tid == 0) {
    const int t = (int)atomicAdd(p.sync, 1u);
    sm.tile_index = t;
    sm.tile = p.tiles[t];
    g_blk_tag[blockIdx.x & 16383] = (unsigned int)t | ((unsigned int)(sm.tile.kind_g & 255) << 24);
    sm.warp_steps[0] = sm.warp_steps[1] = sm.warp_steps[2] = sm.warp_steps[3] = 0;
  }
  __syncthreads();
  const Tile tile = sm.tile;
  const int tile_index = sm.tile_index;
  const int tkind = tile.kind_g & 255, hg0 = (tile.kind_g >> 8) & 255, hg1 = (tile.kind_g >> 16) & 255;
  const bool is_chain = tkind == TILE_CHAIN, is_halo = tkind == TILE_HALO;
  // the tile's chunk schedule: the whole column, its chain part, or the gathered taps of GEMMs [hg0, hg1)
  const ps_lmconv_chunk* sched_src = is_halo ? p.chunks_halo + p.halo_first[hg0] : (is_chain ? p.chunks_chain : p.chunks);
  const int n_sched_total = is_halo ? p.halo_first[hg1] - p.halo_first[hg0] : (is_chain ? p.n_chain_total : p.n_total);
  const int n_sched_body = is_halo ? n_sched_total : (is_chain ? p.n_chain_body : p.n_body);
  const int logit_first = is_chain ? p.n_chain_body : p.logit_first;
  for (int i = tid; i < n_sched_total; i += TC_THREADS) sm.sched[i] = pack_chunk(sched_src[i]);
  const bool traced = tile_index == p.n_tiles - 1;
  const unsigned int* prog = p.sync + 16;
  // Ring geometry: the A tile of a stage holds only the tile's rows (rounded up to 8); the MMA still reads 128 rows
  // and runs on into the stage's weight tile, which only feeds accumulator lanes nobody looks at.
  // A ring stage holds TWO consecutive chunks (slot = chunk & 1), so the issuer pays one barrier wait, one proxy
  // fence and one commit per eight MMAs.
  // Chain tiles (<= 32 rows) run the "split" epilogue: the rows are REPLICATED in the four 32-lane TMEM quadrants (every
  // A operand carries the 32 rows four times), and epilogue warp q computes channel groups q, q + 4, q + 8 of every row
  // instead of one warp computing all 80 / 160 channels: four SM sub-partitions share the transcendental and issue
  // work of the dependent chain.  PS_TC_DEBUG bit 15 keeps the one-thread-per-row epilogue.
  const bool split_epi = is_chain && !(p.debug & 32768);
  const int a_rows = split_epi ? 128 : (tile.nrows + 7) & ~7;
  const int a_bytes = a_rows * 128;
  const int slot_bytes = a_bytes + TC_W_BYTES;
  const int stage_bytes = 2 * slot_bytes;
  // a chain tile keeps two partial-sum buffers behind its (short) ring
  const int ring_cap = TC_STAGES * TC_STAGE_BYTES - (is_chain ? 2 * PART_BUF_BYTES : 0) -
                       (split_epi ? 2 * XBUF_BYTES + STAT_BYTES : 0);
  const int nst = min(TC_MAX_STAGES, ring_cap / stage_bytes);
  unsigned char* pbuf = tiles + ring_cap;
  unsigned char* xbuf = pbuf + 2 * PART_BUF_BYTES;   // split epilogue only
  unsigned char* statbuf = xbuf + 2 * XBUF_BYTES;
  // byte offset of chunk i's slot in the ring
  auto slot_off = [&](int i) { return ((i >> 1) % nst) * stage_bytes + (i & 1) * slot_bytes; };

  if (tid < 128) {
    ps_lmconv_row ri;
    ri.bc = 0;
    ri.w01 = 0;
    ri.w2_flags = 0;
    ri.uidx = 0;
    if (tid < tile.nrows) ri = p.rows[tile.row_begin + tid];
    sm.rows[tid] = ri;
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < nst; ++s) {
        // per chunk: 32 arrivals of the row writers (a gather warp, or the weight producer's lanes for an operand that
        // is not in the ring) + the weight copy; per stage use: one arrival of each of the four gather warps
        mbar_init(&sm.full[s], 70);
        mbar_init(&sm.empty[s], 1);
      }
      for (int i = 0; i < 3; ++i) mbar_init(&sm.acc_full[i], 1);
      mbar_init(&sm.cfull, 4);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&sm.acc_empty[i], 4);
        mbar_init(&sm.pfull[i], 128);
        mbar_init(&sm.pempty[i], 4);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(&sm.tmem_slot, 512);
  }
  tc_fence_before();
  const int need_logits =
      __syncthreads_or(!is_halo && tid < tile.nrows && (p.rows[tile.row_begin + tid].w2_flags & (ROW_SAMPLED | ROW_LOGITS)));
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, sm.tmem_slot, 0);
  if (tid == 0) TC_TRACE(7, 0);
  const int nchunks = need_logits ? n_sched_total : n_sched_body;

  if (warp == 0) {
    // ===== weight producer: whole warp, warp-uniform values, one elected lane issues (see umma_f16_kblock) =====
    int st = 0;
    uint32_t ph = 1;
    for (int i = 0; i < nchunks; ++i) {
      uint2 raw = sm.sched[i];
      raw.x = __shfl_sync(0xffffffffu, raw.x, 0);  // warp-uniform by construction; now also to the compiler
      if (!(i & 1)) mbar_wait(&sm.empty[st], ph);  // the stage's second chunk rides on the same phase
      if (lane == 0) TC_TRACE(3, i);
      const uint32_t bytes = ((raw.x >> 24) & 31u) << 10;  // w_rows * 128
      const uint32_t kind = raw.x >> 29;
      if (kind == A_TMEM || kind == A_REUSE) mbar_arrive(&sm.full[st]);  // nobody writes rows into this slot
      if (p.debug & 1) {
        if (lane == 0) mbar_arrive(&sm.full[st]);
      } else {
        bulk_load_elect(tiles + (size_t)st * stage_bytes + (i & 1) * slot_bytes + a_bytes,
                        p.wblob + (size_t)(raw.x & 0xffffffu) * 16, bytes, &sm.full[st]);
      }
      if (i & 1) {
        if (++st == nst) {
          st = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop with warp-uniform values; one elected lane issues =====
    const uint32_t ring_lo = umma_desc_lo(smem_u32(tiles));
    const uint32_t idesc0 = umma_idesc_f16(0);
    const uint32_t reuse_lo0 = ring_lo + (uint32_t)(slot_off(logit_first) >> 4);
    const uint32_t reuse_lo1 = ring_lo + (uint32_t)(slot_off(logit_first + 1) >> 4);
    int st = 0;
    uint32_t ph = 0, cph = 0;
    uint2 nxt = sm.sched[0];
    for (int i = 0; i < nchunks; i += 2) {
      mbar_wait(&sm.full[st], ph);
      if (!(p.debug & 64)) fence_proxy_async();
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint2 raw = nxt;
        if (i + h + 1 < nchunks) nxt = sm.sched[i + h + 1];
        raw.x = __shfl_sync(0xffffffffu, raw.x, 0);  // warp-uniform by construction; now also to the compiler
        raw.y = __shfl_sync(0xffffffffu, raw.y, 0);
        const uint32_t flags = (raw.y >> 20) & 63u;
        const uint32_t kind = raw.x >> 29;
        if (flags & 16u) {  // first centre chunk of a GEMM: the previous epilogue has finished the operand in TMEM
          mbar_wait(&sm.cfull, cph);
          cph ^= 1u;
          tc_fence_after();
        }
        if (is_halo && !(flags & 1u)) {  // halo tile, first chunk of a GEMM: its accumulator has been drained
          const int n = (int)((raw.y >> 26) & 63u) - hg0;
          if (n >= 2) {
            mbar_wait(&sm.acc_empty[n & 1], (uint32_t)((n >> 1) - 1) & 1u);
            tc_fence_after();
          }
        }
        if (lane == 0) TC_TRACE(0, i + h);
        const uint32_t idesc = idesc0 | (((p.debug & 16) ? 2u : ((raw.x >> 24) & 31u)) << 17);  // N >> 3 = w_rows / 8
        const uint32_t d = tmem_base + (((raw.y >> 15) & 31u) << 4);          // d_col
        const uint32_t a_lo = ring_lo + (uint32_t)((st * stage_bytes + h * slot_bytes) >> 4);
        const uint32_t b_lo = a_lo + (uint32_t)(a_bytes >> 4);
        const uint32_t kc = (raw.y >> 9) & 31u;
        if (p.debug & 32) {
        } else if (kind == A_TMEM) {
          umma_f16_ts_kblock_nc(d, tmem_base + COL_A + kc * 32u, b_lo, idesc, flags & 1u);
        } else if (kind == A_REUSE) {
          // the operand the epilogue wrote for nin_out's first quarter stays in that quarter's slots
          umma_f16_kblock_nc(d, kc ? reuse_lo1 : reuse_lo0, b_lo, idesc, flags & 1u);
        } else {
          umma_f16_kblock_nc(d, a_lo, b_lo, idesc, flags & 1u);
        }
        if (flags & 2u) umma_commit_elect(&sm.acc_full[(flags >> 2) & 3u]);
      }
      umma_commit_elect(&sm.empty[st]);  // frees the stage once these eight MMAs have read it
      if (++st == nst) {
        st = 0;
        ph ^= 1u;
      }
    }
  } else if (warp == 10) {
    // ===== progress publisher: every epilogue warp counts its finished cache tensors in shared memory; this warp makes
    // them visible device-wide (the fence is cumulative over what it observed) and moves the tile's progress word =====
    if (lane == 0) {
      unsigned int* mine = p.sync + 16 + tile_index;
      unsigned int published = 0;
      const unsigned int target = is_halo ? (unsigned int)(hg1 - hg0) : (unsigned int)PROG_DONE;
      while (published < target) {
        // the slowest epilogue warp's count (a sum over the warps would overstate it when one warp lags)
        const unsigned int done =
            min(min(ld_acquire_cta_shared(&sm.warp_steps[0]), ld_acquire_cta_shared(&sm.warp_steps[1])),
                min(ld_acquire_cta_shared(&sm.warp_steps[2]), ld_acquire_cta_shared(&sm.warp_steps[3])));
        if (done > published) {
          __threadfence();
          st_release_gpu(mine, done);
          published = done;
        } else {
          __nanosleep(32);
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
      const ps_lmconv_row ri = sm.rows[split_epi ? (t & 31) : t];
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
    const int npair = a_rows >> 3;  // row pairs (rs + 8m, rs + 4 + 8m) of this lane
    int seen = 0;                              // gathered chunks so far: this warp takes those with seen % 4 == gw
    int cur_gemm = -1;                         // chain tiles: GEMM whose partial sums were requested last
    unsigned int verified = tile.prev_count ? 0u : 0xffffffffu;  // progress every earlier level is known to have reached
    if (verified < (unsigned int)tile.wait_start)
      verified = wait_progress(prog, tile.prev_first, tile.prev_count, (unsigned int)tile.wait_start, tile_index | tkind << 24 | 1 << 28);
    for (int i = 0; i < nchunks; ++i) {
      const int st = (i >> 1) % nst;
      // EVERY gather warp waits for EVERY use of every ring stage, in ring order, whether or not it fills a chunk there.
      // A parity wait is only unambiguous for a waiter that is at most one phase behind the barrier.  A warp that
      // only waited where it owned a chunk could jump 7 chunks ahead (its next chunk lies behind a GEMM's three
      // centre-tap chunks), i.e. 4 stage uses, on a 3-stage ring (128-row tiles): with the MMA warp parked on the
      // epilogue's operand the `empty` barrier was then still TWO phases back, the parity test passed, the warp
      // overwrote a stage that had not been multiplied yet and its 32 arrivals landed in the wrong phase of `full`
      // -- wrong activations, and sooner or later an mbarrier arrival overflow = "unspecified launch failure"
      // (round 1's intermittent fault at batch 128; profiles/r02_launch_failure_rootcause.txt).
      // ... and every gather warp ARRIVES on every use as well: the MMA warp cannot finish a stage use without it, so a
      // gather warp that is held up elsewhere (a chain tile's wait for its halo tile, a progress wait) can never fall a
      // phase BEHIND the ring either -- stage uses whose chunks nobody gathers (centre taps) would otherwise run past it.
      if (!(i & 1)) {
        mbar_wait(&sm.empty[st], ((uint32_t)((i >> 1) / nst) & 1u) ^ 1u);
        if (lane == 0) mbar_arrive(&sm.full[st]);
      }
      const Chunk ch = unpack_chunk(sm.sched[i]);
      if (is_chain && ch.gemm != cur_gemm && ch.gemm < 32 && !(p.debug & 1024)) {
        // Chain tile, first chunk of GEMM gg: bring the rows' partial sums of its gathered taps (written by the halo
        // tile of the GEMM's group) into buffer gg & 1.  All four gather warps: thread t -> row t / 4, every 4th
        // 16-byte group; group k of row r sits at group k ^ (r & 7) so the epilogue's 16-byte row reads spread over
        // the banks.  Every gather thread passes every use of the two buffers in order (parity waits stay unambiguous).
        const int gg = cur_gemm = ch.gemm;
        if (!(p.debug & 2048)) wait_progress(prog, tile.h_first + gg / p.halo_group, 1, (unsigned int)(gg % p.halo_group + 1), tile_index | tkind << 24 | 2 << 28);
        if (!(p.debug & 4096)) mbar_wait(&sm.pempty[gg & 1], ((uint32_t)(gg >> 1) & 1u) ^ 1u);
        const int c0 = p.part_col[gg], ngrp = (p.part_col[gg + 1] - c0) >> 2;
        const int row = t >> 2;
        const float* src = p.part + (size_t)(tile.part_row0 + row) * LMT_PART_COLS + c0;
        const uint32_t dst = smem_u32(pbuf) + (uint32_t)((gg & 1) * PART_BUF_BYTES + row * PART_ROW_BYTES);
        const uint32_t nb = row < tile.nrows ? 16u : 0u;
        for (int k = t & 3; k < ngrp; k += 4)
          cp_async16_zfill(dst + (uint32_t)((k ^ (row & 7)) << 4), nb ? (const void*)(src + 4 * k) : (const void*)p.part, nb);
        if (p.debug & 16384) {
          cp_async_commit();
          cp_async_wait<0>();
          mbar_arrive(&sm.pfull[gg & 1]);
        } else {
          cp_async_arrive_noinc(&sm.pfull[gg & 1]);
        }
      }
      if (ch.a_kind != A_GATHER && ch.a_kind != A_CENTRE) continue;
      if ((seen++ & 3) != gw) continue;
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
        // GEMM j reads tensors the neighbours wrote at step <= j.  Waiting until the previous level is a step further
        // (progress >= j + 2) also proves, by induction over the levels, that every earlier level has reached j + 1.
        // (A halo tile's previous level is the chain of the level before, which started only when everything older was
        // complete: progress >= j + 1 is enough there.)
        const unsigned int need = (unsigned int)min(ch.gemm + (is_halo ? 1 : 2), LMT_TENSORS);
        if (verified < need) verified = wait_progress(prog, tile.prev_first, tile.prev_count, need, tile_index | tkind << 24 | 3 << 28);
      } else {
        bitpos = kg < ch.cin8 ? 27 : 31;
        off = ch.a_tensor * LMT_CELLS * (LMT_ACT * 2) + (ch.ch_off8 + kg) * 16;
      }
      if (lane == 0) TC_TRACE(1, i);
      const uint32_t soff = st * stage_bytes + (i & 1) * slot_bytes;
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
      cp_async_arrive_noinc(&sm.full[st]);
      if (lane == 0) TC_TRACE(2, i);
    }
  } else {
    // ===== epilogue: one thread per row =====
    Epi e;
    e.sm = &sm;
    e.dbg = p.debug;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    e.tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const ps_lmconv_row ri = sm.rows[split_epi ? lane : r];
    e.valid = (ri.w2_flags & ROW_VALID) != 0;
    const int b = ri.bc >> 10, cell = ri.bc & 1023;
    e.actrow = p.act + ((size_t)b * LMT_TENSORS * LMT_CELLS + cell) * LMT_ACT;
    const float* bias = p.bias;
    int g = 0;  // GEMM counter
    // chain tiles: the partial sums of GEMM gi's gathered taps wait in shared-memory buffer gi & 1 (rows 0..31 only)
    const bool prow_ok = is_chain && r < CHAIN_ROWS && !(p.debug & 1024);  // warp-uniform
    const unsigned char* prow = pbuf + r * PART_ROW_BYTES;
    auto part_wait = [&](int gi) {
      if (is_chain && !(p.debug & (1024 | 8192))) mbar_wait(&sm.pfull[gi & 1], (uint32_t)(gi >> 1) & 1u);
    };
    auto part_add16 = [&](float* v, int gi, int c16) {  // v[0..16) += columns [16 c16, 16 c16 + 16) of GEMM gi's partial sums
      if (prow_ok) {
        const unsigned char* b = prow + (gi & 1) * PART_BUF_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 f = *reinterpret_cast<const float4*>(b + (((4 * c16 + q) ^ (r & 7)) << 4));
          v[4 * q] += f.x;
          v[4 * q + 1] += f.y;
          v[4 * q + 2] += f.z;
          v[4 * q + 3] += f.w;
        }
      }
    };
    auto part_release = [&](int gi) {
      if (is_chain && !(p.debug & 1024)) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.pempty[gi & 1]);
      }
    };

    if (is_halo) {
      // ---- halo tile: each GEMM's accumulator (the gathered taps only) goes to the partial-sum buffer as it is ----
      float* out_row = p.part + (size_t)(tile.part_row0 + r) * LMT_PART_COLS;
      for (int gi = hg0; gi < hg1; ++gi) {
        const int n = gi - hg0;
        const uint32_t col0 = (uint32_t)(gi & 1) * 160u;
        mbar_wait(&sm.acc_full[gi & 1], (uint32_t)(n >> 1) & 1u);
        tc_fence_after();
        const int c0 = p.part_col[gi], ncol = p.part_col[gi + 1] - c0;
        for (int c = 0; c < ncol; c += 16) {
          float v[16];
          tmem_ld16_nowait(e.tlane + col0 + c, v);
          tmem_ld_wait();
          if (e.valid) {
            float4* d = reinterpret_cast<float4*>(out_row + c0 + c);
#pragma unroll
            for (int q = 0; q < 4; ++q) __stcg(d + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.acc_empty[gi & 1]);
        e.step_done();
      }
    } else {

    // the first layer reads the neighbours' tokens: a level with sampled cells starts when the previous one has drawn
    if (tile.wait_start && tile.prev_count)
      wait_progress(prog, tile.prev_first, tile.prev_count, (unsigned int)tile.wait_start, tile_index | tkind << 24 | 4 << 28);

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
        // written by other SMs in this launch; the mask keeps the table lookup in bounds even if the watchdog has
        // released a wedged wait and the token is garbage
        const int code = (int)__ldcg(p.codes + (size_t)b * LMT_CELLS + nbr) & (LMT_CLASSES - 1);
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
      {  // K columns 160..191 of the centre operand stay zero for the whole tile
        const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        tmem_st8_nowait(e.tlane + COL_A + 80, z);
        tmem_st8_nowait(e.tlane + COL_A + 88, z);
      }
      const bool row_valid = e.valid;
      if (split_epi && quad != 0) e.valid = false;  // the four replicas of a row compute the same u_init: one writes the cache
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        tmem_st16_nowait(e.tlane + COL_OG + 16 * j, v + 16 * j);
        e.emit16(FORM_PAIR, j, v + 16 * j, 0, (p.raw_mask >> 0) & 1ull);
      }
      e.valid = row_valid;
      e.publish();
      e.step_done();
    }

    if (split_epi) {
      // ================= split epilogue of a chain tile =================
      // Warp `quad` owns channel groups (8 channels each) quad, quad + 4 and, for quad < 2, quad + 8 of all 32 rows; lane =
      // row.  What needs a whole row -- the PONO statistics, the next centre operand (every replica lane needs all of it)
      // -- goes through double-buffered shared memory and one named barrier of the four epilogue warps per exchange.
      const int ng = quad < 2 ? 3 : 2;
      float2* st2 = reinterpret_cast<float2*>(statbuf);
      int sp = 0, xq = 0;  // parities of the statistics / operand exchange buffers
      auto bar_epi = [] { asm volatile("bar.sync 2, 128;" ::: "memory"); };
      const unsigned char* prow_l = pbuf + lane * PART_ROW_BYTES;
      // x[8 i .. 8 i + 8) += 8 fp32 of GEMM gi's partial sums starting at column col8 * 8 (two 16-byte chunks)
      auto part8 = [&](float* x, int gi, int col8) {
        const unsigned char* b = prow_l + (gi & 1) * PART_BUF_BYTES;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 f = *reinterpret_cast<const float4*>(b + (((2 * col8 + h) ^ (lane & 7)) << 4));
          x[4 * h] += f.x;
          x[4 * h + 1] += f.y;
          x[4 * h + 2] += f.z;
          x[4 * h + 3] += f.w;
        }
      };
      auto bias8 = [&](float* x, const float* bp) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(bp)), c = __ldg(reinterpret_cast<const float4*>(bp) + 1);
        x[0] += a.x, x[1] += a.y, x[2] += a.z, x[3] += a.w, x[4] += c.x, x[5] += c.y, x[6] += c.z, x[7] += c.w;
      };
      // positional normalisation over the row's 80 channels, 24 or 16 of them here (layers.py:224-236, unbiased variance)
      auto pono_split = [&](float* x) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (i < ng) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              s1 += x[8 * i + k];
              s2 = fmaf(x[8 * i + k], x[8 * i + k], s2);
            }
          }
        st2[(sp * 4 + quad) * 32 + lane] = make_float2(s1, s2);
        bar_epi();
        float S1 = 0.f, S2 = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 tq = st2[(sp * 4 + q) * 32 + lane];
          S1 += tq.x;
          S2 += tq.y;
        }
        sp ^= 1;
        const float mean = S1 * (1.0f / LMT_F);
        const float inv = rsqrtf(fmaxf(S2 - S1 * mean, 0.f) * (1.0f / (LMT_F - 1)) + 1e-5f);
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (i < ng) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[8 * i + k] = (x[8 * i + k] - mean) * inv;
          }
      };
      // cache rows of this warp's channel groups + (form != NONE) the next GEMM's centre operand for every replica lane
      auto emit_split = [&](int form, const float* x, int tensor, bool raw) {
        uint4* crow = e.valid ? reinterpret_cast<uint4*>(e.actrow + (size_t)tensor * LMT_CELLS * LMT_ACT) : nullptr;
        unsigned char* xrow = xbuf + xq * XBUF_BYTES + lane * XBUF_ROW_BYTES;
        auto xchunk = [&](int c) { return reinterpret_cast<uint4*>(xrow + ((c ^ (lane & 7)) << 4)); };
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (i < ng) {
            const int gr = quad + 4 * i;
            uint32_t pp[4], nn[4], rr[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float p0, n0, p1, n1;
              celu(x[8 * i + 2 * k], p0, n0);
              celu(x[8 * i + 2 * k + 1], p1, n1);
              pp[k] = pack_h2(p0, p1);
              nn[k] = pack_h2(n0, n1);
              rr[k] = pack_h2(x[8 * i + 2 * k], x[8 * i + 2 * k + 1]);
            }
            if (crow) {
              crow[gr] = make_uint4(pp[0], pp[1], pp[2], pp[3]);
              crow[10 + gr] = make_uint4(nn[0], nn[1], nn[2], nn[3]);
              if (raw) crow[20 + gr] = make_uint4(rr[0], rr[1], rr[2], rr[3]);
            }
            if (form == FORM_PAIR) {
              *xchunk(gr) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
              *xchunk(10 + gr) = make_uint4(nn[0], nn[1], nn[2], nn[3]);
            } else if (form == FORM_RAW) {
              *xchunk(gr) = make_uint4(rr[0], rr[1], rr[2], rr[3]);
            }
          }
        if (form == FORM_NONE) return;
        bar_epi();
        const int nchunk2 = form == FORM_PAIR ? 10 : 5;  // pairs of 16-byte chunks = 8 TMEM columns each
#pragma unroll
        for (int j = 0; j < 10; ++j)
          if (j < nchunk2) {
            const uint4 a = *xchunk(2 * j), c = *xchunk(2 * j + 1);
            const uint32_t v8[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
            tmem_st8_nowait(e.tlane + COL_A + 8 * j, v8);
          }
        if (form == FORM_RAW) {  // K = [x 0..79 | 0 .. 127]
          const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
          for (int j = 0; j < 3; ++j) tmem_st8_nowait(e.tlane + COL_A + 40 + 8 * j, z);
        }
        xq ^= 1;
      };

      for (int oi = 0; oi < TC_NOPS; ++oi) {
        const ps_lmconv_op op = p.ops[oi];
        const int next_form = oi + 1 < TC_NOPS ? (p.ops[oi + 1].kind == 0 ? FORM_PAIR : FORM_RAW) : FORM_NONE;
        const bool out_raw = (p.raw_mask >> op.out) & 1ull;
        {  // first GEMM of the op: conv_input (+ nin_skip) of a gated resnet, or the dilated convolution
          const uint32_t col0 = (uint32_t)(g & 1) * 160u;
          mbar_wait(&sm.acc_full[g & 1], (uint32_t)(g >> 1) & 1u);
          if (r == 0) TC_TRACE(4, g);
          tc_fence_after();
          float x[24];
#pragma unroll
          for (int i = 0; i < 3; ++i)
            if (i < ng) tmem_ld8_nowait(e.tlane + col0 + 8 * (quad + 4 * i), x + 8 * i);
          tmem_ld_wait();
          mbar_wait(&sm.pfull[g & 1], (uint32_t)(g >> 1) & 1u);
#pragma unroll
          for (int i = 0; i < 3; ++i)
            if (i < ng) {
              bias8(x + 8 * i, bias + op.b_in + 8 * (quad + 4 * i));
              part8(x + 8 * i, g, quad + 4 * i);
            }
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.pempty[g & 1]);
          pono_split(x);
          if (op.kind == 0) {
            if (op.a >= 0) {
              float sk[24];
#pragma unroll
              for (int i = 0; i < 3; ++i)
                if (i < ng) tmem_ld8_nowait(e.tlane + col0 + 80 + 8 * (quad + 4 * i), sk + 8 * i);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 3; ++i)
                if (i < ng) {
                  bias8(sk + 8 * i, bias + op.b_skip + 8 * (quad + 4 * i));
#pragma unroll
                  for (int k = 0; k < 8; ++k) x[8 * i + k] += sk[8 * i + k];
                }
            }
            if (r == 0) TC_TRACE(6, g);
            emit_split(FORM_PAIR, x, op.mid, false);
            e.publish();
          } else {  // the dilated convolution's PONO output becomes the new residual stream
#pragma unroll
            for (int i = 0; i < 3; ++i)
              if (i < ng) {
                tmem_st8_nowait(e.tlane + COL_OG + 8 * (quad + 4 * i), reinterpret_cast<const uint32_t*>(x + 8 * i));
              }
            emit_split(next_form, x, op.out, out_raw);
            if (next_form != FORM_NONE) e.publish(); else { tmem_st_wait(); tc_fence_before(); }
          }
          e.step_done();
          if (r == 0) TC_TRACE(5, g);
          ++g;
        }
        if (op.kind == 0) {  // y = conv_out(concat_elu(x)); og += PONO(y[:80]) * sigmoid(y[80:])
          const uint32_t col0 = (uint32_t)(g & 1) * 160u;
          mbar_wait(&sm.acc_full[g & 1], (uint32_t)(g >> 1) & 1u);
          if (r == 0) TC_TRACE(4, g);
          tc_fence_after();
          float a[24], gt[24], o[24];
#pragma unroll
          for (int i = 0; i < 3; ++i)
            if (i < ng) {
              tmem_ld8_nowait(e.tlane + col0 + 8 * (quad + 4 * i), a + 8 * i);
              tmem_ld8_nowait(e.tlane + col0 + 80 + 8 * (quad + 4 * i), gt + 8 * i);
              tmem_ld8_nowait(e.tlane + COL_OG + 8 * (quad + 4 * i), o + 8 * i);
            }
          tmem_ld_wait();
          mbar_wait(&sm.pfull[g & 1], (uint32_t)(g >> 1) & 1u);
#pragma unroll
          for (int i = 0; i < 3; ++i)
            if (i < ng) {
              bias8(a + 8 * i, bias + op.b_out + 8 * (quad + 4 * i));
              part8(a + 8 * i, g, quad + 4 * i);
              bias8(gt + 8 * i, bias + op.b_out + 80 + 8 * (quad + 4 * i));
              part8(gt + 8 * i, g, 10 + quad + 4 * i);
            }
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.pempty[g & 1]);
          pono_split(a);
#pragma unroll
          for (int i = 0; i < 3; ++i)
            if (i < ng) {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                o[8 * i + k] = fmaf(a[8 * i + k], __fdividef(1.0f, 1.0f + __expf(-gt[8 * i + k])), o[8 * i + k]);
              tmem_st8_nowait(e.tlane + COL_OG + 8 * (quad + 4 * i), reinterpret_cast<const uint32_t*>(o + 8 * i));
            }
          emit_split(next_form, o, op.out, out_raw);
          if (next_form != FORM_NONE) e.publish(); else { tmem_st_wait(); tc_fence_before(); }
          e.step_done();
          if (r == 0) TC_TRACE(5, g);
          ++g;
        }
      }

      if (need_logits) {
        // ---- logits = nin_out(elu(u)): the operand A = [elu(u) | 0] of every replica row goes into quarter 0's ring slots
        const int first = logit_first;
        {
          unsigned char* xrow = xbuf + xq * XBUF_BYTES + lane * XBUF_ROW_BYTES;
          auto xchunk = [&](int c) { return reinterpret_cast<uint4*>(xrow + ((c ^ (lane & 7)) << 4)); };
          float o[24];
#pragma unroll
          for (int i = 0; i < 3; ++i)
            if (i < ng) tmem_ld8_nowait(e.tlane + COL_OG + 8 * (quad + 4 * i), o + 8 * i);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 3; ++i)
            if (i < ng) {
              uint32_t pp[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float p0, n0, p1, n1;
                celu(o[8 * i + 2 * k], p0, n0);
                celu(o[8 * i + 2 * k + 1], p1, n1);
                pp[k] = pack_h2(p0, p1);
              }
              *xchunk(quad + 4 * i) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
            }
          bar_epi();
#pragma unroll
          for (int kg = 0; kg < 10; ++kg) sts_a(tiles, slot_off(first + (kg >> 3)), kg & 7, r, *xchunk(kg));
#pragma unroll
          for (int kg = 10; kg < 16; ++kg) sts_a(tiles, slot_off(first + 1), kg & 7, r, make_uint4(0, 0, 0, 0));
          xq ^= 1;
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
          if (lane < 8) {
            mbar_arrive(&sm.full[(first >> 1) % nst]);
            mbar_arrive(&sm.full[((first + 1) >> 1) % nst]);
          }
        }
        mbar_wait(&sm.acc_full[2], 0);
        tc_fence_after();
        // ---- the draw, split over the warps: warp `quad` owns classes [128 quad, 128 quad + 128) of every row ----
        const bool sampled = e.valid && (ri.w2_flags & ROW_SAMPLED) && p.uniforms;
        const int cq = quad * 128;
        float mx = -INFINITY, sum = 0.f;
        for (int c0 = 0; c0 < 128; c0 += 16) {
          float l[16];
          tmem_ld16_nowait(e.tlane + cq + c0, l);
          tmem_ld_wait();
          float m2 = mx;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            l[i] = (l[i] + __ldg(bias + p.b_nin + cq + c0 + i)) * p.inv_temperature;
            m2 = fmaxf(m2, l[i]);
          }
          sum *= __expf(mx - m2);
#pragma unroll
          for (int i = 0; i < 16; ++i) sum += __expf(l[i] - m2);
          mx = m2;
        }
        st2[(sp * 4 + quad) * 32 + lane] = make_float2(mx, sum);
        bar_epi();
        float M = -INFINITY;
#pragma unroll
        for (int q = 0; q < 4; ++q) M = fmaxf(M, st2[(sp * 4 + q) * 32 + lane].x);
        float before = 0.f, mine = 0.f, total = 0.f;  // softmax mass of the quarters before this one / of this one / of all
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 tq = st2[(sp * 4 + q) * 32 + lane];
          const float w = tq.y * __expf(tq.x - M);
          if (q < quad) before += w;
          if (q == quad) mine = w;
          total += w;
        }
        sp ^= 1;
        // token = first class whose cumulative mass exceeds u * total; the quarter that holds the crossing finds it
        const float thr = sampled ? p.uniforms[(size_t)b * p.ustride + ri.uidx] * total : 0.f;
        const bool claim = sampled && thr >= before && (quad == 3 || thr < before + mine);
        if (__any_sync(0xffffffffu, claim)) {
          float cum = before;
          int token = cq + 127;
          bool found = !claim;
          for (int c0 = 0; c0 < 128; c0 += 16) {
            if (__all_sync(0xffffffffu, found)) break;
            float l[16];
            tmem_ld16_nowait(e.tlane + cq + c0, l);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              cum += __expf((l[i] + __ldg(bias + p.b_nin + cq + c0 + i)) * p.inv_temperature - M);
              if (!found && cum > thr) {
                token = cq + c0 + i;
                found = true;
              }
            }
          }
          if (claim) __stcg(p.codes + (size_t)b * LMT_CELLS + cell, (long long)token);
        }
        tc_fence_before();
      }
      e.step_done();  // progress PROG_DONE: the tile's tokens (if any) are in `codes`
    } else {

    for (int oi = 0; oi < TC_NOPS; ++oi) {
      const ps_lmconv_op op = p.ops[oi];
      const int next_form = oi + 1 < TC_NOPS ? (p.ops[oi + 1].kind == 0 ? FORM_PAIR : FORM_RAW) : FORM_NONE;
      const bool out_raw = (p.raw_mask >> op.out) & 1ull;
      if (op.kind == 0) {
        {  // x = PONO(conv_input(concat_elu(og))) [+ nin_skip(concat_elu(a))]
          const uint32_t col0 = (uint32_t)(g & 1) * 160u;
          mbar_wait(&sm.acc_full[g & 1], (uint32_t)(g >> 1) & 1u);
          if (r == 0) TC_TRACE(4, g);
          tc_fence_after();
          float x[LMT_F];
#pragma unroll
          for (int j = 0; j < 5; ++j) tmem_ld16_nowait(e.tlane + col0 + 16 * j, x + 16 * j);
          tmem_ld_wait();
          add_bias80(x, bias + op.b_in);
          part_wait(g);
#pragma unroll
          for (int j = 0; j < 5; ++j) part_add16(x + 16 * j, g, j);
          part_release(g);
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
          if (r == 0) TC_TRACE(6, g);
#pragma unroll
          for (int j = 0; j < 5; ++j) e.emit16(FORM_PAIR, j, x + 16 * j, op.mid, false);
          e.publish();
          e.step_done();
          if (r == 0) TC_TRACE(5, g);
          ++g;
        }
        {  // y = conv_out(concat_elu(x)); og += PONO(y[:80]) * sigmoid(y[80:])
          const uint32_t col0 = (uint32_t)(g & 1) * 160u;
          mbar_wait(&sm.acc_full[g & 1], (uint32_t)(g >> 1) & 1u);
          if (r == 0) TC_TRACE(4, g);
          tc_fence_after();
          float a[LMT_F];
#pragma unroll
          for (int j = 0; j < 5; ++j) tmem_ld16_nowait(e.tlane + col0 + 16 * j, a + 16 * j);
          tmem_ld_wait();
          add_bias80(a, bias + op.b_out);
          part_wait(g);
#pragma unroll
          for (int j = 0; j < 5; ++j) part_add16(a + 16 * j, g, j);
          pono80(a);
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            float gt[16], o[16];
            tmem_ld16_nowait(e.tlane + col0 + 80 + 16 * j, gt);
            tmem_ld16_nowait(e.tlane + COL_OG + 16 * j, o);
            tmem_ld_wait();
            add_bias16(gt, bias + op.b_out + 80 + 16 * j);
            part_add16(gt, g, 5 + j);
            if (j == 4) part_release(g);
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = fmaf(a[16 * j + i], __fdividef(1.0f, 1.0f + __expf(-gt[i])), o[i]);
            tmem_st16_nowait(e.tlane + COL_OG + 16 * j, o);
            e.emit16(next_form, j, o, op.out, out_raw);
          }
          if (next_form != FORM_NONE) e.publish(); else { tmem_st_wait(); tc_fence_before(); }
          e.step_done();
          if (r == 0) TC_TRACE(5, g);
          ++g;
        }
      } else {  // dilated masked conv on the raw stream + PONO becomes the new stream
        const uint32_t col0 = (uint32_t)(g & 1) * 160u;
        mbar_wait(&sm.acc_full[g & 1], (uint32_t)(g >> 1) & 1u);
        if (r == 0) TC_TRACE(4, g);
        tc_fence_after();
        float x[LMT_F];
#pragma unroll
        for (int j = 0; j < 5; ++j) tmem_ld16_nowait(e.tlane + col0 + 16 * j, x + 16 * j);
        tmem_ld_wait();
        add_bias80(x, bias + op.b_in);
        part_wait(g);
#pragma unroll
        for (int j = 0; j < 5; ++j) part_add16(x + 16 * j, g, j);
        part_release(g);
        pono80(x);
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          tmem_st16_nowait(e.tlane + COL_OG + 16 * j, x + 16 * j);
          e.emit16(next_form, j, x + 16 * j, op.out, out_raw);
        }
        if (next_form != FORM_NONE) e.publish(); else { tmem_st_wait(); tc_fence_before(); }
        e.step_done();
        if (r == 0) TC_TRACE(5, g);
        ++g;
      }
    }

    if (need_logits) {
      // ---- logits = nin_out(elu(u)): four 128-class quarters share one operand A = [elu(u) | 0], written into
      // the ring stages of quarter 0's two chunks.  Their last users (body chunks) completed before the last
      // accumulator barrier fired, and no later chunk writes rows into them. ----
      {
        const int first = logit_first;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          float o[16], pp[16], nn;
          tmem_ld16_nowait(e.tlane + COL_OG + 16 * j, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) celu(o[i], pp[i], nn);
          if (r < a_rows) {  // rows past the tile's A tile would land in the stage's weights
            const int kg0 = 2 * j, kg1 = 2 * j + 1;
            sts_a(tiles, slot_off(first + (kg0 >> 3)), kg0 & 7, r,
                  make_uint4(pack_h2(pp[0], pp[1]), pack_h2(pp[2], pp[3]), pack_h2(pp[4], pp[5]), pack_h2(pp[6], pp[7])));
            sts_a(tiles, slot_off(first + (kg1 >> 3)), kg1 & 7, r,
                  make_uint4(pack_h2(pp[8], pp[9]), pack_h2(pp[10], pp[11]), pack_h2(pp[12], pp[13]), pack_h2(pp[14], pp[15])));
            sts_a(tiles, slot_off(first + 1), 2 + j, r, make_uint4(0, 0, 0, 0));
            if (j == 0) sts_a(tiles, slot_off(first + 1), 7, r, make_uint4(0, 0, 0, 0));
          }
        }
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane < 8) {
          mbar_arrive(&sm.full[(first >> 1) % nst]);
          mbar_arrive(&sm.full[((first + 1) >> 1) % nst]);
        }
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
        if (sampled) __stcg(p.codes + (size_t)b * LMT_CELLS + cell, (long long)token);
      }
      tc_fence_before();
    }
    e.step_done();  // progress PROG_DONE: the tile's tokens (if any) are in `codes`
    }  // !split_epi
    }  // !is_halo
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

// activation cache + the launch's tile table and progress words
static size_t tc_act_bytes(int B) { return (size_t)(B > 0 ? B : 0) * LMT_TENSORS * LMT_CELLS * LMT_ACT * sizeof(__half); }
// every row has at most one chain tile and 8 halo tiles of its own, usually far fewer
static size_t tc_max_tiles(int B) { return (size_t)(B > 0 ? B : 0) * LMT_CELLS * 2 + 64; }
// rows of one sampled level whose partial sums the buffer holds (two levels in flight); a wider level runs unsplit
static size_t tc_part_rows(int B) { return align_up((size_t)std::max(256, (B > 0 ? B : 0) * 64), 128); }
static size_t tc_part_bytes(int B) { return 2 * tc_part_rows(B) * LMT_PART_COLS * sizeof(float); }
size_t ps_lmconv_tc_cache_bytes(int B) {
  return align_up(tc_act_bytes(B), 256) + align_up(tc_max_tiles(B) * sizeof(Tile), 256) +
         align_up((tc_max_tiles(B) + 16) * sizeof(unsigned int), 256) + tc_part_bytes(B);
}

// Dependency levels (host).  A cell reads, through its three masks, cells generated earlier; it sits one level above
// the highest of them.  Rows come out in two phases:
//   phase A  the known prefix: cells that neither are sampled nor have a sampled cell among their ancestors.  Their
//            tokens are all known, so consecutive levels can run as a pipeline a layer apart;
//   phase B  sampled cells and everything downstream of one, levelled among themselves (a B cell's A neighbours are
//            complete before phase B starts).  A B level needs the previous level's tokens before its first layer.
// level_offsets lists the A levels then the B levels; *first_b_level is the index of the first B level (== *n_levels
// when there is none).  mode 0: sampling -- cells ranked after the image's last sampled cell are dropped, images with
// nothing to sample produce no rows.  mode 1: teacher-forced logits of every cell (all phase A).
int ps_lmconv_levels_host(const int* order, const uint16_t* words, const uint8_t* sample_mask, int B, int mode,
                          ps_lmconv_row* rows_out, int* level_offsets, int max_levels, int* n_levels, int* first_b_level) {
  PS_CHECK_ARG(order && words && rows_out && level_offsets && n_levels && first_b_level && B >= 0 && max_levels >= 2);
  PS_CHECK_ARG(mode == 1 || sample_mask);
  PS_CHECK_ARG(B < (1 << 20));
  // level >= 0: phase A level; level <= -2: phase B level -(level + 2); -1: no row
  std::vector<int> level((size_t)B * LMT_CELLS, -1), uidx((size_t)B * LMT_CELLS, 0);
  // Images are independent: worker c owns the images [c*B/T, (c+1)*B/T).  Rows of a level are emitted in (image, cell)
  // order whatever T is: per-worker level histograms are prefix-summed in worker order before the rows are written.
  unsigned hw = std::thread::hardware_concurrency();
  if (const char* e = getenv("PS_HOST_THREADS")) hw = (unsigned)std::max(1, atoi(e));  // tests pin the worker count
  const int T = std::max(1, std::min({(int)(hw ? hw : 1), 32, B / 4}));
  std::vector<int> tops_a(T, -1), tops_b(T, -1), bad(T, 0);
  // worker c runs fn(c) on the library's persistent host team (hostpool.cuh)
  auto run_workers = [&](auto&& fn) { HostPool::get().run(T, fn); };
  auto image_range = [&](int c, int& lo, int& hi) {
    lo = (int)((long long)B * c / T);
    hi = (int)((long long)B * (c + 1) / T);
  };
  run_workers([&](int c) {
    int lo, hi;
    image_range(c, lo, hi);
    std::vector<int> rank(LMT_CELLS);
    int top_a = -1, top_b = -1;
    for (int b = lo; b < hi && !bad[c]; ++b) {
      const int* ord = order + (size_t)b * LMT_CELLS;
      const uint16_t* w = words + (size_t)b * 3 * LMT_CELLS;
      int* lv = level.data() + (size_t)b * LMT_CELLS;
      std::fill(rank.begin(), rank.end(), -1);
      for (int i = 0; i < LMT_CELLS; ++i) {
        if (ord[i] < 0 || ord[i] >= LMT_CELLS || rank[ord[i]] != -1) {  // out of range, or a cell listed twice
          bad[c] = 1;
          break;
        }
        rank[ord[i]] = i;
      }
      if (bad[c]) break;
      int last = LMT_CELLS - 1;
      const uint8_t* smk = mode == 0 ? sample_mask + (size_t)b * LMT_CELLS : nullptr;
      if (mode == 0) {
        int drawn = 0;
        last = -1;
        for (int i = 0; i < LMT_CELLS; ++i)
          if (smk[ord[i]]) {
            last = i;
            uidx[(size_t)b * LMT_CELLS + ord[i]] = drawn++;
          }
        if (last < 0) continue;  // nothing to sample: the image needs no work
      }
      for (int i = 0; i <= last && !bad[c]; ++i) {
        const int cell = ord[i];
        const int r = cell / 32, cc0 = cell % 32;
        int la = -1, lb = -1;
        bool in_b = smk && smk[cell];
        for (int m = 0; m < 3; ++m) {
          const int dil = m == 2 ? 2 : 1;
          for (int t = 0; t < 9; ++t) {
            if (t == 4 || !((w[m * LMT_CELLS + cell] >> t) & 1)) continue;
            const int rr = r + (t / 3 - 1) * dil, cc = cc0 + (t % 3 - 1) * dil;
            // masks never reach outside the grid, nor forward in the order
            if (rr < 0 || rr >= 32 || cc < 0 || cc >= 32 || rank[rr * 32 + cc] >= i) {
              bad[c] = 1;
              break;
            }
            const int l = lv[rr * 32 + cc];
            if (l >= 0) {
              la = std::max(la, l);
            } else {
              in_b = true;
              lb = std::max(lb, -(l + 2));
            }
          }
          if (bad[c]) break;
        }
        if (bad[c]) break;
        if (in_b) {
          lv[cell] = -(lb + 1 + 2);
          top_b = std::max(top_b, lb + 1);
        } else {
          lv[cell] = la + 1;
          top_a = std::max(top_a, la + 1);
        }
      }
    }
    tops_a[c] = top_a;
    tops_b[c] = top_b;
  });
  for (int c = 0; c < T; ++c)
    if (bad[c]) return fail(PS_EINVAL, "%s: order is not a permutation or a mask reaches outside the grid / forward in the order%s", __func__);
  const int top_a = *std::max_element(tops_a.begin(), tops_a.end()), top_b = *std::max_element(tops_b.begin(), tops_b.end());
  const int na = top_a + 1, nb = top_b + 1;
  if (na + nb > max_levels) return fail(PS_EWORKSPACE, "%s: more dependency levels than level_offsets holds%s", __func__);
  const int nl = na + nb;
  auto slot = [&](int l) { return l >= 0 ? l : na - (l + 2); };
  std::vector<int> hist((size_t)T * (nl + 1), 0);  // hist[c][l]: rows of level l among worker c's images
  run_workers([&](int c) {
    int lo, hi;
    image_range(c, lo, hi);
    int* h = hist.data() + (size_t)c * (nl + 1);
    for (size_t i = (size_t)lo * LMT_CELLS; i < (size_t)hi * LMT_CELLS; ++i)
      if (level[i] != -1) ++h[slot(level[i])];
  });
  // level_offsets, and hist[c][l] := first row of worker c inside level l
  int run = 0;
  for (int l = 0; l < nl; ++l) {
    level_offsets[l] = run;
    for (int c = 0; c < T; ++c) {
      const int n = hist[(size_t)c * (nl + 1) + l];
      hist[(size_t)c * (nl + 1) + l] = run;
      run += n;
    }
  }
  level_offsets[nl] = run;
  run_workers([&](int c) {
    int lo, hi;
    image_range(c, lo, hi);
    int* cursor = hist.data() + (size_t)c * (nl + 1);
    for (int b = lo; b < hi; ++b) {
      const uint16_t* w = words + (size_t)b * 3 * LMT_CELLS;
      for (int cell = 0; cell < LMT_CELLS; ++cell) {
        const int l = level[(size_t)b * LMT_CELLS + cell];
        if (l == -1) continue;
        ps_lmconv_row ri;
        ri.bc = (b << 10) | cell;
        ri.w01 = (uint32_t)w[cell] | ((uint32_t)w[LMT_CELLS + cell] << 16);
        ri.w2_flags = (uint32_t)w[2 * LMT_CELLS + cell] | ROW_VALID;
        if (mode == 1)
          ri.w2_flags |= ROW_LOGITS;
        else if (sample_mask[(size_t)b * LMT_CELLS + cell])
          ri.w2_flags |= ROW_SAMPLED;
        ri.uidx = uidx[(size_t)b * LMT_CELLS + cell];
        rows_out[cursor[slot(l)]++] = ri;
      }
    }
  });
  *n_levels = na + nb;
  *first_b_level = na;
  return PS_OK;
}

int ps_lmconv_tc_run(const ps_lmconv_plan* plan, int B, const ps_lmconv_row* rows_dev, const int* level_offsets_host,
                     int n_levels, int first_b_level, long long* codes, const float* uniforms, int uniforms_stride,
                     float temperature, float* logits_out, void* cache, size_t cache_bytes, void* stream) {
  if (int rcw = wedge_check(__func__)) return rcw;
  PS_WEDGE_ARM();
  PS_CHECK_ARG(plan && plan->wblob && plan->chunks && plan->w_uinit && plan->bias && codes && cache);
  PS_CHECK_ARG(B >= 0 && n_levels >= 0 && temperature > 0.0f);
  PS_CHECK_ARG(plan->n_chunks_body > 0 && plan->n_chunks_total >= plan->n_chunks_body);
  PS_CHECK_ARG(plan->n_chunks_total <= TC_MAX_CHUNKS);
  PS_CHECK_ARG(plan->n_chunks_body % 2 == 0 && plan->n_chunks_total % 2 == 0);  // two chunks per ring stage
  if (B == 0 || n_levels == 0) return PS_OK;
  PS_CHECK_ARG(rows_dev && level_offsets_host && first_b_level >= 0 && first_b_level <= n_levels);
  PS_CHECK_ARG(uniforms || logits_out);  // sampling needs the uniform numbers
  if (cache_bytes < ps_lmconv_tc_cache_bytes(B)) return fail(PS_EWORKSPACE, "%s: activation cache too small%s", __func__);
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.wblob = (const unsigned char*)plan->wblob;
  p.chunks = plan->chunks;
  p.n_body = plan->n_chunks_body;
  p.n_total = plan->n_chunks_total;
  const bool can_split = plan->chunks_chain && plan->chunks_halo && plan->part_col[32] == LMT_PART_COLS &&
                         plan->n_chain_total <= TC_MAX_CHUNKS && plan->n_chain_body % 2 == 0 && plan->n_chain_total % 2 == 0;
  p.chunks_chain = plan->chunks_chain;
  p.n_chain_body = plan->n_chain_body;
  p.n_chain_total = plan->n_chain_total;
  p.chunks_halo = plan->chunks_halo;
  for (int i = 0; i < 33; ++i) {
    p.halo_first[i] = (short)plan->halo_first[i];
    p.part_col[i] = (short)plan->part_col[i];
  }
  p.logit_first = plan->epi_first[32];
  p.w_uinit = (const __half*)plan->w_uinit;
  p.bias = plan->bias;
  p.b_uinit = plan->b_uinit;
  p.b_nin = plan->b_nin;
  p.raw_mask = plan->raw_mask;
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

  int dev = 0;
  PS_CUDA(cudaGetDevice(&dev));
  const int sms = stream_sms((cudaStream_t)stream);
  // GEMMs per halo tile: 4 (a 128-row slab of a sampled level spreads over 8 SMs).  Measured on 16- and 24-SM partitions
  // with 8 and 16: no gain (there the known prefix bounds the launch, not the residency of a level's tiles).
  int halo_group = HALO_GROUP;
  if (const char* e = getenv("PS_TC_HALO_GROUP")) halo_group = std::max(2, std::min(32, atoi(e))) & ~1;  // developer aid
  p.halo_group = halo_group;
  // tiles in level order: a level's rows are split evenly over its tiles
  std::vector<Tile> tiles;
  const int exp_bits = getenv("PS_TC_EXP") ? atoi(getenv("PS_TC_EXP")) : 0;  // developer aid: scheduling experiments
  int prev_first = 0, prev_count = 0;
  if (exp_bits & 8) n_levels = first_b_level;  // timing aid: the known prefix alone
  for (int l = 0; l < n_levels; ++l) {
    const int r0 = level_offsets_host[l], r1 = level_offsets_host[l + 1];
    PS_CHECK_ARG(r1 >= r0);
    if (r1 == r0) continue;
    // A sampled level (sampling mode) is split: halo tiles (<= 128 rows x a group of GEMMs: the gathered neighbour
    // taps, as per-row partial sums) first, then chain tiles (<= 32 rows: what depends on the rows' own column).  Chain
    // tiles wait for the previous level's CHAIN tiles (its tokens) and for their own halo tiles; halo tiles follow the
    // previous level's chain one step behind.  The partial sums of two consecutive levels alternate between the two
    // halves of the buffer: level l + 1's halo tiles start when level l's chain runs, which is after level l - 1's.
    const int rows_l = r1 - r0;
    const bool split = l >= first_b_level && can_split && !(exp_bits & 16) && uniforms && !logits_out &&
                       (size_t)rows_l <= tc_part_rows(B);
    if (split) {
      const int part_base = (int)(((l - first_b_level) & 1) * tc_part_rows(B));
      const int wait = l == first_b_level ? LMT_TENSORS : ((exp_bits & 32) ? PROG_DONE : 0);  // 32: halo tiles do not run ahead
      const int n_hrow = (rows_l + 127) / 128, n_grp = (32 + halo_group - 1) / halo_group;
      const int h_first = (int)tiles.size();
      for (int hr = 0; hr < n_hrow; ++hr)
        for (int gr = 0; gr < n_grp; ++gr) {
          Tile tl;
          memset(&tl, 0, sizeof(tl));
          tl.row_begin = r0 + hr * 128;
          tl.nrows = std::min(128, r1 - tl.row_begin);
          tl.prev_first = prev_first;
          tl.prev_count = prev_count;
          tl.wait_start = wait;
          tl.kind_g = TILE_HALO | (gr * halo_group) << 8 | std::min(32, (gr + 1) * halo_group) << 16;
          tl.part_row0 = part_base + hr * 128;
          tiles.push_back(tl);
        }
      const int c_first = (int)tiles.size();
      for (int hr = 0; hr < n_hrow; ++hr) {
        const int hb = r0 + hr * 128, he = std::min(r1, hb + 128);
        for (int cb = hb; cb < he; cb += CHAIN_ROWS) {
          Tile tl;
          memset(&tl, 0, sizeof(tl));
          tl.row_begin = cb;
          tl.nrows = std::min(CHAIN_ROWS, he - cb);
          tl.prev_first = prev_first;
          tl.prev_count = prev_count;
          tl.wait_start = l == first_b_level ? LMT_TENSORS : PROG_DONE;
          tl.kind_g = TILE_CHAIN;
          tl.part_row0 = part_base + (cb - r0);
          tl.h_first = h_first + hr * n_grp;
          tiles.push_back(tl);
        }
      }
      prev_first = c_first;
      prev_count = (int)tiles.size() - c_first;
      continue;
    }
    // The prefix levels overlap each other, so their tiles are full (fewest SMs per level).  A sampled level runs
    // alone and its time is one tile's latency: small tiles get a deeper operand ring and there are SMs to spare.
    int rpt = 128;
    if (l >= first_b_level) {
      rpt = 32;
      while (rpt < 128 && (r1 - r0 + rpt - 1) / rpt > sms) rpt *= 2;
      if (exp_bits & 2) rpt = 128;
    } else if (exp_bits & 4) {
      rpt = 32;
    }
    const int nt = (r1 - r0 + rpt - 1) / rpt;
    const int first = (int)tiles.size();
    for (int t = 0; t < nt; ++t) {
      Tile tl;
      memset(&tl, 0, sizeof(tl));
      tl.row_begin = r0 + (int)((long long)(r1 - r0) * t / nt);
      tl.nrows = r0 + (int)((long long)(r1 - r0) * (t + 1) / nt) - tl.row_begin;
      tl.prev_first = prev_first;
      tl.prev_count = prev_count;
      // first B level: the whole prefix must be cached; later B levels: the previous level's tokens must be drawn
      tl.wait_start = l < first_b_level ? 0 : (l == first_b_level ? LMT_TENSORS : PROG_DONE);
      if (exp_bits & 1) tl.wait_start = PROG_DONE;  // no pipelining between prefix levels
      tiles.push_back(tl);
    }
    prev_first = first;
    prev_count = nt;
  }
  const int n_tiles = (int)tiles.size();
  if (n_tiles == 0) return PS_OK;
  PS_CHECK_ARG((size_t)n_tiles <= tc_max_tiles(B));
  char* tail = (char*)cache + align_up(tc_act_bytes(B), 256);
  Tile* tiles_dev = (Tile*)tail;
  unsigned int* sync_dev = (unsigned int*)(tail + align_up(tc_max_tiles(B) * sizeof(Tile), 256));
  p.part = (float*)((char*)sync_dev + align_up((tc_max_tiles(B) + 16) * sizeof(unsigned int), 256));
  // The tile table is staged in a pinned buffer that outlives the call (one per host thread, grown on demand): the
  // copy is asynchronous and nothing here waits for the stream.  A second call from the same thread reuses the buffer
  // only after the previous call's copy has left it (event).
  {
    static thread_local Tile* pin = nullptr;
    static thread_local size_t pin_cap = 0;
    static thread_local cudaEvent_t pin_evt = nullptr;
    if (pin_evt) PS_CUDA(cudaEventSynchronize(pin_evt));
    if (pin_cap < (size_t)n_tiles) {
      if (pin) cudaFreeHost(pin);
      pin = nullptr;
      pin_cap = 0;
      const size_t cap = std::max<size_t>(4096, (size_t)n_tiles * 2);
      PS_CUDA(cudaHostAlloc((void**)&pin, cap * sizeof(Tile), cudaHostAllocDefault));
      pin_cap = cap;
    }
    if (!pin_evt) PS_CUDA(cudaEventCreateWithFlags(&pin_evt, cudaEventDisableTiming));
    memcpy(pin, tiles.data(), (size_t)n_tiles * sizeof(Tile));
    PS_CUDA(cudaMemcpyAsync(tiles_dev, pin, (size_t)n_tiles * sizeof(Tile), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    PS_CUDA(cudaEventRecord(pin_evt, (cudaStream_t)stream));
  }
  PS_CUDA(cudaMemsetAsync(sync_dev, 0, ((size_t)n_tiles + 16) * sizeof(unsigned int), (cudaStream_t)stream));
  p.tiles = tiles_dev;
  p.n_tiles = n_tiles;
  p.sync = sync_dev;

  const size_t smem_bytes = 1024 + (size_t)TC_STAGES * TC_STAGE_BYTES + sizeof(TcSmem);
  static_assert(1024 + (size_t)TC_STAGES * TC_STAGE_BYTES + sizeof(TcSmem) <= 227 * 1024, "shared memory budget");
  static thread_local int attr_dev = -1;
  if (attr_dev != dev) {
    PS_CUDA(cudaFuncSetAttribute(lmconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    attr_dev = dev;
  }
  PS_TIME_BEGIN("lmconv_tc_kernel", (cudaStream_t)stream);
  lmconv_tc_kernel<<<n_tiles, TC_THREADS, smem_bytes, (cudaStream_t)stream>>>(p);
  PS_LAUNCHED();
  PS_TIME_END((cudaStream_t)stream);
  return PS_OK;
}

}  // extern "C"
