/*
 * pixelsynth_b200.h -- C ABI of libpixelsynth_b200.so (sm_100a CUDA kernels for the PixelSynth
 * novel-view-synthesis inference hot path).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; tensors are dense, row-major,
 *     in the layouts the reference's PyTorch code uses (NCHW float32 unless stated);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is enqueued
 *     on it, nothing synchronises the device and nothing allocates: the caller owns every buffer,
 *     including the scratch `workspace` whose size the matching *_workspace_bytes() call returns;
 *   - return value: PS_OK (0) or a negative PS_E* code; ps_error_string() names it and
 *     ps_last_error_detail() gives the thread-local detail (e.g. the CUDA error string);
 *   - inputs are never written.  Re-entrant across host threads / devices: no mutable global state.
 *
 * Each entry point cites the reference interface (crockwell/pixelsynth @ cfe18078) it replaces.
 */
#ifndef PIXELSYNTH_B200_H_
#define PIXELSYNTH_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PS_OK 0
#define PS_EINVAL (-1)       /* bad argument (shape, null pointer, unsupported value) */
#define PS_ECUDA (-2)        /* a CUDA runtime call or kernel launch failed */
#define PS_EWORKSPACE (-3)   /* workspace too small */
#define PS_EUNSUPPORTED (-4) /* valid in the reference but not built here (e.g. K > PS_MAX_POINTS_PER_PIXEL) */

#define PS_ABI_VERSION 1
/* PyTorch3D's kMaxPointsPerPixel is 150; the reference uses pp_pixel = 128 (options/train_options.py:104). */
#define PS_MAX_POINTS_PER_PIXEL 128

#define PS_ACCUM_ALPHACOMPOSITE 0
#define PS_ACCUM_WSUM 1
#define PS_ACCUM_WSUMNORM 2

int ps_abi_version(void);
const char* ps_error_string(int code);
const char* ps_last_error_detail(void);

/* ------------------------------------------------------------------------------------------------
 * Splat stage 1-2: unproject the source pixel grid with depth, camera-transform, perspective divide.
 * Replaces PtsManipulator.project_pts (models/projection/z_buffer_manipulator.py:50-83) including the
 * `xyzs` grid buffer of PtsManipulator.__init__ (:38-48).
 *   depth  (B, W*W)   source depth, pixel order p = y*W + x            [pred_pts.view(bs,1,-1)]
 *   mats   (B, 6, 16) row-major 4x4 [K, Kinv, RT1, RT1inv, RT2, RT2inv] (forward_justpts order, :85-87)
 *   pts    (B, W*W, 3) out: `sampler` permuted point-major (as at :103): x right, y down, z>0 in front;
 *                      points with |z_proj| < eps are parked at (-10, 10, 10)
 *   xyproj (B, 4, W*W) out or NULL: pre-division homogeneous coords, the cloud that
 *                      project_pts_cumulative returns (:266)
 * ------------------------------------------------------------------------------------------------ */
int ps_project_pts(const float* depth, const float* mats, int B, int W, float eps, float* pts, float* xyproj,
                   void* stream);

/* Prior-cloud branch of PtsManipulator.project_pts_cumulative (z_buffer_manipulator.py:244-248,253-264).
 *   cloud (B,4,P) homogeneous points in the previous target camera frame; mats3 (B,3,16) = [K, RT2, RT3inv]. */
int ps_project_cloud(const float* cloud, const float* mats3, int B, int P, float eps, float* pts, float* xyproj,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * Splat stage 3-5: K-nearest z-buffer rasterisation + alpha + accumulate + background mask.
 * Replaces RasterizePointsXYsBlending.forward (models/layers/z_buffer_layers.py:55-131), i.e.
 * pytorch3d rasterize_points (:81-84), the alpha formula (:89-98), the 13x13 background dilation
 * (:100-110) and pytorch3d compositing.{alpha_composite,weighted_sum,weighted_sum_norm} (:112-129).
 *   pts    (B,P,3) point cloud in project_pts' frame (NOT yet negated; the kernel applies :71-72)
 *   feat   (B,C,P) per-point features (`src`)
 *   S      output image side; K = points_per_pixel (<= PS_MAX_POINTS_PER_PIXEL)
 *   radius_px = opt.radius (pixels); the NDC radius is radius_px / S * 2 as at :77
 *   out     (B,C,S,S) f32                       [transformed_src_alphas]
 *   bg_mask (B,S,S)   u8 0/1                    [background_mask]
 *   idx     (B,S,S,K) i32 packed b*P+p or -1    [rasterize_points()[0]]   may be NULL
 *   zbuf    (B,S,S,K) f32 or -1                 [rasterize_points()[1]]   may be NULL
 *   dist2   (B,S,S,K) f32 or -1                 [rasterize_points()[2]]   may be NULL
 * Selection is ascending (z, packed index); membership is z >= 0 and dx*dx + dy*dy < r*r evaluated in
 * fp32 without FMA contraction; no point is ever dropped for capacity reasons.
 * ------------------------------------------------------------------------------------------------ */
size_t ps_splat_workspace_bytes(int B, int P, int S, double radius_px);
int ps_splat_points(const float* pts, const float* feat, int B, int P, int C, int S, int K, double radius_px,
                    double tau, int rad_pow, int accumulation, int bg_ksize, float* out, uint8_t* bg_mask,
                    int32_t* idx, float* zbuf, float* dist2, void* workspace, size_t workspace_bytes, void* stream);

/* forward_justpts (z_buffer_manipulator.py:85-107) = ps_project_pts + ps_splat_points with P = W*W.
 * The projected cloud lives in the workspace (ps_splat_fwd_workspace_bytes). */
size_t ps_splat_fwd_workspace_bytes(int B, int W, int S, double radius_px);
int ps_splat_fwd(const float* depth, const float* feat, const float* mats, int B, int W, int C, int S, int K,
                 double radius_px, double tau, int rad_pow, int accumulation, int bg_ksize, float eps, float* out,
                 uint8_t* bg_mask, int32_t* idx, float* zbuf, float* dist2, void* workspace, size_t workspace_bytes,
                 void* stream);

/* Number of kernels this library has launched from the calling thread since the last reset
 * (bench.py's `gpu_launches`). */
long long ps_launch_count(void);
void ps_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------------
 * Dense convolution as a TMA + tcgen05 implicit GEMM (bf16 operands, fp32 accumulation in TMEM).
 * Replaces the cuDNN convolutions behind nn.Conv2d / nn.ConvTranspose2d of the reference's inference
 * networks: Unet (models/networks/architectures.py:191-209,230-279), ResNet_Block of the refinement
 * decoder (models/layers/blocks.py:41-74), VQ-VAE-2 Encoder / Decoder (models/vqvae2/vqvae.py:80-161).
 * Activations are NHWC bf16 with the channel count a multiple of 8.  The convolution is described as a list
 * of taps per input tensor: output pixel (oy, ox) reads input pixel (oy*stride + dy[t], ox*stride + dx[t])
 * (out-of-image reads are zero) against rows wrow[t] .. wrow[t]+cout_pad-1 of the packed weight matrix
 * [w_rows][w_cin_pad] (bf16, input channels contiguous, zero padded to a multiple of 64).  A second input with
 * its own taps accumulates into the same output (the decoder's fused 1x1 skip convolution).
 * Epilogue: v = acc + bias[c] (+ residual); output o = act_o(v * scale_o[c] + shift_o[c]) for up to two NHWC
 * bf16 outputs (channel stride / offset allow writing into a concatenation buffer) and an optional fp32 NCHW
 * copy of output 0.  Output pixel (oy, ox) is stored at (oy*out_sy + out_py, ox*out_sx + out_px) of an
 * out_H x out_W image (the four phases of a stride-2 transposed convolution).
 * ------------------------------------------------------------------------------------------------ */
#define PS_ACT_NONE 0
#define PS_ACT_RELU 1
#define PS_ACT_LEAKY02 2
#define PS_ACT_TANH 3
#define PS_ACT_SIGMOID_AFFINE 4 /* sigmoid(v) * act_param[0] + act_param[1] */
#define PS_ACT_ELU 5

typedef struct {
  const void* ptr; /* NHWC bf16 */
  int H, W, C, cstride;
  int ntaps;
  int dy[16], dx[16];
  int wrow[16];
} ps_conv_input;

typedef struct {
  void* ptr; /* NHWC bf16 or NULL */
  const float* scale;
  const float* shift;
  int per_sample; /* scale/shift are [N][Cout] instead of [Cout] */
  int act;
  int cstride, coffset;
} ps_conv_output;

typedef struct {
  ps_conv_input in[2];
  const void* weights;
  int w_rows, w_cin_pad;
  int N, Hout, Wout, Cout, cout_pad, stride;
  const float* bias;
  const void* residual; /* NHWC bf16 on the full output grid */
  int res_cstride;
  ps_conv_output out[2];
  float* out_f32_nchw;
  float act_param[2];
  int out_H, out_W, out_sy, out_sx, out_py, out_px; /* 0 = same as Hout/Wout, stride 1, phase 0 */
} ps_conv_desc;

int ps_conv_igemm(const ps_conv_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Memory-bound glue between the convolutions (csrc/elementwise.cu).
 * ------------------------------------------------------------------------------------------------ */
/* (N,C,H,W) f32 -> (N,H,W,cstride) bf16, zero padded; if mask (N,H,W u8) is given channel C = float(~mask):
 * the decoder input torch.cat((x, (~background_mask).unsqueeze(1).float()), 1), architectures.py:154. */
int ps_nchw_to_nhwc_bf16(const float* x, int N, int C, int H, int W, const uint8_t* mask, void* out, int cstride,
                         void* stream);
/* NHWC bf16 resampling with up to two outputs y = act(v*scale+shift): mode 0 identity, 1 nn.AvgPool2d(3,2,1)
 * (blocks.py:46), 2 nn.Upsample(scale_factor=2, mode="bilinear") (blocks.py:48, architectures.py:201),
 * 3 F.avg_pool2d(3, 2, 1, count_include_pad=False) (MultiscaleDiscriminator.downsample, discriminators.py:170-177),
 * 4 nn.MaxPool2d(3, 2, 1) (torchvision resnet18, the places365 classifier of z_buffermodel.py:88). */
int ps_resample(const void* in, int N, int H, int W, int C, int in_cstride, int mode, const ps_conv_output* out0,
                const ps_conv_output* out1, void* stream);
/* nn.InstanceNorm2d(affine=False) of the discriminator (normalization.py:78-79) as a per-sample affine: statistics of
 * x (N,HW,cstride) NHWC bf16 per (sample, channel), biased variance -> scale = rsqrt(var+eps), shift = -mean*scale, (N,C). */
int ps_instance_norm_stats(const void* x, int N, int HW, int C, int cstride, float eps, float* scale, float* shift,
                           void* stream);
/* The places365 classifier's input exactly as get_best_sample builds it (z_buffermodel.py:105-110,256-257): image 0 of
 * each of M candidates (f32, img_stride floats apart; (3,256,256) reshaped -- not permuted -- to (256,256,3)), uint8
 * truncation, PIL antialiased bilinear 256->224 bit for bit (tap0[224] first tap, kk[224][4] 22-bit fixed-point weights,
 * built by the host as Pillow's precompute_coeffs does; horizontal then vertical pass, each rounded to uint8), /255,
 * ImageNet normalisation -> (M,224,224,8) NHWC bf16. */
int ps_classifier_input(const float* img, long long img_stride, int M, const int* tap0, const int* kk, void* out,
                        void* stream);
/* LinearNoiseLayer + bn in eval mode (normalization.py:39-47,146-171): per-sample scale/shift (N,cpad) such that
 * bn(x, gain, bias) = x*scale + shift, from noise z (N,Z) and the spectrally normalised (C,Z) gain/bias matrices. */
int ps_noise_affine(const float* z, int N, int Z, const float* Wg, const float* Wb, const float* mean, const float* var,
                    float eps, int C, int cpad, float* scale, float* shift, void* stream);
/* Quantize.forward's code search (vqvae.py:41-48): x (N,D,HW) f32, embed (D,J) -> ids (N,HW) int64. */
int ps_vq_argmin(const float* x, int N, int D, int HW, const float* embed, int J, long long* ids, void* stream);
/* Quantize.embed_code (vqvae.py:76-77) into NHWC bf16 (total, D). */
int ps_embed_codes(const long long* ids, int total, int D, const float* embed, int J, void* out, void* stream);
/* ZbufferModelPts.get_combined (z_buffermodel.py:703-708): a*(1-bg) + b*bg, NCHW f32, bg (N,HW) u8. */
int ps_combine(const float* a, const float* b, const uint8_t* bg, int N, int C, int HW, float* out, void* stream);
/* ResNetDecoder head (architectures.py:157-160): tanh(v + x), or tanh(v) + x when normalize_before_residual. */
int ps_tanh_residual(const float* v, const float* x, long long n, int normalize_before_residual, float* out,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * Generation order and locally-masked-convolution masks (native host code, csrc/glue.cu).  Replaces
 * ZbufferModelPts.get_masks_for_batch (models/z_buffermodel.py:641-701): AvgPool8 + uint8 truncation,
 * cv2.distanceTransform(DIST_L2, 5) x2, int(fd - bd), get_custom_order.pyx custom_idx and
 * masking.get_masks.  ALL POINTERS ARE HOST POINTERS.
 *   bg_mask_host (B,S,S) u8, S = 256           the splat's background mask
 *   dist_host    (B,32,32) i32 or NULL          `distances`
 *   order_host   (B,1024) i32                   generation order as cell index r*32+c
 *   words_host   (B,3,1024) u16                 nine-bit tap masks per cell: [A dil 1, B dil 1, B dil 2],
 *                                               bit t = (dr+1)*3+(dc+1) (F.unfold's tap order)
 *   sample_mask_host (B,32,32) u8               cells to sample: all 64 pixels background (sample.py:29)
 * ------------------------------------------------------------------------------------------------ */
int ps_lmconv_glue_host(const uint8_t* bg_mask_host, int B, int S, int* dist_host, int* order_host,
                        uint16_t* words_host, uint8_t* sample_mask_host);

/* ------------------------------------------------------------------------------------------------
 * Wavefront tensor-core sampler of the locally-masked-convolution PixelCNN (csrc/lmconv_tc.cu).
 * Replaces models/lmconv/sample.py:8-73 (sample) + models/lmconv/model.py:110-155 (OurPixelCNN.forward, called
 * once per token there).  The masks make the network causal in generation order, so every cell's activation
 * column is computed once; cells that are not masked-in neighbours of each other are independent, so the
 * dependency DAG is levelled (ps_lmconv_levels_host) and ONE launch runs all levels: tiles of up to 128 (image,
 * cell) rows are handed out in level order and synchronise through progress words in global memory.
 *
 * ps_lmconv_plan: the network as the kernel consumes it (built by pixelsynth_b200/lmconv.py from the reference's
 * state dict):
 *   wblob   fp16 weight tiles, one per K chunk: [w_rows output channels][64 k] K-major in the 128-byte-swizzle
 *           shared-memory image tcgen05 reads (chunk 16B-group j of row r stored at group j ^ (r & 7))
 *   chunks  the static K-chunk schedule.  a_kind 0: rows gathered from cached tensor a_tensor through mask
 *           `mask` (0 = A dil 1, 1 = B dil 1, 2 = B dil 2), K = (non-centre tap slot, channel), chunk kc of it,
 *           channels start at ch_off8*8 of the 240-wide cache row; a_kind 1: the row's own cell of a_tensor
 *           (nin_skip); a_kind 3: the centre tap, written into tensor memory by the epilogue of the previous
 *           GEMM; a_kind 2: nin_out quarter 0, operand written into the chunk's ring stage by the last epilogue;
 *           a_kind 4: nin_out quarters 1-3, which re-read quarter 0's operand.
 *           d_col = TMEM column of the accumulator, flags bit0 = accumulate, bit1 = last chunk of the GEMM,
 *           bits 2-3 = which accumulator barrier that completes (0/1 ping-pong, 2 = logits), bit4 = first centre
 *           chunk of its GEMM (the issuer waits for the epilogue's operand there); gemm = index of the chunk's
 *           GEMM in execution order
 *   epi_first[g]  index of the first centre chunk of GEMM g (GEMMs in execution order; nin_out = 4 quarters)
 *   ops     the 18 column operations after u_init (14 gated resnets, kind 0; 4 dilated convs + PONO, kind 1) with
 *           the ids (0..32) of the cached tensors they write (mid, out) and their bias offsets
 *   raw_mask  bit t set: cached tensor t is also read un-activated (by a dilated convolution)
 * ps_lmconv_row: bc = image << 10 | cell; w01 = mask word A | mask word B << 16; w2_flags = mask word B-dil-2 |
 *           bit16 sampled | bit17 logits wanted | bit18 valid; uidx = index of the row's uniform number
 * ------------------------------------------------------------------------------------------------ */
#define PS_LMCONV_MAX_GEMMS 40
#define PS_LMCONV_STAGES 6

typedef struct {
  int kind;
  int og, a, mid, out;
  int w_in, b_in, w_skip, b_skip, w_out, b_out;
} ps_lmconv_op;

typedef struct {
  uint32_t w_off16;
  uint16_t w_rows;
  uint8_t a_kind, a_tensor, mask, cin8, kc, ch_off8;
  uint16_t d_col;
  uint8_t flags, gemm;
} ps_lmconv_chunk;

typedef struct {
  int32_t bc;
  uint32_t w01;
  uint32_t w2_flags;
  int32_t uidx;
} ps_lmconv_row;

typedef struct {
  const void* wblob;             /* device */
  const ps_lmconv_chunk* chunks; /* device */
  int n_chunks_body, n_chunks_total; /* without / with the nin_out quarters */
  int epi_first[PS_LMCONV_MAX_GEMMS];
  const void* w_uinit; /* device fp16 [9 taps][513][80] */
  const float* bias;   /* device */
  int b_uinit, b_nin;
  ps_lmconv_op ops[18];
  unsigned long long raw_mask;
  /* The same schedule split in two for the sampled levels.  chain: the chunks whose operand is the row's own column
   * (nin_skip, centre tap, nin_out) -- what a sampled cell's dependent chain really needs; the centre tap's first
   * chunk does not accumulate.  halo: the gathered neighbour taps of GEMM g are chunks [halo_first[g], halo_first[g+1])
   * of chunks_halo, their last chunk completes the accumulator; other CTAs turn them into per-row partial sums,
   * part_col[g] = first of GEMM g's w_rows fp32 columns in a row of the partial-sum buffer (part_col[32] = row length). */
  const ps_lmconv_chunk* chunks_chain; /* device */
  int n_chain_body, n_chain_total;
  const ps_lmconv_chunk* chunks_halo;  /* device */
  int halo_first[33];
  int part_col[33];
} ps_lmconv_plan;

/* device scratch ps_lmconv_tc_run needs for B images: activation cache + tile table + progress words */
size_t ps_lmconv_tc_cache_bytes(int B);

/* Host: rows of every dependency level, level by level (rows_out holds up to B*1024 rows, level_offsets
 * max_levels + 1 ints).  Levels come in two phases: first the known prefix (cells with no sampled cell among their
 * ancestors), then from *first_b_level on the sampled cells and everything downstream of them, levelled among
 * themselves.  mode 0 = sampling (order (B,1024) i32, words (B,3,1024) u16, sample_mask (B,1024) u8 as produced by
 * ps_lmconv_glue_host; cells after an image's last sampled cell are skipped, images with nothing to sample
 * produce no rows); mode 1 = teacher-forced logits of all cells (sample_mask may be NULL; every level is prefix). */
int ps_lmconv_levels_host(const int* order, const uint16_t* words, const uint8_t* sample_mask, int B, int mode,
                          ps_lmconv_row* rows_out, int* level_offsets, int max_levels, int* n_levels,
                          int* first_b_level);

/* Device: runs all levels in one launch on `stream`; nothing synchronises (the tile table goes through a pinned
 * staging buffer of the calling thread; the call waits only for the PREVIOUS call's copy out of that buffer).  Levels at
 * or after first_b_level are split into parallel halo tiles and short chain tiles; the grid is
 * one CTA per tile in dependency order.  Fails with PS_ECUDA if an earlier tensor-core launch wedged (ps_wedge_poll).
 * codes (B,1024) i64: in = known codes, out = sampled cells
 * filled; uniforms (B,stride) f32: the k-th sampled cell (in generation order) of image b takes the first class whose
 * cumulative softmax(logits/temperature) exceeds uniforms[b][k]; logits_out (B,1024,512) f32 or NULL receives the
 * logits of rows flagged bit17; cache: ps_lmconv_tc_cache_bytes(B) bytes of device scratch. */
int ps_lmconv_tc_run(const ps_lmconv_plan* plan, int B, const ps_lmconv_row* rows_dev, const int* level_offsets_host,
                     int n_levels, int first_b_level, long long* codes, const float* uniforms, int uniforms_stride,
                     float temperature, float* logits_out, void* cache, size_t cache_bytes, void* stream);

/* Developer aid: the last tile of every lmconv launch writes clock64 timestamps of its pipeline events into this device
 * buffer of 8 x 1024 int64 (NULL switches it off).  Not part of the product path. */
void ps_lmconv_tc_set_trace(void* dev_buffer);

/* Watchdog of the tensor-core kernels.  Their barrier waits give up after ~1 s (a wedged producer/consumer protocol is
 * a bug, never a legitimate wait), record where in pinned host memory and let the kernel run to completion on garbage
 * instead of hanging the device.  The record is sticky: from then on ps_conv_igemm and ps_lmconv_tc_run fail with
 * PS_ECUDA before launching anything, and ps_wedge_poll returns 1 (info8: [1] block, [2] thread, [3] shared address
 * of the mbarrier or 0xffffffff for a progress-word wait, [4] parity / needed progress), without any device
 * synchronisation.  Callers that hand results to someone else poll after their own synchronisation. */
int ps_wedge_poll(unsigned int* info8_or_null);
int ps_wedge_log(unsigned int* out, int max_words); /* developer aid: raw watchdog words incl. the waiter snapshot */
void ps_wedge_reset(void);

/* SM partitions for pipelined serving (no counterpart in the reference, which runs one batch at a time: this is the
 * scheduling a B200 deployment adds around models/z_buffermodel.py:291-419).  The sampler launch is a latency-bound
 * chain that keeps few SMs busy, so a serving loop keeps two batches in flight: batch k+1's
 * sampler runs on a small partition while batch k's refinement decoder fills the rest.  ps_sm_partition_create splits
 * the device's SMs into two disjoint CUDA green contexts (>= small_sms SMs, rounded up to the hardware granularity,
 * and the remainder) and creates n_small_streams / n_big_streams non-blocking streams on them (cudaStream_t, written
 * to the caller's arrays).  Kernels of this library launched on such a stream size their persistent grids to the
 * partition (ps_stream_sm_count).  PS_EUNSUPPORTED when the driver has no green contexts.  The handle owns the
 * contexts and streams; ps_sm_partition_destroy synchronises and releases them.  A handle is not thread-safe: create,
 * add streams and destroy from one thread (launching on the streams from any thread is fine). */
int ps_sm_partition_create(int device, int small_sms, int n_small_streams, void** small_streams, int n_big_streams,
                           void** big_streams, int* small_count, int* big_count, void** handle);
/* One more stream on the small (big = 0) or large (big = 1) partition of `handle`; high_priority != 0 gives it the
 * device's greatest stream priority, so its kernels' CTAs are placed before those of the partition's other streams
 * whenever SMs free up (the pipeline runs each batch's short front end -- depth net, splat, VQ encoder -- this way, so
 * the host can build the batch's generation order while the previous batch's decoder still has the partition). */
int ps_sm_partition_stream(void* handle, int big, int high_priority, void** stream);
int ps_sm_partition_destroy(void* handle);
int ps_stream_sm_count(void* stream);

/* Per-kernel device timing for bench.py's roofline: while enabled, selected kernels  are bracketed by CUDA events on the launching stream.  ps_timing_collect(name, ...)
 * synchronises those events and returns the summed duration and launch count for `name`;
 * ps_timing_collect(NULL, ...) returns the sum over all names and releases the events. */
void ps_timing_enable(int on);
int ps_timing_collect(const char* kernel, double* total_ms, int* launches);

#ifdef __cplusplus
}
#endif
#endif /* PIXELSYNTH_B200_H_ */
