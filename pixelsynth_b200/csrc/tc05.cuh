// sm_100a PTX wrappers shared by the tensor-core kernels (conv.cu, lmconv_tc.cu): mbarriers, TMA / bulk copies,
// tcgen05 (UMMA) descriptors, MMA issue / commit, TMEM allocation and TMEM <-> register moves.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace ps {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Blocks until the phase with the given parity has completed.  A wait longer than ~1 s means the kernel's barrier
// protocol is wedged (a bug).  The first such wait records who / where in g_wedge AND in the library's pinned, mapped
// host words (g_wedge_host, armed by ps::wedge_arm before every launch of a kernel that uses these waits), and every
// later wait falls through so the kernel terminates instead of hanging the device.  The host words are sticky: every
// C-ABI entry point checks them first and fails with PS_ECUDA from then on (ps_wedge_poll / ps_wedge_reset), so a
// wedged launch can never pass for a result.
static __device__ unsigned int g_wedge[8];
static __device__ unsigned int* g_wedge_host;

static __device__ __noinline__ void wedge_report(unsigned int where, unsigned int what) {
  if (atomicCAS(&g_wedge[0], 0u, 1u) == 0u) {
    g_wedge[1] = blockIdx.x;
    g_wedge[2] = threadIdx.x;
    g_wedge[3] = where;
    g_wedge[4] = what;
    volatile unsigned int* h = g_wedge_host;
    if (h) {
      h[1] = blockIdx.x;
      h[2] = threadIdx.x;
      h[3] = where;
      h[4] = what;
      __threadfence_system();
      h[0] = 1u;
    }
    __threadfence_system();
  }
}

__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  // no suspend-time hint: with one, ptxas emits NANOSLEEP.SYNCS <hint> after a failed check and the wake-up costs
  // hundreds of ns per hand-off, which paced every producer/consumer ring at ~0.3 us per stage
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done;
}

// developer aid: blockIdx -> what the block works on (lmconv: its tile ticket), for the waiter snapshot
static __device__ unsigned int g_blk_tag[16384];

static __device__ __noinline__ void mbar_snapshot(uint32_t bar, uint32_t parity) {
  volatile unsigned int* h = g_wedge_host;
  if (!h) return;
  if ((g_blk_tag[blockIdx.x & 16383] >> 24) != 1u) return;  // chain tiles only
  const unsigned int slot = atomicAdd(&g_wedge[6], 1u);
  if (slot >= 64u) return;
  volatile unsigned int* e = h + 8 + 256 * 8 + slot * 4;
  e[0] = blockIdx.x;
  e[1] = threadIdx.x;
  e[2] = bar;
  e[3] = parity | ((g_blk_tag[blockIdx.x & 16383] & 0xffffffu) << 4) | 0x80000000u;
  __threadfence_system();
}

static __device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  for (int spins = 0;; ++spins) {
    if (mbar_try_wait(bar, parity)) return;
    if ((spins & 63) == 63) {
      if (*(volatile unsigned int*)&g_wedge[0]) {
        if ((threadIdx.x & 31) == 0) mbar_snapshot(smem_u32(bar), parity);
        return;
      }
      if (clock64() - t0 > 2000000000ll) {
        wedge_report(smem_u32(bar), parity);
        return;
      }
    }
  }
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (int spins = 0; spins < 4096; ++spins)
    if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity);  // ~ms without progress: start the wedge watchdog
}

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}
// descriptor-less 1-D bulk copy global -> shared (bytes and both addresses multiples of 16)
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"((uint64_t)src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// whole warp, warp-uniform operands: one elected lane arms the barrier with the byte count and issues the copy
__device__ __forceinline__ void bulk_load_elect(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n\t"
      "@pe cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t"
      "}" ::"r"(smem_u32(dst)),
      "l"((uint64_t)src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// whole warp converged: true in exactly one lane
__device__ __forceinline__ bool elect_one() {
  uint32_t r;
  asm volatile(
      "{\n\t"
      ".reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, pe;\n\t"
      "}"
      : "=r"(r));
  return r != 0;
}
// 16-byte Ampere-style async copy with zero fill (src_bytes = 0 or 16)
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"((uint64_t)src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one (pre-counted) arrival when every cp.async this thread has issued so far has landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start address >> 4 in [0,14), LBO (unused for swizzled K-major) = 1 in [16,30), SBO = 1024 B (8 rows x 128 B) >> 4
// in [32,46), version 1 in [46,48), layout type 2 (SWIZZLE_128B) in [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t umma_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// same with A = B = fp16 (format code 0)
__device__ __forceinline__ uint32_t umma_idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// Same MMA with the shared-memory descriptors given as (low word, shared high word): the issuing thread keeps the
// per-stage low words in registers and adds 2 per 32-byte K step, so a K block costs a handful of instructions.
// desc_hi = SBO 1024 B | version 1 | SWIZZLE_128B; desc_lo = (address >> 4) | LBO 1.
constexpr uint32_t UMMA_DESC_HI_SW128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3fffu) | (1u << 16); }
__device__ __forceinline__ void umma_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(UMMA_DESC_HI_SW128)
      : "memory");
}
// One K block (64 fp16 = four K=16 MMAs over one 128-byte-swizzled stage) + its commit, executed by the WHOLE warp
// with warp-uniform operands; elect.sync predicates the tcgen05 instructions themselves.  Keeping control flow and
// operands uniform lets ptxas hold the descriptors in uniform registers -- under `if (lane == 0)` every operand of
// every UTCHMMA goes through a chain of dependent R2UR moves, ~100 cycles per MMA, which paced the whole ring.
__device__ __forceinline__ void umma_f16_kblock(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc,
                                                uint64_t* commit_bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe, pa, pt;\n\t"
      ".reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pa;\n\t"
      "add.s64 da, da, 2;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.s64 da, da, 2;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.s64 da, da, 2;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(UMMA_DESC_HI_SW128), "r"(smem_u32(commit_bar))
      : "memory");
}
// One K block whose A descriptor has its own high word (stride between 8-row groups, swizzle base offset): the A
// operand of a 3x3 convolution tap read in place from a halo tile (conv.cu).
__device__ __forceinline__ void umma_f16_kblock_ahi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                    uint32_t idesc, uint32_t acc, uint64_t* commit_bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe, pa, pt;\n\t"
      ".reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "mov.b64 da, {%1, %7};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pa;\n\t"
      "add.s64 da, da, 2;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.s64 da, da, 2;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.s64 da, da, 2;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(UMMA_DESC_HI_SW128), "r"(smem_u32(commit_bar)), "r"(a_hi)
      : "memory");
}
// the same without the trailing commit (another tile's MMAs against the same weight stage follow)
__device__ __forceinline__ void umma_f16_kblock_ahi_nc(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                       uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe, pa, pt;\n\t"
      ".reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "mov.b64 da, {%1, %6};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pa;\n\t"
      "add.s64 da, da, 2;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.s64 da, da, 2;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.s64 da, da, 2;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(UMMA_DESC_HI_SW128), "r"(a_hi)
      : "memory");
}
// The same K block with the A operand in tensor memory (A-from-TMEM form): the row's 64 K values are 32 consecutive
// 32-bit columns of its lane (two fp16 per column, even k in the low half), 8 columns per K=16 step.
__device__ __forceinline__ void umma_f16_ts_kblock(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc,
                                                   uint32_t acc, uint64_t* commit_bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe, pa, pt;\n\t"
      ".reg .b64 db;\n\t"
      ".reg .b32 ta;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "mov.b32 ta, %1;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pa;\n\t"
      "add.u32 ta, ta, 8;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
      "add.u32 ta, ta, 8;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
      "add.u32 ta, ta, 8;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(acc), "r"(UMMA_DESC_HI_SW128), "r"(smem_u32(commit_bar))
      : "memory");
}
// the same two K blocks without the trailing commit (the caller commits once per ring stage)
__device__ __forceinline__ void umma_f16_kblock_nc(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe, pa, pt;\n\t"
      ".reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pa;\n\t"
      "add.s64 da, da, 2;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.s64 da, da, 2;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "add.s64 da, da, 2;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n\t"
      "}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(UMMA_DESC_HI_SW128)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ts_kblock_nc(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc,
                                                   uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe, pa, pt;\n\t"
      ".reg .b64 db;\n\t"
      ".reg .b32 ta;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "mov.b32 ta, %1;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pa;\n\t"
      "add.u32 ta, ta, 8;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
      "add.u32 ta, ta, 8;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
      "add.u32 ta, ta, 8;\n\t"
      "add.s64 db, db, 2;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, pt;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(acc), "r"(UMMA_DESC_HI_SW128)
      : "memory");
}
// whole warp, one elected lane commits
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole-warp; ncols a power of two in [32, 512]
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}

// this warp's 32 TMEM lanes x 32 consecutive columns -> 32 registers per thread (waits for completion)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 16 columns, no wait: issue several, then tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* f) {
  uint32_t* v = reinterpret_cast<uint32_t*>(f);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
// 8 columns, no wait
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float* f) {
  uint32_t* v = reinterpret_cast<uint32_t*>(f);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16_nowait(uint32_t taddr, const float* f) {
  const uint32_t* v = reinterpret_cast<const uint32_t*>(f);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8_nowait(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace ps
