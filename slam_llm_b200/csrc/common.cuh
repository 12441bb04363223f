// Common sm_100a device helpers: mbarrier, TMA, tcgen05/TMEM PTX wrappers, warp reductions.
// Hand-written inline PTX (no CUTLASS/CuTe); B200 (sm_100a) only.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#ifndef SLAM_WATCHDOG
#define SLAM_WATCHDOG 1  // bounded mbarrier spins: trap instead of hanging the GPU
#endif

namespace slam {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// SLAM_WAIT_HINT_NS > 0: pass a suspend-time hint to try_wait (the waiting warp may be parked that long before the instruction returns false):
// fewer wake-ups of the 8 epilogue warps + producer while a long MMA main loop runs (experiment knob; 0 = hardware default).
#ifndef SLAM_WAIT_HINT_NS
#define SLAM_WAIT_HINT_NS 0
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
#if SLAM_WAIT_HINT_NS > 0
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(static_cast<uint32_t>(SLAM_WAIT_HINT_NS))
      : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
#endif
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if SLAM_WATCHDOG
  long long t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins == 1024u) t0 = clock64();
    if (spins > 1024u && (spins & 1023u) == 0u && clock64() - t0 > 4000000000LL) {
      printf("slam_b200: mbarrier watchdog block %d thread %d\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load global -> shared, completion on an mbarrier (complete_tx::bytes).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp receives row (lane base + i), 32 consecutive columns.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Instruction descriptor: D=f32, A=B=bf16, both K-major, MxN tile.
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---------------------------------------------------------------- warp-uniform issue helpers
// The MMA-issuing warp of the GEMM kernels is instruction-bound, not data-bound (ncu source page, round 1: ~90 SASS
// instructions and ~830 cycles per 64-wide k-block when the four tcgen05.mma were issued from an `if (lane == 0)` branch -
// the compiler wraps every asm in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop).  These helpers are executed by the
// WHOLE warp with warp-uniform operands; the election happens inside the asm, so the operands live in uniform registers and
// the four UTCHMMA of a k-block issue back to back.
constexpr uint32_t SW128_KMAJOR_DESC_HI = 0x40004040u;   // SBO = 1024 B, descriptor version 1, SWIZZLE_128B (bits 32..63)
// low word of the K-major SW128 descriptor of a tile at shared address `addr` (LBO field = 1)
__device__ __forceinline__ uint32_t sw128_kmajor_desc_lo(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }

#define SLAM_UMMA_STEP(CG, DOFF, AOFF, BOFF, PRED)                                                   \
  "add.u32 alo, %1, " #AOFF ";\n\t add.u32 blo, %2, " #BOFF ";\n\t add.u32 dd, %0, " #DOFF ";\n\t"   \
  "mov.b64 da, {alo, %5};\n\t mov.b64 db, {blo, %5};\n\t"                                             \
  "@lead tcgen05.mma.cta_group::" #CG ".kind::f16 [dd], da, db, %3, " PRED ";\n\t"

// one 64-wide k-block (4 x K=16) into the accumulator at d_tmem; `accumulate` = 0 overwrites on the first step
__device__ __forceinline__ void umma_kblock_1(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, lead;\n\t.reg .b64 da, db;\n\t.reg .b32 alo, blo, dd;\n\t"
      "elect.sync _|lead, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      SLAM_UMMA_STEP(1, 0, 0, 0, "p") SLAM_UMMA_STEP(1, 0, 2, 2, "1") SLAM_UMMA_STEP(1, 0, 4, 4, "1") SLAM_UMMA_STEP(1, 0, 6, 6, "1")
      "}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(SW128_KMAJOR_DESC_HI)
      : "memory");
}
// same with an MN-major B operand (rows of the shared tile run along K): 16 K-rows = two 8-row groups = 2048 B per step
__device__ __forceinline__ void umma_kblock_mnb(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, lead;\n\t.reg .b64 da, db;\n\t.reg .b32 alo, blo, dd;\n\t"
      "elect.sync _|lead, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      SLAM_UMMA_STEP(1, 0, 0, 0, "p") SLAM_UMMA_STEP(1, 0, 2, 128, "1") SLAM_UMMA_STEP(1, 0, 4, 256, "1") SLAM_UMMA_STEP(1, 0, 6, 384, "1")
      "}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(SW128_KMAJOR_DESC_HI)
      : "memory");
}
// A AND B MN-major (both tiles stored with their K dimension along the rows): +2048 B per 16-row K step on both operands
__device__ __forceinline__ void umma_kblock_mna_mnb(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, lead;\n\t.reg .b64 da, db;\n\t.reg .b32 alo, blo, dd;\n\t"
      "elect.sync _|lead, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      SLAM_UMMA_STEP(1, 0, 0, 0, "p") SLAM_UMMA_STEP(1, 0, 128, 128, "1") SLAM_UMMA_STEP(1, 0, 256, 256, "1") SLAM_UMMA_STEP(1, 0, 384, 384, "1")
      "}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(SW128_KMAJOR_DESC_HI)
      : "memory");
}
// low descriptor word of an MN-major SW128 tile made of 64-element MN blocks `lbo_bytes` apart (8-row K groups 1024 B apart)
__device__ __forceinline__ uint32_t sw128_mnmajor_desc_lo(uint32_t addr, uint32_t lbo_bytes) {
  return ((addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}
// CTA-pair variant (cta_group::2, leader CTA only)
__device__ __forceinline__ void umma_kblock_pair(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, lead;\n\t.reg .b64 da, db;\n\t.reg .b32 alo, blo, dd;\n\t"
      "elect.sync _|lead, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      SLAM_UMMA_STEP(2, 0, 0, 0, "p") SLAM_UMMA_STEP(2, 0, 2, 2, "1") SLAM_UMMA_STEP(2, 0, 4, 4, "1") SLAM_UMMA_STEP(2, 0, 6, 6, "1")
      "}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(SW128_KMAJOR_DESC_HI)
      : "memory");
}
// one elected lane commits: mbarrier arrive once all previously issued tcgen05.mma have completed
__device__ __forceinline__ void umma_commit_elect(uint32_t bar_addr) {
  asm volatile(
      "{\n\t.reg .pred lead;\n\telect.sync _|lead, 0xffffffff;\n\t"
      "@lead tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar_addr)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair_elect(uint32_t bar_addr) {
  asm volatile(
      "{\n\t.reg .pred lead;\n\t.reg .b16 m;\n\telect.sync _|lead, 0xffffffff;\n\tmov.b16 m, 3;\n\t"
      "@lead tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(bar_addr)
      : "memory");
}
// TMA tile load issued by one elected lane of a converged warp (operands warp-uniform)
__device__ __forceinline__ void tma_load_2d_elect(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar_addr, int32_t c0, int32_t c1) {
  asm volatile(
      "{\n\t.reg .pred lead;\n\telect.sync _|lead, 0xffffffff;\n\t"
      "@lead cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_elect(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int32_t c0, int32_t c1) {
  asm volatile(
      "{\n\t.reg .pred lead;\n\telect.sync _|lead, 0xffffffff;\n\t"
      "@lead cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}" ::"r"(
          smem_dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_elect(uint32_t bar_addr, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred lead;\n\telect.sync _|lead, 0xffffffff;\n\t"
      "@lead mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar_addr), "r"(bytes)
      : "memory");
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2): two SMs of one TPC run one M=256 MMA
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// wait on a barrier that CTAs of the whole cluster arrive on (acquire at cluster scope); same watchdog as mbar_wait
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  long long t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
#if SLAM_WATCHDOG
    if (++spins == 1024u) t0 = clock64();
    if (spins > 1024u && (spins & 1023u) == 0u && clock64() - t0 > 4000000000LL) {
      printf("slam_b200: cluster mbarrier watchdog block %d thread %d\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
#endif
  }
  (void)t0;
  (void)spins;
}
// TMEM management for a CTA pair: one warp (same warp index) of EACH CTA executes these
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Every kernel of the library lets its dependents start as early as possible (pdl_trigger); kernels launched with the
// programmatic-stream-serialization attribute (the GEMMs) run their prologue (barrier init, TMEM alloc, descriptor prefetch)
// under the tail of the previous kernel and call pdl_wait() before touching memory it produced.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- misc math
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// exact-erf GELU (whisper's F.gelu) with erf from Abramowitz-Stegun 7.1.26: |error| <= 1.5e-7 absolute, i.e. far below one
// bf16 ulp of any activation; 2 MUFU + ~12 FMA instead of libm erff's ~30 instructions with a branch - the GELU epilogue of the
// encoder's fc1 GEMM was epilogue-bound (9.2 us per 128 x 256 tile against 6 us of MMA at K = 1280, tools/epi_probe.py)
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  float ex;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-z * z * 1.4426950408889634f));
  const float erf_abs = fmaf(-poly, ex, 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
// eight bf16 values as one 16-byte vector
struct alignas(16) bf16x8 {
  uint32_t w[4];
};
__device__ __forceinline__ void unpack8(const bf16x8& p, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = unpack_bf16x2(p.w[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float (&f)[8]) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.w[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
  return p;
}


}  // namespace slam
