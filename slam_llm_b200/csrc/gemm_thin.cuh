// Thin products: N <= 64 output columns over a long K - the LoRA down-projections T = x A_cat^T and U = dY (s B_cat) of every adapted linear
// (peft lora.Linear.forward / backward, reference models/slam_model.py:214-218).  They have few 128 x 64 output tiles (13 at M = 1604), so
// the one-CTA-per-tile kernel leaves > 90 % of the SMs idle while 13 CTAs walk 64-96 k-blocks each: 25-37 us per launch, on the critical
// path in front of every fused qkv / d(qkv) GEMM (tools/gemm_trace.py, round 2).
//
// Here a CLUSTER of THIN_SPLIT = 8 CTAs shares one output tile.  CTA r of the cluster accumulates k-blocks [r c, (r + 1) c), c = ceil(nkb / 8),
// in its tensor memory (tcgen05.mma, M = 128, N = 64), writes the fp32 partial tile to its own shared memory, and after a cluster barrier
// every CTA reduces 16 of the tile's 128 rows over the 8 partials through distributed shared memory (ld.shared::cluster) IN RANK ORDER -
// the sum is deterministic - and stores them as bf16.  No workspace, no atomics, no second launch.
//   warp 0      TMA producer (A tile 128 x 64, B tile 64 x 64 per k-block, 128B swizzle)
//   warp 1      TMEM allocation + MMA issuer
//   warps 2-5   partial tile TMEM -> shared memory (thread = row), then the reduce-scatter over the cluster and the bf16 store
#pragma once
#include "gemm_common.cuh"

namespace slam {

constexpr int THIN_SPLIT = 8;                          // CTAs per output tile (portable cluster size)
constexpr int THIN_STAGES = 6;
constexpr int THIN_BN = 64;
constexpr int THIN_A_BYTES = 128 * GEMM_BK * 2;        // 16 KB
constexpr int THIN_B_BYTES = THIN_BN * GEMM_BK * 2;    //  8 KB
constexpr int THIN_PITCH = THIN_BN + 4;                // floats per row of the partial tile: thread = row float4 stores are conflict-free
constexpr int THIN_THREADS = 192;
constexpr int THIN_P_BYTES = 128 * THIN_PITCH * 4;
constexpr int THIN_SMEM_BYTES = THIN_STAGES * (THIN_A_BYTES + THIN_B_BYTES) + THIN_P_BYTES + 256 + 1024;

struct ThinParams {
  int M, N, nkb, kb_per_cta;
  bf16* out;
  long long ldo;
  float alpha;
};

__device__ __forceinline__ float4 ld_shared_cluster_f4(uint32_t cluster_addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr) : "memory");
  return v;
}

__global__ void __cluster_dims__(THIN_SPLIT, 1, 1) __launch_bounds__(THIN_THREADS, 1)
gemm_thin_cluster_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ThinParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + THIN_STAGES * THIN_A_BYTES;
  float* sP = reinterpret_cast<float*>(smem + THIN_STAGES * (THIN_A_BYTES + THIN_B_BYTES));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sP) + THIN_P_BYTES);
  uint64_t* empty_bar = full_bar + THIN_STAGES;
  uint64_t* tfull_bar = empty_bar + THIN_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int m_tile = static_cast<int>(blockIdx.x) / THIN_SPLIT;
  const int kb_begin = static_cast<int>(rank) * p.kb_per_cta;
  const int kb_end = min(p.nkb, kb_begin + p.kb_per_cta);          // (kb_begin >= nkb: this CTA contributes a zero partial)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < THIN_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, THIN_BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB), full_u = smem_u32(full_bar);
    uint32_t stage = 0, ph = 0;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      mbar_wait(&empty_bar[stage], ph ^ 1u);
      const uint32_t fb = full_u + stage * 8;
      mbar_arrive_expect_tx_elect(fb, THIN_A_BYTES + THIN_B_BYTES);
      tma_load_2d_elect(sA_u + stage * THIN_A_BYTES, &tmA, fb, kb * GEMM_BK, m_tile * 128);
      tma_load_2d_elect(sB_u + stage * THIN_B_BYTES, &tmB, fb, kb * GEMM_BK, 0);
      if (++stage == THIN_STAGES) {
        stage = 0;
        ph ^= 1u;
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, THIN_BN);
    const uint32_t my_a = sw128_kmajor_desc_lo(smem_u32(sA) + (lane < THIN_STAGES ? lane : 0) * THIN_A_BYTES);
    const uint32_t my_b = sw128_kmajor_desc_lo(smem_u32(sB) + (lane < THIN_STAGES ? lane : 0) * THIN_B_BYTES);
    const uint32_t empty_u = smem_u32(empty_bar), tfull_u = smem_u32(tfull_bar);
    uint32_t stage = 0, ph = 0;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      mbar_wait(&full_bar[stage], ph);
      tc_fence_after();
      const uint32_t a_lo = __shfl_sync(0xffffffffu, my_a, stage);
      const uint32_t b_lo = __shfl_sync(0xffffffffu, my_b, stage);
      umma_kblock_1(tmem_base, a_lo, b_lo, idesc, kb > kb_begin ? 1u : 0u);
      umma_commit_elect(empty_u + stage * 8);
      if (kb == kb_end - 1) umma_commit_elect(tfull_u);
      if (++stage == THIN_STAGES) {
        stage = 0;
        ph ^= 1u;
      }
    }
  } else {
    // partial tile -> shared memory; thread = accumulator row (TMEM lane quarter = warp % 4)
    const int row = (warp & 3) * 32 + lane;
    float* prow = sP + row * THIN_PITCH;
    if (kb_end > kb_begin) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
#pragma unroll
      for (int c = 0; c < THIN_BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<float4*>(prow + c * 32 + 4 * g) =
              make_float4(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]), __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
      }
      tc_fence_before();
    } else {
#pragma unroll
      for (int g = 0; g < THIN_BN / 4; ++g) *reinterpret_cast<float4*>(prow + 4 * g) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

  cluster_sync_all();                                   // every partial of the cluster is in shared memory (release / acquire at cluster scope)

  if (warp >= 2) {
    // reduce-scatter: this CTA owns rows [16 rank, 16 rank + 16) of the tile; thread -> one row, 8 consecutive columns
    const int t = static_cast<int>(threadIdx.x) - 64;
    const int row_l = static_cast<int>(rank) * 16 + (t >> 3);
    const int col = (t & 7) * 8;
    const uint32_t local = smem_u32(sP + row_l * THIN_PITCH + col);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (uint32_t s = 0; s < THIN_SPLIT; ++s) {        // fixed order: the result does not depend on timing
      const uint32_t ra = mapa_shared(local, s);
      const float4 v0 = ld_shared_cluster_f4(ra), v1 = ld_shared_cluster_f4(ra + 16);
      acc[0] += v0.x; acc[1] += v0.y; acc[2] += v0.z; acc[3] += v0.w;
      acc[4] += v1.x; acc[5] += v1.y; acc[6] += v1.z; acc[7] += v1.w;
    }
    const int grow = m_tile * 128 + row_l;
    if (grow < p.M && col < p.N) {                      // (N is a multiple of 8)
      const uint4 pk = make_uint4(pack_bf16x2(acc[0] * p.alpha, acc[1] * p.alpha), pack_bf16x2(acc[2] * p.alpha, acc[3] * p.alpha),
                                  pack_bf16x2(acc[4] * p.alpha, acc[5] * p.alpha), pack_bf16x2(acc[6] * p.alpha, acc[7] * p.alpha));
      *reinterpret_cast<uint4*>(p.out + static_cast<long long>(grow) * p.ldo + col) = pk;
    }
  }

  cluster_sync_all();                                   // nobody leaves while a peer may still read its partial
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, THIN_BN);
  }
}

}  // namespace slam
