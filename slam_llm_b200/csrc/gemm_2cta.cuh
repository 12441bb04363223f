// CTA-pair (cta_group::2) variant of the tcgen05 GEMM: two CTAs on the two SMs of one TPC compute one 256 x BLOCK_N tile.
// Each CTA stages ITS 128 rows of A and ITS half (BLOCK_N/2 rows) of the B tile; the leader CTA (cluster rank 0) issues
// tcgen05.mma.cta_group::2 with M = 256, which reads A from both CTAs' shared memory, shares the two B halves across the pair
// and accumulates rows [0,128) in the leader's TMEM and rows [128,256) in the peer's.  Per MMA cycle an SM therefore stages
// 2/3 of the bytes of the 128 x 256 single-CTA tile (16 KB A + 16 KB B-half instead of 16 KB + 32 KB per k-block), which is
// what bounds the single-CTA mainloop (shared-memory fill + operand read bandwidth, and bytes in flight per SM).
//
// HALVES = 2 ("pair512"): each CTA owns 256 rows of A (two M = 256 instructions per k-step, same B tile), the pair a 512 x BLOCK_N tile with ONE
// accumulator set (2 x BLOCK_N <= 512 TMEM columns, so no epilogue overlap).  Per MMA cycle an SM stages (256 + BLOCK_N / 2) rows instead of
// 2 x (128 + BLOCK_N / 2): 21 % fewer operand bytes from L2 at BLOCK_N = 192 (cuBLAS runs 256 x 208 per CTA on these shapes:
// profiles/r02_cublas_peek.txt).  Used when the whole GEMM is ONE round of such tiles (weights of 4096 rows x 1604 tokens = 72 tiles on 74 pairs),
// where the second accumulator set of the 256-row schedule hid nothing but the first of two epilogues.
//
// Synchronisation (offsets are identical in both CTAs):
//   full[s]    leader only; 1 arrival (leader producer, expect_tx = both CTAs' stage bytes); both producers' TMA loads
//              complete_tx on it (cp.async.bulk.tensor ... .cta_group::2 with the leader's barrier address);
//   empty[s]   in each CTA; tcgen05.commit.cta_group::2 multicast from the leader frees the slot in both CTAs;
//   tfull[a]   in each CTA; multicast commit publishes the accumulator to both epilogues;
//   tempty[a]  leader only; 16 arrivals = 8 epilogue warps of each CTA (the peer arrives remotely through the cluster window).
#pragma once
#include "gemm_common.cuh"

namespace slam {

template <int BLOCK_N, int HALVES = 1>
struct GemmPairCfg {
  static constexpr int ACC_STAGES = HALVES == 1 ? 2 : 1;
  static constexpr int A_BYTES = HALVES * 128 * GEMM_BK * 2;
  static constexpr int B_BYTES = (BLOCK_N / 2) * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = (200 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int ACC_COLS = ACC_STAGES * HALVES * BLOCK_N;
  static constexpr int TMEM_COLS = ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512);
  static constexpr int BAR_BYTES = 256;
  static constexpr int EPI_BYTES = GEMM_EPI_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + EPI_BYTES + 1024;
  static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
  static_assert((BLOCK_N / 2) % 8 == 0 && BLOCK_N % 16 == 0, "B half must be whole 8-row swizzle groups");
};

template <int BLOCK_N, int HALVES = 1>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const GemmKParams p) {
  using Cfg = GemmPairCfg<BLOCK_N, HALVES>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int ROWS_CTA = 128 * HALVES;                    // rows of A (and of the output tile) per CTA
  constexpr int ACC_STAGES = Cfg::ACC_STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* epi_stage = smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = static_cast<int>(blockIdx.x >> 1);
  const int npairs = static_cast<int>(gridDim.x >> 1);
  if (threadIdx.x == 0) SLAM_TRACE(0);                      // kernel entry

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.kb2 > 0) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * GEMM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();      // barriers of BOTH CTAs are initialised before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) SLAM_TRACE(1);                      // prologue done (barriers, TMEM, cluster sync)
  pdl_trigger();

  const int total_items = p.num_m_tiles * p.num_n_tiles * p.ksplit;   // num_m_tiles counts pair tiles of 2 x ROWS_CTA rows
  const int nkb = p.kb1 + p.kb2;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs; whole warp, elected lane issues)
    const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
    const uint32_t full_leader = mapa_shared(smem_u32(full_bar), 0);
    // tile order: M fastest (consecutive pairs share a B tile = weight tile of the normal orientation).  Swap-AB (transpose_out): the weight
    // is the A operand, so N runs fastest and the pairs that work at the same time share the big operand again.
    auto rows_of = [&](int item, int& row_a, int& row_b) {
      const int tile = item / p.ksplit;
      const int m_tile = p.transpose_out ? tile / p.num_n_tiles : tile % p.num_m_tiles;
      const int n_tile = p.transpose_out ? tile % p.num_n_tiles : tile / p.num_m_tiles;
      row_a = m_tile * 2 * ROWS_CTA + static_cast<int>(rank) * ROWS_CTA;
      row_b = n_tile * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2);
    };
    // Frozen operands (static_ops) do not depend on the preceding kernel: the first ring of their tiles is requested BEFORE
    // griddepcontrol.wait, under that kernel's tail.  Each such stage is armed for its full byte count; the dependent operand of the
    // same stages follows after the wait (`pre` k-blocks of the first work item, all inside the first K segment).
    int pre = 0;
    if (p.static_ops != 0 && pair < total_items) {
      int row_a, row_b;
      rows_of(pair, row_a, row_b);
      const int kb_begin = (pair % p.ksplit) * p.kb_per_split;
      const int kb_end = min(min(nkb, kb_begin + p.kb_per_split), p.kb1);
      pre = max(0, min(STAGES, kb_end - kb_begin));
      for (int s = 0; s < pre; ++s) {
        if (rank == 0) mbar_arrive_expect_tx_elect(smem_u32(&full_bar[s]), 2 * Cfg::STAGE_BYTES);
        const uint32_t fb = full_leader + s * 8;
        if (p.static_ops & 1) tma_load_2d_pair_elect(sA_u + s * Cfg::A_BYTES, &tmA, fb, (kb_begin + s) * GEMM_BK, row_a);
        if (p.static_ops & 2) tma_load_2d_pair_elect(sB_u + s * Cfg::B_BYTES, &tmB, fb, (kb_begin + s) * GEMM_BK, row_b);
      }
    }
    pdl_wait();                                             // everything else may have been written by the preceding kernel
    if (lane == 0) SLAM_TRACE(2);
    uint32_t stage = 0, ph = 0;
    for (int item = pair; item < total_items; item += npairs) {
      const int kb_begin = (item % p.ksplit) * p.kb_per_split;
      const int kb_end = min(nkb, kb_begin + p.kb_per_split);
      int row_a, row_b;
      rows_of(item, row_a, row_b);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const uint32_t fb = full_leader + stage * 8;
        if (pre > 0) {                                      // stage armed above, frozen operand in flight: only the dependent one is missing
          --pre;
          if (!(p.static_ops & 1)) tma_load_2d_pair_elect(sA_u + stage * Cfg::A_BYTES, &tmA, fb, kb * GEMM_BK, row_a);
          if (!(p.static_ops & 2)) tma_load_2d_pair_elect(sB_u + stage * Cfg::B_BYTES, &tmB, fb, kb * GEMM_BK, row_b);
          if (lane == 0 && kb == kb_begin) SLAM_TRACE(3);
        } else {
          mbar_wait(&empty_bar[stage], ph ^ 1u);
          if (lane == 0 && item == pair && kb == kb_begin) SLAM_TRACE(3);   // first TMA load issued
          if (rank == 0) mbar_arrive_expect_tx_elect(smem_u32(&full_bar[stage]), 2 * Cfg::STAGE_BYTES);
          if (kb < p.kb1) {
            tma_load_2d_pair_elect(sA_u + stage * Cfg::A_BYTES, &tmA, fb, kb * GEMM_BK, row_a);
            tma_load_2d_pair_elect(sB_u + stage * Cfg::B_BYTES, &tmB, fb, kb * GEMM_BK, row_b);
          } else {
            const int k2 = kb - p.kb1;
            tma_load_2d_pair_elect(sA_u + stage * Cfg::A_BYTES, &tmA2, fb, k2 * GEMM_BK, row_a);
            tma_load_2d_pair_elect(sB_u + stage * Cfg::B_BYTES, &tmB2, fb, k2 * GEMM_BK, row_b);
          }
        }
        if (++stage == STAGES) {
          stage = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only; whole warp, uniform operands)
    // (touches shared / tensor memory only, behind the full barriers: no griddepcontrol.wait of its own)
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BLOCK_N);
      const uint32_t my_a = sw128_kmajor_desc_lo(smem_u32(sA) + (lane < STAGES ? lane : 0) * Cfg::A_BYTES);
      const uint32_t my_b = sw128_kmajor_desc_lo(smem_u32(sB) + (lane < STAGES ? lane : 0) * Cfg::B_BYTES);
      const uint32_t empty_u = smem_u32(empty_bar), tfull_u = smem_u32(tfull_bar);
      uint32_t stage = 0, ph = 0;
      uint32_t it = 0;
      for (int item = pair; item < total_items; item += npairs, ++it) {
        const int kb_begin = (item % p.ksplit) * p.kb_per_split;
        const int kb_end = min(nkb, kb_begin + p.kb_per_split);
        const uint32_t acc = it % ACC_STAGES;
        const uint32_t aph = (it / ACC_STAGES) & 1u;
        mbar_wait_cluster(&tempty_bar[acc], aph ^ 1u);     // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * HALVES * BLOCK_N;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], ph);
          tc_fence_after();
          if (lane == 0 && it == 0 && kb == kb_begin) SLAM_TRACE(4);      // first operands have landed
          if (lane == 0 && it == 0 && kb == kb_begin + 8) SLAM_TRACE(14); // ... and the 9th k-block (ring refilled once)
          const uint32_t a_lo = __shfl_sync(0xffffffffu, my_a, stage);
          const uint32_t b_lo = __shfl_sync(0xffffffffu, my_b, stage);
          umma_kblock_pair(d_tmem, a_lo, b_lo, idesc, kb > kb_begin ? 1u : 0u);
          if constexpr (HALVES == 2) umma_kblock_pair(d_tmem + BLOCK_N, a_lo + 1024u, b_lo, idesc, kb > kb_begin ? 1u : 0u);   // rows 128..255 of each CTA: A tile + 16 KB
          umma_commit_pair_elect(empty_u + stage * 8);
          if (kb == kb_end - 1) umma_commit_pair_elect(tfull_u + acc * 8);
          if (lane == 0 && kb == kb_end - 1) SLAM_TRACE(it == 0 ? 5 : 6);  // last MMA of the first / of the latest tile issued
          if (++stage == STAGES) {
            stage = 0;
            ph ^= 1u;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (both CTAs, each drains its own 128 rows; 8 warps)
    pdl_wait();                                             // residual / aux reads and the output writes depend on the preceding kernel
    const int e = warp - 4;
    const int q = e & 3, h = e >> 2;
    uint8_t* stg = epi_stage + e * (32 * GEMM_EPI_PITCH);
    constexpr int NCH = BLOCK_N / 32;
    const uint32_t tempty_leader0 = mapa_shared(smem_u32(&tempty_bar[0]), 0);
    uint32_t it = 0;
    for (int item = pair; item < total_items; item += npairs, ++it) {
      const int tile = item / p.ksplit;
      const int m_tile = p.transpose_out ? tile / p.num_n_tiles : tile % p.num_m_tiles;
      const int n_tile = p.transpose_out ? tile % p.num_n_tiles : tile / p.num_m_tiles;
      const uint32_t acc = it % ACC_STAGES;
      const uint32_t aph = (it / ACC_STAGES) & 1u;
      mbar_wait(&tfull_bar[acc], aph);
      tc_fence_after();
      if (e == 0 && lane == 0) SLAM_TRACE(it == 0 ? 7 : 9);             // accumulator of the first / latest tile complete
      const int n0 = n_tile * BLOCK_N;
      const int row_base0 = m_tile * 2 * ROWS_CTA + static_cast<int>(rank) * ROWS_CTA + q * 32;
      const uint32_t taddr0 = tmem_base + acc * HALVES * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
      [[maybe_unused]] const int row_base = row_base0;
      [[maybe_unused]] const uint32_t taddr = taddr0;
      auto release_tmem = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(tempty_leader0 + acc * 8);
      };
      if (p.act == 3) {
        // fused SwiGLU forward: in the blocked-64 layout chunk 4b + h is a gate chunk and 4b + h + 2 its up partner
        if constexpr (BLOCK_N % 128 == 0) {
#pragma unroll 1
          for (int cg = h; cg < NCH; cg += 4) {
            uint32_t rg[32], ru[32];
            tmem_ld_32x32(taddr + cg * 32, rg);
            tmem_ld_32x32(taddr + (cg + 2) * 32, ru);
            tmem_ld_wait();
            if (cg + 4 >= NCH) release_tmem();
            float ag[32], au[32];
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              ag[t] = __uint_as_float(rg[t]);
              au[t] = __uint_as_float(ru[t]);
            }
            gemm_epilogue_swiglu_fwd(p, ag, au, row_base, n0 + cg * 32, stg, lane);
          }
        }
        continue;
      }
      if constexpr (HALVES == 2) {
        // 2 x NCH chunks per lane quarter (accumulator half u / NCH = rows +128), shared by the two warps of the quarter; act 0 only
#pragma unroll 1
        for (int u = h; u < 2 * NCH; u += 2) {
          const int half = u / NCH, c = u % NCH;
          const int rb = row_base0 + half * 128;
          uint4 rsd[4];
          gemm_residual_prefetch(p, rb, lane, n0 + c * 32, rsd);
          uint32_t r[32];
          tmem_ld_32x32(taddr0 + half * BLOCK_N + c * 32, r);
          tmem_ld_wait();
          if (u + 2 >= 2 * NCH) release_tmem();
          float accv[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) accv[t] = __uint_as_float(r[t]);
          gemm_epilogue_chunk(p, accv, rsd, rb, n0 + c * 32, stg, lane);
        }
        if (e == 0 && lane == 0) {
          SLAM_TRACE(it == 0 ? 8 : 10);
          SLAM_TRACE_V(13, it + 1);
        }
        continue;
      }
#pragma unroll 1
      for (int c = h; c < NCH; c += 2) {
        uint4 rsd[4], rsd2[4];
        if (p.act == 4) {
          if (p.transpose_out) gemm_swiglu_bwd_prefetch_t(p, row_base, lane, n0 + c * 32, rsd, rsd2);
          else gemm_swiglu_bwd_prefetch(p, row_base + lane, n0 + c * 32, rsd, rsd2);
        } else {
          gemm_residual_prefetch(p, row_base, lane, n0 + c * 32, rsd);
        }
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        if (c + 2 >= NCH) release_tmem();
        float accv[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) accv[t] = __uint_as_float(r[t]);
        if (p.act == 4) {
          if (p.transpose_out) gemm_epilogue_swiglu_bwd_t(p, accv, rsd, rsd2, row_base, n0 + c * 32, stg, lane);
          else gemm_epilogue_swiglu_bwd(p, accv, rsd, rsd2, row_base, n0 + c * 32, stg, lane);
        } else {
          gemm_epilogue_chunk(p, accv, rsd, row_base, n0 + c * 32, stg, lane);
        }
      }
      if (e == 0 && lane == 0) {
        SLAM_TRACE(it == 0 ? 8 : 10);                                     // epilogue of the first / latest tile done (warp 4's share)
        SLAM_TRACE_V(13, it + 1);                                         // tiles done by this CTA
      }
    }
  }
  if (threadIdx.x == 0) SLAM_TRACE(11);                     // this thread's role finished

  tc_fence_before();
  cluster_sync_all();      // no CTA of the pair may exit (or free TMEM) while the other can still touch its memory
  tc_fence_after();
  if (warp == 2) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  if (threadIdx.x == 0) SLAM_TRACE(12);                     // exit
}

}  // namespace slam
