// tcgen05 / TMEM / TMA GEMM core for the SLAM-LLM training step on B200 (sm_100a).
//
//   out[M,N] = act(alpha * (A·B^T + A2·B2^T) + bias) + residual          (+ fused SwiGLU forward / backward epilogues)
//
// Design (one CTA per SM, persistent over output tiles, warp-specialised, 384 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D loads of BLOCK_M x 64 (A) and BLOCK_N x 64 (B) bf16 k-blocks into a
//               STAGES-deep 128B-swizzled shared-memory ring, completion on mbarriers;
//   warp 1      MMA issuer: the whole warp runs the loop with warp-uniform operands, one elected lane issues the four
//               tcgen05.mma (M=128, N=BLOCK_N, K=16) of a k-block back to back (see "warp-uniform issue helpers" in
//               common.cuh: issuing from an `if (lane == 0)` branch made this warp the bottleneck of the kernel);
//               tcgen05.commit releases ring slots / publishes the accumulator;
//   warp 2      TMEM allocator (BLOCK_M=128: 2 accumulator stages so the epilogue of tile i overlaps tile i+1;
//               BLOCK_M=256: two M=128 accumulators that share every B tile, single accumulator set);
//   warps 4-11  epilogue (gemm_common.cuh): tcgen05.ld 32x32b -> math in the row layout -> bf16 transpose through shared
//               memory -> 16-byte stores; two warps per TMEM lane quarter, interleaved 32-column chunks.
// The K loop runs over TWO operand pairs back to back ("dual K segment"): the base weights and the
// rank-padded LoRA pair, so y = xW^T + (alpha/r)(xA^T)B^T is produced in ONE accumulator tile
// (reference: peft lora.Linear.forward called under models/slam_model.py:400).
// gemm_2cta.cuh holds the CTA-pair (cta_group::2) variant; GemmSchedule (gemm_common.cuh) the tile order and the tail split.
#include <atomic>
#include <chrono>
#include <mutex>

#include "../../include/slam_b200.h"
#include "common.cuh"
#include "gemm_common.cuh"
#include "gemm_2cta.cuh"
#include "gemm_thin.cuh"
#include "host.cuh"

namespace slam {

// (An L2 prefetch of the weight operand 4/8/16 k-blocks ahead measured ~10 % SLOWER than none: profiles/r01_exp_prefetch.log.)

template <int BLOCK_M, int BLOCK_N>
struct GemmCfg {
  static constexpr int HALVES = BLOCK_M / 128;             // M=128 accumulators per tile
  static constexpr int ACC_STAGES = BLOCK_M == 128 ? 2 : 1;
  static constexpr int A_BYTES = BLOCK_M * GEMM_BK * 2;
  static constexpr int B_BYTES = BLOCK_N * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = (200 * 1024) / STAGE_BYTES;
#ifdef SLAM_STAGES_CAP
  static constexpr int STAGES = STAGES_RAW > SLAM_STAGES_CAP ? SLAM_STAGES_CAP : STAGES_RAW;   // experiment: latency- vs bandwidth-bound
#else
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
#endif
  static constexpr int ACC_COLS = ACC_STAGES * HALVES * BLOCK_N;
  static constexpr int TMEM_COLS = ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512);  // power of two
  static constexpr int BAR_BYTES = 256;
  static constexpr int EPI_BYTES = GEMM_EPI_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + EPI_BYTES + 1024;
  static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
};

template <int BLOCK_M, int BLOCK_N>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                    const GemmKParams p) {
  using Cfg = GemmCfg<BLOCK_M, BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int ACC_STAGES = Cfg::ACC_STAGES;
  constexpr int HALVES = Cfg::HALVES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* xchg_bar = tempty_bar + 2;   // [GEMM_EPI_WARPS] tail-split exchange: one per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xchg_bar + GEMM_EPI_WARPS);
  uint8_t* epi_stage = (smem + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES) + Cfg::BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

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
      mbar_init(&tempty_bar[s], GEMM_EPI_WARPS);
    }
    for (int s = 0; s < GEMM_EPI_WARPS; ++s) mbar_init(&xchg_bar[s], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();   // dependents may begin their own prologue now ...
  pdl_wait();      // ... and this grid must not read its operands before the producing grid has completed

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp, one elected lane issues)
    const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB), full_u = smem_u32(full_bar);
    uint32_t stage = 0, ph = 0;
    GemmSchedule sched(p, static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x));
    GemmSeg sg;
    while (sched.next(sg)) {
      const int row_a = (sg.tile % p.num_m_tiles) * BLOCK_M;
      const int row_b = (sg.tile / p.num_m_tiles) * BLOCK_N;
      for (int kb = sg.kb_begin; kb < sg.kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], ph ^ 1u);
        const uint32_t fb = full_u + stage * 8;
        mbar_arrive_expect_tx_elect(fb, Cfg::A_BYTES + Cfg::B_BYTES);
        if (kb < p.kb1) {
          tma_load_2d_elect(sA_u + stage * Cfg::A_BYTES, &tmA, fb, kb * GEMM_BK, row_a);
          tma_load_2d_elect(sB_u + stage * Cfg::B_BYTES, &tmB, fb, kb * GEMM_BK, row_b);
        } else {
          const int k2 = kb - p.kb1;
          tma_load_2d_elect(sA_u + stage * Cfg::A_BYTES, &tmA2, fb, k2 * GEMM_BK, row_a);
          tma_load_2d_elect(sB_u + stage * Cfg::B_BYTES, &tmB2, fb, k2 * GEMM_BK, row_b);
        }
        if (++stage == STAGES) {
          stage = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp; see "warp-uniform issue helpers")
    constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N);
    // lane s keeps the descriptor low words of ring slot s; a shuffle by the (uniform) stage index yields uniform operands
    const uint32_t my_a = sw128_kmajor_desc_lo(smem_u32(sA) + (lane < STAGES ? lane : 0) * Cfg::A_BYTES);
    const uint32_t my_b = sw128_kmajor_desc_lo(smem_u32(sB) + (lane < STAGES ? lane : 0) * Cfg::B_BYTES);
    const uint32_t empty_u = smem_u32(empty_bar), tfull_u = smem_u32(tfull_bar);
    uint32_t stage = 0, ph = 0;
    uint32_t it = 0;
    GemmSchedule sched(p, static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x));
    GemmSeg sg;
    for (; sched.next(sg); ++it) {
      const uint32_t acc = it % ACC_STAGES;
      const uint32_t aph = (it / ACC_STAGES) & 1u;
      mbar_wait(&tempty_bar[acc], aph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * HALVES * BLOCK_N;
      for (int kb = sg.kb_begin; kb < sg.kb_end; ++kb) {
        mbar_wait(&full_bar[stage], ph);
        tc_fence_after();
        const uint32_t a_lo = __shfl_sync(0xffffffffu, my_a, stage);
        const uint32_t b_lo = __shfl_sync(0xffffffffu, my_b, stage);
        const uint32_t accum = kb > sg.kb_begin ? 1u : 0u;
        umma_kblock_1(d_tmem, a_lo, b_lo, idesc, accum);
        if constexpr (HALVES == 2) umma_kblock_1(d_tmem + BLOCK_N, a_lo + 1024u, b_lo, idesc, accum);   // rows 128..255: A tile + 16 KB
        umma_commit_elect(empty_u + stage * 8);
        if (kb == sg.kb_end - 1) umma_commit_elect(tfull_u + acc * 8);
        if (++stage == STAGES) {
          stage = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: 8 warps; warp e drains TMEM lane quarter e % 4,
    // 32-column chunks h, h + 2, h + 4, ... with h = e / 4 (see gemm_epilogue_chunk)
    const int e = warp - 4;
    const int q = e & 3, h = e >> 2;
    uint8_t* stg = epi_stage + e * (32 * GEMM_EPI_PITCH);
    constexpr int NCH = BLOCK_N / 32;
    uint32_t it = 0;
    GemmSchedule sched(p, static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x));
    GemmSeg sg;
    const int cta = static_cast<int>(blockIdx.x);
    for (; sched.next(sg); ++it) {
      const int m_tile = sg.tile % p.num_m_tiles;
      const int n_tile = sg.tile / p.num_m_tiles;
      const uint32_t acc = it % ACC_STAGES;
      const uint32_t aph = (it / ACC_STAGES) & 1u;
      mbar_wait(&tfull_bar[acc], aph);
      tc_fence_after();
      const int n0 = n_tile * BLOCK_N;
      if (sg.kind == 1) {
        if constexpr (HALVES == 1) {
          // ---- tail tile: exchange partial accumulators with the other k-slices of this tile (CTAs base .. base + S - 1).
          // This warp's chunks are c = h + 2 j; slice (j % S) finishes chunk j, the others publish their partial of it.
          const int S = p.tail_slices, slice = cta % S, base = cta - slice;
          const int NL = (NCH - h + 1) / 2;                                          // chunks of this warp
          const uint32_t taddr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
          const int row_base = m_tile * BLOCK_M + q * 32;
          // a slot is [epilogue warp][chunk j][row][32 floats]: 4 KB blocks, 16-byte pieces XOR-swizzled by (row % 8) so that the
          // finisher's thread-per-row reads of the block from shared memory are bank-conflict free
          auto block_of = [&](int cta_id, int j) { return p.sk_partials + ((static_cast<long long>(cta_id) * GEMM_EPI_WARPS + e) * 4 + j) * 1024; };
#pragma unroll 1
          for (int j = 0; j < NL; ++j) {
            if (j % S == slice) continue;
            uint32_t r[32];
            tmem_ld_32x32(taddr + (h + 2 * j) * 32, r);
            tmem_ld_wait();
            float4* dst = reinterpret_cast<float4*>(block_of(cta, j) + lane * 32);
#pragma unroll
            for (int t = 0; t < 8; ++t)
              dst[t ^ (lane & 7)] = make_float4(__uint_as_float(r[4 * t]), __uint_as_float(r[4 * t + 1]), __uint_as_float(r[4 * t + 2]), __uint_as_float(r[4 * t + 3]));
          }
          __threadfence();            // every lane's partial stores are visible device-wide before the flag
          __syncwarp();
          // the tail item is the last work of the CTA: the operand ring is idle, 24 KB of it per epilogue warp receive the peers' blocks
          float* xbuf = reinterpret_cast<float*>(smem + e * (24 * 1024));
          if (lane == 0) {
            sk_flag_publish(p.sk_flags + cta * GEMM_EPI_WARPS + e, p.sk_epoch);
            for (int j = 0; j < S; ++j)
              if (j != slice) sk_flag_wait(p.sk_flags + (base + j) * GEMM_EPI_WARPS + e, p.sk_epoch);
            fence_proxy_async_global();   // peers' generic-proxy stores (acquired above) -> visible to the bulk-copy engine
            uint32_t blocks = 0;
            for (int j = slice; j < NL; j += S) blocks += static_cast<uint32_t>(S - 1);
            mbar_arrive_expect_tx(&xchg_bar[e], blocks * 4096u);
            uint32_t slot = 0;
            for (int j = slice; j < NL; j += S)
              for (int s2 = 0; s2 < S; ++s2)
                if (s2 != slice) bulk_load_1d(xbuf + (slot++) * 1024, block_of(base + s2, j), 4096u, &xchg_bar[e]);
          }
          __syncwarp();
          mbar_wait(&xchg_bar[e], 0);
          // finish this slice's chunks: sum in slice order (own accumulator at position `slice`), then the epilogue
          uint32_t slot = 0;
#pragma unroll 1
          for (int j = slice; j < NL; j += S) {
            const int c = h + 2 * j;
            uint4 rsd[4];
            gemm_residual_prefetch(p, row_base, lane, n0 + c * 32, rsd);
            uint32_t r[32];
            tmem_ld_32x32(taddr + c * 32, r);
            tmem_ld_wait();
            float accv[32];
            bool first = true;
            for (int s2 = 0; s2 < S; ++s2) {
              if (s2 == slice) {
#pragma unroll
                for (int t = 0; t < 32; ++t) accv[t] = first ? __uint_as_float(r[t]) : accv[t] + __uint_as_float(r[t]);
              } else {
                const float4* src = reinterpret_cast<const float4*>(xbuf + (slot++) * 1024 + lane * 32);
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                  const float4 v = src[t ^ (lane & 7)];
                  accv[4 * t] = first ? v.x : accv[4 * t] + v.x;
                  accv[4 * t + 1] = first ? v.y : accv[4 * t + 1] + v.y;
                  accv[4 * t + 2] = first ? v.z : accv[4 * t + 2] + v.z;
                  accv[4 * t + 3] = first ? v.w : accv[4 * t + 3] + v.w;
                }
              }
              first = false;
            }
            gemm_epilogue_chunk(p, accv, rsd, row_base, n0 + c * 32, stg, lane);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        continue;
      }
#pragma unroll 1
      for (int hf = 0; hf < HALVES; ++hf) {
        const int row_base = m_tile * BLOCK_M + hf * 128 + q * 32;
        const uint32_t taddr = tmem_base + (acc * HALVES + hf) * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
        auto release_tmem = [&]() {   // this warp's share of the accumulator is in registers: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
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
              if (hf == HALVES - 1 && cg + 4 >= NCH) release_tmem();
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
#pragma unroll 1
        for (int c = h; c < NCH; c += 2) {
          uint4 rsd[4], rsd2[4];
          if (p.act == 4) gemm_swiglu_bwd_prefetch(p, row_base + lane, n0 + c * 32, rsd, rsd2);
          else gemm_residual_prefetch(p, row_base, lane, n0 + c * 32, rsd);
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (hf == HALVES - 1 && c + 2 >= NCH) release_tmem();
          float accv[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) accv[t] = __uint_as_float(r[t]);
          if (p.act == 4) gemm_epilogue_swiglu_bwd(p, accv, rsd, rsd2, row_base, n0 + c * 32, stg, lane);
          else gemm_epilogue_chunk(p, accv, rsd, row_base, n0 + c * 32, stg, lane);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  });
  return fn;
}

// bf16 row-major [rows, k] with leading dimension ld (elements); box = 64 (K) x box_rows, 128B swizzle.
static int make_tmap(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t k, int64_t ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return -2;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0) {
    set_error("gemm operand must be 16-byte aligned with a leading dimension multiple of 8 (ptr=%p ld=%lld)",
              ptr, (long long)ld);
    return -1;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(GEMM_BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rows=%lld k=%lld ld=%lld box_rows=%d)", (int)r,
              (long long)rows, (long long)k, (long long)ld, box_rows);
    return -3;
  }
  return 0;
}

static void fill_kparams(const slam_gemm_args* g, int block_m, int block_n, GemmKParams& p) {
  p.M = g->m;
  p.N = g->n;
  p.kb1 = static_cast<int>(ceil_div(g->k1, GEMM_BK));
  p.kb2 = g->k2 > 0 ? static_cast<int>(ceil_div(g->k2, GEMM_BK)) : 0;
  p.num_m_tiles = static_cast<int>(ceil_div(g->m, block_m));
  p.num_n_tiles = static_cast<int>(ceil_div(g->n, block_n));
  p.out = g->out;
  p.ldo = g->ldo;
  p.out_f32 = g->out_f32;
  p.act = g->act;
  p.bias = g->bias;
  p.residual = reinterpret_cast<const bf16*>(g->residual);
  p.ldr = g->ldr;
  p.aux = reinterpret_cast<bf16*>(g->aux);
  p.ld_aux = g->ld_aux;
  p.alpha = g->alpha;
  p.transpose_out = g->transpose_out != 0 ? 1 : 0;
  p.static_ops = g->static_operands & 3;
  p.ksplit = g->split_k > 1 ? g->split_k : 1;
  const int nkb_total = p.kb1 + p.kb2;
  if (p.ksplit > nkb_total) p.ksplit = nkb_total;
  p.kb_per_split = static_cast<int>(ceil_div(nkb_total, p.ksplit));
  p.ksplit = static_cast<int>(ceil_div(nkb_total, p.kb_per_split));   // no empty k-slices
  p.tail_tiles = 0;
  p.tail_slices = 1;
  p.tail_kb = nkb_total;
  p.sk_partials = nullptr;
  p.sk_epoch = 0;
  p.sk_flags = nullptr;
}

constexpr int64_t SK_FLAG_BYTES = 8192;                       // [SMs][GEMM_EPI_WARPS] u32 flags at the head of the workspace
static int64_t sk_workspace_bytes() { return SK_FLAG_BYTES + static_cast<int64_t>(num_sms()) * 128 * 256 * 4; }

// Tail-split plan for 128-row tiles: with T tiles on G = #SMs CTAs the last wave has R = T mod G tiles; cutting each into
// s = floor(G / R) k-slices (capped: the owner CTA reads s - 1 partial tiles back from L2) keeps R * s CTAs busy for 1/s of a
// tile time instead of R CTAs for a whole one.
static void plan_tail_split(const slam_gemm_args* g, GemmKParams& p, int block_m, int block_n) {
  if (g->workspace == nullptr || g->tail_split < 0 || block_m != 128 || p.ksplit > 1 || g->act >= 3) return;
  if (g->workspace_bytes < sk_workspace_bytes() || (reinterpret_cast<uintptr_t>(g->workspace) & 15) != 0) return;
  const int G = num_sms();
  const int T = p.num_m_tiles * p.num_n_tiles;
  const int nkb = p.kb1 + p.kb2;
  const int R = T % G;
  if (R == 0) return;
  int s = G / R;
  int max_slices = g->tail_split > 1 ? g->tail_split : 4;
  if (max_slices > 4) max_slices = 4;                   // an epilogue warp owns <= 4 chunks; its 24 KB exchange buffer holds 6 blocks
  if (max_slices > block_n / 64) max_slices = block_n / 64;
  if (s > max_slices) s = max_slices;
  if (s > nkb / 4) s = nkb / 4;              // at least 4 k-blocks per slice
  if (s < 2) return;
  p.tail_tiles = R;
  p.tail_kb = static_cast<int>(ceil_div(nkb, s));
  p.tail_slices = static_cast<int>(ceil_div(nkb, p.tail_kb));
  // flag value of this launch: never 0, distinct from every recent launch (the workspace keeps old epochs between calls);
  // seeded from the clock so that a re-loaded library does not repeat the sequence of a previous instance on the same buffer
  static std::atomic<unsigned int> epoch{static_cast<unsigned int>(std::chrono::steady_clock::now().time_since_epoch().count()) | 1u};
  unsigned int e = epoch.fetch_add(1u) + 1u;
  if (e == 0u) e = epoch.fetch_add(1u) + 1u;
  p.sk_epoch = e;
}

template <int BLOCK_M, int BLOCK_N>
static int launch_gemm(const slam_gemm_args* g, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_M, BLOCK_N>;
  constexpr int GEMM_BM = BLOCK_M;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BLOCK_M, BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("gemm: cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  CUtensorMap tmA, tmB, tmA2, tmB2;
  int rc;
  if ((rc = make_tmap(&tmA, g->a, g->m, g->k1, g->lda, GEMM_BM)) != 0) return rc;
  if ((rc = make_tmap(&tmB, g->b, g->n, g->k1, g->ldb, BLOCK_N)) != 0) return rc;
  if (g->k2 > 0) {
    if ((rc = make_tmap(&tmA2, g->a2, g->m, g->k2, g->lda2, GEMM_BM)) != 0) return rc;
    if ((rc = make_tmap(&tmB2, g->b2, g->n, g->k2, g->ldb2, BLOCK_N)) != 0) return rc;
  } else {
    tmA2 = tmA;
    tmB2 = tmB;
  }
  GemmKParams p;
  fill_kparams(g, GEMM_BM, BLOCK_N, p);
  const int tiles = p.num_m_tiles * p.num_n_tiles * p.ksplit;
  int grid = tiles < num_sms() ? tiles : num_sms();
  plan_tail_split(g, p, BLOCK_M, BLOCK_N);
  if (p.tail_tiles > 0) {
    grid = num_sms();
    p.sk_flags = reinterpret_cast<unsigned int*>(g->workspace);
    p.sk_partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(g->workspace) + SK_FLAG_BYTES);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BLOCK_M, BLOCK_N>, tmA, tmB, tmA2, tmB2, p);
  if (le != cudaSuccess) {
    set_error("slam_gemm_bf16: cudaLaunchKernelEx failed: %s", cudaGetErrorString(le));
    return static_cast<int>(le);
  }
  SLAM_LAUNCH_CHECK("slam_gemm_bf16");
  return 0;
}

#ifdef SLAM_GEMM_TRACE
static int g_trace_host_launch = 0;
#endif

template <int BLOCK_N, int HALVES = 1>
static int launch_gemm_pair(const slam_gemm_args* g, cudaStream_t stream) {
  using Cfg = GemmPairCfg<BLOCK_N, HALVES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_pair_kernel<BLOCK_N, HALVES>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("gemm(pair): cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  CUtensorMap tmA, tmB, tmA2, tmB2;
  int rc;
  if ((rc = make_tmap(&tmA, g->a, g->m, g->k1, g->lda, 128 * HALVES)) != 0) return rc;
  if ((rc = make_tmap(&tmB, g->b, g->n, g->k1, g->ldb, BLOCK_N / 2)) != 0) return rc;
  if (g->k2 > 0) {
    if ((rc = make_tmap(&tmA2, g->a2, g->m, g->k2, g->lda2, 128 * HALVES)) != 0) return rc;
    if ((rc = make_tmap(&tmB2, g->b2, g->n, g->k2, g->ldb2, BLOCK_N / 2)) != 0) return rc;
  } else {
    tmA2 = tmA;
    tmB2 = tmB;
  }
  GemmKParams p;
  fill_kparams(g, 256 * HALVES, BLOCK_N, p);
#ifdef SLAM_GEMM_TRACE
  p.trace_id = g_trace_host_launch++;
#endif
  const int items = p.num_m_tiles * p.num_n_tiles * p.ksplit;
  const int max_pairs = num_sms() / 2;
  const int pairs = items < max_pairs ? items : max_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);                // cluster dims (2,1,1) are compiled into the kernel
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_tcgen05_pair_kernel<BLOCK_N, HALVES>, tmA, tmB, tmA2, tmB2, p);
  if (le != cudaSuccess) {
    set_error("slam_gemm_bf16(pair): cudaLaunchKernelEx failed: %s", cudaGetErrorString(le));
    return static_cast<int>(le);
  }
  SLAM_LAUNCH_CHECK("slam_gemm_bf16.pair");
  return 0;
}

// Thin product (N <= 64, one K segment, plain bf16 output): a cluster of THIN_SPLIT CTAs per 128-row tile (gemm_thin.cuh)
static bool thin_cluster_applies(const slam_gemm_args* g) {
  const int64_t nkb = ceil_div(g->k1, GEMM_BK);
  return g->n <= THIN_BN && g->k2 == 0 && !g->out_f32 && g->act == 0 && g->bias == nullptr && g->residual == nullptr && g->split_k <= 1 &&
         g->transpose_out == 0 && nkb >= 2 * THIN_SPLIT && ceil_div(g->m, 128) * THIN_SPLIT <= 2 * num_sms();
}

static int launch_gemm_thin(const slam_gemm_args* g, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_thin_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, THIN_SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("gemm(thin): cudaFuncSetAttribute(smem=%d) failed: %s", THIN_SMEM_BYTES, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tmap(&tmA, g->a, g->m, g->k1, g->lda, 128)) != 0) return rc;
  if ((rc = make_tmap(&tmB, g->b, g->n, g->k1, g->ldb, THIN_BN)) != 0) return rc;
  ThinParams p;
  p.M = g->m;
  p.N = g->n;
  p.nkb = static_cast<int>(ceil_div(g->k1, GEMM_BK));
  p.kb_per_cta = static_cast<int>(ceil_div(p.nkb, THIN_SPLIT));
  p.out = reinterpret_cast<bf16*>(g->out);
  p.ldo = g->ldo;
  p.alpha = g->alpha;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(THIN_SPLIT * ceil_div(g->m, 128)));   // cluster dims (THIN_SPLIT,1,1) are compiled into the kernel
  cfg.blockDim = dim3(THIN_THREADS);
  cfg.dynamicSmemBytes = THIN_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_thin_cluster_kernel, tmA, tmB, p);
  if (le != cudaSuccess) {
    set_error("slam_gemm_bf16(thin): cudaLaunchKernelEx failed: %s", cudaGetErrorString(le));
    return static_cast<int>(le);
  }
  SLAM_LAUNCH_CHECK("slam_gemm_bf16.thin");
  return 0;
}

// Tile choice.  Fitted to SUSTAINED (power-capped, ~1000 W) B200 measurements of the step's shapes, where throughput is set by
// energy per flop as much as by tensor-pipe occupancy (profiles/r01_gemm_power.log).  Every configuration gives each SM 128
// accumulator rows per tile, so a kernel's time is modelled as
//     waves x BLOCK_N x pen,   waves = full waves + cost of the last partial wave (1, or 1/s + 0.2 with the tail split),
// with one measured relative cost per flop `pen` per configuration: CTA pairs (cta_group::2) move 1/3 fewer operand bytes per
// MMA cycle and are 7-14 % cheaper, but they need M in multiples of 256 (a 1604-row activation wastes 10 % of a pair grid).
// The model reproduces the measured ranking on all 14 GEMM shapes of the step.  Returns the slam_gemm_args.block_n code.
struct TileCand {
  int code, rows, bn;
  double pen;
  bool pair;
};
static int pick_tile(int m, int n, int k, bool tail_split, bool allow_pair, bool need_bn128, bool heavy_epilogue) {
  if (n <= 64) return 128 * 1000 + 64;
  static const TileCand cands[] = {
      {128256, 128, 256, 1.000, false}, {128192, 128, 192, 0.977, false}, {128128, 128, 128, 1.170, false},
      {2000256, 256, 256, 0.882, true}, {2000224, 256, 224, 0.864, true}, {2000192, 256, 192, 0.928, true}, {2000160, 256, 160, 1.010, true},
  };
  const int sms = num_sms();
  const int nkb = static_cast<int>(ceil_div(k, GEMM_BK));
  double best = 1e30;
  int best_code = 128256;
  for (const TileCand& c : cands) {
    if (c.pair && !allow_pair) continue;
    if (need_bn128 && c.bn % 128 != 0) continue;                   // SwiGLU forward pairs chunks 64 columns apart
    if (c.bn >= n + 64) continue;                                  // more than two chunks of padding columns
    const int64_t tiles = ceil_div(m, c.rows) * ceil_div(n, c.bn);
    const int units = c.pair ? sms / 2 : sms;
    const int64_t full = tiles / units, rem = tiles % units;
    double tail = rem == 0 ? 0.0 : 1.0;
    if (!c.pair && tail_split && rem > 0) {
      int64_t sl = units / rem;
      if (sl > 4) sl = 4;
      if (sl > c.bn / 64) sl = c.bn / 64;
      if (sl > nkb / 4) sl = nkb / 4;
      if (sl >= 2) tail = 1.0 / static_cast<double>(sl) + 0.2;
    }
    // the SwiGLU-backward epilogue (two bf16 reads, two exp, two stores per output) outlasts the MMA of a 128 x 256 tile at
    // K = 4096 on one SM; the pair kernels hide it (tools/swiglu_probe.py: 206 us vs 187 us on the d_down shape)
    const double cost = (static_cast<double>(full) + tail) * c.bn * c.pen * (heavy_epilogue && !c.pair ? 1.15 : 1.0);
    if (cost < best) {
      best = cost;
      best_code = c.code;
    }
  }
  return best_code;
}

}  // namespace slam

extern "C" int64_t slam_gemm_workspace_bytes(void) { return slam::sk_workspace_bytes(); }

#ifdef SLAM_GEMM_TRACE
// debug build: buf = device memory of launches x TRACE_CTAS x TRACE_EV x 2 u64 (zeroed by the caller), or NULL to stop; returns the pair launches since the previous call
extern "C" int slam_debug_gemm_trace(void* buf, int launches) {
  const int n = slam::g_trace_host_launch;
  unsigned long long* b = static_cast<unsigned long long*>(buf);
  cudaMemcpyToSymbol(slam::g_trace_buf, &b, sizeof(b));
  cudaMemcpyToSymbol(slam::g_trace_launches, &launches, sizeof(launches));
  slam::g_trace_host_launch = 0;
  return n;
}
#endif

extern "C" int slam_gemm_bf16(const slam_gemm_args* g, void* stream) {
  using namespace slam;
  SLAM_CHECK_ARG(g != nullptr, "gemm: null args");
  SLAM_CHECK_ARG(g->m > 0 && g->n > 0 && g->k1 > 0, "gemm: bad shape m=%d n=%d k1=%d", g->m, g->n, g->k1);
  // 16-byte pieces run along n (normal orientation) or along m (transpose_out: out / residual are [n][m])
  SLAM_CHECK_ARG((g->transpose_out != 0 ? g->m : g->n) % 8 == 0, "gemm: %s=%d must be a multiple of 8", g->transpose_out != 0 ? "m" : "n",
                 g->transpose_out != 0 ? g->m : g->n);
  SLAM_CHECK_ARG(g->ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(g->out) & 15) == 0, "gemm: output must be 16-byte aligned with ldo %% 8 == 0");
  SLAM_CHECK_ARG(g->residual == nullptr || (g->ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(g->residual) & 15) == 0),
                 "gemm: residual must be 16-byte aligned with ldr %% 8 == 0");
  SLAM_CHECK_ARG(g->bias == nullptr || (reinterpret_cast<uintptr_t>(g->bias) & 15) == 0, "gemm: bias must be 16-byte aligned");
  SLAM_CHECK_ARG(g->k2 == 0 || (g->a2 != nullptr && g->b2 != nullptr), "gemm: k2 > 0 needs a2/b2");
  SLAM_CHECK_ARG(g->split_k <= 1 || (g->out_f32 && g->bias == nullptr && g->residual == nullptr && g->act == 0),
                 "gemm: split_k needs a zero-initialised f32 output and no bias/activation/residual");
  SLAM_CHECK_ARG(g->act >= 0 && g->act <= 4, "gemm: unknown act %d", g->act);
  if (g->act >= 3) {
    SLAM_CHECK_ARG(g->aux != nullptr && (reinterpret_cast<uintptr_t>(g->aux) & 15) == 0 && g->ld_aux % 8 == 0 && !g->out_f32 &&
                       g->bias == nullptr && g->residual == nullptr && g->split_k <= 1,
                   "gemm: fused SwiGLU needs a 16-byte aligned aux, a bf16 output and no bias/residual/split_k");
    const int feat = g->transpose_out != 0 ? g->m : g->n;          // the feature dimension of the product (swap-AB: the rows)
    SLAM_CHECK_ARG(g->act == 3 ? feat % 128 == 0 : feat % 64 == 0, "gemm: fused SwiGLU needs whole blocked-64 groups (features=%d)", feat);
  }
  SLAM_CHECK_ARG(g->transpose_out == 0 || (!g->out_f32 && (g->act == 0 || g->act == 4) && g->bias == nullptr && g->split_k <= 1),
                 "gemm: transpose_out needs a bf16 output, act 0 or 4 (SwiGLU backward) and no bias / split_k");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int tile = g->block_n;   // 0 = auto; BLOCK_N alone (64/128/192/256) = 128-row tile; BLOCK_M*1000+BLOCK_N = explicit
  if (tile == 3000064) {   // thin cluster kernel, explicitly
    SLAM_CHECK_ARG(thin_cluster_applies(g), "gemm: tile 3000064 (thin cluster kernel) needs n <= 64, one K segment of >= %d k-blocks, a plain bf16 output",
                   2 * THIN_SPLIT);
    return launch_gemm_thin(g, st);
  }
  if (tile == 0 && thin_cluster_applies(g)) return launch_gemm_thin(g, st);
  if (g->transpose_out != 0 && g->act == 4) {     // swap-AB SwiGLU backward lives in the 256-row CTA-pair kernels only
    if (tile == 0) tile = 2000192;
    SLAM_CHECK_ARG(tile >= 2000000 && tile < 3000000, "gemm: transpose_out with act 4 needs a CTA-pair tile (2000000 + BLOCK_N), got %d", tile);
  }
  if (tile == 0) tile = pick_tile(g->m, g->n, g->k1 + g->k2, g->workspace != nullptr && g->tail_split >= 0 && g->split_k <= 1 && g->act < 3, g->split_k <= 1, g->act == 3, g->act == 4);
  SLAM_CHECK_ARG(g->act != 3 || (tile % 1000) % 128 == 0, "gemm: SwiGLU forward needs a tile of 128 or 256 columns (tile %d)", tile);
  if (tile < 1000) tile += 128 * 1000;   // (2000000 + BLOCK_N = CTA-pair kernel)
  switch (tile) {
    case 128256: return launch_gemm<128, 256>(g, st);
    case 128192: return launch_gemm<128, 192>(g, st);
    case 128128: return launch_gemm<128, 128>(g, st);
    case 128064: return launch_gemm<128, 64>(g, st);
    case 256256: return launch_gemm<256, 256>(g, st);
    case 256224: return launch_gemm<256, 224>(g, st);
    case 2000256: return launch_gemm_pair<256>(g, st);   // CTA-pair kernels (cta_group::2), pair tile 256 x BLOCK_N
    case 2000224: return launch_gemm_pair<224>(g, st);
    case 2000192: return launch_gemm_pair<192>(g, st);
    case 2000160: return launch_gemm_pair<160>(g, st);
    case 2000128: return launch_gemm_pair<128>(g, st);
    case 4000192:            // "pair512": 512 x 192 per SM pair, one accumulator set (plain / residual epilogue only)
      SLAM_CHECK_ARG(g->act == 0 && g->split_k <= 1, "gemm: tile 4000192 supports act 0 without split_k only");
      return launch_gemm_pair<192, 2>(g, st);
    default: set_error("gemm: unsupported tile %d", tile); return -1;
  }
}
