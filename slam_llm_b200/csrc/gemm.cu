// tcgen05 / TMEM / TMA GEMM core for the SLAM-LLM training step on B200 (sm_100a).
//
//   out[M,N] = act(alpha * (A·B^T + A2·B2^T) + bias) + residual
//
// Design (one CTA per SM, persistent over output tiles, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D loads of 128xBK (A) and BNxBK (B) bf16 tiles
//               into a STAGES-deep 128B-swizzled shared-memory ring, completion on mbarriers;
//   warp 1      MMA issuer: one lane issues tcgen05.mma (M=128, N=BLOCK_N, K=16) with both operands
//               read from shared memory through UMMA descriptors, fp32 accumulator in TMEM;
//               tcgen05.commit releases ring slots / publishes the accumulator;
//   warp 2      TMEM allocator (BLOCK_M=128: 2 accumulator stages so the epilogue of tile i overlaps tile i+1;
//               BLOCK_M=256: two M=128 accumulators that SHARE every B tile — the mainloop is bound by the bytes one SM can
//               keep in flight from L2 (~60 B/clk/SM measured), so 256x256 tiles need 1/3 less traffic per MMA cycle);
//   warps 4-7   epilogue: tcgen05.ld 32x32b -> registers -> alpha/bias/activation/residual ->
//               16-byte global stores (bf16 or f32).
// The K loop runs over TWO operand pairs back to back ("dual K segment"): the base weights and the
// rank-padded LoRA pair, so y = xW^T + (alpha/r)(xA^T)B^T is produced in ONE accumulator tile
// (reference: peft lora.Linear.forward called under models/slam_model.py:400).
#include <mutex>

#include "../../include/slam_b200.h"
#include "common.cuh"
#include "host.cuh"

namespace slam {

constexpr int GEMM_BK = 64;
#ifndef SLAM_GEMM_PREFETCH
#define SLAM_GEMM_PREFETCH 0   // L2 prefetch distance of the weight operand in k-blocks; measured on B200: 4/8/16 are ~10 % SLOWER than 0 (profiles/r01_exp_prefetch.log)
#endif
constexpr int GEMM_THREADS = 256;

struct GemmKParams {
  int M, N;
  int kb1, kb2;
  int ksplit, kb_per_split;   // split-K: work item = (tile, k-slice); partial tiles are merged with fp32 atomics
  int num_m_tiles, num_n_tiles;
  void* out;
  long long ldo;
  int out_f32;
  int act;
  const float* bias;
  const bf16* residual;
  long long ldr;
  float alpha;
};

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
  static constexpr int EPI_PITCH = 36;                                    // floats per staged row (32 + 4 pad: conflict-free)
  static constexpr int EPI_BYTES = 4 * 32 * EPI_PITCH * 4;                 // one 32x32 fp32 chunk per epilogue warp
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES) + Cfg::BAR_BYTES);

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
      mbar_init(&tempty_bar[s], 4);
    }
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

  const int total_tiles = p.num_m_tiles * p.num_n_tiles * p.ksplit;   // work items
  const int nkb = p.kb1 + p.kb2;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // The B operand (weights) streams from HBM and every M-tile of an N-tile asks for the same B tile at the same time, so
    // without help each load sees the full HBM latency and the mainloop is bound by bytes-in-flight / latency (measured:
    // halving the stages cuts throughput ~40 %).  The producer therefore prefetches B tiles into L2 PF k-blocks ahead
    // (and the head of its next tile), which costs no shared memory.
    constexpr int PF = SLAM_GEMM_PREFETCH;
    auto prefetch_b = [&](int n_tile_pf, int kb_pf) {
      if (kb_pf < p.kb1) tma_prefetch_l2_2d(&tmB, kb_pf * GEMM_BK, n_tile_pf * BLOCK_N);
      else tma_prefetch_l2_2d(&tmB2, (kb_pf - p.kb1) * GEMM_BK, n_tile_pf * BLOCK_N);
    };
    uint32_t kc = 0;
    if (PF > 0 && p.ksplit == 1 && lane == 0 && static_cast<int>(blockIdx.x) < total_tiles) {
      const int n_first = static_cast<int>(blockIdx.x) / p.num_m_tiles;
      for (int kb = 0; kb < PF && kb < nkb; ++kb) prefetch_b(n_first, kb);
    }
    for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
      const int tile = item / p.ksplit;
      const int kb_begin = (item % p.ksplit) * p.kb_per_split;
      const int kb_end = min(nkb, kb_begin + p.kb_per_split);
      const int m_tile = tile % p.num_m_tiles;
      const int n_tile = tile / p.num_m_tiles;
      const int next_tile = (PF > 0 && p.ksplit == 1) ? item + static_cast<int>(gridDim.x) : total_tiles;
      for (int kb = kb_begin; kb < kb_end; ++kb, ++kc) {
        const uint32_t stage = kc % STAGES;
        const uint32_t ph = (kc / STAGES) & 1u;
        mbar_wait(&empty_bar[stage], ph ^ 1u);
        if (lane == 0) {
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES);
          void* dA = sA + stage * Cfg::A_BYTES;
          void* dB = sB + stage * Cfg::B_BYTES;
          if (kb < p.kb1) {
            tma_load_2d(dA, &tmA, &full_bar[stage], kb * GEMM_BK, m_tile * BLOCK_M);
            tma_load_2d(dB, &tmB, &full_bar[stage], kb * GEMM_BK, n_tile * BLOCK_N);
          } else {
            const int k2 = kb - p.kb1;
            tma_load_2d(dA, &tmA2, &full_bar[stage], k2 * GEMM_BK, m_tile * BLOCK_M);
            tma_load_2d(dB, &tmB2, &full_bar[stage], k2 * GEMM_BK, n_tile * BLOCK_N);
          }
          if (PF > 0 && p.ksplit == 1) {
            if (kb + PF < nkb) prefetch_b(n_tile, kb + PF);
            else if (next_tile < total_tiles) prefetch_b(next_tile / p.num_m_tiles, kb + PF - nkb);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N);
    uint32_t kc = 0;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < total_tiles; item += gridDim.x, ++it) {
      const int kb_begin = (item % p.ksplit) * p.kb_per_split;
      const int kb_end = min(nkb, kb_begin + p.kb_per_split);
      const uint32_t acc = it % ACC_STAGES;
      const uint32_t aph = (it / ACC_STAGES) & 1u;
      mbar_wait(&tempty_bar[acc], aph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * HALVES * BLOCK_N;
      for (int kb = kb_begin; kb < kb_end; ++kb, ++kc) {
        const uint32_t stage = kc % STAGES;
        const uint32_t ph = (kc / STAGES) & 1u;
        mbar_wait(&full_bar[stage], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t a_desc = make_sw128_kmajor_desc(smem_u32(sA + stage * Cfg::A_BYTES));
          const uint64_t b_desc = make_sw128_kmajor_desc(smem_u32(sB + stage * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 bf16 (32 B) along K inside the 128-B swizzle span: +2 in 16-B units;
            // the second M=128 half of a 256-row A tile starts 128 rows * 128 B = 16 KB further (+1024 units)
#pragma unroll
            for (int hf = 0; hf < HALVES; ++hf)
              umma_bf16(d_tmem + hf * BLOCK_N, a_desc + 1024u * hf + 2u * k, b_desc + 2u * k, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == kb_end - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int q = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may read
    uint32_t it = 0;
    for (int item = blockIdx.x; item < total_tiles; item += gridDim.x, ++it) {
      const int tile = item / p.ksplit;
      const int m_tile = tile % p.num_m_tiles;
      const int n_tile = tile / p.num_m_tiles;
      const uint32_t acc = it % ACC_STAGES;
      const uint32_t aph = (it / ACC_STAGES) & 1u;
      mbar_wait(&tfull_bar[acc], aph);
      tc_fence_after();
      const int n0 = n_tile * BLOCK_N;
      // Each chunk of 32 accumulator columns goes TMEM -> registers (thread = row) -> a padded shared-memory tile ->
      // registers again with (8 rows x 4 column-pieces) per warp instruction, so that global stores and residual loads
      // touch whole 32-byte sectors (64 B of bf16 per row) instead of one 16-byte piece of 32 different rows.
      float* stg = epi_stage + q * (32 * Cfg::EPI_PITCH);
      const int piece = lane & 3;
#pragma unroll 1
      for (int hf = 0; hf < HALVES; ++hf) {
        const int row_base = m_tile * BLOCK_M + hf * 128 + q * 32;
        const uint32_t taddr = tmem_base + (acc * HALVES + hf) * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (hf == HALVES - 1 && c == BLOCK_N / 32 - 1) {
            // accumulator fully drained into registers: hand the TMEM stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stg + lane * Cfg::EPI_PITCH + j * 4) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          __syncwarp();
          const int col = n0 + c * 32 + piece * 8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + (lane >> 2);
            const int row = row_base + rr;
            if (row >= p.M || col >= p.N) continue;
            const float4 x0 = *reinterpret_cast<const float4*>(stg + rr * Cfg::EPI_PITCH + piece * 8);
            const float4 x1 = *reinterpret_cast<const float4*>(stg + rr * Cfg::EPI_PITCH + piece * 8 + 4);
            float v[8] = {x0.x * p.alpha, x0.y * p.alpha, x0.z * p.alpha, x0.w * p.alpha,
                          x1.x * p.alpha, x1.y * p.alpha, x1.z * p.alpha, x1.w * p.alpha};
            if (p.bias != nullptr) {
              const float4 b0 = *reinterpret_cast<const float4*>(p.bias + col);
              const float4 b1 = *reinterpret_cast<const float4*>(p.bias + col + 4);
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (p.act == 1) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = gelu_erf(v[e]);
            } else if (p.act == 2) {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.0f);
            }
            if (p.residual != nullptr) {
              const uint4 rsd = *reinterpret_cast<const uint4*>(p.residual + static_cast<long long>(row) * p.ldr + col);
              const float2 r0 = unpack_bf16x2(rsd.x), r1 = unpack_bf16x2(rsd.y), r2 = unpack_bf16x2(rsd.z),
                           r3 = unpack_bf16x2(rsd.w);
              v[0] += r0.x; v[1] += r0.y; v[2] += r1.x; v[3] += r1.y;
              v[4] += r2.x; v[5] += r2.y; v[6] += r3.x; v[7] += r3.y;
            }
            if (p.ksplit > 1) {
              float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col;
#pragma unroll
              for (int e = 0; e < 8; ++e) atomicAdd(o + e, v[e]);      // k-slices merge into the zero-initialised fp32 output
            } else if (p.out_f32) {
              float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col;
              *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
              *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
            } else {
              bf16* o = reinterpret_cast<bf16*>(p.out) + static_cast<long long>(row) * p.ldo + col;
              uint4 pk;
              pk.x = pack_bf16x2(v[0], v[1]);
              pk.y = pack_bf16x2(v[2], v[3]);
              pk.z = pack_bf16x2(v[4], v[5]);
              pk.w = pack_bf16x2(v[6], v[7]);
              *reinterpret_cast<uint4*>(o) = pk;
            }
          }
          __syncwarp();
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
  p.M = g->m;
  p.N = g->n;
  p.kb1 = static_cast<int>(ceil_div(g->k1, GEMM_BK));
  p.kb2 = g->k2 > 0 ? static_cast<int>(ceil_div(g->k2, GEMM_BK)) : 0;
  p.num_m_tiles = static_cast<int>(ceil_div(g->m, GEMM_BM));
  p.num_n_tiles = static_cast<int>(ceil_div(g->n, BLOCK_N));
  p.out = g->out;
  p.ldo = g->ldo;
  p.out_f32 = g->out_f32;
  p.act = g->act;
  p.bias = g->bias;
  p.residual = reinterpret_cast<const bf16*>(g->residual);
  p.ldr = g->ldr;
  p.alpha = g->alpha;
  p.ksplit = g->split_k > 1 ? g->split_k : 1;
  const int nkb_total = p.kb1 + p.kb2;
  if (p.ksplit > nkb_total) p.ksplit = nkb_total;
  p.kb_per_split = static_cast<int>(ceil_div(nkb_total, p.ksplit));
  p.ksplit = static_cast<int>(ceil_div(nkb_total, p.kb_per_split));   // no empty k-slices
  const int tiles = p.num_m_tiles * p.num_n_tiles * p.ksplit;
  const int grid = tiles < num_sms() ? tiles : num_sms();
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

// Tile choice fitted to B200 measurements (profiles/r01_gemm_bench_v2.json).  Returns BLOCK_M * 1000 + BLOCK_N.
//  * 128-row tiles have two TMEM accumulator stages (epilogue overlaps the next tile): pick 256 vs 192 columns by wave
//    quantisation, cost = ceil(tiles / SMs) * BLOCK_N * penalty (192 re-reads A slightly more often; 128 is smem-bound).
//  * 256x256 tiles (two M=128 accumulators sharing each B tile) need 1/3 less L2->SM traffic per MMA cycle but expose their
//    epilogue once per tile: they win only when the whole GEMM is a single wave and K is long enough to amortise it
//    (down_proj, the gate/up dgrad, the encoder fc2: +8..12 %); with several waves they lose 10-20 %.
static int pick_tile(int m, int n, int k) {
  if (n <= 64) return 128 * 1000 + 64;
  if (n < 192) return 128 * 1000 + 128;
  const int sms = num_sms();
  if (k >= 5000 && n >= 256) {   // one well-filled wave of 256-row tiles: prefer the width that uses the most SMs
    const int64_t t224 = ceil_div(m, 256) * ceil_div(n, 224), t256 = ceil_div(m, 256) * ceil_div(n, 256);
    if (t224 <= sms && t224 > t256 && t224 * 10 >= sms * 6) return 256 * 1000 + 224;
    if (t256 <= sms && t256 * 10 >= sms * 6) return 256 * 1000 + 256;
  }
  const int64_t mt = ceil_div(m, 128);
  const int cands[3] = {256, 192, 128};
  const double pen[3] = {1.0, 1.04, 1.5};
  double best = 1e30;
  int best_bn = 256;
  for (int i = 0; i < 3; ++i) {
    if (cands[i] > n && i < 2) continue;
    const double c = static_cast<double>(ceil_div(mt * ceil_div(n, cands[i]), sms)) * cands[i] * pen[i];
    if (c < best) {
      best = c;
      best_bn = cands[i];
    }
  }
  return 128 * 1000 + best_bn;
}

}  // namespace slam

extern "C" int slam_gemm_bf16(const slam_gemm_args* g, void* stream) {
  using namespace slam;
  SLAM_CHECK_ARG(g != nullptr, "gemm: null args");
  SLAM_CHECK_ARG(g->m > 0 && g->n > 0 && g->k1 > 0, "gemm: bad shape m=%d n=%d k1=%d", g->m, g->n, g->k1);
  SLAM_CHECK_ARG(g->n % 8 == 0, "gemm: n=%d must be a multiple of 8", g->n);
  SLAM_CHECK_ARG(g->ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(g->out) & 15) == 0,
                 "gemm: output must be 16-byte aligned with ldo %% 8 == 0");
  SLAM_CHECK_ARG(g->residual == nullptr || (g->ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(g->residual) & 15) == 0),
                 "gemm: residual must be 16-byte aligned with ldr %% 8 == 0");
  SLAM_CHECK_ARG(g->bias == nullptr || (reinterpret_cast<uintptr_t>(g->bias) & 15) == 0, "gemm: bias must be 16-byte aligned");
  SLAM_CHECK_ARG(g->k2 == 0 || (g->a2 != nullptr && g->b2 != nullptr), "gemm: k2 > 0 needs a2/b2");
  SLAM_CHECK_ARG(g->split_k <= 1 || (g->out_f32 && g->bias == nullptr && g->residual == nullptr && g->act == 0),
                 "gemm: split_k needs a zero-initialised f32 output and no bias/activation/residual");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int tile = g->block_n;   // 0 = auto; BLOCK_N alone (64/128/192/256) = 128-row tile; BLOCK_M*1000+BLOCK_N = explicit
  if (tile == 0) tile = pick_tile(g->m, g->n, g->k1 + g->k2);
  if (tile < 1000) tile += 128 * 1000;
  switch (tile) {
    case 128256: return launch_gemm<128, 256>(g, st);
    case 128192: return launch_gemm<128, 192>(g, st);
    case 128128: return launch_gemm<128, 128>(g, st);
    case 128064: return launch_gemm<128, 64>(g, st);
    case 256256: return launch_gemm<256, 256>(g, st);
    case 256224: return launch_gemm<256, 224>(g, st);
    case 256192: return launch_gemm<256, 192>(g, st);
    case 256128: return launch_gemm<256, 128>(g, st);
    default: set_error("gemm: unsupported tile %d", tile); return -1;
  }
}
