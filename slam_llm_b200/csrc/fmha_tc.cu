// tcgen05 / TMEM flash-attention forward: the Whisper encoder shape (dh = 64, non-causal, unmasked; whisper qkv_attention via
// models/encoder.py:26-27) and the Llama decoder shape (dh = 64/128, causal + key-padding mask, GQA, lse
// for the backward; HF LlamaAttention).  The mma.sync kernel in attention.cu remains for other shapes.
//
// One CTA = 128 query rows of one (batch, head); the key/value sequence is walked in tiles of 128 keys.
//   warp 0      TMA producer: Q tile once, then K (2-stage ring) and V tiles (128 x 64 bf16, 128B swizzle)
//   warp 1      MMA issuer (whole warp):  S_j = Q K_j^T  (M=128, N=128, K=64)  into one of two TMEM score buffers,
//                                       O  += P_j V_j  (M=128, N=64,  K=128) with P_j read from shared memory (K-major) and
//                                       V_j used in place as an MN-major operand (no transposed copy of V)
//               (warp 1 also allocates the CTA's 256 TMEM columns: 128 score + 64 output used; two CTAs share an SM)
//   warps 2-5   softmax: thread = query row; tcgen05.ld of the score row, online max / exp2 / sum in fp32, bf16 P written to
//               shared memory in the UMMA 128B-swizzled K-major layout, O rescaled in TMEM (tcgen05.ld/st) when the running
//               max moves; final O / l written to global memory.
// S_{j+1} is issued before P_j V_j, so the tensor pipe computes the next scores while the softmax warps work on the current tile.
#include <math_constants.h>

#include "../../include/slam_b200.h"
#include "common.cuh"
#include "host.cuh"

namespace slam {

constexpr int FA_BM = 128;   // query rows per CTA
constexpr int FA_BN = 128;   // keys per tile
constexpr int FA_THREADS = 192;                       // warp 0: TMA (+ key-mask bits), warp 1: TMEM alloc + MMA, warps 2-5: softmax
constexpr int FA_TILE_BYTES = 128 * 64 * 2;          // one 128 x 64 bf16 block (dh = 128 tiles are two such blocks)
constexpr int FA_KST = 2;                            // K ring stages
constexpr float FA_RESCALE_TH = 8.0f;                // lazy rescale: keep a stale max while it is within 2^8 of the true one

template <int DH>
struct FaCfg {
  static constexpr int NB = DH / 64;                 // 64-wide blocks per head vector
  // dh = 64: Q + K ring + V + P = 96 KB -> TWO CTAs per SM (one CTA's MUFU-bound softmax overlaps the other's waits);
  // dh = 128: 160 KB, one CTA per SM.  TMEM: 128 score columns + DH output columns (256 allocated).
  static constexpr int SMEM = FA_TILE_BYTES * (NB * (1 + FA_KST + 1) + 2) + 512 + 1024;
  static constexpr int CTAS_PER_SM = DH == 64 ? 2 : 1;
};

struct FmhaParams {
  int sq, sk, hq, hkv;
  int causal;
  float scale_log2;
  bf16* out;
  long long ldo;
  float* lse;                 // [B, Hq, Sq] natural-log lse, or NULL
  const uint8_t* key_mask;    // [B, Sk] (1 = attend) or NULL
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// D=f32, A=bf16 K-major, B=bf16 with selectable major-ness
__host__ __device__ constexpr uint32_t fa_idesc(uint32_t M, uint32_t N, uint32_t b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// MASKED = causal and/or key-padding mask and/or lse output (the Llama decoder); otherwise the Whisper encoder shape.
template <int DH, bool MASKED>
__global__ void __launch_bounds__(FA_THREADS, FaCfg<DH>::CTAS_PER_SM)
fmha_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const FmhaParams p) {
  constexpr int NB = FaCfg<DH>::NB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                                 // NB blocks
  uint8_t* sK = sQ + NB * FA_TILE_BYTES;              // FA_KST stages x NB blocks
  uint8_t* sV = sK + FA_KST * NB * FA_TILE_BYTES;     // NB blocks (block nb = head dims [64 nb, 64 nb + 64))
  uint8_t* sP = sV + NB * FA_TILE_BYTES;              // 2 K-blocks of 64 keys
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * FA_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [FA_KST]
  uint64_t* k_empty = bars + 3;   // [FA_KST]
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 6;
  uint64_t* s_full = bars + 7;
  uint64_t* s_empty = bars + 8;
  uint64_t* p_full = bars + 9;
  uint64_t* pv_done = bars + 10;  // P buffer free again / O stable after P_j V_j
  uint64_t* m_full = bars + 11;   // [FA_KST] key-mask bits of the tile are in shared memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  uint32_t* s_maskbits = tmem_slot + 4;               // [FA_KST][4]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid = (heads, batch, query tiles), dispatched x-fastest: with a causal mask the LAST query tile meets the most key tiles,
  // so the tile index is reversed and the heaviest CTAs start first (LPT; the decoder grid is 3.5 uneven waves)
  const int h = blockIdx.x, b = blockIdx.y;
  const int hk = h / (p.hq / p.hkv);
  const int q_tile = (MASKED && p.causal) ? static_cast<int>(gridDim.z - 1 - blockIdx.z) : static_cast<int>(blockIdx.z);
  const int q0 = q_tile * FA_BM;
  int n_tiles = (p.sk + FA_BN - 1) / FA_BN;
  if (MASKED && p.causal) n_tiles = min(n_tiles, (q0 + FA_BM + FA_BN - 1) / FA_BN);   // tiles above the diagonal never contribute
  const bool has_mask = MASKED && p.key_mask != nullptr;

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < FA_KST; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&m_full[s], 1);
    }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                                        // q / k / v come from the preceding GEMM: wait for it before the first load
  const uint32_t tmem_base = *tmem_slot;             // score tile: columns [0, 128)
  const uint32_t tmem_o = tmem_base + 128;           // output accumulator: columns [128, 128 + DH)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (+ key-mask bits)
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, NB * FA_TILE_BYTES);
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) tma_load_2d(sQ + nb * FA_TILE_BYTES, &tmQ, q_full, h * DH + nb * 64, b * p.sq + q0);
    }
    for (int j = 0; j < n_tiles; ++j) {
      const int ks = j % FA_KST;
      mbar_wait(&k_empty[ks], ((j / FA_KST) & 1u) ^ 1u);     // (the softmax warps of tile j - FA_KST are long past their mask bits)
      if (lane == 0) {
        mbar_arrive_expect_tx(&k_full[ks], NB * FA_TILE_BYTES);
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
          tma_load_2d(sK + (ks * NB + nb) * FA_TILE_BYTES, &tmK, &k_full[ks], hk * DH + nb * 64, b * p.sk + j * FA_BN);
      }
      if (has_mask) {
        // 128 key-mask bytes of this tile -> 4 words of bits (lane l covers keys 4l .. 4l+3)
        uint32_t nib = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = j * FA_BN + 4 * lane + e;
          if (col < p.sk && p.key_mask[static_cast<long long>(b) * p.sk + col] != 0) nib |= 1u << e;
        }
        uint32_t v = nib << (4 * (lane & 7));
        v |= __shfl_xor_sync(0xffffffffu, v, 1);
        v |= __shfl_xor_sync(0xffffffffu, v, 2);
        v |= __shfl_xor_sync(0xffffffffu, v, 4);
        if ((lane & 7) == 0) s_maskbits[ks * 4 + (lane >> 3)] = v;
        __syncwarp();
        if (lane == 0) mbar_arrive(&m_full[ks]);
      }
      mbar_wait(v_empty, (j & 1u) ^ 1u);
      if (lane == 0) {
        mbar_arrive_expect_tx(v_full, NB * FA_TILE_BYTES);
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) tma_load_2d(sV + nb * FA_TILE_BYTES, &tmV, v_full, hk * DH + nb * 64, b * p.sk + j * FA_BN);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, warp-uniform operands: common.cuh)
    constexpr uint32_t idesc_s = fa_idesc(128, FA_BN, 0);      // S = Q K^T : both operands K-major
    constexpr uint32_t idesc_o = fa_idesc(128, 64, 1);         // O[:, 64 nb ..] += P V[:, 64 nb ..] : V is MN-major
    constexpr uint32_t BLK = FA_TILE_BYTES >> 4;               // descriptor units between consecutive 64-wide blocks
    const uint32_t lo_q = sw128_kmajor_desc_lo(smem_u32(sQ)), lo_k0 = sw128_kmajor_desc_lo(smem_u32(sK));
    const uint32_t lo_v = sw128_kmajor_desc_lo(smem_u32(sV)), lo_p = sw128_kmajor_desc_lo(smem_u32(sP));
    auto issue_s = [&](int j) {
      const int ks = j % FA_KST;
      mbar_wait(&k_full[ks], (j / FA_KST) & 1u);
      mbar_wait(s_empty, (j & 1u) ^ 1u);                       // softmax of tile j-1 has the scores in registers
      tc_fence_after();
#pragma unroll
      for (uint32_t kb = 0; kb < static_cast<uint32_t>(NB); ++kb)   // K = dh in 64-wide blocks
        umma_kblock_1(tmem_base, lo_q + kb * BLK, lo_k0 + (ks * NB + kb) * BLK, idesc_s, kb);
      umma_commit_elect(smem_u32(&k_empty[ks]));
      umma_commit_elect(smem_u32(s_full));
    };
    mbar_wait(q_full, 0);
    issue_s(0);
    for (int j = 0; j < n_tiles; ++j) {
      if (j + 1 < n_tiles) issue_s(j + 1);                     // next scores overlap this tile's softmax
      mbar_wait(p_full, j & 1u);
      mbar_wait(v_full, j & 1u);
      tc_fence_after();
      // K = 128 keys in two 64-key blocks of P (K-major, 16 KB apart); V block nb: 16 keys = two 8-row groups = 2048 B per step
#pragma unroll
      for (uint32_t kb = 0; kb < 2; ++kb)
#pragma unroll
        for (uint32_t nb = 0; nb < static_cast<uint32_t>(NB); ++nb)
          umma_kblock_mnb(tmem_o + nb * 64, lo_p + kb * BLK, lo_v + nb * BLK + kb * 512u, idesc_o, (j > 0 || kb > 0) ? 1u : 0u);
      umma_commit_elect(smem_u32(v_empty));
      umma_commit_elect(smem_u32(pv_done));
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue (thread = query row)
    const int qd = warp & 3;                                   // TMEM lane quarter this warp may access
    const int row = qd * 32 + lane;
    const int grow = q0 + row;
    const uint32_t lane_base = static_cast<uint32_t>(qd * 32) << 16;
    float m_used = -CUDART_INF_F, l_run = 0.0f;                // exponent offset in use (scaled log2 domain), running sum
    for (int j = 0; j < n_tiles; ++j) {
      const int n0 = j * FA_BN;
      mbar_wait(s_full, j & 1u);
      tc_fence_after();
      const uint32_t ts = tmem_base + lane_base;
      // the whole 128-wide score row lives in registers: one TMEM round trip per tile, and the score buffer is handed back
      // to the MMA warp (which then computes S_{j+1}) before any math
      uint32_t s0[32], s1[32], s2[32], s3[32];
      tmem_ld_32x32(ts, s0);
      tmem_ld_32x32(ts + 32, s1);
      tmem_ld_32x32(ts + 64, s2);
      tmem_ld_32x32(ts + 96, s3);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);
      const uint32_t ninf = __float_as_uint(-CUDART_INF_F);
      if (n0 + FA_BN > p.sk) {                                 // key tail: columns >= sk never contribute
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          if (n0 + e >= p.sk) s0[e] = ninf;
          if (n0 + 32 + e >= p.sk) s1[e] = ninf;
          if (n0 + 64 + e >= p.sk) s2[e] = ninf;
          if (n0 + 96 + e >= p.sk) s3[e] = ninf;
        }
      }
      if (MASKED) {
        if (p.causal && n0 + FA_BN - 1 > q0) {                 // diagonal tile: keys after the query are masked
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (n0 + e > grow) s0[e] = ninf;
            if (n0 + 32 + e > grow) s1[e] = ninf;
            if (n0 + 64 + e > grow) s2[e] = ninf;
            if (n0 + 96 + e > grow) s3[e] = ninf;
          }
        }
        if (has_mask) {
          const int ks = j % FA_KST;
          mbar_wait(&m_full[ks], (j / FA_KST) & 1u);
          const uint32_t b0 = s_maskbits[ks * 4], b1 = s_maskbits[ks * 4 + 1], b2 = s_maskbits[ks * 4 + 2], b3 = s_maskbits[ks * 4 + 3];
          if ((b0 & b1 & b2 & b3) != 0xffffffffu) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              if (!((b0 >> e) & 1u)) s0[e] = ninf;
              if (!((b1 >> e) & 1u)) s1[e] = ninf;
              if (!((b2 >> e) & 1u)) s2[e] = ninf;
              if (!((b3 >> e) & 1u)) s3[e] = ninf;
            }
          }
        }
      }
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int e = 0; e < 32; ++e)
        mx = fmaxf(mx, fmaxf(fmaxf(__uint_as_float(s0[e]), __uint_as_float(s1[e])), fmaxf(__uint_as_float(s2[e]), __uint_as_float(s3[e]))));
      const float m_tile = mx * p.scale_log2;
      // lazy rescale (as in FlashAttention-4): exponents are taken relative to m_used, which is only raised when the true
      // maximum runs more than 2^8 ahead (p <= 256 stays exact enough in bf16 / fp32) or on the first unmasked tile
      const bool need = m_tile > m_used + FA_RESCALE_TH;
      const bool any_need = __any_sync(0xffffffffu, need) != 0;
      float corr = 1.0f;
      if (any_need) {
        const float m_new = fmaxf(m_used, m_tile);
        corr = m_new == -CUDART_INF_F ? 1.0f : ex2_approx(m_used - m_new);   // 0 when the row sees its first unmasked key
        m_used = m_new;
      }
      const float neg_m = m_used == -CUDART_INF_F ? 0.0f : -m_used;          // fully masked so far: exp2(-inf - 0) = 0
      // the single P buffer is free again once P_{j-1} V_{j-1} has completed (also makes O stable for the rescale)
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1u);
        tc_fence_after();
      }
      float rs = 0.0f;
      uint8_t* prow = sP + row * 128;
      auto emit = [&](const uint32_t (&sv)[32], int c) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(sv[2 * e]), p.scale_log2, neg_m));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sv[2 * e + 1]), p.scale_log2, neg_m));
          rs += p0 + p1;
          pk[e] = pack_bf16x2(p0, p1);
        }
        // 32 keys = 4 chunks of 16 B; the chunk index inside the 64-key K-block is XOR-swizzled with (row % 8)
        uint8_t* kblock = prow + (c >> 1) * FA_TILE_BYTES;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int chunk = ((c & 1) * 4 + q4) ^ (row & 7);
          *reinterpret_cast<uint4*>(kblock + chunk * 16) = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
        }
      };
      emit(s0, 0);
      emit(s1, 1);
      emit(s2, 2);
      emit(s3, 3);
      l_run = l_run * corr + rs;
      if (j > 0 && any_need) {                                 // rare after the first tiles: rescale the running output in TMEM
#pragma unroll 1
        for (int c = 0; c < DH / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_o + lane_base + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * corr);
          tmem_st_32x32(tmem_o + lane_base + c * 32, r);
        }
        tmem_st_wait();
      }
      fence_proxy_async();            // generic-proxy writes of P -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue: O / l -> bf16 -> global (+ lse)
    mbar_wait(pv_done, (n_tiles - 1) & 1u);
    tc_fence_after();
    const float inv = l_run > 0.0f ? 1.0f / l_run : 0.0f;
    if (MASKED && p.lse != nullptr && grow < p.sq)
      p.lse[(static_cast<long long>(b) * p.hq + h) * p.sq + grow] = l_run > 0.0f ? m_used * 0.6931471805599453f + logf(l_run) : 0.0f;
#pragma unroll 1
    for (int c = 0; c < DH / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_o + lane_base + c * 32, r);
      tmem_ld_wait();
      if (grow < p.sq) {
        bf16* o = p.out + (static_cast<long long>(b) * p.sq + grow) * p.ldo + h * DH + c * 32;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(r[8 * e]) * inv, __uint_as_float(r[8 * e + 1]) * inv);
          pk.y = pack_bf16x2(__uint_as_float(r[8 * e + 2]) * inv, __uint_as_float(r[8 * e + 3]) * inv);
          pk.z = pack_bf16x2(__uint_as_float(r[8 * e + 4]) * inv, __uint_as_float(r[8 * e + 5]) * inv);
          pk.w = pack_bf16x2(__uint_as_float(r[8 * e + 6]) * inv, __uint_as_float(r[8 * e + 7]) * inv);
          *reinterpret_cast<uint4*>(o + 8 * e) = pk;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

typedef CUresult (*PFN_encodeTiledF)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int fa_make_tmap(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld) {
  static PFN_encodeTiledF enc = nullptr;
  if (enc == nullptr) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) {
      set_error("fmha: cuTensorMapEncodeTiled unavailable");
      return -2;
    }
    enc = reinterpret_cast<PFN_encodeTiledF>(fp);
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("fmha: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
    return -3;
  }
  return 0;
}

template <int DH, bool MASKED>
static int fa_launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const FmhaParams& p, dim3 grid, cudaStream_t st) {
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(fmha_fwd_tc_kernel<DH, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, FaCfg<DH>::SMEM);
    if (e != cudaSuccess) {
      set_error("fmha: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    set = true;
  }
  launch_pdl(fmha_fwd_tc_kernel<DH, MASKED>, grid, FA_THREADS, FaCfg<DH>::SMEM, st, tq, tk, tv, p);
  SLAM_LAUNCH_CHECK("slam_attn_fwd.tcgen05");
  return 0;
}

// Returns 1 if the shape is not handled by the tcgen05 kernels (caller falls through to the mma.sync kernel), 0 on success.
int fmha_fwd_tc_try(const slam_attn_args* a, cudaStream_t st) {
  if ((a->dh != 64 && a->dh != 128) || a->sq != a->sk || a->sk < 64 || a->hkv <= 0 || a->hq % a->hkv != 0) return 1;
  if ((reinterpret_cast<uintptr_t>(a->q) & 15) || (reinterpret_cast<uintptr_t>(a->k) & 15) || (reinterpret_cast<uintptr_t>(a->v) & 15)) return 1;
  CUtensorMap tq, tk, tv;
  const long long rows_q = static_cast<long long>(a->batch) * a->sq, rows_k = static_cast<long long>(a->batch) * a->sk;
  int rc;
  if ((rc = fa_make_tmap(&tq, a->q, rows_q, static_cast<long long>(a->hq) * a->dh, a->ldq)) != 0) return rc;
  if ((rc = fa_make_tmap(&tk, a->k, rows_k, static_cast<long long>(a->hkv) * a->dh, a->ldk)) != 0) return rc;
  if ((rc = fa_make_tmap(&tv, a->v, rows_k, static_cast<long long>(a->hkv) * a->dh, a->ldv)) != 0) return rc;
  FmhaParams p;
  p.sq = a->sq;
  p.sk = a->sk;
  p.hq = a->hq;
  p.hkv = a->hkv;
  p.causal = a->causal;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.out = reinterpret_cast<bf16*>(a->out);
  p.ldo = a->ldo;
  p.lse = a->lse;
  p.key_mask = a->key_mask;
  dim3 grid(a->hq, a->batch, static_cast<unsigned>(ceil_div(a->sq, FA_BM)));
  const bool masked = a->causal || a->key_mask != nullptr || a->lse != nullptr;
  if (a->dh == 64) return masked ? fa_launch<64, true>(tq, tk, tv, p, grid, st) : fa_launch<64, false>(tq, tk, tv, p, grid, st);
  return masked ? fa_launch<128, true>(tq, tk, tv, p, grid, st) : fa_launch<128, false>(tq, tk, tv, p, grid, st);
}

// ================================================================================================== backward (dh = 128)
// tcgen05 / TMEM flash-attention backward for the Llama decoder shape (replaces the mma.sync kernel of attention.cu there):
// one CTA = 128 keys of one (batch, query head), looping over the 128-row query tiles that can see them.  Per tile pair
//     S  = Q K^T              dP = dO V^T                      (both operands K-major, fp32 in TMEM)
//     P  = exp2(S c - lse2)   dS = P o (dP - delta) * scale    (4 warps, thread = query row, straight from TMEM; bf16 tiles [q][key])
//     dV += P^T dO            dK += dS^T Q                     (A = P / dS read as MN-major operands, B = dO / Q as MN-major: no transposes)
//     dQ  = dS K                                               (A = dS K-major; B = K as MN-major) -> fp32 vector reductions into dq_accum
// TMEM (512 columns): dK [0,128) and dV [128,256) live across the whole loop; S [256,384) is reused for dQ once the softmax
// warps have read it; dP [384,512).  Shared memory: K, V, 2 x Q (double-buffered), dO, P, dS = 7 x 32 KB; the next dO is requested
// as soon as the dV MMAs (its last readers) have completed.
constexpr int FB_THREADS = 256;   // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warp 3 idle, warps 4-7 softmax / gradients
constexpr int FB_SMEM = 7 * 2 * FA_TILE_BYTES + 256 + 1024;

struct FmhaBwdParams {
  int sq, sk, hq, hkv, causal;
  float scale, scale_log2;
  const float* lse;          // [B, Hq, Sq] natural log
  const float* delta;        // [B, Hq, Sq] rowsum(dO o O)
  const uint8_t* key_mask;   // [B, Sk] or NULL
  float* dq_accum;           // [B, Sq, Hq, 128] fp32, zeroed
  bf16* dk;                  // dk / dv rows: [B*Sk] x ld, head offset added by the kernel
  bf16* dv;
  long long lddk, lddv;
  int head_stride_is_q;      // 1: dk/dv are per-QUERY-head partials [B, Sk, Hq, 128] (GQA); 0: final [.., hkv, 128] buffers (hq == hkv)
};

__global__ void __launch_bounds__(FB_THREADS, 1)
fmha_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmDO, const FmhaBwdParams p) {
  constexpr int DH = 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  constexpr int T2 = 2 * FA_TILE_BYTES;   // one 128 x 128 bf16 tile = two 64-wide blocks
  uint8_t* sK = smem;
  uint8_t* sV = sK + T2;
  uint8_t* sQ = sV + T2;                  // 2 stages
  uint8_t* sdO = sQ + 2 * T2;
  uint8_t* sP = sdO + T2;                 // [q][key]: K-major over keys (dQ), MN-major over keys with K = query rows (dV, dK)
  uint8_t* sdS = sP + T2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + T2);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;            // [2]
  uint64_t* q_empty = bars + 3;           // [2]
  uint64_t* do_full = bars + 5;
  uint64_t* do_empty = bars + 6;
  uint64_t* s_full = bars + 7;
  uint64_t* dp_full = bars + 8;
  uint64_t* pds_ready = bars + 9;
  uint64_t* dq_full = bars + 10;
  uint64_t* dq_drained = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int hk = h / (p.hq / p.hkv);
  const int j = blockIdx.z;                                  // key tile (LPT: with a causal mask tile 0 has the most work and starts first)
  const int k0 = j * 128;
  const int nq = (p.sq + 127) / 128;
  const int i_start = p.causal ? j : 0;
  const int n_it = nq - i_start;

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    mbar_init(do_full, 1);
    mbar_init(do_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(pds_ready, 4);
    mbar_init(dq_full, 1);
    mbar_init(dq_drained, 4);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_dk = tmem_base, t_dv = tmem_base + 128, t_s = tmem_base + 256, t_dp = tmem_base + 384;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * T2);
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        tma_load_2d(sK + nb * FA_TILE_BYTES, &tmK, kv_full, hk * DH + nb * 64, b * p.sk + k0);
        tma_load_2d(sV + nb * FA_TILE_BYTES, &tmV, kv_full, hk * DH + nb * 64, b * p.sk + k0);
      }
    }
    for (int it = 0; it < n_it; ++it) {
      const int q0 = (i_start + it) * 128;
      const int qs = it & 1;
      if (it >= 2) mbar_wait(&q_empty[qs], ((it - 2) >> 1) & 1u);   // dK MMAs of iteration it - 2 (last readers of this Q stage) completed
      if (lane == 0) {
        mbar_arrive_expect_tx(&q_full[qs], T2);
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) tma_load_2d(sQ + qs * T2 + nb * FA_TILE_BYTES, &tmQ, &q_full[qs], h * DH + nb * 64, b * p.sq + q0);
      }
      if (it >= 1) mbar_wait(do_empty, (it - 1) & 1u);               // dV MMAs of the previous iteration completed
      if (lane == 0) {
        mbar_arrive_expect_tx(do_full, T2);
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) tma_load_2d(sdO + nb * FA_TILE_BYTES, &tmDO, do_full, h * DH + nb * 64, b * p.sq + q0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, warp-uniform operands: common.cuh)
    constexpr uint32_t idesc_kk = fa_idesc(128, 128, 0);               // A, B K-major
    constexpr uint32_t idesc_kmn = fa_idesc(128, 64, 1);               // A K-major, B MN-major, 64 output columns per instruction
    constexpr uint32_t idesc_mnmn = fa_idesc(128, 64, 1) | (1u << 15); // A and B MN-major
    constexpr uint32_t BLK = FA_TILE_BYTES >> 4;                       // descriptor units between the two 64-wide blocks of a tile
    const uint32_t lo_k = sw128_kmajor_desc_lo(smem_u32(sK)), lo_v = sw128_kmajor_desc_lo(smem_u32(sV));
    const uint32_t lo_do = sw128_kmajor_desc_lo(smem_u32(sdO)), lo_ds = sw128_kmajor_desc_lo(smem_u32(sdS));
    // P / dS as MN-major A operands (M = keys): two 64-key blocks FA_TILE_BYTES apart, K = query rows (128 B apart)
    const uint32_t mn_p = sw128_mnmajor_desc_lo(smem_u32(sP), FA_TILE_BYTES), mn_ds = sw128_mnmajor_desc_lo(smem_u32(sdS), FA_TILE_BYTES);
    mbar_wait(kv_full, 0);
    for (int it = 0; it < n_it; ++it) {
      const uint32_t lo_q = sw128_kmajor_desc_lo(smem_u32(sQ + (it & 1) * T2));
      mbar_wait(&q_full[it & 1], (it >> 1) & 1u);
      if (it > 0) mbar_wait(dq_drained, (it - 1) & 1u);      // dQ of the previous tile has left the columns S is written to
      tc_fence_after();
      umma_kblock_1(t_s, lo_q, lo_k, idesc_kk, 0u);          // S = Q K^T   (K = dh: two 64-wide blocks)
      umma_kblock_1(t_s, lo_q + BLK, lo_k + BLK, idesc_kk, 1u);
      umma_commit_elect(smem_u32(s_full));
      mbar_wait(do_full, it & 1u);
      tc_fence_after();
      umma_kblock_1(t_dp, lo_do, lo_v, idesc_kk, 0u);        // dP = dO V^T
      umma_kblock_1(t_dp, lo_do + BLK, lo_v + BLK, idesc_kk, 1u);
      umma_commit_elect(smem_u32(dp_full));
      mbar_wait(pds_ready, it & 1u);
      tc_fence_after();
      // K dimension = 128 query rows (dV, dK) resp. 128 keys (dQ); kb = its 64-row halves (+8192 B on the operands whose rows are K)
#pragma unroll
      for (uint32_t kb = 0; kb < 2; ++kb)
#pragma unroll
        for (uint32_t nb = 0; nb < 2; ++nb)                  // dV += P^T dO
          umma_kblock_mna_mnb(t_dv + nb * 64, mn_p + kb * 512u, lo_do + nb * BLK + kb * 512u, idesc_mnmn, (it > 0 || kb > 0) ? 1u : 0u);
      umma_commit_elect(smem_u32(do_empty));                 // dO may be replaced by the next tile's
#pragma unroll
      for (uint32_t kb = 0; kb < 2; ++kb)
#pragma unroll
        for (uint32_t nb = 0; nb < 2; ++nb)                  // dK += dS^T Q
          umma_kblock_mna_mnb(t_dk + nb * 64, mn_ds + kb * 512u, lo_q + nb * BLK + kb * 512u, idesc_mnmn, (it > 0 || kb > 0) ? 1u : 0u);
      umma_commit_elect(smem_u32(&q_empty[it & 1]));
#pragma unroll
      for (uint32_t kb = 0; kb < 2; ++kb)
#pragma unroll
        for (uint32_t nb = 0; nb < 2; ++nb)                  // dQ = dS K   (into the columns S occupied)
          umma_kblock_mnb(t_s + nb * 64, lo_ds + kb * BLK, lo_k + nb * BLK + kb * 512u, idesc_kmn, kb);
      umma_commit_elect(smem_u32(dq_full));
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax / gradient warps (thread = query row resp. key row)
    const int qd = warp & 3;
    const int rl = qd * 32 + lane;                           // row inside the tile = TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(qd * 32) << 16;
    // key validity of this CTA's 128 keys as four 32-bit words (bit e of word w = key k0 + 32 w + e may be attended)
    uint32_t kbits[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int key = k0 + 32 * w + lane;
      const bool ok = key < p.sk && (p.key_mask == nullptr || p.key_mask[static_cast<long long>(b) * p.sk + key] != 0);
      kbits[w] = __ballot_sync(0xffffffffu, ok);
    }
    for (int it = 0; it < n_it; ++it) {
      const int q0 = (i_start + it) * 128;
      const int row = q0 + rl;
      const bool row_ok = row < p.sq;
      const long long stat = (static_cast<long long>(b) * p.hq + h) * p.sq + (row_ok ? row : 0);
      const float lse2 = row_ok ? p.lse[stat] * 1.4426950408889634f : 0.0f;
      const float dlt = row_ok ? p.delta[stat] : 0.0f;
      mbar_wait(s_full, it & 1u);
      mbar_wait(dp_full, it & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(t_s + lane_base + c * 32, sv);
        tmem_ld_32x32(t_dp + lane_base + c * 32, dv);
        tmem_ld_wait();
        uint32_t valid = row_ok ? kbits[c] : 0u;
        if (p.causal) {                                      // keys after the query are masked
          const int lim = row - (k0 + c * 32);               // key offsets 0 .. lim inside this chunk stay
          valid &= lim >= 31 ? 0xffffffffu : (lim < 0 ? 0u : (0xffffffffu >> (31 - lim)));
        }
        uint32_t pp[16], pds[16];                            // P and dS row chunks, packed bf16 pairs
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float p0 = ((valid >> e) & 1u) ? ex2_approx(fmaf(__uint_as_float(sv[e]), p.scale_log2, -lse2)) : 0.0f;
          const float p1 = ((valid >> (e + 1)) & 1u) ? ex2_approx(fmaf(__uint_as_float(sv[e + 1]), p.scale_log2, -lse2)) : 0.0f;
          pp[e >> 1] = pack_bf16x2(p0, p1);
          pds[e >> 1] = pack_bf16x2(p0 * (__uint_as_float(dv[e]) - dlt) * p.scale, p1 * (__uint_as_float(dv[e + 1]) - dlt) * p.scale);
        }
        // row-major [q][key] tiles: 32 keys = 4 chunks of 16 B inside the 64-key block c / 2, XOR-swizzled by (row % 8)
        const int roff = (c >> 1) * FA_TILE_BYTES + rl * 128;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int ch = (((c & 1) * 4 + q4) ^ (rl & 7)) << 4;
          *reinterpret_cast<uint4*>(sP + roff + ch) = make_uint4(pp[4 * q4], pp[4 * q4 + 1], pp[4 * q4 + 2], pp[4 * q4 + 3]);
          *reinterpret_cast<uint4*>(sdS + roff + ch) = make_uint4(pds[4 * q4], pds[4 * q4 + 1], pds[4 * q4 + 2], pds[4 * q4 + 3]);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_ready);
      // dQ of this tile pair: fp32 reduction into the accumulator every key tile of the row adds to
      mbar_wait(dq_full, it & 1u);
      tc_fence_after();
      float* dqrow = p.dq_accum + ((static_cast<long long>(b) * p.sq + (row_ok ? row : 0)) * p.hq + h) * DH;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_s + lane_base + c * 32, r);
        tmem_ld_wait();
        if (row_ok) {                                        // 16-byte vector reductions: the scalar form is bound by L2 atomic operations
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dqrow + c * 32 + e), "f"(__uint_as_float(r[e])),
                         "f"(__uint_as_float(r[e + 1])), "f"(__uint_as_float(r[e + 2])), "f"(__uint_as_float(r[e + 3]))
                         : "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_drained);
    }
    // dK / dV of this (key tile, query head): TMEM lane = key row -> bf16 rows
    const int key = k0 + rl;
    const int hcol = p.head_stride_is_q ? h : hk;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      bf16* dst = (which == 0 ? p.dk : p.dv) + (static_cast<long long>(b) * p.sk + (key < p.sk ? key : 0)) * (which == 0 ? p.lddk : p.lddv) + hcol * DH;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32((which == 0 ? t_dk : t_dv) + lane_base + c * 32, r);
        tmem_ld_wait();
        if (key < p.sk) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(dst + c * 32 + 8 * g) =
                make_uint4(pack_bf16x2(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1])), pack_bf16x2(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3])),
                           pack_bf16x2(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5])), pack_bf16x2(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7])));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// Returns 1 if the shape is not handled (caller falls through to the mma.sync backward), 0 when launched.  Expects delta computed and
// dq_accum zeroed on the stream (slam_attn_bwd does both); dkv_part non-NULL means per-query-head dK/dV partials (GQA).
int fmha_bwd_tc_try(const slam_attn_args* a, cudaStream_t st) {
  if (a->dh != 128 || a->sq != a->sk || a->sk < 64 || a->hkv <= 0 || a->hq % a->hkv != 0 || a->lse == nullptr) return 1;
  if (a->hq != a->hkv && a->dkv_part == nullptr) return 1;
  if ((reinterpret_cast<uintptr_t>(a->q) & 15) || (reinterpret_cast<uintptr_t>(a->k) & 15) || (reinterpret_cast<uintptr_t>(a->v) & 15) ||
      (reinterpret_cast<uintptr_t>(a->dout) & 15))
    return 1;
  CUtensorMap tq, tk, tv, tdo;
  const long long rows_q = static_cast<long long>(a->batch) * a->sq, rows_k = static_cast<long long>(a->batch) * a->sk;
  int rc;
  if ((rc = fa_make_tmap(&tq, a->q, rows_q, static_cast<long long>(a->hq) * a->dh, a->ldq)) != 0) return rc;
  if ((rc = fa_make_tmap(&tk, a->k, rows_k, static_cast<long long>(a->hkv) * a->dh, a->ldk)) != 0) return rc;
  if ((rc = fa_make_tmap(&tv, a->v, rows_k, static_cast<long long>(a->hkv) * a->dh, a->ldv)) != 0) return rc;
  if ((rc = fa_make_tmap(&tdo, a->dout, rows_q, static_cast<long long>(a->hq) * a->dh, a->lddo)) != 0) return rc;
  FmhaBwdParams p;
  p.sq = a->sq;
  p.sk = a->sk;
  p.hq = a->hq;
  p.hkv = a->hkv;
  p.causal = a->causal;
  p.scale = a->scale;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.lse = a->lse;
  p.delta = a->delta;
  p.key_mask = a->key_mask;
  p.dq_accum = a->dq_accum;
  if (a->dkv_part != nullptr) {
    p.head_stride_is_q = 1;
    p.dk = reinterpret_cast<bf16*>(a->dkv_part);
    p.dv = p.dk + rows_k * a->hq * a->dh;
    p.lddk = p.lddv = static_cast<long long>(a->hq) * a->dh;
  } else {
    p.head_stride_is_q = 0;
    p.dk = reinterpret_cast<bf16*>(a->dk);
    p.dv = reinterpret_cast<bf16*>(a->dv);
    p.lddk = a->lddk;
    p.lddv = a->lddv;
  }
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(fmha_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM);
    if (e != cudaSuccess) {
      set_error("fmha bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    set = true;
  }
  dim3 grid(a->hq, a->batch, static_cast<unsigned>(ceil_div(a->sk, 128)));
  launch_pdl(fmha_bwd_tc_kernel, grid, FB_THREADS, FB_SMEM, st, tq, tk, tv, tdo, p);
  SLAM_LAUNCH_CHECK("slam_attn_bwd.tcgen05");
  return 0;
}

}  // namespace slam
