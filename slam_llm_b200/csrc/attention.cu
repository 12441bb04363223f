// Flash-style multi-head attention forward + backward (bf16 in, fp32 accumulate, online softmax),
// never materialising the [S,S] score matrix (the reference's HF-4.35 eager LlamaAttention and
// whisper's qkv_attention do).
//   - encoder: non-causal, no mask, dh = 64, forward only (models/encoder.py:26-27);
//   - decoder: causal + key-padding mask, GQA (Hq = g * Hkv) without repeat_kv copies, dh = 64/128,
//     forward + backward (dQ, dK, dV) with the log-sum-exp saved by the forward.
// Round-1 implementation: warp-level mma.sync.m16n8k16 tensor-core tiles with ldmatrix-fed
// fragments (attention is ~1.5 % of decoder FLOPs at S≈400; the tcgen05 path is reserved for the GEMMs).
#include <math_constants.h>
#include <stdlib.h>

#include "../../include/slam_b200.h"
#include "common.cuh"
#include "host.cuh"

namespace slam {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct AttnP {
  const bf16* q; long long ldq;
  const bf16* k; long long ldk;
  const bf16* v; long long ldv;
  bf16* out; long long ldo;
  float* lse;
  const uint8_t* key_mask;
  int batch, sq, sk, hq, hkv;
  int causal;
  float scale;
  const bf16* dout; long long lddo;
  bf16* dq; long long lddq;
  bf16* dk; long long lddk;
  bf16* dv; long long lddv;
  float* delta;
  float* dq_accum;
  bf16* dkv_part;   // optional [2][B][Sk][Hq][dh]: per-q-head dK/dV partials (GQA): 4x more CTAs, group-summed afterwards
  const float* rope_cos;   // optional f32 [seq, dh/2]: the backward's finishing kernel applies the INVERSE rotation to dQ / dK
  const float* rope_sin;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// rows x DH bf16 tile, global (row stride ld) -> shared (row pitch DH+8) with cp.async (no register staging, the
// copies of the NEXT tile overlap the MMAs of the current one); rows >= valid are zero-filled (src-size 0)
template <int DH, int ROWS, int THREADS>
__device__ __forceinline__ void load_tile(bf16* dst, const bf16* src, long long ld, int valid) {
  constexpr int PITCH = DH + 8;
  constexpr int VPR = DH / 8;
#pragma unroll
  for (int i = threadIdx.x; i < ROWS * VPR; i += THREADS) {
    const int r = i / VPR, c = (i - r * VPR) * 8;
    const bool ok = r < valid;
    cp_async16(dst + r * PITCH + c, src + (ok ? static_cast<long long>(r) * ld + c : 0), ok ? 16 : 0);
  }
}

// ------------------------------------------------------------------------------------------- forward
// MT = 16-row query tiles per warp.  MT = 2 (128 query rows per CTA) re-uses every K/V fragment read from shared memory for
// two MMAs, halving the ldmatrix traffic per flop (the limiter of mma.sync attention at dh = 64); dh = 128 keeps MT = 1
// (the fp32 output tile alone is 64 registers per 16 rows).
template <int DH, int MT>
__global__ void __launch_bounds__(128, (DH == 64 && MT == 1) ? 4 : 2) attn_fwd_kernel(const AttnP p) {
  pdl_trigger();
  pdl_wait();
  constexpr int BM = 64 * MT, BN = 64, PITCH = DH + 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_attn);
  bf16* sKb = sQ + BM * PITCH;                 // 2 stages of K
  bf16* sVb = sKb + 2 * BN * PITCH;            // 2 stages of V
  uint8_t* sMask = reinterpret_cast<uint8_t*>(sVb + 2 * BN * PITCH);  // 2 x 64 key-mask bytes
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.hq / p.hkv);
  const int q0 = blockIdx.x * BM;
  const int n_end = p.causal ? min(p.sk, q0 + BM) : p.sk;
  const int n_tiles = (n_end + BN - 1) / BN;
  const bf16* kbase = p.k + static_cast<long long>(b) * p.sk * p.ldk + hk * DH;
  const bf16* vbase = p.v + static_cast<long long>(b) * p.sk * p.ldv + hk * DH;
  const uint8_t* mk = p.key_mask != nullptr ? p.key_mask + static_cast<long long>(b) * p.sk : nullptr;

  auto prefetch = [&](int t) {
    const int n0 = t * BN;
    load_tile<DH, BN, 128>(sKb + (t & 1) * BN * PITCH, kbase + static_cast<long long>(n0) * p.ldk, p.ldk, min(BN, p.sk - n0));
    load_tile<DH, BN, 128>(sVb + (t & 1) * BN * PITCH, vbase + static_cast<long long>(n0) * p.ldv, p.ldv, min(BN, p.sk - n0));
    cp_async_commit();
  };
  load_tile<DH, BM, 128>(sQ, p.q + (static_cast<long long>(b) * p.sq + q0) * p.ldq + h * DH, p.ldq, min(BM, p.sq - q0));
  prefetch(0);
  uint8_t mreg = 1;
  if (mk != nullptr && threadIdx.x < BN) {
    const int col = threadIdx.x;
    sMask[col] = col < p.sk ? mk[col] : 0;
  }

  uint32_t qf[MT][DH / 16][4];
  float o[MT][DH / 8][4];
  float mrow[MT][2], lrow[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) o[mt][i][0] = o[mt][i][1] = o[mt][i][2] = o[mt][i][3] = 0.0f;
    mrow[mt][0] = mrow[mt][1] = -CUDART_INF_F;
    lrow[mt][0] = lrow[mt][1] = 0.0f;
  }
  const float scale_log2 = p.scale * kLog2e;
  const int wrow = q0 + warp * 16 * MT;    // first query row of this warp

  for (int t = 0; t < n_tiles; ++t) {
    const int n0 = t * BN;
    cp_async_wait<0>();
    __syncthreads();                       // tile t (and Q on t == 0) landed; everyone is done with tile t-1
    if (t + 1 < n_tiles) {
      prefetch(t + 1);                     // overlaps the MMAs below
      if (mk != nullptr && threadIdx.x < BN) {
        const int col = n0 + BN + threadIdx.x;
        mreg = col < p.sk ? mk[col] : 0;
      }
    }
    if (t == 0) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks)
          ldsm_x4(qf[mt][ks], sQ + (warp * 16 * MT + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + ks * 16 + (lane >> 4) * 8);
    }
    const bf16* sK = sKb + (t & 1) * BN * PITCH;
    const bf16* sV = sVb + (t & 1) * BN * PITCH;
    const uint8_t* sM = sMask + (t & 1) * BN;

    float s[MT][BN / 8][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int i = 0; i < BN / 8; ++i) s[mt][i][0] = s[mt][i][1] = s[mt][i][2] = s[mt][i][3] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) {
#pragma unroll
      for (int nb2 = 0; nb2 < BN / 16; ++nb2) {
        uint32_t kb[4];
        ldsm_x4(kb, sK + (nb2 * 16 + (lane & 7) + (lane >> 4) * 8) * PITCH + ks * 16 + ((lane >> 3) & 1) * 8);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma16816(s[mt][2 * nb2], qf[mt][ks], kb[0], kb[1]);
          mma16816(s[mt][2 * nb2 + 1], qf[mt][ks], kb[2], kb[3]);
        }
      }
    }
    // scale + mask (log2 domain); only boundary / diagonal / masked tiles need the per-element test
    const bool need_mask = (n0 + BN > p.sk) || (p.causal && n0 + BN > q0) || mk != nullptr;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int nb = 0; nb < BN / 8; ++nb) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = s[mt][nb][e] * scale_log2;
          if (need_mask) {
            const int cl = nb * 8 + 2 * (lane & 3) + (e & 1);
            const int col = n0 + cl;
            const int row = wrow + mt * 16 + (lane >> 2) + ((e >> 1) ? 8 : 0);
            bool ok = col < p.sk && (!p.causal || col <= row);
            if (ok && mk != nullptr) ok = sM[cl] != 0;
            v = ok ? v : -CUDART_INF_F;
          }
          s[mt][nb][e] = v;
        }
      }
      // online softmax, two rows per thread and m-tile
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float mx = -CUDART_INF_F;
#pragma unroll
        for (int nb = 0; nb < BN / 8; ++nb) mx = fmaxf(mx, fmaxf(s[mt][nb][2 * r], s[mt][nb][2 * r + 1]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float m_new = fmaxf(mrow[mt][r], mx);
        const float m_safe = m_new == -CUDART_INF_F ? 0.0f : m_new;
        const float corr = exp2f(mrow[mt][r] - m_safe);
        mrow[mt][r] = m_new;
        float rs = 0.0f;
#pragma unroll
        for (int nb = 0; nb < BN / 8; ++nb) {
          const float p0 = exp2f(s[mt][nb][2 * r] - m_safe);
          const float p1 = exp2f(s[mt][nb][2 * r + 1] - m_safe);
          s[mt][nb][2 * r] = p0;
          s[mt][nb][2 * r + 1] = p1;
          rs += p0 + p1;
        }
        lrow[mt][r] = lrow[mt][r] * corr + rs;
#pragma unroll
        for (int i = 0; i < DH / 8; ++i) {
          o[mt][i][2 * r] *= corr;
          o[mt][i][2 * r + 1] *= corr;
        }
      }
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < BN / 16; ++kk) {
      uint32_t a[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        a[mt][0] = pack_bf16x2(s[mt][2 * kk][0], s[mt][2 * kk][1]);
        a[mt][1] = pack_bf16x2(s[mt][2 * kk][2], s[mt][2 * kk][3]);
        a[mt][2] = pack_bf16x2(s[mt][2 * kk + 1][0], s[mt][2 * kk + 1][1]);
        a[mt][3] = pack_bf16x2(s[mt][2 * kk + 1][2], s[mt][2 * kk + 1][3]);
      }
#pragma unroll
      for (int nb2 = 0; nb2 < DH / 16; ++nb2) {
        uint32_t vb[4];
        ldsm_x4_t(vb, sV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + nb2 * 16 + (lane >> 4) * 8);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma16816(o[mt][2 * nb2], a[mt], vb[0], vb[1]);
          mma16816(o[mt][2 * nb2 + 1], a[mt], vb[2], vb[3]);
        }
      }
    }
    if (t + 1 < n_tiles && mk != nullptr && threadIdx.x < BN) sMask[((t + 1) & 1) * BN + threadIdx.x] = mreg;
  }

  // finalize
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float l = lrow[mt][r];
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      const int row = wrow + mt * 16 + (lane >> 2) + r * 8;
      const float inv = l > 0.0f ? 1.0f / l : 0.0f;
      if (row < p.sq) {
        bf16* orow = p.out + (static_cast<long long>(b) * p.sq + row) * p.ldo + h * DH;
#pragma unroll
        for (int i = 0; i < DH / 8; ++i) {
          const uint32_t pk = pack_bf16x2(o[mt][i][2 * r] * inv, o[mt][i][2 * r + 1] * inv);
          *reinterpret_cast<uint32_t*>(orow + i * 8 + 2 * (lane & 3)) = pk;
        }
        if (p.lse != nullptr && (lane & 3) == 0) {
          // fully masked row: lse = 0 keeps exp(s - lse) = 0 in the backward (s = -inf there)
          const float lse = l > 0.0f ? mrow[mt][r] * kLn2 + logf(l) : 0.0f;
          p.lse[(static_cast<long long>(b) * p.hq + h) * p.sq + row] = lse;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------- backward
// delta[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]: dh / 8 lanes per (token, head) row with 16-byte loads, 32 / (dh / 8) rows per warp
__global__ void attn_delta_kernel(const AttnP p, int dh) {
  pdl_trigger();
  pdl_wait();
  const int lpr = dh >> 3;                                   // lanes per row (dh = 64: 8, dh = 128: 16)
  const int rpw = 32 / lpr;                                  // rows per warp
  const int lane = threadIdx.x & 31;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long total = static_cast<long long>(p.batch) * p.sq * p.hq;
  const long long rowi = warp_global * rpw + lane / lpr;
  const bool ok = rowi < total;
  const long long rr = ok ? rowi : 0;
  const int h = static_cast<int>(rr % p.hq);
  const long long tok = rr / p.hq;                           // b * sq + i
  const int c = (lane % lpr) * 8;
  float acc = 0.0f;
  if (ok) {
    const uint4 a = *reinterpret_cast<const uint4*>(p.out + tok * p.ldo + h * dh + c);
    const uint4 g = *reinterpret_cast<const uint4*>(p.dout + tok * p.lddo + h * dh + c);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 x = unpack_bf16x2(aw[t]), y = unpack_bf16x2(gw[t]);
      acc += x.x * y.x + x.y * y.y;
    }
  }
  for (int o = lpr >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (ok && (lane % lpr) == 0) {
    const int bb = static_cast<int>(tok / p.sq), i = static_cast<int>(tok % p.sq);
    p.delta[(static_cast<long long>(bb) * p.hq + h) * p.sq + i] = acc;
  }
}

// One CTA = 64 keys of one (batch, kv head); loops over the q heads of the group and over query blocks.
// Each warp owns 16 keys: dK/dV accumulate in registers (no atomics, GQA-summed in place);
// dQ partials go through fp32 atomics into dq_accum (as FlashAttention-2 does).
template <int DH, int BQ>
__global__ void __launch_bounds__(128) attn_bwd_kernel(const AttnP p) {
  pdl_trigger();
  pdl_wait();
  constexpr int BN = 64, PITCH = DH + 8, SPITCH = BQ + 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  bf16* sK = reinterpret_cast<bf16*>(smem_attn);
  bf16* sV = sK + BN * PITCH;
  bf16* sQb = sV + BN * PITCH;             // 2 stages of Q
  bf16* sdOb = sQb + 2 * BQ * PITCH;       // 2 stages of dO
  bf16* sdS = sdOb + 2 * BQ * PITCH;       // [key][q]
  float* sLseB = reinterpret_cast<float*>(sdS + BN * SPITCH);  // 2 stages of lse (log2 domain) ...
  float* sDeltaB = sLseB + 2 * BQ;                             // ... and delta
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid = (heads, batch, key blocks): CTAs are dispatched x-fastest, so all CTAs of key block 0 start first.  With a causal mask
  // key block j only meets the query rows >= 64 j: heaviest work first (LPT) trims the tail of the 3-wave grid.
  const int b = blockIdx.y;
  const int k0 = blockIdx.z * BN;
  // split_heads: one CTA per (key block, Q head) writing per-head dK/dV partials (more parallelism, shorter critical path);
  // otherwise one CTA per (key block, KV head) looping over the heads of its group with dK/dV kept in registers
  const bool split_heads = p.dkv_part != nullptr;
  const int group_all = p.hq / p.hkv;
  const int hk = split_heads ? static_cast<int>(blockIdx.x) / group_all : static_cast<int>(blockIdx.x);
  const int h_first = split_heads ? static_cast<int>(blockIdx.x) : hk * group_all;
  const int group = split_heads ? 1 : group_all;
  const float scale_log2 = p.scale * kLog2e;

  load_tile<DH, BN, 128>(sK, p.k + (static_cast<long long>(b) * p.sk + k0) * p.ldk + hk * DH, p.ldk, min(BN, p.sk - k0));
  load_tile<DH, BN, 128>(sV, p.v + (static_cast<long long>(b) * p.sk + k0) * p.ldv + hk * DH, p.ldv, min(BN, p.sk - k0));

  float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.0f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.0f;
  }
  // key validity for the two key rows this thread owns in the transposed score tile
  bool key_ok[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int key = k0 + warp * 16 + (lane >> 2) + r * 8;
    key_ok[r] = key < p.sk && (p.key_mask == nullptr || p.key_mask[static_cast<long long>(b) * p.sk + key] != 0);
  }

  const int q_start = p.causal ? (k0 / BQ) * BQ : 0;
  const int nq = (p.sq - q_start + BQ - 1) / BQ;
  const int n_iter = group * nq;

  // stage j of the (head-in-group, query-block) sequence -> shared stage j & 1 (cp.async: overlaps the MMAs of stage j-1)
  auto prefetch = [&](int j) {
    const int h = h_first + j / nq;
    const int q0 = q_start + (j % nq) * BQ;
    const int vq = min(BQ, p.sq - q0);
    const int st = j & 1;
    load_tile<DH, BQ, 128>(sQb + st * BQ * PITCH, p.q + (static_cast<long long>(b) * p.sq + q0) * p.ldq + h * DH, p.ldq, vq);
    load_tile<DH, BQ, 128>(sdOb + st * BQ * PITCH, p.dout + (static_cast<long long>(b) * p.sq + q0) * p.lddo + h * DH, p.lddo, vq);
    if (threadIdx.x < 2 * BQ) {
      const int i = threadIdx.x % BQ;
      const bool is_delta = threadIdx.x >= BQ;
      const long long off = (static_cast<long long>(b) * p.hq + h) * p.sq + q0 + (i < vq ? i : 0);
      cp_async4((is_delta ? sDeltaB : sLseB) + st * BQ + i, (is_delta ? p.delta : p.lse) + off, i < vq ? 4 : 0);
    }
    cp_async_commit();
  };
  if (n_iter > 0) prefetch(0);

  for (int j = 0; j < n_iter; ++j) {
    {
      const int h = h_first + j / nq;
      const int q0 = q_start + (j % nq) * BQ;
      cp_async_wait<0>();
      __syncthreads();                      // stage j landed; everyone is done with stage j-1 (and with sdS)
      if (j + 1 < n_iter) prefetch(j + 1);
      const bf16* sQ = sQb + (j & 1) * BQ * PITCH;
      const bf16* sdO = sdOb + (j & 1) * BQ * PITCH;
      const float* sLse = sLseB + (j & 1) * BQ;
      const float* sDelta = sDeltaB + (j & 1) * BQ;

      // S^T = K_w Q^T and dP^T = V_w dO^T   (16 keys x BQ queries per warp)
      float st[BQ / 8][4], dpt[BQ / 8][4];
#pragma unroll
      for (int i = 0; i < BQ / 8; ++i) {
        st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.0f;
        dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.0f;
      }
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) {
        uint32_t ka[4], va[4];
        const int arow = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int acol = ks * 16 + (lane >> 4) * 8;
        ldsm_x4(ka, sK + arow * PITCH + acol);
        ldsm_x4(va, sV + arow * PITCH + acol);
#pragma unroll
        for (int nb2 = 0; nb2 < BQ / 16; ++nb2) {
          uint32_t qb[4], ob[4];
          const int brow = nb2 * 16 + (lane & 7) + (lane >> 4) * 8;
          const int bcol = ks * 16 + ((lane >> 3) & 1) * 8;
          ldsm_x4(qb, sQ + brow * PITCH + bcol);
          ldsm_x4(ob, sdO + brow * PITCH + bcol);
          mma16816(st[2 * nb2], ka, qb[0], qb[1]);
          mma16816(st[2 * nb2 + 1], ka, qb[2], qb[3]);
          mma16816(dpt[2 * nb2], va, ob[0], ob[1]);
          mma16816(dpt[2 * nb2 + 1], va, ob[2], ob[3]);
        }
      }
      // P^T and dS^T in place; write dS^T (bf16) to shared [key][q]
#pragma unroll
      for (int nb = 0; nb < BQ / 8; ++nb) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1;
          const int key = k0 + warp * 16 + (lane >> 2) + r * 8;
          const int ql = nb * 8 + 2 * (lane & 3) + (e & 1);
          const int qi = q0 + ql;
          const bool ok = key_ok[r] && qi < p.sq && (!p.causal || key <= qi);
          const float pv = ok ? exp2f(st[nb][e] * scale_log2 - sLse[ql] * kLog2e) : 0.0f;
          const float ds = pv * (dpt[nb][e] - sDelta[ql]) * p.scale;
          st[nb][e] = pv;
          dpt[nb][e] = ds;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int krow = warp * 16 + (lane >> 2) + r * 8;
          *reinterpret_cast<uint32_t*>(sdS + krow * SPITCH + nb * 8 + 2 * (lane & 3)) = pack_bf16x2(dpt[nb][2 * r], dpt[nb][2 * r + 1]);
        }
      }
      // dV += P^T dO ; dK += dS^T Q   (contraction over the BQ queries)
#pragma unroll
      for (int kk = 0; kk < BQ / 16; ++kk) {
        uint32_t pa[4], da[4];
        pa[0] = pack_bf16x2(st[2 * kk][0], st[2 * kk][1]);
        pa[1] = pack_bf16x2(st[2 * kk][2], st[2 * kk][3]);
        pa[2] = pack_bf16x2(st[2 * kk + 1][0], st[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(st[2 * kk + 1][2], st[2 * kk + 1][3]);
        da[0] = pack_bf16x2(dpt[2 * kk][0], dpt[2 * kk][1]);
        da[1] = pack_bf16x2(dpt[2 * kk][2], dpt[2 * kk][3]);
        da[2] = pack_bf16x2(dpt[2 * kk + 1][0], dpt[2 * kk + 1][1]);
        da[3] = pack_bf16x2(dpt[2 * kk + 1][2], dpt[2 * kk + 1][3]);
#pragma unroll
        for (int nb2 = 0; nb2 < DH / 16; ++nb2) {
          uint32_t ob[4], qb[4];
          const int brow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          const int bcol = nb2 * 16 + (lane >> 4) * 8;
          ldsm_x4_t(ob, sdO + brow * PITCH + bcol);
          ldsm_x4_t(qb, sQ + brow * PITCH + bcol);
          mma16816(dv[2 * nb2], pa, ob[0], ob[1]);
          mma16816(dv[2 * nb2 + 1], pa, ob[2], ob[3]);
          mma16816(dk[2 * nb2], da, qb[0], qb[1]);
          mma16816(dk[2 * nb2 + 1], da, qb[2], qb[3]);
        }
      }
      __syncthreads();
      // dQ[BQ x DH] += dS[BQ x 64] K[64 x DH]: warp -> (query block of 16, DH slice)
      {
        constexpr int QB = BQ / 16;          // query 16-blocks
        constexpr int SL = 4 / QB;           // DH slices
        constexpr int DSL = DH / SL;         // slice width
        const int qb_i = warp % QB, sl = warp / QB;
        float dq[DSL / 8][4];
#pragma unroll
        for (int i = 0; i < DSL / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.0f;
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
          uint32_t a[4];
          // A = dS (q x key) read transposed from sdS[key][q]
          ldsm_x4_t(a, sdS + (kk * 16 + (lane & 7) + (lane >> 4) * 8) * SPITCH + qb_i * 16 + ((lane >> 3) & 1) * 8);
#pragma unroll
          for (int nb2 = 0; nb2 < DSL / 16; ++nb2) {
            uint32_t kb[4];
            ldsm_x4_t(kb, sK + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + sl * DSL + nb2 * 16 + (lane >> 4) * 8);
            mma16816(dq[2 * nb2], a, kb[0], kb[1]);
            mma16816(dq[2 * nb2 + 1], a, kb[2], kb[3]);
          }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int qi = q0 + qb_i * 16 + (lane >> 2) + r * 8;
          if (qi < p.sq) {
            float* dst = p.dq_accum + ((static_cast<long long>(b) * p.sq + qi) * p.hq + h) * DH + sl * DSL;
#pragma unroll
            for (int i = 0; i < DSL / 8; ++i) {
              atomicAdd(dst + i * 8 + 2 * (lane & 3), dq[i][2 * r]);
              atomicAdd(dst + i * 8 + 2 * (lane & 3) + 1, dq[i][2 * r + 1]);
            }
          }
        }
      }
    }
  }
  // write dK, dV (bf16)
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int key = k0 + warp * 16 + (lane >> 2) + r * 8;
    if (key < p.sk) {
      bf16* dkr = p.dk + (static_cast<long long>(b) * p.sk + key) * p.lddk + hk * DH;
      bf16* dvr = p.dv + (static_cast<long long>(b) * p.sk + key) * p.lddv + hk * DH;
      if (split_heads) {
        const long long part = static_cast<long long>(p.batch) * p.sk * p.hq * DH;
        dkr = p.dkv_part + ((static_cast<long long>(b) * p.sk + key) * p.hq + h_first) * DH;
        dvr = dkr + part;
      }
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) {
        *reinterpret_cast<uint32_t*>(dkr + i * 8 + 2 * (lane & 3)) = pack_bf16x2(dk[i][2 * r], dk[i][2 * r + 1]);
        *reinterpret_cast<uint32_t*>(dvr + i * 8 + 2 * (lane & 3)) = pack_bf16x2(dv[i][2 * r], dv[i][2 * r + 1]);
      }
    }
  }
}

// dK/dV partials [2][rows][Hq][dh] (bf16) -> sum over the G heads of each KV group -> dk/dv [rows][Hkv][dh] with row strides
__global__ void attn_group_sum_kernel(const bf16* __restrict__ part, bf16* __restrict__ dk, long long lddk, bf16* __restrict__ dv, long long lddv,
                                      long long rows, int hq, int hkv, int dh) {
  pdl_trigger();
  pdl_wait();
  const int g = hq / hkv;
  const int vec_per_row = hkv * dh / 8;
  const long long total = 2 * rows * vec_per_row;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const int which = static_cast<int>(i / (rows * vec_per_row));
    const long long rem = i % (rows * vec_per_row);
    const long long r = rem / vec_per_row;
    const int c = static_cast<int>(rem % vec_per_row) * 8;      // column inside [hkv*dh]
    const int hk = c / dh, d0 = c % dh;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bf16* src = part + (static_cast<long long>(which) * rows + r) * hq * dh + static_cast<long long>(hk) * g * dh + d0;
    for (int j = 0; j < g; ++j) {
      const uint4 v = *reinterpret_cast<const uint4*>(src + static_cast<long long>(j) * dh);
      const float2 a0 = unpack_bf16x2(v.x), a1 = unpack_bf16x2(v.y), a2 = unpack_bf16x2(v.z), a3 = unpack_bf16x2(v.w);
      acc[0] += a0.x; acc[1] += a0.y; acc[2] += a1.x; acc[3] += a1.y; acc[4] += a2.x; acc[5] += a2.y; acc[6] += a3.x; acc[7] += a3.y;
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]); o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    bf16* dst = (which == 0 ? dk + r * lddk : dv + r * lddv) + c;
    *reinterpret_cast<uint4*>(dst) = o;
  }
}

// Finishing pass of the attention backward in ONE kernel (was: group sum, fp32 -> bf16 conversion of dQ and a separate inverse-RoPE pass):
//   dQ  = [R^T] dq_accum                       f32 [rows_q, Hq, dh] -> bf16 (row stride lddq)
//   dK  = [R^T] sum over the KV group of dK partials (or dK itself when the kernel wrote it directly)
//   dV  =       sum over the KV group of dV partials (nothing to do when written directly)
// R^T = inverse rotate-half RoPE with position = row % seq (HF apply_rotary_pos_emb transposed), applied when cos/sin are given.
// One thread = 8 "lo" + 8 "hi" elements (d, d + dh/2) of one (row, head).
__global__ void attn_bwd_finish_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, long long lddq, long long rows_q, int seq_q, int hq,
                                       const bf16* __restrict__ part, bf16* __restrict__ dk, long long lddk, bf16* __restrict__ dv, long long lddv,
                                       long long rows_k, int seq_k, int hkv, int dh, const float* __restrict__ cosT, const float* __restrict__ sinT) {
  pdl_trigger();
  pdl_wait();
  const int half = dh / 2;
  const int vph = half / 8;                                           // threads per (row, head)
  const int g = hq / hkv;
  const long long nq = rows_q * hq * vph;
  const long long nk = rows_k * hkv * vph;
  const bool k_work = part != nullptr || cosT != nullptr;           // dK needs a pass (group sum and / or rotation)
  const bool v_work = part != nullptr;
  const long long total = nq + (k_work ? nk : 0) + (v_work ? nk : 0);
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    float lo[8], hi[8];
    bf16* dst;
    int pos;
    bool rotate = cosT != nullptr;
    int v;
    if (i < nq) {
      v = static_cast<int>(i % vph);
      const int h = static_cast<int>((i / vph) % hq);
      const long long r = i / (static_cast<long long>(vph) * hq);
      const float* src = acc + (r * hq + h) * dh + v * 8;
      const float4 a0 = *reinterpret_cast<const float4*>(src), a1 = *reinterpret_cast<const float4*>(src + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(src + half), b1 = *reinterpret_cast<const float4*>(src + half + 4);
      lo[0] = a0.x; lo[1] = a0.y; lo[2] = a0.z; lo[3] = a0.w; lo[4] = a1.x; lo[5] = a1.y; lo[6] = a1.z; lo[7] = a1.w;
      hi[0] = b0.x; hi[1] = b0.y; hi[2] = b0.z; hi[3] = b0.w; hi[4] = b1.x; hi[5] = b1.y; hi[6] = b1.z; hi[7] = b1.w;
      dst = dq + r * lddq + static_cast<long long>(h) * dh + v * 8;
      pos = static_cast<int>(r % seq_q);
    } else {
      long long j = i - nq;
      const bool is_v = k_work ? j >= nk : true;
      if (is_v && k_work) j -= nk;
      v = static_cast<int>(j % vph);
      const int hk = static_cast<int>((j / vph) % hkv);
      const long long r = j / (static_cast<long long>(vph) * hkv);
      bf16* out = (is_v ? dv + r * lddv : dk + r * lddk) + static_cast<long long>(hk) * dh + v * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) lo[e] = hi[e] = 0.f;
      if (part != nullptr) {
        const bf16* src = part + ((is_v ? rows_k : 0) + r) * hq * dh + static_cast<long long>(hk) * g * dh + v * 8;
        for (int t = 0; t < g; ++t) {
          float a[8], b[8];
          unpack8(*reinterpret_cast<const bf16x8*>(src + static_cast<long long>(t) * dh), a);
          unpack8(*reinterpret_cast<const bf16x8*>(src + static_cast<long long>(t) * dh + half), b);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            lo[e] += a[e];
            hi[e] += b[e];
          }
        }
      } else {
        unpack8(*reinterpret_cast<const bf16x8*>(out), lo);
        unpack8(*reinterpret_cast<const bf16x8*>(out + half), hi);
      }
      dst = out;
      pos = static_cast<int>(r % seq_k);
      rotate = rotate && !is_v;
    }
    if (rotate) {
      const float* c = cosT + static_cast<long long>(pos) * half + v * 8;
      const float* sn = sinT + static_cast<long long>(pos) * half + v * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float l = lo[e], h2 = hi[e];
        lo[e] = l * c[e] + h2 * sn[e];
        hi[e] = h2 * c[e] - l * sn[e];
      }
    }
    *reinterpret_cast<bf16x8*>(dst) = pack8(lo);
    *reinterpret_cast<bf16x8*>(dst + half) = pack8(hi);
  }
}

// dq_accum f32 [B,Sq,Hq,DH] -> dq bf16 with row stride lddq
__global__ void attn_dq_convert_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, long long lddq, long long rows, int width) {
  pdl_trigger();
  pdl_wait();
  const int vpr = width / 4;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = rows * vpr;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const long long r = i / vpr;
    const int c = static_cast<int>(i % vpr) * 4;
    const float4 v = *reinterpret_cast<const float4*>(acc + r * width + c);
    uint2 pk;
    pk.x = pack_bf16x2(v.x, v.y);
    pk.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(dq + r * lddq + c) = pk;
  }
}

static int fill_params(const slam_attn_args* a, AttnP& p, bool bwd) {
  SLAM_CHECK_ARG(a != nullptr, "attn: null args");
  SLAM_CHECK_ARG(a->dh == 64 || a->dh == 128, "attn: dh must be 64 or 128 (got %d)", a->dh);
  SLAM_CHECK_ARG(a->hkv > 0 && a->hq % a->hkv == 0, "attn: hq %% hkv != 0");
  SLAM_CHECK_ARG(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0, "attn: row strides must be multiples of 8");
  SLAM_CHECK_ARG(!a->causal || a->sq == a->sk, "attn: causal requires sq == sk");
  p.q = reinterpret_cast<const bf16*>(a->q); p.ldq = a->ldq;
  p.k = reinterpret_cast<const bf16*>(a->k); p.ldk = a->ldk;
  p.v = reinterpret_cast<const bf16*>(a->v); p.ldv = a->ldv;
  p.out = reinterpret_cast<bf16*>(a->out); p.ldo = a->ldo;
  p.lse = a->lse;
  p.key_mask = a->key_mask;
  p.batch = a->batch; p.sq = a->sq; p.sk = a->sk; p.hq = a->hq; p.hkv = a->hkv;
  p.causal = a->causal;
  p.scale = a->scale;
  p.dout = reinterpret_cast<const bf16*>(a->dout); p.lddo = a->lddo;
  p.dq = reinterpret_cast<bf16*>(a->dq); p.lddq = a->lddq;
  p.dk = reinterpret_cast<bf16*>(a->dk); p.lddk = a->lddk;
  p.dv = reinterpret_cast<bf16*>(a->dv); p.lddv = a->lddv;
  p.delta = a->delta;
  p.dq_accum = a->dq_accum;
  p.dkv_part = (bwd && a->hq != a->hkv) ? reinterpret_cast<bf16*>(a->dkv_part) : nullptr;
  p.rope_cos = bwd ? a->rope_cos : nullptr;
  p.rope_sin = bwd ? a->rope_sin : nullptr;
  if (bwd) {
    SLAM_CHECK_ARG(a->lse && a->delta && a->dq_accum && a->dout && a->dq && a->dk && a->dv, "attn_bwd: missing buffers");
    SLAM_CHECK_ARG(a->lddo % 8 == 0 && a->lddq % 8 == 0 && a->lddk % 8 == 0 && a->lddv % 8 == 0, "attn_bwd: row strides must be multiples of 8");
  }
  return 0;
}

template <int DH, int MT>
static int launch_fwd(const AttnP& p, cudaStream_t st) {
  constexpr int SMEM = (4 + MT) * 64 * (DH + 8) * 2 + 128;
  static bool set = false;
  if (!set) {
    cudaFuncSetAttribute(attn_fwd_kernel<DH, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    set = true;
  }
  dim3 grid(static_cast<unsigned>(ceil_div(p.sq, 64 * MT)), p.hq, p.batch);
  launch_pdl(attn_fwd_kernel<DH, MT>, grid, 128, SMEM, st, p);
  SLAM_LAUNCH_CHECK("slam_attn_fwd");
  return 0;
}
template <int DH, int BQ>
static int launch_bwd(const AttnP& p, cudaStream_t st) {
  constexpr int SMEM = (2 * 64 + 4 * BQ) * (DH + 8) * 2 + 64 * (BQ + 8) * 2 + 4 * BQ * 4;
  static bool set = false;
  if (!set) {
    cudaFuncSetAttribute(attn_bwd_kernel<DH, BQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    set = true;
  }
  dim3 grid(p.dkv_part != nullptr ? p.hq : p.hkv, p.batch, static_cast<unsigned>(ceil_div(p.sk, 64)));
  launch_pdl(attn_bwd_kernel<DH, BQ>, grid, 128, SMEM, st, p);
  SLAM_LAUNCH_CHECK("slam_attn_bwd");
  return 0;
}

}  // namespace slam

namespace slam {
int fmha_fwd_tc_try(const slam_attn_args* a, cudaStream_t st);   // fmha_tc.cu: 0 = launched, 1 = shape not handled, else error
int fmha_bwd_tc_try(const slam_attn_args* a, cudaStream_t st);   // same contract; needs delta computed and dq_accum zeroed
}

extern "C" int slam_attn_fwd(const slam_attn_args* a, void* stream) {
  using namespace slam;
  AttnP p;
  int rc = fill_params(a, p, false);
  if (rc != 0) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static const bool use_tc = []() {
    const char* e = getenv("SLAM_ATTN_TC");
    return e == nullptr || e[0] != '0';
  }();
  if (use_tc) {
    rc = fmha_fwd_tc_try(a, st);       // tcgen05 / TMEM kernel for the encoder shape (dh = 64, non-causal, unmasked)
    if (rc != 1) return rc;
  }
  if (a->dh == 128) return launch_fwd<128, 1>(p, st);
  // MT = 2 (32 rows per warp) measured SLOWER on B200 (324 us vs 261 us per whisper-large-v3 layer: 255 registers, 8 warps/SM);
  // it stays instantiated for experiments behind SLAM_ATTN_MT2.
#ifdef SLAM_ATTN_MT2
  if (a->sq >= 256) return launch_fwd<64, 2>(p, st);
#endif
  return launch_fwd<64, 1>(p, st);
}

extern "C" int slam_attn_bwd(const slam_attn_args* a, void* stream) {
  using namespace slam;
  AttnP p;
  int rc = fill_params(a, p, true);
  if (rc != 0) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long rows = static_cast<long long>(p.batch) * p.sq;
  const int width = p.hq * a->dh;
  cudaError_t e = cudaMemsetAsync(p.dq_accum, 0, static_cast<size_t>(rows) * width * sizeof(float), st);
  if (e != cudaSuccess) {
    set_error("attn_bwd: memset failed: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  const long long total_warps = ceil_div(static_cast<long long>(p.batch) * p.sq * p.hq, 32 / (a->dh / 8));
  launch_pdl(attn_delta_kernel, static_cast<unsigned>(ceil_div(total_warps, 8)), 256, 0, st, p, a->dh);
  SLAM_LAUNCH_CHECK("slam_attn_bwd.delta");
  static const bool use_tc = []() {
    const char* e = getenv("SLAM_ATTN_BWD_TC");
    return e == nullptr || e[0] != '0';
  }();
  rc = use_tc ? fmha_bwd_tc_try(a, st) : 1;           // tcgen05 / TMEM kernel for the Llama decoder shape (dh = 128)
  if (rc == 1) rc = a->dh == 64 ? launch_bwd<64, 64>(p, st) : launch_bwd<128, 32>(p, st);
  if (rc != 0) return rc;
  SLAM_CHECK_ARG(p.lddq >= width || p.hq * a->dh <= p.lddq, "attn_bwd: lddq too small");
  SLAM_CHECK_ARG((a->rope_cos == nullptr) == (a->rope_sin == nullptr), "attn_bwd: rope_cos and rope_sin go together");
  SLAM_CHECK_ARG(a->dh % 16 == 0, "attn_bwd: dh %% 16 != 0");
  {
    const long long krows = static_cast<long long>(p.batch) * p.sk;
    const int vph = a->dh / 16;
    long long items = rows * p.hq * vph + 2 * krows * p.hkv * vph;
    long long blocks = ceil_div(items, 256);
    if (blocks > num_sms() * 16) blocks = num_sms() * 16;
    launch_pdl(attn_bwd_finish_kernel, static_cast<unsigned>(blocks), 256, 0, st, p.dq_accum, p.dq, p.lddq, rows, p.sq, p.hq,
               static_cast<const bf16*>(p.dkv_part), p.dk, p.lddk, p.dv, p.lddv, krows, p.sk, p.hkv, a->dh, p.rope_cos, p.rope_sin);
    SLAM_LAUNCH_CHECK("slam_attn_bwd.finish");
  }
  return 0;
}
