// a7: token cross-entropy + argmax accuracy + dlogits on fp32 logits rows (HF ForCausalLMLoss /
//     utils/metric.py:3-20), and the thin weight-gradient product used for LoRA dA/dB, plus the
//     batched strided cast that packs LoRA adapters into the rank-padded bf16 GEMM operands.
#include <math_constants.h>

#include "../../include/slam_b200.h"
#include "common.cuh"
#include "host.cuh"

namespace slam {

// One CTA per row.  Pass 1: online (max, sum-exp, argmax).  Pass 2: dlogits = (softmax - onehot) * gscale.
__global__ void __launch_bounds__(512) cross_entropy_kernel(const float* __restrict__ logits, long long ldl, const int64_t* __restrict__ targets,
                                                            int vocab, float* __restrict__ loss_sum, int* __restrict__ n_valid,
                                                            int* __restrict__ n_correct, bf16* __restrict__ dlogits, long long lddl,
                                                            const float* __restrict__ grad_scale) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_m[16], s_s[16], s_bv[16];
  __shared__ int s_bi[16];
  __shared__ float s_lse;
  const int row = blockIdx.x;
  const float* lr = logits + static_cast<long long>(row) * ldl;
  const int64_t tgt = targets[row];
  const bool valid = tgt >= 0 && tgt < vocab;  // ignore_index = -100
  bf16* dr = dlogits != nullptr ? dlogits + static_cast<long long>(row) * lddl : nullptr;
  const int nvec = vocab / 4;
  if (!valid) {
    if (dr != nullptr) {
      for (int i = threadIdx.x; i < nvec; i += blockDim.x) *reinterpret_cast<uint2*>(dr + i * 4) = make_uint2(0u, 0u);
      for (int i = nvec * 4 + threadIdx.x; i < vocab; i += blockDim.x) dr[i] = __float2bfloat16(0.0f);
    }
    return;
  }
  float m = -CUDART_INF_F, s = 0.0f, bv = -CUDART_INF_F;
  int bi = 0x7fffffff;
  auto upd = [&](float x, int idx) {
    if (x > bv || (x == bv && idx < bi)) {
      bv = x;
      bi = idx;
    }
    if (x > m) {
      s = s * __expf(m - x) + 1.0f;
      m = x;
    } else {
      s += __expf(x - m);
    }
  };
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(lr + i * 4);
    upd(v.x, i * 4);
    upd(v.y, i * 4 + 1);
    upd(v.z, i * 4 + 2);
    upd(v.w, i * 4 + 3);
  }
  for (int i = nvec * 4 + threadIdx.x; i < vocab; i += blockDim.x) upd(lr[i], i);
  // warp then block reduction of (m, s) and (bv, bi)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float bv2 = __shfl_xor_sync(0xffffffffu, bv, o);
    const int bi2 = __shfl_xor_sync(0xffffffffu, bi, o);
    const float mn = fmaxf(m, m2);
    s = (m == -CUDART_INF_F ? 0.0f : s * __expf(m - mn)) + (m2 == -CUDART_INF_F ? 0.0f : s2 * __expf(m2 - mn));
    m = mn;
    if (bv2 > bv || (bv2 == bv && bi2 < bi)) {
      bv = bv2;
      bi = bi2;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) {
    s_m[warp] = m;
    s_s[warp] = s;
    s_bv[warp] = bv;
    s_bi[warp] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = s_m[0], S = s_s[0], BV = s_bv[0];
    int BI = s_bi[0];
    for (int w = 1; w < nw; ++w) {
      const float mn = fmaxf(M, s_m[w]);
      S = (M == -CUDART_INF_F ? 0.0f : S * __expf(M - mn)) + (s_m[w] == -CUDART_INF_F ? 0.0f : s_s[w] * __expf(s_m[w] - mn));
      M = mn;
      if (s_bv[w] > BV || (s_bv[w] == BV && s_bi[w] < BI)) {
        BV = s_bv[w];
        BI = s_bi[w];
      }
    }
    const float lse = M + logf(S);
    s_lse = lse;
    atomicAdd(loss_sum, lse - lr[tgt]);
    atomicAdd(n_valid, 1);
    if (BI == static_cast<int>(tgt)) atomicAdd(n_correct, 1);
  }
  __syncthreads();
  if (dr == nullptr) return;
  const float lse = s_lse;
  const float gs = grad_scale != nullptr ? grad_scale[0] : 1.0f;
  const int t = static_cast<int>(tgt);
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(lr + i * 4);
    float g0 = __expf(v.x - lse), g1 = __expf(v.y - lse), g2 = __expf(v.z - lse), g3 = __expf(v.w - lse);
    const int c = i * 4;
    if (t >= c && t < c + 4) {
      if (t == c) g0 -= 1.0f;
      else if (t == c + 1) g1 -= 1.0f;
      else if (t == c + 2) g2 -= 1.0f;
      else g3 -= 1.0f;
    }
    uint2 pk;
    pk.x = pack_bf16x2(g0 * gs, g1 * gs);
    pk.y = pack_bf16x2(g2 * gs, g3 * gs);
    *reinterpret_cast<uint2*>(dr + c) = pk;
  }
  for (int i = nvec * 4 + threadIdx.x; i < vocab; i += blockDim.x) {
    float g = __expf(lr[i] - lse) - (i == t ? 1.0f : 0.0f);
    dr[i] = __float2bfloat16(g * gs);
  }
}

// C[P,Q] += scale * sum_{m in chunk} A[m,P] * B[m,Q]; block: 64 Q-columns x all P rows; grid.y splits M.
template <int PMAX>
__global__ void __launch_bounds__(256) wgrad_thin_kernel(const bf16* __restrict__ a, long long lda, int pdim, const bf16* __restrict__ bm, long long ldb,
                                                         int qdim, int m, int m_chunk, float scale, float* __restrict__ c, long long ldc) {
  pdl_trigger();
  pdl_wait();
  constexpr int MB = 32;
  __shared__ float sA[MB][PMAX];
  __shared__ float sB[MB][64 + 1];
  const int tq = threadIdx.x & 63, tp = threadIdx.x >> 6;  // 4 row groups
  const int q0 = blockIdx.x * 64;
  const int m_begin = blockIdx.y * m_chunk;
  const int m_end = min(m, m_begin + m_chunk);
  float acc[PMAX / 4];
#pragma unroll
  for (int i = 0; i < PMAX / 4; ++i) acc[i] = 0.0f;
  for (int mb = m_begin; mb < m_end; mb += MB) {
    __syncthreads();
    for (int i = threadIdx.x; i < MB * PMAX; i += 256) {
      const int r = i / PMAX, pc = i - r * PMAX;
      const int mm = mb + r;
      sA[r][pc] = (mm < m_end && pc < pdim) ? __bfloat162float(a[static_cast<long long>(mm) * lda + pc]) : 0.0f;
    }
    for (int i = threadIdx.x; i < MB * 64; i += 256) {
      const int r = i >> 6, qc = i & 63;
      const int mm = mb + r;
      sB[r][qc] = (mm < m_end && q0 + qc < qdim) ? __bfloat162float(bm[static_cast<long long>(mm) * ldb + q0 + qc]) : 0.0f;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < MB; ++r) {
      const float bv = sB[r][tq];
#pragma unroll
      for (int i = 0; i < PMAX / 4; ++i) acc[i] = fmaf(sA[r][tp + 4 * i], bv, acc[i]);
    }
  }
  if (q0 + tq < qdim) {
#pragma unroll
    for (int i = 0; i < PMAX / 4; ++i) {
      const int pr = tp + 4 * i;
      if (pr < pdim) atomicAdd(c + static_cast<long long>(pr) * ldc + q0 + tq, acc[i] * scale);
    }
  }
}

// Tensor-core variant (mma.sync m16n8k16): C[P,Q] += scale * A[m0:m1, :P]^T B[m0:m1, q0:q0+256].  One CTA = 256 output
// columns x an M-chunk; A^T and B fragments come from cp.async-staged shared tiles through ldmatrix.trans.
__device__ __forceinline__ void wg_cp16(void* dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes) : "memory");
}
template <int PT>  // 16-row tiles of P
__global__ void __launch_bounds__(128) wgrad_thin_mma_kernel(const bf16* __restrict__ a, long long lda, int pdim, const bf16* __restrict__ bm,
                                                             long long ldb, int qdim, int m, int m_chunk, float scale, float* __restrict__ c,
                                                             long long ldc) {
  pdl_trigger();
  pdl_wait();
  constexpr int MC = 64, QT = 256, APITCH = PT * 16 + 8, BPITCH = QT + 8;
  __shared__ __align__(16) bf16 sA[MC * APITCH];
  __shared__ __align__(16) bf16 sB[MC * BPITCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QT;
  const int m_begin = blockIdx.y * m_chunk;
  const int m_end = min(m, m_begin + m_chunk);
  float acc[PT][8][4];
#pragma unroll
  for (int i = 0; i < PT; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.0f;

  for (int mb = m_begin; mb < m_end; mb += MC) {
    __syncthreads();
    for (int i = threadIdx.x; i < MC * (PT * 2); i += 128) {          // A chunk: MC rows x (PT*16) cols, 16-byte pieces
      const int r = i / (PT * 2), pc = (i % (PT * 2)) * 8;
      const bool ok = (mb + r) < m_end && pc < pdim;
      wg_cp16(sA + r * APITCH + pc, a + (ok ? static_cast<long long>(mb + r) * lda + pc : 0), ok ? 16 : 0);
    }
    for (int i = threadIdx.x; i < MC * (QT / 8); i += 128) {          // B chunk: MC rows x 256 cols
      const int r = i / (QT / 8), qc = (i % (QT / 8)) * 8;
      const bool ok = (mb + r) < m_end && (q0 + qc) < qdim;
      wg_cp16(sB + r * BPITCH + qc, bm + (ok ? static_cast<long long>(mb + r) * ldb + q0 + qc : 0), ok ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < MC / 16; ++ks) {
      uint32_t af[PT][4];
#pragma unroll
      for (int pt = 0; pt < PT; ++pt) {
        const bf16* ap = sA + (ks * 16 + (lane & 7) + (lane >> 4) * 8) * APITCH + pt * 16 + ((lane >> 3) & 1) * 8;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(af[pt][0]), "=r"(af[pt][1]), "=r"(af[pt][2]), "=r"(af[pt][3])
                     : "r"(smem_u32(ap)));
      }
#pragma unroll
      for (int nb2 = 0; nb2 < 4; ++nb2) {
        uint32_t bf[4];
        const bf16* bp = sB + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * BPITCH + warp * 64 + nb2 * 16 + (lane >> 4) * 8;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(bf[0]), "=r"(bf[1]), "=r"(bf[2]), "=r"(bf[3])
                     : "r"(smem_u32(bp)));
#pragma unroll
        for (int pt = 0; pt < PT; ++pt) {
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                       : "+f"(acc[pt][2 * nb2][0]), "+f"(acc[pt][2 * nb2][1]), "+f"(acc[pt][2 * nb2][2]), "+f"(acc[pt][2 * nb2][3])
                       : "r"(af[pt][0]), "r"(af[pt][1]), "r"(af[pt][2]), "r"(af[pt][3]), "r"(bf[0]), "r"(bf[1]));
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                       : "+f"(acc[pt][2 * nb2 + 1][0]), "+f"(acc[pt][2 * nb2 + 1][1]), "+f"(acc[pt][2 * nb2 + 1][2]), "+f"(acc[pt][2 * nb2 + 1][3])
                       : "r"(af[pt][0]), "r"(af[pt][1]), "r"(af[pt][2]), "r"(af[pt][3]), "r"(bf[2]), "r"(bf[3]));
        }
      }
    }
  }
#pragma unroll
  for (int pt = 0; pt < PT; ++pt)
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int pr = pt * 16 + (lane >> 2) + ((e >> 1) ? 8 : 0);
        const int qc = q0 + warp * 64 + nb * 8 + 2 * (lane & 3) + (e & 1);
        if (pr < pdim && qc < qdim) atomicAdd(c + static_cast<long long>(pr) * ldc + qc, acc[pt][nb][e] * scale);
      }
}

__global__ void zero2d_kernel(float* c, long long ldc, int rows, int cols) {
  pdl_trigger();
  pdl_wait();
  const long long total = static_cast<long long>(rows) * cols;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) c[(i / cols) * ldc + (i % cols)] = 0.0f;
}

// dst[b][i][j] (or dst[b][j][i] when transpose) = scale * src[b][i][j]; f32 -> bf16
__global__ void pack2d_kernel(const float* __restrict__ src, long long src_bs, long long src_ld, bf16* __restrict__ dst, long long dst_bs,
                              long long dst_ld, int rows, int cols, float scale, int transpose) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const float* s = src + b * src_bs;
  bf16* d = dst + b * dst_bs;
  const long long total = static_cast<long long>(rows) * cols;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    if (transpose) {
      // iterate in destination order so that writes are coalesced
      const long long jj = i / rows, ii = i % rows;
      d[jj * dst_ld + ii] = __float2bfloat16(s[ii * src_ld + jj] * scale);
    } else {
      const long long ii = i / cols, jj = i % cols;
      d[ii * dst_ld + jj] = __float2bfloat16(s[ii * src_ld + jj] * scale);
    }
  }
}

}  // namespace slam

extern "C" {

int slam_cross_entropy(const float* logits, int64_t ldl, const int64_t* targets, int32_t rows, int32_t vocab, float* loss_sum, int32_t* n_valid,
                       int32_t* n_correct, void* dlogits, int64_t lddl, const float* grad_scale, void* stream) {
  using namespace slam;
  SLAM_CHECK_ARG(rows >= 0 && vocab > 0, "cross_entropy: bad shape");
  SLAM_CHECK_ARG(ldl % 4 == 0 && (dlogits == nullptr || lddl % 4 == 0), "cross_entropy: leading dims must be multiples of 4");
  if (rows == 0) return 0;
  launch_pdl(cross_entropy_kernel, rows, 512, 0, reinterpret_cast<cudaStream_t>(stream), logits, ldl, targets, vocab, loss_sum, n_valid, n_correct,
                                                                                reinterpret_cast<bf16*>(dlogits), lddl, grad_scale);
  SLAM_LAUNCH_CHECK("slam_cross_entropy");
  return 0;
}

int slam_wgrad_thin(const void* a, int64_t lda, int32_t p, const void* b, int64_t ldb, int32_t q, int32_t m, float scale, float* c, int64_t ldc,
                    int32_t accumulate, void* stream) {
  using namespace slam;
  SLAM_CHECK_ARG(p > 0 && p <= 64 && q > 0 && m > 0, "wgrad_thin: bad shape p=%d q=%d m=%d", p, q, m);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!accumulate) {   // the M-chunks of the product are merged with fp32 atomics: C starts from zero unless the caller accumulates
    long long zb = ceil_div(static_cast<long long>(p) * q, 256);
    if (zb > 1184) zb = 1184;
    launch_pdl(zero2d_kernel, static_cast<unsigned>(zb), 256, 0, st, c, ldc, p, q);
    SLAM_LAUNCH_CHECK("slam_wgrad_thin.zero");
  }
  const bf16* ap = reinterpret_cast<const bf16*>(a);
  const bf16* bp = reinterpret_cast<const bf16*>(b);
  const bool aligned = p % 8 == 0 && q % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(b) & 15) == 0;
  if (aligned) {
    // tensor-core path: 256 output columns x 128 rows of M per CTA
    const int m_chunk = 128;
    dim3 grid(static_cast<unsigned>(ceil_div(q, 256)), static_cast<unsigned>(ceil_div(m, m_chunk)));
    if (p <= 16)
      launch_pdl(wgrad_thin_mma_kernel<1>, grid, 128, 0, st, ap, lda, p, bp, ldb, q, m, m_chunk, scale, c, ldc);
    else if (p <= 32)
      launch_pdl(wgrad_thin_mma_kernel<2>, grid, 128, 0, st, ap, lda, p, bp, ldb, q, m, m_chunk, scale, c, ldc);
    else
      launch_pdl(wgrad_thin_mma_kernel<4>, grid, 128, 0, st, ap, lda, p, bp, ldb, q, m, m_chunk, scale, c, ldc);
    SLAM_LAUNCH_CHECK("slam_wgrad_thin.mma");
    return 0;
  }
  const int qblocks = static_cast<int>(ceil_div(q, 64));
  // ~128 rows of M per block: many short blocks (atomics merge the partial sums) instead of few long latency-bound ones
  int m_chunk = 128;
  int ysplit = static_cast<int>(ceil_div(m, m_chunk));
  if (static_cast<long long>(ysplit) * qblocks > 16LL * num_sms()) {
    ysplit = static_cast<int>(ceil_div(16LL * num_sms(), qblocks));
    m_chunk = static_cast<int>(ceil_div(ceil_div(m, ysplit), 32) * 32);
    ysplit = static_cast<int>(ceil_div(m, m_chunk));
  }
  dim3 grid(qblocks, ysplit);
  if (p <= 16)
    launch_pdl(wgrad_thin_kernel<16>, grid, 256, 0, st, ap, lda, p, bp, ldb, q, m, m_chunk, scale, c, ldc);
  else if (p <= 32)
    launch_pdl(wgrad_thin_kernel<32>, grid, 256, 0, st, ap, lda, p, bp, ldb, q, m, m_chunk, scale, c, ldc);
  else
    launch_pdl(wgrad_thin_kernel<64>, grid, 256, 0, st, ap, lda, p, bp, ldb, q, m, m_chunk, scale, c, ldc);
  SLAM_LAUNCH_CHECK("slam_wgrad_thin");
  return 0;
}

int slam_pack2d(const float* src, int64_t src_batch_stride, int64_t src_ld, void* dst_bf16, int64_t dst_batch_stride, int64_t dst_ld,
                int32_t batch, int32_t rows, int32_t cols, float scale, int32_t transpose, void* stream) {
  using namespace slam;
  SLAM_CHECK_ARG(batch > 0 && rows > 0 && cols > 0, "pack2d: bad shape");
  long long blocks = ceil_div(static_cast<long long>(rows) * cols, 256);
  if (blocks > 592) blocks = 592;
  dim3 grid(static_cast<unsigned>(blocks), batch);
  launch_pdl(pack2d_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream), src, src_batch_stride, src_ld, reinterpret_cast<bf16*>(dst_bf16),
                                                                         dst_batch_stride, dst_ld, rows, cols, scale, transpose);
  SLAM_LAUNCH_CHECK("slam_pack2d");
  return 0;
}

}  // extern "C"
