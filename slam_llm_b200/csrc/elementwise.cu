// HBM-bound kernels of the SLAM-LLM step: casts, transposes, row gathers, norms, RoPE, SwiGLU,
// embedding merge, conv im2col, AdamW.  All are coalesced, 16-byte vectorised, fp32 math inside.
#include "../../include/slam_b200.h"
#include "common.cuh"
#include "host.cuh"

namespace slam {

// block-wide sum of one float (blockDim.x <= 1024, multiple of 32)
__device__ __forceinline__ float block_sum(float v, float* sbuf) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sbuf[warp] = v;
  __syncthreads();
  float t = lane < nw ? sbuf[lane] : 0.0f;
  t = warp_sum(t);
  return t;
}

// ---------------------------------------------------------------- casts / adds
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, int64_t n, float scale) {
  pdl_trigger();
  pdl_wait();
  int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 8;
  for (; i < n; i += stride) {
    if (i + 8 <= n) {
      const float4 a = *reinterpret_cast<const float4*>(x + i);
      const float4 b = *reinterpret_cast<const float4*>(x + i + 4);
      float f[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, b.x * scale, b.y * scale, b.z * scale, b.w * scale};
      *reinterpret_cast<bf16x8*>(y + i) = pack8(f);
    } else {
      for (int64_t j = i; j < n; ++j) y[j] = __float2bfloat16(x[j] * scale);
    }
  }
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ x, float* __restrict__ y, int64_t n) {
  pdl_trigger();
  pdl_wait();
  int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 8;
  for (; i < n; i += stride) {
    if (i + 8 <= n) {
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(x + i), f);
      *reinterpret_cast<float4*>(y + i) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(y + i + 4) = make_float4(f[4], f[5], f[6], f[7]);
    } else {
      for (int64_t j = i; j < n; ++j) y[j] = __bfloat162float(x[j]);
    }
  }
}
__global__ void add_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ y, int64_t n) {
  pdl_trigger();
  pdl_wait();
  int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 8;
  for (; i < n; i += stride) {
    if (i + 8 <= n) {
      float fa[8], fb[8];
      unpack8(*reinterpret_cast<const bf16x8*>(a + i), fa);
      unpack8(*reinterpret_cast<const bf16x8*>(b + i), fb);
#pragma unroll
      for (int e = 0; e < 8; ++e) fa[e] += fb[e];
      *reinterpret_cast<bf16x8*>(y + i) = pack8(fa);
    } else {
      for (int64_t j = i; j < n; ++j) y[j] = __float2bfloat16(__bfloat162float(a[j]) + __bfloat162float(b[j]));
    }
  }
}
__global__ void relu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ yv, bf16* __restrict__ dx, int64_t n) {
  pdl_trigger();
  pdl_wait();
  int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 8;
  for (; i < n; i += stride) {
    if (i + 8 <= n) {
      float fd[8], fy[8];
      unpack8(*reinterpret_cast<const bf16x8*>(dy + i), fd);
      unpack8(*reinterpret_cast<const bf16x8*>(yv + i), fy);
#pragma unroll
      for (int e = 0; e < 8; ++e) fd[e] = fy[e] > 0.0f ? fd[e] : 0.0f;
      *reinterpret_cast<bf16x8*>(dx + i) = pack8(fd);
    } else {
      for (int64_t j = i; j < n; ++j) dx[j] = __bfloat162float(yv[j]) > 0.0f ? dy[j] : __float2bfloat16(0.0f);
    }
  }
}

static inline int ew_grid(int64_t n, int per_thread, int threads) {
  int64_t b = ceil_div(n, static_cast<int64_t>(per_thread) * threads);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

// ---------------------------------------------------------------- transpose (bf16, 64x64 tiles via smem)
__global__ void transpose_bf16_kernel(const bf16* __restrict__ x, int64_t ldx, bf16* __restrict__ y, int64_t ldy, int rows, int cols) {
  pdl_trigger();
  pdl_wait();
  __shared__ bf16 tile[64][66];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 256 threads: 64 x 4
  for (int i = ty; i < 64; i += 4) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? x[static_cast<int64_t>(r) * ldx + c] : __float2bfloat16(0.0f);
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 4) {
    const int c = c0 + i, r = r0 + tx;
    if (c < cols && r < rows) y[static_cast<int64_t>(c) * ldy + r] = tile[tx][i];
  }
}

// batched f32 transpose: dst[b][j][i] = src[b][i][j]   (conv1d weight-gradient layout change, small)
__global__ void transpose_f32_batched_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const float* s = src + static_cast<int64_t>(b) * rows * cols;
  float* d = dst + static_cast<int64_t>(b) * rows * cols;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 256 threads: 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? s[static_cast<int64_t>(r) * cols + c] : 0.0f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < cols && r < rows) d[static_cast<int64_t>(c) * rows + r] = tile[tx][i];
  }
}

// ---------------------------------------------------------------- row gather / scatter (d % 8 == 0)
__global__ void gather_rows_kernel(const bf16* __restrict__ x, const int32_t* __restrict__ idx, bf16* __restrict__ y, int n_idx, int d, int scatter) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x;
  if (i >= n_idx) return;
  const int64_t src = scatter ? i : idx[i];
  const int64_t dst = scatter ? idx[i] : i;
  const bf16x8* s = reinterpret_cast<const bf16x8*>(x + src * d);
  bf16x8* t = reinterpret_cast<bf16x8*>(y + dst * d);
  for (int j = threadIdx.x; j < d / 8; j += blockDim.x) t[j] = s[j];
}

// ---------------------------------------------------------------- column sums (bias grads)
__global__ void colsum_kernel(const bf16* __restrict__ x, int64_t ldx, int rows, int cols, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  // block = 32 columns x 8 row-lanes; grid.y splits rows; atomics merge partial sums
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  const int rows_per = static_cast<int>(ceil_div(rows, gridDim.y));
  const int r_begin = blockIdx.y * rows_per;
  const int r_end = min(rows, r_begin + rows_per);
  float acc = 0.0f;
  if (c < cols)
    for (int r = r_begin + ry; r < r_end; r += 8) acc += __bfloat162float(x[static_cast<int64_t>(r) * ldx + c]);
  part[ry][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += part[k][threadIdx.x & 31];
    atomicAdd(out + c, s);
  }
}
// RMSNorm weight gradient (full fine-tune, freeze_llm=false): dw[c] += sum_r dy[r,c] * x[r,c] * rstd[r]   (y = w * x * rstd)
__global__ void rmsnorm_wgrad_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ rstd, int rows, int cols,
                                     float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  const int rows_per = static_cast<int>(ceil_div(rows, gridDim.y));
  const int r_begin = blockIdx.y * rows_per;
  const int r_end = min(rows, r_begin + rows_per);
  float acc = 0.0f;
  if (c < cols)
    for (int r = r_begin + ry; r < r_end; r += 8) {
      const int64_t i = static_cast<int64_t>(r) * cols + c;
      acc += __bfloat162float(dy[i]) * __bfloat162float(x[i]) * rstd[r];
    }
  part[ry][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += part[k][threadIdx.x & 31];
    atomicAdd(out + c, s);
  }
}

// Embedding-table gradient of the merge (full fine-tune): dE[max(ids[r], 0)] += dx[r] for every row whose modality mask is 0 — the rows
// that took `embed[ids] * (~mask)` in slam_model.py:392.  fp32 atomics (token ids repeat); one block per row.
__global__ void embed_grad_kernel(const int64_t* __restrict__ ids, const uint8_t* __restrict__ mask, const bf16* __restrict__ dx, float* __restrict__ de,
                                  int d, int vocab) {
  pdl_trigger();
  pdl_wait();
  const int64_t r = blockIdx.x;
  if (mask[r]) return;
  int64_t id = ids[r];
  if (id < 0) id = 0;
  if (id >= vocab) return;
  const bf16* src = dx + r * d;
  float* dst = de + id * d;
  for (int k = threadIdx.x; k < d; k += blockDim.x) atomicAdd(dst + k, __bfloat162float(src[k]));
}

__global__ void zero_f32_kernel(float* p, int64_t n) {
  pdl_trigger();
  pdl_wait();
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) p[i] = 0.0f;
}

// ---------------------------------------------------------------- RMSNorm (HF LlamaRMSNorm; modeling_llama.py:62-67)
template <int VPT>  // 8-element vectors per thread held in registers
__global__ void __launch_bounds__(256) rmsnorm_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ y,
                                                          float* __restrict__ rstd_out, int d, float eps) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sbuf[32];
  const int64_t row = blockIdx.x;
  const bf16x8* xr = reinterpret_cast<const bf16x8*>(x + row * d);
  const int nvec = d / 8;
  float v[VPT][8];
  float ss = 0.0f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int j = threadIdx.x + i * blockDim.x;
    if (j < nvec) {
      unpack8(xr[j], v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) ss += v[i][e] * v[i][e];
    }
  }
  ss = block_sum(ss, sbuf);
  const float rstd = rsqrtf(ss / static_cast<float>(d) + eps);
  if (threadIdx.x == 0 && rstd_out != nullptr) rstd_out[row] = rstd;
  const bf16x8* wr = reinterpret_cast<const bf16x8*>(w);
  bf16x8* yr = reinterpret_cast<bf16x8*>(y + row * d);
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int j = threadIdx.x + i * blockDim.x;
    if (j < nvec) {
      float wf[8], o[8];
      unpack8(wr[j], wf);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = v[i][e] * rstd * wf[e];
      yr[j] = pack8(o);
    }
  }
}

template <int VPT>
__global__ void __launch_bounds__(256) rmsnorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                          const float* __restrict__ rstd_in, const bf16* __restrict__ dres,
                                                          bf16* __restrict__ dx, int d) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sbuf[32];
  const int64_t row = blockIdx.x;
  const int nvec = d / 8;
  const bf16x8* xr = reinterpret_cast<const bf16x8*>(x + row * d);
  const bf16x8* gr = reinterpret_cast<const bf16x8*>(dy + row * d);
  const bf16x8* wr = reinterpret_cast<const bf16x8*>(w);
  float xv[VPT][8], gv[VPT][8];
  float dot = 0.0f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int j = threadIdx.x + i * blockDim.x;
    if (j < nvec) {
      float wf[8];
      unpack8(xr[j], xv[i]);
      unpack8(gr[j], gv[i]);
      unpack8(wr[j], wf);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        gv[i][e] *= wf[e];
        dot += gv[i][e] * xv[i][e];
      }
    }
  }
  dot = block_sum(dot, sbuf);
  const float rstd = rstd_in[row];
  const float coef = dot * rstd * rstd * rstd / static_cast<float>(d);
  bf16x8* dxr = reinterpret_cast<bf16x8*>(dx + row * d);
  const bf16x8* rr = dres != nullptr ? reinterpret_cast<const bf16x8*>(dres + row * d) : nullptr;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int j = threadIdx.x + i * blockDim.x;
    if (j < nvec) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = gv[i][e] * rstd - xv[i][e] * coef;
      if (rr != nullptr) {
        float rf[8];
        unpack8(rr[j], rf);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] += rf[e];
      }
      dxr[j] = pack8(o);
    }
  }
}

// ---------------------------------------------------------------- LayerNorm (whisper LayerNorm: fp32 statistics)
template <int VPT>
__global__ void __launch_bounds__(256) layernorm_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                        bf16* __restrict__ y, int d, float eps) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sbuf[32];
  const int64_t row = blockIdx.x;
  const bf16x8* xr = reinterpret_cast<const bf16x8*>(x + row * d);
  const int nvec = d / 8;
  float v[VPT][8];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int j = threadIdx.x + i * blockDim.x;
    if (j < nvec) {
      unpack8(xr[j], v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[i][e];
    }
  }
  const float mean = block_sum(s, sbuf) / static_cast<float>(d);
  float ss = 0.0f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int j = threadIdx.x + i * blockDim.x;
    if (j < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float t = v[i][e] - mean;
        ss += t * t;
      }
    }
  }
  const float rstd = rsqrtf(block_sum(ss, sbuf) / static_cast<float>(d) + eps);
  bf16x8* yr = reinterpret_cast<bf16x8*>(y + row * d);
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int j = threadIdx.x + i * blockDim.x;
    if (j < nvec) {
      float o[8];
      const float4 w0 = *reinterpret_cast<const float4*>(w + j * 8), w1 = *reinterpret_cast<const float4*>(w + j * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(b + j * 8), b1 = *reinterpret_cast<const float4*>(b + j * 8 + 4);
      const float wf[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float bf[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[i][e] - mean) * rstd * wf[e] + bf[e];
      yr[j] = pack8(o);
    }
  }
}

// ---------------------------------------------------------------- warp-per-row norms (no block barriers)
// One warp owns one row; 8 rows per CTA.  Pass 1 reduces with warp shuffles, later passes re-read the row from L1
// (a row is 2.5-8 KB), so HBM traffic stays one read + one write per element.  Replaces the one-CTA-per-row kernels above,
// which spent most of their time in two __syncthreads per row (15 us for 6000 x 1280 = 30 % of the HBM roofline).
__global__ void __launch_bounds__(256) rmsnorm_fwd_warp_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ y,
                                                               float* __restrict__ rstd_out, int rows, int d, float eps) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const bf16x8* xr = reinterpret_cast<const bf16x8*>(x + row * d);
  const int nvec = d / 8;
  float ss = 0.0f;
  for (int j = lane; j < nvec; j += 32) {
    float v[8];
    unpack8(xr[j], v);
#pragma unroll
    for (int e = 0; e < 8; ++e) ss += v[e] * v[e];
  }
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss / static_cast<float>(d) + eps);
  if (lane == 0 && rstd_out != nullptr) rstd_out[row] = rstd;
  const bf16x8* wr = reinterpret_cast<const bf16x8*>(w);
  bf16x8* yr = reinterpret_cast<bf16x8*>(y + row * d);
  for (int j = lane; j < nvec; j += 32) {
    float v[8], wf[8], o[8];
    unpack8(xr[j], v);
    unpack8(wr[j], wf);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = v[e] * rstd * wf[e];
    yr[j] = pack8(o);
  }
}

__global__ void __launch_bounds__(256) rmsnorm_bwd_warp_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                               const float* __restrict__ rstd_in, const bf16* __restrict__ dres,
                                                               bf16* __restrict__ dx, int rows, int d) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = d / 8;
  const bf16x8* xr = reinterpret_cast<const bf16x8*>(x + row * d);
  const bf16x8* gr = reinterpret_cast<const bf16x8*>(dy + row * d);
  const bf16x8* wr = reinterpret_cast<const bf16x8*>(w);
  float dot = 0.0f;
  for (int j = lane; j < nvec; j += 32) {
    float xv[8], gv[8], wf[8];
    unpack8(xr[j], xv);
    unpack8(gr[j], gv);
    unpack8(wr[j], wf);
#pragma unroll
    for (int e = 0; e < 8; ++e) dot += gv[e] * wf[e] * xv[e];
  }
  dot = warp_sum(dot);
  const float rstd = rstd_in[row];
  const float coef = dot * rstd * rstd * rstd / static_cast<float>(d);
  bf16x8* dxr = reinterpret_cast<bf16x8*>(dx + row * d);
  const bf16x8* rr = dres != nullptr ? reinterpret_cast<const bf16x8*>(dres + row * d) : nullptr;
  for (int j = lane; j < nvec; j += 32) {
    float xv[8], gv[8], wf[8], o[8];
    unpack8(xr[j], xv);
    unpack8(gr[j], gv);
    unpack8(wr[j], wf);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = gv[e] * wf[e] * rstd - xv[e] * coef;
    if (rr != nullptr) {
      float rf[8];
      unpack8(rr[j], rf);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] += rf[e];
    }
    dxr[j] = pack8(o);
  }
}

__global__ void __launch_bounds__(256) layernorm_warp_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                             bf16* __restrict__ y, int rows, int d, float eps) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const bf16x8* xr = reinterpret_cast<const bf16x8*>(x + row * d);
  const int nvec = d / 8;
  float s = 0.0f;
  for (int j = lane; j < nvec; j += 32) {
    float v[8];
    unpack8(xr[j], v);
#pragma unroll
    for (int e = 0; e < 8; ++e) s += v[e];
  }
  const float mean = warp_sum(s) / static_cast<float>(d);
  float ss = 0.0f;
  for (int j = lane; j < nvec; j += 32) {
    float v[8];
    unpack8(xr[j], v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float t = v[e] - mean;
      ss += t * t;
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(d) + eps);
  bf16x8* yr = reinterpret_cast<bf16x8*>(y + row * d);
  for (int j = lane; j < nvec; j += 32) {
    float v[8], o[8];
    unpack8(xr[j], v);
    const float4 w0 = *reinterpret_cast<const float4*>(w + j * 8), w1 = *reinterpret_cast<const float4*>(w + j * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(b + j * 8), b1 = *reinterpret_cast<const float4*>(b + j * 8 + 4);
    const float wf[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    const float bf[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = (v[e] - mean) * rstd * wf[e] + bf[e];
    yr[j] = pack8(o);
  }
}

// ---------------------------------------------------------------- register-resident norms for the common hidden sizes
// One row = TPR threads, each holding NV 16-byte vectors of the row in registers: every operand is read from memory exactly
// once, and a 1604 x 4096 activation still spreads over 6416 warps (the warp-per-row kernels above leave the GPU at ~11 warps
// per SM for that shape and re-read the row from L1 for every pass: 26 us for rmsnorm_bwd against ~6 us of memory time).
template <int TPR>
__device__ __forceinline__ float row_sum(float v, float* s_red) {
  v = warp_sum(v);
  if constexpr (TPR > 32) {
    constexpr int WPR = TPR / 32;
    const int warp = threadIdx.x >> 5;
    __syncthreads();                                  // s_red may still be read from a previous reduction
    if ((threadIdx.x & 31) == 0) s_red[warp] = v;
    __syncthreads();
    const int first = (warp / WPR) * WPR;
    v = 0.0f;
#pragma unroll
    for (int i = 0; i < WPR; ++i) v += s_red[first + i];
  }
  return v;
}

template <int NV, int TPR>
__global__ void __launch_bounds__(256) rmsnorm_fwd_reg_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ y,
                                                              float* __restrict__ rstd_out, int rows, float eps) {
  pdl_trigger();
  pdl_wait();
  constexpr int D = NV * TPR * 8, RPC = 256 / TPR;
  __shared__ float s_red[8];
  pdl_trigger();
  const int t = threadIdx.x % TPR;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * RPC + threadIdx.x / TPR;
  const bool ok = row < rows;
  const bf16x8* xr = reinterpret_cast<const bf16x8*>(x + (ok ? row : rows - 1) * D);
  bf16x8 xv[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) xv[j] = xr[j * TPR + t];
  float ss = 0.0f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float v[8];
    unpack8(xv[j], v);
#pragma unroll
    for (int e = 0; e < 8; ++e) ss += v[e] * v[e];
  }
  ss = row_sum<TPR>(ss, s_red);
  const float rstd = rsqrtf(ss / static_cast<float>(D) + eps);
  if (!ok) return;
  if (t == 0 && rstd_out != nullptr) rstd_out[row] = rstd;
  const bf16x8* wr = reinterpret_cast<const bf16x8*>(w);
  bf16x8* yr = reinterpret_cast<bf16x8*>(y + row * D);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float v[8], wf[8], o[8];
    unpack8(xv[j], v);
    unpack8(wr[j * TPR + t], wf);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = v[e] * rstd * wf[e];
    yr[j * TPR + t] = pack8(o);
  }
}

template <int NV, int TPR>
__global__ void __launch_bounds__(256) rmsnorm_bwd_reg_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                              const float* __restrict__ rstd_in, const bf16* __restrict__ dres,
                                                              bf16* __restrict__ dx, int rows) {
  pdl_trigger();
  pdl_wait();
  constexpr int D = NV * TPR * 8, RPC = 256 / TPR;
  __shared__ float s_red[8];
  pdl_trigger();
  const int t = threadIdx.x % TPR;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * RPC + threadIdx.x / TPR;
  const bool ok = row < rows;
  const int64_t rc = ok ? row : rows - 1;
  const bf16x8* xr = reinterpret_cast<const bf16x8*>(x + rc * D);
  const bf16x8* gr = reinterpret_cast<const bf16x8*>(dy + rc * D);
  const bf16x8* wr = reinterpret_cast<const bf16x8*>(w);
  bf16x8 xv[NV], gv[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    xv[j] = xr[j * TPR + t];
    gv[j] = gr[j * TPR + t];
  }
  float gw[NV][8];                                    // dy * w, reused for the output
  float dot = 0.0f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float xf[8], gf[8], wf[8];
    unpack8(xv[j], xf);
    unpack8(gv[j], gf);
    unpack8(wr[j * TPR + t], wf);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      gw[j][e] = gf[e] * wf[e];
      dot += gw[j][e] * xf[e];
    }
  }
  dot = row_sum<TPR>(dot, s_red);
  if (!ok) return;
  const float rstd = rstd_in[row];
  const float coef = dot * rstd * rstd * rstd / static_cast<float>(D);
  bf16x8* dxr = reinterpret_cast<bf16x8*>(dx + row * D);
  const bf16x8* rr = dres != nullptr ? reinterpret_cast<const bf16x8*>(dres + row * D) : nullptr;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float xf[8], o[8];
    unpack8(xv[j], xf);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = gw[j][e] * rstd - xf[e] * coef;
    if (rr != nullptr) {
      float rf[8];
      unpack8(rr[j * TPR + t], rf);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] += rf[e];
    }
    dxr[j * TPR + t] = pack8(o);
  }
}

template <int NV, int TPR>
__global__ void __launch_bounds__(256) layernorm_reg_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                            bf16* __restrict__ y, int rows, float eps) {
  pdl_trigger();
  pdl_wait();
  constexpr int D = NV * TPR * 8, RPC = 256 / TPR;
  __shared__ float s_red[8];
  pdl_trigger();
  const int t = threadIdx.x % TPR;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * RPC + threadIdx.x / TPR;
  const bool ok = row < rows;
  const bf16x8* xr = reinterpret_cast<const bf16x8*>(x + (ok ? row : rows - 1) * D);
  float xf[NV][8];
  float sum = 0.0f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    unpack8(xr[j * TPR + t], xf[j]);
#pragma unroll
    for (int e = 0; e < 8; ++e) sum += xf[j][e];
  }
  const float mean = row_sum<TPR>(sum, s_red) / static_cast<float>(D);
  float ss = 0.0f;
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float c = xf[j][e] - mean;
      ss += c * c;
    }
  const float rstd = rsqrtf(row_sum<TPR>(ss, s_red) / static_cast<float>(D) + eps);
  if (!ok) return;
  bf16x8* yr = reinterpret_cast<bf16x8*>(y + row * D);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c0 = (j * TPR + t) * 8;
    const float4 w0 = *reinterpret_cast<const float4*>(w + c0), w1 = *reinterpret_cast<const float4*>(w + c0 + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(b + c0), b1 = *reinterpret_cast<const float4*>(b + c0 + 4);
    const float wf[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    const float bf[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = (xf[j][e] - mean) * rstd * wf[e] + bf[e];
    yr[j * TPR + t] = pack8(o);
  }
}

// hidden sizes with a register-resident instance: d -> (NV, TPR)
#define SLAM_NORM_DISPATCH(D_, LAUNCH)            \
  switch (D_) {                                   \
    case 4096: LAUNCH(4, 128); break;             \
    case 2048: LAUNCH(4, 64); break;              \
    case 1280: LAUNCH(5, 32); break;              \
    case 1024: LAUNCH(4, 32); break;              \
    case 768: LAUNCH(3, 32); break;               \
    case 512: LAUNCH(2, 32); break;               \
    default: break;                               \
  }

// ---------------------------------------------------------------- RoPE (HF apply_rotary_pos_emb, rotate_half; modeling_llama.py:138-168)
// x viewed as [rows, n_heads, dh] with row stride ld; cos/sin f32 [seq_len, dh/2]; 8 pairs per thread
__global__ void rope_kernel(bf16* __restrict__ x, int64_t ld, int rows, int seq_len, int n_heads, int dh, const float* __restrict__ cosT,
                            const float* __restrict__ sinT, int inverse) {
  pdl_trigger();
  pdl_wait();
  const int half = dh / 2;
  const int vec_per_head = half / 8;
  const int64_t total = static_cast<int64_t>(rows) * n_heads * vec_per_head;
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const int v = static_cast<int>(i % vec_per_head);
    const int h = static_cast<int>((i / vec_per_head) % n_heads);
    const int64_t r = i / (static_cast<int64_t>(vec_per_head) * n_heads);
    const int pos = static_cast<int>(r % seq_len);
    bf16* base = x + r * ld + static_cast<int64_t>(h) * dh + v * 8;
    float lo[8], hi[8];
    unpack8(*reinterpret_cast<const bf16x8*>(base), lo);
    unpack8(*reinterpret_cast<const bf16x8*>(base + half), hi);
    const float* c = cosT + static_cast<int64_t>(pos) * half + v * 8;
    const float* s = sinT + static_cast<int64_t>(pos) * half + v * 8;
    float olo[8], ohi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float cs = c[e];
      const float sn = inverse ? -s[e] : s[e];
      olo[e] = lo[e] * cs - hi[e] * sn;
      ohi[e] = hi[e] * cs + lo[e] * sn;
    }
    *reinterpret_cast<bf16x8*>(base) = pack8(olo);
    *reinterpret_cast<bf16x8*>(base + half) = pack8(ohi);
  }
}

// ---------------------------------------------------------------- SwiGLU (HF LlamaMLP; modeling_llama.py:182-184)
// column of the gate of feature c (a multiple of 8) inside a [rows, 2F] gate/up matrix, and the distance to its up partner
__device__ __forceinline__ void swiglu_cols(int c, int f, int block, int& gcol, int& udist) {
  if (block == 0) {
    gcol = c;
    udist = f;
  } else {                                           // blocked: [block gates | block ups] per group of `block` features
    gcol = (c / block) * 2 * block + (c % block);
    udist = block;
  }
}
__global__ void swiglu_fwd_kernel(const bf16* __restrict__ gu, bf16* __restrict__ h, int rows, int f, int block) {
  pdl_trigger();
  pdl_wait();
  const int vpr = f / 8;
  const int64_t total = static_cast<int64_t>(rows) * vpr;
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const int64_t r = i / vpr;
    const int c = static_cast<int>(i % vpr) * 8;
    float g[8], u[8], o[8];
    int gcol, udist;
    swiglu_cols(c, f, block, gcol, udist);
    unpack8(*reinterpret_cast<const bf16x8*>(gu + r * 2 * f + gcol), g);
    unpack8(*reinterpret_cast<const bf16x8*>(gu + r * 2 * f + gcol + udist), u);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = g[e] / (1.0f + expf(-g[e])) * u[e];
    *reinterpret_cast<bf16x8*>(h + r * f + c) = pack8(o);
  }
}
__global__ void swiglu_bwd_kernel(const bf16* __restrict__ gu, const bf16* __restrict__ dh, bf16* __restrict__ dgu, int rows, int f, int block) {
  pdl_trigger();
  pdl_wait();
  const int vpr = f / 8;
  const int64_t total = static_cast<int64_t>(rows) * vpr;
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; i < total; i += stride) {
    const int64_t r = i / vpr;
    const int c = static_cast<int>(i % vpr) * 8;
    float g[8], u[8], d[8], dg[8], du[8];
    int gcol, udist;
    swiglu_cols(c, f, block, gcol, udist);
    unpack8(*reinterpret_cast<const bf16x8*>(gu + r * 2 * f + gcol), g);
    unpack8(*reinterpret_cast<const bf16x8*>(gu + r * 2 * f + gcol + udist), u);
    unpack8(*reinterpret_cast<const bf16x8*>(dh + r * f + c), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float sg = 1.0f / (1.0f + expf(-g[e]));
      const float silu = g[e] * sg;
      du[e] = d[e] * silu;
      dg[e] = d[e] * u[e] * sg * (1.0f + g[e] * (1.0f - sg));
    }
    *reinterpret_cast<bf16x8*>(dgu + r * 2 * f + gcol) = pack8(dg);
    *reinterpret_cast<bf16x8*>(dgu + r * 2 * f + gcol + udist) = pack8(du);
  }
}

// ---------------------------------------------------------------- embedding gather + modality merge (models/slam_model.py:370-392)
// one block per token row; start/len recomputed per block from the (tiny) mask row — no host sync (.tolist()) needed
__device__ __forceinline__ void mask_span(const uint8_t* __restrict__ mrow, int s, int ta, int* sh, int& start, int& len) {
  // sh: 2 ints of shared memory
  if (threadIdx.x == 0) {
    sh[0] = 0x7fffffff;
    sh[1] = 0;
  }
  __syncthreads();
  int first = 0x7fffffff, cnt = 0;
  for (int j = threadIdx.x; j < s; j += blockDim.x)
    if (mrow[j]) {
      first = min(first, j);
      ++cnt;
    }
  for (int o = 16; o > 0; o >>= 1) {
    first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&sh[0], first);
    atomicAdd(&sh[1], cnt);
  }
  __syncthreads();
  start = sh[0] == 0x7fffffff ? 0 : sh[0];  // argmax of an all-false mask is 0
  len = min(sh[1], ta);
}

__global__ void embed_merge_kernel(const int64_t* __restrict__ ids, const uint8_t* __restrict__ mask, const bf16* __restrict__ audio, int ta,
                                   const bf16* __restrict__ embed, bf16* __restrict__ x, int s, int d) {
  pdl_trigger();
  pdl_wait();
  __shared__ int sh[2];
  const int b = blockIdx.y, r = blockIdx.x;
  const uint8_t* mrow = mask + static_cast<int64_t>(b) * s;
  int start, len;
  mask_span(mrow, s, ta, sh, start, len);
  const bool is_audio = r >= start && r < start + len;
  const bool is_text = mrow[r] == 0;
  int64_t id = ids[static_cast<int64_t>(b) * s + r];
  if (id < 0) id = 0;  // input_ids[input_ids == -1] = 0
  const bf16x8* arow = reinterpret_cast<const bf16x8*>(audio + (static_cast<int64_t>(b) * ta + (r - start)) * d);
  const bf16x8* erow = reinterpret_cast<const bf16x8*>(embed + id * d);
  bf16x8* out = reinterpret_cast<bf16x8*>(x + (static_cast<int64_t>(b) * s + r) * d);
  for (int j = threadIdx.x; j < d / 8; j += blockDim.x) {
    float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (is_audio) unpack8(arow[j], o);
    if (is_text) {
      float e[8];
      unpack8(erow[j], e);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] += e[k];
    }
    out[j] = pack8(o);
  }
}
__global__ void embed_merge_bwd_kernel(const uint8_t* __restrict__ mask, const bf16* __restrict__ dx, bf16* __restrict__ daudio, int ta, int s, int d) {
  pdl_trigger();
  pdl_wait();
  __shared__ int sh[2];
  const int b = blockIdx.y, j = blockIdx.x;  // j: audio row
  int start, len;
  mask_span(mask + static_cast<int64_t>(b) * s, s, ta, sh, start, len);
  bf16x8* out = reinterpret_cast<bf16x8*>(daudio + (static_cast<int64_t>(b) * ta + j) * d);
  const bf16x8* src = reinterpret_cast<const bf16x8*>(dx + (static_cast<int64_t>(b) * s + start + j) * d);
  bf16x8 z;
  z.w[0] = z.w[1] = z.w[2] = z.w[3] = 0u;
  for (int k = threadIdx.x; k < d / 8; k += blockDim.x) out[k] = j < len ? src[k] : z;
}

// ---------------------------------------------------------------- conv stem helpers (models/encoder.py:18-24)
// col[b, t, kk*C + c] = x[b, stride*t + kk - 1, c]; one block per output row; zero fill outside and for cols >= 3C
template <typename TIn>
__global__ void im2col_kernel(const TIn* __restrict__ x, int t_in, int c, int stride, int t_out, bf16* __restrict__ col, int64_t ldk) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y, t = blockIdx.x;
  bf16* out = col + (static_cast<int64_t>(b) * t_out + t) * ldk;
  for (int j = threadIdx.x; j < ldk; j += blockDim.x) {
    float v = 0.0f;
    if (j < 3 * c) {
      const int kk = j / c, cc = j - kk * c;
      const int ti = stride * t + kk - 1;
      if (ti >= 0 && ti < t_in) {
        const TIn raw = x[(static_cast<int64_t>(b) * t_in + ti) * c + cc];
        v = static_cast<float>(raw);
      }
    }
    out[j] = __float2bfloat16(v);
  }
}
__global__ void add_pos_kernel(bf16* __restrict__ x, const float* __restrict__ pos, int t, int d, int64_t total_vec) {
  pdl_trigger();
  pdl_wait();
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int vpr = d / 8;
  for (; i < total_vec; i += stride) {
    const int64_t row = i / vpr;
    const int c = static_cast<int>(i % vpr) * 8;
    const int tt = static_cast<int>(row % t);
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(x + row * d + c), f);
    const float* p = pos + static_cast<int64_t>(tt) * d + c;
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] += p[e];
    *reinterpret_cast<bf16x8*>(x + row * d + c) = pack8(f);
  }
}

// ---------------------------------------------------------------- LoRA-branch dropout (peft lora.Linear: lora_A(dropout(x)))
// Counter-based mask: keep(i) = hash(seed, i) >= p * 2^32 — the same function regenerates the mask in the backward, so no
// mask tensor is stored.  (The RNG stream differs from torch's Philox: only the distribution is reproduced.)
__device__ __forceinline__ uint32_t mix32(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return static_cast<uint32_t>((z ^ (z >> 31)) >> 16);
}
// y = x * keep / (1 - p)
__global__ void dropout_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int64_t n, uint32_t thresh, float inv_keep, uint64_t seed) {
  pdl_trigger();
  pdl_wait();
  int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(x + i), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = mix32(seed, static_cast<uint64_t>(i + e)) >= thresh ? f[e] * inv_keep : 0.0f;
    *reinterpret_cast<bf16x8*>(y + i) = pack8(f);
  }
}
// out = base + lora * keep / (1 - p)   (dX of a LoRA linear under dropout: dY W + mask o ((dY sB) A))
__global__ void dropout_bwd_add_kernel(const bf16* __restrict__ base, const bf16* __restrict__ lora, bf16* __restrict__ out, int64_t n, uint32_t thresh,
                                       float inv_keep, uint64_t seed) {
  pdl_trigger();
  pdl_wait();
  int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    float fb[8], fl[8];
    unpack8(*reinterpret_cast<const bf16x8*>(base + i), fb);
    unpack8(*reinterpret_cast<const bf16x8*>(lora + i), fl);
#pragma unroll
    for (int e = 0; e < 8; ++e) fb[e] += mix32(seed, static_cast<uint64_t>(i + e)) >= thresh ? fl[e] * inv_keep : 0.0f;
    *reinterpret_cast<bf16x8*>(out + i) = pack8(fb);
  }
}

// ---------------------------------------------------------------- AdamW (torch.optim.AdamW semantics, single tensor, fp32)
__device__ __forceinline__ void adamw_elem(float& pv, float gi, float& mi, float& vi, float lr, float beta1, float beta2, float eps, float wd,
                                           float bc1, float bc2_sqrt, float grad_div) {
  const float grad = gi / grad_div;
  pv *= (1.0f - lr * wd);
  mi = beta1 * mi + (1.0f - beta1) * grad;
  vi = beta2 * vi + (1.0f - beta2) * grad * grad;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  pv -= (lr / bc1) * (mi / denom);
}
// 28 B of HBM traffic per parameter: 16-byte vectors (the arena buffers are 16-byte aligned), scalar tail
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                             float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt, float grad_div) {
  pdl_trigger();
  pdl_wait();
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const int64_t n4 = vec ? n / 4 : 0;
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adamw_elem(pv.x, gv.x, mv.x, vv.x, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, grad_div);
    adamw_elem(pv.y, gv.y, mv.y, vv.y, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, grad_div);
    adamw_elem(pv.z, gv.z, mv.z, vv.z, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, grad_div);
    adamw_elem(pv.w, gv.w, mv.w, vv.w, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, grad_div);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (i = 4 * n4 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    adamw_elem(p[i], g[i], m[i], v[i], lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, grad_div);
}

}  // namespace slam

using namespace slam;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<bf16*>(p)
#define CBF(p) reinterpret_cast<const bf16*>(p)

extern "C" {

int slam_cast_f32_to_bf16(const float* x, void* y, int64_t n, float scale, void* stream) {
  SLAM_CHECK_ARG(n >= 0, "cast: n < 0");
  if (n == 0) return 0;
  launch_pdl(cast_f32_bf16_kernel, ew_grid(n, 8, 256), 256, 0, ST(stream), x, BF(y), n, scale);
  SLAM_LAUNCH_CHECK("slam_cast_f32_to_bf16");
  return 0;
}
int slam_cast_bf16_to_f32(const void* x, float* y, int64_t n, void* stream) {
  if (n == 0) return 0;
  launch_pdl(cast_bf16_f32_kernel, ew_grid(n, 8, 256), 256, 0, ST(stream), CBF(x), y, n);
  SLAM_LAUNCH_CHECK("slam_cast_bf16_to_f32");
  return 0;
}
int slam_add_bf16(const void* a, const void* b, void* y, int64_t n, void* stream) {
  if (n == 0) return 0;
  launch_pdl(add_bf16_kernel, ew_grid(n, 8, 256), 256, 0, ST(stream), CBF(a), CBF(b), BF(y), n);
  SLAM_LAUNCH_CHECK("slam_add_bf16");
  return 0;
}
int slam_relu_bwd(const void* dy, const void* y, void* dx, int64_t n, void* stream) {
  if (n == 0) return 0;
  launch_pdl(relu_bwd_kernel, ew_grid(n, 8, 256), 256, 0, ST(stream), CBF(dy), CBF(y), BF(dx), n);
  SLAM_LAUNCH_CHECK("slam_relu_bwd");
  return 0;
}
int slam_transpose_bf16(const void* x, int64_t ldx, void* y, int64_t ldy, int32_t rows, int32_t cols, void* stream) {
  SLAM_CHECK_ARG(rows > 0 && cols > 0, "transpose: bad shape");
  dim3 grid(static_cast<unsigned>(ceil_div(cols, 64)), static_cast<unsigned>(ceil_div(rows, 64)));
  launch_pdl(transpose_bf16_kernel, grid, 256, 0, ST(stream), CBF(x), ldx, BF(y), ldy, rows, cols);
  SLAM_LAUNCH_CHECK("slam_transpose_bf16");
  return 0;
}
int slam_transpose_f32_batched(const float* src, float* dst, int32_t batch, int32_t rows, int32_t cols, void* stream) {
  SLAM_CHECK_ARG(batch > 0 && rows > 0 && cols > 0 && batch <= 65535, "transpose_f32_batched: bad shape");
  dim3 grid(static_cast<unsigned>(ceil_div(cols, 32)), static_cast<unsigned>(ceil_div(rows, 32)), batch);
  launch_pdl(transpose_f32_batched_kernel, grid, 256, 0, ST(stream), src, dst, rows, cols);
  SLAM_LAUNCH_CHECK("slam_transpose_f32_batched");
  return 0;
}
int slam_gather_rows(const void* x, const int32_t* idx, void* y, int32_t n_idx, int32_t d, void* stream) {
  SLAM_CHECK_ARG(d % 8 == 0, "gather_rows: d %% 8 != 0");
  if (n_idx == 0) return 0;
  launch_pdl(gather_rows_kernel, n_idx, 128, 0, ST(stream), CBF(x), idx, BF(y), n_idx, d, 0);
  SLAM_LAUNCH_CHECK("slam_gather_rows");
  return 0;
}
int slam_scatter_rows(const void* x, const int32_t* idx, void* y, int32_t n_idx, int32_t d, void* stream) {
  SLAM_CHECK_ARG(d % 8 == 0, "scatter_rows: d %% 8 != 0");
  if (n_idx == 0) return 0;
  launch_pdl(gather_rows_kernel, n_idx, 128, 0, ST(stream), CBF(x), idx, BF(y), n_idx, d, 1);
  SLAM_LAUNCH_CHECK("slam_scatter_rows");
  return 0;
}
int slam_colsum(const void* x, int64_t ldx, int32_t rows, int32_t cols, float* out, void* stream) {
  SLAM_CHECK_ARG(rows > 0 && cols > 0, "colsum: bad shape");
  launch_pdl(zero_f32_kernel, ew_grid(cols, 1, 256), 256, 0, ST(stream), out, cols);
  SLAM_LAUNCH_CHECK("slam_colsum.zero");
  int ysplit = static_cast<int>(ceil_div(rows, 256));
  if (ysplit > 64) ysplit = 64;
  dim3 grid(static_cast<unsigned>(ceil_div(cols, 32)), static_cast<unsigned>(ysplit));
  launch_pdl(colsum_kernel, grid, 256, 0, ST(stream), CBF(x), ldx, rows, cols, out);
  SLAM_LAUNCH_CHECK("slam_colsum");
  return 0;
}

#define DISPATCH_VPT(d, threads, CALL)                                    \
  do {                                                                    \
    const int vpt__ = static_cast<int>(ceil_div((d) / 8, (threads)));     \
    if (vpt__ <= 1) { CALL(1); }                                          \
    else if (vpt__ <= 2) { CALL(2); }                                     \
    else if (vpt__ <= 4) { CALL(4); }                                     \
    else if (vpt__ <= 8) { CALL(8); }                                     \
    else { set_error("norm: d=%d too large", (int)(d)); return -1; }      \
  } while (0)

int slam_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int32_t rows, int32_t d, float eps, void* stream) {
  SLAM_CHECK_ARG(d % 8 == 0 && rows > 0, "rmsnorm_fwd: bad shape rows=%d d=%d", rows, d);
  bool done = false;
#define SLAM_L(NV, TPR)                                                                                                                  \
  launch_pdl(rmsnorm_fwd_reg_kernel<NV, TPR>, static_cast<unsigned>(ceil_div(rows, 256 / TPR)), 256, 0, ST(stream), CBF(x), CBF(w), BF(y), rstd, rows, eps); \
  done = true
  SLAM_NORM_DISPATCH(d, SLAM_L)
#undef SLAM_L
  if (!done) launch_pdl(rmsnorm_fwd_warp_kernel, static_cast<unsigned>(ceil_div(rows, 8)), 256, 0, ST(stream), CBF(x), CBF(w), BF(y), rstd, rows, d, eps);
  SLAM_LAUNCH_CHECK("slam_rmsnorm_fwd");
  return 0;
}
int slam_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, const void* dres, void* dx, int32_t rows, int32_t d,
                     void* stream) {
  SLAM_CHECK_ARG(d % 8 == 0 && rows > 0, "rmsnorm_bwd: bad shape rows=%d d=%d", rows, d);
  bool done = false;
#define SLAM_L(NV, TPR)                                                                                                                 \
  launch_pdl(rmsnorm_bwd_reg_kernel<NV, TPR>, static_cast<unsigned>(ceil_div(rows, 256 / TPR)), 256, 0, ST(stream), CBF(dy), CBF(x), CBF(w), rstd, CBF(dres), \
                                                                                                         BF(dx), rows);                \
  done = true
  SLAM_NORM_DISPATCH(d, SLAM_L)
#undef SLAM_L
  if (!done)
    launch_pdl(rmsnorm_bwd_warp_kernel, static_cast<unsigned>(ceil_div(rows, 8)), 256, 0, ST(stream), CBF(dy), CBF(x), CBF(w), rstd, CBF(dres), BF(dx), rows, d);
  SLAM_LAUNCH_CHECK("slam_rmsnorm_bwd");
  return 0;
}
int slam_layernorm(const void* x, const float* w, const float* b, void* y, int32_t rows, int32_t d, float eps, void* stream) {
  SLAM_CHECK_ARG(d % 8 == 0 && rows > 0, "layernorm: bad shape rows=%d d=%d", rows, d);
  bool done = false;
#define SLAM_L(NV, TPR)                                                                                                       \
  launch_pdl(layernorm_reg_kernel<NV, TPR>, static_cast<unsigned>(ceil_div(rows, 256 / TPR)), 256, 0, ST(stream), CBF(x), w, b, BF(y), rows, eps); \
  done = true
  SLAM_NORM_DISPATCH(d, SLAM_L)
#undef SLAM_L
  if (!done) launch_pdl(layernorm_warp_kernel, static_cast<unsigned>(ceil_div(rows, 8)), 256, 0, ST(stream), CBF(x), w, b, BF(y), rows, d, eps);
  SLAM_LAUNCH_CHECK("slam_layernorm");
  return 0;
}
int slam_rope(void* x, int64_t ld, int32_t rows, int32_t seq_len, int32_t n_heads, int32_t dh, const float* cos_t, const float* sin_t,
              int32_t inverse, void* stream) {
  SLAM_CHECK_ARG(dh % 16 == 0 && ld % 8 == 0, "rope: dh %% 16 != 0 or ld %% 8 != 0");
  const int64_t total = static_cast<int64_t>(rows) * n_heads * (dh / 16);
  launch_pdl(rope_kernel, ew_grid(total, 1, 256), 256, 0, ST(stream), BF(x), ld, rows, seq_len, n_heads, dh, cos_t, sin_t, inverse);
  SLAM_LAUNCH_CHECK("slam_rope");
  return 0;
}
int slam_swiglu_fwd(const void* gu, void* h, int32_t rows, int32_t f, int32_t block, void* stream) {
  SLAM_CHECK_ARG(f % 8 == 0 && (block == 0 || (block % 8 == 0 && f % block == 0)), "swiglu: f %% 8 != 0 or bad block");
  launch_pdl(swiglu_fwd_kernel, ew_grid(static_cast<int64_t>(rows) * f / 8, 1, 256), 256, 0, ST(stream), CBF(gu), BF(h), rows, f, block);
  SLAM_LAUNCH_CHECK("slam_swiglu_fwd");
  return 0;
}
int slam_swiglu_bwd(const void* gu, const void* dh, void* dgu, int32_t rows, int32_t f, int32_t block, void* stream) {
  SLAM_CHECK_ARG(f % 8 == 0 && (block == 0 || (block % 8 == 0 && f % block == 0)), "swiglu: f %% 8 != 0 or bad block");
  launch_pdl(swiglu_bwd_kernel, ew_grid(static_cast<int64_t>(rows) * f / 8, 1, 256), 256, 0, ST(stream), CBF(gu), CBF(dh), BF(dgu), rows, f, block);
  SLAM_LAUNCH_CHECK("slam_swiglu_bwd");
  return 0;
}
int slam_embed_merge(const int64_t* ids, const uint8_t* mask, const void* audio, int32_t ta, const void* embed, void* x, int32_t batch,
                     int32_t s, int32_t d, void* stream) {
  SLAM_CHECK_ARG(d % 8 == 0 && batch > 0 && s > 0, "embed_merge: bad shape");
  dim3 grid(s, batch);
  launch_pdl(embed_merge_kernel, grid, 128, 0, ST(stream), ids, mask, CBF(audio), ta, CBF(embed), BF(x), s, d);
  SLAM_LAUNCH_CHECK("slam_embed_merge");
  return 0;
}
int slam_rmsnorm_wgrad(const void* dy, const void* x, const float* rstd, int32_t rows, int32_t d, float* dw, void* stream) {
  SLAM_CHECK_ARG(rows > 0 && d > 0, "rmsnorm_wgrad: bad shape");
  int ysplit = static_cast<int>(ceil_div(rows, 256));
  if (ysplit > 64) ysplit = 64;
  dim3 grid(static_cast<unsigned>(ceil_div(d, 32)), static_cast<unsigned>(ysplit));
  launch_pdl(rmsnorm_wgrad_kernel, grid, 256, 0, ST(stream), CBF(dy), CBF(x), rstd, rows, d, dw);
  SLAM_LAUNCH_CHECK("slam_rmsnorm_wgrad");
  return 0;
}
int slam_embed_grad(const int64_t* ids, const uint8_t* mask, const void* dx, float* de, int32_t rows, int32_t d, int32_t vocab, void* stream) {
  SLAM_CHECK_ARG(rows > 0 && d > 0 && vocab > 0, "embed_grad: bad shape");
  launch_pdl(embed_grad_kernel, static_cast<unsigned>(rows), 128, 0, ST(stream), ids, mask, CBF(dx), de, d, vocab);
  SLAM_LAUNCH_CHECK("slam_embed_grad");
  return 0;
}
int slam_embed_merge_bwd(const uint8_t* mask, const void* dx, void* daudio, int32_t ta, int32_t batch, int32_t s, int32_t d, void* stream) {
  SLAM_CHECK_ARG(d % 8 == 0 && batch > 0 && s > 0 && ta > 0, "embed_merge_bwd: bad shape");
  dim3 grid(ta, batch);
  launch_pdl(embed_merge_bwd_kernel, grid, 128, 0, ST(stream), mask, CBF(dx), BF(daudio), ta, s, d);
  SLAM_LAUNCH_CHECK("slam_embed_merge_bwd");
  return 0;
}
int slam_conv_im2col(const void* x, int32_t x_is_f32, int32_t batch, int32_t t_in, int32_t c, int32_t stride, void* col, int64_t ldk,
                     void* stream) {
  SLAM_CHECK_ARG(stride == 1 || stride == 2, "im2col: stride must be 1 or 2");
  SLAM_CHECK_ARG(ldk >= 3 * c && ldk % 8 == 0, "im2col: ldk must be >= 3C and a multiple of 8");
  const int t_out = (t_in + 2 - 3) / stride + 1;
  dim3 grid(t_out, batch);
  if (x_is_f32)
    launch_pdl(im2col_kernel<float>, grid, 256, 0, ST(stream), reinterpret_cast<const float*>(x), t_in, c, stride, t_out, BF(col), ldk);
  else
    launch_pdl(im2col_kernel<bf16>, grid, 256, 0, ST(stream), CBF(x), t_in, c, stride, t_out, BF(col), ldk);
  SLAM_LAUNCH_CHECK("slam_conv_im2col");
  return 0;
}
int slam_add_pos(void* x, const float* pos, int32_t batch, int32_t t, int32_t d, void* stream) {
  SLAM_CHECK_ARG(d % 8 == 0, "add_pos: d %% 8 != 0");
  const int64_t total = static_cast<int64_t>(batch) * t * d / 8;
  launch_pdl(add_pos_kernel, ew_grid(total, 1, 256), 256, 0, ST(stream), BF(x), pos, t, d, total);
  SLAM_LAUNCH_CHECK("slam_add_pos");
  return 0;
}
static inline uint32_t drop_thresh(float p) {
  const double t = static_cast<double>(p) * 4294967296.0;
  return t >= 4294967295.0 ? 4294967295u : static_cast<uint32_t>(t);
}
int slam_dropout(const void* x, void* y, int64_t n, float p, uint64_t seed, void* stream) {
  SLAM_CHECK_ARG(n % 8 == 0 && p >= 0.0f && p < 1.0f, "dropout: n %% 8 != 0 or p outside [0,1)");
  if (n == 0) return 0;
  launch_pdl(dropout_kernel, ew_grid(n, 8, 256), 256, 0, ST(stream), CBF(x), BF(y), n, drop_thresh(p), 1.0f / (1.0f - p), seed);
  SLAM_LAUNCH_CHECK("slam_dropout");
  return 0;
}
int slam_dropout_bwd_add(const void* base, const void* lora, void* out, int64_t n, float p, uint64_t seed, void* stream) {
  SLAM_CHECK_ARG(n % 8 == 0 && p >= 0.0f && p < 1.0f, "dropout_bwd_add: n %% 8 != 0 or p outside [0,1)");
  if (n == 0) return 0;
  launch_pdl(dropout_bwd_add_kernel, ew_grid(n, 8, 256), 256, 0, ST(stream), CBF(base), CBF(lora), BF(out), n, drop_thresh(p), 1.0f / (1.0f - p), seed);
  SLAM_LAUNCH_CHECK("slam_dropout_bwd_add");
  return 0;
}
int slam_adamw(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps,
               float weight_decay, int32_t step_host, float grad_div, void* stream) {
  SLAM_CHECK_ARG(step_host >= 1 && n >= 0, "adamw: step must be >= 1");
  if (n == 0) return 0;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step_host);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step_host);
  launch_pdl(adamw_kernel, ew_grid(n, 4, 256), 256, 0, ST(stream), param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                           static_cast<float>(bc1), static_cast<float>(sqrt(bc2)), grad_div);
  SLAM_LAUNCH_CHECK("slam_adamw");
  return 0;
}

}  // extern "C"
