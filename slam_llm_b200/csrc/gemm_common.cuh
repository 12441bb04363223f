// Shared pieces of the tcgen05 GEMM kernels (gemm.cu: one CTA per tile; gemm_2cta.cu: CTA pairs, cta_group::2).
#pragma once
#include "common.cuh"

namespace slam {

constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_EPI_PITCH = 36;                                    // floats per staged row (32 + 4 pad: conflict-free)

struct GemmKParams {
  int M, N;
  int kb1, kb2;
  int ksplit, kb_per_split;   // split-K: work item = (tile, k-slice); partial tiles are merged with fp32 atomics
  int tail_tiles, tail_slices, tail_kb;   // tail split: the last tail_tiles tiles are cut into tail_slices k-slices of tail_kb k-blocks
  float* sk_partials;         // [gridDim.x] slots of 128 x BLOCK_N fp32, laid out [warp quarter][chunk][row][32 cols, 16-B pieces swizzled]
  unsigned int* sk_flags;     // [gridDim.x][4] flag = epoch of the launch whose partial (of epilogue warp q) is published
  unsigned int sk_epoch;      // distinct per launch, never 0
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

// One unit of work of a persistent CTA: k-blocks [kb_begin, kb_end) of one output tile.
//   kind 0  the accumulation is final for this CTA's purposes (whole tile, or a split-K slice merged with atomics);
//   kind 1  k-slice of a tail tile: the tail_slices CTAs of the tile exchange fp32 partials through the workspace, each
//           finishing (sum in slice order: deterministic; then the normal epilogue) its own share of the 32-column chunks.
struct GemmSeg {
  int tile, kb_begin, kb_end, kind;
};

// Work schedule, iterated identically by the producer, MMA and epilogue warps: whole tiles round-robin (tile = cta + w * G),
// then at most ONE tail item per CTA.  The tail is the last, partial wave: its R tiles are cut into `tail_slices` k-slices so
// that R * tail_slices <= G CTAs are busy instead of R.  CTAs that share a B tile stay in lockstep along K (the L2 -> SM
// operand traffic of this kernel relies on simultaneous requests for the same tile being served together; a schedule that
// desynchronises the CTAs in K - classic stream-K - measured 1.5x SLOWER per k-block on B200).
struct GemmSchedule {
  int nkb, cta, ncta;
  int dp_item, dp_items;
  bool tail_done;
  const GemmKParams* p;
  __device__ __forceinline__ GemmSchedule(const GemmKParams& prm, int cta_, int ncta_) : p(&prm) {
    nkb = prm.kb1 + prm.kb2;
    cta = cta_;
    ncta = ncta_;
    dp_item = cta;
    dp_items = (prm.num_m_tiles * prm.num_n_tiles - prm.tail_tiles) * prm.ksplit;
    tail_done = cta >= prm.tail_tiles * prm.tail_slices;
  }
  __device__ __forceinline__ bool next(GemmSeg& s) {
    if (dp_item < dp_items) {
      s.tile = dp_item / p->ksplit;
      s.kb_begin = (dp_item % p->ksplit) * p->kb_per_split;
      s.kb_end = min(nkb, s.kb_begin + p->kb_per_split);
      s.kind = 0;
      dp_item += ncta;
      return true;
    }
    if (!tail_done) {
      tail_done = true;
      const int slice = cta % p->tail_slices;
      s.tile = dp_items + cta / p->tail_slices;          // (ksplit == 1 whenever there is a tail)
      s.kb_begin = slice * p->tail_kb;
      s.kb_end = min(nkb, s.kb_begin + p->tail_kb);
      s.kind = 1;
      return true;
    }
    return false;
  }
};

__device__ __forceinline__ void sk_flag_publish(unsigned int* f, unsigned int epoch) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
}
__device__ __forceinline__ void sk_flag_wait(const unsigned int* f, unsigned int epoch) {
  unsigned int v = 0;
  long long t0 = 0;
  uint32_t spins = 0;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (v == epoch) break;
    if (++spins == 1024u) t0 = clock64();
    if (spins > 1024u && (spins & 1023u) == 0u && clock64() - t0 > 4000000000LL) {
      printf("slam_b200: tail-split flag watchdog block %d thread %d\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}
// 1-D bulk copy global -> shared (TMA engine), completion bytes on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Epilogue tail shared by both kernels: one warp has staged a 32 x 32 fp32 chunk of the accumulator (thread = row) in a padded
// shared-memory tile; it is read back as (8 rows x 4 column-pieces) per warp instruction so that global stores and residual
// loads touch whole 32-byte sectors, then alpha / bias / activation / residual are applied and 16-byte stores issued.
__device__ __forceinline__ void gemm_epilogue_store_chunk(const GemmKParams& p, const float* stg, int row_base, int col0, int lane) {
  const int piece = lane & 3;
  const int col = col0 + piece * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rr = i * 8 + (lane >> 2);
    const int row = row_base + rr;
    if (row >= p.M || col >= p.N) continue;
    const float4 x0 = *reinterpret_cast<const float4*>(stg + rr * GEMM_EPI_PITCH + piece * 8);
    const float4 x1 = *reinterpret_cast<const float4*>(stg + rr * GEMM_EPI_PITCH + piece * 8 + 4);
    float v[8] = {x0.x * p.alpha, x0.y * p.alpha, x0.z * p.alpha, x0.w * p.alpha, x1.x * p.alpha, x1.y * p.alpha, x1.z * p.alpha, x1.w * p.alpha};
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
      const float2 r0 = unpack_bf16x2(rsd.x), r1 = unpack_bf16x2(rsd.y), r2 = unpack_bf16x2(rsd.z), r3 = unpack_bf16x2(rsd.w);
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
}

}  // namespace slam
