// Shared pieces of the tcgen05 GEMM kernels (gemm.cu: one CTA per tile; gemm_2cta.cu: CTA pairs, cta_group::2).
#pragma once
#include "common.cuh"

namespace slam {

constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = 8;                                     // two warps per TMEM lane quarter, interleaved 32-column chunks
constexpr int GEMM_THREADS = 128 + 32 * GEMM_EPI_WARPS;               // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 spare, 4.. epilogue
constexpr int GEMM_EPI_PITCH = 80;                                    // bytes per staged bf16 row (64 + 16 pad: conflict-free)
constexpr int GEMM_EPI_BYTES = GEMM_EPI_WARPS * 32 * GEMM_EPI_PITCH;  // one 32 x 32 bf16 chunk per epilogue warp

struct GemmKParams {
  int M, N;
  int kb1, kb2;
  int ksplit, kb_per_split;   // split-K: work item = (tile, k-slice); partial tiles are merged with fp32 atomics
  int tail_tiles, tail_slices, tail_kb;   // tail split: the last tail_tiles tiles are cut into tail_slices k-slices of tail_kb k-blocks
  float* sk_partials;         // [gridDim.x] slots of 128 x 256 fp32, laid out [epilogue warp][its chunk j][row][32 cols, 16-B pieces swizzled]
  unsigned int* sk_flags;     // [gridDim.x][8] flag = epoch of the launch whose partial (of epilogue warp e) is published
  unsigned int sk_epoch;      // distinct per launch, never 0
  int num_m_tiles, num_n_tiles;
  void* out;
  long long ldo;
  int out_f32;
  int act;
  const float* bias;
  const bf16* residual;
  long long ldr;
  bf16* aux;                  // fused SwiGLU: act 3 -> h [M, N/2] (written); act 4 -> gu [M, 2N] (read)
  long long ld_aux;
  float alpha;
  int transpose_out;          // 1: the tile is C^T of the logical output: element (row r, col c) of the accumulator goes to out[c * ldo + r] (bf16),
                              //    the residual is read the same way.  Lets the WEIGHT be the 256-row M operand of a CTA pair (rows % 256 == 0,
                              //    no tile padding) while the token dimension becomes the flexible-width N (swap-AB, what cuBLAS does for M = 1604).
  int static_ops;             // slam_gemm_args.static_operands: bit 0 / 1 = the A / B operand may be loaded before griddepcontrol.wait
#ifdef SLAM_GEMM_TRACE
  int trace_id;               // debug build: launch number (host counter)
#endif
};

// Debug build only (SLAM_NVCC_EXTRA=-DSLAM_GEMM_TRACE, tools/gemm_trace.py): time stamps from inside the CTA-pair kernel.
// g_trace_buf[launch % g_trace_launches][cta < TRACE_CTAS][event] = {globaltimer ns, clock64}.
#ifdef SLAM_GEMM_TRACE
constexpr int TRACE_EV = 16, TRACE_CTAS = 160;
static __device__ unsigned long long* g_trace_buf = nullptr;
static __device__ int g_trace_launches = 0;
__device__ __forceinline__ void trace_ev(int launch, int ev, unsigned long long value = ~0ull) {
  if (g_trace_buf == nullptr || blockIdx.x >= TRACE_CTAS) return;
  unsigned long long gt;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
  unsigned long long* slot = g_trace_buf + ((static_cast<size_t>(launch % g_trace_launches) * TRACE_CTAS + blockIdx.x) * TRACE_EV + ev) * 2;
  slot[0] = value != ~0ull ? value : gt;
  slot[1] = static_cast<unsigned long long>(clock64());
}
#define SLAM_TRACE(ev) trace_ev(p.trace_id, ev)
#define SLAM_TRACE_V(ev, v) trace_ev(p.trace_id, ev, static_cast<unsigned long long>(v))
#else
#define SLAM_TRACE(ev) ((void)0)
#define SLAM_TRACE_V(ev, v) ((void)0)
#endif

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

// Epilogue of one 32-column chunk, shared by both kernels.  Measured on B200 (tools/epi_probe.py) the round-1 epilogue - 4 warps,
// fp32 staging, residual loaded after the transpose - took 7.7 us per 128 x 256 tile plain, 14 us with bias + GELU and 23 us with
// a residual: more than the MMA time of a K = 1280 tile (6 us), i.e. the Whisper GEMMs were epilogue-bound.  Now: 8 epilogue
// warps; all math happens in the TMEM register layout (thread = row, 32 consecutive columns), with the row's residual (64
// contiguous bytes = two full sectors) requested BEFORE the TMEM load so its latency overlaps; only the packed bf16 result is
// transposed through a padded shared-memory tile so that a warp store instruction covers 8 rows x 64 B of whole sectors.
// one warp's 32 x 32 bf16 chunk (thread = row, 16 packed pairs) -> padded shared-memory tile -> 16-byte global stores that
// cover 8 rows x 64 B of whole sectors per instruction
__device__ __forceinline__ void gemm_store_chunk_bf16(bf16* out, long long ldo, int M, int N, const uint32_t (&pk)[16], int row_base, int col0,
                                                      uint8_t* stg, int lane) {
#pragma unroll
  for (int g = 0; g < 4; ++g)
    *reinterpret_cast<uint4*>(stg + lane * GEMM_EPI_PITCH + 16 * g) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
  __syncwarp();
  const int piece = lane & 3;
  const int col = col0 + piece * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rr = i * 8 + (lane >> 2);
    const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * GEMM_EPI_PITCH + piece * 16);
    if (row_base + rr < M && col < N) *reinterpret_cast<uint4*>(out + static_cast<long long>(row_base + rr) * ldo + col) = v;
  }
  __syncwarp();
}

// Residual of one 32 x 32 chunk, requested BEFORE the TMEM load.  Normal orientation: rsd = this thread's row, 32 columns.  Swap-AB
// (transpose_out): the residual is [column][row] in memory, so the warp fetches it in the memory layout - rsd[i] = 8 consecutive rows
// (row_base + 8 (lane % 4) ...) of column col0 + 8 i + lane / 4, one 16-byte load each - and gemm_epilogue_chunk transposes it through
// the staging tile.  (Round 2 measured the first version, 32 two-byte loads per thread, at +22 us per 256 x 192 tile: tools/gemm_trace.py.)
__device__ __forceinline__ void gemm_residual_prefetch(const GemmKParams& p, int row_base, int lane, int col0, uint4 (&rsd)[4]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) rsd[g] = make_uint4(0u, 0u, 0u, 0u);
  if (p.residual == nullptr) return;
  if (p.transpose_out) {
    const int r0 = row_base + 8 * (lane & 3);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = col0 + 8 * i + (lane >> 2);
      if (c < p.N && r0 < p.M) rsd[i] = *reinterpret_cast<const uint4*>(p.residual + static_cast<long long>(c) * p.ldr + r0);
    }
    return;
  }
  const int row = row_base + lane;
  if (row < p.M) {
    const bf16* src = p.residual + static_cast<long long>(row) * p.ldr + col0;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (col0 + 8 * g < p.N) rsd[g] = *reinterpret_cast<const uint4*>(src + 8 * g);
  }
}

__device__ __forceinline__ void gemm_epilogue_chunk(const GemmKParams& p, const float (&acc)[32], const uint4 (&rsd)[4], int row_base, int col0,
                                                    uint8_t* stg, int lane) {
  const int row = row_base + lane;
  float v[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) v[e] = acc[e] * p.alpha;
  if (p.bias != nullptr) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (col0 + 4 * g < p.N) {                                   // (N is a multiple of 8: a float4 never straddles the edge)
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * g));   // same address in every lane: one broadcast
        v[4 * g] += b.x; v[4 * g + 1] += b.y; v[4 * g + 2] += b.z; v[4 * g + 3] += b.w;
      }
    }
  }
  if (p.act == 1) {
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = gelu_erf(v[e]);
  } else if (p.act == 2) {
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.0f);
  }
  if (p.residual != nullptr && !p.transpose_out) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float2 r0 = unpack_bf16x2(rsd[g].x), r1 = unpack_bf16x2(rsd[g].y), r2 = unpack_bf16x2(rsd[g].z), r3 = unpack_bf16x2(rsd[g].w);
      v[8 * g] += r0.x; v[8 * g + 1] += r0.y; v[8 * g + 2] += r1.x; v[8 * g + 3] += r1.y;
      v[8 * g + 4] += r2.x; v[8 * g + 5] += r2.y; v[8 * g + 6] += r3.x; v[8 * g + 7] += r3.y;
    }
  }
  if (p.ksplit > 1) {                                             // k-slices merge into the zero-initialised fp32 output
    if (row < p.M) {
      float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col0;
#pragma unroll
      for (int e = 0; e < 32; ++e)
        if (col0 + e < p.N) atomicAdd(o + e, v[e]);
    }
    return;
  }
  if (p.out_f32) {                                                // fp32 outputs (logits, thin products) go out in the row layout
    if (row < p.M) {
      float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo + col0;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        if (col0 + 4 * g < p.N) *reinterpret_cast<float4*>(o + 4 * g) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
    }
    return;
  }
  if (p.transpose_out) {
    // out[column][row] (bf16).  The thread holds one accumulator row x 32 columns; memory wants 8 consecutive ROWS of one column per 16-byte
    // piece.  Both directions go through the warp's staging tile laid out [column][32 rows] (64 B + pad per column): two-byte shared-memory
    // accesses on the thread = row side (a warp touches 64 contiguous bytes: conflict-free), 16-byte accesses on the memory side.
    const int piece = lane & 3, cl = lane >> 2;
    if (p.residual != nullptr) {
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(stg + (8 * i + cl) * GEMM_EPI_PITCH + 16 * piece) = rsd[i];
      __syncwarp();
#pragma unroll
      for (int e = 0; e < 32; ++e)
        v[e] += __bfloat162float(*reinterpret_cast<const bf16*>(stg + e * GEMM_EPI_PITCH + 2 * lane));
      __syncwarp();
    }
#pragma unroll
    for (int e = 0; e < 32; ++e) *reinterpret_cast<bf16*>(stg + e * GEMM_EPI_PITCH + 2 * lane) = __float2bfloat16_rn(v[e]);
    __syncwarp();
    bf16* o = reinterpret_cast<bf16*>(p.out);
    const int r0 = row_base + 8 * piece;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = col0 + 8 * i + cl;
      const uint4 val = *reinterpret_cast<const uint4*>(stg + (8 * i + cl) * GEMM_EPI_PITCH + 16 * piece);
      if (c < p.N && r0 < p.M) *reinterpret_cast<uint4*>(o + static_cast<long long>(c) * p.ldo + r0) = val;
    }
    __syncwarp();
    return;
  }
  uint32_t pk[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) pk[t] = pack_bf16x2(v[2 * t], v[2 * t + 1]);
  gemm_store_chunk_bf16(reinterpret_cast<bf16*>(p.out), p.ldo, p.M, p.N, pk, row_base, col0, stg, lane);
}

// SwiGLU math on bf16-rounded values, exactly as slam_swiglu_fwd / slam_swiglu_bwd compute it (elementwise.cu)
__device__ __forceinline__ float swiglu_fwd_elem(float g, float u) { return g / (1.0f + expf(-g)) * u; }
__device__ __forceinline__ void swiglu_bwd_elem(float g, float u, float d, float& dg, float& du) {
  const float sg = 1.0f / (1.0f + expf(-g));
  du = d * (g * sg);
  dg = d * u * sg * (1.0f + g * (1.0f - sg));
}

// act 3: the thread holds the accumulators of a gate chunk and of its up partner (64 columns further in the blocked-64 layout).
// Writes both chunks of gu and the chunk of h = silu(g) * u.  col_g = global (blocked) column of the gate chunk.
__device__ __forceinline__ void gemm_epilogue_swiglu_fwd(const GemmKParams& p, const float (&ag)[32], const float (&au)[32], int row_base, int col_g,
                                                         uint8_t* stg, int lane) {
  uint32_t pg[16], pu[16], ph[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    pg[t] = pack_bf16x2(ag[2 * t] * p.alpha, ag[2 * t + 1] * p.alpha);
    pu[t] = pack_bf16x2(au[2 * t] * p.alpha, au[2 * t + 1] * p.alpha);
    const float2 g = unpack_bf16x2(pg[t]), u = unpack_bf16x2(pu[t]);
    ph[t] = pack_bf16x2(swiglu_fwd_elem(g.x, u.x), swiglu_fwd_elem(g.y, u.y));
  }
  bf16* out = reinterpret_cast<bf16*>(p.out);
  gemm_store_chunk_bf16(out, p.ldo, p.M, p.N, pg, row_base, col_g, stg, lane);
  gemm_store_chunk_bf16(out, p.ldo, p.M, p.N, pu, row_base, col_g + 64, stg, lane);
  gemm_store_chunk_bf16(p.aux, p.ld_aux, p.M, p.N / 2, ph, row_base, (col_g >> 7) * 64 + (col_g & 127), stg, lane);
}

// act 4: the accumulator chunk is dh for 32 features starting at f0; rg / ru = this row's gate / up values (prefetched from gu).
__device__ __forceinline__ void gemm_swiglu_bwd_prefetch(const GemmKParams& p, int row, int f0, uint4 (&rg)[4], uint4 (&ru)[4]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) rg[g] = ru[g] = make_uint4(0u, 0u, 0u, 0u);
  if (row < p.M && f0 < p.N) {
    const bf16* src = p.aux + static_cast<long long>(row) * p.ld_aux + (f0 >> 6) * 128 + (f0 & 63);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      rg[g] = *reinterpret_cast<const uint4*>(src + 8 * g);
      ru[g] = *reinterpret_cast<const uint4*>(src + 64 + 8 * g);
    }
  }
}
__device__ __forceinline__ void gemm_epilogue_swiglu_bwd(const GemmKParams& p, const float (&acc)[32], const uint4 (&rg)[4], const uint4 (&ru)[4],
                                                         int row_base, int f0, uint8_t* stg, int lane) {
  uint32_t pdg[16], pdu[16];
  const uint32_t* g32 = reinterpret_cast<const uint32_t*>(rg);
  const uint32_t* u32 = reinterpret_cast<const uint32_t*>(ru);
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const float2 d = unpack_bf16x2(pack_bf16x2(acc[2 * t] * p.alpha, acc[2 * t + 1] * p.alpha));   // dh rounded to bf16 like the stored tensor
    const float2 g = unpack_bf16x2(g32[t]), u = unpack_bf16x2(u32[t]);
    float dg0, du0, dg1, du1;
    swiglu_bwd_elem(g.x, u.x, d.x, dg0, du0);
    swiglu_bwd_elem(g.y, u.y, d.y, dg1, du1);
    pdg[t] = pack_bf16x2(dg0, dg1);
    pdu[t] = pack_bf16x2(du0, du1);
  }
  bf16* out = reinterpret_cast<bf16*>(p.out);
  const int col_g = (f0 >> 6) * 128 + (f0 & 63);
  gemm_store_chunk_bf16(out, p.ldo, p.M, 2 * p.N, pdg, row_base, col_g, stg, lane);
  gemm_store_chunk_bf16(out, p.ldo, p.M, 2 * p.N, pdu, row_base, col_g + 64, stg, lane);
}

// act 4 with transpose_out (swap-AB d_down: the WEIGHT W_down^T [F, d] is the M operand): the accumulator chunk is dh for 32 FEATURES (rows
// row_base ..., all inside one 64-feature block) x 32 tokens (columns col0 ...).  gu and d(gu) are [token][2F] (blocked-64) in memory, so - like the
// transposed residual - the warp fetches gate / up in the memory layout (rg[i] / ru[i] = 8 consecutive features of token col0 + 8 i + lane / 4,
// starting at feature row_base + 8 (lane % 4)) and transposes through its staging tile.
__device__ __forceinline__ void gemm_swiglu_bwd_prefetch_t(const GemmKParams& p, int row_base, int lane, int col0, uint4 (&rg)[4], uint4 (&ru)[4]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) rg[g] = ru[g] = make_uint4(0u, 0u, 0u, 0u);
  const int f0 = row_base + 8 * (lane & 3);
  if (f0 >= p.M) return;
  const long long fcol = static_cast<long long>(f0 >> 6) * 128 + (f0 & 63);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = col0 + 8 * i + (lane >> 2);
    if (c < p.N) {
      const bf16* src = p.aux + static_cast<long long>(c) * p.ld_aux + fcol;
      rg[i] = *reinterpret_cast<const uint4*>(src);
      ru[i] = *reinterpret_cast<const uint4*>(src + 64);
    }
  }
}
__device__ __forceinline__ void gemm_epilogue_swiglu_bwd_t(const GemmKParams& p, const float (&acc)[32], const uint4 (&rg)[4], const uint4 (&ru)[4],
                                                           int row_base, int col0, uint8_t* stg, int lane) {
  const int piece = lane & 3, cl = lane >> 2;
  float g[32], u[32];
  // memory layout -> thread = feature: element e = token col0 + e
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(stg + (8 * i + cl) * GEMM_EPI_PITCH + 16 * piece) = rg[i];
  __syncwarp();
#pragma unroll
  for (int e = 0; e < 32; ++e) g[e] = __bfloat162float(*reinterpret_cast<const bf16*>(stg + e * GEMM_EPI_PITCH + 2 * lane));
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(stg + (8 * i + cl) * GEMM_EPI_PITCH + 16 * piece) = ru[i];
  __syncwarp();
#pragma unroll
  for (int e = 0; e < 32; ++e) u[e] = __bfloat162float(*reinterpret_cast<const bf16*>(stg + e * GEMM_EPI_PITCH + 2 * lane));
  __syncwarp();
  // d(gate) goes out first; d(up) overwrites u[] and follows
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    const float d = __bfloat162float(__float2bfloat16_rn(acc[e] * p.alpha));     // dh rounded to bf16 like the stored tensor (slam_swiglu_bwd)
    float dg, du;
    swiglu_bwd_elem(g[e], u[e], d, dg, du);
    *reinterpret_cast<bf16*>(stg + e * GEMM_EPI_PITCH + 2 * lane) = __float2bfloat16_rn(dg);
    u[e] = du;
  }
  __syncwarp();
  bf16* out = reinterpret_cast<bf16*>(p.out);
  const int f0 = row_base + 8 * piece;
  const long long fcol = static_cast<long long>(f0 >> 6) * 128 + (f0 & 63);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = col0 + 8 * i + cl;
    const uint4 val = *reinterpret_cast<const uint4*>(stg + (8 * i + cl) * GEMM_EPI_PITCH + 16 * piece);
    if (c < p.N && f0 < p.M) *reinterpret_cast<uint4*>(out + static_cast<long long>(c) * p.ldo + fcol) = val;
  }
  __syncwarp();
#pragma unroll
  for (int e = 0; e < 32; ++e) *reinterpret_cast<bf16*>(stg + e * GEMM_EPI_PITCH + 2 * lane) = __float2bfloat16_rn(u[e]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = col0 + 8 * i + cl;
    const uint4 val = *reinterpret_cast<const uint4*>(stg + (8 * i + cl) * GEMM_EPI_PITCH + 16 * piece);
    if (c < p.N && f0 < p.M) *reinterpret_cast<uint4*>(out + static_cast<long long>(c) * p.ldo + fcol + 64) = val;
  }
  __syncwarp();
}

}  // namespace slam
