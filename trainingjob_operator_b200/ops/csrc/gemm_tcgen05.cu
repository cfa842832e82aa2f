// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring ->
// tcgen05.mma (one elected thread) -> fp32 accumulators in TMEM (double buffered) ->
// tcgen05.ld epilogue with fused bias / GELU / dGELU / residual / fp32 (split-K, TMA reduce-add) output.
//
//   out[M,N] (+)= opA[M,K] * opB[N,K]^T
//     a_mn == 0 : A stored [M,K] row-major (K contiguous, "K-major")
//     a_mn == 1 : A stored [K,M] row-major (M contiguous, "MN-major")   -> used by wgrad
//     b_mn == 0 : B stored [N,K] row-major                               -> forward  (x @ W^T)
//     b_mn == 1 : B stored [K,N] row-major                               -> dgrad    (dy @ W), wgrad
//
// Two kernels:
//  * gemm_bf16_tcgen05_kernel  -- one CTA per SM, 128 x {128,256} tiles, 320 threads: warp0 = TMA producer, warp1 = TMEM
//    owner + MMA issuer, warps 2..9 = epilogue (TMEM lane group = warp_id % 4, two warps per SM sub-partition).
//  * gemm_bf16_2cta_kernel     -- a 2-CTA cluster (one TPC) computes a 256 x 256 tile with tcgen05.mma.cta_group::2; each
//    CTA loads its 128 rows of A and half of B.  576 threads: 16 epilogue warps (four per sub-partition, 32 rows x 64
//    columns each), 5-stage ring; the residual / pre-GELU tile of an epilogue is prefetched by TMA, outputs leave
//    through swizzled staging + TMA store, the dGELU variant can also emit the bias gradient (column sums).
//    This is the kernel the training step uses for every GEMM with M, N >= 256 (block_n == 512).
// Measurements behind these choices: profiles/ncu_gemm_v1.md, ncu_gemm_v2.md, gemm_role_trace_*.txt.
//
// The reference operator has no GPU code; this kernel belongs to the launched
// workers' training step that BASELINE.json measures (samples/sec).
#include "ptx.cuh"
#include <string.h>

namespace aitj {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 320;   // warp0 TMA, warp1 MMA, warps 2..9 epilogue (2 per TMEM lane group)
constexpr int kEpiWarps = 8;

enum : int {
  EPI_BIAS = 1,       // + bias[N]
  EPI_GELU = 2,       // gelu_tanh(.)
  EPI_RESIDUAL = 4,   // + residual[M,ldc]
  EPI_SAVE_PRE = 8,   // aux[M,ldc] <- value before GELU (bf16)
  EPI_DGELU = 16,     // * gelu'(aux[M,ldc])
  EPI_OUT_F32 = 32,   // out is fp32 (plain store)
  EPI_ACCUM = 64,     // out is fp32, red.global.add (split-K / grad accumulation)
  EPI_MC = 128,       // with EPI_ACCUM: `out` is an NVSwitch multicast address; reduce with multimem.red (GEMM + all-reduce in one kernel)
  EPI_DEBUG_SKIP = 256, // profiling only: the epilogue releases the accumulator without draining it (main loop in isolation)
  EPI_DEBUG_NOTMA = 1024,  // profiling only: stage the tile in shared memory but do not issue the TMA store
  EPI_DEBUG_LDONLY = 2048, // profiling only: read the accumulators (tcgen05.ld) and drop them
  EPI_NO_SIDE_TMA = 8192,  // A/B switch: fetch the residual / pre-GELU tile with per-thread global loads again
  EPI_GROUP_STORE = 4096,  // CTA-pair kernel, bf16 outputs: one 128x64 TMA store per column quarter instead of four 32x64
  EPI_PEER = 16384         // with EPI_ACCUM: `out` lies in the symmetric, owner-sharded gradient buffer; every 32-row group of the
                           // tile is added into the copy of the rank that owns it (red.global.add over NVLink peer memory):
                           // weight-gradient GEMM + reduce-scatter in one kernel, split-K partials included
};

struct GemmArgs {
  int M, N, K;
  int ldc;
  void* out;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* residual;
  __nv_bfloat16* aux;
  int flags;
  int split_k;
  int tiles_m, tiles_n;
  int k_blocks, k_per_split;
  // optional (CTA-pair kernel, bf16 output): colsum[N] += column sums of the stored output, i.e. the bias gradient of
  // the layer whose input gradient this GEMM produces, taken from the staged tile while it waits for its TMA store
  float* colsum;
  // optional per-CTA role timeline (8 x u64 per CTA, see aitj_gemm_set_trace): where each warp role waits
  unsigned long long* trace;
  // EPI_PEER with TMA: one fp32 tensor map of `out` per rank (device memory, 128 B each), indexed by the owner of a
  // 32-row group; null = add from registers (red.global.add.v4.f32 on the peer mapping)
  const CUtensorMap* peer_maps;
  // EPI_PEER with split_k > 1: arrival counters, one per 32x32-aligned sub-block start (self-resetting).  The splits
  // accumulate into the LOCAL copy; the last one to finish a sub-block moves it to its owner and clears the local copy,
  // so a gradient crosses NVLink once instead of split_k times
  unsigned int* peer_counters;
};

// trace slots
enum : int { TR_MMA_TOTAL = 0, TR_MMA_WAIT_FULL = 1, TR_MMA_WAIT_TMEM = 2, TR_TMA_WAIT_EMPTY = 3,
             TR_EPI_WAIT_FULL = 4, TR_EPI_BUSY = 5, TR_EPI_TOTAL = 6, TR_TILES = 7 };
#define TR_BEGIN(tr, t) long long t = (tr) ? clock64() : 0
#define TR_ADD(tr, t, accv) do { if (tr) accv += clock64() - t; } while (0)

struct WorkItem {
  int m_blk, n_blk, kb0, kb1;
};

__device__ __forceinline__ WorkItem decode_work(const GemmArgs& a, int w) {
  const int num_tiles = a.tiles_m * a.tiles_n;
  const int split = w / num_tiles;
  const int t = w - split * num_tiles;
  constexpr int GROUP_M = 8;
  const int group_sz = GROUP_M * a.tiles_n;
  const int g = t / group_sz;
  const int first_m = g * GROUP_M;
  const int gm = min(a.tiles_m - first_m, GROUP_M);
  const int r = t - g * group_sz;
  WorkItem wi;
  wi.m_blk = first_m + (r % gm);
  wi.n_blk = r / gm;
  wi.kb0 = split * a.k_per_split;
  wi.kb1 = min(a.k_blocks, wi.kb0 + a.k_per_split);
  return wi;
}

// ---- epilogue helpers -----------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* x) {
  uint4 o;
  o.x = pack_bf16x2(x[0], x[1]); o.y = pack_bf16x2(x[2], x[3]);
  o.z = pack_bf16x2(x[4], x[5]); o.w = pack_bf16x2(x[6], x[7]);
  return o;
}

// One warp's staging buffer: 32 rows x 128 B, written in the TMA 128B-swizzle pattern (16-byte chunk c of
// row r lives at chunk c ^ (r & 7)), so the st.shared.v4 of a quarter warp hit 32 distinct banks and the tile
// leaves through one coalesced TMA store (or fp32 reduce-add) per chunk.
__device__ __forceinline__ void stage_acquire(int lane) {
  if (lane == 0) tma_store_wait_read<0>();
  __syncwarp();
}
__device__ __forceinline__ void stage_write16(uint8_t* b, int lane, int chunk, const uint4& v) {
  *reinterpret_cast<uint4*>(b + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = v;
}
__device__ __forceinline__ void stage_commit(const CUtensorMap* tm, uint8_t* b, int col0, int row0, int lane,
                                             bool reduce_add) {
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    if (reduce_add) tma_reduce_add_2d(tm, b, col0, row0);
    else tma_store_2d(tm, b, col0, row0);
    tma_store_commit();
  }
}

// EPI_PEER, split-K: a sub-block whose local reduce-adds are in flight.  Its completion (count this split in; the last
// arrival moves the block to its owner) is processed one tile later, when the bulk reduce-adds have long landed, so the
// epilogue never waits for them.
struct PeerPending {
  int row0 = -1, col0 = 0;         // coordinates of the pending sub-block in the output tensor; row0 < 0 = nothing pending
  int cols = 0;                    // its width (32 rows x cols)
  uint32_t bar_phase = 0;          // phase of this warp's move barrier
};

// `bar` != nullptr (and a.peer_maps set): bulk move -- TMA load of the finished block into the warp's staging buffer, TMA
// reduce-add into the owner's copy (bulk packets use NVLink far better than 16-byte red.add's); else move from registers.
__device__ __forceinline__ void peer_finish(const GemmArgs& a, PeerPending& pd, const CUtensorMap* tm_local, uint8_t* sbuf,
                                            uint64_t* bar, int lane) {
  if (pd.row0 < 0) return;
  const int row0 = pd.row0, col0 = pd.col0;
  pd.row0 = -1;
  float* first = reinterpret_cast<float*>(a.out) + static_cast<size_t>(row0) * a.ldc + col0;
  const int owner = peer_owner(first - col0);
  const long long delta = c_peers.delta[owner];
  unsigned int* ctr = a.peer_counters + static_cast<size_t>(row0 >> 5) * ((a.N + 31) >> 5) + (col0 >> 5);
  unsigned int seen = 0;
  if (lane == 0) {
    __threadfence();
    seen = atomicAdd(ctr, 1u);
  }
  seen = __shfl_sync(0xffffffffu, seen, 0);
  if (seen != static_cast<unsigned int>(a.split_k - 1)) return;
  __threadfence();
  // the finished block goes to its owner over NVLink; the local copy is cleared for the next step
  const int rows = min(32, a.M - row0);
  const int cols = min(pd.cols, a.N - col0);
  if (bar != nullptr && a.peer_maps != nullptr) {
#pragma unroll 1
    for (int c = 0; c < cols; c += 32) {
      stage_acquire(lane);                                   // earlier bulk stores have read the staging buffer
      if (lane == 0) {
        mbar_arrive_expect_tx(bar, 4096);
        tma_load_2d(sbuf, tm_local, bar, col0 + c, row0);
      }
      mbar_wait(bar, pd.bar_phase);
      pd.bar_phase ^= 1u;
      if (lane == 0) {
        tma_reduce_add_2d(a.peer_maps + owner, sbuf, col0 + c, row0);
        tma_store_commit();
      }
      // clear the local copy (the load above has completed; nobody else touches this block any more)
      const int ncol4 = min(32, cols - c) >> 2;
      for (int i = lane; i < rows * ncol4; i += 32) {
        const int r = i / ncol4, c4 = i - r * ncol4;
        __stcg(reinterpret_cast<float4*>(first + static_cast<size_t>(r) * a.ldc + c + c4 * 4),
               make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
  } else {
    const int vec_per_row = cols >> 2;
#pragma unroll 8
    for (int i = lane; i < rows * vec_per_row; i += 32) {
      const int r = i / vec_per_row, c4 = i - r * vec_per_row;
      float* p = first + static_cast<size_t>(r) * a.ldc + c4 * 4;
      const float4 v = __ldcg(reinterpret_cast<const float4*>(p));
      red_add_v4_f32(reinterpret_cast<float*>(reinterpret_cast<char*>(p) + delta), v.x, v.y, v.z, v.w);
      __stcg(reinterpret_cast<float4*>(p), make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
  if (lane == 0) *ctr = 0u;
}

// fp32 outputs (plain store, TMA reduce-add, or multimem reduction): raw accumulators, 32 columns per chunk.
template <int kCols>
__device__ __forceinline__ void epilogue_f32(const GemmArgs& a, const CUtensorMap* tm_out, uint32_t tmem_acc, int row0,
                                             int n0, int c_begin, uint8_t* sbuf, int lane, PeerPending* pend = nullptr,
                                             uint64_t* move_bar = nullptr) {
  const int flags = a.flags;
  if ((flags & EPI_PEER) && a.split_k > 1 && a.peer_counters != nullptr) {
    if (row0 >= a.M || n0 + c_begin >= a.N) return;
    float* obase = reinterpret_cast<float*>(a.out);
    float* first = obase + static_cast<size_t>(row0) * a.ldc;
    const long long delta = c_peers.delta[peer_owner(first)];
    // (1) split-K partial -> local copy, exactly like the single-GPU path
    int issued = 0;
#pragma unroll 1
    for (int c = 0; c < kCols / 32; ++c) {
      const int col0 = n0 + c_begin + c * 32;
      if (col0 >= a.N) break;
      uint32_t r[32];
      tmem_ld_32x32(tmem_acc + c_begin + c * 32, r);
      tmem_ld_wait();
      stage_acquire(lane);
#pragma unroll
      for (int q = 0; q < 8; ++q) stage_write16(sbuf, lane, q, make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]));
      stage_commit(tm_out, sbuf, col0, row0, lane, true);
      ++issued;
    }
    // the PREVIOUS tile's reduce-adds have landed once all but this tile's bulk groups are complete
    if (pend != nullptr && pend->row0 >= 0) {
      if (lane == 0) {
        if (issued == kCols / 32) tma_store_wait<kCols / 32>();
        else tma_store_wait<0>();
      }
      __syncwarp();
      peer_finish(a, *pend, tm_out, sbuf, move_bar, lane);
    }
    if (delta == 0) return;                 // this rank owns these rows: they are where they belong
    if (pend != nullptr) {
      pend->row0 = row0;                    // processed with the next tile (or at the end of the kernel)
      pend->col0 = n0 + c_begin;
      pend->cols = kCols;
    } else {
      PeerPending cur;
      cur.row0 = row0; cur.col0 = n0 + c_begin; cur.cols = kCols;
      if (lane == 0) tma_store_wait<0>();
      __syncwarp();
      peer_finish(a, cur, tm_out, sbuf, nullptr, lane);
    }
    return;
  }
  if ((flags & EPI_PEER) && a.peer_maps != nullptr) {
    // fused GEMM -> reduce-scatter through the TMA unit: the same staged 32x32 fp32 chunks as the local split-K path,
    // but the bulk reduce-add targets the OWNER's copy of the gradient (its tensor map covers the peer mapping)
    if (row0 >= a.M) return;
    tm_out = a.peer_maps + peer_owner(reinterpret_cast<const float*>(a.out) + static_cast<size_t>(row0) * a.ldc);
  } else if (flags & (EPI_MC | EPI_PEER)) {
    // fused GEMM -> all-reduce (EPI_MC): the accumulators are reduced into ALL peers' gradient buffers through the
    // NVSwitch (multimem.red).  Fused GEMM -> reduce-scatter (EPI_PEER): they are added into the OWNER's buffer only
    // (ownership is by flat offset with 32-row granularity, so this warp's 32 rows have one owner).  The 32x32 fp32
    // chunk is transposed through the staging buffer so that each warp instruction covers 4 rows x 128 contiguous
    // bytes (full lines on NVLink) instead of 32 rows x 16 bytes.
    float* obase = reinterpret_cast<float*>(a.out);
    const bool mc = (flags & EPI_MC) != 0;
    if (!mc && row0 < a.M) {
      float* first = obase + static_cast<size_t>(row0) * a.ldc;
      obase += peer_ptr(first) - first;
    }
#pragma unroll 1
    for (int c = 0; c < kCols / 32; ++c) {
      const int col0 = n0 + c_begin + c * 32;
      if (col0 >= a.N) break;
      uint32_t r[32];
      tmem_ld_32x32(tmem_acc + c_begin + c * 32, r);
      tmem_ld_wait();
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 8; ++q) stage_write16(sbuf, lane, q, make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]));
      __syncwarp();
      const int chunk = lane & 7;
      if (col0 + chunk * 4 < a.N) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + (lane >> 3);
          const int grow = row0 + rr;
          if (grow < a.M) {
            const uint4 v = *reinterpret_cast<const uint4*>(sbuf + rr * 128 + ((chunk ^ (rr & 7)) << 4));
            float* dst = obase + static_cast<size_t>(grow) * a.ldc + col0 + chunk * 4;
            if (mc) mc_red_add_v4_f32(dst, __uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z),
                                      __uint_as_float(v.w));
            else red_add_v4_f32(dst, __uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z),
                                __uint_as_float(v.w));
          }
        }
      }
    }
    return;
  }
  if (flags & (EPI_OUT_F32 | EPI_ACCUM)) {
#pragma unroll 1
    for (int c = 0; c < kCols / 32; ++c) {
      const int col0 = n0 + c_begin + c * 32;
      if (col0 >= a.N) break;
      uint32_t r[32];
      tmem_ld_32x32(tmem_acc + c_begin + c * 32, r);
      tmem_ld_wait();
      stage_acquire(lane);
#pragma unroll
      for (int q = 0; q < 8; ++q) stage_write16(sbuf, lane, q, make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]));
      stage_commit(tm_out, sbuf, col0, row0, lane, (flags & EPI_ACCUM) != 0);
    }
  }
}

// Drain this warp's share of one accumulator tile: 32 rows x kCols columns starting at tile column c_begin.
// `sbias` holds the tile's bias row (bf16, staged once per tile by the epilogue warps).
template <int kCols>
__device__ __forceinline__ void epilogue_tile(const GemmArgs& a, const CUtensorMap* tm_out, const CUtensorMap* tm_aux,
                                              uint32_t tmem_acc, int row0, int n0, int c_begin,
                                              const __nv_bfloat16* sbias, uint8_t* sbuf, int lane,
                                              PeerPending* pend = nullptr, uint64_t* move_bar = nullptr) {
  const int flags = a.flags;
  const int row = row0 + lane;
  const bool row_ok = row < a.M;
  if (flags & (EPI_MC | EPI_PEER | EPI_OUT_F32 | EPI_ACCUM)) {
    epilogue_f32<kCols>(a, tm_out, tmem_acc, row0, n0, c_begin, sbuf, lane, pend, move_bar);
    return;
  }
  // bf16 output: 64 columns (128 B) per staged chunk
  const bool need_side = (flags & (EPI_DGELU | EPI_RESIDUAL)) != 0;
  const __nv_bfloat16* side = (flags & EPI_DGELU) ? a.aux : a.residual;
#pragma unroll 1
  for (int c = 0; c < kCols / 64; ++c) {
    const int ct = c_begin + c * 64;        // column inside the tile
    const int col0 = n0 + ct;
    if (col0 >= a.N) break;
    // issue the row-scattered side loads (residual / pre-GELU) and both TMEM loads before waiting on anything
    uint4 sv[8];
    if (need_side && row_ok) {
      const __nv_bfloat16* sp = side + static_cast<size_t>(row) * a.ldc + col0;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        sv[q] = (col0 + q * 8 < a.N) ? *reinterpret_cast<const uint4*>(sp + q * 8) : make_uint4(0, 0, 0, 0);
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) sv[q] = make_uint4(0, 0, 0, 0);
    }
    uint32_t r0[32], r1[32];
    tmem_ld_32x32(tmem_acc + ct, r0);
    tmem_ld_32x32(tmem_acc + ct + 32, r1);
    tmem_ld_wait();
    uint4 outv[8], prev[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(q < 4 ? r0[q * 8 + j] : r1[(q - 4) * 8 + j]);
      if (flags & EPI_BIAS) {
        float bv[8];
        unpack8(*reinterpret_cast<const uint4*>(sbias + ct + q * 8), bv);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] += bv[j];
      }
      if (flags & EPI_SAVE_PRE) prev[q] = pack8(x);
      if (flags & EPI_GELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = gelu_tanh(x[j]);
      }
      if (flags & EPI_DGELU) {
        float h[8];
        unpack8(sv[q], h);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] *= gelu_tanh_grad(h[j]);
      }
      if (flags & EPI_RESIDUAL) {
        float h[8];
        unpack8(sv[q], h);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] += h[j];
      }
      outv[q] = pack8(x);
    }
    if (flags & EPI_SAVE_PRE) {
      stage_acquire(lane);
#pragma unroll
      for (int q = 0; q < 8; ++q) stage_write16(sbuf, lane, q, prev[q]);
      stage_commit(tm_aux, sbuf, col0, row0, lane, false);
    }
    stage_acquire(lane);
#pragma unroll
    for (int q = 0; q < 8; ++q) stage_write16(sbuf, lane, q, outv[q]);
    stage_commit(tm_out, sbuf, col0, row0, lane, false);
  }
}

// 16-warp variant of the drain: this warp owns 32 rows x 64 columns (tile columns c_begin .. c_begin+63).  The
// bf16 path works in two 32-column halves so that at most one TMEM load + one side load are live per thread
// (<= 96 registers at 576 threads per CTA); with four epilogue warps per SM sub-partition the TMEM / global /
// TMA-store latencies of one warp are covered by the other three.
//
// Stores: the four warps that share a column quarter (TMEM lane groups 0..3 => CTA rows 0..127) stage into one
// contiguous 16 KB buffer -- warp lg at rows lg*32.., which keeps the 128B-swizzle pattern of a 128-row box -- and ONE
// thread issues ONE 128x64 TMA store for all four (EPI_GROUP_STORE; named barrier 2+quarter, 128 threads).  The TMA
// stores of a K=768 GEMM cost ~10 us next to the operand ring (profiles/ncu_gemm_v2.md); fewer, larger stores did
// not change that (it is bytes, not store count), so the per-warp 4 KB stores stay the default.
struct StoreGroup {
  uint8_t* buf;      // 16 KB staging of the group; this warp's rows start at buf + lg * 4096
  int bar_id;        // named barrier of the group
  int row0_cta;      // first output row of this CTA's 128-row slab
  bool issuer;       // the one thread that talks to the TMA unit
  bool grouped;
};

__device__ __forceinline__ void group_acquire(const StoreGroup& g, int lane) {
  if (g.grouped) {
    if (g.issuer) tma_store_wait_read<0>();
    asm volatile("bar.sync %0, 128;" ::"r"(g.bar_id) : "memory");
  } else {
    stage_acquire(lane);
  }
}
__device__ __forceinline__ void group_commit(const StoreGroup& g, const CUtensorMap* tm32, const CUtensorMap* tm128,
                                             uint8_t* wbuf, int col0, int row0, int lane) {
  if (g.grouped) {
    fence_proxy_async_smem();
    asm volatile("bar.sync %0, 128;" ::"r"(g.bar_id) : "memory");
    if (g.issuer) {
      tma_store_2d(tm128, g.buf, col0, g.row0_cta);
      tma_store_commit();
    }
  } else {
    stage_commit(tm32, wbuf, col0, row0, lane, false);
  }
}

__device__ __forceinline__ void epilogue_cols64(const GemmArgs& a, const CUtensorMap* tm_out, const CUtensorMap* tm_aux,
                                                const CUtensorMap* tm_out128, const CUtensorMap* tm_aux128,
                                                uint32_t tmem_acc, int row0, int n0, int c_begin,
                                                const __nv_bfloat16* sbias, const StoreGroup& g, int lg, int lane,
                                                uint64_t* side_bar = nullptr, uint32_t side_parity = 0,
                                                PeerPending* pend = nullptr, uint64_t* move_bar = nullptr) {
  // side_bar != nullptr: the residual / pre-GELU tile of this warp was prefetched by TMA into its staging buffer
  // (swizzled like the output); it is read from there and overwritten in place by the result
  const int flags = a.flags;
  uint8_t* sbuf = g.buf + lg * 4096;
  if (flags & (EPI_MC | EPI_OUT_F32 | EPI_ACCUM)) {
    epilogue_f32<64>(a, tm_out, tmem_acc, row0, n0, c_begin, sbuf, lane, pend, move_bar);
    return;
  }
  const int col0 = n0 + c_begin;
  if (col0 >= a.N) return;          // uniform over the store group
  const int row = row0 + lane;
  const bool row_ok = row < a.M;
  const bool need_side = (flags & (EPI_DGELU | EPI_RESIDUAL)) != 0;
  const __nv_bfloat16* side = (flags & EPI_DGELU) ? a.aux : a.residual;
  const __nv_bfloat16* sp = side + static_cast<size_t>(row) * a.ldc + col0;
  if (flags & EPI_SAVE_PRE) {
    // pass 1: the pre-activation (acc + bias) goes out through the aux tensor map
    group_acquire(g, lane);
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_acc + c_begin + h * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float x[8], bv[8];
        unpack8(*reinterpret_cast<const uint4*>(sbias + c_begin + h * 32 + q * 8), bv);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(r[q * 8 + j]) + ((flags & EPI_BIAS) ? bv[j] : 0.f);
        stage_write16(sbuf, lane, h * 4 + q, pack8(x));
      }
    }
    group_commit(g, tm_aux, tm_aux128, sbuf, col0, row0, lane);
  }
  if (side_bar != nullptr) {
    mbar_wait(side_bar, side_parity);
  } else {
    group_acquire(g, lane);
  }
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    uint4 sv[4];
    if (side_bar != nullptr) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        sv[q] = *reinterpret_cast<const uint4*>(sbuf + lane * 128 + (((h * 4 + q) ^ (lane & 7)) << 4));
    } else if (need_side && row_ok) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        sv[q] = (col0 + h * 32 + q * 8 < a.N) ? *reinterpret_cast<const uint4*>(sp + h * 32 + q * 8)
                                             : make_uint4(0, 0, 0, 0);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) sv[q] = make_uint4(0, 0, 0, 0);
    }
    uint32_t r[32];
    tmem_ld_32x32(tmem_acc + c_begin + h * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(r[q * 8 + j]);
      if (flags & EPI_BIAS) {
        float bv[8];
        unpack8(*reinterpret_cast<const uint4*>(sbias + c_begin + h * 32 + q * 8), bv);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] += bv[j];
      }
      if (flags & EPI_GELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = gelu_tanh(x[j]);
      }
      if (flags & EPI_DGELU) {
        float hh[8];
        unpack8(sv[q], hh);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] *= gelu_tanh_grad(hh[j]);
      }
      if (flags & EPI_RESIDUAL) {
        float hh[8];
        unpack8(sv[q], hh);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] += hh[j];
      }
      if (flags & EPI_DEBUG_LDONLY) {
        if (x[0] == 1.2345e30f) sbuf[0] = 1;       // keep the loads alive
      } else {
        stage_write16(sbuf, lane, h * 4 + q, pack8(x));
      }
    }
  }
  if (flags & (EPI_DEBUG_NOTMA | EPI_DEBUG_LDONLY)) {
    if (g.grouped) {   // keep the barrier sequence of the group balanced
      asm volatile("bar.sync %0, 128;" ::"r"(g.bar_id) : "memory");
    }
    return;
  }
  group_commit(g, tm_out, tm_out128, sbuf, col0, row0, lane);
  if (a.colsum != nullptr) {
    // lane l sums columns 2l, 2l+1 of this warp's 32 staged rows (bf16, exactly what was stored); one row of the
    // swizzled buffer is read by the 32 lanes as 32 distinct words -> conflict free
    __syncwarp();
    const int rows = min(32, a.M - row0);
    const int chunk = lane >> 2, within = (lane & 3) << 2;
    float s0 = 0.f, s1 = 0.f;
    for (int r = 0; r < rows; ++r) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(sbuf + r * 128 + ((chunk ^ (r & 7)) << 4) + within);
      const float2 f = unpack_bf16x2(w);
      s0 += f.x; s1 += f.y;
    }
    const int c = col0 + 2 * lane;
    if (c < a.N) atomicAdd(a.colsum + c, s0);
    if (c + 1 < a.N) atomicAdd(a.colsum + c + 1, s1);
  }
}

// Epilogue warps stage the tile's bias row (kBlockN bf16) into shared memory; 256 threads, named barrier 1.
template <int kBlockN, int kEpiThreads = 256>
__device__ __forceinline__ void stage_bias(const GemmArgs& a, __nv_bfloat16* sbias, int n0, int epi_tid) {
  if (a.flags & EPI_BIAS) {
    if (epi_tid < kBlockN) {
      const int col = n0 + epi_tid;
      sbias[epi_tid] = col < a.N ? a.bias[col] : __float2bfloat16(0.f);
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
}

template <int kBlockN, bool kAMN, bool kBMN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_aux,
                         const GemmArgs args) {
  constexpr int kStageA = BLOCK_M * BLOCK_K * 2;
  constexpr int kStageB = kBlockN * BLOCK_K * 2;
  constexpr int kStageBytes = kStageA + kStageB;
  constexpr int kStages = (kBlockN == 256) ? 4 : 6;
  constexpr uint32_t kTmemCols = 2 * kBlockN;
  constexpr uint32_t kIdesc = make_idesc_bf16(BLOCK_M, kBlockN, kAMN ? 1u : 0u, kBMN ? 1u : 0u);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  constexpr int kStagingBytes = kEpiWarps * 4096;  // one 32x128B chunk per epilogue warp
  uint8_t* staging = smem + kStages * kStageBytes;
  __nv_bfloat16* sbias = reinterpret_cast<__nv_bfloat16*>(staging + kStagingBytes);   // 2 x 256 bf16
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + kStagingBytes + 1024);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    if (args.flags & EPI_SAVE_PRE) tma_prefetch_desc(&tmap_aux);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_work = args.tiles_m * args.tiles_n * args.split_k;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const WorkItem wi = decode_work(args, w);
        for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kStageA;
          if constexpr (!kAMN) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, wi.m_blk * BLOCK_M);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              tma_load_2d(sa + j * 8192, &tmap_a, &full_bar[stage], wi.m_blk * BLOCK_M + j * 64, kb * BLOCK_K);
          }
          if constexpr (!kBMN) {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, wi.n_blk * kBlockN);
          } else {
#pragma unroll
            for (int j = 0; j < kBlockN / 64; ++j)
              tma_load_2d(sb + j * 8192, &tmap_b, &full_bar[stage], wi.n_blk * kBlockN + j * 64, kb * BLOCK_K);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const WorkItem wi = decode_work(args, w);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBlockN;
        for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
          const uint32_t b_addr = a_addr + kStageA;
#pragma unroll
          for (int ks = 0; ks < BLOCK_K / UMMA_K; ++ks) {
            const uint64_t da = kAMN ? make_sw128_desc(a_addr + ks * 2048, BLOCK_K * 128, 1024)
                                     : make_sw128_desc(a_addr + ks * UMMA_K * 2, 16, 1024);
            const uint64_t db = kBMN ? make_sw128_desc(b_addr + ks * 2048, BLOCK_K * 128, 1024)
                                     : make_sw128_desc(b_addr + ks * UMMA_K * 2, 16, 1024);
            umma_bf16(d_tmem, da, db, kIdesc, (kb > wi.kb0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;
    const int epi_tid = threadIdx.x - 64;
    uint8_t* sbuf = staging + (warp - 2) * 4096;
    int acc = 0;
    uint32_t acc_phase = 0;
    PeerPending pend;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
      const WorkItem wi = decode_work(args, w);
      stage_bias<kBlockN>(args, sbias + acc * 256, wi.n_blk * kBlockN, epi_tid);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (wi.kb1 > wi.kb0) {
        epilogue_tile<kBlockN / 2>(args, &tmap_out, &tmap_aux,
                                   tmem_base + acc * kBlockN + (static_cast<uint32_t>(lg * 32) << 16),
                                   wi.m_blk * BLOCK_M + lg * 32, wi.n_blk * kBlockN, half * (kBlockN / 2),
                                   sbias + acc * 256, sbuf, lane, &pend);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) tma_store_wait<0>();
    __syncwarp();
    peer_finish(args, pend, &tmap_out, sbuf, nullptr, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------ CTA-pair kernel (cta_group::2)
// Two CTAs of a cluster (one TPC) compute one 256x256 tile: each loads its own 128 rows of A and HALF of
// B (128 of the 256 N rows), the leader issues tcgen05.mma.cta_group::2 (M=256) that reads both CTAs'
// shared memory, and each CTA drains its own 128 accumulator rows from its TMEM.  Per-CTA operand traffic
// drops from 48 KB to 32 KB per k-block (the 1-CTA kernel is L2->SM bandwidth bound at K=768..3072) and the
// freed shared memory buys a 6-stage ring.
// kPeer: instantiated for the weight-gradient layout only -- carries the "pending finished block" state of the
// owner-sharded split-K protocol across tiles (registers the other instantiations must not pay for)
template <bool kAMN, bool kBMN, int kEpiW, bool kPeer = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * kEpiW, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_aux,
                      const __grid_constant__ CUtensorMap tmap_out128, const __grid_constant__ CUtensorMap tmap_aux128,
                      const __grid_constant__ CUtensorMap tmap_side, const GemmArgs args) {
  constexpr int kPairM = 256, kPairN = 256;
  constexpr int kStageA = BLOCK_M * BLOCK_K * 2;       // this CTA's 128 rows of A
  constexpr int kStageB = (kPairN / 2) * BLOCK_K * 2;  // this CTA's half of B
  constexpr int kStageBytes = kStageA + kStageB;
  constexpr int kStages = kEpiW == 16 ? 5 : 6;   // 16 staging buffers cost one ring stage
  constexpr uint32_t kTmemCols = 2 * kPairN;
  constexpr uint32_t kIdesc = make_idesc_bf16(kPairM, kPairN, kAMN ? 1u : 0u, kBMN ? 1u : 0u);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  constexpr int kStagingBytes = kEpiW * 4096;
  uint8_t* staging = smem + kStages * kStageBytes;
  __nv_bfloat16* sbias = reinterpret_cast<__nv_bfloat16*>(staging + kStagingBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + kStagingBytes + 1024);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* side_bar = tmem_empty + 3;          // [kEpiW] one per epilogue warp: its prefetched residual / pre-GELU tile

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    if (args.flags & EPI_SAVE_PRE) tma_prefetch_desc(&tmap_aux);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);    // leader's copy is the one in use: 1 arrive(expect_tx) + bytes of both CTAs
      mbar_init(&empty_bar[i], 1);   // per CTA, released by the leader's multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);   // per CTA, multicast commit
      mbar_init(&tmem_empty[i], 2 * kEpiW);  // leader's copy: all epilogue warps of both CTAs
    }
    for (int i = 0; i < kEpiW; ++i) mbar_init(&side_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta<kTmemCols>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_work = args.tiles_m * args.tiles_n * args.split_k;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long tr_wait = 0;
      for (int w = cluster_id; w < num_work; w += num_clusters) {
        const WorkItem wi = decode_work(args, w);
        const int m0 = wi.m_blk * kPairM + static_cast<int>(rank) * BLOCK_M;
        const int n0 = wi.n_blk * kPairN + static_cast<int>(rank) * (kPairN / 2);
        for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
          TR_BEGIN(args.trace, t0);
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          TR_ADD(args.trace, t0, tr_wait);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);
          const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kStageA;
          if constexpr (!kAMN) {
            tma_load_2d_2sm(sa, &tmap_a, full_leader, kb * BLOCK_K, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              tma_load_2d_2sm(sa + j * 8192, &tmap_a, full_leader, m0 + j * 64, kb * BLOCK_K);
          }
          if constexpr (!kBMN) {
            tma_load_2d_2sm(sb, &tmap_b, full_leader, kb * BLOCK_K, n0);
          } else {
#pragma unroll
            for (int j = 0; j < (kPairN / 2) / 64; ++j)
              tma_load_2d_2sm(sb + j * 8192, &tmap_b, full_leader, n0 + j * 64, kb * BLOCK_K);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
      if (args.trace) args.trace[blockIdx.x * 8 + TR_TMA_WAIT_EMPTY] = tr_wait;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long tr_full = 0, tr_tmem = 0, tr_tiles = 0;
      TR_BEGIN(args.trace, tr_start);
      for (int w = cluster_id; w < num_work; w += num_clusters) {
        const WorkItem wi = decode_work(args, w);
        TR_BEGIN(args.trace, t0);
        mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1u);
        TR_ADD(args.trace, t0, tr_tmem);
        ++tr_tiles;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kPairN;
        for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
          TR_BEGIN(args.trace, t1);
          mbar_wait_cluster(&full_bar[stage], phase);
          TR_ADD(args.trace, t1, tr_full);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
          const uint32_t b_addr = a_addr + kStageA;
#pragma unroll
          for (int ks = 0; ks < BLOCK_K / UMMA_K; ++ks) {
            const uint64_t da = kAMN ? make_sw128_desc(a_addr + ks * 2048, BLOCK_K * 128, 1024)
                                     : make_sw128_desc(a_addr + ks * UMMA_K * 2, 16, 1024);
            const uint64_t db = kBMN ? make_sw128_desc(b_addr + ks * 2048, BLOCK_K * 128, 1024)
                                     : make_sw128_desc(b_addr + ks * UMMA_K * 2, 16, 1024);
            umma_bf16_2cta(d_tmem, da, db, kIdesc, (kb > wi.kb0 || ks > 0) ? 1u : 0u);
          }
          umma_commit_2cta(&empty_bar[stage], 3);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2cta(&tmem_full[acc], 3);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      if (args.trace) {
        unsigned long long* tr = args.trace + blockIdx.x * 8;
        tr[TR_MMA_TOTAL] = clock64() - tr_start;
        tr[TR_MMA_WAIT_FULL] = tr_full;
        tr[TR_MMA_WAIT_TMEM] = tr_tmem;
        tr[TR_TILES] = tr_tiles;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2.., both CTAs)
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;     // column slice: halves (8 warps) or quarters (16 warps)
    const int epi_tid = threadIdx.x - 64;
    // staging: warps of one column slice are contiguous in TMEM-lane-group order (rows 0..127 of the CTA's slab)
    uint8_t* sbuf = kEpiW == 16 ? staging + (half * 4 + lg) * 4096 : staging + (warp - 2) * 4096;
    StoreGroup sg;
    sg.buf = staging + half * 16384;
    sg.bar_id = 2 + half;
    sg.issuer = lg == 0 && lane == 0;
    sg.grouped = kEpiW == 16 && (args.flags & EPI_GROUP_STORE) != 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    // The residual / pre-GELU tile this warp's epilogue needs is fetched by TMA into the warp's staging buffer while
    // the tile's MMAs are still running (the row-scattered 16-byte global loads it replaces cost the dGELU epilogue
    // ~4000 cycles per tile, more than the K=768 main loop leaves).
    const int side_kind = args.flags & (EPI_DGELU | EPI_RESIDUAL);
    const bool side_tma = kEpiW == 16 && !sg.grouped &&
                          !(args.flags & (EPI_SAVE_PRE | EPI_MC | EPI_OUT_F32 | EPI_ACCUM | EPI_NO_SIDE_TMA)) &&
                          (side_kind == EPI_DGELU || side_kind == EPI_RESIDUAL);
    uint64_t* my_side_bar = &side_bar[warp - 2];
    uint32_t side_phase = 0;
    PeerPending pend;
    PeerPending* const pp = kPeer ? &pend : nullptr;
    // the warp's side barrier is free for the bulk move of finished gradient blocks (no side prefetch with fp32 outputs)
    uint64_t* const move_bar = (kPeer && kEpiW == 16 && !side_tma) ? my_side_bar : nullptr;
    long long tr_wait = 0, tr_busy = 0;
    TR_BEGIN(args.trace, tr_start);
    for (int w = cluster_id; w < num_work; w += num_clusters) {
      const WorkItem wi = decode_work(args, w);
      bool side_now = false;
      if constexpr (kEpiW == 16) {
        if (side_tma && wi.kb1 > wi.kb0 && wi.n_blk * kPairN + half * 64 < args.N) {
          side_now = true;
          stage_acquire(lane);             // the previous tile's TMA store has finished reading this buffer
          if (lane == 0) {
            mbar_arrive_expect_tx(my_side_bar, 4096);
            tma_load_2d(sbuf, &tmap_side, my_side_bar, wi.n_blk * kPairN + half * 64,
                        wi.m_blk * kPairM + static_cast<int>(rank) * BLOCK_M + lg * 32);
          }
        }
      }
      stage_bias<kPairN, 32 * kEpiW>(args, sbias + acc * 256, wi.n_blk * kPairN, epi_tid);
      TR_BEGIN(args.trace, t0);
      mbar_wait_cluster(&tmem_full[acc], acc_phase);
      TR_ADD(args.trace, t0, tr_wait);
      TR_BEGIN(args.trace, t1);
      tc_fence_after();
      if (wi.kb1 > wi.kb0 && !(args.flags & EPI_DEBUG_SKIP)) {
        const uint32_t t_acc = tmem_base + acc * kPairN + (static_cast<uint32_t>(lg * 32) << 16);
        const int row0 = wi.m_blk * kPairM + static_cast<int>(rank) * BLOCK_M + lg * 32;
        if constexpr (kEpiW == 16) {
          sg.row0_cta = wi.m_blk * kPairM + static_cast<int>(rank) * BLOCK_M;
          epilogue_cols64(args, &tmap_out, &tmap_aux, &tmap_out128, &tmap_aux128, t_acc, row0, wi.n_blk * kPairN,
                          half * 64, sbias + acc * 256, sg, lg, lane, side_now ? my_side_bar : nullptr, side_phase,
                          pp, move_bar);
          if (side_now) side_phase ^= 1u;
        } else {
          epilogue_tile<kPairN / 2>(args, &tmap_out, &tmap_aux, t_acc, row0, wi.n_blk * kPairN, half * (kPairN / 2),
                                    sbias + acc * 256, sbuf, lane, pp, move_bar);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tmem_empty[acc], 0);
      TR_ADD(args.trace, t1, tr_busy);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) tma_store_wait<0>();
    if constexpr (kPeer) {
      __syncwarp();
      peer_finish(args, pend, &tmap_out, sbuf, move_bar, lane);
    }
    if (args.trace && warp == 2 && lane == 0) {
      unsigned long long* tr = args.trace + blockIdx.x * 8;
      tr[TR_EPI_WAIT_FULL] = tr_wait;
      tr[TR_EPI_BUSY] = tr_busy;
      tr[TR_EPI_TOTAL] = clock64() - tr_start;
    }
  }

  tc_fence_before();
  cluster_sync_all();   // the peer's shared memory / TMEM must stay alive until the leader's last MMA retired
  if (warp == 1) tmem_dealloc_2cta<kTmemCols>(tmem_base);
}


// ------------------------------------------------------------------ 4-CTA cluster: two CTA pairs sharing B by TMA multicast
// The CTA-pair kernel above moves 64 B/clk into each SM's shared memory at full MMA rate, and all of it is read from
// L2 -- whose slices together deliver ~6300 B/clk (B300_MICROARCH.md: LTS throughput cap), i.e. ~43 B/clk per SM: the
// K=768 GEMMs of the step are L2-output bound at about two thirds of the tensor peak.  Here two pairs of one cluster
// work on vertically adjacent 256x256 tiles (same n_blk, m_blk = 2i and 2i+1), so they need the same B tile: each of
// the four CTAs fetches only HALF of the B half it needs (64 of 128 rows) and multicasts it to the CTA with the same
// role in the other pair.  L2 reads per CTA and k-block drop from 32 KB to 24 KB.
//
// Protocol differences to the pair kernel:
//  * B arrives by cta_group::2 multicast loads: every destination CTA's bytes are signalled on the full barrier of the
//    leader of ITS pair, so a leader still waits for one barrier carrying both CTAs' A and B (2 x 32 KB);
//  * empty_bar collects the commits of BOTH pairs' leaders (a producer's multicast also writes the other pair's smem).
// (AITJ_GEMM_QUAD_RELAY=1 selects the first version of the protocol: plain multicast loads signalling each CTA's own
//  barrier and a relay warp forwarding "my stage landed" to the leader -- kept for A/B.)
template <bool kAMN, bool kBMN>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(64 + 32 * 16, 1)
gemm_bf16_quad_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_aux,
                      const __grid_constant__ CUtensorMap tmap_side, const GemmArgs args, const int relay) {
  constexpr int kEpiW = 16;
  constexpr int kPairM = 256, kPairN = 256;
  constexpr int kStageA = BLOCK_M * BLOCK_K * 2;
  constexpr int kStageB = (kPairN / 2) * BLOCK_K * 2;
  constexpr int kStageBytes = kStageA + kStageB;
  constexpr int kStages = 5;
  constexpr uint32_t kTmemCols = 2 * kPairN;
  constexpr uint32_t kIdesc = make_idesc_bf16(kPairM, kPairN, kAMN ? 1u : 0u, kBMN ? 1u : 0u);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  constexpr int kStagingBytes = kEpiW * 4096;
  uint8_t* staging = smem + kStages * kStageBytes;
  __nv_bfloat16* sbias = reinterpret_cast<__nv_bfloat16*>(staging + kStagingBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + kStagingBytes + 1024);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* peer_full = empty_bar + kStages;     // leader's copy: the pair's other CTA has its stage
  uint64_t* tmem_full = peer_full + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* side_bar = tmem_empty + 3;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();      // 0..3
  const uint32_t pair = crank >> 1;              // which of the two tiles of the cluster
  const uint32_t rank = crank & 1u;              // role inside the pair
  const uint32_t leader_rank = pair << 1;
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_out);
    if (args.flags & EPI_SAVE_PRE) tma_prefetch_desc(&tmap_aux);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);    // own producer's arrive(expect_tx); bytes from own loads + the partner's multicast
      mbar_init(&empty_bar[i], 2);   // both pairs' leaders commit to every CTA of the cluster
      mbar_init(&peer_full[i], 1);   // (leader) arrive of the other CTA's relay
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * kEpiW);
    }
    for (int i = 0; i < kEpiW; ++i) mbar_init(&side_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta<kTmemCols>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_m2 = (args.tiles_m + 1) >> 1;                 // super-rows of two 256-row tiles
  const int num_work = tiles_m2 * args.tiles_n * args.split_k;
  const int cluster_id = blockIdx.x >> 2;
  const int num_clusters = gridDim.x >> 2;
  GemmArgs sched = args;
  sched.tiles_m = tiles_m2;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (all four CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint16_t mask = static_cast<uint16_t>((1u << rank) | (1u << (rank + 2)));   // same role, both pairs
      for (int w = cluster_id; w < num_work; w += num_clusters) {
        const WorkItem wi = decode_work(sched, w);
        const int m0 = (wi.m_blk * 2 + static_cast<int>(pair)) * kPairM + static_cast<int>(rank) * BLOCK_M;
        const int n0 = wi.n_blk * kPairN + static_cast<int>(rank) * (kPairN / 2) + static_cast<int>(pair) * 64;
        for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
          mbar_wait_cluster(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kStageA + pair * 8192;            // my 64 rows of this role's 128-row B half
          if (relay) {
            mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
            if constexpr (!kAMN) {
              tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j)
                tma_load_2d(sa + j * 8192, &tmap_a, &full_bar[stage], m0 + j * 64, kb * BLOCK_K);
            }
            if constexpr (!kBMN) {
              tma_load_2d_mcast(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n0, mask);    // box 64 (k) x 64 (n rows)
            } else {
              tma_load_2d_mcast(sb, &tmap_b, &full_bar[stage], n0, kb * BLOCK_K, mask);    // box 64 (n) x 64 (k rows)
            }
          } else {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);
            const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), leader_rank);
            if constexpr (!kAMN) {
              tma_load_2d_2sm(sa, &tmap_a, full_leader, kb * BLOCK_K, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j)
                tma_load_2d_2sm(sa + j * 8192, &tmap_a, full_leader, m0 + j * 64, kb * BLOCK_K);
            }
            if constexpr (!kBMN) {
              tma_load_2d_2sm_mcast(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, n0, mask);
            } else {
              tma_load_2d_2sm_mcast(sb, &tmap_b, &full_bar[stage], n0, kb * BLOCK_K, mask);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (leader) {
        // -------------------------------------------------------- MMA issuer (leader of each pair)
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint16_t pair_mask = static_cast<uint16_t>(3u << leader_rank);
        for (int w = cluster_id; w < num_work; w += num_clusters) {
          const WorkItem wi = decode_work(sched, w);
          mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kPairN;
          for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
            mbar_wait_cluster(&full_bar[stage], phase);
            if (relay) mbar_wait_cluster(&peer_full[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
            const uint32_t b_addr = a_addr + kStageA;
#pragma unroll
            for (int ks = 0; ks < BLOCK_K / UMMA_K; ++ks) {
              const uint64_t da = kAMN ? make_sw128_desc(a_addr + ks * 2048, BLOCK_K * 128, 1024)
                                       : make_sw128_desc(a_addr + ks * UMMA_K * 2, 16, 1024);
              const uint64_t db = kBMN ? make_sw128_desc(b_addr + ks * 2048, BLOCK_K * 128, 1024)
                                       : make_sw128_desc(b_addr + ks * UMMA_K * 2, 16, 1024);
              umma_bf16_2cta(d_tmem, da, db, kIdesc, (kb > wi.kb0 || ks > 0) ? 1u : 0u);
            }
            umma_commit_2cta(&empty_bar[stage], 0xF);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          umma_commit_2cta(&tmem_full[acc], pair_mask);
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
      } else if (relay) {
        // -------------------------------------------------------- relay: "this CTA's stage has landed" -> leader
        for (int w = cluster_id; w < num_work; w += num_clusters) {
          const WorkItem wi = decode_work(sched, w);
          for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
            mbar_wait_cluster(&full_bar[stage], phase);
            mbar_arrive_remote(&peer_full[stage], leader_rank);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..17, all four CTAs)
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;
    const int epi_tid = threadIdx.x - 64;
    uint8_t* sbuf = staging + (half * 4 + lg) * 4096;
    StoreGroup sg;
    sg.buf = staging + half * 16384;
    sg.bar_id = 2 + half;
    sg.issuer = lg == 0 && lane == 0;
    sg.grouped = false;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int side_kind = args.flags & (EPI_DGELU | EPI_RESIDUAL);
    const bool side_tma = !(args.flags & (EPI_SAVE_PRE | EPI_MC | EPI_OUT_F32 | EPI_ACCUM | EPI_NO_SIDE_TMA)) &&
                          (side_kind == EPI_DGELU || side_kind == EPI_RESIDUAL);
    uint64_t* my_side_bar = &side_bar[warp - 2];
    uint32_t side_phase = 0;
    for (int w = cluster_id; w < num_work; w += num_clusters) {
      const WorkItem wi = decode_work(sched, w);
      const int m_blk = wi.m_blk * 2 + static_cast<int>(pair);
      bool side_now = false;
      if (side_tma && wi.kb1 > wi.kb0 && wi.n_blk * kPairN + half * 64 < args.N) {
        side_now = true;
        stage_acquire(lane);
        if (lane == 0) {
          mbar_arrive_expect_tx(my_side_bar, 4096);
          tma_load_2d(sbuf, &tmap_side, my_side_bar, wi.n_blk * kPairN + half * 64,
                      m_blk * kPairM + static_cast<int>(rank) * BLOCK_M + lg * 32);
        }
      }
      stage_bias<kPairN, 32 * kEpiW>(args, sbias + acc * 256, wi.n_blk * kPairN, epi_tid);
      mbar_wait_cluster(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (wi.kb1 > wi.kb0 && !(args.flags & EPI_DEBUG_SKIP)) {
        const uint32_t t_acc = tmem_base + acc * kPairN + (static_cast<uint32_t>(lg * 32) << 16);
        const int row0 = m_blk * kPairM + static_cast<int>(rank) * BLOCK_M + lg * 32;
        sg.row0_cta = m_blk * kPairM + static_cast<int>(rank) * BLOCK_M;
        epilogue_cols64(args, &tmap_out, &tmap_aux, &tmap_out, &tmap_aux, t_acc, row0, wi.n_blk * kPairN, half * 64,
                        sbias + acc * 256, sg, lg, lane, side_now ? my_side_bar : nullptr, side_phase);
        if (side_now) side_phase ^= 1u;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tmem_empty[acc], leader_rank);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta<kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D tensor map (128B swizzle): `inner` contiguous elements per row, `outer` rows, row stride `ld` elements.
static int encode_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                     uint32_t box_inner, uint32_t box_outer, bool f32 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -10;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * (f32 ? 4u : 2u)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 100;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

template <int kBlockN, bool kAMN, bool kBMN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tx,
                       const GemmArgs& args, int max_ctas, cudaStream_t stream) {
  constexpr int kStages = (kBlockN == 256) ? 4 : 6;
  constexpr int kStageBytes = BLOCK_M * BLOCK_K * 2 + kBlockN * BLOCK_K * 2;
  constexpr int kSmem = kStages * kStageBytes + kEpiWarps * 4096 + 1024 + 1024 + 256;
  static bool configured = false;
  auto kern = gemm_bf16_tcgen05_kernel<kBlockN, kAMN, kBMN>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return -20;
    configured = true;
  }
  const int num_work = args.tiles_m * args.tiles_n * args.split_k;
  int grid = num_work < num_sms() ? num_work : num_sms();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  kern<<<grid, kGemmThreads, kSmem, stream>>>(ta, tb, to, tx, args);
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -30;
}

static int pair_epilogue_warps() {
  // 16 (default): four epilogue warps per SM sub-partition, 5-stage ring.  8: two per sub-partition, 6 stages.
  static int w = 0;
  if (!w) {
    const char* e = getenv("AITJ_GEMM_EPI_WARPS");
    w = (e && atoi(e) == 8) ? 8 : 16;
  }
  return w;
}

template <bool kAMN, bool kBMN, int kEpiW, bool kPeer = false>
static int launch_gemm_2cta(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tx,
                            const CUtensorMap& to128, const CUtensorMap& tx128, const CUtensorMap& tside,
                            const GemmArgs& args, int max_ctas, cudaStream_t stream) {
  constexpr int kSmem = (kEpiW == 16 ? 5 : 6) * (BLOCK_M * BLOCK_K * 2 + 128 * BLOCK_K * 2) + kEpiW * 4096 + 1024 +
                        1024 + 256;
  static bool configured = false;
  auto kern = gemm_bf16_2cta_kernel<kAMN, kBMN, kEpiW, kPeer>;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess) return -20;
    configured = true;
  }
  const int num_work = args.tiles_m * args.tiles_n * args.split_k;
  int pairs = num_sms() / 2;
  if (num_work < pairs) pairs = num_work;
  if (max_ctas > 1 && pairs > max_ctas / 2) pairs = max_ctas / 2;
  if (pairs < 1) pairs = 1;
  kern<<<pairs * 2, 64 + 32 * kEpiW, kSmem, stream>>>(ta, tb, to, tx, to128, tx128, tside, args);
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -30;
}


static int g_quad_clusters = 0;

template <bool kAMN, bool kBMN>
static int launch_gemm_quad(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tx,
                            const CUtensorMap& tside, const GemmArgs& args, int max_ctas, cudaStream_t stream) {
  constexpr int kSmem = 5 * (BLOCK_M * BLOCK_K * 2 + 128 * BLOCK_K * 2) + 16 * 4096 + 1024 + 1024 + 256;
  static int max_clusters = 0;
  auto kern = gemm_bf16_quad_kernel<kAMN, kBMN>;
  if (!max_clusters) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess) return -20;
    // how many 4-CTA clusters the GPU can hold at once (GPCs whose SM count is not a multiple of 4 strand a few SMs)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms() / 4 * 4);
    cfg.blockDim = dim3(64 + 32 * 16);
    cfg.dynamicSmemBytes = kSmem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = num_sms() / 4; }
    max_clusters = n;
    g_quad_clusters = n;
    if (getenv("AITJ_GEMM_QUAD_CLUSTERS")) max_clusters = atoi(getenv("AITJ_GEMM_QUAD_CLUSTERS"));
  }
  const int num_work = ((args.tiles_m + 1) / 2) * args.tiles_n * args.split_k;
  int clusters = max_clusters;
  if (num_work < clusters) clusters = num_work;
  if (max_ctas > 3 && clusters > max_ctas / 4) clusters = max_ctas / 4;
  if (clusters < 1) clusters = 1;
  static const int relay = getenv("AITJ_GEMM_QUAD_RELAY") && atoi(getenv("AITJ_GEMM_QUAD_RELAY")) != 0;
  kern<<<clusters * 4, 64 + 32 * 16, kSmem, stream>>>(ta, tb, to, tx, tside, args, relay);
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -30;
}



}  // namespace aitj

// ---- EPI_PEER: per-rank tensor maps of a gradient tensor, built once per (tensor, shape) and kept in device memory
#include <map>
#include <tuple>
static aitj::PeerTable g_peer_host;          // host copy of the table installed by aitj_gemm_set_peers
static unsigned int* g_peer_counters = nullptr;      // arrival counters of the split-K "last one moves it" protocol
constexpr long long kPeerCounterSlots = 1 << 16;
static bool g_peer_tma = false;
static std::map<std::tuple<const void*, int, int, int>, CUtensorMap*> g_peer_map_cache;

static const CUtensorMap* peer_maps_for(const void* out, int M, int N, int ldc) {
  auto key = std::make_tuple(out, M, N, ldc);
  auto it = g_peer_map_cache.find(key);
  if (it != g_peer_map_cache.end()) return it->second;
  // first use of this tensor: must happen outside stream capture (the warm-up steps before a graph is captured do it)
  CUtensorMap host[8];
  for (int r = 0; r < g_peer_host.n; ++r) {
    const char* base = reinterpret_cast<const char*>(out) + g_peer_host.delta[r];
    if (aitj::encode_2d(&host[r], base, N, M, ldc, 32, 32, true)) return nullptr;
  }
  CUtensorMap* dev = nullptr;
  if (cudaMalloc(&dev, sizeof(CUtensorMap) * 8) != cudaSuccess) return nullptr;
  if (cudaMemcpy(dev, host, sizeof(CUtensorMap) * g_peer_host.n, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  g_peer_map_cache[key] = dev;
  return dev;
}

static unsigned long long* g_gemm_trace = nullptr;
static float* g_gemm_colsum = nullptr;   // consumed by the next aitj_gemm_bf16 call

extern "C" {

// Debug/profiling: when non-null, the CTA-pair kernel writes 8 x u64 per CTA (TR_* slots, clock64 cycles) showing
// where the MMA issuer, the TMA producer and epilogue warp 2 waited.  The buffer needs 8*8*num_SMs bytes.
int aitj_gemm_set_trace(void* buf) { g_gemm_trace = reinterpret_cast<unsigned long long*>(buf); return 0; }

// The next aitj_gemm_bf16 call (CTA-pair kernel, bf16 output, 16 epilogue warps) also accumulates the column sums of
// its output into buf[N] (fp32).  One-shot: cleared by that call.
int aitj_gemm_set_colsum(void* buf) { g_gemm_colsum = reinterpret_cast<float*>(buf); return 0; }

// Returns 0 on success. See file header for operand conventions. lda/ldb/ldc in elements.
// block_n: 128 or 256 (0 = auto). split_k > 1 requires EPI_ACCUM. max_ctas: 0 = all SMs.
int aitj_gemm_bf16(const void* A, const void* B, void* out, int M, int N, int K, int lda, int ldb, int ldc,
                   int a_mn, int b_mn, const void* bias, const void* residual, void* aux, int flags, int split_k,
                   int block_n, int max_ctas, void* stream_ptr) {
  // block_n == 512 selects the CTA-pair (cta_group::2) kernel: 256x256 tile per 2-CTA cluster.
  using namespace aitj;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((N & 7) || (ldc & 7) || (lda & 7) || (ldb & 7)) return -1;
  if ((reinterpret_cast<uintptr_t>(out) & 15) || ((flags & EPI_SAVE_PRE) && (reinterpret_cast<uintptr_t>(aux) & 15))) return -4;
  if (split_k > 1 && !(flags & EPI_ACCUM)) return -2;
  if ((flags & (EPI_MC | EPI_PEER)) && !(flags & EPI_ACCUM)) return -5;
  // block_n == 1024: two CTA pairs per 4-CTA cluster sharing B by TMA multicast (needs the 16-warp epilogue)
  const bool quad = block_n == 1024 && pair_epilogue_warps() == 16 && !(flags & EPI_GROUP_STORE);
  const bool pair = block_n == 512 || block_n == 1024;
  if (pair) block_n = 256;
  if (block_n == 0) block_n = (N > 128) ? 256 : 128;
  if (block_n != 128 && block_n != 256) return -3;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_ptr);

  GemmArgs args;
  args.M = M; args.N = N; args.K = K; args.ldc = ldc; args.out = out;
  args.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  args.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  args.aux = reinterpret_cast<__nv_bfloat16*>(aux);
  {
    // measured: no gain for plain epilogues, slower for the dGELU one (profiles/ncu_gemm_v2.md) -> opt-in only
    static const bool group_store = getenv("AITJ_GEMM_GROUP_STORE") && atoi(getenv("AITJ_GEMM_GROUP_STORE")) != 0;
    if (group_store && pair && !(flags & (EPI_OUT_F32 | EPI_ACCUM))) flags |= EPI_GROUP_STORE;
    static const bool no_side_tma = getenv("AITJ_GEMM_SIDE_TMA") && atoi(getenv("AITJ_GEMM_SIDE_TMA")) == 0;
    if (no_side_tma) flags |= EPI_NO_SIDE_TMA;
  }
  args.flags = flags;
  args.colsum = g_gemm_colsum;
  g_gemm_colsum = nullptr;
  if (args.colsum && !(pair && pair_epilogue_warps() == 16 && !(flags & (EPI_OUT_F32 | EPI_ACCUM | EPI_GROUP_STORE))))
    return -6;   // only the 16-warp CTA-pair bf16 epilogue implements it
  args.trace = g_gemm_trace;
  args.peer_maps = nullptr;
  args.peer_counters = nullptr;
  args.tiles_m = pair ? (M + 255) / 256 : (M + BLOCK_M - 1) / BLOCK_M;
  args.tiles_n = (N + block_n - 1) / block_n;
  args.k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  if (split_k < 1) split_k = 1;
  if (split_k > args.k_blocks) split_k = args.k_blocks;
  args.k_per_split = (args.k_blocks + split_k - 1) / split_k;
  args.split_k = (args.k_blocks + args.k_per_split - 1) / args.k_per_split;

  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = encode_2d(&ta, A, K, M, lda, BLOCK_K, BLOCK_M);
  else       rc = encode_2d(&ta, A, M, K, lda, 64, BLOCK_K);
  if (rc) return rc;
  if (!b_mn) rc = encode_2d(&tb, B, K, N, ldb, BLOCK_K, quad ? 64 : (pair ? 128 : block_n));
  else       rc = encode_2d(&tb, B, N, K, ldb, 64, BLOCK_K);
  if (rc) return rc - 1000;
  CUtensorMap to, tx;
  const bool out_f32 = (flags & (EPI_OUT_F32 | EPI_ACCUM)) != 0;
  const bool peer_move = (flags & EPI_PEER) && args.split_k > 1 && g_peer_counters != nullptr &&
                         static_cast<long long>((M + 31) / 32) * ((N + 31) / 32) <= kPeerCounterSlots;
  if ((flags & EPI_PEER) && !peer_move && g_peer_tma) {
    const CUtensorMap* pm = peer_maps_for(out, M, N, ldc);
    if (!pm) return -7;
    args.peer_maps = pm;
  }
  if (peer_move) {
    args.peer_counters = g_peer_counters;
    if (g_peer_tma) {
      args.peer_maps = peer_maps_for(out, M, N, ldc);
      if (!args.peer_maps) return -7;
    }
    rc = encode_2d(&to, out, N, M, ldc, 32, 32, true);
    if (rc) return rc - 2000;
  } else if (flags & (EPI_MC | EPI_PEER)) {
    to = ta;   // the multicast / direct peer paths never dereference the local output map
  } else {
    rc = encode_2d(&to, out, N, M, ldc, out_f32 ? 32 : 64, 32, out_f32);
    if (rc) return rc - 2000;
  }
  tx = to;
  if (flags & EPI_SAVE_PRE) {
    rc = encode_2d(&tx, aux, N, M, ldc, 64, 32, false);
    if (rc) return rc - 3000;
  }

  if (pair) {
    CUtensorMap to128 = to, tx128 = tx;
    if (flags & EPI_GROUP_STORE) {
      rc = encode_2d(&to128, out, N, M, ldc, 64, 128, false);
      if (rc) return rc - 4000;
      tx128 = to128;
      if (flags & EPI_SAVE_PRE) {
        rc = encode_2d(&tx128, aux, N, M, ldc, 64, 128, false);
        if (rc) return rc - 5000;
      }
    }
    CUtensorMap tside = to;
    if (!(flags & (EPI_OUT_F32 | EPI_ACCUM)) && (flags & (EPI_DGELU | EPI_RESIDUAL))) {
      const void* side_ptr = (flags & EPI_DGELU) ? static_cast<const void*>(aux) : residual;
      rc = encode_2d(&tside, side_ptr, N, M, ldc, 64, 32, false);
      if (rc) return rc - 6000;
    }
    if (quad) {
      if (!a_mn && !b_mn) return launch_gemm_quad<false, false>(ta, tb, to, tx, tside, args, max_ctas, stream);
      if (!a_mn && b_mn) return launch_gemm_quad<false, true>(ta, tb, to, tx, tside, args, max_ctas, stream);
      if (a_mn && !b_mn) return launch_gemm_quad<true, false>(ta, tb, to, tx, tside, args, max_ctas, stream);
      return launch_gemm_quad<true, true>(ta, tb, to, tx, tside, args, max_ctas, stream);
    }
#define AITJ_PAIR(W)                                                                                  \
    if (!a_mn && !b_mn) return launch_gemm_2cta<false, false, W>(ta, tb, to, tx, to128, tx128, tside, args, max_ctas, stream); \
    if (!a_mn && b_mn) return launch_gemm_2cta<false, true, W>(ta, tb, to, tx, to128, tx128, tside, args, max_ctas, stream);   \
    if (a_mn && !b_mn) return launch_gemm_2cta<true, false, W>(ta, tb, to, tx, to128, tx128, tside, args, max_ctas, stream);   \
    return launch_gemm_2cta<true, true, W>(ta, tb, to, tx, to128, tx128, tside, args, max_ctas, stream);
    if (peer_move && a_mn && b_mn && pair_epilogue_warps() == 16)
      return launch_gemm_2cta<true, true, 16, true>(ta, tb, to, tx, to128, tx128, tside, args, max_ctas, stream);
    if (pair_epilogue_warps() == 16) { AITJ_PAIR(16) }
    AITJ_PAIR(8)
#undef AITJ_PAIR
  }
#define AITJ_DISPATCH(BN)                                                                     \
  if (!a_mn && !b_mn) return launch_gemm<BN, false, false>(ta, tb, to, tx, args, max_ctas, stream);   \
  if (!a_mn && b_mn) return launch_gemm<BN, false, true>(ta, tb, to, tx, args, max_ctas, stream);     \
  if (a_mn && !b_mn) return launch_gemm<BN, true, false>(ta, tb, to, tx, args, max_ctas, stream);     \
  return launch_gemm<BN, true, true>(ta, tb, to, tx, args, max_ctas, stream);
  if (block_n == 256) { AITJ_DISPATCH(256) }
  AITJ_DISPATCH(128)
#undef AITJ_DISPATCH
}

// Ownership table of the symmetric gradient buffer for EPI_PEER (see ptx.cuh: PeerTable); one copy per translation unit.
int aitj_gemm_set_peers(const void* base, const long long* delta, const long long* bound, int n) {
  if (n < 1 || n > 8) return -1;
  aitj::PeerTable t;
  memset(&t, 0, sizeof(t));
  t.base = reinterpret_cast<const float*>(base);
  for (int i = 0; i < n; ++i) t.delta[i] = delta[i];
  for (int i = 0; i <= n; ++i) t.bound[i] = bound[i];
  t.n = n;
  g_peer_host = t;
  if (!g_peer_counters) {
    if (cudaMalloc(&g_peer_counters, sizeof(unsigned int) * kPeerCounterSlots) != cudaSuccess) return -3;
    cudaMemset(g_peer_counters, 0, sizeof(unsigned int) * kPeerCounterSlots);
  }
  {
    const char* e = getenv("AITJ_RS_MOVE");            // 0: every split-K partial crosses NVLink (A/B arm)
    if (e && atoi(e) == 0) { cudaFree(g_peer_counters); g_peer_counters = nullptr; }
  }
  for (auto& kv : g_peer_map_cache) cudaFree(kv.second);     // maps of a previous symmetric allocation
  g_peer_map_cache.clear();
  {
    // AITJ_RS_TMA=0: add from registers instead of bulk reduce-adds through the TMA unit
    const char* e = getenv("AITJ_RS_TMA");
    g_peer_tma = !(e && atoi(e) == 0);
  }
  return cudaMemcpyToSymbol(aitj::c_peers, &t, sizeof(t)) == cudaSuccess ? 0 : -2;
}

int aitj_num_sms() { return aitj::num_sms(); }
// what cudaOccupancyMaxActiveClusters reported for the 4-CTA cluster kernel (0 until it has been launched once)
int aitj_gemm_quad_clusters() { return aitj::g_quad_clusters; }

}  // extern "C"
