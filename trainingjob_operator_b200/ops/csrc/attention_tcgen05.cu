// Flash attention forward for sm_100a (head dim 64, bf16): S = Q K^T and O_j = P_j V_j on tcgen05 tensor cores with the
// accumulators in TMEM, online softmax in registers (one thread per query row, so no shuffles), operands moved by TMA
// straight out of the packed [B*T, 3*H*64] qkv activation -- no split / transpose / contiguous copies -- and the
// output written as [B*T, H*64], the layout the projection GEMM consumes.
//
// One CTA = one 128-row query tile of one (batch, head); two CTAs per SM so that one CTA's exponentials (the MUFU
// pipe is the bound at D=64: 128x128 exps vs 2x256 tensor cycles per key block) overlap the other's MMAs.
//   warp 0    TMA producer: Q once, K double-buffered, V single-buffered (its slot frees when P.V retires)
//   warp 1    TMEM owner + MMA issuer:  S[j+1] = Q K[j+1]^T is issued before P[j] V[j], so the next block's scores are
//             ready as soon as the softmax warps are
//   warps 2-5 softmax: tcgen05.ld of their row of S, P = 2^(s*scale - R) -> bf16 -> 128B-swizzled smem (the A operand of
//             the second MMA).  O accumulates IN TMEM across key blocks (the MMA adds into it); the softmax reference R
//             is allowed to lag: it only moves -- and O / l are only rescaled, by a tcgen05.ld / st round trip -- when
//             a row's running maximum grew by more than 2^8, which after the first blocks is rare.
//
// The reference operator has no GPU code; this kernel belongs to the launched workers' step.
#ifdef AITJ_ATTN_DEBUG
#define AITJ_MBAR_DEBUG 1
#endif
#include "ptx.cuh"

namespace aitj {

constexpr int ATT_BM = 128;
constexpr int ATT_BN = 128;
constexpr int ATT_D = 64;
constexpr int kAttThreads = 192;

struct AttnArgs {
  int B, T, H;
  int causal;
  float scale_log2;   // softmax scale * log2(e)
  float* lse;         // [B, H, T] natural-log sum-exp of the scaled scores (what the backward needs)
  unsigned long long* trace;   // optional: 8 x u64 per CTA, clock64 spent by softmax warp 2 lane 0 in each phase
  int stagger_cycles;          // start delay of the second CTA of every SM
  unsigned int* sm_arrivals;   // [#SMs] never-reset arrival counters: odd arrival on an SM == its second CTA
};
#define ATR_BEGIN(t) long long t = kTrace ? clock64() : 0
#define ATR_ADD(t, accv) do { if (kTrace) accv += clock64() - t; } while (0)

// One 32-column chunk of the exponential pass for one row: p = 2^(min(x*sl2 - R, 64)), row-sum and row-max partials,
// bf16 P into the swizzled A tile.  Written in 8-wide stages -- arguments, then 8 independent MUFU.EX2, then a tree
// sum -- so that eight exponentials are always in flight (ptxas otherwise rotates three registers through
// MUFU -> FSEL and every exponential pays the MUFU latency).  Masked (key > query) entries get argument -1e30 BEFORE
// the exponential, so there is no select behind the MUFU, and only the diagonal block instantiates that code.
template <bool kDiag>
__device__ __forceinline__ void attn_exp_chunk(const uint32_t (&cur)[32], int c, int row, float sl2, float R, float& sum,
                                               float& bmx, uint8_t* prow) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float x[8], t[8], e[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x[k] = __uint_as_float(cur[q * 8 + k]);
      t[k] = fminf(fmaf(x[k], sl2, -R), 64.f);
      if (kDiag && c * 32 + q * 8 + k > row) { t[k] = -1.0e30f; x[k] = -1.0e30f; }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = fast_exp2(t[k]);
    bmx = fmaxf(bmx, fmaxf(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), fmaxf(fmaxf(x[4], x[5]), fmaxf(x[6], x[7]))));
    sum += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
    asm volatile("" : "+f"(sum), "+f"(bmx));    // keep the partial reductions here: ptxas otherwise defers all 128 adds
                                                // to the end of the block and spills the exponentials it needs for them
    const int chunk = (c & 1) * 4 + q;
    *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) =
        make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
  }
}

template <bool B>
struct BoolTag { static constexpr bool value = B; };

struct AttnTile {
  int qt, b, h, row0, n_blocks;
};
__device__ __forceinline__ AttnTile attn_tile(const AttnArgs& a, int t) {
  // tiles are enumerated heaviest first (causal: the last query tile of every (b,h) sees the most key blocks), so the
  // static round-robin over persistent CTAs is a longest-processing-time-first schedule
  const int BH = a.B * a.H, n_qt = a.T / ATT_BM;
  AttnTile x;
  const int w = t / BH, bh = t - w * BH;
  x.qt = a.causal ? n_qt - 1 - w : w;
  x.b = bh / a.H;
  x.h = bh - x.b * a.H;
  x.row0 = x.b * a.T + x.qt * ATT_BM;
  x.n_blocks = a.causal ? x.qt + 1 : a.T / ATT_BN;
  return x;
}

// Static schedule of the persistent CTAs: round r of the heaviest-first tile list is dealt out forward for even r and
// backward for odd r ("snake"), which pairs heavy with light tiles: per-CTA load 21..24 key blocks instead of 19..27 for
// plain round-robin at B=16, H=12, T=1024 on 296 CTAs.
__device__ __forceinline__ int attn_first_tile() { return blockIdx.x; }
__device__ __forceinline__ int attn_next_tile(int, int round) {
  const int n = gridDim.x;
  return round * n + ((round & 1) ? n - 1 - static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x));
}

template <bool kTrace>
__global__ void __launch_bounds__(kAttThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_out,
                const AttnArgs args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;                 // 128 x 64 bf16
  uint8_t* sK = sQ + 16384;           // 2 stages of 128 x 64
  uint8_t* sV = sK + 32768;           // 128 keys x 64
  uint8_t* sP = sV + 16384;           // 128 x 128 bf16 as two K-major 64-key atoms; atom 0 doubles as output staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 32768);
  uint64_t* q_full = bars;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;        // [2]
  uint64_t* k_empty = bars + 4;       // [2]
  uint64_t* v_full = bars + 6;
  uint64_t* v_empty = bars + 7;
  uint64_t* s_full = bars + 8;
  uint64_t* p_ready = bars + 9;
  uint64_t* o_full = bars + 10;       // [2] P.V of block g has been added into O (TMEM columns 128..191): o_full[g & 1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int C = args.H * ATT_D;
  const int n_tiles = args.B * args.H * (args.T / ATT_BM);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_out);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);
    mbar_init(&o_full[0], 1);
    mbar_init(&o_full[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // The two CTAs of an SM are launched together and would run in lock step -- both in their exponential phase (each at
  // half MUFU rate), then both waiting for their next S.  Starting the second one half a block period late makes them
  // alternate instead: one CTA's exponentials overlap the other's MMA round trip.
  if (args.stagger_cycles > 0) {
    __shared__ unsigned int s_second;
    if (threadIdx.x == 0) {
      unsigned int smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      s_second = atomicAdd(args.sm_arrivals + smid, 1u) & 1u;
    }
    __syncthreads();
    if (s_second) {
      const long long t0 = clock64();
      while (clock64() - t0 < args.stagger_cycles) {}
    }
  }

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int g = 0, ti = 0;
      for (int t = attn_first_tile(); t < n_tiles; t = attn_next_tile(t, ++ti)) {
        const AttnTile x = attn_tile(args, t);
        mbar_wait(q_empty, (ti & 1) ^ 1u, 1);            // the previous tile's last S = Q K^T retired
        mbar_arrive_expect_tx(q_full, 16384);
        tma_load_2d(sQ, &tmap_qkv, q_full, x.h * ATT_D, x.row0);
        for (int j = 0; j < x.n_blocks; ++j, ++g) {
          const int s = g & 1;
          const int krow = x.b * args.T + j * ATT_BN;
          mbar_wait(&k_empty[s], ((g >> 1) & 1) ^ 1u, 2);
          mbar_arrive_expect_tx(&k_full[s], 16384);
          tma_load_2d(sK + s * 16384, &tmap_qkv, &k_full[s], C + x.h * ATT_D, krow);
          mbar_wait(v_empty, (g & 1) ^ 1u, 3);
          mbar_arrive_expect_tx(v_full, 16384);
          tma_load_2d(sV, &tmap_qkv, v_full, 2 * C + x.h * ATT_D, krow);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t kIdescS = make_idesc_bf16(ATT_BM, ATT_BN, 0u, 0u);   // Q (K-major) x K (K-major)
      constexpr uint32_t kIdescO = make_idesc_bf16(ATT_BM, ATT_D, 0u, 1u);    // P (K-major) x V (MN-major)
      const uint32_t t_S = tmem_base, t_O = tmem_base + ATT_BN;
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP), v_addr = smem_u32(sV);
      auto issue_s = [&](int gg) {
        const int s = gg & 1;
        mbar_wait(&k_full[s], (gg >> 1) & 1, 4);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + s * 16384);
#pragma unroll
        for (int ks = 0; ks < ATT_D / 16; ++ks)
          umma_bf16(t_S, make_sw128_desc(q_addr + ks * 32, 16, 1024), make_sw128_desc(k_addr + ks * 32, 16, 1024),
                    kIdescS, ks > 0 ? 1u : 0u);
        umma_commit(&k_empty[s]);
        umma_commit(s_full);
      };
      int g = 0, ti = 0;
      for (int t = attn_first_tile(); t < n_tiles; t = attn_next_tile(t, ++ti)) {
        const AttnTile x = attn_tile(args, t);
        mbar_wait(q_full, ti & 1, 5);
        issue_s(g);
        for (int j = 0; j < x.n_blocks; ++j) {
          mbar_wait(p_ready, (g + j) & 1, 6);      // P[j] is in smem and S[j] has been read out of TMEM
          tc_fence_after();
          if (j + 1 < x.n_blocks) issue_s(g + j + 1);
          else umma_commit(q_empty);            // no more S for this tile: Q may be overwritten once they retired
          mbar_wait(v_full, (g + j) & 1, 7);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < ATT_BN / 16; ++ks)
            umma_bf16(t_O, make_sw128_desc(p_addr + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                      make_sw128_desc(v_addr + ks * 2048, 16384, 1024), kIdescO, (j > 0 || ks > 0) ? 1u : 0u);
          umma_commit(v_empty);
          umma_commit(&o_full[(g + j) & 1]);
        }
        g += x.n_blocks;
      }
    }
  } else {
    // ------------------------------------------------------------ softmax (warps 2..5, thread = row)
    const int lg = warp & 3;
    const int row = lg * 32 + lane;                       // row inside the tile == TMEM lane
    const uint32_t t_S = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    const uint32_t t_O = t_S + ATT_BN;
    const float sl2 = args.scale_log2;
    const bool issuer = warp == 2 && lane == 0;           // the one thread that talks to the TMA unit for stores
    long long tr_ws = 0, tr_p1 = 0, tr_p2 = 0, tr_wo = 0, tr_acc = 0, tr_blocks = 0;
    ATR_BEGIN(tr_start);
    // P.V completions alternate between two barriers and every one is waited for, in order, at the latest two blocks
    // late: by then it has happened (the tensor pipe retires in order) and the same barrier cannot have completed
    // again (its next P.V needs this thread's p_ready), so the parity test can neither block nor alias.  With a single
    // barrier a waiter two phases behind sees the parity it is waiting for as "not yet" and deadlocks.
    int o_waited = 0;
    auto wait_o_until = [&](int k) {
      while (o_waited <= k) { mbar_wait(&o_full[o_waited & 1], (o_waited >> 1) & 1, 8); ++o_waited; }
    };
    // rescale this thread's row of O in TMEM by `a` once P.V of global block gb retired -- the rare path
    auto rescale_o = [&](int gb, float a) {
      ATR_BEGIN(t3);
      wait_o_until(gb);
      ATR_ADD(t3, tr_wo);
      ATR_BEGIN(t4);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < ATT_D / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_O + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * a);
        tmem_st_32x32(t_O + c * 32, r);
      }
      tmem_st_wait();
      ATR_ADD(t4, tr_acc);
    };

    int g = 0, ti = 0;
    for (int t = attn_first_tile(); t < n_tiles; t = attn_next_tile(t, ++ti)) {
      const AttnTile x = attn_tile(args, t);
      if (ti > 0) {
        // the previous tile's output was staged in sP: its TMA store must have read it before P is written again
        if (issuer) tma_store_wait_read<0>();
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      // (l, O) are kept relative to the reference `ref` (log2 domain, scaled); pend = a reference move decided at the
      // end of a block and applied once that block's P.V has retired
      float ref = 0.f, l = 0.f, pend_ref = 0.f;
      bool pend = false;
      // the block body exists twice -- plain, and with the key > query mask of the diagonal block (the last one of
      // a causal tile) -- as separate code regions, so that the plain path carries no mask instructions at all
      auto block = [&](int j, auto diag_tag) {
        constexpr bool diag = decltype(diag_tag)::value;
        const int gj = g + j;
        ATR_BEGIN(t0);
        mbar_wait(s_full, gj & 1, 9);
        ATR_ADD(t0, tr_ws);
        ATR_BEGIN(t1);
        tc_fence_after();
        if (j == 0) {
          // first block: exact row maximum as the reference (one extra sweep over TMEM)
          float mx = -1.0e30f;
#pragma unroll 1
          for (int c = 0; c < ATT_BN / 64; ++c) {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(t_S + c * 64, r0);
            tmem_ld_32x32(t_S + c * 64 + 32, r1);
            tmem_ld_wait();
            if (diag) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if (c * 64 + i <= row) mx = fmaxf(mx, __uint_as_float(r0[i]));
                if (c * 64 + 32 + i <= row) mx = fmaxf(mx, __uint_as_float(r1[i]));
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(r0[i]), __uint_as_float(r1[i])));
            }
          }
          ref = mx * sl2;
        } else if (__any_sync(0xffffffffu, pend)) {
          // the previous block pushed some row's maximum more than 2^8 above the reference: move it now
          const float a = pend ? fast_exp2(ref - pend_ref) : 1.f;
          rescale_o(gj - 1, a);
          l *= a;
          if (pend) ref = pend_ref;
          pend = false;
        }
        ATR_ADD(t1, tr_p1);
        ATR_BEGIN(t2);
        // p = 2^(s*scale - ref) (clamped at 2^64: bf16 and fp32 share the exponent range, so a stale reference costs
        // no accuracy as long as nothing overflows), row sum, the block's own maximum, bf16 P into the swizzled A
        // tile.  The TMEM load of the next 32 columns is in flight while the current 32 go through the MUFU pipe.
        float sum, bmx;
#pragma unroll 1
        for (int attempt = 0; attempt < 2; ++attempt) {
          const float R = ref;
          sum = 0.f;
          bmx = -1.0e30f;
          uint32_t ra[32], rb[32];
          tmem_ld_32x32(t_S, ra);
          // P[gj-1] is still being read by its P.V MMA (issued after this block's S, so S-ready does not imply it has
          // retired): it must have before the first store into the P tile.  The wait hides behind the TMEM load.
          if (gj >= 1) wait_o_until(gj - 1);
#pragma unroll
          for (int c = 0; c < ATT_BN / 32; ++c) {
            tmem_ld_wait();
            uint32_t (&cur)[32] = (c & 1) ? rb : ra;
            if (c + 1 < ATT_BN / 32) tmem_ld_32x32(t_S + (c + 1) * 32, (c & 1) ? ra : rb);
            uint8_t* prow = sP + (c >> 1) * 16384 + row * 128;
            attn_exp_chunk<diag>(cur, c, row, sl2, R, sum, bmx, prow);
          }
          if (!__any_sync(0xffffffffu, bmx * sl2 - R > 60.f)) break;
          // some row outgrew its reference by more than 2^60 within this block (the clamp would bite): move the
          // reference to the true maximum first -- (l, O) move with it -- and redo the block exactly
          const float r2 = fmaxf(R, bmx * sl2);
          const float a = fast_exp2(R - r2);
          if (j > 0) rescale_o(gj - 1, a);
          l *= a;
          ref = r2;
        }
        l += sum;
        if (bmx * sl2 - ref > 8.f) { pend = true; pend_ref = bmx * sl2; }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
        ATR_ADD(t2, tr_p2);
      };
      for (int j = 0; j + 1 < x.n_blocks; ++j) block(j, BoolTag<false>{});
      if (args.causal) block(x.n_blocks - 1, BoolTag<true>{});
      else block(x.n_blocks - 1, BoolTag<false>{});
      g += x.n_blocks;
      tr_blocks += x.n_blocks;
      // tile epilogue: O / l, log-sum-exp, bf16 tile staged in P's first atom, one TMA store (not waited for here)
      wait_o_until(g - 1);
      tc_fence_after();
      const float inv = 1.0f / l;
      args.lse[(static_cast<size_t>(x.b) * args.H + x.h) * args.T + x.qt * ATT_BM + row] =
          ref * 0.6931471805599453f + logf(l);
#pragma unroll
      for (int c = 0; c < ATT_D / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_O + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = c * 4 + q;
          *reinterpret_cast<uint4*>(sP + row * 128 + ((chunk ^ (row & 7)) << 4)) =
              make_uint4(pack_bf16x2(__uint_as_float(r[q * 8]) * inv, __uint_as_float(r[q * 8 + 1]) * inv),
                         pack_bf16x2(__uint_as_float(r[q * 8 + 2]) * inv, __uint_as_float(r[q * 8 + 3]) * inv),
                         pack_bf16x2(__uint_as_float(r[q * 8 + 4]) * inv, __uint_as_float(r[q * 8 + 5]) * inv),
                         pack_bf16x2(__uint_as_float(r[q * 8 + 6]) * inv, __uint_as_float(r[q * 8 + 7]) * inv));
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (issuer) {
        tma_store_2d(&tmap_out, sP, x.h * ATT_D, x.row0);
        tma_store_commit();
      }
    }
    if (issuer) tma_store_wait<0>();
    if (kTrace && issuer) {
      unsigned long long* tr = args.trace + static_cast<size_t>(blockIdx.x) * 8;
      tr[0] = clock64() - tr_start; tr[1] = tr_ws; tr[2] = tr_p1; tr[3] = tr_p2; tr[4] = tr_wo; tr[5] = tr_acc;
      tr[6] = tr_blocks;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFnA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int attn_encode_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                          uint32_t box_inner, uint32_t box_outer) {
  static EncodeTiledFnA fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return -10;
    fn = reinterpret_cast<EncodeTiledFnA>(p);
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2u};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 100;
}

}  // namespace aitj

static unsigned long long* g_attn_trace = nullptr;

extern "C" {

int aitj_attn_set_trace(void* buf) { g_attn_trace = reinterpret_cast<unsigned long long*>(buf); return 0; }

// qkv: bf16 [B*T, 3*H*64] (q | k | v, heads contiguous inside each third);  out: bf16 [B*T, H*64];
// lse: fp32 [B, H, T].  T must be a multiple of 128.  scale <= 0 selects 1/sqrt(64).
int aitj_attn_fwd(const void* qkv, void* out, void* lse, int B, int T, int H, int causal, float scale, void* stream_ptr) {
  using namespace aitj;
  if (B <= 0 || T <= 0 || H <= 0) return 0;
  if (T % ATT_BM) return -1;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return -4;
  const int C = H * ATT_D;
  CUtensorMap tq, to;
  int rc = attn_encode_2d(&tq, qkv, 3ull * C, static_cast<uint64_t>(B) * T, 3ull * C, ATT_D, ATT_BN);
  if (rc) return rc;
  rc = attn_encode_2d(&to, out, C, static_cast<uint64_t>(B) * T, C, ATT_D, ATT_BM);
  if (rc) return rc - 1000;
  AttnArgs a;
  a.B = B; a.T = T; a.H = H; a.causal = causal;
  if (scale <= 0.f) scale = 0.125f;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.lse = reinterpret_cast<float*>(lse);
  a.trace = g_attn_trace;
  constexpr int kSmem = 16384 + 32768 + 16384 + 32768 + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess ||
        cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess)
      return -20;
    configured = true;
  }
  static int n_sms = 0;
  if (!n_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  static const int stagger = getenv("AITJ_ATTN_STAGGER") ? atoi(getenv("AITJ_ATTN_STAGGER")) : 1500;
  a.stagger_cycles = stagger;
  static unsigned int* sm_arrivals = nullptr;
  if (!sm_arrivals) {
    if (cudaMalloc(&sm_arrivals, 1024 * sizeof(unsigned int)) != cudaSuccess) return -21;
    cudaMemset(sm_arrivals, 0, 1024 * sizeof(unsigned int));
  }
  a.sm_arrivals = sm_arrivals;
  const int n_tiles = B * H * (T / ATT_BM);
  static const int max_ctas = getenv("AITJ_ATTN_MAX_CTAS") ? atoi(getenv("AITJ_ATTN_MAX_CTAS")) : 0;   // tests
  int n_ctas = n_tiles < 2 * n_sms ? n_tiles : 2 * n_sms;
  if (max_ctas > 0 && n_ctas > max_ctas) n_ctas = max_ctas;
  dim3 grid(n_ctas);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_ptr);
  if (a.trace) attn_fwd_kernel<true><<<grid, kAttThreads, kSmem, st>>>(tq, to, a);
  else attn_fwd_kernel<false><<<grid, kAttThreads, kSmem, st>>>(tq, to, a);
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -30;
}

}  // extern "C"
