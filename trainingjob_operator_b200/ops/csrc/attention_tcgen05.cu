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
};
#define ATR_BEGIN(t) long long t = kTrace ? clock64() : 0
#define ATR_ADD(t, accv) do { if (kTrace) accv += clock64() - t; } while (0)

template <bool kTrace>
__global__ void __launch_bounds__(kAttThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_out,
                const AttnArgs args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = smem;                 // 128 x 64 bf16 (reused as the output staging tile)
  uint8_t* sK = sQ + 16384;           // 2 stages of 128 x 64
  uint8_t* sV = sK + 32768;           // 128 keys x 64
  uint8_t* sP = sV + 16384;           // 128 x 128 bf16 as two K-major 64-key atoms
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 32768);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;        // [2]
  uint64_t* k_empty = bars + 3;       // [2]
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 6;
  uint64_t* s_full = bars + 7;
  uint64_t* p_ready = bars + 8;
  uint64_t* o_full = bars + 9;        // P.V of block j has been added into O (TMEM columns 128..191)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int C = args.H * ATT_D;
  const int n_qt = args.T / ATT_BM;
  const int qt = args.causal ? (n_qt - 1 - static_cast<int>(blockIdx.x)) : static_cast<int>(blockIdx.x);  // heavy tiles first
  const int b = blockIdx.y / args.H, h = blockIdx.y - b * args.H;
  const int row0 = b * args.T + qt * ATT_BM;
  const int n_blocks = args.causal ? qt + 1 : args.T / ATT_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_out);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 16384);
      tma_load_2d(sQ, &tmap_qkv, q_full, h * ATT_D, row0);
      for (int j = 0; j < n_blocks; ++j) {
        const int s = j & 1;
        const int krow = b * args.T + j * ATT_BN;
        mbar_wait(&k_empty[s], ((j >> 1) & 1) ^ 1u);
        mbar_arrive_expect_tx(&k_full[s], 16384);
        tma_load_2d(sK + s * 16384, &tmap_qkv, &k_full[s], C + h * ATT_D, krow);
        mbar_wait(v_empty, (j & 1) ^ 1u);
        mbar_arrive_expect_tx(v_full, 16384);
        tma_load_2d(sV, &tmap_qkv, v_full, 2 * C + h * ATT_D, krow);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t kIdescS = make_idesc_bf16(ATT_BM, ATT_BN, 0u, 0u);   // Q (K-major) x K (K-major)
      constexpr uint32_t kIdescO = make_idesc_bf16(ATT_BM, ATT_D, 0u, 1u);    // P (K-major) x V (MN-major)
      const uint32_t t_S = tmem_base, t_O = tmem_base + ATT_BN;
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP), v_addr = smem_u32(sV);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        mbar_wait(&k_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + s * 16384);
#pragma unroll
        for (int ks = 0; ks < ATT_D / 16; ++ks)
          umma_bf16(t_S, make_sw128_desc(q_addr + ks * 32, 16, 1024), make_sw128_desc(k_addr + ks * 32, 16, 1024),
                    kIdescS, ks > 0 ? 1u : 0u);
        umma_commit(&k_empty[s]);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_blocks; ++j) {
        mbar_wait(p_ready, j & 1);            // P[j] is in smem and S[j] has been read out of TMEM
        tc_fence_after();
        if (j + 1 < n_blocks) issue_s(j + 1);
        mbar_wait(v_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < ATT_BN / 16; ++ks)
          umma_bf16(t_O, make_sw128_desc(p_addr + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                    make_sw128_desc(v_addr + ks * 2048, 16384, 1024), kIdescO, (j > 0 || ks > 0) ? 1u : 0u);
        umma_commit(v_empty);
        umma_commit(o_full);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax / accumulate (warps 2..5, thread = row)
    const int lg = warp & 3;
    const int row = lg * 32 + lane;                       // row inside the tile == TMEM lane
    const uint32_t t_S = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    const uint32_t t_O = t_S + ATT_BN;
    const float sl2 = args.scale_log2;
    // (l, O) are kept relative to the reference `ref` (log2 domain, already scaled); pend > 0 is a reference move that
    // still has to be applied to them (decided at the end of a block, applied once that block's P.V has retired)
    float ref = 0.f, l = 0.f, pend_ref = 0.f;
    bool pend = false;
    long long tr_ws = 0, tr_p1 = 0, tr_p2 = 0, tr_wo = 0, tr_acc = 0;
    ATR_BEGIN(tr_start);

    // rescale this thread's row of O in TMEM by `a` (after P.V of block jb retired) -- the rare path
    auto rescale_o = [&](int jb, float a) {
      ATR_BEGIN(t3);
      mbar_wait(o_full, jb & 1);
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

    for (int j = 0; j < n_blocks; ++j) {
      const bool diag = args.causal && j == n_blocks - 1;   // key block == query tile: mask key > query
      ATR_BEGIN(t0);
      mbar_wait(s_full, j & 1);
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
        rescale_o(j - 1, a);
        l *= a;
        if (pend) ref = pend_ref;
        pend = false;
      }
      ATR_ADD(t1, tr_p1);
      ATR_BEGIN(t2);
      // p = 2^(s*scale - ref) (clamped at 2^64: bf16 and fp32 share the exponent range, so a stale reference costs no
      // accuracy as long as nothing overflows), row sum, the block's own maximum, bf16 P into the swizzled A tile.
      // The TMEM load of the next 32 columns is in flight while the current 32 go through the MUFU pipe.
      float sum, bmx;
#pragma unroll 1
      for (int attempt = 0; attempt < 2; ++attempt) {
        const float R = ref;
        sum = 0.f;
        bmx = -1.0e30f;
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(t_S, ra);
#pragma unroll
        for (int c = 0; c < ATT_BN / 32; ++c) {
          tmem_ld_wait();
          uint32_t (&cur)[32] = (c & 1) ? rb : ra;
          if (c + 1 < ATT_BN / 32) tmem_ld_32x32(t_S + (c + 1) * 32, (c & 1) ? ra : rb);
          uint8_t* prow = sP + (c >> 1) * 16384 + row * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float e[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int i = q * 8 + k;
              const float x = __uint_as_float(cur[i]);
              float v = fast_exp2(fminf(fmaf(x, sl2, -R), 64.f));
              if (diag && c * 32 + i > row) v = 0.f;
              else bmx = fmaxf(bmx, x);
              e[k] = v;
              sum += v;
            }
            const int chunk = (c & 1) * 4 + q;
            *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) =
                make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]),
                           pack_bf16x2(e[6], e[7]));
          }
        }
        if (!__any_sync(0xffffffffu, bmx * sl2 - R > 60.f)) break;
        // some row outgrew its reference by more than 2^60 within this block (the clamp would bite): move the reference
        // to the true maximum first -- (l, O) move with it -- and redo the block exactly
        const float r2 = fmaxf(R, bmx * sl2);
        const float a = fast_exp2(R - r2);
        if (j > 0) rescale_o(j - 1, a);
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
    }
    // epilogue: O / l, log-sum-exp, bf16 tile through the (now idle) Q buffer and one TMA store
    mbar_wait(o_full, (n_blocks - 1) & 1);
    tc_fence_after();
    if (kTrace && warp == 2 && lane == 0) {
      unsigned long long* tr = args.trace + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 8;
      tr[0] = clock64() - tr_start; tr[1] = tr_ws; tr[2] = tr_p1; tr[3] = tr_p2; tr[4] = tr_wo; tr[5] = tr_acc;
      tr[6] = n_blocks;
    }
    const float inv = 1.0f / l;
    args.lse[(static_cast<size_t>(b) * args.H + h) * args.T + qt * ATT_BM + row] = ref * 0.6931471805599453f + logf(l);
#pragma unroll
    for (int c = 0; c < ATT_D / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(t_O + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int chunk = c * 4 + q;
        *reinterpret_cast<uint4*>(sQ + row * 128 + ((chunk ^ (row & 7)) << 4)) =
            make_uint4(pack_bf16x2(__uint_as_float(r[q * 8]) * inv, __uint_as_float(r[q * 8 + 1]) * inv),
                       pack_bf16x2(__uint_as_float(r[q * 8 + 2]) * inv, __uint_as_float(r[q * 8 + 3]) * inv),
                       pack_bf16x2(__uint_as_float(r[q * 8 + 4]) * inv, __uint_as_float(r[q * 8 + 5]) * inv),
                       pack_bf16x2(__uint_as_float(r[q * 8 + 6]) * inv, __uint_as_float(r[q * 8 + 7]) * inv));
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (warp == 2 && lane == 0) {
      tma_store_2d(&tmap_out, sQ, h * ATT_D, row0);
      tma_store_commit();
      tma_store_wait<0>();
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
  dim3 grid(T / ATT_BM, B * H);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_ptr);
  if (a.trace) attn_fwd_kernel<true><<<grid, kAttThreads, kSmem, st>>>(tq, to, a);
  else attn_fwd_kernel<false><<<grid, kAttThreads, kSmem, st>>>(tq, to, a);
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -30;
}

}  // extern "C"
