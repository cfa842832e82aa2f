// Flash attention backward for sm_100a (head dim 64, bf16): all five GEMMs of the backward on tcgen05 tensor cores with
// their accumulators in TMEM, operands moved by TMA straight out of the packed [B*T, 3*H*64] qkv activation and the
// [B*T, H*64] output gradient, and dK / dV written straight INTO the packed d_qkv gradient that the qkv weight- and
// input-gradient GEMMs consume (no gather / transpose kernel afterwards).
//
// One work item = one 128-key block j of one (batch, head); the CTA walks the query blocks i that see it (i >= j when
// causal) and keeps dK_j and dV_j in TMEM for the whole walk:
//
//   S  = Q_i K_j^T                 P  = 2^(S*scale*log2e - lse_i*log2e)        (lse from the forward kernel)
//   dP = dO_i V_j^T                dS = P o (dP - delta_i) * scale             (delta_i = rowsum(dO_i o O_i))
//   dV_j += P^T dO_i       dK_j += dS^T Q_i       dQ_i = dS K_j  --> fp32 TMA reduce-add into dq_acc
//
//   warp 0      TMA producer: K_j, V_j once per item; Q_i, dO_i double buffered
//   warp 1      TMEM owner + MMA issuer.  S(i+1) is issued as soon as the compute warps have read S(i) out of TMEM and
//               dP(i+1) as soon as they have read dP(i), so the tensor pipe works on the next block's scores while the
//               exponentials / dS of the current one are being computed
//   warps 2-9   compute: thread = query row (TMEM lane), two warps per lane group split the 128 key columns; P and dS go
//               to shared memory as bf16 in the 128B-swizzled layout that serves both as a K-major A operand (dQ = dS K)
//               and as an MN-major A operand (dV = P^T dO, dK = dS^T Q); dQ(i-1) is drained while the MMAs of block i run
//
// TMEM: S 128 | dP 128 | dV 64 | dK 64 | dQ 64 columns (448 of 512).  dQ needs a reduction over key blocks, which is
// done by bulk fp32 reduce-adds into dq_acc (L2 resident: 50 MB per layer); attn_dq_finish_kernel turns it into the bf16
// q-third of d_qkv and clears it for the next layer.
//
// The reference operator has no GPU code; this kernel belongs to the launched workers' step.
#include "ptx.cuh"

namespace aitj {

constexpr int AB_BM = 128;      // queries per block
constexpr int AB_BN = 128;      // keys per block
constexpr int AB_D = 64;
constexpr int kAbComputeWarps = 8;
constexpr int kAbThreads = 64 + 32 * kAbComputeWarps;

struct AttnBwdArgs {
  int B, T, H;
  int causal;
  float scale;          // softmax scale
  float scale_log2;     // scale * log2(e)
  const float* lse;     // [B, H, T] natural-log sum-exp from the forward
  const float* delta;   // [B, H, T] rowsum(dO o O)
};

// TMEM column offsets
constexpr uint32_t AB_TS = 0, AB_TDP = 128, AB_TDV = 256, AB_TDK = 320, AB_TDQ = 384;

struct AbItem {
  int b, h, j, i0, n_it;
};
__device__ __forceinline__ AbItem ab_item(const AttnBwdArgs& a, int t) {
  // heaviest first: under a causal mask key block 0 is seen by every query block
  const int BH = a.B * a.H, n_blk = a.T / AB_BN;
  AbItem x;
  const int w = t / BH, bh = t - w * BH;
  x.j = w;
  x.b = bh / a.H;
  x.h = bh - x.b * a.H;
  x.i0 = a.causal ? x.j : 0;
  x.n_it = n_blk - x.i0;
  return x;
}
__device__ __forceinline__ int ab_next(int round) {
  const int n = gridDim.x;
  return round * n + ((round & 1) ? n - 1 - static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x));
}

__global__ void __launch_bounds__(kAbThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do,
                const __grid_constant__ CUtensorMap tmap_dq, const __grid_constant__ CUtensorMap tmap_dqkv,
                const AttnBwdArgs args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sK = smem;                      // 128 keys x 64
  uint8_t* sV = sK + 16384;
  uint8_t* sQ = sV + 16384;                // 2 stages of 128 queries x 64
  uint8_t* sdO = sQ + 32768;               // 2 stages
  uint8_t* sP = sdO + 32768;               // 2 atoms of 64 keys: [atom][128 q rows][64 keys]
  uint8_t* sdS = sP + 32768;
  uint8_t* staging = sdS + 32768;          // 8 x 4 KB, one per compute warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kAbComputeWarps * 4096);
  uint64_t* kv_full = bars;
  uint64_t* kv_empty = bars + 1;
  uint64_t* qd_full = bars + 2;            // [2]
  uint64_t* qd_empty = bars + 4;           // [2]
  uint64_t* s_full = bars + 6;
  uint64_t* s_free = bars + 7;
  uint64_t* dp_full = bars + 8;
  uint64_t* dp_free = bars + 9;
  uint64_t* p_ready = bars + 10;
  uint64_t* dv_done = bars + 11;
  uint64_t* ds_ready = bars + 12;
  uint64_t* dkq_done = bars + 13;
  uint64_t* dq_free = bars + 14;
  uint64_t* acc_full = bars + 15;
  uint64_t* acc_free = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int C = args.H * AB_D;
  const int n_items = args.B * args.H * (args.T / AB_BN);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_dq);
    tma_prefetch_desc(&tmap_dqkv);
    mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&qd_full[i], 1); mbar_init(&qd_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(s_free, kAbComputeWarps);
    mbar_init(dp_full, 1); mbar_init(dp_free, kAbComputeWarps);
    mbar_init(p_ready, kAbComputeWarps); mbar_init(dv_done, 1);
    mbar_init(ds_ready, kAbComputeWarps); mbar_init(dkq_done, 1);
    mbar_init(dq_free, kAbComputeWarps);
    mbar_init(acc_full, 1); mbar_init(acc_free, kAbComputeWarps);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0, wk = 0;
      for (int t = blockIdx.x; t < n_items; t = ab_next(++wk)) {
        const AbItem x = ab_item(args, t);
        const int krow = x.b * args.T + x.j * AB_BN;
        mbar_wait(kv_empty, (wk & 1) ^ 1u, 1);
        mbar_arrive_expect_tx(kv_full, 32768);
        tma_load_2d(sK, &tmap_qkv, kv_full, C + x.h * AB_D, krow);
        tma_load_2d(sV, &tmap_qkv, kv_full, 2 * C + x.h * AB_D, krow);
        for (int ii = 0; ii < x.n_it; ++ii, ++it) {
          const int s = it & 1;
          const int qrow = x.b * args.T + (x.i0 + ii) * AB_BM;
          mbar_wait(&qd_empty[s], ((it >> 1) & 1) ^ 1u, 2);
          mbar_arrive_expect_tx(&qd_full[s], 32768);
          tma_load_2d(sQ + s * 16384, &tmap_qkv, &qd_full[s], x.h * AB_D, qrow);
          tma_load_2d(sdO + s * 16384, &tmap_do, &qd_full[s], x.h * AB_D, qrow);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t kIdS = make_idesc_bf16(AB_BM, AB_BN, 0u, 0u);     // K-major x K-major (S, dP)
      constexpr uint32_t kIdT = make_idesc_bf16(AB_BN, AB_D, 1u, 1u);      // A^T (MN-major) x MN-major (dV, dK)
      constexpr uint32_t kIdQ = make_idesc_bf16(AB_BM, AB_D, 0u, 1u);      // K-major x MN-major (dQ)
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP), ds_addr = smem_u32(sdS);
      const uint32_t t_S = tmem_base + AB_TS, t_dP = tmem_base + AB_TDP, t_dV = tmem_base + AB_TDV,
                     t_dK = tmem_base + AB_TDK, t_dQ = tmem_base + AB_TDQ;
      auto issue_scores = [&](uint32_t t_dst, uint32_t a_addr, uint32_t b_addr) {     // [128 x 64] x [128 x 64]^T
#pragma unroll
        for (int ks = 0; ks < AB_D / 16; ++ks)
          umma_bf16(t_dst, make_sw128_desc(a_addr + ks * 32, 16, 1024), make_sw128_desc(b_addr + ks * 32, 16, 1024), kIdS,
                    ks > 0 ? 1u : 0u);
      };
      auto issue_transposed = [&](uint32_t t_dst, uint32_t at_addr, uint32_t b_addr, bool acc) {   // A^T[128x128] x B[128x64]
#pragma unroll
        for (int ks = 0; ks < AB_BM / 16; ++ks)
          umma_bf16(t_dst, make_sw128_desc(at_addr + ks * 2048, 16384, 1024), make_sw128_desc(b_addr + ks * 2048, 16384, 1024),
                    kIdT, (acc || ks > 0) ? 1u : 0u);
      };
      int it = 0, wk = 0;
      for (int t = blockIdx.x; t < n_items; t = ab_next(++wk)) {
        const AbItem x = ab_item(args, t);
        mbar_wait(kv_full, wk & 1, 3);
        mbar_wait(acc_free, (wk & 1) ^ 1u, 4);          // the previous item's dK / dV have been read out of TMEM
        tc_fence_after();
        for (int ii = 0; ii < x.n_it; ++ii, ++it) {
          const int s = it & 1;
          const uint32_t q_addr = smem_u32(sQ + s * 16384), do_addr = smem_u32(sdO + s * 16384);
          if (ii == 0) {
            mbar_wait(&qd_full[s], (it >> 1) & 1, 5);
            mbar_wait(s_free, (it & 1) ^ 1u, 6);
            tc_fence_after();
            issue_scores(t_S, q_addr, k_addr);
            umma_commit(s_full);
            mbar_wait(dp_free, (it & 1) ^ 1u, 7);
            tc_fence_after();
            issue_scores(t_dP, do_addr, v_addr);
            umma_commit(dp_full);
          }
          const bool more = ii + 1 < x.n_it;
          const int s2 = (it + 1) & 1;
          const uint32_t q2 = smem_u32(sQ + s2 * 16384), do2 = smem_u32(sdO + s2 * 16384);
          if (more) {
            mbar_wait(&qd_full[s2], ((it + 1) >> 1) & 1, 8);
            mbar_wait(s_free, it & 1, 9);                 // the compute warps hold S(it) in registers now
            tc_fence_after();
            issue_scores(t_S, q2, k_addr);
            umma_commit(s_full);
          }
          mbar_wait(p_ready, it & 1, 10);
          tc_fence_after();
          issue_transposed(t_dV, p_addr, do_addr, ii > 0);            // dV += P^T dO
          umma_commit(dv_done);
          if (more) {
            mbar_wait(dp_free, it & 1, 11);
            tc_fence_after();
            issue_scores(t_dP, do2, v_addr);
            umma_commit(dp_full);
          }
          mbar_wait(ds_ready, it & 1, 12);
          mbar_wait(dq_free, (it & 1) ^ 1u, 13);          // dQ(it-1) has been drained
          tc_fence_after();
          issue_transposed(t_dK, ds_addr, q_addr, ii > 0);            // dK += dS^T Q
#pragma unroll
          for (int ks = 0; ks < AB_BN / 16; ++ks)                     // dQ = dS K
            umma_bf16(t_dQ, make_sw128_desc(ds_addr + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                      make_sw128_desc(k_addr + ks * 2048, 16384, 1024), kIdQ, ks > 0 ? 1u : 0u);
          umma_commit(dkq_done);
          umma_commit(&qd_empty[s]);
        }
        umma_commit(acc_full);
        umma_commit(kv_empty);
      }
    }
  } else {
    // ------------------------------------------------------------ compute warps (2..9)
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;                    // which 64 key columns (and which 32 of the 64 d columns)
    const int row = lg * 32 + lane;                      // query row inside the block == TMEM lane
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    uint8_t* sbuf = staging + (warp - 2) * 4096;
    const float sl2 = args.scale_log2, scale = args.scale;
    constexpr float LOG2E = 1.4426950408889634f;

    // dQ(it) -> fp32 staging -> bulk reduce-add into dq_acc; this warp: 32 rows x 32 d-columns
    auto drain_dq = [&](int qrow0, int h, uint32_t par) {
      mbar_wait(dkq_done, par, 20);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32(lane_base + AB_TDQ + half * 32, r);
      tmem_ld_wait();
      tc_fence_before();
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(sbuf + lane * 128 + ((q ^ (lane & 7)) << 4)) =
            make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(dq_free);
        tma_reduce_add_2d(&tmap_dq, sbuf, h * AB_D + half * 32, qrow0 + lg * 32);
        tma_store_commit();
      }
    };

    int it = 0, wk = 0;
    for (int t = blockIdx.x; t < n_items; t = ab_next(++wk)) {
      const AbItem x = ab_item(args, t);
      const size_t stat_base = (static_cast<size_t>(x.b) * args.H + x.h) * args.T;
      int prev_qrow0 = 0;
      for (int ii = 0; ii < x.n_it; ++ii, ++it) {
        const int i = x.i0 + ii;
        const bool diag = args.causal && i == x.j;
        const float lse2 = args.lse[stat_base + i * AB_BM + row] * LOG2E;
        const float dl = args.delta[stat_base + i * AB_BM + row];
        const int qrow0 = x.b * args.T + i * AB_BM;
        // ---- P = 2^(S*scale*log2e - lse*log2e), bf16, into sP (and kept packed in registers for dS)
        mbar_wait(s_full, it & 1, 21);
        tc_fence_after();
        uint32_t pk[32];                                  // 64 bf16 probabilities of this thread's half row
        {
          uint32_t ra[32], rb[32];
          tmem_ld_32x32(lane_base + AB_TS + half * 64, ra);
          tmem_ld_32x32(lane_base + AB_TS + half * 64 + 32, rb);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              const uint32_t* src = c ? rb : ra;
              float e0 = fast_exp2(fmaf(__uint_as_float(src[k]), sl2, -lse2));
              float e1 = fast_exp2(fmaf(__uint_as_float(src[k + 1]), sl2, -lse2));
              if (diag) {
                const int key = half * 64 + c * 32 + k;
                if (key > row) e0 = 0.f;
                if (key + 1 > row) e1 = 0.f;
              }
              pk[c * 16 + (k >> 1)] = pack_bf16x2(e0, e1);
            }
          }
        }
        mbar_wait(dv_done, (it & 1) ^ 1u, 22);            // dV(it-1) no longer reads sP
        {
          uint8_t* prow = sP + half * 16384 + row * 128;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(prow + ((q ^ (row & 7)) << 4)) =
                make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
        // ---- dQ of the previous block leaves while the tensor pipe works on dV(it) / dP(it+1)
        if (ii > 0) drain_dq(prev_qrow0, x.h, (it - 1) & 1);
        prev_qrow0 = qrow0;
        // ---- dS = P o (dP - delta) * scale, bf16, into sdS
        mbar_wait(dp_full, it & 1, 23);
        tc_fence_after();
        uint32_t dk[32];
        {
          uint32_t ra[32], rb[32];
          tmem_ld_32x32(lane_base + AB_TDP + half * 64, ra);
          tmem_ld_32x32(lane_base + AB_TDP + half * 64 + 32, rb);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dp_free);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              const uint32_t* src = c ? rb : ra;
              const float2 p = unpack_bf16x2(pk[c * 16 + (k >> 1)]);
              const float d0 = p.x * (__uint_as_float(src[k]) - dl) * scale;
              const float d1 = p.y * (__uint_as_float(src[k + 1]) - dl) * scale;
              dk[c * 16 + (k >> 1)] = pack_bf16x2(d0, d1);
            }
          }
        }
        if (ii == 0) mbar_wait(dkq_done, (it & 1) ^ 1u, 24);   // (ii > 0: drain_dq above already waited for it)
        {
          uint8_t* drow = sdS + half * 16384 + row * 128;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(drow + ((q ^ (row & 7)) << 4)) =
                make_uint4(dk[q * 4], dk[q * 4 + 1], dk[q * 4 + 2], dk[q * 4 + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(ds_ready);
      }
      drain_dq(prev_qrow0, x.h, (it - 1) & 1);
      // ---- item epilogue: dV (warps of half 0) and dK (half 1) -> bf16 -> packed d_qkv
      mbar_wait(acc_full, wk & 1, 25);
      tc_fence_after();
      {
        const uint32_t t_src = lane_base + (half == 0 ? AB_TDV : AB_TDK);
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(t_src, r0);
        tmem_ld_32x32(t_src + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_free);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t* src = q < 4 ? r0 : r1;
          const int o = (q & 3) * 8;
          *reinterpret_cast<uint4*>(sbuf + lane * 128 + ((q ^ (lane & 7)) << 4)) =
              make_uint4(pack_bf16x2(__uint_as_float(src[o]), __uint_as_float(src[o + 1])),
                         pack_bf16x2(__uint_as_float(src[o + 2]), __uint_as_float(src[o + 3])),
                         pack_bf16x2(__uint_as_float(src[o + 4]), __uint_as_float(src[o + 5])),
                         pack_bf16x2(__uint_as_float(src[o + 6]), __uint_as_float(src[o + 7])));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int krow0 = x.b * args.T + x.j * AB_BN + lg * 32;
          tma_store_2d(&tmap_dqkv, sbuf, (half == 0 ? 2 * C : C) + x.h * AB_D, krow0);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// delta[b, h, t] = sum_d dO[b*T+t, h*64+d] * O[b*T+t, h*64+d]; one thread per (row, head): consecutive threads read
// consecutive 128-byte segments
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ dO,
                                                         const __nv_bfloat16* __restrict__ O, float* __restrict__ delta,
                                                         int B, int T, int H) {
  const size_t n = static_cast<size_t>(B) * T * H;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t rowi = idx / H;
    const int h = static_cast<int>(idx - rowi * H);
    const uint4* a = reinterpret_cast<const uint4*>(dO + idx * AB_D);
    const uint4* b = reinterpret_cast<const uint4*>(O + idx * AB_D);
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint4 u = a[q], v = b[q];
      const uint32_t uw[4] = {u.x, u.y, u.z, u.w}, vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 x = unpack_bf16x2(uw[j]), y = unpack_bf16x2(vw[j]);
        s = fmaf(x.x, y.x, fmaf(x.y, y.y, s));
      }
    }
    const size_t bi = rowi / T;
    const int tt = static_cast<int>(rowi - bi * T);
    delta[(bi * H + h) * T + tt] = s;
  }
}

// d_qkv[:, 0:C] <- bf16(dq_acc); dq_acc <- 0 (ready for the next layer's backward)
__global__ void __launch_bounds__(256) attn_dq_finish_kernel(float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ d_qkv,
                                                             size_t rows, int C) {
  const int vec_per_row = C / 8;
  const size_t total = rows * vec_per_row;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = idx / vec_per_row;
    const int c = static_cast<int>(idx - r * vec_per_row) * 8;
    float4* src = reinterpret_cast<float4*>(dq_acc + r * C + c);
    const float4 a = src[0], b = src[1];
    src[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    src[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<uint4*>(d_qkv + r * 3 * C + c) =
        make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
  }
}

typedef CUresult (*EncodeTiledFnB)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int ab_encode_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                        uint32_t box_inner, uint32_t box_outer, bool f32) {
  static EncodeTiledFnB fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return -10;
    fn = reinterpret_cast<EncodeTiledFnB>(p);
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * (f32 ? 4u : 2u)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr),
                  dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 100;
}

}  // namespace aitj

extern "C" {

// qkv bf16 [B*T, 3*H*64]; out, d_out bf16 [B*T, H*64]; lse fp32 [B,H,T] (natural log, from aitj_attn_fwd);
// delta fp32 [B,H,T] scratch; dq_acc fp32 [B*T, H*64] scratch that MUST be zero on entry (it is zero again on exit);
// d_qkv bf16 [B*T, 3*H*64] receives dq | dk | dv.  T % 128 == 0.  scale <= 0 selects 1/sqrt(64).
int aitj_attn_bwd(const void* qkv, const void* out, const void* d_out, const void* lse, void* delta, void* dq_acc,
                  void* d_qkv, int B, int T, int H, int causal, float scale, void* stream_ptr) {
  using namespace aitj;
  if (B <= 0 || T <= 0 || H <= 0) return 0;
  if (T % AB_BM) return -1;
  const int C = H * AB_D;
  const uint64_t rows = static_cast<uint64_t>(B) * T;
  CUtensorMap tq, tdo, tdq, tdqkv;
  int rc = ab_encode_2d(&tq, qkv, 3ull * C, rows, 3ull * C, AB_D, AB_BN, false);
  if (rc) return rc;
  rc = ab_encode_2d(&tdo, d_out, C, rows, C, AB_D, AB_BM, false);
  if (rc) return rc - 1000;
  rc = ab_encode_2d(&tdq, dq_acc, C, rows, C, 32, 32, true);
  if (rc) return rc - 2000;
  rc = ab_encode_2d(&tdqkv, d_qkv, 3ull * C, rows, 3ull * C, AB_D, 32, false);
  if (rc) return rc - 3000;
  AttnBwdArgs a;
  a.B = B; a.T = T; a.H = H; a.causal = causal;
  if (scale <= 0.f) scale = 0.125f;
  a.scale = scale;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.lse = reinterpret_cast<const float*>(lse);
  a.delta = reinterpret_cast<const float*>(delta);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_ptr);
  constexpr int kSmem = 16384 * 2 + 32768 * 4 + kAbComputeWarps * 4096 + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem) != cudaSuccess) return -20;
    configured = true;
  }
  static int n_sms = 0;
  if (!n_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const size_t n_stat = static_cast<size_t>(B) * T * H;
  int blocks = static_cast<int>((n_stat + 255) / 256);
  if (blocks > n_sms * 8) blocks = n_sms * 8;
  attn_delta_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(d_out),
                                            reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<float*>(delta), B, T, H);
  const int n_items = B * H * (T / AB_BN);
  const int grid = n_items < n_sms ? n_items : n_sms;
  attn_bwd_kernel<<<grid, kAbThreads, kSmem, st>>>(tq, tdo, tdq, tdqkv, a);
  const size_t nvec = rows * (C / 8);
  blocks = static_cast<int>((nvec + 255) / 256);
  if (blocks > n_sms * 8) blocks = n_sms * 8;
  attn_dq_finish_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<float*>(dq_acc), reinterpret_cast<__nv_bfloat16*>(d_qkv),
                                                rows, C);
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -30;
}

}  // extern "C"
