// rp_body_fused.cu - fused passes over the tokens for the part of a SASRec block BEFORE the attention (training and inference):
//
//   rp_ln_qkv_fused      q_in = LayerNorm1(x) ;  Q = q_in Wq^T + bq ;  [K | V] = x Wkv^T + bkv            (one pass, x read once)
//   rp_pre_attn_bwd      dq_in = dQ Wq + dh ;  t = LayerNorm1-backward(dq_in) ;  dx = [dK | dV] Wkv + t     (one pass)
//
// Replaces  attention_layernorms[i] + the packed in-projection of torch.nn.MultiheadAttention(query = LN(x), key = value = x)
//   replay/nn/sequential/sasrec/transformer.py:99-106 ; replay/models/nn/sequential/sasrec/model.py:434-435
// and autograd's backward of both.  Round 1 ran LayerNorm + two GEMM launches (x and q_in read three times, 7 [T, d] passes)
// forward and two GEMMs + LayerNorm-backward (9 passes) backward; here every activation tile is read once and written once.
// The d x d weights stay resident in shared memory, 128-token tiles stream through a TMA ring, the LayerNorm runs in the
// epilogue warps on the staged tile (statistics of a row are exchanged between its two column-half threads through shared
// memory), its result goes to TMEM as the packed bf16 A operand of the Q GEMM.  d in {64, 128}.
#include "rp_host.h"
#include "rp_sm100.cuh"

namespace rp {

static constexpr int kBfEpiWarps = 8;
static constexpr int kBfThreads = 64 + kBfEpiWarps * 32;

struct LnQkvParams {
  const float* ln_w;
  const float* ln_b;
  const float* b_in;          // [3d] packed in_proj_bias (q | k | v)
  float eps;
  int T;
  int hd_valid;               // > 0: padded feature slots (rp_sm100.cuh feat_valid): statistics over the real features only
  __nv_bfloat16* q_in;        // [T, d]   LayerNorm output (residual of the block, saved for the backward)
  __nv_bfloat16* Q;           // [T, d]
  __nv_bfloat16* KV;          // [T, 2d]
  float* mean_out;            // [T] or null
  float* rstd_out;
  int kv_only;                // predict, final block: only [K | V] = x Wkv^T + bkv (no LayerNorm, no Q: those run on the B last rows)
};

// pack 8 fp32 accumulator words (+ bias) into 4 bf16 pairs
__device__ __forceinline__ uint4 pack8_bias(const uint32_t* r, const float* bb) {
  return make_uint4(pack_bf16(__uint_as_float(r[0]) + bb[0], __uint_as_float(r[1]) + bb[1]),
                    pack_bf16(__uint_as_float(r[2]) + bb[2], __uint_as_float(r[3]) + bb[3]),
                    pack_bf16(__uint_as_float(r[4]) + bb[4], __uint_as_float(r[5]) + bb[5]),
                    pack_bf16(__uint_as_float(r[6]) + bb[6], __uint_as_float(r[7]) + bb[7]));
}

// Output tiles leave through shared memory and TMA stores: a thread owns one ROW of the tile, so direct global stores are 32
// different 128-byte lines per warp instruction (ncu r2c: l1tex LSU wavefronts 60 % busy = the limiter of round 2's first
// version, 31 sectors per store request); staged in a swizzled [128 x 64] tile the four warps of a column half write
// conflict-free 16-byte pieces and ONE thread hands the tile to the TMA unit.
template <int KCH, int NA>
__global__ void __launch_bounds__(kBfThreads, 1)
ln_qkv_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWq,
                    const __grid_constant__ CUtensorMap tmWkv, const __grid_constant__ CUtensorMap tmOq,
                    const __grid_constant__ CUtensorMap tmOQ, const __grid_constant__ CUtensorMap tmOKV, const LnQkvParams p) {
  constexpr int D = KCH * 64;
  constexpr int WQ_BYTES = KCH * D * 128;        // [D x D] as KCH chunks of [D rows x 64]
  constexpr int WKV_BYTES = KCH * 2 * D * 128;   // [2D x D]
  constexpr int X_STAGE = KCH * 128 * 128;       // [128 x D]
  constexpr int STG = 128 * 128;                 // one [128 rows x 64 columns] bf16 staging tile per column half
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sWq = smem;
  uint8_t* sWkv = smem + WQ_BYTES;
  uint8_t* sX = smem + WQ_BYTES + WKV_BYTES;
  uint8_t* sOut = sX + NA * X_STAGE;
  __shared__ uint64_t bar_w, x_full[NA], x_empty[NA], kv_full, kv_free, q_ready, q_full, q_free;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_lnw[D], s_lnb[D], s_bias[3 * D];
  __shared__ float2 s_stat[2][128];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.T + 127) / 128;
  const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar_w, 1);
    for (int i = 0; i < NA; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1 + kBfEpiWarps);   // the KV GEMM (tcgen05.commit) and the LayerNorm readers
    }
    mbar_init(&kv_full, 1);
    mbar_init(&kv_free, kBfEpiWarps);
    mbar_init(&q_ready, kBfEpiWarps);
    mbar_init(&q_full, 1);
    mbar_init(&q_free, kBfEpiWarps);
    fence_barrier_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWq);
    tma_prefetch_desc(&tmWkv);
    tma_prefetch_desc(&tmOq);
    tma_prefetch_desc(&tmOQ);
    tma_prefetch_desc(&tmOKV);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  if (threadIdx.x >= 64)
    for (int i = threadIdx.x - 64; i < 3 * D; i += kBfEpiWarps * 32) {
      if (i < D && !p.kv_only) {
        s_lnw[i] = p.ln_w[i];
        s_lnb[i] = p.ln_b[i];
      }
      s_bias[i] = p.b_in[i];
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t TA = tmem, TQ = tmem + 128, TKV = tmem + 256;   // packed q_in | Q accumulator | [K | V] accumulator

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bar_w, WQ_BYTES + WKV_BYTES);
      for (int kc = 0; kc < KCH; ++kc) {
        tma_load_2d(sWq + kc * (D * 128), &tmWq, &bar_w, kc * 64, 0);
        tma_load_2d(sWkv + kc * (2 * D * 128), &tmWkv, &bar_w, kc * 64, 0);
      }
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const uint32_t s = it % NA, ph = (it / NA) & 1;
        mbar_wait(&x_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&x_full[s], X_STAGE);
        for (int kc = 0; kc < KCH; ++kc) tma_load_2d(sX + s * X_STAGE + kc * 16384, &tmX, &x_full[s], kc * 64, t * 128);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_kv = umma_idesc_bf16(128, 2 * D);
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, D);
      mbar_wait(&bar_w, 0);
      tc_fence_after();
      for (int it = 0; it < my_tiles; ++it) {
        const uint32_t s = it % NA, ph = (it / NA) & 1, tp = it & 1;
        // [K | V] = x . Wkv^T : needs only the staged tile
        mbar_wait(&x_full[s], ph);
        if (it > 0) mbar_wait(&kv_free, tp ^ 1);
        tc_fence_after();
        {
          const uint32_t a0 = smem_u32(sX + s * X_STAGE), b0 = smem_u32(sWkv);
#pragma unroll
          for (int kc = 0; kc < KCH; ++kc)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_ss(TKV, umma_desc_sw128(a0 + kc * 16384 + ks * 32, 16, 1024),
                      umma_desc_sw128(b0 + kc * (2 * D * 128) + ks * 32, 16, 1024), idesc_kv, (kc | ks) != 0);
        }
        umma_commit(&x_empty[s]);
        umma_commit(&kv_full);
        if (p.kv_only) continue;
        // Q = LN(x) . Wq^T : A operand = packed bf16 q_in written to TMEM by the epilogue warps
        mbar_wait(&q_ready, tp);
        if (it > 0) mbar_wait(&q_free, tp ^ 1);
        tc_fence_after();
        {
          const uint32_t b0 = smem_u32(sWq);
#pragma unroll
          for (int kc = 0; kc < KCH; ++kc)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_ts(TQ, TA + kc * 64 + ks * 8, umma_desc_sw128(b0 + kc * (D * 128) + ks * 32, 16, 1024), idesc_q, (kc | ks) != 0);
        }
        umma_commit(&q_full);
      }
    }
  } else {
    // ------------------------------------------------ epilogue warps: warp%4 = TMEM lane quarter, (warp-2)/4 = column half
    const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const bool has_half = half * 64 < D;
    const int c0 = half * 64;
    uint8_t* stg = sOut + half * STG;                      // this column half's staging tile
    const bool leader = (ew & 3) == 0 && lane == 0;        // issues / tracks this half's TMA stores
    // the row fragment `pk` (64 bf16 of this thread's row) -> staging tile -> global [128 x 64] box at (x, tile * 128)
    auto stage_store = [&](const uint32_t(&pk)[32], const CUtensorMap* tm, int x, int y) {
      if (leader) tma_store_wait_read();                   // the previous store has finished reading the staging tile
      named_bar_sync(2 + half, 128);
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8)
        *reinterpret_cast<uint4*>(stg + sw128_off((uint32_t)row, (uint32_t)c8)) =
            make_uint4(pk[c8 * 4], pk[c8 * 4 + 1], pk[c8 * 4 + 2], pk[c8 * 4 + 3]);
      fence_proxy_async();
      named_bar_sync(2 + half, 128);
      if (leader) {
        tma_store_2d(tm, stg, x, y);
        tma_store_commit();
      }
    };
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t s = it % NA, ph = (it / NA) & 1, tp = it & 1;
      const int t = (int)blockIdx.x + it * (int)gridDim.x;
      const int m = t * 128 + row;
      const bool row_ok = m < p.T;
      if (p.kv_only) {
        // ---- [K | V] only: wait for the accumulator (its GEMM has also released the staged tile on x_empty: arriving after
        // this wait keeps every warp inside the current phase of that barrier), drain, store
        mbar_wait(&kv_full, tp);
        tc_fence_after();
        __syncwarp();
        if (lane == 0) mbar_arrive(&x_empty[s]);
#pragma unroll 1
        for (int cc = 0; cc < D; cc += 64) {
          const int col = half * D + cc;
          uint32_t r0[32], r1[32], pk[32];
          tmem_ld32(TKV + lane_base + col, r0);
          tmem_ld32(TKV + lane_base + col + 32, r1);
          tmem_ld_wait();
          if (cc + 64 >= D) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&kv_free);
          }
          const float* bb = &s_bias[D + col];
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            const uint4 a = pack8_bias(&r0[c8 * 8], bb + c8 * 8), b = pack8_bias(&r1[c8 * 8], bb + 32 + c8 * 8);
            pk[c8 * 4] = a.x; pk[c8 * 4 + 1] = a.y; pk[c8 * 4 + 2] = a.z; pk[c8 * 4 + 3] = a.w;
            pk[16 + c8 * 4] = b.x; pk[16 + c8 * 4 + 1] = b.y; pk[16 + c8 * 4 + 2] = b.z; pk[16 + c8 * 4 + 3] = b.w;
          }
          stage_store(pk, &tmOKV, col, t * 128);
        }
        continue;
      }
      // ---- LayerNorm of this thread's 64 columns of the staged x tile
      mbar_wait(&x_full[s], ph);
      float xv[64];
      float sum = 0.f, sq = 0.f;
      if (has_half) {
        const uint8_t* xt = sX + s * X_STAGE + half * 16384;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          const uint4 v = *reinterpret_cast<const uint4*>(xt + sw128_off((uint32_t)row, (uint32_t)c8));
          const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(v2[e]);
            xv[c8 * 8 + 2 * e] = f.x;
            xv[c8 * 8 + 2 * e + 1] = f.y;
            sum += f.x + f.y;
            sq = fmaf(f.x, f.x, fmaf(f.y, f.y, sq));
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&x_empty[s]);   // this warp no longer reads the staged tile
      s_stat[half][row] = make_float2(sum, sq);
      asm volatile("bar.sync 1, %0;" ::"r"(kBfEpiWarps * 32) : "memory");
      const float2 sa = s_stat[0][row], sb = (D > 64) ? s_stat[1][row] : make_float2(0.f, 0.f);
      const float inv_d = 1.f / (float)feat_count(D, p.hd_valid);   // padded columns are zero: sums need no mask
      const float mean = (sa.x + sb.x) * inv_d;
      const float var = fmaxf((sa.y + sb.y) * inv_d - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.eps);
      asm volatile("bar.sync 1, %0;" ::"r"(kBfEpiWarps * 32) : "memory");   // s_stat is rewritten by the next tile
      if (has_half) {
        uint32_t pk[32];
#pragma unroll
        for (int q = 0; q < 64; q += 2)
          pk[q >> 1] = pack_bf16((xv[q] - mean) * rstd * s_lnw[c0 + q] + s_lnb[c0 + q],
                                 (xv[q + 1] - mean) * rstd * s_lnw[c0 + q + 1] + s_lnb[c0 + q + 1]);
        // TA is free: its last reader (the previous tile's Q GEMM) completed before q_full, which this warp waited for below
        tmem_st16(TA + lane_base + c0, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
        tmem_st16(TA + lane_base + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&pk[16]));
        if (row_ok && half == 0 && p.mean_out) {
          p.mean_out[m] = mean;
          p.rstd_out[m] = rstd;
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_ready);
        stage_store(pk, &tmOq, c0, t * 128);
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_ready);
      }
      // ---- [K | V] = acc + bias: this thread drains columns [half*D, half*D + D) of its row, 64 at a time
      mbar_wait(&kv_full, tp);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < D; cc += 64) {
        const int col = half * D + cc;
        uint32_t r0[32], r1[32], pk[32];
        tmem_ld32(TKV + lane_base + col, r0);
        tmem_ld32(TKV + lane_base + col + 32, r1);
        tmem_ld_wait();
        if (cc + 64 >= D) {   // everything this warp needs of the accumulator sits in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&kv_free);
        }
        const float* bb = &s_bias[D + col];
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const uint4 a = pack8_bias(&r0[c8 * 8], bb + c8 * 8), b = pack8_bias(&r1[c8 * 8], bb + 32 + c8 * 8);
          pk[c8 * 4] = a.x; pk[c8 * 4 + 1] = a.y; pk[c8 * 4 + 2] = a.z; pk[c8 * 4 + 3] = a.w;
          pk[16 + c8 * 4] = b.x; pk[16 + c8 * 4 + 1] = b.y; pk[16 + c8 * 4 + 2] = b.z; pk[16 + c8 * 4 + 3] = b.w;
        }
        stage_store(pk, &tmOKV, col, t * 128);
      }
      // ---- Q = acc + bq
      mbar_wait(&q_full, tp);
      tc_fence_after();
      if (has_half) {
        uint32_t r0[32], r1[32], pk[32];
        tmem_ld32(TQ + lane_base + c0, r0);
        tmem_ld32(TQ + lane_base + c0 + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_free);
        const float* bb = &s_bias[c0];
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const uint4 a = pack8_bias(&r0[c8 * 8], bb + c8 * 8), b = pack8_bias(&r1[c8 * 8], bb + 32 + c8 * 8);
          pk[c8 * 4] = a.x; pk[c8 * 4 + 1] = a.y; pk[c8 * 4 + 2] = a.z; pk[c8 * 4 + 3] = a.w;
          pk[16 + c8 * 4] = b.x; pk[16 + c8 * 4 + 1] = b.y; pk[16 + c8 * 4 + 2] = b.z; pk[16 + c8 * 4 + 3] = b.w;
        }
        stage_store(pk, &tmOQ, c0, t * 128);
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_free);
      }
    }
    if (leader) tma_store_wait_all();   // the staging tile must outlive its last store
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int KCH>
static int launch_ln_qkv(const CUtensorMap& tmX, const CUtensorMap& tmWq, const CUtensorMap& tmWkv, const CUtensorMap& tmOq,
                         const CUtensorMap& tmOQ, const CUtensorMap& tmOKV, const LnQkvParams& p, cudaStream_t st) {
  constexpr int D = KCH * 64;
  constexpr int NA = 2;
  const int smem = 3 * KCH * D * 128 + NA * KCH * 128 * 128 + 2 * 128 * 128 + 1024;
  auto kern = ln_qkv_fused_kernel<KCH, NA>;
  RP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int n_tiles = (p.T + 127) / 128;
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  kern<<<grid, kBfThreads, smem, st>>>(tmX, tmWq, tmWkv, tmOq, tmOQ, tmOKV, p);
  RP_LAUNCH_CHECK();
  return RP_OK;
}

}  // namespace rp

using namespace rp;

// x bf16 [T, d]; w_in bf16 [3d, d] (packed in_proj_weight: rows [0,d) = Wq, [d,3d) = Wk | Wv), b_in fp32 [3d]; ln_w / ln_b fp32 [d].
// Outputs: q_in bf16 [T, d] = LayerNorm(x), Q bf16 [T, d] = q_in Wq^T + bq, KV bf16 [T, 2d] = x [Wk | Wv]^T + [bk | bv],
// mean / rstd fp32 [T] (optional, both or none).  No output may alias x.  d in {64, 128}.
RP_API int rp_ln_qkv_fused(const void* x, const float* ln_w, const float* ln_b, float eps, const void* w_in, const float* b_in,
                           int T, int d, void* q_in, void* Q, void* KV, float* mean_out, float* rstd_out, int hd_valid,
                           void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const bool kv_only = q_in == nullptr && Q == nullptr;   // [K | V] projection alone (LayerNorm parameters not read)
  if (!x || !w_in || !b_in || !KV || T <= 0) return RP_EINVAL;
  if (!kv_only && (!ln_w || !ln_b || !q_in || !Q)) return RP_EINVAL;
  if ((mean_out == nullptr) != (rstd_out == nullptr)) return RP_EINVAL;
  if (kv_only) {
    q_in = KV;   // tensor maps need a valid address; nothing is stored through them in this mode
    Q = KV;
    mean_out = rstd_out = nullptr;
  }
  if (d != 64 && d != 128) return RP_ESHAPE;
  if (hd_valid < 0 || hd_valid > 128 || (hd_valid > 0 && d % (hd_valid <= 64 ? 64 : 128))) return RP_ESHAPE;
  if ((!kv_only && (q_in == x || Q == x)) || KV == x) return RP_EINVAL;
  CUtensorMap tmX, tmWq, tmWkv, tmOq, tmOQ, tmOKV;
  int rc;
  if ((rc = make_tmap_bf16(&tmX, x, T, d, d, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmWq, w_in, d, d, d, d)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmWkv, reinterpret_cast<const __nv_bfloat16*>(w_in) + (size_t)d * d, 2 * d, d, d, 2 * d)) != RP_OK)
    return rc;
  if ((rc = make_tmap_bf16(&tmOq, q_in, T, d, d, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmOQ, Q, T, d, d, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmOKV, KV, T, 2 * d, 2 * d, 128)) != RP_OK) return rc;
  LnQkvParams p;
  p.ln_w = ln_w; p.ln_b = ln_b; p.b_in = b_in; p.eps = eps; p.T = T; p.hd_valid = hd_valid;
  p.q_in = reinterpret_cast<__nv_bfloat16*>(q_in); p.Q = reinterpret_cast<__nv_bfloat16*>(Q);
  p.KV = reinterpret_cast<__nv_bfloat16*>(KV); p.mean_out = mean_out; p.rstd_out = rstd_out;
  p.kv_only = kv_only ? 1 : 0;
  return d == 64 ? launch_ln_qkv<1>(tmX, tmWq, tmWkv, tmOq, tmOQ, tmOKV, p, stream)
                 : launch_ln_qkv<2>(tmX, tmWq, tmWkv, tmOq, tmOQ, tmOKV, p, stream);
}

// ------------------------------------------------------------------------------------------------------------------
// Backward of the pre-attention part:   dq_in = dQ Wq + dh ;  t = LN1-backward(dq_in ; x, mean, rstd, w) ;  dx = dKV Wkv + t
// (q_in = LN1(x) feeds the Q projection AND is the residual of the block: `x = q + attention(q, x, x)`,
//  replay/nn/sequential/sasrec/transformer.py:99-107, so dh - the gradient of h = q_in + attn - adds to dq_in directly;
//  K and V are projected from the un-normalised x, so their gradient by-passes the LayerNorm.)
// Per 128-token tile two independent GEMMs (contraction over the projection outputs: the weights are read MN-major in place),
// both accumulators in TMEM (two tiles in flight), A operands streamed in 64-column chunks through one TMA ring.  The
// LayerNorm parameter gradients (column sums over the tokens of dq and dq * xhat) are reduced inside each warp with a
// reduce-scatter butterfly (62 shuffles per quantity and tile), accumulated in registers over the CTA's tiles and added to
// the gradient buffers once per CTA.
// ------------------------------------------------------------------------------------------------------------------
struct PreAttnBwdParams {
  const __nv_bfloat16* dh;    // [T, d] gradient of h = q_in + attn wrt h (residual branch into q_in)
  const __nv_bfloat16* x;     // [T, d] input of LayerNorm1
  const float* mean;
  const float* rstd;
  const float* ln_w;
  __nv_bfloat16* dx;          // [T, d]
  float* dln_w;               // [d] +=
  float* dln_b;               // [d] +=
  int T;
  int hd_valid;               // > 0: padded feature slots - statistics over the real features, no gradient into padded inputs
};

// column sums over the 32 rows of a warp: on return lane l holds the sums of columns 2l and 2l+1 in v[0], v[1]
__device__ __forceinline__ void warp_colsum64(float (&v)[64], int lane) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const bool up = lane & 16;
    const float send = up ? v[i] : v[i + 32], keep = up ? v[i + 32] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const bool up = lane & 8;
    const float send = up ? v[i] : v[i + 16], keep = up ? v[i + 16] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 4;
    const float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 2;
    const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 1;
    const float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
}

template <int KCH, int NA>
__global__ void __launch_bounds__(kBfThreads, 1)
pre_attn_bwd_kernel(const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDKV,
                    const __grid_constant__ CUtensorMap tmWq, const __grid_constant__ CUtensorMap tmWkv,
                    const __grid_constant__ CUtensorMap tmDX, const PreAttnBwdParams p) {
  constexpr int D = KCH * 64;
  constexpr int NCH = 3 * KCH;                  // A chunks ([128 x 64]) per tile: KCH of dQ, then 2 KCH of [dK | dV]
  constexpr int WQ_BYTES = KCH * KCH * 8192;    // MN-major B: K chunks (64 output features) x N chunks (64 input features)
  constexpr int WKV_BYTES = 2 * KCH * KCH * 8192;
  constexpr int A_CHUNK = 128 * 128;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sWq = smem;
  uint8_t* sWkv = smem + WQ_BYTES;
  uint8_t* sA = smem + WQ_BYTES + WKV_BYTES;
  uint8_t* sOut = sA + NA * A_CHUNK;   // one [128 x 64] bf16 staging tile per column half: dx leaves through TMA stores
  __shared__ uint64_t bar_w, a_full[NA], a_empty[NA], acc_full[2], acc_free[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_lnw[D];
  __shared__ float2 s_stat[2][128];
  __shared__ float s_red[2][kBfEpiWarps][64];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.T + 127) / 128;
  const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar_w, 1);
    for (int i = 0; i < NA; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_free[i], kBfEpiWarps);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmDQ);
    tma_prefetch_desc(&tmDKV);
    tma_prefetch_desc(&tmWq);
    tma_prefetch_desc(&tmWkv);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  if (threadIdx.x >= 64)
    for (int i = threadIdx.x - 64; i < D; i += kBfEpiWarps * 32) s_lnw[i] = p.ln_w[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;   // parity pp: acc1 (dQ Wq) at pp*256, acc2 (dKV Wkv) at pp*256 + 128

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bar_w, WQ_BYTES + WKV_BYTES);
      for (int kc = 0; kc < KCH; ++kc)
        for (int nc = 0; nc < KCH; ++nc) tma_load_2d(sWq + (kc * KCH + nc) * 8192, &tmWq, &bar_w, nc * 64, kc * 64);
      for (int kc = 0; kc < 2 * KCH; ++kc)
        for (int nc = 0; nc < KCH; ++nc) tma_load_2d(sWkv + (kc * KCH + nc) * 8192, &tmWkv, &bar_w, nc * 64, kc * 64);
      uint32_t it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
        for (int c = 0; c < NCH; ++c, ++it) {
          const uint32_t s = it % NA, ph = (it / NA) & 1;
          mbar_wait(&a_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&a_full[s], A_CHUNK);
          if (c < KCH) tma_load_2d(sA + s * A_CHUNK, &tmDQ, &a_full[s], c * 64, t * 128);
          else tma_load_2d(sA + s * A_CHUNK, &tmDKV, &a_full[s], (c - KCH) * 64, t * 128);
        }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, D, false, true);
      mbar_wait(&bar_w, 0);
      tc_fence_after();
      uint32_t it = 0;
      for (int n = 0; n < my_tiles; ++n) {
        const uint32_t pp = n & 1, pph = (n >> 1) & 1;
        if (n >= 2) mbar_wait(&acc_free[pp], pph ^ 1);
        tc_fence_after();
        for (int c = 0; c < NCH; ++c, ++it) {
          const uint32_t s = it % NA, ph = (it / NA) & 1;
          mbar_wait(&a_full[s], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sA + s * A_CHUNK);
          const bool first = (c < KCH);
          const uint32_t b0 = first ? smem_u32(sWq) + c * (KCH * 8192) : smem_u32(sWkv) + (c - KCH) * (KCH * 8192);
          const uint32_t dcol = tmem + pp * 256 + (first ? 0 : 128);
          const bool fresh = first ? (c == 0) : (c == KCH);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ss(dcol, umma_desc_sw128(a0 + ks * 32, 16, 1024), umma_desc_sw128(b0 + ks * 2048, 8192, 1024), idesc,
                    !(fresh && ks == 0));
          umma_commit(&a_empty[s]);
        }
        umma_commit(&acc_full[pp]);
      }
    }
  } else {
    const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const bool has_half = half * 64 < D;
    const int c0 = half * 64;
    float acc_w0 = 0.f, acc_w1 = 0.f, acc_b0 = 0.f, acc_b1 = 0.f;   // columns c0 + 2*lane, +1 over this warp's rows, all tiles
    uint8_t* stg = sOut + half * (128 * 128);
    const bool leader = (ew & 3) == 0 && lane == 0;
    for (int n = 0; n < my_tiles; ++n) {
      const uint32_t pp = n & 1, pph = (n >> 1) & 1;
      const int t = (int)blockIdx.x + n * (int)gridDim.x;
      const int m = t * 128 + row;
      const bool row_ok = m < p.T;
      // row-side operands first (they do not depend on the GEMMs); x stays packed (32 registers) and is expanded where used
      uint32_t xp[32];
      float dq[64];
      float mean = 0.f, rstd = 0.f;
      if (has_half) {
        if (row_ok) {
          mean = p.mean[m];
          rstd = p.rstd[m];
        }
        const uint4* xr = reinterpret_cast<const uint4*>(p.x + (size_t)(row_ok ? m : 0) * D + c0);
        const uint4* hr = reinterpret_cast<const uint4*>(p.dh + (size_t)(row_ok ? m : 0) * D + c0);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          const uint4 xv = row_ok ? __ldg(xr + c8) : make_uint4(0u, 0u, 0u, 0u);
          const uint4 hv = row_ok ? __ldg(hr + c8) : make_uint4(0u, 0u, 0u, 0u);
          xp[c8 * 4] = xv.x; xp[c8 * 4 + 1] = xv.y; xp[c8 * 4 + 2] = xv.z; xp[c8 * 4 + 3] = xv.w;
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&hv);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 hf = __bfloat1622float2(h2[e]);
            dq[c8 * 8 + 2 * e] = hf.x;
            dq[c8 * 8 + 2 * e + 1] = hf.y;
          }
        }
      }
      const float nmr = row_ok ? -mean * rstd : 0.f, rs_ok = row_ok ? rstd : 0.f;   // xhat = x * rstd - mean * rstd (0 beyond T)
      mbar_wait(&acc_full[pp], pph);
      tc_fence_after();
      float s1 = 0.f, s2 = 0.f;
      if (has_half) {
        {
          uint32_t r0[32];
          tmem_ld32(tmem + lane_base + pp * 256 + c0, r0);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; ++q) dq[q] += __uint_as_float(r0[q]);
          tmem_ld32(tmem + lane_base + pp * 256 + c0 + 32, r0);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; ++q) dq[q + 32] += __uint_as_float(r0[q]);
        }
#pragma unroll
        for (int q = 0; q < 64; q += 2) {
          const float2 xf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xp[q >> 1]));
          const float g0 = dq[q] * s_lnw[c0 + q], g1 = dq[q + 1] * s_lnw[c0 + q + 1];
          s1 += g0 + g1;
          s2 = fmaf(g0, fmaf(xf.x, rs_ok, nmr), fmaf(g1, fmaf(xf.y, rs_ok, nmr), s2));
        }
      }
      s_stat[half][row] = make_float2(s1, s2);
      asm volatile("bar.sync 1, %0;" ::"r"(kBfEpiWarps * 32) : "memory");
      const float2 sa = s_stat[0][row], sb = (D > 64) ? s_stat[1][row] : make_float2(0.f, 0.f);
      const float inv_d = 1.f / (float)feat_count(D, p.hd_valid);
      const float m1 = (sa.x + sb.x) * inv_d, m2 = (sa.y + sb.y) * inv_d;
      asm volatile("bar.sync 1, %0;" ::"r"(kBfEpiWarps * 32) : "memory");
      if (has_half) {
        // dx = dKV Wkv (second accumulator, read 32 columns at a time) + LayerNorm-backward(dq), staged for one TMA store
        if (leader) tma_store_wait_read();   // the previous tile's store has finished reading the staging tile
        named_bar_sync(2 + half, 128);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r2[32];
          tmem_ld32(tmem + lane_base + pp * 256 + 128 + c0 + hh * 32, r2);
          tmem_ld_wait();
          if (hh == 1) {   // both accumulators of this tile have been read
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[pp]);
          }
          {
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              uint32_t w32[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int q = hh * 32 + c8 * 8 + 2 * e;
                const float2 xf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xp[q >> 1]));
                float t0 = rstd * (dq[q] * s_lnw[c0 + q] - m1 - fmaf(xf.x, rs_ok, nmr) * m2);
                float t1 = rstd * (dq[q + 1] * s_lnw[c0 + q + 1] - m1 - fmaf(xf.y, rs_ok, nmr) * m2);
                if (p.hd_valid > 0) {   // padded inputs of the LayerNorm do not exist: no gradient
                  if (!feat_valid(c0 + q, p.hd_valid)) t0 = 0.f;
                  if (!feat_valid(c0 + q + 1, p.hd_valid)) t1 = 0.f;
                }
                w32[e] = pack_bf16(t0 + __uint_as_float(r2[q - hh * 32]), t1 + __uint_as_float(r2[q + 1 - hh * 32]));
              }
              *reinterpret_cast<uint4*>(stg + sw128_off((uint32_t)row, (uint32_t)(hh * 4 + c8))) = make_uint4(w32[0], w32[1], w32[2], w32[3]);
            }
          }
        }
        fence_proxy_async();
        named_bar_sync(2 + half, 128);
        if (leader) {
          tma_store_2d(&tmDX, stg, c0, t * 128);
          tma_store_commit();
        }
        // LayerNorm parameter gradients: column sums over this warp's 32 rows (rows >= T hold zeros)
        {
          float pw[64];
#pragma unroll
          for (int q = 0; q < 64; q += 2) {
            const float2 xf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xp[q >> 1]));
            pw[q] = dq[q] * fmaf(xf.x, rs_ok, nmr);
            pw[q + 1] = dq[q + 1] * fmaf(xf.y, rs_ok, nmr);
          }
          warp_colsum64(pw, lane);
          acc_w0 += pw[0];
          acc_w1 += pw[1];
        }
        warp_colsum64(dq, lane);
        acc_b0 += dq[0];
        acc_b1 += dq[1];
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_free[pp]);
      }
    }
    if (leader) tma_store_wait_all();   // the staging tile must outlive its last store
    // CTA-level reduction over the four lane quarters, then one atomic per column and CTA
    s_red[0][ew][2 * lane] = acc_w0;
    s_red[0][ew][2 * lane + 1] = acc_w1;
    s_red[1][ew][2 * lane] = acc_b0;
    s_red[1][ew][2 * lane + 1] = acc_b1;
    asm volatile("bar.sync 1, %0;" ::"r"(kBfEpiWarps * 32) : "memory");
    const int tid = threadIdx.x - 64;   // 0..255: (quantity, column)
    if (tid < 2 * D) {
      const int qty = tid / D, col = tid % D, hh = col / 64, cc = col % 64;
      float tot = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) tot += s_red[qty][hh * 4 + k][cc];
      atomicAdd((qty == 0 ? p.dln_w : p.dln_b) + col, tot);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int KCH>
static int launch_pre_attn_bwd(const CUtensorMap& tmDQ, const CUtensorMap& tmDKV, const CUtensorMap& tmWq,
                               const CUtensorMap& tmWkv, const CUtensorMap& tmDX, const PreAttnBwdParams& p, cudaStream_t st) {
  constexpr int NA = 5;
  const int smem = 3 * KCH * KCH * 8192 + NA * 128 * 128 + 2 * 128 * 128 + 1024;
  auto kern = pre_attn_bwd_kernel<KCH, NA>;
  RP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int n_tiles = (p.T + 127) / 128;
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  kern<<<grid, kBfThreads, smem, st>>>(tmDQ, tmDKV, tmWq, tmWkv, tmDX, p);
  RP_LAUNCH_CHECK();
  return RP_OK;
}

// dQ bf16 [T, d]; dKV bf16 [T, 2d]; dh, x bf16 [T, d]; mean, rstd fp32 [T] (LayerNorm1 statistics of x); ln_w fp32 [d];
// w_in bf16 [3d, d] (packed in_proj_weight).  Outputs: dx bf16 [T, d] (no aliasing with the inputs); dln_w / dln_b fp32 [d] are
// ACCUMULATED (+=, fp32 atomics: one per column and CTA).  d in {64, 128}.
RP_API int rp_pre_attn_bwd(const void* dQ, const void* dKV, const void* dh, const void* x, const float* mean, const float* rstd,
                           const float* ln_w, const void* w_in, int T, int d, void* dx, float* dln_w, float* dln_b,
                           int hd_valid, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!dQ || !dKV || !dh || !x || !mean || !rstd || !ln_w || !w_in || !dx || !dln_w || !dln_b || T <= 0) return RP_EINVAL;
  if (d != 64 && d != 128) return RP_ESHAPE;
  if (hd_valid < 0 || hd_valid > 128 || (hd_valid > 0 && d % (hd_valid <= 64 ? 64 : 128))) return RP_ESHAPE;
  if (dx == dQ || dx == dKV || dx == dh || dx == x) return RP_EINVAL;
  CUtensorMap tmDQ, tmDKV, tmWq, tmWkv, tmDX;
  int rc;
  if ((rc = make_tmap_bf16(&tmDX, dx, T, d, d, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmDQ, dQ, T, d, d, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmDKV, dKV, T, 2 * d, 2 * d, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmWq, w_in, d, d, d, 64)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmWkv, reinterpret_cast<const __nv_bfloat16*>(w_in) + (size_t)d * d, 2 * d, d, d, 64)) != RP_OK)
    return rc;
  PreAttnBwdParams p;
  p.dh = reinterpret_cast<const __nv_bfloat16*>(dh); p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.mean = mean; p.rstd = rstd; p.ln_w = ln_w; p.dx = reinterpret_cast<__nv_bfloat16*>(dx);
  p.dln_w = dln_w; p.dln_b = dln_b; p.T = T; p.hd_valid = hd_valid;
  return d == 64 ? launch_pre_attn_bwd<1>(tmDQ, tmDKV, tmWq, tmWkv, tmDX, p, stream)
                 : launch_pre_attn_bwd<2>(tmDQ, tmDKV, tmWq, tmWkv, tmDX, p, stream);
}
