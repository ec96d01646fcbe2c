// rp_ffn.cu - fused point-wise feed-forward for inference / predict():  out = relu(y W1^T + b1) W2^T + b2 + y  in ONE pass.
//   replaces  SasRecPointWiseFeedForward.forward (eval)   replay/models/nn/sequential/sasrec/model.py:496-506
//             PointWiseFeedForward.forward (eval)          replay/nn/ffn.py:43-57
// predict() is HBM-bound on [T, d] activation passes (T = users x L tokens); two GEMM launches read y, write u, read u, read y
// again (residual) and write the result.  Here both d x d weights stay resident in shared memory, a persistent CTA streams
// 128-token tiles of y through a TMA ring, the hidden activation u never leaves the SM (bf16 written back into TMEM over the
// first accumulator and consumed from there as the A operand of the second MMA) and the residual is read from the y tile that
// is already in shared memory: y is read once, the result written once.  d in {64, 128}.
#include "rp_host.h"
#include "rp_philox.cuh"
#include "rp_sm100.cuh"

namespace rp {

static constexpr int kFfnEpiWarps = 8;
static constexpr int kFfnThreads = 64 + kFfnEpiWarps * 32;

struct FfnParams {
  const float* b1;
  const float* b2;
  const uint8_t* rowmask;   // optional [T]: rows with 0 are written as zeros (legacy SASRec pad rows)
  __nv_bfloat16* out;       // [T, d]
  int T;
};

template <int KCH /* d / 64 */, int NA>
__global__ void __launch_bounds__(kFfnThreads, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const FfnParams p) {
  constexpr int D = KCH * 64;
  constexpr int W_BYTES = KCH * D * 128;      // [D x D] bf16 as KCH chunks of [D rows x 64]
  constexpr int Y_STAGE = KCH * 128 * 128;    // [128 x D]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW1 = smem;
  uint8_t* sW2 = smem + W_BYTES;
  uint8_t* sY = smem + 2 * W_BYTES;
  __shared__ uint64_t bar_w, y_full[NA], y_empty[NA], s1_full[2], u_ready[2], acc2_full[2], acc2_free[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_b1[D], s_b2[D];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.T + 127) / 128;
  if (threadIdx.x == 0) {
    mbar_init(&bar_w, 1);
    for (int i = 0; i < NA; ++i) {
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], kFfnEpiWarps);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s1_full[i], 1);
      mbar_init(&u_ready[i], kFfnEpiWarps);
      mbar_init(&acc2_full[i], 1);
      mbar_init(&acc2_free[i], kFfnEpiWarps);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  if (threadIdx.x >= 64)
    for (int i = threadIdx.x - 64; i < D; i += kFfnEpiWarps * 32) {
      s_b1[i] = p.b1[i];
      s_b2[i] = p.b2[i];
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // TMEM columns: acc1[p] at p*128 (its first D/2 columns per 64-wide half later hold u as packed bf16), acc2[p] at 256+p*128

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bar_w, 2 * W_BYTES);
      for (int kc = 0; kc < KCH; ++kc) {
        tma_load_2d(sW1 + kc * (D * 128), &tmW1, &bar_w, kc * 64, 0);
        tma_load_2d(sW2 + kc * (D * 128), &tmW2, &bar_w, kc * 64, 0);
      }
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const uint32_t s = it % NA, ph = (it / NA) & 1;
        mbar_wait(&y_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&y_full[s], Y_STAGE);
        for (int kc = 0; kc < KCH; ++kc) tma_load_2d(sY + s * Y_STAGE + kc * 16384, &tmY, &y_full[s], kc * 64, t * 128);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, D);
      mbar_wait(&bar_w, 0);
      tc_fence_after();
      const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
      for (int it = 0; it <= my_tiles; ++it) {
        if (it < my_tiles) {  // first GEMM of tile `it`:  acc1 = y . W1^T
          const uint32_t s = it % NA, ph = (it / NA) & 1, pp = it & 1;
          mbar_wait(&y_full[s], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sY + s * Y_STAGE), b0 = smem_u32(sW1);
#pragma unroll
          for (int kc = 0; kc < KCH; ++kc)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_ss(tmem + pp * 128, umma_desc_sw128(a0 + kc * 16384 + ks * 32, 16, 1024),
                      umma_desc_sw128(b0 + kc * (D * 128) + ks * 32, 16, 1024), idesc, (kc | ks) != 0);
          umma_commit(&s1_full[pp]);
        }
        if (it >= 1) {  // second GEMM of tile `it - 1`:  acc2 = u . W2^T, u read from TMEM
          const int jt = it - 1;
          const uint32_t q = jt & 1, qph = (jt >> 1) & 1;
          mbar_wait(&u_ready[q], qph);
          mbar_wait(&acc2_free[q], qph ^ 1);
          tc_fence_after();
          const uint32_t b0 = smem_u32(sW2);
#pragma unroll
          for (int kc = 0; kc < KCH; ++kc)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)  // K step kc*64 + ks*16: packed pairs at column kc*64 + ks*8 of acc1[q]
              umma_ts(tmem + 256 + q * 128, tmem + q * 128 + kc * 64 + ks * 8,
                      umma_desc_sw128(b0 + kc * (D * 128) + ks * 32, 16, 1024), idesc, (kc | ks) != 0);
          umma_commit(&acc2_full[q]);
        }
      }
    }
  } else {
    // ------------------------------------------------ epilogue warps: warp%4 = lane quarter, (warp-2)/4 = 64-column half
    const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const bool has_half = half * 64 < D;  // d = 64: only half 0 carries columns
    const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    for (int it = 0; it <= my_tiles; ++it) {
      if (it < my_tiles) {  // u = relu(acc1 + b1) -> bf16, packed in place over this warp's own (already read) columns
        const uint32_t pp = it & 1, pph = (it >> 1) & 1;
        mbar_wait(&s1_full[pp], pph);
        tc_fence_after();
        if (has_half) {
          const uint32_t base = tmem + lane_base + pp * 128 + half * 64;
          uint32_t r0[32], r1[32];
          tmem_ld32(base, r0);
          tmem_ld32(base + 32, r1);
          tmem_ld_wait();
          uint32_t pk[32];
#pragma unroll
          for (int q = 0; q < 32; q += 2) {
            const float a0 = fmaxf(__uint_as_float(r0[q]) + s_b1[half * 64 + q], 0.f);
            const float a1 = fmaxf(__uint_as_float(r0[q + 1]) + s_b1[half * 64 + q + 1], 0.f);
            const float c0 = fmaxf(__uint_as_float(r1[q]) + s_b1[half * 64 + 32 + q], 0.f);
            const float c1 = fmaxf(__uint_as_float(r1[q + 1]) + s_b1[half * 64 + 32 + q + 1], 0.f);
            pk[q >> 1] = pack_bf16(a0, a1);
            pk[16 + (q >> 1)] = pack_bf16(c0, c1);
          }
          tmem_st16(base, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
          tmem_st16(base + 16, *reinterpret_cast<uint32_t(*)[16]>(&pk[16]));
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&u_ready[pp]);
      }
      if (it >= 1) {  // out = acc2 + b2 + y  (y from the shared-memory tile), rows >= T are not stored
        const int jt = it - 1;
        const uint32_t q = jt & 1, qph = (jt >> 1) & 1, s = jt % NA;
        const int t = (int)blockIdx.x + jt * (int)gridDim.x;
        const int m = t * 128 + row;
        mbar_wait(&acc2_full[q], qph);
        tc_fence_after();
        if (has_half) {
          uint32_t r0[32], r1[32];
          tmem_ld32(tmem + lane_base + 256 + q * 128 + half * 64, r0);
          tmem_ld32(tmem + lane_base + 256 + q * 128 + half * 64 + 32, r1);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc2_free[q]);
          const uint8_t* ytile = sY + s * Y_STAGE + half * 16384;  // 64-column chunk `half` of the y tile
          const float keep = (p.rowmask == nullptr || (m < p.T && p.rowmask[m])) ? 1.f : 0.f;
          if (m < p.T) {
            __nv_bfloat16* o = p.out + (size_t)m * D + half * 64;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {  // 8 columns (16 bytes) per step
              const uint4 yv = *reinterpret_cast<const uint4*>(ytile + sw128_off((uint32_t)row, (uint32_t)c8));
              const __nv_bfloat162* y2 = reinterpret_cast<const __nv_bfloat162*>(&yv);
              uint4 w;
              uint32_t* w32 = reinterpret_cast<uint32_t*>(&w);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int col = c8 * 8 + 2 * e;
                const float2 yf = __bfloat1622float2(y2[e]);
                const float v0 = __uint_as_float(col < 32 ? r0[col] : r1[col - 32]) + s_b2[half * 64 + col] + yf.x;
                const float v1 = __uint_as_float(col + 1 < 32 ? r0[col + 1] : r1[col + 1 - 32]) + s_b2[half * 64 + col + 1] + yf.y;
                w32[e] = pack_bf16(v0 * keep, v1 * keep);
              }
              *reinterpret_cast<uint4*>(o + c8 * 8) = w;
            }
          }
        } else {
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc2_free[q]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&y_empty[s]);  // the y tile (A operand of GEMM 1 and residual) is no longer needed
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// Inference: everything after the attention of one SASRec block in ONE pass over the tokens:
//   h = O Wo^T + bo + q_in ;  y = LayerNorm(h) ;  out = relu(y W1^T + b1) W2^T + b2 + y
// (replaces out-projection GEMM + LayerNorm + the FFN above: h and y never reach HBM - 3 tensors move instead of 7).
// Per 128-token tile three chained MMAs; y and u live in TMEM as packed bf16 over their own accumulators and feed the next MMA
// as its A operand; LayerNorm statistics of a row are split over the two column-half threads and exchanged through shared
// memory.  Two tiles are in flight per CTA (2 x 256 TMEM columns: the third accumulator reuses the first one's columns once y
// has been copied next to u), their MMAs and epilogues interleave; O tiles are prefetched through a TMA ring.
// ------------------------------------------------------------------------------------------------------------------
struct PostAttnParams {
  const float* bo;
  const float* ln_w;
  const float* ln_b;
  const float* b1;
  const float* b2;
  const __nv_bfloat16* q_in;   // residual of the out-projection, [T, d]
  const uint8_t* rowmask;
  __nv_bfloat16* out;
  float eps;
  int T;
  // TRAIN: activations saved for the backward (bf16 [T, d]; u is stored AFTER its dropout, as the un-fused path does) and
  // the two dropouts of the FFN (replay/nn/ffn.py:49-55), regenerated in the backward from (seed, site offset, element index)
  __nv_bfloat16* h_save;
  __nv_bfloat16* y_save;
  __nv_bfloat16* u_save;
  float* mean_out;
  float* rstd_out;
  float drop_p;
  unsigned long long seed, off1, off2;
  const unsigned long long* seed_ptr;
  int hd_valid;                // > 0: padded feature slots (rp_sm100.cuh): LayerNorm statistics over the real features only
};

// keep/scale 4 consecutive elements of one row (activation dropout, rp_philox.cuh): row key of (seed, site, row), the column
// keys of the 4 columns from the shared-memory table
__device__ __forceinline__ void drop4(float& a, float& b, float& c, float& d, uint32_t row_key, const uint32_t* col_keys,
                                      uint32_t thr, float ks) {
  const uint4 k = *reinterpret_cast<const uint4*>(col_keys);
  a = drop_mix(row_key, k.x) >= thr ? a * ks : 0.f;
  b = drop_mix(row_key, k.y) >= thr ? b * ks : 0.f;
  c = drop_mix(row_key, k.z) >= thr ? c * ks : 0.f;
  d = drop_mix(row_key, k.w) >= thr ? d * ks : 0.f;
}

// the four output tensor maps (box [128 rows x 64 columns]): h, y, u (training only) and the block output
struct PostAttnOutMaps {
  CUtensorMap h, y, u, out;
};

template <int KCH, int NA, bool TRAIN>
__global__ void __launch_bounds__(kFfnThreads, 1)
post_attn_fused_kernel(const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmWo,
                       const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                       const __grid_constant__ PostAttnOutMaps om, const PostAttnParams p) {
  constexpr int D = KCH * 64;
  constexpr int W_BYTES = KCH * D * 128;
  constexpr int O_STAGE = KCH * 128 * 128;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sWo = smem;
  uint8_t* sW1 = smem + W_BYTES;
  uint8_t* sW2 = smem + 2 * W_BYTES;
  uint8_t* sO = smem + 3 * W_BYTES;
  uint8_t* sOut = sO + NA * O_STAGE;   // one [128 x 64] bf16 staging tile per column half: outputs leave through TMA stores
  // TWO tiles in flight per CTA (parity p = tile & 1): the three MMAs of one tile alternate with those of the other, so every
  // epilogue (LayerNorm, ReLU, output) overlaps an MMA of the other tile instead of sitting on a serial chain.
  __shared__ uint64_t bar_w, o_full[NA], o_empty[NA], g0_full[2], y_ready[2], g1_full[2], u_ready[2], g2_full[2], tile_done[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_vec[5][D];     // bo, ln_w, ln_b, b1, b2
  __shared__ __align__(16) uint32_t s_ck[D];      // dropout column keys
  __shared__ float2 s_stat[2][128];               // (sum, sum of squares) of each row's column half

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.T + 127) / 128;
  const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int n_pairs = (my_tiles + 1) / 2;
  if (threadIdx.x == 0) {
    mbar_init(&bar_w, 1);
    for (int i = 0; i < NA; ++i) {
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&g0_full[i], 1);
      mbar_init(&y_ready[i], kFfnEpiWarps);
      mbar_init(&g1_full[i], 1);
      mbar_init(&u_ready[i], kFfnEpiWarps);
      mbar_init(&g2_full[i], 1);
      mbar_init(&tile_done[i], kFfnEpiWarps);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmO);
    tma_prefetch_desc(&tmWo);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  if (threadIdx.x >= 64)
    for (int i = threadIdx.x - 64; i < D; i += kFfnEpiWarps * 32) {
      s_vec[0][i] = p.bo[i];
      s_vec[1][i] = p.ln_w[i];
      s_vec[2][i] = p.ln_b[i];
      s_vec[3][i] = p.b1[i];
      s_vec[4][i] = p.b2[i];
      s_ck[i] = drop_col_key((uint32_t)i);
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // TMEM per parity p (256 columns): R0 = p*256 holds acc0, then y (packed bf16 in the first 32 columns of each 64-wide half),
  // finally acc2; R1 = p*256 + 128 holds acc1, then u (packed, first 32 columns of each half) and a copy of y (next 32 columns)

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bar_w, 3 * W_BYTES);
      for (int kc = 0; kc < KCH; ++kc) {
        tma_load_2d(sWo + kc * (D * 128), &tmWo, &bar_w, kc * 64, 0);
        tma_load_2d(sW1 + kc * (D * 128), &tmW1, &bar_w, kc * 64, 0);
        tma_load_2d(sW2 + kc * (D * 128), &tmW2, &bar_w, kc * 64, 0);
      }
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const uint32_t s = it % NA, ph = (it / NA) & 1;
        mbar_wait(&o_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&o_full[s], O_STAGE);
        for (int kc = 0; kc < KCH; ++kc) tma_load_2d(sO + s * O_STAGE + kc * 16384, &tmO, &o_full[s], kc * 64, t * 128);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, D);
      mbar_wait(&bar_w, 0);
      tc_fence_after();
      for (int pi = 0; pi < n_pairs; ++pi) {
        const uint32_t pph = pi & 1;
#pragma unroll 1
        for (int stage = 0; stage < 3; ++stage) {
#pragma unroll 1
          for (int pp = 0; pp < 2; ++pp) {
            const int it = 2 * pi + pp;
            if (it >= my_tiles) continue;
            const uint32_t R0 = tmem + pp * 256, R1 = R0 + 128;
            if (stage == 0) {  // h~ = O . Wo^T
              const uint32_t s = it % NA, ph = (it / NA) & 1;
              if (pi > 0) mbar_wait(&tile_done[pp], pph ^ 1);  // tile it-2 has left R0 / R1
              mbar_wait(&o_full[s], ph);
              tc_fence_after();
              const uint32_t a0 = smem_u32(sO + s * O_STAGE), b0 = smem_u32(sWo);
#pragma unroll
              for (int kc = 0; kc < KCH; ++kc)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_ss(R0, umma_desc_sw128(a0 + kc * 16384 + ks * 32, 16, 1024),
                          umma_desc_sw128(b0 + kc * (D * 128) + ks * 32, 16, 1024), idesc, (kc | ks) != 0);
              umma_commit(&o_empty[s]);
              umma_commit(&g0_full[pp]);
            } else if (stage == 1) {  // y . W1^T, y read from TMEM (R0)
              mbar_wait(&y_ready[pp], pph);
              tc_fence_after();
              const uint32_t b0 = smem_u32(sW1);
#pragma unroll
              for (int kc = 0; kc < KCH; ++kc)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_ts(R1, R0 + kc * 64 + ks * 8, umma_desc_sw128(b0 + kc * (D * 128) + ks * 32, 16, 1024), idesc, (kc | ks) != 0);
              umma_commit(&g1_full[pp]);
            } else {  // u . W2^T, u read from TMEM (R1); the result goes to R0, whose y has been copied to R1 by then
              mbar_wait(&u_ready[pp], pph);
              tc_fence_after();
              const uint32_t b0 = smem_u32(sW2);
#pragma unroll
              for (int kc = 0; kc < KCH; ++kc)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_ts(R0, R1 + kc * 64 + ks * 8, umma_desc_sw128(b0 + kc * (D * 128) + ks * 32, 16, 1024), idesc, (kc | ks) != 0);
              umma_commit(&g2_full[pp]);
            }
          }
        }
      }
    }
  } else {
    const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const bool has_half = half * 64 < D;
    const int c0 = half * 64;
    const float keep_scale = (TRAIN && p.drop_p > 0.f) ? 1.f / (1.f - p.drop_p) : 1.f;
    const uint32_t drop_thr = (TRAIN && p.drop_p > 0.f) ? (uint32_t)(p.drop_p * 4294967296.0) : 0u;
    const unsigned long long seed_eff = TRAIN ? p.seed + ((p.drop_p > 0.f && p.seed_ptr) ? *p.seed_ptr : 0ull) : 0ull;
    // a thread owns one ROW: direct global stores would be 32 different 128-byte lines per warp instruction (the L1 LSU
    // wavefront limit measured in profiles/r2c_body_ncu.md); the row pieces go to a swizzled staging tile instead and one
    // thread per column half hands the [128 x 64] tile to the TMA unit (rows beyond T are clipped by the hardware)
    uint8_t* stg = sOut + half * (128 * 128);
    const bool leader = (ew & 3) == 0 && lane == 0;
    auto stage_store = [&](const uint32_t(&pk)[32], const CUtensorMap* tm, int y) {
      if (leader) tma_store_wait_read();
      named_bar_sync(2 + half, 128);
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8)
        *reinterpret_cast<uint4*>(stg + sw128_off((uint32_t)row, (uint32_t)c8)) =
            make_uint4(pk[c8 * 4], pk[c8 * 4 + 1], pk[c8 * 4 + 2], pk[c8 * 4 + 3]);
      fence_proxy_async();
      named_bar_sync(2 + half, 128);
      if (leader) {
        tma_store_2d(tm, stg, c0, y);
        tma_store_commit();
      }
    };
    for (int pi = 0; pi < n_pairs; ++pi) {
      const uint32_t pph = pi & 1;
#pragma unroll 1
      for (int stage = 0; stage < 3; ++stage) {
#pragma unroll 1
        for (int pp = 0; pp < 2; ++pp) {
          const int it = 2 * pi + pp;
          if (it >= my_tiles) continue;
          const uint32_t R0 = tmem + lane_base + pp * 256, R1 = R0 + 128;
          const int t = (int)blockIdx.x + it * (int)gridDim.x;
          const int m = t * 128 + row;
          const bool row_ok = m < p.T;
          if (stage == 0) {
            // ---- h = O Wo^T + bo + q_in ; LayerNorm ; y -> TMEM (bf16, over this warp's own accumulator columns)
            mbar_wait(&g0_full[pp], pph);
            tc_fence_after();
            float hv[64];
            float sum = 0.f, sq = 0.f;
            if (has_half) {
              uint32_t r0[32], r1[32];
              tmem_ld32(R0 + c0, r0);
              tmem_ld32(R0 + c0 + 32, r1);
              tmem_ld_wait();
              const uint4* qrow = reinterpret_cast<const uint4*>(p.q_in + (size_t)(row_ok ? m : 0) * D + c0);
#pragma unroll
              for (int c8 = 0; c8 < 8; ++c8) {
                const uint4 qv = row_ok ? __ldg(qrow + c8) : make_uint4(0u, 0u, 0u, 0u);
                const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&qv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int col = c8 * 8 + 2 * e;
                  const float2 qf = __bfloat1622float2(q2[e]);
                  float v0 = __uint_as_float(col < 32 ? r0[col] : r1[col - 32]) + s_vec[0][c0 + col] + qf.x;
                  float v1 = __uint_as_float(col + 1 < 32 ? r0[col + 1] : r1[col + 1 - 32]) + s_vec[0][c0 + col + 1] + qf.y;
                  if (TRAIN) {  // the statistics describe exactly the bf16 h the backward will read
                    v0 = __bfloat162float(__float2bfloat16(v0));
                    v1 = __bfloat162float(__float2bfloat16(v1));
                  }
                  hv[col] = v0;
                  hv[col + 1] = v1;
                  sum += v0 + v1;
                  sq = fmaf(v0, v0, fmaf(v1, v1, sq));
                }
              }
              if (TRAIN) {
                uint32_t ph[32];
#pragma unroll
                for (int q = 0; q < 64; q += 2) ph[q >> 1] = pack_bf16(hv[q], hv[q + 1]);
                stage_store(ph, &om.h, t * 128);
              }
            }
            s_stat[half][row] = make_float2(sum, sq);
            asm volatile("bar.sync 1, %0;" ::"r"(kFfnEpiWarps * 32) : "memory");
            const float2 sa = s_stat[0][row], sb = (D > 64) ? s_stat[1][row] : make_float2(0.f, 0.f);
            const float inv_d = 1.f / (float)feat_count(D, p.hd_valid);   // padded columns are zero: sums need no mask
            const float mean = (sa.x + sb.x) * inv_d;
            const float var = fmaxf((sa.y + sb.y) * inv_d - mean * mean, 0.f);
            const float rstd = rsqrtf(var + p.eps);
            asm volatile("bar.sync 1, %0;" ::"r"(kFfnEpiWarps * 32) : "memory");  // s_stat is rewritten by the next tile
            if (has_half) {
              uint32_t pk[32];
#pragma unroll
              for (int q = 0; q < 64; q += 2) {
                const float y0 = (hv[q] - mean) * rstd * s_vec[1][c0 + q] + s_vec[2][c0 + q];
                const float y1 = (hv[q + 1] - mean) * rstd * s_vec[1][c0 + q + 1] + s_vec[2][c0 + q + 1];
                pk[q >> 1] = pack_bf16(y0, y1);
              }
              tmem_st16(R0 + c0, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
              tmem_st16(R0 + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&pk[16]));
              if (TRAIN && row_ok && half == 0) {
                p.mean_out[m] = mean;
                p.rstd_out[m] = rstd;
              }
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&y_ready[pp]);   // the next GEMM starts while y travels to HBM
              if (TRAIN) stage_store(pk, &om.y, t * 128);
            } else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&y_ready[pp]);
            }
          } else if (stage == 1) {
            // ---- u = relu(y W1^T + b1) -> TMEM (R1, first 32 columns of the half); y copied next to it, R0 becomes free
            mbar_wait(&g1_full[pp], pph);
            tc_fence_after();
            if (has_half) {
              uint32_t r0[32], r1[32], yk[32];
              tmem_ld32(R1 + c0, r0);
              tmem_ld32(R1 + c0 + 32, r1);
              tmem_ld32(R0 + c0, yk);
              tmem_ld_wait();
              uint32_t pk[32];
              if (TRAIN && p.drop_p > 0.f) {
                const uint32_t rk1 = drop_row_key(seed_eff, p.off1, (unsigned long long)m);
#pragma unroll
                for (int q = 0; q < 32; q += 4) {
                  float a0 = fmaxf(__uint_as_float(r0[q]) + s_vec[3][c0 + q], 0.f), a1 = fmaxf(__uint_as_float(r0[q + 1]) + s_vec[3][c0 + q + 1], 0.f);
                  float a2 = fmaxf(__uint_as_float(r0[q + 2]) + s_vec[3][c0 + q + 2], 0.f), a3 = fmaxf(__uint_as_float(r0[q + 3]) + s_vec[3][c0 + q + 3], 0.f);
                  float b0 = fmaxf(__uint_as_float(r1[q]) + s_vec[3][c0 + 32 + q], 0.f), b1 = fmaxf(__uint_as_float(r1[q + 1]) + s_vec[3][c0 + 32 + q + 1], 0.f);
                  float b2 = fmaxf(__uint_as_float(r1[q + 2]) + s_vec[3][c0 + 32 + q + 2], 0.f), b3 = fmaxf(__uint_as_float(r1[q + 3]) + s_vec[3][c0 + 32 + q + 3], 0.f);
                  drop4(a0, a1, a2, a3, rk1, s_ck + c0 + q, drop_thr, keep_scale);
                  drop4(b0, b1, b2, b3, rk1, s_ck + c0 + 32 + q, drop_thr, keep_scale);
                  pk[q >> 1] = pack_bf16(a0, a1);
                  pk[(q >> 1) + 1] = pack_bf16(a2, a3);
                  pk[16 + (q >> 1)] = pack_bf16(b0, b1);
                  pk[16 + (q >> 1) + 1] = pack_bf16(b2, b3);
                }
              } else {
#pragma unroll
                for (int q = 0; q < 32; q += 2) {
                  pk[q >> 1] = pack_bf16(fmaxf(__uint_as_float(r0[q]) + s_vec[3][c0 + q], 0.f),
                                         fmaxf(__uint_as_float(r0[q + 1]) + s_vec[3][c0 + q + 1], 0.f));
                  pk[16 + (q >> 1)] = pack_bf16(fmaxf(__uint_as_float(r1[q]) + s_vec[3][c0 + 32 + q], 0.f),
                                                fmaxf(__uint_as_float(r1[q + 1]) + s_vec[3][c0 + 32 + q + 1], 0.f));
                }
              }
              tmem_st16(R1 + c0, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
              tmem_st16(R1 + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&pk[16]));
              tmem_st16(R1 + c0 + 32, *reinterpret_cast<uint32_t(*)[16]>(&yk[0]));
              tmem_st16(R1 + c0 + 48, *reinterpret_cast<uint32_t(*)[16]>(&yk[16]));
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&u_ready[pp]);
              if (TRAIN) stage_store(pk, &om.u, t * 128);
            } else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&u_ready[pp]);
            }
          } else {
            // ---- out = u W2^T + b2 + y   (acc2 in R0, y re-read from its copy in R1)
            mbar_wait(&g2_full[pp], pph);
            tc_fence_after();
            if (has_half) {
              uint32_t r0[32], r1[32], yk[32];
              tmem_ld32(R0 + c0, r0);
              tmem_ld32(R0 + c0 + 32, r1);
              tmem_ld32(R1 + c0 + 32, yk);
              tmem_ld_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tile_done[pp]);  // everything of this tile is in registers
              const float keep = (p.rowmask == nullptr || (row_ok && p.rowmask[m])) ? 1.f : 0.f;
              uint32_t po[32];
              const uint32_t rk2 = (TRAIN && p.drop_p > 0.f) ? drop_row_key(seed_eff, p.off2, (unsigned long long)m) : 0u;
#pragma unroll
              for (int c8 = 0; c8 < 8; ++c8) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int col = c8 * 8 + 2 * e;
                  const float2 yf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yk[col >> 1]));
                  float f0 = __uint_as_float(col < 32 ? r0[col] : r1[col - 32]) + s_vec[4][c0 + col];
                  float f1 = __uint_as_float(col + 1 < 32 ? r0[col + 1] : r1[col + 1 - 32]) + s_vec[4][c0 + col + 1];
                  if (TRAIN && p.drop_p > 0.f) {
                    const uint2 ck = *reinterpret_cast<const uint2*>(s_ck + c0 + col);
                    f0 = drop_mix(rk2, ck.x) >= drop_thr ? f0 * keep_scale : 0.f;
                    f1 = drop_mix(rk2, ck.y) >= drop_thr ? f1 * keep_scale : 0.f;
                  }
                  po[c8 * 4 + e] = pack_bf16((f0 + yf.x) * keep, (f1 + yf.y) * keep);
                }
              }
              stage_store(po, &om.out, t * 128);
            } else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tile_done[pp]);
            }
          }
        }
      }
    }
    if (leader) tma_store_wait_all();   // the staging tile must outlive its last store
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int KCH, bool TRAIN>
static int launch_post_attn(const CUtensorMap& tmO, const CUtensorMap& tmWo, const CUtensorMap& tmW1, const CUtensorMap& tmW2,
                            const PostAttnParams& p, cudaStream_t st) {
  constexpr int D = KCH * 64;
  constexpr int NA = KCH == 1 ? 4 : 2;
  const int smem = 3 * KCH * D * 128 + NA * KCH * 128 * 128 + 2 * 128 * 128 + 1024;
  PostAttnOutMaps om;
  int rc;
  if ((rc = make_tmap_bf16(&om.out, p.out, p.T, D, D, 128)) != RP_OK) return rc;
  om.h = om.y = om.u = om.out;
  if (TRAIN) {
    if ((rc = make_tmap_bf16(&om.h, p.h_save, p.T, D, D, 128)) != RP_OK) return rc;
    if ((rc = make_tmap_bf16(&om.y, p.y_save, p.T, D, D, 128)) != RP_OK) return rc;
    if ((rc = make_tmap_bf16(&om.u, p.u_save, p.T, D, D, 128)) != RP_OK) return rc;
  }
  auto kern = post_attn_fused_kernel<KCH, NA, TRAIN>;
  RP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int n_tiles = (p.T + 127) / 128;
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  kern<<<grid, kFfnThreads, smem, st>>>(tmO, tmWo, tmW1, tmW2, om, p);
  RP_LAUNCH_CHECK();
  return RP_OK;
}

template <int KCH>
static int launch_ffn(const CUtensorMap& tmY, const CUtensorMap& tmW1, const CUtensorMap& tmW2, const FfnParams& p,
                      cudaStream_t st) {
  constexpr int NA = 3;
  constexpr int D = KCH * 64;
  const int smem = 2 * KCH * D * 128 + NA * KCH * 128 * 128 + 1024;
  auto kern = ffn_fused_kernel<KCH, NA>;
  RP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int n_tiles = (p.T + 127) / 128;
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  kern<<<grid, kFfnThreads, smem, st>>>(tmY, tmW1, tmW2, p);
  RP_LAUNCH_CHECK();
  return RP_OK;
}

}  // namespace rp

using namespace rp;

// y, out bf16 [T, d] (out may not alias y); w1, w2 bf16 [d, d] (row = output feature, as torch Linear / Conv1d(k=1) weights);
// b1, b2 fp32 [d]; rowmask optional uint8 [T].  d in {64, 128}.
RP_API int rp_ffn_fused(const void* y, const void* w1, const float* b1, const void* w2, const float* b2,
                        const uint8_t* rowmask, int T, int d, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!y || !w1 || !b1 || !w2 || !b2 || !out || T <= 0) return RP_EINVAL;
  if (d != 64 && d != 128) return RP_ESHAPE;
  if (y == out) return RP_EINVAL;
  CUtensorMap tmY, tmW1, tmW2;
  int rc;
  if ((rc = make_tmap_bf16(&tmY, y, T, d, d, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmW1, w1, d, d, d, d)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmW2, w2, d, d, d, d)) != RP_OK) return rc;
  FfnParams p;
  p.b1 = b1; p.b2 = b2; p.rowmask = rowmask; p.out = reinterpret_cast<__nv_bfloat16*>(out); p.T = T;
  return d == 64 ? launch_ffn<1>(tmY, tmW1, tmW2, p, stream) : launch_ffn<2>(tmY, tmW1, tmW2, p, stream);
}

// Inference: out-projection + residual + LayerNorm + FFN of one SASRec block in one pass (see post_attn_fused_kernel).
//   o, q_in, out bf16 [T, d] (out may not alias o / q_in); wo, w1, w2 bf16 [d, d]; bo, ln_w, ln_b, b1, b2 fp32 [d]; d in {64,128}.
//   replaces (eval)  out_proj + "x = q + a" + LayerNorm + FFN   replay/nn/sequential/sasrec/transformer.py:99-110 ;
//                                                              replay/models/nn/sequential/sasrec/model.py:435-441
RP_API int rp_post_attn_fused(const void* o, const void* q_in, const void* wo, const float* bo, const float* ln_w,
                              const float* ln_b, float eps, const void* w1, const float* b1, const void* w2, const float* b2,
                              const uint8_t* rowmask, int T, int d, void* out, int hd_valid, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!o || !q_in || !wo || !bo || !ln_w || !ln_b || !w1 || !b1 || !w2 || !b2 || !out || T <= 0) return RP_EINVAL;
  if (d != 64 && d != 128) return RP_ESHAPE;
  if (hd_valid < 0 || hd_valid > 128 || (hd_valid > 0 && d % (hd_valid <= 64 ? 64 : 128))) return RP_ESHAPE;
  if (out == o || out == q_in) return RP_EINVAL;
  CUtensorMap tmO, tmWo, tmW1, tmW2;
  int rc;
  if ((rc = make_tmap_bf16(&tmO, o, T, d, d, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmWo, wo, d, d, d, d)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmW1, w1, d, d, d, d)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmW2, w2, d, d, d, d)) != RP_OK) return rc;
  PostAttnParams p;
  p.bo = bo; p.ln_w = ln_w; p.ln_b = ln_b; p.b1 = b1; p.b2 = b2;
  p.q_in = reinterpret_cast<const __nv_bfloat16*>(q_in); p.rowmask = rowmask;
  p.out = reinterpret_cast<__nv_bfloat16*>(out); p.eps = eps; p.T = T;
  p.h_save = p.y_save = p.u_save = nullptr; p.mean_out = p.rstd_out = nullptr;
  p.drop_p = 0.f; p.seed = p.off1 = p.off2 = 0ull; p.seed_ptr = nullptr; p.hd_valid = hd_valid;
  return d == 64 ? launch_post_attn<1, false>(tmO, tmWo, tmW1, tmW2, p, stream)
                 : launch_post_attn<2, false>(tmO, tmWo, tmW1, tmW2, p, stream);
}

// Training forward of everything after the attention of one SASRec block, one pass over the tokens:
//   h = O Wo^T + bo + q_in ; y = LayerNorm(h) ; u = dropout1(relu(y W1^T + b1)) ; out = (y + dropout2(u W2^T + b2)) [* rowmask]
// and the activations the backward needs are written on the way: h, y, u (bf16 [T, d]) and the LayerNorm statistics
// (fp32 [T]) - 2 tensors read, 4 written, against 14 [T, d] passes of the four separate launches
// (out-projection GEMM, LayerNorm, two FFN GEMMs).  Dropout element e of site s uses word (e & 3) of
// drop_mix(drop_row_key(seed + *seed_ptr, off_s, row), drop_col_key(column)): the same stream rp_gemm's epilogue and rp_dropout_bwd use.
//   replaces (train)  replay/nn/sequential/sasrec/transformer.py:99-110 ; replay/nn/ffn.py:43-57 ;
//                     replay/models/nn/sequential/sasrec/model.py:435-441,496-506
RP_API int rp_post_attn_train(const void* o, const void* q_in, const void* wo, const float* bo, const float* ln_w,
                              const float* ln_b, float eps, const void* w1, const float* b1, const void* w2, const float* b2,
                              const uint8_t* rowmask, int T, int d, float drop_p, unsigned long long seed,
                              unsigned long long drop_off1, unsigned long long drop_off2, const unsigned long long* seed_ptr,
                              void* h_save, void* y_save, void* u_save, float* mean_out, float* rstd_out, void* out,
                              int hd_valid, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!o || !q_in || !wo || !bo || !ln_w || !ln_b || !w1 || !b1 || !w2 || !b2 || !out || T <= 0) return RP_EINVAL;
  if (!h_save || !y_save || !u_save || !mean_out || !rstd_out) return RP_EINVAL;
  if (d != 64 && d != 128) return RP_ESHAPE;
  if (drop_p < 0.f || drop_p >= 1.f || (drop_off1 & 3) || (drop_off2 & 3)) return RP_EINVAL;
  if (hd_valid < 0 || hd_valid > 128 || (hd_valid > 0 && d % (hd_valid <= 64 ? 64 : 128))) return RP_ESHAPE;
  if (out == o || out == q_in) return RP_EINVAL;
  CUtensorMap tmO, tmWo, tmW1, tmW2;
  int rc;
  if ((rc = make_tmap_bf16(&tmO, o, T, d, d, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmWo, wo, d, d, d, d)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmW1, w1, d, d, d, d)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmW2, w2, d, d, d, d)) != RP_OK) return rc;
  PostAttnParams p;
  p.bo = bo; p.ln_w = ln_w; p.ln_b = ln_b; p.b1 = b1; p.b2 = b2;
  p.q_in = reinterpret_cast<const __nv_bfloat16*>(q_in); p.rowmask = rowmask;
  p.out = reinterpret_cast<__nv_bfloat16*>(out); p.eps = eps; p.T = T;
  p.h_save = reinterpret_cast<__nv_bfloat16*>(h_save); p.y_save = reinterpret_cast<__nv_bfloat16*>(y_save);
  p.u_save = reinterpret_cast<__nv_bfloat16*>(u_save); p.mean_out = mean_out; p.rstd_out = rstd_out;
  p.drop_p = drop_p; p.seed = seed; p.off1 = drop_off1; p.off2 = drop_off2; p.seed_ptr = seed_ptr; p.hd_valid = hd_valid;
  return d == 64 ? launch_post_attn<1, true>(tmO, tmWo, tmW1, tmW2, p, stream)
                 : launch_post_attn<2, true>(tmO, tmWo, tmW1, tmW2, p, stream);
}
