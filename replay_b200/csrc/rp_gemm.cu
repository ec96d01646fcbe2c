// rp_gemm.cu - generic batched bf16 GEMM on tcgen05 with a fused epilogue; the workhorse of the transformer body.
//
//   C[m, n] = epilogue( alpha * sum_k A(m, k) * B(n, k) )            (per batch element)
//
// Operands may be K-major (stored [rows = M/N, cols = K]) or MN-major (stored [rows = K, cols = M/N]); both are fed to
// the tensor core straight from TMA-written 128B-swizzled shared memory (no transposes in HBM).  Replaces, on the body:
//   torch.nn.MultiheadAttention in/out projections, Conv1d(d,d,1)/Linear FFN layers and their autograd backward
//   (replay/nn/sequential/sasrec/transformer.py:36-46,99-106 ; replay/nn/ffn.py:43-57 ;
//    replay/models/nn/sequential/sasrec/model.py:407-414,490-506 ; replay/models/nn/sequential/bert4rec/model.py:471-527).
//
// CTA = one 128 x BN output tile (x one K split).  warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue.
#include <stdlib.h>

#include "rp_host.h"
#include "rp_philox.cuh"
#include "rp_sm100.cuh"

namespace rp {

struct GemmParams {
  int M, N, K;                 // per-batch problem size
  int inner;                   // batch index bz = outer * inner + in
  int a_r0, a_ro, a_ri;        // A row offset = a_r0 + outer*a_ro + in*a_ri   (rows of the stored 2-D array)
  int a_c0, a_co, a_ci;        // A col offset
  int b_r0, b_ro, b_ri, b_c0, b_co, b_ci;
  void* C;                     // output
  long long ldc, c_off0, c_oo, c_oi;   // element offsets: c_off0 + outer*c_oo + in*c_oi, row pitch ldc
  int out_mode;                // 0: bf16 store, 1: fp32 atomic add, 2: fp32 store, 3: fp32 store of the split-K partial at
                               //    C + ksplit * c_split_stride (reduced afterwards by rp_reduce_splits; no atomics)
  long long c_split_stride;
  float alpha;
  const float* bias;           // [N] or null
  int act;                     // 0 none, 1 relu, 2 gelu(erf)
  const __nv_bfloat16* residual;  // same geometry as C (bf16) or null
  const uint8_t* rowmask;      // [>= rows] multiply row m by rowmask[row_index] (row_index = c row) or null
  long long rowmask_off0, rowmask_oo;   // row index base per batch: rowmask_off0 + outer*rowmask_oo
  float drop_p;                // dropout prob applied after act, before residual (0 = off)
  unsigned long long seed, drop_offset;
  const unsigned long long* seed_ptr;  // optional device counter added to seed (CUDA-graph replays get fresh masks)
  int split_k;                 // number of K splits (out_mode 1 only)
  const __nv_bfloat16* gate;   // same geometry as C: x *= (gate != 0) ? gate_scale : 0   (ReLU+dropout backward) or null
  float gate_scale;
  __nv_bfloat16* C2;           // optional second output (bf16, geometry of C): the value after bias, before the activation
  int gate_mode;               // 0: gate != 0 ? gate_scale : 0 ;  1: gelu'(gate) * gate_scale (gate = saved pre-activation)
  float post_drop_p;           // second dropout applied AFTER the residual add (BERT4Rec block output), 0 = off
  unsigned long long post_drop_offset;
  const float* row_exp2_offset;  // act 3: x = exp2(x * log2(e) + row_exp2_offset[m])   (softmax numerators from stored lse)
  const int32_t* m_limit_dev;    // optional device scalar: rows m with m_limit_base + m >= *m_limit_dev are not computed
  int m_limit_base;
  const int32_t* k_limit_dev;    // optional device scalar: the contraction stops at *k_limit_dev - k_limit_base (whole 64-chunks)
  int k_limit_base;
};

// Epilogue of one [1 row x 32 columns] strip held in registers (shared by the tile kernel and the persistent kernel).
struct EpiRow {
  long long c_base;   // element offset of this output row in C
  long long drop_row; // row index of the activation dropout stream (rp_philox.cuh): element (drop_row, column)
  float rm;           // row-mask factor
  float exp_off;      // act 3: per-row exponent offset
  float keep_scale;
  uint32_t drop_thr;
  unsigned long long seed_eff;
};

__device__ __forceinline__ void gemm_epilogue_chunk(const GemmParams& p, const float* __restrict__ s_bias, const EpiRow& er,
                                                    const uint32_t (&raw)[32], int n0, int c) {
  const long long c_base = er.c_base;
  const float rm = er.rm, keep_scale = er.keep_scale;
  const uint32_t drop_thr = er.drop_thr;
  const unsigned long long seed_eff = er.seed_eff;
  float x[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) x[q] = __uint_as_float(raw[q]) * p.alpha;
      if (p.bias) {
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[c + q]);
          x[q] += b4.x; x[q + 1] += b4.y; x[q + 2] += b4.z; x[q + 3] += b4.w;
        }
      }
      const bool full = (n0 + c + 32 <= p.N);
      if (p.C2) {
        __nv_bfloat16* o2 = p.C2 + c_base + n0 + c;
#pragma unroll
        for (int q = 0; q < 32; q += 2)
          if (n0 + c + q < p.N) *reinterpret_cast<uint32_t*>(o2 + q) = pack_bf16(x[q], x[q + 1]);
      }
      if (p.act == 1) {
#pragma unroll
        for (int q = 0; q < 32; ++q) x[q] = fmaxf(x[q], 0.f);
      } else if (p.act == 2) {
#pragma unroll
        for (int q = 0; q < 32; ++q) x[q] = 0.5f * x[q] * (1.f + erff(x[q] * 0.70710678118654752f));
      } else if (p.act == 3) {
#pragma unroll
        for (int q = 0; q < 32; ++q) x[q] = ex2f(fmaf(x[q], 1.4426950408889634f, er.exp_off));
      }
      if (p.drop_p > 0.f) {
        const uint32_t rk = drop_row_key(seed_eff, p.drop_offset, (unsigned long long)er.drop_row);
#pragma unroll
        for (int q = 0; q < 32; ++q)
          x[q] = drop_mix(rk, drop_col_key((uint32_t)(n0 + c + q))) >= drop_thr ? x[q] * keep_scale : 0.f;
      }
      if (p.gate) {
        const __nv_bfloat16* gp = p.gate + c_base + n0 + c;
        if (p.gate_mode == 0) {
#pragma unroll
          for (int q = 0; q < 32; ++q)
            if (n0 + c + q < p.N) x[q] = (__bfloat162float(gp[q]) != 0.f) ? x[q] * p.gate_scale : 0.f;
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q)
            if (n0 + c + q < p.N) {
              const float z = __bfloat162float(gp[q]);
              const float dg = 0.5f * (1.f + erff(z * 0.70710678118654752f)) + z * 0.3989422804014327f * __expf(-0.5f * z * z);
              x[q] *= dg * p.gate_scale;
            }
        }
      }
      if (p.residual) {
        const __nv_bfloat16* rp_ = p.residual + c_base + n0 + c;
        if (full) {
#pragma unroll
          for (int q = 0; q < 32; q += 8) {
            const uint4 rv = *reinterpret_cast<const uint4*>(rp_ + q);
            const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 f = __bfloat1622float2(r2[t]);
              x[q + 2 * t] += f.x;
              x[q + 2 * t + 1] += f.y;
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q)
            if (n0 + c + q < p.N) x[q] += __bfloat162float(rp_[q]);
        }
      }
      if (p.post_drop_p > 0.f) {
        const float ks2 = 1.f / (1.f - p.post_drop_p);
        const uint32_t thr2 = (uint32_t)(p.post_drop_p * 4294967296.0);
        const uint32_t rk = drop_row_key(p.seed + (p.seed_ptr ? *p.seed_ptr : 0ull), p.post_drop_offset, (unsigned long long)er.drop_row);
#pragma unroll
        for (int q = 0; q < 32; ++q)
          x[q] = drop_mix(rk, drop_col_key((uint32_t)(n0 + c + q))) >= thr2 ? x[q] * ks2 : 0.f;
      }
      if (p.rowmask) {
#pragma unroll
        for (int q = 0; q < 32; ++q) x[q] *= rm;
      }
      if (p.out_mode == 0) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.C) + c_base + n0 + c;
        if (full) {
#pragma unroll
          for (int q = 0; q < 32; q += 8) {
            uint4 w;
            w.x = pack_bf16(x[q], x[q + 1]);
            w.y = pack_bf16(x[q + 2], x[q + 3]);
            w.z = pack_bf16(x[q + 4], x[q + 5]);
            w.w = pack_bf16(x[q + 6], x[q + 7]);
            *reinterpret_cast<uint4*>(o + q) = w;
          }
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q)
            if (n0 + c + q < p.N) o[q] = __float2bfloat16(x[q]);
        }
      } else {
        float* o = reinterpret_cast<float*>(p.C) + c_base + n0 + c;
        if (p.out_mode == 1) {
#pragma unroll
          for (int q = 0; q < 32; ++q)
            if (n0 + c + q < p.N) atomicAdd(o + q, x[q]);
        } else if (p.out_mode == 4) {  // C += x, plain read-modify-write (every element has exactly one owner: split_k == 1)
          if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int q = 0; q < 32; q += 4) {
              float4 v = *reinterpret_cast<float4*>(o + q);
              v.x += x[q]; v.y += x[q + 1]; v.z += x[q + 2]; v.w += x[q + 3];
              *reinterpret_cast<float4*>(o + q) = v;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (n0 + c + q < p.N) o[q] += x[q];
          }
        } else if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
          for (int q = 0; q < 32; q += 4) *reinterpret_cast<float4*>(o + q) = make_float4(x[q], x[q + 1], x[q + 2], x[q + 3]);
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q)
            if (n0 + c + q < p.N) o[q] = x[q];
        }
      }
}

static constexpr int kGemmThreads = 192;

template <int BN, bool A_MN, bool B_MN, int NSTAGE>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  constexpr int A_BYTES = 128 * 128;       // [128 x 64] bf16
  constexpr int B_BYTES = BN * 128;        // [BN x 64] bf16
  constexpr int STAGE = A_BYTES + B_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_full[NSTAGE], bar_empty[NSTAGE], bar_acc;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[BN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x;
  const int m_tile = blockIdx.y / p.split_k, ksplit = blockIdx.y % p.split_k;
  const int bz = blockIdx.z, outer = bz / p.inner, in = bz % p.inner;
  const int m0 = m_tile * 128, n0 = n_tile * BN;
  if (p.m_limit_dev != nullptr && m0 + p.m_limit_base >= *p.m_limit_dev) return;  // whole tile beyond the dynamic row count
  int k_eff = p.K;
  if (p.k_limit_dev != nullptr) k_eff = max(0, min(p.K, *p.k_limit_dev - p.k_limit_base));
  const int k_chunks = (k_eff + 63) / 64;
  const int kc_begin = (int)(((long long)k_chunks * ksplit) / p.split_k);
  const int kc_end = (int)(((long long)k_chunks * (ksplit + 1)) / p.split_k);
  const int a_r = p.a_r0 + outer * p.a_ro + in * p.a_ri, a_c = p.a_c0 + outer * p.a_co + in * p.a_ci;
  const int b_r = p.b_r0 + outer * p.b_ro + in * p.b_ri, b_c = p.b_c0 + outer * p.b_co + in * p.b_ci;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(&bar_acc, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, BN < 32 ? 32 : BN);
  if (p.bias != nullptr && threadIdx.x >= 64) {
    for (int i = threadIdx.x - 64; i < BN; i += 128) s_bias[i] = (n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int kc = kc_begin, it = 0; kc < kc_end; ++kc, ++it) {
        const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1;
        mbar_wait(&bar_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&bar_full[s], STAGE);
        uint8_t* sa = smem + s * STAGE;
        uint8_t* sb = sa + A_BYTES;
        if (A_MN) {  // stored [K rows, M cols]: two boxes of [64 k-rows x 64 m]
          tma_load_2d(sa, &tmA, &bar_full[s], a_c + m0, a_r + kc * 64);
          tma_load_2d(sa + 8192, &tmA, &bar_full[s], a_c + m0 + 64, a_r + kc * 64);
        } else {     // stored [M rows, K cols]: one box of [128 rows x 64 k]
          tma_load_2d(sa, &tmA, &bar_full[s], a_c + kc * 64, a_r + m0);
        }
        if (B_MN) {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c)
            tma_load_2d(sb + c * 8192, &tmB, &bar_full[s], b_c + n0 + c * 64, b_r + kc * 64);
        } else {
          tma_load_2d(sb, &tmB, &bar_full[s], b_c + kc * 64, b_r + n0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, A_MN, B_MN);
      for (int kc = kc_begin, it = 0; kc < kc_end; ++kc, ++it) {
        const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1;
        mbar_wait(&bar_full[s], ph);
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem + s * STAGE), b0 = a0 + A_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ad = A_MN ? umma_desc_sw128(a0 + ks * 2048, 8192, 1024) : umma_desc_sw128(a0 + ks * 32, 16, 1024);
          const uint64_t bd = B_MN ? umma_desc_sw128(b0 + ks * 2048, 8192, 1024) : umma_desc_sw128(b0 + ks * 32, 16, 1024);
          umma_ss(tmem, ad, bd, idesc, (it | ks) != 0);
        }
        umma_commit(&bar_empty[s]);
      }
      umma_commit(&bar_acc);
    }
  } else {
    // ------------------------------------------------ epilogue: thread = output row
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int m = m0 + row;
    mbar_wait(&bar_acc, 0);
    tc_fence_after();
    const bool row_ok = m < p.M;
    const long long c_base = p.c_off0 + (long long)outer * p.c_oo + (long long)in * p.c_oi + (long long)m * p.ldc +
                             (p.out_mode == 3 ? (long long)ksplit * p.c_split_stride : 0ll);
    float rm = 1.f;
    if (p.rowmask && row_ok) rm = p.rowmask[p.rowmask_off0 + (long long)outer * p.rowmask_oo + m] ? 1.f : 0.f;
    EpiRow er;
    er.c_base = c_base;
    er.drop_row = (long long)bz * p.M + m;
    er.rm = rm;
    er.exp_off = (p.act == 3 && row_ok) ? p.row_exp2_offset[m] : 0.f;
    er.keep_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    er.drop_thr = p.drop_p > 0.f ? (uint32_t)(p.drop_p * 4294967296.0) : 0u;
    er.seed_eff = p.seed + ((p.drop_p > 0.f && p.seed_ptr) ? *p.seed_ptr : 0ull);
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t raw[32];
      tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + c, raw);
      tmem_ld_wait();
      if (!row_ok || n0 + c >= p.N) continue;
      if (kc_begin >= kc_end) {  // empty contraction (dynamic K limit): the accumulator was never written
#pragma unroll
        for (int q = 0; q < 32; ++q) raw[q] = 0u;
      }
      gemm_epilogue_chunk(p, s_bias, er, raw, n0, c);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, BN < 32 ? 32 : BN);
}

template <int BN, bool A_MN, bool B_MN, int NSTAGE>
static int launch_gemm_n(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int batch, cudaStream_t st) {
  const int smem = NSTAGE * (128 * 128 + BN * 128) + 1024;
  auto kern = gemm_kernel<BN, A_MN, B_MN, NSTAGE>;
  RP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((p.N + BN - 1) / BN, ((p.M + 127) / 128) * p.split_k, batch);
  kern<<<grid, kGemmThreads, smem, st>>>(tmA, tmB, p);
  RP_LAUNCH_CHECK();
  return RP_OK;
}


// ------------------------------------------------------------------------------------------------------------------
// Weight-stationary persistent variant for the tall-skinny projections of the body (M = tokens, N, K <= 256):
// the CTA keeps its [BN x K] slice of the weight in shared memory, streams 128-row activation tiles through a TMA ring,
// double-buffers the accumulator in TMEM and overlaps the epilogue of tile i with the loads + MMAs of tile i+1.
// grid = (ctas_per_n_tile, n_tiles); CTA x handles M tiles x, x + gridDim.x, ...
// ------------------------------------------------------------------------------------------------------------------
static constexpr int kWsEpiWarps = 8;
static constexpr int kWsThreads = 64 + kWsEpiWarps * 32;

template <int BN, int KCH, bool B_MN, int NA>
__global__ void __launch_bounds__(kWsThreads, 1)
gemm_ws_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  constexpr int A_STAGE = KCH * 128 * 128;          // [128 x K]
  constexpr int B_BYTES = KCH * BN * 128;           // [BN x K]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;
  uint8_t* sA = smem + B_BYTES;
  __shared__ uint64_t bar_b, bar_full[NA], bar_empty[NA], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[BN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BN;
  const int m_tiles = (p.M + 127) / 128;

  if (threadIdx.x == 0) {
    mbar_init(&bar_b, 1);
    for (int i = 0; i < NA; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_tfull[i], 1);
      mbar_init(&bar_tempty[i], kWsEpiWarps);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 2 * BN);
  if (p.bias != nullptr && threadIdx.x >= 64)
    for (int i = threadIdx.x - 64; i < BN; i += kWsEpiWarps * 32) s_bias[i] = (n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bar_b, B_BYTES);
      for (int kc = 0; kc < KCH; ++kc) {
        if (B_MN) {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c)
            tma_load_2d(sB + kc * (BN * 128) + c * 8192, &tmB, &bar_b, p.b_c0 + n0 + c * 64, p.b_r0 + kc * 64);
        } else {
          tma_load_2d(sB + kc * (BN * 128), &tmB, &bar_b, p.b_c0 + kc * 64, p.b_r0 + n0);
        }
      }
      int it = 0;
      for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++it) {
        const uint32_t s = it % NA, ph = (it / NA) & 1;
        mbar_wait(&bar_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&bar_full[s], A_STAGE);
        for (int kc = 0; kc < KCH; ++kc)
          tma_load_2d(sA + s * A_STAGE + kc * 16384, &tmA, &bar_full[s], p.a_c0 + kc * 64, p.a_r0 + mt * 128);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, false, B_MN);
      mbar_wait(&bar_b, 0);
      int it = 0;
      for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++it) {
        const uint32_t s = it % NA, ph = (it / NA) & 1, as = it & 1, aph = (it >> 1) & 1;
        mbar_wait(&bar_tempty[as], aph ^ 1);
        mbar_wait(&bar_full[s], ph);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA + s * A_STAGE), b0 = smem_u32(sB);
#pragma unroll
        for (int kc = 0; kc < KCH; ++kc)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = umma_desc_sw128(a0 + kc * 16384 + ks * 32, 16, 1024);
            const uint64_t bd = B_MN ? umma_desc_sw128(b0 + kc * (BN * 128) + ks * 2048, 8192, 1024)
                                     : umma_desc_sw128(b0 + kc * (BN * 128) + ks * 32, 16, 1024);
            umma_ss(tmem + as * BN, ad, bd, idesc, (kc | ks) != 0);
          }
        umma_commit(&bar_empty[s]);
        umma_commit(&bar_tfull[as]);
      }
    }
  } else {
    // ------------------------------------------------ epilogue: 8 warps, warp%4 = lane quarter, (warp-2)/4 = column half
    const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
    const int row = quarter * 32 + lane;
    constexpr int HALF = BN / 2;
    EpiRow er;
    er.keep_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    er.drop_thr = p.drop_p > 0.f ? (uint32_t)(p.drop_p * 4294967296.0) : 0u;
    er.seed_eff = p.seed + ((p.drop_p > 0.f && p.seed_ptr) ? *p.seed_ptr : 0ull);
    int it = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++it) {
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      const int m = mt * 128 + row;
      const bool row_ok = m < p.M;
      er.c_base = p.c_off0 + (long long)m * p.ldc;
      er.drop_row = m;
      er.rm = 1.f;
      er.exp_off = 0.f;
      if (p.rowmask && row_ok) er.rm = p.rowmask[p.rowmask_off0 + m] ? 1.f : 0.f;
      mbar_wait(&bar_tfull[as], aph);
      tc_fence_after();
      const uint32_t tbase = tmem + ((uint32_t)(quarter * 32) << 16) + as * BN + half * HALF;
      if (HALF == 128) {   // BN = 256 (both halves of a fused K | V projection in one pass over the activations)
        uint32_t r0[32], r1[32];
        tmem_ld32(tbase, r0);
        tmem_ld32(tbase + 32, r1);
        tmem_ld_wait();
        if (row_ok && n0 + half * HALF < p.N) gemm_epilogue_chunk(p, s_bias, er, r0, n0, half * HALF);
        if (row_ok && n0 + half * HALF + 32 < p.N) gemm_epilogue_chunk(p, s_bias, er, r1, n0, half * HALF + 32);
        tmem_ld32(tbase + 64, r0);
        tmem_ld32(tbase + 96, r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_tempty[as]);
        if (row_ok && n0 + half * HALF + 64 < p.N) gemm_epilogue_chunk(p, s_bias, er, r0, n0, half * HALF + 64);
        if (row_ok && n0 + half * HALF + 96 < p.N) gemm_epilogue_chunk(p, s_bias, er, r1, n0, half * HALF + 96);
      } else if (HALF == 64) {
        uint32_t r0[32], r1[32];
        tmem_ld32(tbase, r0);
        tmem_ld32(tbase + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_tempty[as]);  // accumulator stage is free once its values sit in registers
        if (row_ok && n0 + half * HALF < p.N) gemm_epilogue_chunk(p, s_bias, er, r0, n0, half * HALF);
        if (row_ok && n0 + half * HALF + 32 < p.N) gemm_epilogue_chunk(p, s_bias, er, r1, n0, half * HALF + 32);
      } else {
        uint32_t r0[32];
        tmem_ld32(tbase, r0);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_tempty[as]);
        if (row_ok && n0 + half * HALF < p.N) gemm_epilogue_chunk(p, s_bias, er, r0, n0, half * HALF);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 2 * BN);
}

template <int BN, int KCH, bool B_MN>
static int launch_gemm_ws(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t st) {
  constexpr int NA = 2;  // B + 2 A stages: <= 96 KB for K <= 128 so that two CTAs share an SM
  const int smem = KCH * BN * 128 + NA * KCH * 128 * 128 + 1024;
  auto kern = gemm_ws_kernel<BN, KCH, B_MN, NA>;
  RP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int n_tiles = (p.N + BN - 1) / BN, m_tiles = (p.M + 127) / 128;
  int per_n = ((KCH <= 2 && BN <= 128) ? 2 : 1) * sm_count() / n_tiles;
  if (per_n < 1) per_n = 1;
  if (per_n > m_tiles) per_n = m_tiles;
  dim3 grid(per_n, n_tiles);
  kern<<<grid, kWsThreads, smem, st>>>(tmA, tmB, p);
  RP_LAUNCH_CHECK();
  return RP_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent streaming variant for mid-size and large single GEMMs (BERT4Rec's d -> 4d FFN, the d = 512 CE backward,
// predict-sized projections): one CTA per SM walks 128 x 128 output tiles (n fastest, so that CTAs running side by side
// share the A rows in L2); A and B k-chunks stream through an NSTAGE-deep TMA ring; FOUR accumulator stages in TMEM
// (4 x 128 = 512 columns) let the 8 epilogue warps trail the tensor core by up to three tiles, so short-K tiles with heavy
// epilogues (bias + GELU + dropout) and long-K tiles both keep the MMA pipe fed.  batch == 1, split_k == 1.
// ------------------------------------------------------------------------------------------------------------------
static constexpr int kPsEpiWarps = 8;
static constexpr int kPsThreads = 64 + kPsEpiWarps * 32;

// BN = 128: four accumulator stages; BN = 256 (wide outputs): two stages, and an M=128 x N=256 MMA reads 12 KB of shared
// memory per 128 tensor-core cycles instead of 8 KB per 64, i.e. it is no longer shared-memory-bandwidth bound.
template <int BN, bool A_MN, bool B_MN, int NSTAGE>
__global__ void __launch_bounds__(kPsThreads, 1)
gemm_ps_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  constexpr int kPsAcc = 512 / BN;
  constexpr int A_BYTES = 128 * 128, B_BYTES = BN * 128, STAGE = A_BYTES + B_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_full[NSTAGE], bar_empty[NSTAGE], bar_tfull[kPsAcc], bar_tempty[kPsAcc];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int m_eff = p.M, k_eff = p.K;
  if (p.m_limit_dev != nullptr) m_eff = max(0, min(p.M, *p.m_limit_dev - p.m_limit_base));
  if (p.k_limit_dev != nullptr) k_eff = max(0, min(p.K, *p.k_limit_dev - p.k_limit_base));
  const int m_tiles = (m_eff + 127) / 128, n_tiles = (p.N + BN - 1) / BN;
  const int k_chunks = (k_eff + 63) / 64;
  const long long total = (long long)m_tiles * n_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    for (int i = 0; i < kPsAcc; ++i) {
      mbar_init(&bar_tfull[i], 1);
      mbar_init(&bar_tempty[i], kPsEpiWarps);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, kPsAcc * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0;
      for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const int m0 = (int)(t / n_tiles) * 128, n0 = (int)(t % n_tiles) * BN;
        for (int kc = 0; kc < k_chunks; ++kc, ++it) {
          const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1;
          mbar_wait(&bar_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&bar_full[s], STAGE);
          uint8_t* sa = smem + s * STAGE;
          uint8_t* sb = sa + A_BYTES;
          if (A_MN) {
            tma_load_2d(sa, &tmA, &bar_full[s], p.a_c0 + m0, p.a_r0 + kc * 64);
            tma_load_2d(sa + 8192, &tmA, &bar_full[s], p.a_c0 + m0 + 64, p.a_r0 + kc * 64);
          } else {
            tma_load_2d(sa, &tmA, &bar_full[s], p.a_c0 + kc * 64, p.a_r0 + m0);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(sb + c * 8192, &tmB, &bar_full[s], p.b_c0 + n0 + c * 64, p.b_r0 + kc * 64);
          } else {
            tma_load_2d(sb, &tmB, &bar_full[s], p.b_c0 + kc * 64, p.b_r0 + n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, A_MN, B_MN);
      uint32_t it = 0, tile = 0;
      for (long long t = blockIdx.x; t < total; t += gridDim.x, ++tile) {
        const uint32_t as = tile % kPsAcc, aph = (tile / kPsAcc) & 1;
        mbar_wait(&bar_tempty[as], aph ^ 1);
        tc_fence_after();
        for (int kc = 0; kc < k_chunks; ++kc, ++it) {
          const uint32_t s = it % NSTAGE, ph = (it / NSTAGE) & 1;
          mbar_wait(&bar_full[s], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + s * STAGE), b0 = a0 + A_BYTES;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = A_MN ? umma_desc_sw128(a0 + ks * 2048, 8192, 1024) : umma_desc_sw128(a0 + ks * 32, 16, 1024);
            const uint64_t bd = B_MN ? umma_desc_sw128(b0 + ks * 2048, 8192, 1024) : umma_desc_sw128(b0 + ks * 32, 16, 1024);
            umma_ss(tmem + as * BN, ad, bd, idesc, (kc | ks) != 0);
          }
          umma_commit(&bar_empty[s]);
        }
        umma_commit(&bar_tfull[as]);
      }
    }
  } else {
    // ------------------------------------------------ epilogue: 8 warps, warp%4 = lane quarter, (warp-2)/4 = column half
    const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
    const int row = quarter * 32 + lane;
    constexpr int HALF = BN / 2;
    EpiRow er;
    er.keep_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    er.drop_thr = p.drop_p > 0.f ? (uint32_t)(p.drop_p * 4294967296.0) : 0u;
    er.seed_eff = p.seed + ((p.drop_p > 0.f && p.seed_ptr) ? *p.seed_ptr : 0ull);
    uint32_t tile = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x, ++tile) {
      const uint32_t as = tile % kPsAcc, aph = (tile / kPsAcc) & 1;
      const int m = (int)(t / n_tiles) * 128 + row, n0 = (int)(t % n_tiles) * BN;
      const bool row_ok = m < p.M;
      er.c_base = p.c_off0 + (long long)m * p.ldc;
      er.drop_row = m;
      er.rm = 1.f;
      if (p.rowmask && row_ok) er.rm = p.rowmask[p.rowmask_off0 + m] ? 1.f : 0.f;
      er.exp_off = (p.act == 3 && row_ok) ? p.row_exp2_offset[m] : 0.f;
      mbar_wait(&bar_tfull[as], aph);
      tc_fence_after();
      const uint32_t tbase = tmem + ((uint32_t)(quarter * 32) << 16) + as * BN + half * HALF;
      const float* bias_n0 = p.bias ? p.bias + n0 : nullptr;  // N % 32 == 0 is required with a bias (launcher)
#pragma unroll 1
      for (int c0 = 0; c0 < HALF; c0 += 64) {  // 64 columns per round
        uint32_t r0[32], r1[32];
        tmem_ld32(tbase + c0, r0);
        tmem_ld32(tbase + c0 + 32, r1);
        tmem_ld_wait();
        if (c0 + 64 == HALF) {  // accumulator stage is free once its last values sit in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_tempty[as]);
        }
        if (k_chunks == 0) {  // empty contraction (dynamic K limit): the accumulator was never written
#pragma unroll
          for (int q = 0; q < 32; ++q) r0[q] = r1[q] = 0u;
        }
        const int c = half * HALF + c0;
        if (row_ok && n0 + c < p.N) gemm_epilogue_chunk(p, bias_n0, er, r0, n0, c);
        if (row_ok && n0 + c + 32 < p.N) gemm_epilogue_chunk(p, bias_n0, er, r1, n0, c + 32);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kPsAcc * BN);
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm_ps(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t st) {
  constexpr int NSTAGE = BN == 128 ? 6 : 4;
  const int smem = NSTAGE * (128 * 128 + BN * 128) + 1024;
  auto kern = gemm_ps_kernel<BN, A_MN, B_MN, NSTAGE>;
  RP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long long tiles = (long long)((p.M + 127) / 128) * ((p.N + BN - 1) / BN);
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  kern<<<grid, kPsThreads, smem, st>>>(tmA, tmB, p);
  RP_LAUNCH_CHECK();
  return RP_OK;
}

// short-K problems (the body's d x d projections) use 2 stages so that 3 CTAs share an SM and their prologues,
// main loops and epilogues overlap; long-K problems (weight gradients) use a 4-deep ring.
template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int batch, cudaStream_t st) {
  const int chunks_per_cta = ((p.K + 63) / 64 + p.split_k - 1) / p.split_k;
  if (chunks_per_cta <= 2) return launch_gemm_n<BN, A_MN, B_MN, 2>(tmA, tmB, p, batch, st);
  return launch_gemm_n<BN, A_MN, B_MN, 4>(tmA, tmB, p, batch, st);
}

}  // namespace rp

using namespace rp;

#include "rp_gemm_desc.h"

RP_API int rp_gemm(const rp_gemm_desc* g, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!g || !g->A || !g->B || !g->C) return RP_EINVAL;
  if (g->M <= 0 || g->N <= 0 || g->K <= 0 || g->batch <= 0 || g->inner <= 0) return RP_ESHAPE;
  if (g->split_k < 1 || (g->split_k > 1 && g->out_mode != 1 && g->out_mode != 3)) return RP_EINVAL;
  if (g->out_mode == 0 && (g->ldc % 8 != 0)) return RP_EALIGN;
  GemmParams p;
  p.M = g->M; p.N = g->N; p.K = g->K; p.inner = g->inner;
  p.a_r0 = g->a_r0; p.a_ro = g->a_ro; p.a_ri = g->a_ri; p.a_c0 = g->a_c0; p.a_co = g->a_co; p.a_ci = g->a_ci;
  p.b_r0 = g->b_r0; p.b_ro = g->b_ro; p.b_ri = g->b_ri; p.b_c0 = g->b_c0; p.b_co = g->b_co; p.b_ci = g->b_ci;
  p.C = g->C; p.ldc = g->ldc; p.c_off0 = g->c_off0; p.c_oo = g->c_oo; p.c_oi = g->c_oi; p.out_mode = g->out_mode;
  p.alpha = g->alpha; p.bias = g->bias; p.act = g->act;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(g->residual);
  p.rowmask = g->rowmask; p.rowmask_off0 = g->rowmask_off0; p.rowmask_oo = g->rowmask_oo;
  p.drop_p = g->drop_p; p.seed = g->seed; p.drop_offset = g->drop_offset; p.seed_ptr = g->seed_ptr; p.split_k = g->split_k;
  p.gate = reinterpret_cast<const __nv_bfloat16*>(g->gate); p.gate_scale = g->gate_scale;
  p.C2 = reinterpret_cast<__nv_bfloat16*>(g->C2); p.gate_mode = g->gate_mode; p.post_drop_p = g->post_drop_p;
  p.post_drop_offset = g->post_drop_offset;
  p.c_split_stride = g->c_split_stride;
  p.row_exp2_offset = g->row_exp2_offset;
  p.m_limit_dev = g->m_limit_dev; p.m_limit_base = g->m_limit_base;
  p.k_limit_dev = g->k_limit_dev; p.k_limit_base = g->k_limit_base;
  if (g->act == 3 && !g->row_exp2_offset) return RP_EINVAL;
  if (g->out_mode == 4 && g->split_k != 1) return RP_EINVAL;
  CUtensorMap tmA, tmB;
  int rc;
  // K-major operand: box [128 (or BN) rows x 64 cols]; MN-major operand: box [64 k-rows x 64 cols]
  const int bn = (g->N <= 64) ? 64 : 128;
  if ((rc = make_tmap_bf16(&tmA, g->A, g->a_rows, g->a_cols, g->lda, g->a_mn ? 64 : 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmB, g->B, g->b_rows, g->b_cols, g->ldb, g->b_mn ? 64 : bn)) != RP_OK) return rc;
  // weight-stationary persistent kernel: activation [M, K] K-major, K in {64,128,256}, one batch, no split-K
  static const int ws_min_m = getenv("RP_GEMM_WS_MIN_M") ? atoi(getenv("RP_GEMM_WS_MIN_M")) : 131072;  // measured: pays for M >~ 100K rows (predict), neutral at 51K
  const bool ws_ok = g->act != 3 && g->out_mode != 4 && !g->m_limit_dev && !g->k_limit_dev && g->M >= ws_min_m && !g->a_mn && g->batch == 1 && g->split_k == 1 && g->M >= 1024 && g->N >= 64 &&
                     (g->K == 64 || g->K == 128 || g->K == 256) && g->a_ro == 0 && g->a_ri == 0 && g->b_ro == 0 && g->b_ri == 0 &&
                     g->a_co == 0 && g->a_ci == 0 && g->b_co == 0 && g->b_ci == 0 && g->c_oo == 0 && g->c_oi == 0;
  if (ws_ok && bn == 128 && g->N == 256 && !g->b_mn && g->K <= 128 && !getenv("RP_GEMM_WS_NO_WIDE")) {
    // both 128-column halves in one CTA: the activations are read once (two n-tiles each streamed the whole [M, K] matrix:
    // 421 MB of DRAM reads for a 210 MB input in the predict body's K | V projection, ncu r2j)
    if ((rc = make_tmap_bf16(&tmB, g->B, g->b_rows, g->b_cols, g->ldb, 256)) != RP_OK) return rc;
    return g->K == 64 ? launch_gemm_ws<256, 1, false>(tmA, tmB, p, stream) : launch_gemm_ws<256, 2, false>(tmA, tmB, p, stream);
  }
  if (ws_ok && bn == 128) {
#define RP_WS_CASE(KCH_)                                                             \
  return g->b_mn ? launch_gemm_ws<128, KCH_, true>(tmA, tmB, p, stream) : launch_gemm_ws<128, KCH_, false>(tmA, tmB, p, stream)
    if (g->K == 64) { RP_WS_CASE(1); }
    if (g->K == 128) { RP_WS_CASE(2); }
    RP_WS_CASE(4);
#undef RP_WS_CASE
  }
  // persistent streaming kernel: single GEMMs with enough tiles for every SM and a contraction / width beyond the body's
  // d x d projections (those are launch/latency-bound and stay on the tile kernel, which co-schedules 3 CTAs per SM)
  static const long long ps_min_flop = getenv("RP_GEMM_PS_MIN_FLOP") ? atoll(getenv("RP_GEMM_PS_MIN_FLOP")) : 4000000000ll;
  static const bool ps_small = getenv("RP_GEMM_PS_SMALL") && atoi(getenv("RP_GEMM_PS_SMALL")) != 0;  // experiment: d x d projections too
  const long long tiles = (long long)((g->M + 127) / 128) * ((g->N + 127) / 128);
  const bool ps_ok = g->batch == 1 && g->split_k == 1 && g->out_mode != 1 && g->out_mode != 3 && bn == 128 &&
                     tiles >= sm_count() && 2ll * g->M * g->N * g->K >= ps_min_flop && (g->K > 128 || g->N > 128 || ps_small) &&
                     (!g->bias || g->N % 32 == 0) && g->a_ro == 0 && g->a_ri == 0 && g->b_ro == 0 && g->b_ri == 0 &&
                     g->a_co == 0 && g->a_ci == 0 && g->b_co == 0 && g->b_ci == 0 && g->c_oo == 0 && g->c_oi == 0 &&
                     g->rowmask_oo == 0;
  if (ps_ok) {
    // wide outputs: 128 x 256 tiles when they still give every SM a tile
    const bool wide = g->N >= 256 && (long long)((g->M + 127) / 128) * ((g->N + 255) / 256) >= sm_count();
    if (wide && !g->b_mn) {  // K-major B: the TMA box grows to 256 rows
      if ((rc = make_tmap_bf16(&tmB, g->B, g->b_rows, g->b_cols, g->ldb, 256)) != RP_OK) return rc;
    }
#define RP_PS_CASE(BN_)                                                                                  \
  do {                                                                                                   \
    if (!g->a_mn && !g->b_mn) return launch_gemm_ps<BN_, false, false>(tmA, tmB, p, stream);             \
    if (!g->a_mn && g->b_mn) return launch_gemm_ps<BN_, false, true>(tmA, tmB, p, stream);               \
    if (g->a_mn && !g->b_mn) return launch_gemm_ps<BN_, true, false>(tmA, tmB, p, stream);               \
    return launch_gemm_ps<BN_, true, true>(tmA, tmB, p, stream);                                         \
  } while (0)
    if (wide) RP_PS_CASE(256);
    RP_PS_CASE(128);
#undef RP_PS_CASE
  }
#define RP_GEMM_CASE(BN_, AMN_, BMN_) return launch_gemm<BN_, AMN_, BMN_>(tmA, tmB, p, g->batch, stream)
  if (bn == 64) {
    if (!g->a_mn && !g->b_mn) RP_GEMM_CASE(64, false, false);
    if (!g->a_mn && g->b_mn) RP_GEMM_CASE(64, false, true);
    if (g->a_mn && !g->b_mn) RP_GEMM_CASE(64, true, false);
    RP_GEMM_CASE(64, true, true);
  } else {
    if (!g->a_mn && !g->b_mn) RP_GEMM_CASE(128, false, false);
    if (!g->a_mn && g->b_mn) RP_GEMM_CASE(128, false, true);
    if (g->a_mn && !g->b_mn) RP_GEMM_CASE(128, true, false);
    RP_GEMM_CASE(128, true, true);
  }
#undef RP_GEMM_CASE
}

// dst[i] (+)= sum_s src[s * stride + i]   - second stage of the split-K weight-gradient GEMMs (out_mode 3).
// The sums are short (n = d*d elements) but deep (~100 splits): a block covers 32 float4 columns x 8 split groups so that
// ~100 independent 16-byte loads per column are in flight instead of one serial chain; fixed summation order (deterministic).
__global__ void __launch_bounds__(256) reduce_splits_kernel(const float* __restrict__ src, int n_splits, long long stride,
                                                            long long n, float* __restrict__ dst, int accumulate) {
  __shared__ float4 part[8][32];
  const int col = threadIdx.x & 31, sg = threadIdx.x >> 5;
  for (long long i0 = (long long)blockIdx.x * 128; i0 < n; i0 += (long long)gridDim.x * 128) {
    const long long i = i0 + col * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) {
      for (int s2 = sg; s2 < n_splits; s2 += 8) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + (long long)s2 * stride + i));
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
    }
    part[sg][col] = a;
    __syncthreads();
    if (sg == 0 && i < n) {
      float4 t = accumulate ? *reinterpret_cast<const float4*>(dst + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 v = part[g][col];
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      *reinterpret_cast<float4*>(dst + i) = t;
    }
    __syncthreads();
  }
}

RP_API int rp_reduce_splits(const float* src, int n_splits, long long stride, long long n, float* dst, int accumulate,
                            void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!src || !dst || n_splits <= 0 || n <= 0 || (n & 3) || (stride & 3)) return RP_EINVAL;
  long long blocks = (n + 127) / 128;
  if (blocks > rp::sm_count() * 8) blocks = rp::sm_count() * 8;
  reduce_splits_kernel<<<(int)blocks, 256, 0, stream>>>(src, n_splits, stride, n, dst, accumulate);
  RP_LAUNCH_CHECK();
  return RP_OK;
}
