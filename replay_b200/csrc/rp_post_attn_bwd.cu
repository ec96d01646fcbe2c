// rp_post_attn_bwd.cu - backward of everything AFTER the attention of one SASRec block, one pass over the tokens:
//
//   forward (rp_post_attn_train):  h = O Wo^T + bo + q_in ; y = LN2(h) ; u = drop1(relu(y W1^T + b1)) ; x' = (y + drop2(u W2^T + b2)) [* pad]
//   here, given dz = d loss / d x':
//     dzm = dz [* pad] ;  d_t = drop2'(dzm) ;  du = (d_t W2) * relu'/drop1'(u) ;  dy = du W1 + dzm ;
//     dh  = LN2-backward(dy ; h, mean, rstd, w) ;  d_o = dh Wo          (+ dLN2.weight, dLN2.bias)
//
// Replaces autograd's backward of  replay/nn/sequential/sasrec/transformer.py:107-110 + replay/nn/ffn.py:43-57
// (legacy: replay/models/nn/sequential/sasrec/model.py:436-441,496-506), which round 1 ran as dropout-backward + three GEMMs +
// LayerNorm-backward (5 launches, 15 [T, d] passes).  Here dz, u and h are read once; d_t, du, dh (operands of the grouped
// weight-gradient launch, dh also the residual gradient into the pre-attention part) and d_o are written once.
// Three chained tcgen05 GEMMs per 128-token tile whose A operands (d_t, du, dh) never leave the SM: the epilogue warps write
// them into TMEM as packed bf16 over accumulators that are no longer needed (two 128-column regions per tile, two tiles in
// flight); the weights are read MN-major in place (contraction over their output features) and stay resident in shared memory.
#include "rp_host.h"
#include "rp_philox.cuh"
#include "rp_sm100.cuh"

namespace rp {

static constexpr int kPbEpiWarps = 8;
static constexpr int kPbThreads = 64 + kPbEpiWarps * 32;

struct PostAttnBwdParams {
  const __nv_bfloat16* u;      // [T, d] saved FFN hidden activation AFTER its dropout (zero = ReLU-clipped or dropped)
  const __nv_bfloat16* h;      // [T, d] saved LayerNorm2 input
  const float* mean;
  const float* rstd;
  const float* ln_w;
  const uint8_t* rowmask;      // legacy: the block output was multiplied by the pad mask (or null)
  __nv_bfloat16* d_t;          // [T, d] or null (then d_t == dz: no dropout, no row mask)
  __nv_bfloat16* du;
  __nv_bfloat16* dh;
  __nv_bfloat16* d_o;
  float* dln_w;
  float* dln_b;
  float drop_p;
  unsigned long long seed, off2;
  const unsigned long long* seed_ptr;
  int T;
  int hd_valid;                // > 0: padded feature slots - LN statistics over the real features, no gradient into padded inputs
};

// column sums over the 32 rows of a warp: on return lane l holds the sums of columns 2l and 2l+1 in v[0], v[1]
__device__ __forceinline__ void warp_colsum64_pb(float (&v)[64], int lane) {
#pragma unroll
  for (int w = 32, bit = 16; w >= 2; w >>= 1, bit >>= 1) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i < w) {
        const bool up = lane & bit;
        const float send = up ? v[i] : v[i + w], keep = up ? v[i + w] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
      }
    }
  }
}

// output tensor maps (box [128 rows x 64 columns]): d_t (valid only when p.d_t != null), du, dh, d_o
struct PostAttnBwdOutMaps {
  CUtensorMap d_t, du, dh, d_o;
};

template <int KCH, int NA>
__global__ void __launch_bounds__(kPbThreads, 1)
post_attn_bwd_kernel(const __grid_constant__ CUtensorMap tmDZ, const __grid_constant__ CUtensorMap tmW2,
                     const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmWo,
                     const __grid_constant__ PostAttnBwdOutMaps om, const PostAttnBwdParams p) {
  constexpr int D = KCH * 64;
  constexpr int W_BYTES = KCH * KCH * 8192;   // MN-major B: K chunks (64 output features) x N chunks (64 input features)
  constexpr int Z_STAGE = KCH * 128 * 128;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW2 = smem;
  uint8_t* sW1 = smem + W_BYTES;
  uint8_t* sWo = smem + 2 * W_BYTES;
  uint8_t* sZ = smem + 3 * W_BYTES;
  uint8_t* sOut = sZ + NA * Z_STAGE;   // one [128 x 64] bf16 staging tile per column half: outputs leave through TMA stores
  __shared__ uint64_t bar_w, z_full[NA], z_empty[NA], a_ready[3][2], g_full[3][2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_lnw[D];
  __shared__ __align__(16) uint32_t s_ck[D];      // dropout column keys (rp_philox.cuh)
  __shared__ float2 s_stat[2][128];
  __shared__ float s_red[2][kPbEpiWarps][64];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.T + 127) / 128;
  const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int n_pairs = (my_tiles + 1) / 2;
  if (threadIdx.x == 0) {
    mbar_init(&bar_w, 1);
    for (int i = 0; i < NA; ++i) {
      mbar_init(&z_full[i], 1);
      mbar_init(&z_empty[i], kPbEpiWarps);
    }
    for (int k = 0; k < 3; ++k)
      for (int i = 0; i < 2; ++i) {
        mbar_init(&a_ready[k][i], kPbEpiWarps);
        mbar_init(&g_full[k][i], 1);
      }
    fence_barrier_init();
    tma_prefetch_desc(&tmDZ);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmWo);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  if (threadIdx.x >= 64)
    for (int i = threadIdx.x - 64; i < D; i += kPbEpiWarps * 32) {
      s_lnw[i] = p.ln_w[i];
      s_ck[i] = drop_col_key((uint32_t)i);
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // TMEM per parity pp (256 columns): R0 = pp*256: d_t (packed bf16) -> accumulator 2 (du W1) -> dh (packed);
  //                                   R1 = R0 + 128: accumulator 1 (d_t W2) -> du (packed) -> accumulator 3 (dh Wo)

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bar_w, 3 * W_BYTES);
      for (int kc = 0; kc < KCH; ++kc)
        for (int nc = 0; nc < KCH; ++nc) {
          tma_load_2d(sW2 + (kc * KCH + nc) * 8192, &tmW2, &bar_w, nc * 64, kc * 64);
          tma_load_2d(sW1 + (kc * KCH + nc) * 8192, &tmW1, &bar_w, nc * 64, kc * 64);
          tma_load_2d(sWo + (kc * KCH + nc) * 8192, &tmWo, &bar_w, nc * 64, kc * 64);
        }
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const uint32_t s = it % NA, ph = (it / NA) & 1;
        mbar_wait(&z_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&z_full[s], Z_STAGE);
        for (int kc = 0; kc < KCH; ++kc) tma_load_2d(sZ + s * Z_STAGE + kc * 16384, &tmDZ, &z_full[s], kc * 64, t * 128);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, D, false, true);
      mbar_wait(&bar_w, 0);
      tc_fence_after();
      for (int pi = 0; pi < n_pairs; ++pi) {
        const uint32_t pph = pi & 1;
#pragma unroll 1
        for (int stage = 0; stage < 3; ++stage) {
#pragma unroll 1
          for (int pp = 0; pp < 2; ++pp) {
            if (2 * pi + pp >= my_tiles) continue;
            const uint32_t R0 = tmem + pp * 256, R1 = R0 + 128;
            mbar_wait(&a_ready[stage][pp], pph);
            tc_fence_after();
            const uint32_t b0 = smem_u32(stage == 0 ? sW2 : (stage == 1 ? sW1 : sWo));
            const uint32_t a_t = (stage == 1) ? R1 : R0, d_t = (stage == 1) ? R0 : R1;
#pragma unroll
            for (int kc = 0; kc < KCH; ++kc)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_ts(d_t, a_t + kc * 64 + ks * 8, umma_desc_sw128(b0 + kc * (KCH * 8192) + ks * 2048, 8192, 1024), idesc,
                        (kc | ks) != 0);
            umma_commit(&g_full[stage][pp]);
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
    const float ks_ = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    const uint32_t thr = p.drop_p > 0.f ? (uint32_t)(p.drop_p * 4294967296.0) : 0u;
    const unsigned long long seed_eff = p.seed + ((p.drop_p > 0.f && p.seed_ptr) ? *p.seed_ptr : 0ull);
    float acc_w0 = 0.f, acc_w1 = 0.f, acc_b0 = 0.f, acc_b1 = 0.f;
    // row pieces -> swizzled staging tile -> one TMA store per [128 x 64] tile (a thread owns a ROW: direct global stores are
    // 32 different lines per warp instruction, the L1 LSU wavefront limit of profiles/r2c_body_ncu.md); rows >= T are clipped
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
      for (int stage = 0; stage < 4; ++stage) {
#pragma unroll 1
        for (int pp = 0; pp < 2; ++pp) {
          const int it = 2 * pi + pp;
          if (it >= my_tiles) continue;
          const uint32_t R0 = tmem + lane_base + pp * 256, R1 = R0 + 128;
          const uint32_t s = it % NA, zph = (it / NA) & 1;
          const int t = (int)blockIdx.x + it * (int)gridDim.x;
          const int m = t * 128 + row;
          const bool row_ok = m < p.T;
          const float rm = (p.rowmask == nullptr || (row_ok && p.rowmask[m])) ? 1.f : 0.f;
          const uint8_t* ztile = sZ + s * Z_STAGE + half * 16384;
          if (stage == 0) {
            // ---- d_t = dropout2'(dz * pad) -> TMEM (A operand of GEMM 1) and HBM (operand of the W2 / b2 gradients)
            mbar_wait(&z_full[s], zph);
            if (has_half) {
              uint32_t pk[32];
              const uint32_t rk2 = p.drop_p > 0.f ? drop_row_key(seed_eff, p.off2, (unsigned long long)m) : 0u;
#pragma unroll
              for (int c8 = 0; c8 < 8; ++c8) {
                const uint4 zv = *reinterpret_cast<const uint4*>(ztile + sw128_off((uint32_t)row, (uint32_t)c8));
                const __nv_bfloat162* z2 = reinterpret_cast<const __nv_bfloat162*>(&zv);
                float v[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(z2[e]);
                  v[2 * e] = f.x * rm;
                  v[2 * e + 1] = f.y * rm;
                }
                if (p.drop_p > 0.f) {
                  const uint4 ka = *reinterpret_cast<const uint4*>(s_ck + c0 + c8 * 8), kb = *reinterpret_cast<const uint4*>(s_ck + c0 + c8 * 8 + 4);
                  v[0] = drop_mix(rk2, ka.x) >= thr ? v[0] * ks_ : 0.f; v[1] = drop_mix(rk2, ka.y) >= thr ? v[1] * ks_ : 0.f;
                  v[2] = drop_mix(rk2, ka.z) >= thr ? v[2] * ks_ : 0.f; v[3] = drop_mix(rk2, ka.w) >= thr ? v[3] * ks_ : 0.f;
                  v[4] = drop_mix(rk2, kb.x) >= thr ? v[4] * ks_ : 0.f; v[5] = drop_mix(rk2, kb.y) >= thr ? v[5] * ks_ : 0.f;
                  v[6] = drop_mix(rk2, kb.z) >= thr ? v[6] * ks_ : 0.f; v[7] = drop_mix(rk2, kb.w) >= thr ? v[7] * ks_ : 0.f;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) pk[c8 * 4 + e] = pack_bf16(v[2 * e], v[2 * e + 1]);
              }
              tmem_st16(R0 + c0, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
              tmem_st16(R0 + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&pk[16]));
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&a_ready[0][pp]);
              if (p.d_t != nullptr) stage_store(pk, &om.d_t, t * 128);
            } else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&a_ready[0][pp]);
            }
          } else if (stage == 1) {
            // ---- du = (d_t W2) * [u != 0] / keep  -> TMEM (in place over this warp's own accumulator columns) and HBM
            uint32_t uk[32];
            if (has_half) {
              const uint4* ur = reinterpret_cast<const uint4*>(p.u + (size_t)(row_ok ? m : 0) * D + c0);
#pragma unroll
              for (int c8 = 0; c8 < 8; ++c8) {
                const uint4 uv = row_ok ? __ldg(ur + c8) : make_uint4(0u, 0u, 0u, 0u);
                uk[c8 * 4] = uv.x; uk[c8 * 4 + 1] = uv.y; uk[c8 * 4 + 2] = uv.z; uk[c8 * 4 + 3] = uv.w;
              }
            }
            mbar_wait(&g_full[0][pp], pph);
            tc_fence_after();
            if (has_half) {
              uint32_t r0[32], r1[32], pk[32];
              tmem_ld32(R1 + c0, r0);
              tmem_ld32(R1 + c0 + 32, r1);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 64; q += 2) {
                const uint32_t w = uk[q >> 1];
                const float a = __uint_as_float(q < 32 ? r0[q] : r1[q - 32]), b = __uint_as_float(q + 1 < 32 ? r0[q + 1] : r1[q + 1 - 32]);
                // bf16 zero test on the raw halves (-0 cannot occur after ReLU)
                pk[q >> 1] = pack_bf16((w & 0x7fffu) ? a * ks_ : 0.f, (w & 0x7fff0000u) ? b * ks_ : 0.f);
              }
              tmem_st16(R1 + c0, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
              tmem_st16(R1 + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&pk[16]));
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&a_ready[1][pp]);
              stage_store(pk, &om.du, t * 128);
            } else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&a_ready[1][pp]);
            }
          } else if (stage == 2) {
            // ---- dy = du W1 + dz*pad ; dh = LayerNorm2-backward(dy) -> TMEM (in place) and HBM ; LN parameter gradients
            uint32_t hp[32];
            float mean = 0.f, rstd = 0.f;
            if (has_half) {
              if (row_ok) {
                mean = p.mean[m];
                rstd = p.rstd[m];
              }
              const uint4* hr = reinterpret_cast<const uint4*>(p.h + (size_t)(row_ok ? m : 0) * D + c0);
#pragma unroll
              for (int c8 = 0; c8 < 8; ++c8) {
                const uint4 hv = row_ok ? __ldg(hr + c8) : make_uint4(0u, 0u, 0u, 0u);
                hp[c8 * 4] = hv.x; hp[c8 * 4 + 1] = hv.y; hp[c8 * 4 + 2] = hv.z; hp[c8 * 4 + 3] = hv.w;
              }
            }
            const float nmr = row_ok ? -mean * rstd : 0.f, rs_ok = row_ok ? rstd : 0.f;
            mbar_wait(&g_full[1][pp], pph);
            tc_fence_after();
            float dy[64];
            float s1 = 0.f, s2 = 0.f;
            if (has_half) {
              {
                uint32_t r0[32];
                tmem_ld32(R0 + c0, r0);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 32; ++q) dy[q] = __uint_as_float(r0[q]);
                tmem_ld32(R0 + c0 + 32, r0);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 32; ++q) dy[q + 32] = __uint_as_float(r0[q]);
              }
#pragma unroll
              for (int c8 = 0; c8 < 8; ++c8) {  // residual branch: the (masked) incoming gradient, still staged in shared memory
                const uint4 zv = *reinterpret_cast<const uint4*>(ztile + sw128_off((uint32_t)row, (uint32_t)c8));
                const __nv_bfloat162* z2 = reinterpret_cast<const __nv_bfloat162*>(&zv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(z2[e]);
                  dy[c8 * 8 + 2 * e] = fmaf(f.x, rm, dy[c8 * 8 + 2 * e]);
                  dy[c8 * 8 + 2 * e + 1] = fmaf(f.y, rm, dy[c8 * 8 + 2 * e + 1]);
                }
              }
#pragma unroll
              for (int q = 0; q < 64; q += 2) {
                const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hp[q >> 1]));
                const float g0 = dy[q] * s_lnw[c0 + q], g1 = dy[q + 1] * s_lnw[c0 + q + 1];
                s1 += g0 + g1;
                s2 = fmaf(g0, fmaf(hf.x, rs_ok, nmr), fmaf(g1, fmaf(hf.y, rs_ok, nmr), s2));
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&z_empty[s]);   // the staged dz tile is no longer needed by this warp
            s_stat[half][row] = make_float2(s1, s2);
            asm volatile("bar.sync 1, %0;" ::"r"(kPbEpiWarps * 32) : "memory");
            const float2 sa = s_stat[0][row], sb = (D > 64) ? s_stat[1][row] : make_float2(0.f, 0.f);
            const float inv_d = 1.f / (float)feat_count(D, p.hd_valid);
            const float m1 = (sa.x + sb.x) * inv_d, m2 = (sa.y + sb.y) * inv_d;
            asm volatile("bar.sync 1, %0;" ::"r"(kPbEpiWarps * 32) : "memory");
            if (has_half) {
              uint32_t pk[32];
#pragma unroll
              for (int q = 0; q < 64; q += 2) {
                const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hp[q >> 1]));
                float t0 = rstd * (dy[q] * s_lnw[c0 + q] - m1 - fmaf(hf.x, rs_ok, nmr) * m2);
                float t1 = rstd * (dy[q + 1] * s_lnw[c0 + q + 1] - m1 - fmaf(hf.y, rs_ok, nmr) * m2);
                if (p.hd_valid > 0) {   // padded inputs of the LayerNorm do not exist: no gradient
                  if (!feat_valid(c0 + q, p.hd_valid)) t0 = 0.f;
                  if (!feat_valid(c0 + q + 1, p.hd_valid)) t1 = 0.f;
                }
                pk[q >> 1] = pack_bf16(t0, t1);
              }
              tmem_st16(R0 + c0, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
              tmem_st16(R0 + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&pk[16]));
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&a_ready[2][pp]);   // GEMM 3 may start while dh travels and the column sums are formed
              stage_store(pk, &om.dh, t * 128);
            } else {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&a_ready[2][pp]);
            }
            if (has_half) {
              float pw[64];
#pragma unroll
              for (int q = 0; q < 64; q += 2) {
                const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hp[q >> 1]));
                pw[q] = dy[q] * fmaf(hf.x, rs_ok, nmr);
                pw[q + 1] = dy[q + 1] * fmaf(hf.y, rs_ok, nmr);
              }
              warp_colsum64_pb(pw, lane);
              acc_w0 += pw[0];
              acc_w1 += pw[1];
              warp_colsum64_pb(dy, lane);
              acc_b0 += dy[0];
              acc_b1 += dy[1];
            }
          } else {
            // ---- d_o = dh Wo -> HBM (gradient of the attention output)
            mbar_wait(&g_full[2][pp], pph);
            tc_fence_after();
            if (has_half) {
              uint32_t r0[32], r1[32];
              tmem_ld32(R1 + c0, r0);
              tmem_ld32(R1 + c0 + 32, r1);
              tmem_ld_wait();
              tc_fence_before();   // the next pair's stage 0 overwrites R0 / GEMM 1 overwrites R1: ordered after these loads
              uint32_t po[32];
#pragma unroll
              for (int q = 0; q < 32; q += 2) {
                po[q >> 1] = pack_bf16(__uint_as_float(r0[q]), __uint_as_float(r0[q + 1]));
                po[16 + (q >> 1)] = pack_bf16(__uint_as_float(r1[q]), __uint_as_float(r1[q + 1]));
              }
              stage_store(po, &om.d_o, t * 128);
            }
          }
        }
      }
    }
    if (leader) tma_store_wait_all();   // the staging tile must outlive its last store
    s_red[0][ew][2 * lane] = acc_w0;
    s_red[0][ew][2 * lane + 1] = acc_w1;
    s_red[1][ew][2 * lane] = acc_b0;
    s_red[1][ew][2 * lane + 1] = acc_b1;
    asm volatile("bar.sync 1, %0;" ::"r"(kPbEpiWarps * 32) : "memory");
    const int tid = threadIdx.x - 64;
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
static int launch_post_attn_bwd(const CUtensorMap& tmDZ, const CUtensorMap& tmW2, const CUtensorMap& tmW1,
                                const CUtensorMap& tmWo, const PostAttnBwdParams& p, cudaStream_t st) {
  constexpr int D = KCH * 64;
  constexpr int NA = 2;   // two tiles in flight keep their dz tile until their LayerNorm stage
  const int smem = 3 * KCH * KCH * 8192 + NA * KCH * 128 * 128 + 2 * 128 * 128 + 1024;
  PostAttnBwdOutMaps om;
  int rc;
  if ((rc = make_tmap_bf16(&om.du, p.du, p.T, D, D, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&om.dh, p.dh, p.T, D, D, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&om.d_o, p.d_o, p.T, D, D, 128)) != RP_OK) return rc;
  om.d_t = om.du;
  if (p.d_t && (rc = make_tmap_bf16(&om.d_t, p.d_t, p.T, D, D, 128)) != RP_OK) return rc;
  auto kern = post_attn_bwd_kernel<KCH, NA>;
  RP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int n_tiles = (p.T + 127) / 128;
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  kern<<<grid, kPbThreads, smem, st>>>(tmDZ, tmW2, tmW1, tmWo, om, p);
  RP_LAUNCH_CHECK();
  return RP_OK;
}

}  // namespace rp

using namespace rp;

// dz, u, h bf16 [T, d]; mean / rstd fp32 [T] (LayerNorm2 statistics saved by rp_post_attn_train); ln_w fp32 [d]; w2, w1, wo bf16
// [d, d] (row = output feature); rowmask optional uint8 [T].  Dropout site 2 is regenerated from (seed + *seed_ptr, drop_off2,
// element index) exactly as rp_post_attn_train / rp_gemm drew it; site 1 is encoded in the zeros of u.
// Outputs bf16 [T, d]: d_t (may be NULL when drop_p == 0 and rowmask == NULL: then d_t == dz), du, dh, d_o (none may alias an
// input); dln_w / dln_b fp32 [d] are ACCUMULATED (one atomic per column and CTA).  d in {64, 128}.
RP_API int rp_post_attn_bwd(const void* dz, const void* u, const void* h, const float* mean, const float* rstd, const float* ln_w,
                            const void* w2, const void* w1, const void* wo, const uint8_t* rowmask, int T, int d, float drop_p,
                            unsigned long long seed, unsigned long long drop_off2, const unsigned long long* seed_ptr, void* d_t,
                            void* du, void* dh, void* d_o, float* dln_w, float* dln_b, int hd_valid, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!dz || !u || !h || !mean || !rstd || !ln_w || !w2 || !w1 || !wo || !du || !dh || !d_o || !dln_w || !dln_b || T <= 0)
    return RP_EINVAL;
  if (d != 64 && d != 128) return RP_ESHAPE;
  if (drop_p < 0.f || drop_p >= 1.f || (drop_off2 & 3)) return RP_EINVAL;
  if (hd_valid < 0 || hd_valid > 128 || (hd_valid > 0 && d % (hd_valid <= 64 ? 64 : 128))) return RP_ESHAPE;
  if (!d_t && (drop_p > 0.f || rowmask)) return RP_EINVAL;
  if (d_t == dz || du == dz || dh == dz || d_o == dz) return RP_EINVAL;
  CUtensorMap tmDZ, tmW2, tmW1, tmWo;
  int rc;
  if ((rc = make_tmap_bf16(&tmDZ, dz, T, d, d, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmW2, w2, d, d, d, 64)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmW1, w1, d, d, d, 64)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmWo, wo, d, d, d, 64)) != RP_OK) return rc;
  PostAttnBwdParams p;
  p.u = reinterpret_cast<const __nv_bfloat16*>(u); p.h = reinterpret_cast<const __nv_bfloat16*>(h);
  p.mean = mean; p.rstd = rstd; p.ln_w = ln_w; p.rowmask = rowmask;
  p.d_t = reinterpret_cast<__nv_bfloat16*>(d_t); p.du = reinterpret_cast<__nv_bfloat16*>(du);
  p.dh = reinterpret_cast<__nv_bfloat16*>(dh); p.d_o = reinterpret_cast<__nv_bfloat16*>(d_o);
  p.dln_w = dln_w; p.dln_b = dln_b; p.drop_p = drop_p; p.seed = seed; p.off2 = drop_off2; p.seed_ptr = seed_ptr; p.T = T;
  p.hd_valid = hd_valid;
  return d == 64 ? launch_post_attn_bwd<1>(tmDZ, tmW2, tmW1, tmWo, p, stream)
                 : launch_post_attn_bwd<2>(tmDZ, tmW2, tmW1, tmWo, p, stream);
}
