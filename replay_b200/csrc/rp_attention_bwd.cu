// rp_attention_bwd.cu - fused attention backward on tcgen05 for L <= 256, head_dim 64: one CTA per (sequence, head).
//
// Replaces autograd's backward of torch.nn.MultiheadAttention's SDPA core
// (replay/nn/sequential/sasrec/transformer.py:99-106, models/nn/sequential/sasrec/model.py:435, bert4rec/model.py:494)
// without materialising the [B*H, L, L] probability / gradient matrices the un-fused path needs.
//
// Keys are the M (TMEM lane) dimension so that dK and dV accumulate with A operands taken straight from TMEM:
//   for each 128-key tile kt, for each 128-query tile qt (skipped when causally empty):
//     S^T  = K_kt . Q_qt^T          (SS, fp32 in TMEM)            dP^T = V_kt . dO_qt^T        (SS)
//     thread = key row j:  P = exp2(s*sl2 - m_i) * inv_i (masked), dP = dP^T * dropmask/keep,
//                          dS = P * (dP - delta_i) * scale,  Pd = P * dropmask/keep
//       Pd^T (bf16) -> TMEM over S^T,  dS^T (bf16) -> TMEM over dP^T  and -> shared memory (swizzled, MN-major)
//     dV_kt += Pd^T . dO_qt   (A from TMEM, B = dO tile read MN-major)
//     dK_kt += dS^T . Q_qt    (A from TMEM, B = Q tile read MN-major)
//     dQ_qt += dS  . K_kt     (A = dS^T tile in smem read MN-major, B = K tile read MN-major)
//   row statistics m_i (max in exp2 units) and inv_i (1/rowsum) come from the forward; delta_i = sum_c dO[i,c] O[i,c].
// TMEM: S^T 128 | dP^T 128 | dK 64 | dV 64 | dQ (2 x 64) = 512 columns.
#include "rp_host.h"
#include "rp_philox.cuh"
#include "rp_sm100.cuh"

namespace rp {

struct AttnBwdParams {
  int B, H, L, Lp;
  int causal, mask_pad_keys;
  float scale;
  const uint8_t* pad_mask;
  const __nv_bfloat16* O;      // [T, ldo] forward output (for delta)
  const __nv_bfloat16* dO;     // [T, ld_do]
  int ldo, ld_do;
  const float* m_save;         // [B*H, Lp]
  const float* inv_sum;        // [B*H, Lp]
  __nv_bfloat16* dQ; int ld_dq, dq_c0;   // outputs: rows b*L + i, columns x_c0 + h*64
  __nv_bfloat16* dK; int ld_dk, dk_c0;
  __nv_bfloat16* dV; int ld_dv, dv_c0;
  int q_c0, k_c0, v_c0;        // column offsets of head 0 inside the Q / K / V arrays (tensor maps)
  float drop_p;
  unsigned long long seed, drop_off;
  const unsigned long long* seed_ptr;
};

static constexpr float kL2e = 1.4426950408889634f;

static constexpr int kAbThreads = 32 + 8 * 32;  // issuer warp + 8 softmax warps (lane quarter x 64-query column half)

__global__ void __launch_bounds__(kAbThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, const AttnBwdParams p) {
  constexpr int HD = 64;
  constexpr int TILE = 128 * 128;  // bytes of one [128 rows x 64 bf16] swizzled tile
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;               // 2 tiles (queries 0-127, 128-255)
  uint8_t* sdO = sQ + 2 * TILE;
  uint8_t* sK = sdO + 2 * TILE;     // 2 tiles (keys)
  uint8_t* sV = sK + 2 * TILE;
  uint8_t* sdS = sV + 2 * TILE;     // [128 keys x 128 queries] bf16 as two 64-query chunks
  __shared__ float s_m[256], s_inv[256], s_delta[256];
  __shared__ uint64_t bar_load, bar_s, bar_p, bar_acc;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % p.H, b = blockIdx.x / p.H;
  const int bz = b * p.H + h;
  const int L = p.L;
  const int n_t = (L + 127) / 128;  // tiles along queries and along keys

  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_p, 8);
    mbar_init(&bar_acc, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  // row statistics of this (sequence, head): m, 1/sum from the forward, delta from O and dO
  if (threadIdx.x >= 32) {
    for (int i = threadIdx.x - 32; i < 256; i += 256) {
      float m = 0.f, inv = 0.f, dl = 0.f;
      if (i < L) {
        m = p.m_save[(size_t)bz * p.Lp + i];
        inv = p.inv_sum[(size_t)bz * p.Lp + i];
        const uint4* o4 = reinterpret_cast<const uint4*>(p.O + ((size_t)b * L + i) * p.ldo + h * HD);
        const uint4* d4 = reinterpret_cast<const uint4*>(p.dO + ((size_t)b * L + i) * p.ld_do + h * HD);
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) {
          const uint4 ov = o4[c], dv = d4[c];
          const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&ov);
          const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&dv);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 a = __bfloat1622float2(o2[t]), g = __bfloat1622float2(d2[t]);
            dl = fmaf(a.x, g.x, fmaf(a.y, g.y, dl));
          }
        }
      }
      s_m[i] = m;
      s_inv[i] = inv;
      s_delta[i] = dl;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_st = tmem, t_dpt = tmem + 128, t_dk = tmem + 256, t_dv = tmem + 320, t_dq = tmem + 384;

  // which (kt, qt) pairs carry any visible entry
  auto pair_active = [&](int kt, int qt) { return !(p.causal && qt * 128 + 127 < kt * 128); };

  if (warp == 0) {
    if (elect_one()) {
      const int row0 = b * L;
      mbar_arrive_expect_tx(&bar_load, 4 * n_t * TILE);
      for (int t = 0; t < n_t; ++t) {
        tma_load_2d(sQ + t * TILE, &tmQ, &bar_load, p.q_c0 + h * HD, row0 + t * 128);
        tma_load_2d(sdO + t * TILE, &tmdO, &bar_load, h * HD, row0 + t * 128);
        tma_load_2d(sK + t * TILE, &tmK, &bar_load, p.k_c0 + h * HD, row0 + t * 128);
        tma_load_2d(sV + t * TILE, &tmV, &bar_load, p.v_c0 + h * HD, row0 + t * 128);
      }
      mbar_wait(&bar_load, 0);
      tc_fence_after();
      constexpr uint32_t id_s = umma_idesc_bf16(128, 128);                 // S^T, dP^T: SS K-major
      constexpr uint32_t id_kv = umma_idesc_bf16(128, HD, false, true);    // dK, dV: A TMEM, B MN-major
      constexpr uint32_t id_q = umma_idesc_bf16(128, HD, true, true);      // dQ: A smem MN-major, B MN-major
      uint32_t pp = 0;  // phase counter of bar_p (one phase per pair step and one per dK/dV drain)
      bool dq_started[2] = {false, false};
      for (int kt = 0; kt < n_t; ++kt) {
        bool kv_started = false;
        for (int qt = 0; qt < n_t; ++qt) {
          if (!pair_active(kt, qt)) continue;
          // tcgen05.mma ops execute in issue order, so these may overwrite S^T / dP^T behind the previous pair's
          // second-stage MMAs; the softmax threads only touch them after bar_s, which also covers those MMAs
          const uint32_t k0 = smem_u32(sK + kt * TILE), v0 = smem_u32(sV + kt * TILE);
          const uint32_t q0 = smem_u32(sQ + qt * TILE), g0 = smem_u32(sdO + qt * TILE);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma_ss(t_st, umma_desc_sw128(k0 + ks * 32, 16, 1024), umma_desc_sw128(q0 + ks * 32, 16, 1024), id_s, ks != 0);
            umma_ss(t_dpt, umma_desc_sw128(v0 + ks * 32, 16, 1024), umma_desc_sw128(g0 + ks * 32, 16, 1024), id_s, ks != 0);
          }
          umma_commit(&bar_s);
          mbar_wait(&bar_p, pp & 1);
          ++pp;
          tc_fence_after();
          const uint32_t ds0 = smem_u32(sdS);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {  // contraction over the 128 queries of this tile, 16 at a time
            const uint32_t acol = (ks >> 2) * 64 + (ks & 3) * 8;  // queries 0-63 packed at +0, 64-127 at +64
            umma_ts(t_dv, t_st + acol, umma_desc_sw128(g0 + ks * 2048, 16, 1024), id_kv, kv_started || ks != 0);
            umma_ts(t_dk, t_dpt + acol, umma_desc_sw128(q0 + ks * 2048, 16, 1024), id_kv, kv_started || ks != 0);
          }
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)    // contraction over the 128 keys of this tile
            umma_ss(t_dq + qt * HD, umma_desc_sw128(ds0 + ks * 2048, 16384, 1024), umma_desc_sw128(k0 + ks * 2048, 16, 1024), id_q,
                    dq_started[qt] || ks != 0);
          kv_started = true;
          dq_started[qt] = true;
        }
        // dK / dV of this key tile are complete once everything issued so far has retired: hand them to the epilogue
        // threads and wait until they are drained before the next key tile restarts the accumulators
        umma_commit(&bar_acc);
        mbar_wait(&bar_p, pp & 1);
        ++pp;
        tc_fence_after();
      }
    }
  } else {
    // ------------------------------------------------ 8 warps: thread = key row (then query row for dQ) x column half
    const int quarter = warp & 3, cgp = (warp - 1) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale * kL2e;
    const uint32_t thr = p.drop_p > 0.f ? (uint32_t)(p.drop_p * 4294967296.0) : 0u;
    const float ks_drop = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    const unsigned long long seed_eff = p.seed + ((p.drop_p > 0.f && p.seed_ptr) ? *p.seed_ptr : 0ull);
    uint32_t np = 0, acc_it = 0;  // pairs seen (bar_s phase), key tiles seen (bar_acc phase)
    for (int kt = 0; kt < n_t; ++kt) {
      const int j = kt * 128 + row;  // key position
      bool key_ok = j < L;
      if (key_ok && p.mask_pad_keys) key_ok = p.pad_mask[(size_t)b * L + j] != 0;
      for (int qt = 0; qt < n_t; ++qt) {
        if (!pair_active(kt, qt)) continue;
        mbar_wait(&bar_s, np & 1);
        ++np;
        tc_fence_after();
#pragma unroll 1
        for (int c = cgp * 64; c < cgp * 64 + 64; c += 32) {
          uint32_t rs[32], rd[32];
          tmem_ld32(t_st + lane_base + c, rs);
          tmem_ld32(t_dpt + lane_base + c, rd);
          tmem_ld_wait();
          uint32_t pk_p[16], pk_s[16];
#pragma unroll
          for (int q = 0; q < 32; q += 2) {
            float pd2[2], ds2[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int i = qt * 128 + c + q + e;  // query position
              const bool vis = key_ok && i < L && (!p.causal || j <= i);
              float pr = vis ? ex2f(fmaf(__uint_as_float(rs[q + e]), sl2, -s_m[i])) * s_inv[i] : 0.f;
              float keep = ks_drop;
              if (p.drop_p > 0.f) {
                const unsigned long long idx = p.drop_off + ((unsigned long long)bz * p.Lp + (unsigned long long)i) * p.Lp + j;
                keep = drop_hash32(seed_eff, idx) >= thr ? ks_drop : 0.f;
              }
              const float dp = __uint_as_float(rd[q + e]) * keep;
              ds2[e] = pr * (dp - s_delta[i]) * p.scale;
              pd2[e] = pr * keep;
            }
            pk_p[q >> 1] = pack_bf16(pd2[0], pd2[1]);
            pk_s[q >> 1] = pack_bf16(ds2[0], ds2[1]);
          }
          // bf16 operands go into the first half of THIS warp's own (already consumed) 64 columns
          tmem_st16(t_st + lane_base + cgp * 64 + ((c & 63) >> 1), pk_p);    // Pd^T over S^T
          tmem_st16(t_dpt + lane_base + cgp * 64 + ((c & 63) >> 1), pk_s);   // dS^T over dP^T
          // dS^T also to shared memory as the MN-major A operand of dQ: row = key (K index), 64-query chunks
          uint8_t* dst = sdS + (c >> 6) * 16384;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const uint32_t chunk16 = (uint32_t)(((c & 63) >> 3) + g8);
            *reinterpret_cast<uint4*>(dst + sw128_off((uint32_t)row, chunk16)) =
                make_uint4(pk_s[g8 * 4], pk_s[g8 * 4 + 1], pk_s[g8 * 4 + 2], pk_s[g8 * 4 + 3]);
          }
        }
        tmem_st_wait();
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_p);
      }
      // ---- dK / dV of this key tile
      mbar_wait(&bar_acc, acc_it & 1);
      tc_fence_after();
      ++acc_it;
      {
        const int which = cgp;  // column-half 0 drains dK, column-half 1 drains dV
        const uint32_t src = (which == 0 ? t_dk : t_dv) + lane_base;
        __nv_bfloat16* outp = which == 0 ? p.dK + ((size_t)b * L + j) * p.ld_dk + p.dk_c0 + h * HD
                                         : p.dV + ((size_t)b * L + j) * p.ld_dv + p.dv_c0 + h * HD;
#pragma unroll
        for (int c = 0; c < HD; c += 32) {
          uint32_t r[32];
          tmem_ld32(src + c, r);
          tmem_ld_wait();
          if (j < L) {
#pragma unroll
            for (int q = 0; q < 32; q += 8) {
              uint4 w;
              w.x = pack_bf16(__uint_as_float(r[q]), __uint_as_float(r[q + 1]));
              w.y = pack_bf16(__uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
              w.z = pack_bf16(__uint_as_float(r[q + 4]), __uint_as_float(r[q + 5]));
              w.w = pack_bf16(__uint_as_float(r[q + 6]), __uint_as_float(r[q + 7]));
              *reinterpret_cast<uint4*>(outp + c + q) = w;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_p);  // accumulators drained: the issuer may start the next key tile
    }
    // ---- dQ: all pairs done (the last bar_acc wait above ordered every MMA); thread = query row
    for (int qt = 0; qt < n_t; ++qt) {
      const int i = qt * 128 + row;
      {
        const int c = cgp * 32;
        uint32_t r[32];
        tmem_ld32(t_dq + qt * HD + lane_base + c, r);
        tmem_ld_wait();
        if (i < L) {
          __nv_bfloat16* outp = p.dQ + ((size_t)b * L + i) * p.ld_dq + p.dq_c0 + h * HD + c;
#pragma unroll
          for (int q = 0; q < 32; q += 8) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(r[q]), __uint_as_float(r[q + 1]));
            w.y = pack_bf16(__uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
            w.z = pack_bf16(__uint_as_float(r[q + 4]), __uint_as_float(r[q + 5]));
            w.w = pack_bf16(__uint_as_float(r[q + 6]), __uint_as_float(r[q + 7]));
            *reinterpret_cast<uint4*>(outp + q) = w;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace rp

using namespace rp;

struct rp_attn_bwd_desc {
  const void* q; long long q_rows, q_cols, ldq; int q_c0;
  const void* k; long long k_rows, k_cols, ldk; int k_c0;
  const void* v; long long v_rows, v_cols, ldv; int v_c0;
  const void* d_out; long long do_rows, do_cols, ld_do;
  const void* out; int ldo;
  int B, H, L, head_dim;
  int causal, mask_pad_keys;
  const uint8_t* pad_mask;
  const float* m_save; const float* inv_sum;
  void* dq; int ld_dq, dq_c0;
  void* dk; int ld_dk, dk_c0;
  void* dv; int ld_dv, dv_c0;
  float drop_p; unsigned long long seed, drop_off; const unsigned long long* seed_ptr;
  float scale;
};

RP_API int rp_attn_bwd(const rp_attn_bwd_desc* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->q || !a->k || !a->v || !a->d_out || !a->out || !a->pad_mask || !a->m_save || !a->inv_sum || !a->dq || !a->dk ||
      !a->dv)
    return RP_EINVAL;
  if (a->L <= 0 || a->L > 256 || a->B <= 0 || a->H <= 0 || a->head_dim != 64) return RP_ESHAPE;
  if ((a->ld_dq & 7) || (a->ld_dk & 7) || (a->ld_dv & 7) || (a->ldo & 7) || (a->ld_do & 7) || (a->dq_c0 & 7) || (a->dk_c0 & 7) ||
      (a->dv_c0 & 7))
    return RP_EALIGN;
  AttnBwdParams p;
  p.B = a->B; p.H = a->H; p.L = a->L; p.Lp = (a->L + 63) & ~63;
  p.causal = a->causal; p.mask_pad_keys = a->mask_pad_keys;
  p.scale = a->scale > 0.f ? a->scale : 1.f / sqrtf((float)a->head_dim);
  p.pad_mask = a->pad_mask;
  p.O = reinterpret_cast<const __nv_bfloat16*>(a->out); p.ldo = a->ldo;
  p.dO = reinterpret_cast<const __nv_bfloat16*>(a->d_out); p.ld_do = (int)a->ld_do;
  p.m_save = a->m_save; p.inv_sum = a->inv_sum;
  p.dQ = reinterpret_cast<__nv_bfloat16*>(a->dq); p.ld_dq = a->ld_dq; p.dq_c0 = a->dq_c0;
  p.dK = reinterpret_cast<__nv_bfloat16*>(a->dk); p.ld_dk = a->ld_dk; p.dk_c0 = a->dk_c0;
  p.dV = reinterpret_cast<__nv_bfloat16*>(a->dv); p.ld_dv = a->ld_dv; p.dv_c0 = a->dv_c0;
  p.q_c0 = a->q_c0; p.k_c0 = a->k_c0; p.v_c0 = a->v_c0;
  p.drop_p = a->drop_p; p.seed = a->seed; p.drop_off = a->drop_off; p.seed_ptr = a->seed_ptr;
  CUtensorMap tmQ, tmK, tmV, tmdO;
  int rc;
  if ((rc = make_tmap_bf16(&tmQ, a->q, a->q_rows, a->q_cols, a->ldq, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmK, a->k, a->k_rows, a->k_cols, a->ldk, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmV, a->v, a->v_rows, a->v_cols, a->ldv, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmdO, a->d_out, a->do_rows, a->do_cols, a->ld_do, 128)) != RP_OK) return rc;
  const int smem = 10 * 128 * 128 + 1024;
  RP_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  attn_bwd_kernel<<<a->B * a->H, kAbThreads, smem, stream>>>(tmQ, tmK, tmV, tmdO, p);
  RP_LAUNCH_CHECK();
  return RP_OK;
}
