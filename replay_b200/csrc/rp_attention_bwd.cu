// rp_attention_bwd.cu - fused attention backward on tcgen05 for L <= 256, head_dim 64: one CTA per (sequence, head).
//
// Replaces autograd's backward of torch.nn.MultiheadAttention's SDPA core
// (replay/nn/sequential/sasrec/transformer.py:99-106, models/nn/sequential/sasrec/model.py:435, bert4rec/model.py:494)
// without materialising the [B*H, L, L] probability / gradient matrices the un-fused path needs.
//
// Keys are the M (TMEM lane) dimension so that dK and dV accumulate with A operands taken straight from TMEM.  The work
// is cut into STEPS of (128-key tile kt) x (64 queries qs); causally empty steps are skipped.  Per step:
//     S^T  = K_kt . Q_qs^T          (SS, fp32 in TMEM, N = live queries)     dP^T = V_kt . dO_qs^T        (SS)
//     thread = key row j, 32 of the 64 query columns:
//                          P = exp2(s*sl2 - m_i) * inv_i (masked), dP = dP^T * dropmask/keep,
//                          dS = P * (dP - delta_i) * scale,  Pd = P * dropmask/keep
//       Pd^T (bf16) -> TMEM over S^T,  dS^T (bf16) -> TMEM over dP^T  and -> shared memory (swizzled, MN-major)
//     dV_kt += Pd^T . dO_qs   (A from TMEM, B = dO rows read MN-major)
//     dK_kt += dS^T . Q_qs    (A from TMEM, B = Q rows read MN-major)
//   and once both 64-query halves of a 128-query tile are done:
//     dQ_qt += dS  . K_kt     (A = dS^T tile in smem read MN-major, B = K tile read MN-major)
// S^T / dP^T are DOUBLE BUFFERED: tcgen05.mma executes in issue order, so the issuer queues the first-stage MMAs of step
// s+1 before it waits for the element-wise warps of step s - the tensor pipe and the 8 element-wise warps overlap instead of
// alternating (round 1: one buffer, 187 us per launch at config 2, issue slots 38 % busy).  32 x 32 chunks that are entirely
// masked (above the causal diagonal, padded keys, rows / columns beyond L) are zero-filled without loading or computing.
//   row statistics m_i (max in exp2 units) and inv_i (1/rowsum) come from the forward; delta_i = sum_c dO[i,c] O[i,c].
// TMEM: 2 x (S^T 64 | dP^T 64) | dK 64 | dV 64 | dQ (2 x 64) = 512 columns.
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

#ifdef RP_ATTN_TRACE  // diagnostic build (RP_NVCC_EXTRA=-DRP_ATTN_TRACE): per-CTA phase timestamps of the first compute warp
__device__ unsigned long long g_attn_trace[1024 * 32];
#define RP_TR(k) do { if (threadIdx.x == 32 && (k) < 32) g_attn_trace[blockIdx.x * 32 + (k)] = clock64(); } while (0)
#else
#define RP_TR(k) do { } while (0)
#endif

static constexpr int kAbThreads = 32 + 8 * 32;  // issuer warp + 8 element-wise warps (lane quarter x 32-query column half)

// the eight 3-D tensor maps [B][L][columns] (box 128 rows x 64 columns, rows >= L out of bounds: zero-filled / not stored)
struct AttnBwdMaps {
  CUtensorMap q, k, v, d_o, o, dq, dk, dv;
};

// One 32 x 32 chunk of a step: thread = key row, 32 query columns starting at i0.  GENERAL: per-element visibility
// (diagonal / ragged chunks); otherwise every element is visible.  DROP: probability dropout on.
template <bool GENERAL, bool DROP>
__device__ __forceinline__ void attn_bwd_chunk(const uint32_t (&rs)[32], const uint32_t (&rd)[32], uint32_t (&pk_p)[16],
                                               uint32_t (&pk_s)[16], const float4* __restrict__ s_stat, int i0, int j, int L,
                                               bool key_ok, bool causal, float sl2, float scale, uint32_t thr, float ks_drop,
                                               uint32_t col_key) {
  const float keep_s = ks_drop * scale;
#pragma unroll
  for (int q = 0; q < 32; q += 2) {
    float pd2[2], ds2[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = i0 + q + e;                   // query position
      const float4 st = s_stat[i];                // m, 1/sum, delta * scale, row key (bits)
      const float pr = ex2f(fmaf(__uint_as_float(rs[q + e]), sl2, -st.x)) * st.y;
      float kp = ks_drop, kps = keep_s;
      if (DROP) {
        const bool keep = drop_mix(__float_as_uint(st.w), col_key) >= thr;
        kp = keep ? ks_drop : 0.f;
        kps = keep ? keep_s : 0.f;
      }
      ds2[e] = pr * fmaf(__uint_as_float(rd[q + e]), kps, -st.z);
      pd2[e] = pr * kp;
      if (GENERAL) {  // selects, not multiplies: columns beyond the MMA's N extent hold stale TMEM contents
        const bool vis = key_ok && i < L && (!causal || j <= i);
        ds2[e] = vis ? ds2[e] : 0.f;
        pd2[e] = vis ? pd2[e] : 0.f;
      }
    }
    pk_p[q >> 1] = pack_bf16(pd2[0], pd2[1]);
    pk_s[q >> 1] = pack_bf16(ds2[0], ds2[1]);
  }
}

__global__ void __launch_bounds__(kAbThreads, 1)
attn_bwd_kernel(const __grid_constant__ AttnBwdMaps tm, const AttnBwdParams p) {
  constexpr int HD = 64;
  constexpr int TILE = 128 * 128;  // bytes of one [128 rows x 64 bf16] swizzled tile
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;               // 2 tiles (queries 0-127, 128-255)
  uint8_t* sdO = sQ + 2 * TILE;
  uint8_t* sK = sdO + 2 * TILE;     // 2 tiles (keys)
  uint8_t* sV = sK + 2 * TILE;
  uint8_t* sdS = sV + 2 * TILE;     // 2 buffers x [128 keys x 128 queries] bf16 as two 64-query chunks each; at the start the
                                    // first buffer receives the two O tiles (for delta)
  uint8_t* sStage = sdS + 4 * TILE; // one output tile on its way to HBM (TMA store) while the main loop still owns the rest
  __shared__ float4 s_stat[256];    // per query: m, 1/sum, delta * scale, dropout row key
  __shared__ uint64_t bar_load[2], bar_s[2], bar_p[2], bar_acc;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % p.H, b = blockIdx.x / p.H;
  const int bz = b * p.H + h;
  const int L = p.L;
  const int n_t = (L + 127) / 128;  // 128-key tiles / 128-query tiles
  const int n_qs = (L + 63) / 64;   // 64-query steps
  RP_TR(0);

  // The steps, in the same order in every thread: (kt, qs) with the causally empty ones (qs < 2 kt) left out.
  //   code = (kt << 4) | qs, -1 past the end
  auto first_step = [&]() -> int { return 0; };
  auto next_step = [&](int code) -> int {
    const int kt = code >> 4, qs = code & 15;
    if (qs + 1 < n_qs) return code + 1;
    if (kt + 1 < n_t) return ((kt + 1) << 4) | (p.causal ? 2 * (kt + 1) : 0);
    return -1;
  };
  auto same_pair = [](int a, int c) { return c >= 0 && ((a ^ c) & ~1) == 0; };   // same kt and same 128-query tile
  auto same_kt = [](int a, int c) { return c >= 0 && (a >> 4) == (c >> 4); };

  if (threadIdx.x == 0) {
    mbar_init(&bar_load[0], 1);
    mbar_init(&bar_load[1], 1);
    mbar_init(&bar_s[0], 1);
    mbar_init(&bar_s[1], 1);
    mbar_init(&bar_p[0], 8);
    mbar_init(&bar_p[1], 8);
    mbar_init(&bar_acc, 1);
    fence_barrier_init();
    // the loads go out before anything else: tile 0 of everything first (the first steps need nothing more)
    for (int t = 0; t < n_t; ++t) {
      mbar_arrive_expect_tx(&bar_load[t], 5 * TILE);
      tma_load_3d(sK + t * TILE, &tm.k, &bar_load[t], p.k_c0 + h * HD, t * 128, b);
      tma_load_3d(sQ + t * TILE, &tm.q, &bar_load[t], p.q_c0 + h * HD, t * 128, b);
      tma_load_3d(sV + t * TILE, &tm.v, &bar_load[t], p.v_c0 + h * HD, t * 128, b);
      tma_load_3d(sdO + t * TILE, &tm.d_o, &bar_load[t], h * HD, t * 128, b);
      tma_load_3d(sdS + t * TILE, &tm.o, &bar_load[t], h * HD, t * 128, b);
    }
  }
  auto live_q = [&](int qs) { return min(64, L - qs * 64); };   // live queries of a step (>= 1)
  // first-stage MMAs of one step into buffer buf (issuer thread only; tmem = TMEM base)
  auto issue_first = [&](uint32_t tmem, int code, int buf) {
    const int kt = code >> 4, qs = code & 15, qt = qs >> 1;
    const int n16 = (live_q(qs) + 15) & ~15;
    const uint32_t id_s = umma_idesc_bf16(128, n16);                   // S^T, dP^T: SS K-major
    const uint32_t k0 = smem_u32(sK + kt * TILE), v0 = smem_u32(sV + kt * TILE);
    const uint32_t q0 = smem_u32(sQ + qt * TILE) + (qs & 1) * 8192, g0 = smem_u32(sdO + qt * TILE) + (qs & 1) * 8192;
    const uint32_t t_st = tmem + (uint32_t)buf * 128u, t_dpt = t_st + 64u;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      umma_ss(t_st, umma_desc_sw128(k0 + ks * 32, 16, 1024), umma_desc_sw128(q0 + ks * 32, 16, 1024), id_s, ks != 0);
      umma_ss(t_dpt, umma_desc_sw128(v0 + ks * 32, 16, 1024), umma_desc_sw128(g0 + ks * 32, 16, 1024), id_s, ks != 0);
    }
    umma_commit(&bar_s[buf]);
  };

  bool loaded1 = false;
  bool key_ok2[2] = {false, false};
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncwarp();
    tc_fence_after();
    if (elect_one()) {
      // the first two steps go out before the block-wide sync: they only need the tiles, not the row statistics
      const uint32_t tmem = tmem_slot;
      mbar_wait(&bar_load[0], 0);
      tc_fence_after();
      const int c0 = first_step(), c1 = next_step(c0);
      issue_first(tmem, c0, 0);
      if (c1 >= 0) {
        if ((c1 >> 4) > 0 || ((c1 & 15) >> 1) > 0) {
          mbar_wait(&bar_load[1], 0);
          loaded1 = true;
        }
        issue_first(tmem, c1, 1);
      }
    }
    __syncwarp();
  } else {
    // row statistics of this (sequence, head): m, 1/sum from the forward, delta = sum_c dO[i,c] O[i,c] from the tiles the
    // TMA just delivered (row-per-thread global reads of O and dO cost 11 000 cycles per CTA here, trace r2f), the dropout key
    const int i = threadIdx.x - 32, ti = i >> 7, r = i & 127;
    // key validity of this thread's two key rows in the main loop (TMEM lane quarter = warp % 4): row, 128 + row
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int j = t * 128 + (warp & 3) * 32 + lane;
      key_ok2[t] = j < L && (!p.mask_pad_keys || p.pad_mask[(size_t)b * L + j] != 0);
    }
    const unsigned long long seed_eff = p.seed + ((p.drop_p > 0.f && p.seed_ptr) ? *p.seed_ptr : 0ull);
    float m = 0.f, inv = 0.f, dl = 0.f;
    if (i < L) {
      m = p.m_save[(size_t)bz * p.Lp + i];
      inv = p.inv_sum[(size_t)bz * p.Lp + i];
    }
    const uint32_t rk = drop_row_key(seed_eff, p.drop_off, (unsigned long long)bz * p.Lp + (unsigned long long)i);
    if (ti < n_t) {
      mbar_wait(&bar_load[ti], 0);
#pragma unroll
      for (int c = 0; c < HD / 8; ++c) {
        const uint4 ov = *reinterpret_cast<const uint4*>(sdS + ti * TILE + sw128_off((uint32_t)r, (uint32_t)c));
        const uint4 dv = *reinterpret_cast<const uint4*>(sdO + ti * TILE + sw128_off((uint32_t)r, (uint32_t)c));
        const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&ov);
        const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&dv);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 a = __bfloat1622float2(o2[t]), g = __bfloat1622float2(d2[t]);
          dl = fmaf(a.x, g.x, fmaf(a.y, g.y, dl));
        }
      }
    }
    s_stat[i] = make_float4(m, inv, dl * p.scale, __uint_as_float(rk));
  }
  RP_TR(1);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  RP_TR(2);
  const uint32_t tmem = tmem_slot;
  const uint32_t t_dk = tmem + 256, t_dv = tmem + 320, t_dq = tmem + 384;

  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t id_kv = umma_idesc_bf16(128, HD, false, true);    // dK, dV: A TMEM, B MN-major
      constexpr uint32_t id_q = umma_idesc_bf16(128, HD, true, true);      // dQ: A smem MN-major, B MN-major
      bool dq_started[2] = {false, false};
      bool kv_started = false;
      int pair_idx = 0;
      int cur = first_step(), nxt = next_step(cur);
      for (int s = 0; cur >= 0; ++s) {
        const int kt = cur >> 4, qs = cur & 15, qt = qs >> 1, buf = s & 1;
        mbar_wait(&bar_p[buf], (s >> 1) & 1);
        tc_fence_after();
        const uint32_t q0 = smem_u32(sQ + qt * TILE), g0 = smem_u32(sdO + qt * TILE), k0 = smem_u32(sK + kt * TILE);
        const uint32_t t_st = tmem + (uint32_t)buf * 128u, t_dpt = t_st + 64u;
        const int nks = (live_q(qs) + 15) >> 4;   // contraction over the live queries of this step, 16 at a time
        for (int ks = 0; ks < nks; ++ks) {
          const uint32_t acol = (ks >> 1) * 32 + (ks & 1) * 8;  // queries 0-31 packed at +0, 32-63 at +32
          const uint32_t boff = ((qs & 1) * 4 + ks) * 2048;
          umma_ts(t_dv, t_st + acol, umma_desc_sw128(g0 + boff, 16, 1024), id_kv, kv_started || ks != 0);
          umma_ts(t_dk, t_dpt + acol, umma_desc_sw128(q0 + boff, 16, 1024), id_kv, kv_started || ks != 0);
        }
        kv_started = true;
        if (!same_pair(cur, nxt)) {
          const uint32_t ds0 = smem_u32(sdS) + (pair_idx & 1) * 2 * TILE;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)    // contraction over the 128 keys of this tile
            umma_ss(t_dq + qt * HD, umma_desc_sw128(ds0 + ks * 2048, 16384, 1024), umma_desc_sw128(k0 + ks * 2048, 16, 1024), id_q,
                    dq_started[qt] || ks != 0);
          dq_started[qt] = true;
          ++pair_idx;
        }
        if (!same_kt(cur, nxt)) {
          // dK / dV of this key tile are complete once everything issued so far has retired; the element-wise warps drain
          // them before they arrive for the next step, which is what gates the next accumulate = 0 MMA
          umma_commit(&bar_acc);
          kv_started = false;
        }
        // tcgen05.mma executes in issue order: the first stage of step s + 2 reuses this step's buffer behind its second stage
        const int nn = nxt >= 0 ? next_step(nxt) : -1;
        if (nn >= 0) {
          if (!loaded1 && ((nn >> 4) > 0 || ((nn & 15) >> 1) > 0)) {
            mbar_wait(&bar_load[1], 0);
            loaded1 = true;
          }
          issue_first(tmem, nn, buf);
        }
        cur = nxt;
        nxt = nn;
      }
    }
  } else {
    // ------------------------------------------------ 8 warps: thread = key row (then query row for dQ) x column half
    const int quarter = warp & 3, cgp = (warp - 1) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale * kL2e;
    const bool drop = p.drop_p > 0.f;
    const uint32_t thr = drop ? (uint32_t)(p.drop_p * 4294967296.0) : 0u;
    const float ks_drop = drop ? 1.f / (1.f - p.drop_p) : 1.f;
    const bool leader = threadIdx.x == 32;
    // one [128 x 64] fp32 accumulator -> bf16 tile in shared memory (swizzled, ready for a TMA store); no synchronisation
    auto drain_to = [&](uint32_t src, uint8_t* stage) {
      uint32_t r[32];
      tmem_ld32(src + lane_base + cgp * 32, r);
      tmem_ld_wait();
      uint32_t w[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) w[q] = pack_bf16(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]));
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8)
        *reinterpret_cast<uint4*>(stage + sw128_off((uint32_t)row, (uint32_t)(cgp * 4 + g8))) =
            make_uint4(w[g8 * 4], w[g8 * 4 + 1], w[g8 * 4 + 2], w[g8 * 4 + 3]);
    };
    int pair_idx = 0;
    int lent_buf = -1;   // dS buffer lent to a TMA store of dV (mid-kernel drain): its next writer waits for the store's read
    int cur = first_step(), nxt = next_step(cur);
    for (int s = 0; cur >= 0; ++s) {
      const int kt = cur >> 4, qs = cur & 15, buf = s & 1;
      const int j = kt * 128 + row;  // key position
      const bool key_ok = kt == 0 ? key_ok2[0] : key_ok2[1];
      const uint32_t col_key = drop_col_key((uint32_t)j);
      const int i0 = qs * 64 + cgp * 32;          // first query column of this warp's chunk
      const int j0 = kt * 128 + quarter * 32;     // first key row of this warp
      const uint32_t t_st = tmem + (uint32_t)buf * 128u, t_dpt = t_st + 64u;
      mbar_wait(&bar_s[buf], (s >> 1) & 1);
      tc_fence_after();
      RP_TR(3 + 3 * s);
      if (lent_buf == (pair_idx & 1)) {
        if (leader) tma_store_wait_read();
        named_bar_sync(1, 256);
        lent_buf = -1;
      }
      if (i0 < L) {   // (columns beyond L are neither contracted over nor stored)
        uint32_t pk_p[16], pk_s[16];
        const bool any_key = __any_sync(0xffffffffu, key_ok);
        const bool empty = !any_key || (p.causal && j0 > i0 + 31);
        if (empty) {
#pragma unroll
          for (int q = 0; q < 16; ++q) pk_p[q] = pk_s[q] = 0u;
        } else {
          uint32_t rs[32], rd[32];
          tmem_ld32(t_st + lane_base + cgp * 32, rs);
          tmem_ld32(t_dpt + lane_base + cgp * 32, rd);
          tmem_ld_wait();
          const bool full = __all_sync(0xffffffffu, key_ok) && i0 + 31 < L && (!p.causal || j0 + 31 <= i0);
          if (full) {
            if (drop) attn_bwd_chunk<false, true>(rs, rd, pk_p, pk_s, s_stat, i0, j, L, key_ok, p.causal, sl2, p.scale, thr, ks_drop, col_key);
            else attn_bwd_chunk<false, false>(rs, rd, pk_p, pk_s, s_stat, i0, j, L, key_ok, p.causal, sl2, p.scale, thr, ks_drop, col_key);
          } else {
            if (drop) attn_bwd_chunk<true, true>(rs, rd, pk_p, pk_s, s_stat, i0, j, L, key_ok, p.causal, sl2, p.scale, thr, ks_drop, col_key);
            else attn_bwd_chunk<true, false>(rs, rd, pk_p, pk_s, s_stat, i0, j, L, key_ok, p.causal, sl2, p.scale, thr, ks_drop, col_key);
          }
        }
        // bf16 operands go into the first half of THIS warp's own (already consumed) 32 columns
        tmem_st16(t_st + lane_base + cgp * 32, pk_p);    // Pd^T over S^T
        tmem_st16(t_dpt + lane_base + cgp * 32, pk_s);   // dS^T over dP^T
        // dS^T also to shared memory as the MN-major A operand of dQ: row = key (K index), 64-query chunks
        uint8_t* dst = sdS + (pair_idx & 1) * 2 * TILE + (qs & 1) * TILE;
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8)
          *reinterpret_cast<uint4*>(dst + sw128_off((uint32_t)row, (uint32_t)(cgp * 4 + g8))) =
              make_uint4(pk_s[g8 * 4], pk_s[g8 * 4 + 1], pk_s[g8 * 4 + 2], pk_s[g8 * 4 + 3]);
        tmem_st_wait();
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
      }
      if (!same_pair(cur, nxt)) ++pair_idx;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_p[buf]);
      RP_TR(4 + 3 * s);
      if (!same_kt(cur, nxt)) {
        // ---- dK / dV of this key tile leave through shared memory + TMA stores.  After the LAST key tile every operand
        //      tile is free, so dK, dV and the dQ tiles each take their own region: one barrier, then all stores go out.
        //      Mid-kernel: dK takes the staging tile, dV the dS buffer of the pair that just finished (its next writer waits).
        mbar_wait(&bar_acc, kt & 1);
        tc_fence_after();
        RP_TR(5 + 3 * s);
        const bool last = nxt < 0;
        uint8_t* dv_tile = last ? sQ : sdS + ((pair_idx - 1) & 1) * 2 * TILE;
        drain_to(t_dk, sStage);
        drain_to(t_dv, dv_tile);
        if (last)
          for (int qt = 0; qt < n_t; ++qt) drain_to(t_dq + qt * HD, sdO + qt * TILE);
        fence_proxy_async();
        tc_fence_before();
        named_bar_sync(1, 256);
        if (leader) {
          tma_store_3d(&tm.dk, sStage, p.dk_c0 + h * HD, kt * 128, b);
          tma_store_3d(&tm.dv, dv_tile, p.dv_c0 + h * HD, kt * 128, b);
          if (last)
            for (int qt = 0; qt < n_t; ++qt) tma_store_3d(&tm.dq, sdO + qt * TILE, p.dq_c0 + h * HD, qt * 128, b);
          tma_store_commit();
          if (last) tma_store_wait_read();
        }
        if (!last) lent_buf = (pair_idx - 1) & 1;
      }
      cur = nxt;
      nxt = nxt >= 0 ? next_step(nxt) : -1;
    }
    RP_TR(28);
  }
  RP_TR(29);
  tc_fence_before();
  __syncthreads();
  RP_TR(30);
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

#ifdef RP_ATTN_TRACE
RP_API int rp_debug_attn_trace(unsigned long long* host_out, int n_words) {
  RP_CUDA_CHECK(cudaDeviceSynchronize());
  RP_CUDA_CHECK(cudaMemcpyFromSymbol(host_out, rp::g_attn_trace, sizeof(unsigned long long) * (size_t)n_words));
  return RP_OK;
}
#endif

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
  AttnBwdMaps tm;
  int rc;
  const uint64_t B = (uint64_t)a->B, L = (uint64_t)a->L;
  if ((uint64_t)a->q_rows < B * L || (uint64_t)a->k_rows < B * L || (uint64_t)a->v_rows < B * L || (uint64_t)a->do_rows < B * L)
    return RP_ESHAPE;
  if ((rc = make_tmap_bf16_seq(&tm.q, a->q, B, L, a->q_cols, a->ldq, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16_seq(&tm.k, a->k, B, L, a->k_cols, a->ldk, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16_seq(&tm.v, a->v, B, L, a->v_cols, a->ldv, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16_seq(&tm.d_o, a->d_out, B, L, a->do_cols, a->ld_do, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16_seq(&tm.o, a->out, B, L, (uint64_t)a->H * 64, a->ldo, 128)) != RP_OK) return rc;
  // the gradient arrays are addressed like their forward counterparts: columns x_c0 + h*64 of rows b*L + i
  if ((rc = make_tmap_bf16_seq(&tm.dq, a->dq, B, L, (uint64_t)a->dq_c0 + (uint64_t)a->H * 64, a->ld_dq, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16_seq(&tm.dk, a->dk, B, L, (uint64_t)a->dk_c0 + (uint64_t)a->H * 64, a->ld_dk, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16_seq(&tm.dv, a->dv, B, L, (uint64_t)a->dv_c0 + (uint64_t)a->H * 64, a->ld_dv, 128)) != RP_OK) return rc;
  const int smem = 13 * 128 * 128 + 1024;
  RP_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  attn_bwd_kernel<<<a->B * a->H, kAbThreads, smem, stream>>>(tm, p);
  RP_LAUNCH_CHECK();
  return RP_OK;
}
