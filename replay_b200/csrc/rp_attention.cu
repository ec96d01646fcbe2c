// rp_attention.cu - padded-sequence multi-head attention for L <= 512 (L > 256: head_dim 64 only) (SASRec causal + key-padding, BERT4Rec key-padding):
// fused forward on tcgen05 (S = Q.K^T in TMEM -> masked softmax in registers -> P bf16 back into TMEM -> O = P.V),
// and the row-wise softmax backward that sits between the batched backward GEMMs (rp_gemm).
//
// Replaces torch.nn.MultiheadAttention's scaled-dot-product core as configured by the reference:
//   replay/nn/sequential/sasrec/transformer.py:36-46,99-106 + replay/nn/mask.py:18-51 (float [B*H,L,L] mask never built)
//   replay/models/nn/sequential/sasrec/model.py:407-414,435 (bool causal mask, pad keys NOT masked)
//   replay/models/nn/sequential/bert4rec/model.py:471,494 (key_padding_mask only)
// Fully masked query rows produce a zero output (torch >= 2.5 safe-softmax semantics of the training path).
#include "rp_host.h"
#include "rp_philox.cuh"
#include "rp_sm100.cuh"

namespace rp {

struct AttnParams {
  int B, H, L, Lp;             // Lp = round_up(L, 64): row pitch / row count of the saved probability buffers
  int causal, mask_pad_keys;
  float scale;                 // 1/sqrt(head_dim)
  const uint8_t* pad_mask;     // [B*L] 1 = real token
  __nv_bfloat16* out;          // [B*L, ldo] ; head h writes columns [h*HD, (h+1)*HD)
  int ldo;
  __nv_bfloat16* p_save;       // [B*H, Lp, Lp] unnormalised exp(s - max) (bf16) or null
  float* inv_sum;              // [B*H, Lp] 1 / row sum (0 for fully masked rows)
  float* m_save;               // [B*H, Lp] row max in exp2 units (max * scale * log2e), for the fused backward; or null
  int q_c0, k_c0, v_c0;        // column offsets of head 0 inside the Q / K / V 2-D arrays
  float drop_p;
  unsigned long long seed, drop_off;
  const unsigned long long* seed_ptr;
};

static constexpr float kLog2eA = 1.4426950408889634f;

static constexpr int kAfThreads = 32 + 8 * 32;  // issuer warp + 8 softmax warps (lane quarter x even / odd 32-key chunks)

// KB = number of 256-key blocks the CTA keeps resident (1: L <= 256, 2: L <= 512).  S [128 x 256*KB] fp32 fills 256*KB
// TMEM columns; P (bf16) is written back over its first half and O accumulates behind it.
// A query row is shared by TWO threads (even / odd 32-key chunks; row max and row sum meet in shared memory): 8 softmax
// warps per CTA, 16 per SM - the kernel is bound by the integer / exp work of the element-wise passes, not by the MMAs, and
// 4 warps per CTA left the ALU pipe half idle (ncu r2c).  Warps whose 32 rows lie beyond L do nothing; 32 x 32 chunks above
// the causal diagonal are zero-filled without loads, exponentials or dropout draws.
template <int HD, int KB>
__global__ void __launch_bounds__(kAfThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  constexpr int HC = HD / 64;                 // 64-wide head-dim chunks
  constexpr int Q_BYTES = HC * 128 * 128;     // [128 x HD]
  constexpr int KV_CHUNK = KB * 256 * 128;    // one 64-wide head-dim chunk of all resident keys
  constexpr int KV_BYTES = HC * KV_CHUNK;     // [256*KB x HD]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + KV_BYTES;
  __shared__ __align__(16) uint32_t s_colkey[256 * KB];   // dropout key of every key position
  __shared__ float s_red[2][2][128];                        // [max | sum][chunk parity][row]
  __shared__ uint64_t bar_load, bar_v, bar_s, bar_p, bar_o;
  __shared__ uint64_t bar_pair[4][2][2];                    // [lane quarter][chunk parity of the loader][iteration parity]
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int bz = b * p.H + h;
  const int L = p.L;
  // Inference: a query tile whose rows are ALL padding (left-padded windows: the first tile of every user with at most L - 128
  // items, ~40 % of the MovieLens-shaped users at L = 200) produces nothing anybody reads - the new path masks pad positions as
  // keys and never queries them, the legacy path zeroes pad rows after the block.  Such a CTA writes zeros and leaves before
  // any load or MMA.  (Training keeps the rows: the backward reads their saved statistics.)
  if (p.pad_mask && p.inv_sum == nullptr && p.p_save == nullptr && p.m_save == nullptr) {
    const int r = q0 + (int)threadIdx.x;
    const int live = (threadIdx.x < 128 && r < L) ? (p.pad_mask[(size_t)b * L + r] != 0) : 0;
    if (!__syncthreads_or(live)) {
      // 128 rows x HD bf16 of this head: 16-byte stores, HD / 8 per row
      constexpr int PER_ROW = HD / 8;
      for (int i = threadIdx.x; i < 128 * PER_ROW; i += blockDim.x) {
        const int rr = q0 + i / PER_ROW;
        if (rr < L)
          *reinterpret_cast<uint4*>(p.out + ((size_t)b * L + rr) * p.ldo + h * HD + (i % PER_ROW) * 8) = make_uint4(0u, 0u, 0u, 0u);
      }
      return;
    }
  }
  int nk = p.causal ? min(L, q0 + 128) : L;        // keys that can be visible to this query tile
  const int nk32 = (nk + 31) & ~31;                // MMA N / K extent (<= 256 per block, KB blocks)
  const int n_boxes = (nk32 + 127) / 128;

  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_v, 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_p, 8);
    mbar_init(&bar_o, 1);
    for (int t = 0; t < 16; ++t) mbar_init(&bar_pair[0][0][0] + t, 1);
    fence_barrier_init();
    const int row0 = b * L;
    mbar_arrive_expect_tx(&bar_load, HC * 128 * 128 + HC * n_boxes * 128 * 128);
    mbar_arrive_expect_tx(&bar_v, HC * n_boxes * 128 * 128);
    for (int c = 0; c < HC; ++c) {
      tma_load_2d(sQ + c * 16384, &tmQ, &bar_load, p.q_c0 + h * HD + c * 64, row0 + q0);
      for (int bx = 0; bx < n_boxes; ++bx)
        tma_load_2d(sK + c * KV_CHUNK + bx * 16384, &tmK, &bar_load, p.k_c0 + h * HD + c * 64, row0 + bx * 128);
    }
    for (int c = 0; c < HC; ++c)
      for (int bx = 0; bx < n_boxes; ++bx)
        tma_load_2d(sV + c * KV_CHUNK + bx * 16384, &tmV, &bar_v, p.v_c0 + h * HD + c * 64, row0 + bx * 128);
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(&tmem_slot, 256 * KB);
  } else if (p.drop_p > 0.f) {
    for (int j = threadIdx.x - 32; j < 256 * KB; j += 256) s_colkey[j] = drop_col_key((uint32_t)j);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tmem_o = tmem + 128 * KB;   // reuses the S columns behind the packed P once P is complete

  if (warp == 0) {
    if (elect_one()) {
      mbar_wait(&bar_load, 0);
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {           // one MMA N extent is at most 256 keys
        const int nb = min(256, nk32 - kb * 256);
        if (nb <= 0) break;
        const uint32_t idesc1 = umma_idesc_bf16(128, nb);
#pragma unroll
        for (int c = 0; c < HC; ++c)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ss(tmem + kb * 256, umma_desc_sw128(smem_u32(sQ) + c * 16384 + ks * 32, 16, 1024),
                    umma_desc_sw128(smem_u32(sK) + c * KV_CHUNK + kb * 32768 + ks * 32, 16, 1024), idesc1, (c | ks) != 0);
      }
      umma_commit(&bar_s);
      // ---- second GEMM once the softmax warps have written P
      mbar_wait(&bar_v, 0);
      mbar_wait(&bar_p, 0);
      tc_fence_after();
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, HD, false, true);
      for (int ks = 0; ks < nk32 / 16; ++ks)
        umma_ts(tmem_o, tmem + ks * 8, umma_desc_sw128(smem_u32(sV) + ks * 2048, KV_CHUNK, 1024), idesc2, ks != 0);
      umma_commit(&bar_o);
    }
  } else {
    // ------------------------------------------------ softmax warps: thread = (query row, chunk parity)
    const int quarter = warp & 3, par = (warp - 1) >> 2;
    const int row = quarter * 32 + lane;
    const int i = q0 + row;                 // query position inside the sequence
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const bool warp_live = q0 + quarter * 32 < L;        // some row of this warp is a real query
    const int n_chunks = nk32 / 32;
    // chunks at or below the causal diagonal of this warp's LAST row; the rest of the tile's key extent is masked for
    // every row of the warp
    const int n_vis = p.causal ? min(n_chunks, (q0 + quarter * 32 + 31) / 32 + 1) : n_chunks;
    // key-visibility bit masks of this thread's chunks (chunk c = 2 t + par), 32 keys per word
    uint32_t kmask[4 * KB];
    {
      uint8_t pm[4 * KB];   // all loads first: a ballot per load serialises their latencies
#pragma unroll
      for (int t = 0; t < 4 * KB; ++t) {
        const int j = (2 * t + par) * 32 + lane;
        pm[t] = (j < L && p.mask_pad_keys) ? p.pad_mask[(size_t)b * L + j] : (uint8_t)(j < L);
      }
#pragma unroll
      for (int t = 0; t < 4 * KB; ++t) kmask[t] = __ballot_sync(0xffffffffu, pm[t] != 0);
    }
    const float sl2 = p.scale * kLog2eA;
    const bool drop = p.drop_p > 0.f;
    const uint32_t thr = drop ? (uint32_t)(p.drop_p * 4294967296.0) : 0u;
    const float ks_drop = drop ? 1.f / (1.f - p.drop_p) : 1.f;
    const unsigned long long seed_eff = p.seed + ((drop && p.seed_ptr) ? *p.seed_ptr : 0ull);
    const uint32_t row_key = drop ? drop_row_key(seed_eff, p.drop_off, (unsigned long long)bz * p.Lp + (unsigned long long)i) : 0u;
    auto vis_mask = [&](int c) -> uint32_t {
      uint32_t m = 0;
#pragma unroll
      for (int t = 0; t < 4 * KB; ++t)
        if (2 * t + par == c) m = kmask[t];
      if (p.causal) {
        const int rel = i - c * 32;
        const uint32_t cm = rel >= 31 ? 0xffffffffu : (rel < 0 ? 0u : ((2u << rel) - 1u));
        m &= cm;
      }
      return m;
    };
    mbar_wait_relaxed(&bar_s, 0);
    tc_fence_after();
    // pass 1: row max over the visible keys of this thread's chunks
    float mx = -INFINITY;
    if (warp_live) {
      for (int c = par; c < n_vis; c += 2) {
        const uint32_t vm = vis_mask(c);
        if (__all_sync(0xffffffffu, vm == 0u)) continue;
        uint32_t raw[32];
        tmem_ld32(tmem + lane_base + c * 32, raw);
        tmem_ld_wait();
        if (vm == 0xffffffffu) {  // chunk entirely visible (the common case below the diagonal): no per-element selects
          float m0 = __uint_as_float(raw[0]), m1 = __uint_as_float(raw[1]), m2 = __uint_as_float(raw[2]), m3 = __uint_as_float(raw[3]);
#pragma unroll
          for (int q = 4; q < 32; q += 4) {
            m0 = fmaxf(m0, __uint_as_float(raw[q]));
            m1 = fmaxf(m1, __uint_as_float(raw[q + 1]));
            m2 = fmaxf(m2, __uint_as_float(raw[q + 2]));
            m3 = fmaxf(m3, __uint_as_float(raw[q + 3]));
          }
          mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
        } else if (vm != 0u) {
#pragma unroll
          for (int q = 0; q < 32; ++q)
            if (vm & (1u << q)) mx = fmaxf(mx, __uint_as_float(raw[q]));
        }
      }
    }
    s_red[0][par][row] = mx;
    named_bar_sync(1 + quarter, 64);   // the two warps of this lane quarter
    mx = fmaxf(mx, s_red[0][par ^ 1][row]);
    const float moff = (mx == -INFINITY) ? 0.f : mx * sl2;
    // pass 2: exponentials, row sum, P (bf16) back into TMEM over S, optional copy of the un-dropped P to global
    float sum = 0.f;
    __nv_bfloat16* prow =
        (p.p_save && i < L) ? p.p_save + ((size_t)bz * p.Lp + i) * p.Lp : nullptr;
    // The packed P chunk c lands on columns [16 c, 16 c + 16) = inside S chunk c / 2, which belongs to this row's OTHER
    // thread for every second chunk.  S chunk t is loaded in iteration t / 2 and stored over in iteration t, so each warp
    // announces "my loads of iteration t are done" (mbarrier, before its exponentials) and checks the partner warp's
    // announcement of the same iteration right before its store - by then it has long been made.  Two mbarriers per
    // direction alternate by iteration: a warp can be at most one announcement ahead of its partner's check.
    for (int t = 0; 2 * t < n_chunks; ++t) {
      const int c = 2 * t + par;
      const bool active = warp_live && c < n_chunks;
      const uint32_t vm = (active && c < n_vis) ? vis_mask(c) : 0u;
      const bool blank = !active || __all_sync(0xffffffffu, vm == 0u);  // masked for the whole warp: nothing loaded or drawn
      uint32_t raw[32];
      if (!blank) {
        tmem_ld32(tmem + lane_base + c * 32, raw);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_pair[quarter][par][t & 1]);
      if (!active) continue;
      uint32_t pk[16];
      if (blank) {
#pragma unroll
        for (int q = 0; q < 16; ++q) pk[q] = 0u;
        if (prow) {
#pragma unroll
          for (int q = 0; q < 32; q += 8) *reinterpret_cast<uint4*>(prow + c * 32 + q) = make_uint4(0u, 0u, 0u, 0u);
        }
      } else {
        float e[32];
        if (vm == 0xffffffffu) {
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            e[q] = ex2f(fmaf(__uint_as_float(raw[q]), sl2, -moff));
            sum += e[q];
          }
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const float v = ex2f(fmaf(__uint_as_float(raw[q]), sl2, -moff));
            e[q] = (vm & (1u << q)) ? v : 0.f;
            sum += e[q];
          }
        }
        if (prow) {
#pragma unroll
          for (int q = 0; q < 32; q += 8) {
            uint4 w;
            w.x = pack_bf16(e[q], e[q + 1]);
            w.y = pack_bf16(e[q + 2], e[q + 3]);
            w.z = pack_bf16(e[q + 4], e[q + 5]);
            w.w = pack_bf16(e[q + 6], e[q + 7]);
            *reinterpret_cast<uint4*>(prow + c * 32 + q) = w;
          }
        }
        if (drop) {
          const uint4* ck = reinterpret_cast<const uint4*>(s_colkey + c * 32);
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            const uint4 k4 = ck[q >> 2];
            e[q] = drop_mix(row_key, k4.x) >= thr ? e[q] * ks_drop : 0.f;
            e[q + 1] = drop_mix(row_key, k4.y) >= thr ? e[q + 1] * ks_drop : 0.f;
            e[q + 2] = drop_mix(row_key, k4.z) >= thr ? e[q + 2] * ks_drop : 0.f;
            e[q + 3] = drop_mix(row_key, k4.w) >= thr ? e[q + 3] * ks_drop : 0.f;
          }
        }
#pragma unroll
        for (int q = 0; q < 32; q += 2) pk[q >> 1] = pack_bf16(e[q], e[q + 1]);
      }
      mbar_wait(&bar_pair[quarter][par ^ 1][t & 1], (t >> 1) & 1);
      tc_fence_after();
      tmem_st16(tmem + lane_base + c * 16, pk);
    }
    if (warp_live) tmem_st_wait();
    s_red[1][par][row] = sum;
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_p);
    named_bar_sync(1 + quarter, 64);   // the two warps of this lane quarter
    sum += s_red[1][par ^ 1][row];
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    if (par == 0 && i < L) {
      if (p.inv_sum) p.inv_sum[(size_t)bz * p.Lp + i] = inv;
      if (p.m_save) p.m_save[(size_t)bz * p.Lp + i] = moff;
    }
    // ---- O = (P.V) / sum : chunk parity 0 stores the even 32-column groups of the head, parity 1 the odd ones
    mbar_wait_relaxed(&bar_o, 0);
    tc_fence_after();
    if (warp_live) {
#pragma unroll
      for (int c = par * 32; c < HD; c += 64) {
        uint32_t raw[32];
        tmem_ld32(tmem_o + lane_base + c, raw);
        tmem_ld_wait();
        if (i < L) {
          __nv_bfloat16* o = p.out + ((size_t)b * L + i) * p.ldo + h * HD + c;
#pragma unroll
          for (int q = 0; q < 32; q += 8) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(raw[q]) * inv, __uint_as_float(raw[q + 1]) * inv);
            w.y = pack_bf16(__uint_as_float(raw[q + 2]) * inv, __uint_as_float(raw[q + 3]) * inv);
            w.z = pack_bf16(__uint_as_float(raw[q + 4]) * inv, __uint_as_float(raw[q + 5]) * inv);
            w.w = pack_bf16(__uint_as_float(raw[q + 6]) * inv, __uint_as_float(raw[q + 7]) * inv);
            *reinterpret_cast<uint4*>(o + q) = w;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256 * KB);
}

// Row-wise softmax backward between the batched GEMMs.  One warp per (batch*head, query) row.
//   in : p_save  = exp(s - max) (bf16), inv_sum, dpd = dO.V^T (bf16, w.r.t. the dropped & rescaled probabilities)
//   out: ds (over dpd) = P * (dP - sum_j P_j dP_j) * scale      with P = p_save * inv_sum, dP = dpd * mask / keep
//        pd (over p_save) = P * mask / keep                       (A operand of dV = Pd^T . dO)
template <int NK>  // 128-column blocks per row: Lp <= 128 * NK
__global__ void attn_softmax_bwd_kernel(__nv_bfloat16* __restrict__ p_save, __nv_bfloat16* __restrict__ dpd,
                                        const float* __restrict__ inv_sum, int BH, int L, int Lp, float scale,
                                        float drop_p, unsigned long long seed, unsigned long long drop_off,
                                        const unsigned long long* __restrict__ seed_ptr) {
  if (drop_p > 0.f && seed_ptr) seed += *seed_ptr;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const uint32_t thr = drop_p > 0.f ? (uint32_t)(drop_p * 4294967296.0) : 0u;
  const float ksd = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const long long n_rows = (long long)BH * L;
  uint32_t colkey[4 * NK];  // the lane's columns are the same in every row
#pragma unroll
  for (int k = 0; k < NK; ++k)
#pragma unroll
    for (int q = 0; q < 4; ++q) colkey[k * 4 + q] = drop_col_key((uint32_t)(k * 128 + lane * 4 + q));
  for (long long r = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); r < n_rows; r += (long long)gridDim.x * wpb) {
    const int bz = (int)(r / L), i = (int)(r % L);
    const size_t base = ((size_t)bz * Lp + i) * Lp;
    const uint32_t row_key = drop_row_key(seed, drop_off, (unsigned long long)bz * Lp + (unsigned long long)i);
    const float inv = inv_sum[(size_t)bz * Lp + i];
    // lane owns columns [4*lane + 128*k, +4), k < NK   (columns >= L hold zeros and are never written)
    float P[4 * NK], dP[4 * NK], keep[4 * NK];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int j0 = k * 128 + lane * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) { P[k * 4 + q] = 0.f; dP[k * 4 + q] = 0.f; keep[k * 4 + q] = 0.f; }
      if (j0 < L) {
        const uint2 pr = *reinterpret_cast<const uint2*>(p_save + base + j0);
        const uint2 dr = *reinterpret_cast<const uint2*>(dpd + base + j0);
        const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(&pr);
        const __nv_bfloat162* dh = reinterpret_cast<const __nv_bfloat162*>(&dr);
        const float2 p0 = __bfloat1622float2(ph[0]), p1 = __bfloat1622float2(ph[1]);
        const float2 d0 = __bfloat1622float2(dh[0]), d1 = __bfloat1622float2(dh[1]);
        const float pv[4] = {p0.x, p0.y, p1.x, p1.y}, dv[4] = {d0.x, d0.y, d1.x, d1.y};
        uint32_t rv[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        if (drop_p > 0.f) {
#pragma unroll
          for (int q = 0; q < 4; ++q) rv[q] = drop_mix(row_key, colkey[k * 4 + q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const bool in = (j0 + q) < L;
          keep[k * 4 + q] = (in && rv[q] >= thr) ? ksd : 0.f;
          P[k * 4 + q] = in ? pv[q] * inv : 0.f;
          dP[k * 4 + q] = dv[q] * keep[k * 4 + q];
          dot += P[k * 4 + q] * dP[k * 4 + q];
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int j0 = k * 128 + lane * 4;
      if (j0 < L) {
        float ds[4], pd[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          ds[q] = P[k * 4 + q] * (dP[k * 4 + q] - dot) * scale;
          pd[q] = P[k * 4 + q] * keep[k * 4 + q];
        }
        uint2 w0, w1;
        w0.x = pack_bf16(ds[0], ds[1]); w0.y = pack_bf16(ds[2], ds[3]);
        w1.x = pack_bf16(pd[0], pd[1]); w1.y = pack_bf16(pd[2], pd[3]);
        *reinterpret_cast<uint2*>(dpd + base + j0) = w0;
        *reinterpret_cast<uint2*>(p_save + base + j0) = w1;
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// predict(): only the LAST position of every sequence is scored, so in the last transformer block a single query row per
// (sequence, head) attends to the keys - a memory-bound GEMV pair, one warp per (b, h).
//   q    bf16 [B, H*HD]   (compact: one row per sequence)
//   k, v bf16 rows b*L + j of 2-D arrays with pitches ldk / ldv, head h at columns x_c0 + h*HD
//   out  bf16 [B, H*HD]
// Visible keys: j < L (the query is the last position, so causal masking is a no-op) and, if mask_pad_keys, pad[b*L+j].
// ------------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void attn_last_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                 const __nv_bfloat16* __restrict__ v, long long ldk, long long ldv, int k_c0, int v_c0,
                                 const uint8_t* __restrict__ pad_mask, int B, int H, int L, int mask_pad_keys, float scale,
                                 __nv_bfloat16* __restrict__ out) {
  constexpr int PER = HD / 32;  // output columns per lane
  __shared__ float s_q[8][HD];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int bh = blockIdx.x * (blockDim.x >> 5) + w;
  if (bh >= B * H) return;
  const int b = bh / H, h = bh % H;
  for (int c = lane; c < HD; c += 32) s_q[w][c] = __bfloat162float(q[(size_t)b * H * HD + h * HD + c]) * scale;
  __syncwarp();
  // scores: lane owns keys lane, lane+32, ...
  float sc[16];  // L <= 512
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int j = i * 32 + lane;
    sc[i] = -INFINITY;
    if (j < L && (!mask_pad_keys || pad_mask[(size_t)b * L + j])) {
      const uint4* kr = reinterpret_cast<const uint4*>(k + ((size_t)b * L + j) * ldk + k_c0 + h * HD);
      float acc = 0.f;
#pragma unroll
      for (int c8 = 0; c8 < HD / 8; ++c8) {
        const uint4 raw = kr[c8];
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __bfloat1622float2(h2[t]);
          acc = fmaf(f.x, s_q[w][c8 * 8 + 2 * t], fmaf(f.y, s_q[w][c8 * 8 + 2 * t + 1], acc));
        }
      }
      sc[i] = acc;
      mx = fmaxf(mx, acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    sc[i] = (sc[i] == -INFINITY) ? 0.f : __expf(sc[i] - mx);
    sum += sc[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = sum > 0.f ? 1.f / sum : 0.f;  // no visible key -> zero output (safe-softmax semantics)
  // P.V: every lane accumulates the value rows of ITS keys (independent 16-byte loads, nothing serialised on a broadcast of
  // p_j), then a reduce-scatter butterfly over the lanes leaves HD/32 finished output columns per lane
  float o[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) o[c] = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int j = i * 32 + lane;
    if (j < L && sc[i] != 0.f) {
      const uint4* vr = reinterpret_cast<const uint4*>(v + ((size_t)b * L + j) * ldv + v_c0 + h * HD);
      const float pj = sc[i];
#pragma unroll
      for (int c8 = 0; c8 < HD / 8; ++c8) {
        const uint4 raw = vr[c8];
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __bfloat1622float2(h2[t]);
          o[c8 * 8 + 2 * t] = fmaf(pj, f.x, o[c8 * 8 + 2 * t]);
          o[c8 * 8 + 2 * t + 1] = fmaf(pj, f.y, o[c8 * 8 + 2 * t + 1]);
        }
      }
    }
  }
  int base = 0;
#define RP_RS_STEP(OFF, N)                                                     \
  {                                                                            \
    const bool up = (lane & OFF) != 0;                                         \
    _Pragma("unroll") for (int c = 0; c < N; ++c) {                            \
      const float send = up ? o[c] : o[c + N];                                 \
      const float recv = __shfl_xor_sync(0xffffffffu, send, OFF);              \
      o[c] = (up ? o[c + N] : o[c]) + recv;                                    \
    }                                                                          \
    base += up ? N : 0;                                                        \
  }
  RP_RS_STEP(16, HD / 2)
  RP_RS_STEP(8, HD / 4)
  RP_RS_STEP(4, HD / 8)
  RP_RS_STEP(2, HD / 16)
  RP_RS_STEP(1, HD / 32)
#undef RP_RS_STEP
  __nv_bfloat16* op = out + (size_t)b * H * HD + h * HD + base;
#pragma unroll
  for (int c = 0; c < PER; c += 2) *reinterpret_cast<uint32_t*>(op + c) = pack_bf16(o[c] * inv, o[c + 1] * inv);
}

}  // namespace rp

using namespace rp;

RP_API int rp_attn_last(const void* q, const void* k, const void* v, long long ldk, long long ldv, int k_c0, int v_c0,
                        const uint8_t* pad_mask, int B, int H, int L, int head_dim, int mask_pad_keys, void* out,
                        float scale_in, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!q || !k || !v || !pad_mask || !out) return RP_EINVAL;
  if (B <= 0 || H <= 0 || L <= 0 || L > 512) return RP_ESHAPE;
  if ((ldk & 7) || (ldv & 7) || (k_c0 & 7) || (v_c0 & 7)) return RP_EALIGN;
  const float scale = scale_in > 0.f ? scale_in : 1.f / sqrtf((float)head_dim);
  const int blocks = (B * H + 7) / 8;
  const __nv_bfloat16 *qq = reinterpret_cast<const __nv_bfloat16*>(q), *kk = reinterpret_cast<const __nv_bfloat16*>(k),
                      *vv = reinterpret_cast<const __nv_bfloat16*>(v);
  if (head_dim == 64)
    attn_last_kernel<64><<<blocks, 256, 0, stream>>>(qq, kk, vv, ldk, ldv, k_c0, v_c0, pad_mask, B, H, L, mask_pad_keys, scale,
                                                     reinterpret_cast<__nv_bfloat16*>(out));
  else if (head_dim == 128)
    attn_last_kernel<128><<<blocks, 256, 0, stream>>>(qq, kk, vv, ldk, ldv, k_c0, v_c0, pad_mask, B, H, L, mask_pad_keys, scale,
                                                      reinterpret_cast<__nv_bfloat16*>(out));
  else
    return RP_ESHAPE;
  RP_LAUNCH_CHECK();
  return RP_OK;
}

struct rp_attn_desc {
  const void* q; long long q_rows, q_cols, ldq; int q_c0;
  const void* k; long long k_rows, k_cols, ldk; int k_c0;
  const void* v; long long v_rows, v_cols, ldv; int v_c0;
  int B, H, L, head_dim;
  int causal, mask_pad_keys;
  const uint8_t* pad_mask;
  void* out; int ldo;
  void* p_save; float* inv_sum;
  float drop_p; unsigned long long seed, drop_off; const unsigned long long* seed_ptr;
  float* m_save;
  float scale;
};

RP_API int rp_attn_fwd(const rp_attn_desc* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->q || !a->k || !a->v || !a->out || !a->pad_mask) return RP_EINVAL;
  if (a->L <= 0 || a->L > 512 || a->B <= 0 || a->H <= 0) return RP_ESHAPE;
  if (a->head_dim != 64 && a->head_dim != 128) return RP_ESHAPE;
  if (a->L > 256 && a->head_dim != 64) return RP_ESHAPE;  // 512 resident keys x 128 head dims do not fit shared memory
  if (a->ldo % 8 != 0) return RP_EALIGN;
  AttnParams p;
  p.B = a->B; p.H = a->H; p.L = a->L; p.Lp = (a->L + 63) & ~63;
  p.causal = a->causal; p.mask_pad_keys = a->mask_pad_keys;
  p.scale = a->scale > 0.f ? a->scale : 1.f / sqrtf((float)a->head_dim);
  p.pad_mask = a->pad_mask;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out); p.ldo = a->ldo;
  p.p_save = reinterpret_cast<__nv_bfloat16*>(a->p_save); p.inv_sum = a->inv_sum; p.m_save = a->m_save;
  p.q_c0 = a->q_c0; p.k_c0 = a->k_c0; p.v_c0 = a->v_c0;
  p.drop_p = a->drop_p; p.seed = a->seed; p.drop_off = a->drop_off; p.seed_ptr = a->seed_ptr;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_tmap_bf16(&tmQ, a->q, a->q_rows, a->q_cols, a->ldq, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmK, a->k, a->k_rows, a->k_cols, a->ldk, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmV, a->v, a->v_rows, a->v_cols, a->ldv, 128)) != RP_OK) return rc;
  dim3 grid((a->L + 127) / 128, a->H, a->B);
  if (a->L > 256) {
    const int smem = 16384 + 2 * 65536 + 1024;
    RP_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attn_fwd_kernel<64, 2><<<grid, kAfThreads, smem, stream>>>(tmQ, tmK, tmV, p);
  } else if (a->head_dim == 64) {
    const int smem = 16384 + 2 * 32768 + 1024;
    RP_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attn_fwd_kernel<64, 1><<<grid, kAfThreads, smem, stream>>>(tmQ, tmK, tmV, p);
  } else {
    const int smem = 2 * (16384 + 2 * 32768) + 1024;
    RP_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attn_fwd_kernel<128, 1><<<grid, kAfThreads, smem, stream>>>(tmQ, tmK, tmV, p);
  }
  RP_LAUNCH_CHECK();
  return RP_OK;
}

RP_API int rp_attn_softmax_bwd(void* p_save, void* dpd, const float* inv_sum, int BH, int L, float scale, float drop_p,
                               unsigned long long seed, unsigned long long drop_off, const unsigned long long* seed_ptr,
                               void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!p_save || !dpd || !inv_sum || BH <= 0 || L <= 0 || L > 512) return RP_EINVAL;
  const int Lp = (L + 63) & ~63;
  const long long rows = (long long)BH * L;
  long long blocks = (rows + 7) / 8;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (L > 256)
    attn_softmax_bwd_kernel<4><<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<__nv_bfloat16*>(p_save),
                                                                reinterpret_cast<__nv_bfloat16*>(dpd), inv_sum, BH, L, Lp, scale,
                                                                drop_p, seed, drop_off, seed_ptr);
  else
    attn_softmax_bwd_kernel<2><<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<__nv_bfloat16*>(p_save),
                                                                reinterpret_cast<__nv_bfloat16*>(dpd), inv_sum, BH, L, Lp, scale,
                                                                drop_p, seed, drop_off, seed_ptr);
  RP_LAUNCH_CHECK();
  return RP_OK;
}
