// rp_elementwise.cu - the HBM-bound kernels of the transformer body and of the optimizer: batch preparation / target
// compaction, embedding gather + positional add (+dropout) and its backward, LayerNorm forward/backward (optionally
// gathering / scattering the valid-target rows), dropout backward, bias-gradient column sums, Adam.
// All are coalesced, vectorised (8/16-byte accesses) row-per-warp kernels sized in multiples of the SM count.
//
// Reference call sites: replay/nn/sequential/sasrec/agg.py:37-53, replay/models/nn/sequential/sasrec/model.py:346-357
// (embedding), transformer.py:47-49,60-62 + model.py:415-417,463 (LayerNorm eps 1e-8 / 1e-5),
// replay/models/nn/optimizer_utils/optimizer_factory.py:71-87 (torch.optim.Adam).
#include "rp_host.h"
#include "rp_philox.cuh"
#include "rp_sm100.cuh"

namespace rp {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------------------------
// batch preparation: int64 ids / bool masks from the data layer -> int32 ids (pads replaced by pad_id), compacted list of
// valid targets (ascending token index), their labels, and the count (device memory).
// Single block: T is at most a few hundred thousand tokens.
// ------------------------------------------------------------------------------------------------------------------
// pass 1 (many blocks): ids conversion + number of valid targets per block of 1024 tokens
__global__ void __launch_bounds__(1024, 1)
prepare_count_kernel(const int64_t* __restrict__ ids64, const uint8_t* __restrict__ pad_mask,
                     const int64_t* __restrict__ labels64, const uint8_t* __restrict__ target_mask, int T, int pad_id,
                     int n_items, int32_t* __restrict__ ids32, int32_t* __restrict__ block_counts) {
  __shared__ int warp_tot[32];
  const int t = blockIdx.x * 1024 + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool valid = false;
  if (t < T) {
    const int64_t id = ids64[t];
    ids32[t] = (pad_mask[t] && id >= 0 && id < n_items) ? (int32_t)id : pad_id;
    if (target_mask) {
      const int64_t y = labels64[t];
      valid = target_mask[t] != 0 && y >= 0 && y < n_items;
    }
  }
  if (!target_mask) return;
  const unsigned bal = __ballot_sync(0xffffffffu, valid);
  if (lane == 0) warp_tot[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (lane == 0) block_counts[blockIdx.x] = w;
  }
}

// pass 2 (same grid): every block sums the counts of the blocks before it (<= a few hundred) and writes its survivors
__global__ void __launch_bounds__(1024, 1)
prepare_write_kernel(const int64_t* __restrict__ labels64, const uint8_t* __restrict__ target_mask, int T, int n_items,
                     const int32_t* __restrict__ block_counts, int32_t* __restrict__ valid_idx,
                     int32_t* __restrict__ labels_c, int32_t* __restrict__ n_valid) {
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  const int t = blockIdx.x * 1024 + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int part = 0;
  for (int b = threadIdx.x; b < (int)blockIdx.x; b += 1024) part += block_counts[b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) warp_tot[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s0 = 0;
    for (int w = 0; w < 32; ++w) s0 += warp_tot[w];
    base_s = s0;
  }
  __syncthreads();
  bool valid = false;
  int64_t y = 0;
  if (t < T) {
    y = labels64[t];
    valid = target_mask[t] != 0 && y >= 0 && y < n_items;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, valid);
  const int pre = __popc(bal & ((1u << lane) - 1));
  __syncthreads();
  if (lane == 0) warp_tot[warp] = __popc(bal);
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += warp_tot[w];
  const int pos = base_s + woff + pre;
  if (valid) {
    valid_idx[pos] = t;
    labels_c[pos] = (int32_t)y;
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 1023) *n_valid = pos + (valid ? 1 : 0);
}

// ------------------------------------------------------------------------------------------------------------------
// embedding:  x[t] = E[ids[t]] * scale + P[pos0 + t % L]  -> dropout -> (* pad)        one warp per token
// ------------------------------------------------------------------------------------------------------------------
template <int VEC /* bf16 elements per lane = d / 32 */>
__global__ void embed_fwd_kernel(const __nv_bfloat16* __restrict__ table, const float* __restrict__ pos,
                                 const int32_t* __restrict__ ids, const uint8_t* __restrict__ pad_mask, int T, int L,
                                 int pos0, float scale, int zero_pad_rows, float drop_p, unsigned long long seed,
                                 unsigned long long drop_off, const unsigned long long* __restrict__ seed_ptr,
                                 const uint8_t* __restrict__ tok_mask, const __nv_bfloat16* __restrict__ mask_emb,
                                 __nv_bfloat16* __restrict__ out) {
  if (drop_p > 0.f && seed_ptr) seed += *seed_ptr;
  constexpr int D = VEC * 32;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const uint32_t thr = drop_p > 0.f ? (uint32_t)(drop_p * 4294967296.0) : 0u;
  const float ks = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  uint32_t ck[VEC];  // dropout column keys of this lane's columns (rp_philox.cuh)
#pragma unroll
  for (int i = 0; i < VEC; ++i) ck[i] = drop_col_key((uint32_t)(lane * VEC + i));
  // TOK tokens per warp and iteration: the id loads, then the row loads of all of them are in flight together (one token at a
  // time the loop was a chain of two dependent DRAM round trips per token, 43 us for 102 400 tokens at d = 128).  The TOK
  // tokens sit at the SAME position of TOK consecutive sequences, so the fp32 position row (2 x the bytes of the bf16 item
  // row) is fetched once per iteration instead of once per token: the predict body's 819 200 tokens moved 420 MB of position
  // rows through L2 next to 210 MB of item rows and 210 MB of output (102 us, L2-bound).
  constexpr int TOK = VEC <= 4 ? 8 : 4;
  const int n_seq = T / L;                      // T is a multiple of L (whole sequences)
  const int n_grp = (n_seq + TOK - 1) / TOK;
  const long long n_work = (long long)n_grp * L;
  for (long long u = blockIdx.x * wpb + (threadIdx.x >> 5); u < n_work; u += (long long)gridDim.x * wpb) {
    const int pidx = (int)(u % L), s0 = (int)(u / L) * TOK;
    int id[TOK];
    bool use_mask[TOK];
#pragma unroll
    for (int k = 0; k < TOK; ++k) {
      const int t = min(s0 + k, n_seq - 1) * L + pidx;
      id[k] = ids[t];
      // BERT4Rec: positions with token_mask == 0 (<MASK> and pads) take the single mask embedding (bert4rec/model.py:285-288)
      use_mask[k] = tok_mask && !tok_mask[t];
    }
    __nv_bfloat162 ev[TOK][VEC / 2];
#pragma unroll
    for (int k = 0; k < TOK; ++k) {
      const __nv_bfloat16* e = use_mask[k] ? mask_emb + lane * VEC : table + (size_t)id[k] * D + lane * VEC;
#pragma unroll
      for (int i = 0; i < VEC; i += 2) ev[k][i >> 1] = *reinterpret_cast<const __nv_bfloat162*>(e + i);
    }
    float pv[VEC];
    {
      const float* p = pos + (size_t)(pos0 + pidx) * D + lane * VEC;
#pragma unroll
      for (int i = 0; i < VEC; ++i) pv[i] = p[i];
    }
#pragma unroll
    for (int k = 0; k < TOK; ++k) {
      if (s0 + k >= n_seq) break;
      const int t = (s0 + k) * L + pidx;
      float v[VEC];
#pragma unroll
      for (int i = 0; i < VEC; i += 2) {
        const float2 f = __bfloat1622float2(ev[k][i >> 1]);
        v[i] = f.x * scale + pv[i];
        v[i + 1] = f.y * scale + pv[i + 1];
      }
      if (drop_p > 0.f) {
        const uint32_t rk = drop_row_key(seed, drop_off, (unsigned long long)t);
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = drop_mix(rk, ck[i]) >= thr ? v[i] * ks : 0.f;
      }
      if (zero_pad_rows && !pad_mask[t]) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = 0.f;
      }
      __nv_bfloat16* o = out + (size_t)t * D + lane * VEC;
#pragma unroll
      for (int i = 0; i < VEC; i += 2) *reinterpret_cast<uint32_t*>(o + i) = pack_bf16(v[i], v[i + 1]);
    }
  }
}

// backward of the gather: dE[ids[t]] += dx[t] * scale * mask   (fp32 atomics, pad row frozen)
template <int VEC>
__global__ void embed_bwd_table_kernel(const __nv_bfloat16* __restrict__ dx, const int32_t* __restrict__ ids,
                                       const uint8_t* __restrict__ pad_mask, int T, int pad_id, float scale,
                                       int zero_pad_rows, float drop_p, unsigned long long seed,
                                       unsigned long long drop_off, const unsigned long long* __restrict__ seed_ptr,
                                       const uint8_t* __restrict__ tok_mask, float* __restrict__ d_mask_emb,
                                       float* __restrict__ dE) {
  if (drop_p > 0.f && seed_ptr) seed += *seed_ptr;
  constexpr int D = VEC * 32;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const uint32_t thr = drop_p > 0.f ? (uint32_t)(drop_p * 4294967296.0) : 0u;
  const float ks = (drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f) * scale;
  uint32_t ck[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) ck[i] = drop_col_key((uint32_t)(lane * VEC + i));
  constexpr int TOK = 4;  // tokens per warp and iteration: independent loads in flight, vector reductions (red.v4.f32)
  for (int tb = (blockIdx.x * wpb + (threadIdx.x >> 5)) * TOK; tb < T; tb += gridDim.x * wpb * TOK) {
    int id[TOK];
    bool on[TOK];
    __nv_bfloat162 gv[TOK][VEC / 2];
#pragma unroll
    for (int k = 0; k < TOK; ++k) {
      const int t = min(tb + k, T - 1);
      id[k] = ids[t];
      on[k] = tb + k < T && id[k] != pad_id && !((zero_pad_rows || tok_mask) && !pad_mask[t]);  // pads receive no gradient
    }
#pragma unroll
    for (int k = 0; k < TOK; ++k) {
      const __nv_bfloat16* g = dx + (size_t)min(tb + k, T - 1) * D + lane * VEC;
#pragma unroll
      for (int i = 0; i < VEC; i += 2) gv[k][i >> 1] = *reinterpret_cast<const __nv_bfloat162*>(g + i);
    }
#pragma unroll
    for (int k = 0; k < TOK; ++k) {
      if (!on[k]) continue;
      const int t = tb + k;
      float* dst = (tok_mask && !tok_mask[t]) ? d_mask_emb + lane * VEC : dE + (size_t)id[k] * D + lane * VEC;
      float v[VEC];
#pragma unroll
      for (int i = 0; i < VEC; i += 2) {
        const float2 f = __bfloat1622float2(gv[k][i >> 1]);
        v[i] = f.x * ks;
        v[i + 1] = f.y * ks;
      }
      if (drop_p > 0.f) {
        const uint32_t rk = drop_row_key(seed, drop_off, (unsigned long long)t);
#pragma unroll
        for (int i = 0; i < VEC; ++i)
          if (drop_mix(rk, ck[i]) < thr) v[i] = 0.f;
      }
      if constexpr (VEC % 4 == 0) {
#pragma unroll
        for (int i = 0; i < VEC; i += 4) atomicAdd(reinterpret_cast<float4*>(dst + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
      } else {
#pragma unroll
        for (int i = 0; i < VEC; i += 2) atomicAdd(reinterpret_cast<float2*>(dst + i), make_float2(v[i], v[i + 1]));
      }
    }
  }
}

// dP[pos0 + l] += sum_b dx[b*L + l] * mask.   grid = (L, G): block (l, g) sums a slab of the batch; thread = 4 columns x
// one batch lane; smem reduce over the batch lanes, then one atomic per column per block.
__global__ void embed_bwd_pos_kernel(const __nv_bfloat16* __restrict__ dx, const uint8_t* __restrict__ pad_mask, int B,
                                     int L, int D, int pos0, int zero_pad_rows, float drop_p, unsigned long long seed,
                                     unsigned long long drop_off, const unsigned long long* __restrict__ seed_ptr,
                                     float* __restrict__ dP) {
  if (drop_p > 0.f && seed_ptr) seed += *seed_ptr;
  extern __shared__ float red4[];  // [rows_per_iter][D]
  const int l = blockIdx.x;
  const int tpr = D / 4;                       // threads per row
  const int rlanes = blockDim.x / tpr;         // batch lanes
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  const int per = (B + gridDim.y - 1) / gridDim.y;
  const int b0 = blockIdx.y * per, b1 = min(B, b0 + per);
  const uint32_t thr = drop_p > 0.f ? (uint32_t)(drop_p * 4294967296.0) : 0u;
  const float ks = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t ck[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) ck[k] = drop_col_key((uint32_t)(cg * 4 + k));
  for (int b = b0 + rl; b < b1; b += rlanes) {
    const int t = b * L + l;
    if (zero_pad_rows && !pad_mask[t]) continue;
    const uint2 raw = *reinterpret_cast<const uint2*>(dx + (size_t)t * D + cg * 4);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
    const float2 a = __bfloat1622float2(h[0]), c = __bfloat1622float2(h[1]);
    float v[4] = {a.x, a.y, c.x, c.y};
    if (drop_p > 0.f) {
      const uint32_t rk = drop_row_key(seed, drop_off, (unsigned long long)t);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = drop_mix(rk, ck[k]) >= thr ? v[k] * ks : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += v[k];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) red4[rl * D + cg * 4 + k] = acc[k];
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float sum = 0.f;
    for (int r = 0; r < rlanes; ++r) sum += red4[r * D + c];
    atomicAdd(dP + (size_t)(pos0 + l) * D + c, sum);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm forward.  One warp per output row.  gather != null: output row r reads input row gather[r] and only
// r < *n_rows_dev rows are produced (compaction of the valid targets for the CE head).
// ------------------------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                     const float* __restrict__ b, float eps, int n_rows,
                                     const int32_t* __restrict__ n_rows_dev, const int32_t* __restrict__ gather,
                                     __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out,
                                     float* __restrict__ rstd_out, int hd_valid) {
  constexpr int D = VEC * 32;
  const float inv_d = 1.f / (float)feat_count(D, hd_valid);
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int rows = n_rows_dev ? min(n_rows, *n_rows_dev) : n_rows;
  float wv[VEC], bv[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    wv[i] = w[lane * VEC + i];
    bv[i] = b[lane * VEC + i];
  }
  constexpr int RB = VEC <= 8 ? 4 : 2;   // rows per warp and iteration: their index and row loads are in flight together
  for (int rb = (blockIdx.x * wpb + (threadIdx.x >> 5)) * RB; rb < rows; rb += gridDim.x * wpb * RB) {
    int src[RB];
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      const int rr = min(rb + k, rows - 1);
      src[k] = gather ? gather[rr] : rr;
    }
    __nv_bfloat162 xv[RB][VEC / 2];
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      const __nv_bfloat16* xr = x + (size_t)src[k] * D + lane * VEC;
#pragma unroll
      for (int i = 0; i < VEC; i += 2) xv[k][i >> 1] = *reinterpret_cast<const __nv_bfloat162*>(xr + i);
    }
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      const int r = rb + k;
      if (r >= rows) break;
      float v[VEC];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; i += 2) {
        const float2 f = __bfloat1622float2(xv[k][i >> 1]);
        v[i] = f.x;
        v[i + 1] = f.y;
        s += f.x + f.y;
      }
      const float mean = warp_sum(s) * inv_d;   // padded columns hold zeros: they add nothing to the sum
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float dlt = feat_valid(lane * VEC + i, hd_valid) ? v[i] - mean : 0.f;
        q += dlt * dlt;
      }
      const float var = warp_sum(q) * inv_d;
      const float rstd = rsqrtf(var + eps);
      __nv_bfloat16* yr = y + (size_t)r * D + lane * VEC;
#pragma unroll
      for (int i = 0; i < VEC; i += 2)
        *reinterpret_cast<uint32_t*>(yr + i) =
            pack_bf16((v[i] - mean) * rstd * wv[i] + bv[i], (v[i + 1] - mean) * rstd * wv[i + 1] + bv[i + 1]);
      if (lane == 0) {
        mean_out[r] = mean;
        rstd_out[r] = rstd;
      }
    }
  }
}

// LayerNorm backward.  dy row r (compact index when gather != null) -> dx row (gather ? gather[r] : r).
//   dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w ;  dw += sum dy * xhat ; db += sum dy
// add_to != null: dx += add_to[row] (fuses the residual-branch gradient).  Rows not covered by a gather stay untouched
// (the caller zero-fills dx first when it scatters).
template <int VEC>
__global__ void layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                     const float* __restrict__ w, const float* __restrict__ mean_in,
                                     const float* __restrict__ rstd_in, int n_rows,
                                     const int32_t* __restrict__ n_rows_dev, const int32_t* __restrict__ gather,
                                     const __nv_bfloat16* __restrict__ add_to, __nv_bfloat16* __restrict__ dx,
                                     float* __restrict__ dw, float* __restrict__ db, int hd_valid) {
  constexpr int D = VEC * 32;
  const float inv_d = 1.f / (float)feat_count(D, hd_valid);
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int rows = n_rows_dev ? min(n_rows, *n_rows_dev) : n_rows;
  float wv[VEC], dw_acc[VEC], db_acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    wv[i] = w[lane * VEC + i];
    dw_acc[i] = 0.f;
    db_acc[i] = 0.f;
  }
  constexpr int RB = VEC <= 8 ? 2 : 1;   // rows per warp and iteration (loads of both rows in flight together)
  for (int rb = (blockIdx.x * wpb + (threadIdx.x >> 5)) * RB; rb < rows; rb += gridDim.x * wpb * RB) {
    int rowi[RB];
    float mean_[RB], rstd_[RB];
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      const int rr = min(rb + k, rows - 1);
      rowi[k] = gather ? gather[rr] : rr;
      mean_[k] = mean_in[rr];
      rstd_[k] = rstd_in[rr];
    }
    __nv_bfloat162 xv[RB][VEC / 2], gv[RB][VEC / 2], av[RB][VEC / 2];
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      const __nv_bfloat16* xr = x + (size_t)rowi[k] * D + lane * VEC;
      const __nv_bfloat16* gr = dy + (size_t)min(rb + k, rows - 1) * D + lane * VEC;
#pragma unroll
      for (int i = 0; i < VEC; i += 2) {
        xv[k][i >> 1] = *reinterpret_cast<const __nv_bfloat162*>(xr + i);
        gv[k][i >> 1] = *reinterpret_cast<const __nv_bfloat162*>(gr + i);
        if (add_to) av[k][i >> 1] = *reinterpret_cast<const __nv_bfloat162*>(add_to + (size_t)rowi[k] * D + lane * VEC + i);
      }
    }
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      if (rb + k >= rows) break;
      const int row = rowi[k];
      const float mean = mean_[k], rstd = rstd_[k];
      float xh[VEC], g[VEC];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; i += 2) {
        const float2 xf = __bfloat1622float2(xv[k][i >> 1]);
        const float2 gf = __bfloat1622float2(gv[k][i >> 1]);
        xh[i] = feat_valid(lane * VEC + i, hd_valid) ? (xf.x - mean) * rstd : 0.f;
        xh[i + 1] = feat_valid(lane * VEC + i + 1, hd_valid) ? (xf.y - mean) * rstd : 0.f;
        dw_acc[i] += gf.x * xh[i];
        dw_acc[i + 1] += gf.y * xh[i + 1];
        db_acc[i] += gf.x;
        db_acc[i + 1] += gf.y;
        g[i] = gf.x * wv[i];
        g[i + 1] = gf.y * wv[i + 1];
        s1 += g[i] + g[i + 1];
        s2 += g[i] * xh[i] + g[i + 1] * xh[i + 1];
      }
      s1 = warp_sum(s1) * inv_d;
      s2 = warp_sum(s2) * inv_d;
      __nv_bfloat16* o = dx + (size_t)row * D + lane * VEC;
#pragma unroll
      for (int i = 0; i < VEC; i += 2) {
        float o0 = rstd * (g[i] - s1 - xh[i] * s2), o1 = rstd * (g[i + 1] - s1 - xh[i + 1] * s2);
        if (!feat_valid(lane * VEC + i, hd_valid)) o0 = 0.f;       // padded inputs do not exist: no gradient
        if (!feat_valid(lane * VEC + i + 1, hd_valid)) o1 = 0.f;
        if (add_to) {
          const float2 af = __bfloat1622float2(av[k][i >> 1]);
          o0 += af.x;
          o1 += af.y;
        }
        *reinterpret_cast<uint32_t*>(o + i) = pack_bf16(o0, o1);
      }
    }
  }
  // block reduction of dw/db through shared memory, then one atomic per column per block
  extern __shared__ float red[];  // [2][wpb][D]
  float* rw = red;
  float* rb = red + (size_t)wpb * D;
  const int wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    rw[wid * D + lane * VEC + i] = dw_acc[i];
    rb[wid * D + lane * VEC + i] = db_acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float a = 0.f, bsum = 0.f;
    for (int k = 0; k < wpb; ++k) {
      a += rw[k * D + c];
      bsum += rb[k * D + c];
    }
    atomicAdd(dw + c, a);
    atomicAdd(db + c, bsum);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// out = in * dropout_mask / keep  (the mask is regenerated from the forward's (seed, offset, geometry)); optional row mask
// ------------------------------------------------------------------------------------------------------------------
__global__ void dropout_bwd_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n,
                                   int cols, const uint8_t* __restrict__ rowmask, float drop_p,
                                   unsigned long long seed, unsigned long long drop_off,
                                   const unsigned long long* __restrict__ seed_ptr) {
  if (drop_p > 0.f && seed_ptr) seed += *seed_ptr;
  const uint32_t thr = drop_p > 0.f ? (uint32_t)(drop_p * 4294967296.0) : 0u;
  const float ks = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n;
       i += (long long)gridDim.x * blockDim.x * 4) {
    const uint2 raw = *reinterpret_cast<const uint2*>(in + i);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
    float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    float v[4] = {a.x, a.y, b.x, b.y};
    if (drop_p > 0.f) {
      const long long row = i / cols;
      const uint32_t col = (uint32_t)(i - row * cols);   // cols % 4 == 0: the 4 elements share the row
      const uint32_t rk = drop_row_key(seed, drop_off, (unsigned long long)row);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = drop_mix(rk, drop_col_key(col + k)) >= thr ? v[k] * ks : 0.f;
    }
    if (rowmask && !rowmask[i / cols]) v[0] = v[1] = v[2] = v[3] = 0.f;
    uint2 w;
    w.x = pack_bf16(v[0], v[1]);
    w.y = pack_bf16(v[2], v[3]);
    *reinterpret_cast<uint2*>(out + i) = w;
  }
}

// db[c] += sum_r dY[r, c]        dY bf16 [rows, cols] with pitch ld.  thread = 4 columns x one row lane (8-byte loads,
// 4 rows in flight per thread), smem reduce over the row lanes, one atomic per column per block.
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ dy, int rows, int cols, long long ld,
                              float* __restrict__ db) {
  extern __shared__ float red4[];  // [rlanes][cols]
  const int tpr = cols / 4, rlanes = blockDim.x / tpr;
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (rl < rlanes) {
    const int stride = gridDim.x * rlanes;
    int r = blockIdx.x * rlanes + rl;
    for (; r + 3 * stride < rows; r += 4 * stride) {
      uint2 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) raw[u] = *reinterpret_cast<const uint2*>(dy + (size_t)(r + u * stride) * ld + cg * 4);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[u]);
        const float2 a = __bfloat1622float2(h[0]), c = __bfloat1622float2(h[1]);
        acc[0] += a.x; acc[1] += a.y; acc[2] += c.x; acc[3] += c.y;
      }
    }
    for (; r < rows; r += stride) {
      const uint2 raw = *reinterpret_cast<const uint2*>(dy + (size_t)r * ld + cg * 4);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
      const float2 a = __bfloat1622float2(h[0]), c = __bfloat1622float2(h[1]);
      acc[0] += a.x; acc[1] += a.y; acc[2] += c.x; acc[3] += c.y;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) red4[rl * cols + cg * 4 + k] = acc[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float sum = 0.f;
    for (int r = 0; r < rlanes; ++r) sum += red4[r * cols + c];
    atomicAdd(db + c, sum);
  }
}

// Several column sums in ONE launch (blockIdx.y selects the tensor): the five bias gradients of a transformer block's
// backward are small, independent reductions whose launch / drain overhead would otherwise be paid five times.
struct ColsumBatch {
  const __nv_bfloat16* dy[6];
  float* db[6];
  long long ld[6];
  int cols[6];
  int rows;
};
__global__ void colsum_multi_kernel(const ColsumBatch b) {
  extern __shared__ float red4[];  // [rlanes][cols]
  const int which = blockIdx.y;
  const __nv_bfloat16* __restrict__ dy = b.dy[which];
  float* __restrict__ db = b.db[which];
  const int cols = b.cols[which], rows = b.rows;
  const long long ld = b.ld[which];
  const int tpr = cols / 4, rlanes = blockDim.x / tpr;
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (rl < rlanes) {
    const int stride = gridDim.x * rlanes;
    int r = blockIdx.x * rlanes + rl;
    for (; r + 3 * stride < rows; r += 4 * stride) {
      uint2 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) raw[u] = *reinterpret_cast<const uint2*>(dy + (size_t)(r + u * stride) * ld + cg * 4);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[u]);
        const float2 a = __bfloat1622float2(h[0]), c = __bfloat1622float2(h[1]);
        acc[0] += a.x; acc[1] += a.y; acc[2] += c.x; acc[3] += c.y;
      }
    }
    for (; r < rows; r += stride) {
      const uint2 raw = *reinterpret_cast<const uint2*>(dy + (size_t)r * ld + cg * 4);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
      const float2 a = __bfloat1622float2(h[0]), c = __bfloat1622float2(h[1]);
      acc[0] += a.x; acc[1] += a.y; acc[2] += c.x; acc[3] += c.y;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) red4[rl * cols + cg * 4 + k] = acc[k];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float sum = 0.f;
    for (int r = 0; r < rlanes; ++r) sum += red4[r * cols + c];
    atomicAdd(db + c, sum);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam, no weight decay / amsgrad) over one flat fp32 parameter buffer; also refreshes the bf16 shadow
// copy the kernels consume and zeroes the gradient for the next step.  The step counter and lr live in device memory so
// the launch is CUDA-graph replayable.  grad_scale multiplies the gradient (1/world_size after a sum all-reduce).
// ------------------------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            __nv_bfloat16* __restrict__ shadow, long long n, const float* __restrict__ lr_dev,
                            const int32_t* __restrict__ step_dev, float beta1, float beta2, float eps, float grad_scale,
                            const uint8_t* __restrict__ frozen /* per-element freeze mask or null */, int zero_grad) {
  const float lr = *lr_dev;
  const int step = *step_dev;  // 1-based, already incremented by adam_tick_kernel
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n;
       i += (long long)gridDim.x * blockDim.x * 4) {
    float4 pp = *reinterpret_cast<float4*>(p + i), gg = *reinterpret_cast<float4*>(g + i);
    float4 mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
    float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = ga[k] * grad_scale;
      ma[k] = beta1 * ma[k] + (1.f - beta1) * gk;
      va[k] = beta2 * va[k] + (1.f - beta2) * gk * gk;
      const float denom = sqrtf(va[k]) * inv_sqrt_bc2 + eps;
      const float upd = step_size * ma[k] / denom;
      if (!frozen || !frozen[i + k]) pa[k] -= upd;
    }
    *reinterpret_cast<float4*>(p + i) = pp;
    *reinterpret_cast<float4*>(m + i) = mm;
    *reinterpret_cast<float4*>(v + i) = vv;
    if (zero_grad) *reinterpret_cast<float4*>(g + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (shadow) {
      uint2 w;
      w.x = pack_bf16(pp.x, pp.y);
      w.y = pack_bf16(pp.z, pp.w);
      *reinterpret_cast<uint2*>(shadow + i) = w;
    }
  }
}

__global__ void adam_tick_kernel(int32_t* step_dev) { *step_dev += 1; }

__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n;
       i += (long long)gridDim.x * blockDim.x * 4) {
    const float4 f = *reinterpret_cast<const float4*>(src + i);
    uint2 w;
    w.x = pack_bf16(f.x, f.y);
    w.y = pack_bf16(f.z, f.w);
    *reinterpret_cast<uint2*>(dst + i) = w;
  }
}

static inline int grid_for(long long work_items, int per_block) {
  long long b = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace rp

using namespace rp;

RP_API int rp_prepare_batch(const int64_t* ids, const uint8_t* pad_mask, const int64_t* labels, const uint8_t* target_mask,
                            int T, int pad_id, int n_items, int32_t* ids32, int32_t* valid_idx, int32_t* labels_c,
                            int32_t* n_valid, int32_t* scratch /* >= ceil(T/1024) ints */, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!ids || !pad_mask || !ids32 || T <= 0) return RP_EINVAL;
  if (target_mask && (!labels || !valid_idx || !labels_c || !n_valid)) return RP_EINVAL;
  const int n_blocks = (T + 1023) / 1024;
  if (target_mask && !scratch) return RP_EINVAL;
  prepare_count_kernel<<<n_blocks, 1024, 0, stream>>>(ids, pad_mask, labels, target_mask, T, pad_id, n_items, ids32, scratch);
  RP_LAUNCH_CHECK();
  if (target_mask) {
    prepare_write_kernel<<<n_blocks, 1024, 0, stream>>>(labels, target_mask, T, n_items, scratch, valid_idx, labels_c, n_valid);
    RP_LAUNCH_CHECK();
  }
  return RP_OK;
}

#define RP_DISPATCH_D(d, CALL)            \
  switch (d) {                            \
    case 64: { constexpr int VEC = 2; CALL; } break;   \
    case 128: { constexpr int VEC = 4; CALL; } break;  \
    case 256: { constexpr int VEC = 8; CALL; } break;  \
    case 512: { constexpr int VEC = 16; CALL; } break; \
    default: return RP_ESHAPE;            \
  }

RP_API int rp_embed_fwd(const void* table, const float* pos, const int32_t* ids, const uint8_t* pad_mask, int T, int L,
                        int d, int pos0, float scale, int zero_pad_rows, float drop_p, unsigned long long seed,
                        unsigned long long drop_off, const unsigned long long* seed_ptr, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!table || !pos || !ids || !out || T <= 0 || L <= 0) return RP_EINVAL;
  if (T % L != 0) return RP_ESHAPE;   // whole sequences (the kernel walks position by position)
  const int grid = grid_for(T, 8);
  RP_DISPATCH_D(d, (embed_fwd_kernel<VEC><<<grid, 256, 0, stream>>>(
                       reinterpret_cast<const __nv_bfloat16*>(table), pos, ids, pad_mask, T, L, pos0, scale, zero_pad_rows,
                       drop_p, seed, drop_off, seed_ptr, nullptr, nullptr, reinterpret_cast<__nv_bfloat16*>(out))));
  RP_LAUNCH_CHECK();
  return RP_OK;
}

RP_API int rp_embed_bwd(const void* dx, const int32_t* ids, const uint8_t* pad_mask, int B, int L, int d, int pad_id,
                        int pos0, float scale, int zero_pad_rows, float drop_p, unsigned long long seed,
                        unsigned long long drop_off, const unsigned long long* seed_ptr, float* d_table, float* d_pos,
                        void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!dx || !ids || !d_table || !d_pos || B <= 0 || L <= 0) return RP_EINVAL;
  const int T = B * L;
  const int grid = grid_for(T, 8);
  RP_DISPATCH_D(d, (embed_bwd_table_kernel<VEC><<<grid, 256, 0, stream>>>(
                       reinterpret_cast<const __nv_bfloat16*>(dx), ids, pad_mask, T, pad_id, scale, zero_pad_rows, drop_p,
                       seed, drop_off, seed_ptr, nullptr, nullptr, d_table)));
  RP_LAUNCH_CHECK();
  {
    const int rlanes = 256 / (d / 4);
    int G = (B + 31) / 32;
    if (G < 1) G = 1;
    if (G > 16) G = 16;
    embed_bwd_pos_kernel<<<dim3(L, G), 256, (size_t)rlanes * d * sizeof(float), stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(dx), pad_mask, B, L, d, pos0, zero_pad_rows, drop_p, seed, drop_off, seed_ptr,
        d_pos);
  }
  RP_LAUNCH_CHECK();
  return RP_OK;
}

RP_API int rp_layernorm_fwd(const void* x, const float* w, const float* b, float eps, int n_rows, int d,
                            const int32_t* n_rows_dev, const int32_t* gather, void* y, float* mean, float* rstd,
                            int hd_valid, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !w || !b || !y || !mean || !rstd || n_rows <= 0) return RP_EINVAL;
  if (hd_valid < 0 || hd_valid > 128 || (hd_valid > 0 && d % (hd_valid <= 64 ? 64 : 128))) return RP_ESHAPE;
  const int grid = grid_for(n_rows, 8);
  RP_DISPATCH_D(d, (layernorm_fwd_kernel<VEC><<<grid, 256, 0, stream>>>(
                       reinterpret_cast<const __nv_bfloat16*>(x), w, b, eps, n_rows, n_rows_dev, gather,
                       reinterpret_cast<__nv_bfloat16*>(y), mean, rstd, hd_valid)));
  RP_LAUNCH_CHECK();
  return RP_OK;
}

RP_API int rp_layernorm_bwd(const void* dy, const void* x, const float* w, const float* mean, const float* rstd,
                            int n_rows, int d, const int32_t* n_rows_dev, const int32_t* gather, const void* add_to,
                            void* dx, float* dw, float* db, int hd_valid, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!dy || !x || !w || !mean || !rstd || !dx || !dw || !db || n_rows <= 0) return RP_EINVAL;
  if (hd_valid < 0 || hd_valid > 128 || (hd_valid > 0 && d % (hd_valid <= 64 ? 64 : 128))) return RP_ESHAPE;
  int grid = grid_for(n_rows, 8 * 16);  // each warp walks ~16 rows so the dw/db atomics stay few
  const size_t smem = (size_t)2 * 8 * d * sizeof(float);
  RP_DISPATCH_D(d, (layernorm_bwd_kernel<VEC><<<grid, 256, smem, stream>>>(
                       reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(x), w, mean, rstd,
                       n_rows, n_rows_dev, gather, reinterpret_cast<const __nv_bfloat16*>(add_to),
                       reinterpret_cast<__nv_bfloat16*>(dx), dw, db, hd_valid)));
  RP_LAUNCH_CHECK();
  return RP_OK;
}

RP_API int rp_dropout_bwd(const void* in, void* out, long long rows, int cols, const uint8_t* rowmask, float drop_p,
                          unsigned long long seed, unsigned long long drop_off, const unsigned long long* seed_ptr,
                          void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!in || !out || rows <= 0 || cols <= 0 || (cols & 3)) return RP_EINVAL;
  const long long n = rows * cols;
  dropout_bwd_kernel<<<grid_for(n / 4, 256), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(in),
                                                                reinterpret_cast<__nv_bfloat16*>(out), n, cols, rowmask,
                                                                drop_p, seed, drop_off, seed_ptr);
  RP_LAUNCH_CHECK();
  return RP_OK;
}

RP_API int rp_colsum(const void* dy, int rows, int cols, long long ld, float* db, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!dy || !db || rows <= 0 || cols <= 0 || (cols & 3) || cols > 1024 || (ld & 3)) return RP_EINVAL;
  const int rlanes = 256 / (cols / 4);
  int grid = sm_count() * 4;
  if (grid > (rows + rlanes - 1) / rlanes) grid = (rows + rlanes - 1) / rlanes;
  colsum_kernel<<<grid, 256, (size_t)rlanes * cols * sizeof(float), stream>>>(reinterpret_cast<const __nv_bfloat16*>(dy), rows,
                                                                             cols, ld, db);
  RP_LAUNCH_CHECK();
  return RP_OK;
}

// db[i][c] += sum_r dy[i][r, c] for n <= 6 tensors that share the row count (one launch)
RP_API int rp_colsum_multi(int n, const void* const* dy, const int* cols, const long long* ld, float* const* db, int rows,
                           void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (n <= 0 || n > 6 || !dy || !cols || !ld || !db || rows <= 0) return RP_EINVAL;
  ColsumBatch b;
  int max_cols = 0, min_cols = 1 << 30;
  for (int i = 0; i < n; ++i) {
    if (!dy[i] || !db[i] || cols[i] <= 0 || (cols[i] & 3) || cols[i] > 1024 || (ld[i] & 3)) return RP_EINVAL;
    b.dy[i] = reinterpret_cast<const __nv_bfloat16*>(dy[i]);
    b.db[i] = db[i];
    b.ld[i] = ld[i];
    b.cols[i] = cols[i];
    max_cols = cols[i] > max_cols ? cols[i] : max_cols;
    min_cols = cols[i] < min_cols ? cols[i] : min_cols;
  }
  b.rows = rows;
  const int rlanes_max = 256 / (min_cols / 4), rlanes_min = 256 / (max_cols / 4);
  if (rlanes_min < 1) return RP_ESHAPE;
  int gx = sm_count() * 4 / n;
  if (gx < 1) gx = 1;
  if (gx > (rows + rlanes_min - 1) / rlanes_min) gx = (rows + rlanes_min - 1) / rlanes_min;
  const size_t smem = (size_t)(rlanes_max > rlanes_min ? rlanes_max : rlanes_min) * max_cols * sizeof(float);
  colsum_multi_kernel<<<dim3(gx, n), 256, smem, stream>>>(b);
  RP_LAUNCH_CHECK();
  return RP_OK;
}

RP_API int rp_adam_step(float* p, float* g, float* m, float* v, void* shadow_bf16, long long n, const float* lr_dev,
                        int32_t* step_dev, float beta1, float beta2, float eps, float grad_scale, const uint8_t* frozen,
                        int zero_grad, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!p || !g || !m || !v || !lr_dev || !step_dev || n <= 0 || (n & 3)) return RP_EINVAL;
  adam_tick_kernel<<<1, 1, 0, stream>>>(step_dev);
  adam_kernel<<<grid_for(n / 4, 256), 256, 0, stream>>>(p, g, m, v, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), n, lr_dev,
                                                        step_dev, beta1, beta2, eps, grad_scale, frozen, zero_grad);
  RP_LAUNCH_CHECK();
  return RP_OK;
}

RP_API int rp_cast_bf16(const float* src, void* dst, long long n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!src || !dst || n <= 0 || (n & 3)) return RP_EINVAL;
  cast_bf16_kernel<<<grid_for(n / 4, 256), 256, 0, stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  RP_LAUNCH_CHECK();
  return RP_OK;
}

__global__ void counter_add_kernel(unsigned long long* c, unsigned long long inc) { *c += inc; }

RP_API int rp_counter_add(unsigned long long* counter, unsigned long long inc, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!counter) return RP_EINVAL;
  counter_add_kernel<<<1, 1, 0, stream>>>(counter, inc);
  RP_LAUNCH_CHECK();
  return RP_OK;
}

// ---- BERT4Rec embedding: where(token_mask, E[ids], mask_emb) + P[t % L], no sqrt(d) scaling (bert4rec/model.py:239-296)
RP_API int rp_bert_embed_fwd(const void* table, const void* mask_emb, const float* pos, const int32_t* ids,
                             const uint8_t* tok_mask, int T, int L, int d, float drop_p, unsigned long long seed,
                             unsigned long long drop_off, const unsigned long long* seed_ptr, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!table || !mask_emb || !pos || !ids || !tok_mask || !out || T <= 0 || L <= 0) return RP_EINVAL;
  if (T % L != 0) return RP_ESHAPE;
  const int grid = grid_for(T, 8);
  RP_DISPATCH_D(d, (embed_fwd_kernel<VEC><<<grid, 256, 0, stream>>>(
                       reinterpret_cast<const __nv_bfloat16*>(table), pos, ids, tok_mask, T, L, 0, 1.f, 0, drop_p, seed, drop_off,
                       seed_ptr, tok_mask, reinterpret_cast<const __nv_bfloat16*>(mask_emb), reinterpret_cast<__nv_bfloat16*>(out))));
  RP_LAUNCH_CHECK();
  return RP_OK;
}

RP_API int rp_bert_embed_bwd(const void* dx, const int32_t* ids, const uint8_t* pad_mask, const uint8_t* tok_mask, int B, int L,
                             int d, float drop_p, unsigned long long seed, unsigned long long drop_off,
                             const unsigned long long* seed_ptr, float* d_table, float* d_mask_emb, float* d_pos,
                             void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!dx || !ids || !pad_mask || !tok_mask || !d_table || !d_mask_emb || !d_pos || B <= 0 || L <= 0) return RP_EINVAL;
  const int T = B * L;
  const int grid = grid_for(T, 8);
  RP_DISPATCH_D(d, (embed_bwd_table_kernel<VEC><<<grid, 256, 0, stream>>>(
                       reinterpret_cast<const __nv_bfloat16*>(dx), ids, pad_mask, T, -1, 1.f, 0, drop_p, seed, drop_off, seed_ptr,
                       tok_mask, d_mask_emb, d_table)));
  RP_LAUNCH_CHECK();
  {
    const int rlanes = 256 / (d / 4);
    int G = (B + 31) / 32;
    if (G < 1) G = 1;
    if (G > 16) G = 16;
    // positional gradient: pad positions carry an exactly-zero dx, so no masking is needed
    embed_bwd_pos_kernel<<<dim3(L, G), 256, (size_t)rlanes * d * sizeof(float), stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(dx), pad_mask, B, L, d, 0, 1, drop_p, seed, drop_off, seed_ptr, d_pos);
  }
  RP_LAUNCH_CHECK();
  return RP_OK;
}

// ---- row gather / scatter with a device-side row count (valid-target compaction without a LayerNorm)
template <int VEC>
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, const int32_t* __restrict__ idx, int n_max,
                                   const int32_t* __restrict__ n_dev, __nv_bfloat16* __restrict__ dst, int scatter) {
  constexpr int D = VEC * 32;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int n = n_dev ? min(n_max, *n_dev) : n_max;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n; r += gridDim.x * wpb) {
    const size_t a = (size_t)(scatter ? r : idx[r]) * D + lane * VEC, b = (size_t)(scatter ? idx[r] : r) * D + lane * VEC;
#pragma unroll
    for (int i = 0; i < VEC; i += 2) *reinterpret_cast<uint32_t*>(dst + b + i) = *reinterpret_cast<const uint32_t*>(src + a + i);
  }
}

// scatter = 0: dst[r] = src[idx[r]] ; scatter = 1: dst[idx[r]] = src[r]   (r < min(n_max, *n_dev))
RP_API int rp_gather_rows(const void* src, const int32_t* idx, int n_max, const int32_t* n_dev, int d, void* dst, int scatter,
                          void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!src || !idx || !dst || n_max <= 0) return RP_EINVAL;
  const int grid = grid_for(n_max, 8);
  RP_DISPATCH_D(d, (gather_rows_kernel<VEC><<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(src), idx, n_max,
                                                                      n_dev, reinterpret_cast<__nv_bfloat16*>(dst), scatter)));
  RP_LAUNCH_CHECK();
  return RP_OK;
}
