// rp_selftest.cu - single-CTA tcgen05 bring-up test: D[128,128] = A[128,128] * B[128,128]^T in the operand modes the
// production kernels rely on.  Exposed through the C ABI as rp_selftest_umma so a GPU test can pin the descriptor
// encodings (K-major / MN-major shared-memory operands, A operand from TMEM) against a plain matmul.
#include "rp_host.h"
#include "rp_sm100.cuh"

namespace rp {

// mode bit 0: B is MN-major (global Bt[K,N])   bit 1: A from TMEM   bit 2: A is MN-major (global At[K,M])
template <int MODE>
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __nv_bfloat16* __restrict__ Araw, float* __restrict__ D) {
  constexpr bool B_MN = (MODE & 1) != 0;
  constexpr bool A_TMEM = (MODE & 2) != 0;
  constexpr bool A_MN = (MODE & 4) != 0;
  constexpr int M = 128, N = 128, K = 128;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;            // 32 KB
  uint8_t* sB = smem + 32768;    // 32 KB
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tmem_d = tmem;        // 128 fp32 columns
  const uint32_t tmem_a = tmem + 128;  // 64 columns of packed bf16 pairs

  if (threadIdx.x == 0) {
    uint32_t bytes = 32768 + (A_TMEM ? 0 : 32768);
    mbar_arrive_expect_tx(&bar_full, bytes);
    if (!A_TMEM) {
      // two boxes of [128 rows x 64 cols]; for K-major these are the two K chunks, for MN-major the two M chunks
      tma_load_2d(sA, &tmA, &bar_full, 0, 0);
      tma_load_2d(sA + 16384, &tmA, &bar_full, 64, 0);
    }
    tma_load_2d(sB, &tmB, &bar_full, 0, 0);
    tma_load_2d(sB + 16384, &tmB, &bar_full, 64, 0);
  }
  if (A_TMEM) {
    // thread t owns TMEM lane t = row t of A; pack (k, k+1) into one 32-bit column
    const int row = warp * 32 + lane;
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(Araw + (size_t)row * K);
#pragma unroll
    for (int c = 0; c < 64; c += 16) {
      uint32_t v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = arow[c + j];
      tmem_st16(tmem_a + ((uint32_t)(warp * 32) << 16) + c, v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  if (threadIdx.x == 0) {
    mbar_wait(&bar_full, 0);
    tc_fence_after();
    constexpr uint32_t idesc = umma_idesc_bf16(M, N, A_MN, B_MN);
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t bdesc;
      if (B_MN)
        bdesc = umma_desc_sw128(smem_u32(sB) + ks * 2048, /*lbo*/ 16384, /*sbo*/ 1024);
      else
        bdesc = umma_desc_sw128(smem_u32(sB) + (ks / 4) * 16384 + (ks % 4) * 32, 16, 1024);
      if (A_TMEM) {
        umma_ts(tmem_d, tmem_a + ks * 8, bdesc, idesc, ks > 0);
      } else {
        uint64_t adesc;
        if (A_MN)
          adesc = umma_desc_sw128(smem_u32(sA) + ks * 2048, 16384, 1024);
        else
          adesc = umma_desc_sw128(smem_u32(sA) + (ks / 4) * 16384 + (ks % 4) * 32, 16, 1024);
        umma_ss(tmem_d, adesc, bdesc, idesc, ks > 0);
      }
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  {
    const int row = warp * 32 + lane;
#pragma unroll
    for (int c = 0; c < N; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) D[(size_t)row * N + c + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// Probe (tools/probe_tma.py): how fast can ONE SM pull [box_rows x 64] bf16 boxes of a K-major [rows, d] table through TMA
// into a 6-stage ring when nothing consumes them?  Every CTA streams `tiles` row tiles (all d/64 chunks each), CTAs start
// at different tiles.  The ceiling this gives bounds the B-operand feed of score_topk / CE kernels (32 KB per 128x128 tile).
__global__ void __launch_bounds__(64, 1) tma_probe_kernel(const __grid_constant__ CUtensorMap tm, int n_row_tiles, int kch,
                                                          int tiles, int box_bytes, int same_tile) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NST = 6;
  __shared__ uint64_t bar_full[NST];
  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) mbar_init(&bar_full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t it = 0;
    for (int t = 0; t < tiles; ++t) {
      const int rt = same_tile ? (int)(blockIdx.x % n_row_tiles) : (int)((blockIdx.x * 7 + t) % n_row_tiles);
      for (int kc = 0; kc < kch; ++kc, ++it) {
        const uint32_t s = it % NST;
        if (it >= NST) mbar_wait(&bar_full[s], ((it / NST) - 1) & 1);  // the previous load into this slot has landed
        mbar_arrive_expect_tx(&bar_full[s], box_bytes);
        tma_load_2d(smem + s * 32768, &tm, &bar_full[s], kc * 64, rt * (box_bytes / 128));
      }
    }
    for (uint32_t k = (it > NST ? it - NST : 0); k < it; ++k) mbar_wait(&bar_full[k % NST], (k / NST) & 1);
  }
  __syncthreads();
}

// Probe (tools/probe_mma.py): issue rate of tcgen05.mma 128x128x16 bf16 for the operand forms the kernels use, with nothing
// else going on: one elected thread issues `iters` groups of 8 MMAs (K = 128) on fixed shared-memory / TMEM operands
// (contents irrelevant) and the CTA reports elapsed SM cycles.  mode bit0: B MN-major, bit1: A from TMEM, bit2: A MN-major,
// bits 3-4: N = 128 / 256 / 64.
__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int mode, int iters, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0 && elect_one()) {
    const bool b_mn = mode & 1, a_tmem = mode & 2, a_mn = mode & 4;
    const int nsel = (mode >> 3) & 3, N = nsel == 1 ? 256 : (nsel == 2 ? 64 : 128);
    const uint32_t idesc = umma_idesc_bf16(128, N, a_mn, b_mn);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32768);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t ad = a_mn ? umma_desc_sw128(a0 + ks * 2048, 16384, 1024) : umma_desc_sw128(a0 + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
        const uint64_t bd = b_mn ? umma_desc_sw128(b0 + ks * 2048, 16384, 1024) : umma_desc_sw128(b0 + (ks >> 2) * (N * 128) + (ks & 3) * 32, 16, 1024);
        if (a_tmem) umma_ts(tmem + (N == 256 ? 0 : (it & 1) * 128), tmem + 256 + ks * 8, bd, idesc, ks != 0);
        else umma_ss(tmem + (N == 256 ? 0 : (it & 1) * 128), ad, bd, idesc, ks != 0);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace rp

RP_API int rp_selftest_mma_probe(int mode, int iters, int grid, long long* cycles_out, void* stream_) {
  using namespace rp;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (mode < 0 || mode > 31 || iters <= 0 || grid <= 0 || !cycles_out) return RP_EINVAL;
  const int smem = 98304 + 1024;
  RP_CUDA_CHECK(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mma_probe_kernel<<<grid, 128, smem, stream>>>(mode, iters, cycles_out);
  RP_LAUNCH_CHECK();
  return RP_OK;
}

RP_API int rp_selftest_tma_probe(const void* table, long long rows, int d, int box_rows, int tiles, int same_tile, int grid,
                                 void* stream_) {
  using namespace rp;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CUtensorMap tm;
  int rc;
  if (!table || rows <= 0 || d % 64 || (box_rows != 64 && box_rows != 128 && box_rows != 256)) return RP_EINVAL;
  if ((rc = make_tmap_bf16(&tm, table, rows, d, d, box_rows)) != RP_OK) return rc;
  const int smem = 6 * 32768 + 1024;
  RP_CUDA_CHECK(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tma_probe_kernel<<<grid, 64, smem, stream>>>(tm, (int)(rows / box_rows), d / 64, tiles, box_rows * 128, same_tile);
  RP_LAUNCH_CHECK();
  return RP_OK;
}

RP_API int rp_selftest_umma(int mode, const void* A, const void* B, float* D, void* stream_) {
  using namespace rp;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CUtensorMap tmA, tmB;
  int rc;
  // every operand is a [128,128] bf16 row-major array; what the rows mean depends on the mode
  if ((rc = make_tmap_bf16(&tmA, A, 128, 128, 128, 128)) != RP_OK) return rc;
  if ((rc = make_tmap_bf16(&tmB, B, 128, 128, 128, 128)) != RP_OK) return rc;
  const int smem = 65536 + 1024;
  const __nv_bfloat16* a = reinterpret_cast<const __nv_bfloat16*>(A);
#define RP_ST_CASE(m)                                                                                      \
  case m:                                                                                                  \
    RP_CUDA_CHECK(cudaFuncSetAttribute(umma_selftest_kernel<m>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    umma_selftest_kernel<m><<<1, 128, smem, stream>>>(tmA, tmB, a, D);                                     \
    break;
  switch (mode) {
    RP_ST_CASE(0)
    RP_ST_CASE(1)
    RP_ST_CASE(2)
    RP_ST_CASE(3)
    RP_ST_CASE(4)
    RP_ST_CASE(5)
    default:
      return RP_EINVAL;
  }
#undef RP_ST_CASE
  RP_LAUNCH_CHECK();
  return RP_OK;
}
