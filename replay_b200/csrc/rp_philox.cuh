// rp_philox.cuh - counter-based RNGs: the forward and the backward regenerate the same dropout mask from
// (seed, element index / 4) instead of storing it.  Element e uses word (e & 3) of rng4x32(seed, e >> 2).
#pragma once
#include <stdint.h>

namespace rp {

__host__ __device__ __forceinline__ uint4 philox4x32(unsigned long long seed, unsigned long long ctr) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0u, c3 = 0u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
    const unsigned long long p0 = (unsigned long long)M0 * c0, p1 = (unsigned long long)M1 * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return make_uint4(c0, c1, c2, c3);
}

__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {  // murmur3 finaliser: full avalanche in 5 steps
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}

// Activation-dropout generator: four 32-bit words for the 4 consecutive elements of counter `ctr` (= element index / 4).
// Dropout only needs independent-looking Bernoulli draws that the backward can regenerate; Philox4x32-10 costs ~110
// integer instructions per call and showed up as ~5 % of the training step, this counter hash (one mixed 32-bit key per
// (seed, ctr), four finalisers on a Weyl sequence) costs ~30.  Philox stays in use where a reference-grade stream matters
// (the BERT4Rec token masker).
__host__ __device__ __forceinline__ uint4 rng4x32(unsigned long long seed, unsigned long long ctr) {
  const uint32_t key = fmix32((uint32_t)ctr * 0x9E3779B1u ^ (uint32_t)seed) ^
                       fmix32((uint32_t)(ctr >> 32) * 0x85EBCA77u + (uint32_t)(seed >> 32) + 0x27D4EB2Fu);
  return make_uint4(fmix32(key), fmix32(key + 0x9E3779B9u), fmix32(key + 0x3C6EF372u), fmix32(key + 0xDAA66D2Bu));
}

// keep decision for element e with drop threshold thr = p * 2^32
__device__ __forceinline__ bool philox_keep(unsigned long long seed, unsigned long long e, uint32_t thr) {
  const uint4 r = rng4x32(seed, e >> 2);
  const uint32_t w = (e & 3) == 0 ? r.x : (e & 3) == 1 ? r.y : (e & 3) == 2 ? r.z : r.w;
  return w >= thr;
}

// Cheap per-element hash (murmur3 finaliser) for the attention-probability dropout: the fused attention backward walks
// the [query, key] matrix transposed, so the mask must be computable per element in any order (Philox is used where the
// forward and backward visit 4 consecutive elements together).
__device__ __forceinline__ uint32_t drop_hash32(unsigned long long seed, unsigned long long idx) {
  uint32_t h = ((uint32_t)idx * 0x9E3779B1u) ^ (uint32_t)seed ^ ((uint32_t)(idx >> 32) * 0x85EBCA77u) ^ (uint32_t)(seed >> 32);
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}

}  // namespace rp
