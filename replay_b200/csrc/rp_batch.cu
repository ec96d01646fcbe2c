// rp_batch.cu - device-side batch construction: all user histories live in HBM as one CSR store
// (offsets + item ids); one launch cuts, left-pads, shifts and masks the windows of a whole batch.  Replaces the per-sample
// host path of the reference: TorchSequentialDataset.__getitem__ / _pad_sequence / _generate_padding_mask
// (replay/data/nn/torch_sequential_dataset.py:69-136), SasRecTrainingDataset.__getitem__
// (replay/models/nn/sequential/sasrec/dataset.py:104-126), Bert4RecUniformMasker.mask + Bert4RecTrainingDataset.__getitem__
// (replay/models/nn/sequential/bert4rec/dataset.py:71-92,163-177), _shift_features (bert4rec/dataset.py:322-351) and the
// default collate that stacks the samples; and the new path's torch-op version of the same thing: Array1DColumn.__getitem__
// (replay/data/nn/parquet/impl/array_1d_column.py:70-84, indexing.py:42-78) + NextTokenTransform
// (replay/nn/transform/next_token.py:65-96), ~10 small kernels there.
//
// HBM-bound integer work: per row one CSR offset pair + <= W item ids are read (coalesced along the window) and W x
// (8 + 1 [+ 8 + 1]) bytes are written; one CTA per batch row so the BERT masker's row-wide all()/any() fix-ups are block
// reductions.
#include "rp_host.h"
#include "rp_philox.cuh"

namespace rp {

enum { kSasrecTrain = 0, kPredict = 1, kBertTrain = 2, kBertPredict = 3 };

struct BatchArgs {
  const int64_t* offsets;
  const int32_t* items;
  const int32_t* seq_index;
  const int32_t* seq_offset;
  const int64_t* query_ids;
  const float* uniforms;
  int64_t* ids;
  uint8_t* pad_mask;
  int64_t* labels;
  uint8_t* aux_mask;
  int64_t* query_out;
  long long n_seq;
  int B, L, mode, pad_value;
  float mask_prob;
  unsigned long long seed, draw0;
};

// window position w of a W-wide left-padded window holding the n items [first, first+n) of the history
__device__ __forceinline__ int64_t window_item(const int32_t* __restrict__ items, long long first, int n, int W, int w,
                                               int pad_value) {
  const int k = w - (W - n);
  return k >= 0 ? (int64_t)items[first + k] : (int64_t)pad_value;
}

__global__ void __launch_bounds__(128) build_batch_kernel(const BatchArgs a) {
  const int b = blockIdx.x;
  const int s = a.seq_index[b];
  const long long beg = a.offsets[s], end = a.offsets[s + 1];
  const int len = (int)(end - beg);
  const int shift = a.mode == kSasrecTrain ? 1 : 0;
  const int L = a.L, W = L + shift;
  int off = a.seq_offset ? a.seq_offset[b] : max(0, len - W);
  off = min(max(off, 0), len);
  const int n = min(len - off, W);  // items inside the window; the mask has exactly n trailing ones
  const long long first = beg + off;
  if (threadIdx.x == 0 && a.query_out) a.query_out[b] = a.query_ids ? a.query_ids[s] : (int64_t)s;

  int64_t* ids = a.ids + (size_t)b * L;
  uint8_t* pm = a.pad_mask + (size_t)b * L;

  if (a.mode == kSasrecTrain) {
    // inputs = window[:-1], labels = window[1:], masks likewise (sasrec/dataset.py:107-118)
    int64_t* lab = a.labels + (size_t)b * L;
    uint8_t* tm = a.aux_mask + (size_t)b * L;
    for (int p = threadIdx.x; p < L; p += blockDim.x) {
      ids[p] = window_item(a.items, first, n, W, p, a.pad_value);
      pm[p] = p >= W - n;
      lab[p] = window_item(a.items, first, n, W, p + 1, a.pad_value);
      tm[p] = p + 1 >= W - n;
    }
    return;
  }
  if (a.mode == kPredict) {
    for (int p = threadIdx.x; p < L; p += blockDim.x) {
      ids[p] = window_item(a.items, first, n, W, p, a.pad_value);
      pm[p] = p >= W - n;
    }
    return;
  }
  if (a.mode == kBertPredict) {
    // roll the window left by one, last slot = padding; token_mask = shifted pad mask, pad_mask = that with last = 1
    uint8_t* tk = a.aux_mask + (size_t)b * L;
    for (int p = threadIdx.x; p < L; p += blockDim.x) {
      const bool last = p == L - 1;
      ids[p] = last ? (int64_t)a.pad_value : window_item(a.items, first, n, W, p + 1, a.pad_value);
      const uint8_t t = last ? 0 : (uint8_t)(p + 1 >= W - n);
      tk[p] = t;
      pm[p] = last ? 1 : t;
    }
    return;
  }
  // kBertTrain: inputs = positive_labels = window; token_mask[p] = (u[p] * pad[p]) >= mask_prob, then the two corner-case
  // fix-ups of Bert4RecUniformMasker.mask (all kept -> mask the last token; none kept -> un-mask the one before last)
  int64_t* lab = a.labels + (size_t)b * L;
  uint8_t* tk = a.aux_mask + (size_t)b * L;
  int all_kept = 1, any_kept = 0;
  for (int p = threadIdx.x; p < L; p += blockDim.x) {
    const int64_t v = window_item(a.items, first, n, W, p, a.pad_value);
    const bool real = p >= W - n;
    ids[p] = v;
    lab[p] = v;
    pm[p] = real;
    float u;
    if (a.uniforms) {
      u = a.uniforms[(size_t)b * L + p];
    } else {
      // uniform in [0,1) with 24 random bits, like torch.rand(float32); one Philox block per 4 positions of one draw
      const uint4 r = philox4x32(a.seed, (a.draw0 + (unsigned long long)b) * (unsigned long long)((L + 3) / 4) + (p >> 2));
      const uint32_t w = (p & 3) == 0 ? r.x : (p & 3) == 1 ? r.y : (p & 3) == 2 ? r.z : r.w;
      u = (float)(w >> 8) * (1.0f / 16777216.0f);
    }
    const bool keep = (u * (real ? 1.f : 0.f)) >= a.mask_prob;
    tk[p] = keep;
    all_kept &= keep ? 1 : 0;
    any_kept |= keep ? 1 : 0;
  }
  all_kept = __syncthreads_and(all_kept);
  any_kept = __syncthreads_or(any_kept);
  if (threadIdx.x == 0) {
    if (all_kept) tk[L - 1] = 0;
    else if (!any_kept && L > 1) tk[L - 2] = 1;
  }
}

}  // namespace rp

using namespace rp;

RP_API int rp_build_batch(const int64_t* offsets, const int32_t* items, long long n_seq, const int32_t* seq_index,
                          const int32_t* seq_offset, int B, int L, int mode, int pad_value, float mask_prob,
                          const float* uniforms, unsigned long long seed, unsigned long long draw0, const int64_t* query_ids,
                          int64_t* ids, uint8_t* pad_mask, int64_t* labels, uint8_t* aux_mask, int64_t* query_out,
                          void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!offsets || !items || !seq_index || !ids || !pad_mask || n_seq <= 0 || B < 0 || L <= 0) return RP_EINVAL;
  if (mode < kSasrecTrain || mode > kBertPredict) return RP_EINVAL;
  if ((mode == kSasrecTrain || mode == kBertTrain) && (!labels || !aux_mask)) return RP_EINVAL;
  if (mode == kBertPredict && !aux_mask) return RP_EINVAL;
  if (mode == kBertTrain && !(mask_prob >= 0.f)) return RP_EINVAL;
  if (B == 0) return RP_OK;
  BatchArgs a;
  a.offsets = offsets; a.items = items; a.seq_index = seq_index; a.seq_offset = seq_offset; a.query_ids = query_ids;
  a.uniforms = uniforms; a.ids = ids; a.pad_mask = pad_mask; a.labels = labels; a.aux_mask = aux_mask;
  a.query_out = query_out; a.n_seq = n_seq; a.B = B; a.L = L; a.mode = mode; a.pad_value = pad_value;
  a.mask_prob = mask_prob; a.seed = seed; a.draw0 = draw0;
  build_batch_kernel<<<B, 128, 0, stream>>>(a);
  RP_LAUNCH_CHECK();
  return RP_OK;
}
