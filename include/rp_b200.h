/* rp_b200.h - C ABI of librp_b200.so: the B200 (sm_100a) kernels behind RePlay's sequential-recommender hot path.
 *
 * The reference (sb-ai-lab/RePlay @ b4e051e8) has NO FFI on this path: its extension points are Python protocols and
 * Lightning hooks.  Each entry point below therefore cites the reference *Python* call it replaces;
 * INTEGRATION.md shows the ctypes stub a RePlay maintainer would add at that call site.
 *
 * Conventions (all functions):
 *   - caller owns all memory; pointers are device pointers unless the name says host; no allocation inside;
 *   - asynchronous with respect to the host, ordered on `stream` (a cudaStream_t / CUstream passed as void*);
 *   - scratch memory is passed in by the caller, sized by the matching *_workspace() function;
 *   - return value: 0 = ok, < 0 = argument / shape / alignment error (RP_E*), > 0 = a cudaError_t;
 *   - never throws, keeps no global mutable state besides cached driver entry points / device attributes;
 *   - bf16 tensors are row-major with 16-byte aligned base and row pitch.
 */
#ifndef RP_B200_H
#define RP_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RP_OK 0
#define RP_EINVAL (-1)     /* null pointer / unsupported flag */
#define RP_ESHAPE (-2)     /* unsupported size */
#define RP_EALIGN (-3)     /* pointer or pitch not 16-byte aligned */
#define RP_EDRIVER (-4)    /* CUDA driver entry point unavailable / tensor-map encode failed */
#define RP_EWORKSPACE (-5) /* workspace too small */

/* library / build info: returns a static string such as "rp_b200 0.1 sm_100a" */
const char* rp_version(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Predict head:  logits = hq . table^T  ->  seen-item mask  ->  top-K        (one fused pass, logits never stored)
 *   replaces  EmbeddingTyingHead.forward          replay/nn/head.py:29-34
 *                                                  replay/models/nn/sequential/sasrec/model.py:286-307 (legacy)
 *             SeenItemsFilter._compute_scores     replay/nn/lightning/postprocessor/seen_items.py:56-83
 *             RemoveSeenItems._compute_scores     replay/models/nn/sequential/postprocessors/postprocessors.py:55-95
 *             torch.topk(logits, k, dim=1)        replay/nn/lightning/callback/predictions_callback.py:90
 *                                                  replay/models/nn/sequential/callbacks/prediction_callbacks.py:93
 * ------------------------------------------------------------------------------------------------------------- */

/* seen_ids int64 [n_users, S] (any order, duplicates allowed, ids outside [0,item_count) are padding)
 *   -> out_sorted int32 [n_users, S], ascending, padding = INT32_MAX.
 * inv_map (optional, int32 [item_count]): position of each item in candidates_to_score, -1 if absent; when given the
 * output holds candidate positions instead of item ids (seen_items.py:68-71,80-81). */
int rp_seen_prepare(const int64_t* seen_ids, int n_users, int S, int item_count, const int32_t* inv_map,
                    int32_t* out_sorted, void* stream);

size_t rp_score_topk_workspace(int n_users, int n_items, int d, int K);

/* hq bf16 [n_users, d]; table bf16 [n_items, d] (the rows that are scored: all items, or the gathered candidates);
 * bias fp32 [n_items] or NULL; seen_sorted from rp_seen_prepare or NULL (no filter); candidates int64 [n_items] or NULL
 * (maps a scored column back to an item id, predictions_callback.py:91-92).
 * out_ids int64 [n_users, K], out_scores fp32 [n_users, K], sorted by (score desc, column asc).
 * d in {64,128,256,512}; 1 <= K <= 32. */
int rp_score_topk(const void* hq, const void* table, const float* bias, const int32_t* seen_sorted, int S, int n_users,
                  int n_items, int d, int K, const int64_t* candidates, int64_t* out_ids, float* out_scores,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Training head: full-catalog cross entropy fused with the logits GEMM, forward and backward
 *   replaces  logits = hidden . E^T                 replay/nn/head.py:29-34 ; replay/nn/sequential/sasrec/model.py:258-265
 *             torch.nn.CrossEntropyLoss (mean)       replay/nn/loss/ce.py:49-81
 *                                                    replay/models/nn/sequential/sasrec/lightning.py:335-355
 *                                                    replay/models/nn/sequential/bert4rec/lightning.py:332-351
 *             and autograd's backward of both.
 * hc bf16 [capacity, d]: hidden rows of the VALID targets, compacted (rows >= *n_valid are ignored but must be finite);
 * table bf16 [n_items, d]; labels int32 [capacity]; n_valid int32 [1] IN DEVICE MEMORY (keeps the step graph-capturable).
 * ------------------------------------------------------------------------------------------------------------- */
size_t rp_ce_head_workspace(int capacity_tokens, int n_items, int d);

/* loss_out fp32 [2] = { mean CE over the valid targets, 1 / n_valid }; lse fp32 [capacity];
 * cvec fp32 [round_up(capacity,128)] (per-token exponent offsets for the backward; entries >= capacity must be -inf). */
int rp_ce_head_fwd(const void* hc, const void* table, const int32_t* labels, const int32_t* n_valid, int capacity,
                   int n_items, int d, float* loss_out, float* lse, float* cvec, void* workspace, size_t workspace_bytes,
                   void* stream);

/* gradients of the mean CE for d(loss) = 1:  d_hc bf16 [capacity, d] (rows < *n_valid written);
 * d_table fp32 [n_items, d] is OVERWRITTEN (softmax part) and then atomically corrected by the one-hot part.
 * d in {64,128,256}. */
int rp_ce_head_bwd(const void* hc, const void* table, const int32_t* labels, const int32_t* n_valid, int capacity,
                   int n_items, int d, const float* loss_out, const float* cvec, void* d_hc, float* d_table, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Bring-up self test of the tcgen05 operand encodings (used by tests/, not by the product path).
 * A, B: bf16 [128,128]; D: fp32 [128,128].  mode bit0: B given as Bt[K,N]; bit1: A staged through TMEM;
 * bit2: A given as At[K,M].  D = A . B^T in every mode.
 * ------------------------------------------------------------------------------------------------------------- */
int rp_selftest_umma(int mode, const void* A, const void* B, float* D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RP_B200_H */
