"""TEST INFRASTRUCTURE - CPU restatement of the reference's per-sample input-layout producers.

Pure-Python loops, one sample at a time, exactly as the reference does it; pinned against the real reference classes by
``oracle/gen_golden.py::gen_dataset_layout`` -> ``tests/golden/dataset_layout.npz``.

Follows
  * TorchSequentialDataset._iter_with_window / __getitem__ / _pad_sequence / _generate_padding_mask
    (replay/data/nn/torch_sequential_dataset.py:69-171),
  * SasRecTrainingDataset.__getitem__ (replay/models/nn/sequential/sasrec/dataset.py:104-126),
  * Bert4RecUniformMasker.mask (replay/models/nn/sequential/bert4rec/dataset.py:71-92),
  * _shift_features / _shift_seq (replay/models/nn/sequential/bert4rec/dataset.py:322-351).
"""
from __future__ import annotations

import numpy as np


def window_index(lengths, window: int, sliding_window_step=None):
    """[(sequence_index, offset)] in the reference's iteration order (torch_sequential_dataset.py:154-171)."""
    out = []
    for i, n in enumerate(lengths):
        left = int(n) - window
        if sliding_window_step is not None:
            off = left
            while off > 0:
                out.append((i, off))
                off -= sliding_window_step
            out.append((i, 0))
        else:
            out.append((i, max(0, left)))
    return out


def padded_window(seq, offset: int, window: int, pad_value: int):
    """(ids [window] int64, mask [window] bool): seq[offset:offset+window] left-padded; the mask is built from the FULL
    history length like the reference (_generate_padding_mask, 98-106) - identical for every offset the index produces."""
    seq = np.asarray(seq, dtype=np.int64)
    cut = seq[offset:offset + window]
    ids = np.full(window, pad_value, dtype=np.int64)
    if len(cut):
        ids[window - len(cut):] = cut
    mask = np.ones(window, dtype=bool)
    if len(seq) < window:
        mask[:window - len(seq)] = False  # note: reference writes mask[:-len]; len == 0 leaves all ones there (never occurs)
    return ids, mask


def sasrec_training_sample(seq, offset: int, max_len: int, pad_value: int):
    ids, mask = padded_window(seq, offset, max_len + 1, pad_value)
    return {"item_id": ids[:-1], "padding_mask": mask[:-1], "positive_labels": ids[1:], "target_padding_mask": mask[1:]}


def prediction_sample(seq, max_len: int, pad_value: int):
    ids, mask = padded_window(seq, max(0, len(seq) - max_len), max_len, pad_value)
    return {"item_id": ids, "padding_mask": mask}


def bert_token_mask(pad_mask, uniforms, mask_prob: float):
    """Bert4RecUniformMasker.mask with the uniform draws made explicit (float32 arithmetic as in torch)."""
    m = (np.asarray(uniforms, dtype=np.float32) * np.asarray(pad_mask, dtype=np.float32)) >= np.float32(mask_prob)
    if m.all():
        m[-1] = False
    elif (not m.any()) and len(m) > 1:
        m[-2] = True
    return m


def bert_training_sample(seq, offset: int, max_len: int, pad_value: int, uniforms, mask_prob: float):
    ids, mask = padded_window(seq, offset, max_len, pad_value)
    return {"item_id": ids, "pad_mask": mask, "token_mask": bert_token_mask(mask, uniforms, mask_prob), "positive_labels": ids}


def bert_prediction_sample(seq, max_len: int, pad_value: int):
    ids, mask = padded_window(seq, max(0, len(seq) - max_len), max_len, pad_value)
    sh = np.roll(ids, -1)
    sh[-1] = pad_value
    tok = np.roll(mask, -1)
    tok[-1] = False
    pad = tok.copy()
    pad[-1] = True
    return {"item_id": sh, "pad_mask": pad, "token_mask": tok}
