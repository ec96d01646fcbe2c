"""MovieLens-shaped synthetic interaction sequences: log-normal history lengths, Zipf-like item popularity,
windows of the last L+1 items, LEFT padded; inputs = window[:-1], labels = window[1:]
(layout of SasRecTrainingDataset.__getitem__, replay/models/nn/sequential/sasrec/dataset.py:104-126)."""
from __future__ import annotations

import torch


def make_sequences(n_users: int, n_items: int, seq_len: int, seed: int = 1234, pad_value: int | None = None):
    """Returns (ids [U,L] int64, pad_mask [U,L] bool, labels [U,L] int64, target_mask [U,L] bool) on the CPU."""
    g = torch.Generator().manual_seed(seed)
    pad = n_items if pad_value is None else pad_value
    L = seq_len
    n_u = torch.exp(torch.randn(n_users, generator=g) * 0.95 + 4.56).round().clamp(20, 2314).long()
    ranks = torch.arange(n_items, dtype=torch.float64)
    prob = (ranks + 10.0) ** -0.8
    prob = prob[torch.randperm(n_items, generator=g)]
    win = torch.full((n_users, L + 1), pad, dtype=torch.int64)
    msk = torch.zeros(n_users, L + 1, dtype=torch.bool)
    keep = n_u.clamp(max=L + 1)
    total = int(keep.sum())
    items = torch.multinomial(prob, total, replacement=True, generator=g)
    col = torch.arange(L + 1).unsqueeze(0)
    real = col >= (L + 1 - keep).unsqueeze(1)
    win[real] = items
    msk[real] = True
    return win[:, :-1].contiguous(), msk[:, :-1].contiguous(), win[:, 1:].contiguous(), msk[:, 1:].contiguous()


def make_histories(n_users: int, n_items: int, seed: int = 1234):
    """Full (un-windowed) histories as CSR: (offsets int64 [U+1], items int64 [sum n_u]) on the CPU, same length and
    popularity model as ``make_sequences`` - the input of the device-side batch construction (replay_b200.device_data)."""
    g = torch.Generator().manual_seed(seed)
    n_u = torch.exp(torch.randn(n_users, generator=g) * 0.95 + 4.56).round().clamp(20, 2314).long()
    prob = (torch.arange(n_items, dtype=torch.float64) + 10.0) ** -0.8
    prob = prob[torch.randperm(n_items, generator=g)]
    offsets = torch.zeros(n_users + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(n_u, 0)
    items = torch.multinomial(prob, int(offsets[-1]), replacement=True, generator=g)
    return offsets, items
