"""Host-side input-layout producers with the reference's semantics: left padding, shift-by-one
labels (replay/models/nn/sequential/sasrec/dataset.py:104-126, replay/data/nn/torch_sequential_dataset.py:115-136) for
SASRec, uniform token masking for BERT4Rec (bert4rec/dataset.py:71-92).  Vectorised over a list of id sequences."""
from __future__ import annotations

import torch


def left_pad(sequences, length: int, pad_value: int):
    """[n] variable-length id lists -> (ids [n, length] int64, mask [n, length] bool): keep the LAST ``length`` items."""
    n = len(sequences)
    ids = torch.full((n, length), pad_value, dtype=torch.int64)
    mask = torch.zeros(n, length, dtype=torch.bool)
    for r, s in enumerate(sequences):
        s = torch.as_tensor(s, dtype=torch.int64)[-length:]
        k = s.numel()
        if k:
            ids[r, length - k:] = s
            mask[r, length - k:] = True
    return ids, mask


def sasrec_training_batch(sequences, max_len: int, pad_value: int, query_ids=None):
    """Batch dict with the legacy key names (sasrec/dataset.py:120-126)."""
    full, msk = left_pad(sequences, max_len + 1, pad_value)
    q = torch.arange(len(sequences)) if query_ids is None else torch.as_tensor(query_ids)
    return {"query_id": q.view(-1, 1), "feature_tensor": {"item_id": full[:, :-1].contiguous()},
            "padding_mask": msk[:, :-1].contiguous(), "positive_labels": full[:, 1:].contiguous(),
            "target_padding_mask": msk[:, 1:].contiguous()}


def sasrec_prediction_batch(sequences, max_len: int, pad_value: int, query_ids=None):
    ids, msk = left_pad(sequences, max_len, pad_value)
    q = torch.arange(len(sequences)) if query_ids is None else torch.as_tensor(query_ids)
    return {"query_id": q.view(-1, 1), "feature_tensor": {"item_id": ids}, "padding_mask": msk}


def to_new_path_batch(b: dict, with_seen: bool = True) -> dict:
    """legacy batch dict -> the new path's model inputs (make_default_sasrec_transforms, nn/transform/template/sasrec.py:9-42):
    feature_tensors, padding_mask, positive_labels [B,L,1], target_padding_mask [B,L,1] (+ seen_ids = the window)."""
    out = {"query_id": b["query_id"], "feature_tensors": b["feature_tensor"], "padding_mask": b["padding_mask"]}
    if "positive_labels" in b:
        out["positive_labels"] = b["positive_labels"].unsqueeze(-1)
        out["target_padding_mask"] = b["target_padding_mask"].unsqueeze(-1)
    if with_seen:
        out["seen_ids"] = b["feature_tensor"]["item_id"]
    return out


def balanced_rank_shards(work: torch.Tensor, world: int) -> torch.Tensor:
    """Deal the samples of ONE global batch to ``world`` data-parallel ranks so that every rank gets the same number of
    samples and (almost) the same amount of ``work`` (per-sample cost, e.g. the number of valid targets of a window: the
    full-catalog CE head costs one logit row per valid target, so ranks with longer windows are slower and - the gradient
    exchange being a barrier - every step runs at the pace of the slowest rank; MovieLens-shaped windows spread 2.5 % between
    ranks at 512 windows each, the slowest of 8 is 3.7 % above the mean).  Samples are sorted by work and dealt in
    boustrophedon order (0..W-1, W-1..0, ...).  ``work``: [n] with n divisible by world.  Returns int64 [world, n // world]:
    row r = sample indices of rank r.  The reference's samplers shard by index only (replay/data/nn/parquet/info/
    partitioning.py:102-122, torch DistributedSampler); dealing by length is what a length-grouped sampler does."""
    n = work.numel()
    if n % world != 0:
        raise ValueError(f"global batch of {n} samples is not divisible by {world} ranks")
    order = torch.argsort(work.reshape(-1), descending=True, stable=True)
    pos = torch.arange(n)
    r = pos % (2 * world)
    rank_of = torch.where(r < world, r, 2 * world - 1 - r)
    return torch.stack([order[rank_of == k] for k in range(world)])


def replica_partition(length: int, curr_replica: int, num_replicas: int, generator: torch.Generator | None = None,
                      device="cpu") -> torch.Tensor:
    """Row indices of one data-parallel replica, exactly as the reference's parquet reader assigns them
    (``Partitioning.generate``, replay/data/nn/parquet/info/partitioning.py:64-122): the index range is padded to a multiple
    of the replica count, optionally permuted with ``generator`` (a permutation of the PADDED range), replica r takes every
    ``num_replicas``-th entry starting at r, and padded indices wrap around modulo ``length`` - so the last replicas may see
    a few rows twice (``trainer.user_shard`` is the exact, duplicate-free partition used for predict()).
    Returns int64 [ceil(length / num_replicas)]."""
    if length < 1:
        raise ValueError(f"Length is invalid. Got {length}.")
    if num_replicas < 1:
        raise ValueError(f"Num Replicas is invalid. Got {num_replicas}.")
    if curr_replica < 0 or num_replicas <= curr_replica:
        raise ValueError(f"Curr Replicas is invalid. Got {curr_replica}.")
    per_replica = -(-length // num_replicas)
    full = per_replica * num_replicas
    if generator is None:
        raw = torch.arange(full, dtype=torch.int64, device=device)
    else:
        raw = torch.randperm(full, dtype=torch.int64, generator=generator).to(device)
    return torch.remainder(raw[curr_replica:full:num_replicas], length)
