"""Device-resident sequence store and batch construction.

All user histories are uploaded ONCE as a CSR store (offsets int64 + item ids int32); every batch is then cut, left-padded,
shifted and masked by one CUDA launch (``rp_build_batch``, csrc/rp_batch.cu) instead of the reference's per-sample host
path (``TorchSequentialDataset.__getitem__`` -> ``SasRecTrainingDataset.__getitem__`` / ``Bert4RecTrainingDataset`` ->
default collate -> H2D copy; replay/data/nn/torch_sequential_dataset.py:69-171, sasrec/dataset.py:104-126,
bert4rec/dataset.py:71-92,163-177,322-351).  The produced batch dictionaries carry the reference's key names, dtypes and
shapes, so they feed ``training_step`` / ``predict_step`` of the modules in ``replay_b200.models`` / ``replay_b200.nn``
unchanged.  There is no CPU fallback: building a batch needs the CUDA library."""
from __future__ import annotations

import numpy as np
import torch

from ._lib import check, lib

SASREC_TRAIN, PREDICT, BERT_TRAIN, BERT_PREDICT = 0, 1, 2, 3


def window_index(lengths, window: int, sliding_window_step: int | None = None):
    """(sequence_index int32 [n], offset int32 [n]) in the reference's iteration order
    (TorchSequentialDataset._iter_with_window, torch_sequential_dataset.py:154-171), vectorised: without a step one
    window per history at offset max(0, len - window); with a step the offsets len-window, len-window-step, ... (> 0)
    followed by 0."""
    lengths = np.asarray(lengths, dtype=np.int64)
    left = lengths - window
    if sliding_window_step is None:
        return np.arange(len(lengths), dtype=np.int32), np.maximum(left, 0).astype(np.int32)
    step = int(sliding_window_step)
    if step <= 0:
        raise ValueError("sliding_window_step must be positive")
    extra = np.where(left > 0, (left + step - 1) // step, 0)  # windows with a positive offset
    counts = extra + 1
    seq = np.repeat(np.arange(len(lengths), dtype=np.int64), counts)
    starts = np.cumsum(counts) - counts
    k = np.arange(counts.sum(), dtype=np.int64) - np.repeat(starts, counts)  # 0.. within one history
    off = np.repeat(left, counts) - k * step
    off = np.where(k == np.repeat(extra, counts), 0, off)  # the closing (i, 0) window
    return seq.astype(np.int32), off.astype(np.int32)


class DeviceSequenceStore:
    """CSR store of item-id histories in HBM.  ``sequences``: list of 1-D integer arrays (one per query, item ids already
    label-encoded to 0..|I|-1 as the reference's SequenceTokenizer does), or pass ``offsets``/``items`` directly."""

    def __init__(self, sequences=None, *, offsets=None, items=None, query_ids=None, device="cuda"):
        if sequences is not None:
            lens = np.fromiter((len(s) for s in sequences), dtype=np.int64, count=len(sequences))
            offsets = np.zeros(len(lens) + 1, dtype=np.int64)
            np.cumsum(lens, out=offsets[1:])
            items = np.concatenate([np.asarray(s, dtype=np.int64) for s in sequences]) if len(lens) else np.zeros(0, np.int64)
        offsets = np.asarray(offsets, dtype=np.int64)
        items = np.asarray(items)
        if offsets.ndim != 1 or len(offsets) < 2 or offsets[0] != 0 or offsets[-1] != len(items) or np.any(np.diff(offsets) < 0):
            raise ValueError("offsets must be a non-decreasing CSR row pointer starting at 0 and ending at len(items)")
        if len(items) and (items.min() < 0 or items.max() >= 2 ** 31 - 1):
            raise ValueError("item ids must fit int32")
        self.device = torch.device(device)
        self.lengths = np.diff(offsets)
        self.n_seq = len(self.lengths)
        self.offsets = torch.from_numpy(np.array(offsets, dtype=np.int64)).to(self.device)
        self.items = torch.from_numpy(items.astype(np.int32)).to(self.device)
        if len(items) == 0:  # keep a valid pointer
            self.items = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.query_ids = None if query_ids is None else torch.from_numpy(np.array(query_ids, dtype=np.int64)).to(self.device)

    @classmethod
    def from_sequential_dataset(cls, sequential, feature_name: str | None = None, device="cuda"):
        """Duck-typed ``SequentialDataset`` (replay/data/nn/sequential_dataset.py:18-105): ``__len__``, ``get_query_id``,
        ``get_sequence`` and ``schema.item_id_feature_name``."""
        name = feature_name or sequential.schema.item_id_feature_name
        n = len(sequential)
        return cls([np.asarray(sequential.get_sequence(i, name)) for i in range(n)],
                   query_ids=[sequential.get_query_id(i) for i in range(n)], device=device)

    @classmethod
    def from_parquet(cls, source, item_column: str = "item_id", query_column: str | None = None, device="cuda"):
        """Sequence-per-row parquet (the layout the reference's ParquetDataset / ParquetModule reads: one row per query, the
        item ids in a list<int> column; replay/data/nn/parquet/impl/array_1d_column.py:87-140): the list column's offsets and
        flat values become the CSR store without a Python loop (null lists count as empty).  ``source``: a path, a list of
        paths or a ``pyarrow.Table``."""
        import pyarrow as pa
        import pyarrow.compute as pc
        import pyarrow.parquet as pq

        cols = [item_column] + ([query_column] if query_column else [])
        if isinstance(source, pa.Table):
            table = source.select(cols)
        elif isinstance(source, (list, tuple)):
            table = pa.concat_tables([pq.read_table(p, columns=cols) for p in source])
        else:
            table = pq.read_table(source, columns=cols)
        col = table.column(item_column).combine_chunks()
        if not (pa.types.is_list(col.type) or pa.types.is_large_list(col.type)):
            raise ValueError(f"column {item_column!r} must be a list column, got {col.type}")
        lengths = pc.fill_null(pc.list_value_length(col), 0).to_numpy(zero_copy_only=False).astype(np.int64)
        values = pc.list_flatten(col)
        if values.null_count:
            raise ValueError(f"column {item_column!r} holds null item ids")
        offsets = np.zeros(len(lengths) + 1, dtype=np.int64)
        np.cumsum(lengths, out=offsets[1:])
        q = table.column(query_column).combine_chunks().to_numpy(zero_copy_only=False) if query_column else None
        return cls(offsets=offsets, items=values.to_numpy(zero_copy_only=False), query_ids=q, device=device)

    def __len__(self):
        return self.n_seq

    # --------------------------------------------------------------------------------------------- batch builders
    def _build(self, mode, seq_index, seq_offset, L, pad_value, *, mask_prob=0.0, uniforms=None, seed=0, draw0=0,
               with_labels=False, with_aux=False):
        dev = self.device
        seq_index = torch.as_tensor(seq_index, device=dev).to(torch.int32).contiguous()
        B = seq_index.numel()
        if seq_offset is not None:
            seq_offset = torch.as_tensor(seq_offset, device=dev).to(torch.int32).contiguous()
            if seq_offset.numel() != B:
                raise ValueError("seq_offset must have one entry per batch row")
        ids = torch.empty(B, L, dtype=torch.int64, device=dev)
        pad = torch.empty(B, L, dtype=torch.bool, device=dev)
        labels = torch.empty(B, L, dtype=torch.int64, device=dev) if with_labels else None
        aux = torch.empty(B, L, dtype=torch.bool, device=dev) if with_aux else None
        q = torch.empty(B, dtype=torch.int64, device=dev)
        if uniforms is not None:
            uniforms = torch.as_tensor(uniforms, device=dev, dtype=torch.float32).contiguous()
            if tuple(uniforms.shape) != (B, L):
                raise ValueError("uniforms must be [B, L]")
        p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        check(lib().rp_build_batch(self.offsets.data_ptr(), self.items.data_ptr(), self.n_seq, seq_index.data_ptr(),
                                   p(seq_offset), B, L, mode, int(pad_value), float(mask_prob), p(uniforms), int(seed),
                                   int(draw0), p(self.query_ids), ids.data_ptr(), pad.data_ptr(), p(labels), p(aux),
                                   q.data_ptr(), torch.cuda.current_stream().cuda_stream), "rp_build_batch")
        return ids, pad, labels, aux, q.view(-1, 1)

    def sasrec_training_batch(self, seq_index, max_len: int, pad_value: int, seq_offset=None, feature_name="item_id"):
        """Reference keys (sasrec/dataset.py:120-126): query_id [B,1], feature_tensor{item_id [B,L]}, padding_mask,
        positive_labels, target_padding_mask - all on the device."""
        ids, pad, labels, tmask, q = self._build(SASREC_TRAIN, seq_index, seq_offset, max_len, pad_value, with_labels=True,
                                                 with_aux=True)
        return {"query_id": q, "feature_tensor": {feature_name: ids}, "padding_mask": pad, "positive_labels": labels,
                "target_padding_mask": tmask}

    def sasrec_new_path_batch(self, seq_index, max_len: int, pad_value: int, feature_name="item_id", with_seen: bool = True):
        """The new path's model inputs: Array1DColumn.__getitem__ with shape max_len + 1 (left-padded gather of the last
        elements, parquet/impl/indexing.py:42-78) + NextTokenTransform(shift=1) (nn/transform/next_token.py:65-96) +
        the unsqueeze of the default SASRec transform template -> feature_tensors, padding_mask, positive_labels [B,L,1],
        target_padding_mask [B,L,1] (+ seen_ids = the window)."""
        ids, pad, labels, tmask, q = self._build(SASREC_TRAIN, seq_index, None, max_len, pad_value, with_labels=True,
                                                 with_aux=True)
        out = {"query_id": q, "feature_tensors": {feature_name: ids}, "padding_mask": pad,
               "positive_labels": labels.unsqueeze(-1), "target_padding_mask": tmask.unsqueeze(-1)}
        if with_seen:
            out["seen_ids"] = ids
        return out

    def sasrec_prediction_batch(self, seq_index, max_len: int, pad_value: int, feature_name="item_id"):
        ids, pad, _, _, q = self._build(PREDICT, seq_index, None, max_len, pad_value)
        return {"query_id": q, "padding_mask": pad, "feature_tensor": {feature_name: ids}}

    def bert4rec_training_batch(self, seq_index, max_len: int, pad_value: int, mask_prob: float = 0.15, seq_offset=None,
                                seed: int = 0, draw0: int = 0, uniforms=None, feature_name="item_id"):
        """Reference keys (bert4rec/dataset.py:167-173).  ``token_mask`` False = masked.  Random draws: Philox keyed by
        (seed, draw0 + row); pass ``uniforms`` [B, L] to reproduce a given masker stream exactly."""
        ids, pad, labels, tok, q = self._build(BERT_TRAIN, seq_index, seq_offset, max_len, pad_value, mask_prob=mask_prob,
                                               uniforms=uniforms, seed=seed, draw0=draw0, with_labels=True, with_aux=True)
        return {"query_id": q, "pad_mask": pad, "inputs": {feature_name: ids}, "token_mask": tok, "positive_labels": labels}

    def bert4rec_prediction_batch(self, seq_index, max_len: int, pad_value: int, feature_name="item_id"):
        ids, pad, _, tok, q = self._build(BERT_PREDICT, seq_index, None, max_len, pad_value, with_aux=True)
        return {"query_id": q, "pad_mask": pad, "inputs": {feature_name: ids}, "token_mask": tok}


class DeviceBatchLoader:
    """Iterates device-built training batches the way ``DataLoader(SasRecTrainingDataset(...), shuffle=True)`` +
    ``DistributedSampler`` would: the window index is built once (host, vectorised), permuted per epoch with a seeded
    generator shared by all ranks, padded by wrap-around to a multiple of the world size (DistributedSampler semantics) and
    strided over the ranks; every batch is then one kernel launch on HBM-resident data."""

    def __init__(self, store: DeviceSequenceStore, max_len: int, batch_size: int, pad_value: int, kind: str = "sasrec",
                 sliding_window_step: int | None = None, shuffle: bool = True, drop_last: bool = False, seed: int = 0,
                 rank: int = 0, world_size: int = 1, mask_prob: float = 0.15, partitioning: str = "sampler"):
        if kind not in ("sasrec", "bert4rec"):
            raise ValueError(f"unknown kind {kind!r}")
        if partitioning not in ("sampler", "replay"):
            raise ValueError(f"unknown partitioning {partitioning!r}")
        self.partitioning = partitioning
        self.store, self.L, self.bs, self.pad, self.kind = store, int(max_len), int(batch_size), int(pad_value), kind
        window = self.L + (1 if kind == "sasrec" else 0)
        seq, off = window_index(store.lengths, window, sliding_window_step)
        self.win_seq = torch.from_numpy(seq).to(store.device)
        self.win_off = torch.from_numpy(off).to(store.device)
        self.n = len(seq)
        self.shuffle, self.drop_last, self.seed, self.rank, self.world = shuffle, drop_last, int(seed), rank, world_size
        self.mask_prob = float(mask_prob)
        self.epoch = 0
        self.per_rank = -(-self.n // world_size)

    def set_epoch(self, epoch: int):
        self.epoch = int(epoch)

    def __len__(self):
        return self.per_rank // self.bs if self.drop_last else -(-self.per_rank // self.bs)

    def epoch_indices(self) -> torch.Tensor:
        """This rank's window indices for the current epoch (on the store's device): seeded permutation shared by all ranks,
        wrap-around padding to a multiple of the world size, strided over the ranks (DistributedSampler semantics).
        ``partitioning="replay"``: the reference parquet reader's assignment instead (``replay_b200.data.replica_partition`` =
        ``Partitioning.generate``, replay/data/nn/parquet/info/partitioning.py:64-122: permutation of the PADDED range, modulo)."""
        dev = self.store.device
        if getattr(self, "partitioning", "sampler") == "replay":
            from .data import replica_partition

            g = torch.Generator(device="cpu").manual_seed(self.seed + self.epoch) if self.shuffle else None
            return replica_partition(self.n, self.rank, self.world, generator=g, device=dev)
        if self.shuffle:
            g = torch.Generator(device="cpu").manual_seed(self.seed + self.epoch)
            order = torch.randperm(self.n, generator=g).to(dev)
        else:
            order = torch.arange(self.n, device=dev)
        total = self.per_rank * self.world
        if total > self.n:  # wrap-around padding
            order = torch.cat([order, order[: total - self.n]])
        return order[self.rank: total: self.world]

    def __iter__(self):
        mine = self.epoch_indices()
        for i in range(len(self)):
            idx = mine[i * self.bs: (i + 1) * self.bs]
            s, o = self.win_seq[idx], self.win_off[idx]
            if self.kind == "sasrec":
                yield self.store.sasrec_training_batch(s, self.L, self.pad, seq_offset=o)
            else:
                draw0 = (self.epoch * self.per_rank * self.world) + self.rank * self.per_rank + i * self.bs
                yield self.store.bert4rec_training_batch(s, self.L, self.pad, self.mask_prob, seq_offset=o,
                                                         seed=self.seed, draw0=draw0)
