"""GPU tests of device-side batch construction (rp_build_batch): bit-exact against samples produced by
the reference's own dataset classes (tests/golden/dataset_layout.npz), against the loop restatement (oracle/dataset.py) on
random histories, and through size-independent properties at full batch size."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda")


def _layout(golden_dir):
    z = np.load(os.path.join(golden_dir, "dataset_layout.npz"))
    off = np.concatenate([[0], np.cumsum(z["lengths"])])
    return z, [z["items"][off[i]:off[i + 1]] for i in range(len(z["lengths"]))]


def _np(t):
    return t.cpu().numpy()


def test_golden_samples_of_the_reference_datasets(golden_dir, cuda):
    from replay_b200.device_data import DeviceSequenceStore, window_index
    z, seqs = _layout(golden_dir)
    L, pad, step, prob = int(z["L"]), int(z["pad"]), int(z["step"]), float(z["mask_prob"])
    st = DeviceSequenceStore(seqs, query_ids=z["query_ids"])
    for tag, sw in (("slide", step), ("last", None)):
        s, o = window_index(st.lengths, L + 1, sw)
        b = st.sasrec_training_batch(s, L, pad, seq_offset=o)
        assert np.array_equal(_np(b["feature_tensor"]["item_id"]), z[f"sas_{tag}_ids"])
        assert np.array_equal(_np(b["padding_mask"]), z[f"sas_{tag}_pad"])
        assert np.array_equal(_np(b["positive_labels"]), z[f"sas_{tag}_labels"])
        assert np.array_equal(_np(b["target_padding_mask"]), z[f"sas_{tag}_tmask"])
        assert np.array_equal(_np(b["query_id"])[:, 0], z[f"sas_{tag}_query"])
        assert b["feature_tensor"]["item_id"].dtype == torch.int64 and b["padding_mask"].dtype == torch.bool
        s, o = window_index(st.lengths, L, sw)
        bb = st.bert4rec_training_batch(s, L, pad, prob, seq_offset=o, uniforms=z[f"bert_{tag}_uniforms"])
        assert np.array_equal(_np(bb["inputs"]["item_id"]), z[f"bert_{tag}_ids"])
        assert np.array_equal(_np(bb["pad_mask"]), z[f"bert_{tag}_pad"])
        assert np.array_equal(_np(bb["token_mask"]), z[f"bert_{tag}_tok"])
        assert np.array_equal(_np(bb["positive_labels"]), z[f"bert_{tag}_labels"])
    # default offset (no seq_offset) = the reference's "last window" index
    b = st.sasrec_training_batch(np.arange(len(seqs)), L, pad)
    assert np.array_equal(_np(b["feature_tensor"]["item_id"]), z["sas_last_ids"])
    p = st.sasrec_prediction_batch(np.arange(len(seqs)), L, pad)
    assert np.array_equal(_np(p["feature_tensor"]["item_id"]), z["pred_ids"])
    assert np.array_equal(_np(p["padding_mask"]), z["pred_pad"])
    bp = st.bert4rec_prediction_batch(np.arange(len(seqs)), L, pad)
    assert np.array_equal(_np(bp["inputs"]["item_id"]), z["bertpred_ids"])
    assert np.array_equal(_np(bp["pad_mask"]), z["bertpred_pad"])
    assert np.array_equal(_np(bp["token_mask"]), z["bertpred_tok"])
    # new path: Array1DColumn gather + NextTokenTransform of the reference on an arbitrary (repeating) row order
    nb = st.sasrec_new_path_batch(z["newpath_order"], L, pad)
    assert np.array_equal(_np(nb["feature_tensors"]["item_id"]), z["newpath_ids"])
    assert np.array_equal(_np(nb["padding_mask"]), z["newpath_pad"])
    assert np.array_equal(_np(nb["positive_labels"])[..., 0], z["newpath_labels"])
    assert np.array_equal(_np(nb["target_padding_mask"])[..., 0], z["newpath_tmask"])
    assert np.array_equal(_np(nb["seen_ids"]), z["newpath_ids"])
    # masker corner cases (draw independent)
    for tag, pr in (("p0", 0.0), ("p2", 2.0)):
        bb = st.bert4rec_training_batch(np.arange(len(seqs)), L, pad, pr, seed=5)
        assert np.array_equal(_np(bb["token_mask"]), z[f"bert_{tag}_tok"])


@pytest.mark.parametrize("L,step", [(50, None), (200, 37), (1, 1), (512, 100)])
def test_random_histories_against_loop_restatement(cuda, L, step):
    from oracle import dataset as od
    from replay_b200.device_data import DeviceSequenceStore, window_index
    rng = np.random.default_rng(L)
    n_items = 100_000
    lens = np.clip(np.round(np.exp(rng.normal(4.56, 0.95, 600))), 0, 2314).astype(np.int64)
    lens[:4] = [0, 1, L, L + 1]
    seqs = [rng.integers(0, n_items, n) for n in lens]
    st = DeviceSequenceStore(seqs)
    s, o = window_index(lens, L + 1, step)
    b = st.sasrec_training_batch(s, L, n_items, seq_offset=o)
    ids, pm, lab, tm = (_np(b["feature_tensor"]["item_id"]), _np(b["padding_mask"]), _np(b["positive_labels"]),
                        _np(b["target_padding_mask"]))
    pick = rng.choice(len(s), size=min(len(s), 300), replace=False)
    for r in pick:
        if lens[s[r]] == 0:
            continue  # the reference's mask[:-0] quirk; empty histories never reach it (tokenizer drops them)
        ref = od.sasrec_training_sample(seqs[s[r]], int(o[r]), L, n_items)
        assert np.array_equal(ids[r], ref["item_id"]) and np.array_equal(pm[r], ref["padding_mask"])
        assert np.array_equal(lab[r], ref["positive_labels"]) and np.array_equal(tm[r], ref["target_padding_mask"])
    # properties over ALL rows: left padding, shift-by-one, pad positions hold the padding value
    assert (np.diff(pm.astype(np.int8), axis=1) >= 0).all() and (np.diff(tm.astype(np.int8), axis=1) >= 0).all()
    if L > 1:
        assert np.array_equal(ids[:, 1:], lab[:, :-1]) and np.array_equal(pm[:, 1:], tm[:, :-1])
    assert (ids[~pm] == n_items).all() and (lab[~tm] == n_items).all() and (ids[pm] < n_items).all()
    assert np.array_equal(tm.sum(1), np.minimum(lens[s] - o, L + 1).clip(0, L))
    # empty history: everything padded
    e = np.where(lens[s] == 0)[0]
    assert (~pm[e]).all() and (~tm[e]).all()
    # BERT with explicit uniforms against the loop restatement
    s2, o2 = window_index(lens, L, step)
    u = rng.random((len(s2), L), dtype=np.float32)
    bb = st.bert4rec_training_batch(s2, L, n_items, 0.15, seq_offset=o2, uniforms=u)
    tok, bpm = _np(bb["token_mask"]), _np(bb["pad_mask"])
    for r in rng.choice(len(s2), size=min(len(s2), 300), replace=False):
        if lens[s2[r]] == 0:
            continue
        ref = od.bert_training_sample(seqs[s2[r]], int(o2[r]), L, n_items, u[r], 0.15)
        assert np.array_equal(tok[r], ref["token_mask"]) and np.array_equal(bpm[r], ref["pad_mask"])
        assert np.array_equal(_np(bb["inputs"]["item_id"])[r], ref["item_id"])


def test_bert_masker_philox_stream(cuda):
    from replay_b200.device_data import DeviceSequenceStore
    rng = np.random.default_rng(1)
    L, n_items, n = 200, 1000, 4096
    seqs = [rng.integers(0, n_items, 250) for _ in range(n)]
    st = DeviceSequenceStore(seqs)
    idx = np.arange(n)
    a = st.bert4rec_training_batch(idx, L, n_items, 0.15, seed=9, draw0=100)
    b = st.bert4rec_training_batch(idx, L, n_items, 0.15, seed=9, draw0=100)
    c = st.bert4rec_training_batch(idx, L, n_items, 0.15, seed=9, draw0=101)
    d = st.bert4rec_training_batch(idx, L, n_items, 0.15, seed=10, draw0=100)
    ta, tb, tc, td = (_np(x["token_mask"]) for x in (a, b, c, d))
    assert np.array_equal(ta, tb)                       # counter based: same (seed, draw) -> same mask
    assert np.array_equal(ta[1:], tc[:-1])              # draw index = draw0 + row
    assert (ta != td).mean() > 0.1
    frac = 1.0 - ta.mean()                              # masked share of real tokens (all rows are full)
    assert abs(frac - 0.15) < 0.004, frac
    per_pos = 1.0 - ta.mean(0)
    assert abs(per_pos - 0.15).max() < 0.03             # no positional bias
    assert ((~ta).sum(1) >= 1).all()                    # every row has at least one masked token


def test_loader_epoch_covers_every_window_once_and_trains(cuda):
    """DeviceBatchLoader: two ranks' shards together visit every window exactly once (up to the wrap-around padding), and
    its batches drive the legacy SasRec module's fused training_step."""
    from replay_b200.device_data import DeviceBatchLoader, DeviceSequenceStore
    from replay_b200.models.nn.sequential import SasRec
    from replay_b200.schema import TensorFeatureInfo, TensorSchema
    rng = np.random.default_rng(2)
    n_items, L, U = 500, 32, 301
    seqs = []
    for _ in range(U):
        n, s0 = int(rng.integers(5, 60)), int(rng.integers(0, n_items))
        seqs.append((s0 + np.arange(n)) % n_items)
    st = DeviceSequenceStore(seqs, query_ids=np.arange(U) + 7)
    seen = []
    for rank in range(2):
        ld = DeviceBatchLoader(st, L, 64, n_items, sliding_window_step=8, seed=3, rank=rank, world_size=2)
        ld.set_epoch(1)
        for b in ld:
            win = torch.cat([b["feature_tensor"]["item_id"], b["positive_labels"][:, -1:]], 1)
            seen.append(torch.cat([b["query_id"], win], 1).cpu())
    seen = torch.cat(seen)
    ld1 = DeviceBatchLoader(st, L, 10_000, n_items, sliding_window_step=8, shuffle=False)
    allw = next(iter(ld1))
    allw = torch.cat([allw["query_id"], allw["feature_tensor"]["item_id"], allw["positive_labels"][:, -1:]], 1).cpu()
    assert len(seen) in (len(allw), len(allw) + 1)
    assert {tuple(r) for r in seen.tolist()} == {tuple(r) for r in allw.tolist()}
    # drive the fused training step of the legacy module with loader batches
    model = SasRec(TensorSchema(TensorFeatureInfo("item_id", n_items, 0, 64)), block_count=1, head_count=1, hidden_size=64,
                   max_seq_len=L, dropout_rate=0.0)
    ld = DeviceBatchLoader(st, L, 64, 0, seed=1)
    losses = []
    for ep in range(8):
        ld.set_epoch(ep)
        for i, b in enumerate(ld):
            losses.append(float(model.training_step(b, i)))
    assert losses[-1] < losses[0] - 1.0, (losses[0], losses[-1])
