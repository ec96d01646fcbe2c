"""Plain-torch CPU restatement of RePlay's SASRec hot path (TEST INFRASTRUCTURE - see oracle/__init__.py).

Canonical parameter dict ``P`` (the layout ``replay_b200`` uses too)::

    item_emb [I+1, d]   row ``pad_id`` (= I for the recommended schema) is the padding row
    pos_emb  [Lmax, d]
    blocks   list of dicts: ln1_w ln1_b in_w[3d,d] in_b[3d] out_w[d,d] out_b[d] ln2_w ln2_b w1[d,d] b1 w2[d,d] b2
    lnf_w lnf_b         final LayerNorm
Everything is computed in the dtype of the parameters handed in (fp32 for parity with the reference, fp64 to adjudicate).

Reference files restated (all under /root/reference/replay):
  new path   nn/sequential/sasrec/model.py:85-113,258-307 ; nn/sequential/sasrec/transformer.py:74-110 ;
             nn/sequential/sasrec/agg.py:37-53 ; nn/mask.py:18-51 ; nn/ffn.py:43-57 ; nn/head.py:16-34 ;
             nn/embedding.py:176-221 ; nn/loss/ce.py:49-81
  legacy     models/nn/sequential/sasrec/model.py:159-180,216-245,346-357,419-442,465-473,286-307,496-506 ;
             models/nn/sequential/sasrec/lightning.py:335-355
  predict    nn/lightning/postprocessor/seen_items.py:56-83 ; nn/lightning/callback/predictions_callback.py:80-96 ;
             models/nn/sequential/postprocessors/postprocessors.py:55-95
  optimizer  models/nn/optimizer_utils/optimizer_factory.py:71-87 (torch.optim.Adam lr 1e-3 betas (0.9,0.98))
"""
from __future__ import annotations

import math

import torch

# ----------------------------------------------------------------------------------------------------------------------
# parameter conversion from the reference's state_dict key names
# ----------------------------------------------------------------------------------------------------------------------


def _blocks_from_sd(sd, prefix, n):
    blocks = []
    for i in range(n):
        a = f"{prefix}attention_layers.{i}."
        blocks.append(
            {
                "ln1_w": sd[f"{prefix}attention_layernorms.{i}.weight"],
                "ln1_b": sd[f"{prefix}attention_layernorms.{i}.bias"],
                "in_w": sd[a + "in_proj_weight"],
                "in_b": sd[a + "in_proj_bias"],
                "out_w": sd[a + "out_proj.weight"],
                "out_b": sd[a + "out_proj.bias"],
                "ln2_w": sd[f"{prefix}forward_layernorms.{i}.weight"],
                "ln2_b": sd[f"{prefix}forward_layernorms.{i}.bias"],
                "w1": sd[f"{prefix}forward_layers.{i}.conv1.weight"][:, :, 0],
                "b1": sd[f"{prefix}forward_layers.{i}.conv1.bias"],
                "w2": sd[f"{prefix}forward_layers.{i}.conv2.weight"][:, :, 0],
                "b2": sd[f"{prefix}forward_layers.{i}.conv2.bias"],
            }
        )
    return blocks


def _count_blocks(sd, prefix):
    n = 0
    while f"{prefix}attention_layers.{n}.in_proj_weight" in sd:
        n += 1
    return n


def params_from_new_state_dict(sd, item_feature="item_id"):
    """``replay.nn.sequential.SasRec.state_dict()`` -> canonical dict (keys: SURVEY Appendix B, new path)."""
    sd = {k: v.detach().clone() for k, v in sd.items()}
    n = _count_blocks(sd, "body.encoder.")
    return {
        "item_emb": sd[f"body.embedder.feature_embedders.{item_feature}.emb.weight"],
        "pos_emb": sd["body.embedding_aggregator.pe.weight"],
        "blocks": _blocks_from_sd(sd, "body.encoder.", n),
        "lnf_w": sd["body.output_normalization.weight"],
        "lnf_b": sd["body.output_normalization.bias"],
    }


def params_from_legacy_state_dict(sd):
    """legacy ``SasRecModel.state_dict()`` -> canonical dict (keys: SURVEY Appendix B, legacy)."""
    sd = {k: v.detach().clone() for k, v in sd.items()}
    n = _count_blocks(sd, "sasrec_layers.")
    return {
        "item_emb": sd["item_embedder.item_emb.weight"],
        "pos_emb": sd["item_embedder.pos_emb.pe.weight"],
        "blocks": _blocks_from_sd(sd, "sasrec_layers.", n),
        "lnf_w": sd["output_normalization.last_layernorm.weight"],
        "lnf_b": sd["output_normalization.last_layernorm.bias"],
    }


def params_to(P, dtype):
    out = {}
    for k, v in P.items():
        if k == "blocks":
            out[k] = [{kk: vv.to(dtype) for kk, vv in b.items()} for b in v]
        else:
            out[k] = v.to(dtype)
    return out


def random_params(n_items, d, l_max, n_blocks, seed=0, dtype=torch.float32, bias_scale=0.02):
    """Reference-style init (xavier-normal on >=2-D) with *randomised* biases/LN so tests see them."""
    g = torch.Generator().manual_seed(seed)

    def xn(*shape):
        fan_out, fan_in = shape[0], shape[1]
        std = math.sqrt(2.0 / (fan_in + fan_out))
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    def small(n):
        return torch.randn(n, generator=g, dtype=torch.float32) * bias_scale

    item = xn(n_items + 1, d)
    item[n_items].zero_()  # nn/embedding.py:198-200 zero-fills the pad row
    P = {"item_emb": item, "pos_emb": xn(l_max, d), "blocks": [], "lnf_w": 1 + small(d), "lnf_b": small(d)}
    for _ in range(n_blocks):
        P["blocks"].append(
            {
                "ln1_w": 1 + small(d), "ln1_b": small(d),
                "in_w": xn(3 * d, d), "in_b": small(3 * d),
                "out_w": xn(d, d), "out_b": small(d),
                "ln2_w": 1 + small(d), "ln2_b": small(d),
                "w1": xn(d, d), "b1": small(d), "w2": xn(d, d), "b2": small(d),
            }
        )
    return params_to(P, dtype)


# ----------------------------------------------------------------------------------------------------------------------
# body
# ----------------------------------------------------------------------------------------------------------------------


def layer_norm(x, w, b, eps):
    mean = x.mean(-1, keepdim=True)
    var = x.var(-1, unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * w + b


def mha(q_in, kv_in, blk, n_heads, visible, uniform_rows=None):
    """torch.nn.MultiheadAttention as configured at nn/sequential/sasrec/transformer.py:36-46 (dropout off).

    ``visible`` [B, L, L] bool: key j may be attended by query i.  Rows with no visible key give a zero
    attention output (torch>=2.5 "safe softmax" on an all -inf row; verified on torch 2.11).
    """
    B, L, d = q_in.shape
    hd = d // n_heads
    wq, wk, wv = blk["in_w"][:d], blk["in_w"][d : 2 * d], blk["in_w"][2 * d :]
    bq, bk, bv = blk["in_b"][:d], blk["in_b"][d : 2 * d], blk["in_b"][2 * d :]
    q = (q_in @ wq.T + bq).view(B, L, n_heads, hd).transpose(1, 2)
    k = (kv_in @ wk.T + bk).view(B, L, n_heads, hd).transpose(1, 2)
    v = (kv_in @ wv.T + bv).view(B, L, n_heads, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if uniform_rows is not None:  # eval-mode pad query rows: every unmasked score is finfo.min -> uniform weights
        s = torch.where(uniform_rows[:, None, :, None], torch.zeros_like(s), s)
    s = s.masked_fill(~visible[:, None], float("-inf"))
    m = s.max(-1, keepdim=True).values
    m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    e = torch.exp(s - m)
    den = e.sum(-1, keepdim=True)
    p = torch.where(den > 0, e / den.clamp_min(1e-300), torch.zeros_like(e))
    o = (p @ v).transpose(1, 2).reshape(B, L, d)
    return o @ blk["out_w"].T + blk["out_b"]


def sasrec_body(P, ids, pad_mask, n_heads, variant="new", lnf_eps=None, return_all=False, mode="train"):
    """Hidden states [B, L, d] of the SASRec body with dropout off.

    variant "new":    nn/sequential/sasrec/model.py:85-113 in *training-mask* semantics (-inf masks): pad keys are
                      masked, pad query rows get attention output = out_proj.bias; real rows are identical in eval.
    variant "legacy": models/nn/sequential/sasrec/model.py:159-180: causal mask only, pad rows zeroed after the
                      embedding and after every block.
    ``ids`` int64 [B, L] (pad positions hold any id; they are replaced by the pad row), ``pad_mask`` bool [B, L]
    (True = real).  Left padding is assumed by the data layer but not required here.
    """
    item_emb, pos = P["item_emb"], P["pos_emb"]
    dt = item_emb.dtype
    B, L = ids.shape
    d = item_emb.shape[1]
    pad_id = item_emb.shape[0] - 1
    ids = ids.masked_fill(~pad_mask, pad_id)  # legacy model.py:236-239 ; new path: data already holds padding_value
    x = item_emb[ids] * (d ** 0.5)
    if variant == "new":
        x = x + pos[pos.shape[0] - L :].unsqueeze(0)  # agg.py:51 (last L rows)
        if lnf_eps is None:
            lnf_eps = 1e-5  # from_params: torch.nn.LayerNorm default (model.py:248)
    else:
        assert L == pos.shape[0], "legacy SASRec needs L == max_len (sasrec/model.py:528-529)"
        x = x + pos.unsqueeze(0)
        x = x * pad_mask.unsqueeze(-1).to(dt)  # model.py:357
        if lnf_eps is None:
            lnf_eps = 1e-8  # model.py:463
    causal = torch.tril(torch.ones(L, L, dtype=torch.bool))
    if variant == "new":
        visible = causal.unsqueeze(0) & pad_mask.unsqueeze(1)  # mask.py:31-41 + transformer.py:94-96 (diag of pads re-masked)
        if mode == "eval":
            # eval uses finfo(fp32).min instead of -inf (mask.py:41, transformer.py:95): for a PAD query row i the sums are
            # -inf for pad keys j != i and finfo.min (+score, absorbed) for j == i and for real keys j > i  ->  uniform
            # attention over {i} U {j > i real}.  Real query rows are unchanged.  Only matters for pad rows, i.e. for
            # predict() only when a user has no real item at all.
            eye = torch.eye(L, dtype=torch.bool)
            pad_q = ~pad_mask
            vis_pad = eye.unsqueeze(0) | ((~causal).unsqueeze(0) & pad_mask.unsqueeze(1))
            visible = torch.where(pad_q.unsqueeze(-1), vis_pad, visible)
            uniform_rows = pad_q
    else:
        visible = causal.unsqueeze(0).expand(B, L, L)
    if not (variant == "new" and mode == "eval"):
        uniform_rows = None
    acts = []
    for blk in P["blocks"]:
        q = layer_norm(x, blk["ln1_w"], blk["ln1_b"], 1e-8)
        a = mha(q, x, blk, n_heads, visible, uniform_rows)  # K,V from the un-normalised x (transformer.py:99-106)
        x = q + a
        x = layer_norm(x, blk["ln2_w"], blk["ln2_b"], 1e-8)
        hmid = torch.relu(x @ blk["w1"].T + blk["b1"])
        x = x + (hmid @ blk["w2"].T + blk["b2"])  # ffn.py:49-55 / legacy model.py:502-504
        if variant == "legacy":
            x = x * pad_mask.unsqueeze(-1).to(dt)  # model.py:441
        acts.append(x)
    h = layer_norm(x, P["lnf_w"], P["lnf_b"], lnf_eps)
    return (h, acts) if return_all else h


# ----------------------------------------------------------------------------------------------------------------------
# loss / training step
# ----------------------------------------------------------------------------------------------------------------------


def ce_loss(hidden, table, labels, valid):
    """Full-catalog CE, mean over valid targets (nn/loss/ce.py:49-81 ; legacy sasrec/lightning.py:335-355).

    hidden [..., d], table [I, d] (item rows only), labels int64 [...], valid bool [...].
    """
    h = hidden.reshape(-1, hidden.shape[-1])[valid.reshape(-1)]
    y = labels.reshape(-1)[valid.reshape(-1)]
    logits = h @ table.T
    lse = torch.logsumexp(logits, dim=-1)
    tgt = logits.gather(1, y[:, None])[:, 0]
    return (lse - tgt).mean()


def train_loss(P, ids, pad_mask, labels, target_mask, n_heads, variant="new", lnf_eps=None):
    h = sasrec_body(P, ids, pad_mask, n_heads, variant, lnf_eps)
    n_items = P["item_emb"].shape[0] - 1
    return ce_loss(h, P["item_emb"][:n_items], labels, target_mask)


def flat_param_list(P):
    out = [P["item_emb"], P["pos_emb"]]
    for b in P["blocks"]:
        out += [b[k] for k in ("ln1_w", "ln1_b", "in_w", "in_b", "out_w", "out_b", "ln2_w", "ln2_b", "w1", "b1", "w2", "b2")]
    out += [P["lnf_w"], P["lnf_b"]]
    return out


def loss_and_grads(P, ids, pad_mask, labels, target_mask, n_heads, variant="new", lnf_eps=None):
    """Loss and d(loss)/d(param) via autograd on the restatement; the pad row's gradient is zeroed
    (torch.nn.Embedding(padding_idx=...) freezes it: nn/embedding.py:170-175, legacy model.py:339)."""
    Pg = {}
    for k, v in P.items():
        if k == "blocks":
            Pg[k] = [{kk: vv.detach().clone().requires_grad_(True) for kk, vv in b.items()} for b in v]
        else:
            Pg[k] = v.detach().clone().requires_grad_(True)
    loss = train_loss(Pg, ids, pad_mask, labels, target_mask, n_heads, variant, lnf_eps)
    loss.backward()
    G = {}
    for k, v in Pg.items():
        if k == "blocks":
            G[k] = [{kk: (vv.grad if vv.grad is not None else torch.zeros_like(vv)) for kk, vv in b.items()} for b in v]
        else:
            G[k] = v.grad if v.grad is not None else torch.zeros_like(v)
    G["item_emb"][-1].zero_()
    return loss.detach(), G


def adam_step(p, g, m, v, step, lr=1e-3, b1=0.9, b2=0.98, eps=1e-8):
    """torch.optim.Adam (no weight decay, no amsgrad) single-tensor update; ``step`` is 1-based."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


# ----------------------------------------------------------------------------------------------------------------------
# predict head
# ----------------------------------------------------------------------------------------------------------------------


def seen_filter(scores, seen_ids, item_count):
    """SeenItemsFilter._compute_scores (nn/lightning/postprocessor/seen_items.py:56-83): ids outside [0,item_count)
    are padding; scores[b, seen] = -inf on a clone."""
    scores = scores.clone()
    ok = (seen_ids >= 0) & (seen_ids < item_count)
    rows = torch.arange(scores.shape[0])[:, None].expand_as(seen_ids)
    scores[rows[ok], seen_ids[ok]] = float("-inf")
    return scores


def score_topk(hq, table, seen_ids, k, candidates=None, acc_dtype=torch.float64):
    """Predict head: scores = hq @ table.T, seen filter, top-k sorted descending.

    predictions_callback.py:80-96 (torch.topk(k, dim=1)); torch's tie order is unspecified, the oracle (and the CUDA
    kernel) use (score desc, scored column asc).  With ``candidates`` the scores are over table[candidates] in the given
    order and the returned ids are mapped back through ``candidates`` (predictions_callback.py:91-92); the seen filter
    applies to the real item ids (seen_items.py:68-71,80-81).
    Returns (ids int64 [B,k], scores acc_dtype [B,k]).
    """
    item_count = table.shape[0]
    hq = hq.to(acc_dtype)
    tb = table.to(acc_dtype)
    if candidates is not None:
        tb = tb[candidates]
    scores = hq @ tb.T
    if seen_ids is not None:
        if candidates is None:
            scores = seen_filter(scores, seen_ids, item_count)
        else:
            full = torch.full((scores.shape[0], item_count), float("-inf"), dtype=acc_dtype)
            full[:, candidates] = scores
            full = seen_filter(full, seen_ids, item_count)
            scores = full[:, candidates]
    # ties: smaller column (position in the scored list) first - torch.argsort(stable) on -score
    cols = torch.argsort(-scores, dim=1, stable=True)[:, :k]
    top = torch.gather(scores, 1, cols)
    ids = candidates[cols] if candidates is not None else cols
    return ids, top


# ----------------------------------------------------------------------------------------------------------------------
# input-layout producers (host side, known-answer tests in the reference)
# ----------------------------------------------------------------------------------------------------------------------


def sasrec_training_example(sequence, max_len, pad_value):
    """SasRecTrainingDataset.__getitem__ (models/nn/sequential/sasrec/dataset.py:104-126) on one item-id sequence:
    left-pad the last max_len+1 ids (torch_sequential_dataset.py:115-136), inputs = [:-1], labels = [1:].
    Returns (ids, pad_mask, labels, target_mask) for one row."""
    seq = torch.as_tensor(sequence, dtype=torch.int64)[-(max_len + 1) :]
    n = seq.numel()
    full = torch.full((max_len + 1,), pad_value, dtype=torch.int64)
    msk = torch.zeros(max_len + 1, dtype=torch.bool)
    if n:
        full[max_len + 1 - n :] = seq
        msk[max_len + 1 - n :] = True
    return full[:-1], msk[:-1], full[1:], msk[1:]
