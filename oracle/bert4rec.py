"""Plain-torch CPU restatement of RePlay's BERT4Rec hot path (TEST INFRASTRUCTURE - see oracle/__init__.py).

Reference files restated (under /root/reference/replay/models/nn/sequential/bert4rec):
  model.py:121-143 (forward_step), 239-296 (BertEmbedding), 363-382 (BaseHead = F.linear), 481-500 (TransformerBlock),
  521-527 (PositionwiseFeedForward, exact-erf GELU) ; lightning.py:332-351 (CE over pad & ~tok positions) ;
  dataset.py:71-92 (uniform masker), 322-345 (_shift_features for predict).

Canonical parameter dict::
    item_emb [I, d] (no pad row: padding_value 0 is a valid row), mask_emb [1, d], pos_emb [L, d]
    blocks: ln1_w ln1_b in_w in_b out_w out_b ln2_w ln2_b w1[4d,d] b1[4d] w2[d,4d] b2[d]
    head_w [I, d] + head_b [I]   (untied ``ClassificationHead``)   or   head_b only (tied: head_w is item_emb)
"""
from __future__ import annotations

import math

import torch

from .sasrec import layer_norm, mha


def params_from_state_dict(sd, item_feature="item_id"):
    sd = {k: v.detach().clone() for k, v in sd.items()}
    n = 0
    while f"transformer_blocks.{n}.attention.in_proj_weight" in sd:
        n += 1
    P = {
        "item_emb": sd[f"item_embedder.cat_embeddings.{item_feature}.weight"],
        "mask_emb": sd["item_embedder.mask_embedding.weight"],
        "pos_emb": sd["item_embedder.position.pe.weight"],
        "blocks": [],
    }
    for i in range(n):
        t = f"transformer_blocks.{i}."
        P["blocks"].append(
            {
                "ln1_w": sd[t + "attention_norm.weight"], "ln1_b": sd[t + "attention_norm.bias"],
                "in_w": sd[t + "attention.in_proj_weight"], "in_b": sd[t + "attention.in_proj_bias"],
                "out_w": sd[t + "attention.out_proj.weight"], "out_b": sd[t + "attention.out_proj.bias"],
                "ln2_w": sd[t + "pff_norm.weight"], "ln2_b": sd[t + "pff_norm.bias"],
                "w1": sd[t + "pff.w_1.weight"], "b1": sd[t + "pff.w_1.bias"],
                "w2": sd[t + "pff.w_2.weight"], "b2": sd[t + "pff.w_2.bias"],
            }
        )
    if "_head.linear.weight" in sd:
        P["head_w"] = sd["_head.linear.weight"]
        P["head_b"] = sd["_head.linear.bias"]
    else:
        P["head_b"] = sd["_head.out_bias"]
    return P


def gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def bert4rec_body(P, ids, pad_mask, token_mask, n_heads, num_passes=1):
    """Hidden states [B, L, d], dropout off (model.py:121-143)."""
    B, L = ids.shape
    x = torch.where(token_mask.unsqueeze(-1), P["item_emb"][ids], P["mask_emb"].expand(B, L, -1))  # model.py:285-288
    x = x + P["pos_emb"][:L].unsqueeze(0)  # no sqrt(d) scaling
    visible = pad_mask.unsqueeze(1).expand(B, L, L)  # key_padding_mask=~pad only (model.py:494)
    for blk in P["blocks"]:
        for _ in range(num_passes):
            xn = layer_norm(x, blk["ln1_w"], blk["ln1_b"], 1e-5)
            a = mha(xn, xn, blk, n_heads, visible)
            y = x + a
            yn = layer_norm(y, blk["ln2_w"], blk["ln2_b"], 1e-5)
            x = y + (gelu_erf(yn @ blk["w1"].T + blk["b1"]) @ blk["w2"].T + blk["b2"])
    return x


def head_weights(P):
    return (P["head_w"] if "head_w" in P else P["item_emb"]), P["head_b"]


def train_loss(P, ids, pad_mask, token_mask, labels, n_heads):
    """CE over positions that are real and masked (lightning.py:344-351)."""
    h = bert4rec_body(P, ids, pad_mask, token_mask, n_heads)
    w, b = head_weights(P)
    sel = pad_mask & ~token_mask
    logits = h[sel] @ w.T + b
    y = labels[sel]
    return (torch.logsumexp(logits, -1) - logits.gather(1, y[:, None])[:, 0]).mean()


def shift_for_predict(ids, pad_mask, token_mask, pad_value=0):
    """_shift_features (dataset.py:322-345): roll left by one; the last position becomes <MASK> with pad=True."""
    ids2 = torch.roll(ids, -1, dims=-1)
    ids2[..., -1] = pad_value
    pm = torch.roll(pad_mask, -1, dims=-1)
    pm[..., -1] = True
    tm = torch.roll(token_mask, -1, dims=-1)
    tm[..., -1] = False
    return ids2, pm, tm


def uniform_masker(pad_mask, mask_prob, generator):
    """Bert4RecUniformMasker.mask (dataset.py:71-92): token_mask = rand > p, & pad; if nothing is masked among the real
    tokens mask the last one; if everything is masked unmask the first real one... (corner cases per reference)."""
    tok = torch.rand(pad_mask.shape, generator=generator) > mask_prob
    tok = tok & pad_mask
    return tok
