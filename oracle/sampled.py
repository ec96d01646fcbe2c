"""TEST INFRASTRUCTURE - CPU restatement of the reference's sampled losses, plain torch autograd.

Follows
  * SampledLossBase.get_sampled_logits (replay/nn/loss/base.py:49-154): boolean-mask the valid targets, logits of the
    positive and of the negatives through the tying head (replay/nn/head.py:16-50),
  * mask_negative_logits (replay/nn/loss/base.py:157-196): -1e9 where a negative equals the positive / the ignore index,
  * CESampled.forward (replay/nn/loss/ce.py:199-249), BCESampled.forward (replay/nn/loss/bce.py:154-218),
  * legacy _compute_loss_ce_sampled / _compute_loss_bce_sampled (replay/models/nn/sequential/sasrec/lightning.py:310-376)
    with the negatives made an explicit argument (the reference draws them inside the loss with torch's RNG).
Single positive per position.  Pinned against the real reference classes by oracle/gen_golden.py::gen_sampled_losses ->
tests/golden/sampled_losses.npz.
"""
from __future__ import annotations

import math

import torch


def sampled_logits(hidden, table, positive_labels, negative_labels, target_mask):
    """hidden [B, L, d], table [|I|(+1), d], positive_labels [B, L], target_mask [B, L] bool, negative_labels [N] | [B, N] |
    [B, L, N]  ->  (z_pos [M, 1], z_neg [M, N], pos [M, 1], neg [N] or [M, N]) over the M valid targets."""
    B, L, d = hidden.shape
    neg = negative_labels
    if neg.dim() == 2:
        neg = neg.unsqueeze(1).repeat(1, L, 1)
    h = hidden[target_mask]                       # [M, d]
    pos = positive_labels[target_mask].unsqueeze(-1)
    z_pos = (h * table[pos[:, 0]]).sum(-1, keepdim=True)
    if neg.dim() == 1:
        z_neg = h @ table[neg].T
    else:
        neg = neg[target_mask]                    # [M, N]
        z_neg = torch.einsum("md,mnd->mn", h, table[neg])
    return z_pos, z_neg, pos, neg


def mask_negative_logits(z_neg, neg, pos, ignore_index):
    z_neg = z_neg.clone()
    if ignore_index >= 0:
        z_neg = z_neg.masked_fill(neg == ignore_index if neg.dim() > 1 else (neg == ignore_index).unsqueeze(0).expand_as(z_neg), -1e9)
    n = neg.unsqueeze(-2) if neg.dim() > 1 else neg
    collide = (pos.unsqueeze(-1) == n).sum(-2).bool()
    return z_neg.masked_fill(collide, -1e9)


def ce_sampled(hidden, table, positive_labels, negative_labels, target_mask, ignore_index=-100):
    z_pos, z_neg, pos, neg = sampled_logits(hidden, table, positive_labels, negative_labels, target_mask)
    z_neg = mask_negative_logits(z_neg, neg, pos, ignore_index)
    logits = torch.cat((z_pos, z_neg), dim=-1)
    return torch.nn.functional.cross_entropy(logits, torch.zeros(len(logits), dtype=torch.long))


def bce_sampled(hidden, table, positive_labels, negative_labels, target_mask, log_eps=1e-6, clamp=100.0, ignore_index=-100,
                mask_collisions=True):
    z_pos, z_neg, pos, neg = sampled_logits(hidden, table, positive_labels, negative_labels, target_mask)
    if mask_collisions:
        z_neg = mask_negative_logits(z_neg, neg, pos, ignore_index)
    pp, npb = torch.sigmoid(z_pos), torch.sigmoid(z_neg)
    pl = torch.clamp(torch.log(pp + log_eps), -clamp, clamp).sum()
    nl = torch.clamp(torch.log((1 - npb) + log_eps), -clamp, clamp).sum()
    return -(pl + nl) / z_pos.size(0)


def legacy_ce_sampled(hidden, table, positive_labels, negative_labels, target_mask, vocab_size):
    """negative_labels [B, L, N] (one independent draw per valid target, as torch.randint in the reference)."""
    z_pos, z_neg, pos, neg = sampled_logits(hidden, table, positive_labels, negative_labels, target_mask)
    n_neg = min(z_neg.size(1), vocab_size)
    reject = pos == neg
    z_neg = z_neg + math.log(vocab_size - 1)
    z_neg = z_neg - 1e6 * reject
    z_neg = z_neg - torch.log((n_neg - reject.sum(dim=-1, keepdim=True)).float())
    logits = torch.cat([z_pos, z_neg], dim=1).float()
    return torch.nn.functional.cross_entropy(logits, torch.zeros(len(logits), dtype=torch.long))


def legacy_bce_sampled(hidden, table, positive_labels, negative_labels, target_mask):
    return bce_sampled(hidden, table, positive_labels, negative_labels, target_mask, 1e-6, 100.0, mask_collisions=False)


def sasrec_sampled_loss(P, ids, pad_mask, labels, target_mask, negatives, n_heads, kind, variant="new", **kw):
    """Body of oracle.sasrec + one of the sampled heads; returns the scalar loss (autograd-capable)."""
    from . import sasrec as osr
    h = osr.sasrec_body(P, ids, pad_mask, n_heads, variant)
    table = P["item_emb"]
    fn = {"ce": ce_sampled, "bce": bce_sampled, "legacy_ce": legacy_ce_sampled, "legacy_bce": legacy_bce_sampled}[kind]
    return fn(h, table, labels, negatives, target_mask, **kw)


def loss_and_grads(P, ids, pad_mask, labels, target_mask, negatives, n_heads, kind, variant="new", **kw):
    Pg = {}
    for k, v in P.items():
        Pg[k] = [{kk: vv.detach().clone().requires_grad_(True) for kk, vv in b.items()} for b in v] if k == "blocks" \
            else v.detach().clone().requires_grad_(True)
    loss = sasrec_sampled_loss(Pg, ids, pad_mask, labels, target_mask, negatives, n_heads, kind, variant, **kw)
    loss.backward()
    G = {}
    for k, v in Pg.items():
        if k == "blocks":
            G[k] = [{kk: (vv.grad if vv.grad is not None else torch.zeros_like(vv)) for kk, vv in b.items()} for b in v]
        else:
            G[k] = v.grad if v.grad is not None else torch.zeros_like(v)
    G["item_emb"][-1].zero_()
    return loss.detach(), G


# ----------------------------------------------------------------------------------------------------------------------
# full-catalog per-row losses, one positive label per position
#   LogOutCE          replay/nn/loss/logout_ce.py:74-145   CrossEntropyLoss over [positive | catalog with the positive masked]
#   LogOutCEWeighted  replay/nn/loss/logout_ce.py:198-228  mean(loss_t * w_t), w masked by the target padding mask
#   CEWeighted        replay/nn/loss/ce.py:111-143         (loss [B * L] * w [B, L, 1]).mean() - the broadcast product, i.e.
#                                                          sum of the valid rows' CE / (B * L) * mean(w over ALL positions);
#                                                          restated as the reference computes it
#   LogInCE           replay/nn/loss/login_ce.py:170-239   -clamp(log(p + eps), -c, c), p = softmax prob of the positive
# Pinned against the real classes by oracle/gen_golden.py::gen_row_losses -> tests/golden/row_losses.npz.
# ----------------------------------------------------------------------------------------------------------------------
def row_loss(hidden, table, labels, target_mask, kind, weights=None, log_eps=1e-6, clamp=100.0):
    """hidden [B, L, d], table [|I|(+1), d] (rows >= |I| ignored via n_items = labels' range), labels [B, L], target_mask
    [B, L] bool, weights [B, L, 1] or None."""
    h = hidden[target_mask]
    y = labels[target_mask]
    logits = h @ table.T
    ce = torch.nn.functional.cross_entropy(logits, y, reduction="none")
    if kind == "logout":
        return ce.mean()
    if kind == "logout_weighted":
        return (ce * weights[..., 0][target_mask]).mean()
    if kind == "ce_weighted":
        # CE.forward keeps all B * L positions (ignore_index, reduction="none": zeros at the ignored ones); the product with
        # the [B, L, 1] weights broadcasts to [B, L, B * L] and its mean is mean(loss over B * L) * mean(w) - as the reference
        full = torch.zeros(target_mask.numel(), dtype=ce.dtype)
        full = full.masked_scatter(target_mask.reshape(-1), ce)
        return (full * weights).mean()
    if kind == "login":
        p = torch.exp(-ce)
        return (-torch.clamp(torch.log(p + log_eps), -clamp, clamp)).mean()
    raise ValueError(kind)


def row_loss_and_grads(P, ids, pad_mask, labels, target_mask, n_heads, kind, weights=None, **kw):
    from .sasrec import sasrec_body

    Pg = {}
    for k, v in P.items():
        Pg[k] = [{kk: vv.detach().clone().requires_grad_(True) for kk, vv in b.items()} for b in v] if k == "blocks" \
            else v.detach().clone().requires_grad_(True)
    hidden = sasrec_body(Pg, ids, pad_mask, n_heads, variant="new")
    n_items = Pg["item_emb"].shape[0] - 1
    loss = row_loss(hidden, Pg["item_emb"][:n_items], labels, target_mask, kind, weights, **kw)
    loss.backward()
    G = {}
    for k, v in Pg.items():
        if k == "blocks":
            G[k] = [{kk: (vv.grad if vv.grad is not None else torch.zeros_like(vv)) for kk, vv in b.items()} for b in v]
        else:
            G[k] = v.grad if v.grad is not None else torch.zeros_like(v)
    G["item_emb"][-1].zero_()
    return loss.detach(), G
