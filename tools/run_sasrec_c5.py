"""SASRec at BASELINE configs[4] shape on ONE GPU (L=512, d=512, H=8, |I|=1M, full CE + Adam): step time per batch size,
the share of the CE head, and predict (top-10 over 1M items).  The config is quoted for 8 GPUs data-parallel; per-GPU work is
the same (weak scaling), the all-reduce of the 2 GB gradient is not included here."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from replay_b200 import ops
from replay_b200.engine import EncoderConfig, SasRecEngine
from replay_b200.synthetic import make_sequences

L, d, H, I = 512, 512, 8, 1_000_000
batches = [int(x) for x in (sys.argv[1:] or ["16", "32"])]


def ev_time(fn, n):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for B in batches:
    cfg = EncoderConfig(n_items=I, d=d, n_heads=H, n_blocks=2, max_len=L, dropout=0.2, variant="new")
    eng = SasRecEngine(cfg, B, L, "cuda", seed=1)
    ids, pm, lab, tm = make_sequences(B, I, L, seed=1234)
    eng.set_batch(ids.cuda(), pm.cuda(), lab.cuda(), tm.cuda())
    losses = [float(eng.train_step()[0])]
    eng.n_valid_hint = int(eng.n_valid)  # the loader knows how many targets a batch holds (load balance only)
    losses += [float(eng.train_step()[0]) for _ in range(2)]
    ms = ev_time(eng.train_step, 5)
    nv = int(eng.n_valid)
    table16 = eng.params16["item_emb"][:I]
    t_f = ev_time(lambda: ops.ce_head_fwd(eng.ce, eng.hc, table16, eng.labels_c, eng.n_valid, d_hc=eng.s["dhc"], n_valid_hint=eng.n_valid_hint), 3)
    t_b = ev_time(lambda: ops.ce_head_bwd(eng.ce, eng.hc, table16, eng.labels_c, eng.n_valid, eng.s["dhc"], eng.grads["item_emb"], n_valid_hint=eng.n_valid_hint), 3)
    t_adam = ev_time(lambda: eng.optimizer_step(), 3)
    flops = 3 * 2.0 * nv * I * d
    print(f"c5 B={B}: n_valid {nv} losses {[round(x, 3) for x in losses]} step {ms:.1f} ms -> {B / ms * 1e3:.1f} seq/s | "
          f"CE fwd {t_f:.1f} ms bwd {t_b:.1f} ms ({flops / ((t_f + t_b) * 1e-3) / 1e12:.0f} TFLOP/s credited) adam {t_adam:.2f} ms | "
          f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del eng
    torch.cuda.empty_cache()

# predict: 512 users per call (small-batch, HBM-bound head: 1 GB table per call) and 4096 users per call
for Bu in (512, 4096):
    cfg = EncoderConfig(n_items=I, d=d, n_heads=H, n_blocks=2, max_len=L, variant="new")
    es = SasRecEngine(cfg, Bu, L, "cuda", seed=7, with_grad=False)
    uid, upm, _, _ = make_sequences(Bu, I, L, seed=7)
    uid, upm = uid.cuda(), upm.cuda()
    tab = es.params16["item_emb"][:I]

    def predict():
        es.set_batch(uid, upm)
        hq = es.forward_last_hidden()
        return ops.score_topk(hq, tab, 10, ops.seen_prepare(uid, I))

    predict()
    ms = ev_time(predict, 3)
    hq = es.hq
    seen = ops.seen_prepare(uid, I)
    th = ev_time(lambda: ops.score_topk(hq, tab, 10, seen), 3)
    print(f"c5 predict {Bu} users/call: {ms:.2f} ms -> {Bu / ms * 1e3:.0f} users/s ; head {th:.2f} ms "
          f"({2.0 * Bu * I * d / (th * 1e-3) / 1e12:.0f} TFLOP/s, table {I * d * 2 / (th * 1e-3) / 1e9:.0f} GB/s)", flush=True)
    del es
    torch.cuda.empty_cache()
