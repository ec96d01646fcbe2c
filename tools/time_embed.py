"""Time rp_embed_fwd at the predict body's shape (4096 x 200 tokens, d = 128, |I| = 500 k) and the training shape."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from replay_b200.engine import EncoderConfig, SasRecEngine
from replay_b200.synthetic import make_sequences
for Bu, I in ((4096, 500_000), (512, 50_000)):
    L, d = 200, 128
    es = SasRecEngine(EncoderConfig(n_items=I, d=d, n_heads=2, n_blocks=2, max_len=L, variant="new"), Bu, L, "cuda", seed=7, with_grad=False)
    uid, upm, _, _ = make_sequences(Bu, I, L, seed=7)
    es.set_batch(uid.cuda(), upm.cuda()); es._prepare(False)
    cfg = es.cfg
    def run():
        from replay_b200._lib import check
        check(es.lib.rp_embed_fwd(es.params16["item_emb"].data_ptr(), es.params["pos_emb"].data_ptr(), es.ids32.data_ptr(), es.in_pad.data_ptr(),
                                  es.T, L, d, cfg.max_len - L, math.sqrt(cfg.d), 0, 0.0, es.seed, 0, es.rng_counter.data_ptr(), es.x[0].data_ptr(), es._stream()), "embed")
    for _ in range(3): run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): run()
    b.record(); torch.cuda.synchronize()
    print(f"embed_fwd {Bu} x {L}: {a.elapsed_time(b) / 20 * 1e3:.1f} us")
    del es
