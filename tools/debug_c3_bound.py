"""Diagnostic: the device-side |logit| bound of the fused CE pass at BERT4Rec config 3 (does the fused path run in the step?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from replay_b200 import ops
from replay_b200.trainer import Trainer

c = dict(bench.CONFIGS[3]) if hasattr(bench, "CONFIGS") else None
print("config", c)
dev = torch.device("cuda")
mod, core, to_batch = bench.build_module(c, dev)
B, L = c["per_gpu_batch"], c["seq_len"]
eng = core.ensure_engine(B, L, with_grad=True)
tr = Trainer(eng, use_graph=False)
data = bench.make_batches(c, B * 2, seed=1234)
devb = [t.reshape(2, B, L).to(dev) for t in data]
for i in range(4):
    tr.step(*(t[i % 2] for t in devb))
    torch.cuda.synchronize()
    st = eng.ce
    off = st.capacity * 32 * 2 * 8 + 4096
    words = st.ws[off:off + 20].view(torch.int32)
    fl = st.ws[off + 4:off + 16].view(torch.float32)
    print("step", i, "ticket", int(words[0]), "max|h|^2, max|e|^2, max|b| =", [float(x) for x in fl], "flag", int(words[4]),
          "n_valid", int(eng.n_valid.item()), "loss", float(eng.ce.loss[0]))
W16, bias = eng._head()
print("head rows", W16.shape, "bias", None if bias is None else (bias.shape, float(bias.abs().max())), "table row norm max",
      float(W16.float().norm(dim=1).max()), "hc row norm max", float(eng.hc[: int(eng.n_valid.item())].float().norm(dim=1).max()))
