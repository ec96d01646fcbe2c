"""Time the CE-head fwd/bwd for each compiled variant in replay_b200/build/variants (A/B of compile-time knobs)."""
import ctypes, glob, os, sys, json
import torch
P, ci, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
T, nv_, I, d = (102400, 55574, 50000, 128) if "--b256" not in sys.argv else (51200, 26263, 50000, 128)
if "--c3" in sys.argv:   # BERT4Rec config 3: 256 sequences, ~4000 masked positions, 100 k items, d = 256, biased head
    T, nv_, I, d = 51200, 3997, 100000, 256
BIAS = "--c3" in sys.argv
g = torch.Generator(device="cuda").manual_seed(0)
hc = torch.randn(T, d, device="cuda", generator=g).bfloat16(); hc[nv_:] = 0
table = (torch.randn(I, d, device="cuda", generator=g) * 0.3).bfloat16()
labels = torch.randint(0, I, (T,), device="cuda", generator=g).int()
nv = torch.tensor([nv_], dtype=torch.int32, device="cuda")
loss = torch.zeros(2, device="cuda"); lse = torch.zeros(T, device="cuda")
cvec = torch.full((T,), float("-inf"), device="cuda")
d_hc = torch.zeros(T, d, device="cuda", dtype=torch.bfloat16); d_tab = torch.zeros(I + 1, d, device="cuda")
bias = (torch.randn(I + 127, device="cuda", generator=g) * 0.1) if BIAS else None
d_bias = torch.zeros(I + 127, device="cuda") if BIAS else None
bp, dbp = (bias.data_ptr(), d_bias.data_ptr()) if BIAS else (None, None)
st = torch.cuda.current_stream().cuda_stream
res = {}
for path in sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "replay_b200", "build", "variants", "*.so"))):
    L = ctypes.CDLL(path)
    L.rp_ce_head_workspace.restype = sz; L.rp_ce_head_workspace.argtypes = [ci, ci, ci]
    L.rp_ce_head_fwd.argtypes = [P, P, P, P, P, ci, ci, ci, P, P, P, P, ci, P, sz, P]
    L.rp_ce_head_bwd.argtypes = [P, P, P, P, P, ci, ci, ci, P, P, P, P, P, ci, ci, P, sz, P]
    wsb = L.rp_ce_head_workspace(T, I, d); ws = torch.zeros(wsb, dtype=torch.uint8, device="cuda")
    fwd = lambda: L.rp_ce_head_fwd(hc.data_ptr(), table.data_ptr(), bp, labels.data_ptr(), nv.data_ptr(), T, I, d, loss.data_ptr(), lse.data_ptr(), cvec.data_ptr(), d_hc.data_ptr(), nv_, ws.data_ptr(), wsb, st)
    bwd = lambda: L.rp_ce_head_bwd(hc.data_ptr(), table.data_ptr(), bp, labels.data_ptr(), nv.data_ptr(), T, I, d, loss.data_ptr(), cvec.data_ptr(), d_hc.data_ptr(), d_tab.data_ptr(), dbp, 1, 0, ws.data_ptr(), wsb, st)
    def t(fn, n=10):
        for _ in range(3): assert fn() == 0
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    r = dict(fwd_ms=round(t(fwd), 4), bwd_ms=round(t(bwd), 4), loss=float(loss[0]))
    # checksums of both gradients: every variant must agree with the round-1 issue order to accumulation noise
    r["dh_abs_sum"] = float(d_hc[:nv_].float().abs().sum()); r["de_abs_sum"] = float(d_tab[:I].abs().sum())
    r["dh_probe"] = float(d_hc[nv_ // 2].float().sum()); r["de_probe"] = float(d_tab[I // 3].sum())
    res[os.path.basename(path)] = r
    print(os.path.basename(path), json.dumps(r), file=sys.stderr, flush=True)   # progress: a variant that hangs is the next one
print(json.dumps(res, indent=1))
