#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native RePlay sequential-recommender hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (torchrun launches N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on host cores

Workload (BASELINE.json configs[1]): SASRec seq_len=200 d=128 H=2 2 blocks |items|=50 000, full-catalog CE, Adam,
dropout 0.2, bf16 compute / fp32 master, MovieLens-shaped synthetic sequences (replay_b200/synthetic.py, seed 1234),
data parallel over N GPUs (weak scaling: 512 sequences per GPU per step; 128 / 256 / 512 / 1024 give 92 / 109 / 124 / 130 k
seq/s on one B200, profiles/README.md).  One step = forward + backward + gradient
all-reduce + Adam over one batch.  Metric: training sequences/s (whole job).  The same JSON line also carries the scoring
leg of BASELINE's metric (users/s, top-K@10 with seen-item filter, |items| = 500 000) under "scoring".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(model="SASRec(new path)", seq_len=200, d=128, heads=2, blocks=2, n_items=50_000, dropout=0.2, per_gpu_batch=512)
SCORE_CFG = dict(n_items=500_000, d=128, seq_len=200, k=10, users_per_call=4096)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            z = json.load(fh)
        return dict(hbm=z["hbm_gbs"], tc_burst=z["bf16_tflops"], tc_sustained=z["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = int(self.rows[0][1]) if self.rows and self.rows[0][1].isdigit() else None
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def train_flops_per_seq(L, d, n_blocks, n_items, n_valid_per_seq):
    """SURVEY.md §8d dense upper bound: body N_b(12 d^2 + 4 d L) per token, head 2 d |I| per VALID target, x3 for train."""
    body = L * n_blocks * (12 * d * d + 4 * d * L)
    head = n_valid_per_seq * 2 * d * n_items
    return 3.0 * (body + head)


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port (plain torch fp32 on the host cores)
# ----------------------------------------------------------------------------------------------------------------------
def cpu_train_seq_per_s(batch=8, steps=3, warmup=1):
    from oracle import sasrec as osr
    from replay_b200.synthetic import make_sequences

    torch.set_num_threads(min(os.cpu_count() or 1, 32))  # torch CPU GEMMs stop scaling (and regress) past ~32 threads
    c = CFG
    P = osr.random_params(c["n_items"], c["d"], c["seq_len"], c["blocks"], seed=0)
    flat = [p.requires_grad_(True) for p in osr.flat_param_list(P)]
    opt = torch.optim.Adam(flat, lr=1e-3, betas=(0.9, 0.98))
    ids, pm, lab, tm = make_sequences(batch * (steps + warmup), c["n_items"], c["seq_len"], seed=1234)
    ts = []
    for s in range(steps + warmup):
        sl = slice(s * batch, (s + 1) * batch)
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = osr.train_loss(P, ids[sl], pm[sl], lab[sl], tm[sl], c["heads"], "new")
        loss.backward()
        P["item_emb"].grad[-1].zero_()
        opt.step()
        ts.append(time.perf_counter() - t0)
    ts = sorted(ts[warmup:])
    med = ts[len(ts) // 2]
    return batch / med, med, float(loss.detach())


def cpu_predict_users_per_s(users=64, reps=3):
    """Reference predict path on the host cores (oracle port): body forward (eval) -> last hidden -> [U, |I|] logits ->
    SeenItemsFilter (clone + scatter -inf) -> torch.topk(10), fp32, at the scoring leg's shape (|I| = 500K, L = 200, d = 128)."""
    from oracle import sasrec as osr
    from replay_b200.synthetic import make_sequences

    torch.set_num_threads(min(os.cpu_count() or 1, 32))
    sc = SCORE_CFG
    P = osr.random_params(sc["n_items"], sc["d"], sc["seq_len"], 2, seed=7)
    ids, pm, _, _ = make_sequences(users, sc["n_items"], sc["seq_len"], seed=7)
    ts = []
    with torch.no_grad():
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            h = osr.sasrec_body(P, ids, pm, 2, "new", mode="eval")[:, -1]
            scores = h @ P["item_emb"][: sc["n_items"]].T
            scores = osr.seen_filter(scores, ids, sc["n_items"])
            torch.topk(scores, sc["k"], dim=1)
            ts.append(time.perf_counter() - t0)
    ts = sorted(ts[1:])
    return users / ts[len(ts) // 2]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 8
    n_timed = max(1, min(args.steps, 5))  # bounded sample: a CPU step of this workload takes seconds
    v, med, _ = cpu_train_seq_per_s(batch=batch, steps=n_timed, warmup=1)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "sasrec_train_seq_per_s", "value": v, "unit": "seq/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SASRec L=200 d=128 |I|=50K full-CE train step, CPU oracle port of the reference algorithm",
                   **{k: CFG[k] for k in ("seq_len", "d", "heads", "blocks", "n_items")}, "global_batch": batch},
        "cpu_baseline": {"value": v, "unit": "seq/s", "cores": cores, "kind": "port",
                         "sample": f"{n_timed} timed steps of batch {batch} (fwd+bwd+Adam, dropout off), torch fp32, {cores} threads"},
        "e2e": {"value": v, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    from replay_b200 import ops
    from replay_b200.engine import EncoderConfig, SasRecEngine
    from replay_b200.synthetic import make_sequences
    from replay_b200.trainer import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: replay_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    c = dict(CFG)
    if args.dropout is not None:
        c["dropout"] = args.dropout
    if args.batch is not None:
        c["per_gpu_batch"] = args.batch
    B, L, d, I = c["per_gpu_batch"], c["seq_len"], c["d"], c["n_items"]
    cfg = EncoderConfig(n_items=I, d=d, n_heads=c["heads"], n_blocks=c["blocks"], max_len=L, dropout=c["dropout"], variant="new")
    eng = SasRecEngine(cfg, B, L, dev, seed=1234)
    tr = Trainer(eng, use_graph=not args.no_graph)
    n_batches = 6
    ids, pm, lab, tm = make_sequences(B * n_batches * world, I, L, seed=1234)
    sh = slice(rank * B * n_batches, (rank + 1) * B * n_batches)
    host = [t[sh].view(n_batches, B, L).pin_memory() for t in (ids, pm, lab, tm)]
    devb = [t.to(dev) for t in host]
    valid_per_seq = float(tm.sum()) / tm.shape[0]
    eng.n_valid_hint = int(valid_per_seq * B)  # the data loader knows how many targets a batch holds (load balance only)
    PK = peaks()

    def step_dev(i):
        j = i % n_batches
        return tr.step(devb[0][j], devb[1][j], devb[2][j], devb[3][j])

    def step_e2e(i):  # host buffers -> device every step, loss read back every step
        j = i % n_batches
        loss = tr.step(*(h[j].to(dev, non_blocking=True) for h in host))
        return float(loss[0].item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W, K = max(args.warmup, 3), args.steps
    for i in range(W + 3):  # +3: two eager steps and the graph capture
        step_dev(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        loss = step_dev(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    final_loss = float(loss[0].item())
    # ---- e2e: same step through host buffers
    for i in range(2):
        step_e2e(i)
    barrier()
    e0.record()
    for i in range(K):
        step_e2e(i)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    h2d = sum(h[0].numel() * h[0].element_size() for h in host)
    # ---- same step fed by device-side batch construction: histories resident in HBM as CSR, one rp_build_batch launch
    # per step cuts / left-pads / shifts the windows of B randomly drawn users, loss read back every step
    dev_batches = None
    if not args.no_device_batches:
        from replay_b200.device_data import DeviceSequenceStore
        from replay_b200.synthetic import make_histories

        n_hist = 65536
        off_h, items_h = make_histories(n_hist, I, seed=1234 + rank)
        store = DeviceSequenceStore(offsets=off_h.numpy(), items=items_h.numpy(), device=dev)
        picks = torch.randint(0, n_hist, (n_batches, B), generator=torch.Generator().manual_seed(rank), dtype=torch.int32).to(dev)

        def step_store(i):
            b = store.sasrec_training_batch(picks[i % n_batches], L, I)
            loss = tr.step(b["feature_tensor"]["item_id"], b["padding_mask"], b["positive_labels"], b["target_padding_mask"])
            return float(loss[0].item())

        for i in range(2):
            step_store(i)
        barrier()
        e0.record()
        for i in range(K):
            step_store(i)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_batches = {"value": world * B * K / float(t.item()) * 1e3, "unit": "seq/s", "ms_per_step": float(t.item()) / K,
                       "histories_per_gpu": n_hist, "store_bytes": int(items_h.numel() * 4 + off_h.numel() * 8),
                       "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4,
                       "note": "batches cut on the GPU from the HBM-resident CSR history store (rp_build_batch), no host input"}
        del store
    def time_kernel(fn, iters=10):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    # ---- scoring leg (every rank scores its own contiguous shard of the users, no collective: SURVEY 8e):
    # body forward (eval, last block evaluated for the last position only) + fused score/seen-mask/top-K at |I| = 500K
    sc = SCORE_CFG
    scoring = None
    if not args.no_scoring:
        from replay_b200.trainer import user_shard

        torch.cuda.empty_cache()
        Bu = sc["users_per_call"]
        cfg_s = EncoderConfig(n_items=sc["n_items"], d=sc["d"], n_heads=2, n_blocks=2, max_len=sc["seq_len"], variant="new")
        es = SasRecEngine(cfg_s, Bu, sc["seq_len"], dev, seed=7, with_grad=False)
        n_calls = 6
        lo, hi = user_shard(world * Bu * 2, rank, world)  # 2 distinct calls' worth of users per rank
        uid, upm, _, _ = make_sequences(world * Bu * 2, sc["n_items"], sc["seq_len"], seed=7)
        uid, upm = uid[lo:hi].view(2, Bu, -1).to(dev), upm[lo:hi].view(2, Bu, -1).to(dev)
        tab = es.params16["item_emb"][: sc["n_items"]]

        def predict(i):
            j = i % 2
            es.set_batch(uid[j], upm[j])
            hq = es.forward_last_hidden()
            seen = ops.seen_prepare(uid[j], sc["n_items"])
            return ops.score_topk(hq, tab, sc["k"], seen)

        for i in range(3):
            predict(i)
        barrier()
        e0.record()
        for i in range(n_calls):
            predict(i)
        e1.record()
        barrier()
        tp = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        ms_p = float(tp.item()) / n_calls
        if rank == 0:
            seen = ops.seen_prepare(uid[0], sc["n_items"])
            hq = es.hq
            t_head = time_kernel(lambda: ops.score_topk(hq, tab, sc["k"], seen))
            head_flops = 2.0 * Bu * sc["n_items"] * sc["d"]
            scoring = {
                "metric": "sasrec_predict_topk10_users_per_s", "value": world * Bu / ms_p * 1e3, "unit": "users/s", "n_gpus": world,
                "config": {"workload": "SASRec predict(): body fwd + fused score+seen-filter+top-10, users sharded over the GPUs",
                           **sc},
                "ms_per_call": ms_p,
                "cpu_baseline": None if args.no_cpu else {
                    "value": cpu_predict_users_per_s(), "unit": "users/s", "cores": torch.get_num_threads(), "kind": "port",
                    "sample": "64 users, 3 timed calls: oracle body + full logits + seen filter + torch.topk, torch fp32 CPU"},
                "roofline": {"bound": "tensor", "kernel": "score_topk_kernel", "achieved": head_flops / (t_head * 1e-3) / 1e12,
                             "peak": PK["tc_burst"], "unit": "TFLOP/s", "frac": head_flops / (t_head * 1e-3) / 1e12 / PK["tc_burst"],
                             "head_ms": t_head, "head_users_per_s": Bu / t_head * 1e3, "traffic": None},
            }
        del es
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    seq_s = world * B * K / ms * 1e3
    seq_s_e2e = world * B * K / ms_e2e * 1e3
    fl_seq = train_flops_per_seq(L, d, c["blocks"], I, valid_per_seq)
    step_tflops = seq_s / world * fl_seq / 1e12

    # ---- roofline of the dominant kernels: the three tcgen05 CE-head kernels, timed live with CUDA events (standalone,
    # same buffers as the last step; each launch streams > L2 worth of operands through TMEM/SMEM)
    n_valid = int(eng.n_valid.item())
    table16 = eng.params16["item_emb"][:I]
    t_fwd = time_kernel(lambda: ops.ce_head_fwd(eng.ce, eng.hc, table16, eng.labels_c, eng.n_valid, d_hc=eng.s["dhc"],
                                                n_valid_hint=eng.n_valid_hint))
    t_bwd = time_kernel(lambda: ops.ce_head_bwd(eng.ce, eng.hc, table16, eng.labels_c, eng.n_valid, eng.s["dhc"], eng.grads["item_emb"]))
    gemm_flops = 2.0 * n_valid * I * d
    ce_ms = t_fwd + t_bwd
    n_exec = 4 if eng.fused_ce else 5  # GEMM-equivalents executed: fused fwd+dH (S, dH) + dE pass (S, dE)
    roof = {
        "bound": "tensor", "kernel": "ce_bwd_kernel<FUSED> (fwd+dH) + ce_bwd_kernel<COL> (dE): logits GEMM + softmax-CE, fwd+bwd",
        "achieved": 3 * gemm_flops / (ce_ms * 1e-3) / 1e12, "peak": PK["tc_burst"], "unit": "TFLOP/s",
        "frac": 3 * gemm_flops / (ce_ms * 1e-3) / 1e12 / PK["tc_burst"],
        # dram__bytes_read+write of the dE pass from the committed ncu capture (profiles/r1_ce_head_ncu.md): equals the
        # algorithmic bytes (bf16 table 12.8 MB + compacted hidden rows 6.7 MB); its fp32 dE output stays in L2
        "traffic": 19.7e6,
        "peak_source": PK["src"] + " burst (kernels timed alone)",
        "detail": {"ce_fwd_ms": t_fwd, "ce_bwd_ms": t_bwd, "n_valid_targets": n_valid,
                   "algorithmic_flops_per_launch_pair": 3 * gemm_flops,
                   "executed_tflops": n_exec * gemm_flops / (ce_ms * 1e-3) / 1e12, "fused_fwd_dh": bool(eng.fused_ce),
                   "share_of_step": ce_ms / (ms / K)},
    }
    # ---- CPU baseline (bounded sample, rank 0)
    cpu = None
    if not args.no_cpu:
        v, med, _ = cpu_train_seq_per_s(batch=8, steps=3, warmup=1)
        cpu = {"value": v, "unit": "seq/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "3 timed steps of batch 8 (fwd+bwd+Adam, dropout off) of the oracle port, torch fp32 CPU"}
    line = {
        "metric": "sasrec_train_seq_per_s", "value": seq_s, "unit": "seq/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: SASRec L=200 d=128 H=2 blocks=2 |I|=50K, full-catalog CE + Adam, "
                               f"dropout {c['dropout']}, MovieLens-shaped synthetic windows (inputs > L2: ~2 GB of activations per step)",
                   "global_batch": world * B, "per_gpu_batch": B, "seq_len": L, "d": d, "n_items": I,
                   "parallelism": f"dp{world}", "valid_targets_per_seq": valid_per_seq, "cuda_graph": not args.no_graph},
        "e2e": {"value": seq_s_e2e, "unit": "seq/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / K},
        "e2e_device_batches": dev_batches,
        "gpu_launches": (tr.launches_per_step or 0) * K,
        "clocks": clocks,
        "roofline": roof,
        "step_roofline": {"credited_flops_per_seq": fl_seq, "achieved_tflops_per_gpu": step_tflops,
                          "peak": PK["tc_sustained"], "frac": step_tflops / PK["tc_sustained"],
                          "note": "whole step vs sustained bf16 peak; FLOPs per SURVEY 8d (valid targets only, x3 for train)"},
        "cpu_baseline": cpu,
        "scoring": scoring,
        "final_loss": final_loss,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-scoring", action="store_true")
    ap.add_argument("--no-device-batches", action="store_true", help="skip the device-side batch construction leg")
    ap.add_argument("--batch", type=int, default=None, help="sequences per GPU and step (default 512; SURVEY 8d sweeps {128, 256, 512})")
    ap.add_argument("--dropout", type=float, default=None, help="diagnostic override of the workload's dropout (0.2); "
                    "a run with this flag is not the benchmark configuration")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
