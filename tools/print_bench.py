"""Pretty-print the headline fields of a bench.py JSON line (file path or stdin)."""
import json
import sys

d = json.load(open(sys.argv[1])) if len(sys.argv) > 1 else json.loads(sys.stdin.read())
print(f"train   {d['value']:.0f} {d['unit']} ({d['ms_per_step']:.3f} ms/step, {d['n_gpus']} GPU, batch {d['config'].get('global_batch')})")
if d.get("e2e"):
    print(f"e2e     {d['e2e']['value']:.0f}   device-built batches {(d.get('e2e_device_batches') or {}).get('value', float('nan')):.0f}")
if d.get("roofline"):
    print(f"CE head {d['roofline']['achieved']:.0f} TFLOP/s = {d['roofline']['frac']:.3f} of {d['roofline']['peak']}; step {d['step_roofline']['frac']:.3f}")
s = d.get("scoring")
if s:
    print(f"predict {s['value']:.0f} {s['unit']} ({s['ms_per_call']:.3f} ms/call), head {s['roofline']['head_ms']:.3f} ms = {s['roofline']['frac']:.3f}")
    if s.get("cpu_baseline"):
        print(f"cpu     {d['cpu_baseline']['value']:.1f} seq/s, {s['cpu_baseline']['value']:.0f} users/s ({d['cpu_baseline']['cores']} threads)")
print("launches", d.get("gpu_launches"), "clocks", d.get("clocks"))
