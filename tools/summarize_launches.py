"""Launch list (ncu --metrics gpu__time_duration.sum --csv) -> per-kernel totals of ONE training step: the rows between the
last two `prepare_count_kernel` launches (the first kernel of every step)."""
import collections
import csv
import re
import sys

src, out_md, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = []
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("rp::", "")
    rows.append((name, us))
starts = [i for i, (n, _) in enumerate(rows) if n.startswith("prepare_count_kernel")]
assert len(starts) >= 2, "need two step starts in the capture window"
seg = rows[starts[-2]:starts[-1]]
tot = collections.OrderedDict()
cnt = collections.Counter()
for n, us in seg:
    tot[n] = tot.get(n, 0.0) + us
    cnt[n] += 1
total = sum(tot.values())
with open(out_md, "w") as f:
    f.write(f"# {title}\n\n")
    f.write(f"`{src}`: one step = {len(seg)} launches, {total:.0f} us of kernel time (cold-cache, serialised under ncu: use the shares)\n\n")
    f.write("| kernel | total us | share | launches |\n|---|---|---|---|\n")
    for n, us in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write(f"| {n} | {us:.1f} | {100 * us / total:.1f}% | {cnt[n]} |\n")
print(open(out_md).read())
