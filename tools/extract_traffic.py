"""dram__bytes_read.sum + dram__bytes_write.sum per launch out of .ncu-rep captures -> profiles/r2_traffic.json, the file
bench.py reads `roofline.traffic` from (key = workload of the capture, e.g. ce_head_c2_b512).

    python tools/extract_traffic.py KEY rep1.ncu-rep[:kernel-regex] [rep2.ncu-rep[:kernel-regex] ...]

The value is the SUM over the listed captures' matching launches (first match per capture): the CE-head roofline is quoted
per launch PAIR (fused forward + dH pass, dE pass)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r2_traffic.json")


def dram_bytes(rep, pattern):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        if pattern and not re.search(pattern, r[kn]):
            continue
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(k)
            tot += float(r[i].replace(",", "")) * scale[units[i]]
        return tot, r[kn][:90]
    raise SystemExit(f"no launch matching {pattern!r} in {rep}")


def main():
    key, specs = sys.argv[1], sys.argv[2:]
    total, notes = 0.0, []
    for sp in specs:
        rep, _, pat = sp.partition(":")
        b, name = dram_bytes(rep, pat)
        total += b
        notes.append(f"{os.path.basename(rep)}: {name} = {b / 1e6:.1f} MB")
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    data[key] = total
    data.setdefault("_sources", {})[key] = notes
    with open(OUT, "w") as fh:
        json.dump(data, fh, indent=1, sort_keys=True)
    print(key, total, notes)


if __name__ == "__main__":
    main()
