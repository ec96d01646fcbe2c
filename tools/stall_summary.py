"""Per-kernel warp-stall breakdown (stall cycles per issued instruction) from an .ncu-rep: python tools/stall_summary.py REP"""
import csv, io, subprocess, sys

def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    kn = hdr.index("Kernel Name")
    cols = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    extra = [h for h in ("smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
                         "gpu__time_duration.sum") if h in hdr]
    for r in rows[2:]:
        vals = sorted(((float(r[i].replace(",", "") or 0), hdr[i][len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")])
                       for i in cols), reverse=True)[:7]
        print(r[kn].split("(")[0][-44:], " ".join(f"{h.split('__')[1][:14]}={r[hdr.index(h)]}" for h in extra))
        print("    " + "  ".join(f"{n}={v:.2f}" for v, n in vals))

if __name__ == "__main__":
    main(sys.argv[1])
