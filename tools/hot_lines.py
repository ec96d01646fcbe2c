"""Hottest source lines of one kernel in an .ncu-rep (needs -lineinfo + --import-source on):
   python tools/hot_lines.py REP KERNEL_REGEX [N]   -> file:line, share of stall samples, share of executed warp instructions"""
import csv, io, subprocess, sys

def main(rep, kre, n=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}",
                          "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    recs, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1]; continue
        if r and r[0] == "Line No":
            hdr = r; continue
        if hdr and len(r) == len(hdr) and r[0].strip():
            recs.append((cur, r))
    if not recs:
        print("no correlated source"); return
    ia = next(i for i, k in enumerate(hdr) if k.startswith("Warp Stall Sampling (All"))
    ii = hdr.index("Instructions Executed")
    def num(x):
        try: return float(x.replace(",", "") or 0)
        except ValueError: return 0.0
    tot = sum(num(r[ia]) for _, r in recs) or 1.0
    tot_i = sum(num(r[ii]) for _, r in recs) or 1.0
    print(f"total samples {tot:.0f}, warp instructions {tot_i:.0f}")
    for f, r in sorted(recs, key=lambda x: -num(x[1][ia]))[:n]:
        print(f"{(f or '?').split('/')[-1]}:{r[0]:>4} {100*num(r[ia])/tot:5.1f}% smp  {100*num(r[ii])/tot_i:5.1f}% ins  | {r[1].strip()[:105]}")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
