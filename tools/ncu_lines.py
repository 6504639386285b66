"""Per-source-line stall samples of an .ncu-rep (needs -lineinfo): python tools/ncu_lines.py rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; agg = {}; cur = None
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"): cur = r[1]; continue
    if len(r) > 2 and "Warp Stall Sampling (All Samples)" in r: hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    try:
        agg[(cur, r[0], r[1][:100])] = (int(r[hdr.index("Warp Stall Sampling (All Samples)")] or 0), int(r[hdr.index("Instructions Executed")] or 0))
    except Exception: pass
tot = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print("total samples", tot, "total warp instructions", ti)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v[0]:7d} {100*v[0]/max(tot,1):5.1f}% inst={v[1]:10d} {100*v[1]/max(ti,1):5.1f}%  {(k[0] or '').split('/')[-1]}:{k[1]}  {k[2]}")
