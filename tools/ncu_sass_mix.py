"""Aggregate an ncu source page (SASS view) by opcode: warp instructions, active-lane average, stall samples.
usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_sass_mix.py [events]"""
import csv, sys, collections
rows = list(csv.reader(sys.stdin))
hdr = None; agg = collections.defaultdict(lambda: [0, 0, 0]); tot = [0, 0, 0]
conv = [0, 0]
for r in rows:
    if r and r[0] == "Address": hdr = {k: i for i, k in enumerate(r)}; continue
    if hdr is None or len(r) < len(hdr) - 2 or not r[0].startswith("0x"): continue
    sass = r[hdr["Source"]].strip()
    parts = sass.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.rstrip(";")
    key = ".".join(op.split(".")[:2]) if op.split(".")[0] in ("STG", "LDG", "STS", "LDS", "LDGSTS", "MUFU", "ST", "LD", "ATOMG", "RED", "REDG") else op.split(".")[0]
    ie = int(r[hdr["Instructions Executed"]]); te = int(r[hdr["Thread Instructions Executed"]]); sm = int(r[hdr["# Samples"]])
    agg[key][0] += ie; agg[key][1] += te; agg[key][2] += sm
    tot[0] += ie; tot[1] += te; tot[2] += sm
    if ie and te / ie > 30: conv[0] += ie
    else: conv[1] += ie
ev = float(sys.argv[1]) if len(sys.argv) > 1 else None
print(f"total warp-inst {tot[0]:.4g} thread-inst {tot[1]:.4g} avg lanes {tot[1]/max(tot[0],1):.2f} samples {tot[2]}")
print(f"warp-inst in converged (>30 lanes) code {conv[0]:.4g}  divergent {conv[1]:.4g}")
if ev: print(f"per event: warp-inst {tot[0]/ev:.3f} thread-inst {tot[1]/ev:.1f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{k:14s} warp-inst {v[0]:12.4g} ({100*v[0]/tot[0]:5.1f}%) lanes {v[1]/max(v[0],1):5.1f} samples {100*v[2]/max(tot[2],1):5.1f}%")
