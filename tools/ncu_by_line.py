"""Join an ncu SASS source page with nvdisasm -g line info (same build) and aggregate per source line.
usage: python tools/ncu_by_line.py X.ncu-rep kernel.sass [events] [min_pct]
  kernel.sass = the kernel's section of `nvdisasm -g lib.cubin` (instruction order must match)."""
import csv, re, subprocess, sys, collections
rep, sass = sys.argv[1], sys.argv[2]
ev = float(sys.argv[3]) if len(sys.argv) > 3 else None
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.4
lines = []; cur = None
for l in open(sass):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);', l)
    if m: lines.append((cur, m.group(2)))
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr = None; ins = []
for r in rows:
    if r and r[0] == "Address": hdr = {k: i for i, k in enumerate(r)}; continue
    if hdr and r and r[0].startswith("0x"): ins.append(r)
assert len(ins) == len(lines), (len(ins), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0, 0]); tot = [0, 0, 0]
for (src, _), r in zip(lines, ins):
    ie = int(r[hdr["Instructions Executed"]]); te = int(r[hdr["Thread Instructions Executed"]]); sm = int(r[hdr["# Samples"]])
    a = agg[src]; a[0] += ie; a[1] += te; a[2] += sm; a[3] += 1
    tot[0] += ie; tot[1] += te; tot[2] += sm
print(f"total warp-inst {tot[0]:.4g}" + (f" = {tot[0]/ev:.3f}/event" if ev else "") + f", avg lanes {tot[1]/tot[0]:.1f}")
for src, a in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if 100 * a[0] / tot[0] >= minpct or 100 * a[2] / tot[2] >= 2 * minpct:
        print(f"{src[0]:22s}:{src[1]:4d}  sass {a[3]:4d}  warp-inst {100*a[0]/tot[0]:5.2f}%" + (f" {a[0]/ev:6.3f}/ev" if ev else "") + f"  lanes {a[1]/max(a[0],1):5.1f}  samples {100*a[2]/tot[2]:5.2f}%")
