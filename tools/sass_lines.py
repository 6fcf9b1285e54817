"""Per-source-line SASS instruction counts of one kernel: nvdisasm -g output on stdin.
usage: nvdisasm -g -fun <mangled> X.cubin | python tools/sass_lines.py [file-substring] [lo hi]"""
import sys, re, collections
want = sys.argv[1] if len(sys.argv) > 1 else "hs_lane_engine"
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
cur = None; cnt = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for line in sys.stdin:
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1), int(m.group(2))); continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        cnt[cur] += 1; ops[cur][m.group(2).split(".")[0]] += 1
tot = 0; region = collections.Counter()
for (f, l), c in sorted(cnt.items()):
    if want in f and lo <= l <= hi:
        tot += c; region.update(ops[(f, l)])
        print(f"{l:5d} {c:4d}  " + " ".join(f"{k}:{v}" for k, v in ops[(f, l)].most_common(6)))
print("total in range", tot, dict(region.most_common(15)))
