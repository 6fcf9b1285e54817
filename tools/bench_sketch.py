"""Thread engine on the sketch-sink farm (lb-rr64 -> HLL | CMS) and the plain farm: A/B across library variants."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import happysim_b200 as hs
from bench_configs import run
K = 10000
b = hs.ModelBuilder()
src = b.source(rate=512.0, key_population=K)
servers = [b.server(f"S{i}", mean_service_s=0.1) for i in range(64)]
hll = b.sketch_hll("uniques", precision=12, table=hs.hll_table(12, 1, K))
cms = b.sketch_cms("freq", width=272, depth=5, table=hs.cms_table(272, 5, 2, K))
lb = b.load_balancer(backends=servers)
b.set_target(src, lb)
for k, sv in enumerate(servers):
    b.set_target(sv, hll if k % 2 else cms)
print(os.path.basename(os.environ.get("HS_B200_LIB", "default")))
run("lb-rr64 -> HLL(p=12) | CMS(272x5) sinks", b.build(), 16384, 10.0, engine=3)
t = hs.ModelBuilder()
sk = t.sink()
s2 = t.server("B", mean_service_s=0.05, target=sk) if False else None
stages = [t.server(f"T{i}", mean_service_s=0.05) for i in range(3)]
so = t.source(rate=12.0)
t.set_target(so, stages[0])
for a, c in zip(stages, stages[1:] + [sk]):
    t.set_target(a, c)
run("tandem of 3 servers (generic path)", t.build(), 65536, 50.0, engine=3)
