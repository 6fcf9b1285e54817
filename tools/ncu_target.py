"""Small fixed workload for ncu captures: python tools/ncu_target.py [summary|record|warp|thread] [end_s] [replicas]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import happysim_b200 as hs
from happysim_b200 import engine
mode = sys.argv[1] if len(sys.argv) > 1 else "summary"
end_s = float(sys.argv[2]) if len(sys.argv) > 2 else 500.0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
caps = dict(record_cap=1024, sample_cap=128, service_cap=128) if mode == "record" else {}
eng = engine.Engine(0)
model = hs.lb_round_robin(64, 512.0) if mode in ("warp", "thread") else hs.mm1()
if mode == "warp":
    n = min(n, 4096)
eng.upload(model)
for i in range(3):
    eng.run(engine.make_params(seed=1234, end_ns=int(end_s * 1e9), n_replicas=n, flags=0,
                               engine={"warp": 1, "thread": 3}.get(mode, 0), **caps))
    eng.sync()
    print(mode, "ms", eng.last_run_ms())
