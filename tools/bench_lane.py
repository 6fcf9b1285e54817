"""Lane-engine A/B on configs[1]: python tools/bench_lane.py  (HS_B200_LIB selects the library variant)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import happysim_b200 as hs
from happysim_b200 import engine

def run(name, end_s, hash=False, **kw):
    eng = engine.Engine(0)
    eng.upload(hs.mm1())
    best = None
    for _ in range(4):
        eng.run(engine.make_params(seed=1234, end_ns=int(end_s * 1e9), n_replicas=65536, flags=(1 if hash else 0), **kw))
        eng.sync()
        ms = eng.last_run_ms(); best = ms if best is None else min(best, ms)
    ev = int(eng.read_outputs()["summaries"]["events_processed"].sum())
    print(f"{os.path.basename(os.environ.get('HS_B200_LIB', 'default')):14s} {name:10s} ev={ev:.3e} {best:9.2f} ms {ev / best / 1e6:8.3f} Gev/s", flush=True)
    eng.close()

run("summary", 2000.0)
run("record", 2000.0, record_cap=1024, sample_cap=128, service_cap=128)
run("hash+rec", 2000.0, record_cap=1024, sample_cap=128, service_cap=128, hash=True)
