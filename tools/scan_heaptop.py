"""Thread engine: shared-memory heap-top size scan on configs[2] (HS_THREAD_HEAPTOP override)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import happysim_b200 as hs
from happysim_b200 import engine

def run(n, top, rpw=None):
    if top is None: os.environ.pop("HS_THREAD_HEAPTOP", None)
    else: os.environ["HS_THREAD_HEAPTOP"] = str(top)
    if rpw is None: os.environ.pop("HS_THREAD_RPW", None)
    else: os.environ["HS_THREAD_RPW"] = str(rpw)
    eng = engine.Engine(0); eng.upload(hs.lb_round_robin(64, 512.0)); best = None
    for _ in range(3):
        eng.run(engine.make_params(seed=1234, end_ns=int(10e9), n_replicas=n, flags=0, engine=3)); eng.sync()
        ms = eng.last_run_ms(); best = ms if best is None else min(best, ms)
    ev = int(eng.read_outputs()["summaries"]["events_processed"].sum())
    print(f"configs[2] n={n:6d} rpw={rpw} heap_top={top}: {best:8.2f} ms {ev / best / 1e6:7.3f} Gev/s", flush=True)
    eng.close()

for top in (None, 5, 21, 85): run(16384, top)
for top in (5, 21, 85): run(16384, top, rpw=16)
for top in (None, 0, 5): run(65536, top)
for top in (None, 5, 21, 85): run(32768, top)
