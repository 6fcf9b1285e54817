"""Quick device-time probe of the engine (not the bench): events/s for a few shapes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import happysim_b200 as hs
from happysim_b200 import engine

def probe(name, model, n, end_s, flags=0, reps=3, **kw):
    eng = engine.Engine(0)
    eng.upload(model)
    best = None
    for i in range(reps):
        eng.run(engine.make_params(seed=1234, end_ns=int(end_s * 1e9), n_replicas=n, flags=flags, **kw))
        eng.sync()
        ms = eng.last_run_ms()
        best = ms if best is None else min(best, ms)
    out = eng.read_outputs()
    ev = int(out["summaries"]["events_processed"].sum())
    bad = int((out["summaries"]["status"] != 0).sum())
    print(f"{name:28s} n={n:7d} end={end_s:8.0f}s flags={flags} events={ev:.3e} best={best:9.3f} ms  "
          f"{ev / best / 1e6:9.2f} Gev/s  flagged={bad}", flush=True)
    eng.close()

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "lane"
    if which == "lane":
        for n in (65536, 262144):
            for fl in (0, 1):
                probe("mm1", hs.mm1(), n, 200.0, flags=fl)
        probe("mm1 rec", hs.mm1(), 65536, 2000.0, flags=0, record_cap=1024, sample_cap=128, service_cap=128)
        probe("mm1 rec-only", hs.mm1(), 65536, 2000.0, flags=0, record_cap=1024)
        probe("mm1 long", hs.mm1(), 65536, 2000.0, flags=0)
