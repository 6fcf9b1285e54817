"""Throughput of a linked ParallelSimulation ensemble on one GPU: the tandem of the fixture
linked_tandem_const (A: Source -> Server -> [link, 50 ms] -> B: Server(c=2) -> Sink, 50 ms windows) and the three-partition
lossy fan-out, many replicas, timed with CUDA events around the whole window loop.

    python tools/bench_linked.py [replicas ...]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import happysim_b200  # noqa: F401,E402
import golden_lib as G  # noqa: E402
from happysim_b200.linked import LinkedRun  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [4096, 16384, 65536]
    for name, end_s in (("linked_tandem_const", 20.0), ("linked_lossy_fanout", 10.0)):
        lm, kw, z = G.load_linked(name)
        for n in sizes:
            run = LinkedRun(lm)
            try:
                end_ns = int(end_s * 1e9)
                run.run(seed=kw["seed"], end_ns=int(1e9), n_replicas=n, flags=0)            # warm-up (allocations, module load)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                outs, (delivered, lost, over) = run.run(seed=kw["seed"], end_ns=end_ns, n_replicas=n, flags=0)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            finally:
                run.close()
            ev = sum(int(o["summaries"]["events_processed"].sum()) for o in outs)
            bad = sum(int((o["summaries"]["status"] != 0).sum()) for o in outs)
            bits = 0
            for o in outs:
                for x in o["summaries"]["status"]:
                    bits |= int(x)
            who = [(q, int(r)) for q, o in enumerate(outs) for r in (o["summaries"]["status"] != 0).nonzero()[0][:3]]
            print(json.dumps({"model": name, "partitions": lm.n_partitions, "replicas": n, "sim_s": end_s, "windows": run.windows,
                              "events": ev, "cross_partition_events": int(delivered.sum()), "lost": int(lost.sum()),
                              "inbox_overflows": int(over.sum()), "flagged": bad, "status_bits": bits, "flagged_where": who, "wall_ms": round(dt * 1e3, 2),
                              "events_per_s": round(ev / dt, 1), "us_per_window": round(dt * 1e6 / run.windows, 1)}), flush=True)


if __name__ == "__main__":
    main()
