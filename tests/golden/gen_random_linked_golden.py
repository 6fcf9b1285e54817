"""Reference results for the random linked ParallelSimulations of tests/random_models.random_linked_model:
the unmodified reference's ParallelSimulation + WindowedCoordinator, Philox streams injected
(ref_harness.run_reference_linked).  Per (seed, partition): events processed, final time, order hash over every
processed event, events left in the heap, sample counts, a digest of the entity statistics; per seed: windows and
delivered cross-partition events.

    python tests/golden/gen_random_linked_golden.py     # needs /root/reference; writes tests/golden/random_linked.npz
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import conftest  # noqa: F401,E402
import random_models as RM  # noqa: E402
import ref_harness as RH  # noqa: E402

ROW = np.dtype([("seed", "<i4"), ("part", "<i4"), ("events_processed", "<i8"), ("final_time_ns", "<i8"), ("order_hash", "<u8"),
                ("heap_left", "<i4"), ("n_sink_samples", "<i8"), ("n_service_samples", "<i8"), ("stats_digest", "<u8")])
TOP = np.dtype([("seed", "<i4"), ("windows", "<i4"), ("delivered", "<i8"), ("total_events", "<i8")])


def digest(a) -> int:
    return int.from_bytes(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], "little")


def main():
    rows, tops = [], []
    for seed in range(RM.LINKED_SEEDS):
        lm, end_s, what = RM.random_linked_model(seed)
        outs, summ = RH.run_reference_linked(lm, seed=1000 + seed, end_ns=int(end_s * 1e9))
        for q, o in enumerate(outs):
            s = o["summaries"][0]
            rows.append((seed, q, int(s["events_processed"]), int(s["final_time_ns"]), int(s["order_hash"]), int(s["heap_left"]),
                         int(s["n_sink_samples"]), int(s["n_service_samples"]), digest(o["entity_stats"][0])))
        tops.append((seed, summ.total_windows, summ.total_cross_partition_events, summ.total_events_processed))
        print(what, "->", summ.total_events_processed, "events,", summ.total_cross_partition_events, "delivered")
    np.savez_compressed(os.path.join(HERE, "random_linked.npz"), rows=np.array(rows, dtype=ROW), tops=np.array(tops, dtype=TOP))


if __name__ == "__main__":
    main()
