"""tests/golden/random_models.npz: what the UNMODIFIED reference does on the seeded random models of
tests/random_models.py (run in the build container, needs /root/reference):

    python tests/golden/gen_random_golden.py

Per seed: the summary (event count, final time, order hash over every processed event, pending events, sample
counts), the per-entity statistics and the sketch states.  Models whose load balancer uses an arbitrary random
key table have no reference counterpart (ConsistentHash computes its own ring) and are left out."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]

import ref_harness as RH                      # noqa: E402
from random_models import random_lane_model, random_model        # noqa: E402
from test_random_models import REF_LANE_SEEDS as LANE_SEEDS, REF_SEEDS as SEEDS   # noqa: E402

out = {}
kept = []
for seed in SEEDS:
    m, end_s, what, ex = random_model(seed, with_extras=True)
    if ex["random_key_table"]:
        continue
    ref = RH.run_reference(m, seed=1000 + seed, rid=0, end_ns=int(end_s * 1e9), sketch_seeds=ex["sketch_seeds"],
                           zipf_s=ex["zipf_s"])
    kept.append(seed)
    out[f"s{seed}_summary"] = ref["summaries"]
    out[f"s{seed}_stats"] = ref["entity_stats"]
    out[f"s{seed}_sketches"] = m.canonical_sketches(ref["sketches"])[0] if "sketches" in ref else np.zeros(0, np.uint8)
    print(what, "->", int(ref["summaries"]["events_processed"][0]), "events")
for seed in LANE_SEEDS:               # the single-server topology (exact tick/completion ties included)
    m, end_s, what = random_lane_model(seed)
    ref = RH.run_reference(m, seed=77 + seed, rid=0, end_ns=int(end_s * 1e9))
    out[f"lane{seed}_summary"] = ref["summaries"]
    out[f"lane{seed}_stats"] = ref["entity_stats"]
    print(what, "->", int(ref["summaries"]["events_processed"][0]), "events")
out["seeds"] = np.array(kept)
np.savez_compressed(os.path.join(HERE, "random_models.npz"), **out)
print("wrote random_models.npz for seeds", kept)
