"""tests/golden/random_models_v2.npz: what the UNMODIFIED reference does on the seeded random models of
tests/random_models.py:random_model_v2 -- user-defined step profiles (the reference evaluates the StepProfile
object's own get_rate, the oracle / device the table) and CachingServer farms (the example's own class, its
own ConsistentHash ring).  Run in the build container (needs /root/reference):

    python tests/golden/gen_random_golden_v2.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]

import ref_harness as RH                      # noqa: E402
from random_models import random_model_v2     # noqa: E402
from test_random_models import REF_SEEDS_V2 as SEEDS   # noqa: E402

out = {}
for seed in SEEDS:
    m, end_s, what, ex = random_model_v2(seed, with_extras=True)
    ref = RH.run_reference(m, seed=2000 + seed, rid=0, end_ns=int(end_s * 1e9), chash_vnodes=ex["chash_vnodes"],
                           profile_objects=ex["profile_objects"])
    out[f"s{seed}_summary"] = ref["summaries"]
    out[f"s{seed}_stats"] = ref["entity_stats"]
    out[f"s{seed}_sketches"] = ref["sketches"] if "sketches" in ref else np.zeros(0, np.uint8)
    print(what, "->", int(ref["summaries"]["events_processed"][0]), "events")
out["seeds"] = np.array(SEEDS)
np.savez_compressed(os.path.join(HERE, "random_models_v2.npz"), **out)
print("wrote random_models_v2.npz for", len(SEEDS), "seeds")
