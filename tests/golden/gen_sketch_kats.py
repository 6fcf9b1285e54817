"""Known answers of the reference's own sketch classes on fixed streams (run in the build container):

    python tests/golden/gen_sketch_kats.py          # needs /root/reference

-> tests/golden/sketch_kats.npz: for each case the input stream and what the UNMODIFIED reference classes
(happysimulator/sketching/*.py) answer after add() and after merge().  tests/test_sketch_kats.py replays the
streams into the host mirrors (happy-simulator_b200/sketching.py) and compares -- floats bitwise."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get("HS_REFERENCE_ROOT", "/root/reference"))

from happysimulator.sketching.bloom_filter import BloomFilter          # noqa: E402
from happysimulator.sketching.count_min_sketch import CountMinSketch   # noqa: E402
from happysimulator.sketching.hyperloglog import HyperLogLog           # noqa: E402
from happysimulator.sketching.tdigest import TDigest                   # noqa: E402
from happysimulator.sketching.topk import TopK                         # noqa: E402

QS = [0.0, 0.001, 0.01, 0.1, 0.25, 0.5, 0.75, 0.9, 0.99, 0.999, 1.0]
out = {}
rng = np.random.RandomState(20260922)


def bits(x):
    return np.array(x, dtype=np.float64).view(np.int64)


# ---- integer-keyed sketches: a skewed stream, split in two halves that are also merged
keys = np.minimum((rng.pareto(1.1, size=3000) * 3).astype(np.int64), 499)
out["keys"] = keys
a, b = keys[:1700], keys[1700:]
for p in (4, 9, 14):
    h, h1, h2 = HyperLogLog(p, seed=p), HyperLogLog(p, seed=p), HyperLogLog(p, seed=p)
    for k in keys: h.add(int(k))
    for k in a: h1.add(int(k))
    for k in b: h2.add(int(k))
    h1.merge(h2)
    out[f"hll{p}_regs"] = np.array(h._registers, dtype=np.uint8)
    out[f"hll{p}_card"] = np.array([h.cardinality(), h1.cardinality(), HyperLogLog(p).cardinality(), h1.item_count])
    assert h1._registers == h._registers
c, c1, c2 = CountMinSketch(37, 4, seed=3), CountMinSketch(37, 4, seed=3), CountMinSketch(37, 4, seed=3)
for k in keys: c.add(int(k))
for k in a: c1.add(int(k))
for k in b: c2.add(int(k))
c1.merge(c2)
out["cms_counters"] = np.array(c._counters, dtype=np.int64)
out["cms_est"] = np.array([c1.estimate(k) for k in range(500)], dtype=np.int64)
out["cms_dims"] = np.array([CountMinSketch.from_error_rate(0.01, 0.01).width, CountMinSketch.from_error_rate(0.01, 0.01).depth,
                            CountMinSketch.from_error_rate(0.2, 0.5).width, CountMinSketch.from_error_rate(0.2, 0.5).depth])
bf, b1, b2 = (BloomFilter.from_expected_items(300, 0.02, seed=5) for _ in range(3))
for k in keys: bf.add(int(k))
for k in a: b1.add(int(k))
for k in b: b2.add(int(k))
b1.merge(b2)
out["bloom_cfg"] = np.array([bf.size_bits, bf.num_hashes, BloomFilter.from_expected_items(0, 0.5).size_bits,
                             BloomFilter.from_expected_items(0, 0.5).num_hashes, BloomFilter(100).num_hashes])
out["bloom_bits"] = np.array(bf._bits, dtype=np.uint64)
out["bloom_contains"] = np.array([int(b1.contains(k)) for k in range(600)], dtype=np.int64)
out["bloom_stats"] = bits([bf.fill_ratio, bf.false_positive_rate, float(b1._bits_set), float(b1.item_count)])
for kk in (3, 25, 600):
    t, t1, t2 = TopK(kk), TopK(kk), TopK(kk)
    for k in keys: t.add(int(k))
    for k in a: t1.add(int(k))
    for k in b: t2.add(int(k))
    out[f"topk{kk}_state"] = np.array([[cn.item, cn.count, cn.error] for cn in t._counters.values()], dtype=np.int64)
    t1.merge(t2)
    out[f"topk{kk}_merged"] = np.array([[fe.item, fe.count, fe.error] for fe in t1.top()] +
                                       [[t1.item_count, t1.max_error(), t1.guaranteed_threshold()]], dtype=np.int64)
    fe = t.estimate_with_error(499)
    out[f"topk{kk}_misc"] = np.array([t.estimate(0), t.estimate(498), fe.count, fe.error, int(0 in t), t.tracked_count])

# ---- TDigest: latencies-like floats, ties, sorted runs; compression small and default; merge
vals = np.concatenate([rng.exponential(0.1, 2500), np.full(300, 0.25), np.sort(rng.uniform(0, 2, 400)), [0.0, 5.0]])
rng.shuffle(vals)
out["vals"] = vals
for comp in (10.0, 100.0):
    d, d1, d2 = TDigest(comp), TDigest(comp), TDigest(comp)
    for v in vals: d.add(float(v))
    for v in vals[:1234]: d1.add(float(v))
    for v in vals[1234:]: d2.add(float(v))
    pre = (len(d._centroids), len(d._buffer))
    out[f"td{int(comp)}_centroids"] = np.array([[c.mean for c in d._centroids], bits([float(c.count) for c in d._centroids]).view(np.float64)])
    out[f"td{int(comp)}_buffer"] = np.array(d._buffer, dtype=np.float64)
    out[f"td{int(comp)}_q"] = bits([d.quantile(q) for q in QS] + [d.cdf(v) for v in (-1.0, 0.0, 0.05, 0.25, 0.3, 1.0, 5.0, 6.0)] +
                                   [float(d.centroid_count), float(pre[0]), float(pre[1]), d.min, d.max, d.percentile(99.9)])
    d1.merge(d2)
    out[f"td{int(comp)}_merged"] = bits([d1.quantile(q) for q in QS] + [float(d1.centroid_count), float(d1.item_count), d1.min, d1.max])
one = TDigest(50.0); one.add(3.5)
out["td_single"] = bits([one.quantile(0.0), one.quantile(0.3), one.quantile(1.0), one.cdf(3.5), one.cdf(1.0)])
np.savez_compressed(os.path.join(HERE, "sketch_kats.npz"), **out)
print("wrote sketch_kats.npz:", len(out), "arrays")
