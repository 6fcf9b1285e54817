"""The host mirrors of the sketch classes against known answers recorded from the UNMODIFIED reference classes
(tests/golden/sketch_kats.npz, written by tests/golden/gen_sketch_kats.py): add(), merge() and every query,
floats compared bitwise."""
import os

import numpy as np
import pytest

import happysim_b200 as hs

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sketch_kats.npz"))
KEYS = [int(k) for k in Z["keys"]]
A, B = KEYS[:1700], KEYS[1700:]
QS = [0.0, 0.001, 0.01, 0.1, 0.25, 0.5, 0.75, 0.9, 0.99, 0.999, 1.0]


def bits(x):
    return np.array(x, dtype=np.float64).view(np.int64).tolist()


def fed(make, items):
    s = make()
    for k in items:
        s.add(k)
    return s


@pytest.mark.parametrize("p", [4, 9, 14])
def test_hyperloglog(p):
    h, h1, h2 = (fed(lambda: hs.HyperLogLog(p, seed=p), it) for it in (KEYS, A, B))
    h1.merge(h2)
    assert np.array_equal(h._registers, Z[f"hll{p}_regs"]) and np.array_equal(h1._registers, h._registers)
    assert [h.cardinality(), h1.cardinality(), hs.HyperLogLog(p).cardinality(), h1.item_count] == Z[f"hll{p}_card"].tolist()
    assert h.standard_error() == 1.04 / np.sqrt(1 << p)


def test_count_min_sketch():
    mk = lambda: hs.CountMinSketch(37, 4, seed=3)
    c, c1, c2 = fed(mk, KEYS), fed(mk, A), fed(mk, B)
    c1.merge(c2)
    assert np.array_equal(c._counters.astype(np.int64), Z["cms_counters"])
    assert [c1.estimate(k) for k in range(500)] == Z["cms_est"].tolist()
    d1, d2 = hs.CountMinSketch.from_error_rate(0.01, 0.01), hs.CountMinSketch.from_error_rate(0.2, 0.5)
    assert [d1.width, d1.depth, d2.width, d2.depth] == Z["cms_dims"].tolist()


def test_bloom_filter():
    mk = lambda: hs.BloomFilter.from_expected_items(300, 0.02, seed=5)
    bf, b1, b2 = fed(mk, KEYS), fed(mk, A), fed(mk, B)
    b1.merge(b2)
    e = hs.BloomFilter.from_expected_items(0, 0.5)
    assert [bf.size_bits, bf.num_hashes, e.size_bits, e.num_hashes, hs.BloomFilter(100).num_hashes] == Z["bloom_cfg"].tolist()
    assert np.array_equal(bf._bits, Z["bloom_bits"])
    assert [int(b1.contains(k)) for k in range(600)] == Z["bloom_contains"].tolist()
    assert bits([bf.fill_ratio, bf.false_positive_rate, float(b1._bits_set), float(b1.item_count)]) == Z["bloom_stats"].tolist()
    assert all(k in bf for k in KEYS[:50])


@pytest.mark.parametrize("k", [3, 25, 600])
def test_topk_space_saving(k):
    t, t1, t2 = (fed(lambda: hs.TopK(k), it) for it in (KEYS, A, B))
    assert [[it, c[0], c[1]] for it, c in t._counters.items()] == Z[f"topk{k}_state"].tolist()
    t1.merge(t2)
    got = [[fe.item, fe.count, fe.error] for fe in t1.top()] + [[t1.item_count, t1.max_error(), t1.guaranteed_threshold()]]
    assert got == Z[f"topk{k}_merged"].tolist()
    fe = t.estimate_with_error(499)
    assert [t.estimate(0), t.estimate(498), fe.count, fe.error, int(0 in t), t.tracked_count] == Z[f"topk{k}_misc"].tolist()


@pytest.mark.parametrize("comp", [10.0, 100.0])
def test_tdigest(comp):
    vals = [float(v) for v in Z["vals"]]
    d, d1, d2 = (fed(lambda: hs.TDigest(comp), it) for it in (vals, vals[:1234], vals[1234:]))
    cen = Z[f"td{int(comp)}_centroids"]
    assert bits(d._means) == bits(cen[0]) and d._counts == [int(x) for x in cen[1]]
    assert bits(d._buffer) == bits(Z[f"td{int(comp)}_buffer"])
    pre = (len(d._means), len(d._buffer))
    got = [d.quantile(q) for q in QS] + [d.cdf(v) for v in (-1.0, 0.0, 0.05, 0.25, 0.3, 1.0, 5.0, 6.0)] + \
          [float(d.centroid_count), float(pre[0]), float(pre[1]), d.min, d.max, d.percentile(99.9)]
    assert bits(got) == Z[f"td{int(comp)}_q"].tolist()
    d1.merge(d2)
    got = [d1.quantile(q) for q in QS] + [float(d1.centroid_count), float(d1.item_count), d1.min, d1.max]
    assert bits(got) == Z[f"td{int(comp)}_merged"].tolist()


def test_tdigest_edges():
    one = hs.TDigest(50.0); one.add(3.5)
    assert bits([one.quantile(0.0), one.quantile(0.3), one.quantile(1.0), one.cdf(3.5), one.cdf(1.0)]) == Z["td_single"].tolist()
    with pytest.raises(ValueError):
        hs.TDigest(10.0).quantile(0.5)
    with pytest.raises(ValueError):
        one.quantile(1.5)
    with pytest.raises(ValueError):
        hs.TDigest(0)
    assert hs.TDigest(10.0).cdf(1.0) == 0.0


# ---- the same streams through the shared C steps (csrc/hs_sketch.h: what the kernels and the oracle execute)
import oracle_lib as O                                   # noqa: E402
from happysim_b200 import _abi as ABI                    # noqa: E402


def c_fed(algo, state, tab, p_or_depth, width, K, items):
    L = O.lib()
    tp = tab.ctypes.data if tab is not None else None
    for k in items:
        L.hs_cpu_sketch_add(state.ctypes.data, tp, algo, p_or_depth, width, K, k)
    return state


def test_shared_c_steps_on_the_kat_streams():
    K = 500
    for p in (4, 9, 14):
        regs = c_fed(ABI.HS_SK_HLL, np.zeros(1 << p, np.uint8), hs.hll_table(p, p, K), p, 0, K, KEYS)
        assert np.array_equal(regs, Z[f"hll{p}_regs"])
    cnt = c_fed(ABI.HS_SK_CMS, np.zeros(4 * 37, np.uint32), hs.cms_table(37, 4, 3, K), 4, 37, K, KEYS)
    assert np.array_equal(cnt.reshape(4, 37).astype(np.int64), Z["cms_counters"])
    m, nh = int(Z["bloom_cfg"][0]), int(Z["bloom_cfg"][1])
    words = c_fed(ABI.HS_SK_BLOOM, np.zeros((m + 63) // 64, np.uint64), hs.bloom_table(m, nh, 5, K), nh, m, K, KEYS)
    assert np.array_equal(words, Z["bloom_bits"])
    for k in (3, 25, 600):
        st = c_fed(ABI.HS_SK_TOPK, np.zeros(16 + 12 * k, np.uint8), None, k, 0, K, KEYS)
        n = int(st[:4].view(np.uint32)[0])
        assert st[16: 16 + 12 * n].view(np.int32).reshape(n, 3).tolist() == Z[f"topk{k}_state"].tolist()


@pytest.mark.parametrize("comp", [10.0, 100.0])
def test_shared_c_tdigest_on_the_kat_stream(comp):
    """3 200 values with ties and sorted runs: buffer sort, the stable merge rule for equal means and the
    compression arithmetic of csrc/hs_sketch.h give the reference's centroids bit for bit."""
    L = O.lib()
    buf, cap = int(comp * 2), 2 * int(comp * 2)
    state = np.zeros(32 + cap * 16 + (buf * 8 + 15) // 16 * 16, np.uint8)
    for v in Z["vals"]:
        assert L.hs_cpu_tdigest_add(state.ctypes.data, comp, buf, cap, float(v)) == 1
    d = hs.TDigest(comp); d._load_device_state(state, capacity=cap)
    cen = Z[f"td{int(comp)}_centroids"]
    assert bits(d._means) == bits(cen[0]) and d._counts == [int(x) for x in cen[1]]
    assert bits(d._buffer) == bits(Z[f"td{int(comp)}_buffer"]) and d.item_count == len(Z["vals"])
    assert bits([d.min, d.max]) == Z[f"td{int(comp)}_q"].tolist()[-3:-1]
