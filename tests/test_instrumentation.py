"""Data / BucketedData mirrors against the reference's own classes (when the checkout is present)
and against hand-checked values."""
import os
import random
import sys

import pytest

import happysim_b200 as hs


def test_data_aggregations_hand_checked():
    d = hs.Data()
    d._samples = [(0.1, 1.0), (0.6, 3.0), (1.2, 2.0), (2.9, 10.0)]
    assert d.count() == 4 and d.sum() == 16.0 and d.mean() == 4.0 and d.min() == 1.0 and d.max() == 10.0
    assert d.percentile(0.0) == 1.0 and d.percentile(1.0) == 10.0 and d.percentile(0.5) == 2.5
    b = d.bucket(1.0)
    assert b.times() == [0.0, 1.0, 2.0] and b.counts() == [2, 1, 1] and b.sums() == [4.0, 2.0, 10.0]
    assert d.between(0.5, 2.0).raw_values() == [3.0, 2.0]
    assert d.rate(1.0).raw_values() == [2.0, 1.0, 1.0]
    assert hs.Data().mean() == 0.0 and hs.Data().percentile(0.9) == 0.0 and not hs.Data()


@pytest.mark.skipif(not os.path.isdir("/root/reference/happysimulator"), reason="reference checkout not present")
def test_data_matches_the_reference_class():
    sys.path.insert(0, "/root/reference")
    from happysimulator.instrumentation.data import Data as RefData
    rnd = random.Random(3)
    samples = sorted((rnd.random() * 30, rnd.expovariate(2.0)) for _ in range(2000))
    a, b = hs.Data(), RefData()
    a._samples, b._samples = list(samples), list(samples)
    for f in ("mean", "min", "max", "count", "sum", "std"):
        assert getattr(a, f)() == getattr(b, f)()
    for p in (0.0, 0.5, 0.9, 0.99, 1.0):
        assert a.percentile(p) == b.percentile(p)
    assert a.bucket(2.5).to_dict() == b.bucket(2.5).to_dict()
    assert a.rate(5.0).values == b.rate(5.0).values


def test_percentile_helper_equals_the_reference_helper_float_for_float():
    import random
    import sys
    import pytest
    for d in ("/root/reference", __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))), "baseline", "_ref")):
        if __import__("os").path.isdir(__import__("os").path.join(d, "happysimulator")) and d not in sys.path:
            sys.path.insert(0, d)
    ref = pytest.importorskip("happysimulator.instrumentation.data")
    from happysim_b200.instrumentation import _percentile_sorted as mine
    rnd = random.Random(1)
    for n in (0, 1, 2, 3, 7, 100, 1001):
        v = sorted(rnd.random() * 10 for _ in range(n))
        for p in (-1, 0, 1e-9, 0.25, 0.5, 0.99, 0.999, 1, 2, rnd.random(), rnd.random()):
            assert mine(v, p) == ref._percentile_sorted(v, p), (n, p)
