"""The shared sampler (happy-simulator_b200/csrc/hs_sampler.h) through its CPU twins."""
import ctypes as C
import math
import random

import numpy as np

import oracle_lib as O


def philox(ctr, key):
    out = (C.c_uint32 * 4)()
    O.lib().hs_cpu_philox(*ctr, *key, out)
    return [int(x) for x in out]


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors for philox4x32-10 (Salmon et al., SC'11)."""
    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_uniform_is_genrand_res53_of_the_block():
    L = O.lib()
    for seed, rid, sid, n in [(0, 0, 0, 0), (42, 7, 1 | (3 << 8), 5), (2**40 + 3, 65535, 2, 123456789012)]:
        x = philox([(n >> 1) & 0xffffffff, (n >> 1) >> 32, rid, sid], [seed & 0xffffffff, seed >> 32])
        a, b = (x[2], x[3]) if n & 1 else (x[0], x[1])
        want = ((a >> 5) * 67108864.0 + (b >> 6)) / 9007199254740992.0
        assert L.hs_cpu_uniform(seed, rid, sid, n) == want
        assert 0.0 <= want < 1.0


def test_log_within_one_ulp_of_libm():
    L = O.lib()
    rnd = random.Random(5)
    worst = 0.0
    for _ in range(100000):
        x = 1.0 - rnd.random()
        a, b = L.hs_cpu_log(x), math.log(x)
        if b != 0.0:
            worst = max(worst, abs(a - b) / math.ulp(b))
    assert worst <= 1.0
    assert L.hs_cpu_log(1.0) == 0.0
    assert L.hs_cpu_log(2.0 ** -53) == math.log(2.0 ** -53)
    for x in (0.5, 0.25, 0.7071067811865476, 0.7071067811865475, 1e-300, 5e-324):
        assert abs(L.hs_cpu_log(x) - math.log(x)) <= math.ulp(math.log(x))


def test_time_arithmetic_matches_python_semantics():
    """T1-T3 / L2 / D1 of SURVEY.md 8(a) against the reference's Python expressions."""
    L = O.lib()
    rnd = random.Random(9)
    for _ in range(20000):
        ns = rnd.randrange(0, 10**15)
        s = rnd.random() * 10 ** rnd.randrange(-9, 6)
        assert L.hs_cpu_seconds_to_ns(s) == int(s * 1_000_000_000)            # temporal.py:58-62
        assert L.hs_cpu_ns_to_seconds(ns) == float(ns) / 1_000_000_000        # temporal.py:66-68
        target, rate = rnd.expovariate(1.0), rnd.choice([8.0, 10.0, 50.0, 512.0, 0.37])
        t_next = float(ns) / 1_000_000_000 + target / rate                    # arrival_time_provider.py:70-78
        assert L.hs_cpu_next_arrival_ns(ns, target, rate) == int(t_next * 1_000_000_000)
        u, lam = rnd.random(), 1 / rnd.choice([0.1, 0.01, 0.25])
        sample = L.hs_cpu_exp1(u) / lam                                       # exponential.py:43-45
        assert L.hs_cpu_exp_latency_ns(u, lam) == int(sample * 1_000_000_000)


def test_constant_rate_tick_drift_golden_vector():
    """tests/regression/test_arrival_time_regression.py:20-31 of the reference: rate 50 ticks."""
    L = O.lib()
    t, got = 0, []
    for _ in range(10):
        t = L.hs_cpu_next_arrival_ns(t, 1.0, 50.0)
        got.append(t / 1e9)
    want = [0.02, 0.04, 0.06, 0.08, 0.1, 0.12, 0.14, 0.16, 0.18, 0.199999999]
    assert all(abs(a - b) < 1e-8 for a, b in zip(got, want))
    assert got[-1] == 0.199999999
    t, got = 0, []
    for _ in range(10):
        t = L.hs_cpu_next_arrival_ns(t, 1.0, 100.0)
        got.append(t)
    assert got[-1] == 99999999


def test_nonconstant_profile_golden_vectors_of_the_reference():
    """tests/regression/test_arrival_time_regression.py:46-103 of the reference: the first arrivals of
    ConstantArrivalTimeProvider over LinearRampProfile / SpikeProfile (adaptive Simpson + Brent)."""
    L = O.lib()

    def series(kind, p, n):
        t, out = 0, []
        for _ in range(n):
            t = L.hs_cpu_next_arrival_profile_ns(kind, *p, t, 1.0)
            out.append(t / 1e9)
        return out
    up = [0.095864499, 0.184655976, 0.267741515, 0.346097448, 0.420449859, 0.49135612, 0.55925515, 0.624499924,
          0.687379335, 0.748133387]
    down = [0.010004504, 0.020018032, 0.030040609, 0.040072259, 0.050113007, 0.060162878, 0.070221897, 0.080290089,
            0.090367479, 0.100454092]
    spike = [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.799999999, 0.899999998, 0.999999998, 1.099999998, 1.199999998,
             1.299999998, 1.399999998, 1.499999998, 1.599999998, 1.699999998, 1.799999998, 1.899999998, 1.999999997,
             2.009999997, 2.019999996, 2.029999996, 2.039999995, 2.049999994, 2.059999994, 2.069999993, 2.079999993,
             2.089999992, 2.099999991]
    assert all(abs(a - b) < 1e-8 for a, b in zip(series(1, (10.0, 10.0, 100.0, 0.0), 10), up))
    assert all(abs(a - b) < 1e-8 for a, b in zip(series(1, (10.0, 100.0, 10.0, 0.0), 10), down))
    assert all(abs(a - b) < 1e-8 for a, b in zip(series(2, (10.0, 100.0, 2.0, 1.0), 30), spike))


def test_zipf_routing_keys_follow_the_reference_inverse_transform():
    """distributions/zipf.py:96-123: cum_probs as the reference computes them (weights 1/(k+1)^s, float sum(),
    running sum, last forced to 1.0) and bisect_left over them, clamped -- hs_routing_key is that bisect."""
    import bisect
    import ctypes as C
    import happysim_b200 as hs
    L = O.lib()
    for K, s in ((1, 1.0), (7, 0.0), (50, 1.0), (1000, 1.3)):
        cum = hs.zipf_cdf(K, s)
        assert cum[-1] == 1.0 and np.all(np.diff(cum) >= 0) and len(cum) == K
        if s == 0.0:
            assert cum[0] == 1.0 / K
        else:
            w = [1.0 / ((k + 1) ** s) for k in range(K)]
            assert cum[0] == w[0] / sum(w)
        cp = cum.ctypes.data_as(C.POINTER(C.c_double))
        us = list(np.random.RandomState(K).random_sample(400)) + [0.0, float(cum[0]), float(np.nextafter(cum[0], 1)), 0.9999999999999999]
        for u in us:
            want = min(bisect.bisect_left(list(cum), u), K - 1)
            assert L.hs_cpu_routing_key(u, K, cp) == want
    assert L.hs_cpu_routing_key(0.37, 10, None) == 3          # uniform: int(u * n)
    # rank 0 is the hottest key
    cum = hs.zipf_cdf(100, 1.0)
    keys = [L.hs_cpu_routing_key(float(u), 100, cum.ctypes.data_as(C.POINTER(C.c_double))) for u in np.random.RandomState(3).random_sample(5000)]
    counts = np.bincount(keys, minlength=100)
    assert counts[0] > counts[1] > counts[5] > counts[50]


def test_ns_to_seconds_is_the_ieee_division():
    """hs_ns_to_seconds evaluates float(ns) / 1e9 as q = x * RN(1e-9), r = fma(-q, 1e9, x), fma(r, RN(1e-9), q)
    (hs_div_recip, three fp64 instructions on the device).  It must equal the IEEE quotient the reference
    computes (core/temporal.py:66-68) for every ns count: random magnitudes, exact multiples of 1e9 and
    their neighbours, powers of two and their neighbours.  (tools/divcheck.c: 5e9 more operands, 0 mismatches.)"""
    L = O.lib()
    rng = np.random.default_rng(7)
    vals = [int(v) >> int(s) for v, s in zip(rng.integers(0, 2**62, 120_000), rng.integers(0, 62, 120_000))]
    vals += [m * 10**9 + k for m in rng.integers(0, 9_000_000, 20_000).tolist() for k in (-1, 0, 1)]
    vals += [(1 << e) + k for e in range(1, 63) for k in range(-40, 41)]
    for v in vals:
        if v < 0:
            continue
        assert L.hs_cpu_ns_to_seconds(v) == float(v) / 1e9, v
