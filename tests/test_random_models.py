"""Randomised differential testing.  CPU part: every generated model validates and runs on the oracle (and is
deterministic); the GPU part (tests/test_gpu_random_models.py) runs the same models on both general engines."""
import pytest

import oracle_lib as O
from happysim_b200 import engine
from random_models import random_model

SEEDS = list(range(64))            # run on the device engines too (tests/test_gpu_random_models.py)
REF_SEEDS = list(range(256))       # oracle vs the unmodified reference (CPU only)


@pytest.mark.parametrize("seed", SEEDS)
def test_random_model_validates_and_runs_on_the_oracle(seed):
    model, end_s, what = random_model(seed)
    engine.validate_model(model)
    kw = dict(seed=1000 + seed, end_ns=int(end_s * 1e9), n_replicas=3, record_cap=12000, sample_cap=1500, service_cap=1500)
    a = O.oracle_run(model, O.make_params(**kw))
    b = O.oracle_run(random_model(seed)[0], O.make_params(**kw))
    assert a["summaries"].tobytes() == b["summaries"].tobytes(), what
    assert int(a["summaries"]["events_processed"].min()) > 0, what
    assert int(a["summaries"]["events_processed"].max()) < 12000, what      # the recorder holds the whole run


# ---- the same models on the UNMODIFIED reference (tests/golden/random_models.npz, gen_random_golden.py) ---------
import os                                    # noqa: E402

import numpy as np                           # noqa: E402

REF = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "random_models.npz"))


def reference_answer(seed):
    """(summary row, entity stats, canonical sketch bytes) the reference produced for replica word 0, or None for
    the models that have no reference counterpart (arbitrary key table)."""
    if seed not in REF["seeds"]:
        return None
    return REF[f"s{seed}_summary"][0], REF[f"s{seed}_stats"][0], REF[f"s{seed}_sketches"]


def check_against_reference(model, out, seed, r=0):
    ans = reference_answer(seed)
    if ans is None:
        return False
    ws, wstats, wsk = ans
    s = out["summaries"][r]
    for f in ("events_processed", "final_time_ns", "order_hash", "heap_left", "n_sink_samples", "n_service_samples"):
        assert int(s[f]) == int(ws[f]), (seed, f, int(s[f]), int(ws[f]))
    assert out["entity_stats"][r].tobytes() == wstats.tobytes(), (seed, "entity statistics")
    if len(wsk):
        assert model.canonical_sketches(out["sketches"])[r].tobytes() == wsk.tobytes(), (seed, "sketch states")
    return True


@pytest.mark.parametrize("seed", REF_SEEDS)
def test_oracle_matches_the_reference_on_random_models(seed):
    """Order hash over every processed event, counts, statistics and sketch states of the random models, as the
    unmodified reference produced them with the Philox plug-ins."""
    model, end_s, what = random_model(seed)
    out = O.oracle_run(model, O.make_params(seed=1000 + seed, end_ns=int(end_s * 1e9), n_replicas=1))
    if not check_against_reference(model, out, seed):
        pytest.skip("arbitrary key table: no reference counterpart")


LANE_SEEDS = list(range(48))
REF_LANE_SEEDS = list(range(160))


@pytest.mark.parametrize("seed", LANE_SEEDS)
def test_random_lane_model_runs_on_the_oracle(seed):
    from random_models import random_lane_model
    model, end_s, what = random_lane_model(seed)
    engine.validate_model(model)
    out = O.oracle_run(model, O.make_params(seed=77 + seed, end_ns=int(end_s * 1e9), n_replicas=2))
    assert 0 < int(out["summaries"]["events_processed"].max()) < 12000, what


@pytest.mark.parametrize("seed", REF_LANE_SEEDS)
def test_oracle_matches_the_reference_on_random_single_server_models(seed):
    from random_models import random_lane_model
    model, end_s, what = random_lane_model(seed)
    out = O.oracle_run(model, O.make_params(seed=77 + seed, end_ns=int(end_s * 1e9), n_replicas=1))
    s, ws = out["summaries"][0], REF[f"lane{seed}_summary"][0]
    for f in ("events_processed", "final_time_ns", "order_hash", "heap_left", "n_sink_samples", "n_service_samples"):
        assert int(s[f]) == int(ws[f]), (what, f, int(s[f]), int(ws[f]))
    assert out["entity_stats"][0].tobytes() == REF[f"lane{seed}_stats"][0].tobytes(), what


# ---- second generator: step profiles and CachingServer farms (tests/golden/random_models_v2.npz) -------------------
SEEDS_V2 = list(range(48))          # also on the device engines (tests/test_gpu_random_models.py)
REF_SEEDS_V2 = list(range(120))

_v2_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "random_models_v2.npz")
REF_V2 = np.load(_v2_path) if os.path.exists(_v2_path) else None


def check_against_reference_v2(out, seed, r=0):
    ws, wstats, wsk = REF_V2[f"s{seed}_summary"][0], REF_V2[f"s{seed}_stats"][0], REF_V2[f"s{seed}_sketches"]
    s = out["summaries"][r]
    for f in ("events_processed", "final_time_ns", "order_hash", "heap_left", "n_sink_samples", "n_service_samples"):
        assert int(s[f]) == int(ws[f]), (seed, f, int(s[f]), int(ws[f]))
    assert out["entity_stats"][r].tobytes() == wstats.tobytes(), (seed, "entity statistics")
    if len(wsk):
        assert out["sketches"][r].tobytes() == wsk.tobytes(), (seed, "TTL cache states")


@pytest.mark.parametrize("seed", REF_SEEDS_V2)
def test_oracle_matches_the_reference_on_random_step_profile_and_cache_models(seed):
    from random_models import random_model_v2
    model, end_s, what = random_model_v2(seed)
    engine.validate_model(model)
    out = O.oracle_run(model, O.make_params(seed=2000 + seed, end_ns=int(end_s * 1e9), n_replicas=1))
    check_against_reference_v2(out, seed)
