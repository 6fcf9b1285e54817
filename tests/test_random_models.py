"""Randomised differential testing.  CPU part: every generated model validates and runs on the oracle (and is
deterministic); the GPU part (tests/test_gpu_random_models.py) runs the same models on both general engines."""
import pytest

import oracle_lib as O
from happysim_b200 import engine
from random_models import random_model

SEEDS = list(range(24))


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
