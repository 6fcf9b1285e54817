"""72 random ParallelSimulations with PartitionLinks (tests/random_models.random_linked_model): the oracle's coordinator
against the unmodified reference's (tests/golden/random_linked.npz, gen_random_linked_golden.py) -- per partition the
event count, the final clock, the order hash over every processed event, what is left in the heap, sample counts and the
entity statistics; per run the windows and the delivered cross-partition events."""
import hashlib
import os

import numpy as np
import pytest

import golden_lib as G
import oracle_lib as O
import random_models as RM

Z = np.load(os.path.join(G.GOLDEN_DIR, "random_linked.npz"))


def digest(a) -> int:
    return int.from_bytes(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], "little")


def oracle(lm, seed, end_ns, **kw):
    nP = lm.n_partitions
    ps = [O.make_params(seed=1000 + seed, end_ns=end_ns, rid_base=q, rid_stride=nP + 1, **kw) for q in range(nP)]
    return O.oracle_run_linked(lm, ps, end_ns=end_ns, cseed=1000 + seed)


@pytest.mark.parametrize("seed", range(RM.LINKED_SEEDS))
def test_oracle_equals_the_reference_on_random_linked_models(seed):
    lm, end_s, what = RM.random_linked_model(seed)
    outs, delivered, lost, ends = oracle(lm, seed, int(end_s * 1e9))
    top = Z["tops"][Z["tops"]["seed"] == seed][0]
    assert (len(ends), int(delivered[0])) == (int(top["windows"]), int(top["delivered"])), what
    rows = Z["rows"][Z["rows"]["seed"] == seed]
    assert len(rows) == lm.n_partitions
    for q, o in enumerate(outs):
        s, w = o["summaries"][0], rows[rows["part"] == q][0]
        for f in ("events_processed", "final_time_ns", "order_hash", "heap_left", "n_sink_samples", "n_service_samples"):
            assert int(s[f]) == int(w[f]), (what, q, f, int(s[f]), int(w[f]))
        assert int(s["status"]) & ~4 == 0 and digest(o["entity_stats"][0]) == int(w["stats_digest"]), (what, q)


def test_the_generator_covers_what_it_is_meant_to():
    kinds = {"grid": 0, "loss": 0, "expo": 0, "shared": 0, "back": 0, "lb_dest": 0, "four": 0}
    for seed in range(RM.LINKED_SEEDS):
        lm, _, what = RM.random_linked_model(seed)
        links = [l for ls in lm.links for l in ls]
        kinds["grid"] += "grid" in what
        kinds["loss"] += any(l.packet_loss > 0 for l in links)
        kinds["expo"] += any(l.latency_kind == 1 for l in links)
        kinds["shared"] += len({l.stream for l in links}) < len(links)
        kinds["back"] += any(l.dest < p for p, ls in enumerate(lm.links) for l in ls)
        kinds["four"] += lm.n_partitions == 4
        for p, m in enumerate(lm.models):
            for i in m.ids_of(9):
                d = lm.models[lm.links[p][int(m.entities["i0"][i])].dest]
                kinds["lb_dest"] += int(d.entities["kind"][int(m.entities["i1"][i])]) == 5
    assert all(v >= 5 for v in kinds.values()), kinds
    assert int(Z["tops"]["delivered"].min()) >= 0 and int((Z["tops"]["delivered"] > 50).sum()) > 50
