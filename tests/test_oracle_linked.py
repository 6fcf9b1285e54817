"""SURVEY 8(f) row 4, first half -- partitions joined by PartitionLinks: the oracle's restatement of
ParallelSimulation / WindowedCoordinator (oracle/hs_oracle.c, hs_oracle_run_linked) against fixtures produced by the
unmodified reference (tests/golden/gen_linked_golden.py): per partition the event order, every record, the entity
statistics, Sink samples and service times; the coordinator's window and delivery counts."""
import numpy as np
import pytest

import golden_lib as G
import oracle_lib as O
from happysim_b200 import _abi as A
from happysim_b200.linked import LinkedModel, LinkSpec


def run_oracle(lm, kw, z, n_replicas=1, replica_index_base=0):
    nP = lm.n_partitions
    ps = [O.make_params(seed=kw["seed"], end_ns=kw["end_ns"], n_replicas=n_replicas, rid_base=q, rid_stride=nP + 1,
                        replica_index_base=replica_index_base, **G.linked_caps(z, q)) for q in range(nP)]
    return O.oracle_run_linked(lm, ps, end_ns=kw["end_ns"], cseed=kw["seed"])


@pytest.mark.parametrize("name", G.case_names("linked_"))
def test_oracle_reproduces_the_reference_coordinator(name):
    lm, kw, z = G.load_linked(name)
    lm.validate()
    outs, delivered, lost, ends = run_oracle(lm, kw, z)
    assert len(ends) == int(z["total_windows"])
    assert int(delivered[0]) == int(z["cross_events"])
    assert sum(int(o["summaries"]["events_processed"][0]) for o in outs) == int(z["total_events"])
    for q in range(lm.n_partitions):
        G.check_linked_partition(z, q, outs[q])
    sent = sum(int(o["entity_stats"][0][i]["c0"]) for m, o in zip(lm.models, outs) for i in m.ids_of(A.HS_ENT_REMOTE))
    assert sent == int(delivered[0]) + int(lost[0])


def test_the_fixtures_exercise_time_travel_loss_and_cross_counter_ties():
    """What makes the coordinator's semantics visible: delivered events that land behind the receiver's clock are
    skipped (not processed), lossy links drop some, and in the aligned ring requests from two partitions' counters
    tie on the nanosecond."""
    lm, kw, z = G.load_linked("linked_tandem_lossy_exp")          # exponential link latency, mean = the window
    outs, delivered, lost, _ = run_oracle(lm, kw, z)
    b = outs[1]
    arrived = int(b["entity_stats"][0][0]["c0"]) + int(b["entity_stats"][0][0]["c1"])       # accepted + dropped at B.server
    left = int((b["summaries"]["heap_left"][0]))
    assert int(lost[0]) > 10 and arrived + 10 < int(delivered[0]) - left                   # many were skipped
    lm, kw, z = G.load_linked("linked_lossy_fanout")
    outs, delivered, lost, _ = run_oracle(lm, kw, z)
    assert int(lost[0]) > 20 and int(delivered[0]) > 200
    for name, exact in (("linked_aligned_ring", True), ("linked_aligned_ring_spread", False)):
        lm, kw, z = G.load_linked(name)
        rec = z["p1_records"]
        req = rec[rec["kind"] == A.HS_EV_REQ_ENQUEUE]
        t, c = np.unique(req["time_ns"], return_counts=True)
        assert (c > 1).sum() > 20                                                         # same-nanosecond requests at B.server
        # ties on time AND index between two different events (the payload a worker takes over keeps its index, and
        # the bootstrap tick and the first run-time event both carry index 0: neither is a pair of rivals)
        other = rec[(rec["kind"] != A.HS_EV_REQ_WORKER) & (rec["sort_index"] > 0)]
        _, cc = np.unique(np.stack([other["time_ns"], other["sort_index"].astype(np.int64)], axis=1), axis=0, return_counts=True)
        assert ((cc > 1).sum() > 0) == exact


def test_replicas_are_independent_and_keyed_by_their_global_index():
    lm, kw, z = G.load_linked("linked_lossy_fanout")
    outs3, d3, l3, _ = run_oracle(lm, kw, z, n_replicas=3)
    for q in range(lm.n_partitions):
        G.check_linked_partition(z, q, outs3[q], r=0)
    outs1, d1, l1, _ = run_oracle(lm, kw, z, n_replicas=1, replica_index_base=2)
    for q in range(lm.n_partitions):
        assert outs1[q]["summaries"][0].tobytes() == outs3[q]["summaries"][2].tobytes()
        assert outs1[q]["entity_stats"][0].tobytes() == outs3[q]["entity_stats"][2].tobytes()
    assert (int(d1[0]), int(l1[0])) == (int(d3[2]), int(l3[2])) != (int(d3[0]), int(l3[0]))


def test_window_ends_follow_the_coordinators_float_arithmetic():
    lm = LinkedModel([], [], [], window_s=0.05)
    ends = lm.window_ends(2 * 10**9)
    assert ends[:3] == [50_000_000, 100_000_000, 150_000_000] and ends[-1] == 2 * 10**9
    assert len(ends) == 41          # 0.05 accumulates in float seconds: the 40th window ends 1 ns short of 2 s
    with pytest.raises(ValueError, match="does not advance"):
        LinkedModel([], [], [], window_s=1e-12).window_ends(10)


def test_validation_of_remote_rows():
    lm, kw, z = G.load_linked("linked_tandem_const")
    lm.models[0].entities = lm.models[0].entities.copy()
    lm.models[0].entities["i1"][2] = 7                       # no such entity in B
    with pytest.raises(ValueError, match="cannot receive"):
        lm.validate()
    lm, kw, z = G.load_linked("linked_tandem_const")
    lm.models[1].inbox_cap = 0
    with pytest.raises(ValueError, match="no inbox"):
        lm.validate()
    lm, kw, z = G.load_linked("linked_tandem_const")
    lm.links[0][0] = LinkSpec(0, 0, 0.05)
    with pytest.raises(ValueError, match="destination"):
        lm.validate()
