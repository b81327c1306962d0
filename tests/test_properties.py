"""Property tests (hypothesis, CPU): invariants every successful run of the reference algorithm satisfies, checked on the
structure-faithful oracle, and three-way agreement of the restatements (C++ TreeMap-style, pure Python, flat-array C++).
These are the generalisation of the reference's helper verifyPartitionsAndBuildReplicaCounts
(KafkaTopicAssignerTest.java:159-187) plus the rack-exclusivity and capacity rules of
KafkaAssignmentStrategy.java:65-71, 320-324, 346-348."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import oracle_lib as ol
from oracle import py_oracle as po
from tests import util


@st.composite
def clusters(draw):
    n_brokers = draw(st.integers(2, 12))
    brokers = sorted(draw(st.lists(st.integers(1, 60), min_size=n_brokers, max_size=n_brokers, unique=True)))
    n_racks = draw(st.integers(1, n_brokers))
    racks = {}
    for b in brokers:
        if draw(st.booleans()) or draw(st.booleans()):
            racks[b] = "r%d" % draw(st.integers(0, n_racks - 1))
    rf = draw(st.integers(1, min(3, n_brokers)))
    universe = brokers + [97, 98]
    topics = []
    for ti in range(draw(st.integers(1, 3))):
        n_parts = draw(st.integers(1, 10))
        cur = {}
        for p in range(n_parts):
            cur[p] = draw(st.lists(st.sampled_from(universe), min_size=rf, max_size=rf, unique=True))
        topics.append(("pt%d" % ti, cur))
    return dict(topics=topics, brokers=brokers, racks=racks, desired_rf=-1), rf


@settings(max_examples=200, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(clusters())
def test_invariants_and_three_way_agreement(case_rf):
    case, rf = case_rf
    ol.build()
    got = util.run_oracle_case(ol, case)
    # (1) pure-Python restatement agrees (records or identical exception)
    try:
        recs = po.run_topics(case["topics"], case["brokers"], case["racks"], case["desired_rf"])
        assert got.get("records") == [[n, p, r] for n, p, r in recs]
    except po.JavaError as e:
        assert got["error"]["message"] == e.message
        return
    rack_key = {b: case["racks"].get(b, str(b)) for b in case["brokers"]}   # KAS:81-86
    by_topic = {}
    for name, p, reps in got["records"]:
        by_topic.setdefault(name, {})[p] = reps
    for name, cur in case["topics"]:
        new = by_topic[name]
        assert sorted(new) == sorted(cur)                                   # every partition answered (TreeMap order)
        cap = -(-len(cur) * rf // len(case["brokers"]))                     # KAS:65-71
        load = {}
        for p, reps in new.items():
            assert len(reps) == rf and len(set(reps)) == rf                 # TEST:168
            assert all(b in rack_key for b in reps)                         # only live brokers
            assert len({rack_key[b] for b in reps}) == rf                   # one replica per rack (KAS:346-348)
            for b in reps:
                load[b] = load.get(b, 0) + 1
        assert max(load.values()) <= cap                                    # per-topic capacity
        # stickiness: a current replica on a live broker is only dropped for capacity or rack reasons, so a partition
        # whose current brokers are all live, on distinct racks, in a topic with room keeps them when cap is not binding
        if all(load.get(b, 0) < cap for b in rack_key):
            for p, reps in cur.items():
                if all(b in rack_key for b in reps) and len({rack_key[b] for b in reps}) == len(reps):
                    assert set(reps) <= set(new[p]) or max(load.values()) == cap
    # (2) flat-array CPU solver agrees when the case is dense (same P and RF for every topic handled one by one)
    fctx = ol.FastContext()
    ids = np.array(case["brokers"], dtype=np.int32)
    keys = {}
    ridx = np.array([keys.setdefault(rack_key[b], len(keys)) for b in case["brokers"]], dtype=np.int32)
    for name, cur in case["topics"]:
        arr = np.array([cur[p] for p in sorted(cur)], dtype=np.int32)[None, :, :]
        out, out_len, fst = ol.fast_run_dense(fctx, np.array([po.java_string_hash(name)], dtype=np.int32), arr, ids, ridx)
        assert fst.code == 0
        assert [[int(x) for x in row] for row in out] == [by_topic[name][p] for p in sorted(cur)]


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(clusters())
def test_replica_sets_are_a_fixed_point(case_rf):
    """Re-running the assigner on its own output moves nothing: every replica sticks (loads <= cap, one replica per
    rack already hold), so the replica SETS are unchanged — only leadership order may differ (fresh Context)."""
    case, rf = case_rf
    first = util.run_oracle_case(ol, case)
    if "error" in first:
        return
    by_topic = {}
    for name, p, reps in first["records"]:
        by_topic.setdefault(name, {})[p] = reps
    again = dict(case, topics=[(name, by_topic[name]) for name, _ in case["topics"]])
    second = util.run_oracle_case(ol, again)
    assert "records" in second
    assert [(n, p, sorted(r)) for n, p, r in second["records"]] == [(n, p, sorted(r)) for n, p, r in first["records"]]
