"""CPU tests (-m "not gpu"): pin the oracle. The two restatements (C++ structure-faithful, pure Python)
must agree with each other, with the committed golden vectors, and with every assertion of the
reference's own JUnit tests (KafkaTopicAssignerTest.java:18-187)."""
import random

import numpy as np
import pytest

from oracle import py_oracle as po
from tests import util


def _counts(new):
    c = {}
    for reps in new.values():
        assert len(reps) == len(set(reps))  # TEST:168 no broker twice in a partition
        for b in reps:
            c[b] = c.get(b, 0) + 1
    return c


def _sticky(cur, new, k=1):
    for p, reps in new.items():  # TEST:181-184 minimal movement
        assert len(set(reps) & set(cur[p])) >= k


CUR_A = {0: [10, 11], 1: [11, 12], 2: [12, 10], 3: [10, 12]}


def _both(oracle, topic, cur, brokers, racks, desired=-1):
    py = po.KafkaTopicAssigner().generate_assignment(topic, cur, brokers, racks, desired)
    case = dict(topics=[(topic, cur)], brokers=brokers, racks=racks, desired_rf=desired)
    cpp = util.run_oracle_case(oracle, case)
    assert cpp["records"] == [[topic, p, py[p]] for p in sorted(py)]
    return py


def test_ref_rack_aware_expansion(oracle):  # TEST:18-57
    new = _both(oracle, "test", CUR_A, [10, 11, 12, 13, 14], {10: "a", 11: "b", 12: "c", 13: "a", 14: "b"})
    _sticky(CUR_A, new)
    c = _counts(new)
    assert sorted(c.values()).count(1) == 2 and sorted(c.values()).count(2) == 3
    assert new == {0: [10, 11], 1: [11, 12], 2: [12, 13], 3: [14, 10]}  # SURVEY §8c hand trace


def test_ref_cluster_expansion(oracle):  # TEST:59-82
    new = _both(oracle, "test", CUR_A, [10, 11, 12, 13], {})
    _sticky(CUR_A, new)
    assert all(v == 2 for v in _counts(new).values())
    assert new == {0: [10, 11], 1: [11, 12], 2: [12, 13], 3: [13, 10]}


def test_ref_decommission(oracle):  # TEST:84-122
    cur = {0: [10, 11], 1: [11, 12], 2: [12, 13], 3: [13, 10]}
    new = _both(oracle, "test", cur, [10, 11, 13], {})
    _sticky(cur, new)
    c = _counts(new)
    assert 12 not in c
    assert sorted(c.values()) == [2, 3, 3]
    assert new == {0: [10, 11], 1: [11, 13], 2: [13, 10], 3: [10, 13]}


def test_ref_replacement(oracle):  # TEST:124-157 — holds the reference's only exact pin (TEST:143-144)
    new = _both(oracle, "test", CUR_A, [10, 11, 13], {})
    _sticky(CUR_A, new)
    assert 12 not in _counts(new)
    assert new[0] == CUR_A[0] == [10, 11]
    assert 11 in new[1] and (10 in new[1] or 13 in new[1])
    assert 10 in new[2] and (11 in new[2] or 13 in new[2])
    assert 10 in new[3] and (11 in new[3] or 13 in new[3])


def test_java_string_hash(oracle):
    for s, h in [("test", 3556498), ("", 0), ("a", 97), ("polygenelubricants", -2**31), ("topic-000000", None),
                 ("héllo-日本", None), ("\U0001F600x", None)]:
        assert oracle.java_string_hash(s) == po.java_string_hash(s)
        if h is not None:
            assert po.java_string_hash(s) == h


def test_golden_cases_both_oracles(oracle):
    cases = util.load_golden()
    assert len(cases) >= 20
    for c in cases:
        got = util.run_oracle_case(oracle, c)
        exp = c["expected"]
        if "error" in exp:
            assert "error" in got, c["name"]
            assert got["error"] == exp["error"], c["name"]
        else:
            assert got.get("records") == exp["records"], c["name"]


def test_random_cross_check_py_vs_cpp(oracle):
    rng = random.Random(7)
    n_ok = 0
    for _ in range(150):
        nb = rng.randint(2, 14)
        brokers = sorted(rng.sample(range(1, 60), nb))
        racks = {b: "k%d" % rng.randrange(max(2, nb // 2)) for b in brokers if rng.random() < 0.7}
        universe = brokers + [100, 101, 102]
        topics = []
        for ti in range(rng.randint(1, 4)):
            rf = rng.randint(1, min(4, nb))
            ragged = rng.random() < 0.2
            cur = {}
            for p in sorted(rng.sample(range(0, 30), rng.randint(1, 12))):  # entry order = ascending (flat ABI order)
                k = rng.randint(1, 4) if ragged else rf
                cur[p] = rng.sample(universe, k)
            topics.append(("rt%d" % rng.randrange(1000), cur))
        desired = rng.choice([-1, -1, -1, 1, 2, 3])
        case = dict(topics=topics, brokers=brokers, racks=racks, desired_rf=desired)
        cpp = util.run_oracle_case(oracle, case)
        try:
            recs = po.run_topics(topics, brokers, racks, desired)
            assert cpp.get("records") == [[n, p, r] for n, p, r in recs]
            n_ok += 1
        except po.JavaError as e:
            assert cpp["error"]["message"] == e.message and cpp["error"]["kind"] == e.kind
    assert n_ok > 20


def test_synth_generator_is_deterministic_and_feasible(oracle):
    import kafka_assigner_b200 as kab
    a = kab.synth.make_config("c1", "mixed")
    b = kab.synth.make_config("c1", "mixed")
    assert np.array_equal(a.cur, b.cur) and a.cur.shape == (10, 8, 3)
    out, ln, st = util.oracle_dense(oracle, a)
    assert st.code == 0
    # every partition: RF distinct brokers on distinct racks, all live
    rack = dict(zip(a.broker_id.tolist(), a.rack_index.tolist()))
    for row in out:
        assert len(set(row.tolist())) == 3 and len({rack[int(x)] for x in row}) == 3
    c2 = kab.synth.make_config("c2", "mixed")
    out2, ln2, st2 = util.oracle_dense(oracle, c2)
    assert st2.code == 0 and (ln2 == 3).all()
    cap = -(-c2.P * c2.RF // c2.N)
    per_topic = out2.reshape(c2.T, -1)
    for t in range(0, c2.T, 97):
        assert np.bincount(per_topic[t] - 1000, minlength=c2.N).max() <= cap  # KAS:65-71 capacity bound


def test_fast_cpu_solver_agrees_with_structure_faithful_oracle(oracle):
    """Third restatement (oracle/fast_oracle.cpp, flat arrays) == kafka_oracle.cpp, successes and failures."""
    import kafka_assigner_b200 as kab
    shapes = [dict(T=10, P=8, RF=3, N=6, R=3, n_old=6), dict(T=40, P=33, RF=3, N=64, R=8), dict(T=16, P=100, RF=3, N=30, R=6),
              dict(T=64, P=17, RF=1, N=11, R=11), dict(T=12, P=96, RF=5, N=35, R=7), dict(T=30, P=48, RF=3, N=60, R=6, remove_frac=0.2, n_old=60),
              dict(T=30, P=44, RF=3, N=60, R=6, remove_frac=0.2, n_old=60, rack_aware=False)]
    n_ok = n_err = 0
    for sh in shapes:
        for kind in ("structured", "random", "mixed"):
            cl = kab.synth.make_cluster(seed=0xF00 + sh["T"], kind=kind, **sh)
            exp, exp_len, est = util.oracle_dense(oracle, cl)
            got, got_len, st = oracle.fast_run_dense(oracle.FastContext(), cl.topic_hash, cl.cur, cl.broker_id, cl.rack_index)
            assert (st.code, st.topic_index) == (est.code, est.topic_index), (sh, kind)
            if est.code == 0:
                assert np.array_equal(got, exp) and np.array_equal(got_len, exp_len), (sh, kind)
                n_ok += 1
            else:
                assert st.partition == est.partition
                n_err += 1
    assert n_ok >= 12
    c2 = kab.synth.make_config("c2", "mixed")
    exp, _, _ = util.oracle_dense(oracle, c2)
    got, _, st = oracle.fast_run_dense(oracle.FastContext(), c2.topic_hash, c2.cur, c2.broker_id, c2.rack_index)
    assert st.code == 0 and np.array_equal(got, exp)


def test_fast_solver_pinned_on_baseline_shapes(oracle):
    """The flat-array solver is what the -m gpu tests and bench.py compare EVERY row of the big configs against; pin it to the
    structure-faithful oracle on topic prefixes of exactly those shapes (c3, c4, c5 at three decommission fractions)."""
    import kafka_assigner_b200 as kab
    cases = [("c3", dict(T=60), k) for k in ("structured", "random", "mixed")] + [("c4shard", dict(T=16), "mixed")] + \
            [("c5", dict(T=4, remove_frac=f), "mixed") for f in (0.01, 0.2, 0.5)]
    for key, over, kind in cases:
        cl = kab.synth.make_config(key, kind, **over)
        exp, exp_len, est = util.oracle_dense(oracle, cl)
        got, got_len, st = oracle.fast_run_dense(oracle.FastContext(), cl.topic_hash, cl.cur, cl.broker_id, cl.rack_index)
        assert st.code == est.code == 0, (key, kind)
        assert np.array_equal(got, exp) and np.array_equal(got_len, exp_len), (key, kind)
