"""Generates tests/golden/*.json — run in the build container:  python tests/golden/make_golden.py

The reference is Java and no JVM exists here, so the vectors are produced by the pure-Python
restatement (oracle/py_oracle.py) and cross-checked against the C++ restatement before being written.
Inputs of the first four cases are the reference's own JUnit inputs
(src/test/java/siftscience/kafka/tools/KafkaTopicAssignerTest.java:21-35, 62-68, 87-93, 127-133); the
expected outputs agree with the hand traces of SURVEY.md §8c and satisfy every assertion of those tests.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import py_oracle as po  # noqa: E402


def run_case(topics, brokers, racks, desired_rf=-1):
    """topics: [(name, {partition: [brokers]})] through ONE assigner."""
    try:
        recs = po.run_topics([(n, {int(k): v for k, v in c.items()}) for n, c in topics], brokers, racks, desired_rf)
        return {"records": [[n, p, r] for n, p, r in recs]}
    except po.JavaError as e:
        return {"error": {"kind": e.kind, "message": e.message, "partition": e.partition, "a": e.a, "b": e.b}}


def main():
    cases = []
    cur_a = {0: [10, 11], 1: [11, 12], 2: [12, 10], 3: [10, 12]}
    cases.append(dict(name="ref_testRackAwareExpansion", topics=[["test", cur_a]], brokers=[10, 11, 12, 13, 14],
                      racks={10: "a", 11: "b", 12: "c", 13: "a", 14: "b"}, desired_rf=-1))
    cases.append(dict(name="ref_testClusterExpansion", topics=[["test", cur_a]], brokers=[10, 11, 12, 13], racks={}, desired_rf=-1))
    cases.append(dict(name="ref_testDecommission", topics=[["test", {0: [10, 11], 1: [11, 12], 2: [12, 13], 3: [13, 10]}]],
                      brokers=[10, 11, 13], racks={}, desired_rf=-1))
    cases.append(dict(name="ref_testReplacement", topics=[["test", cur_a]], brokers=[10, 11, 13], racks={}, desired_rf=-1))
    # shared-Context order dependence
    A = {0: [3, 1], 1: [4, 3], 2: [1, 4]}
    B = {0: [1, 2], 1: [1, 3], 2: [2, 1]}
    cases.append(dict(name="two_topics_ab", topics=[["a", A], ["b", B]], brokers=[1, 2, 3, 4], racks={}, desired_rf=-1))
    cases.append(dict(name="two_topics_ba", topics=[["b", B], ["a", A]], brokers=[1, 2, 3, 4], racks={}, desired_rf=-1))
    # error paths
    cases.append(dict(name="err_rf_mismatch", topics=[["t", {0: [1, 2], 1: [1]}]], brokers=[1, 2, 3], racks={}, desired_rf=-1))
    cases.append(dict(name="err_rf_not_positive", topics=[["t", {}]], brokers=[1, 2, 3], racks={}, desired_rf=-1))
    cases.append(dict(name="err_rf_gt_brokers", topics=[["t", {0: [1, 2, 3]}]], brokers=[1, 2], racks={}, desired_rf=-1))
    cases.append(dict(name="err_unassignable_racks", topics=[["t", {0: [1, 2], 1: [2, 1]}]], brokers=[1, 2, 3],
                      racks={1: "x", 2: "x", 3: "y"}, desired_rf=3))
    cases.append(dict(name="err_hash_min_value", topics=[["polygenelubricants", {0: [1, 2, 3]}]], brokers=[1, 2, 3], racks={}, desired_rf=-1))
    cases.append(dict(name="second_topic_fails", topics=[["ok", {0: [1, 2]}], ["bad", {0: [1, 2, 3, 4]}]], brokers=[1, 2, 3], racks={}, desired_rf=-1))
    # replication-factor changes, rack-less brokers, string-collision quirk (rack named like a broker id)
    cases.append(dict(name="rf_increase", topics=[["grow", {0: [1], 1: [2], 2: [3], 3: [1]}]], brokers=[1, 2, 3, 4],
                      racks={1: "a", 2: "b", 3: "a", 4: "b"}, desired_rf=2))
    cases.append(dict(name="rf_decrease_keeps_extra", topics=[["shrink", {0: [1, 2, 3], 1: [2, 3, 4], 2: [3, 4, 1]}]], brokers=[1, 2, 3, 4],
                      racks={}, desired_rf=2))
    cases.append(dict(name="rack_name_collides_with_id", topics=[["q", {0: [13, 14], 1: [14, 15], 2: [15, 13]}]], brokers=[13, 14, 15, 16],
                      racks={14: "13", 16: "z"}, desired_rf=-1))
    # seeded random multi-topic cases
    rng = random.Random(20260922)
    for ci in range(12):
        nb = rng.randint(3, 12)
        brokers = rng.sample(range(1, 40), nb)
        nr = rng.randint(2, nb)
        racks = {b: "rack%d" % rng.randrange(nr) for b in brokers if rng.random() < 0.8}
        old = brokers + rng.sample(range(40, 60), rng.randint(0, 3))
        topics = []
        for ti in range(rng.randint(1, 5)):
            rf = rng.randint(1, min(3, nb))
            topics.append(["t%d_%d" % (ci, ti), {p: rng.sample(old, rf) for p in range(rng.randint(1, 9))}])
        cases.append(dict(name="random_%02d" % ci, topics=topics, brokers=sorted(brokers), racks=racks, desired_rf=-1))
    out = []
    for c in cases:
        res = run_case([(n, cur) for n, cur in c["topics"]], c["brokers"], c["racks"], c["desired_rf"])
        c = dict(c)
        c["topics"] = [[n, {str(k): v for k, v in cur.items()}] for n, cur in c["topics"]]
        c["racks"] = {str(k): v for k, v in c["racks"].items()}
        c["expected"] = res
        out.append(c)
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote %d cases" % len(out))
    for c in out:
        print(c["name"], "->", "ERROR " + c["expected"]["error"]["message"] if "error" in c["expected"] else "%d records" % len(c["expected"]["records"]))


if __name__ == "__main__":
    main()
