"""Small end-to-end cases for compute-sanitizer (memcheck): dense + ragged + error + 8-slot rows + pipelined."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kafka_assigner_b200 as kab  # noqa: E402

cl = kab.synth.make_cluster(T=12, P=19, RF=3, N=30, R=5, seed=5, kind="mixed")
out, out_len, st = kab.Solver(0).solve_cluster(cl)
assert st.code == 0
a = kab.KafkaTopicAssigner()
print(a.generate_assignment("test", {0: [10, 11], 1: [11, 12], 2: [12, 10], 3: [10, 12]}, {10, 11, 13}, {}, -1))
print(a.generate_assignment("wide", {p: [1 + (p + i) % 9 for i in range(6)] for p in range(7)}, set(range(1, 12)), {}, -1))
try:
    a.generate_assignment("t", {0: [1, 2], 1: [2, 1]}, {1, 2, 3}, {1: "x", 2: "x", 3: "y"}, 3)
except kab.IllegalStateException as e:
    print("expected:", e)
big = kab.synth.make_cluster(T=40, P=16, RF=3, N=2000, R=20, seed=4, kind="random")   # global-LUT-free, larger table
print(kab.Solver(0).solve_cluster(big)[2].code)
lv = kab.synth.make_cluster(T=30, P=90, RF=3, N=60, R=6, seed=9, kind="mixed")      # capacity 5: conflict levels, chunk table, window mode
print(kab.Solver(0).solve_cluster(lv)[2].code)
wide = kab.synth.make_cluster(T=40, P=160, RF=3, N=600, R=6, seed=10, kind="mixed")  # capacity 1, 160-wide topics: bounds-free chain loop
s = kab.Solver(0)
s.set_brokers(wide.broker_id, wide.rack_index)
text, st = s.solve_dense_json(wide.topic_names, wide.topic_hash, wide.cur)            # + the device-side JSON emitter
print(st.code, len(text))
print("sanitize cases done")
