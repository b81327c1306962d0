"""Randomised GPU-vs-oracle parity fuzz through the C ABI (ragged ka_solve). Every case must match bit-for-bit: records,
or the same exception (kind, topic index, partition, operands).   python tests/tools/fuzz_parity.py --cases 4000 --seed 1"""
import argparse
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import kafka_assigner_b200 as kab  # noqa: E402
from oracle import oracle_lib as ol  # noqa: E402
from tests import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cases", type=int, default=2000)
ap.add_argument("--seed", type=int, default=1)
a = ap.parse_args()
rng = random.Random(a.seed)
solver = kab.Solver(0)
t0 = time.time()
n_ok = n_err = 0
kinds = {}
for it in range(a.cases):
    solver.reset()
    nb = rng.choice([1, 2, 3, 5, 8, 13, 30, 64, 100, 257, 600])
    spread = rng.choice([1, 1, 3, 1000, 10**6])
    brokers = sorted(rng.sample(range(-50 * spread, 50 * spread + nb * spread + 1000), nb))
    n_rack = rng.choice([2, 3, 5, max(1, nb // 2), nb, nb])
    racks = {b: "k%d" % rng.randrange(n_rack) for b in brokers if rng.random() < rng.choice([0.0, 0.6, 1.0])}
    universe = brokers + [brokers[-1] + 1, brokers[0] - 1, 7_000_000]
    desired = -1 if rng.random() < 0.8 else rng.choice([0, 1, 2, 3, 4, 6])
    topics = []
    max_rf = rng.choice([1, 2, 3, 3, 3, 4, 6, 8])
    if rng.random() < 0.7:  # mostly feasible shapes: RF no larger than the number of racks / brokers
        max_rf = max(1, min(max_rf, n_rack if racks else nb, nb))
    for ti in range(rng.randint(1, 5)):
        rf = rng.randint(1, max(1, min(max_rf, len(universe))))
        ragged = rng.random() < (0.4 if desired >= 0 else 0.03)
        cur = {}
        npart = rng.choice([1, 2, 7, 31, 32, 33, 64, 90, 200]) if rng.random() < 0.9 else rng.randint(0, 400)
        for p in sorted(rng.sample(range(0, 1000), npart)):
            k = rng.randint(0, min(8, len(universe))) if ragged else rf
            cur[p] = rng.sample(universe, k)
        name = "polygenelubricants" if rng.random() < 0.03 else rng.choice(["t%d_%d" % (it, ti), "x" * rng.randint(1, 40), "日本-%d" % ti])
        topics.append((name, cur))
    case = dict(topics=topics, brokers=brokers, racks=racks, desired_rf=desired)
    if util.stride_for(topics, desired) > 8:
        continue
    exp = util.run_oracle_case(ol, case)
    got = util.run_gpu_case(kab, case, solver)
    if got != exp:
        print("MISMATCH at case", it, "seed", a.seed)
        print(" exp:", str(exp)[:600])
        print(" got:", str(got)[:600])
        print(" case:", str(case)[:1500])
        sys.exit(1)
    if "records" in exp:
        n_ok += 1
    else:
        n_err += 1
        kinds[exp["error"]["kind"]] = kinds.get(exp["error"]["kind"], 0) + 1
print("fuzz ok: %d cases (%d solved, %d reference exceptions %s) in %.1f s" % (n_ok + n_err, n_ok, n_err, kinds, time.time() - t0))
