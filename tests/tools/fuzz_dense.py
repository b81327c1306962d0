"""Randomised parity fuzz of the DENSE path (ka_solve_dense, incl. the pipelined super-chunk mode) against the CPU
restatements: the flat-array solver for every case, the structure-faithful oracle for the small ones.
   KA_PIPELINE_STAGES=3 python tests/tools/fuzz_dense.py --cases 300 --seed 1"""
import argparse
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import kafka_assigner_b200 as kab  # noqa: E402
from oracle import oracle_lib as ol  # noqa: E402
from tests import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cases", type=int, default=200)
ap.add_argument("--seed", type=int, default=1)
a = ap.parse_args()
rng = random.Random(a.seed)
t0 = time.time()
n_ok = n_err = 0
for it in range(a.cases):
    R = rng.choice([3, 4, 5, 8, 10, 20, 50])
    RF = rng.choice([1, 2, 3, 3, 3])
    R = max(R, RF)
    N = R * rng.choice([1, 2, 3, 5, 10, 40])
    P = rng.choice([1, 3, 8, 31, 32, 33, 64, 100, 128, 257])
    T = rng.choice([1, 2, 5, 17, 64, 200])
    if T * P * RF > 400000:
        T = max(1, 400000 // (P * RF))
    kind = rng.choice(["structured", "random", "mixed"])
    remove = rng.choice([0.0, 0.0, 0.1, 0.3])
    cl = kab.synth.make_cluster(T=T, P=P, RF=RF, N=N, R=R, seed=rng.randrange(1 << 30), kind=kind, remove_frac=remove,
                                rack_aware=rng.random() < 0.8, n_old=rng.choice([None, N]), t_offset=rng.choice([0, 1000]))
    s = kab.Solver(0)
    out, out_len, st = s.solve_cluster(cl, check=False)
    fo, fl, fst = ol.fast_run_dense(ol.FastContext(), cl.topic_hash, cl.cur, cl.broker_id, cl.rack_index)
    ok = (st.code, st.topic_index, st.partition) == (fst.code, fst.topic_index, fst.partition)
    if ok and fst.code == 0:
        ok = np.array_equal(out.reshape(-1, RF), fo) and np.array_equal(out_len.reshape(-1), fl)
    if ok and cl.replicas <= 30000:
        eo, el, est = util.oracle_dense(ol, cl)
        ok = (st.code, st.topic_index, st.partition) == (est.code, est.topic_index, est.partition)
        if ok and est.code == 0:
            ok = np.array_equal(out.reshape(-1, RF), eo)
    if not ok:
        print("MISMATCH case", it, "seed", a.seed, cl.meta, "gpu", (st.code, st.topic_index, st.partition), "cpu", (fst.code, fst.topic_index, fst.partition))
        sys.exit(1)
    n_ok += fst.code == 0
    n_err += fst.code != 0
    s.close()
print("dense fuzz ok: %d cases (%d solved, %d reference exceptions), KA_PIPELINE_STAGES=%s, %.1f s" % (
    n_ok + n_err, n_ok, n_err, os.environ.get("KA_PIPELINE_STAGES", "default"), time.time() - t0))
