"""Dev tool: per-phase device times of one workload (CUDA events inside the library), for tuning knobs
passed through the environment (e.g. KA_ORDER_THREADS). Verifies against the oracle on a topic prefix."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import kafka_assigner_b200 as kab  # noqa: E402
from oracle import oracle_lib as ol  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--kind", default="mixed")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--verify-topics", type=int, default=300)
a = ap.parse_args()
cl = kab.synth.make_config(a.workload, a.kind)
s = kab.Solver(0)
s.set_brokers(cl.broker_id, cl.rack_index)
s.set_timing(True)
d_hash = torch.from_numpy(cl.topic_hash).cuda()
d_cur = torch.from_numpy(cl.cur).cuda()
d_out = torch.empty((cl.T, cl.P, cl.RF), dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rows = []
for i in range(a.steps + 2):
    s.reset()
    flush.fill_(i)
    torch.cuda.synchronize()
    st = s.solve_dense_device(cl.T, d_hash.data_ptr(), cl.P, cl.RF, d_cur.data_ptr(), -1, cl.RF, 0, d_out.data_ptr(),
                              stream=torch.cuda.current_stream().cuda_stream)
    assert st.code == 0, st.code
    if i >= 2:
        rows.append(s.last_timing())
avg = {k: float(np.mean([r[k] for r in rows])) for k in rows[0]}
n = min(cl.T, a.verify_topics)
sub = cl.subset(0, n)
po, pid, ro, cur = sub.ragged()
_, _, exp, _ = ol.run(ol.OracleContext(), sub.topic_names, po, pid, ro, cur, sub.broker_id, sub.rack_name, -1, cl.RF)
ok = np.array_equal(d_out.cpu().numpy()[:n].reshape(-1, cl.RF), exp)
knobs = " ".join("%s=%s" % (k[9:], v) for k, v in sorted(os.environ.items()) if k.startswith("KA_ORDER_") or k.startswith("KA_CHAIN") or k.startswith("KA_PIPE"))
print("%s [%s] KA_ORDER_THREADS=%s A=%.3fms T=%.3fms B1=%.3fms B2+emit=%.3fms chains_wall=%.3fms total=%.3fms  rate=%.3g/s  verified(first %d topics)=%s" % (
    a.workload, knobs, os.environ.get("KA_ORDER_THREADS", "-"), avg["sticky_spread_ms"], avg["level_tables_ms"], avg["leader_order_ms"],
    avg["slot1_emit_ms"], avg["chains_wall_ms"],
    avg["total_ms"], cl.replicas / (avg["total_ms"] * 1e-3), n, ok))
