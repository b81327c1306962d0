"""BASELINE config 5: decommission sweep — remove 1..50 % of 10 000 brokers (50 racks), 1M partitions RF=3
(1000 topics x 1000 partitions), one GPU. Prints a markdown table (device-resident solve, CUDA events inside the
library) with the oracle timed on a topic prefix beside it; verifies that prefix bit-exactly."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import kafka_assigner_b200 as kab  # noqa: E402
from oracle import oracle_lib as ol  # noqa: E402

print("| removed | live brokers | cap | A ms | chunk tables ms | slot-0 chain ms | total ms | assignments/s | oracle (1 core) assignments/s | verified |")
print("|---|---|---|---|---|---|---|---|---|---|")
for f in (0.01, 0.02, 0.05, 0.10, 0.20, 0.30, 0.40, 0.50):
    cl = kab.synth.make_config("c5", "mixed", remove_frac=f)
    s = kab.Solver(0)
    s.set_brokers(cl.broker_id, cl.rack_index)
    s.set_timing(True)
    d_hash = torch.from_numpy(cl.topic_hash).cuda()
    d_cur = torch.from_numpy(cl.cur).cuda()
    d_out = torch.empty((cl.T, cl.P, cl.RF), dtype=torch.int32, device="cuda")
    rows = []
    for i in range(4):
        s.reset()
        torch.cuda.synchronize()
        st = s.solve_dense_device(cl.T, d_hash.data_ptr(), cl.P, cl.RF, d_cur.data_ptr(), -1, cl.RF, 0, d_out.data_ptr(),
                                  stream=torch.cuda.current_stream().cuda_stream)
        assert st.code == 0, (f, st.code, st.topic_index)
        if i:
            rows.append(s.last_timing())
    avg = {k: float(np.mean([r[k] for r in rows])) for k in rows[0]}
    n = 24
    sub = cl.subset(0, n)
    po, pid, ro, cur = sub.ragged()
    t0 = time.perf_counter()
    _, _, exp, _ = ol.run(ol.OracleContext(), sub.topic_names, po, pid, ro, cur, sub.broker_id, sub.rack_name, -1, cl.RF)
    dt = time.perf_counter() - t0
    ok = np.array_equal(d_out.cpu().numpy()[:n].reshape(-1, cl.RF), exp)
    print("| %d%% | %d | %d | %.3f | %.3f | %.3f | %.3f | %.3g | %.3g | %s |" % (
        round(f * 100), cl.N, -(-cl.P * cl.RF // cl.N), avg["sticky_spread_ms"], avg["level_tables_ms"], avg["leader_order_ms"],
        avg["total_ms"], cl.replicas / (avg["total_ms"] * 1e-3), sub.replicas / dt, ok))
    del s
