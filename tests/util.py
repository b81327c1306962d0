"""Shared helpers for the parity tests: flatten {topic: {partition: [brokers]}} cases into the flat
layout of include/kassign.h, and run them through the C++ oracle or the CUDA library."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_golden():
    with open(os.path.join(HERE, "golden", "cases.json")) as f:
        cases = json.load(f)
    for c in cases:
        c["topics"] = [(n, {int(k): v for k, v in cur.items()}) for n, cur in c["topics"]]
        c["racks"] = {int(k): v for k, v in c["racks"].items()}
    return cases


def flatten(topics):
    """topics: [(name, {partition: [brokers]})] -> names, part_off, part_id, rep_off, cur (ascending partitions)."""
    names, part_off, part_id, rep_off, cur = [], [0], [], [0], []
    for name, asg in topics:
        names.append(name)
        for p in sorted(asg):
            part_id.append(p)
            cur.extend(asg[p])
            rep_off.append(len(cur))
        part_off.append(len(part_id))
    return (names, np.array(part_off, dtype=np.int64), np.array(part_id, dtype=np.int32),
            np.array(rep_off, dtype=np.int64), np.array(cur, dtype=np.int32))


def stride_for(topics, desired_rf):
    m = max([len(v) for _, a in topics for v in a.values()], default=0)
    return max(1, m, desired_rf if desired_rf >= 0 else 0)


def records_from_flat(names, part_off, part_id, out, out_len):
    recs = []
    for t, n in enumerate(names):
        for g in range(int(part_off[t]), int(part_off[t + 1])):
            recs.append([n, int(part_id[g]), [int(x) for x in out[g, :out_len[g]]]])
    return recs


def run_oracle_case(ol, case):
    names, part_off, part_id, rep_off, cur = flatten(case["topics"])
    brokers = sorted(case["brokers"])
    racks = [case["racks"].get(b) for b in brokers]
    stride = stride_for(case["topics"], case["desired_rf"])
    ln, pid, out, st = ol.run(ol.OracleContext(), names, part_off, part_id, rep_off, cur, brokers, racks,
                              case["desired_rf"], stride, raise_on_error=False)
    if st.code != 0:
        return {"error": {"kind": st.code, "message": st.message.decode(), "partition": st.partition, "a": st.a, "b": st.b},
                "topic_index": st.topic_index}
    return {"records": records_from_flat(names, part_off, pid, out, ln)}


def run_gpu_case(kab, case, solver=None):
    """Through the C ABI (ka_solve, ragged form). Returns records or the re-thrown reference exception."""
    names, part_off, part_id, rep_off, cur = flatten(case["topics"])
    s = solver or kab.Solver(0)
    s.set_brokers_with_racks(case["brokers"], case["racks"])
    stride = stride_for(case["topics"], case["desired_rf"])
    th = np.array([kab.java_string_hash(n) for n in names], dtype=np.int32)
    out, out_len, st = s.solve_ragged(th, part_off, part_id, rep_off, cur, case["desired_rf"], stride, check=False)
    if st.code != 0:
        try:
            kab.raise_for_status(st, names)
        except (kab.IllegalStateException, kab.ArrayIndexOutOfBoundsException) as e:
            return {"error": {"kind": st.code, "message": str(e), "partition": st.partition, "a": st.a, "b": st.b},
                    "topic_index": st.topic_index}
    return {"records": records_from_flat(names, part_off, part_id, out, out_len)}


def oracle_dense(ol, cl, ctx=None):
    """Run a synth.Cluster through the C++ oracle; returns (out [T*P, RF], out_len, status)."""
    part_off, part_id, rep_off, cur = cl.ragged()
    ln, _, out, st = ol.run(ctx or ol.OracleContext(), cl.topic_names, part_off, part_id, rep_off, cur, cl.broker_id,
                            cl.rack_name, cl.desired_rf, max(cl.RF, cl.desired_rf, 1), raise_on_error=False)
    return out, ln, st
