"""CPU model of kernel B's dataflow schedule (analysis tool, no GPU): given a solved run, simulate W warps taking
32-partition windows in order (K windows per warp in flight), each warp iterating with period `iter_cycles` scaled by the
number of warps sharing its scheduler; a lane commits at the end of an iteration if all its predecessors had committed
when the iteration started; a warp pays `switch_cycles` when it moves to its next window(s). Prints predicted cycles for a
few (threads, K) choices so the measured kernel times can be explained / the next tuning step chosen."""
import argparse
import heapq
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import kafka_assigner_b200 as kab  # noqa: E402
from oracle import oracle_lib as ol  # noqa: E402


def predecessors(broker_id, out):
    """pred[q] = list of the previous partition on each of q's brokers (or -1)."""
    idx = np.searchsorted(broker_id, out)
    last = np.full(len(broker_id), -1, dtype=np.int64)
    pred = np.full(idx.shape, -1, dtype=np.int64)
    for q in range(idx.shape[0]):
        for i, b in enumerate(idx[q]):
            pred[q, i] = last[b]
            last[b] = q
    return pred


def simulate(pred, warps, K, iter_cycles, switch_cycles, contention):
    Q = pred.shape[0]
    nwin = (Q + 31) // 32
    commit_time = np.full(Q, np.inf)
    per_sched = max(1.0, warps / 4.0)
    period = iter_cycles * (1.0 + contention * (per_sched - 1.0)) * (1.0 + 0.6 * (K - 1))
    # event queue of (time, warp); each warp holds K consecutive windows of its own sequence
    next_win = [w for w in range(warps)]
    held = [[] for _ in range(warps)]
    heap = []
    for w in range(warps):
        for _ in range(K):
            if next_win[w] < nwin:
                held[w].append(next_win[w])
                next_win[w] += warps
        heapq.heappush(heap, (switch_cycles, w))
    pending = {}
    for w in range(warps):
        for win in held[w]:
            pending[win] = list(range(win * 32, min(Q, win * 32 + 32)))
    t_end = 0.0
    while heap:
        t, w = heapq.heappop(heap)
        if not held[w]:
            continue
        start = t
        done_any_window = False
        for win in list(held[w]):
            still = []
            for q in pending[win]:
                ok = True
                for p in pred[q]:
                    if p >= 0 and not (commit_time[p] <= start):
                        ok = False
                        break
                if ok:
                    commit_time[q] = start + period
                else:
                    still.append(q)
            pending[win] = still
            if not still:
                held[w].remove(win)
                del pending[win]
                done_any_window = True
                if next_win[w] < nwin:
                    nw = next_win[w]
                    next_win[w] += warps
                    held[w].append(nw)
                    pending[nw] = list(range(nw * 32, min(Q, nw * 32 + 32)))
        t_next = start + period + (switch_cycles if done_any_window else 0.0)
        t_end = max(t_end, start + period)
        if held[w]:
            heapq.heappush(heap, (t_next, w))
    return t_end


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--topics", type=int, default=0, help="topic prefix (0 = all)")
    a = ap.parse_args()
    cl = kab.synth.make_config(a.workload, "mixed")
    if a.topics:
        cl = cl.subset(0, a.topics)
    out, _, st = ol.fast_run_dense(ol.FastContext(), cl.topic_hash, cl.cur, cl.broker_id, cl.rack_index)
    assert st.code == 0
    pred = predecessors(cl.broker_id, out)
    print("%s: %d partitions" % (cl.name, pred.shape[0]))
    for warps, K in [(4, 1), (8, 1), (16, 1), (32, 1), (8, 2), (16, 2), (8, 4)]:
        cyc = simulate(pred, warps, K, iter_cycles=239.0, switch_cycles=250.0, contention=0.35)
        print("  threads=%4d K=%d  -> %.2f M cycles = %.3f ms at 1.965 GHz" % (warps * 32, K, cyc / 1e6, cyc / 1.965e6))


def simulate_refill(pred, warps, iter_cycles, contention, refill_overhead=0.10):
    """Variant: a lane that committed takes the next unassigned partition immediately (no per-window barrier)."""
    Q = pred.shape[0]
    commit_time = np.full(Q, np.inf)
    per_sched = max(1.0, warps / 4.0)
    period = iter_cycles * (1.0 + contention * (per_sched - 1.0)) * (1.0 + refill_overhead)
    nxt = 0
    lanes = [[] for _ in range(warps)]
    heap = []
    for w in range(warps):
        take = list(range(nxt, min(Q, nxt + 32)))
        nxt += len(take)
        lanes[w] = take
        heapq.heappush(heap, (0.0, w))
    t_end = 0.0
    while heap:
        t, w = heapq.heappop(heap)
        if not lanes[w]:
            continue
        still = []
        for q in lanes[w]:
            ok = True
            for p in pred[q]:
                if p >= 0 and not (commit_time[p] <= t):
                    ok = False
                    break
            if ok:
                commit_time[q] = t + period
            else:
                still.append(q)
        free = 32 - len(still)
        take = list(range(nxt, min(Q, nxt + free)))
        nxt += len(take)
        lanes[w] = still + take
        t_end = max(t_end, t + period)
        if lanes[w]:
            heapq.heappush(heap, (t + period, w))
    return t_end
