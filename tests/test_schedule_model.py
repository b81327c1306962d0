"""CPU model of the round-2 leader-order design, checked against the structure-faithful oracle (no GPU needed).

The CUDA path replaces the reference's single pass over all partitions (KafkaAssignmentStrategy.java:217-237, one shared
`Context.counter`) by
  1. a per-topic CONFLICT-LEVEL schedule (partitions of one level share no broker; levels run in order, topics one after the other),
  2. records that list a row's brokers in the order the reference's rotated scan visits them (KAS:263-278, 188-200) plus the
     precomputed tie-breaks e_pq of the slot-1 scan, short rows padded with a dummy broker whose counters are "infinite",
  3. ONE CHAIN PER REPLICA SLOT: slot 0 only reads/bumps counter[.][0], slot 1 counter[.][1] (given the slot-0 winner), and
     counter[.][2] is a plain sum for rows of <= 3 replicas.
This file restates exactly that in Python — the same record layout and decision rules as kassign_stage.cuh / kassign_order.cuh —
and asserts that it reproduces the oracle's ordered lists AND its final Context, even when the partitions of a level are
processed in a scrambled order and the whole slot-0 chain runs before the slot-1 chain starts.
"""
import random

import numpy as np
import pytest

import kafka_assigner_b200 as kab
from tests import util

INF = 0x3FFFFFFF


def java_abs_hash(h):
    return int(np.int64(abs(int(h))) if h != -2**31 else 2**31)


def build_records(cl, sets):
    """What kernel A emits for every row: ([a0, a1, a2] in slot-0 scan order with the dummy N for missing slots, len, e01, e02, e12)."""
    N = cl.N
    idx_of = {int(b): i for i, b in enumerate(cl.broker_id)}
    recs = []
    for t in range(cl.T):
        habs = java_abs_hash(cl.topic_hash[t])
        s2, s3 = habs % 2, habs % 3
        for p in range(cl.P):
            ix = sorted(idx_of[int(b)] for b in sets[t][p])          # ascending index == ascending id (KAS:205-214)
            k = len(ix)
            a, e = [N, N, N], (0, 0, 0)
            if k == 1:
                a[0] = ix[0]
            elif k == 2:
                a[0], a[1] = ix[s2], ix[1 - s2]                       # |hash| % 2 == 1: the higher id is scanned first
            elif k == 3:
                i = [(3 - s3) % 3, (4 - s3) % 3, (5 - s3) % 3]        # list position at scan position 0, 1, 2
                a = [ix[i[0]], ix[i[1]], ix[i[2]]]
                e = tuple(s2 if i[x] < i[y] else 1 - s2 for x, y in ((0, 1), (0, 2), (1, 2)))
            recs.append((a, k, e))
    return recs


def levels_of_topic(rows):
    """Conflict level of every partition of one topic (kernel A's LEVELS pass)."""
    last, lv = {}, []
    for a, k, _ in rows:
        real = [b for b in a[:max(k, 0)]]
        l = 1 + max([last.get(b, 0) for b in real] or [0])
        for b in real:
            last[b] = l
        lv.append(l)
    return lv


def run_model(cl, sets, rng):
    N, P = cl.N, cl.P
    recs = build_records(cl, sets)
    # schedule: topic by topic, inside a topic by level; inside a level ANY order (scrambled here)
    order = []
    for t in range(cl.T):
        rows = recs[t * P:(t + 1) * P]
        lv = levels_of_topic(rows)
        for level in range(1, max(lv + [0]) + 1):
            members = [t * P + p for p in range(P) if lv[p] == level]
            used = [b for q in members for b in recs[q][0][:recs[q][1]]]
            assert len(used) == len(set(used)), "partitions of one level must not share a broker"
            rng.shuffle(members)
            order.extend(members)
    assert sorted(order) == list(range(cl.T * P))
    c0 = [0] * N + [INF]
    c1 = [0] * N + [INF]
    c2 = [0] * (N + 1)
    # ---- slot-0 chain over ALL rows first (it never needs a slot-1 decision) ----
    mid = {}
    for q in order:
        a, k, e = recs[q]
        x = [c0[a[0]], c0[a[1]], c0[a[2]]]
        L10, L20, L21 = x[1] < x[0], x[2] < x[0], x[2] < x[1]       # strict '<' in scan order: ties to the earlier position
        is2 = L21 if L10 else L20
        is1 = L10 and not L21
        w = 2 if is2 else (1 if is1 else 0)
        c0[a[w]] += 1
        p_, q_ = (1, 2) if w == 0 else ((0, 2) if w == 1 else (0, 1))
        mid[q] = (a[p_], a[q_], e[{(0, 1): 0, (0, 2): 1, (1, 2): 2}[(p_, q_)]], a[w], k)
    # ---- slot-1 chain ----
    out = {}
    for q in order:
        op, oq, e, oA, k = mid[q]
        pick = c1[oq] < c1[op] + e
        o1, o2 = (oq, op) if pick else (op, oq)
        c1[o1] += 1
        if k > 2:
            c2[o2] += 1                                              # slot 2: a plain sum (the emit kernel's atomicAdd)
        out[q] = [oA, o1, o2][:k]
    return out, c0, c1, c2


@pytest.mark.parametrize("shape", [dict(T=30, P=24, RF=3, N=40, R=5), dict(T=12, P=90, RF=3, N=60, R=6),      # capacity 2 / 5: levels
                                   dict(T=40, P=16, RF=3, N=120, R=6), dict(T=25, P=20, RF=2, N=30, R=5),   # capacity 1; RF 2
                                   dict(T=50, P=9, RF=1, N=12, R=4), dict(T=10, P=8, RF=3, N=6, R=3, n_old=6)])
def test_level_schedule_and_per_slot_chains_reproduce_the_reference(oracle, shape):
    rng = random.Random(1234)
    ran = 0
    for kind in ("structured", "random", "mixed"):
        cl = kab.synth.make_cluster(seed=0x51D + shape["T"], kind=kind, **shape)
        octx = oracle.OracleContext()
        exp, exp_len, est = util.oracle_dense(oracle, cl, octx)
        if est.code != 0:
            continue
        exp = exp.reshape(cl.T, cl.P, -1)
        sets = [[[int(b) for b in exp[t, p, :exp_len[t * cl.P + p]]] for p in range(cl.P)] for t in range(cl.T)]
        out, c0, c1, c2 = run_model(cl, sets, rng)
        for t in range(cl.T):
            for p in range(cl.P):
                got = [int(cl.broker_id[i]) for i in out[t * cl.P + p]]
                assert got == sets[t][p], (shape, kind, t, p)
        for i, b in enumerate(cl.broker_id):
            assert (c0[i], c1[i], c2[i]) == tuple(octx.counter(int(b), s) for s in range(3)), (shape, kind, int(b))
        ran += 1
    assert ran >= 1, shape
