"""Topic-sharded multi-GPU driver: one process per GPU, torch.distributed (NCCL) for the plumbing.

What shards and what does not:
  * capacity / sticky fill / orphan spread (KafkaAssignmentStrategy.java:65-200) carry no cross-topic
    state, so rank g stages its own contiguous topic block with NO data-path collective;
  * leader-preference ordering (KAS:202-239) reads and bumps `Context.counter`, which the reference keeps
    in ONE KafkaTopicAssigner for the whole run (KafkaTopicAssigner.java:19-23,
    KafkaAssignmentGenerator.java:172) — a strict serial chain over all topics in order. The only exact
    distribution is a ring hand-off: rank g orders its block after receiving the counter table from rank
    g-1 (N x slots int32, <= 320 KB) and forwards it to rank g+1. A final broadcast from the last rank
    leaves every rank's Context equal to the reference's Context after the whole run.

The compute callbacks are injected, so the protocol itself is testable on CPU with the gloo backend and
the oracle as the stand-in backend (tests/test_multi_gloo.py).
"""


def shard_range(total_topics, world, rank):
    """Contiguous topic block of `rank`: [t0, t1). Blocks differ by at most one topic."""
    base, extra = divmod(total_topics, world)
    t0 = rank * base + min(rank, extra)
    return t0, t0 + base + (1 if rank < extra else 0)


class RunAborted(RuntimeError):
    """A topic of the run failed on some rank: the reference aborts the whole run at the first failing topic
    (KafkaAssignmentGenerator.java:173-186 never prints), so every rank raises, with the lowest failing topic index."""

    def __init__(self, topic_index):
        super().__init__("run aborted: first failing topic %d" % topic_index)
        self.topic_index = topic_index


NO_FAILURE = 2**31 - 1


def agree_on_failure(local_first_bad, dist, tensor_factory):
    """All-reduce (MIN) the first failing GLOBAL topic index of every rank; NO_FAILURE when a rank saw none."""
    t = tensor_factory([NO_FAILURE if local_first_bad is None else int(local_first_bad)])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return int(t[0])


def _p2p(dist, op, tensor, peer, group):
    """One send or recv as a batched P2P op (NCCL: enqueued on the current stream, no host block; gloo: completes here)."""
    if hasattr(dist, "batch_isend_irecv") and hasattr(dist, "P2POp"):
        fn = dist.isend if op == "send" else dist.irecv
        for req in dist.batch_isend_irecv([dist.P2POp(fn, tensor, peer, group)]):
            req.wait()
    elif op == "send":
        dist.send(tensor, dst=peer)
    else:
        dist.recv(tensor, src=peer)


def ring_solve_phases(rank, world, stage, phases, dist, finish=None, final_broadcast=True, final_sums=(), status=None,
                      tensor_factory=None, group=None):
    """Topic-sharded solve as a pipeline of serial chains handed from rank to rank.

    stage():   context-free stage of this rank's block (enqueue only).
    phases:    list (or a callable returning the list, evaluated after stage()) of (run, export, import_, buf): every phase is
               ONE serial chain over all topics of the run. Rank g receives buf from g-1, imports it, runs its part of the
               chain, exports and sends to g+1 — then moves on to the next phase, so chain p+1 of rank g overlaps chain p of
               the ranks behind it. With the per-slot leader-order chains of libkassign (slot r touches only counter[.][r])
               the critical path is  world x slot0 + slot1  instead of  world x (slot0 + slot1).
    finish():  after the last phase (e.g. the parallel emit).
    final_sums: (export_delta, add_total, buf) triples for state that is a commutative SUM over the ranks' blocks — the slot-2
               counters of rows <= 3: one all-reduce, the only collective of the path (and it never feeds a decision).
    status / tensor_factory: as in ring_solve.
    """
    stage()
    if callable(phases):
        phases = phases()
    for run, export, import_, buf in phases:
        if rank > 0:
            _p2p(dist, "recv", buf, rank - 1, group)
            import_(buf)
        run()
        if rank < world - 1:
            export(buf)
            _p2p(dist, "send", buf, rank + 1, group)
    if finish is not None:
        finish()
    if status is not None:
        bad = agree_on_failure(status(), dist, tensor_factory)
        if bad != NO_FAILURE:
            raise RunAborted(bad)
    if final_broadcast and world > 1:
        for run, export, import_, buf in phases:
            if rank == world - 1:
                export(buf)
            dist.broadcast(buf, src=world - 1, group=group)
            if rank != world - 1:
                import_(buf)
        for export_delta, add_total, buf in final_sums:
            export_delta(buf)
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
            add_total(buf)


def ring_solve(rank, world, stage, order, export_counters, import_counters, ctr_buf, dist, final_broadcast=True, status=None,
               tensor_factory=None):
    """Single-chain form (the whole counter table travels with one leader-order stage per rank): used for rows of 4..8
    replicas and by backends that cannot split the slots. See ring_solve_phases."""
    return ring_solve_phases(rank, world, stage, [(order, export_counters, import_counters, ctr_buf)], dist,
                             final_broadcast=final_broadcast, status=status, tensor_factory=tensor_factory)
