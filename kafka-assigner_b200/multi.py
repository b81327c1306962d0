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


def ring_solve(rank, world, stage, order, export_counters, import_counters, ctr_buf, dist, final_broadcast=True, status=None,
               tensor_factory=None):
    """Run one topic-sharded solve.

    stage():            context-free stage of this rank's block (enqueue only)
    order():            leader-order stage of the staged block against the local Context
    export_counters(t)/import_counters(t): copy the local Context counters to / from tensor t
    ctr_buf:            a tensor [N*slots] int32 on this rank's device, identical shape on all ranks
    dist:               torch.distributed (or a stand-in with send/recv/broadcast)
    status():           optional; returns None or the GLOBAL index of this rank's first failing topic (synchronises the
                        rank). When given, the ranks agree on the lowest failing topic of the run, every rank raises
                        RunAborted and the final broadcast is skipped (counters are undefined after an error). Callers that
                        keep the solve asynchronous pass None and call agree_on_failure() themselves after synchronising.
    """
    stage()
    if rank > 0:
        dist.recv(ctr_buf, src=rank - 1)
        import_counters(ctr_buf)
    order()
    if rank < world - 1:
        export_counters(ctr_buf)
        dist.send(ctr_buf, dst=rank + 1)
    if status is not None:
        bad = agree_on_failure(status(), dist, tensor_factory)
        if bad != NO_FAILURE:
            raise RunAborted(bad)
    if final_broadcast and world > 1:
        if rank == world - 1:
            export_counters(ctr_buf)
        dist.broadcast(ctr_buf, src=world - 1)
        if rank != world - 1:
            import_counters(ctr_buf)
