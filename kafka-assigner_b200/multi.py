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


def ring_solve(rank, world, stage, order, export_counters, import_counters, ctr_buf, dist, final_broadcast=True):
    """Run one topic-sharded solve.

    stage():            context-free stage of this rank's block (enqueue only)
    order():            leader-order stage of the staged block against the local Context
    export_counters(t)/import_counters(t): copy the local Context counters to / from tensor t
    ctr_buf:            a tensor [N*slots] int32 on this rank's device, identical shape on all ranks
    dist:               torch.distributed (or a stand-in with send/recv/broadcast)
    """
    stage()
    if rank > 0:
        dist.recv(ctr_buf, src=rank - 1)
        import_counters(ctr_buf)
    order()
    if rank < world - 1:
        export_counters(ctr_buf)
        dist.send(ctr_buf, dst=rank + 1)
    if final_broadcast and world > 1:
        if rank == world - 1:
            export_counters(ctr_buf)
        dist.broadcast(ctr_buf, src=world - 1)
        if rank != world - 1:
            import_counters(ctr_buf)
