"""py_oracle.py — second, independent CPU restatement of the reference hot path (TEST INFRASTRUCTURE).

Pure-Python, small cases only. Written separately from oracle/kafka_oracle.cpp (different data
structures: dicts + explicit sorted() where the Java code relies on TreeMap/TreeSet order) so that
agreement between the two restatements is evidence that each reads the Java correctly.

Follows (paths relative to /root/reference/src/main/java/siftscience/kafka/tools/):
  KafkaAssignmentStrategy.java (KAS) :40-369, KafkaTopicAssigner.java (KTA) :42-72,
  KafkaAssignmentGenerator.java (KAG) :172-184 (the topic loop with ONE shared assigner).

PARITY STATUS: "parity unpinned" against a live JVM (no Java toolchain in the build container);
pinned by the reference's own JUnit inputs/assertions (KafkaTopicAssignerTest.java:18-157).

Only tests/ may import this module.
"""
import math

INT_MIN = -(1 << 31)


class JavaError(Exception):
    """An exception the reference would throw. kind mirrors oracle ErrKind codes."""

    def __init__(self, kind, message, partition=-1, a=0, b=0):
        super().__init__(message)
        self.kind, self.message, self.partition, self.a, self.b = kind, message, partition, a, b


ERR_RF_MISMATCH, ERR_RF_NOT_POSITIVE, ERR_RF_GT_BROKERS, ERR_UNASSIGNABLE, ERR_INDEX = 1, 2, 3, 4, 5


def _i32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def java_string_hash(s: str) -> int:
    """java.lang.String.hashCode: s[0]*31^(n-1)+... over UTF-16 code units, int32 wrap."""
    h = 0
    data = s.encode("utf-16-be", "surrogatepass")
    for i in range(0, len(data), 2):
        h = (31 * h + ((data[i] << 8) | data[i + 1])) & 0xFFFFFFFF
    return _i32(h)


def java_abs(v: int) -> int:
    return v if v >= 0 or v == INT_MIN else -v


def java_rem(a: int, b: int) -> int:
    """Java % : remainder truncated toward zero (sign of dividend)."""
    return int(math.fmod(a, b))


class Context:
    """KAS:360-369. counter[node][replica_slot] -> count."""

    def __init__(self):
        self.counter = {}


def node_processing_order(topic_hash, node_ids_sorted):
    """KAS:188-200. node_ids_sorted: iteration order of the (sorted) Java collection."""
    n = len(node_ids_sorted)
    order = [None] * n
    index = java_rem(java_abs(topic_hash), n)
    for node_id in node_ids_sorted:
        if index < 0 or index >= n:
            raise JavaError(ERR_INDEX, str(index), -1, index, n)  # ArrayIndexOutOfBoundsException
        order[index] = node_id
        index += 1
        if index == n:
            index = 0
    return order


def rack_aware_assignment(topic_hash, current, node_rack, nodes, partitions, rf, context):
    """KAS:40-63. current: list of (partition, [brokers]) in the input map's entry order."""
    n_nodes = len(nodes)
    # KAS:65-71 — int multiply widened to double, ceil of the double quotient, (int) cast.
    total = float(_i32(len(partitions) * rf))
    cap = int(math.ceil(total / n_nodes))

    # KAS:73-99 — node table; rack key = rack string, or the node id's decimal string when absent.
    rack_of = {}
    for nid in nodes:
        r = node_rack.get(nid)
        rack_of[nid] = str(nid) if r is None else r
    node_parts = {nid: set() for nid in nodes}
    rack_parts = {r: set() for r in set(rack_of.values())}

    def can_accept(nid, p):  # KAS:320-324 + 346-348
        return p not in node_parts[nid] and len(node_parts[nid]) < cap and p not in rack_parts[rack_of[nid]]

    def accept(nid, p):  # KAS:326-331 + 350-354
        assert can_accept(nid, p)
        node_parts[nid].add(p)
        rack_parts[rack_of[nid]].add(p)

    # KAS:101-131 — sticky fill: round-robin over replica slots, partitions ascending.
    lists = {}
    for p, reps in current:
        lists[p] = list(reps)
    live = sorted(lists)
    pos = {p: 0 for p in live}
    while live:
        nxt = []
        for p in live:
            if pos[p] < len(lists[p]):
                nid = lists[p][pos[p]]
                pos[p] += 1
                if nid in node_parts and can_accept(nid, p):
                    accept(nid, p)
                nxt.append(p)
            # else: exhausted iterator is removed from the round-robin (KAS:125-127)
        live = nxt

    # KAS:133-160 — orphans per partition (ascending), only if > 0.
    held = {}
    for nid in sorted(node_parts):
        for p in node_parts[nid]:
            held[p] = held.get(p, 0) + 1
    orphans = []
    for p in sorted(partitions):
        rem = rf - held.get(p, 0)
        if rem > 0:
            orphans.append((p, rem))

    # KAS:162-186 — first-fit in the rotated order, restart at j=0 for every orphan.
    order = node_processing_order(topic_hash, sorted(node_parts))
    for p, rem in orphans:
        for nid in order:
            if rem <= 0:
                break
            if can_accept(nid, p):
                accept(nid, p)
                rem -= 1
        if rem != 0:
            raise JavaError(ERR_UNASSIGNABLE, "Partition %d could not be fully assigned!" % p, p)

    # KAS:202-239 — leader-preference ordering with the cross-topic counters.
    unordered = {}
    for nid in sorted(node_parts):
        for p in sorted(node_parts[nid]):
            unordered.setdefault(p, []).append(nid)
    counter = context.counter

    def ensure(nid, slot):  # KAS:289-301
        return counter.setdefault(nid, {}).setdefault(slot, 0)

    prefs = {}
    for p in sorted(unordered):
        remaining = sorted(set(unordered[p]))
        k = len(unordered[p])
        ordered = []
        for slot in range(k):
            best, best_c = None, None
            for nid in node_processing_order(topic_hash, remaining):  # KAS:263-278
                c = ensure(nid, slot)
                if best_c is None or c < best_c:
                    best, best_c = nid, c
            remaining.remove(best)
            ordered.append(best)
        for slot, nid in enumerate(ordered):  # KAS:254-261
            counter[nid][slot] = _i32(ensure(nid, slot) + 1)
        prefs[p] = ordered
    return prefs


class KafkaTopicAssigner:
    """KTA:18-72. One instance owns one Context (KTA:19-23)."""

    def __init__(self):
        self.context = Context()

    def generate_assignment(self, topic, current_assignment, brokers, rack_assignment, desired_rf):
        """current_assignment: dict or list of (partition, [brokers]) — iterated in the given order."""
        entries = list(current_assignment.items()) if isinstance(current_assignment, dict) else list(current_assignment)
        rf = desired_rf
        partitions = set()
        for p, reps in entries:
            partitions.add(p)
            if rf < 0:
                rf = len(reps)
            elif desired_rf < 0:
                if rf != len(reps):
                    raise JavaError(ERR_RF_MISMATCH,
                                    "Topic %s has partition %d with unexpected replication factor %d" % (topic, p, len(reps)),
                                    p, len(reps))
        if not rf > 0:
            raise JavaError(ERR_RF_NOT_POSITIVE, "Topic %s does not have a positive replication factor!" % topic)
        if not rf <= len(brokers):
            raise JavaError(ERR_RF_GT_BROKERS,
                            "Topic %s has a higher replication factor (%d) than available brokers!" % (topic, rf), -1, rf)
        return rack_aware_assignment(java_string_hash(topic), entries, rack_assignment, list(brokers),
                                     partitions, rf, self.context)


def run_topics(topics, brokers, rack_assignment, desired_rf=-1, assigner=None):
    """KAG:172-184: topics = [(name, current_assignment)], ONE assigner for the whole loop.
    Returns the flat record stream [(topic, partition, [replicas])] in output order."""
    assigner = assigner or KafkaTopicAssigner()
    out = []
    for name, cur in topics:
        fin = assigner.generate_assignment(name, cur, brokers, rack_assignment, desired_rf)
        for p in sorted(fin):
            out.append((name, p, fin[p]))
    return out
