// kafka_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A structure-faithful C++17 restatement of the reference's hot path:
//   siftscience.kafka.tools.KafkaAssignmentStrategy   (KAS = src/main/java/siftscience/kafka/tools/KafkaAssignmentStrategy.java)
//   siftscience.kafka.tools.KafkaTopicAssigner        (KTA = .../KafkaTopicAssigner.java)
//   the per-topic loop of KafkaAssignmentGenerator    (KAG = .../KafkaAssignmentGenerator.java:172-184)
// java.util.TreeMap/TreeSet are mirrored by std::map/std::set so every visit order below is the
// reference's by construction; java.util.HashMap is mirrored by std::unordered_map (only get/put
// semantics are used on those, never iteration order).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this file's library. The product library (libkassign.so) never links or calls it.
//
// PARITY STATUS: "parity unpinned" against a live JVM — no java/javac/jar exists in the build
// container, so the real reference cannot be executed. The oracle is pinned instead by
//   (1) the reference's own four JUnit inputs (TEST = src/test/.../KafkaTopicAssignerTest.java:18-157),
//       whose assertions (including the single exact pin TEST:143-144, p0 == [10,11]) it satisfies,
//   (2) an independent second restatement (oracle/py_oracle.py) that must agree bit-for-bit,
//   (3) the hand traces recorded in SURVEY.md §8c.
//
// Build: g++ -O2 -std=c++17 -shared -fPIC -o oracle/liboracle.so oracle/kafka_oracle.cpp

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

// ---- Java semantics helpers -------------------------------------------------------------------

// java.lang.String.hashCode over UTF-16 code units with int32 wrap-around (used at KAS:190).
int32_t java_string_hash_utf8(const char* s, size_t n) {
    uint32_t h = 0;
    size_t i = 0;
    while (i < n) {
        uint32_t c = (unsigned char)s[i];
        uint32_t cp;
        int extra;
        if (c < 0x80) { cp = c; extra = 0; }
        else if ((c >> 5) == 0x6) { cp = c & 0x1F; extra = 1; }
        else if ((c >> 4) == 0xE) { cp = c & 0x0F; extra = 2; }
        else if ((c >> 3) == 0x1E) { cp = c & 0x07; extra = 3; }
        else { cp = 0xFFFD; extra = 0; }
        ++i;
        for (int k = 0; k < extra && i < n; ++k, ++i) cp = (cp << 6) | ((unsigned char)s[i] & 0x3F);
        if (cp >= 0x10000) {  // surrogate pair: two UTF-16 units
            cp -= 0x10000;
            h = 31u * h + (0xD800u + (cp >> 10));
            h = 31u * h + (0xDC00u + (cp & 0x3FF));
        } else {
            h = 31u * h + cp;
        }
    }
    return (int32_t)h;
}

// java.lang.Math.abs(int): abs(Integer.MIN_VALUE) == Integer.MIN_VALUE.
int32_t java_abs(int32_t v) { return v < 0 ? (int32_t)(0u - (uint32_t)v) : v; }

// Java (int) cast of a double: saturating, NaN -> 0.
int32_t java_d2i(double d) {
    if (std::isnan(d)) return 0;
    if (d >= 2147483647.0) return INT32_MAX;
    if (d <= -2147483648.0) return INT32_MIN;
    return (int32_t)d;
}

// Exceptions the reference path can raise (SURVEY §8b "Errors").
enum ErrKind {
    ERR_NONE = 0,
    ERR_RF_MISMATCH = 1,      // IllegalStateException KTA:58-60
    ERR_RF_NOT_POSITIVE = 2,  // IllegalStateException KTA:65-66
    ERR_RF_GT_BROKERS = 3,    // IllegalStateException KTA:67-69
    ERR_UNASSIGNABLE = 4,     // IllegalStateException KAS:183-184
    ERR_INDEX = 5,            // ArrayIndexOutOfBoundsException KAS:190-192 (hashCode == MIN_VALUE)
    ERR_ARG = 6,              // IllegalArgumentException KAS:327-328 / 351-352 (unreachable)
};

struct JavaException : std::runtime_error {
    ErrKind kind;
    int32_t partition, a, b;
    JavaException(ErrKind k, const std::string& m, int32_t p = -1, int32_t a_ = 0, int32_t b_ = 0)
        : std::runtime_error(m), kind(k), partition(p), a(a_), b(b_) {}
};

// ---- KAS:337-355 Rack ---------------------------------------------------------------------------
struct Rack {
    std::string id;
    std::set<int> assignedPartitions;
    explicit Rack(std::string i) : id(std::move(i)) {}
    bool canAccept(int partition) const { return assignedPartitions.count(partition) == 0; }  // KAS:346-348
    void accept(int partition) {                                                              // KAS:350-354
        if (!canAccept(partition))
            throw JavaException(ERR_ARG, "Attempted to accept unacceptable partition " + std::to_string(partition), partition);
        assignedPartitions.insert(partition);
    }
};

// ---- KAS:307-332 Node ---------------------------------------------------------------------------
struct Node {
    int id;
    int capacity;
    Rack* rack;
    std::set<int> assignedPartitions;
    Node(int i, int c, Rack* r) : id(i), capacity(c), rack(r) {}
    bool canAccept(int partition) const {  // KAS:320-324
        return assignedPartitions.count(partition) == 0 &&
               (int)assignedPartitions.size() < capacity &&
               rack->canAccept(partition);
    }
    void accept(int partition) {  // KAS:326-331
        if (!canAccept(partition))
            throw JavaException(ERR_ARG, "Attempted to accept unacceptable partition " + std::to_string(partition), partition);
        assignedPartitions.insert(partition);
        rack->accept(partition);
    }
};

// ---- KAS:360-369 Context ------------------------------------------------------------------------
struct Context {
    std::unordered_map<int, std::unordered_map<int, int>> counter;
};

using Assignment = std::map<int, std::vector<int>>;          // TreeMap<Integer, List<Integer>>
using InputAssignment = std::vector<std::pair<int, std::vector<int>>>;  // Map in its entrySet() order

struct NodeMap {
    std::map<std::string, std::unique_ptr<Rack>> rackMap;  // Maps.newTreeMap() KAS:77
    std::map<int, std::unique_ptr<Node>> nodeMap;          // Maps.newTreeMap() KAS:78
};

// ---- KAS:65-71 ----------------------------------------------------------------------------------
int getMaxReplicasPerNode(size_t nNodes, size_t nPartitions, int replicationFactor) {
    // `partitions.size() * replicationFactor` is an int multiply widened to double.
    double totalReplicas = (double)(int32_t)((uint32_t)nPartitions * (uint32_t)replicationFactor);
    return java_d2i(std::ceil(totalReplicas / (double)nNodes));
}

// ---- KAS:73-99 ----------------------------------------------------------------------------------
void createNodeMap(NodeMap& nm, const std::unordered_map<int, std::string>& nodeRackAssignment,
                   const std::vector<int>& nodes, int maxReplicas) {
    for (int nodeId : nodes) {
        if (nm.nodeMap.count(nodeId)) throw JavaException(ERR_ARG, "duplicate node id");  // checkState KAS:80
        std::string rackId;
        auto it = nodeRackAssignment.find(nodeId);
        if (it == nodeRackAssignment.end()) rackId = std::to_string(nodeId);  // KAS:82-86
        else rackId = it->second;
        Rack* rack;
        auto rit = nm.rackMap.find(rackId);
        if (rit == nm.rackMap.end()) {
            rack = new Rack(rackId);
            nm.rackMap.emplace(rackId, std::unique_ptr<Rack>(rack));
        } else {
            rack = rit->second.get();
        }
        nm.nodeMap.emplace(nodeId, std::make_unique<Node>(nodeId, maxReplicas, rack));
    }
}

// ---- KAS:101-131 --------------------------------------------------------------------------------
void fillNodesFromAssignment(const InputAssignment& assignment, NodeMap& nm) {
    // TreeMap<partition, Iterator>: ascending partition order (KAS:107-110).
    std::map<int, std::pair<const std::vector<int>*, size_t>> assignmentIterators;
    for (const auto& e : assignment) assignmentIterators[e.first] = {&e.second, 0};
    bool filled = false;
    while (!filled) {
        for (auto rr = assignmentIterators.begin(); rr != assignmentIterators.end();) {
            int partition = rr->first;
            auto& nodeIt = rr->second;
            if (nodeIt.second < nodeIt.first->size()) {
                int nodeId = (*nodeIt.first)[nodeIt.second++];
                auto nit = nm.nodeMap.find(nodeId);
                if (nit != nm.nodeMap.end() && nit->second->canAccept(partition)) nit->second->accept(partition);
                ++rr;
            } else {
                rr = assignmentIterators.erase(rr);  // roundRobin.remove() KAS:126
            }
        }
        filled = assignmentIterators.empty();
    }
}

// ---- KAS:133-160 --------------------------------------------------------------------------------
std::map<int, int> getOrphanedReplicas(const NodeMap& nm, const std::set<int>& partitions, int replicationFactor) {
    std::map<int, int> partitionCounter;
    for (const auto& kv : nm.nodeMap)
        for (int partition : kv.second->assignedPartitions) partitionCounter[partition] += 1;
    std::map<int, int> orphanedReplicas;
    for (int partition : partitions) {
        int remainingReplicas = replicationFactor;
        auto it = partitionCounter.find(partition);
        if (it != partitionCounter.end()) remainingReplicas -= it->second;
        if (remainingReplicas > 0) orphanedReplicas[partition] = remainingReplicas;
    }
    return orphanedReplicas;
}

// ---- KAS:188-200 --------------------------------------------------------------------------------
// nodeIds must be in the iteration order of the Java collection passed in (always a sorted set).
std::vector<int> getNodeProcessingOrder(int32_t topicHash, const std::vector<int>& nodeIds) {
    int32_t len = (int32_t)nodeIds.size();
    std::vector<int> order(len);
    // Java: Math.abs(hash) % len, with truncated remainder; len == 0 would be ArithmeticException.
    if (len == 0) throw JavaException(ERR_ARG, "/ by zero");  // ArithmeticException; unreachable (RF <= N, RF > 0)
    int32_t index = java_abs(topicHash) % len;
    for (int nodeId : nodeIds) {
        if (index < 0 || index >= len)
            throw JavaException(ERR_INDEX, std::to_string(index), -1, index, len);  // AIOOBE
        order[index] = nodeId;
        if (++index == len) index = 0;
    }
    return order;
}

// ---- KAS:162-186 --------------------------------------------------------------------------------
void assignOrphans(int32_t topicHash, NodeMap& nm, const std::map<int, int>& orphanedReplicas) {
    std::vector<int> keys;
    keys.reserve(nm.nodeMap.size());
    for (const auto& kv : nm.nodeMap) keys.push_back(kv.first);
    std::vector<int> nodeProcessingOrder = getNodeProcessingOrder(topicHash, keys);
    for (const auto& e : orphanedReplicas) {
        int partition = e.first;
        int remainingReplicas = e.second;
        auto nodeIt = nodeProcessingOrder.begin();
        while (nodeIt != nodeProcessingOrder.end() && remainingReplicas > 0) {
            Node* node = nm.nodeMap.find(*nodeIt++)->second.get();
            if (node->canAccept(partition)) {
                node->accept(partition);
                remainingReplicas--;
            }
        }
        if (remainingReplicas != 0)
            throw JavaException(ERR_UNASSIGNABLE, "Partition " + std::to_string(partition) + " could not be fully assigned!", partition);
    }
}

// ---- KAS:244-302 --------------------------------------------------------------------------------
struct PreferenceListOrderTracker {
    int32_t topicHash;
    std::unordered_map<int, std::unordered_map<int, int>>& nodeAssignmentCounters;

    int ensureCount(int nodeId, int replicaId) {  // KAS:289-301 (lazy zero, materialised on read)
        auto& replicaCount = nodeAssignmentCounters[nodeId];
        auto it = replicaCount.find(replicaId);
        if (it == replicaCount.end()) { replicaCount[replicaId] = 0; return 0; }
        return it->second;
    }
    void updateCountersFromList(const std::vector<int>& preferences) {  // KAS:254-261
        int replica = 0;
        for (int nodeId : preferences) {
            int currentCount = ensureCount(nodeId, replica);
            nodeAssignmentCounters[nodeId][replica] = (int32_t)((uint32_t)currentCount + 1u);
            replica++;
        }
    }
    int getLeastSeenNodeForReplicaId(int replicaId, const std::set<int>& nodes) {  // KAS:263-278
        bool have = false;
        int minCount = 0, minNode = 0;
        std::vector<int> v(nodes.begin(), nodes.end());
        std::vector<int> order = getNodeProcessingOrder(topicHash, v);
        for (int nodeId : order) {
            int count = ensureCount(nodeId, replicaId);
            if (!have || count < minCount) { have = true; minCount = count; minNode = nodeId; }
        }
        return minNode;
    }
};

// ---- KAS:202-239 --------------------------------------------------------------------------------
Assignment computePreferenceLists(int32_t topicHash, const NodeMap& nm, Context& context) {
    Assignment unorderedPreferences;
    for (const auto& kv : nm.nodeMap) {
        int nodeId = kv.second->id;
        for (int partition : kv.second->assignedPartitions) unorderedPreferences[partition].push_back(nodeId);
    }
    PreferenceListOrderTracker tracker{topicHash, context.counter};
    Assignment preferences;
    for (const auto& e : unorderedPreferences) {
        int partitionId = e.first;
        const std::vector<int>& preferenceList = e.second;
        std::vector<int> ordered;
        int replicationFactor = (int)preferenceList.size();
        std::set<int> nodeSet(preferenceList.begin(), preferenceList.end());
        for (int replica = 0; replica < replicationFactor; replica++) {
            int nodeToSelect = tracker.getLeastSeenNodeForReplicaId(replica, nodeSet);
            nodeSet.erase(nodeToSelect);
            ordered.push_back(nodeToSelect);
        }
        tracker.updateCountersFromList(ordered);
        preferences[partitionId] = std::move(ordered);
    }
    return preferences;
}

// ---- KAS:40-63 ----------------------------------------------------------------------------------
Assignment getRackAwareAssignment(int32_t topicHash, const InputAssignment& currentAssignment,
                                  const std::unordered_map<int, std::string>& nodeRackAssignment,
                                  const std::vector<int>& nodes, const std::set<int>& partitions,
                                  int replicationFactor, Context& context) {
    int maxReplicas = getMaxReplicasPerNode(nodes.size(), partitions.size(), replicationFactor);
    NodeMap nm;
    createNodeMap(nm, nodeRackAssignment, nodes, maxReplicas);
    fillNodesFromAssignment(currentAssignment, nm);
    std::map<int, int> orphanedReplicas = getOrphanedReplicas(nm, partitions, replicationFactor);
    assignOrphans(topicHash, nm, orphanedReplicas);
    return computePreferenceLists(topicHash, nm, context);
}

// ---- KTA:42-72 ----------------------------------------------------------------------------------
Assignment generateAssignment(Context& ctx, const std::string& topic, int32_t topicHash,
                              const InputAssignment& currentAssignment, const std::vector<int>& brokers,
                              const std::unordered_map<int, std::string>& rackAssignment,
                              int desiredReplicationFactor) {
    int replicationFactor = desiredReplicationFactor;
    std::set<int> partitions;
    for (const auto& entry : currentAssignment) {
        int partition = entry.first;
        const auto& replicas = entry.second;
        partitions.insert(partition);
        if (replicationFactor < 0) {
            replicationFactor = (int)replicas.size();
        } else if (desiredReplicationFactor < 0) {
            if (replicationFactor != (int)replicas.size())
                throw JavaException(ERR_RF_MISMATCH,
                                    "Topic " + topic + " has partition " + std::to_string(partition) +
                                        " with unexpected replication factor " + std::to_string(replicas.size()),
                                    partition, (int)replicas.size());
        }
    }
    if (!(replicationFactor > 0))
        throw JavaException(ERR_RF_NOT_POSITIVE, "Topic " + topic + " does not have a positive replication factor!");
    if (!(replicationFactor <= (int)brokers.size()))
        throw JavaException(ERR_RF_GT_BROKERS,
                            "Topic " + topic + " has a higher replication factor (" +
                                std::to_string(replicationFactor) + ") than available brokers!",
                            -1, replicationFactor);
    return getRackAwareAssignment(topicHash, currentAssignment, rackAssignment, brokers, partitions,
                                  replicationFactor, ctx);
}

}  // namespace

// ==================================================================================================
// C interface (ctypes). Flat arrays in the same layout as include/kassign.h so a test can feed the
// oracle and the CUDA library the very same buffers.
// ==================================================================================================
extern "C" {

struct oracle_status {
    int32_t code;         // ErrKind
    int32_t topic_index;  // failing topic (loop order, KAG:173)
    int32_t partition;    // partition id where meaningful, else -1
    int32_t a, b;         // operands of the message (size / RF / index,len)
    char message[256];    // the Java exception message text
};

void* oracle_ctx_create() { return new Context(); }
void oracle_ctx_destroy(void* c) { delete (Context*)c; }
void oracle_ctx_reset(void* c) { ((Context*)c)->counter.clear(); }

int32_t oracle_java_string_hash(const char* utf8) { return java_string_hash_utf8(utf8, std::strlen(utf8)); }

// Read Context.counter[brokerId][slot] (0 if absent) — for tests that compare counters with the GPU.
int32_t oracle_ctx_get_counter(void* c, int32_t broker_id, int32_t slot) {
    auto& m = ((Context*)c)->counter;
    auto it = m.find(broker_id);
    if (it == m.end()) return 0;
    auto jt = it->second.find(slot);
    return jt == it->second.end() ? 0 : jt->second;
}

// Seed Context.counter[brokerId][slot] — lets a test hand a Context from one process to another.
void oracle_ctx_set_counter(void* c, int32_t broker_id, int32_t slot, int32_t value) {
    ((Context*)c)->counter[broker_id][slot] = value;
}

// The KAG:172-184 loop: topics in order through ONE assigner/Context; stops at the first exception.
//   topic_names: T NUL-terminated UTF-8 strings, concatenated; name_off[T+1] byte offsets (incl. NULs)
//   part_off[T+1]: partition ranges; part_id[ΣP]: partition ids in the ENTRY ORDER of the input map
//   rep_off[ΣP+1]: replica ranges into cur_broker
//   brokers[N] in any order; rack_off[N+1]/rack_blob: rack names, rack_off[i]==rack_off[i+1] ⇒ no rack
//   out: out_len[ΣP] list lengths and out_broker[ΣP*out_stride] (row g = g-th partition in ASCENDING
//   partition-id order within its topic, i.e. the TreeMap order of KAG:177); unused tail = -1.
int oracle_run(void* cptr, int32_t T, const char* topic_names, const int64_t* name_off,
               const int64_t* part_off, const int32_t* part_id, const int64_t* rep_off,
               const int32_t* cur_broker, int32_t N, const int32_t* brokers, const char* rack_blob,
               const int64_t* rack_off, int32_t desired_rf, int32_t out_stride, int32_t* out_len,
               int32_t* out_part_id, int32_t* out_broker, oracle_status* st) {
    Context& ctx = *(Context*)cptr;
    std::memset(st, 0, sizeof(*st));
    st->topic_index = -1;
    st->partition = -1;
    std::vector<int> brokerVec(brokers, brokers + N);
    std::unordered_map<int, std::string> rackAssignment;
    for (int i = 0; i < N; ++i)
        if (rack_off && rack_off[i + 1] > rack_off[i])
            rackAssignment[brokers[i]] = std::string(rack_blob + rack_off[i], (size_t)(rack_off[i + 1] - rack_off[i]));
    for (int32_t t = 0; t < T; ++t) {
        std::string topic(topic_names + name_off[t]);
        int32_t h = java_string_hash_utf8(topic.data(), topic.size());
        InputAssignment cur;
        for (int64_t g = part_off[t]; g < part_off[t + 1]; ++g)
            cur.emplace_back(part_id[g], std::vector<int>(cur_broker + rep_off[g], cur_broker + rep_off[g + 1]));
        try {
            Assignment fin = generateAssignment(ctx, topic, h, cur, brokerVec, rackAssignment, desired_rf);
            // Flatten in TreeMap order (KAG:177-183). A partition with zero accepted replicas does not
            // appear in `preferences` at all (KAS:205-214) — impossible after a successful assignOrphans
            // because RF > 0, but keep the general shape: rows are emitted for the map's entries only.
            int64_t g = part_off[t];
            for (const auto& e : fin) {
                if ((int)e.second.size() > out_stride) {
                    st->code = -1;
                    std::snprintf(st->message, sizeof(st->message), "out_stride too small");
                    st->topic_index = t;
                    return -1;
                }
                out_part_id[g] = e.first;
                out_len[g] = (int32_t)e.second.size();
                for (int r = 0; r < out_stride; ++r)
                    out_broker[g * out_stride + r] = r < (int)e.second.size() ? e.second[r] : -1;
                ++g;
            }
            for (; g < part_off[t + 1]; ++g) {  // duplicate partition ids in the input collapse in the map
                out_part_id[g] = -1;
                out_len[g] = 0;
                for (int r = 0; r < out_stride; ++r) out_broker[g * out_stride + r] = -1;
            }
        } catch (const JavaException& e) {
            st->code = e.kind;
            st->topic_index = t;
            st->partition = e.partition;
            st->a = e.a;
            st->b = e.b;
            std::snprintf(st->message, sizeof(st->message), "%s", e.what());
            return e.kind;
        }
    }
    return 0;
}

}  // extern "C"
