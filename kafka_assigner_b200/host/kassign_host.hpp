// kassign_host.hpp — C++ host-side mirror of the reference's interface for the hot path, over the C ABI
// (include/kassign.h). Header-only; link with -lkassign.
//
//   kassign::KafkaTopicAssigner::generateAssignment   <->  siftscience.kafka.tools.KafkaTopicAssigner.generateAssignment
//                                                          (reference KafkaTopicAssigner.java:42-72)
//   kassign::solveTopics                              <->  the per-topic loop with ONE shared assigner
//                                                          (reference KafkaAssignmentGenerator.java:172-184)
//   kassign::newAssignmentJson                        <->  the org.json emitter (KafkaAssignmentGenerator.java:169-186)
//
// Same argument meaning and error behaviour: failures are re-thrown as IllegalStateException /
// ArrayIndexOutOfBoundsException with the reference's message texts (KTA:58-60, 65-66, 67-69; KAS:183-184, 190-192).
// All compute runs in libkassign.so's CUDA kernels; there is no CPU fallback — without a GPU the constructor throws.
#pragma once
#include <cstdint>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/kassign.h"

namespace kassign {

struct IllegalStateException : std::logic_error { using std::logic_error::logic_error; };
struct ArrayIndexOutOfBoundsException : std::out_of_range { using std::out_of_range::out_of_range; };
struct KassignError : std::runtime_error {
    int code;
    KassignError(int c, const std::string& what) : std::runtime_error("kassign error " + std::to_string(c) + " " + what), code(c) {}
};

using Assignment = std::map<int, std::vector<int>>;  // partition -> replicas (leader first); TreeMap order

// Re-throw a ka_status as the reference's exception with the identical message.
inline void throwForStatus(const ka_status& st, const std::vector<std::string>& topicNames) {
    if (st.code == KA_OK) return;
    const std::string topic = (st.topic_index >= 0 && (size_t)st.topic_index < topicNames.size()) ? topicNames[st.topic_index] : "?";
    switch (st.code) {
    case KA_ERR_RF_MISMATCH:
        throw IllegalStateException("Topic " + topic + " has partition " + std::to_string(st.partition) +
                                    " with unexpected replication factor " + std::to_string(st.a));
    case KA_ERR_RF_NOT_POSITIVE:
        throw IllegalStateException("Topic " + topic + " does not have a positive replication factor!");
    case KA_ERR_RF_GT_BROKERS:
        throw IllegalStateException("Topic " + topic + " has a higher replication factor (" + std::to_string(st.a) +
                                    ") than available brokers!");
    case KA_ERR_UNASSIGNABLE:
        throw IllegalStateException("Partition " + std::to_string(st.partition) + " could not be fully assigned!");
    case KA_ERR_HASH_INDEX:
        throw ArrayIndexOutOfBoundsException(std::to_string(st.a));
    default:
        throw KassignError(st.code, "(topic_index=" + std::to_string(st.topic_index) + ")");
    }
}

// One topic of a run: name + current assignment.
struct TopicInput {
    std::string name;
    Assignment current;
};

struct TopicOutput {
    std::string name;
    Assignment assignment;
};

class KafkaTopicAssigner {
public:
    // `new KafkaTopicAssigner()` (KTA:21-23): one Context per instance.
    explicit KafkaTopicAssigner(int device = 0) : ctx_(ka_ctx_create(device)) {
        if (!ctx_) throw KassignError(KA_ERR_NO_DEVICE, "no usable CUDA device: kassign has no CPU fallback");
    }
    ~KafkaTopicAssigner() { ka_ctx_destroy(ctx_); }
    KafkaTopicAssigner(const KafkaTopicAssigner&) = delete;
    KafkaTopicAssigner& operator=(const KafkaTopicAssigner&) = delete;

    // generateAssignment(topic, currentAssignment, brokers, rackAssignment, desiredReplicationFactor) — KTA:42-44.
    Assignment generateAssignment(const std::string& topic, const Assignment& currentAssignment, const std::set<int>& brokers,
                                  const std::map<int, std::string>& rackAssignment, int desiredReplicationFactor) {
        std::vector<TopicInput> one{{topic, currentAssignment}};
        return solveTopics(one, brokers, rackAssignment, desiredReplicationFactor)[0].assignment;
    }

    // The KAG:172-184 loop as ONE batched device call: topics in order through this instance's Context.
    std::vector<TopicOutput> solveTopics(const std::vector<TopicInput>& topics, const std::set<int>& brokers,
                                         const std::map<int, std::string>& rackAssignment, int desiredReplicationFactor) {
        setBrokers(brokers, rackAssignment);
        const int T = (int)topics.size();
        std::vector<int32_t> hash(T), partId, cur;
        std::vector<int64_t> partOff(T + 1, 0), repOff(1, 0);
        std::vector<std::string> names(T);
        int maxLen = 0;
        for (int t = 0; t < T; ++t) {
            names[t] = topics[t].name;
            hash[t] = ka_java_string_hash(topics[t].name.c_str());
            for (const auto& e : topics[t].current) {  // std::map: ascending partition == TreeMap order (KAS:107-110)
                partId.push_back(e.first);
                for (int b : e.second) cur.push_back(b);
                repOff.push_back((int64_t)cur.size());
                maxLen = std::max(maxLen, (int)e.second.size());
            }
            partOff[t + 1] = (int64_t)partId.size();
        }
        const int stride = std::max(1, std::max(maxLen, std::max(desiredReplicationFactor, 0)));
        const size_t Q = partId.size();
        std::vector<int32_t> outLen(Q, 0), out(Q * (size_t)stride, -1);
        ka_status st{};
        ka_solve(ctx_, T, hash.data(), partOff.data(), partId.data(), repOff.data(), cur.data(), desiredReplicationFactor, stride,
                 outLen.data(), out.data(), &st);
        throwForStatus(st, names);
        std::vector<TopicOutput> res(T);
        for (int t = 0; t < T; ++t) {
            res[t].name = names[t];
            for (int64_t g = partOff[t]; g < partOff[t + 1]; ++g)
                res[t].assignment[partId[g]] = std::vector<int>(out.begin() + g * stride, out.begin() + g * stride + outLen[g]);
        }
        return res;
    }

    ka_ctx* handle() { return ctx_; }

private:
    void setBrokers(const std::set<int>& brokers, const std::map<int, std::string>& racks) {
        std::vector<int32_t> ids(brokers.begin(), brokers.end());  // std::set: ascending == TreeMap order (KAS:78)
        if (ids == ids_ && racks == racks_) return;
        std::vector<const char*> names(ids.size(), nullptr);
        for (size_t i = 0; i < ids.size(); ++i) {
            auto it = racks.find(ids[i]);
            if (it != racks.end()) names[i] = it->second.c_str();
        }
        std::vector<int32_t> rackIdx(ids.size());
        int rc = ka_rack_indices((int32_t)ids.size(), ids.data(), names.data(), rackIdx.data());
        if (rc == KA_OK) rc = ka_ctx_set_brokers(ctx_, (int32_t)ids.size(), ids.data(), rackIdx.data());
        if (rc != KA_OK) throw KassignError(rc, "ka_ctx_set_brokers");
        ids_ = ids;
        racks_ = racks;
    }
    ka_ctx* ctx_;
    std::vector<int32_t> ids_;
    std::map<int, std::string> racks_;
};

// ---- JSON emission ------------------------------------------------------------------------------------------------
inline void appendInt(std::string& s, long long v) {
    char buf[24];
    int n = 0;
    unsigned long long u = v < 0 ? 0ULL - (unsigned long long)v : (unsigned long long)v;
    do { buf[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) s.push_back('-');
    while (n) s.push_back(buf[--n]);
}

// org.json JSONObject.quote(): escapes ", \, control chars and "</" (Kafka topic names never need it).
inline void appendQuoted(std::string& s, const std::string& v) {
    s.push_back('"');
    char prev = 0;
    for (unsigned char c : v) {
        switch (c) {
        case '\\': case '"': s.push_back('\\'); s.push_back((char)c); break;
        case '/': if (prev == '<') s.push_back('\\'); s.push_back('/'); break;
        case '\b': s += "\\b"; break;
        case '\t': s += "\\t"; break;
        case '\n': s += "\\n"; break;
        case '\f': s += "\\f"; break;
        case '\r': s += "\\r"; break;
        default:
            if (c < 0x20) { static const char* hx = "0123456789abcdef"; s += "\\u00"; s.push_back(hx[c >> 4]); s.push_back(hx[c & 15]); }
            else s.push_back((char)c);
        }
        prev = (char)c;
    }
    s.push_back('"');
}

// Key order of org.json 20131018 objects == java.util.HashMap iteration order of the keys (predicted for
// JDK >= 8, unverified without a JVM — isolated here so it can be corrected in one place):
//   top level: "partitions" (bucket 0) before "version" (13); per record: "partition" (3), "replicas" (6), "topic" (9).
inline std::string newAssignmentJson(const std::vector<TopicOutput>& topics) {
    std::string s = "{\"partitions\":[";
    bool first = true;
    for (const auto& t : topics)
        for (const auto& e : t.assignment) {  // ascending partition, topics in loop order (KAG:173-183)
            if (!first) s.push_back(',');
            first = false;
            s += "{\"partition\":";
            appendInt(s, e.first);
            s += ",\"replicas\":[";
            for (size_t i = 0; i < e.second.size(); ++i) {
                if (i) s.push_back(',');
                appendInt(s, e.second[i]);
            }
            s += "],\"topic\":";
            appendQuoted(s, t.name);
            s.push_back('}');
        }
    s += "],\"version\":1}";  // KAFKA_FORMAT_VERSION (KAG:49)
    return s;
}

// Kafka 0.10 ZkUtils.formatAsReassignmentJson shape (used for "CURRENT ASSIGNMENT:", KAG:103-111): scala Map literals keep
// insertion order for <= 4 entries: version, partitions / topic, partition, replicas.
inline std::string kafkaReassignmentJson(const std::vector<TopicInput>& topics) {
    std::string s = "{\"version\":1,\"partitions\":[";
    bool first = true;
    for (const auto& t : topics)
        for (const auto& e : t.current) {
            if (!first) s.push_back(',');
            first = false;
            s += "{\"topic\":";
            appendQuoted(s, t.name);
            s += ",\"partition\":";
            appendInt(s, e.first);
            s += ",\"replicas\":[";
            for (size_t i = 0; i < e.second.size(); ++i) {
                if (i) s.push_back(',');
                appendInt(s, e.second[i]);
            }
            s += "]}";
        }
    s += "]}";
    return s;
}

}  // namespace kassign
