// test_kafka_topic_assigner.cpp — the reference's JUnit class, re-expressed against the C++ host mirror
// (reference src/test/java/siftscience/kafka/tools/KafkaTopicAssignerTest.java:18-187). Same inputs, same
// assertions (load histograms, stickiness, the exact pin newAssignment.get(0) == [10, 11]), plus the exception texts.
// Needs a GPU (kassign has no CPU fallback). Exit code 0 = all passed.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "kassign_host.hpp"

using kassign::Assignment;

static int failures = 0;
#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) { std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } \
    } while (0)

// KafkaTopicAssignerTest.verifyPartitionsAndBuildReplicaCounts (TEST:159-187)
static std::map<int, int> verifyPartitionsAndBuildReplicaCounts(const Assignment& cur, const Assignment& neu, int minimalMovementThreshold) {
    std::map<int, int> brokerReplicaCounts;
    for (const auto& e : neu) {
        std::set<int> replicaSet(e.second.begin(), e.second.end());
        CHECK(replicaSet.size() == e.second.size());  // no broker twice in a partition
        for (int b : e.second) brokerReplicaCounts[b]++;
        int stuck = 0;
        for (int b : cur.at(e.first)) stuck += (int)replicaSet.count(b);
        CHECK(stuck >= minimalMovementThreshold);       // movement was minimal
    }
    return brokerReplicaCounts;
}

static const Assignment CUR_A = {{0, {10, 11}}, {1, {11, 12}}, {2, {12, 10}}, {3, {10, 12}}};

static void testRackAwareExpansion() {  // TEST:18-57
    kassign::KafkaTopicAssigner assigner;
    Assignment neu = assigner.generateAssignment("test", CUR_A, {10, 11, 12, 13, 14}, {{10, "a"}, {11, "b"}, {12, "c"}, {13, "a"}, {14, "b"}}, -1);
    auto counts = verifyPartitionsAndBuildReplicaCounts(CUR_A, neu, 1);
    int one = 0, two = 0;
    for (auto& kv : counts) { one += kv.second == 1; two += kv.second == 2; }
    CHECK(one == 2);
    CHECK(two == 3);
}

static void testClusterExpansion() {  // TEST:59-82
    kassign::KafkaTopicAssigner assigner;
    Assignment neu = assigner.generateAssignment("test", CUR_A, {10, 11, 12, 13}, {}, -1);
    for (auto& kv : verifyPartitionsAndBuildReplicaCounts(CUR_A, neu, 1)) CHECK(kv.second == 2);
}

static void testDecommission() {  // TEST:84-122
    const Assignment cur = {{0, {10, 11}}, {1, {11, 12}}, {2, {12, 13}}, {3, {13, 10}}};
    kassign::KafkaTopicAssigner assigner;
    Assignment neu = assigner.generateAssignment("test", cur, {10, 11, 13}, {}, -1);
    auto counts = verifyPartitionsAndBuildReplicaCounts(cur, neu, 1);
    CHECK(counts.count(12) == 0);
    int servingTwo = 0, servingThree = 0;
    for (auto& kv : counts) {
        if (kv.second == 2) servingTwo++;
        else if (kv.second == 3) servingThree++;
        else CHECK(!"No broker should serve fewer than two or greater than 3 replicas");
    }
    CHECK(servingTwo == 1);
    CHECK(servingThree == 2);
}

static void testReplacement() {  // TEST:124-157
    kassign::KafkaTopicAssigner assigner;
    Assignment neu = assigner.generateAssignment("test", CUR_A, {10, 11, 13}, {}, -1);
    auto counts = verifyPartitionsAndBuildReplicaCounts(CUR_A, neu, 1);
    CHECK(counts.count(12) == 0);
    CHECK(neu.at(0) == CUR_A.at(0));  // TEST:143-144 — the reference's only exact pin
    auto has = [&](int p, int b) { return std::count(neu.at(p).begin(), neu.at(p).end(), b) > 0; };
    CHECK(has(1, 11) && (has(1, 10) || has(1, 13)));
    CHECK(has(2, 10) && (has(2, 11) || has(2, 13)));
    CHECK(has(3, 10) && (has(3, 11) || has(3, 13)));
}

static void testExceptionsCarryTheReferenceMessages() {  // KTA:58-60, 65-66, 67-69; KAS:183-184, 190-192
    kassign::KafkaTopicAssigner a;
    auto expect = [&](const char* what, auto fn) {
        try { fn(); CHECK(!"expected an exception"); }
        catch (const std::exception& e) { if (std::string(e.what()) != what) { std::fprintf(stderr, "got '%s' want '%s'\n", e.what(), what); ++failures; } }
    };
    expect("Topic t has partition 1 with unexpected replication factor 1", [&] { a.generateAssignment("t", {{0, {1, 2}}, {1, {1}}}, {1, 2, 3}, {}, -1); });
    expect("Topic t does not have a positive replication factor!", [&] { a.generateAssignment("t", {}, {1, 2, 3}, {}, -1); });
    expect("Topic t has a higher replication factor (3) than available brokers!", [&] { a.generateAssignment("t", {{0, {1, 2, 3}}}, {1, 2}, {}, -1); });
    expect("Partition 0 could not be fully assigned!", [&] { a.generateAssignment("t", {{0, {1, 2}}, {1, {2, 1}}}, {1, 2, 3}, {{1, "x"}, {2, "x"}, {3, "y"}}, 3); });
    expect("-2", [&] { a.generateAssignment("polygenelubricants", {{0, {1, 2, 3}}}, {1, 2, 3}, {}, -1); });
}

int main() {
    try {
        testRackAwareExpansion();
        testClusterExpansion();
        testDecommission();
        testReplacement();
        testExceptionsCarryTheReferenceMessages();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "unexpected exception: %s\n", e.what());
        return 2;
    }
    std::printf("%s (%d failure%s)\n", failures ? "FAILED" : "OK", failures, failures == 1 ? "" : "s");
    return failures ? 1 : 0;
}
