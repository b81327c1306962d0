// fast_oracle.cpp — OPTIMISED single-thread CPU solver of the same path (TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT).
//
// BASELINE.md "B1": what a careful CPU implementation of the reference algorithm costs when the Java collection
// machinery is replaced by flat arrays — the fair CPU yardstick next to the structure-faithful restatement
// (kafka_oracle.cpp, which mirrors TreeMap/TreeSet and therefore the reference's own cost profile). It is also a third,
// independently written restatement: tests require it to agree bit-for-bit with kafka_oracle.cpp and py_oracle.py.
//
// Follows KafkaAssignmentStrategy.java:40-369 / KafkaTopicAssigner.java:42-72 / KafkaAssignmentGenerator.java:172-184
// (same visit orders: sticky fill by (slot, partition ascending); orphans ascending with first-fit over the sorted broker
// ids rotated by |hashCode| % N; leader ordering by least counter[broker][slot] with ties to the rotated scan order).
// Only tests/ and bench.py's cpu legs may load this. The product library never links or calls it.
//
// Build: g++ -O2 -std=c++17 -shared -fPIC -o oracle/libfastoracle.so oracle/fast_oracle.cpp
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

extern "C" {

struct fast_status {
    int32_t code, topic_index, partition, a, b;  // same codes as include/kassign.h (1..5)
};

struct fast_ctx {
    std::vector<int32_t> ctr;  // [N][8] for the CURRENT broker table (tests reset between tables)
    int N = 0;
};

void* fast_ctx_create() { return new fast_ctx(); }
void fast_ctx_destroy(void* c) { delete (fast_ctx*)c; }
void fast_ctx_reset(void* c) { ((fast_ctx*)c)->ctr.assign(((fast_ctx*)c)->ctr.size(), 0); }

// Dense form only (what the BASELINE configs use): cur[T][P][RF] broker ids -> out[T][P][S], out_len[T][P].
// broker_id ascending; broker_rack dense rack index. Returns status code (0 ok).
int fast_solve_dense(void* cptr, int32_t T, const int32_t* topic_hash, int32_t P, int32_t RF, const int32_t* cur, int32_t N,
                     const int32_t* broker_id, const int32_t* broker_rack, int32_t desired_rf, int32_t S, int32_t* out_len,
                     int32_t* out, fast_status* st) {
    fast_ctx& ctx = *(fast_ctx*)cptr;
    if (ctx.N != N) { ctx.N = N; ctx.ctr.assign((size_t)N * 8, 0); }
    std::memset(st, 0, sizeof(*st));
    st->topic_index = -1;
    st->partition = -1;
    // id -> index: dense table when the id range is small, else binary search
    const int64_t lo = N ? broker_id[0] : 0, hi = N ? broker_id[N - 1] : -1;
    std::vector<int32_t> lut;
    if (N && hi - lo < (1 << 22)) {
        lut.assign((size_t)(hi - lo + 1), -1);
        for (int i = 0; i < N; ++i) lut[(size_t)(broker_id[i] - lo)] = i;
    }
    auto lookup = [&](int32_t id) -> int {
        if (!lut.empty()) return (id < lo || id > hi) ? -1 : lut[(size_t)(id - lo)];
        const int32_t* p = std::lower_bound(broker_id, broker_id + N, id);
        return (p != broker_id + N && *p == id) ? (int)(p - broker_id) : -1;
    };
    std::vector<int32_t> load(N), acc((size_t)P * S), cnt(P);
    for (int32_t t = 0; t < T; ++t) {
        int rf = desired_rf >= 0 ? desired_rf : (P > 0 ? RF : -1);
        auto fail = [&](int code, int part, int a, int b) {
            st->code = code; st->topic_index = t; st->partition = part; st->a = a; st->b = b;
            return code;
        };
        if (!(rf > 0)) return fail(2, -1, 0, 0);
        if (!(rf <= N)) return fail(3, -1, rf, 0);
        const int64_t tot = (int64_t)P * rf;
        const int cap = (int)((tot + N - 1) / N);
        std::fill(load.begin(), load.end(), 0);
        std::fill(cnt.begin(), cnt.end(), 0);
        const int32_t* c = cur + (size_t)t * P * RF;
        // sticky fill (KAS:101-131)
        for (int r = 0; r < RF; ++r)
            for (int p = 0; p < P; ++p) {
                const int idx = lookup(c[(size_t)p * RF + r]);
                if (idx < 0 || load[idx] >= cap) continue;
                const int rk = broker_rack[idx];
                int32_t* row = &acc[(size_t)p * S];
                bool clash = false;
                for (int i = 0; i < cnt[p]; ++i) clash |= broker_rack[row[i]] == rk;
                if (clash) continue;
                row[cnt[p]++] = idx;
                load[idx]++;
            }
        // rotated order + orphans (KAS:133-200)
        const int32_t h = topic_hash[t];
        int start;
        if (h == INT32_MIN) {
            const uint32_t r = 0x80000000u % (uint32_t)N;
            if (r) return fail(5, -1, -(int)r, N);
            start = 0;
        } else {
            start = (int)((uint32_t)(h < 0 ? -h : h) % (uint32_t)N);
        }
        const int i0 = (N - start) % N;
        int head = 0;
        for (int p = 0; p < P; ++p) {
            int rem = rf - cnt[p];
            if (rem <= 0) continue;
            int32_t* row = &acc[(size_t)p * S];
            while (head < N) {  // positions < head hold full nodes for good
                int idx = i0 + head; if (idx >= N) idx -= N;
                if (load[idx] < cap) break;
                ++head;
            }
            for (int j = head; j < N && rem > 0; ++j) {
                int idx = i0 + j; if (idx >= N) idx -= N;
                if (load[idx] >= cap) continue;
                const int rk = broker_rack[idx];
                bool clash = false;
                for (int i = 0; i < cnt[p]; ++i) clash |= broker_rack[row[i]] == rk;
                if (clash) continue;
                row[cnt[p]++] = idx;
                load[idx]++;
                --rem;
            }
            if (rem > 0) return fail(4, p, 0, 0);
        }
        // leader ordering (KAS:202-302)
        const uint32_t habs = h == INT32_MIN ? 0x80000000u : (uint32_t)(h < 0 ? -h : h);
        for (int p = 0; p < P; ++p) {
            int32_t* row = &acc[(size_t)p * S];
            const int k0 = cnt[p];
            std::sort(row, row + k0);
            int rem[8];
            for (int i = 0; i < k0; ++i) rem[i] = row[i];
            int32_t* o = out + ((size_t)t * P + p) * S;
            for (int r = 0; r < k0; ++r) {
                const int k = k0 - r;
                int s;
                if (h == INT32_MIN) {
                    const uint32_t m = 0x80000000u % (uint32_t)k;
                    if (m) return fail(5, -1, -(int)m, k);
                    s = 0;
                } else {
                    s = (int)(habs % (uint32_t)k);
                }
                int best = -1, bestc = 0;
                for (int j = 0; j < k; ++j) {  // order[j] = rem[(j - s) mod k]
                    int m = j - s; if (m < 0) m += k;
                    const int cv = ctx.ctr[(size_t)rem[m] * 8 + r];
                    if (best < 0 || cv < bestc) { best = m; bestc = cv; }
                }
                o[r] = rem[best];
                for (int i = best; i + 1 < k; ++i) rem[i] = rem[i + 1];
            }
            for (int r = 0; r < k0; ++r) { ctx.ctr[(size_t)o[r] * 8 + r]++; o[r] = broker_id[o[r]]; }
            for (int r = k0; r < S; ++r) o[r] = -1;
            out_len[(size_t)t * P + p] = k0;
        }
    }
    return 0;
}

}  // extern "C"
