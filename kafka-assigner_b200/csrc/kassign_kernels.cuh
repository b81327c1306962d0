// kassign_kernels.cuh — hand-written sm_100a kernels for the kafka-assigner hot path.
//
// Reference being replaced (KAS = KafkaAssignmentStrategy.java, KTA = KafkaTopicAssigner.java):
//   kernel A  ka_sticky_spread_kernel   KTA:49-69 (RF inference/validation), KAS:65-71 (capacity),
//                                       KAS:73-99 (node/rack table), KAS:101-131 (sticky fill),
//                                       KAS:133-160 (orphans), KAS:162-200 (rotated first-fit spread)
//   kernels T ka_ticket_*               no reference counterpart: they number, per broker, the partitions
//                                       that contain it in global (topic, partition) order so that the
//                                       serial chain of KAS:202-239 can run as an exact dataflow
//   kernel B  ka_leader_order_kernel    KAS:202-239 + PreferenceListOrderTracker KAS:244-302 against the
//                                       cross-topic Context.counter (KAS:360-369, KTA:19-23)
//
// Everything is integer indexing: no tensor cores. Broker table staged into shared memory with one TMA
// bulk copy per CTA (cp.async.bulk + mbarrier), per-topic state (broker loads, replica slab) lives in
// shared memory, decisions that depend on visit order are taken with warp ballots / match / shuffles in
// the reference's order — never by an atomics race.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define KA_MAX_SLOTS 8      // max replicas per partition row (out_stride)
#define KA_DEAD 0xFFFFu     // "broker not in the live set" marker in 16-bit index space
#define KA_FULL 0xFFFFFFFFu

// Error codes (mirror include/kassign.h)
#define KA_E_RF_MISMATCH 1
#define KA_E_RF_NOT_POSITIVE 2
#define KA_E_RF_GT_BROKERS 3
#define KA_E_UNASSIGNABLE 4
#define KA_E_HASH_INDEX 5
#define KA_E_INTERNAL_SPIN -5

enum { KA_LUT_SMEM = 0, KA_LUT_GLOBAL = 1, KA_LUT_BSEARCH = 2 };

struct KaSolveParams {
    // problem
    int T;
    int topic_base;             // index of this block's first topic in the whole run (status reporting)
    const int32_t* topic_hash;  // [T]
    const int64_t* part_off;    // [T+1] or nullptr (dense: P partitions per topic)
    int P;
    const int64_t* rep_off;     // [Q+1] or nullptr (dense: RF replicas per row)
    int RF;
    const int32_t* cur;         // current replica lists (broker IDs)
    int desired_rf;
    int S;                      // row stride of slab / set / out
    int Pmax;                   // max partitions of any topic (smem sizing)
    // broker table
    int N;
    const uint16_t* blob;       // global: rack16[Npad] | lut16[range_pad] (16B aligned, multiple of 16B)
    int blob_bytes;             // bytes staged into smem (rack, plus lut when lut_mode == SMEM)
    int lut_off;                // element offset (uint16) of lut16 inside blob
    int R;                      // number of distinct racks (compact ids 0..R-1 in rack16)
    int rackptr;                // 1: spread phase uses per-rack first-free pointers (R small); 0: window scan
    int roff_off;               // element offset of rack_off16[R+1] inside blob
    int memb_off;               // element offset of members16[N] (sorted indices grouped by rack, ascending) inside blob
    int rp_bytes;               // per-warp bytes of each of the three rack-pointer arrays
    int lut_mode;
    int min_id;
    uint32_t range;
    const uint16_t* glut;       // global lut16 (lut_mode == GLOBAL)
    const int32_t* broker_id;   // [N] ascending (global)
    // outputs
    int32_t* set;               // [Q*S] accepted broker indices, ascending per row, -1 padded
    uint32_t* meta;             // [Q] len | rotation bits
    int4* tstatus;              // [T] per-topic error record (written only on error)
    unsigned* err_topic;        // unsigned atomicMin of the failing topic index (init = 0xFFFFFFFF)
};

// ------------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ka_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ka_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ka_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ka_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void ka_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void ka_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ka_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ka_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(ka_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared (non-tensor form). bytes % 16 == 0, both addresses 16B aligned.
__device__ __forceinline__ void ka_tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     ka_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(ka_smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ int4 ka_lds_v4_volatile(const int* p) {
    int4 v;
    asm volatile("ld.volatile.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ka_smem_u32(p)));
    return v;
}
__device__ __forceinline__ void ka_sts_volatile(int* p, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(ka_smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int4 ka_ldg_stream_v4(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t ka_lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Rotation bits of one topic: (|hash| % k) for k = 2..8 packed above the 4-bit length (KAS:190 applied
// to the remaining-set sizes of KAS:267). Layout: len[0:4) k2[4] k3[5:7) k4[7:9) k5[9:12) k6[12:15) k7[15:18) k8[18:21)
__device__ __forceinline__ uint32_t ka_rot_bits(uint32_t habs) {
    return ((habs % 2u) << 4) | ((habs % 3u) << 5) | ((habs % 4u) << 7) | ((habs % 5u) << 9) | ((habs % 6u) << 12) |
           ((habs % 7u) << 15) | ((habs % 8u) << 18);
}
template <int RS>
__device__ __forceinline__ int ka_rot_of(uint32_t meta, int k) {
    // k in [1,RS]; select chain instead of a table so nothing lands in local memory
    int s = 0;
    if (k == 2) s = (meta >> 4) & 1u;
    if (k == 3) s = (meta >> 5) & 3u;
    if (k == 4) s = (meta >> 7) & 3u;
    if (RS > 4) {
        if (k == 5) s = (meta >> 9) & 7u;
        if (k == 6) s = (meta >> 12) & 7u;
        if (k == 7) s = (meta >> 15) & 7u;
        if (k == 8) s = (meta >> 18) & 7u;
    }
    return s;
}

// ------------------------------------------------------------------------------------------------
// Kernel A: sticky fill + orphan spread, one topic per warp, persistent CTAs.
// ------------------------------------------------------------------------------------------------
struct KaTab {               // CTA-shared views into the staged broker table
    const uint16_t* rack;    // [N] compact rack id of each broker (sorted-index order)
    const uint16_t* lut;     // [range] (lut_mode == SMEM)
    const uint16_t* roff;    // [R+1] member ranges per rack
    const uint16_t* memb;    // [N] sorted indices grouped by rack, ascending inside a rack
};

__device__ __forceinline__ uint32_t ka_lookup(int id, const KaTab& tab, const KaSolveParams& p) {
    if (p.lut_mode == KA_LUT_SMEM) {
        uint32_t off = (uint32_t)id - (uint32_t)p.min_id;
        return off < p.range ? (uint32_t)tab.lut[off] : KA_DEAD;
    } else if (p.lut_mode == KA_LUT_GLOBAL) {
        uint32_t off = (uint32_t)id - (uint32_t)p.min_id;
        return off < p.range ? (uint32_t)__ldg(&p.glut[off]) : KA_DEAD;
    } else {
        int lo = 0, hi = p.N - 1;
        while (lo <= hi) {
            int mid = (lo + hi) >> 1;
            int v = __ldg(&p.broker_id[mid]);
            if (v == id) return (uint32_t)mid;
            if (v < id) lo = mid + 1; else hi = mid - 1;
        }
        return KA_DEAD;
    }
}

template <typename LoadT>
__device__ void ka_solve_topic(const KaSolveParams& p, const KaTab& tab, int t, LoadT* load, uint16_t* slab, uint8_t* cnt,
                               uint16_t* rpos, uint16_t* rst, uint16_t* rkk) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = ka_lanemask_lt();
    const int S = p.S;
    const int N = p.N;

    int64_t g0;
    int P;
    if (p.part_off) {
        g0 = p.part_off[t];
        P = (int)(p.part_off[t + 1] - g0);
    } else {
        P = p.P;
        g0 = (int64_t)t * P;
    }

    int err = 0, errp = -1, erra = 0, errb = 0;

    // ---- KTA:49-61 replication-factor inference / validation ------------------------------------
    int rf = p.desired_rf;
    int maxlen = 0;
    if (!p.rep_off) {
        maxlen = P > 0 ? p.RF : 0;
        if (rf < 0 && P > 0) rf = p.RF;
    } else {
        const int64_t* ro = p.rep_off + g0;
        int first = P > 0 ? (int)(ro[1] - ro[0]) : -1;
        int mism = 0x7FFFFFFF;
        for (int pp = lane; pp < P; pp += 32) {
            int sz = (int)(ro[pp + 1] - ro[pp]);
            maxlen = max(maxlen, sz);
            if (sz != first) mism = min(mism, pp);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            maxlen = max(maxlen, __shfl_xor_sync(KA_FULL, maxlen, o));
            mism = min(mism, __shfl_xor_sync(KA_FULL, mism, o));
        }
        if (rf < 0 && P > 0) {
            rf = first;
            if (mism != 0x7FFFFFFF) {  // first entry (ascending) whose size differs (KTA:57-60)
                err = KA_E_RF_MISMATCH;
                errp = mism;
                erra = (int)(ro[mism + 1] - ro[mism]);
            }
        }
    }
    if (!err && !(rf > 0)) err = KA_E_RF_NOT_POSITIVE;          // KTA:65-66
    if (!err && !(rf <= N)) { err = KA_E_RF_GT_BROKERS; erra = rf; }  // KTA:67-69

    const int32_t h = p.topic_hash[t];
    const bool hmin = (h == (int32_t)0x80000000);
    const uint32_t habs = hmin ? 0x80000000u : (uint32_t)(h < 0 ? -h : h);

    if (!err) {
        // ---- KAS:65-71 capacity: (int)ceil((double)(P*rf) / N) == integer ceil-div for int32 operands
        const int64_t tot = (int64_t)(int32_t)((uint32_t)P * (uint32_t)rf);
        const int cap = tot > 0 ? (int)((tot + N - 1) / N) : 0;

        // ---- KAS:73-99 node table: loads start empty -------------------------------------------
        {
            uint32_t* lw = reinterpret_cast<uint32_t*>(load);
            const int words = (N * (int)sizeof(LoadT) + 3) >> 2;
            for (int i = lane; i < words; i += 32) lw[i] = 0u;
            uint32_t* cw = reinterpret_cast<uint32_t*>(cnt);
            for (int i = lane; i < ((P + 3) >> 2); i += 32) cw[i] = 0u;
        }

        // ---- stage the topic's current assignment as 16-bit broker indices ----------------------
        if (!p.rep_off) {
            const int RF = p.RF;
            const int32_t* src = p.cur + g0 * RF;
            const int n = P * RF;
            if (RF == S) {
                // coalesced, 128-bit vectorised when the slab is 16B aligned
                const bool al = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((n & 3) == 0);
                if (al) {
                    const int4* s4 = reinterpret_cast<const int4*>(src);
                    for (int e = lane; e < (n >> 2); e += 32) {
                        int4 v = ka_ldg_stream_v4(s4 + e);
                        uint32_t a = ka_lookup(v.x, tab, p), b = ka_lookup(v.y, tab, p);
                        uint32_t c = ka_lookup(v.z, tab, p), d = ka_lookup(v.w, tab, p);
                        uint2 pk = make_uint2(a | (b << 16), c | (d << 16));
                        *reinterpret_cast<uint2*>(slab + 4 * e) = pk;
                    }
                } else {
                    for (int e = lane; e < n; e += 32) slab[e] = (uint16_t)ka_lookup(__ldg(src + e), tab, p);
                }
            } else {
                for (int e = lane; e < n; e += 32) {
                    int pp = e / RF, r = e - pp * RF;
                    slab[pp * S + r] = (uint16_t)ka_lookup(__ldg(src + e), tab, p);
                }
                for (int e = lane; e < P * (S - RF); e += 32) {
                    int pp = e / (S - RF), r = RF + (e - pp * (S - RF));
                    slab[pp * S + r] = (uint16_t)KA_DEAD;
                }
            }
        } else {
            const int64_t* ro = p.rep_off + g0;
            for (int pp = lane; pp < P; pp += 32) {
                int64_t off = ro[pp];
                int sz = (int)(ro[pp + 1] - off);
                for (int r = 0; r < S; ++r)
                    slab[pp * S + r] = r < sz ? (uint16_t)ka_lookup(__ldg(p.cur + off + r), tab, p) : (uint16_t)KA_DEAD;
            }
        }
        __syncwarp();

        // ---- KAS:101-131 sticky fill: visit order (slot r, partition ascending) ------------------
        for (int r = 0; r < maxlen; ++r) {
            for (int c0 = 0; c0 < P; c0 += 32) {
                const int pp = c0 + lane;
                const bool valid = pp < P;
                uint32_t idx = valid ? (uint32_t)slab[pp * S + r] : KA_DEAD;
                const int k = valid ? (int)cnt[pp] : 0;
                bool feas = idx != KA_DEAD;
                if (feas) {
                    const uint32_t rk = tab.rack[idx];
                    for (int i = 0; i < k; ++i)  // rack exclusivity (also covers "node already has p")
                        if (tab.rack[slab[pp * S + i]] == rk) feas = false;
                }
                const uint32_t fm = __ballot_sync(KA_FULL, feas);
                int rank = 0, gsz = 0, l = 0;
                if (feas) {
                    // rank among this pass's candidates of the same broker, ascending partition
                    const uint32_t m = __match_any_sync(fm, idx);
                    rank = __popc(m & lt);
                    gsz = __popc(m);
                    l = (int)load[idx];
                }
                __syncwarp();  // every candidate has read the broker's load before anyone updates it
                if (feas) {
                    if (l + rank < cap) {
                        slab[pp * S + k] = (uint16_t)idx;  // in-place compaction: k <= r
                        cnt[pp] = (uint8_t)(k + 1);
                    }
                    if (rank == 0) load[idx] = (LoadT)(l + min(gsz, max(cap - l, 0)));
                }
                __syncwarp();
            }
        }

        // ---- KAS:188-200 rotated processing order ------------------------------------------------
        uint32_t start = 0;
        if (!hmin) {
            start = habs % (uint32_t)N;
        } else {
            uint32_t rmd = 0x80000000u % (uint32_t)N;  // Math.abs(MIN_VALUE) % N == -(2^31 % N)
            if (rmd != 0) { err = KA_E_HASH_INDEX; erra = -(int)rmd; errb = N; }
        }
        const int i0 = (int)(((uint32_t)N - start) % (uint32_t)N);  // sorted index at order position 0

        // ---- KAS:133-186 orphans, ascending partition; first-fit from position 0 each time -------
        if (p.rackptr) {
            // Exact reformulation of the walk: a position is acceptable iff its node is not full and its rack does
            // not hold the partition yet, and loads / used racks only grow — so "first acceptable position from
            // j = 0" == min over the racks not yet used of that rack's first non-full member (in rotated order).
            // Each rack keeps a monotone pointer into its member list; placing a replica is one warp min-reduction.
            const int R = p.R;
            for (int r = lane; r < R; r += 32) {
                const int o = tab.roff[r], sz = (int)tab.roff[r + 1] - o;
                int lo = 0, hi = sz;
                while (lo < hi) {  // first member with sorted index >= i0 starts the rack's rotated order
                    const int mid = (lo + hi) >> 1;
                    if ((int)tab.memb[o + mid] < i0) lo = mid + 1; else hi = mid;
                }
                const int st = (lo == sz) ? 0 : lo;
                int kk = 0, pos = 0xFFFF;
                while (kk < sz) {
                    const int m = tab.memb[o + (st + kk >= sz ? st + kk - sz : st + kk)];
                    if ((int)load[m] < cap) { pos = m - i0; if (pos < 0) pos += N; break; }
                    ++kk;
                }
                rst[r] = (uint16_t)st; rkk[r] = (uint16_t)kk; rpos[r] = (uint16_t)pos;
            }
            __syncwarp();
            for (int c0 = 0; c0 < P && !err; c0 += 32) {
                const int pp0 = c0 + lane;
                const int need = pp0 < P ? rf - (int)cnt[pp0] : 0;
                uint32_t ob = __ballot_sync(KA_FULL, need > 0);
                while (ob && !err) {
                    const int src = __ffs(ob) - 1;
                    ob &= ob - 1;
                    const int pp = c0 + src;
                    int rem = __shfl_sync(KA_FULL, need, src);
                    int k = (int)cnt[pp];
                    uint32_t ur[KA_MAX_SLOTS];  // racks already holding this partition (warp-uniform)
#pragma unroll
                    for (int i = 0; i < KA_MAX_SLOTS; ++i) ur[i] = i < k ? (uint32_t)tab.rack[slab[pp * S + i]] : 0xFFFFFFFFu;
                    while (rem > 0) {
                        uint32_t best = 0xFFFFFFFFu;
                        for (int r = lane; r < R; r += 32) {
                            bool used = false;
#pragma unroll
                            for (int i = 0; i < KA_MAX_SLOTS; ++i) used = used || (ur[i] == (uint32_t)r);
                            const uint32_t cnd = used ? 0xFFFFFFFFu : (((uint32_t)rpos[r] << 16) | (uint32_t)r);
                            best = min(best, cnd);
                        }
                        best = __reduce_min_sync(KA_FULL, best);
                        if ((best >> 16) == 0xFFFFu) break;  // no rack can take it: stranded (KAS:183-184)
                        const int pos = (int)(best >> 16), r = (int)(best & 0xFFFFu);
                        int idx = i0 + pos;
                        if (idx >= N) idx -= N;
                        const int nl = (int)load[idx] + 1;
                        __syncwarp();
                        if (lane == 0) {
                            load[idx] = (LoadT)nl;
                            slab[pp * S + k] = (uint16_t)idx;
                        }
#pragma unroll
                        for (int i = 0; i < KA_MAX_SLOTS; ++i)
                            if (i == k) ur[i] = (uint32_t)r;
                        ++k;
                        --rem;
                        if (nl >= cap) {  // the rack's first-free member just filled up: advance its pointer
                            const int o = tab.roff[r], sz = (int)tab.roff[r + 1] - o, st = rst[r];
                            int kk = (int)rkk[r] + 1, np = 0xFFFF;
                            while (kk < sz) {
                                const int m = tab.memb[o + (st + kk >= sz ? st + kk - sz : st + kk)];
                                if ((int)load[m] < cap) { np = m - i0; if (np < 0) np += N; break; }
                                ++kk;
                            }
                            if (lane == 0) { rkk[r] = (uint16_t)kk; rpos[r] = (uint16_t)np; }
                        }
                        __syncwarp();
                    }
                    if (lane == 0) cnt[pp] = (uint8_t)k;
                    __syncwarp();
                    if (rem > 0 && !err) { err = KA_E_UNASSIGNABLE; errp = pp; }  // KAS:183-184
                }
            }
        } else {
        int head = 0;  // all order positions < head hold full nodes (loads never decrease)
        for (int c0 = 0; c0 < P && !err; c0 += 32) {
            const int pp0 = c0 + lane;
            const int need = pp0 < P ? rf - (int)cnt[pp0] : 0;
            uint32_t ob = __ballot_sync(KA_FULL, need > 0);
            while (ob && !err) {
                const int src = __ffs(ob) - 1;
                ob &= ob - 1;
                const int pp = c0 + src;
                int rem = __shfl_sync(KA_FULL, need, src);
                int k = (int)cnt[pp];
                uint32_t ur[KA_MAX_SLOTS];  // racks already holding this partition (warp-uniform)
#pragma unroll
                for (int i = 0; i < KA_MAX_SLOTS; ++i) ur[i] = i < k ? (uint32_t)tab.rack[slab[pp * S + i]] : 0xFFFFFFFFu;
                bool adv = true;
                for (int j = head; j < N && rem > 0; j += 32) {
                    const int pos = j + lane;
                    int idx = i0 + pos;
                    if (idx >= N) idx -= N;
                    bool nonfull = false;
                    uint32_t rk = 0xFFFFFFFEu;
                    if (pos < N) {
                        nonfull = (int)load[idx] < cap;
                        rk = tab.rack[idx];
                    }
                    if (adv) {
                        const uint32_t nb = __ballot_sync(KA_FULL, nonfull);
                        if (nb == 0) head = min(j + 32, N);
                        else { head = j + __ffs(nb) - 1; adv = false; }
                    }
                    bool feas = nonfull;
#pragma unroll
                    for (int i = 0; i < KA_MAX_SLOTS; ++i) feas = feas && (ur[i] != rk);
                    uint32_t fb = __ballot_sync(KA_FULL, feas);
                    while (fb && rem > 0) {
                        const int f = __ffs(fb) - 1;
                        const int cidx = __shfl_sync(KA_FULL, idx, f);
                        const uint32_t crk = __shfl_sync(KA_FULL, rk, f);
                        if (lane == f) {
                            load[idx] = (LoadT)((int)load[idx] + 1);
                            slab[pp * S + k] = (uint16_t)cidx;
                        }
#pragma unroll
                        for (int i = 0; i < KA_MAX_SLOTS; ++i)
                            if (i == k) ur[i] = crk;
                        ++k;
                        --rem;
                        fb &= ~((2u << f) - 1u);                          // only later positions
                        fb &= ~__ballot_sync(KA_FULL, rk == crk);        // that rack is now taken
                    }
                    __syncwarp();
                }
                if (lane == 0) cnt[pp] = (uint8_t)k;
                __syncwarp();
                if (rem > 0 && !err) { err = KA_E_UNASSIGNABLE; errp = pp; }  // KAS:183-184
            }
        }

        }

        // ---- per-partition finalisation: ascending broker order (KAS:205-214) --------------------
        if (!err) {
            int firstbad = 0x7FFFFFFF, badk = 0;
            for (int pp = lane; pp < P; pp += 32) {
                const int k = (int)cnt[pp];
                uint16_t* row = slab + pp * S;
                for (int i = 1; i < k; ++i) {  // insertion sort, k <= 8
                    uint16_t v = row[i];
                    int j = i - 1;
                    while (j >= 0 && row[j] > v) { row[j + 1] = row[j]; --j; }
                    row[j + 1] = v;
                }
                if (hmin && k >= 3 && pp < firstbad) { firstbad = pp; badk = k; }
            }
            if (hmin) {  // KAS:267 with Math.abs(MIN_VALUE): first remaining-set size that does not divide 2^31
                int fb2 = firstbad;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) fb2 = min(fb2, __shfl_xor_sync(KA_FULL, fb2, o));
                if (fb2 != 0x7FFFFFFF) {
                    const uint32_t bm = __ballot_sync(KA_FULL, firstbad == fb2);
                    const int kk = __shfl_sync(KA_FULL, badk, __ffs(bm) - 1);
                    const int kfail = (kk & (kk - 1)) ? kk : kk - 1;
                    err = KA_E_HASH_INDEX;
                    errp = -1;
                    erra = -(int)(0x80000000u % (uint32_t)kfail);
                    errb = kfail;
                }
            }
            __syncwarp();
        }
    }

    // ---- emit ------------------------------------------------------------------------------------
    const uint32_t rot = (err || hmin) ? 0u : ka_rot_bits(habs);
    int32_t* oset = p.set + g0 * S;
    if (!err) {
        for (int e = lane; e < P * S; e += 32) {
            const int pp = e / S, i = e - pp * S;
            oset[e] = i < (int)cnt[pp] ? (int32_t)slab[e] : -1;
        }
        for (int pp = lane; pp < P; pp += 32) p.meta[g0 + pp] = (uint32_t)cnt[pp] | rot;
    } else {
        for (int e = lane; e < P * S; e += 32) oset[e] = -1;
        for (int pp = lane; pp < P; pp += 32) p.meta[g0 + pp] = 0u;
        if (lane == 0) {
            p.tstatus[p.topic_base + t] = make_int4(err, errp, erra, errb);
            atomicMin(p.err_topic, (unsigned)(p.topic_base + t));
        }
    }
    __syncwarp();
}

template <typename LoadT>
__global__ void __launch_bounds__(512) ka_sticky_spread_kernel(const KaSolveParams p, int load_bytes, int slab_bytes, int cnt_bytes) {
    extern __shared__ __align__(16) unsigned char ka_smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(ka_smem);
    unsigned char* blob = ka_smem + 16;
    unsigned char* warp_base = blob + p.blob_bytes;

    // TMA bulk-stage the broker table (rack indices + id->index LUT) once per CTA.
    if (threadIdx.x == 0) {
        ka_mbar_init(bar, 1);
        ka_fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0 && p.blob_bytes > 0) {
        ka_mbar_expect_tx(bar, (uint32_t)p.blob_bytes);
        ka_tma_bulk_g2s(blob, p.blob, (uint32_t)p.blob_bytes, bar);
    }
    if (p.blob_bytes > 0) ka_mbar_wait(bar, 0);

    KaTab tab;
    tab.rack = reinterpret_cast<const uint16_t*>(blob);
    tab.lut = reinterpret_cast<const uint16_t*>(blob) + p.lut_off;
    tab.roff = reinterpret_cast<const uint16_t*>(blob) + p.roff_off;
    tab.memb = reinterpret_cast<const uint16_t*>(blob) + p.memb_off;

    const int warp = threadIdx.x >> 5;
    const int nwarp = blockDim.x >> 5;
    const int per_warp = load_bytes + slab_bytes + cnt_bytes + 3 * p.rp_bytes;
    unsigned char* mine = warp_base + (size_t)warp * per_warp;
    LoadT* load = reinterpret_cast<LoadT*>(mine);
    uint16_t* slab = reinterpret_cast<uint16_t*>(mine + load_bytes);
    uint8_t* cnt = reinterpret_cast<uint8_t*>(mine + load_bytes + slab_bytes);
    uint16_t* rpos = reinterpret_cast<uint16_t*>(mine + load_bytes + slab_bytes + cnt_bytes);
    uint16_t* rst = reinterpret_cast<uint16_t*>(mine + load_bytes + slab_bytes + cnt_bytes + p.rp_bytes);
    uint16_t* rkk = reinterpret_cast<uint16_t*>(mine + load_bytes + slab_bytes + cnt_bytes + 2 * p.rp_bytes);

    const int total_warps = gridDim.x * nwarp;
    for (int t = blockIdx.x * nwarp + warp; t < p.T; t += total_warps) ka_solve_topic<LoadT>(p, tab, t, load, slab, cnt, rpos, rst, rkk);
}

// ------------------------------------------------------------------------------------------------
// Kernels T: per-broker occurrence numbering ("tickets") in global partition order.
//   ticket(q, b) = base(b) + #{ q' < q : b in set(q') },  base(b) = sum of b's counters at entry.
// Kernel B below may order partition q as soon as, for each of its brokers, the broker's counter row
// sums to the ticket — i.e. every earlier partition on that broker has committed its increment.
// Chunks are contiguous ranges of L partitions (L % 32 == 0), one warp per chunk.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) ka_ticket_hist_kernel(const int32_t* __restrict__ set, int64_t Q, int S, int N, int64_t L,
                                                              int num_chunks, int32_t* __restrict__ hist) {
    extern __shared__ __align__(16) unsigned char ka_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    int* h = reinterpret_cast<int*>(ka_smem) + (size_t)warp * N;
    for (int c = blockIdx.x * nwarp + warp; c < num_chunks; c += gridDim.x * nwarp) {
        for (int i = lane; i < N; i += 32) h[i] = 0;
        __syncwarp();
        const int64_t e0 = (int64_t)c * L * S;
        const int64_t e1 = min((int64_t)(c + 1) * L, Q) * S;
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const int idx = set[e];
            if (idx >= 0) atomicAdd(&h[idx], 1);
        }
        __syncwarp();
        for (int i = lane; i < N; i += 32) hist[(size_t)c * N + i] = h[i];
        __syncwarp();
    }
}

// seed[b] = sum of broker b's counter row = number of partitions that have committed on b so far (every commit adds
// exactly one to exactly one slot of each of its brokers): the ticket base of the first block of a solve.
__global__ void ka_seed_init_kernel(const int32_t* __restrict__ ctr8, int RS, int N, int32_t* __restrict__ seed) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= N) return;
    int run = 0;
    for (int r = 0; r < RS; ++r) run += ctr8[b * KA_MAX_SLOTS + r];
    seed[b] = run;
}

// Exclusive scan over chunks, per broker, seeded with seed[b]; leaves seed[b] + (block total of b) in seed[b] so the
// next block of a pipelined solve continues the numbering without reading live counters. One WARP per broker: lane l
// owns chunks l, l+32, ...; every load of the column is in flight at once, then a shuffle scan per 32-chunk tile carries
// the running total (num_chunks <= KA_SCAN_MAX_CHUNKS).
#define KA_SCAN_MAX_CHUNKS 2048
__global__ void __launch_bounds__(256) ka_ticket_scan_kernel(int32_t* hist, int num_chunks, int N, int32_t* seed) {
    const int lane = threadIdx.x & 31;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= N) return;
    int run = seed[b];
    for (int c0 = 0; c0 < num_chunks; c0 += 256) {  // 8 tiles of 32 chunks per round: 8 independent loads per lane
        int v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int c = c0 + t * 32 + lane;
            v[t] = c < num_chunks ? hist[(size_t)c * N + b] : 0;
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            int x = v[t];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(KA_FULL, x, o);
                if (lane >= o) x += y;
            }
            const int c = c0 + t * 32 + lane;
            if (c < num_chunks) hist[(size_t)c * N + b] = run + x - v[t];  // exclusive
            run += __shfl_sync(KA_FULL, x, 31);
        }
    }
    if (lane == 0) seed[b] = run;
}

// When tick4/idx01 are given (rows of <= 3 replicas) the pass also emits the packed per-partition record kernel B
// consumes with one 128-bit and one 32-bit load:  tick4[q] = {t0, t1, t2, idx2 | (meta & 0xFFFF) << 16},
// idx01[q] = idx0 | idx1 << 16.
__global__ void __launch_bounds__(1024) ka_ticket_rank_kernel(const int32_t* __restrict__ set, int64_t Q, int S, int N, int64_t L,
                                                              int num_chunks, const int32_t* __restrict__ base,
                                                              int32_t* __restrict__ ticket, const uint32_t* __restrict__ meta,
                                                              int4* __restrict__ tick4, uint32_t* __restrict__ idx01) {
    extern __shared__ __align__(16) unsigned char ka_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const uint32_t lt = ka_lanemask_lt();
    int* issued = reinterpret_cast<int*>(ka_smem) + (size_t)warp * 2 * N;
    uint32_t* owner = reinterpret_cast<uint32_t*>(issued + N);
    for (int c = blockIdx.x * nwarp + warp; c < num_chunks; c += gridDim.x * nwarp) {
        for (int i = lane; i < N; i += 32) { issued[i] = base[(size_t)c * N + i]; owner[i] = 0u; }
        __syncwarp();
        const int64_t q0 = (int64_t)c * L, q1 = min((int64_t)(c + 1) * L, Q);
        for (int64_t qb = q0; qb < q1; qb += 32) {
            const int64_t q = qb + lane;
            const bool valid = q < q1;
            const int32_t* row = set + q * S;
            int idx[KA_MAX_SLOTS];
#pragma unroll
            for (int i = 0; i < KA_MAX_SLOTS; ++i) idx[i] = (valid && i < S) ? row[i] : -1;
#pragma unroll
            for (int i = 0; i < KA_MAX_SLOTS; ++i)
                if (idx[i] >= 0) atomicOr(&owner[idx[i]], 1u << lane);
            __syncwarp();
            uint32_t own[KA_MAX_SLOTS];
            int tkv[KA_MAX_SLOTS];
#pragma unroll
            for (int i = 0; i < KA_MAX_SLOTS; ++i) {
                own[i] = 0u;
                tkv[i] = 0;
                if (idx[i] >= 0) {
                    own[i] = owner[idx[i]];
                    tkv[i] = issued[idx[i]] + __popc(own[i] & lt);
                }
                if (!tick4 && valid && i < S) ticket[q * S + i] = tkv[i];
            }
            if (tick4 && valid) {
                const uint32_t m = meta[q];
                tick4[q] = make_int4(tkv[0], tkv[1], tkv[2], (int)(((uint32_t)max(idx[2], 0) & 0xFFFFu) | ((m & 0xFFFFu) << 16)));
                idx01[q] = ((uint32_t)max(idx[0], 0) & 0xFFFFu) | (((uint32_t)max(idx[1], 0) & 0xFFFFu) << 16);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < KA_MAX_SLOTS; ++i)
                if (idx[i] >= 0 && (own[i] & lt) == 0u) {  // lowest partition holding this broker in the window
                    issued[idx[i]] += __popc(own[i]);
                    owner[idx[i]] = 0u;
                }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Kernel B: leader-preference ordering (KAS:202-239) as an exact dataflow over the counter table.
// One CTA; counters in shared memory, row-major [N][RS]. Each lane owns one partition of a 32-partition
// window; a partition commits when all of its brokers' rows have reached its tickets. Windows are taken
// in global order (warp w: windows w, w+W, ...), so the lowest uncommitted partition is always in flight
// and always ready: the loop cannot deadlock.
// ------------------------------------------------------------------------------------------------
struct KaOrderParams {
    int64_t Q;
    int S;
    int N;
    const int32_t* set;       // [Q*S]
    const int32_t* ticket;    // [Q*S]
    const uint32_t* meta;     // [Q]
    const int32_t* broker_id; // [N]
    int32_t* ctr8;            // [N*8] global counters (in/out)
    int32_t* out;             // [Q*S] broker ids, leader first
    int32_t* out_len;         // [Q] or nullptr
    int* err_flag;            // set to KA_E_INTERNAL_SPIN if a window spins beyond the guard
    const int4* tick4;        // order3: packed records from ka_ticket_rank_kernel
    const uint32_t* idx01;
    uint8_t* pcode;           // order3: chosen positions (b0 | b1 << 2) per partition, consumed by ka_emit_kernel
};

// ------------------------------------------------------------------------------------------------
// Kernel B, specialised for rows of 4 slots (every partition list <= 4 replicas: all BASELINE configs).
// Written with explicit scalars — no arrays — so that every selection compiles to SEL/predication and
// nothing is spilled to (or dynamically indexed in) local memory.
// ------------------------------------------------------------------------------------------------
#define KA_BIG 0x7FFFFFFF

__device__ __forceinline__ int ka_sel4(int k, int a0, int a1, int a2, int a3) {
    return k == 0 ? a0 : (k == 1 ? a1 : (k == 2 ? a2 : a3));
}

// One selection round of KAS:263-278 over the remaining positions `rem` (bitmask over the ascending
// broker list), counters cv0..cv3 = counter[broker at pos][slot of this round], rotation s, k = popc(rem).
__device__ __forceinline__ int ka_pick4(uint32_t rem, int k, int s, int cv0, int cv1, int cv2, int cv3) {
    long long best = 0x7FFFFFFFFFFFFFFFLL;
    int bpos = 0;
    {
        int j = s; if (j >= k) j -= k;
        const long long key = (long long)cv0 * 8 + j;
        if ((rem & 1u) && key < best) { best = key; bpos = 0; }
    }
    {
        int j = __popc(rem & 1u) + s; if (j >= k) j -= k;
        const long long key = (long long)cv1 * 8 + j;
        if ((rem & 2u) && key < best) { best = key; bpos = 1; }
    }
    {
        int j = __popc(rem & 3u) + s; if (j >= k) j -= k;
        const long long key = (long long)cv2 * 8 + j;
        if ((rem & 4u) && key < best) { best = key; bpos = 2; }
    }
    {
        int j = __popc(rem & 7u) + s; if (j >= k) j -= k;
        const long long key = (long long)cv3 * 8 + j;
        if ((rem & 8u) && key < best) { best = key; bpos = 3; }
    }
    return bpos;
}

template <int NT>
__global__ void __launch_bounds__(NT, 1) ka_leader_order4_kernel(const KaOrderParams p) {
    extern __shared__ __align__(16) unsigned char ka_smem[];
    int* ctr = reinterpret_cast<int*>(ka_smem);  // [N][4]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int S = p.S;

    for (int i = threadIdx.x; i < p.N * 4; i += blockDim.x) ctr[i] = p.ctr8[(i >> 2) * KA_MAX_SLOTS + (i & 3)];
    __syncthreads();

    const int64_t nwin = (p.Q + 31) >> 5;
    // software prefetch of the next window's inputs: global latency stays off the dependency chain
    uint32_t n_meta = 0u;
    int n_i0 = 0, n_i1 = 0, n_i2 = 0, n_i3 = 0, n_t0 = 0, n_t1 = 0, n_t2 = 0, n_t3 = 0;
    auto fetch = [&](int64_t win) {
        const int64_t q = win * 32 + lane;
        n_meta = 0u;
        if (win < nwin && q < p.Q) {
            n_meta = __ldg(p.meta + q);
            const int32_t* sp = p.set + q * S;
            const int32_t* tp = p.ticket + q * S;
            n_i0 = __ldg(sp); n_t0 = __ldg(tp);
            if (S > 1) { n_i1 = __ldg(sp + 1); n_t1 = __ldg(tp + 1); }
            if (S > 2) { n_i2 = __ldg(sp + 2); n_t2 = __ldg(tp + 2); }
            if (S > 3) { n_i3 = __ldg(sp + 3); n_t3 = __ldg(tp + 3); }
        }
    };
    fetch(warp);

    for (int64_t win = warp; win < nwin; win += nwarp) {
        const int64_t q = win * 32 + lane;
        const bool valid = q < p.Q;
        const uint32_t meta = n_meta;
        const int len = (int)(meta & 15u);
        // shared-memory byte offsets of this partition's broker rows (clamped for unused slots)
        const int a0 = (len > 0 ? max(n_i0, 0) : 0) * 16, a1 = (len > 1 ? max(n_i1, 0) : 0) * 16;
        const int a2 = (len > 2 ? max(n_i2, 0) : 0) * 16, a3 = (len > 3 ? max(n_i3, 0) : 0) * 16;
        const int t0 = n_t0, t1 = n_t1, t2 = n_t2, t3 = n_t3;
        fetch(win + nwarp);

        const int s2 = (int)((meta >> 4) & 1u), s3 = (int)((meta >> 5) & 3u), s4 = (int)((meta >> 7) & 3u);
        // RF=3 tie-breaks, independent of the counters (computed while waiting): rotated scan position of list
        // position i is j_i = (i + s3) % 3; t_xy = [j_x < j_y].
        const int j0 = s3, j1 = (s3 + 1 >= 3) ? s3 - 2 : s3 + 1, j2 = (s3 + 2 >= 3) ? s3 - 1 : s3 + 2;
        const int t10 = j1 < j0 ? 1 : 0, t20 = j2 < j0 ? 1 : 0, t21 = j2 < j1 ? 1 : 0;
        int b0 = 0, b1 = 1, b2 = 2, b3 = 3;  // chosen list position per slot
        bool pending = valid && len > 0;
        uint32_t spins = 0;
        const char* cb = reinterpret_cast<const char*>(ctr);
        for (;;) {
            int d = KA_BIG;
            if (pending) {
                // Read every broker row of the partition back-to-back (rows that already reached the ticket
                // are stable until this partition commits, so re-reading them is harmless) ...
                const int4 r0 = ka_lds_v4_volatile(reinterpret_cast<const int*>(cb + a0));
                const int4 r1 = ka_lds_v4_volatile(reinterpret_cast<const int*>(cb + a1));
                const int4 r2 = ka_lds_v4_volatile(reinterpret_cast<const int*>(cb + a2));
                int4 r3 = make_int4(0, 0, 0, 0);
                if (S > 3) r3 = ka_lds_v4_volatile(reinterpret_cast<const int*>(cb + a3));
                // ... distance = commits still to land on the slowest broker before it is this partition's turn
                d = t0 - (r0.x + r0.y + r0.z + r0.w);
                if (len > 1) d = max(d, t1 - (r1.x + r1.y + r1.z + r1.w));
                if (len > 2) d = max(d, t2 - (r2.x + r2.y + r2.z + r2.w));
                if (len > 3) d = max(d, t3 - (r3.x + r3.y + r3.z + r3.w));
                if (d == 0) {
                    // every broker of this partition has reached its ticket: order it (KAS:226-234) and commit
                    // counter[list[r]][r] += 1 (KAS:254-261) — one store per broker row.
                    if (len == 3) {
                        // RF = 3, branch-free. (c, j) lexicographic compare == c_x < c_y + [j_x < j_y].
                        const bool lt10 = r1.x < r0.x + t10;
                        const int m = lt10 ? 1 : 0;
                        const int cm = lt10 ? r1.x : r0.x;
                        const bool lt2m = r2.x < cm + (lt10 ? t21 : t20);
                        b0 = lt2m ? 2 : m;
                        const int v0 = lt2m ? r2.x : cm;
                        // slot 1: remaining positions pa < pb; rotation s2 decides who is scanned first (wins ties)
                        const int pa = (b0 == 0) ? 1 : 0, pb = (b0 == 2) ? 1 : 2;
                        const int ca = (b0 == 0) ? r1.y : r0.y;
                        const int cbv = (b0 == 2) ? r1.y : r2.y;
                        const bool pickb = s2 ? !(ca < cbv) : (cbv < ca);
                        b1 = pickb ? pb : pa;
                        const int v1 = pickb ? cbv : ca;
                        b2 = 3 - b0 - b1;
                        const int ad0 = (b0 == 0) ? a0 : ((b0 == 1) ? a1 : a2);
                        const int ad1 = (b1 == 0) ? a0 : ((b1 == 1) ? a1 : a2);
                        const int ad2 = (b2 == 0) ? a0 : ((b2 == 1) ? a1 : a2);
                        const int v2 = (b2 == 0) ? r0.z : ((b2 == 1) ? r1.z : r2.z);
                        ka_sts_volatile(reinterpret_cast<int*>(const_cast<char*>(cb) + ad0), v0 + 1);
                        ka_sts_volatile(reinterpret_cast<int*>(const_cast<char*>(cb) + ad1 + 4), v1 + 1);
                        ka_sts_volatile(reinterpret_cast<int*>(const_cast<char*>(cb) + ad2 + 8), v2 + 1);
                    } else {
                        uint32_t rem = (1u << len) - 1u;
                        int k = len;
                        b0 = ka_pick4(rem, k, k == 4 ? s4 : (k == 3 ? s3 : (k == 2 ? s2 : 0)), r0.x, r1.x, r2.x, r3.x);
                        rem &= ~(1u << b0); --k;
                        if (k > 0) {
                            b1 = ka_pick4(rem, k, k == 3 ? s3 : (k == 2 ? s2 : 0), r0.y, r1.y, r2.y, r3.y);
                            rem &= ~(1u << b1); --k;
                        }
                        if (k > 0) {
                            b2 = ka_pick4(rem, k, k == 2 ? s2 : 0, r0.z, r1.z, r2.z, r3.z);
                            rem &= ~(1u << b2); --k;
                        }
                        if (k > 0) b3 = __ffs(rem) - 1;
                        {
                            const int ad = ka_sel4(b0, a0, a1, a2, a3);
                            const int cv = ka_sel4(b0, r0.x, r1.x, r2.x, r3.x);
                            ka_sts_volatile(reinterpret_cast<int*>(const_cast<char*>(cb) + ad), cv + 1);
                        }
                        if (len > 1) {
                            const int ad = ka_sel4(b1, a0, a1, a2, a3);
                            const int cv = ka_sel4(b1, r0.y, r1.y, r2.y, r3.y);
                            ka_sts_volatile(reinterpret_cast<int*>(const_cast<char*>(cb) + ad + 4), cv + 1);
                        }
                        if (len > 2) {
                            const int ad = ka_sel4(b2, a0, a1, a2, a3);
                            const int cv = ka_sel4(b2, r0.z, r1.z, r2.z, r3.z);
                            ka_sts_volatile(reinterpret_cast<int*>(const_cast<char*>(cb) + ad + 8), cv + 1);
                        }
                        if (len > 3) {
                            const int ad = ka_sel4(b3, a0, a1, a2, a3);
                            const int cv = ka_sel4(b3, r0.w, r1.w, r2.w, r3.w);
                            ka_sts_volatile(reinterpret_cast<int*>(const_cast<char*>(cb) + ad + 12), cv + 1);
                        }
                    }
                    pending = false;
                }
            }
            if (!__any_sync(KA_FULL, pending)) break;  // every lane of the window has committed
            // (a __nanosleep back-off for warps far from the frontier was measured: no gain on any config)
            if (++spins > (1u << 22)) {  // guard: a ticket/set inconsistency must not hang the GPU
                if (pending) atomicExch(p.err_flag, KA_E_INTERNAL_SPIN);
                break;
            }
        }
        if (valid) {
            const int i0 = ka_sel4(b0, a0, a1, a2, a3) >> 4, i1 = ka_sel4(b1, a0, a1, a2, a3) >> 4;
            const int i2 = ka_sel4(b2, a0, a1, a2, a3) >> 4, i3 = ka_sel4(b3, a0, a1, a2, a3) >> 4;
            int32_t* o = p.out + q * S;
            o[0] = len > 0 ? __ldg(&p.broker_id[i0]) : -1;
            if (S > 1) o[1] = len > 1 ? __ldg(&p.broker_id[i1]) : -1;
            if (S > 2) o[2] = len > 2 ? __ldg(&p.broker_id[i2]) : -1;
            if (S > 3) o[3] = len > 3 ? __ldg(&p.broker_id[i3]) : -1;
            if (p.out_len) p.out_len[q] = len;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.N * 4; i += blockDim.x) p.ctr8[(i >> 2) * KA_MAX_SLOTS + (i & 3)] = ctr[i];
}

// ------------------------------------------------------------------------------------------------
// Kernel B, speculative straight-line variant for rows of <= 3 replicas (the RF=3 case of every BASELINE
// config). Each poll iteration reads the three counter rows, evaluates readiness AND the ordering decision
// side by side (independent instruction chains overlap inside the single warp), and commits with
// predicated stores. No divergent "ready path": an iteration costs the same whether 0 or 32 lanes commit.
// ------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT, 1) ka_leader_order3_kernel(const KaOrderParams p) {
    extern __shared__ __align__(16) unsigned char ka_smem[];
    int* ctr = reinterpret_cast<int*>(ka_smem);  // [N][4]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;

    for (int i = threadIdx.x; i < p.N * 4; i += blockDim.x) ctr[i] = p.ctr8[(i >> 2) * KA_MAX_SLOTS + (i & 3)];
    __syncthreads();

    const int64_t nwin = (p.Q + 31) >> 5;
    // software prefetch of the next window's packed records: global latency stays off the dependency chain
    int4 n_tk = make_int4(0, 0, 0, 0);
    uint32_t n_i01 = 0u;
    auto fetch = [&](int64_t win) {
        const int64_t q = win * 32 + lane;
        n_tk = make_int4(0, 0, 0, 0);
        n_i01 = 0u;
        if (win < nwin && q < p.Q) {
            n_tk = __ldg(p.tick4 + q);
            n_i01 = __ldg(p.idx01 + q);
        }
    };
    fetch(warp);
    const uint32_t cbase = ka_smem_u32(ctr);

    for (int64_t win = warp; win < nwin; win += nwarp) {
        const int64_t q = win * 32 + lane;
        const bool valid = q < p.Q;
        const uint32_t meta = (uint32_t)n_tk.w >> 16;
        const int len = (int)(meta & 15u);
        // byte offsets of the broker rows; unused slots alias slot 0 so the readiness test stays uniform
        const int o0 = (int)(n_i01 & 0xFFFFu) * 16;
        const int o1 = len > 1 ? (int)(n_i01 >> 16) * 16 : o0;
        const int o2 = len > 2 ? (int)((uint32_t)n_tk.w & 0xFFFFu) * 16 : o0;
        const int t0 = n_tk.x;
        const int t1 = len > 1 ? n_tk.y : t0, t2 = len > 2 ? n_tk.z : t0;
        fetch(win + nwarp);

        const int s2 = (int)((meta >> 4) & 1u), s3 = (int)((meta >> 5) & 3u);
        const int j0 = s3, j1 = (s3 + 1 >= 3) ? s3 - 2 : s3 + 1, j2 = (s3 + 2 >= 3) ? s3 - 1 : s3 + 2;
        const int t10 = j1 < j0 ? 1 : 0, t20 = j2 < j0 ? 1 : 0, t21 = j2 < j1 ? 1 : 0;

        bool pending = valid && len > 0;
        // lanes with nothing to do poll one fixed row (a broadcast: one shared-memory wavefront)
        uint32_t a0 = cbase + (pending ? o0 : 0), a1 = cbase + (pending ? o1 : 0), a2 = cbase + (pending ? o2 : 0);
        int pb0 = 0, pb1 = 1, pb2 = 2;
        uint32_t spins = 0;
        const bool all3 = !__any_sync(KA_FULL, pending && len != 3);
        if (all3) {
            // ---- tight loop: every partition of the window has exactly 3 replicas -------------------------------
            // Per iteration: 3 x LDS.128, readiness = OR of the three ticket differences, and the KAS:226-234
            // decision from three parallel pairwise compares per slot combined with predicate logic; predicated
            // commit. (c, j) lexicographic "x beats y"  <=>  c_x < c_y + [j_x < j_y]  (strict <, ties to the earlier
            // position of the rotated scan, KAS:263-278).
            int pcode = 0 | (1 << 2);  // b0 | b1 << 2
            do {
                int4 r0, r1, r2;
                asm volatile("ld.volatile.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w) : "r"(a0));
                asm volatile("ld.volatile.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r1.x), "=r"(r1.y), "=r"(r1.z), "=r"(r1.w) : "r"(a1));
                asm volatile("ld.volatile.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r2.x), "=r"(r2.y), "=r"(r2.z), "=r"(r2.w) : "r"(a2));
                const int e = (t0 - r0.x - r0.y - (r0.z + r0.w)) | (t1 - r1.x - r1.y - (r1.z + r1.w)) | (t2 - r2.x - r2.y - (r2.z + r2.w));
                const bool commit = pending && e == 0;
                // slot 0 (leader): pairwise "beats" among the three counters, then combine
                const bool L10 = r1.x < r0.x + t10, L20 = r2.x < r0.x + t20, L21 = r2.x < r1.x + t21;
                const bool is2 = L10 ? L21 : L20;           // position 2 wins
                const bool is1 = L10 && !L21;               // position 1 wins
                const int b0 = is2 ? 2 : (is1 ? 1 : 0);
                const int v0 = is2 ? r2.x : (is1 ? r1.x : r0.x);
                const uint32_t ad0 = is2 ? a2 : (is1 ? a1 : a0);
                // counter[list[r]][r] += 1 (KAS:254-261): one store per broker row; each store is issued as soon as
                // its operands exist so the next partition on that broker wakes up as early as possible
                if (commit) asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(ad0), "r"(v0 + 1) : "memory");
                // slot 1: for each possible remaining pair (lo, hi): does hi get picked? (rotation s2 = who is scanned first)
                const bool P01 = r1.y < r0.y + s2, P02 = r2.y < r0.y + s2, P12 = r2.y < r1.y + s2;
                const bool pickhi = is2 ? P01 : (is1 ? P02 : P12);
                // remaining pair: b0==0 -> (1,2), b0==1 -> (0,2), b0==2 -> (0,1)
                const int lo = (b0 == 0) ? 1 : 0, hi = is2 ? 1 : 2;
                const int b1 = pickhi ? hi : lo;
                const int ylo = (b0 == 0) ? r1.y : r0.y, yhi = is2 ? r1.y : r2.y;
                const int v1 = pickhi ? yhi : ylo;
                const uint32_t alo = (b0 == 0) ? a1 : a0, ahi = is2 ? a1 : a2;
                const uint32_t ad1 = pickhi ? ahi : alo;
                if (commit) asm volatile("st.volatile.shared.s32 [%0+4], %1;" ::"r"(ad1), "r"(v1 + 1) : "memory");
                const int zlo = (b0 == 0) ? r1.z : r0.z, zhi = is2 ? r1.z : r2.z;
                const uint32_t ad2 = pickhi ? alo : ahi;
                const int v2 = pickhi ? zlo : zhi;
                if (commit) {
                    asm volatile("st.volatile.shared.s32 [%0+8], %1;" ::"r"(ad2), "r"(v2 + 1) : "memory");
                    pcode = b0 | (b1 << 2);
                    pending = false;
                    a0 = a1 = a2 = cbase;
                }
            } while (__any_sync(KA_FULL, pending) && ++spins < (1u << 22));
            if (pending) atomicExch(p.err_flag, KA_E_INTERNAL_SPIN);  // guard tripped: ticket/set inconsistency
            pb0 = pcode & 3; pb1 = pcode >> 2; pb2 = 3 - pb0 - pb1;
        } else {
        for (;;) {
            int4 r0, r1, r2;
            asm volatile("ld.volatile.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w) : "r"(a0));
            asm volatile("ld.volatile.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r1.x), "=r"(r1.y), "=r"(r1.z), "=r"(r1.w) : "r"(a1));
            asm volatile("ld.volatile.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r2.x), "=r"(r2.y), "=r"(r2.z), "=r"(r2.w) : "r"(a2));
            // readiness: every broker row sums to this partition's ticket
            const int e = (t0 - r0.x - r0.y - (r0.z + r0.w)) | (t1 - r1.x - r1.y - (r1.z + r1.w)) | (t2 - r2.x - r2.y - (r2.z + r2.w));
            // speculative RF=3 decision (KAS:226-234; meaningful when e == 0 and len == 3)
            const bool lt10 = r1.x < r0.x + t10;
            const int m = lt10 ? 1 : 0;
            const int cm = lt10 ? r1.x : r0.x;
            const bool lt2m = r2.x < cm + (lt10 ? t21 : t20);
            const int b0 = lt2m ? 2 : m;
            const int v0 = lt2m ? r2.x : cm;
            const int ca = (b0 == 0) ? r1.y : r0.y;
            const int cbv = (b0 == 2) ? r1.y : r2.y;
            const int pa = (b0 == 0) ? 1 : 0, pb = (b0 == 2) ? 1 : 2;
            const bool pickb = s2 ? !(ca < cbv) : (cbv < ca);
            const int b1 = pickb ? pb : pa;
            const int v1 = pickb ? cbv : ca;
            const int b2 = 3 - b0 - b1;
            const uint32_t ad0 = (b0 == 0) ? a0 : ((b0 == 1) ? a1 : a2);
            const uint32_t ad1 = (b1 == 0) ? a0 : ((b1 == 1) ? a1 : a2);
            const uint32_t ad2 = (b2 == 0) ? a0 : ((b2 == 1) ? a1 : a2);
            const int v2 = (b2 == 0) ? r0.z : ((b2 == 1) ? r1.z : r2.z);
            const bool commit = pending && e == 0;
            if (commit) {
                if (len == 3) {
                    // counter[list[r]][r] += 1 (KAS:254-261): one store per broker row commits the partition
                    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(ad0), "r"(v0 + 1) : "memory");
                    asm volatile("st.volatile.shared.s32 [%0+4], %1;" ::"r"(ad1), "r"(v1 + 1) : "memory");
                    asm volatile("st.volatile.shared.s32 [%0+8], %1;" ::"r"(ad2), "r"(v2 + 1) : "memory");
                    pb0 = b0; pb1 = b1; pb2 = b2;
                } else if (len == 2) {
                    // two replicas: slot 0 over {0,1} rotated by s2, slot 1 is the other
                    const bool pick1 = s2 ? !(r0.x < r1.x) : (r1.x < r0.x);
                    pb0 = pick1 ? 1 : 0; pb1 = 1 - pb0;
                    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(pick1 ? a1 : a0), "r"((pick1 ? r1.x : r0.x) + 1) : "memory");
                    asm volatile("st.volatile.shared.s32 [%0+4], %1;" ::"r"(pick1 ? a0 : a1), "r"((pick1 ? r0.y : r1.y) + 1) : "memory");
                } else {
                    pb0 = 0;
                    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(a0), "r"(r0.x + 1) : "memory");
                }
                pending = false;
                a0 = a1 = a2 = cbase;
            }
            if (!__any_sync(KA_FULL, pending)) break;  // every lane of the window has committed
            if (++spins > (1u << 22)) {  // guard: a ticket/set inconsistency must not hang the GPU
                if (pending) atomicExch(p.err_flag, KA_E_INTERNAL_SPIN);
                break;
            }
        }
        }
        // one byte per partition: the chosen list position of slot 0 and slot 1 (slot 2 is the remaining one);
        // ka_emit_kernel turns it into broker ids, fully parallel, off the serial chain
        if (valid) p.pcode[q] = (uint8_t)(pb0 | (pb1 << 2));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.N * 4; i += blockDim.x) p.ctr8[(i >> 2) * KA_MAX_SLOTS + (i & 3)] = ctr[i];
}

// ------------------------------------------------------------------------------------------------
// Kernel B, generic row width (lists of 5..8 replicas — rare): array-based, correctness first.
// ------------------------------------------------------------------------------------------------
template <int RS>
__device__ __forceinline__ void ka_order_generic(const int (&c)[RS][RS], int len, uint32_t meta, int (&perm)[RS]) {
    uint32_t remmask = (1u << len) - 1u;
#pragma unroll
    for (int r = 0; r < RS; ++r) {
        if (r < len) {
            const int k = len - r;
            const int s = ka_rot_of<RS>(meta, k);
            long long best = 0x7FFFFFFFFFFFFFFFLL;
            int bpos = 0;
#pragma unroll
            for (int pos = 0; pos < RS; ++pos) {
                if ((remmask >> pos) & 1u) {
                    int j = __popc(remmask & ((1u << pos) - 1u)) + s;
                    if (j >= k) j -= k;
                    const long long key = (long long)c[pos][r] * 8 + j;
                    if (key < best) { best = key; bpos = pos; }
                }
            }
            perm[r] = bpos;
            remmask &= ~(1u << bpos);
        }
    }
}

template <int RS, int NT>
__global__ void __launch_bounds__(NT, 1) ka_leader_order_kernel(const KaOrderParams p) {
    extern __shared__ __align__(16) unsigned char ka_smem[];
    int* ctr = reinterpret_cast<int*>(ka_smem);  // [N][RS]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int S = p.S;

    for (int i = threadIdx.x; i < p.N * RS; i += blockDim.x) ctr[i] = p.ctr8[(i / RS) * KA_MAX_SLOTS + (i % RS)];
    __syncthreads();

    const int64_t nwin = (p.Q + 31) >> 5;
    for (int64_t win = warp; win < nwin; win += nwarp) {
        const int64_t q = win * 32 + lane;
        const bool valid = q < p.Q;
        const uint32_t meta = valid ? p.meta[q] : 0u;
        const int len = (int)(meta & 15u);
        int idx[RS], tk[RS];
#pragma unroll
        for (int i = 0; i < RS; ++i) {
            idx[i] = (valid && i < len) ? max(p.set[q * S + i], 0) : 0;
            tk[i] = (valid && i < len) ? p.ticket[q * S + i] : 0;
        }
        int perm[RS];
#pragma unroll
        for (int i = 0; i < RS; ++i) perm[i] = i;
        bool pending = valid && len > 0;
        uint32_t spins = 0;
        while (__any_sync(KA_FULL, pending)) {
            if (pending) {
                int c[RS][RS];
                bool ready = true;
#pragma unroll
                for (int i = 0; i < RS; ++i) {
                    if (i < len) {
                        int sum = 0;
#pragma unroll
                        for (int v = 0; v < RS; v += 4) {
                            int4 rw = ka_lds_v4_volatile(ctr + idx[i] * RS + v);
                            c[i][v] = rw.x; c[i][v + 1] = rw.y; c[i][v + 2] = rw.z; c[i][v + 3] = rw.w;
                            sum += rw.x + rw.y + rw.z + rw.w;
                        }
                        ready = ready && (sum == tk[i]);
                    }
                }
                if (ready) {
                    ka_order_generic<RS>(c, len, meta, perm);
#pragma unroll
                    for (int r = 0; r < RS; ++r) {
                        if (r < len) {
                            int bi = 0, cv = 0;
#pragma unroll
                            for (int pos = 0; pos < RS; ++pos)
                                if (perm[r] == pos) { bi = idx[pos]; cv = c[pos][r]; }
                            ka_sts_volatile(ctr + bi * RS + r, cv + 1);
                        }
                    }
                    pending = false;
                }
            }
            if (++spins > (1u << 22)) {
                if (pending) atomicExch(p.err_flag, KA_E_INTERNAL_SPIN);
                pending = false;
            }
        }
        if (valid) {
#pragma unroll
            for (int r = 0; r < RS; ++r) {
                if (r < S) {
                    int bi = -1;
#pragma unroll
                    for (int pos = 0; pos < RS; ++pos)
                        if (r < len && perm[r] == pos) bi = idx[pos];
                    p.out[q * S + r] = bi >= 0 ? __ldg(&p.broker_id[bi]) : -1;
                }
            }
            if (p.out_len) p.out_len[q] = len;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.N * RS; i += blockDim.x) p.ctr8[(i / RS) * KA_MAX_SLOTS + (i % RS)] = ctr[i];
}

// ------------------------------------------------------------------------------------------------
// Emit: (ascending index set, permutation code) -> ordered broker ids + list length. One thread per partition,
// fully parallel; keeps the id lookups and the 4 B/replica output stream off kernel B's serial chain.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ka_emit_kernel(const int32_t* __restrict__ set, const uint32_t* __restrict__ meta,
                                                      const uint8_t* __restrict__ pcode, const int32_t* __restrict__ broker_id, int64_t Q,
                                                      int S, int32_t* __restrict__ out, int32_t* __restrict__ out_len) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const int len = (int)(meta[q] & 15u);
    const int pc = pcode[q];
    const int b0 = pc & 3, b1 = (pc >> 2) & 3, b2 = 3 - b0 - b1;
    const int32_t* row = set + q * S;
    int32_t* o = out + q * S;
    const int i0 = len > 0 ? row[b0] : -1;
    o[0] = i0 >= 0 ? __ldg(broker_id + i0) : -1;
    if (S > 1) { const int i1 = len > 1 ? row[b1] : -1; o[1] = i1 >= 0 ? __ldg(broker_id + i1) : -1; }
    if (S > 2) { const int i2 = len > 2 ? row[b2] : -1; o[2] = i2 >= 0 ? __ldg(broker_id + i2) : -1; }
    if (out_len) out_len[q] = len;
}
