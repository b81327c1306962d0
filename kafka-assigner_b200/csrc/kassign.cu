// kassign.cu — C ABI (include/kassign.h) over the sm_100a kernels in kassign_kernels.cuh.
//
// Reference boundary: KafkaTopicAssigner.generateAssignment (KafkaTopicAssigner.java:42-72) batched over
// the topic loop of KafkaAssignmentGenerator.java:172-184. No CPU fallback exists in this library.
#include "kassign_kernels.cuh"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/kassign.h"

namespace {

constexpr int KA_SM_COUNT_FALLBACK = 148;
constexpr size_t KA_SMEM_BUDGET = 200 * 1024;   // per-CTA dynamic smem we allow ourselves (of 227 KB)
constexpr uint32_t KA_LUT_SMEM_MAX_RANGE = 32768;
constexpr uint32_t KA_LUT_GLOBAL_MAX_RANGE = 1u << 25;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct HostPinned {
    int err_topic;
    int spin_flag;
    int4 tstatus;
};

}  // namespace

struct ka_ctx {
    int device = 0;
    int sm_count = KA_SM_COUNT_FALLBACK;
    cudaStream_t stream = nullptr;  // used by the host-buffer entry points
    // broker table
    int N = 0;
    std::vector<int32_t> broker_id;
    std::vector<int32_t> broker_rack;
    int lut_mode = KA_LUT_SMEM;
    int min_id = 0;
    uint32_t range = 0;
    int blob_bytes = 16;
    int lut_off = 0;
    int R = 0, roff_off = 0, memb_off = 0;
    DevBuf d_blob, d_glut, d_broker_id, d_ctr8;
    // counters of brokers not in the current table (Context.counter is keyed by broker id)
    std::unordered_map<int32_t, std::vector<int32_t>> parked;
    // scratch
    DevBuf d_hash, d_part_off, d_rep_off, d_cur, d_set, d_meta, d_ticket, d_out, d_out_len, d_hist, d_tstatus, d_flags;
    DevBuf d_tick4, d_idx01, d_pcode;
    HostPinned* h_pin = nullptr;
    // bookkeeping
    bool timing = false;
    cudaEvent_t ev[10] = {};
    float last_ms[8] = {};
    bool ev_valid = false;
    int64_t launches = 0;
    int order_threads = 0;  // leader-order CTA size override (0 = heuristic from N); env KA_ORDER_THREADS wins
    // staged problem (between the context-free stage and the leader-order stage)
    bool staged = false;
    int64_t st_Q = 0;
    int st_S = 0, st_T = 0;
    int64_t st_L = 0;
    int st_chunks = 0, st_RS = 4, st_rank_warps = 1, st_rank_grid = 1;
    size_t st_rank_smem = 0, st_b_smem = 0;
    // async status
    cudaStream_t last_stream = nullptr;
    bool pending_status = false;
    const int32_t* last_part_id = nullptr;  // host pointer (ragged API) for status translation
    const int64_t* last_part_off = nullptr;
    ka_status last{};
};

namespace {

#define KA_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            std::fprintf(stderr, "[kassign] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e), __FILE__, __LINE__, \
                         cudaGetErrorString(_e));                                                  \
            return KA_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

int set_status(ka_status* st, int code, int topic = -1, int part = -1, int a = 0, int b = 0) {
    if (st) {
        st->code = code;
        st->topic_index = topic;
        st->partition = part;
        st->a = a;
        st->b = b;
    }
    return code;
}

inline size_t align16(size_t v) { return (v + 15) & ~size_t(15); }

// download current device counters into ctx->parked keyed by id
int park_counters(ka_ctx* c) {
    if (c->N == 0 || !c->d_ctr8.p) return KA_OK;
    std::vector<int32_t> h((size_t)c->N * KA_MAX_SLOTS);
    KA_CUDA(cudaMemcpy(h.data(), c->d_ctr8.p, h.size() * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < c->N; ++i) {
        const int32_t* row = h.data() + (size_t)i * KA_MAX_SLOTS;
        bool nz = false;
        for (int r = 0; r < KA_MAX_SLOTS; ++r) nz |= row[r] != 0;
        if (nz) c->parked[c->broker_id[i]] = std::vector<int32_t>(row, row + KA_MAX_SLOTS);
        else c->parked.erase(c->broker_id[i]);
    }
    return KA_OK;
}

struct Plan {
    // kernel A
    int a_warps, a_grid, a_load_bytes, a_slab_bytes, a_cnt_bytes, a_load_kind;  // kind 0=u8 1=u16 2=u32
    int a_rackptr, a_rp_bytes;
    size_t a_smem;
    // tickets
    int64_t L;
    int num_chunks;
    int t_warps_hist, t_warps_rank, t_grid_hist, t_grid_rank;
    size_t t_smem_hist, t_smem_rank;
    // order
    int RS;
    size_t b_smem;
};

int make_plan(ka_ctx* c, int64_t Q, int S, int Pmax, int64_t capmax, Plan& pl, ka_status* st) {
    const int N = c->N;
    // ---- kernel A
    pl.a_load_kind = capmax <= 255 ? 0 : (capmax <= 65535 ? 1 : 2);
    const int lsz = pl.a_load_kind == 0 ? 1 : (pl.a_load_kind == 1 ? 2 : 4);
    pl.a_load_bytes = (int)align16((size_t)std::max(N, 1) * lsz);
    pl.a_slab_bytes = (int)align16((size_t)std::max(Pmax, 1) * S * 2);
    pl.a_cnt_bytes = (int)align16((size_t)std::max(Pmax, 1));
    // spread phase: window scan over the rotated order by default (measured faster on every BASELINE config: the
    // monotone head finds a slot within ~1 window); KA_SPREAD_RACKPTR=1 selects the per-rack first-free-pointer
    // variant (exact too; pays off only when walks are long: many full nodes AND tight rack constraints).
    pl.a_rackptr = 0;
    if (const char* e = std::getenv("KA_SPREAD_RACKPTR")) pl.a_rackptr = std::atoi(e) && c->R > 0 && c->R <= 4096;
    pl.a_rp_bytes = pl.a_rackptr ? (int)align16((size_t)c->R * 2) : 0;
    const size_t per_warp = (size_t)pl.a_load_bytes + pl.a_slab_bytes + pl.a_cnt_bytes + 3 * (size_t)pl.a_rp_bytes;
    const size_t shared = 16 + (size_t)c->blob_bytes;
    if (shared + per_warp > KA_SMEM_BUDGET) return set_status(st, KA_ERR_LIMIT, -1, -1, Pmax, N);
    pl.a_warps = (int)std::min<size_t>(16, (KA_SMEM_BUDGET - shared) / per_warp);
    // prefer >= 2 CTAs/SM when the table is small: cap warps so that two CTAs fit
    pl.a_smem = shared + per_warp * pl.a_warps;
    // ---- tickets
    const int64_t windows = (Q + 31) / 32;
    const size_t hist_pw = (size_t)std::max(N, 1) * 4, rank_pw = hist_pw * 2;
    if (rank_pw > KA_SMEM_BUDGET) return set_status(st, KA_ERR_LIMIT, -1, -1, N, 0);
    pl.t_warps_hist = (int)std::max<size_t>(1, std::min<size_t>(32, KA_SMEM_BUDGET / hist_pw));
    pl.t_warps_rank = (int)std::max<size_t>(1, std::min<size_t>(32, KA_SMEM_BUDGET / rank_pw));
    const int64_t max_conc = (int64_t)c->sm_count * pl.t_warps_rank;
    int64_t nc = (int64_t)std::ceil(std::sqrt(5.0 * (double)std::max<int64_t>(windows, 1)));
    nc = std::max<int64_t>(1, std::min<int64_t>(nc, std::min<int64_t>(max_conc, std::max<int64_t>(windows, 1))));
    int64_t wpc = (std::max<int64_t>(windows, 1) + nc - 1) / nc;  // windows per chunk
    pl.L = wpc * 32;
    pl.num_chunks = (int)((std::max<int64_t>(windows, 1) + wpc - 1) / wpc);
    pl.t_smem_hist = hist_pw * pl.t_warps_hist;
    pl.t_smem_rank = rank_pw * pl.t_warps_rank;
    pl.t_grid_hist = (pl.num_chunks + pl.t_warps_hist - 1) / pl.t_warps_hist;
    pl.t_grid_rank = (pl.num_chunks + pl.t_warps_rank - 1) / pl.t_warps_rank;
    // ---- order
    pl.RS = S <= 4 ? 4 : 8;
    pl.b_smem = (size_t)std::max(N, 1) * pl.RS * 4;
    if (pl.b_smem > 220 * 1024) return set_status(st, KA_ERR_LIMIT, -1, -1, N, pl.RS);
    return KA_OK;
}

template <typename K>
cudaError_t allow_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// Stage 1 (context-free, shards across GPUs): kernel A + per-chunk broker histograms. All pointers are
// device pointers. Leaves the sorted replica sets in ctx scratch for enqueue_order().
int enqueue_stage(ka_ctx* c, cudaStream_t s, int T, const int32_t* d_hash, const int64_t* d_part_off, int P,
                  const int64_t* d_rep_off, int RF, const int32_t* d_cur, int desired_rf, int S, int Pmax,
                  int64_t capmax, int64_t Q, ka_status* st) {
    Plan pl;
    c->staged = false;
    int rc = make_plan(c, Q, S, Pmax, capmax, pl, st);
    if (rc != KA_OK) return rc;
    const int N = c->N;

    KA_CUDA(c->d_set.reserve((size_t)std::max<int64_t>(Q, 1) * S * 4));
    KA_CUDA(c->d_meta.reserve((size_t)std::max<int64_t>(Q, 1) * 4));
    KA_CUDA(c->d_ticket.reserve((size_t)std::max<int64_t>(Q, 1) * S * 4));
    KA_CUDA(c->d_hist.reserve((size_t)pl.num_chunks * std::max(N, 1) * 4));
    KA_CUDA(c->d_tstatus.reserve((size_t)std::max(T, 1) * sizeof(int4)));
    KA_CUDA(c->d_flags.reserve(64));

    // flags: [0] err_topic = INT_MAX, [1] spin flag = 0
    c->h_pin->err_topic = INT_MAX;
    c->h_pin->spin_flag = 0;
    static const int init_flags[2] = {INT_MAX, 0};
    KA_CUDA(cudaMemcpyAsync(c->d_flags.p, init_flags, sizeof(init_flags), cudaMemcpyHostToDevice, s));

    if (c->timing) KA_CUDA(cudaEventRecord(c->ev[1], s));

    if (T > 0) {
        KaSolveParams p{};
        p.T = T;
        p.topic_hash = d_hash;
        p.part_off = d_part_off;
        p.P = P;
        p.rep_off = d_rep_off;
        p.RF = RF;
        p.cur = d_cur;
        p.desired_rf = desired_rf;
        p.S = S;
        p.Pmax = Pmax;
        p.N = N;
        p.blob = c->d_blob.as<uint16_t>();
        p.blob_bytes = c->blob_bytes;
        p.lut_off = c->lut_off;
        p.R = c->R;
        p.rackptr = pl.a_rackptr;
        p.roff_off = c->roff_off;
        p.memb_off = c->memb_off;
        p.rp_bytes = pl.a_rp_bytes;
        p.lut_mode = c->lut_mode;
        p.min_id = c->min_id;
        p.range = c->range;
        p.glut = c->d_glut.as<uint16_t>();
        p.broker_id = c->d_broker_id.as<int32_t>();
        p.set = c->d_set.as<int32_t>();
        p.meta = c->d_meta.as<uint32_t>();
        p.tstatus = c->d_tstatus.as<int4>();
        p.err_topic = c->d_flags.as<int>();
        const int threads = pl.a_warps * 32;
        int grid = (T + pl.a_warps - 1) / pl.a_warps;
        int occ = 1;
        cudaError_t e;
        if (pl.a_load_kind == 0) {
            e = allow_smem(ka_sticky_spread_kernel<uint8_t>, pl.a_smem);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ka_sticky_spread_kernel<uint8_t>, threads, pl.a_smem);
        } else if (pl.a_load_kind == 1) {
            e = allow_smem(ka_sticky_spread_kernel<uint16_t>, pl.a_smem);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ka_sticky_spread_kernel<uint16_t>, threads, pl.a_smem);
        } else {
            e = allow_smem(ka_sticky_spread_kernel<uint32_t>, pl.a_smem);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ka_sticky_spread_kernel<uint32_t>, threads, pl.a_smem);
        }
        KA_CUDA(e);
        grid = std::min(grid, std::max(1, occ) * c->sm_count);
        if (pl.a_load_kind == 0)
            ka_sticky_spread_kernel<uint8_t><<<grid, threads, pl.a_smem, s>>>(p, pl.a_load_bytes, pl.a_slab_bytes, pl.a_cnt_bytes);
        else if (pl.a_load_kind == 1)
            ka_sticky_spread_kernel<uint16_t><<<grid, threads, pl.a_smem, s>>>(p, pl.a_load_bytes, pl.a_slab_bytes, pl.a_cnt_bytes);
        else
            ka_sticky_spread_kernel<uint32_t><<<grid, threads, pl.a_smem, s>>>(p, pl.a_load_bytes, pl.a_slab_bytes, pl.a_cnt_bytes);
        KA_CUDA(cudaGetLastError());
        c->launches++;
    }
    if (c->timing) KA_CUDA(cudaEventRecord(c->ev[2], s));

    if (Q > 0 && N > 0) {
        KA_CUDA(allow_smem(ka_ticket_hist_kernel, pl.t_smem_hist));
        ka_ticket_hist_kernel<<<pl.t_grid_hist, pl.t_warps_hist * 32, pl.t_smem_hist, s>>>(c->d_set.as<int32_t>(), Q, S, N, pl.L, pl.num_chunks,
                                                                                            c->d_hist.as<int32_t>());
        KA_CUDA(cudaGetLastError());
        c->launches++;
    }
    if (c->timing) KA_CUDA(cudaEventRecord(c->ev[6], s));
    c->staged = true;
    c->st_Q = Q;
    c->st_S = S;
    c->st_T = T;
    c->st_L = pl.L;
    c->st_chunks = pl.num_chunks;
    c->st_RS = pl.RS;
    c->st_rank_warps = pl.t_warps_rank;
    c->st_rank_grid = pl.t_grid_rank;
    c->st_rank_smem = pl.t_smem_rank;
    c->st_b_smem = pl.b_smem;
    return KA_OK;
}

// Stage 2 (the serial chain through Context.counter, KAS:202-239): ticket scan + rank, then kernel B.
int enqueue_order(ka_ctx* c, cudaStream_t s, int32_t* d_out, int32_t* d_out_len, ka_status* st) {
    if (!c->staged) return set_status(st, KA_ERR_BAD_ARG);
    const int N = c->N, S = c->st_S;
    const int64_t Q = c->st_Q;
    if (c->timing) KA_CUDA(cudaEventRecord(c->ev[7], s));
    int spec = 1;
    if (const char* e = std::getenv("KA_ORDER_SPEC")) spec = std::atoi(e);
    const bool packed = c->st_RS == 4 && S <= 3 && spec;  // rows of <= 3 replicas: packed records + emit kernel
    if (Q > 0 && N > 0) {
        KA_CUDA(allow_smem(ka_ticket_rank_kernel, c->st_rank_smem));
        if (packed) {
            KA_CUDA(c->d_tick4.reserve((size_t)Q * 16));
            KA_CUDA(c->d_idx01.reserve((size_t)Q * 4));
            KA_CUDA(c->d_pcode.reserve((size_t)Q));
        }
        ka_ticket_scan_kernel<<<(N + 255) / 256, 256, 0, s>>>(c->d_hist.as<int32_t>(), c->st_chunks, N, c->d_ctr8.as<int32_t>(), c->st_RS);
        KA_CUDA(cudaGetLastError());
        ka_ticket_rank_kernel<<<c->st_rank_grid, c->st_rank_warps * 32, c->st_rank_smem, s>>>(
            c->d_set.as<int32_t>(), Q, S, N, c->st_L, c->st_chunks, c->d_hist.as<int32_t>(), c->d_ticket.as<int32_t>(),
            c->d_meta.as<uint32_t>(), packed ? c->d_tick4.as<int4>() : nullptr, packed ? c->d_idx01.as<uint32_t>() : nullptr);
        KA_CUDA(cudaGetLastError());
        c->launches += 2;
    }
    if (c->timing) KA_CUDA(cudaEventRecord(c->ev[3], s));

    if (Q > 0 && N > 0) {
        KaOrderParams o{};
        o.Q = Q;
        o.S = S;
        o.N = N;
        o.set = c->d_set.as<int32_t>();
        o.ticket = c->d_ticket.as<int32_t>();
        o.meta = c->d_meta.as<uint32_t>();
        o.broker_id = c->d_broker_id.as<int32_t>();
        o.ctr8 = c->d_ctr8.as<int32_t>();
        o.out = d_out;
        o.out_len = d_out_len;
        o.err_flag = c->d_flags.as<int>() + 1;
        // CTA size of the leader-order kernel: the dependency DAG is ~N/RF wide and every extra polling warp costs the
        // frontier warps issue slots. Measured optimum (tools/phase_times.py, packed-record kernel): 128 threads at
        // N=100, 256 at N=1000, 512 at N=5000, 1024 at N=10000.
        int nt = N < 400 ? 128 : (N < 2500 ? 256 : (N < 7500 ? 512 : 1024));
        if (c->order_threads > 0) nt = c->order_threads;
        if (const char* e = std::getenv("KA_ORDER_THREADS")) nt = std::atoi(e);
        o.sleep_ns = 0;
        o.near_dist = 1;
        if (const char* e = std::getenv("KA_ORDER_SLEEP_NS")) o.sleep_ns = (unsigned)std::atoi(e);
        if (const char* e = std::getenv("KA_ORDER_NEAR")) o.near_dist = std::atoi(e);
        o.idle_polls = 4;
        if (const char* e = std::getenv("KA_ORDER_IDLE")) o.idle_polls = (unsigned)std::atoi(e);
        o.tick4 = c->d_tick4.as<int4>();
        o.idx01 = c->d_idx01.as<uint32_t>();
        o.pcode = c->d_pcode.as<uint8_t>();
        if (packed) {
#define KA_LAUNCH_ORDER3(NT)                                                        \
    do {                                                                            \
        KA_CUDA(allow_smem(ka_leader_order3_kernel<NT>, c->st_b_smem));             \
        ka_leader_order3_kernel<NT><<<1, NT, c->st_b_smem, s>>>(o);                 \
    } while (0)
            if (nt >= 1024) KA_LAUNCH_ORDER3(1024);
            else if (nt >= 512) KA_LAUNCH_ORDER3(512);
            else if (nt >= 256) KA_LAUNCH_ORDER3(256);
            else if (nt >= 128) KA_LAUNCH_ORDER3(128);
            else KA_LAUNCH_ORDER3(64);
#undef KA_LAUNCH_ORDER3
        } else if (c->st_RS == 4) {
#define KA_LAUNCH_ORDER4(NT)                                                        \
    do {                                                                            \
        KA_CUDA(allow_smem(ka_leader_order4_kernel<NT>, c->st_b_smem));             \
        ka_leader_order4_kernel<NT><<<1, NT, c->st_b_smem, s>>>(o);                 \
    } while (0)
            if (nt >= 1024) KA_LAUNCH_ORDER4(1024);
            else if (nt >= 512) KA_LAUNCH_ORDER4(512);
            else if (nt >= 256) KA_LAUNCH_ORDER4(256);
            else KA_LAUNCH_ORDER4(128);
#undef KA_LAUNCH_ORDER4
        } else {
            KA_CUDA(allow_smem(ka_leader_order_kernel<8, 256>, c->st_b_smem));
            ka_leader_order_kernel<8, 256><<<1, 256, c->st_b_smem, s>>>(o);
        }
        KA_CUDA(cudaGetLastError());
        c->launches++;
        if (packed) {
            ka_emit_kernel<<<(unsigned)((Q + 255) / 256), 256, 0, s>>>(c->d_set.as<int32_t>(), c->d_meta.as<uint32_t>(), c->d_pcode.as<uint8_t>(),
                                                                     c->d_broker_id.as<int32_t>(), Q, S, d_out, d_out_len);
            KA_CUDA(cudaGetLastError());
            c->launches++;
        }
    }
    if (c->timing) KA_CUDA(cudaEventRecord(c->ev[4], s));
    c->staged = false;

    // status words back to pinned host memory (async)
    KA_CUDA(cudaMemcpyAsync(&c->h_pin->err_topic, c->d_flags.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    return KA_OK;
}

int enqueue_pipeline(ka_ctx* c, cudaStream_t s, int T, const int32_t* d_hash, const int64_t* d_part_off, int P,
                     const int64_t* d_rep_off, int RF, const int32_t* d_cur, int desired_rf, int S, int Pmax,
                     int64_t capmax, int64_t Q, int32_t* d_out, int32_t* d_out_len, ka_status* st) {
    int rc = enqueue_stage(c, s, T, d_hash, d_part_off, P, d_rep_off, RF, d_cur, desired_rf, S, Pmax, capmax, Q, st);
    if (rc != KA_OK) return rc;
    return enqueue_order(c, s, d_out, d_out_len, st);
}

// Wait for the stream, translate device flags into a ka_status.
int finish_status(ka_ctx* c, cudaStream_t s, ka_status* st) {
    KA_CUDA(cudaStreamSynchronize(s));
    c->pending_status = false;
    ka_status r{};
    r.code = KA_OK;
    r.topic_index = -1;
    r.partition = -1;
    if (c->h_pin->err_topic != INT_MAX) {
        const int t = c->h_pin->err_topic;
        KA_CUDA(cudaMemcpy(&c->h_pin->tstatus, c->d_tstatus.as<int4>() + t, sizeof(int4), cudaMemcpyDeviceToHost));
        r.code = c->h_pin->tstatus.x;
        r.topic_index = t;
        int ord = c->h_pin->tstatus.y;
        r.partition = ord;
        if (ord >= 0 && c->last_part_id && c->last_part_off) r.partition = c->last_part_id[c->last_part_off[t] + ord];
        r.a = c->h_pin->tstatus.z;
        r.b = c->h_pin->tstatus.w;
    } else if (c->h_pin->spin_flag != 0) {
        r.code = KA_ERR_CUDA;
        std::fprintf(stderr, "[kassign] internal error: leader-order dataflow guard tripped\n");
    }
    if (c->timing && c->ev_valid) {
        for (int i = 0; i < 8; ++i) c->last_ms[i] = 0.f;
        cudaEventElapsedTime(&c->last_ms[3], c->ev[0], c->ev[1]);  // H2D
        cudaEventElapsedTime(&c->last_ms[0], c->ev[1], c->ev[2]);  // kernel A
        float t1 = 0.f, t2 = 0.f;
        cudaEventElapsedTime(&t1, c->ev[2], c->ev[6]);             // ticket histogram (stage 1)
        cudaEventElapsedTime(&t2, c->ev[7], c->ev[3]);             // ticket scan + rank (stage 2)
        c->last_ms[1] = t1 + t2;
        cudaEventElapsedTime(&c->last_ms[2], c->ev[3], c->ev[4]);  // kernel B
        cudaEventElapsedTime(&c->last_ms[4], c->ev[4], c->ev[5]);  // D2H
        cudaEventElapsedTime(&c->last_ms[5], c->ev[0], c->ev[5]);  // total
    }
    c->last = r;
    if (st) *st = r;
    return r.code;
}

}  // namespace

// =================================================================================================
extern "C" {

const char* ka_version(void) { return "kassign-b200 0.1 (sm_100a)"; }

int32_t ka_java_string_hash(const char* s) {
    // java.lang.String.hashCode over UTF-16 code units (KAS:190)
    uint32_t h = 0;
    const unsigned char* u = reinterpret_cast<const unsigned char*>(s);
    while (*u) {
        uint32_t cp;
        int extra;
        unsigned char b = *u++;
        if (b < 0x80) { cp = b; extra = 0; }
        else if ((b & 0xE0) == 0xC0) { cp = b & 0x1F; extra = 1; }
        else if ((b & 0xF0) == 0xE0) { cp = b & 0x0F; extra = 2; }
        else if ((b & 0xF8) == 0xF0) { cp = b & 0x07; extra = 3; }
        else { cp = 0xFFFD; extra = 0; }
        while (extra-- > 0 && *u) cp = (cp << 6) | (*u++ & 0x3F);
        if (cp >= 0x10000) {
            cp -= 0x10000;
            h = h * 31u + (0xD800u + (cp >> 10));
            h = h * 31u + (0xDC00u + (cp & 0x3FFu));
        } else {
            h = h * 31u + cp;
        }
    }
    return (int32_t)h;
}

int32_t ka_rack_indices(int32_t N, const int32_t* broker_id, const char* const* rack_name, int32_t* broker_rack) {
    if (N < 0 || (N > 0 && (!broker_id || !broker_rack))) return KA_ERR_BAD_ARG;
    // rack key = the rack string, or Integer.toString(id) when no rack is defined (KAS:81-86); brokers
    // share a Rack object iff their keys are equal strings (KAS:90-94).
    std::map<std::string, int32_t> key2idx;
    for (int i = 0; i < N; ++i) {
        std::string key = (rack_name && rack_name[i]) ? std::string(rack_name[i]) : std::to_string(broker_id[i]);
        auto it = key2idx.find(key);
        if (it == key2idx.end()) it = key2idx.emplace(key, (int32_t)key2idx.size()).first;
        broker_rack[i] = it->second;
    }
    return KA_OK;
}

ka_ctx* ka_ctx_create(int32_t device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    ka_ctx* c = new (std::nothrow) ka_ctx();
    if (!c) return nullptr;
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return nullptr; }
    if (cudaHostAlloc(reinterpret_cast<void**>(&c->h_pin), sizeof(HostPinned), cudaHostAllocDefault) != cudaSuccess) { delete c; return nullptr; }
    for (auto& e : c->ev) cudaEventCreate(&e);
    return c;
}

void ka_ctx_destroy(ka_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (DevBuf* b : {&c->d_blob, &c->d_glut, &c->d_broker_id, &c->d_ctr8, &c->d_hash, &c->d_part_off, &c->d_rep_off, &c->d_cur, &c->d_set,
                      &c->d_meta, &c->d_ticket, &c->d_tick4, &c->d_idx01, &c->d_pcode, &c->d_out, &c->d_out_len, &c->d_hist, &c->d_tstatus, &c->d_flags})
        b->release();
    for (auto& e : c->ev)
        if (e) cudaEventDestroy(e);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int32_t ka_ctx_reset(ka_ctx* c) {
    if (!c) return KA_ERR_NO_DEVICE;
    KA_CUDA(cudaSetDevice(c->device));
    c->parked.clear();
    if (c->N > 0 && c->d_ctr8.p) KA_CUDA(cudaMemset(c->d_ctr8.p, 0, (size_t)c->N * KA_MAX_SLOTS * 4));
    return KA_OK;
}

int32_t ka_ctx_set_brokers(ka_ctx* c, int32_t N, const int32_t* broker_id, const int32_t* broker_rack) {
    if (!c) return KA_ERR_NO_DEVICE;
    if (N < 0 || (N > 0 && (!broker_id || !broker_rack))) return KA_ERR_BAD_ARG;
    if (N > 65535) return KA_ERR_LIMIT;
    for (int i = 0; i < N; ++i) {
        if (i > 0 && broker_id[i] <= broker_id[i - 1]) return KA_ERR_BAD_ARG;  // strictly ascending
        if (broker_rack[i] < 0 || broker_rack[i] >= 65535) return KA_ERR_BAD_ARG;
    }
    KA_CUDA(cudaSetDevice(c->device));
    if (c->pending_status) finish_status(c, c->last_stream, nullptr);
    int rc = park_counters(c);
    if (rc != KA_OK) return rc;

    c->N = N;
    c->broker_id.assign(broker_id, broker_id + N);
    c->broker_rack.assign(broker_rack, broker_rack + N);
    c->min_id = N > 0 ? broker_id[0] : 0;
    const uint64_t range64 = N > 0 ? (uint64_t)((int64_t)broker_id[N - 1] - (int64_t)broker_id[0]) + 1 : 0;
    const size_t npad = align16((size_t)std::max(N, 1) * 2) / 2;  // uint16 elements, 16B multiple
    // compact rack ids in order of first appearance (rack identity is all that matters, KAS:90-94)
    std::vector<uint16_t> rackc(std::max(N, 1), 0);
    {
        std::unordered_map<int32_t, int> seen;
        for (int i = 0; i < N; ++i) {
            auto it = seen.find(broker_rack[i]);
            if (it == seen.end()) it = seen.emplace(broker_rack[i], (int)seen.size()).first;
            rackc[i] = (uint16_t)it->second;
        }
        c->R = (int)seen.size();
    }
    std::vector<uint16_t> blob;
    size_t lut_elems = 0;
    if (range64 <= KA_LUT_SMEM_MAX_RANGE) {
        c->lut_mode = KA_LUT_SMEM;
        c->range = (uint32_t)range64;
        lut_elems = align16((size_t)std::max<uint64_t>(range64, 1) * 2) / 2;
    } else if (range64 <= KA_LUT_GLOBAL_MAX_RANGE) {
        c->lut_mode = KA_LUT_GLOBAL;
        c->range = (uint32_t)range64;
        std::vector<uint16_t> g((size_t)range64, (uint16_t)KA_DEAD);
        for (int i = 0; i < N; ++i) g[(size_t)((int64_t)broker_id[i] - c->min_id)] = (uint16_t)i;
        KA_CUDA(c->d_glut.reserve(g.size() * 2));
        KA_CUDA(cudaMemcpy(c->d_glut.p, g.data(), g.size() * 2, cudaMemcpyHostToDevice));
    } else {
        c->lut_mode = KA_LUT_BSEARCH;
        c->range = 0;
    }
    const size_t roff_elems = align16((size_t)(c->R + 1) * 2) / 2;
    c->lut_off = (int)npad;
    c->roff_off = (int)(npad + lut_elems);
    c->memb_off = (int)(npad + lut_elems + roff_elems);
    blob.assign(npad + lut_elems + roff_elems + npad, (uint16_t)KA_DEAD);
    for (int i = 0; i < N; ++i) {
        blob[i] = rackc[i];
        if (c->lut_mode == KA_LUT_SMEM) blob[npad + (size_t)((int64_t)broker_id[i] - c->min_id)] = (uint16_t)i;
    }
    {   // rack member lists (CSR): sorted indices ascending inside each rack
        std::vector<int> cntr(c->R + 1, 0);
        for (int i = 0; i < N; ++i) cntr[rackc[i] + 1]++;
        for (int r = 0; r < c->R; ++r) cntr[r + 1] += cntr[r];
        for (int r = 0; r <= c->R; ++r) blob[c->roff_off + r] = (uint16_t)cntr[r];
        std::vector<int> fill(cntr.begin(), cntr.end() - 1);
        for (int i = 0; i < N; ++i) blob[c->memb_off + fill[rackc[i]]++] = (uint16_t)i;
    }
    c->blob_bytes = (int)(blob.size() * 2);
    KA_CUDA(c->d_blob.reserve(blob.size() * 2));
    KA_CUDA(cudaMemcpy(c->d_blob.p, blob.data(), blob.size() * 2, cudaMemcpyHostToDevice));
    KA_CUDA(c->d_broker_id.reserve((size_t)std::max(N, 1) * 4));
    if (N > 0) KA_CUDA(cudaMemcpy(c->d_broker_id.p, broker_id, (size_t)N * 4, cudaMemcpyHostToDevice));
    // counters for the new table
    std::vector<int32_t> h((size_t)std::max(N, 1) * KA_MAX_SLOTS, 0);
    for (int i = 0; i < N; ++i) {
        auto it = c->parked.find(broker_id[i]);
        if (it != c->parked.end()) std::copy(it->second.begin(), it->second.end(), h.begin() + (size_t)i * KA_MAX_SLOTS);
    }
    KA_CUDA(c->d_ctr8.reserve(h.size() * 4));
    KA_CUDA(cudaMemcpy(c->d_ctr8.p, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    return KA_OK;
}

int32_t ka_ctx_counter_slots(ka_ctx*) { return KA_MAX_SLOTS; }

int32_t ka_ctx_get_counters(ka_ctx* c, int32_t* counter) {
    if (!c) return KA_ERR_NO_DEVICE;
    if (!counter) return KA_ERR_BAD_ARG;
    KA_CUDA(cudaSetDevice(c->device));
    if (c->pending_status) finish_status(c, c->last_stream, nullptr);
    if (c->N > 0) KA_CUDA(cudaMemcpy(counter, c->d_ctr8.p, (size_t)c->N * KA_MAX_SLOTS * 4, cudaMemcpyDeviceToHost));
    return KA_OK;
}

int32_t ka_ctx_set_counters(ka_ctx* c, const int32_t* counter) {
    if (!c) return KA_ERR_NO_DEVICE;
    if (!counter) return KA_ERR_BAD_ARG;
    KA_CUDA(cudaSetDevice(c->device));
    if (c->pending_status) finish_status(c, c->last_stream, nullptr);
    if (c->N > 0) KA_CUDA(cudaMemcpy(c->d_ctr8.p, counter, (size_t)c->N * KA_MAX_SLOTS * 4, cudaMemcpyHostToDevice));
    return KA_OK;
}

int32_t ka_ctx_export_counters_device(ka_ctx* c, int32_t* d_counter, void* stream) {
    if (!c) return KA_ERR_NO_DEVICE;
    if (!d_counter) return KA_ERR_BAD_ARG;
    KA_CUDA(cudaSetDevice(c->device));
    if (c->N > 0)
        KA_CUDA(cudaMemcpyAsync(d_counter, c->d_ctr8.p, (size_t)c->N * KA_MAX_SLOTS * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return KA_OK;
}

int32_t ka_ctx_import_counters_device(ka_ctx* c, const int32_t* d_counter, void* stream) {
    if (!c) return KA_ERR_NO_DEVICE;
    if (!d_counter) return KA_ERR_BAD_ARG;
    KA_CUDA(cudaSetDevice(c->device));
    if (c->N > 0)
        KA_CUDA(cudaMemcpyAsync(c->d_ctr8.p, d_counter, (size_t)c->N * KA_MAX_SLOTS * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return KA_OK;
}

int32_t ka_ctx_set_timing(ka_ctx* c, int32_t enabled) {
    if (!c) return KA_ERR_NO_DEVICE;
    c->timing = enabled != 0;
    return KA_OK;
}

int32_t ka_ctx_last_timing(ka_ctx* c, float* ms) {
    if (!c) return KA_ERR_NO_DEVICE;
    if (!ms) return KA_ERR_BAD_ARG;
    for (int i = 0; i < 8; ++i) ms[i] = c->last_ms[i];
    return KA_OK;
}

int64_t ka_ctx_launch_count(ka_ctx* c) { return c ? c->launches : 0; }

int32_t ka_last_status(ka_ctx* c, ka_status* st) {
    if (!c) return set_status(st, KA_ERR_NO_DEVICE);
    if (cudaSetDevice(c->device) != cudaSuccess) return set_status(st, KA_ERR_CUDA);
    if (c->pending_status) return finish_status(c, c->last_stream, st);
    if (st) *st = c->last;
    return c->last.code;
}

static int validate_dense(ka_ctx* c, int32_t T, int32_t P, int32_t RF, int32_t desired_rf, int32_t S, ka_status* st) {
    if (!c) return set_status(st, KA_ERR_NO_DEVICE);
    if (T < 0 || P < 0 || RF < 0) return set_status(st, KA_ERR_BAD_ARG);
    if (S < 1 || S > KA_MAX_SLOTS) return set_status(st, KA_ERR_LIMIT, -1, -1, S);
    const int rf_t = desired_rf >= 0 ? desired_rf : RF;
    if (S < RF || (rf_t <= c->N && S < rf_t)) return set_status(st, KA_ERR_BAD_ARG, -1, -1, S);
    return KA_OK;
}

int32_t ka_solve_dense_device(ka_ctx* c, int32_t T, const int32_t* d_topic_hash, int32_t P, int32_t RF,
                              const int32_t* d_cur_broker, int32_t desired_rf, int32_t out_stride,
                              int32_t* d_out_len, int32_t* d_out_broker, void* stream, ka_status* st) {
    int rc = validate_dense(c, T, P, RF, desired_rf, out_stride, st);
    if (rc != KA_OK) return rc;
    if (cudaSetDevice(c->device) != cudaSuccess) return set_status(st, KA_ERR_CUDA);
    if (c->pending_status) finish_status(c, c->last_stream, nullptr);
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t Q = (int64_t)T * P;
    const int rf_t = desired_rf >= 0 ? desired_rf : RF;
    const int64_t capmax = c->N > 0 ? ((int64_t)P * std::max(rf_t, 0) + c->N - 1) / c->N : 0;
    c->last_part_id = nullptr;
    c->last_part_off = nullptr;
    if (c->timing) { cudaEventRecord(c->ev[0], s); }
    rc = enqueue_pipeline(c, s, T, d_topic_hash, nullptr, P, nullptr, RF, d_cur_broker, desired_rf, out_stride, P, capmax, Q,
                          d_out_broker, d_out_len, st);
    if (rc != KA_OK) return rc;
    if (c->timing) { cudaEventRecord(c->ev[5], s); c->ev_valid = true; }
    c->last_stream = s;
    c->pending_status = true;
    if (st) return finish_status(c, s, st);
    return KA_OK;
}

int32_t ka_stage_dense_device(ka_ctx* c, int32_t T, const int32_t* d_topic_hash, int32_t P, int32_t RF,
                              const int32_t* d_cur_broker, int32_t desired_rf, int32_t out_stride, void* stream) {
    ka_status lst;
    int rc = validate_dense(c, T, P, RF, desired_rf, out_stride, &lst);
    if (rc != KA_OK) return rc;
    if (cudaSetDevice(c->device) != cudaSuccess) return KA_ERR_CUDA;
    if (c->pending_status) finish_status(c, c->last_stream, nullptr);
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t Q = (int64_t)T * P;
    const int rf_t = desired_rf >= 0 ? desired_rf : RF;
    const int64_t capmax = c->N > 0 ? ((int64_t)P * std::max(rf_t, 0) + c->N - 1) / c->N : 0;
    c->last_part_id = nullptr;
    c->last_part_off = nullptr;
    if (c->timing) { cudaEventRecord(c->ev[0], s); }
    return enqueue_stage(c, s, T, d_topic_hash, nullptr, P, nullptr, RF, d_cur_broker, desired_rf, out_stride, P, capmax, Q, &lst);
}

int32_t ka_order_device(ka_ctx* c, int32_t* d_out_len, int32_t* d_out_broker, void* stream, ka_status* st) {
    if (!c) return set_status(st, KA_ERR_NO_DEVICE);
    if (cudaSetDevice(c->device) != cudaSuccess) return set_status(st, KA_ERR_CUDA);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = enqueue_order(c, s, d_out_broker, d_out_len, st);
    if (rc != KA_OK) return rc;
    if (c->timing) { cudaEventRecord(c->ev[5], s); c->ev_valid = true; }
    c->last_stream = s;
    c->pending_status = true;
    if (st) return finish_status(c, s, st);
    return KA_OK;
}

int32_t ka_solve_dense(ka_ctx* c, int32_t T, const int32_t* topic_hash, int32_t P, int32_t RF,
                       const int32_t* cur_broker, int32_t desired_rf, int32_t out_stride,
                       int32_t* out_len, int32_t* out_broker, ka_status* st) {
    int rc = validate_dense(c, T, P, RF, desired_rf, out_stride, st);
    if (rc != KA_OK) return rc;
    if (cudaSetDevice(c->device) != cudaSuccess) return set_status(st, KA_ERR_CUDA);
    if (c->pending_status) finish_status(c, c->last_stream, nullptr);
    cudaStream_t s = c->stream;
    const int64_t Q = (int64_t)T * P, R = Q * RF;
    if ((T > 0 && !topic_hash) || (R > 0 && !cur_broker) || (Q > 0 && !out_broker)) return set_status(st, KA_ERR_BAD_ARG);
    KA_CUDA(c->d_hash.reserve((size_t)std::max(T, 1) * 4));
    KA_CUDA(c->d_cur.reserve((size_t)std::max<int64_t>(R, 1) * 4));
    KA_CUDA(c->d_out.reserve((size_t)std::max<int64_t>(Q, 1) * out_stride * 4));
    KA_CUDA(c->d_out_len.reserve((size_t)std::max<int64_t>(Q, 1) * 4));
    if (c->timing) KA_CUDA(cudaEventRecord(c->ev[0], s));
    if (T > 0) KA_CUDA(cudaMemcpyAsync(c->d_hash.p, topic_hash, (size_t)T * 4, cudaMemcpyHostToDevice, s));
    if (R > 0) KA_CUDA(cudaMemcpyAsync(c->d_cur.p, cur_broker, (size_t)R * 4, cudaMemcpyHostToDevice, s));
    const int rf_t = desired_rf >= 0 ? desired_rf : RF;
    const int64_t capmax = c->N > 0 ? ((int64_t)P * std::max(rf_t, 0) + c->N - 1) / c->N : 0;
    c->last_part_id = nullptr;
    c->last_part_off = nullptr;
    rc = enqueue_pipeline(c, s, T, c->d_hash.as<int32_t>(), nullptr, P, nullptr, RF, c->d_cur.as<int32_t>(), desired_rf, out_stride, P,
                          capmax, Q, c->d_out.as<int32_t>(), out_len ? c->d_out_len.as<int32_t>() : nullptr, st);
    if (rc != KA_OK) return rc;
    if (Q > 0) {
        if (c->N > 0) {
            KA_CUDA(cudaMemcpyAsync(out_broker, c->d_out.p, (size_t)Q * out_stride * 4, cudaMemcpyDeviceToHost, s));
            if (out_len) KA_CUDA(cudaMemcpyAsync(out_len, c->d_out_len.p, (size_t)Q * 4, cudaMemcpyDeviceToHost, s));
        }
    }
    if (c->timing) { KA_CUDA(cudaEventRecord(c->ev[5], s)); c->ev_valid = true; }
    c->last_stream = s;
    c->pending_status = true;
    ka_status local;
    return finish_status(c, s, st ? st : &local);
}

int32_t ka_solve(ka_ctx* c, int32_t T, const int32_t* topic_hash, const int64_t* part_off,
                 const int32_t* part_id, const int64_t* rep_off, const int32_t* cur_broker,
                 int32_t desired_rf, int32_t out_stride, int32_t* out_len, int32_t* out_broker,
                 ka_status* st) {
    if (!c) return set_status(st, KA_ERR_NO_DEVICE);
    if (T < 0 || (T > 0 && (!topic_hash || !part_off))) return set_status(st, KA_ERR_BAD_ARG);
    const int S = out_stride;
    if (S < 1 || S > KA_MAX_SLOTS) return set_status(st, KA_ERR_LIMIT, -1, -1, S);
    if (cudaSetDevice(c->device) != cudaSuccess) return set_status(st, KA_ERR_CUDA);
    if (c->pending_status) finish_status(c, c->last_stream, nullptr);
    const int64_t Q = T > 0 ? part_off[T] : 0;
    if (Q < 0 || (T > 0 && part_off[0] != 0) || (Q > 0 && (!rep_off || !out_broker))) return set_status(st, KA_ERR_BAD_ARG);
    const int64_t R = Q > 0 ? rep_off[Q] : 0;
    if (R < 0 || (Q > 0 && rep_off[0] != 0) || (R > 0 && !cur_broker)) return set_status(st, KA_ERR_BAD_ARG);
    // host-side sizing scan: largest topic, largest current list, largest capacity (KAS:65-71)
    int Pmax = 0;
    int64_t capmax = 0, maxsz = 0;
    for (int t = 0; t < T; ++t) {
        const int64_t a = part_off[t], b = part_off[t + 1];
        if (b < a) return set_status(st, KA_ERR_BAD_ARG, t);
        const int64_t Pn = b - a;
        if (Pn > INT_MAX / 16) return set_status(st, KA_ERR_LIMIT, t, -1, (int)std::min<int64_t>(Pn, INT_MAX));
        Pmax = std::max<int>(Pmax, (int)Pn);
        int64_t rf_t = desired_rf;
        if (rf_t < 0 && Pn > 0) rf_t = rep_off[a + 1] - rep_off[a];
        if (rf_t > 0 && c->N > 0 && rf_t <= c->N) {
            capmax = std::max<int64_t>(capmax, (Pn * rf_t + c->N - 1) / c->N);
            if (rf_t > S) return set_status(st, KA_ERR_BAD_ARG, t, -1, S);
        }
    }
    for (int64_t g = 0; g < Q; ++g) {
        const int64_t sz = rep_off[g + 1] - rep_off[g];
        if (sz < 0) return set_status(st, KA_ERR_BAD_ARG);
        maxsz = std::max(maxsz, sz);
    }
    if (maxsz > S) return set_status(st, KA_ERR_BAD_ARG, -1, -1, S);

    cudaStream_t s = c->stream;
    KA_CUDA(c->d_hash.reserve((size_t)std::max(T, 1) * 4));
    KA_CUDA(c->d_part_off.reserve((size_t)(T + 1) * 8));
    KA_CUDA(c->d_rep_off.reserve((size_t)(Q + 1) * 8));
    KA_CUDA(c->d_cur.reserve((size_t)std::max<int64_t>(R, 1) * 4));
    KA_CUDA(c->d_out.reserve((size_t)std::max<int64_t>(Q, 1) * S * 4));
    KA_CUDA(c->d_out_len.reserve((size_t)std::max<int64_t>(Q, 1) * 4));
    if (c->timing) KA_CUDA(cudaEventRecord(c->ev[0], s));
    if (T > 0) {
        KA_CUDA(cudaMemcpyAsync(c->d_hash.p, topic_hash, (size_t)T * 4, cudaMemcpyHostToDevice, s));
        KA_CUDA(cudaMemcpyAsync(c->d_part_off.p, part_off, (size_t)(T + 1) * 8, cudaMemcpyHostToDevice, s));
    }
    if (Q > 0) KA_CUDA(cudaMemcpyAsync(c->d_rep_off.p, rep_off, (size_t)(Q + 1) * 8, cudaMemcpyHostToDevice, s));
    if (R > 0) KA_CUDA(cudaMemcpyAsync(c->d_cur.p, cur_broker, (size_t)R * 4, cudaMemcpyHostToDevice, s));
    c->last_part_id = part_id;
    c->last_part_off = part_off;
    int rc = enqueue_pipeline(c, s, T, c->d_hash.as<int32_t>(), c->d_part_off.as<int64_t>(), 0, c->d_rep_off.as<int64_t>(), 0,
                              c->d_cur.as<int32_t>(), desired_rf, S, Pmax, capmax, Q, c->d_out.as<int32_t>(),
                              out_len ? c->d_out_len.as<int32_t>() : nullptr, st);
    if (rc != KA_OK) return rc;
    if (Q > 0 && c->N > 0) {
        KA_CUDA(cudaMemcpyAsync(out_broker, c->d_out.p, (size_t)Q * S * 4, cudaMemcpyDeviceToHost, s));
        if (out_len) KA_CUDA(cudaMemcpyAsync(out_len, c->d_out_len.p, (size_t)Q * 4, cudaMemcpyDeviceToHost, s));
    }
    if (c->timing) { KA_CUDA(cudaEventRecord(c->ev[5], s)); c->ev_valid = true; }
    c->last_stream = s;
    c->pending_status = true;
    ka_status local;
    rc = finish_status(c, s, st ? st : &local);
    c->last_part_id = nullptr;
    c->last_part_off = nullptr;
    return rc;
}

}  // extern "C"
