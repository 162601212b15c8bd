// api.cu — the C ABI of libotters_b200.so (include/otters_b200.h): contexts, device-resident stores,
// query orchestration.  Everything that touches rows runs in the CUDA kernels of scan_kernel.cuh / select.cu
// / meta.cu / store.cu; there is no CPU fallback.
#include <errno.h>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <memory>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "internal.h"

namespace otters {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

static inline double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline uint64_t round_up(uint64_t x, uint64_t m) { return (x + m - 1) / m * m; }
static inline uint64_t pow2_at_least(uint64_t x) {
    uint64_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

}  // namespace otters

using namespace otters;

namespace otters {

struct FusedFilter {  // device-side lowered filter evaluated inside the scan kernel
    const DevLeaf* leaves = nullptr;
    const uint32_t* clause_off = nullptr;
    uint32_t n_clauses = 0, n_leaves = 0;
    const uint32_t* chunk_keep = nullptr;  // null: the scan kernel prunes chunks itself (lazy pruning)
    uint32_t chunk_size = 1;
    uint32_t n_chunks = 0;
    static size_t smem_bytes_for(uint32_t n_leaves, uint32_t n_clauses) {
        return ((size_t)n_leaves * sizeof(DevLeaf) + ((size_t)n_clauses + 1) * 4 + 127) / 128 * 128;
    }
    size_t smem_bytes() const { return smem_bytes_for(n_leaves, n_clauses); }
};
constexpr size_t kMaxFusedFilterBytes = 16 * 1024;

struct QueryRun {
    uint32_t result_list = 0;  // index into ctx->d_list holding the final ordered candidates
    uint64_t k_eff = 0;
    bool big = false;          // result lives in ctx->d_emit-sized list (emit-all path)
    bool prefetched = false;   // header + candidates already sit in ctx->h_result (batched path)
    bool zero_copy = false;    // the selection writes header + candidates into ctx->h_zc (mapped host memory)
};

// A query that was enqueued by otters_query_submit and has not been waited for yet (one per lane).
struct Pending {
    bool active = false;
    uint64_t ticket = 0;
    bool meta = false;
    otters_metastore* ms = nullptr;
    QueryRun run;
    bool take_max = true;
    bool scan = false;           // a scan was enqueued (else: only statistics come back)
    bool fetched = false;        // the blocking path already ran: results sit in h_idx / h_score / h_qid
    uint32_t nq = 0;
    bool has_filter = false;
    uint64_t leaf_zm = 0, leaf_row = 0, n_leaves = 0;  // algorithmic metadata bytes per chunk / per evaluated row
    double t_submit = 0.0;
    std::vector<uint64_t> h_idx;
    std::vector<float> h_score;
    std::vector<uint32_t> h_qid;
    otters_query_stats stats{};
};

}  // namespace otters

// =================================================================================================
// context
// =================================================================================================
struct otters_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    size_t smem_optin = 0;
    size_t smem_per_sm = 0;
    otters_scan_tuning tuning{};
    otters_last_work last{};

    // device scratch
    float* d_query = nullptr;           // the staged queries inside d_io
    size_t d_query_off = 0;
    // one 64-byte control block, zeroed by a single memset at the start of every query:
    //   +0 unit counter, +4 emit count, +8 list counts[2], +16 tau, +24 rows scored, +32 stats[4]
    uint8_t* d_ctrl = nullptr;
    uint32_t* d_counter = nullptr;      // [0] unit counter, [1] emit count
    uint32_t* d_list_count = nullptr;   // [2]
    uint64_t* d_tau = nullptr;
    unsigned long long* d_rows_scored = nullptr;
    unsigned long long* d_stats = nullptr;  // [0] evaluated chunks, [1] vectors_compared
    uint64_t* d_cta_keys = nullptr;     // [grid_max][kMaxFusedK]
    uint32_t* d_cta_counts = nullptr;
    uint32_t grid_max = 0;
    uint8_t* d_list_raw[2] = {nullptr, nullptr};  // ResultHeader + entries
    Cand* d_list[2] = {nullptr, nullptr};  // running / result lists (entries)
    size_t list_cap = 0;                // entries of d_list[0] (d_list[1] always holds kMaxFusedK)
    uint64_t* d_scratch_keys = nullptr;
    uint32_t* d_scratch_src = nullptr;
    uint32_t scratch_elems = 0;
    uint32_t scan_smem_configured[6] = {0, 0, 0, 0, 0, 0};
    uint32_t planner_smem_configured[3] = {0, 0, 0};
    uint32_t* d_mask = nullptr;         // uploaded VecStore row mask
    size_t d_mask_words = 0;
    Cand* d_emit = nullptr;             // emit-all path
    size_t emit_cap = 0;
    // batched (tensor-core) path
    float* d_qh = nullptr;              // hi / lo tf32 split of the staged queries, padded to 256-query tiles
    float* d_ql = nullptr;
    size_t d_qh_floats = 0, d_ql_floats = 0;
    uint16_t* d_qb = nullptr;           // bf16 copy of the staged queries (bf16 rung), padded to 256-query tiles
    size_t d_qb_elems = 0;
    float* d_qscal = nullptr;           // per-query scalar: 1/|q| (cosine) or |q|^2 (euclidean)
    size_t d_qscal_floats = 0;
    uint32_t* d_cta_qids = nullptr;     // [grid_max][kMaxFusedK]
    unsigned long long* d_batch_info = nullptr;  // d_ctrl + 64: [0] shared threshold key, [1] flags | max error bits << 32, [2] best excluded
    uint32_t batch_smem_configured[6] = {0, 0, 0, 0, 0, 0};
    // fused peer exchange requested by otters_query_exchange for the query being enqueued
    bool ex_active = false;
    bool ex_published = false;  // a kernel that publishes this rank's records for the current query has been launched
    uint32_t ex_world = 0, ex_rank = 0, ex_kmax = 0, ex_k = 0, ex_seq = 0, ex_slot = 0;
    otters_topk_record* ex_records[kMaxPeers] = {};
    uint32_t* ex_flags[kMaxPeers] = {};

    // Per-query input image: [control block 256 B | lowered filter | queries], assembled in pinned memory and sent with
    // ONE host-to-device copy per query (it also resets the control block, so there is no memset on the query path).
    // Two pinned slots alternate so the host can assemble query i+1 while the copy of query i is still in flight.
    uint8_t* d_io = nullptr;
    size_t d_io_bytes = 0;
    uint8_t* h_io[2] = {nullptr, nullptr};
    size_t h_io_bytes[2] = {0, 0};
    cudaEvent_t ev_io[2]{};
    bool io_pending[2] = {false, false};
    int io_slot = 0;
    size_t io_used = 0;
    bool io_open = false;
    // pinned staging: row masks, results
    uint8_t* h_stage = nullptr;
    size_t h_stage_bytes = 0;
    uint8_t* h_result = nullptr;
    size_t h_result_bytes = 0;
    // zero-copy result area of the fused-selection path: mapped pinned memory the selection kernel writes the final
    // header + candidates into, so a blocking query ends with a stream sync instead of a D2H copy + sync
    uint8_t* h_zc = nullptr;
    uint8_t* d_zc = nullptr;     // device alias of h_zc
    bool want_host_result = false;  // the entry point being served will read the result on the host

    cudaEvent_t ev[8]{};
    bool timing = false;          // record the phase events for the query being enqueued
    bool timed_single = false;    // ev[3]/ev[4] bracket the single scan kernel of the last query
    bool timed_meta = false;      // ev[0]/ev[1] bracket the prune kernel of the last query
    bool timed_rowmask = false;   // ev[1]/ev[6] bracket the stand-alone row-mask kernel of the last query
    uint32_t last_dim = 0;        // of the last scan (for the algorithmic-bytes figure)
    uint32_t last_esz = 4;        // bytes per stored element of the last scan (4, or 2 for a bf16 store)
    int32_t last_metric = 0;

    // per-query MetaStore scratch (stores are immutable after build: everything a query writes lives in the context
    // that runs it, so two lanes can search the same store at the same time)
    uint32_t* d_chunk_keep = nullptr;   // stand-alone prune kernel: one bit per chunk
    size_t chunk_keep_words = 0;
    uint8_t* d_gather = nullptr;        // result-column gather staging
    size_t d_gather_bytes = 0;
    uint32_t* d_meta_mask = nullptr;    // stand-alone row-mask kernel: one bit per row
    size_t meta_mask_words = 0;
    FusedFilter cur_filter;             // where the last lowered filter lives on the device

    // lanes: queries in flight (otters_query_submit / otters_query_wait).  lane[0] is this context itself, lane[1..] are
    // child contexts with their own stream and scratch, created on first use.
    bool pipelined = false;             // the query being planned was submitted on a lane (otters_query_submit)
    otters_ctx* parent = nullptr;
    otters_ctx* lane[kMaxLanes] = {};
    uint32_t next_lane = 0;
    uint64_t tickets = 0;
    Pending pend;
    // what otters_ctx_last_work reports after otters_query_wait: the waited query's counters (lane 0 shares `last` with the
    // query it has in flight, so the report is kept apart); a blocking call on this context switches back to `last`
    otters_last_work reported{};
    bool report_waited = false;
};

namespace otters {

static int ensure_stage(otters_ctx* c, size_t bytes) {
    if (bytes <= c->h_stage_bytes) return OTTERS_OK;
    if (c->h_stage) cudaFreeHost(c->h_stage);
    c->h_stage = nullptr;
    c->h_stage_bytes = 0;
    size_t nb = std::max<size_t>(round_up(bytes, 4096), 1 << 16);
    OTTERS_CUDA(cudaMallocHost((void**)&c->h_stage, nb));
    c->h_stage_bytes = nb;
    return OTTERS_OK;
}

// elapsed time between two recorded, completed events; never leaves a CUDA error behind
static bool elapsed_ms(cudaEvent_t a, cudaEvent_t b, float* ms) {
    if (cudaEventElapsedTime(ms, a, b) == cudaSuccess) return true;
    cudaGetLastError();
    *ms = 0.f;
    return false;
}

static int ensure_pinned(uint8_t** ptr, size_t* cap, size_t bytes) {
    if (bytes <= *cap) return OTTERS_OK;
    if (*ptr) cudaFreeHost(*ptr);
    *ptr = nullptr;
    *cap = 0;
    size_t nb = std::max<size_t>(round_up(bytes, 4096), 1 << 14);
    OTTERS_CUDA(cudaMallocHost((void**)ptr, nb));
    *cap = nb;
    return OTTERS_OK;
}

// resets the scan/selection state but keeps the prune kernel's stats (fallback from the batched path)
static int reset_scan_state(otters_ctx* c) {
    OTTERS_CUDA(cudaMemsetAsync(c->d_ctrl, 0, 32, c->stream));
    OTTERS_CUDA(cudaMemsetAsync(c->d_ctrl + 64, 0, 64, c->stream));
    return OTTERS_OK;
}

static void bind_io(otters_ctx* c) {
    c->d_ctrl = c->d_io;
    c->d_counter = reinterpret_cast<uint32_t*>(c->d_ctrl);
    c->d_list_count = reinterpret_cast<uint32_t*>(c->d_ctrl + 8);
    c->d_tau = reinterpret_cast<uint64_t*>(c->d_ctrl + 16);
    c->d_rows_scored = reinterpret_cast<unsigned long long*>(c->d_ctrl + 24);
    c->d_stats = reinterpret_cast<unsigned long long*>(c->d_ctrl + 32);
    c->d_batch_info = reinterpret_cast<unsigned long long*>(c->d_ctrl + 64);
}

static int io_reserve(otters_ctx* c, size_t total) {
    const int sl = c->io_slot;
    if (total > c->h_io_bytes[sl]) {
        const size_t nb = std::max<size_t>(round_up(total + total / 2, 4096), 1 << 16);
        uint8_t* nh = nullptr;
        OTTERS_CUDA(cudaMallocHost((void**)&nh, nb));
        if (c->h_io[sl]) {
            memcpy(nh, c->h_io[sl], c->io_used);
            cudaFreeHost(c->h_io[sl]);
        }
        c->h_io[sl] = nh;
        c->h_io_bytes[sl] = nb;
    }
    if (total > c->d_io_bytes) {
        OTTERS_CUDA(cudaStreamSynchronize(c->stream));  // earlier queries still use the old buffer
        const size_t nb = std::max<size_t>(round_up(total + total / 2, 4096), 1 << 16);
        uint8_t* nd = nullptr;
        if (cudaMalloc((void**)&nd, nb) != cudaSuccess) return fail(OTTERS_ERR_NOMEM, "device allocation for query inputs failed");
        if (c->d_io) {
            OTTERS_CUDA(cudaMemcpy(nd, c->d_io, 256, cudaMemcpyDeviceToDevice));
            cudaFree(c->d_io);
        }
        c->d_io = nd;
        c->d_io_bytes = nb;
        bind_io(c);
    }
    return OTTERS_OK;
}

// Opens the input image of a new query: takes the next pinned slot (waiting for the copy that last used it) and
// zeroes the control block image (unit counter, list counts, tau, rows scored, stats, batch info).
static int begin_query(otters_ctx* c) {
    const int sl = c->io_slot;
    if (c->io_pending[sl]) {
        OTTERS_CUDA(cudaEventSynchronize(c->ev_io[sl]));
        c->io_pending[sl] = false;
    }
    c->io_used = 0;
    int rc = io_reserve(c, 256);
    if (rc) return rc;
    memset(c->h_io[sl], 0, 256);
    c->io_used = 256;
    c->io_open = true;
    c->last = otters_last_work{};
    c->timed_single = c->timed_meta = c->timed_rowmask = false;
    c->timing = c->tuning.timing == 1;
    c->want_host_result = false;
    if (!c->pipelined) c->report_waited = false;  // a blocking call: otters_ctx_last_work reports it
    return OTTERS_OK;
}

// appends `bytes` (256-byte aligned) to the input image; returns the host pointer to fill and the device address
static int io_push(otters_ctx* c, size_t bytes, uint8_t** host, size_t* dev_off) {
    const size_t off = round_up(c->io_used, 256);
    int rc = io_reserve(c, off + bytes + 256);
    if (rc) return rc;
    *host = c->h_io[c->io_slot] + off;
    *dev_off = off;
    c->io_used = off + bytes;
    return OTTERS_OK;
}

// sends the input image with one copy; kernels enqueued afterwards see the control block reset and all inputs
static int io_flush(otters_ctx* c) {
    if (!c->io_open) return OTTERS_OK;
    const int sl = c->io_slot;
    OTTERS_CUDA(cudaMemcpyAsync(c->d_io, c->h_io[sl], c->io_used, cudaMemcpyHostToDevice, c->stream));
    c->last.h2d_bytes += c->io_used;
    OTTERS_CUDA(cudaEventRecord(c->ev_io[sl], c->stream));
    c->io_pending[sl] = true;
    c->io_slot ^= 1;
    c->io_open = false;
    c->d_query = reinterpret_cast<float*>(c->d_io + c->d_query_off);  // d_io may have moved while the image grew
    return OTTERS_OK;
}

static inline void rec_event(otters_ctx* c, int i) {
    if (c->timing) cudaEventRecord(c->ev[i], c->stream);
}

static int alloc_list(otters_ctx* c, int i, size_t entries) {
    if (c->d_list_raw[i]) cudaFree(c->d_list_raw[i]);
    c->d_list_raw[i] = nullptr;
    c->d_list[i] = nullptr;
    if (cudaMalloc((void**)&c->d_list_raw[i], sizeof(ResultHeader) + entries * sizeof(Cand)) != cudaSuccess)
        return fail(OTTERS_ERR_NOMEM, "device allocation for result lists failed");
    c->d_list[i] = reinterpret_cast<Cand*>(c->d_list_raw[i] + sizeof(ResultHeader));
    return OTTERS_OK;
}

static int ensure_list0(otters_ctx* c, size_t entries) {
    entries = std::max<size_t>(entries, kMaxFusedK);
    if (entries <= c->list_cap && c->d_list[0]) return OTTERS_OK;
    OTTERS_CUDA(cudaStreamSynchronize(c->stream));
    int rc = alloc_list(c, 0, entries);
    if (rc) return rc;
    c->list_cap = entries;
    return OTTERS_OK;
}

static inline ResultHeader* list_hdr(otters_ctx* c, int i) { return reinterpret_cast<ResultHeader*>(c->d_list_raw[i]); }

template <typename T>
static int ensure_dev(T** ptr, size_t* cap, size_t need, cudaStream_t s) {
    if (need <= *cap && *ptr) return OTTERS_OK;
    if (*ptr) {
        OTTERS_CUDA(cudaStreamSynchronize(s));
        cudaFree(*ptr);
        *ptr = nullptr;
        *cap = 0;
    }
    size_t n = std::max<size_t>(need, 16);
    OTTERS_CUDA(cudaMalloc((void**)ptr, n * sizeof(T)));
    *cap = n;
    return OTTERS_OK;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- device vector storage shared by VecStore and MetaStore ---------------------------------------
constexpr uint32_t kSinglePassBackoff = 16;

struct VecStorage {
    otters_ctx* ctx = nullptr;
    uint32_t single_pass_backoff = 0;  // batches left that skip the single-pass tf32 selection (see run_queries)
    uint32_t dim = 0;
    uint32_t pitch = 0;  // floats per stored row: dim rounded up to 4 (16-byte rows for TMA bulk copies)
    uint64_t n = 0, cap = 0;
    float* d_rows = nullptr;
    float* d_inv = nullptr;
    uint32_t* d_minv_bits = nullptr;  // smallest positive inverse row norm (float bits), for the batched path's error bound
    bool minv_valid = false;
    // bf16 SHADOW of the rows for K2's bf16 rung: selection only (scores are always re-computed from the fp32 rows), built on
    // the first query batch that wants it when memory allows (ensure_bf16_shadow), dropped whenever the rows change
    uint16_t* d_rows_h = nullptr;
    uint64_t pitch_h = 0, h_cap_rows = 0, h_declined_n = 0;
    bool h_valid = false;
    uint32_t bf16_backoff = 0;         // batches left that skip the bf16 rung (its certificate failed on this store)

    bool half = false;  // rows are bf16 (OTTERS_VECTORS_FMT_BF16): d_rows then points at uint16_t elements, pitch counts bf16
    size_t esz() const { return half ? 2 : 4; }
    void set_format(uint32_t d, bool h) {
        dim = d;
        half = h;
        pitch = (uint32_t)round_up(std::max<uint32_t>(d, 1), h ? 8 : 4);  // 16-byte rows for the bulk copies
    }
    uint8_t* row_ptr(uint64_t row) const { return reinterpret_cast<uint8_t*>(d_rows) + row * pitch * esz(); }

    int reserve(uint64_t want) {
        if (want <= cap) return OTTERS_OK;
        uint64_t ncap = std::max<uint64_t>(want, cap ? cap + cap / 2 : 0);
        if (ncap == want && cap != 0 && want < cap * 2) ncap = want;
        float *nr = nullptr, *ni = nullptr;
        if (cudaMalloc((void**)&nr, std::max<uint64_t>(ncap * pitch * esz(), 16)) != cudaSuccess)
            return fail(OTTERS_ERR_NOMEM, "device allocation for vectors failed");
        if (cudaMalloc((void**)&ni, std::max<uint64_t>(ncap, 4) * sizeof(float)) != cudaSuccess) {
            cudaFree(nr);
            return fail(OTTERS_ERR_NOMEM, "device allocation for inverse norms failed");
        }
        if (n) {
            OTTERS_CUDA(cudaMemcpyAsync(nr, d_rows, n * pitch * esz(), cudaMemcpyDeviceToDevice, ctx->stream));
            OTTERS_CUDA(cudaMemcpyAsync(ni, d_inv, n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        }
        OTTERS_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_rows);
        cudaFree(d_inv);
        d_rows = nr;
        d_inv = ni;
        cap = ncap;
        return OTTERS_OK;
    }
    // bf16 stores: rows arrive as fp32 (host, device or generated), pass through an fp32 staging slab and are rounded to
    // nearest even on the device; the inverse norms are those of the ROUNDED rows (what the store holds is what is scored)
    template <typename Fill>
    int add_half(uint64_t cnt, Fill fill) {
        const uint64_t pitch4 = round_up(std::max<uint32_t>(dim, 1), 4);
        const uint64_t slab = std::max<uint64_t>(1, std::min<uint64_t>(cnt, ((uint64_t)256 << 20) / (pitch4 * 4)));
        float* tmp = nullptr;
        if (cudaMalloc((void**)&tmp, slab * pitch4 * 4) != cudaSuccess) return fail(OTTERS_ERR_NOMEM, "device allocation for the staging slab failed");
        int rc = OTTERS_OK;
        for (uint64_t done = 0; done < cnt && !rc; done += slab) {
            const uint64_t m = std::min(slab, cnt - done);
            rc = fill(tmp, pitch4, done, m);
            if (!rc) rc = launch_convert_bf16(tmp, pitch4, m, dim, reinterpret_cast<uint16_t*>(d_rows) + (n + done) * pitch, pitch, m, ctx->stream);
        }
        if (!rc) rc = launch_inv_norms_bf16(reinterpret_cast<const uint16_t*>(d_rows), pitch, dim, n, cnt, d_inv, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        cudaFree(tmp);
        return rc;
    }
    int add(const float* rows, uint64_t cnt, cudaMemcpyKind kind) {
        if (cnt == 0) return OTTERS_OK;
        if (n + cnt >= 0xFFFFFFF0ull) return fail(OTTERS_ERR_UNSUPPORTED, "a store shard is limited to 2^32-16 rows");
        int rc = reserve(n + cnt);
        if (rc) return rc;
        // large host inputs are pinned in place for the duration of the copy (cudaHostRegister): the H2D copy then runs at
        // PCIe rate instead of being staged through the driver's bounce buffers; silently skipped when the range cannot be pinned
        struct Pin {
            void* p = nullptr;
            ~Pin() {
                if (p) cudaHostUnregister(p);
            }
        } pin;
        const size_t in_bytes = (size_t)cnt * dim * sizeof(float);
        if (kind == cudaMemcpyHostToDevice && in_bytes >= ((size_t)64 << 20) && !getenv("OTTERS_NO_PIN")) {
            if (cudaHostRegister((void*)rows, in_bytes, cudaHostRegisterDefault) == cudaSuccess) pin.p = (void*)rows;
            else cudaGetLastError();
        }
        if (half) {
            rc = add_half(cnt, [&](float* tmp, uint64_t pitch4, uint64_t first, uint64_t m) -> int {
                OTTERS_CUDA(cudaMemcpy2DAsync(tmp, pitch4 * sizeof(float), rows + first * dim, dim * sizeof(float), dim * sizeof(float), m, kind,
                                              ctx->stream));
                return OTTERS_OK;
            });
            if (rc) return rc;
            n += cnt;
            minv_valid = false;
            h_valid = false;
            return OTTERS_OK;
        }
        if (pitch == dim) {
            OTTERS_CUDA(cudaMemcpyAsync(d_rows + n * pitch, rows, cnt * dim * sizeof(float), kind, ctx->stream));
        } else {
            OTTERS_CUDA(cudaMemsetAsync(d_rows + n * pitch, 0, cnt * pitch * sizeof(float), ctx->stream));
            OTTERS_CUDA(cudaMemcpy2DAsync(d_rows + n * pitch, pitch * sizeof(float), rows, dim * sizeof(float),
                                          dim * sizeof(float), cnt, kind, ctx->stream));
        }
        rc = launch_inv_norms(d_rows, pitch, dim, n, cnt, d_inv, ctx->stream);
        if (rc) return rc;
        OTTERS_CUDA(cudaStreamSynchronize(ctx->stream));  // the caller's buffer may be reused after return
        n += cnt;
        minv_valid = false;
        h_valid = false;
        return OTTERS_OK;
    }
    int add_synth(ShardMap gen_map, uint64_t cnt, uint64_t seed) {
        if (cnt == 0) return OTTERS_OK;
        if (n + cnt >= 0xFFFFFFF0ull) return fail(OTTERS_ERR_UNSUPPORTED, "a store shard is limited to 2^32-16 rows");
        int rc = reserve(n + cnt);
        if (rc) return rc;
        if (half) {
            // generated row ids continue where the store ends: slab `first` rows further into the map
            rc = add_half(cnt, [&](float* tmp, uint64_t pitch4, uint64_t first, uint64_t m) -> int {
                ShardMap mm = gen_map;
                return launch_synth_fill_at(tmp, pitch4, dim, mm, first, m, seed, ctx->stream);
            });
            if (rc) return rc;
            n += cnt;
            minv_valid = false;
            h_valid = false;
            return OTTERS_OK;
        }
        rc = launch_synth_fill(d_rows, pitch, dim, n, gen_map, cnt, seed, ctx->stream);
        if (rc) return rc;
        rc = launch_inv_norms(d_rows, pitch, dim, n, cnt, d_inv, ctx->stream);
        if (rc) return rc;
        OTTERS_CUDA(cudaStreamSynchronize(ctx->stream));
        n += cnt;
        minv_valid = false;
        h_valid = false;
        return OTTERS_OK;
    }
    // overwrites individual rows (host data, dim floats each) and recomputes their inverse norms
    int set_rows(const uint64_t* rows, const float* data, uint64_t cnt) {
        for (uint64_t i = 0; i < cnt; ++i)
            if (rows[i] >= n) return fail(OTTERS_ERR_INVALID, "row index out of bounds");
        if (half) {
            const uint64_t pitch4 = round_up(std::max<uint32_t>(dim, 1), 4);
            float* tmp = nullptr;
            if (cnt && cudaMalloc((void**)&tmp, cnt * pitch4 * 4) != cudaSuccess) return fail(OTTERS_ERR_NOMEM, "device allocation for the staging slab failed");
            int rc = OTTERS_OK;
            if (cnt && cudaMemcpy2DAsync(tmp, pitch4 * sizeof(float), data, dim * sizeof(float), dim * sizeof(float), cnt, cudaMemcpyHostToDevice,
                                         ctx->stream) != cudaSuccess)
                rc = fail(OTTERS_ERR_CUDA, "copy of the replacement rows failed");
            for (uint64_t i = 0; i < cnt && !rc; ++i) {
                rc = launch_convert_bf16(tmp + i * pitch4, pitch4, 1, dim, reinterpret_cast<uint16_t*>(d_rows) + rows[i] * pitch, pitch, 1, ctx->stream);
                if (!rc) rc = launch_inv_norms_bf16(reinterpret_cast<const uint16_t*>(d_rows), pitch, dim, rows[i], 1, d_inv, ctx->stream);
            }
            cudaStreamSynchronize(ctx->stream);
            cudaFree(tmp);
            minv_valid = false;
            h_valid = false;
            return rc;
        }
        for (uint64_t i = 0; i < cnt; ++i) {
            OTTERS_CUDA(cudaMemcpyAsync(d_rows + rows[i] * pitch, data + i * dim, (size_t)dim * sizeof(float), cudaMemcpyHostToDevice,
                                        ctx->stream));
            int rc = launch_inv_norms(d_rows, pitch, dim, rows[i], 1, d_inv, ctx->stream);
            if (rc) return rc;
        }
        OTTERS_CUDA(cudaStreamSynchronize(ctx->stream));
        minv_valid = false;
        h_valid = false;
        return OTTERS_OK;
    }
    void release() {
        cudaFree(d_rows);
        cudaFree(d_inv);
        cudaFree(d_minv_bits);
        cudaFree(d_rows_h);
        d_rows = d_inv = nullptr;
        d_minv_bits = nullptr;
        d_rows_h = nullptr;
        h_cap_rows = h_declined_n = 0;
        minv_valid = false;
        h_valid = false;
        n = cap = 0;
    }
};

// ---- scan planning ---------------------------------------------------------------------------------
struct ScanPlan {
    ScanLaunch launch;
    uint32_t off_filter;
    uint32_t kc, nkc, pitch_s, slots, unit_rows, unit_small, n_big, n_units, cap;
    uint32_t off_query, off_warps, warp_bytes, off_w_rows, off_w_info, off_w_inv, off_w_list, off_w_slots;
    uint32_t planners, off_ring;  // planner front-end (scan_planner.cu): planner warps per CTA (0 = autonomous warps)
};

static int plan_scan(const otters_ctx* c, uint32_t dim_pad, uint64_t n_rows, uint32_t k_fused, size_t filter_bytes, bool fuse_select,
                     ScanPlan* out, bool half = false) {
    ScanPlan pl{};
    const otters_scan_tuning& t = c->tuning;
    pl.cap = k_fused ? (uint32_t)pow2_at_least(std::max<uint32_t>(2 * k_fused, 64)) : 0;
    const uint32_t hdr = (uint32_t)round_up(16 + (uint64_t)pl.cap * 8, 128);
    pl.off_query = hdr;
    pl.off_filter = (uint32_t)round_up((uint64_t)hdr + (uint64_t)dim_pad * 4, 128);
    // front-end: planner warps feeding worker warps (fused top-k only), or autonomous warps
    // automatic choice (profiles/r1_planner_ab.log): the planner front-end wins where rows are wide or no predicate has to
    // be evaluated (10Mx768 unfiltered 4.07 vs 4.18 ms, 5Mx1536 filtered 2.17 vs 2.26 ms) and loses on narrow filtered rows
    // (10Mx128: 0.79 vs 0.50 ms), where the per-tile ring handshake and the planners' metadata round trips dominate
    const uint32_t row_words = half ? dim_pad / 2 : dim_pad;  // row length in 4-byte words: what the front-end choices depend on
    const bool planner_auto = filter_bytes == 0 ? row_words >= 256 : row_words >= 1024;
    const bool planner = k_fused && t.scan_mode != 1 && (t.scan_mode == 2 || planner_auto);
    pl.off_ring = (uint32_t)(pl.off_filter + filter_bytes);
    pl.off_warps = pl.off_ring + (planner ? kPlannerRingBytes : 0);
    // planner warps per CTA: a filtered unit costs a planner one memory round trip (~2 µs under load); narrow rows are
    // consumed faster, so they need more planners to stay ahead of the workers
    if (planner) pl.planners = t.planners ? std::min<uint32_t>(t.planners, 4) : (filter_bytes ? (row_words <= 256 ? 4 : 2) : 1);
    static const size_t margin = getenv("OTTERS_SMEM_MARGIN") ? (size_t)atoi(getenv("OTTERS_SMEM_MARGIN")) : 128;
    // shared memory of one CTA: the opt-in maximum, or an equal share of the SM (minus the 1 KB the system reserves per
    // CTA) when several CTAs are to be co-resident per SM (ctas_per_sm)
    const uint32_t cps = t.ctas_per_sm ? std::min<uint32_t>(t.ctas_per_sm, 4) : 1;
    size_t cta_smem = c->smem_optin;
    if (cps > 1) cta_smem = std::min<size_t>(cta_smem, c->smem_per_sm / cps - 1024);
    const size_t budget = cta_smem > 2048 ? cta_smem - margin : 0;  // (the kernels' static shared memory is < 64 bytes)
    if (pl.off_warps + 4096 > budget) return fail(OTTERS_ERR_UNSUPPORTED, "vector dimension too large for the scan kernel");

    uint32_t dim8 = (uint32_t)round_up(dim_pad, 8);
    // Measured on B200 (profiles/r1_sweep_*.log): many autonomous warps with one small slot each beat
    // fewer warps with deep rings — 12-16 warps x 1 slot x 256 columns reads 7.1-7.3 TB/s at dim 768.
    // (bf16 rows: twice the columns per slot, the same 1 KB per staged row)
    uint32_t kc_target = t.kc_floats ? (uint32_t)round_up(t.kc_floats, 8) : (half ? 512 : 256);
    uint32_t nkc = (dim_pad + kc_target - 1) / kc_target;
    if (nkc < 1) nkc = 1;
    uint32_t kc = (uint32_t)round_up((dim8 + nkc - 1) / nkc, 8);
    nkc = (dim_pad + kc - 1) / kc;
    pl.kc = kc;
    pl.nkc = nkc;
    pl.pitch_s = (uint32_t)round_up(half ? kc / 2 : kc, 32) + 8;
    const uint32_t slot_bytes = kTileRows * pl.pitch_s * 4;

    auto warp_bytes_for = [&](uint32_t S, uint32_t* o_rows, uint32_t* o_info, uint32_t* o_inv, uint32_t* o_list, uint32_t* o_slots) {
        uint32_t o = 0;
        o += S * 8;
        *o_rows = o;
        o += S * kTileRows * 4;
        *o_info = o;
        o += (uint32_t)round_up(S * 4, 16);
        *o_inv = o;
        o += S * kTileRows * 4;
        *o_list = o;
        o += 2 * kMaxUnitRows;  // two row lists: the next unit's is built while the current one streams
        o = (uint32_t)round_up(o, 128);
        *o_slots = o;
        o += S * slot_bytes;
        return (uint32_t)round_up(o, 128);
    };
    uint32_t dummy[5];
    const size_t avail = budget - pl.off_warps;
    uint32_t total_slots = (uint32_t)(avail / (slot_bytes + 256));
    if (total_slots < 1) return fail(OTTERS_ERR_UNSUPPORTED, "vector dimension too large for the scan kernel");
    uint32_t S = t.slots_per_warp ? t.slots_per_warp : (total_slots >= 32 ? 2 : 1);
    if (planner) S = 1;
    uint32_t W = t.warps_per_cta ? t.warps_per_cta : std::min<uint32_t>(total_slots / S, 16);
    if (W < 1) W = 1;
    if (W > 16) W = 16;
    while (W > 1 && (size_t)W * warp_bytes_for(S, dummy, dummy + 1, dummy + 2, dummy + 3, dummy + 4) > avail) --W;
    while (S > 1 && (size_t)W * warp_bytes_for(S, dummy, dummy + 1, dummy + 2, dummy + 3, dummy + 4) > avail) --S;
    if ((size_t)W * warp_bytes_for(S, dummy, dummy + 1, dummy + 2, dummy + 3, dummy + 4) > avail)
        return fail(OTTERS_ERR_UNSUPPORTED, "scan tuning does not fit in shared memory");
    pl.slots = S;
    pl.warp_bytes = warp_bytes_for(S, &pl.off_w_rows, &pl.off_w_info, &pl.off_w_inv, &pl.off_w_list, &pl.off_w_slots);

    uint32_t ctas_per_sm = cps;
    uint32_t grid = (uint32_t)c->sm_count * ctas_per_sm;
    uint32_t unit_rows = t.unit_rows ? t.unit_rows : kMaxUnitRows;
    if (unit_rows != 32 && unit_rows != 64 && unit_rows != 128) unit_rows = kMaxUnitRows;
    // small stores (e.g. one shard of a row-sharded search) need finer units, or the last units of the dynamic schedule
    // leave most warps idle: measured on a 1.25M x 768 shard, 32-row units scan in 0.293 ms against 0.330 ms for 128-row
    // units (profiles/r1_unit_rows_shard.log); large stores keep 128-row units (fewer unit boundaries)
    if (!t.unit_rows) {
        // planner front-end: a unit is spread over all warps of its CTA, so only the per-CTA unit count matters.
        // Blocking calls want ~16 units per warp so that the tail of the dynamic schedule stays short; queries submitted on
        // the lanes keep larger units — the tail of one query's scan is filled by the head of the next one's, so fewer unit
        // boundaries win (1.25M x 768 shard: 3598 vs 3478 queries/s) — as long as every warp still gets a couple of units
        // (a 100k-row store cut into 128-row units would leave two thirds of the warps without work).
        const uint64_t per = c->pipelined ? 2 : 16;
        const uint64_t want = planner ? (uint64_t)grid * per : (uint64_t)grid * W * per;
        while (unit_rows > 32 && n_rows / unit_rows < want) unit_rows >>= 1;
    }
    pl.unit_rows = unit_rows;
    // guided schedule: the last ~2 units' worth of rows per warp (per CTA with the planner front-end) are cut into small
    // units, so the warps finish within one small unit of each other instead of one big one (a 128-row unit of 768-d rows
    // is 130 µs of one warp's bandwidth)
    pl.unit_small = t.unit_rows ? unit_rows : (unit_rows > 32 ? 32 : 16);
    {
        const uint64_t tail_rows = std::min<uint64_t>(n_rows, (uint64_t)grid * (planner ? 4 : W * 2) * unit_rows);
        pl.n_big = (uint32_t)((n_rows - tail_rows) / unit_rows);
        const uint64_t rest = n_rows - (uint64_t)pl.n_big * unit_rows;
        pl.n_units = pl.n_big + (uint32_t)((rest + pl.unit_small - 1) / pl.unit_small);
    }
    uint32_t need_ctas = planner ? pl.n_units : (pl.n_units + W - 1) / W;
    if (need_ctas < 1) need_ctas = 1;
    if (grid > need_ctas) grid = need_ctas;
    if (grid > c->grid_max) grid = c->grid_max;
    pl.launch.grid = grid;
    pl.launch.block = (W + pl.planners) * 32;
    pl.launch.smem_bytes = pl.off_warps + W * pl.warp_bytes;
    // the last CTA runs the selection in the same shared memory (select_body.cuh)
    if (fuse_select && pl.launch.smem_bytes < kSelectSmemBytes) pl.launch.smem_bytes = kSelectSmemBytes;
    *out = pl;
    return OTTERS_OK;
}

static void fill_exchange(const otters_ctx* c, SelectParams* se) {
    if (!c->ex_active) return;
    se->ex_world = c->ex_world;
    se->ex_rank = c->ex_rank;
    se->ex_kmax = c->ex_kmax;
    se->ex_seq = c->ex_seq;
    se->ex_slot = c->ex_slot;
    se->ex_k = c->ex_k;
    for (uint32_t i = 0; i < c->ex_world; ++i) {
        se->ex_records[i] = c->ex_records[i];
        se->ex_flags[i] = c->ex_flags[i];
    }
    se->records = nullptr;
}

// ---- the query core: scan + select for every query of the batch --------------------------------------
static float host_inv_norm(const float* v, uint32_t dim) {
    // src/vec.rs:390-397: serial f32 sum of squares, sqrt, reciprocal (0 for a zero vector).  The host code is
    // built with -ffp-contract=off and without fast-math, so the loop is neither contracted nor reassociated.
    float s = -0.0f;
    for (uint32_t i = 0; i < dim; ++i) {
        const float p = v[i] * v[i];
        s = s + p;
    }
    float norm = sqrtf(s);
    return norm != 0.0f ? 1.0f / norm : 0.0f;
}


// appends the queries, zero padded to the stored row pitch, to the input image of the current query
static int stage_queries(otters_ctx* c, const otters_vec_query* q, uint32_t dim_pad) {
    const size_t qfloats = (size_t)q->nq * dim_pad;
    uint8_t* h = nullptr;
    size_t off = 0;
    int rc = io_push(c, std::max<size_t>(qfloats * 4, 16), &h, &off);
    if (rc) return rc;
    float* hq = reinterpret_cast<float*>(h);
    if (dim_pad == q->dim) {
        memcpy(hq, q->queries, qfloats * 4);
    } else {
        for (uint32_t i = 0; i < q->nq; ++i) {
            memcpy(hq + (size_t)i * dim_pad, q->queries + (size_t)i * q->dim, (size_t)q->dim * 4);
            for (uint32_t j = q->dim; j < dim_pad; ++j) hq[(size_t)i * dim_pad + j] = 0.f;
        }
    }
    c->d_query = reinterpret_cast<float*>(c->d_io + off);  // io_push may have moved d_io: take the address afterwards
    c->d_query_off = off;
    return OTTERS_OK;
}

// ---- query batches on the tensor cores (K2, batched.cu) ---------------------------------------------------
// Selection runs on 3xTF32 tensor-core scores; every selected (row, query) pair is re-scored in the
// reference's exact arithmetic and the result is accepted only if no excluded pair can reach it:
//   (approximate cut score) +- delta must lie strictly outside the exact k-th score,
// with delta a bound on |tensor-core score - exact score|.  Otherwise the caller falls back to K1 per query.
static int store_min_inv_norm(otters_ctx* c, VecStorage* st) {
    if (st->minv_valid) return OTTERS_OK;
    if (!st->d_minv_bits && cudaMalloc((void**)&st->d_minv_bits, 16) != cudaSuccess)
        return fail(OTTERS_ERR_NOMEM, "device allocation failed");
    int rc = launch_min_inv_norm(st->d_inv, st->n, st->d_minv_bits, c->stream);
    if (rc) return rc;
    st->minv_valid = true;
    return OTTERS_OK;
}

// The bf16 shadow costs half of the store again; in automatic mode it is built only when that leaves the device half empty.
static int ensure_bf16_shadow(otters_ctx* c, VecStorage* st, bool force, bool* ok) {
    *ok = false;
    if (st->half || st->h_valid) {  // a bf16 store IS its own bf16 operand
        *ok = true;
        return OTTERS_OK;
    }
    const uint64_t pitch_h = round_up(st->dim, 8);
    const size_t bytes = (size_t)st->n * pitch_h * 2;
    if (st->d_rows_h && (st->h_cap_rows < st->n || st->pitch_h != pitch_h)) {
        OTTERS_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(st->d_rows_h);
        st->d_rows_h = nullptr;
        st->h_cap_rows = 0;
    }
    if (!st->d_rows_h) {
        if (!force) {
            if (st->h_declined_n == st->n) return OTTERS_OK;
            size_t fr = 0, tot = 0;
            if (cudaMemGetInfo(&fr, &tot) != cudaSuccess || fr < 2 * bytes + ((size_t)1 << 30) || fr - bytes < tot / 2) {
                cudaGetLastError();
                st->h_declined_n = st->n;
                return OTTERS_OK;
            }
        }
        if (cudaMalloc((void**)&st->d_rows_h, std::max<size_t>(bytes, 16)) != cudaSuccess) {
            cudaGetLastError();
            st->d_rows_h = nullptr;
            st->h_declined_n = st->n;
            return OTTERS_OK;
        }
        st->h_cap_rows = st->n;
        st->pitch_h = pitch_h;
    }
    int rc = launch_convert_bf16(st->d_rows, st->pitch, st->n, st->dim, st->d_rows_h, pitch_h, st->n, c->stream);
    if (rc) return rc;
    OTTERS_CUDA(cudaStreamSynchronize(c->stream));  // other contexts / lanes may use the shadow from their own streams
    st->h_valid = true;
    *ok = true;
    return OTTERS_OK;
}

static bool batch_eligible(const otters_ctx* c, const otters_vec_query* q, uint64_t n_rows, uint64_t k_eff) {
    const uint32_t mode = c->tuning.batch_mode;
    if (mode == 2 || q->nq < 2 || k_eff == 0 || k_eff > kMaxFusedK || n_rows == 0) return false;
    const uint32_t cap = (uint32_t)pow2_at_least(std::max<uint64_t>(2 * k_eff, 64));
    if (batch_smem_bytes(cap) > c->smem_optin) return false;
    if (mode == 1) return true;
    // one tensor-core pass costs about as much as a handful of streaming scans of the store
    return q->nq >= 8 && n_rows >= 4096;
}

static int run_batched(otters_ctx* c, VecStorage* st, const otters_vec_query* q,
                       const uint32_t* d_row_mask, uint32_t row_mask_words, otters_topk_record* d_records_out, ShardMap map,
                       const unsigned long long* stats_src, QueryRun* run, uint32_t passes, bool* accepted) {
    *accepted = false;
    c->last.batch_attempts += 1;
    cudaStream_t s = c->stream;
    const uint32_t dim_pad = st->pitch;
    const uint64_t k_eff = run->k_eff;
    // every CTA keeps a few more candidates than asked for: the best EXCLUDED pair then lies well below the k-th score,
    // so the certificate below rarely fails on a near-tie between the k-th and (k+1)-th pair
    const uint32_t k = (uint32_t)std::min<uint64_t>(kMaxFusedK, k_eff + std::max<uint64_t>(k_eff / 8, 16));
    const uint32_t cap = (uint32_t)pow2_at_least(std::max<uint32_t>(2 * k, 64));
    const uint32_t nq_pad = (uint32_t)round_up(q->nq, kBatchQueries);
    const bool take_max = q->take_type == OTTERS_TAKE_MAX;

    // per-query scalars on the device: 1/|q| exactly as the reference computes it (cosine; also used by the exact
    // re-scoring) and |q|^2 (euclidean); the batch's error bound follows from the largest norms
    int rc = ensure_dev(&c->d_qscal, &c->d_qscal_floats, (size_t)q->nq * 2, s);
    if (rc) return rc;
    rc = ensure_dev(&c->d_qh, &c->d_qh_floats, (size_t)nq_pad * dim_pad, s);
    if (rc) return rc;
    rc = ensure_dev(&c->d_ql, &c->d_ql_floats, (size_t)nq_pad * dim_pad, s);
    if (rc) return rc;
    float* d_qinv = c->d_qscal;
    float* d_qn2 = c->d_qscal + q->nq;
    uint32_t* d_excl = reinterpret_cast<uint32_t*>(c->d_ctrl + 80);
    float* d_delta = reinterpret_cast<float*>(c->d_ctrl + 84);
    uint32_t* d_qmax2 = reinterpret_cast<uint32_t*>(c->d_ctrl + 88);
    if (q->metric == OTTERS_METRIC_COSINE) {
        rc = launch_inv_norms(c->d_query, dim_pad, st->dim, 0, q->nq, d_qinv, s);
        if (rc) return rc;
    } else {
        rc = store_min_inv_norm(c, st);
        if (rc) return rc;
    }
    rc = launch_split_queries(c->d_query, q->nq, nq_pad, dim_pad, c->d_qh, c->d_ql, d_qn2, d_qmax2, s);
    if (rc) return rc;
    rc = launch_batch_delta(q->metric, st->dim, passes, d_qmax2, st->d_minv_bits, d_delta, s);
    if (rc) return rc;
    const uint32_t q_pitch_h = (uint32_t)round_up(st->dim, 8);
    if (passes == 2) {
        if (!st->half && !st->h_valid) return fail(OTTERS_ERR_INVALID, "the bf16 rung was requested without its shadow rows");
        rc = ensure_dev(&c->d_qb, &c->d_qb_elems, (size_t)nq_pad * q_pitch_h, s);
        if (rc) return rc;
        rc = launch_convert_bf16(c->d_query, dim_pad, q->nq, st->dim, c->d_qb, q_pitch_h, nq_pad, s);
        if (rc) return rc;
    }

    // the single-pass rung runs as CTA pairs (tcgen05 cta_group::2: each CTA stages half of the query tile, both CTAs' TMA loads
    // credit the leader's barrier directly, three k-blocks per stage): 2.59 vs 3.03 ms on 1M x 768 x 1024 queries, same box
    // (profiles/r2_k2_variants.log); the 3xTF32 rung stays on single CTAs (5.74 vs 5.99 ms, round 1)
    const uint32_t cg_auto = passes != 3 ? 2u : 1u;
    const uint32_t cg_want = c->tuning.batch_cta_group == 0 ? cg_auto : c->tuning.batch_cta_group;
    const uint32_t cg = (cg_want == 2 && c->sm_count >= 2) ? 2 : 1;
    const uint32_t tile_rows = kBatchRows * cg;
    const uint32_t n_rowtiles = (uint32_t)((st->n + tile_rows - 1) / tile_rows);
    const uint64_t n_tiles = (uint64_t)n_rowtiles * (nq_pad / kBatchQueries);
    BatchLaunch bl{};
    bl.vectors = st->d_rows;
    bl.n_rows = st->n;
    bl.pitch_g = st->pitch;
    bl.dim = st->dim;
    bl.dim_pad = dim_pad;
    bl.q_hi = c->d_qh;
    bl.q_lo = c->d_ql;
    bl.v_half = st->half ? reinterpret_cast<const uint16_t*>(st->d_rows) : st->d_rows_h;
    bl.pitch_h = st->half ? st->pitch : st->pitch_h;
    bl.q_half = c->d_qb;
    bl.q_pitch_h = q_pitch_h;
    bl.nq_pad = nq_pad;
    bl.cta_group = cg;
    bl.grid = (uint32_t)std::min<uint64_t>((uint64_t)(c->sm_count / cg), n_tiles) * cg;
    BatchParams bp{};
    bp.n_rows = (uint32_t)st->n;
    bp.nq = q->nq;
    bp.inv_norms = st->d_inv;
    bp.q_scal = q->metric == OTTERS_METRIC_COSINE ? d_qinv : d_qn2;
    bp.take_max = take_max;
    bp.has_filter = q->has_filter;
    bp.thr = q->thr;
    bp.cmp = q->cmp;
    bp.delta = d_delta;
    bp.row_mask = d_row_mask;
    bp.row_mask_words = row_mask_words;
    bp.k = k;
    bp.cap = cap;
    bp.g_tau = c->d_batch_info;
    bp.g_flags = reinterpret_cast<uint32_t*>(c->d_batch_info + 1);
    bp.g_excl = d_excl;
    bp.pairs_scored = c->d_rows_scored;
    bp.cta_keys = c->d_cta_keys;
    bp.cta_qids = c->d_cta_qids;
    bp.cta_counts = c->d_cta_counts;
    bp.passes = passes;
    {
        static const int kps = getenv("OTTERS_K2_KPS") ? atoi(getenv("OTTERS_K2_KPS")) : 3;  // clamped in the kernel to leave two stages
        static const int direct = getenv("OTTERS_K2_DIRECT") ? atoi(getenv("OTTERS_K2_DIRECT")) : 1;
        static const int epi = getenv("OTTERS_K2_EPI_WARPS") ? atoi(getenv("OTTERS_K2_EPI_WARPS")) : 8;  // A/B: 4 = one epilogue warp per scheduler
        bp.kps = (uint32_t)kps;
        bp.pair_direct = (uint32_t)direct;
        bp.epi_warps = epi == 4 ? 4u : 8u;
        static const int redo = getenv("OTTERS_K2_REDO") ? atoi(getenv("OTTERS_K2_REDO")) : 0;  // A/B: 1 = redo chunks on the general path
        bp.redo_general = redo ? 1u : 0u;
    }
#ifdef OTTERS_K2_EXPERIMENTS
    // timing experiments only (role-by-role timing of K2, scripts/dbg_passes_roles.py): results are garbage, so the hooks are
    // compiled out of the shipped library and a run with them on is never accepted (see below)
    if (const char* e = getenv("OTTERS_BATCH_DBG")) bp.dbg = (uint32_t)atoi(e);
#endif
    rec_event(c, 2);
    rec_event(c, 3);
    rc = launch_batch(bl, bp, q->metric, c->batch_smem_configured, s);
    if (rc) return rc;
    rec_event(c, 4);
    c->timed_single = true;

    // exact re-scoring of every selected pair, then a full sort of the exact candidates
    const uint32_t total = bl.grid * k;
    const uint64_t n_sort = std::max<uint64_t>(pow2_at_least(total), 2048);
    if (c->emit_cap < n_sort) {
        OTTERS_CUDA(cudaStreamSynchronize(s));
        cudaFree(c->d_emit);
        c->d_emit = nullptr;
        c->emit_cap = 0;
        if (cudaMalloc((void**)&c->d_emit, n_sort * sizeof(Cand)) != cudaSuccess)
            return fail(OTTERS_ERR_NOMEM, "device allocation for the candidate sort failed");
        c->emit_cap = n_sort;
    }
    rc = ensure_list0(c, k_eff);
    if (rc) return rc;
    RescoreParams rp{};
    rp.half = st->half ? 1u : 0u;
    rp.vectors = st->d_rows;
    rp.inv_norms = st->d_inv;
    rp.queries = c->d_query;
    rp.q_inv = d_qinv;
    rp.pitch_g = st->pitch;
    rp.dim = st->dim;
    rp.dim_pad = dim_pad;
    rp.cta_keys = c->d_cta_keys;
    rp.cta_qids = c->d_cta_qids;
    rp.cta_counts = c->d_cta_counts;
    rp.n_lists = bl.grid;
    rp.k = k;
    rp.take_max = take_max;
    rp.has_filter = q->has_filter;
    rp.thr = q->thr;
    rp.cmp = q->cmp;
    rp.out = c->d_emit;
    rp.out_slots = total;
    rp.out_count = c->d_counter + 1;
    rp.max_err_bits = reinterpret_cast<uint32_t*>(c->d_batch_info + 1) + 1;
    rc = launch_rescore(rp, q->metric, (uint32_t)n_sort, s);
    if (rc) return rc;
    rc = launch_global_sort(c->d_emit, n_sort, s);
    if (rc) return rc;
    rc = launch_take_sorted(c->d_emit, c->d_counter + 1, nullptr, k_eff, c->d_list[0], c->d_list_count, c->d_tau, list_hdr(c, 0),
                            c->d_rows_scored, stats_src, s, c->d_batch_info + 1);
    if (rc) return rc;
    if (d_records_out) {
        rc = launch_cands_to_records(c->d_list[0], c->d_list_count, (uint32_t)k_eff, map, take_max, d_records_out, s);
        if (rc) return rc;
    }
    rec_event(c, 5);
    c->last.kernel_launches += 8 + (d_records_out ? 1 : 0) + (passes == 2 ? 1 : 0);

    // fetch header + candidates and verify the selection
    const size_t bytes = sizeof(ResultHeader) + (size_t)k_eff * sizeof(Cand);
    rc = ensure_pinned(&c->h_result, &c->h_result_bytes, bytes);
    if (rc) return rc;
    OTTERS_CUDA(cudaMemcpyAsync(c->h_result, c->d_list_raw[0], bytes, cudaMemcpyDeviceToHost, s));
    OTTERS_CUDA(cudaStreamSynchronize(s));
    c->last.d2h_bytes += bytes;
    const ResultHeader* hdr = reinterpret_cast<const ResultHeader*>(c->h_result);
    const Cand* list = reinterpret_cast<const Cand*>(c->h_result + sizeof(ResultHeader));
    const uint32_t flags = (uint32_t)hdr->extra[0];
    const uint32_t err_bits = (uint32_t)(hdr->extra[0] >> 32);
    const uint32_t excl = (uint32_t)hdr->extra[1];
    const uint32_t delta_bits = (uint32_t)(hdr->extra[1] >> 32);
    float max_err, delta;
    memcpy(&max_err, &err_bits, 4);
    memcpy(&delta, &delta_bits, 4);
    bool ok = (flags & 1u) == 0 && delta <= FLT_MAX && max_err <= delta && bp.dbg == 0;
    if (ok && excl != 0) {
        // some pair was left out of the candidate lists: its exact score is within delta of its tensor-core
        // score, so it cannot belong to the result iff even that bound stays strictly outside the k-th score
        if (hdr->count < k_eff) ok = false;
        else {
            const float e_k = key_score(list[hdr->count - 1].key, take_max);
            const float x = key_score((uint64_t)excl << 32, take_max);
            ok = take_max ? (x + delta < e_k) : (x - delta > e_k);
        }
    }
    if (getenv("OTTERS_BATCH_TRACE"))
        fprintf(stderr, "[otters batch] passes=%u k=%llu count=%u flags=%u max_err=%g delta=%g excl=%g e_k=%g verified=%d\n",
                passes, (unsigned long long)k_eff, hdr->count, flags, max_err, delta, excl ? key_score((uint64_t)excl << 32, take_max) : NAN,
                hdr->count ? key_score(list[hdr->count - 1].key, take_max) : NAN, (int)ok);
    c->last.batch_max_err = max_err;
    c->last.batch_delta = delta;
    c->last.batch_candidates = 0;
    if (!ok) {
        c->last.batch_fallback = 1;
        c->timed_single = false;
        return OTTERS_OK;
    }
    c->last.batch_used = 1;
    c->last.batch_passes = passes;
    run->result_list = 0;
    run->big = false;
    run->prefetched = true;
    *accepted = true;
    return OTTERS_OK;
}

static int run_queries(otters_ctx* c, VecStorage* st, const otters_vec_query* q, const uint32_t* d_row_mask,
                       uint32_t row_mask_words, otters_topk_record* d_records_out, ShardMap map,
                       const unsigned long long* stats_src, const FusedFilter* ff, QueryRun* run) {
    const uint32_t dim_pad = st->pitch;
    const uint64_t n_rows = st->n;
    const uint64_t k_eff = std::min<uint64_t>(q->k, n_rows * (uint64_t)q->nq);
    run->k_eff = k_eff;
    run->result_list = 0;
    c->last_dim = st->dim;
    c->last_esz = (uint32_t)st->esz();
    c->last_metric = q->metric;
    cudaStream_t s = c->stream;

    // the queries were staged (zero padded to the stored pitch) by stage_queries(); send the input image now
    int rc = io_flush(c);
    if (rc) return rc;

    if (!c->ex_active && batch_eligible(c, q, n_rows, k_eff)) {
        // selection runs single-pass tf32 first (a third of the MMAs and of the operand traffic; error bound 2^-9 |q||v|);
        // when its certificate fails the batch is redone with the 3xTF32 split (2^-15), and only then query by query.
        // A store whose single-pass certificate failed goes straight to 3xTF32 for its next kSinglePassBackoff batches.
        // The bf16 rung (a bf16 shadow of the rows, half the operand bytes and twice the MMA rate, bound 2^-7 |q||v|) goes first
        // when the shadow exists or can be built; every rung is selection only and carries the same kind of certificate.
        // A bf16 STORE has one rung: its rows are the bf16 operand (no shadow, no operand error on the row side) and the exact
        // re-scoring reads the same rows; the tf32 rungs would need fp32 rows, so a declined certificate goes to K1.
        const bool half_auto = st->half && c->tuning.batch_passes != 2;  // (its declined certificates back off like the ladder's)
        const uint32_t want = st->half ? 2u : c->tuning.batch_passes;
        bool accepted = false;
        if (half_auto && st->bf16_backoff) {
            st->bf16_backoff -= 1;
        } else if (want == 2 || (want == 0 && st->bf16_backoff == 0)) {
            bool have = false;
            rc = ensure_bf16_shadow(c, st, want == 2, &have);
            if (rc) return rc;
            if (want == 2 && !have) return fail(OTTERS_ERR_NOMEM, "device allocation for the bf16 shadow rows failed");
            if (have) {
                rc = run_batched(c, st, q, d_row_mask, row_mask_words, d_records_out, map, stats_src, run, 2, &accepted);
                if (rc) return rc;
                if (accepted) return OTTERS_OK;
                if (want == 0 || half_auto) st->bf16_backoff = kSinglePassBackoff;
                rc = reset_scan_state(c);
                if (rc) return rc;
                c->last.batch_fallback = want == 2 ? 1 : 0;
            }
        } else if (want == 0 && st->bf16_backoff) {
            st->bf16_backoff -= 1;
        }
        const bool try_single = want == 1 || (want == 0 && st->single_pass_backoff == 0);
        if (want == 0 && st->single_pass_backoff) st->single_pass_backoff -= 1;
        if (try_single) {
            rc = run_batched(c, st, q, d_row_mask, row_mask_words, d_records_out, map, stats_src, run, 1, &accepted);
            if (rc) return rc;
            if (accepted) return OTTERS_OK;
            if (want == 0) st->single_pass_backoff = kSinglePassBackoff;
            rc = reset_scan_state(c);
            if (rc) return rc;
        }
        if (want != 1 && want != 2) {
            c->last.batch_fallback = 0;
            rc = run_batched(c, st, q, d_row_mask, row_mask_words, d_records_out, map, stats_src, run, 3, &accepted);
            if (rc) return rc;
            if (accepted) return OTTERS_OK;
            rc = reset_scan_state(c);  // selection could not be verified: exact path, query by query
            if (rc) return rc;
        }
    }

    // per-query inverse norms for the streaming kernel (a by-value kernel parameter)
    std::vector<float> q_inv(q->nq);
    for (uint32_t i = 0; i < q->nq; ++i) q_inv[i] = host_inv_norm(q->queries + (size_t)i * q->dim, q->dim);

    const bool fused = k_eff <= kMaxFusedK;
    // single queries: the selection (K3) runs in the last CTA of the scan kernel — one launch per query
    const bool fuse_select = fused && q->nq == 1 && !c->tuning.separate_select;
    ScanPlan pl;
    rc = plan_scan(c, dim_pad, n_rows, fused ? (uint32_t)k_eff : 0, ff ? ff->smem_bytes() : 0, fuse_select, &pl, st->half);
    if (rc) return rc;

    ScanParams sp{};
    sp.vectors = st->d_rows;
    sp.inv_norms = st->d_inv;
    sp.pitch_g = st->pitch;
    sp.half = st->half ? 1u : 0u;
    sp.dim = st->dim;
    sp.dim_pad = dim_pad;
    sp.n_rows = (uint32_t)n_rows;
    sp.row_mask = d_row_mask;
    sp.row_mask_words = row_mask_words;
    if (ff) {
        sp.flt_leaves = ff->leaves;
        sp.flt_clause_off = ff->clause_off;
        sp.flt_n_clauses = ff->n_clauses;
        sp.flt_n_leaves = ff->n_leaves;
        sp.chunk_keep = ff->chunk_keep;  // null: lazy pruning inside the scan kernel
        sp.chunk_size = ff->chunk_size;
        sp.n_chunks = ff->n_chunks;
        sp.nq_stats = q->nq;
        sp.stats = c->d_stats;
    }
    sp.off_filter = pl.off_filter;
    sp.n_units = pl.n_units;
    sp.unit_rows = pl.unit_rows;
    sp.unit_small = pl.unit_small;
    sp.n_big = pl.n_big;
    sp.unit_counter = c->d_counter;
    sp.k = (uint32_t)std::min<uint64_t>(k_eff, kMaxFusedK);
    sp.cap = pl.cap;
    sp.take_max = q->take_type == OTTERS_TAKE_MAX;
    sp.has_filter = q->has_filter;
    sp.thr = q->thr;
    sp.cmp = q->cmp;
    sp.kc = pl.kc;
    sp.nkc = pl.nkc;
    sp.pitch_s = pl.pitch_s;
    sp.slots = pl.slots;
    sp.planners = pl.planners;
    sp.off_ring = pl.off_ring;
    sp.off_query = pl.off_query;
    sp.off_warps = pl.off_warps;
    sp.warp_bytes = pl.warp_bytes;
    sp.off_w_rows = pl.off_w_rows;
    sp.off_w_info = pl.off_w_info;
    sp.off_w_inv = pl.off_w_inv;
    sp.off_w_list = pl.off_w_list;
    sp.off_w_slots = pl.off_w_slots;
    sp.cta_keys = c->d_cta_keys;
    sp.cta_counts = c->d_cta_counts;
    sp.rows_scored = c->d_rows_scored;
    sp.g_tau = reinterpret_cast<unsigned long long*>(c->d_ctrl + 96);  // zeroed with the control block
    {
        // unit ids claimed ahead per claiming warp: two (same-box A/B, profiles/r2_claim_depth.txt: 10M x 128 filtered 2054 -> 2111
        // q/s, 100k x 128 44.6k -> 47.5k; four ids are no better and starve late warps on small stores)
        static const int depth_env = getenv("OTTERS_CLAIM_DEPTH") ? atoi(getenv("OTTERS_CLAIM_DEPTH")) : 0;
        sp.claim_depth = depth_env ? (uint32_t)std::min(std::max(depth_env, 1), 4) : 2u;
    }
    {
        static const int pred_seq = getenv("OTTERS_PRED_SEQ") ? atoi(getenv("OTTERS_PRED_SEQ")) : 1;
        static const int no_prefetch = getenv("OTTERS_NO_PREFETCH") ? atoi(getenv("OTTERS_NO_PREFETCH")) : 0;
        sp.pred_seq = pred_seq;
        sp.no_prefetch = no_prefetch;
    }

    uint32_t cur = 0;
    if (fused) {
        rc = ensure_list0(c, kMaxFusedK);
        if (rc) return rc;
        rec_event(c, 2);
        for (uint32_t qi = 0; qi < q->nq; ++qi) {
            if (qi) OTTERS_CUDA(cudaMemsetAsync(c->d_counter, 0, sizeof(uint32_t), s));  // qi == 0: begin_query()
            sp.query = c->d_query + (size_t)qi * dim_pad;
            sp.q_inv = q_inv[qi];
            sp.qid = qi;
            sp.tau_in = qi ? c->d_tau : nullptr;
            SelectParams se{};
            se.cta_keys = c->d_cta_keys;
            se.cta_counts = c->d_cta_counts;
            se.n_lists = pl.launch.grid;
            se.list_stride = sp.k;
            se.qid = qi;
            se.prev = qi ? c->d_list[cur] : nullptr;
            se.prev_count = qi ? c->d_list_count + cur : nullptr;
            se.out = c->d_list[cur ^ 1];
            se.out_count = c->d_list_count + (cur ^ 1);
            se.tau_out = c->d_tau;
            se.k = sp.k;
            se.scratch_keys = c->d_scratch_keys;
            se.scratch_src = c->d_scratch_src;
            se.scratch_elems = c->scratch_elems;
            se.records = (qi + 1 == q->nq) ? d_records_out : nullptr;
            se.map = map;
            se.take_max = sp.take_max;
            se.hdr = list_hdr(c, cur ^ 1);
            se.rows_scored_src = c->d_rows_scored;
            se.stats_src = stats_src;
            if (qi + 1 == q->nq) {
                fill_exchange(c, &se);
                if (c->want_host_result && !d_records_out) {
                    se.host_out = c->d_zc;
                    run->zero_copy = true;
                }
            }
            sp.fuse_select = fuse_select ? 1u : 0u;
            sp.done_counter = reinterpret_cast<uint32_t*>(c->d_ctrl + 104);  // zeroed with the control block
            if (fuse_select) sp.sel = se;
            if (qi == 0 && q->nq == 1) rec_event(c, 3);
            rc = pl.planners ? launch_scan_planner(sp, pl.launch, q->metric, c->planner_smem_configured, s)
                             : launch_scan(sp, pl.launch, q->metric, false, c->scan_smem_configured, s);
            if (rc) return rc;
            if (qi == 0 && q->nq == 1) {
                rec_event(c, 4);
                c->timed_single = true;
            }
            c->last.kernel_launches += 1;
            if (!fuse_select) {
                rc = launch_select(se, s);
                if (rc) return rc;
                c->last.kernel_launches += 1;
            }
            if (c->ex_active && qi + 1 == q->nq) c->ex_published = true;
            cur ^= 1;
        }
        rec_event(c, 5);
        run->result_list = cur;
        run->big = false;
        if (c->ex_active) run->k_eff = c->ex_k;  // the merged list may be longer than this shard's own
    } else {
        // large k: emit every passing candidate, sort everything, keep the best k (per query, with the
        // running list appended before the sort)
        if (k_eff >= 0x7FFFFFF0ull) return fail(OTTERS_ERR_UNSUPPORTED, "take count too large");
        const uint64_t n_sort = std::max<uint64_t>(pow2_at_least(n_rows + k_eff), 2048);
        if (c->emit_cap < n_sort) {
            OTTERS_CUDA(cudaStreamSynchronize(s));
            cudaFree(c->d_emit);
            c->d_emit = nullptr;
            c->emit_cap = 0;
            if (cudaMalloc((void**)&c->d_emit, n_sort * sizeof(Cand)) != cudaSuccess)
                return fail(OTTERS_ERR_NOMEM, "device allocation for the candidate sort failed");
            c->emit_cap = n_sort;
        }
        rc = ensure_list0(c, k_eff);
        if (rc) return rc;
        sp.emit = c->d_emit;
        sp.emit_count = c->d_counter + 1;
        sp.emit_cap = (uint32_t)std::min<uint64_t>(n_sort, 0xFFFFFFFFull);
        rec_event(c, 2);
        for (uint32_t qi = 0; qi < q->nq; ++qi) {
            if (qi) OTTERS_CUDA(cudaMemsetAsync(c->d_counter, 0, 2 * sizeof(uint32_t), s));
            sp.query = c->d_query + (size_t)qi * dim_pad;
            sp.q_inv = q_inv[qi];
            sp.qid = qi;
            sp.tau_in = qi ? c->d_tau : nullptr;
            rc = launch_scan(sp, pl.launch, q->metric, true, c->scan_smem_configured, s);
            if (rc) return rc;
            // the running list lives in d_list[0] (single buffer: it is copied into the sort array first); its length
            // alternates between the two count slots, so that take_sorted's block 0 never writes the count other blocks read
            const uint32_t* prev_cnt = qi ? c->d_list_count + (qi & 1u) : nullptr;
            uint32_t* out_cnt = c->d_list_count + ((qi + 1u) & 1u);
            rc = launch_append_prev(c->d_emit, c->d_counter + 1, qi ? c->d_list[0] : nullptr, prev_cnt, n_sort, s);
            if (rc) return rc;
            rc = launch_global_sort(c->d_emit, n_sort, s);
            if (rc) return rc;
            rc = launch_take_sorted(c->d_emit, c->d_counter + 1, prev_cnt, k_eff, c->d_list[0], out_cnt, c->d_tau, list_hdr(c, 0),
                                    c->d_rows_scored, stats_src, s);
            if (rc) return rc;
            c->last.kernel_launches += 4;
        }
        if (d_records_out) {
            rc = launch_cands_to_records(c->d_list[0], c->d_list_count + (q->nq & 1u), (uint32_t)k_eff, map, sp.take_max, d_records_out, s);
            if (rc) return rc;
        }
        rec_event(c, 5);
        run->result_list = 0;
        run->big = true;
    }
    return OTTERS_OK;
}

// one D2H copy (header + ordered candidates), one stream sync, decode into the caller's arrays
static int fetch_results(otters_ctx* c, const QueryRun& run, bool take_max, uint64_t row_base, uint64_t* out_idx,
                         float* out_score, uint32_t* out_qid, uint64_t cap, uint64_t* out_len, unsigned long long* stats_out) {
    cudaStream_t s = c->stream;
    const size_t bytes = sizeof(ResultHeader) + (size_t)run.k_eff * sizeof(Cand);
    const uint8_t* src = nullptr;
    if (run.zero_copy) {
        OTTERS_CUDA(cudaStreamSynchronize(s));  // the selection kernel has written the result into mapped host memory
        c->last.d2h_bytes += bytes;
        src = c->h_zc;
    } else if (!run.prefetched) {
        int rc = ensure_pinned(&c->h_result, &c->h_result_bytes, bytes);
        if (rc) return rc;
        OTTERS_CUDA(cudaMemcpyAsync(c->h_result, c->d_list_raw[run.result_list], bytes, cudaMemcpyDeviceToHost, s));
        OTTERS_CUDA(cudaStreamSynchronize(s));
        c->last.d2h_bytes += bytes;
    }
    if (!src) src = c->h_result;  // (after ensure_pinned, which may have reallocated it)
    const ResultHeader* hdr = reinterpret_cast<const ResultHeader*>(src);
    const uint32_t n = hdr->count;
    c->last.rows_scored = hdr->rows_scored;
    if (stats_out) {
        stats_out[0] = hdr->stats[0];
        stats_out[1] = hdr->stats[1];
    }
    const Cand* list = reinterpret_cast<const Cand*>(src + sizeof(ResultHeader));
    uint64_t m = std::min<uint64_t>(n, cap);
    for (uint64_t i = 0; i < m; ++i) {
        if (out_idx) out_idx[i] = row_base + key_row(list[i].key);
        if (out_score) out_score[i] = key_score(list[i].key, take_max);
        if (out_qid) out_qid[i] = list[i].qid;
    }
    *out_len = n;
    return OTTERS_OK;
}

static void finish_work_stats(otters_ctx* c) {
    // only called after the stream was synchronised, so all events have completed
    float ms = 0.f;
    if (c->timing) {
        if (c->timed_single && elapsed_ms(c->ev[3], c->ev[4], &ms)) c->last.scan_ms = ms;
        else if (elapsed_ms(c->ev[2], c->ev[5], &ms)) c->last.scan_ms = ms;
        if (c->timed_single && elapsed_ms(c->ev[4], c->ev[5], &ms)) c->last.select_ms = ms;
    }
    const uint64_t per_row = (uint64_t)c->last_dim * c->last_esz + (c->last_metric == OTTERS_METRIC_COSINE ? 4 : 0);
    c->last.scan_bytes = c->last.rows_scored * per_row;
}

static int validate_query(const otters_vec_query* q, uint32_t store_dim, bool meta) {
    if (!q) return fail(OTTERS_ERR_INVALID, "Query vectors or their norms are not set");
    if (q->metric < 0 || q->metric > 2) return fail(OTTERS_ERR_INVALID, "Search metric is not set");
    if (q->take_type != OTTERS_TAKE_MIN && q->take_type != OTTERS_TAKE_MAX) return fail(OTTERS_ERR_INVALID, "invalid take type");
    if (q->has_filter && (q->cmp < 0 || q->cmp > 4)) return fail(OTTERS_ERR_INVALID, "invalid filter comparator");
    if (meta) return OTTERS_OK;  // MetaStore swallows the per-chunk validation errors (src/meta_compute.rs:182)
    if (q->nq == 0) return fail(OTTERS_ERR_INVALID, "No queries provided");  // src/vec.rs:186-188
    if (!q->queries) return fail(OTTERS_ERR_INVALID, "Query vectors or their norms are not set");
    if (q->dim != store_dim)  // src/vec.rs:190-198
        return fail(OTTERS_ERR_INVALID, "Query vector length " + std::to_string(q->dim) + " does not match expected dimension " +
                                            std::to_string(store_dim));
    return OTTERS_OK;
}

}  // namespace otters

// =================================================================================================
// context API
// =================================================================================================
extern "C" const char* otters_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* otters_version(void) { return "otters_b200 0.1.0 (sm_100a)"; }

namespace otters {
static int ctx_create_impl(int device, void* cuda_stream, otters_ctx** out) {
    if (!out) return fail(OTTERS_ERR_INVALID, "null output pointer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(OTTERS_ERR_CUDA, "no CUDA device available: libotters_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(OTTERS_ERR_INVALID, "invalid CUDA device index");
    OTTERS_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    OTTERS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(OTTERS_ERR_UNSUPPORTED, std::string("libotters_b200 is built for sm_100a only; device is sm_") +
                                                std::to_string(prop.major) + std::to_string(prop.minor));
    std::unique_ptr<otters_ctx> c(new otters_ctx());
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    c->smem_per_sm = prop.sharedMemPerMultiprocessor;
    if (cuda_stream) {
        c->stream = (cudaStream_t)cuda_stream;
    } else {
        OTTERS_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    c->grid_max = (uint32_t)c->sm_count * 2;
    if (const char* e = getenv("OTTERS_SCAN_MODE")) c->tuning.scan_mode = (uint32_t)atoi(e);
    c->d_io_bytes = 1 << 16;
    OTTERS_CUDA(cudaMalloc((void**)&c->d_io, c->d_io_bytes));
    OTTERS_CUDA(cudaMemset(c->d_io, 0, 256));
    bind_io(c.get());
    for (auto& e : c->ev_io) OTTERS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    OTTERS_CUDA(cudaHostAlloc((void**)&c->h_zc, sizeof(ResultHeader) + kMaxFusedK * sizeof(Cand), cudaHostAllocMapped));
    OTTERS_CUDA(cudaHostGetDevicePointer((void**)&c->d_zc, c->h_zc, 0));
    OTTERS_CUDA(cudaMalloc((void**)&c->d_cta_keys, (size_t)c->grid_max * kMaxFusedK * sizeof(uint64_t)));
    OTTERS_CUDA(cudaMalloc((void**)&c->d_cta_counts, (size_t)c->grid_max * sizeof(uint32_t)));
    OTTERS_CUDA(cudaMalloc((void**)&c->d_cta_qids, (size_t)c->grid_max * kMaxFusedK * sizeof(uint32_t)));
    {
        int rc0 = alloc_list(c.get(), 0, kMaxFusedK);
        if (rc0) return rc0;
        rc0 = alloc_list(c.get(), 1, kMaxFusedK);
        if (rc0) return rc0;
        c->list_cap = kMaxFusedK;
    }
    c->scratch_elems = (uint32_t)pow2_at_least((uint64_t)(c->grid_max + 1) * kMaxFusedK);
    OTTERS_CUDA(cudaMalloc((void**)&c->d_scratch_keys, (size_t)c->scratch_elems * sizeof(uint64_t)));
    OTTERS_CUDA(cudaMalloc((void**)&c->d_scratch_src, (size_t)c->scratch_elems * sizeof(uint32_t)));
    for (auto& e : c->ev) OTTERS_CUDA(cudaEventCreate(&e));
    *out = c.release();
    return OTTERS_OK;
}
}  // namespace otters

extern "C" int otters_ctx_create(int device, void* cuda_stream, otters_ctx** out) { return ctx_create_impl(device, cuda_stream, out); }

extern "C" int otters_ctx_destroy(otters_ctx* c) {
    if (!c) return OTTERS_OK;
    DeviceGuard g(c->device);
    for (uint32_t i = 1; i < kMaxLanes; ++i) {
        if (c->lane[i]) otters_ctx_destroy(c->lane[i]);
        c->lane[i] = nullptr;
    }
    cudaStreamSynchronize(c->stream);
    cudaFree(c->d_chunk_keep);
    cudaFree(c->d_meta_mask);
    cudaFree(c->d_gather);
    cudaFree(c->d_io);
    for (auto& e : c->ev_io)
        if (e) cudaEventDestroy(e);
    for (auto& h : c->h_io)
        if (h) cudaFreeHost(h);
    cudaFree(c->d_cta_keys);
    cudaFree(c->d_cta_counts);
    cudaFree(c->d_list_raw[0]);
    cudaFree(c->d_list_raw[1]);
    cudaFree(c->d_scratch_keys);
    cudaFree(c->d_scratch_src);
    cudaFree(c->d_mask);
    cudaFree(c->d_emit);
    cudaFree(c->d_qh);
    cudaFree(c->d_ql);
    cudaFree(c->d_qb);
    cudaFree(c->d_qscal);
    cudaFree(c->d_cta_qids);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->h_result) cudaFreeHost(c->h_result);
    if (c->h_zc) cudaFreeHost(c->h_zc);
    for (auto& e : c->ev)
        if (e) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return OTTERS_OK;
}

extern "C" int otters_ctx_synchronize(otters_ctx* c) {
    if (!c) return fail(OTTERS_ERR_INVALID, "null context");
    DeviceGuard g(c->device);
    OTTERS_CUDA(cudaStreamSynchronize(c->stream));
    for (uint32_t i = 1; i < kMaxLanes; ++i)
        if (c->lane[i]) OTTERS_CUDA(cudaStreamSynchronize(c->lane[i]->stream));
    return OTTERS_OK;
}

extern "C" int otters_ctx_join(otters_ctx* c) {
    if (!c) return fail(OTTERS_ERR_INVALID, "null context");
    DeviceGuard g(c->device);
    for (uint32_t i = 1; i < kMaxLanes; ++i) {
        otters_ctx* l = c->lane[i];
        if (!l) continue;
        OTTERS_CUDA(cudaEventRecord(l->ev_io[0], l->stream));  // (ev_io events carry no timing: cheap to record)
        OTTERS_CUDA(cudaStreamWaitEvent(c->stream, l->ev_io[0], 0));
    }
    return OTTERS_OK;
}

extern "C" int otters_ctx_set_tuning(otters_ctx* c, const otters_scan_tuning* t) {
    if (!c) return fail(OTTERS_ERR_INVALID, "null context");
    c->tuning = t ? *t : otters_scan_tuning{};
    // test hook: OTTERS_SCAN_MODE selects the K1 front-end wherever the caller left it automatic
    if (c->tuning.scan_mode == 0)
        if (const char* e = getenv("OTTERS_SCAN_MODE")) c->tuning.scan_mode = (uint32_t)atoi(e);
    for (uint32_t i = 1; i < kMaxLanes; ++i)
        if (c->lane[i]) c->lane[i]->tuning = c->tuning;
    return OTTERS_OK;
}

extern "C" int otters_ctx_last_work(otters_ctx* c, otters_last_work* out) {
    if (!c || !out) return fail(OTTERS_ERR_INVALID, "null argument");
    *out = c->report_waited ? c->reported : c->last;
    return OTTERS_OK;
}

// =================================================================================================
// VecStore
// =================================================================================================
struct otters_vecstore {
    VecStorage st;
};

extern "C" int otters_vecstore_create_fmt(otters_ctx* c, uint32_t dim, int32_t vector_format, otters_vecstore** out) {
    if (!c || !out) return fail(OTTERS_ERR_INVALID, "null argument");
    if (vector_format != OTTERS_VECTORS_FMT_F32 && vector_format != OTTERS_VECTORS_FMT_BF16)
        return fail(OTTERS_ERR_INVALID, "unknown vector format");
    auto* vs = new otters_vecstore();
    vs->st.ctx = c;
    vs->st.set_format(dim, vector_format == OTTERS_VECTORS_FMT_BF16);
    *out = vs;
    return OTTERS_OK;
}

extern "C" int otters_vecstore_create(otters_ctx* c, uint32_t dim, otters_vecstore** out) {
    return otters_vecstore_create_fmt(c, dim, OTTERS_VECTORS_FMT_F32, out);
}

extern "C" int32_t otters_vecstore_format(const otters_vecstore* vs) {
    return vs && vs->st.half ? OTTERS_VECTORS_FMT_BF16 : OTTERS_VECTORS_FMT_F32;
}

extern "C" int otters_vecstore_destroy(otters_vecstore* vs) {
    if (!vs) return OTTERS_OK;
    DeviceGuard g(vs->st.ctx->device);
    cudaStreamSynchronize(vs->st.ctx->stream);
    vs->st.release();
    delete vs;
    return OTTERS_OK;
}

extern "C" int otters_vecstore_reserve(otters_vecstore* vs, uint64_t n_rows) {
    if (!vs) return fail(OTTERS_ERR_INVALID, "null store");
    DeviceGuard g(vs->st.ctx->device);
    return vs->st.reserve(n_rows);
}

extern "C" int otters_vecstore_add(otters_vecstore* vs, const float* rows, uint64_t n) {
    if (!vs) return fail(OTTERS_ERR_INVALID, "null store");
    if (n && !rows) return fail(OTTERS_ERR_INVALID, "null rows");
    if (n && vs->st.dim == 0) return fail(OTTERS_ERR_INVALID, "Input vector length 0 does not match expected dimension 0");
    DeviceGuard g(vs->st.ctx->device);
    return vs->st.add(rows, n, cudaMemcpyHostToDevice);
}

extern "C" int otters_vecstore_add_device(otters_vecstore* vs, const float* d_rows, uint64_t n) {
    if (!vs) return fail(OTTERS_ERR_INVALID, "null store");
    if (n && !d_rows) return fail(OTTERS_ERR_INVALID, "null rows");
    DeviceGuard g(vs->st.ctx->device);
    return vs->st.add(d_rows, n, cudaMemcpyDeviceToDevice);
}

extern "C" int otters_vecstore_add_synthetic(otters_vecstore* vs, uint64_t first_row, uint64_t n, uint64_t seed) {
    if (!vs) return fail(OTTERS_ERR_INVALID, "null store");
    DeviceGuard g(vs->st.ctx->device);
    ShardMap m;
    m.row_base = first_row;
    return vs->st.add_synth(m, n, seed);
}

static ShardMap to_map(const otters_shard_map* in) {
    ShardMap m;
    if (in) {
        m.row_base = in->row_base;
        m.world = in->world;
        m.rank = in->rank;
        m.block = in->block_rows;
    }
    return m;
}

extern "C" int otters_vecstore_add_synthetic_sharded(otters_vecstore* vs, const otters_shard_map* map, uint64_t n_local, uint64_t seed) {
    if (!vs) return fail(OTTERS_ERR_INVALID, "null store");
    DeviceGuard g(vs->st.ctx->device);
    return vs->st.add_synth(to_map(map), n_local, seed);
}

extern "C" int otters_vecstore_set_rows(otters_vecstore* vs, const uint64_t* rows, const float* data, uint64_t n) {
    if (!vs) return fail(OTTERS_ERR_INVALID, "null store");
    if (n && (!rows || !data)) return fail(OTTERS_ERR_INVALID, "null rows");
    DeviceGuard g(vs->st.ctx->device);
    OTTERS_CUDA(otters_ctx_synchronize(vs->st.ctx) == OTTERS_OK ? cudaSuccess : cudaErrorUnknown);  // no query may be in flight
    return vs->st.set_rows(rows, data, n);
}

extern "C" uint64_t otters_vecstore_len(const otters_vecstore* vs) { return vs ? vs->st.n : 0; }
extern "C" uint32_t otters_vecstore_dim(const otters_vecstore* vs) { return vs ? vs->st.dim : 0; }

extern "C" int otters_vecstore_inv_norms(const otters_vecstore* vs, uint64_t first, uint64_t n, float* out) {
    if (!vs || !out) return fail(OTTERS_ERR_INVALID, "null argument");
    if (first + n > vs->st.n) return fail(OTTERS_ERR_INVALID, "row range out of bounds");
    DeviceGuard g(vs->st.ctx->device);
    OTTERS_CUDA(cudaStreamSynchronize(vs->st.ctx->stream));
    if (n) OTTERS_CUDA(cudaMemcpy(out, vs->st.d_inv + first, n * sizeof(float), cudaMemcpyDeviceToHost));
    return OTTERS_OK;
}

namespace otters {
// uploads a user row mask (Lsb0 u64 words) as 32-bit words, bits past row_mask_bits set to "keep"
static int upload_row_mask(otters_ctx* c, const otters_vec_query* q, uint64_t n_rows, const uint32_t** d_mask, uint32_t* words) {
    *d_mask = nullptr;
    *words = 0;
    if (!q->row_mask_words) return OTTERS_OK;
    const uint64_t bits = std::min<uint64_t>(q->row_mask_bits, n_rows);
    const uint64_t w32 = (bits + 31) / 32;
    if (w32 == 0) return OTTERS_OK;
    int rc = ensure_stage(c, w32 * 4 + 64);
    if (rc) return rc;
    rc = ensure_dev(&c->d_mask, &c->d_mask_words, w32, c->stream);
    if (rc) return rc;
    uint32_t* h = reinterpret_cast<uint32_t*>(c->h_stage);
    for (uint64_t i = 0; i < w32; ++i) {
        uint64_t w = q->row_mask_words[i >> 1];
        h[i] = (uint32_t)(i & 1 ? (w >> 32) : w);
    }
    if (bits & 31) h[w32 - 1] |= ~0u << (bits & 31);  // rows >= mask length are kept (src/vec.rs:234,297)
    OTTERS_CUDA(cudaMemcpyAsync(c->d_mask, h, w32 * 4, cudaMemcpyHostToDevice, c->stream));
    c->last.h2d_bytes += w32 * 4;
    // the staging buffer is reused by run_queries: make sure the copy has been consumed
    OTTERS_CUDA(cudaStreamSynchronize(c->stream));
    *d_mask = c->d_mask;
    *words = (uint32_t)w32;
    return OTTERS_OK;
}
}  // namespace otters

namespace otters {
static int vec_enqueue(otters_ctx* c, otters_vecstore* vs, const otters_vec_query* q, otters_topk_record* d_records, ShardMap map,
                       bool want_host, Pending* pd);
static int vec_finish(otters_ctx* c, Pending* pd, bool fetch, uint64_t* out_idx, float* out_score, uint32_t* out_qid, uint64_t cap,
                      uint64_t* out_len);
}  // namespace otters

extern "C" int otters_vecstore_query(otters_vecstore* vs, const otters_vec_query* q, uint64_t* out_idx, float* out_score,
                                     uint32_t* out_qid, uint64_t cap, uint64_t* out_len) {
    if (!vs || !out_len) return fail(OTTERS_ERR_INVALID, "null argument");
    otters_ctx* c = vs->st.ctx;
    DeviceGuard g(c->device);
    *out_len = 0;
    Pending pd;
    int rc = vec_enqueue(c, vs, q, nullptr, ShardMap{}, true, &pd);
    if (rc) return rc;
    return vec_finish(c, &pd, true, out_idx, out_score, out_qid, cap, out_len);
}

// =================================================================================================
// MetaStore
// =================================================================================================
namespace otters {

struct MetaColumn {
    std::string name;
    int32_t dtype = 0;
    // device
    void* d_values = nullptr;
    uint32_t* d_nulls = nullptr;
    void* d_zmin = nullptr;
    void* d_zmax = nullptr;
    uint32_t* d_non_null = nullptr;
    uint64_t* d_bloom = nullptr;
    uint64_t bloom_stride = 0;
    uint32_t bloom_k0 = 0;              // probes of a full chunk's filter
    uint64_t* d_bloom_mbits = nullptr;
    uint32_t* d_bloom_k = nullptr;
    // host copies kept for the parity exports
    std::vector<int64_t> zmin_i, zmax_i;
    std::vector<double> zmin_f, zmax_f;
    std::vector<uint32_t> non_null;
    // string dictionary
    std::unordered_map<std::string, uint32_t> dict;
    std::vector<std::string> dict_strings;  // code -> bytes (result-column gather)
    size_t value_bytes = 0;
};

// Bloom filter spec of this repo (fastbloom's layout is not reproducible offline)
static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void bloom_hash(const uint8_t* s, uint64_t len, uint64_t* h1, uint64_t* h2) {
    uint64_t h = 0xCBF29CE484222325ull;
    for (uint64_t i = 0; i < len; ++i) {
        h ^= s[i];
        h *= 0x100000001B3ull;
    }
    *h1 = mix64(h);
    *h2 = mix64(h ^ 0x9E3779B97F4A7C15ull) | 1ull;
}
static void bloom_params(uint64_t n_items, int mode, double fpr, uint64_t bits, uint64_t* m_bits, uint32_t* k_hashes) {
    const uint64_t n = n_items ? n_items : 1;
    uint64_t m;
    if (mode == 0) {
        const double ln2 = 0.6931471805599453;
        double mm = ceil(-(double)n * log(fpr) / (ln2 * ln2));
        m = mm < 64.0 ? 64 : (uint64_t)mm;
    } else {
        m = bits < 64 ? 64 : bits;
    }
    m = (m + 63) / 64 * 64;
    double kk = floor((double)m / (double)n * 0.6931471805599453 + 0.5);
    *k_hashes = kk < 1.0 ? 1u : (kk > 16.0 ? 16u : (uint32_t)kk);
    *m_bits = m;
}

// Rust `as` casts used by the reference when it narrows literals (src/meta_compute.rs:244-288)
static inline int32_t i64_as_i32(int64_t v) { return (int32_t)(uint32_t)(uint64_t)v; }

}  // namespace otters

struct otters_metastore {
    otters_ctx* ctx = nullptr;
    VecStorage st;
    uint64_t chunk_size = 1024;
    uint64_t n_chunks = 0;
    std::vector<MetaColumn> cols;
    DevColumn* d_cols = nullptr;
    bool has_stats = false;
    otters_query_stats last{};
    std::vector<uint8_t> user_blob;  // opaque caller bytes of the file this store was loaded from (otters_metastore_load)
};

extern "C" int otters_metastore_destroy(otters_metastore* ms) {
    if (!ms) return OTTERS_OK;
    DeviceGuard g(ms->ctx->device);
    cudaStreamSynchronize(ms->ctx->stream);
    for (auto& c : ms->cols) {
        cudaFree(c.d_values);
        cudaFree(c.d_nulls);
        cudaFree(c.d_zmin);
        cudaFree(c.d_zmax);
        cudaFree(c.d_non_null);
        cudaFree(c.d_bloom);
        cudaFree(c.d_bloom_mbits);
        cudaFree(c.d_bloom_k);
    }
    cudaFree(ms->d_cols);
    ms->st.release();
    delete ms;
    return OTTERS_OK;
}

namespace otters {

template <typename T>
static int upload(T** dptr, const void* src, size_t bytes) {
    *dptr = nullptr;
    if (cudaMalloc((void**)dptr, std::max<size_t>(bytes, 16)) != cudaSuccess)
        return fail(OTTERS_ERR_NOMEM, "device allocation for metadata failed");
    if (bytes) OTTERS_CUDA(cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice));
    return OTTERS_OK;
}

static inline bool host_is_null(const otters_column& c, uint64_t row) {
    return c.null_words ? ((c.null_words[row >> 6] >> (row & 63)) & 1ull) != 0 : false;
}

// build_zone_stat_for_range (reference src/meta_compute.rs:32-132) + packed ranges (src/meta.rs:237-271)
static int build_column(otters_metastore* ms, const otters_column& in, const otters_build_params* p, double bloom_fpr,
                        uint64_t bloom_bits, MetaColumn* mc) {
    const uint64_t n = p->n_rows, cs = ms->chunk_size, nc = ms->n_chunks;
    mc->name = in.name ? in.name : "";
    mc->dtype = in.dtype;
    mc->non_null.assign(nc, 0);
    int rc;
    // null bitmap: same bytes, viewed as 32-bit words on the device
    if (in.null_words) {
        const size_t w64 = (n + 63) / 64;
        rc = upload(&mc->d_nulls, in.null_words, w64 * 8);
        if (rc) return rc;
    }
    auto chunk_range = [&](uint64_t ch, uint64_t* s, uint64_t* e) {
        *s = ch * cs;
        *e = std::min<uint64_t>(*s + cs, n);
    };
    switch (in.dtype) {
    case OTTERS_DTYPE_INT32: {
        if (n && !in.values) return fail(OTTERS_ERR_INVALID, "expected Int32 column");
        const int32_t* v = (const int32_t*)in.values;
        std::vector<int32_t> mn(nc), mx(nc);
        mc->zmin_i.resize(nc);
        mc->zmax_i.resize(nc);
        for (uint64_t ch = 0; ch < nc; ++ch) {
            uint64_t s, e;
            chunk_range(ch, &s, &e);
            int64_t lo = INT64_MAX, hi = INT64_MIN;
            uint32_t cnt = 0;
            for (uint64_t i = s; i < e; ++i)
                if (!host_is_null(in, i)) {
                    lo = std::min<int64_t>(lo, v[i]);
                    hi = std::max<int64_t>(hi, v[i]);
                    ++cnt;
                }
            mn[ch] = i64_as_i32(lo);  // `*min as i32` (src/meta.rs:254-255)
            mx[ch] = i64_as_i32(hi);
            mc->zmin_i[ch] = mn[ch];
            mc->zmax_i[ch] = mx[ch];
            mc->non_null[ch] = cnt;
        }
        if ((rc = upload(&mc->d_values, v, n * 4))) return rc;
        if ((rc = upload(&mc->d_zmin, mn.data(), nc * 4))) return rc;
        if ((rc = upload(&mc->d_zmax, mx.data(), nc * 4))) return rc;
        mc->value_bytes = 4;
        break;
    }
    case OTTERS_DTYPE_INT64:
    case OTTERS_DTYPE_DATETIME: {
        if (n && !in.values) return fail(OTTERS_ERR_INVALID, "expected Int64/DateTime column");
        const int64_t* v = (const int64_t*)in.values;
        mc->zmin_i.resize(nc);
        mc->zmax_i.resize(nc);
        for (uint64_t ch = 0; ch < nc; ++ch) {
            uint64_t s, e;
            chunk_range(ch, &s, &e);
            int64_t lo = INT64_MAX, hi = INT64_MIN;
            uint32_t cnt = 0;
            for (uint64_t i = s; i < e; ++i)
                if (!host_is_null(in, i)) {
                    lo = std::min(lo, v[i]);
                    hi = std::max(hi, v[i]);
                    ++cnt;
                }
            mc->zmin_i[ch] = lo;
            mc->zmax_i[ch] = hi;
            mc->non_null[ch] = cnt;
        }
        if ((rc = upload(&mc->d_values, v, n * 8))) return rc;
        if ((rc = upload(&mc->d_zmin, mc->zmin_i.data(), nc * 8))) return rc;
        if ((rc = upload(&mc->d_zmax, mc->zmax_i.data(), nc * 8))) return rc;
        mc->value_bytes = 8;
        break;
    }
    case OTTERS_DTYPE_FLOAT32: {
        if (n && !in.values) return fail(OTTERS_ERR_INVALID, "expected Float32 column");
        const float* v = (const float*)in.values;
        std::vector<float> mn(nc), mx(nc);
        mc->zmin_f.resize(nc);
        mc->zmax_f.resize(nc);
        for (uint64_t ch = 0; ch < nc; ++ch) {
            uint64_t s, e;
            chunk_range(ch, &s, &e);
            double lo = INFINITY, hi = -INFINITY;  // f64::min/max ignore NaN (src/meta_compute.rs:69-83)
            uint32_t cnt = 0;
            for (uint64_t i = s; i < e; ++i)
                if (!host_is_null(in, i)) {
                    lo = fmin(lo, (double)v[i]);
                    hi = fmax(hi, (double)v[i]);
                    ++cnt;
                }
            mn[ch] = (float)lo;
            mx[ch] = (float)hi;
            mc->zmin_f[ch] = mn[ch];
            mc->zmax_f[ch] = mx[ch];
            mc->non_null[ch] = cnt;
        }
        if ((rc = upload(&mc->d_values, v, n * 4))) return rc;
        if ((rc = upload(&mc->d_zmin, mn.data(), nc * 4))) return rc;
        if ((rc = upload(&mc->d_zmax, mx.data(), nc * 4))) return rc;
        mc->value_bytes = 4;
        break;
    }
    case OTTERS_DTYPE_FLOAT64: {
        if (n && !in.values) return fail(OTTERS_ERR_INVALID, "expected Float64 column");
        const double* v = (const double*)in.values;
        mc->zmin_f.resize(nc);
        mc->zmax_f.resize(nc);
        for (uint64_t ch = 0; ch < nc; ++ch) {
            uint64_t s, e;
            chunk_range(ch, &s, &e);
            double lo = INFINITY, hi = -INFINITY;
            uint32_t cnt = 0;
            for (uint64_t i = s; i < e; ++i)
                if (!host_is_null(in, i)) {
                    lo = fmin(lo, v[i]);
                    hi = fmax(hi, v[i]);
                    ++cnt;
                }
            mc->zmin_f[ch] = lo;
            mc->zmax_f[ch] = hi;
            mc->non_null[ch] = cnt;
        }
        if ((rc = upload(&mc->d_values, v, n * 8))) return rc;
        if ((rc = upload(&mc->d_zmin, mc->zmin_f.data(), nc * 8))) return rc;
        if ((rc = upload(&mc->d_zmax, mc->zmax_f.data(), nc * 8))) return rc;
        mc->value_bytes = 8;
        break;
    }
    case OTTERS_DTYPE_STRING: {
        if (n && (!in.str_offsets || (!in.str_bytes && in.str_offsets[n] != 0)))
            return fail(OTTERS_ERR_INVALID, "expected String column");
        // dictionary codes (device compares u32 codes; byte equality <=> code equality)
        std::vector<uint32_t> codes(n);
        for (uint64_t i = 0; i < n; ++i) {
            if (host_is_null(in, i)) {
                codes[i] = 0xFFFFFFFFu;
                continue;
            }
            std::string sv((const char*)in.str_bytes + in.str_offsets[i], in.str_offsets[i + 1] - in.str_offsets[i]);
            auto it = mc->dict.find(sv);
            if (it == mc->dict.end()) {
                mc->dict_strings.push_back(sv);
                it = mc->dict.emplace(std::move(sv), (uint32_t)mc->dict.size()).first;
            }
            codes[i] = it->second;
        }
        // per-chunk Bloom filters sized for the chunk length (src/meta_compute.rs:99-116)
        uint64_t m0;
        uint32_t k0;
        bloom_params(std::min<uint64_t>(cs, std::max<uint64_t>(n, 1)), p->bloom_mode, bloom_fpr, bloom_bits, &m0, &k0);
        mc->bloom_stride = m0 / 64;
        mc->bloom_k0 = k0;
        std::vector<uint64_t> words(std::max<uint64_t>(nc, 1) * mc->bloom_stride, 0);
        std::vector<uint64_t> mbits(std::max<uint64_t>(nc, 1), 64);
        std::vector<uint32_t> kh(std::max<uint64_t>(nc, 1), 1);
        for (uint64_t ch = 0; ch < nc; ++ch) {
            uint64_t s, e;
            chunk_range(ch, &s, &e);
            bloom_params(e - s, p->bloom_mode, bloom_fpr, bloom_bits, &mbits[ch], &kh[ch]);
            uint64_t* w = words.data() + ch * mc->bloom_stride;
            uint32_t cnt = 0;
            for (uint64_t i = s; i < e; ++i)
                if (!host_is_null(in, i)) {
                    uint64_t h1, h2;
                    bloom_hash(in.str_bytes + in.str_offsets[i], in.str_offsets[i + 1] - in.str_offsets[i], &h1, &h2);
                    uint64_t bit = h1 % mbits[ch], step = h2 % mbits[ch];  // probe j = (a + j*b) mod m, stepped
                    if (step == 0) step = 1;
                    for (uint32_t j = 0; j < kh[ch]; ++j) {
                        w[bit >> 6] |= 1ull << (bit & 63);
                        bit += step;
                        if (bit >= mbits[ch]) bit -= mbits[ch];
                    }
                    ++cnt;
                }
            mc->non_null[ch] = cnt;
        }
        if ((rc = upload(&mc->d_values, codes.data(), n * 4))) return rc;
        if ((rc = upload(&mc->d_bloom, words.data(), words.size() * 8))) return rc;
        if ((rc = upload(&mc->d_bloom_mbits, mbits.data(), mbits.size() * 8))) return rc;
        if ((rc = upload(&mc->d_bloom_k, kh.data(), kh.size() * 4))) return rc;
        mc->value_bytes = 4;
        break;
    }
    default: return fail(OTTERS_ERR_INVALID, "unknown column dtype");
    }
    if ((rc = upload(&mc->d_non_null, mc->non_null.data(), nc * 4))) return rc;
    return OTTERS_OK;
}

// ---- device-side column build (build.cu): same tables as build_column above, computed by kernels over the uploaded column ----
struct DevTmp {  // frees temporary device buffers on scope exit
    std::vector<void*> ptrs;
    template <typename T>
    int alloc(T** p, size_t bytes) {
        *p = nullptr;
        if (cudaMalloc((void**)p, std::max<size_t>(bytes, 16)) != cudaSuccess) return fail(OTTERS_ERR_NOMEM, "device allocation for the store build failed");
        ptrs.push_back(*p);
        return OTTERS_OK;
    }
    ~DevTmp() {
        for (void* p : ptrs) cudaFree(p);
    }
};

template <typename T>
static int dev_alloc(T** dptr, size_t bytes) {
    *dptr = nullptr;
    if (cudaMalloc((void**)dptr, std::max<size_t>(bytes, 16)) != cudaSuccess) return fail(OTTERS_ERR_NOMEM, "device allocation for metadata failed");
    return OTTERS_OK;
}

// returns OTTERS_OK with *used = false when the column must take the host path (hash collision in the dictionary)
static int build_column_device(otters_metastore* ms, const otters_column& in, const otters_build_params* p, double bloom_fpr,
                               uint64_t bloom_bits, MetaColumn* mc, bool* used) {
    *used = true;
    const uint64_t n = p->n_rows, cs = ms->chunk_size, nc = ms->n_chunks;
    cudaStream_t s = ms->ctx->stream;
    mc->name = in.name ? in.name : "";
    mc->dtype = in.dtype;
    mc->non_null.assign(nc, 0);
    int rc;
    if (in.null_words) {
        if ((rc = upload(&mc->d_nulls, in.null_words, (n + 63) / 64 * 8))) return rc;
    }
    if ((rc = dev_alloc(&mc->d_non_null, nc * 4))) return rc;
    if (in.dtype != OTTERS_DTYPE_STRING) {
        const size_t w = (in.dtype == OTTERS_DTYPE_INT32 || in.dtype == OTTERS_DTYPE_FLOAT32) ? 4 : 8;
        if (in.dtype < 0 || in.dtype > OTTERS_DTYPE_DATETIME) return fail(OTTERS_ERR_INVALID, "unknown column dtype");
        if (n && !in.values) return fail(OTTERS_ERR_INVALID, "column values are missing");
        if ((rc = upload(&mc->d_values, in.values, n * w))) return rc;
        if ((rc = dev_alloc(&mc->d_zmin, nc * w))) return rc;
        if ((rc = dev_alloc(&mc->d_zmax, nc * w))) return rc;
        if ((rc = launch_zonemap(in.dtype, mc->d_values, mc->d_nulls, n, cs, nc, mc->d_zmin, mc->d_zmax, mc->d_non_null, s))) return rc;
        mc->value_bytes = w;
        // host copies of the (small) tables for the parity exports
        std::vector<uint8_t> hmin(nc * w), hmax(nc * w);
        OTTERS_CUDA(cudaMemcpyAsync(hmin.data(), mc->d_zmin, nc * w, cudaMemcpyDeviceToHost, s));
        OTTERS_CUDA(cudaMemcpyAsync(hmax.data(), mc->d_zmax, nc * w, cudaMemcpyDeviceToHost, s));
        OTTERS_CUDA(cudaMemcpyAsync(mc->non_null.data(), mc->d_non_null, nc * 4, cudaMemcpyDeviceToHost, s));
        OTTERS_CUDA(cudaStreamSynchronize(s));
        const bool is_f = in.dtype == OTTERS_DTYPE_FLOAT32 || in.dtype == OTTERS_DTYPE_FLOAT64;
        if (is_f) {
            mc->zmin_f.resize(nc);
            mc->zmax_f.resize(nc);
        } else {
            mc->zmin_i.resize(nc);
            mc->zmax_i.resize(nc);
        }
        for (uint64_t ch = 0; ch < nc; ++ch) {
            switch (in.dtype) {
            case OTTERS_DTYPE_INT32: mc->zmin_i[ch] = ((int32_t*)hmin.data())[ch]; mc->zmax_i[ch] = ((int32_t*)hmax.data())[ch]; break;
            case OTTERS_DTYPE_FLOAT32: mc->zmin_f[ch] = ((float*)hmin.data())[ch]; mc->zmax_f[ch] = ((float*)hmax.data())[ch]; break;
            case OTTERS_DTYPE_FLOAT64: mc->zmin_f[ch] = ((double*)hmin.data())[ch]; mc->zmax_f[ch] = ((double*)hmax.data())[ch]; break;
            default: mc->zmin_i[ch] = ((int64_t*)hmin.data())[ch]; mc->zmax_i[ch] = ((int64_t*)hmax.data())[ch]; break;
            }
        }
        return OTTERS_OK;
    }
    // ---- String: hashes -> Bloom filters + dictionary codes -------------------------------------------------------------
    if (n && (!in.str_offsets || (!in.str_bytes && in.str_offsets[n] != 0))) return fail(OTTERS_ERR_INVALID, "expected String column");
    if (n >= 0xFFFFFFF0ull) return fail(OTTERS_ERR_UNSUPPORTED, "a store shard is limited to 2^32-16 rows");
    DevTmp tmp;
    uint8_t* d_bytes = nullptr;
    uint64_t *d_offs = nullptr, *d_hash = nullptr;
    const uint64_t n_bytes = n ? in.str_offsets[n] : 0;
    if ((rc = tmp.alloc(&d_bytes, n_bytes))) return rc;
    if ((rc = tmp.alloc(&d_offs, (n + 1) * 8))) return rc;
    if ((rc = tmp.alloc(&d_hash, n * 8))) return rc;
    if (n_bytes) OTTERS_CUDA(cudaMemcpyAsync(d_bytes, in.str_bytes, n_bytes, cudaMemcpyHostToDevice, s));
    if (n) OTTERS_CUDA(cudaMemcpyAsync(d_offs, in.str_offsets, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    if ((rc = launch_string_hash(d_bytes, d_offs, mc->d_nulls, n, d_hash, s))) return rc;
    // Bloom geometry per chunk (host: one entry per chunk), filters built on the device
    uint64_t m0;
    uint32_t k0;
    bloom_params(std::min<uint64_t>(cs, std::max<uint64_t>(n, 1)), p->bloom_mode, bloom_fpr, bloom_bits, &m0, &k0);
    mc->bloom_stride = m0 / 64;
    mc->bloom_k0 = k0;
    std::vector<uint64_t> mbits(std::max<uint64_t>(nc, 1), 64);
    std::vector<uint32_t> kh(std::max<uint64_t>(nc, 1), 1);
    for (uint64_t ch = 0; ch < nc; ++ch) {
        const uint64_t cs0 = ch * cs, ce = std::min<uint64_t>(cs0 + cs, n);
        bloom_params(ce - cs0, p->bloom_mode, bloom_fpr, bloom_bits, &mbits[ch], &kh[ch]);
    }
    const size_t n_words = std::max<uint64_t>(nc, 1) * mc->bloom_stride;
    if ((rc = dev_alloc(&mc->d_bloom, n_words * 8))) return rc;
    OTTERS_CUDA(cudaMemsetAsync(mc->d_bloom, 0, std::max<size_t>(n_words * 8, 16), s));
    if ((rc = upload(&mc->d_bloom_mbits, mbits.data(), mbits.size() * 8))) return rc;
    if ((rc = upload(&mc->d_bloom_k, kh.data(), kh.size() * 4))) return rc;
    if ((rc = launch_bloom_build(d_hash, mc->d_nulls, n, cs, nc, mc->d_bloom_mbits, mc->d_bloom_k, mc->bloom_stride, mc->d_bloom, mc->d_non_null, s)))
        return rc;
    // dictionary: distinct hashes claim table slots; the host numbers them in order of first occurrence
    const uint64_t T = std::max<uint64_t>(pow2_at_least(2 * std::max<uint64_t>(n, 1)), 1024);
    uint64_t* d_keys = nullptr;
    uint32_t *d_rep = nullptr, *d_slot_code = nullptr, *d_list_slot = nullptr, *d_list_rep = nullptr, *d_flags = nullptr;
    const uint32_t list_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(n, 1), 0xFFFFFFF0ull);
    if ((rc = tmp.alloc(&d_keys, T * 8))) return rc;
    if ((rc = tmp.alloc(&d_rep, T * 4))) return rc;
    if ((rc = tmp.alloc(&d_slot_code, T * 4))) return rc;
    if ((rc = tmp.alloc(&d_list_slot, (size_t)list_cap * 4))) return rc;
    if ((rc = tmp.alloc(&d_list_rep, (size_t)list_cap * 4))) return rc;
    if ((rc = tmp.alloc(&d_flags, 16))) return rc;
    OTTERS_CUDA(cudaMemsetAsync(d_keys, 0xFF, T * 8, s));
    OTTERS_CUDA(cudaMemsetAsync(d_rep, 0xFF, T * 4, s));
    OTTERS_CUDA(cudaMemsetAsync(d_flags, 0, 16, s));
    if ((rc = launch_dict_insert(d_hash, mc->d_nulls, n, d_keys, d_rep, T, s))) return rc;
    if ((rc = launch_dict_collect(d_keys, d_rep, T, d_list_slot, d_list_rep, d_flags, list_cap, s))) return rc;
    uint32_t n_distinct = 0;
    OTTERS_CUDA(cudaMemcpyAsync(&n_distinct, d_flags, 4, cudaMemcpyDeviceToHost, s));
    OTTERS_CUDA(cudaStreamSynchronize(s));
    if (n_distinct > list_cap) return fail(OTTERS_ERR_UNSUPPORTED, "dictionary build: inconsistent distinct count");
    std::vector<uint32_t> l_slot(n_distinct), l_rep(n_distinct), order(n_distinct), l_code(n_distinct);
    if (n_distinct) {
        OTTERS_CUDA(cudaMemcpy(l_slot.data(), d_list_slot, (size_t)n_distinct * 4, cudaMemcpyDeviceToHost));
        OTTERS_CUDA(cudaMemcpy(l_rep.data(), d_list_rep, (size_t)n_distinct * 4, cudaMemcpyDeviceToHost));
    }
    for (uint32_t i = 0; i < n_distinct; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return l_rep[a] < l_rep[b]; });
    mc->dict.clear();
    mc->dict_strings.clear();
    mc->dict_strings.reserve(n_distinct);
    bool collision = false;
    for (uint32_t code = 0; code < n_distinct; ++code) {
        const uint32_t i = order[code], r = l_rep[i];
        l_code[i] = code;
        std::string sv((const char*)in.str_bytes + in.str_offsets[r], in.str_offsets[r + 1] - in.str_offsets[r]);
        if (!mc->dict.emplace(sv, code).second) collision = true;  // (two slots, one string: impossible — same hash, same slot)
        mc->dict_strings.push_back(std::move(sv));
    }
    uint32_t* d_list_code = nullptr;
    if ((rc = tmp.alloc(&d_list_code, (size_t)std::max<uint32_t>(n_distinct, 1) * 4))) return rc;
    if (n_distinct) OTTERS_CUDA(cudaMemcpyAsync(d_list_code, l_code.data(), (size_t)n_distinct * 4, cudaMemcpyHostToDevice, s));
    if ((rc = dev_alloc((uint32_t**)&mc->d_values, n * 4))) return rc;
    if ((rc = launch_dict_encode(d_list_slot, d_list_code, n_distinct, d_slot_code, d_hash, mc->d_nulls, d_bytes, d_offs, n, d_keys, d_rep, T,
                                 (uint32_t*)mc->d_values, d_flags + 1, s)))
        return rc;
    uint32_t mismatch = 0;
    OTTERS_CUDA(cudaMemcpyAsync(&mismatch, d_flags + 1, 4, cudaMemcpyDeviceToHost, s));
    OTTERS_CUDA(cudaMemcpyAsync(mc->non_null.data(), mc->d_non_null, nc * 4, cudaMemcpyDeviceToHost, s));
    OTTERS_CUDA(cudaStreamSynchronize(s));
    mc->value_bytes = 4;
    if (mismatch || collision) {  // two different strings share a 64-bit hash: let the host build this column
        *used = false;
        cudaFree(mc->d_values); cudaFree(mc->d_nulls); cudaFree(mc->d_non_null); cudaFree(mc->d_bloom); cudaFree(mc->d_bloom_mbits); cudaFree(mc->d_bloom_k);
        *mc = MetaColumn{};
    }
    return OTTERS_OK;
}

}  // namespace otters

extern "C" int otters_metastore_build(otters_ctx* c, const otters_build_params* p, otters_metastore** out,
                                      otters_build_stats* stats) {
    if (!c || !p || !out) return fail(OTTERS_ERR_INVALID, "null argument");
    DeviceGuard g(c->device);
    const double t_build = now_s();
    if (p->vectors_kind != OTTERS_VECTORS_SYNTHETIC && p->n_rows && !p->vectors)
        return fail(OTTERS_ERR_INVALID, "vectors must be provided to build MetaStore");  // src/meta.rs:153-155
    if (p->dim == 0 && p->n_rows > 0) return fail(OTTERS_ERR_INVALID, "vector dimension cannot be zero");  // :177-179
    if (p->n_columns && !p->columns) return fail(OTTERS_ERR_INVALID, "null columns");
    std::unique_ptr<otters_metastore> ms(new otters_metastore());
    ms->ctx = c;
    ms->chunk_size = std::max<uint64_t>(p->chunk_size, 1);  // src/meta.rs:86-89
    if (ms->chunk_size > 0xFFFFFFFFull) ms->chunk_size = 0xFFFFFFFFull;
    ms->n_chunks = (p->n_rows + ms->chunk_size - 1) / ms->chunk_size;
    double fpr = p->bloom_fpr;
    uint64_t bits = p->bloom_bits;
    if (p->bloom_mode == 0) {  // src/meta.rs:92-101
        if (!std::isfinite(fpr)) fpr = 0.01;
        fpr = std::min(std::max(fpr, 1e-2), 0.5);
    } else {
        bits = std::max<uint64_t>(bits, 64);  // src/meta.rs:106-110
    }
    if (p->vector_format != OTTERS_VECTORS_FMT_F32 && p->vector_format != OTTERS_VECTORS_FMT_BF16)
        return fail(OTTERS_ERR_INVALID, "unknown vector format");
    ms->st.ctx = c;
    ms->st.set_format(p->dim, p->vector_format == OTTERS_VECTORS_FMT_BF16);

    auto cleanup = [&](int rc) {
        otters_metastore_destroy(ms.release());
        return rc;
    };
    // vectors ingest (VecStore::add_vectors per chunk in the reference, src/meta.rs:209-212)
    const double t_ing = now_s();
    int rc = OTTERS_OK;
    if (p->vectors_kind == OTTERS_VECTORS_HOST) rc = ms->st.add(p->vectors, p->n_rows, cudaMemcpyHostToDevice);
    else if (p->vectors_kind == OTTERS_VECTORS_DEVICE) rc = ms->st.add(p->vectors, p->n_rows, cudaMemcpyDeviceToDevice);
    else if (p->vectors_kind == OTTERS_VECTORS_SYNTHETIC) {
        ShardMap m;
        if (p->synthetic_map) {
            const otters_shard_map* in = (const otters_shard_map*)p->synthetic_map;
            m.row_base = in->row_base;
            m.world = in->world;
            m.rank = in->rank;
            m.block = in->block_rows;
        } else {
            m.row_base = p->synthetic_first_row;
        }
        rc = ms->st.add_synth(m, p->n_rows, p->synthetic_seed);
    }
    else rc = fail(OTTERS_ERR_INVALID, "invalid vectors_kind");
    if (rc) return cleanup(rc);
    const double ingest_s = now_s() - t_ing;

    // zonemaps, Bloom filters, dictionary codes
    const double t_zone = now_s();
    ms->cols.resize(p->n_columns);
    std::vector<DevColumn> dcols(std::max<uint32_t>(p->n_columns, 1));
    for (uint32_t i = 0; i < p->n_columns; ++i) {
        // tables are built by kernels over the uploaded column (build.cu); OTTERS_BUILD_HOST=1 keeps the host loops
        // (the two are bit-identical: tests/test_gpu_build.py)
        const char* host_env = getenv("OTTERS_BUILD_HOST");
        bool on_device = !(host_env && atoi(host_env) != 0);
        if (on_device) {
            rc = build_column_device(ms.get(), p->columns[i], p, fpr, bits, &ms->cols[i], &on_device);
            if (rc) return cleanup(rc);
        }
        if (!on_device) rc = build_column(ms.get(), p->columns[i], p, fpr, bits, &ms->cols[i]);
        if (rc) return cleanup(rc);
        const MetaColumn& mc = ms->cols[i];
        DevColumn& d = dcols[i];
        d.dtype = mc.dtype;
        d.values = mc.d_values;
        d.null_words = mc.d_nulls;
        d.zmin = mc.d_zmin;
        d.zmax = mc.d_zmax;
        d.non_null = mc.d_non_null;
        d.bloom = mc.d_bloom;
        d.bloom_stride = mc.bloom_stride;
        d.bloom_mbits = mc.d_bloom_mbits;
        d.bloom_k = mc.d_bloom_k;
    }
    if ((rc = upload(&ms->d_cols, dcols.data(), dcols.size() * sizeof(DevColumn)))) return cleanup(rc);
    const double zone_s = now_s() - t_zone;

    if (stats) {  // src/meta.rs:292-299
        stats->n_rows = p->n_rows;
        stats->dim = p->dim;
        stats->n_chunks = ms->n_chunks;
        stats->vectors_ingest_s = ingest_s;
        stats->zonemap_build_s = zone_s;
        stats->build_total_s = now_s() - t_build;
    }
    *out = ms.release();
    return OTTERS_OK;
}

extern "C" int otters_metastore_set_rows(otters_metastore* ms, const uint64_t* rows, const float* data, uint64_t n) {
    if (!ms) return fail(OTTERS_ERR_INVALID, "null store");
    if (n && (!rows || !data)) return fail(OTTERS_ERR_INVALID, "null rows");
    DeviceGuard g(ms->ctx->device);
    OTTERS_CUDA(otters_ctx_synchronize(ms->ctx) == OTTERS_OK ? cudaSuccess : cudaErrorUnknown);  // no query may be in flight
    return ms->st.set_rows(rows, data, n);
}

extern "C" uint64_t otters_metastore_n_chunks(const otters_metastore* ms) { return ms ? ms->n_chunks : 0; }
extern "C" uint64_t otters_metastore_chunk_size(const otters_metastore* ms) { return ms ? ms->chunk_size : 0; }
extern "C" uint64_t otters_metastore_len(const otters_metastore* ms) { return ms ? ms->st.n : 0; }

namespace otters {

// Lowers the caller's CompiledFilter into device leaves: literal casts follow the reference
// (src/meta.rs:431-521 for zonemaps, src/meta_compute.rs:244-288 for rows).  Leaf/column type
// combinations that Expr::compile can never produce (src/expr.rs:385-466) are rejected.
static int lower_filter(otters_metastore* ms, const otters_filter* f, std::vector<uint32_t>* offs, std::vector<DevLeaf>* leaves) {
    offs->clear();
    leaves->clear();
    if (!f->clause_offsets || (!f->leaves && f->n_clauses && f->clause_offsets[f->n_clauses]))
        return fail(OTTERS_ERR_INVALID, "malformed filter");
    for (uint32_t ci = 0; ci <= f->n_clauses; ++ci) offs->push_back(f->clause_offsets[ci]);
    const uint32_t nl = f->clause_offsets[f->n_clauses];
    for (uint32_t li = 0; li < nl; ++li) {
        const otters_leaf& in = f->leaves[li];
        if (in.col >= ms->cols.size()) return fail(OTTERS_ERR_INVALID, "Unknown column '" + std::to_string(in.col) + "'");
        if (in.op < 0 || in.op > 5) return fail(OTTERS_ERR_INVALID, "invalid comparison operator");
        MetaColumn& mc = ms->cols[in.col];
        DevLeaf d{};
        d.col = in.col;
        d.op = in.op;
        {   // Eq, Neq, Lt, Lte, Gt, Gte over the states {>, <, ==, unordered}
            static const uint32_t kTruth[6] = {0x4u, 0xBu, 0x2u, 0x6u, 0x1u, 0x5u};
            d.tt = kTruth[in.op];
        }
        d.values = mc.d_values;
        d.null_words = mc.d_nulls;
        d.zmin = mc.d_zmin;
        d.zmax = mc.d_zmax;
        d.non_null = mc.d_non_null;
        d.bloom = mc.d_bloom;
        d.bloom_stride = mc.bloom_stride;
        d.bloom_mbits = mc.d_bloom_mbits;
        d.bloom_k = mc.d_bloom_k;
        switch (mc.dtype) {
        case OTTERS_DTYPE_INT32:
            if (in.kind != OTTERS_LIT_I64)
                return fail(OTTERS_ERR_INVALID, "Type mismatch for column '" + mc.name + "': expected Int32, got literal " +
                                                    (in.kind == OTTERS_LIT_F64 ? "float" : "string"));
            d.exec = LEAF_I32;
            d.i32 = i64_as_i32(in.i);
            break;
        case OTTERS_DTYPE_INT64:
        case OTTERS_DTYPE_DATETIME:
            if (in.kind != OTTERS_LIT_I64)
                return fail(OTTERS_ERR_INVALID, "Type mismatch for column '" + mc.name + "': expected " +
                                                    (mc.dtype == OTTERS_DTYPE_INT64 ? "Int64" : "DateTime") + ", got literal " +
                                                    (in.kind == OTTERS_LIT_F64 ? "float" : "string"));
            d.exec = LEAF_I64;
            d.i64 = in.i;
            break;
        case OTTERS_DTYPE_FLOAT32:
            if (in.kind != OTTERS_LIT_F64)
                return fail(OTTERS_ERR_INVALID, "Type mismatch for column '" + mc.name + "': expected Float32 literal widened to f64");
            d.exec = LEAF_F32;
            d.f32 = (float)in.f;
            break;
        case OTTERS_DTYPE_FLOAT64:
            if (in.kind != OTTERS_LIT_F64)
                return fail(OTTERS_ERR_INVALID, "Type mismatch for column '" + mc.name + "': expected Float64 literal widened to f64");
            d.exec = LEAF_F64;
            d.f64 = in.f;
            break;
        case OTTERS_DTYPE_STRING: {
            if (in.kind != OTTERS_LIT_STR)
                return fail(OTTERS_ERR_INVALID, "Type mismatch for column '" + mc.name + "': expected String, got literal string");
            if (in.op != OTTERS_OP_EQ && in.op != OTTERS_OP_NEQ)
                return fail(OTTERS_ERR_INVALID, "Unsupported comparator for string column '" + mc.name + "'");
            d.exec = LEAF_STR;
            std::string lit((const char*)in.s, in.slen);
            auto it = mc.dict.find(lit);
            d.code_valid = it != mc.dict.end();
            d.code = d.code_valid ? it->second : 0;
            bloom_hash((const uint8_t*)lit.data(), lit.size(), &d.h1, &d.h2);
            // every full chunk has the same filter size m0: its start / step are precomputed so that the prune kernel
            // only divides for the (shorter) last chunk
            d.bloom_m0 = mc.bloom_stride * 64;
            d.bloom_a0 = d.bloom_m0 ? d.h1 % d.bloom_m0 : 0;
            d.bloom_b0 = d.bloom_m0 ? d.h2 % d.bloom_m0 : 0;
            if (d.bloom_b0 == 0) d.bloom_b0 = 1;
            d.bloom_k0 = mc.bloom_k0;
            d.bloom_full_chunks = ms->st.n / ms->chunk_size;
            break;
        }
        default: return fail(OTTERS_ERR_INVALID, "unknown column dtype");
        }
        leaves->push_back(d);
    }
    return OTTERS_OK;
}

// What a query needs from the stand-alone metadata kernels.
enum MetaMode {
    META_LAZY = 0,     // nothing: the scan kernel prunes chunks (zonemaps + Bloom) and evaluates the row predicate itself
    META_PRUNE = 1,    // K0 only: chunk mask + statistics (rows are evaluated inside the scan kernel, or nothing is scanned)
    META_ROWMASK = 2,  // K0 + K0b: chunk mask, statistics and the row bitmask in ctx->d_meta_mask
};

static int ensure_meta_scratch(otters_ctx* c, const otters_metastore* ms, bool rows) {
    int rc = ensure_dev(&c->d_chunk_keep, &c->chunk_keep_words, (size_t)((ms->n_chunks + 31) / 32 + 1), c->stream);
    if (rc) return rc;
    if (rows) rc = ensure_dev(&c->d_meta_mask, &c->meta_mask_words, (size_t)((ms->st.n + 31) / 32 + 1), c->stream);
    return rc;
}

// Lowers the filter into the input image of the current query, sends the image, and enqueues the stand-alone metadata
// kernels `mode` asks for; leaves ctx->cur_filter pointing at the lowered filter on the device.
static int run_meta_filter(otters_ctx* c, otters_metastore* ms, const otters_filter* f, uint32_t nq, MetaMode mode) {
    cudaStream_t s = c->stream;
    if (!f) {
        // no meta_filter: every chunk is evaluated (src/meta.rs:658) — the statistics travel inside the control block image
        unsigned long long* hs = reinterpret_cast<unsigned long long*>(c->h_io[c->io_slot] + 32);
        hs[0] = ms->n_chunks;
        hs[1] = (unsigned long long)ms->st.n * nq;
        return io_flush(c);
    }
    std::vector<uint32_t> offs;
    std::vector<DevLeaf> leaves;
    int rc = lower_filter(ms, f, &offs, &leaves);
    if (rc) return rc;
    if (mode != META_LAZY) {
        rc = ensure_meta_scratch(c, ms, mode == META_ROWMASK);
        if (rc) return rc;
    }
    const size_t off_bytes = round_up(offs.size() * 4, 16);
    const size_t total = off_bytes + leaves.size() * sizeof(DevLeaf);
    uint8_t* hf = nullptr;
    size_t f_off = 0;
    rc = io_push(c, total + 16, &hf, &f_off);
    if (rc) return rc;
    memcpy(hf, offs.data(), offs.size() * 4);
    if (!leaves.empty()) memcpy(hf + off_bytes, leaves.data(), leaves.size() * sizeof(DevLeaf));
    rc = io_flush(c);  // control block reset + lowered filter + staged queries: one copy
    if (rc) return rc;
    FusedFilter& cf = c->cur_filter;
    cf.clause_off = reinterpret_cast<const uint32_t*>(c->d_io + f_off);
    cf.leaves = reinterpret_cast<const DevLeaf*>(c->d_io + f_off + off_bytes);
    cf.n_clauses = f->n_clauses;
    cf.n_leaves = (uint32_t)leaves.size();
    cf.chunk_keep = mode == META_LAZY ? nullptr : c->d_chunk_keep;
    cf.chunk_size = (uint32_t)ms->chunk_size;
    cf.n_chunks = (uint32_t)ms->n_chunks;
    if (mode == META_LAZY) return OTTERS_OK;
    MetaKernelParams mp{};
    mp.cols = ms->d_cols;
    mp.n_rows = (uint32_t)ms->st.n;
    mp.chunk_size = (uint32_t)ms->chunk_size;
    mp.n_chunks = (uint32_t)ms->n_chunks;
    mp.nq = nq;
    mp.chunk_keep = c->d_chunk_keep;
    mp.row_mask = c->d_meta_mask;
    mp.stats = c->d_stats;
    mp.clause_off = cf.clause_off;
    mp.leaves = cf.leaves;
    mp.n_clauses = f->n_clauses;
    rec_event(c, 0);
    rc = launch_prune(mp, (uint32_t)leaves.size(), s);
    if (rc) return rc;
    rec_event(c, 1);
    c->timed_meta = true;
    c->last.kernel_launches += 1;
    if (mode == META_ROWMASK) {
        rc = launch_rowmask(mp, (uint32_t)leaves.size(), s);
        if (rc) return rc;
        c->last.kernel_launches += 1;
        rec_event(c, 6);
        c->timed_rowmask = true;
    }
    return OTTERS_OK;
}

static int exchange_only(otters_ctx* c, int take_max, const unsigned long long* stats_src, QueryRun* run);

// ---- MetaQueryPlan::collect in two halves: enqueue (everything up to the last kernel launch) and finish (wait, copy the
// result out, assemble the statistics).  The blocking entry points run both back to back; otters_query_submit runs the
// first and otters_query_wait the second.
static int meta_enqueue(otters_ctx* c, otters_metastore* ms, const otters_vec_query* q, const otters_filter* filter,
                        otters_topk_record* d_records, ShardMap map, bool want_host, bool want_stats, Pending* pd) {
    *pd = Pending{};
    pd->t_submit = now_s();
    pd->meta = true;
    pd->ms = ms;
    int rc = validate_query(q, ms->st.dim, true);
    if (rc) return rc;
    if (q->row_mask_words) return fail(OTTERS_ERR_INVALID, "row masks are not part of MetaQueryPlan");
    cudaStream_t s = c->stream;
    rc = begin_query(c);
    if (rc) return rc;
    // durations of otters_query_stats come from per-phase events: recorded only when the caller waits for stats
    if (c->tuning.timing == 0 && want_stats && want_host) c->timing = true;
    c->want_host_result = want_host;
    // per-chunk collect() errors are swallowed by the reference (src/meta_compute.rs:182): an empty
    // batch or a wrong-dimension query returns no rows but still reports stats
    const bool chunk_err = q->nq == 0 || q->dim != ms->st.dim || !q->queries;
    const bool scan = !chunk_err && q->k > 0 && ms->st.n > 0;
    if (scan) {
        rc = stage_queries(c, q, ms->st.pitch);
        if (rc) return rc;
    }
    // the batched tensor-core kernel gates rows with a precomputed mask (K0b) instead of the fused predicate
    const bool batched = scan && !c->ex_active && batch_eligible(c, q, ms->st.n, std::min<uint64_t>(q->k, ms->st.n * (uint64_t)q->nq));
    const uint32_t n_leaves = filter && filter->clause_offsets ? filter->clause_offsets[filter->n_clauses] : 0;
    // Row predicate: by default its own kernel (K0b writes the surviving-row bitmask at HBM bandwidth, the scan's producer then
    // reads one mask word per 32 rows); the scan kernel can also evaluate the CNF itself per work unit (fused K0b:
    // disable_fused_predicate == 2, or lazy_prune) — one launch less, but every unit then pays dependent metadata round trips
    // inside the streaming warps: measured 3-6 % slower on every filtered workload (profiles/r2_predicate_ab.txt)
    const bool want_fused = c->tuning.disable_fused_predicate == 2 || c->tuning.lazy_prune;
    const bool fuse = scan && filter && !batched && want_fused &&
                      FusedFilter::smem_bytes_for(n_leaves, filter->n_clauses) <= kMaxFusedFilterBytes;
    // ... and can prune the chunks itself as well (lazy K0, opt-in: measured slower than the stand-alone kernel, whose 9-12 us
    // hide under the other lane's scan when two queries are in flight)
    const bool lazy = fuse && q->nq == 1 && c->tuning.lazy_prune;
    rc = run_meta_filter(c, ms, filter, q->nq, lazy ? META_LAZY : ((scan && filter && !fuse) ? META_ROWMASK : META_PRUNE));
    if (rc) return rc;
    pd->scan = scan;
    pd->take_max = q->take_type == OTTERS_TAKE_MAX;
    pd->nq = q->nq;
    pd->has_filter = filter != nullptr;
    if (filter) {
        for (uint32_t li = 0; li < n_leaves; ++li) {
            const MetaColumn& mc = ms->cols[filter->leaves[li].col];
            pd->leaf_zm += 2 * mc.value_bytes + 4;
            pd->leaf_row += mc.value_bytes;
        }
        pd->n_leaves = n_leaves;
    }
    if (scan) {
        rc = run_queries(c, &ms->st, q, (filter && !fuse) ? c->d_meta_mask : nullptr,
                         (filter && !fuse) ? (uint32_t)((ms->st.n + 31) / 32) : 0, d_records, map, c->d_stats,
                         fuse ? &c->cur_filter : nullptr, &pd->run);
        if (rc) return rc;
    } else if (c->ex_active) {
        // nothing to scan on this rank, but it still takes part in the exchange
        rc = exchange_only(c, pd->take_max, c->d_stats, &pd->run);
        if (rc) return rc;
    } else if (d_records) {
        // no scan: the record buffer must still hold k empty slots
        const uint64_t k_eff = std::min<uint64_t>(q->k, ms->st.n * (uint64_t)q->nq);
        if (k_eff) {
            rc = launch_cands_to_records(c->d_list[0], c->d_list_count, (uint32_t)k_eff, map, 1, d_records, s);
            if (rc) return rc;
        }
    }
    pd->active = true;
    return OTTERS_OK;
}

// `fetch`: copy the ordered result to the caller (else it stays on the device and only the statistics come back, and
// those only when `stats` is given — the one case that does not synchronise).
static int meta_finish(otters_ctx* c, Pending* pd, bool fetch, uint64_t* out_idx, float* out_score, uint32_t* out_qid, uint64_t cap,
                       uint64_t* out_len, otters_query_stats* stats) {
    otters_metastore* ms = pd->ms;
    cudaStream_t s = c->stream;
    unsigned long long hstats[4] = {0, 0, 0, 0};
    uint64_t n_out = 0;
    bool synced = false;
    auto fetch_stats_only = [&]() -> int {
        int r2 = ensure_pinned(&c->h_result, &c->h_result_bytes, 64);
        if (r2) return r2;
        OTTERS_CUDA(cudaMemcpyAsync(c->h_result, c->d_stats, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        OTTERS_CUDA(cudaStreamSynchronize(s));
        c->last.d2h_bytes += 2 * sizeof(unsigned long long);
        memcpy(hstats, c->h_result, 2 * sizeof(unsigned long long));
        synced = true;
        return OTTERS_OK;
    };
    const bool have_list = pd->scan || pd->run.k_eff > 0 || pd->run.zero_copy;  // a selection ran (scan or exchange-only)
    int rc;
    if (have_list && fetch) {
        rc = fetch_results(c, pd->run, pd->take_max, 0, out_idx, out_score, out_qid, cap, &n_out, hstats);
        if (rc) return rc;
        synced = true;
        if (pd->scan) finish_work_stats(c);
    } else if (!have_list || stats) {
        rc = fetch_stats_only();
        if (rc) return rc;
    }
    if (out_len) *out_len = n_out;
    // stats (src/meta.rs:711-721)
    otters_query_stats st{};
    st.total_chunks = ms->n_chunks;
    st.evaluated_chunks = hstats[0];
    st.pruned_chunks = st.total_chunks - st.evaluated_chunks;
    st.vectors_compared = hstats[1];
    const bool timed = c->timing && synced;
    float ms_f = 0.f;
    if (timed && c->timed_meta && elapsed_ms(c->ev[0], c->ev[1], &ms_f)) {
        st.prune_s = ms_f * 1e-3;
        c->last.prune_ms = ms_f;
    }
    if (timed && c->timed_rowmask && elapsed_ms(c->ev[1], c->ev[6], &ms_f)) c->last.rowmask_ms = ms_f;
    if (timed && pd->scan) {
        if (elapsed_ms(c->ev[2], c->ev[5], &ms_f)) st.score_s = ms_f * 1e-3 + c->last.rowmask_ms * 1e-3;
        if (c->timed_single && elapsed_ms(c->ev[4], c->ev[5], &ms_f)) {
            st.merge_s = ms_f * 1e-3;
            st.score_s -= st.merge_s;
            c->last.select_ms = ms_f;
        }
    }
    // algorithmic bytes of the metadata evaluation
    if (pd->has_filter) {
        const uint64_t rows_eval = pd->nq ? st.vectors_compared / pd->nq : 0;
        c->last.meta_bytes = ms->n_chunks * pd->leaf_zm +
                             (pd->scan ? rows_eval * pd->leaf_row + rows_eval / 8 * pd->n_leaves + ms->st.n / 8 : 0);
    }
    st.total_s = now_s() - pd->t_submit;
    if (synced) {
        ms->last = st;
        ms->has_stats = true;
    }
    if (stats) *stats = st;
    pd->active = false;
    return OTTERS_OK;
}

static int meta_query_impl(otters_ctx* c, otters_metastore* ms, const otters_vec_query* q, const otters_filter* filter,
                           otters_topk_record* d_records, ShardMap map, uint64_t* out_idx, float* out_score,
                           uint32_t* out_qid, uint64_t cap, uint64_t* out_len, otters_query_stats* stats) {
    if (out_len) *out_len = 0;
    const bool want_host = !d_records && out_len != nullptr;
    Pending pd;
    int rc = meta_enqueue(c, ms, q, filter, d_records, map, want_host, stats != nullptr, &pd);
    if (rc) return rc;
    return meta_finish(c, &pd, want_host, out_idx, out_score, out_qid, cap, out_len, stats);
}

}  // namespace otters

extern "C" int otters_metastore_query(otters_metastore* ms, const otters_vec_query* q, const otters_filter* filter,
                                      uint64_t* out_idx, float* out_score, uint32_t* out_qid, uint64_t cap, uint64_t* out_len,
                                      otters_query_stats* stats) {
    if (!ms || !out_len) return fail(OTTERS_ERR_INVALID, "null argument");
    DeviceGuard g(ms->ctx->device);
    return meta_query_impl(ms->ctx, ms, q, filter, nullptr, ShardMap{}, out_idx, out_score, out_qid, cap, out_len, stats);
}

extern "C" int otters_metastore_last_stats(const otters_metastore* ms, otters_query_stats* out) {
    if (!ms || !out) return fail(OTTERS_ERR_INVALID, "null argument");
    if (!ms->has_stats) return fail(OTTERS_ERR_INVALID, "(no query stats)");
    *out = ms->last;
    return OTTERS_OK;
}

namespace otters {
static int export_mask(otters_metastore* ms, const otters_filter* filter, bool rows, uint8_t* keep) {
    otters_ctx* c = ms->ctx;
    int rc = begin_query(c);
    if (rc) return rc;
    rc = run_meta_filter(c, ms, filter, 1, rows ? META_ROWMASK : META_PRUNE);
    if (rc) return rc;
    const uint64_t n = rows ? ms->st.n : ms->n_chunks;
    if (!filter) {
        OTTERS_CUDA(cudaStreamSynchronize(c->stream));
        memset(keep, 1, n);
        return OTTERS_OK;
    }
    std::vector<uint32_t> words((n + 31) / 32 + 1);
    OTTERS_CUDA(cudaMemcpyAsync(words.data(), rows ? c->d_meta_mask : c->d_chunk_keep, ((n + 31) / 32) * 4,
                                cudaMemcpyDeviceToHost, c->stream));
    OTTERS_CUDA(cudaStreamSynchronize(c->stream));
    for (uint64_t i = 0; i < n; ++i) keep[i] = (words[i >> 5] >> (i & 31)) & 1u;
    return OTTERS_OK;
}
}  // namespace otters

extern "C" int otters_metastore_chunk_mask(otters_metastore* ms, const otters_filter* filter, uint8_t* keep) {
    if (!ms || !keep) return fail(OTTERS_ERR_INVALID, "null argument");
    DeviceGuard g(ms->ctx->device);
    return export_mask(ms, filter, false, keep);
}
extern "C" int otters_metastore_row_mask(otters_metastore* ms, const otters_filter* filter, uint8_t* keep) {
    if (!ms || !keep) return fail(OTTERS_ERR_INVALID, "null argument");
    DeviceGuard g(ms->ctx->device);
    return export_mask(ms, filter, true, keep);
}

extern "C" int otters_metastore_zonemap_i64(const otters_metastore* ms, uint32_t col, int64_t* mn, int64_t* mx, uint64_t* non_null) {
    if (!ms || col >= ms->cols.size() || ms->cols[col].zmin_i.size() != ms->n_chunks || !mn || !mx || !non_null)
        return fail(OTTERS_ERR_INVALID, "column has no integer zonemap");
    for (uint64_t i = 0; i < ms->n_chunks; ++i) {
        mn[i] = ms->cols[col].zmin_i[i];
        mx[i] = ms->cols[col].zmax_i[i];
        non_null[i] = ms->cols[col].non_null[i];
    }
    return OTTERS_OK;
}
extern "C" int otters_metastore_zonemap_f64(const otters_metastore* ms, uint32_t col, double* mn, double* mx, uint64_t* non_null) {
    if (!ms || col >= ms->cols.size() || ms->cols[col].zmin_f.size() != ms->n_chunks || !mn || !mx || !non_null)
        return fail(OTTERS_ERR_INVALID, "column has no float zonemap");
    for (uint64_t i = 0; i < ms->n_chunks; ++i) {
        mn[i] = ms->cols[col].zmin_f[i];
        mx[i] = ms->cols[col].zmax_f[i];
        non_null[i] = ms->cols[col].non_null[i];
    }
    return OTTERS_OK;
}
extern "C" int otters_metastore_inv_norms(const otters_metastore* ms, uint64_t first, uint64_t n, float* out) {
    if (!ms || !out) return fail(OTTERS_ERR_INVALID, "null argument");
    if (first + n > ms->st.n) return fail(OTTERS_ERR_INVALID, "row range out of bounds");
    DeviceGuard g(ms->ctx->device);
    OTTERS_CUDA(cudaStreamSynchronize(ms->ctx->stream));
    if (n) OTTERS_CUDA(cudaMemcpy(out, ms->st.d_inv + first, n * sizeof(float), cudaMemcpyDeviceToHost));
    return OTTERS_OK;
}

// =================================================================================================
// row-sharded multi-GPU helpers, non-blocking queries
// =================================================================================================
namespace otters {

// VecQueryPlan::collect in two halves (see meta_enqueue / meta_finish)
static int vec_enqueue(otters_ctx* c, otters_vecstore* vs, const otters_vec_query* q, otters_topk_record* d_records, ShardMap map,
                       bool want_host, Pending* pd) {
    *pd = Pending{};
    pd->t_submit = now_s();
    int rc = validate_query(q, vs->st.dim, false);
    if (rc) return rc;
    pd->take_max = q->take_type == OTTERS_TAKE_MAX;
    pd->nq = q->nq;
    c->last = otters_last_work{};
    const uint64_t k_eff = std::min<uint64_t>(q->k, vs->st.n * (uint64_t)q->nq);
    if (k_eff == 0 && !c->ex_active) {  // take(0) / empty store (tests/vec_store_tests.rs:430-445,488-499)
        pd->active = true;
        return OTTERS_OK;
    }
    rc = begin_query(c);
    if (rc) return rc;
    c->want_host_result = want_host;
    if (vs->st.n == 0) {  // (only with an exchange attached: the rank still contributes k empty records)
        rc = exchange_only(c, pd->take_max, nullptr, &pd->run);
    } else {
        const uint32_t* d_mask = nullptr;
        uint32_t mask_words = 0;
        rc = upload_row_mask(c, q, vs->st.n, &d_mask, &mask_words);
        if (rc) return rc;
        rc = stage_queries(c, q, vs->st.pitch);
        if (rc) return rc;
        rc = run_queries(c, &vs->st, q, d_mask, mask_words, d_records, map, nullptr, nullptr, &pd->run);
        pd->scan = true;
    }
    if (rc) return rc;
    pd->active = true;
    return OTTERS_OK;
}

static int vec_finish(otters_ctx* c, Pending* pd, bool fetch, uint64_t* out_idx, float* out_score, uint32_t* out_qid, uint64_t cap,
                      uint64_t* out_len) {
    if (out_len) *out_len = 0;
    pd->active = false;
    const bool have_list = pd->scan || pd->run.k_eff > 0 || pd->run.zero_copy;
    if (!have_list || !fetch) return OTTERS_OK;
    uint64_t n = 0;
    int rc = fetch_results(c, pd->run, pd->take_max, 0, out_idx, out_score, out_qid, cap, &n, nullptr);
    if (rc) return rc;
    if (out_len) *out_len = n;
    if (pd->scan) finish_work_stats(c);
    return OTTERS_OK;
}

// no local scan on this rank (empty shard, take(0) locally impossible, swallowed per-chunk error): the rank still
// contributes k empty records and merges everybody else's
static int exchange_only(otters_ctx* c, int take_max, const unsigned long long* stats_src, QueryRun* run) {
    int rc0 = io_flush(c);
    if (rc0) return rc0;
    SelectParams se{};
    se.cta_keys = c->d_cta_keys;
    se.cta_counts = c->d_cta_counts;
    se.n_lists = 0;
    se.list_stride = 1;
    se.out = c->d_list[1];
    se.out_count = c->d_list_count + 1;
    se.tau_out = c->d_tau;
    se.k = 0;
    se.scratch_keys = c->d_scratch_keys;
    se.scratch_src = c->d_scratch_src;
    se.scratch_elems = c->scratch_elems;
    se.take_max = take_max;
    se.hdr = list_hdr(c, 1);
    se.rows_scored_src = c->d_rows_scored;
    se.stats_src = stats_src;
    fill_exchange(c, &se);
    if (c->want_host_result) {
        se.host_out = c->d_zc;
        run->zero_copy = true;
    }
    int rc = launch_select(se, c->stream);
    if (rc) return rc;
    c->ex_published = true;
    c->last.kernel_launches += 1;
    run->result_list = 1;
    run->k_eff = c->ex_k;
    run->big = false;
    return OTTERS_OK;
}

static int check_exchange(const otters_vec_query* q, const otters_peer_exchange* ex, uint64_t seq) {
    if (ex->world < 2 || ex->world > kMaxPeers || ex->rank >= ex->world || !ex->peer_records || !ex->peer_flags)
        return fail(OTTERS_ERR_INVALID, "peer exchange: world must be 2..8 with mapped record and flag areas");
    if (seq == 0) return fail(OTTERS_ERR_INVALID, "peer exchange: sequence numbers start at 1");
    if (q->k == 0 || q->k > kMaxFusedK || q->k > ex->k_max)
        return fail(OTTERS_ERR_UNSUPPORTED, "peer exchange serves take counts 1..min(1024, k_max)");
    return OTTERS_OK;
}

static void attach_exchange(otters_ctx* c, const otters_vec_query* q, const otters_peer_exchange* ex, uint64_t seq) {
    c->ex_active = true;
    c->ex_published = false;
    c->ex_world = ex->world;
    c->ex_rank = ex->rank;
    c->ex_kmax = (uint32_t)ex->k_max;
    c->ex_k = (uint32_t)q->k;
    c->ex_seq = (uint32_t)(seq % 0xFFFFFFFFull) + 1u;  // the flag value: never 0 (the areas start zeroed), unique within any
                                                      // window of kExchangeSlots consecutive queries
    c->ex_slot = (uint32_t)(seq % kExchangeSlots);
    for (uint32_t i = 0; i < ex->world; ++i) {
        c->ex_records[i] = (otters_topk_record*)ex->peer_records[i];
        c->ex_flags[i] = ex->peer_flags[i];
    }
}

// A rank that fails after the exchange was attached but before a kernel published its records would leave the peers
// spinning on its flag: publish k empty records instead (best effort; the caller still gets the original error).
static void exchange_bail_out(otters_ctx* c, bool take_max) {
    if (!c->ex_active || c->ex_published) return;
    const std::string err = g_last_error;
    QueryRun run;
    c->want_host_result = false;
    c->io_open = false;
    if (exchange_only(c, take_max, nullptr, &run) != OTTERS_OK) cudaGetLastError();
    g_last_error = err;
}

// the enqueue half of every query entry point: exactly one of vs / ms; ex nullable
static int enqueue_any(otters_ctx* c, otters_vecstore* vs, otters_metastore* ms, const otters_vec_query* q, const otters_filter* filter,
                       const otters_shard_map* map_in, const otters_peer_exchange* ex, uint64_t seq, bool want_host, bool want_stats,
                       Pending* pd) {
    if (ex) {
        int rc = check_exchange(q, ex, seq);
        if (rc) return rc;
        attach_exchange(c, q, ex, seq);
    }
    struct Reset {
        otters_ctx* c;
        ~Reset() { c->ex_active = false; }
    } reset{c};
    const ShardMap map = to_map(map_in);
    int rc = ms ? meta_enqueue(c, ms, q, filter, nullptr, map, want_host, want_stats, pd)
                : vec_enqueue(c, vs, q, nullptr, map, want_host, pd);
    if (rc) exchange_bail_out(c, q->take_type == OTTERS_TAKE_MAX);
    return rc;
}

static int ctx_create_impl(int device, void* cuda_stream, otters_ctx** out);

static otters_ctx* get_lane(otters_ctx* parent, uint32_t i) {
    if (i == 0) return parent;
    if (!parent->lane[i]) {
        otters_ctx* child = nullptr;
        if (ctx_create_impl(parent->device, nullptr, &child) != OTTERS_OK) return nullptr;
        child->parent = parent;
        child->tuning = parent->tuning;
        parent->lane[i] = child;
    }
    return parent->lane[i];
}

}  // namespace otters

extern "C" int otters_query_local_device(otters_vecstore* vs, otters_metastore* ms, const otters_vec_query* q,
                                         const otters_filter* filter, const otters_shard_map* map_in, void* d_records,
                                         otters_query_stats* stats) {
    if ((!vs && !ms) || (vs && ms) || !d_records) return fail(OTTERS_ERR_INVALID, "pass exactly one store and a record buffer");
    const ShardMap map = to_map(map_in);
    if (ms) {
        DeviceGuard g(ms->ctx->device);
        return meta_query_impl(ms->ctx, ms, q, filter, (otters_topk_record*)d_records, map, nullptr, nullptr, nullptr, 0, nullptr, stats);
    }
    if (filter) return fail(OTTERS_ERR_INVALID, "meta_filter needs a MetaStore");
    otters_ctx* c = vs->st.ctx;
    DeviceGuard g(c->device);
    Pending pd;
    return vec_enqueue(c, vs, q, (otters_topk_record*)d_records, map, false, &pd);
}

extern "C" int otters_query_exchange(otters_vecstore* vs, otters_metastore* ms, const otters_vec_query* q, const otters_filter* filter,
                                     const otters_shard_map* map_in, const otters_peer_exchange* ex, uint64_t seq, uint64_t* out_idx,
                                     float* out_score, uint32_t* out_qid, uint64_t cap, uint64_t* out_len, otters_query_stats* stats) {
    if ((!vs && !ms) || (vs && ms) || !q || !ex) return fail(OTTERS_ERR_INVALID, "pass exactly one store, a query and an exchange");
    if (vs && filter) return fail(OTTERS_ERR_INVALID, "meta_filter needs a MetaStore");
    otters_ctx* c = vs ? vs->st.ctx : ms->ctx;
    DeviceGuard g(c->device);
    const bool want_fetch = out_idx || out_score || out_qid || cap;
    if (want_fetch && !out_len) return fail(OTTERS_ERR_INVALID, "null out_len");
    if (out_len) *out_len = 0;
    Pending pd;
    int rc = enqueue_any(c, vs, ms, q, filter, map_in, ex, seq, want_fetch, stats != nullptr, &pd);
    if (rc) return rc;
    if (ms) return meta_finish(c, &pd, want_fetch, out_idx, out_score, out_qid, cap, out_len, stats);
    return vec_finish(c, &pd, want_fetch, out_idx, out_score, out_qid, cap, out_len);
}

extern "C" int otters_query_submit(otters_vecstore* vs, otters_metastore* ms, const otters_vec_query* q, const otters_filter* filter,
                                   const otters_shard_map* map_in, const otters_peer_exchange* ex, uint64_t seq, uint64_t* ticket) {
    if ((!vs && !ms) || (vs && ms) || !q || !ticket) return fail(OTTERS_ERR_INVALID, "pass exactly one store, a query and a ticket");
    if (vs && filter) return fail(OTTERS_ERR_INVALID, "meta_filter needs a MetaStore");
    otters_ctx* parent = vs ? vs->st.ctx : ms->ctx;
    DeviceGuard g(parent->device);
    const uint32_t li = parent->next_lane % kMaxLanes;
    otters_ctx* c = get_lane(parent, li);
    if (!c) return OTTERS_ERR_CUDA;
    // a ticket of this lane that was never waited for is abandoned here: its result buffers are reused
    c->pend.active = false;
    Pending pd;
    c->pipelined = true;
    int rc = enqueue_any(c, vs, ms, q, filter, map_in, ex, seq, true, false, &pd);
    c->pipelined = false;
    if (rc) return rc;
    parent->next_lane += 1;
    parent->tickets += 1;
    pd.ticket = (parent->tickets << 8) | li;
    c->pend = pd;
    *ticket = pd.ticket;
    return OTTERS_OK;
}

namespace otters {
// ---- per-query top-k for a batch (extension beyond src/vec.rs:217-219's merged list) ------------------------------------
// Every query of the batch gets its own list — exactly what the single-query call returns for it (bit-identical by
// construction: each list comes from one streaming scan in the reference's arithmetic).  The queries are pipelined over the
// context's two lanes, so the launch, the input copy and the selection of query i+1 overlap the scan of query i.
static int query_batch_impl(otters_vecstore* vs, otters_metastore* ms, const otters_vec_query* q, const otters_filter* filter,
                            uint64_t* out_idx, float* out_score, uint64_t* out_len, otters_query_stats* stats) {
    otters_ctx* parent = vs ? vs->st.ctx : ms->ctx;
    DeviceGuard g(parent->device);
    if (!q) return fail(OTTERS_ERR_INVALID, "Query vectors or their norms are not set");
    if (!out_len || ((!out_idx || !out_score) && q->nq && q->k)) return fail(OTTERS_ERR_INVALID, "null output buffers");
    const uint32_t store_dim = vs ? vs->st.dim : ms->st.dim;
    if (vs) {
        int rc = validate_query(q, store_dim, false);
        if (rc) return rc;
    } else if (q->nq == 0 || q->dim != store_dim || !q->queries) {
        // MetaStore swallows per-chunk errors (src/meta_compute.rs:182): no rows, statistics only — same as the merged call
        uint64_t n = 0;
        return meta_query_impl(parent, ms, q, filter, nullptr, ShardMap{}, nullptr, nullptr, nullptr, 0, &n, stats);
    }
    Pending pd[kMaxLanes];
    otters_query_stats st_first{};
    otters_last_work total{};
    int rc = OTTERS_OK;
    for (uint32_t i = 0; i <= q->nq && rc == OTTERS_OK; ++i) {
        if (i < q->nq) {
            otters_ctx* c = get_lane(parent, i % kMaxLanes);
            if (!c) return OTTERS_ERR_CUDA;
            c->pend.active = false;
            otters_vec_query one = *q;
            one.queries = q->queries + (size_t)i * q->dim;
            one.nq = 1;
            c->pipelined = true;
            rc = enqueue_any(c, vs, ms, &one, filter, nullptr, nullptr, 0, true, false, &pd[i % kMaxLanes]);
            c->pipelined = false;
            if (rc) break;
        }
        if (i >= 1) {
            const uint32_t j = i - 1;
            otters_ctx* c = get_lane(parent, j % kMaxLanes);
            uint64_t n = 0;
            otters_query_stats st{};
            rc = ms ? meta_finish(c, &pd[j % kMaxLanes], true, out_idx + (size_t)j * q->k, out_score + (size_t)j * q->k, nullptr, q->k, &n, &st)
                    : vec_finish(c, &pd[j % kMaxLanes], true, out_idx + (size_t)j * q->k, out_score + (size_t)j * q->k, nullptr, q->k, &n);
            out_len[j] = std::min<uint64_t>(n, q->k);
            if (j == 0) st_first = st;
            total.kernel_launches += c->last.kernel_launches;
            total.rows_scored += c->last.rows_scored;
            total.scan_bytes += c->last.scan_bytes;
            total.h2d_bytes += c->last.h2d_bytes;
            total.d2h_bytes += c->last.d2h_bytes;
        }
    }
    if (rc) {
        otters_ctx_synchronize(parent);
        return rc;
    }
    parent->last = total;
    parent->report_waited = false;
    if (ms) {
        // statistics of the batch as the reference reports them: chunks once, vectors_compared = sum over chunks of len * Q
        st_first.vectors_compared *= q->nq;
        ms->last = st_first;
        ms->has_stats = true;
        if (stats) *stats = st_first;
    }
    return OTTERS_OK;
}

}  // namespace otters

extern "C" int otters_vecstore_query_batch(otters_vecstore* vs, const otters_vec_query* q, uint64_t* out_idx, float* out_score,
                                           uint64_t* out_len) {
    if (!vs) return fail(OTTERS_ERR_INVALID, "null store");
    return query_batch_impl(vs, nullptr, q, nullptr, out_idx, out_score, out_len, nullptr);
}

extern "C" int otters_metastore_query_batch(otters_metastore* ms, const otters_vec_query* q, const otters_filter* filter,
                                            uint64_t* out_idx, float* out_score, uint64_t* out_len, otters_query_stats* stats) {
    if (!ms) return fail(OTTERS_ERR_INVALID, "null store");
    if (q && q->row_mask_words) return fail(OTTERS_ERR_INVALID, "row masks are not part of MetaQueryPlan");
    return query_batch_impl(nullptr, ms, q, filter, out_idx, out_score, out_len, stats);
}

// MetaQueryResults.data (src/meta.rs:723-821): the metadata of the result rows, gathered on the device
extern "C" int otters_metastore_gather(otters_metastore* ms, uint32_t col, const uint64_t* rows, uint64_t n, void* out_values,
                                       uint8_t* out_nulls) {
    if (!ms || col >= ms->cols.size()) return fail(OTTERS_ERR_INVALID, "unknown column");
    if (n == 0) return OTTERS_OK;
    if (!rows || !out_values || !out_nulls) return fail(OTTERS_ERR_INVALID, "null argument");
    otters_ctx* c = ms->ctx;
    DeviceGuard g(c->device);
    const MetaColumn& mc = ms->cols[col];
    const size_t w = mc.value_bytes;
    for (uint64_t i = 0; i < n; ++i)
        if (rows[i] >= ms->st.n) return fail(OTTERS_ERR_INVALID, "row index out of bounds");
    // staging: [rows u32 | values | nulls] in one pinned buffer and one device buffer
    const size_t off_v = round_up(n * 4, 16), off_n = off_v + round_up(n * w, 16), total = off_n + round_up(n, 16);
    int rc = ensure_stage(c, total);
    if (rc) return rc;
    rc = ensure_dev(&c->d_gather, &c->d_gather_bytes, total, c->stream);
    if (rc) return rc;
    uint32_t* h_rows = reinterpret_cast<uint32_t*>(c->h_stage);
    for (uint64_t i = 0; i < n; ++i) h_rows[i] = (uint32_t)rows[i];
    cudaStream_t s = c->stream;
    OTTERS_CUDA(cudaMemcpyAsync(c->d_gather, h_rows, n * 4, cudaMemcpyHostToDevice, s));
    rc = launch_gather(mc.d_values, mc.d_nulls, (uint32_t)w, reinterpret_cast<const uint32_t*>(c->d_gather), (uint32_t)n, c->d_gather + off_v,
                       c->d_gather + off_n, s);
    if (rc) return rc;
    OTTERS_CUDA(cudaMemcpyAsync(c->h_stage + off_v, c->d_gather + off_v, total - off_v, cudaMemcpyDeviceToHost, s));
    OTTERS_CUDA(cudaStreamSynchronize(s));
    memcpy(out_values, c->h_stage + off_v, n * w);
    memcpy(out_nulls, c->h_stage + off_n, n);
    return OTTERS_OK;
}

extern "C" int otters_metastore_dict_entry(const otters_metastore* ms, uint32_t col, uint32_t code, const uint8_t** bytes, uint64_t* len) {
    if (!ms || col >= ms->cols.size() || !bytes || !len) return fail(OTTERS_ERR_INVALID, "unknown column");
    const MetaColumn& mc = ms->cols[col];
    if (mc.dtype != OTTERS_DTYPE_STRING || code >= mc.dict_strings.size()) return fail(OTTERS_ERR_INVALID, "not a dictionary code of this column");
    *bytes = reinterpret_cast<const uint8_t*>(mc.dict_strings[code].data());
    *len = mc.dict_strings[code].size();
    return OTTERS_OK;
}

// =================================================================================================
// Persistence — the reference's roadmap item "Persistence (save/load MetaStore to/from disk)" (README.md:206;
// SURVEY.md §8f rank 3).  A built MetaStore is immutable, so the file is simply its HBM image: the rows (fp32 or bf16 as
// stored), the inverse norms, every column's values / null words / zonemap tables / Bloom filters / dictionary, plus an
// opaque caller blob (the host mirror keeps its row-order permutation there).  Loading allocates and copies — nothing is
// recomputed, so a loaded store answers with the same bytes, statistics included.
// =================================================================================================
namespace otters {
namespace {

constexpr char kFileMagic[8] = {'O', 'T', 'T', 'E', 'R', 'S', 'B', '2'};
constexpr uint32_t kFileVersion = 1;
constexpr size_t kIoSlab = (size_t)64 << 20;

struct FileHeader {
    char magic[8];
    uint32_t version, half;
    uint64_t n_rows;
    uint32_t dim, pitch;
    uint64_t chunk_size, n_chunks;
    uint32_t n_cols, reserved;
    uint64_t user_bytes;
    uint64_t total_bytes;  // of the whole file: a truncated or padded file is rejected before anything is allocated
};

struct FileIo {
    FILE* f = nullptr;
    std::vector<uint8_t> slab;
    uint64_t pos = 0;
    ~FileIo() {
        if (f) fclose(f);
    }
    int put(const void* p, size_t n) {
        if (n && fwrite(p, 1, n, f) != n) return fail(OTTERS_ERR_INVALID, std::string("write failed: ") + strerror(errno));
        pos += n;
        return OTTERS_OK;
    }
    int get(void* p, size_t n) {
        if (n && fread(p, 1, n, f) != n) return fail(OTTERS_ERR_INVALID, "store file is truncated");
        pos += n;
        return OTTERS_OK;
    }
    template <typename T>
    int put_pod(const T& v) { return put(&v, sizeof(T)); }
    template <typename T>
    int get_pod(T* v) { return get(v, sizeof(T)); }
    int put_dev(const void* d, size_t n) {  // device -> file through a bounce slab
        if (n && !d) return fail(OTTERS_ERR_INVALID, "store is missing a device array");
        if (slab.empty()) slab.resize(kIoSlab);
        for (size_t done = 0; done < n; done += kIoSlab) {
            const size_t m = std::min(kIoSlab, n - done);
            OTTERS_CUDA(cudaMemcpy(slab.data(), (const uint8_t*)d + done, m, cudaMemcpyDeviceToHost));
            int rc = put(slab.data(), m);
            if (rc) return rc;
        }
        return OTTERS_OK;
    }
    int get_dev(void* d, size_t n) {
        if (slab.empty()) slab.resize(kIoSlab);
        for (size_t done = 0; done < n; done += kIoSlab) {
            const size_t m = std::min(kIoSlab, n - done);
            int rc = get(slab.data(), m);
            if (rc) return rc;
            OTTERS_CUDA(cudaMemcpy((uint8_t*)d + done, slab.data(), m, cudaMemcpyHostToDevice));
        }
        return OTTERS_OK;
    }
    template <typename T>
    int put_vec(const std::vector<T>& v) {
        const uint64_t n = v.size();
        int rc = put_pod(n);
        return rc ? rc : put(v.data(), n * sizeof(T));
    }
    template <typename T>
    int get_vec(std::vector<T>* v, uint64_t max_elems) {
        uint64_t n = 0;
        int rc = get_pod(&n);
        if (rc) return rc;
        if (n > max_elems) return fail(OTTERS_ERR_INVALID, "store file is corrupt (table larger than the store)");
        v->resize(n);
        return get(v->data(), n * sizeof(T));
    }
};

struct ColHeader {
    int32_t dtype;
    uint32_t has_nulls, has_zonemap, has_bloom;
    uint64_t value_bytes, bloom_stride;
    uint32_t bloom_k0, name_len;
};

template <typename T>
int dev_new(T** dptr, size_t bytes) {
    *dptr = nullptr;
    if (cudaMalloc((void**)dptr, std::max<size_t>(bytes, 16)) != cudaSuccess) {
        cudaGetLastError();
        return fail(OTTERS_ERR_NOMEM, "device allocation for the loaded store failed");
    }
    return OTTERS_OK;
}

}  // namespace
}  // namespace otters

extern "C" int otters_metastore_save(otters_metastore* ms, const char* path, const void* user, uint64_t user_bytes) {
    if (!ms || !path) return fail(OTTERS_ERR_INVALID, "null argument");
    if (user_bytes && !user) return fail(OTTERS_ERR_INVALID, "null user blob");
    DeviceGuard g(ms->ctx->device);
    if (otters_ctx_synchronize(ms->ctx) != OTTERS_OK) return OTTERS_ERR_CUDA;
    FileIo io;
    io.f = fopen(path, "wb");
    if (!io.f) return fail(OTTERS_ERR_INVALID, std::string("cannot open '") + path + "' for writing: " + strerror(errno));
    const VecStorage& st = ms->st;
    const uint64_t n = st.n, nc = ms->n_chunks;
    FileHeader h{};
    memcpy(h.magic, kFileMagic, 8);
    h.version = kFileVersion;
    h.half = st.half ? 1u : 0u;
    h.n_rows = n;
    h.dim = st.dim;
    h.pitch = st.pitch;
    h.chunk_size = ms->chunk_size;
    h.n_chunks = nc;
    h.n_cols = (uint32_t)ms->cols.size();
    h.user_bytes = user_bytes;
    int rc = io.put_pod(h);  // total_bytes is patched at the end
    if (!rc) rc = io.put(user, user_bytes);
    if (!rc) rc = io.put_dev(st.d_rows, (size_t)n * st.pitch * st.esz());
    if (!rc) rc = io.put_dev(st.d_inv, (size_t)n * 4);
    for (size_t i = 0; i < ms->cols.size() && !rc; ++i) {
        const MetaColumn& mc = ms->cols[i];
        ColHeader ch{};
        ch.dtype = mc.dtype;
        ch.has_nulls = mc.d_nulls ? 1u : 0u;
        ch.has_zonemap = (mc.d_zmin && mc.d_zmax) ? 1u : 0u;
        ch.has_bloom = mc.d_bloom ? 1u : 0u;
        ch.value_bytes = mc.value_bytes;
        ch.bloom_stride = mc.bloom_stride;
        ch.bloom_k0 = mc.bloom_k0;
        ch.name_len = (uint32_t)mc.name.size();
        rc = io.put_pod(ch);
        if (!rc) rc = io.put(mc.name.data(), mc.name.size());
        if (!rc) rc = io.put_dev(mc.d_values, (size_t)n * mc.value_bytes);
        if (!rc && ch.has_nulls) rc = io.put_dev(mc.d_nulls, (size_t)((n + 63) / 64) * 8);
        if (!rc && ch.has_zonemap) rc = io.put_dev(mc.d_zmin, (size_t)nc * mc.value_bytes);
        if (!rc && ch.has_zonemap) rc = io.put_dev(mc.d_zmax, (size_t)nc * mc.value_bytes);
        if (!rc) rc = io.put_dev(mc.d_non_null, (size_t)nc * 4);
        if (!rc && ch.has_bloom) rc = io.put_dev(mc.d_bloom, (size_t)nc * mc.bloom_stride * 8);
        if (!rc && ch.has_bloom) rc = io.put_dev(mc.d_bloom_mbits, (size_t)nc * 8);
        if (!rc && ch.has_bloom) rc = io.put_dev(mc.d_bloom_k, (size_t)nc * 4);
        if (!rc) rc = io.put_vec(mc.zmin_i);
        if (!rc) rc = io.put_vec(mc.zmax_i);
        if (!rc) rc = io.put_vec(mc.zmin_f);
        if (!rc) rc = io.put_vec(mc.zmax_f);
        if (!rc) rc = io.put_vec(mc.non_null);
        const uint64_t nd = mc.dict_strings.size();
        if (!rc) rc = io.put_pod(nd);
        for (uint64_t d = 0; d < nd && !rc; ++d) {
            const uint64_t len = mc.dict_strings[d].size();
            rc = io.put_pod(len);
            if (!rc) rc = io.put(mc.dict_strings[d].data(), len);
        }
    }
    if (!rc) {
        h.total_bytes = io.pos;
        if (fseek(io.f, 0, SEEK_SET) != 0 || fwrite(&h, 1, sizeof(h), io.f) != sizeof(h)) rc = fail(OTTERS_ERR_INVALID, "write failed");
    }
    if (fclose(io.f) != 0 && !rc) rc = fail(OTTERS_ERR_INVALID, std::string("write failed: ") + strerror(errno));
    io.f = nullptr;
    if (rc) remove(path);
    return rc;
}

extern "C" int otters_metastore_load(otters_ctx* c, const char* path, otters_metastore** out) {
    if (!c || !path || !out) return fail(OTTERS_ERR_INVALID, "null argument");
    *out = nullptr;
    DeviceGuard g(c->device);
    FileIo io;
    io.f = fopen(path, "rb");
    if (!io.f) return fail(OTTERS_ERR_INVALID, std::string("cannot open '") + path + "': " + strerror(errno));
    FileHeader h{};
    int rc = io.get_pod(&h);
    if (rc) return rc;
    if (memcmp(h.magic, kFileMagic, 8) != 0) return fail(OTTERS_ERR_INVALID, "not an otters_b200 store file");
    if (h.version != kFileVersion) return fail(OTTERS_ERR_UNSUPPORTED, "store file version " + std::to_string(h.version) + " is not supported");
    if (fseek(io.f, 0, SEEK_END) != 0) return fail(OTTERS_ERR_INVALID, "cannot seek in the store file");
    const uint64_t actual = (uint64_t)ftell(io.f);
    if (actual != h.total_bytes) return fail(OTTERS_ERR_INVALID, "store file is truncated or corrupt (size does not match its header)");
    fseek(io.f, (long)sizeof(h), SEEK_SET);
    const uint64_t want_pitch = round_up(std::max<uint32_t>(h.dim, 1), h.half ? 8 : 4);
    const uint64_t cs = std::max<uint64_t>(h.chunk_size, 1);
    if (h.half > 1 || h.pitch != want_pitch || h.n_chunks != (h.n_rows + cs - 1) / cs || h.n_rows >= 0xFFFFFFF0ull || h.n_cols > 65536 ||
        h.user_bytes > actual)
        return fail(OTTERS_ERR_INVALID, "store file is corrupt (inconsistent header)");

    std::unique_ptr<otters_metastore> ms(new otters_metastore());
    auto cleanup = [&](int r) {
        otters_metastore_destroy(ms.release());
        return r;
    };
    ms->ctx = c;
    ms->chunk_size = cs;
    ms->n_chunks = h.n_chunks;
    ms->st.ctx = c;
    ms->st.set_format(h.dim, h.half != 0);
    ms->user_blob.resize(h.user_bytes);
    if ((rc = io.get(ms->user_blob.data(), h.user_bytes))) return cleanup(rc);
    const uint64_t n = h.n_rows, nc = h.n_chunks;
    if ((rc = ms->st.reserve(n))) return cleanup(rc);
    if ((rc = io.get_dev(ms->st.d_rows, (size_t)n * ms->st.pitch * ms->st.esz()))) return cleanup(rc);
    if ((rc = io.get_dev(ms->st.d_inv, (size_t)n * 4))) return cleanup(rc);
    ms->st.n = n;
    ms->cols.resize(h.n_cols);
    std::vector<DevColumn> dcols(h.n_cols);
    for (uint32_t i = 0; i < h.n_cols; ++i) {
        MetaColumn& mc = ms->cols[i];
        ColHeader ch{};
        if ((rc = io.get_pod(&ch))) return cleanup(rc);
        const bool is_str = ch.dtype == OTTERS_DTYPE_STRING;
        if (ch.dtype < 0 || ch.dtype > OTTERS_DTYPE_DATETIME || (ch.value_bytes != 4 && ch.value_bytes != 8) || ch.name_len > 4096 ||
            (ch.has_bloom && (ch.bloom_stride == 0 || ch.bloom_stride > ((uint64_t)1 << 24))) || (is_str && ch.value_bytes != 4))
            return cleanup(fail(OTTERS_ERR_INVALID, "store file is corrupt (column header)"));
        mc.dtype = ch.dtype;
        mc.value_bytes = ch.value_bytes;
        mc.bloom_stride = ch.bloom_stride;
        mc.bloom_k0 = ch.bloom_k0;
        mc.name.resize(ch.name_len);
        if ((rc = io.get(&mc.name[0], ch.name_len))) return cleanup(rc);
        if ((rc = dev_new((uint8_t**)&mc.d_values, (size_t)n * mc.value_bytes))) return cleanup(rc);
        if ((rc = io.get_dev(mc.d_values, (size_t)n * mc.value_bytes))) return cleanup(rc);
        if (ch.has_nulls) {
            if ((rc = dev_new(&mc.d_nulls, (size_t)((n + 63) / 64) * 8))) return cleanup(rc);
            if ((rc = io.get_dev(mc.d_nulls, (size_t)((n + 63) / 64) * 8))) return cleanup(rc);
        }
        if (ch.has_zonemap) {
            if ((rc = dev_new((uint8_t**)&mc.d_zmin, (size_t)nc * mc.value_bytes))) return cleanup(rc);
            if ((rc = io.get_dev(mc.d_zmin, (size_t)nc * mc.value_bytes))) return cleanup(rc);
            if ((rc = dev_new((uint8_t**)&mc.d_zmax, (size_t)nc * mc.value_bytes))) return cleanup(rc);
            if ((rc = io.get_dev(mc.d_zmax, (size_t)nc * mc.value_bytes))) return cleanup(rc);
        }
        if ((rc = dev_new(&mc.d_non_null, (size_t)nc * 4))) return cleanup(rc);
        if ((rc = io.get_dev(mc.d_non_null, (size_t)nc * 4))) return cleanup(rc);
        if (ch.has_bloom) {
            if ((rc = dev_new(&mc.d_bloom, (size_t)nc * mc.bloom_stride * 8))) return cleanup(rc);
            if ((rc = io.get_dev(mc.d_bloom, (size_t)nc * mc.bloom_stride * 8))) return cleanup(rc);
            if ((rc = dev_new(&mc.d_bloom_mbits, (size_t)nc * 8))) return cleanup(rc);
            if ((rc = io.get_dev(mc.d_bloom_mbits, (size_t)nc * 8))) return cleanup(rc);
            if ((rc = dev_new(&mc.d_bloom_k, (size_t)nc * 4))) return cleanup(rc);
            if ((rc = io.get_dev(mc.d_bloom_k, (size_t)nc * 4))) return cleanup(rc);
        }
        if ((rc = io.get_vec(&mc.zmin_i, nc))) return cleanup(rc);
        if ((rc = io.get_vec(&mc.zmax_i, nc))) return cleanup(rc);
        if ((rc = io.get_vec(&mc.zmin_f, nc))) return cleanup(rc);
        if ((rc = io.get_vec(&mc.zmax_f, nc))) return cleanup(rc);
        if ((rc = io.get_vec(&mc.non_null, nc))) return cleanup(rc);
        uint64_t nd = 0;
        if ((rc = io.get_pod(&nd))) return cleanup(rc);
        if (nd > n + 1) return cleanup(fail(OTTERS_ERR_INVALID, "store file is corrupt (dictionary larger than the store)"));
        mc.dict_strings.resize(nd);
        for (uint64_t d = 0; d < nd; ++d) {
            uint64_t len = 0;
            if ((rc = io.get_pod(&len))) return cleanup(rc);
            if (len > actual) return cleanup(fail(OTTERS_ERR_INVALID, "store file is corrupt (dictionary entry)"));
            mc.dict_strings[d].resize(len);
            if ((rc = io.get(&mc.dict_strings[d][0], len))) return cleanup(rc);
            mc.dict.emplace(mc.dict_strings[d], (uint32_t)d);
        }
        DevColumn& d = dcols[i];
        d.dtype = mc.dtype;
        d.values = mc.d_values;
        d.null_words = mc.d_nulls;
        d.zmin = mc.d_zmin;
        d.zmax = mc.d_zmax;
        d.non_null = mc.d_non_null;
        d.bloom = mc.d_bloom;
        d.bloom_stride = mc.bloom_stride;
        d.bloom_mbits = mc.d_bloom_mbits;
        d.bloom_k = mc.d_bloom_k;
    }
    if (io.pos != actual) return cleanup(fail(OTTERS_ERR_INVALID, "store file is corrupt (trailing bytes)"));
    if ((rc = upload(&ms->d_cols, dcols.data(), dcols.size() * sizeof(DevColumn)))) return cleanup(rc);
    *out = ms.release();
    return OTTERS_OK;
}

extern "C" int otters_metastore_user_blob(const otters_metastore* ms, const void** bytes, uint64_t* len) {
    if (!ms || !bytes || !len) return fail(OTTERS_ERR_INVALID, "null argument");
    *bytes = ms->user_blob.data();
    *len = ms->user_blob.size();
    return OTTERS_OK;
}

extern "C" uint32_t otters_metastore_n_columns(const otters_metastore* ms) { return ms ? (uint32_t)ms->cols.size() : 0; }
extern "C" uint32_t otters_metastore_dim(const otters_metastore* ms) { return ms ? ms->st.dim : 0; }
extern "C" int32_t otters_metastore_format(const otters_metastore* ms) {
    return ms && ms->st.half ? OTTERS_VECTORS_FMT_BF16 : OTTERS_VECTORS_FMT_F32;
}
extern "C" int otters_metastore_column_info(const otters_metastore* ms, uint32_t col, const char** name, int32_t* dtype) {
    if (!ms || col >= ms->cols.size() || !name || !dtype) return fail(OTTERS_ERR_INVALID, "unknown column");
    *name = ms->cols[col].name.c_str();
    *dtype = ms->cols[col].dtype;
    return OTTERS_OK;
}

extern "C" int otters_query_wait(otters_ctx* parent, uint64_t ticket, uint64_t* out_idx, float* out_score, uint32_t* out_qid,
                                 uint64_t cap, uint64_t* out_len, otters_query_stats* stats) {
    if (!parent || !out_len) return fail(OTTERS_ERR_INVALID, "null argument");
    *out_len = 0;
    const uint32_t li = (uint32_t)(ticket & 0xFFu);
    otters_ctx* c = li < kMaxLanes ? (li == 0 ? parent : parent->lane[li]) : nullptr;
    if (!c || !c->pend.active || c->pend.ticket != ticket)
        return fail(OTTERS_ERR_INVALID, "unknown or expired ticket (a lane holds one query: wait for a ticket before submitting two more)");
    DeviceGuard g(c->device);
    int rc = c->pend.meta ? meta_finish(c, &c->pend, true, out_idx, out_score, out_qid, cap, out_len, stats)
                          : vec_finish(c, &c->pend, true, out_idx, out_score, out_qid, cap, out_len);
    parent->reported = c->last;
    parent->report_waited = true;
    return rc;
}

extern "C" int otters_topk_merge_device(otters_ctx* c, const void* d_records, uint64_t n_records, uint64_t k, int32_t take_type,
                                        uint64_t* out_idx, float* out_score, uint32_t* out_qid, uint64_t cap, uint64_t* out_len) {
    if (!c || !out_len) return fail(OTTERS_ERR_INVALID, "null argument");
    DeviceGuard g(c->device);
    *out_len = 0;
    if (n_records == 0 || k == 0) return OTTERS_OK;
    if (!d_records) return fail(OTTERS_ERR_INVALID, "null records");
    if (n_records > c->scratch_elems) return fail(OTTERS_ERR_UNSUPPORTED, "merge: too many records for the device merge");
    const uint64_t k_eff = std::min<uint64_t>(k, n_records);
    cudaStream_t s = c->stream;
    int rc = ensure_list0(c, k_eff);
    if (rc) return rc;
    rc = launch_merge_records((const otters_topk_record*)d_records, (uint32_t)n_records, (uint32_t)k_eff, take_type == OTTERS_TAKE_MAX,
                              c->d_list[0], c->d_list_count, list_hdr(c, 0), c->d_rows_scored, c->d_scratch_keys, c->d_scratch_src,
                              c->scratch_elems, s);
    if (rc) return rc;
    c->last.kernel_launches += 1;
    QueryRun run;
    run.result_list = 0;
    run.k_eff = k_eff;
    if (!out_idx && !out_score && !out_qid && cap == 0) {
        // asynchronous form: the merged list stays on the device (used to time the device-only pipeline)
        *out_len = k_eff;
        return OTTERS_OK;
    }
    rc = fetch_results(c, run, take_type == OTTERS_TAKE_MAX, 0, out_idx, out_score, out_qid, cap, out_len, nullptr);
    if (rc) return rc;
    finish_work_stats(c);  // timing of this rank's local scan (the events completed with the sync above)
    return OTTERS_OK;
}
