// internal.h — host-side structures and kernel launch prototypes of libotters_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/otters_b200.h"
#include "common.cuh"

namespace otters {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
#define OTTERS_CUDA(expr)                                                                                   \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            return ::otters::fail(OTTERS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
    } while (0)

// local row -> global row id (otters_shard_map)
struct ShardMap {
    uint64_t row_base = 0;
    uint32_t world = 1;
    uint32_t rank = 0;
    uint64_t block = 0;
    __host__ __device__ __forceinline__ uint64_t global_row(uint64_t local) const {
        if (world <= 1 || block == 0) return row_base + local;
        return row_base + ((local / block) * world + rank) * block + local % block;
    }
};

constexpr uint32_t kTileRows = 16;     // rows per staged tile (2 threads per row, one warp per tile)
constexpr uint32_t kMaxUnitRows = 128; // rows per dynamically scheduled work unit
constexpr uint32_t kMaxFusedK = 1024;  // largest k served by the fused per-CTA top-k buffers
constexpr uint32_t kSelectSmemElems = 8192;
constexpr uint32_t kSelectSmemBytes = kSelectSmemElems * 12;  // working set of K3: 64-bit keys + 32-bit source tags
constexpr uint32_t kPlannerRingTiles = 64;   // 16-row tiles in flight between the planner and the worker warps of K1
constexpr uint32_t kPlannerRingBytes = 16 + 3 * 4 * 64 + 64 * 16 * 4 + 112;  // sizeof(Ring) rounded up to 128
constexpr uint32_t kMaxPeers = 8;       // GPUs of one NVSwitch domain taking part in a row-sharded search
constexpr uint32_t kExchangeSlots = 4;  // depth of the peer-exchange record / flag areas: twice the queries a context keeps in flight
constexpr uint32_t kMaxLanes = 2;       // queries a context keeps in flight (otters_query_submit / otters_query_wait)
constexpr uint32_t kBatchRows = 128;    // store rows per tile of the batched kernel (UMMA M)
constexpr uint32_t kBatchQueries = 256; // queries per tile of the batched kernel (UMMA N)

struct DevLeaf;

// ---- selection (K3; stand-alone kernel or fused into the last CTA of a scan kernel) --------------
// Every result list in device memory is preceded by this 64-byte header, so that one D2H copy returns
// the count, the instrumentation counters and the ordered candidates together.
struct ResultHeader {
    uint32_t count;
    uint32_t pad;
    unsigned long long rows_scored;
    unsigned long long stats[4];  // [0] evaluated chunks, [1] vectors_compared
    unsigned long long extra[2];  // batched path: [0] flags | max |approx - exact| bits << 32, [1] goodness of the best excluded pair
};
static_assert(sizeof(ResultHeader) == 64, "ResultHeader must be 64 bytes");

struct SelectParams {
    const uint64_t* cta_keys;
    const uint32_t* cta_counts;
    uint32_t n_lists;
    uint32_t list_stride;
    uint32_t qid;
    const Cand* prev;           // running list of a batch (may be null)
    const uint32_t* prev_count;
    Cand* out;
    uint32_t* out_count;
    uint64_t* tau_out;
    uint32_t k;
    // global scratch for the rare case where the working set exceeds shared memory
    uint64_t* scratch_keys;
    uint32_t* scratch_src;
    uint32_t scratch_elems;     // power of two
    // optional fixed-size record output (sharded search)
    otters_topk_record* records;
    ShardMap map;
    int32_t take_max;
    // header of the output list + the counters copied into it
    ResultHeader* hdr;
    const unsigned long long* rows_scored_src;
    const unsigned long long* stats_src;
    // optional zero-copy result: mapped host memory (ResultHeader + k candidates) the final list is also written to
    uint8_t* host_out;
    // fused peer exchange of the row-sharded search (ex_world > 1): this rank's k records are stored straight into
    // every peer's record area over NVLink, a per-(query parity, rank) flag publishes them, and the same kernel
    // waits for the other ranks' records and merges all of them into `out`
    uint32_t ex_world, ex_rank, ex_kmax, ex_seq;  // ex_seq != 0: the flag value that publishes this query's records
    uint32_t ex_slot;                           // which of the kExchangeSlots record / flag areas this query uses
    uint32_t ex_k;                              // records every rank contributes (the caller's take count)
    otters_topk_record* ex_records[kMaxPeers];  // peer p's record area as mapped here: [kExchangeSlots][world][k_max]
    uint32_t* ex_flags[kMaxPeers];              // peer p's flag area: [kExchangeSlots][world]
};
// The opt-in dynamic shared-memory limit belongs to (kernel function, device) and is shared by every context of the
// process (a context's lanes included): it is tracked per device in a static of each launch wrapper and only ever raised.
// (Tracking it per context let a second context LOWER the limit under the first one's feet: invalid-argument launches.)
inline uint32_t& smem_limit_slot(uint32_t (&slots)[64]) {
    int dev = 0;
    cudaGetDevice(&dev);
    return slots[dev & 63];
}

// ---- scan kernel ------------------------------------------------------------------------------
struct ScanParams {
    const float* vectors;       // [n_rows][pitch_g] fp32 — or bf16 (half != 0: the pointer is then a uint16_t array)
    const float* inv_norms;     // [n_rows]
    const float* query;         // [dim_pad] zero padded, device
    float q_inv;                // 1/|q| (0 for a zero query), computed like src/vec.rs:390-397
    uint64_t pitch_g;           // elements per stored row (dim rounded up to 4 floats / 8 bf16: 16-byte rows)
    uint32_t half;              // rows are bf16 (OTTERS_VECTORS_FMT_BF16)
    uint32_t dim;
    uint32_t dim_pad;
    uint32_t n_rows;
    const uint32_t* row_mask;   // Lsb0 32-bit words, bit=1 keep; null = all rows
    uint32_t row_mask_words;    // words available; rows past them are kept
    // fused predicate (MetaStore): when flt_leaves != null the producer evaluates the CNF itself for the
    // rows of every unit whose chunk survived pruning, instead of reading a precomputed row mask
    const DevLeaf* flt_leaves;
    const uint32_t* flt_clause_off;
    uint32_t flt_n_clauses;
    uint32_t flt_n_leaves;
    const uint32_t* chunk_keep; // Lsb0 words, one bit per chunk (stand-alone prune kernel); null with flt_leaves set = LAZY
                                // pruning: every warp evaluates the zonemap / Bloom rules of its unit's chunks itself and the
                                // unit holding a chunk's first row accounts it in `stats` (no K0 launch on the query path)
    uint32_t chunk_size;
    uint32_t n_chunks;
    uint32_t nq_stats;          // lazy pruning: queries of the batch (vectors_compared = sum over kept chunks of len * nq)
    unsigned long long* stats;  // lazy pruning: [0] evaluated chunks, [1] vectors_compared
    uint32_t off_filter;        // shared-memory offset of the staged filter
    uint32_t n_units;
    uint32_t unit_rows;         // 32, 64 or 128: rows of the first n_big work units
    uint32_t unit_small;        // 16 or 32: rows of the remaining units (the tail of the store)
    uint32_t n_big;
    uint32_t* unit_counter;
    // selection
    uint32_t k;
    uint32_t cap;               // per-CTA candidate buffer capacity (power of two >= 2k)
    int32_t take_max;
    int32_t has_filter;
    float thr;
    int32_t cmp;
    const uint64_t* tau_in;     // device: running threshold key from earlier queries of a batch (0 = none)
    unsigned long long* g_tau;  // device: grid-wide threshold shared by the CTAs of this launch (zeroed per query; may be null)
    uint32_t qid;
    // staging layout
    uint32_t kc;                // columns per slot
    uint32_t nkc;               // slots steps per tile
    uint32_t pitch_s;           // floats per staged row in shared memory (== 8 mod 32)
    uint32_t slots;             // slots per warp
    uint32_t planners;          // planner front-end (scan_planner.cu): planner warps per CTA (0 = autonomous warps)
    uint32_t off_ring;          // planner front-end: shared-memory offset of the tile ring
    uint32_t off_query, off_warps, warp_bytes, off_w_rows, off_w_info, off_w_inv, off_w_list, off_w_slots;
    // outputs (fused mode)
    uint64_t* cta_keys;         // [grid][k]
    uint32_t* cta_counts;       // [grid]
    // outputs (emit-all mode)
    Cand* emit;                 // candidate array
    uint32_t* emit_count;
    uint32_t emit_cap;
    // instrumentation
    unsigned long long* rows_scored;
    uint32_t claim_depth;       // unit ids every claiming warp keeps in flight (1..4)
    uint32_t pred_seq, no_prefetch;  // experiment switches (OTTERS_PRED_SEQ / OTTERS_NO_PREFETCH)
    // fused selection: the last CTA to publish its list (ticket from done_counter) runs K3 itself — one launch per query
    uint32_t fuse_select;
    uint32_t* done_counter;
    SelectParams sel;
};

struct ScanLaunch {
    uint32_t grid, block, smem_bytes;
};

int launch_scan(const ScanParams& p, const ScanLaunch& l, int metric, bool emit_all, uint32_t* smem_configured, cudaStream_t s);
int launch_scan_planner(const ScanParams& p, const ScanLaunch& l, int metric, uint32_t* smem_configured, cudaStream_t s);

// ---- selection kernels ------------------------------------------------------------------------
int launch_select(const SelectParams& p, cudaStream_t s);

// records (after all-gather) -> global best k
int launch_merge_records(const otters_topk_record* recs, uint32_t n, uint32_t k, int take_max, Cand* out,
                         uint32_t* out_count, ResultHeader* hdr, const unsigned long long* rows_scored_src, uint64_t* scratch_keys,
                         uint32_t* scratch_src, uint32_t scratch_elems, cudaStream_t s);
// full sort of a candidate array (emit-all path); n_pow2 elements, padded with key = 0
int launch_global_sort(Cand* buf, uint64_t n_pow2, cudaStream_t s);
int launch_append_prev(Cand* buf, const uint32_t* emit_count, const Cand* prev, const uint32_t* prev_count, uint64_t n_pow2,
                       cudaStream_t s);
int launch_take_sorted(const Cand* buf, const uint32_t* emit_count, const uint32_t* prev_count, uint64_t k, Cand* out,
                       uint32_t* out_count, uint64_t* tau_out, ResultHeader* hdr, const unsigned long long* rows_scored_src,
                       const unsigned long long* stats_src, cudaStream_t s, const unsigned long long* extra_src = nullptr);
int launch_cands_to_records(const Cand* cands, const uint32_t* count, uint32_t k, ShardMap map, int take_max,
                            otters_topk_record* recs, cudaStream_t s);

// ---- batched (tensor-core) path: batched.cu -------------------------------------------------------------
struct BatchParams {
    uint32_t n_rows, nq;
    uint32_t n_qtiles, nkb;      // filled by launch_batch
    const float* inv_norms;
    const float* q_scal;         // per query: 1/|q| (cosine) or |q|^2 (euclidean); unused for dot
    int32_t take_max;
    int32_t has_filter;
    float thr;
    int32_t cmp;
    const float* delta;          // device: bound on |approximate - exact| score (loosens the vec_filter)
    const uint32_t* row_mask;    // Lsb0 words, bit = 1 keep; null = all rows
    uint32_t row_mask_words;
    uint32_t k, cap;
    unsigned long long* g_tau;   // grid-wide threshold key shared by the CTAs (selection speed only)
    uint32_t* g_flags;           // bit 0: a non-finite approximate score was seen
    uint32_t* g_excl;            // max over CTAs of the goodness (key >> 32) of the best excluded pair; 0 = none
    unsigned long long* pairs_scored;
    uint64_t* cta_keys;          // [grid][k] approximate keys, best first
    uint32_t* cta_qids;          // [grid][k]
    uint32_t* cta_counts;        // [grid]
    uint32_t passes;             // arithmetic rung: 3 (3xTF32 split, fp32-faithful), 1 (one tf32 MMA per product: selection only, wider
                                 // error bound) or 2 (one bf16 MMA per product on the bf16 shadow arrays: widest bound, fastest)
    uint32_t kps;                // single pass: k-blocks per pipeline stage (1 or 2: one barrier round trip and one commit for two)
    uint32_t pair_direct;        // CTA pairs, single pass: both CTAs' TMA loads credit the leader's barrier themselves (no relay warp)
    uint32_t redo_general;       // lean epilogue: a chunk holding a candidate is redone on the general straight-line path (1) or column by column (0)
    uint32_t epi_warps;          // epilogue warps per CTA: 8 (default: two per TMEM lane quarter, even / odd column chunks) or 4
    uint32_t dbg;                // timing experiments only (OTTERS_BATCH_DBG): 1 no loads, 2 no split, 4 no epilogue, 8 no MMAs
};
struct BatchLaunch {
    const float* vectors;
    uint64_t n_rows, pitch_g;
    uint32_t dim, dim_pad;
    const float* q_hi;
    const float* q_lo;
    uint32_t nq_pad;             // multiple of kBatchQueries
    const uint16_t* v_half;      // bf16 rung: bf16 shadow of the store rows [n_rows][pitch_h] ...
    const uint16_t* q_half;      // ... and of the (zero padded) queries [nq_pad][q_pitch_h]
    uint64_t pitch_h;
    uint32_t q_pitch_h;
    uint32_t grid;               // CTAs (a multiple of cta_group)
    uint32_t cta_group;          // 1: one CTA per 128-row tile; 2: CTA pairs (tcgen05 cta_group::2) on 256-row tiles
};
struct RescoreParams {
    uint32_t half;               // the store's rows are bf16 (vectors then points at uint16_t elements, pitch_g counts them)
    const float* vectors;
    const float* inv_norms;
    const float* queries;        // [nq][dim_pad] raw fp32
    const float* q_inv;          // per query 1/|q| (cosine)
    uint64_t pitch_g;
    uint32_t dim, dim_pad;
    const uint64_t* cta_keys;
    const uint32_t* cta_qids;
    const uint32_t* cta_counts;
    uint32_t n_lists, k;
    int32_t take_max, has_filter;
    float thr;
    int32_t cmp;
    Cand* out;                   // [out_slots] exact candidates (key 0 = dropped)
    uint32_t out_slots;
    uint32_t* out_count;         // number of surviving candidates
    uint32_t* max_err_bits;
};
uint32_t batch_smem_bytes(uint32_t cap);
int launch_split_queries(const float* q, uint32_t nq, uint32_t nq_pad, uint32_t dim_pad, float* qh, float* ql, float* qn2,
                         uint32_t* qmax2_bits, cudaStream_t s);
int launch_batch_delta(int metric, uint32_t dim, uint32_t passes, const uint32_t* qmax2_bits, const uint32_t* vmin_inv_bits, float* delta,
                       cudaStream_t s);
int launch_batch(const BatchLaunch& l, BatchParams p, int metric, uint32_t* smem_configured, cudaStream_t s);
int launch_rescore(const RescoreParams& p, int metric, uint32_t n_sort, cudaStream_t s);
int launch_min_inv_norm(const float* inv, uint64_t n, uint32_t* out_bits, cudaStream_t s);
int launch_convert_bf16(const float* src, uint64_t src_pitch, uint64_t n_src, uint32_t dim, uint16_t* dst, uint64_t dst_pitch, uint64_t n_dst,
                        cudaStream_t s);

// ---- store kernels ----------------------------------------------------------------------------
int launch_inv_norms(const float* rows, uint64_t pitch_g, uint32_t dim, uint64_t first, uint64_t n, float* out,
                     cudaStream_t s);
int launch_inv_norms_bf16(const uint16_t* rows, uint64_t pitch_g, uint32_t dim, uint64_t first, uint64_t n, float* out, cudaStream_t s);
int launch_synth_fill(float* rows, uint64_t pitch_g, uint32_t dim, uint64_t dst_first, ShardMap gen_map, uint64_t n,
                      uint64_t seed, cudaStream_t s);
int launch_synth_fill_at(float* rows, uint64_t pitch_g, uint32_t dim, ShardMap gen_map, uint64_t gen_first, uint64_t n, uint64_t seed,
                         cudaStream_t s);

// ---- device-side MetaStore build: build.cu ---------------------------------------------------------------------------
int launch_zonemap(int dtype, const void* values, const uint32_t* null_words, uint64_t n_rows, uint64_t chunk_size, uint64_t n_chunks,
                   void* zmin, void* zmax, uint32_t* non_null, cudaStream_t s);
int launch_string_hash(const uint8_t* bytes, const uint64_t* offsets, const uint32_t* null_words, uint64_t n_rows, uint64_t* hashes, cudaStream_t s);
int launch_bloom_build(const uint64_t* hashes, const uint32_t* null_words, uint64_t n_rows, uint64_t chunk_size, uint64_t n_chunks,
                       const uint64_t* mbits, const uint32_t* khash, uint64_t stride_words, uint64_t* words, uint32_t* non_null, cudaStream_t s);
int launch_dict_insert(const uint64_t* hashes, const uint32_t* null_words, uint64_t n_rows, uint64_t* keys, uint32_t* rep, uint64_t table_size,
                       cudaStream_t s);
int launch_dict_collect(const uint64_t* keys, const uint32_t* rep, uint64_t table_size, uint32_t* list_slot, uint32_t* list_rep, uint32_t* count,
                        uint32_t cap, cudaStream_t s);
int launch_dict_encode(const uint32_t* list_slot, const uint32_t* list_code, uint32_t n_distinct, uint32_t* slot_code, const uint64_t* hashes,
                       const uint32_t* null_words, const uint8_t* bytes, const uint64_t* offsets, uint64_t n_rows, const uint64_t* keys,
                       const uint32_t* rep, uint64_t table_size, uint32_t* codes, uint32_t* mismatch, cudaStream_t s);

// ---- metadata kernels -------------------------------------------------------------------------
struct DevColumn {
    int32_t dtype;
    int32_t pad;
    const void* values;          // i32 / i64 / f32 / f64 / u32 dictionary codes
    const uint32_t* null_words;  // Lsb0 32-bit words, bit=1 null; may be null
    const void* zmin;            // per chunk, typed like the reference's PackedRanges (src/meta.rs:71-76)
    const void* zmax;
    const uint32_t* non_null;    // per chunk
    const uint64_t* bloom;       // per chunk Bloom words, stride bloom_stride
    uint64_t bloom_stride;
    const uint64_t* bloom_mbits; // per chunk
    const uint32_t* bloom_k;     // per chunk
};

enum LeafExec : int32_t { LEAF_I32 = 0, LEAF_I64 = 1, LEAF_F32 = 2, LEAF_F64 = 3, LEAF_STR = 4 };

struct DevLeaf {         // one lowered leaf, self-contained: carries the device pointers of its column
    uint32_t col;
    int32_t op;
    int32_t exec;        // LeafExec
    int32_t code_valid;  // LEAF_STR: literal present in the dictionary
    uint32_t tt;         // 4-bit truth table of the operator over the compare states (predicate.cuh): bit 0 value > literal,
                         // 1 value < literal, 2 equal, 3 unordered
    uint32_t pad0;
    int64_t i64;
    double f64;
    float f32;
    int32_t i32;
    uint32_t code;
    uint32_t bloom_k0;   // LEAF_STR: probes of a full chunk's filter
    uint64_t h1, h2;     // LEAF_STR: Bloom probe hashes of the literal
    uint64_t bloom_m0, bloom_a0, bloom_b0;  // filter size of a full chunk and h1 % m0, h2 % m0 (step, never 0)
    uint64_t bloom_full_chunks;             // chunks [0, n_rows / chunk_size) are full: m = m0, k = k0 without a load
    const void* values;
    const uint32_t* null_words;
    const void* zmin;
    const void* zmax;
    const uint32_t* non_null;
    const uint64_t* bloom;
    uint64_t bloom_stride;
    const uint64_t* bloom_mbits;
    const uint32_t* bloom_k;
};

struct DevFilter {       // variable-length: header, clause offsets, leaves (all in one device buffer)
    uint32_t n_clauses;
    uint32_t n_leaves;
};

struct MetaKernelParams {
    const DevColumn* cols;
    const uint32_t* clause_off;  // n_clauses + 1
    const DevLeaf* leaves;
    uint32_t n_clauses;
    uint32_t n_rows;
    uint32_t chunk_size;
    uint32_t n_chunks;
    uint32_t nq;
    uint32_t* chunk_keep;        // words
    uint32_t* row_mask;          // words
    unsigned long long* stats;   // [0] evaluated chunks, [1] vectors_compared
};
int launch_prune(const MetaKernelParams& p, uint32_t n_leaves, cudaStream_t s);
int launch_rowmask(const MetaKernelParams& p, uint32_t n_leaves, cudaStream_t s);
int launch_count_all_chunks(const MetaKernelParams& p, cudaStream_t s);
int launch_gather(const void* values, const uint32_t* null_words, uint32_t width, const uint32_t* rows, uint32_t n, void* out_values,
                  uint8_t* out_nulls, cudaStream_t s);

}  // namespace otters
