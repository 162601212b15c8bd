// meta.cu — K0 (chunk pruning over zonemaps + Bloom filters) and K0b (per-row CNF predicate ->
// surviving-row bitmask that gates the scan kernel).
//
// K0  replaces MetaStore::build_chunk_mask_for_plan and its leaf helpers (reference
//     src/meta.rs:407-544) and the range kernels of src/type_utils.rs:446-584,739-889.
// K0b replaces build_row_mask_for_chunk and its leaf helpers (src/meta_compute.rs:194-318) and the
//     row kernels of src/type_utils.rs:306-444,586-736.
//
// Semantics: filter = AND over clauses of OR over leaves; a NULL row
// fails every leaf (including Neq); NaN satisfies only Neq; literals are cast to the column width
// on the host (api.cu) exactly as the reference does.
#include "internal.h"
#include "predicate.cuh"

namespace otters {
namespace {

// K0: one thread per chunk.  The lowered filter is staged in shared memory first and every leaf is evaluated
// unconditionally, so the zonemap / Bloom loads of all leaves are in flight together (three dependent memory
// round trips in total instead of one chain per leaf).
constexpr uint32_t kPruneSmemLeaves = 64;

__global__ void __launch_bounds__(64) prune_kernel(const __grid_constant__ MetaKernelParams p, uint32_t n_leaves) {
    __shared__ __align__(16) DevLeaf s_leaves[kPruneSmemLeaves];
    __shared__ uint32_t s_off[kPruneSmemLeaves + 1];
    const DevLeaf* leaves = p.leaves;
    const uint32_t* clause_off = p.clause_off;
    if (n_leaves <= kPruneSmemLeaves && p.n_clauses <= kPruneSmemLeaves) {
        const uint32_t words = n_leaves * (uint32_t)(sizeof(DevLeaf) / 4);
        for (uint32_t i = threadIdx.x; i < words; i += blockDim.x)
            reinterpret_cast<uint32_t*>(s_leaves)[i] = __ldg(reinterpret_cast<const uint32_t*>(p.leaves) + i);
        for (uint32_t i = threadIdx.x; i <= p.n_clauses; i += blockDim.x) s_off[i] = __ldg(p.clause_off + i);
        __syncthreads();
        leaves = s_leaves;
        clause_off = s_off;
    }
    const uint32_t ch = blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    uint32_t len = 0;
    if (ch < p.n_chunks) {
        keep = true;
        for (uint32_t ci = 0; ci < p.n_clauses; ++ci) {
            bool any = false;
            for (uint32_t li = clause_off[ci]; li < clause_off[ci + 1]; ++li) any |= chunk_leaf_sat(leaves[li], ch);
            keep &= any;
        }
        const uint64_t base = (uint64_t)ch * p.chunk_size;
        len = (uint32_t)(base + p.chunk_size <= p.n_rows ? p.chunk_size : p.n_rows - base);
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
    const int lane = threadIdx.x & 31;
    if (lane == 0 && (ch >> 5) < ((p.n_chunks + 31) >> 5)) p.chunk_keep[ch >> 5] = m;
    // stats: evaluated chunks and vectors_compared = sum over evaluated chunks of len * nq
    // (src/meta_compute.rs:166, src/meta.rs:666-669)
    unsigned long long vc = keep ? (unsigned long long)len * p.nq : 0ull;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) vc += __shfl_xor_sync(0xFFFFFFFFu, vc, d);
    if (lane == 0 && m) {
        atomicAdd(&p.stats[0], (unsigned long long)__popc(m));
        atomicAdd(&p.stats[1], vc);
    }
}

// K0, leaf-parallel form (CNFs of up to 32 leaves): a block of 256 threads owns 32 chunks = one word of the keep
// bitmask; 8 threads share a chunk and evaluate its leaves side by side, so the zonemap / Bloom loads of ALL leaves are
// in flight at once (the thread-per-chunk form walks the leaves one after the other: one memory round trip each).
__global__ void __launch_bounds__(256) prune_leafpar_kernel(const __grid_constant__ MetaKernelParams p, uint32_t n_leaves) {
    __shared__ __align__(16) DevLeaf s_leaves[32];
    __shared__ uint32_t s_off[33];
    __shared__ uint32_t s_word;
    __shared__ unsigned long long s_vc;
    const uint32_t words = n_leaves * (uint32_t)(sizeof(DevLeaf) / 4);
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x)
        reinterpret_cast<uint32_t*>(s_leaves)[i] = __ldg(reinterpret_cast<const uint32_t*>(p.leaves) + i);
    for (uint32_t i = threadIdx.x; i <= p.n_clauses && i <= 32; i += blockDim.x) s_off[i] = __ldg(p.clause_off + i);
    if (threadIdx.x == 0) {
        s_word = 0;
        s_vc = 0;
    }
    __syncthreads();
    const uint32_t sub = threadIdx.x & 7u;              // which leaves of the chunk this thread evaluates
    const uint32_t cl = threadIdx.x >> 3;               // chunk within the block's word
    const uint32_t ch = blockIdx.x * 32 + cl;
    uint32_t mask = 0;
    if (ch < p.n_chunks)
        for (uint32_t li = sub; li < n_leaves; li += 8)
            if (chunk_leaf_sat(s_leaves[li], ch)) mask |= 1u << li;
    mask |= __shfl_xor_sync(0xFFFFFFFFu, mask, 1);
    mask |= __shfl_xor_sync(0xFFFFFFFFu, mask, 2);
    mask |= __shfl_xor_sync(0xFFFFFFFFu, mask, 4);
    if (sub == 0 && ch < p.n_chunks) {
        bool keep = true;
        for (uint32_t ci = 0; ci < p.n_clauses; ++ci) {
            const uint32_t a = s_off[ci], b = s_off[ci + 1];
            const uint32_t cm = b > a ? (uint32_t)(((1ull << (b - a)) - 1ull) << a) : 0u;
            keep &= (mask & cm) != 0;
        }
        if (keep) {
            const uint64_t base = (uint64_t)ch * p.chunk_size;
            const uint32_t len = (uint32_t)(base + p.chunk_size <= p.n_rows ? p.chunk_size : p.n_rows - base);
            atomicOr(&s_word, 1u << cl);
            atomicAdd(&s_vc, (unsigned long long)len * p.nq);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        p.chunk_keep[blockIdx.x] = s_word;
        // stats: evaluated chunks and vectors_compared = sum over evaluated chunks of len * nq
        // (src/meta_compute.rs:166, src/meta.rs:666-669)
        if (s_word) {
            atomicAdd(&p.stats[0], (unsigned long long)__popc(s_word));
            atomicAdd(&p.stats[1], s_vc);
        }
    }
}

// no meta_filter: every chunk is evaluated (src/meta.rs:658)
__global__ void count_all_kernel(const __grid_constant__ MetaKernelParams p) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        p.stats[0] = p.n_chunks;
        p.stats[1] = (unsigned long long)p.n_rows * p.nq;
    }
}

// K0b: the surviving-row bitmask.  A warp owns groups of 128 rows (four mask words); a lane owns row `lane` of each of the
// four words, so every load is coalesced (32 consecutive values; warp-uniform null and chunk words), the four rows' loads of
// a leaf are in flight together, and the type / operator dispatch is paid once per leaf (rows_pass_x4 + truth tables).
// Round 1's form (one row per lane, a switch per row and leaf) was bound by instruction issue and by one dependent round trip
// per 32 rows: 94 us for the 10M-row target.  The chunk bits of the NEXT group are loaded while the current one is evaluated.
// Rows of pruned chunks are never read.
__device__ __forceinline__ uint32_t chunk_live4(const MetaKernelParams& p, uint32_t r0, uint32_t lane) {
    uint32_t live = 0;
#pragma unroll
    for (uint32_t u = 0; u < 4; ++u) {
        const uint32_t row = r0 + 32u * u + lane;
        if (row < p.n_rows) {
            const uint32_t ch = row / p.chunk_size;
            live |= ((__ldg(p.chunk_keep + (ch >> 5)) >> (ch & 31)) & 1u) << u;
        }
    }
    return live;
}

__global__ void __launch_bounds__(256) rowmask_kernel(const __grid_constant__ MetaKernelParams p, uint32_t n_leaves, int use_smem) {
    extern __shared__ __align__(16) uint8_t fsm[];
    const DevLeaf* leaves = p.leaves;
    const uint32_t* clause_off = p.clause_off;
    if (use_smem) {
        DevLeaf* sl = reinterpret_cast<DevLeaf*>(fsm);
        uint32_t* so = reinterpret_cast<uint32_t*>(fsm + (size_t)n_leaves * sizeof(DevLeaf));
        const uint32_t words = n_leaves * (uint32_t)(sizeof(DevLeaf) / 4);
        for (uint32_t i = threadIdx.x; i < words; i += blockDim.x)
            reinterpret_cast<uint32_t*>(sl)[i] = reinterpret_cast<const uint32_t*>(p.leaves)[i];
        for (uint32_t i = threadIdx.x; i <= p.n_clauses; i += blockDim.x) so[i] = p.clause_off[i];
        __syncthreads();
        leaves = sl;
        clause_off = so;
    }
    const uint32_t n_words = (p.n_rows + 31) >> 5;
    const uint32_t n_groups = (p.n_rows + 127) >> 7;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t live = g < n_groups ? chunk_live4(p, g << 7, lane) : 0;
    for (; g < n_groups; g += warps_total) {
        const uint32_t g_next = g + warps_total;
        const uint32_t live_next = g_next < n_groups ? chunk_live4(p, g_next << 7, lane) : 0;  // in flight during the evaluation
        const uint32_t pass = __any_sync(0xFFFFFFFFu, live != 0) ? rows_pass_x4(leaves, clause_off, p.n_clauses, g << 7, lane, live) : 0u;
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u) {
            const unsigned m = __ballot_sync(0xFFFFFFFFu, (pass >> u) & 1u);
            const uint32_t w = (g << 2) + u;
            if (lane == 0 && w < n_words) p.row_mask[w] = m;
        }
        live = live_next;
    }
}

// Result-column gather (MetaQueryPlan::collect's `data`, reference src/meta.rs:723-821): values and NULL flags of the result
// rows of one column.  W = value width in bytes (4: Int32 / Float32 / dictionary codes, 8: Int64 / Float64 / DateTime).
template <typename T>
__global__ void gather_kernel(const T* values, const uint32_t* null_words, const uint32_t* rows, uint32_t n, T* out_values, uint8_t* out_nulls) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t row = rows[i];
    out_values[i] = values[row];
    out_nulls[i] = null_words ? (uint8_t)((null_words[row >> 5] >> (row & 31)) & 1u) : (uint8_t)0;
}

}  // namespace

int launch_gather(const void* values, const uint32_t* null_words, uint32_t width, const uint32_t* rows, uint32_t n, void* out_values,
                  uint8_t* out_nulls, cudaStream_t s) {
    if (n == 0) return OTTERS_OK;
    const unsigned blocks = (n + 255) / 256;
    if (width == 8)
        gather_kernel<uint64_t><<<blocks, 256, 0, s>>>((const uint64_t*)values, null_words, rows, n, (uint64_t*)out_values, out_nulls);
    else
        gather_kernel<uint32_t><<<blocks, 256, 0, s>>>((const uint32_t*)values, null_words, rows, n, (uint32_t*)out_values, out_nulls);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_prune(const MetaKernelParams& p, uint32_t n_leaves, cudaStream_t s) {
    if (p.n_chunks == 0) return OTTERS_OK;
    if (n_leaves >= 1 && n_leaves <= 32 && p.n_clauses <= 32)
        prune_leafpar_kernel<<<(p.n_chunks + 31) / 32, 256, 0, s>>>(p, n_leaves);
    else
        prune_kernel<<<(p.n_chunks + 63) / 64, 64, 0, s>>>(p, n_leaves);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_rowmask(const MetaKernelParams& p, uint32_t n_leaves, cudaStream_t s) {
    if (p.n_rows == 0) return OTTERS_OK;
    uint32_t n_groups = (p.n_rows + 127) >> 7;
    uint32_t blocks = (n_groups + 7) / 8;  // 8 warps per block, one 128-row group per warp and iteration
    if (blocks > 148 * 8) blocks = 148 * 8;
    size_t smem = (size_t)n_leaves * sizeof(DevLeaf) + ((size_t)p.n_clauses + 1) * 4;
    int use_smem = smem <= 40 * 1024;
    rowmask_kernel<<<blocks, 256, use_smem ? smem : 0, s>>>(p, n_leaves, use_smem);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_count_all_chunks(const MetaKernelParams& p, cudaStream_t s) {
    count_all_kernel<<<1, 32, 0, s>>>(p);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

}  // namespace otters
