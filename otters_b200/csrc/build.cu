// build.cu — device-side MetaStore build (MetaStoreBuilder::build, reference src/meta.rs:151-305): per-chunk zonemaps
// (min / max / non-null, src/meta_compute.rs:32-132), per-chunk string Bloom filters (:99-116) and the dictionary encoding
// of string columns, as kernels over the uploaded columns.  Every table is bit-identical to what the host path (api.cu:
// build_column) builds: min / max / counts and bit-set unions do not depend on the order of evaluation.
#include "internal.h"

namespace otters {
namespace {

__device__ __forceinline__ bool is_null_bit(const uint32_t* null_words, uint64_t row) {
    return null_words && ((null_words[row >> 5] >> (row & 31)) & 1u);
}

// ---- zonemaps: one warp per chunk ------------------------------------------------------------------------------------
// Integers are reduced in int64 starting from (INT64_MAX, INT64_MIN) and narrowed with Rust's wrapping `as` for Int32
// columns (src/meta.rs:254-255), floats in double with fmin / fmax (NaN ignored, f64::min / max, src/meta_compute.rs:69-83)
// starting from (+inf, -inf) and narrowed to f32 for Float32 columns: an all-NULL chunk keeps the start values.
template <typename T, typename OUT>
__global__ void __launch_bounds__(256) zonemap_int_kernel(const T* values, const uint32_t* null_words, uint64_t n_rows, uint64_t chunk_size,
                                                          uint64_t n_chunks, OUT* zmin, OUT* zmax, uint32_t* non_null) {
    const uint64_t ch = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (ch >= n_chunks) return;
    const uint64_t s = ch * chunk_size, e = s + chunk_size < n_rows ? s + chunk_size : n_rows;
    long long lo = 0x7FFFFFFFFFFFFFFFll, hi = (long long)0x8000000000000000ull;
    uint32_t cnt = 0;
    for (uint64_t i = s + lane; i < e; i += 32) {
        if (!is_null_bit(null_words, i)) {
            const long long v = (long long)values[i];
            lo = v < lo ? v : lo;
            hi = v > hi ? v : hi;
            ++cnt;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const long long olo = __shfl_xor_sync(0xFFFFFFFFu, lo, d), ohi = __shfl_xor_sync(0xFFFFFFFFu, hi, d);
        lo = olo < lo ? olo : lo;
        hi = ohi > hi ? ohi : hi;
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
    }
    if (lane == 0) {
        zmin[ch] = (OUT)(unsigned long long)lo;  // wrapping narrowing for OUT = int32_t
        zmax[ch] = (OUT)(unsigned long long)hi;
        non_null[ch] = cnt;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) zonemap_float_kernel(const T* values, const uint32_t* null_words, uint64_t n_rows, uint64_t chunk_size,
                                                            uint64_t n_chunks, T* zmin, T* zmax, uint32_t* non_null) {
    const uint64_t ch = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (ch >= n_chunks) return;
    const uint64_t s = ch * chunk_size, e = s + chunk_size < n_rows ? s + chunk_size : n_rows;
    double lo = __longlong_as_double(0x7FF0000000000000ll), hi = __longlong_as_double((long long)0xFFF0000000000000ull);
    uint32_t cnt = 0;
    for (uint64_t i = s + lane; i < e; i += 32) {
        if (!is_null_bit(null_words, i)) {
            const double v = (double)values[i];
            lo = fmin(lo, v);
            hi = fmax(hi, v);
            ++cnt;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xFFFFFFFFu, lo, d));
        hi = fmax(hi, __shfl_xor_sync(0xFFFFFFFFu, hi, d));
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
    }
    if (lane == 0) {
        zmin[ch] = (T)lo;
        zmax[ch] = (T)hi;
        non_null[ch] = cnt;
    }
}

// ---- strings ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64_dev(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// FNV-1a 64 of every non-NULL string (the base hash of the Bloom spec; also the dictionary key)
__global__ void string_hash_kernel(const uint8_t* bytes, const uint64_t* offsets, const uint32_t* null_words, uint64_t n_rows, uint64_t* hashes) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    uint64_t h = 0xCBF29CE484222325ull;
    if (!is_null_bit(null_words, i))
        for (uint64_t b = offsets[i]; b < offsets[i + 1]; ++b) {
            h ^= bytes[b];
            h *= 0x100000001B3ull;
        }
    hashes[i] = h;
}

// Bloom insert: one block per chunk, the chunk's filter is assembled in shared memory (m <= kBloomSmemBits) or straight in
// global memory with atomics.  Probe i = (a + i*b) mod m, a = mix64(h) mod m, b = (mix64(h ^ C) | 1) mod m (1 if 0).
constexpr uint32_t kBloomSmemWords = 4096;  // 256 Kbit

__global__ void __launch_bounds__(256) bloom_build_kernel(const uint64_t* hashes, const uint32_t* null_words, uint64_t n_rows, uint64_t chunk_size,
                                                          const uint64_t* mbits, const uint32_t* khash, uint64_t stride_words, uint64_t* words,
                                                          uint32_t* non_null) {
    __shared__ unsigned long long s_words[kBloomSmemWords];
    __shared__ uint32_t s_cnt;
    const uint64_t ch = blockIdx.x;
    const uint64_t s = ch * chunk_size, e = s + chunk_size < n_rows ? s + chunk_size : n_rows;
    const uint64_t m = mbits[ch];
    const uint32_t kh = khash[ch];
    const uint64_t m_words = (m + 63) / 64;
    const bool in_smem = m_words <= kBloomSmemWords;
    unsigned long long* out = reinterpret_cast<unsigned long long*>(words + ch * stride_words);
    if (in_smem)
        for (uint64_t w = threadIdx.x; w < m_words; w += blockDim.x) s_words[w] = 0ull;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    uint32_t cnt = 0;
    for (uint64_t i = s + threadIdx.x; i < e; i += blockDim.x) {
        if (is_null_bit(null_words, i)) continue;
        ++cnt;
        const uint64_t h = hashes[i];
        const uint64_t h1 = mix64_dev(h), h2 = mix64_dev(h ^ 0x9E3779B97F4A7C15ull) | 1ull;
        uint64_t bit = h1 % m, step = h2 % m;
        if (step == 0) step = 1;
        for (uint32_t j = 0; j < kh; ++j) {
            if (in_smem) atomicOr(&s_words[bit >> 6], 1ull << (bit & 63));
            else atomicOr(&out[bit >> 6], 1ull << (bit & 63));
            bit += step;
            if (bit >= m) bit -= m;
        }
    }
    if (cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (in_smem)
        for (uint64_t w = threadIdx.x; w < m_words; w += blockDim.x) out[w] = s_words[w];
    if (threadIdx.x == 0) non_null[ch] = s_cnt;
}

// ---- dictionary encoding ---------------------------------------------------------------------------------------------
// Open-addressing table keyed by the 64-bit string hash (linear probing, table size a power of two >= 2 * rows).  Every
// distinct hash claims one slot; the slot remembers the SMALLEST row holding it (its representative).  The host numbers the
// occupied slots by representative row — i.e. in order of first occurrence, exactly like the host path's dictionary — and a
// second pass writes the rows' codes.  A verification pass compares every row's bytes with its representative's: two
// different strings sharing a 64-bit hash raise a flag and the build falls back to the host dictionary (never observed).
constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

__global__ void dict_insert_kernel(const uint64_t* hashes, const uint32_t* null_words, uint64_t n_rows, unsigned long long* keys, uint32_t* rep,
                                   uint64_t mask) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows || is_null_bit(null_words, i)) return;
    unsigned long long h = hashes[i];
    if (h == kEmptyKey) h = kEmptyKey - 1;  // (reserved value; the verification pass still guards exactness)
    uint64_t slot = mix64_dev(h) & mask;
    for (;;) {
        const unsigned long long old = atomicCAS(&keys[slot], kEmptyKey, h);
        if (old == kEmptyKey || old == h) break;
        slot = (slot + 1) & mask;
    }
    atomicMin(&rep[slot], (uint32_t)i);
}

// occupied slots -> compact list (slot, representative row); order is arbitrary, the host sorts it
__global__ void dict_collect_kernel(const unsigned long long* keys, const uint32_t* rep, uint64_t table_size, uint32_t* list_slot, uint32_t* list_rep,
                                    uint32_t* count, uint32_t cap) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= table_size || keys[s] == kEmptyKey) return;
    const uint32_t at = atomicAdd(count, 1u);
    if (at < cap) {
        list_slot[at] = (uint32_t)s;
        list_rep[at] = rep[s];
    }
}

__global__ void dict_scatter_codes_kernel(const uint32_t* list_slot, const uint32_t* list_code, uint32_t n, uint32_t* slot_code) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slot_code[list_slot[i]] = list_code[i];
}

__global__ void dict_encode_kernel(const uint64_t* hashes, const uint32_t* null_words, const uint8_t* bytes, const uint64_t* offsets, uint64_t n_rows,
                                   const unsigned long long* keys, const uint32_t* rep, const uint32_t* slot_code, uint64_t mask, uint32_t* codes,
                                   uint32_t* mismatch) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    if (is_null_bit(null_words, i)) {
        codes[i] = 0xFFFFFFFFu;
        return;
    }
    unsigned long long h = hashes[i];
    if (h == kEmptyKey) h = kEmptyKey - 1;
    uint64_t slot = mix64_dev(h) & mask;
    while (keys[slot] != h) slot = (slot + 1) & mask;
    codes[i] = slot_code[slot];
    // byte equality with the representative (exactness does not rest on the hash)
    const uint64_t r = rep[slot];
    const uint64_t a0 = offsets[i], a1 = offsets[i + 1], b0 = offsets[r], b1 = offsets[r + 1];
    bool same = (a1 - a0) == (b1 - b0);
    for (uint64_t k = 0; same && k < a1 - a0; ++k) same = bytes[a0 + k] == bytes[b0 + k];
    if (!same) atomicExch(mismatch, 1u);
}

}  // namespace

int launch_zonemap(int dtype, const void* values, const uint32_t* null_words, uint64_t n_rows, uint64_t chunk_size, uint64_t n_chunks,
                   void* zmin, void* zmax, uint32_t* non_null, cudaStream_t s) {
    if (n_chunks == 0) return OTTERS_OK;
    const unsigned blocks = (unsigned)((n_chunks * 32 + 255) / 256);
    switch (dtype) {
    case OTTERS_DTYPE_INT32:
        zonemap_int_kernel<int32_t, int32_t><<<blocks, 256, 0, s>>>((const int32_t*)values, null_words, n_rows, chunk_size, n_chunks, (int32_t*)zmin, (int32_t*)zmax, non_null);
        break;
    case OTTERS_DTYPE_INT64:
    case OTTERS_DTYPE_DATETIME:
        zonemap_int_kernel<int64_t, int64_t><<<blocks, 256, 0, s>>>((const int64_t*)values, null_words, n_rows, chunk_size, n_chunks, (int64_t*)zmin, (int64_t*)zmax, non_null);
        break;
    case OTTERS_DTYPE_FLOAT32:
        zonemap_float_kernel<float><<<blocks, 256, 0, s>>>((const float*)values, null_words, n_rows, chunk_size, n_chunks, (float*)zmin, (float*)zmax, non_null);
        break;
    case OTTERS_DTYPE_FLOAT64:
        zonemap_float_kernel<double><<<blocks, 256, 0, s>>>((const double*)values, null_words, n_rows, chunk_size, n_chunks, (double*)zmin, (double*)zmax, non_null);
        break;
    default: return fail(OTTERS_ERR_INVALID, "column type has no zonemap");
    }
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_string_hash(const uint8_t* bytes, const uint64_t* offsets, const uint32_t* null_words, uint64_t n_rows, uint64_t* hashes, cudaStream_t s) {
    if (n_rows == 0) return OTTERS_OK;
    string_hash_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, s>>>(bytes, offsets, null_words, n_rows, hashes);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_bloom_build(const uint64_t* hashes, const uint32_t* null_words, uint64_t n_rows, uint64_t chunk_size, uint64_t n_chunks,
                       const uint64_t* mbits, const uint32_t* khash, uint64_t stride_words, uint64_t* words, uint32_t* non_null, cudaStream_t s) {
    if (n_chunks == 0) return OTTERS_OK;
    bloom_build_kernel<<<(unsigned)n_chunks, 256, 0, s>>>(hashes, null_words, n_rows, chunk_size, mbits, khash, stride_words, words, non_null);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_dict_insert(const uint64_t* hashes, const uint32_t* null_words, uint64_t n_rows, uint64_t* keys, uint32_t* rep, uint64_t table_size,
                       cudaStream_t s) {
    if (n_rows == 0) return OTTERS_OK;
    dict_insert_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, s>>>(hashes, null_words, n_rows, reinterpret_cast<unsigned long long*>(keys), rep,
                                                                        table_size - 1);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_dict_collect(const uint64_t* keys, const uint32_t* rep, uint64_t table_size, uint32_t* list_slot, uint32_t* list_rep, uint32_t* count,
                        uint32_t cap, cudaStream_t s) {
    dict_collect_kernel<<<(unsigned)((table_size + 255) / 256), 256, 0, s>>>(reinterpret_cast<const unsigned long long*>(keys), rep, table_size, list_slot,
                                                                             list_rep, count, cap);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_dict_encode(const uint32_t* list_slot, const uint32_t* list_code, uint32_t n_distinct, uint32_t* slot_code, const uint64_t* hashes,
                       const uint32_t* null_words, const uint8_t* bytes, const uint64_t* offsets, uint64_t n_rows, const uint64_t* keys,
                       const uint32_t* rep, uint64_t table_size, uint32_t* codes, uint32_t* mismatch, cudaStream_t s) {
    if (n_distinct) dict_scatter_codes_kernel<<<(n_distinct + 255) / 256, 256, 0, s>>>(list_slot, list_code, n_distinct, slot_code);
    if (n_rows)
        dict_encode_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, s>>>(hashes, null_words, bytes, offsets, n_rows,
                                                                            reinterpret_cast<const unsigned long long*>(keys), rep, slot_code,
                                                                            table_size - 1, codes, mismatch);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

}  // namespace otters
