// store.cu — device-side store maintenance kernels: per-row inverse norms and the synthetic
// row generator used by tests and bench.py.
#include "internal.h"

namespace otters {
namespace {

// VecStore::add_vector (reference src/vec.rs:357-371): norm = sqrt(serial f32 sum of x*x);
// inv = norm != 0 ? 1/norm : 0.  One thread per row, strictly in column order, multiply then add
// (no FMA) so the value is bit-identical to the CPU path.
__global__ void inv_norms_kernel(const float* __restrict__ rows, uint64_t pitch_g, uint32_t dim, uint64_t first, uint64_t n,
                                 float* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* v = rows + (first + i) * pitch_g;
    float s = -0.0f;  // Rust's f32 Sum identity
    uint32_t c = 0;
    for (; c + 4 <= dim; c += 4) {
        const float4 x = *reinterpret_cast<const float4*>(v + c);
        s = __fadd_rn(s, __fmul_rn(x.x, x.x));
        s = __fadd_rn(s, __fmul_rn(x.y, x.y));
        s = __fadd_rn(s, __fmul_rn(x.z, x.z));
        s = __fadd_rn(s, __fmul_rn(x.w, x.w));
    }
    for (; c < dim; ++c) s = __fadd_rn(s, __fmul_rn(v[c], v[c]));
    const float norm = __fsqrt_rn(s);
    out[first + i] = norm != 0.0f ? __fdiv_rn(1.0f, norm) : 0.0f;
}

// the same for bf16 rows: every element widened (exactly) to fp32 first — the norm of what the store actually holds
__global__ void inv_norms_bf16_kernel(const uint16_t* __restrict__ rows, uint64_t pitch_g, uint32_t dim, uint64_t first, uint64_t n,
                                      float* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint16_t* v = rows + (first + i) * pitch_g;
    float s = -0.0f;
    uint32_t c = 0;
    for (; c + 8 <= dim; c += 8) {
        const uint4 w = *reinterpret_cast<const uint4*>(v + c);
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = __uint_as_float(ww[j] << 16), b = __uint_as_float(ww[j] & 0xFFFF0000u);
            s = __fadd_rn(s, __fmul_rn(a, a));
            s = __fadd_rn(s, __fmul_rn(b, b));
        }
    }
    for (; c < dim; ++c) {
        const float a = __uint_as_float((uint32_t)v[c] << 16);
        s = __fadd_rn(s, __fmul_rn(a, a));
    }
    const float norm = __fsqrt_rn(s);
    out[first + i] = norm != 0.0f ? __fdiv_rn(1.0f, norm) : 0.0f;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// x = (splitmix64(seed ^ (row*dim + col)) >> 40) * 2^-23 - 1   in [-1, 1)   
__global__ void synth_fill_kernel(float* __restrict__ rows, uint64_t pitch_g, uint32_t dim, uint64_t dst_first,
                                  ShardMap gen_map, uint64_t gen_first, uint64_t n, uint64_t seed) {
    const uint64_t total = n * pitch_g;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        const uint64_t r = i / pitch_g;
        const uint32_t c = (uint32_t)(i - r * pitch_g);
        float x = 0.f;
        if (c < dim) {
            const uint64_t u = splitmix64(seed ^ (gen_map.global_row(gen_first + r) * (uint64_t)dim + c));
            x = (float)(u >> 40) * (1.0f / 8388608.0f) - 1.0f;
        }
        rows[(dst_first + r) * pitch_g + c] = x;
    }
}

}  // namespace

int launch_inv_norms(const float* rows, uint64_t pitch_g, uint32_t dim, uint64_t first, uint64_t n, float* out,
                     cudaStream_t s) {
    if (n == 0) return OTTERS_OK;
    inv_norms_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(rows, pitch_g, dim, first, n, out);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_inv_norms_bf16(const uint16_t* rows, uint64_t pitch_g, uint32_t dim, uint64_t first, uint64_t n, float* out, cudaStream_t s) {
    if (n == 0) return OTTERS_OK;
    inv_norms_bf16_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(rows, pitch_g, dim, first, n, out);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

int launch_synth_fill(float* rows, uint64_t pitch_g, uint32_t dim, uint64_t dst_first, ShardMap gen_map, uint64_t n,
                      uint64_t seed, cudaStream_t s) {
    if (n == 0) return OTTERS_OK;
    synth_fill_kernel<<<148 * 16, 256, 0, s>>>(rows, pitch_g, dim, dst_first, gen_map, 0, n, seed);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

// rows [gen_first, gen_first + n) of the generator's local numbering, written to the START of `rows` (a staging slab)
int launch_synth_fill_at(float* rows, uint64_t pitch_g, uint32_t dim, ShardMap gen_map, uint64_t gen_first, uint64_t n, uint64_t seed,
                         cudaStream_t s) {
    if (n == 0) return OTTERS_OK;
    synth_fill_kernel<<<148 * 16, 256, 0, s>>>(rows, pitch_g, dim, 0, gen_map, gen_first, n, seed);
    OTTERS_CUDA(cudaGetLastError());
    return OTTERS_OK;
}

}  // namespace otters
