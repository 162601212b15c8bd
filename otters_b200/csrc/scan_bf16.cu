// scan_bf16.cu — K1 (scan_kernel.cuh) instantiated for bf16 rows (OTTERS_VECTORS_FMT_BF16).
#include "scan_kernel.cuh"

namespace otters {

int launch_scan_bf16(const ScanParams& p, const ScanLaunch& l, int metric, bool emit_all, uint32_t* smem_configured, cudaStream_t s) {
    return scan_impl::launch_scan_fmt<true>(p, l, metric, emit_all, smem_configured, s);
}

}  // namespace otters
