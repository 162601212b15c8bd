// scan.cu — K1 (scan_kernel.cuh) instantiated for fp32 rows, and the entry point that picks the row format.
#include "scan_kernel.cuh"

namespace otters {

int launch_scan_bf16(const ScanParams& p, const ScanLaunch& l, int metric, bool emit_all, uint32_t* smem_configured, cudaStream_t s);  // scan_bf16.cu

int launch_scan(const ScanParams& p, const ScanLaunch& l, int metric, bool emit_all, uint32_t* smem_configured, cudaStream_t s) {
    if (p.half) return launch_scan_bf16(p, l, metric, emit_all, smem_configured, s);
    return scan_impl::launch_scan_fmt<false>(p, l, metric, emit_all, smem_configured, s);
}

}  // namespace otters
