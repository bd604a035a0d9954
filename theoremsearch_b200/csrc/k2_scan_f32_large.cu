// K2 kernel instantiations: fp32 rows of 769..2048 elements
#include "k2_scan_impl.cuh"

namespace ts {

int launch_scan_f32_large(const ts_index* ix, const ScanParams& p, int nchunk, int nq, int nparts, cudaStream_t s,
        cudaEvent_t ev0, cudaEvent_t ev1) {
    if (nchunk <= 8) return launch_k<4, 8>(ix, p, nq, nparts, s, ev0, ev1);
    if (nchunk <= 16) return launch_k<4, 16>(ix, p, nq, nparts, s, ev0, ev1);
    set_error("scan: no kernel for %d chunks per row", nchunk);
    return TS_ERR_UNSUPPORTED;
}

}  // namespace ts
