// K2 kernel instantiations: bf16 rows up to 768 elements
#include "k2_scan_impl.cuh"

namespace ts {

int launch_scan_bf16_small(const ts_index* ix, const ScanParams& p, int nchunk, int nq, int nparts, cudaStream_t s,
        cudaEvent_t ev0, cudaEvent_t ev1) {
    if (nchunk == 1) return launch_k<2, 1>(ix, p, nq, nparts, s, ev0, ev1);
    if (nchunk == 2) return launch_k<2, 2>(ix, p, nq, nparts, s, ev0, ev1);
    if (nchunk == 3) return launch_k<2, 3>(ix, p, nq, nparts, s, ev0, ev1);
    set_error("scan: no kernel for %d chunks per row", nchunk);
    return TS_ERR_UNSUPPORTED;
}

}  // namespace ts
