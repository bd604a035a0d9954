// K2 host side: grid sizing and dispatch to the kernel instantiations, which are compiled in four
// translation units (k2_scan_{bf16,f32}_{small,large}.cu) so that the build parallelises.
#include "k2_scan_impl.cuh"

#include <atomic>

namespace ts {

int launch_scan_bf16_small(const ts_index*, const ScanParams&, int nchunk, int nq, int nparts, cudaStream_t, cudaEvent_t, cudaEvent_t);
int launch_scan_bf16_large(const ts_index*, const ScanParams&, int nchunk, int nq, int nparts, cudaStream_t, cudaEvent_t, cudaEvent_t);
int launch_scan_f32_small(const ts_index*, const ScanParams&, int nchunk, int nq, int nparts, cudaStream_t, cudaEvent_t, cudaEvent_t);
int launch_scan_f32_large(const ts_index*, const ScanParams&, int nchunk, int nq, int nparts, cudaStream_t, cudaEvent_t, cudaEvent_t);

// "scan.timeline": a ring of 8 launches x 1024 CTAs x 8 stamps (ts_debug_scan_timeline)
__device__ unsigned long long g_scan_timeline[8 * 1024 * 8];
static std::atomic<unsigned> g_timeline_launch{0};

int debug_scan_timeline(uint64_t* out_host, int launches_back, int n_ctas) {
    TS_REQUIRE(out_host != nullptr && launches_back >= 0 && launches_back < 8 && n_ctas >= 1 && n_ctas <= 1024,
               TS_ERR_BAD_ARG, "debug_scan_timeline: bad argument");
    TS_CHECK_CUDA(cudaDeviceSynchronize());
    const unsigned slot = (g_timeline_launch.load() - 1u - (unsigned)launches_back) & 7u;
    TS_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_scan_timeline, (size_t)n_ctas * 8 * sizeof(uint64_t),
                                       (size_t)slot * 1024 * 8 * sizeof(uint64_t)));
    return TS_OK;
}

int scan_nparts(const ts_index* ix) {
    int ctas = tunables().scan_ctas_per_sm < 1 ? 1 : tunables().scan_ctas_per_sm;
    return sm_count(ix->device) * ctas;
}

int launch_scan_topk(const ts_index* ix, const void* data, int data_dtype, int64_t n_rows,
                     const float* queries_f32, int nq, int k, const uint32_t* allow_mask,
                     uint64_t* part_keys, int nparts, cudaStream_t s, cudaEvent_t ev0,
                     cudaEvent_t ev1, const int* qlist, const int* qcount, const ScanFused* fused) {
    TS_REQUIRE(k >= 1 && k <= TS_MAX_K, TS_ERR_BAD_ARG, "scan: k=%d out of range [1, %d]", k, TS_MAX_K);
    TS_REQUIRE(n_rows < (int64_t)0xFFFFFFFFll, TS_ERR_UNSUPPORTED, "scan: more than 2^32-1 rows per shard");
    TS_REQUIRE(nparts == scan_nparts(ix), TS_ERR_BAD_ARG, "scan: workspace sized for another grid");
    ScanParams p;
    p.data = (const uint8_t*)data;
    p.n_rows = n_rows;
    p.dim_pad = ix->dim_pad;
    p.queries = queries_f32;
    p.k = k;
    p.mask = allow_mask;
    p.part_keys = part_keys;
    p.stages = 0;
    p.nq = nq;
    p.qlist = qlist;
    p.qcount = qcount;
    p.q_raw = nullptr;
    p.q_dtype = TS_F32;
    p.q_normalize = 0;
    p.dim = ix->dim;
    p.tickets = nullptr;
    memset(&p.fin, 0, sizeof(p.fin));
    memset(&p.xchg, 0, sizeof(p.xchg));
    p.pdl = 0;
    p.ring_gate = nullptr;
    p.ring_need = 0;
    p.q_out = nullptr;
    p.timeline = nullptr;
    if (tunables().scan_timeline) {
        unsigned long long* base = nullptr;
        TS_CHECK_CUDA(cudaGetSymbolAddress((void**)&base, g_scan_timeline));
        p.timeline = base + (size_t)(g_timeline_launch.fetch_add(1) & 7u) * 1024 * 8;
    }
    if (fused != nullptr) {
        p.pdl = fused->pdl;
        p.ring_gate = fused->ring_gate;
        p.ring_need = fused->ring_need;
        p.q_raw = fused->q_raw;
        p.q_dtype = fused->q_dtype;
        p.q_normalize = fused->q_normalize;
        p.tickets = fused->tickets;
        p.fin.nq = nq;
        p.fin.id_map = fused->id_map;
        p.fin.out_keys = fused->out_keys;
        p.fin.out_scores = fused->out_scores;
        p.fin.out_ids = fused->out_ids;
        p.fin.out_stride = k;
        p.q_out = fused->q_out;
        p.fin.done_flag = fused->done_flag;
        p.fin.done_value = fused->done_value;
        p.xchg = fused->xchg;
    }
    if (data_dtype == TS_BF16) {
        p.row_bytes = (uint32_t)ix->dim_pad * 2;
        const int nchunk = (ix->dim_pad + 255) / 256;
        if (nchunk <= 3) return launch_scan_bf16_small(ix, p, nchunk, nq, nparts, s, ev0, ev1);
        if (nchunk <= 8) return launch_scan_bf16_large(ix, p, nchunk, nq, nparts, s, ev0, ev1);
    } else if (data_dtype == TS_F32) {
        p.row_bytes = (uint32_t)ix->dim_pad * 4;
        const int nchunk = (ix->dim_pad + 127) / 128;
        if (nchunk <= 6) return launch_scan_f32_small(ix, p, nchunk, nq, nparts, s, ev0, ev1);
        if (nchunk <= 16) return launch_scan_f32_large(ix, p, nchunk, nq, nparts, s, ev0, ev1);
    }
    set_error("scan: no kernel for dtype %d dim %d", data_dtype, ix->dim);
    return TS_ERR_UNSUPPORTED;
}

}  // namespace ts
