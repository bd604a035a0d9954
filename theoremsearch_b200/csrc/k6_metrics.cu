// K6 — retrieval-evaluation metrics straight from the batched top-k, on the device.
//
// Reference: compare_embeddings.py:55-92 (`evaluate_retrieval`) computes six metrics — precision_at_k :95,
// hit_at_k :120, mrr_at_k :143, ndcg_at_k :216, err_at_k :257, q_measure_at_k :315 — each of which re-sorts the
// full [Q, N] similarity matrix on the host and then walks a {doc: relevance} dict per query per rank.
// Here the ranking never leaves the GPU: K3 writes ids[Q, k]; this kernel joins every ranked id with the
// query's relevance judgements (binary search in a per-query sorted table built once by ts_eval_create) and
// accumulates all six metrics in ONE left-to-right pass over the ranks, in fp64, in the reference's order of
// operations; a second kernel averages over the queries in a fixed order. 6 doubles come back.
//
// One thread per query: the walk is k (<= 64) dependent steps of a few flops each, the batch supplies the
// parallelism. Bytes: 8·k per query of ranking + ~log2(judgements) probes per rank — negligible next to the search.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "ts_common.cuh"

struct ts_eval {
    int device = 0;
    int nq = 0;
    int64_t nnz = 0;
    int ideal_width = 0;         // judged relevances kept per query for the ideal ranking (sorted descending)
    bool ideal_complete = false; // no query has more judgements than ideal_width: nDCG@k is available for every k
    double max_rel = 0.0;        // largest relevance judged anywhere (0 when nothing is judged)
    int64_t missing_correct = -1;   // first query without a document of relevance exactly 1, or -1
    int64_t* offsets = nullptr;  // [nq + 1]
    int64_t* docs = nullptr;     // [nnz] ascending within each query
    double* rels = nullptr;      // [nnz] relevance of docs[i]
    int64_t* correct = nullptr;  // [nq] first judged doc (judgement order) of relevance exactly 1, -1 if none
    double* ideal = nullptr;     // [nq, ideal_width] zero padded
    double* total_gain = nullptr;   // [nq] sum over the judged docs of 2^rel - 1
    double* per_query = nullptr;    // [nq, 6] scratch when the caller passes none
};

namespace ts {

enum { M_PRECISION = 0, M_HIT = 1, M_MRR = 2, M_NDCG = 3, M_ERR = 4, M_QMEASURE = 5, M_COUNT = 6 };

struct EvalParams {
    const int64_t* ranked;
    int64_t stride;
    int width;
    int nq;
    int k[M_COUNT];
    int gain_exp;
    double scale;          // 2^max_rel; <= 0: nothing is relevant anywhere, ERR and Q-measure are 0
    const int64_t* offsets;
    const int64_t* docs;
    const double* rels;
    const int64_t* correct;
    const double* ideal;
    int ideal_width;
    const double* total_gain;
    double* per_query;
};

__device__ __forceinline__ double judged_relevance(const int64_t* docs, const double* rels, int64_t lo, int64_t hi, int64_t doc) {
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        const int64_t d = docs[mid];
        if (d == doc) return rels[mid];
        if (d < doc) lo = mid + 1;
        else hi = mid;
    }
    return 0.0;   // unjudged documents have relevance 0 (compare_embeddings.py:229,283,343: dict.get(doc, 0.0))
}

__global__ void __launch_bounds__(128) eval_metrics_kernel(const EvalParams p) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= p.nq) return;
    const int64_t lo = p.offsets[q], hi = p.offsets[q + 1];
    const bool judged = hi > lo;
    const int64_t correct = p.correct[q];
    const int64_t* row = p.ranked + (int64_t)q * p.stride;
    int kmax = 0;
    for (int m = 0; m < M_COUNT; ++m) kmax = max(kmax, p.k[m]);
    kmax = min(kmax, p.width);

    int pos = 0;                       // 1-based rank among the valid (non-padding) entries
    int found[3] = {0, 0, 0};          // rank of the correct doc inside the cut of precision / hit / mrr
    double dcg = 0.0;
    double err = 0.0, not_satisfied = 1.0;
    bool walking = true;               // ERR: the cascade ends once the user has almost surely stopped
    double cum_gain = 0.0, blended = 0.0;
    for (int i = 0; i < kmax; ++i) {
        const int64_t doc = row[i];
        if (doc < 0) continue;         // padding (k > corpus rows)
        ++pos;
        const double rel = judged ? judged_relevance(p.docs, p.rels, lo, hi, doc) : 0.0;
        if (doc == correct) {
#pragma unroll
            for (int m = 0; m < 3; ++m)
                if (i < p.k[m] && found[m] == 0) found[m] = pos;
        }
        const double exp_gain = exp2(rel) - 1.0;
        if (i < p.k[M_NDCG]) dcg += (p.gain_exp ? exp_gain : rel) / log2((double)pos + 1.0);
        if (p.scale > 0.0) {
            const double g = exp_gain / p.scale;
            if (i < p.k[M_ERR] && walking) {
                if (g > 0.0) err += not_satisfied * g / (double)pos;
                not_satisfied *= 1.0 - g;
                if (g > 0.0 && not_satisfied <= 1e-12) walking = false;
            }
            if (i < p.k[M_QMEASURE] && g > 0.0) {
                cum_gain += g;
                blended += g * (cum_gain / (double)pos);
            }
        }
    }
    double idcg = 0.0;
    const int ideal_n = min(p.k[M_NDCG], p.ideal_width);
    for (int j = 0; j < ideal_n; ++j) {
        const double rel = p.ideal[(int64_t)q * p.ideal_width + j];
        idcg += (p.gain_exp ? exp2(rel) - 1.0 : rel) / log2((double)j + 2.0);
    }
    double* out = p.per_query + (int64_t)q * M_COUNT;
    out[M_PRECISION] = (found[0] > 0 ? 1.0 : 0.0) / (double)p.k[M_PRECISION];
    out[M_HIT] = found[1] > 0 ? 1.0 : 0.0;
    out[M_MRR] = found[2] > 0 ? 1.0 / (double)found[2] : 0.0;
    out[M_NDCG] = idcg == 0.0 ? 0.0 : dcg / idcg;
    out[M_ERR] = (judged && p.scale > 0.0) ? err : 0.0;
    const double total = p.scale > 0.0 ? p.total_gain[q] / p.scale : 0.0;
    out[M_QMEASURE] = (judged && total > 0.0) ? blended / total : 0.0;
}

// mean over the queries of each metric: thread t adds queries t, t+256, ... in order, then a fixed tree
__global__ void __launch_bounds__(256) eval_mean_kernel(const double* __restrict__ per_query, int nq, double* __restrict__ means) {
    __shared__ double part[256];
    for (int m = 0; m < M_COUNT; ++m) {
        double s = 0.0;
        for (int q = threadIdx.x; q < nq; q += 256) s += per_query[(int64_t)q * M_COUNT + m];
        part[threadIdx.x] = s;
        __syncthreads();
        for (int w = 128; w > 0; w >>= 1) {
            if ((int)threadIdx.x < w) part[threadIdx.x] += part[threadIdx.x + w];
            __syncthreads();
        }
        if (threadIdx.x == 0) means[m] = part[0] / (double)nq;
        __syncthreads();
    }
}

template <typename T>
static int upload(T** dst, const std::vector<T>& src) {
    TS_CHECK_CUDA(cudaMalloc(dst, std::max<size_t>(src.size(), 1) * sizeof(T)));
    if (!src.empty()) TS_CHECK_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return TS_OK;
}

}  // namespace ts

using namespace ts;

extern "C" {

void ts_eval_destroy(ts_eval* ev) {
    if (!ev) return;
    DeviceGuard g(ev->device);
    cudaFree(ev->offsets);
    cudaFree(ev->docs);
    cudaFree(ev->rels);
    cudaFree(ev->correct);
    cudaFree(ev->ideal);
    cudaFree(ev->total_gain);
    cudaFree(ev->per_query);
    delete ev;
}

int ts_eval_create(ts_eval** out, int device, int nq, const int64_t* offsets, const int64_t* docs, const double* rels,
                   int max_k) {
    TS_REQUIRE(out != nullptr, TS_ERR_BAD_ARG, "eval_create: out is NULL");
    *out = nullptr;
    TS_REQUIRE(nq >= 1 && offsets != nullptr, TS_ERR_BAD_ARG, "eval_create: nq=%d / offsets", nq);
    TS_REQUIRE(max_k >= 1 && max_k <= 64, TS_ERR_BAD_ARG, "eval_create: max_k=%d outside [1, 64]", max_k);
    TS_REQUIRE(offsets[0] == 0, TS_ERR_BAD_ARG, "eval_create: offsets[0] must be 0");
    for (int q = 0; q < nq; ++q)
        TS_REQUIRE(offsets[q + 1] >= offsets[q], TS_ERR_BAD_ARG, "eval_create: offsets decrease at query %d", q);
    const int64_t nnz = offsets[nq];
    TS_REQUIRE(nnz == 0 || (docs != nullptr && rels != nullptr), TS_ERR_BAD_ARG, "eval_create: docs / rels is NULL");
    DeviceGuard g(device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "eval_create: cannot select CUDA device %d (no CPU fallback)", device);

    // per query: the lookup table sorted by doc, the correct doc, the ideal ranking's relevances, the total gain
    int64_t widest = 0;
    for (int q = 0; q < nq; ++q) widest = std::max(widest, offsets[q + 1] - offsets[q]);
    const int ideal_width = (int)std::max<int64_t>(1, std::min<int64_t>(widest, max_k));
    std::vector<int64_t> h_docs((size_t)nnz), h_correct((size_t)nq, -1), order;
    std::vector<double> h_rels((size_t)nnz), h_ideal((size_t)nq * ideal_width, 0.0), h_total((size_t)nq, 0.0), sorted_rels;
    double max_rel = 0.0;
    int64_t missing = -1;
    for (int q = 0; q < nq; ++q) {
        const int64_t lo = offsets[q], n = offsets[q + 1] - lo;
        order.resize((size_t)n);
        std::iota(order.begin(), order.end(), (int64_t)0);
        std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return docs[lo + a] < docs[lo + b]; });
        double total = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            const double r = rels[lo + j];
            TS_REQUIRE(docs[lo + j] >= 0 && std::isfinite(r), TS_ERR_BAD_ARG, "eval_create: query %d judgement %lld: doc %lld, relevance %g", q,
                       (long long)j, (long long)docs[lo + j], r);
            if (h_correct[(size_t)q] < 0 && r == 1.0) h_correct[(size_t)q] = docs[lo + j];   // compare_embeddings.py:111
            total += std::exp2(r) - 1.0;
            max_rel = std::max(max_rel, r);
            h_docs[(size_t)(lo + j)] = docs[lo + order[(size_t)j]];
            h_rels[(size_t)(lo + j)] = rels[lo + order[(size_t)j]];
            TS_REQUIRE(j == 0 || h_docs[(size_t)(lo + j)] != h_docs[(size_t)(lo + j - 1)], TS_ERR_BAD_ARG,
                       "eval_create: query %d judges doc %lld twice", q, (long long)h_docs[(size_t)(lo + j)]);
        }
        h_total[(size_t)q] = total;
        if (h_correct[(size_t)q] < 0 && missing < 0) missing = q;
        sorted_rels.assign(rels + lo, rels + lo + n);
        const size_t keep = (size_t)std::min<int64_t>(n, ideal_width);
        std::partial_sort(sorted_rels.begin(), sorted_rels.begin() + keep, sorted_rels.end(), std::greater<double>());
        std::copy(sorted_rels.begin(), sorted_rels.begin() + keep, h_ideal.begin() + (size_t)q * ideal_width);
    }

    ts_eval* ev = new ts_eval();
    ev->device = device;
    ev->nq = nq;
    ev->nnz = nnz;
    ev->ideal_width = ideal_width;
    ev->ideal_complete = widest <= (int64_t)ideal_width;
    ev->max_rel = max_rel;
    ev->missing_correct = missing;
    std::vector<int64_t> h_offsets(offsets, offsets + nq + 1);
    int rc = upload(&ev->offsets, h_offsets);
    if (!rc) rc = upload(&ev->docs, h_docs);
    if (!rc) rc = upload(&ev->rels, h_rels);
    if (!rc) rc = upload(&ev->correct, h_correct);
    if (!rc) rc = upload(&ev->ideal, h_ideal);
    if (!rc) rc = upload(&ev->total_gain, h_total);
    if (!rc && cudaMalloc(&ev->per_query, (size_t)nq * M_COUNT * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        set_error("eval_create: cudaMalloc of the per-query scratch failed");
        rc = TS_ERR_OOM;
    }
    if (rc) {
        ts_eval_destroy(ev);
        return rc;
    }
    *out = ev;
    return TS_OK;
}

int ts_eval_num_queries(const ts_eval* ev) { return ev ? ev->nq : -1; }
double ts_eval_max_relevance(const ts_eval* ev) { return ev ? ev->max_rel : -1.0; }
int64_t ts_eval_first_query_without_correct_doc(const ts_eval* ev) { return ev ? ev->missing_correct : -2; }

int ts_eval_rankings(ts_eval* ev, const int64_t* ranked, int64_t stride, int width, const int* k6, int gain_exp,
                     double max_rel, double* out_means, double* out_per_query, void* stream) {
    TS_REQUIRE(ev != nullptr && ranked != nullptr && k6 != nullptr && out_means != nullptr, TS_ERR_BAD_ARG,
               "eval_rankings: NULL argument");
    TS_REQUIRE(width >= 1 && stride >= width, TS_ERR_BAD_ARG, "eval_rankings: width=%d stride=%lld", width, (long long)stride);
    EvalParams p;
    // k <= 0: no cut, the whole ranking (mrr_at_k(k=None), compare_embeddings.py:152-160)
    for (int m = 0; m < M_COUNT; ++m) p.k[m] = k6[m] <= 0 ? width : k6[m];
    // the ideal DCG sums the k largest judged relevances of the query, whatever the ranking's width
    TS_REQUIRE(ev->ideal_complete || p.k[M_NDCG] <= ev->ideal_width, TS_ERR_UNSUPPORTED,
               "eval_rankings: nDCG@%d needs a table created with max_k >= %d (it keeps %d relevances per query)",
               p.k[M_NDCG], p.k[M_NDCG], ev->ideal_width);
    DeviceGuard g(ev->device);
    TS_REQUIRE(g.ok, TS_ERR_CUDA, "eval_rankings: cannot select CUDA device %d", ev->device);
    if (max_rel < 0.0) max_rel = ev->max_rel;          // compare_embeddings.py:272-279: the largest judged relevance
    p.ranked = ranked;
    p.stride = stride;
    p.width = width;
    p.nq = ev->nq;
    p.gain_exp = gain_exp ? 1 : 0;
    p.scale = max_rel > 0.0 ? std::exp2(max_rel) : 0.0;
    p.offsets = ev->offsets;
    p.docs = ev->docs;
    p.rels = ev->rels;
    p.correct = ev->correct;
    p.ideal = ev->ideal;
    p.ideal_width = ev->ideal_width;
    p.total_gain = ev->total_gain;
    p.per_query = out_per_query ? out_per_query : ev->per_query;
    cudaStream_t s = (cudaStream_t)stream;
    eval_metrics_kernel<<<(ev->nq + 127) / 128, 128, 0, s>>>(p);
    TS_LAUNCH_CHECK();
    eval_mean_kernel<<<1, 256, 0, s>>>(p.per_query, ev->nq, out_means);
    TS_LAUNCH_CHECK();
    return TS_OK;
}

}  // extern "C"
