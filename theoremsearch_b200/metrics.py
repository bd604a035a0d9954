"""Retrieval-evaluation metrics straight from the batched top-k.

The reference's notebook evaluation (``compare_embeddings.py:58-371``) materialises the full
[Q, N] cosine matrix on the host and runs ``np.argsort(-sim_matrix, axis=1)`` once inside EACH of
its six metrics, although every metric only ever reads the first k ranks.  Here ONE batched search
(K3: tcgen05 GEMM + fused top-k) produces the [Q, k] ranking and all six metrics consume it — on the
GPU when the ranking is a CUDA tensor (K6 ``eval_metrics_kernel`` through ``ts_eval_rankings``: the ids
never leave the device, six doubles come back), as array arithmetic on the host when the ranking already
is a host array (e.g. read from a file).

Function names, argument meaning and defaults follow the reference; the first argument is the
ranking (``ranked[q]`` = doc ids, best first, at least k of them; -1 = padding) instead of the
similarity matrix.  Ties: lower doc id first (BASELINE.json) where the reference's argsort order is
unspecified.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from .api import build_index
from .index import TheoremIndex, _stream_ptr

METRIC_ORDER = ("precision", "hit", "mrr", "ndcg", "err", "q_measure")     # ts_eval_rankings' k6 / output order


def _as_ranked(ranked) -> np.ndarray:
    if isinstance(ranked, torch.Tensor):
        ranked = ranked.detach().cpu().numpy()
    r = np.asarray(ranked)
    if r.ndim != 2:
        raise ValueError(f"ranked must be [num_queries, >=k] doc ids, got shape {r.shape}")
    return r.astype(np.int64, copy=False)


class _Judged:
    """The [Q, k] ranking joined ONCE with the relevance judgements — every metric below is array arithmetic on
    these tables, where the reference walks a dict per query per rank inside each of its six metric loops.

    ``qrels``: {query -> {doc -> relevance}} (compare_embeddings.py:175-182). All (query, doc, relevance) triples
    are flattened into one sorted key table; the relevance of every ranked entry is a single ``searchsorted``.
    ``pos`` is each entry's 1-based rank among the VALID entries of its row (-1 padding and ranks past ``k``
    are invalid and carry relevance 0)."""

    def __init__(self, ranked, qrels, k: Optional[int]):
        r = _as_ranked(ranked)
        self.nq = r.shape[0]
        r = r if k is None else r[:, :k]
        self.docs = r
        self.valid = r >= 0
        self.pos = np.cumsum(self.valid, axis=1)
        per_query = [qrels.get(q) or {} for q in range(self.nq)]
        self.judged = np.array([bool(d) for d in per_query], dtype=bool)
        counts = np.array([len(d) for d in per_query], dtype=np.int64)
        self.owner = np.repeat(np.arange(self.nq), counts)
        self.doc_of = np.fromiter((doc for d in per_query for doc in d), dtype=np.int64, count=int(counts.sum()))
        self.rel_of = np.fromiter((v for d in per_query for v in d.values()), dtype=float, count=int(counts.sum()))
        self.starts = np.concatenate([[0], np.cumsum(counts)])
        span = int(max(self.doc_of.max(initial=0), r.max(initial=0))) + 2
        keys = self.owner * span + self.doc_of
        order = np.argsort(keys, kind="stable")
        keys, rel_sorted = keys[order], self.rel_of[order]
        want = (np.arange(self.nq)[:, None] * span + np.where(self.valid, r, span - 1)).reshape(-1)
        at = np.minimum(np.searchsorted(keys, want), max(len(keys) - 1, 0))
        found = (keys[at] == want) if len(keys) else np.zeros(want.shape, dtype=bool)
        rel = np.where(found, rel_sorted[at] if len(keys) else 0.0, 0.0).reshape(r.shape)
        self.rel = np.where(self.valid, rel, 0.0)

    def correct_doc(self) -> np.ndarray:
        """Per query, the first judged doc (in judgement order) with relevance exactly 1 (compare_embeddings.py:111)."""
        is_one = self.rel_of == 1
        first = np.full(self.nq, -1, dtype=np.int64)
        idx = np.flatnonzero(is_one)
        if idx.size:
            q_of, where = np.unique(self.owner[idx], return_index=True)
            first[q_of] = self.doc_of[idx[where]]
        if (first < 0).any():
            raise StopIteration(f"query {int(np.flatnonzero(first < 0)[0])} has no document of relevance 1")
        return first

    def found_rank(self) -> np.ndarray:
        """1-based rank of each query's correct doc inside the cut ranking, 0 where it is absent."""
        match = self.valid & (self.docs == self.correct_doc()[:, None])
        return np.where(match.any(axis=1), self.pos[np.arange(self.nq), match.argmax(axis=1)], 0)

    def ideal(self, k: Optional[int]) -> np.ndarray:
        """[Q, width] judged relevances of each query sorted descending and cut at k (zero padded)."""
        width = int((self.starts[1:] - self.starts[:-1]).max(initial=0))
        width = width if k is None else min(width, k)
        out = np.zeros((self.nq, max(width, 1)))
        order = np.lexsort((-self.rel_of, self.owner))              # by query, then relevance descending
        slot = np.arange(len(order)) - self.starts[self.owner[order]]
        keep = slot < out.shape[1]
        out[self.owner[order][keep], slot[keep]] = self.rel_of[order][keep]
        return out


def _seq_sum(terms: np.ndarray) -> np.ndarray:
    """Row sums accumulated left to right (the reference adds term by term; np.sum would add pairwise)."""
    return np.cumsum(terms, axis=1)[:, -1] if terms.shape[1] else np.zeros(terms.shape[0])


def _gain(rel: np.ndarray, gain: str) -> np.ndarray:
    if gain == "exp":
        return np.exp2(rel) - 1.0
    if gain == "linear":
        return rel
    raise ValueError(f"Unknown gain scheme: {gain}")


class JudgedTable:
    """The relevance judgements resident on the GPU (``ts_eval_create``): ``qrels`` = {query -> {doc -> relevance}}
    (compare_embeddings.py:175-182) flattened ONCE into per-query sorted lookup tables, each query's correct
    document, its ideal relevances and its total gain. ``evaluate`` then costs two small kernels per ranking."""

    def __init__(self, qrels, num_queries: int, device, max_k: int = 64):
        self.device = torch.device(device)
        per_query = [qrels.get(q) or {} for q in range(int(num_queries))]
        counts = np.array([len(d) for d in per_query], dtype=np.int64)
        offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        docs = np.fromiter((doc for d in per_query for doc in d), dtype=np.int64, count=int(counts.sum()))
        rels = np.fromiter((v for d in per_query for v in d.values()), dtype=np.float64, count=int(counts.sum()))
        self.handle = C.c_void_p()
        check(lib.ts_eval_create(C.byref(self.handle), self.device.index or 0, int(num_queries), offsets.ctypes.data,
                                 docs.ctypes.data, rels.ctypes.data, int(max_k)))
        self.num_queries = int(num_queries)

    def close(self) -> None:
        if self.handle:
            lib.ts_eval_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def max_relevance(self) -> float:
        return float(lib.ts_eval_max_relevance(self.handle))

    def evaluate(self, ranked: torch.Tensor, k: dict, gain: str = "exp", max_rel: Optional[float] = None,
                 per_query: bool = False):
        """``ranked``: CUDA int64 [Q, width] (what ``TheoremIndex.search`` returned). ``k``: {metric name -> cut},
        names from ``METRIC_ORDER``; metrics left out are not reported, ``None`` = no cut. Returns {name -> mean
        over the queries} (and the [Q, 6] per-query table with ``per_query=True``)."""
        if gain not in ("exp", "linear"):
            raise ValueError(f"Unknown gain scheme: {gain}")
        if not (isinstance(ranked, torch.Tensor) and ranked.is_cuda and ranked.dtype == torch.int64 and ranked.dim() == 2):
            raise _lib.TheoremSearchError(-1, "JudgedTable.evaluate takes the CUDA int64 [Q, k] ids of a batched search")
        if ranked.device != self.device or ranked.shape[0] != self.num_queries:
            raise _lib.TheoremSearchError(-1, f"ranking {tuple(ranked.shape)} on {ranked.device}: the table judges "
                                              f"{self.num_queries} queries on {self.device}")
        if any(m in k for m in ("precision", "hit", "mrr")):
            q = int(lib.ts_eval_first_query_without_correct_doc(self.handle))
            if q >= 0:       # the reference's `next(d for d, v in ... if v == 1)` (compare_embeddings.py:111)
                raise StopIteration(f"query {q} has no document of relevance 1")
        if ranked.stride(1) != 1:
            ranked = ranked.contiguous()
        k6 = (C.c_int * 6)(*[int(k.get(m) or 0) if m in k else 1 for m in METRIC_ORDER])
        means = torch.empty(6, dtype=torch.float64, device=self.device)
        table = torch.empty((self.num_queries, 6), dtype=torch.float64, device=self.device) if per_query else None
        with torch.cuda.device(self.device):
            check(lib.ts_eval_rankings(self.handle, ranked.data_ptr(), int(ranked.stride(0)), int(ranked.shape[1]), k6,
                                       1 if gain == "exp" else 0, -1.0 if max_rel is None else float(max_rel),
                                       means.data_ptr(), table.data_ptr() if per_query else None,
                                       _stream_ptr(self.device)))
        host = means.cpu().numpy()
        out = {m: float(host[i]) for i, m in enumerate(METRIC_ORDER) if m in k}
        return (out, table) if per_query else out


def _on_device(ranked) -> bool:
    return isinstance(ranked, torch.Tensor) and ranked.is_cuda


def _device_metric(ranked: torch.Tensor, qrels, name: str, k, **kw) -> float:
    """One metric of a CUDA ranking. ``qrels``: the reference's dict, or a ``JudgedTable`` built once for many calls."""
    if isinstance(qrels, JudgedTable):
        return qrels.evaluate(ranked, {name: k}, **kw)[name]
    table = JudgedTable(qrels, ranked.shape[0], ranked.device, max_k=min(64, max(1, int(k or ranked.shape[1]))))
    try:
        return table.evaluate(ranked, {name: k}, **kw)[name]
    finally:
        table.close()


def rank_concepts(q_emb, corpus, k: int, dtype: str = "f32", as_tensor: bool = False):
    """``rank_concepts`` (compare_embeddings.py:47-52) truncated to the k ranks anything downstream
    reads: one batched exact search instead of Q full argsorts.  ``corpus``: a ``TheoremIndex`` or a
    raw [N, D] embedding matrix, indexed on the fly as ``util.cos_sim`` would consume it — stored fp32 by
    default so the ranking is the reference's own (evaluation sets are small); ``dtype="bf16"`` puts a
    large corpus on the tensor-core path."""
    index = corpus if isinstance(corpus, TheoremIndex) else build_index(corpus, dtype=dtype)
    k = max(1, min(int(k), len(index)))
    _, ids = index.search(q_emb, k, normalize=True)
    return ids if as_tensor else ids.cpu().numpy()      # as_tensor: the ranking stays on the GPU for K6


def precision_at_k(ranked, qrels, k: int = 5) -> float:
    """compare_embeddings.py:95-119 — (correct doc among the first k) / k, averaged over queries."""
    if _on_device(ranked):
        return _device_metric(ranked, qrels, "precision", k)
    return float(np.mean((_Judged(ranked, qrels, k).found_rank() > 0) / k))


def hit_at_k(ranked, qrels, k: int = 5) -> float:
    """compare_embeddings.py:122-141."""
    if _on_device(ranked):
        return _device_metric(ranked, qrels, "hit", k)
    return float(np.mean((_Judged(ranked, qrels, k).found_rank() > 0).astype(np.int64)))


def mrr_at_k(ranked, qrels, k: Optional[int] = None) -> float:
    """compare_embeddings.py:143-173.  With ``k=None`` the reference walks the full ranking; here the
    walk ends at the width of ``ranked`` (reciprocal ranks below 1/width count as 0)."""
    if _on_device(ranked):
        return _device_metric(ranked, qrels, "mrr", k)
    rank = _Judged(ranked, qrels, k).found_rank()
    return float(np.mean(np.where(rank > 0, 1.0 / np.maximum(rank, 1), 0.0)))


def _generate_qrels(queries, slogans):
    """compare_embeddings.py:175-182: 0.5 for every slogan of the query's paper, else 0."""
    paper_of_slogan = np.array([s[1] for s in slogans], dtype=object)
    return {i: dict(enumerate(np.where(paper_of_slogan == q[1], 0.5, 0).tolist())) for i, q in enumerate(queries)}


def _discounted(gains: np.ndarray, pos: np.ndarray) -> np.ndarray:
    """sum_i gain_i / log2(rank_i + 1), rank 1-based (compare_embeddings.py:196-213)."""
    return _seq_sum(gains / np.log2(np.maximum(pos, 1) + 1.0))


def ndcg_at_k(ranked, qrels, k: int = 10, gain: str = "exp") -> float:
    """compare_embeddings.py:216-243: DCG of the ranking over the DCG of the judged relevances sorted descending."""
    if _on_device(ranked):
        return _device_metric(ranked, qrels, "ndcg", k, gain=gain)
    j = _Judged(ranked, qrels, k)
    dcg = _discounted(np.where(j.valid, _gain(j.rel, gain), 0.0), j.pos)
    ideal = j.ideal(k)
    idcg = _discounted(_gain(ideal, gain), np.arange(1, ideal.shape[1] + 1)[None, :])
    return float(np.mean(np.where(idcg == 0.0, 0.0, dcg / np.where(idcg == 0.0, 1.0, idcg))))


def _max_rel(qrels) -> float:
    return max((max(d.values()) for d in qrels.values() if d), default=0.0) or 0.0


def _scale(qrels, max_rel: Optional[float]) -> Optional[float]:
    """2^max_rel, the normaliser of the graded gains; None when nothing is relevant (the metric is then 0)."""
    if max_rel is None:
        max_rel = max(0.0, _max_rel(qrels))
        if max_rel <= 0.0:
            return None
    return 2.0 ** max_rel


def err_at_k(ranked, qrels, k: int = 10, max_rel: Optional[float] = None) -> float:
    """compare_embeddings.py:257-311 — expected reciprocal rank under the cascade model: the user stops at rank i
    with probability p_i = (2^rel_i - 1) / 2^max_rel having passed every earlier rank; the walk is abandoned once
    the probability of still reading drops to 1e-12."""
    if _on_device(ranked):
        return _device_metric(ranked, qrels, "err", k, max_rel=max_rel)
    scale = _scale(qrels, max_rel)
    if scale is None:
        return 0.0
    j = _Judged(ranked, qrels, k)
    if j.nq == 0:
        return 0.0
    stop = (np.exp2(j.rel) - 1.0) / scale
    still_reading = np.cumprod(1.0 - stop, axis=1)
    reached = np.concatenate([np.ones((j.nq, 1)), still_reading[:, :-1]], axis=1)
    gave_up = (stop > 0.0) & (still_reading <= 1e-12)
    live = (np.cumsum(gave_up, axis=1) - gave_up) == 0            # nothing before this rank ended the walk
    terms = np.where(live & (stop > 0.0), reached * stop * (1.0 / np.maximum(j.pos, 1)), 0.0)
    return float(np.mean(np.where(j.judged, _seq_sum(terms), 0.0)))


def q_measure_at_k(ranked, qrels, k: int = 10, max_rel: Optional[float] = None) -> float:
    """compare_embeddings.py:315-371 — sum over relevant ranks of gain_i * (cumulative gain_i / i), over the total
    gain of everything judged for the query."""
    if _on_device(ranked):
        return _device_metric(ranked, qrels, "q_measure", k, max_rel=max_rel)
    scale = _scale(qrels, max_rel)
    if scale is None:
        return 0.0
    j = _Judged(ranked, qrels, k)
    if j.nq == 0:
        return 0.0
    total = np.zeros(j.nq)
    np.add.at(total, j.owner, (np.exp2(j.rel_of) - 1.0) / scale)
    g = (np.exp2(j.rel) - 1.0) / scale
    g = np.where(g > 0.0, g, 0.0)
    blended = _seq_sum(g * (np.cumsum(g, axis=1) / np.maximum(j.pos, 1)))
    ok = j.judged & (total > 0.0)
    return float(np.mean(np.where(ok, blended / np.where(ok, total, 1.0), 0.0)))


def evaluate_rankings(ranked, qrels, top_k_report: int = 3) -> dict:
    """The six numbers ``evaluate_retrieval`` prints (compare_embeddings.py:69-92), from one ranking. A CUDA
    ranking is evaluated on the device in one pass (K6); ``qrels`` may then be a ``JudgedTable``."""
    k = top_k_report
    if _on_device(ranked):
        table = qrels if isinstance(qrels, JudgedTable) else JudgedTable(qrels, ranked.shape[0], ranked.device,
                                                                         max_k=min(64, max(1, k)))
        try:
            got = table.evaluate(ranked, {"precision": 1, "hit": k, "mrr": k, "ndcg": k, "err": k, "q_measure": k})
        finally:
            if table is not qrels:
                table.close()
        return {"P@1": got["precision"], f"H@{k}": got["hit"], f"MRR@{k}": got["mrr"], f"nDCG@{k}": got["ndcg"],
                f"ERR@{k}": got["err"], f"Q-measure@{k}": got["q_measure"]}
    return {
        "P@1": precision_at_k(ranked, qrels, k=1),
        f"H@{k}": hit_at_k(ranked, qrels, k=k),
        f"MRR@{k}": mrr_at_k(ranked, qrels, k=k),
        f"nDCG@{k}": ndcg_at_k(ranked, qrels, k=k),
        f"ERR@{k}": err_at_k(ranked, qrels, k=k),
        f"Q-measure@{k}": q_measure_at_k(ranked, qrels, k=k),
    }


def evaluate_retrieval(model, theorems, queries, qrels, top_k_report: int = 3, verbose: bool = True,
                       dtype: str = "f32") -> dict:
    """``evaluate_retrieval`` (compare_embeddings.py:55-92), same arguments: encode both sides with the
    caller's model, ONE batched exact search on the GPU, six metrics from its top-k computed on the GPU as well
    (the ranking is never copied to the host).  Returns the dict the reference only prints."""
    s_emb = model.encode([item[0] for item in theorems], convert_to_tensor=True)
    q_emb = model.encode([item[0] for item in queries], convert_to_tensor=True)
    ranked = rank_concepts(q_emb, s_emb, max(1, top_k_report), dtype=dtype, as_tensor=True)
    res = evaluate_rankings(ranked, qrels, top_k_report)
    if verbose:
        for name, val in res.items():
            print(f"{name} | {val}")
    return res
