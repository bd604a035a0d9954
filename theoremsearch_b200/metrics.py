"""Retrieval-evaluation metrics straight from the batched top-k.

The reference's notebook evaluation (``compare_embeddings.py:58-371``) materialises the full
[Q, N] cosine matrix on the host and runs ``np.argsort(-sim_matrix, axis=1)`` once inside EACH of
its six metrics, although every metric only ever reads the first k ranks.  Here ONE batched search
(K3: tcgen05 GEMM + fused top-k) produces the [Q, k] ranking and all six metrics consume it.

Function names, argument meaning and defaults follow the reference; the first argument is the
ranking (``ranked[q]`` = doc ids, best first, at least k of them; -1 = padding) instead of the
similarity matrix.  Ties: lower doc id first (BASELINE.json) where the reference's argsort order is
unspecified.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .api import build_index
from .index import TheoremIndex


def _as_ranked(ranked) -> np.ndarray:
    if isinstance(ranked, torch.Tensor):
        ranked = ranked.detach().cpu().numpy()
    r = np.asarray(ranked)
    if r.ndim != 2:
        raise ValueError(f"ranked must be [num_queries, >=k] doc ids, got shape {r.shape}")
    return r


def _correct_doc(rels_dict: dict) -> int:
    return next(d for d, v in rels_dict.items() if v == 1)  # compare_embeddings.py:111


def _cut(order: np.ndarray, k: Optional[int]) -> np.ndarray:
    order = order if k is None else order[:k]
    return order[order >= 0]


def rank_concepts(q_emb, corpus, k: int, dtype: str = "f32") -> np.ndarray:
    """``rank_concepts`` (compare_embeddings.py:47-52) truncated to the k ranks anything downstream
    reads: one batched exact search instead of Q full argsorts.  ``corpus``: a ``TheoremIndex`` or a
    raw [N, D] embedding matrix, indexed on the fly as ``util.cos_sim`` would consume it — stored fp32 by
    default so the ranking is the reference's own (evaluation sets are small); ``dtype="bf16"`` puts a
    large corpus on the tensor-core path."""
    index = corpus if isinstance(corpus, TheoremIndex) else build_index(corpus, dtype=dtype)
    k = max(1, min(int(k), len(index)))
    _, ids = index.search(q_emb, k, normalize=True)
    return ids.cpu().numpy()


def precision_at_k(ranked, qrels, k: int = 5) -> float:
    """compare_embeddings.py:95-119 — hit / k, averaged."""
    r = _as_ranked(ranked)
    vals = []
    for q in range(r.shape[0]):
        hit = 1 if _correct_doc(qrels[q]) in _cut(r[q], k) else 0
        vals.append(hit / k)
    return float(np.mean(vals))


def hit_at_k(ranked, qrels, k: int = 5) -> float:
    """compare_embeddings.py:122-141."""
    r = _as_ranked(ranked)
    return float(np.mean([1 if _correct_doc(qrels[q]) in _cut(r[q], k) else 0 for q in range(r.shape[0])]))


def mrr_at_k(ranked, qrels, k: Optional[int] = None) -> float:
    """compare_embeddings.py:143-173.  With ``k=None`` the reference walks the full ranking; here the
    walk ends at the width of ``ranked`` (reciprocal ranks below 1/width count as 0)."""
    r = _as_ranked(ranked)
    rrs = []
    for q in range(r.shape[0]):
        row = _cut(r[q], k)
        m = np.where(row == _correct_doc(qrels[q]))[0]
        rrs.append(1.0 / (int(m[0]) + 1) if m.size else 0.0)
    return float(np.mean(rrs))


def _generate_qrels(queries, slogans):
    """compare_embeddings.py:175-182: 0.5 for every slogan of the query's paper, else 0."""
    return {i: {j: 0.5 if slogans[j][1] == queries[i][1] else 0 for j in range(len(slogans))}
            for i in range(len(queries))}


def _rels(order: np.ndarray, rels_dict: dict, k: Optional[int], default: float = 0.0) -> np.ndarray:
    return np.array([rels_dict.get(int(d), default) for d in _cut(order, k)], dtype=float)


def _dcg_from_rels(rels: np.ndarray, gain: str = "exp") -> float:
    """compare_embeddings.py:196-213."""
    if rels.size == 0:
        return 0.0
    if gain == "exp":
        gains = np.power(2.0, rels) - 1.0
    elif gain == "linear":
        gains = rels
    else:
        raise ValueError(f"Unknown gain scheme: {gain}")
    return float(np.sum(gains / np.log2(np.arange(2, rels.size + 2))))


def ndcg_at_k(ranked, qrels, k: int = 10, gain: str = "exp") -> float:
    """compare_embeddings.py:216-243."""
    r = _as_ranked(ranked)
    out = []
    for q in range(r.shape[0]):
        rels_dict = qrels.get(q, {})
        dcg = _dcg_from_rels(_rels(r[q], rels_dict, k), gain)
        ideal = np.sort(np.array(list(rels_dict.values()), dtype=float))[::-1]
        if k is not None:
            ideal = ideal[:k]
        idcg = _dcg_from_rels(ideal, gain)
        out.append(0.0 if idcg == 0.0 else dcg / idcg)
    return float(np.mean(out))


def _max_rel(qrels) -> float:
    m = 0.0
    for rels_dict in qrels.values():
        if rels_dict:
            m = max(m, max(rels_dict.values()))
    return m


def err_at_k(ranked, qrels, k: int = 10, max_rel: Optional[float] = None) -> float:
    """compare_embeddings.py:257-311 (expected reciprocal rank, cascade model)."""
    r = _as_ranked(ranked)
    if max_rel is None:
        max_rel = _max_rel(qrels)
        if max_rel <= 0.0:
            return 0.0
    denom = 2.0 ** max_rel
    errs = []
    for q in range(r.shape[0]):
        rels_dict = qrels.get(q, None)
        if not rels_dict:
            errs.append(0.0)
            continue
        rels = _rels(r[q], rels_dict, k)
        if rels.size == 0:
            errs.append(0.0)
            continue
        ps = (np.power(2.0, rels) - 1.0) / denom
        err_q, not_sat = 0.0, 1.0
        for i, p in enumerate(ps, start=1):
            if p > 0.0:
                err_q += not_sat * p * (1.0 / i)
            not_sat *= (1.0 - p)
            if p > 0.0 and not_sat <= 1e-12:
                break
        errs.append(err_q)
    return float(np.mean(errs)) if errs else 0.0


def q_measure_at_k(ranked, qrels, k: int = 10, max_rel: Optional[float] = None) -> float:
    """compare_embeddings.py:315-371."""
    r = _as_ranked(ranked)
    if max_rel is None:
        max_rel = _max_rel(qrels)
        if max_rel <= 0.0:
            return 0.0
    denom = 2.0 ** max_rel
    scores = []
    for q in range(r.shape[0]):
        rels_dict = qrels.get(q, None)
        if not rels_dict:
            scores.append(0.0)
            continue
        gains_all = (np.power(2.0, np.array(list(rels_dict.values()), dtype=float)) - 1.0) / denom
        cg_star = gains_all.sum()
        if cg_star <= 0.0:
            scores.append(0.0)
            continue
        gains_k = (np.power(2.0, _rels(r[q], rels_dict, k)) - 1.0) / denom
        cg = q_sum = 0.0
        for i, g in enumerate(gains_k, start=1):
            if g <= 0.0:
                continue
            cg += g
            q_sum += g * (cg / i)
        scores.append(q_sum / cg_star)
    return float(np.mean(scores)) if scores else 0.0


def evaluate_rankings(ranked, qrels, top_k_report: int = 3) -> dict:
    """The six numbers ``evaluate_retrieval`` prints (compare_embeddings.py:69-92), from one ranking."""
    k = top_k_report
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
    caller's model, ONE batched exact search on the GPU, six metrics from its top-k.  Returns the dict
    the reference only prints."""
    s_emb = model.encode([item[0] for item in theorems], convert_to_tensor=True)
    q_emb = model.encode([item[0] for item in queries], convert_to_tensor=True)
    ranked = rank_concepts(q_emb, s_emb, max(1, top_k_report), dtype=dtype)
    res = evaluate_rankings(ranked, qrels, top_k_report)
    if verbose:
        for name, val in res.items():
            print(f"{name} | {val}")
    return res
