"""CPU oracle for TheoremSearch's retrieval hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``theoremsearch_b200/`` imports this module; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` do, and there only as the checker / the CPU baseline — never as the product path.

Parity status: **PARTIALLY PINNED.**  The reference (uw-math-ai/TheoremSearch) has no tests, no
golden files and no recorded metrics for this path, and its arithmetic lives in two unpinned
third-party packages that are not installed here:

* ``sentence-transformers`` (unpinned; ``requirements.txt``, ``ec2/requirements.txt:4``):
  ``util.cos_sim(a, b)`` = ``torch.mm(F.normalize(a, p=2, dim=1), F.normalize(b, p=2, dim=1).T)``
  after promoting both to 2-D tensors (published behaviour of the library).
* ``pgvector`` (unpinned; ``requirements.txt:2``, ``rds_schema.sql:43``): ``a <#> b`` is the
  NEGATIVE inner product, accumulated in fp32, exact sequential scan when no index exists.

What IS pinned: the reference's own Python on either side of those calls is executed, from the
files where they lie, by ``tests/golden/make_golden.py`` (``test_app.search_theorems``,
``app_showcase_model.search_and_display``, ``streamlit_app.search_and_display`` result-row
assembly, and every metric in ``compare_embeddings.py``) with ``cos_sim`` / the SQL cursor
replaced by the restatements below; the outputs are committed under ``tests/golden/`` and
``tests/test_oracle.py`` checks this module against them.  ``torch.mm`` / ``torch.sort`` /
``numpy.argsort`` are the same libraries the reference calls.

Every function cites the reference file:line it restates.
"""
from __future__ import annotations

import math
from typing import Iterable, Sequence

import numpy as np
import torch

EPS = 1e-12  # torch.nn.functional.normalize default eps, used by util.cos_sim


# --------------------------------------------------------------------------------------
# a1/a2: L2 normalisation  (streamlit_app.py:173 normalize_embeddings=True;
#        ec2/generate_embeddings/embeddings.py:27,35; the F.normalize inside util.cos_sim)
# --------------------------------------------------------------------------------------
def normalize(x: np.ndarray | torch.Tensor, eps: float = EPS) -> torch.Tensor:
    """``F.normalize(x, p=2, dim=1, eps)``: x / max(||x||_2, eps), fp32 in, fp32 out."""
    t = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x)
    t = t.to(torch.float32)
    if t.dim() == 1:
        t = t.unsqueeze(0)
    return torch.nn.functional.normalize(t, p=2, dim=1, eps=eps)


def normalize_f64(x: np.ndarray, eps: float = EPS) -> np.ndarray:
    """The normalisation as K1 defines it bit-for-bit: ||x|| accumulated in fp64, rounded to
    fp32, clamped at eps, then an fp32 division per element.  Differs from ``normalize`` by
    at most one fp32 ulp of the norm (torch sums in fp32, in its own order)."""
    x = np.asarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[None, :]
    out = np.empty_like(x)
    for lo in range(0, x.shape[0], 65536):  # blocked: keeps the fp64 temporaries small
        blk = x[lo: lo + 65536]
        nrm = np.sqrt(np.sum(blk.astype(np.float64) ** 2, axis=1)).astype(np.float32)
        nrm = np.maximum(nrm, np.float32(eps))
        out[lo: lo + 65536] = blk / nrm[:, None]
    return out


# --------------------------------------------------------------------------------------
# quantisation of the stored corpus (K1's cast): bf16 round-to-nearest-even
# --------------------------------------------------------------------------------------
def to_bf16(x: np.ndarray | torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 with round-to-nearest-even (what ``__float2bfloat16_rn`` does)."""
    return torch.as_tensor(np.asarray(x, dtype=np.float32) if not isinstance(x, torch.Tensor)
                           else x).to(torch.float32).to(torch.bfloat16)


def bf16_round(x: np.ndarray) -> np.ndarray:
    """fp32 array rounded through bf16 and back (the values the index actually holds)."""
    return to_bf16(x).to(torch.float32).numpy()


def quantize_fp8_e4m3(x: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Per-row scaled e4m3 as the IVF lists store it: scale = max|x| / 448 (1 if the row is
    zero), q = e4m3_rn(x / scale).  Returns (dequantised fp32 rows, scales)."""
    x = np.asarray(x, dtype=np.float32)
    amax = np.abs(x).max(axis=1)
    scale = np.where(amax > 0, amax / np.float32(448.0), np.float32(1.0)).astype(np.float32)
    q = torch.from_numpy(x / scale[:, None]).to(torch.float8_e4m3fn).to(torch.float32).numpy()
    return q * scale[:, None], scale


# --------------------------------------------------------------------------------------
# a4/a7: cosine scores
# --------------------------------------------------------------------------------------
def cos_sim(a, b) -> torch.Tensor:
    """``sentence_transformers.util.cos_sim`` restated (test_app.py:76,
    app_showcase_model.py:93, compare_embeddings.py:61): promote to 2-D float tensors,
    normalise both, ``torch.mm(a_n, b_n.T)``.  Returns [Qa, Nb] fp32."""
    a = torch.as_tensor(np.asarray(a)) if not isinstance(a, torch.Tensor) else a
    b = torch.as_tensor(np.asarray(b)) if not isinstance(b, torch.Tensor) else b
    if a.dim() == 1:
        a = a.unsqueeze(0)
    if b.dim() == 1:
        b = b.unsqueeze(0)
    a_n = torch.nn.functional.normalize(a.to(torch.float32), p=2, dim=1)
    b_n = torch.nn.functional.normalize(b.to(torch.float32), p=2, dim=1)
    return torch.mm(a_n, b_n.transpose(0, 1))


def neg_inner_product(q: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """pgvector ``rows <#> q`` (streamlit_app.py:275,281): -(<row, q>), fp32 accumulate."""
    return -(np.asarray(rows, np.float32) @ np.asarray(q, np.float32)).astype(np.float32)


def scores_f64(q: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """Exact-arithmetic reference scores: fp64 dot products of the SAME (already quantised /
    already normalised) inputs the kernel sees.  q: [D] or [Q, D]; rows: [N, D]."""
    q = np.asarray(q, dtype=np.float64)
    rows = np.asarray(rows, dtype=np.float64)
    return q @ rows.T


# --------------------------------------------------------------------------------------
# a5/a8: ranking.  The reference's tie order is unspecified (np.argsort default quicksort /
# torch.topk); BASELINE.json fixes it: score descending, then LOWER id first.
# --------------------------------------------------------------------------------------
def rank_desc(scores: np.ndarray, k: int | None = None, ids: np.ndarray | None = None) -> np.ndarray:
    """Positions of the top-k entries of a 1-D score vector under (score desc, id asc).
    Restates ``np.argsort(-cosine_scores)[:5]`` (test_app.py:77) with the tie rule added."""
    s = np.asarray(scores)
    n = s.shape[0]
    tie = np.arange(n) if ids is None else np.asarray(ids)
    order = np.lexsort((tie, -s))  # last key is primary
    return order if k is None else order[: min(k, n)]


def topk(scores: np.ndarray, k: int, ids: np.ndarray | None = None) -> tuple[np.ndarray, np.ndarray]:
    """Batched top-k of a [Q, N] (or [N]) score matrix -> (scores [Q, k'], positions [Q, k'])
    with k' = min(k, N).  Restates ``np.argsort(-sim_matrix, axis=1)[:, :k]``
    (compare_embeddings.py:105) and ``torch.topk(scores, k=min(200, N), sorted=True)``
    (app_showcase_model.py:96)."""
    s = np.asarray(scores)
    squeeze = s.ndim == 1
    if squeeze:
        s = s[None, :]
    kk = min(k, s.shape[1])
    idx = np.stack([rank_desc(row, kk, ids) for row in s]) if s.shape[0] else np.zeros((0, kk), np.int64)
    val = np.take_along_axis(s, idx, axis=1) if s.shape[0] else np.zeros((0, kk), s.dtype)
    return (val[0], idx[0]) if squeeze else (val, idx)


def exact_search(queries: np.ndarray, rows: np.ndarray, k: int, ids: np.ndarray | None = None,
                 allow: np.ndarray | None = None) -> tuple[np.ndarray, np.ndarray]:
    """The parity oracle proper: fp64 scores of already-prepared inputs, top-k under
    (score desc, id asc), optional boolean ``allow`` row filter (SQL WHERE before LIMIT,
    streamlit_app.py:280-282).  Returns (scores fp64 [Q, k], ids int64 [Q, k]) padded with
    (-inf, -1) when fewer than k rows are eligible."""
    q = np.asarray(queries, dtype=np.float64)
    if q.ndim == 1:
        q = q[None, :]
    rows = np.asarray(rows)
    n = rows.shape[0]
    rid = np.arange(n, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
    out_s = np.full((q.shape[0], k), -np.inf, dtype=np.float64)
    out_i = np.full((q.shape[0], k), -1, dtype=np.int64)
    if n == 0:
        return out_s, out_i
    s = scores_f64(q, rows)
    if allow is not None:
        s = np.where(np.asarray(allow, bool)[None, :], s, -np.inf)
    n_ok = n if allow is None else int(np.count_nonzero(allow))
    for qi in range(q.shape[0]):
        # ties by ROW position first == ties by id when ids are increasing; the kernels break
        # ties on the row, so the oracle does too and maps to ids afterwards.
        order = rank_desc(s[qi], min(k, n))
        order = order[: min(k, n_ok)]
        out_s[qi, : order.size] = s[qi, order]
        out_i[qi, : order.size] = rid[order]
    return out_s, out_i


def ids_match_within_eps(got_ids: np.ndarray, ref_scores_all: np.ndarray, ref_ids: np.ndarray,
                         eps: float) -> bool:
    """fp32 re-association tolerance (SURVEY §7 hard part 1): ``got_ids`` equals the oracle's
    ``ref_ids`` except that entries whose oracle scores lie within ``eps`` of each other may
    be permuted, and the tail may swap with an outside row whose score is within ``eps`` of
    the k-th.  ``ref_scores_all``: the oracle's fp64 score of every row (1-D)."""
    got_ids = np.asarray(got_ids)
    ref_ids = np.asarray(ref_ids)
    if got_ids.shape != ref_ids.shape:
        return False
    valid = got_ids >= 0
    if not np.array_equal(valid, ref_ids >= 0):
        return False
    g = got_ids[valid]
    r = ref_ids[valid]
    if g.size == 0:
        return True
    if len(set(g.tolist())) != g.size:
        return False
    gs = ref_scores_all[g]
    rs = ref_scores_all[r]
    # position-wise: the row we returned at rank j must score within eps of the oracle's rank j
    return bool(np.all(np.abs(gs - rs) <= eps))


# --------------------------------------------------------------------------------------
# corpus writes (ec2/generate_embeddings/__main__.py:84-101 -> ec2/rds/upsert.py:29-52)
# --------------------------------------------------------------------------------------
def upsert_rows(table: dict, ids: Sequence[int], rows: np.ndarray) -> int:
    """``INSERT INTO theorem_embedding_x (slogan_id, embedding) VALUES ... ON CONFLICT (slogan_id) DO UPDATE SET
    embedding = EXCLUDED.embedding`` run by ``cur.executemany`` — one statement per row, in order: an existing
    id has its embedding replaced (its place in the table kept), a new id is appended, a repeated id ends up with
    its last embedding.  ``table``: {id: row} in insertion order (a Python dict).  Returns how many distinct
    PRE-EXISTING ids were overwritten."""
    before = set(table)
    hit = set()
    for i, r in zip(ids, rows):
        i = int(i)
        if i in before:
            hit.add(i)
        table[i] = np.asarray(r, dtype=np.float32)
    return len(hit)


def delete_rows(table: dict, ids: Sequence[int]) -> int:
    """``DELETE FROM theorem WHERE paper_id = ANY(%s)`` (ec2/parse_arxiv_papers/__main__.py:271-274) reaches the
    embedding table through ``ON DELETE CASCADE`` (rds_schema.sql:35,46,51): the rows of the named ids vanish, ids
    that are not stored match nothing.  ``table``: {id: row}.  Returns how many rows were deleted."""
    gone = 0
    for i in ids:
        if table.pop(int(i), None) is not None:
            gone += 1
    return gone


# --------------------------------------------------------------------------------------
# reference call shapes
# --------------------------------------------------------------------------------------
def search_theorems_topk(query_embedding, embeddings_db, k: int = 5):
    """test_app.py:75-77 — ``cosine_scores = util.cos_sim(q, db)[0]``;
    ``top = np.argsort(-cosine_scores.cpu())[:5]``.  Returns (positions [k], scores [k])."""
    cosine_scores = cos_sim(query_embedding, embeddings_db)[0].numpy()
    top = rank_desc(cosine_scores, k)
    return top, cosine_scores[top]


def reference_single_query_verbatim(query_embedding, embeddings_db, k: int = 5):
    """The reference expression with nothing added (test_app.py:75-77): ``np.argsort`` of a
    torch tensor dispatches to ``Tensor.argsort`` (unstable), so tie order is whatever torch
    gives.  This is the form the CPU baseline TIMES; parity uses ``search_theorems_topk``."""
    cosine_scores = cos_sim(query_embedding, embeddings_db)[0]
    top_results_indices = np.argsort(-cosine_scores.cpu())[:k]
    return top_results_indices, cosine_scores


def showcase_pool(query_embedding, embeddings_db, pool: int = 200):
    """app_showcase_model.py:92-96 — ``torch.topk(cos_sim(q, db)[0], k=min(200, N), sorted=True)``."""
    cosine_scores = cos_sim(query_embedding, embeddings_db)[0].numpy()
    top = rank_desc(cosine_scores, min(pool, cosine_scores.shape[0]))
    return top, cosine_scores[top]


def batched_ranking(q_emb, s_emb, k: int | None = None):
    """compare_embeddings.py:58-61,105 — ``sim = util.cos_sim(q_emb, s_emb).cpu().numpy()``;
    ``np.argsort(-sim, axis=1)``.  Returns (sim [Q, N], ranked positions [Q, k or N])."""
    sim = cos_sim(q_emb, s_emb).numpy()
    ranked = np.stack([rank_desc(row, k) for row in sim]) if sim.shape[0] else np.zeros((0, 0), np.int64)
    return sim, ranked


def pgvector_search(query_vec, stored_rows, k: int, allow=None):
    """streamlit_app.py:253-286 — rows were normalised when written
    (ec2/generate_embeddings/embeddings.py:27,35), the query by ``normalize_embeddings=True``
    (:173); ``ORDER BY e.embedding <#> q ASC LIMIT k``; ``similarity = 1.0 - (<#>)`` (:275),
    i.e. 1 + <e, q>.  Returns (positions [k'], similarity float64 [k'])."""
    d = neg_inner_product(query_vec, stored_rows).astype(np.float64)
    n = d.shape[0]
    eligible = np.arange(n) if allow is None else np.nonzero(np.asarray(allow, bool))[0]
    order = eligible[np.lexsort((eligible, d[eligible]))][:k]  # ASC distance, then lower row
    return order, 1.0 - d[order]


def citation_rerank(similarity: Sequence[float], citations: Sequence, weight: float, k: int):
    """streamlit_app.py:316-364 — candidates are the top ``max(50, 10*k)`` by similarity;
    ``weighted = similarity + w * ln(citations)`` when citations is not NULL and > 0, else
    ``similarity``; ``ORDER BY weighted DESC, similarity DESC LIMIT k``.  Input order (the
    candidate pool's similarity order) breaks remaining ties.  Returns (pool positions [k'],
    weighted scores [k'])."""
    sim = np.asarray(similarity, dtype=np.float64)
    w = np.array([sim[i] + weight * (math.log(float(c)) if (c is not None and c > 0) else 0.0)
                  for i, c in enumerate(citations)], dtype=np.float64)
    order = np.lexsort((np.arange(sim.size), -sim, -w))[:k]
    return order, w[order]


def pool_size(top_k: int) -> int:
    """streamlit_app.py:317 — ``max(50, int(top_k) * 10)``."""
    return max(50, int(top_k) * 10)


# --------------------------------------------------------------------------------------
# K5: sharded merge, and the packed key the GPUs exchange
# --------------------------------------------------------------------------------------
def pack_key(score: float, row: int) -> int:
    """(orderable_u32(fp32 score) << 32) | (0xFFFFFFFF - row): unsigned max == score desc,
    row asc.  -0.0 is canonicalised to +0.0, NaN to the lowest key above 'empty' (0)."""
    f = np.float32(score)
    if np.isnan(f):
        hi = 1
    else:
        f = np.float32(f + np.float32(0.0))
        u = int(np.frombuffer(np.float32(f).tobytes(), dtype=np.uint32)[0])
        hi = (u ^ 0xFFFFFFFF) if (u & 0x80000000) else (u | 0x80000000)
        hi = max(hi, 1)
    return (hi << 32) | (0xFFFFFFFF - int(row))


def unpack_key(key: int) -> tuple[float, int]:
    hi = (key >> 32) & 0xFFFFFFFF
    row = 0xFFFFFFFF - (key & 0xFFFFFFFF)
    u = (hi & 0x7FFFFFFF) if (hi & 0x80000000) else (hi ^ 0xFFFFFFFF)
    return float(np.frombuffer(np.uint32(u).tobytes(), dtype=np.float32)[0]), row


def merge_shards(shard_scores: Iterable[np.ndarray], shard_rows: Iterable[np.ndarray],
                 shard_base: Sequence[int], k: int) -> tuple[np.ndarray, np.ndarray]:
    """Top-k of the union of per-shard top-k lists: entries are (score, local
    row); global row = shard_base[g] + local row; order (score desc, global row asc).
    Each input is [Q, k_g]; rows < 0 are padding.  Returns (scores [Q, k], global rows [Q, k])."""
    ss = [np.asarray(s, dtype=np.float64) for s in shard_scores]
    rr = [np.asarray(r, dtype=np.int64) for r in shard_rows]
    nq = ss[0].shape[0]
    out_s = np.full((nq, k), -np.inf)
    out_r = np.full((nq, k), -1, dtype=np.int64)
    for qi in range(nq):
        sc = np.concatenate([s[qi] for s in ss])
        gr = np.concatenate([np.where(r[qi] >= 0, r[qi] + int(b), -1) for r, b in zip(rr, shard_base)])
        keep = gr >= 0
        sc, gr = sc[keep], gr[keep]
        order = np.lexsort((gr, -sc))[:k]
        out_s[qi, : order.size] = sc[order]
        out_r[qi, : order.size] = gr[order]
    return out_s, out_r


def shard_bounds(n: int, world: int) -> list[tuple[int, int]]:
    """Contiguous row shards [g*N/G, (g+1)*N/G)."""
    return [((g * n) // world, ((g + 1) * n) // world) for g in range(world)]


# --------------------------------------------------------------------------------------
# K4: IVF-Flat reference (pgvector ivfflat semantics: probe the nprobe nearest lists,
# exact scores inside them)
# --------------------------------------------------------------------------------------
def ivf_assign(rows: np.ndarray, centroids: np.ndarray) -> np.ndarray:
    """Nearest centroid by inner product (spherical k-means assignment), ties -> lower list."""
    s = np.asarray(rows, np.float64) @ np.asarray(centroids, np.float64).T
    return np.argmax(s, axis=1)  # argmax returns the first maximum == lower list id


def ivf_search(query: np.ndarray, rows: np.ndarray, centroids: np.ndarray, assign: np.ndarray,
               k: int, nprobe: int) -> tuple[np.ndarray, np.ndarray]:
    """Single-query IVF-Flat: top-nprobe lists by <q, centroid>, exact top-k among their rows."""
    q = np.asarray(query, np.float64)
    cs = np.asarray(centroids, np.float64) @ q
    probe = rank_desc(cs, nprobe)
    member = np.isin(assign, probe)
    s, i = exact_search(q, rows, k, allow=member)
    return s[0], i[0]


def recall_at_k(got_ids: np.ndarray, exact_ids: np.ndarray) -> float:
    """|got ∩ exact| / |exact| averaged over queries (BASELINE.json: recall@10 vs exact)."""
    got_ids = np.atleast_2d(got_ids)
    exact_ids = np.atleast_2d(exact_ids)
    hits = 0
    tot = 0
    for g, e in zip(got_ids, exact_ids):
        e = e[e >= 0]
        hits += len(set(g.tolist()) & set(e.tolist()))
        tot += e.size
    return hits / max(tot, 1)


# --------------------------------------------------------------------------------------
# evaluation metrics from a ranking (compare_embeddings.py:95-371), restated so that they
# consume top-k ids instead of the full [Q, N] similarity matrix.
# --------------------------------------------------------------------------------------
def _correct_doc(qrels_q: dict) -> int:
    return next(d for d, v in qrels_q.items() if v == 1)  # compare_embeddings.py:111


def precision_at_k(ranked: np.ndarray, qrels: dict, k: int = 5) -> float:
    """compare_embeddings.py:95-117 (hit / k, one relevant doc per query)."""
    return float(np.mean([(1 if _correct_doc(qrels[q]) in ranked[q, :k] else 0) / k
                          for q in range(ranked.shape[0])]))


def hit_at_k(ranked: np.ndarray, qrels: dict, k: int = 5) -> float:
    """compare_embeddings.py:120-140."""
    return float(np.mean([1.0 if _correct_doc(qrels[q]) in ranked[q, :k] else 0.0
                          for q in range(ranked.shape[0])]))


def mrr_at_k(ranked: np.ndarray, qrels: dict, k: int | None = None) -> float:
    """compare_embeddings.py:143-172."""
    rrs = []
    for q in range(ranked.shape[0]):
        row = ranked[q] if k is None else ranked[q, :k]
        m = np.where(row == _correct_doc(qrels[q]))[0]
        rrs.append(1.0 / (int(m[0]) + 1) if m.size else 0.0)
    return float(np.mean(rrs))


def _dcg(rels: np.ndarray, gain: str = "exp") -> float:
    """compare_embeddings.py:195-213."""
    if rels.size == 0:
        return 0.0
    gains = np.power(2.0, rels) - 1.0 if gain == "exp" else rels
    return float(np.sum(gains / np.log2(np.arange(2, rels.size + 2))))


def ndcg_at_k(ranked: np.ndarray, qrels: dict, k: int = 10, gain: str = "exp") -> float:
    """compare_embeddings.py:216-244."""
    out = []
    for q in range(ranked.shape[0]):
        rd = qrels.get(q, {})
        rels = np.array([rd.get(d, 0.0) for d in ranked[q, :k]], dtype=float)
        ideal = np.sort(np.array(list(rd.values()), dtype=float))[::-1][:k]
        idcg = _dcg(ideal, gain)
        out.append(0.0 if idcg == 0.0 else _dcg(rels, gain) / idcg)
    return float(np.mean(out))


def _max_rel(qrels: dict) -> float:
    m = 0.0
    for rd in qrels.values():
        if rd:
            m = max(m, max(rd.values()))
    return m


def err_at_k(ranked: np.ndarray, qrels: dict, k: int = 10, max_rel: float | None = None) -> float:
    """compare_embeddings.py:257-311."""
    if max_rel is None:
        max_rel = _max_rel(qrels)
        if max_rel <= 0.0:
            return 0.0
    denom = 2.0 ** max_rel
    errs = []
    for q in range(ranked.shape[0]):
        rd = qrels.get(q)
        if not rd:
            errs.append(0.0)
            continue
        rels = np.array([rd.get(int(d), 0.0) for d in ranked[q, :k]], dtype=float)
        ps = (np.power(2.0, rels) - 1.0) / denom
        e, not_sat = 0.0, 1.0
        for i, p in enumerate(ps, start=1):
            if p > 0.0:
                e += not_sat * p / i
            not_sat *= 1.0 - p
            if p > 0.0 and not_sat <= 1e-12:
                break
        errs.append(e)
    return float(np.mean(errs)) if errs else 0.0


def q_measure_at_k(ranked: np.ndarray, qrels: dict, k: int = 10, max_rel: float | None = None) -> float:
    """compare_embeddings.py:315-371."""
    if max_rel is None:
        max_rel = _max_rel(qrels)
        if max_rel <= 0.0:
            return 0.0
    denom = 2.0 ** max_rel
    out = []
    for q in range(ranked.shape[0]):
        rd = qrels.get(q)
        if not rd:
            out.append(0.0)
            continue
        cg_star = ((np.power(2.0, np.array(list(rd.values()), dtype=float)) - 1.0) / denom).sum()
        if cg_star <= 0.0:
            out.append(0.0)
            continue
        rels = np.array([rd.get(int(d), 0.0) for d in ranked[q, :k]], dtype=float)
        gains = (np.power(2.0, rels) - 1.0) / denom
        cg = qs = 0.0
        for i, g in enumerate(gains, start=1):
            if g <= 0.0:
                continue
            cg += g
            qs += g * (cg / i)
        out.append(qs / cg_star)
    return float(np.mean(out)) if out else 0.0


# --------------------------------------------------------------------------------------
# synthetic workloads: seeded, chunked, content independent of GPU count
# --------------------------------------------------------------------------------------
CHUNK_ROWS = 1 << 20
QUERY_SEED = 1_000_000


def synthetic_rows(first: int, n: int, dim: int, seed: int = 0) -> np.ndarray:
    """Rows [first, first+n) of the synthetic corpus: chunk c (rows c*2^20 ...) is
    ``torch.randn`` from ``torch.Generator(cpu).manual_seed(seed + c)``; raw (un-normalised)."""
    out = np.empty((n, dim), dtype=np.float32)
    pos = 0
    while pos < n:
        r = first + pos
        c, off = divmod(r, CHUNK_ROWS)
        take = min(n - pos, CHUNK_ROWS - off)
        g = torch.Generator(device="cpu").manual_seed(seed + c)
        # generate the chunk prefix up to off+take so content does not depend on `first`
        blk = torch.randn((off + take, dim), generator=g, dtype=torch.float32)
        out[pos: pos + take] = blk[off:].numpy()
        pos += take
    return out


def synthetic_queries(nq: int, dim: int, seed: int = QUERY_SEED) -> np.ndarray:
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn((nq, dim), generator=g, dtype=torch.float32).numpy()
