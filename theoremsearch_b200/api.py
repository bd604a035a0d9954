"""Reference-shaped search API: the call shapes of TheoremSearch's retrieval path, backed by
``libtheoremsearch.so``.

Tensor level:
  * ``build_index``      <- the corpus tensor of ``test_app.py:129-130`` / ``torch.load`` at
                            ``app_showcase_model.py:52`` / the rows of ``theorem_embedding_qwen``.
  * ``cos_sim_topk``     <- ``util.cos_sim(q, corpus)`` + ``argsort`` / ``torch.topk``
                            (``test_app.py:76-77``, ``app_showcase_model.py:93-96``,
                            ``compare_embeddings.py:61,105``).
  * ``search_theorems``  <- ``test_app.py:67`` (same signature; returns rows instead of rendering).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from ._lib import TheoremSearchError
from .index import TheoremIndex

_FP8_NAMES = ("fp8", "fp8_e4m3", "e4m3")
TS_IVF_MAX_CANDIDATES = 256    # include/theoremsearch.h


def build_index(embeddings, ids=None, dtype: str = "bf16", normalize: bool = True,
                device: int | str | torch.device = 0, capacity: Optional[int] = None) -> TheoremIndex:
    """Corpus [N, D] (torch tensor on any device, or numpy) -> resident index.  ``normalize=True``
    performs once, at build time, the ``F.normalize`` that ``util.cos_sim`` repeats on every
    query (test_app.py:76) — and that the production writer already applies
    (ec2/generate_embeddings/embeddings.py:27,35).  ``dtype``: "bf16" (default), "f32", or "fp8" = bf16 rows
    plus an e4m3 scan copy: ``cos_sim_topk`` then scans half the bytes and re-scores the best
    ``max(128, 2k)`` candidates exactly."""
    n, d = int(embeddings.shape[0]), int(embeddings.shape[1])
    fp8 = dtype in _FP8_NAMES
    ix = TheoremIndex(d, capacity if capacity is not None else max(n, 1), dtype="bf16" if fp8 else dtype,
                      device=device)
    if n:
        if isinstance(embeddings, torch.Tensor) and embeddings.is_cuda and embeddings.device != ix.device:
            embeddings = embeddings.to(ix.device)
        ix.add(embeddings, ids=ids, normalize=normalize)
    if fp8:
        # ``dtype='fp8'`` (BASELINE "optionally fp8-e4m3, rescored in fp32"): the scan reads an
        # e4m3 copy (1 byte / element + a per-row scale); the bf16 rows stay resident for the exact re-score.
        if n == 0:
            raise TheoremSearchError(-1, "build_index(dtype='fp8') needs the rows up front (the e4m3 copy is built once)")
        ix.build_fp8_shadow()
        ix.scan_dtype = "fp8"
    return ix


def cos_sim_topk(queries, corpus_index: TheoremIndex, k: int, normalize_queries: bool = True,
                 allow_mask: Optional[torch.Tensor] = None):
    """(scores float32 [Q, k], ids int64 [Q, k]) sorted by score desc then id asc.  A 1-D query
    returns 1-D rows so ``idx.item()`` / ``scores[i].item()`` consumers (test_app.py:85-88) work."""
    one_d = (queries.ndim if hasattr(queries, "ndim") else np.asarray(queries).ndim) == 1
    fp8 = getattr(corpus_index, "scan_dtype", None) == "fp8" and 2 * k <= TS_IVF_MAX_CANDIDATES
    if fp8 and not corpus_index.fp8_scan_ready:
        # rows were added since the e4m3 copy was made: refresh it (one pass over the corpus) unless real IVF
        # lists have taken its place, in which case the exact bf16 scan answers
        if corpus_index.nlist <= 1:
            corpus_index.build_fp8_shadow()
        else:
            fp8 = False
    if fp8:
        scores, ids = corpus_index.search_fp8(queries, k, rescore_k=min(TS_IVF_MAX_CANDIDATES, max(128, 2 * k)),
                                              normalize=normalize_queries, allow_mask=allow_mask)
    else:
        scores, ids = corpus_index.search(queries, k, normalize=normalize_queries, allow_mask=allow_mask)
    return (scores[0], ids[0]) if one_d else (scores, ids)


def search_theorems(query, model, theorems_data, embeddings_db, k: int = 5):
    """``test_app.py:67-88`` with the same arguments: encode the query with the caller's model,
    score against the corpus, take the top 5.  ``embeddings_db`` is a ``TheoremIndex`` (or a
    raw [N, D] tensor, indexed on the fly).  Returns one dict per hit —
    ``{"rank", "index", "similarity", "theorem"}`` — where the reference rendered an expander."""
    if not query:
        return []
    index = embeddings_db if isinstance(embeddings_db, TheoremIndex) else build_index(embeddings_db)
    query_embedding = model.encode(query, convert_to_tensor=True)  # test_app.py:75
    k = min(k, len(index))
    if k == 0:
        return []
    scores, ids = cos_sim_topk(query_embedding.reshape(-1), index, k)
    scores = scores.cpu()
    ids = ids.cpu()
    out = []
    for rank in range(k):
        idx = int(ids[rank].item())
        if idx < 0:
            break
        out.append({"rank": rank + 1, "index": idx, "similarity": float(scores[rank].item()),
                    "theorem": theorems_data[idx]})
    return out
