"""App-level search: ``streamlit_app.search_and_display`` (reference ``streamlit_app.py:165-399``)
minus the rendering, over an in-memory table instead of RDS.

The reference runs one SQL statement: paper ⨝ theorem ⨝ latest_slogan ⨝ theorem_embedding_qwen,
``WHERE <filters>``, ``ORDER BY e.embedding <#> q ASC LIMIT k`` (:253-286) — or, with a citation
weight, a ``max(50, 10k)`` candidate pool re-ranked by ``similarity + w*ln(citations)``
(:316-364).  Here the joined table is ``TheoremStore.rows`` (one row per theorem, carrying its
latest slogan — the ``DISTINCT ON (theorem_id) ... ORDER BY slogan_id DESC`` of :254-259 is
applied when the store is built), the WHERE clause becomes an allow-bitmask consumed INSIDE the
scan kernel (so LIMIT k is exact among eligible rows, as in SQL), and the ORDER BY/LIMIT is K2.
"""
from __future__ import annotations

import math
from typing import Any, Optional, Sequence

import numpy as np

from .index import pack_allow_mask

ALLOWED_TYPES = ["theorem", "lemma", "proposition", "corollary"]  # streamlit_app.py:40-42

# column order of the reference's SELECT (streamlit_app.py:261-274)
COLUMNS = ("paper_id", "title", "authors", "link", "last_updated", "summary", "journal_ref",
           "primary_category", "categories", "citations", "theorem_id", "theorem_name", "theorem_body",
           "theorem_slogan")


def _contains(haystacks: np.ndarray, needle: str) -> np.ndarray:
    """Row-wise ``needle in haystack`` (SQL ``ILIKE '%needle%'`` on pre-lowered text), vectorised."""
    if haystacks.size == 0:
        return np.zeros(0, dtype=bool)
    return np.char.find(haystacks, needle) >= 0


def infer_type(name: Optional[str]) -> str:
    """streamlit_app.py:61-68."""
    if not name:
        return "theorem"
    lower = name.lower()
    for t in ALLOWED_TYPES:
        if t in lower:
            return t
    return "theorem"


def pool_size(top_k: int) -> int:
    """streamlit_app.py:317."""
    return max(50, int(top_k) * 10)


def latest_slogan_rows(slogans: Sequence[tuple[int, int]]) -> list[int]:
    """``SELECT DISTINCT ON (theorem_id) ... ORDER BY theorem_id, slogan_id DESC``
    (streamlit_app.py:254-259): input (theorem_id, slogan_id) per embedding row, output the
    positions of the rows to keep (the highest slogan_id of each theorem), in theorem_id order."""
    best: dict[int, tuple[int, int]] = {}
    for pos, (tid, sid) in enumerate(slogans):
        if tid not in best or sid > best[tid][0]:
            best[tid] = (sid, pos)
    return [best[t][1] for t in sorted(best)]


class TheoremStore:
    """The joined paper/theorem/slogan table plus the embedding index over the same rows.

    ``rows[i]`` is a tuple in ``COLUMNS`` order and describes index row ``i``.  ``index`` needs
    ``search_host(queries, k, normalize, allow_mask)`` and ``device`` (a ``TheoremIndex``)."""

    def __init__(self, rows: Sequence[Sequence[Any]], index, ann: Optional[dict] = None):
        """``ann``: None = exact scan (what the reference's index-less table does, rds_schema.sql); or
        ``{"nprobe": 32, "rescore_k": 100}`` = the IVF-Flat path (pgvector ``ivfflat`` with
        ``SET ivfflat.probes``) on an index whose lists are built — filters still apply inside the scan."""
        self.rows = [tuple(r) for r in rows]
        self.index = index
        self.ann = dict(ann) if ann else None
        n = len(self.rows)
        col = {c: [r[j] for r in self.rows] for j, c in enumerate(COLUMNS)}
        # every filter column is encoded ONCE into numpy form so the per-query WHERE is vectorised
        # (a Python loop over 10M rows per query would cost more than the scan it gates)
        link_l = np.array([(l or "").lower() for l in col["link"]], dtype=str)
        self._has_link = np.array([l is not None for l in col["link"]], dtype=bool)
        self._is_arxiv = _contains(link_l, "arxiv.org")
        self._link_l = link_l
        self._has_title = np.array([t is not None for t in col["title"]], dtype=bool)
        self._title_l = np.array([(t or "").lower() for t in col["title"]], dtype=str)
        self._author_rows: dict[Any, list[int]] = {}                   # author -> rows (p.authors && %s)
        for i, a in enumerate(col["authors"]):
            for name in set(a or ()):
                self._author_rows.setdefault(name, []).append(i)
        cats = {}
        self._category_code = np.array([cats.setdefault(c, len(cats)) for c in col["primary_category"]],
                                       dtype=np.int64).reshape(n)
        self._category_of = cats
        self._year = np.array([(d.year if d is not None else -1) for d in col["last_updated"]], dtype=np.int64)
        self._has_journal = np.array([j is not None for j in col["journal_ref"]], dtype=bool)
        self._has_name = np.array([nm is not None for nm in col["theorem_name"]], dtype=bool)
        self._name_l = np.array([(nm or "").lower() for nm in col["theorem_name"]], dtype=str)
        self._cit_known = np.array([c is not None for c in col["citations"]], dtype=bool)
        self._cit = np.array([(c if c is not None else 0) for c in col["citations"]], dtype=np.int64)
        assert n == len(self._cit)

    def __len__(self) -> int:
        return len(self.rows)

    # ---------------------------------------------------------------------------- WHERE
    def build_allow(self, filters: dict) -> np.ndarray:
        """The WHERE clause of streamlit_app.py:175-243 as a boolean row mask (SQL three-valued
        logic: a NULL operand makes the predicate not-true)."""
        n = len(self.rows)
        allow = np.ones(n, dtype=bool)
        arxiv = self._is_arxiv & self._has_link
        not_arxiv = (~self._is_arxiv) & self._has_link
        sources = filters.get("sources") or []
        if sources:                                                            # :179-186
            m = np.zeros(n, dtype=bool)
            hit = False
            if "arXiv" in sources:
                m |= arxiv
                hit = True
            if "Stacks Project" in sources:
                m |= not_arxiv
                hit = True
            if hit:
                allow &= m
        if filters.get("authors"):                                             # :189-191  p.authors && %s
            m = np.zeros(n, dtype=bool)
            for name in set(filters["authors"]):
                m[self._author_rows.get(name, [])] = True
            allow &= m
        if filters.get("tags"):                                                # :194-196
            codes = [self._category_of[c] for c in set(filters["tags"]) if c is not None and c in self._category_of]
            allow &= np.isin(self._category_code, np.array(codes, dtype=np.int64))
        if filters.get("year_range"):                                          # :199-205
            y0, y1 = filters["year_range"]
            allow &= (arxiv & (self._year >= y0) & (self._year <= y1)) | not_arxiv
        js = filters.get("journal_status", "All")                              # :208-212
        if js == "Journal Article":
            allow &= arxiv & self._has_journal
        elif js == "Preprint Only":
            allow &= arxiv & ~self._has_journal
        pf = filters.get("paper_filter") or {"ids": set(), "titles": set()}    # :215-227
        ids = [str(i).lower() for i in pf.get("ids", ())]
        titles = [str(t).lower() for t in pf.get("titles", ())]
        if ids or titles:
            m = np.zeros(n, dtype=bool)
            for i in ids:
                m |= self._has_link & _contains(self._link_l, i)
            for t in titles:
                m |= self._has_title & _contains(self._title_l, t)
            allow &= m
        if filters.get("types"):                                               # :230-233
            m = np.zeros(n, dtype=bool)
            for t in filters["types"]:
                m |= _contains(self._name_l, str(t).lower())
            allow &= m & self._has_name
        low, high = filters["citation_range"]                                  # :236-245
        between = self._cit_known & (self._cit >= low) & (self._cit <= high)
        if filters["include_unknown_citations"]:
            allow &= between | ~self._cit_known
        else:
            allow &= between
        return allow

    # ---------------------------------------------------------------------------- rows
    def result_row(self, i: int, similarity: float, score: float) -> dict:
        """The 16-key dict of streamlit_app.py:297-314 / :379-396."""
        (paper_id, title, authors, link, last_updated, _summary, journal_ref, primary_category, _categories,
         citations, theorem_id, theorem_name, theorem_body, theorem_slogan) = self.rows[i]
        link_str = link or ""
        return {
            "paper_id": paper_id,
            "authors": authors,
            "paper_title": title,
            "paper_url": link,
            "year": last_updated.year if last_updated else None,
            "primary_category": primary_category,
            "source": "arXiv" if "arxiv.org" in link_str else "Stacks Project",
            "type": infer_type(theorem_name or ""),
            "journal_published": bool(journal_ref),
            "citations": citations,
            "theorem_id": theorem_id,
            "theorem_name": theorem_name,
            "theorem_slogan": theorem_slogan,
            "theorem_body": theorem_body,
            "similarity": float(similarity),
            "score": float(score),
        }

    # ---------------------------------------------------------------------------- search
    def _topk(self, query_vec, k: int, mask):
        if self.ann is None:
            return self.index.search_host(query_vec, k, normalize=False, allow_mask=mask)
        return self.index.ivf_search_host(query_vec, k, nprobe=int(self.ann.get("nprobe", 32)),
                                          rescore_k=max(k, int(self.ann.get("rescore_k", 100))),
                                          normalize=False, allow_mask=mask)

    def search(self, query, model, filters: dict) -> list[dict]:
        """``search_and_display(query, model, filters)`` up to (not including) rendering."""
        if not filters["sources"]:                                             # :166-168
            return []
        citation_weight = float(filters["citation_weight"])                    # :170
        if isinstance(query, np.ndarray) or hasattr(query, "detach"):
            # an already-embedded query (``query: str or ndarray``): what ``model.encode`` would have
            # returned, L2-normalised here the way ``normalize_embeddings=True`` does (x / max(||x||, 1e-12))
            query_vec = np.asarray(query.detach().cpu().numpy() if hasattr(query, "detach") else query,
                                   dtype=np.float32).reshape(-1)
            query_vec = query_vec / max(float(np.linalg.norm(query_vec.astype(np.float64))), 1e-12)
        else:
            query_vec = model.encode(query or "", normalize_embeddings=True, convert_to_numpy=True)  # :173
        query_vec = np.asarray(query_vec, dtype=np.float32).reshape(-1)
        top_k = int(filters["top_k"])
        allow = self.build_allow(filters)
        mask = None if bool(allow.all()) else pack_allow_mask(allow, self.index.device)
        n_ok = int(allow.sum())
        if n_ok == 0 or top_k <= 0:
            return []
        if citation_weight == 0.0:                                             # :252-314
            k = min(top_k, n_ok)
            scores, rows = self._topk(query_vec, k, mask)
            out = []
            for s, r in zip(scores[0], rows[0]):
                if r < 0:
                    break
                sim = 1.0 + float(s)                                           # :275  1.0 - (e <#> q)
                out.append(self.result_row(int(r), sim, sim))
            return out
        pool = min(pool_size(top_k), n_ok)                                     # :317
        scores, rows = self._topk(query_vec, pool, mask)
        cand = [(1.0 + float(s), int(r)) for s, r in zip(scores[0], rows[0]) if r >= 0]
        weighted = []
        for sim, r in cand:                                                    # :351-360
            c = self.rows[r][9]
            weighted.append(sim + citation_weight * (math.log(float(c)) if (c is not None and c > 0) else 0.0))
        order = sorted(range(len(cand)), key=lambda j: (-weighted[j], -cand[j][0], j))[:top_k]   # :362-363
        return [self.result_row(cand[j][1], cand[j][0], weighted[j]) for j in order]


def _showcase_keep(items: Sequence[dict], filters: dict) -> np.ndarray:
    """Which pool entries pass the showcase app's post-filter (app_showcase_model.py:104-121), as one boolean
    vector: every criterion is evaluated column-wise over the whole pool, the criteria are AND-ed."""
    n = len(items)
    keep = np.ones(n, dtype=bool)
    source = np.array([it["source"] for it in items], dtype=object)
    on_arxiv = source == "arXiv"
    keep &= np.isin(source, list(filters["sources"]))
    if filters["types"]:
        keep &= np.isin(np.array([it["type"].lower() for it in items], dtype=object), list(filters["types"]))
    if filters["tags"]:
        keep &= np.isin(np.array([it["primary_math_tag"] for it in items], dtype=object), list(filters["tags"]))
    if filters["authors"]:
        wanted = list(filters["authors"])
        keep &= np.fromiter((any(w in it["authors"] for w in wanted) for it in items), dtype=bool, count=n)
    lo, hi = filters["citation_range"]
    cites = np.array([it["citations"] for it in items])
    keep &= (cites >= lo) & (cites <= hi)
    if filters["year_range"]:                      # the year window binds arXiv entries only
        y0, y1 = filters["year_range"]
        year = np.array([it.get("year", 0) for it in items])
        keep &= ~on_arxiv | ((year >= y0) & (year <= y1))
    status = filters["journal_status"]
    if status in ("Journal Article", "Preprint Only"):   # so does the journal / preprint switch
        published = np.array([bool(it.get("journal_published", False)) for it in items], dtype=bool)
        keep &= ~on_arxiv | (published if status == "Journal Article" else ~published)
    return keep


def search_showcase(query, model, theorems_data, embeddings_db, filters: dict, pool: int = 200) -> list[dict]:
    """``app_showcase_model.search_and_display`` (reference :82-129) up to rendering: the top ``min(200, N)``
    rows by cosine form a pool, the pool is POST-filtered and its first ``top_k`` survivors (in rank order) are
    returned.  Kept for drop-in fidelity; ``TheoremStore.search`` pre-filters instead (exact among eligible rows)."""
    if not query or not filters["sources"]:
        return []
    query_emb = model.encode(query, convert_to_tensor=True)                    # :92
    q = np.asarray(query_emb.detach().cpu().numpy() if hasattr(query_emb, "detach") else query_emb,
                   dtype=np.float32).reshape(-1)
    scores, idxs = embeddings_db.search_host(q, min(pool, len(theorems_data)), normalize=True)   # :93-96
    valid = idxs[0] >= 0
    rows, sims = idxs[0][valid], scores[0][valid]
    if rows.size == 0:
        return []
    items = [theorems_data[int(r)] for r in rows]
    keep = _showcase_keep(items, filters)
    top_k = int(filters["top_k"])
    # the reference stops once it HAS top_k results, testing after each pool entry: with top_k < 1 that is after
    # the first entry, whatever it was
    chosen = np.flatnonzero(keep)[:top_k] if top_k >= 1 else np.flatnonzero(keep[:1])
    return [{"info": items[j], "similarity": float(sims[j])} for j in chosen]
