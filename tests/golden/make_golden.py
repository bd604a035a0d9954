#!/usr/bin/env python
"""Generate tests/golden/*.json by EXECUTING the reference's own functions.

Run once, in the build container (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

How: each reference file is parsed with ``ast``; only the named ``def``s are compiled (module
top levels are Streamlit scripts / notebook cells that need models, S3 and Postgres, so they are
never executed) and run with stand-ins for what is not installed here:

  * ``streamlit`` / ``streamlit_antd_components`` -> a recorder that logs every call;
  * ``sentence_transformers.util.cos_sim``        -> ``oracle.cos_sim`` (the library's published
    behaviour restated: F.normalize both sides, torch.mm);
  * ``model.encode``                              -> a table lookup returning fixed vectors;
  * ``get_rds_connection``                        -> a cursor whose ``fetchall`` returns rows
    scored by ``oracle.pgvector_search`` (pgvector's ``<#>`` restated) and ranked by the SQL's
    ORDER BY as restated in ``oracle.citation_rerank``.

So the goldens pin the reference's OWN logic on both sides of the third-party calls — argsort /
topk / slicing / ``.item()`` consumption, the post-filter loop, the 16-key result rows, the
candidate-pool size and parameter order, and all six evaluation metrics — and the third-party
arithmetic enters only through the restatements named above (oracle header: "partially pinned").
No reference source is copied into this repository; only outputs are stored.
"""
from __future__ import annotations

import ast
import contextlib
import datetime
import json
import os
import re
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

REF = os.environ.get("TS_REFERENCE", "/root/reference")


# ----------------------------------------------------------------------------- helpers
def load_defs(path: str, names: list[str], glb: dict) -> dict:
    """Compile only the requested top-level function definitions of a reference file."""
    with open(path, encoding="utf-8") as f:
        tree = ast.parse(f.read(), filename=path)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    missing = set(names) - {n.name for n in wanted}
    assert not missing, f"{path}: missing {missing}"
    for n in wanted:
        n.decorator_list = []  # @st.cache_* decorators need a live Streamlit
    mod = ast.Module(body=wanted, type_ignores=[])
    exec(compile(mod, path, "exec"), glb)
    return glb


class Recorder:
    """Stands in for the `st` / `sac` modules: records (name, args, kwargs) of every call."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def fn(*a, **k):
            self.calls.append((name, a, k))
            return contextlib.nullcontext() if name in ("expander", "sidebar", "spinner", "container") else None
        return fn

    def titles(self):
        return [a[0] for n, a, k in self.calls if n == "expander"]


class TableModel:
    def __init__(self, table):
        self.table = table

    def encode(self, text, convert_to_tensor=False, normalize_embeddings=False, convert_to_numpy=False, **kw):
        v = np.asarray(self.table[text], dtype=np.float32)
        if normalize_embeddings:
            v = oracle.normalize(v).numpy()[0]
        return torch.from_numpy(v) if convert_to_tensor else v


def util_shim():
    u = types.SimpleNamespace()
    u.cos_sim = oracle.cos_sim
    return u


def capture_local(fn, local_name, *args, **kwargs):
    """Run fn and return the value a local variable had when fn returned (non-invasive)."""
    box = {}

    def tracer(frame, event, arg):
        if frame.f_code is fn.__code__:
            def local_trace(frame, event, arg):
                if event == "return" and local_name in frame.f_locals:
                    box["v"] = frame.f_locals[local_name]
                return local_trace
            return local_trace
        return None

    sys.settrace(tracer)
    try:
        ret = fn(*args, **kwargs)
    finally:
        sys.settrace(None)
    return box.get("v"), ret


def rng_rows(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((n, d), generator=g, dtype=torch.float32).numpy()


# ----------------------------------------------------------------------------- 1. test_app.search_theorems
def golden_test_app():
    st = Recorder()
    glb = {"st": st, "util": util_shim(), "np": np, "re": re}
    load_defs(os.path.join(REF, "test_app.py"), ["search_theorems"], glb)
    n, d = 60, 32
    corpus = rng_rows(n, d, 100) * 2.0          # un-normalised, as test_app.py:130
    corpus[41] = corpus[7]                      # exact duplicate -> reference tie order is unspecified
    theorems = [{"type": ["theorem", "lemma", "proposition"][i % 3], "paper_url": f"https://arxiv.org/abs/0000.{i:05d}",
                 "global_context": "", "content": f"content {i}"} for i in range(n)]
    queries = {f"q{j}": (corpus[(11 * j) % n] + 0.3 * rng_rows(1, d, 200 + j)[0]) for j in range(6)}
    queries["q_dup"] = corpus[7] * 1.5
    model = TableModel(queries)
    out = []
    for name in queries:
        st.calls.clear()
        glb["search_theorems"](name, model, theorems, torch.from_numpy(corpus))
        hits = []
        srcs = [a[0] for n_, a, k in st.calls if n_ == "markdown" and a and str(a[0]).startswith("**Source:**")]
        for title, src in zip(st.titles(), srcs):
            m = re.match(r"\*\*Result (\d+) \| Similarity: (-?\d+\.\d+) \| Type: (\w+)\*\*", title)
            idx = int(re.search(r"0000\.(\d+)", src).group(1))
            hits.append({"rank": int(m.group(1)), "index": idx, "similarity_4dp": m.group(2), "type": m.group(3)})
        out.append({"query": name, "hits": hits})
    return {"n": n, "d": d, "corpus_seed": 100, "corpus": corpus.tolist(), "queries": {k: v.tolist() for k, v in queries.items()},
            "theorem_types": [t["type"] for t in theorems], "results": out,
            "source": "test_app.py:67-88 search_theorems executed from the reference file"}


# ----------------------------------------------------------------------------- 2. app_showcase_model.search_and_display
def golden_showcase():
    st = Recorder()
    glb = {"st": st, "util": util_shim(), "np": np, "re": re, "torch": torch,
           "clean_latex_for_display": lambda s: s}
    load_defs(os.path.join(REF, "app_showcase_model.py"), ["search_and_display"], glb)
    n, d = 400, 24
    corpus = rng_rows(n, d, 300)
    r = np.random.default_rng(5)
    sources = ["arXiv", "Stacks Project"]
    tags = ["math.AG", "math.NT", "math.CO"]
    theorems = []
    for i in range(n):
        theorems.append({
            "type": ["theorem", "lemma", "proposition", "corollary"][int(r.integers(4))],
            "primary_math_tag": tags[int(r.integers(3))],
            "authors": [f"A{int(r.integers(6))}", f"A{int(r.integers(6))}"],
            "source": sources[int(r.random() < 0.25)],
            "citations": int(r.integers(0, 200)),
            "year": int(r.integers(1995, 2025)),
            "journal_published": bool(r.random() < 0.4),
            "paper_title": f"Paper {i}", "paper_url": f"https://example.org/{i:05d}",
            "global_context": "", "content": f"content {i}",
        })
    queries = {f"s{j}": rng_rows(1, d, 400 + j)[0] for j in range(4)}
    model = TableModel(queries)
    filter_sets = [
        {"authors": [], "types": [], "tags": [], "sources": ["arXiv", "Stacks Project"], "year_range": None,
         "journal_status": "All", "citation_range": (0, 1000000), "top_k": 5},
        {"authors": ["A1"], "types": ["lemma", "theorem"], "tags": ["math.AG"], "sources": ["arXiv"],
         "year_range": (2000, 2022), "journal_status": "All", "citation_range": (10, 190), "top_k": 7},
        {"authors": [], "types": ["corollary"], "tags": [], "sources": ["Stacks Project"], "year_range": (2010, 2012),
         "journal_status": "Preprint Only", "citation_range": (0, 1000000), "top_k": 20},
        {"authors": ["A0", "A5"], "types": [], "tags": ["math.NT", "math.CO"], "sources": ["arXiv"],
         "year_range": (1991, 2025), "journal_status": "Preprint Only", "citation_range": (40, 160), "top_k": 10},
    ]
    out = []
    for qname in queries:
        for fi, filters in enumerate(filter_sets):
            st.calls.clear()
            filtered, _ = capture_local(glb["search_and_display"], "filtered_results", qname, model, theorems,
                                        torch.from_numpy(corpus), filters)
            hits = [{"index": int(f["info"]["paper_url"][-5:]), "similarity": float(f["similarity"])}
                    for f in (filtered or [])]
            out.append({"query": qname, "filters": fi, "hits": hits, "titles": st.titles()})
    return {"n": n, "d": d, "corpus": corpus.tolist(), "queries": {k: v.tolist() for k, v in queries.items()},
            "theorems": theorems, "filter_sets": filter_sets, "results": out,
            "source": "app_showcase_model.py:82-160 search_and_display executed from the reference file"}


# ----------------------------------------------------------------------------- 3. compare_embeddings metrics
def golden_metrics():
    names = ["rank_concepts", "precision_at_k", "hit_at_k", "mrr_at_k", "_generate_qrels", "_get_rels_for_query",
             "_dcg_from_rels", "ndcg_at_k", "_get_rels_sparse", "err_at_k", "q_measure_at_k"]
    glb = {"np": np}
    load_defs(os.path.join(REF, "compare_embeddings.py"), names, glb)
    nq, n, d = 14, 90, 16
    docs = rng_rows(n, d, 500)
    # query i's correct document is doc 3*i; queries are noisy copies so ranks vary
    queries = np.stack([docs[3 * i] + 1.2 * rng_rows(1, d, 600 + i)[0] for i in range(nq)])
    sim = oracle.cos_sim(queries, docs).numpy()                       # compare_embeddings.py:61
    paper_of_doc = [j // 5 for j in range(n)]
    q_list = [(f"q{i}", paper_of_doc[3 * i]) for i in range(nq)]
    s_list = [(f"s{j}", paper_of_doc[j]) for j in range(n)]
    qrels = glb["_generate_qrels"](q_list, s_list)                    # 0.5 for same paper
    for i in range(nq):
        qrels[i][3 * i] = 1                                           # the exact match (driver cell :452-453 intent)
    metrics = {}
    for k in (1, 3, 5, 10):
        metrics[str(k)] = {
            "precision": glb["precision_at_k"](sim, qrels, k=k),
            "hit": glb["hit_at_k"](sim, qrels, k=k),
            "mrr": glb["mrr_at_k"](sim, qrels, k=k),
            "ndcg": glb["ndcg_at_k"](sim, qrels, k=k),
            "err": glb["err_at_k"](sim, qrels, k=k),
            "q_measure": glb["q_measure_at_k"](sim, qrels, k=k),
        }
    ranked = np.stack(glb["rank_concepts"](sim))
    return {"nq": nq, "n": n, "d": d, "docs": docs.tolist(), "queries": queries.tolist(),
            "qrels": {str(q): {str(dd): v for dd, v in rd.items()} for q, rd in qrels.items()},
            "ranked_top10": ranked[:, :10].tolist(), "metrics": metrics,
            "source": "compare_embeddings.py:47-371 executed from the reference file"}


# ----------------------------------------------------------------------------- 4. streamlit_app.search_and_display rows
class FakeCursor:
    """Answers the two SQL shapes of streamlit_app.py:253-286 / :319-366 from an in-memory
    table, using the oracle's pgvector restatement; records the SQL and parameters."""

    def __init__(self, db, log):
        self.db, self.log, self.rows = db, log, []

    def execute(self, sql, params):
        self.log.append({"sql_has_candidates_cte": "candidates AS" in sql, "n_params": len(params),
                         "limit_literal": (re.search(r"LIMIT (\d+)\s*\)", sql) or [None, None])[1]})
        qv = np.asarray(params[0], dtype=np.float32)
        assert np.array_equal(qv, np.asarray(params[-3 if "candidates AS" in sql else -2], dtype=np.float32))
        top_k = int(params[-1])
        emb = self.db["embeddings"]
        if "candidates AS" in sql:
            weight = float(params[-2])
            pool = int(re.search(r"LIMIT (\d+)\s*\)", sql).group(1))
            order, sim = oracle.pgvector_search(qv, emb, pool)
            cits = [self.db["rows"][i][9] for i in order]
            sel, w = oracle.citation_rerank(sim, cits, weight, top_k)
            self.rows = [tuple(self.db["rows"][order[j]]) + (float(sim[j]), float(w[jj]))
                         for jj, j in enumerate(sel)]
        else:
            order, sim = oracle.pgvector_search(qv, emb, top_k)
            self.rows = [tuple(self.db["rows"][i]) + (float(s),) for i, s in zip(order, sim)]

    def fetchall(self):
        return self.rows

    def close(self):
        pass


class FakeConn:
    def __init__(self, db, log):
        self.db, self.log = db, log

    def cursor(self):
        return FakeCursor(self.db, self.log)

    def close(self):
        pass


def golden_streamlit_rows():
    st, sac = Recorder(), Recorder()
    log = []
    n, d = 120, 32
    emb = oracle.normalize(rng_rows(n, d, 700)).numpy()               # written normalised (embeddings.py:27,35)
    r = np.random.default_rng(9)
    rows = []
    for i in range(n):
        arxiv = r.random() < 0.7
        link = f"https://arxiv.org/abs/2401.{i:05d}" if arxiv else f"https://stacks.math.columbia.edu/tag/{i:04X}"
        name = ["Theorem 1.2", "Lemma 3", "Proposition 2.1 (main)", "Corollary 5", "Remark 7", None][int(r.integers(6))]
        cit = None if r.random() < 0.2 else int(r.integers(0, 500))
        rows.append((f"paper{i}", f"Title {i}", [f"Author {int(r.integers(9))}"], link,
                     datetime.datetime(int(r.integers(1999, 2025)), 5, 17) if r.random() < 0.9 else None,
                     f"summary {i}", "J. Math 1" if r.random() < 0.3 else None, "math.AG", ["math.AG"], cit,
                     1000 + i, name, f"body {i}", f"slogan {i}"))
    db = {"embeddings": emb, "rows": rows}
    glb = {"st": st, "sac": sac, "re": re, "json": json, "clean_latex_for_display": lambda s: s,
           "get_rds_connection": lambda: FakeConn(db, log), "EMBED_TABLE": "theorem_embedding_qwen",
           "ALLOWED_TYPES": ["theorem", "lemma", "proposition", "corollary"]}
    load_defs(os.path.join(REF, "streamlit_app.py"), ["search_and_display", "infer_type"], glb)
    queries = {f"p{j}": rng_rows(1, d, 800 + j)[0] * 3 for j in range(3)}
    model = TableModel(queries)
    base = {"sources": ["arXiv", "Stacks Project"], "authors": [], "tags": [], "year_range": None,
            "journal_status": "All", "types": [], "citation_range": (0, 10**9), "include_unknown_citations": True,
            "paper_filter": {"ids": set(), "titles": set()}}
    out = []
    for qname in queries:
        for top_k, w in ((5, 0.0), (3, 0.0), (5, 0.05), (20, 0.01), (2, -0.02)):
            log.clear()
            filters = dict(base, top_k=top_k, citation_weight=w)
            results, _ = capture_local(glb["search_and_display"], "results", qname, model, filters)
            out.append({"query": qname, "top_k": top_k, "citation_weight": w, "sql": list(log),
                        "results": json.loads(json.dumps(results, default=str))})
    ser_rows = [[(x.isoformat() if isinstance(x, datetime.datetime) else x) for x in row] for row in rows]
    return {"n": n, "d": d, "embeddings": emb.tolist(), "rows": ser_rows,
            "queries": {k: v.tolist() for k, v in queries.items()}, "results": out,
            "source": "streamlit_app.py:165-399 search_and_display executed from the reference file; "
                      "SQL answered by oracle.pgvector_search / oracle.citation_rerank"}


def main():
    os.makedirs(HERE, exist_ok=True)
    for name, fn in [("test_app_search_theorems", golden_test_app), ("showcase_search", golden_showcase),
                     ("compare_embeddings_metrics", golden_metrics), ("streamlit_rows", golden_streamlit_rows)]:
        data = fn()
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(data, f)
        print(name, "->", os.path.getsize(os.path.join(HERE, name + ".json")), "bytes")


if __name__ == "__main__":
    main()
