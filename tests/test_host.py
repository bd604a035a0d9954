"""Host-side logic of the drop-in boundary (filters -> mask, result rows, citation re-rank, showcase
post-filter) against goldens produced by the reference's own functions. The scorer is the test-only
OracleIndex (no GPU here); tests/test_gpu_store.py runs the same goldens through the CUDA index."""
import datetime
import os

import numpy as np
import pytest
import torch

import theoremsearch_b200 as ts
from oracle import oracle
from theoremsearch_b200 import store as st
from tests.helpers import BASE_FILTERS, OracleIndex, TableModel, golden_store_rows, load_golden, unpack_allow_mask


def test_search_rows_match_reference_run():
    g = load_golden("streamlit_rows")
    store = st.TheoremStore(golden_store_rows(g), OracleIndex(np.array(g["embeddings"], np.float32)))
    model = TableModel({k: np.array(v, np.float32) for k, v in g["queries"].items()})
    for res in g["results"]:
        filters = dict(BASE_FILTERS, top_k=res["top_k"], citation_weight=res["citation_weight"])
        got = store.search(res["query"], model, filters)
        want = res["results"]
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert list(a.keys()) == list(b.keys())          # the 16 keys, in the reference's order
            for key in a:
                if key in ("similarity", "score"):
                    assert a[key] == pytest.approx(b[key], abs=2e-6), key
                else:
                    assert a[key] == b[key], key


def test_showcase_post_filter_matches_reference_run():
    g = load_golden("showcase_search")
    index = OracleIndex(np.array(g["corpus"], np.float32))
    # the reference normalises inside cos_sim on every call; OracleIndex does it in search_host(normalize=True)
    index.rows = __import__("oracle.oracle", fromlist=["x"]).normalize_f64(index.rows)
    model = TableModel({k: np.array(v, np.float32) for k, v in g["queries"].items()})
    for res in g["results"]:
        filters = dict(g["filter_sets"][res["filters"]])
        filters["citation_range"] = tuple(filters["citation_range"])
        got = st.search_showcase(res["query"], model, g["theorems"], index, filters)
        assert [int(h["info"]["paper_url"][-5:]) for h in got] == [h["index"] for h in res["hits"]]
        assert np.allclose([h["similarity"] for h in got], [h["similarity"] for h in res["hits"]], atol=2e-6)


def _mini_store():
    D = datetime.datetime
    rows = [
        # paper_id title authors link last_updated summary journal_ref cat cats citations tid name body slogan
        ("p0", "Optimal Transport I", ["Ann", "Bob"], "https://arxiv.org/abs/2401.00001", D(2024, 1, 2), "", None, "math.AP", [], 10, 1, "Theorem 1", "b", "s"),
        ("p1", "Schemes", ["Cat"], "https://stacks.math.columbia.edu/tag/0001", None, "", None, "math.AG", [], None, 2, "Lemma 2.3", "b", "s"),
        ("p2", "Graphs", ["Bob"], "https://ARXIV.org/abs/1905.12345", D(2019, 5, 1), "", "J. Comb.", "math.CO", [], 0, 3, "Proposition A", "b", "s"),
        ("p3", "No link", ["Dan"], None, D(2020, 1, 1), "", None, "math.CO", [], 5, 4, None, "b", "s"),
        ("p4", "Transport II", ["Eve"], "https://arxiv.org/abs/2001.00002", None, "", None, "math.AP", [], 300, 5, "Main Corollary", "b", "s"),
    ]
    emb = np.eye(5, 8, dtype=np.float32)
    return st.TheoremStore(rows, OracleIndex(emb))


def test_where_clause_three_valued_logic():
    s = _mini_store()
    f = dict(BASE_FILTERS)
    assert s.build_allow(f).tolist() == [True, True, True, False, True]          # NULL link fails both source tests
    assert s.build_allow(dict(f, sources=["arXiv"])).tolist() == [True, False, True, False, True]   # ILIKE is case-insensitive
    assert s.build_allow(dict(f, sources=["Stacks Project"])).tolist() == [False, True, False, False, False]
    assert s.build_allow(dict(f, authors=["Bob", "Zed"])).tolist() == [True, False, True, False, False]
    assert s.build_allow(dict(f, tags=["math.AP"])).tolist() == [True, False, False, False, True]
    # year: arXiv rows need a year in range (NULL date fails), non-arXiv rows pass
    assert s.build_allow(dict(f, year_range=(2019, 2024))).tolist() == [True, True, True, False, False]
    assert s.build_allow(dict(f, year_range=(2020, 2024))).tolist() == [True, True, False, False, False]
    assert s.build_allow(dict(f, journal_status="Journal Article")).tolist() == [False, False, True, False, False]
    assert s.build_allow(dict(f, journal_status="Preprint Only")).tolist() == [True, False, False, False, True]
    assert s.build_allow(dict(f, types=["lemma", "corollary"])).tolist() == [False, True, False, False, True]
    assert s.build_allow(dict(f, paper_filter={"ids": {"1905.12345"}, "titles": {"transport"}})).tolist() == \
        [True, False, True, False, True]
    assert s.build_allow(dict(f, citation_range=(1, 100))).tolist() == [True, True, False, False, False]
    assert s.build_allow(dict(f, citation_range=(1, 100), include_unknown_citations=False)).tolist() == \
        [True, False, False, False, False]


def _sql_where_row(row, f):
    """The WHERE clause of streamlit_app.py:175-243 evaluated for ONE row the way Postgres would
    (a NULL operand makes the predicate not-true) — the slow, obvious form ``build_allow`` must equal."""
    (_pid, title, authors, link, last_updated, _s, journal_ref, category, _c, citations, _tid, name, _b, _sl) = row
    arxiv = link is not None and "arxiv.org" in link.lower()
    not_arxiv = link is not None and "arxiv.org" not in link.lower()
    ok = True
    src = f.get("sources") or []
    if "arXiv" in src or "Stacks Project" in src:
        ok &= ("arXiv" in src and arxiv) or ("Stacks Project" in src and not_arxiv)
    if f.get("authors"):
        ok &= bool(set(authors or ()) & set(f["authors"]))
    if f.get("tags"):
        ok &= category is not None and category in f["tags"]
    if f.get("year_range"):
        y0, y1 = f["year_range"]
        ok &= (arxiv and last_updated is not None and y0 <= last_updated.year <= y1) or not_arxiv
    if f.get("journal_status") == "Journal Article":
        ok &= arxiv and journal_ref is not None
    elif f.get("journal_status") == "Preprint Only":
        ok &= arxiv and journal_ref is None
    pf = f.get("paper_filter") or {}
    ids = [str(i).lower() for i in pf.get("ids", ())]
    titles = [str(t).lower() for t in pf.get("titles", ())]
    if ids or titles:
        ok &= (link is not None and any(i in link.lower() for i in ids)) or \
              (title is not None and any(t in title.lower() for t in titles))
    if f.get("types"):
        ok &= name is not None and any(str(t).lower() in name.lower() for t in f["types"])
    low, high = f["citation_range"]
    between = citations is not None and low <= citations <= high
    ok &= between or (f["include_unknown_citations"] and citations is None)
    return bool(ok)


def test_vectorised_where_equals_row_by_row_sql_semantics():
    rng = np.random.default_rng(5)
    authors = ["Ann", "Bob", "Cy", "Dee", "Eli", "Fay"]
    cats = ["math.AP", "math.AG", "math.NT", "cs.LG", None]
    names = ["Theorem 1.2", "Lemma 3", "Main Proposition", "Corollary A", "Remark", None, "lemma (Zorn)"]
    links = ["http://arxiv.org/abs/1905.12345v1", "https://stacks.math.columbia.edu/tag/00AB", None,
             "HTTP://ARXIV.ORG/abs/2001.00001", "http://arxiv.org/abs/2203.04567"]
    titles = ["Optimal transport and curvature", "Stacks", None, "On the Riemann zeta function", "TRANSPORT maps"]
    rows = []
    for i in range(400):
        pick = lambda xs: xs[int(rng.integers(len(xs)))]
        au = None if rng.random() < 0.1 else [authors[j] for j in rng.choice(len(authors), int(rng.integers(0, 4)), replace=False)]
        date = None if rng.random() < 0.15 else datetime.datetime(int(rng.integers(2005, 2026)), 1, 1)
        cit = None if rng.random() < 0.3 else int(rng.integers(0, 300))
        rows.append((f"p{i}", pick(titles), au, pick(links), date, "s", pick([None, "J. Math. 1"]), pick(cats),
                     "c", cit, i, pick(names), "body", "slogan"))
    s = st.TheoremStore(rows, OracleIndex(np.zeros((len(rows), 4), np.float32)))
    for trial in range(200):
        f = dict(BASE_FILTERS)
        if rng.random() < 0.5:
            f["sources"] = [x for x in ("arXiv", "Stacks Project") if rng.random() < 0.6] or ["arXiv"]
        if rng.random() < 0.4:
            f["authors"] = [authors[j] for j in rng.choice(len(authors), 2, replace=False)] + ["Nobody"]
        if rng.random() < 0.4:
            f["tags"] = [c for c in cats[:4] if rng.random() < 0.5] + ["math.ZZ"]
        if rng.random() < 0.4:
            y0 = int(rng.integers(2000, 2026))
            f["year_range"] = (y0, y0 + int(rng.integers(0, 12)))
        f["journal_status"] = ["All", "Journal Article", "Preprint Only"][int(rng.integers(3))]
        if rng.random() < 0.3:
            f["paper_filter"] = {"ids": {"1905.12345", "tag/00"} if rng.random() < 0.5 else set(),
                                 "titles": {"Transport"} if rng.random() < 0.7 else set()}
        if rng.random() < 0.4:
            f["types"] = [t for t in ("Lemma", "corollary", "theorem", "proposition") if rng.random() < 0.5]
        lo = int(rng.integers(0, 100))
        f["citation_range"] = (lo, lo + int(rng.integers(0, 250)))
        f["include_unknown_citations"] = bool(rng.random() < 0.5)
        want = [_sql_where_row(r, f) for r in rows]
        assert s.build_allow(f).tolist() == want, f


def test_filters_apply_before_limit():
    s = _mini_store()
    model = TableModel({"q": np.array([0.1, 0.9, 0.5, 0.4, 0.3, 0, 0, 0], np.float32)})
    f = dict(BASE_FILTERS, top_k=2, citation_weight=0.0, sources=["arXiv"])
    got = s.search("q", model, f)
    assert [r["theorem_id"] for r in got] == [3, 5]        # row 1 (best) is filtered out BEFORE the limit
    # reference quirk kept: the SQL filter is ILIKE (case-insensitive, :181) but the row's "source" is a
    # case-sensitive Python `"arxiv.org" in link` (:293) -> an upper-case host passes the arXiv filter yet
    # is labelled Stacks Project.
    assert got[0]["source"] == "Stacks Project" and got[0]["type"] == "proposition" and got[0]["journal_published"] is True
    assert got[1]["source"] == "arXiv"
    assert got[1]["type"] == "corollary" and got[1]["year"] is None
    assert s.search("q", model, dict(f, sources=[])) == []
    # similarity is 1 + cosine (streamlit_app.py:275)
    qn = model.table["q"] / np.linalg.norm(model.table["q"])
    assert got[0]["similarity"] == pytest.approx(1.0 + qn[2], abs=1e-6)


def test_search_accepts_an_embedded_query():
    """``search(query: str or ndarray, model, filters)`` — a raw embedding gives the rows the
    text query gives (it is normalised like ``normalize_embeddings=True``), with no model call."""
    import torch
    s = _mini_store()
    raw = np.array([0.1, 0.9, 0.5, 0.4, 0.3, 0, 0, 0], np.float32)
    model = TableModel({"q": raw})
    f = dict(BASE_FILTERS, top_k=3, citation_weight=0.0)
    want = s.search("q", model, f)
    for emb in (raw * 5.0, torch.from_numpy(raw * 0.01)):
        got = s.search(emb, None, f)
        assert [r["theorem_id"] for r in got] == [r["theorem_id"] for r in want]
        assert [r["similarity"] for r in got] == pytest.approx([r["similarity"] for r in want], abs=1e-6)
    assert s.search(raw, None, dict(f, citation_weight=0.5))[0]["score"] >= want[0]["similarity"] - 1e-6


def test_citation_weight_reranks_pool():
    s = _mini_store()
    model = TableModel({"q": np.array([0.5, 0.5, 0.5, 0.5, 0.49, 0, 0, 0], np.float32)})
    f = dict(BASE_FILTERS, top_k=3, citation_weight=0.1)
    got = s.search("q", model, f)
    # ln(300) lifts row 4 to the top, ln(10) row 0 next; citations 0 / NULL add nothing
    assert [r["theorem_id"] for r in got] == [5, 1, 2]
    assert got[0]["score"] == pytest.approx(got[0]["similarity"] + 0.1 * np.log(300.0))
    assert got[2]["score"] == got[2]["similarity"]


def test_latest_slogan_rows_distinct_on():
    # (theorem_id, slogan_id) per embedding row -> keep the highest slogan_id of each theorem
    assert st.latest_slogan_rows([(7, 1), (3, 2), (7, 5), (3, 1), (9, 4)]) == [1, 2, 4]


def test_allow_mask_bit_layout():
    allow = np.zeros(70, dtype=bool)
    allow[[0, 31, 32, 69]] = True
    m = ts.pack_allow_mask(allow)
    w = m.numpy().view(np.uint32)
    assert w.tolist() == [0x80000001, 0x00000001, 1 << 5]
    assert np.array_equal(unpack_allow_mask(m, 70), allow)


def test_infer_type_and_pool_size():
    assert st.infer_type("Lemma 3.1") == "lemma" and st.infer_type(None) == "theorem"
    assert st.infer_type("Remark") == "theorem" and st.infer_type("Main Theorem (Corollary)") == "theorem"
    assert st.pool_size(1) == 50 and st.pool_size(5) == 50 and st.pool_size(20) == 200


# --------------------------------------------------------------------------------------- f4: metrics
def test_product_metrics_match_reference_run():
    """theoremsearch_b200.metrics (top-k in) against the numbers the reference's own functions produced
    (tests/golden/make_golden.py ran compare_embeddings.py's six metrics on the full sim matrix)."""
    from theoremsearch_b200 import metrics
    g = load_golden("compare_embeddings_metrics")
    qrels = {int(q): {int(d): v for d, v in rd.items()} for q, rd in g["qrels"].items()}
    ranked = np.array(g["ranked_top10"], dtype=np.int64)
    for k_str, want in g["metrics"].items():
        k = int(k_str)
        got = {"precision": metrics.precision_at_k(ranked, qrels, k), "hit": metrics.hit_at_k(ranked, qrels, k),
               "mrr": metrics.mrr_at_k(ranked, qrels, k), "ndcg": metrics.ndcg_at_k(ranked, qrels, k),
               "err": metrics.err_at_k(ranked, qrels, k), "q_measure": metrics.q_measure_at_k(ranked, qrels, k)}
        for name in want:
            assert got[name] == pytest.approx(want[name], abs=1e-12), (k, name)
        # identical to the oracle's restatement, and padding (-1) is ignored
        assert got["ndcg"] == pytest.approx(oracle.ndcg_at_k(ranked, qrels, k), abs=1e-15)
        padded = np.concatenate([ranked[:, :k], np.full((ranked.shape[0], 3), -1)], axis=1)
        assert metrics.mrr_at_k(padded, qrels, None) == pytest.approx(want["mrr"], abs=1e-12)
    rep = metrics.evaluate_rankings(ranked, qrels, 3)
    assert set(rep) == {"P@1", "H@3", "MRR@3", "nDCG@3", "ERR@3", "Q-measure@3"}
    assert rep["MRR@3"] == pytest.approx(g["metrics"]["3"]["mrr"], abs=1e-12)
    q = metrics._generate_qrels([("a", "p1"), ("b", "p2")], [("s", "p1"), ("t", "p2"), ("u", "p1")])
    assert q == {0: {0: 0.5, 1: 0, 2: 0.5}, 1: {0: 0, 1: 0.5, 2: 0}}


# --------------------------------------------------------------------------------------- f2: formats
def test_pgvector_text_and_latest_slogan():
    from theoremsearch_b200 import formats
    v = formats.parse_pgvector_text("[0.5,-1,2e-3]")
    assert v.dtype == np.float32 and np.allclose(v, [0.5, -1.0, 2e-3])
    assert np.array_equal(formats.parse_pgvector_text(b"[1,2]"), np.array([1, 2], np.float32))
    assert np.array_equal(formats.parse_pgvector_text([1.5, 2.5]), np.array([1.5, 2.5], np.float32))
    with pytest.raises(ValueError):
        formats.parse_pgvector_text("1,2,3")
    # DISTINCT ON (theorem_id) ORDER BY theorem_id, slogan_id DESC  (streamlit_app.py:254-259)
    theorem = [7, 3, 7, 3, 9, 7]
    slogan = [10, 11, 12, 13, 14, 5]
    keep, t, s = formats.latest_slogan_per_theorem(theorem, slogan)
    assert t.tolist() == [3, 7, 9] and s.tolist() == [13, 12, 14] and keep.tolist() == [3, 2, 4]


def test_embedding_library_files_roundtrip(tmp_path):
    """The pair of files app_create_embeddings.py:85-93 writes and app_showcase_model.py:41-58 reads."""
    import pickle
    import torch
    from theoremsearch_b200 import formats
    emb = torch.from_numpy(oracle.synthetic_rows(0, 12, 16, seed=4))
    meta = [{"paper_title": f"P{i}", "type": "theorem", "content": f"c{i}"} for i in range(12)]
    formats.save_embedding_library(str(tmp_path), emb, meta)
    assert sorted(os.listdir(tmp_path)) == ["corpus_embeddings.pt", "theorems_data.pkl"]
    assert torch.equal(torch.load(tmp_path / "corpus_embeddings.pt"), emb)            # what the reference loads
    with open(tmp_path / "theorems_data.pkl", "rb") as f:
        assert pickle.load(f) == meta
    e2, m2 = formats.load_embedding_library(str(tmp_path), as_index=False)
    assert torch.equal(e2, emb) and m2 == meta
    assert formats.load_embedding_library(str(tmp_path / "missing")) == (None, None)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) works without a GPU and prints
    exactly one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1", "--rows", "200000", "--cpu-sample-rows", "20000"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("configs[1]") and d["gpu_launches"] == 0


# --------------------------------------------------------------------------------------- round 2 additions
def test_oracle_upsert_rows_is_row_by_row_on_conflict_do_update():
    """ec2/rds/upsert.py:29-52 runs one INSERT ... ON CONFLICT DO UPDATE per row (executemany): existing ids keep their
    place and take the new embedding, new ids append in order, a repeated id ends with its last embedding."""
    table = {}
    assert oracle.upsert_rows(table, [7, 3, 9], np.array([[1.0], [2.0], [3.0]])) == 0
    assert list(table) == [7, 3, 9]
    n = oracle.upsert_rows(table, [3, 11, 3, 7], np.array([[20.0], [40.0], [21.0], [10.0]]))
    assert n == 2                                        # ids 3 and 7 existed
    assert list(table) == [7, 3, 9, 11]                  # places kept, 11 appended
    assert [float(table[i][0]) for i in table] == [10.0, 21.0, 3.0, 40.0]


def test_hierarchical_corpus_generator_is_deterministic_and_structured():
    from theoremsearch_b200 import synthetic
    m = synthetic.HierarchicalCorpus(64, n_leaves=80, device="cpu", seed=0, n_blobs=4, rank=8, n_bases=2)
    a, b = m.rows(500, seed=5), m.rows(500, seed=5)
    assert torch.equal(a, b) and a.shape == (500, 64)
    q = m.queries(20)
    assert q.shape == (20, 64) and not torch.equal(q[:10], m.queries(10, seed=1))
    # rows of one blob are far closer to each other than to rows of another blob
    x = torch.nn.functional.normalize(m.rows(2000, seed=9), dim=1)
    sims = x @ x.T
    assert float(sims.topk(2, dim=1).values[:, 1].mean()) > 0.5 > float(sims.mean())


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` on a tiny corpus: one JSON line, the steps / warm-up it was given, nothing
    extrapolated, the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--rows", "20000",
                          "--steps", "3", "--warmup", "4"], capture_output=True, text=True, check=True, timeout=300).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["steps"] == 3 and d["warmup"] == 4 and d["extrapolated"] is False
    assert d["rows_timed"] == 20000 and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-6


def test_host_metrics_on_random_graded_judgements_match_the_oracle():
    """The array-arithmetic (host) evaluation of a ranking against the oracle's per-query loops: graded relevances,
    rows padded with -1, unjudged queries, every cut."""
    from theoremsearch_b200 import metrics
    rng = np.random.default_rng(4)
    nq, n_docs, width = 120, 300, 12
    ranked = np.stack([rng.permutation(n_docs)[:width] for _ in range(nq)]).astype(np.int64)
    for q in np.flatnonzero(rng.random(nq) < 0.2):
        ranked[q, rng.integers(1, width):] = -1
    qrels = {}
    for q in range(nq):
        pool = np.concatenate([ranked[q][ranked[q] >= 0], rng.integers(0, n_docs, size=20)])
        docs = list(dict.fromkeys(int(d) for d in rng.permutation(pool)))[: int(rng.integers(1, 25))]
        rels = rng.choice([0.0, 0.5, 2.0, 3.0], size=len(docs)).tolist()
        rels[int(rng.integers(0, len(docs)))] = 1.0
        qrels[q] = dict(zip(docs, rels))
    for k in (1, 3, 5, 10, 12):
        assert metrics.precision_at_k(ranked, qrels, k) == pytest.approx(oracle.precision_at_k(ranked, qrels, k), abs=1e-12)
        assert metrics.hit_at_k(ranked, qrels, k) == pytest.approx(oracle.hit_at_k(ranked, qrels, k), abs=1e-12)
        assert metrics.mrr_at_k(ranked, qrels, k) == pytest.approx(oracle.mrr_at_k(ranked, qrels, k), abs=1e-12)
        assert metrics.ndcg_at_k(ranked, qrels, k) == pytest.approx(oracle.ndcg_at_k(ranked, qrels, k), abs=1e-12)
        assert metrics.ndcg_at_k(ranked, qrels, k, gain="linear") == pytest.approx(
            oracle.ndcg_at_k(ranked, qrels, k, gain="linear"), abs=1e-12)
        for max_rel in (None, 4.0):
            assert metrics.err_at_k(ranked, qrels, k, max_rel) == pytest.approx(oracle.err_at_k(ranked, qrels, k, max_rel), abs=1e-12)
            assert metrics.q_measure_at_k(ranked, qrels, k, max_rel) == pytest.approx(
                oracle.q_measure_at_k(ranked, qrels, k, max_rel), abs=1e-12)
    assert metrics.mrr_at_k(ranked, qrels, None) == pytest.approx(oracle.mrr_at_k(ranked, qrels, None), abs=1e-12)
    some = {q: v for q, v in qrels.items() if q % 5}                 # every fifth query unjudged
    for k in (3, 10):
        assert metrics.ndcg_at_k(ranked, some, k) == pytest.approx(oracle.ndcg_at_k(ranked, some, k), abs=1e-12)
        assert metrics.err_at_k(ranked, some, k) == pytest.approx(oracle.err_at_k(ranked, some, k), abs=1e-12)
        assert metrics.q_measure_at_k(ranked, some, k) == pytest.approx(oracle.q_measure_at_k(ranked, some, k), abs=1e-12)


def test_oracle_table_writes():
    """oracle.upsert_rows / delete_rows: the dict model of the reference's writer statements."""
    t = {}
    assert oracle.upsert_rows(t, [5, 7, 5], np.array([[1.0], [2.0], [3.0]], np.float32)) == 0
    assert list(t) == [5, 7] and float(t[5][0]) == 3.0                # a repeated id keeps its last row, first place
    assert oracle.upsert_rows(t, [7, 9], np.array([[4.0], [5.0]], np.float32)) == 1
    assert list(t) == [5, 7, 9] and float(t[7][0]) == 4.0
    assert oracle.delete_rows(t, [7, 7, 100]) == 1 and list(t) == [5, 9]
    assert oracle.delete_rows(t, []) == 0
