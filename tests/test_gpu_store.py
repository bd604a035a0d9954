"""The reference-run goldens through the REAL CUDA index (C ABI): app-level rows, citation re-rank,
showcase post-filter, test_app search, and the batched ranking that feeds the evaluation metrics."""
import numpy as np
import pytest
import torch

from oracle import oracle
from tests.helpers import BASE_FILTERS, TableModel, golden_store_rows, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ts():
    import theoremsearch_b200 as ts
    return ts


@pytest.mark.parametrize("dtype,tol", [("f32", 1e-5), ("bf16", 1e-3)])
def test_streamlit_rows_golden(ts, dtype, tol):
    from theoremsearch_b200 import store as st
    g = load_golden("streamlit_rows")
    index = ts.build_index(np.array(g["embeddings"], np.float32), dtype=dtype, normalize=False)
    store = st.TheoremStore(golden_store_rows(g), index)
    model = TableModel({k: np.array(v, np.float32) for k, v in g["queries"].items()})
    for res in g["results"]:
        filters = dict(BASE_FILTERS, top_k=res["top_k"], citation_weight=res["citation_weight"])
        got = store.search(res["query"], model, filters)
        want = res["results"]
        assert [r["theorem_id"] for r in got] == [r["theorem_id"] for r in want]
        for a, b in zip(got, want):
            assert list(a.keys()) == list(b.keys())
            for key in a:
                if key in ("similarity", "score"):
                    assert a[key] == pytest.approx(b[key], abs=tol), key
                else:
                    assert a[key] == b[key], key


def test_showcase_golden(ts):
    from theoremsearch_b200 import store as st
    g = load_golden("showcase_search")
    index = ts.build_index(np.array(g["corpus"], np.float32), dtype="f32", normalize=True)
    model = TableModel({k: np.array(v, np.float32) for k, v in g["queries"].items()})
    for res in g["results"]:
        filters = dict(g["filter_sets"][res["filters"]])
        got = st.search_showcase(res["query"], model, g["theorems"], index, filters)
        assert [int(h["info"]["paper_url"][-5:]) for h in got] == [h["index"] for h in res["hits"]]
        assert np.allclose([h["similarity"] for h in got], [h["similarity"] for h in res["hits"]], atol=1e-5)


def test_test_app_golden(ts):
    g = load_golden("test_app_search_theorems")
    corpus = np.array(g["corpus"], np.float32)
    theorems = [{"type": t} for t in g["theorem_types"]]
    model = TableModel({k: np.array(v, np.float32) for k, v in g["queries"].items()})
    index = ts.build_index(corpus, dtype="f32", normalize=True)
    for res in g["results"]:
        hits = ts.search_theorems(res["query"], model, theorems, index)
        assert [h["index"] for h in hits] == [h["index"] for h in res["hits"]]
        assert [f"{h['similarity']:.4f}" for h in hits] == [h["similarity_4dp"] for h in res["hits"]]


def test_metrics_golden_from_gpu_topk(ts):
    g = load_golden("compare_embeddings_metrics")
    docs = np.array(g["docs"], np.float32)
    queries = np.array(g["queries"], np.float32)
    qrels = {int(q): {int(d): v for d, v in rd.items()} for q, rd in g["qrels"].items()}
    index = ts.build_index(docs, dtype="f32", normalize=True)
    scores, ids = ts.cos_sim_topk(torch.from_numpy(queries), index, 10)
    ranked = ids.cpu().numpy()
    assert ranked.tolist() == g["ranked_top10"]
    for k_str, want in g["metrics"].items():
        k = int(k_str)
        assert oracle.ndcg_at_k(ranked, qrels, k) == pytest.approx(want["ndcg"], abs=1e-12)
        assert oracle.mrr_at_k(ranked, qrels, k) == pytest.approx(want["mrr"], abs=1e-12)
        assert oracle.hit_at_k(ranked, qrels, k) == pytest.approx(want["hit"], abs=1e-12)


def test_where_mask_through_kernel(ts):
    from theoremsearch_b200 import store as st
    g = load_golden("streamlit_rows")
    emb = np.array(g["embeddings"], np.float32)
    index = ts.build_index(emb, dtype="f32", normalize=False)
    store = st.TheoremStore(golden_store_rows(g), index)
    model = TableModel({k: np.array(v, np.float32) for k, v in g["queries"].items()})
    filters = dict(BASE_FILTERS, top_k=7, citation_weight=0.0, sources=["Stacks Project"], citation_range=(0, 250),
                   include_unknown_citations=False, types=["lemma", "theorem"])
    allow = store.build_allow(filters)
    assert 0 < allow.sum() < len(store)
    got = store.search("p0", model, filters)
    q = oracle.normalize(np.array(g["queries"]["p0"], np.float32)).numpy()[0]
    order, sim = oracle.pgvector_search(q, emb, 7, allow=allow)
    assert [r["theorem_id"] for r in got] == [1000 + int(i) for i in order]
    assert np.allclose([r["similarity"] for r in got], sim, atol=1e-5)
    assert all(r["source"] == "Stacks Project" for r in got)


def test_streamlit_rows_golden_through_ivf_store(ts):
    """The app-level search over the IVF path (all lists probed => same rows as the exact scan), filters and
    citation re-rank included: what switching the production table to an ivfflat index would serve."""
    from theoremsearch_b200 import store as st
    g = load_golden("streamlit_rows")
    emb = np.array(g["embeddings"], np.float32)
    index = ts.build_index(emb, dtype="bf16", normalize=False)
    nlist = max(2, min(8, emb.shape[0] // 4))
    index.ivf_train(nlist, iters=3, seed=0)
    index.ivf_build("bf16")
    exact = st.TheoremStore(golden_store_rows(g), index)
    ann = st.TheoremStore(golden_store_rows(g), index, ann={"nprobe": nlist, "rescore_k": 64})
    model = TableModel({k: np.array(v, np.float32) for k, v in g["queries"].items()})
    for res in g["results"]:
        filters = dict(BASE_FILTERS, top_k=res["top_k"], citation_weight=res["citation_weight"])
        a, b = ann.search(res["query"], model, filters), exact.search(res["query"], model, filters)
        assert a == b
        assert [r["theorem_id"] for r in a] == [r["theorem_id"] for r in res["results"]]
    # a filter that removes most rows still fills top_k from the eligible ones
    filters = dict(BASE_FILTERS, top_k=3, citation_weight=0.0, sources=["arXiv"])
    q = next(iter(g["queries"]))
    assert ann.search(q, model, filters) == exact.search(q, model, filters)
