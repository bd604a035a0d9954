"""The oracle against the golden vectors produced by running the reference's own functions
(tests/golden/make_golden.py), plus its internal consistency. CPU only."""
import numpy as np
import pytest
import torch

from oracle import oracle
from tests.helpers import load_golden


def test_search_theorems_matches_reference_run():
    g = load_golden("test_app_search_theorems")
    corpus = np.array(g["corpus"], dtype=np.float32)
    for res in g["results"]:
        q = np.array(g["queries"][res["query"]], dtype=np.float32)
        top, scores = oracle.search_theorems_topk(q, corpus, 5)
        assert top.tolist() == [h["index"] for h in res["hits"]]
        assert [f"{s:.4f}" for s in scores] == [h["similarity_4dp"] for h in res["hits"]]
        assert [g["theorem_types"][i].capitalize() for i in top] == [h["type"] for h in res["hits"]]


def test_showcase_pool_matches_reference_run():
    g = load_golden("showcase_search")
    corpus = np.array(g["corpus"], dtype=np.float32)
    # unfiltered filter set 0: the hits are the head of the reference's torch.topk pool
    for res in g["results"]:
        if res["filters"] != 0:
            continue
        q = np.array(g["queries"][res["query"]], dtype=np.float32)
        top, scores = oracle.showcase_pool(q, corpus, 200)
        assert top[: len(res["hits"])].tolist() == [h["index"] for h in res["hits"]]
        assert np.allclose(scores[: len(res["hits"])], [h["similarity"] for h in res["hits"]], atol=1e-6)


def test_metrics_match_reference_run():
    g = load_golden("compare_embeddings_metrics")
    docs = np.array(g["docs"], dtype=np.float32)
    queries = np.array(g["queries"], dtype=np.float32)
    qrels = {int(q): {int(d): v for d, v in rd.items()} for q, rd in g["qrels"].items()}
    sim, ranked = oracle.batched_ranking(queries, docs)
    assert ranked[:, :10].tolist() == g["ranked_top10"]
    for k_str, want in g["metrics"].items():
        k = int(k_str)
        got = {
            "precision": oracle.precision_at_k(ranked, qrels, k), "hit": oracle.hit_at_k(ranked, qrels, k),
            "mrr": oracle.mrr_at_k(ranked, qrels, k), "ndcg": oracle.ndcg_at_k(ranked, qrels, k),
            "err": oracle.err_at_k(ranked, qrels, k), "q_measure": oracle.q_measure_at_k(ranked, qrels, k),
        }
        for name in want:
            assert got[name] == pytest.approx(want[name], abs=1e-12), (k, name)
        # the metrics only ever look at the first k ranks: top-k ids are sufficient input
        got_topk = oracle.ndcg_at_k(ranked[:, :k], qrels, k)
        assert got_topk == pytest.approx(want["ndcg"], abs=1e-12)


def test_pgvector_rows_match_reference_run():
    g = load_golden("streamlit_rows")
    emb = np.array(g["embeddings"], dtype=np.float32)
    for res in g["results"]:
        q = oracle.normalize(np.array(g["queries"][res["query"]], dtype=np.float32)).numpy()[0]
        k, w = res["top_k"], res["citation_weight"]
        if w == 0.0:
            order, sim = oracle.pgvector_search(q, emb, k)
            assert [r["theorem_id"] for r in res["results"]] == [1000 + int(i) for i in order]
            assert np.allclose([r["similarity"] for r in res["results"]], sim, atol=1e-12)
            assert res["sql"][0]["sql_has_candidates_cte"] is False
        else:
            assert res["sql"][0]["limit_literal"] == str(oracle.pool_size(k))
            order, sim = oracle.pgvector_search(q, emb, oracle.pool_size(k))
            cits = [g["rows"][i][9] for i in order]
            sel, ws = oracle.citation_rerank(sim, cits, w, k)
            assert [r["theorem_id"] for r in res["results"]] == [1000 + int(order[j]) for j in sel]
            assert np.allclose([r["score"] for r in res["results"]], ws, atol=1e-12)


# ------------------------------------------------------------------------------ self-consistency
def test_tie_rule_and_padding():
    rows = np.array([[1, 0], [0, 1], [1, 0], [0, 1], [-1, 0]], dtype=np.float32)
    s, i = oracle.exact_search(np.array([1, 0], np.float32), rows, 7)
    assert i[0].tolist() == [0, 2, 1, 3, 4, -1, -1]
    assert np.isneginf(s[0][5:]).all()
    s, i = oracle.exact_search(np.array([1, 0], np.float32), rows, 2, allow=np.array([0, 1, 1, 1, 1], bool))
    assert i[0].tolist() == [2, 1]
    s, i = oracle.exact_search(np.array([1, 0], np.float32), rows, 3, ids=np.array([50, 40, 30, 20, 10]))
    assert i[0].tolist() == [50, 30, 40]   # ties break on ROW order (rows 0,2 then 1), ids are labels


def test_pack_key_order_is_score_desc_then_row_asc():
    rng = np.random.default_rng(0)
    scores = np.concatenate([rng.standard_normal(200).astype(np.float32), [0.0, -0.0, 1.0, 1.0, -np.inf, np.inf]])
    rows = rng.permutation(scores.size)
    keys = [oracle.pack_key(s, r) for s, r in zip(scores, rows)]
    by_key = sorted(range(len(keys)), key=lambda j: -keys[j])
    by_rule = sorted(range(len(keys)), key=lambda j: (-float(scores[j]) if scores[j] == scores[j] else np.inf, rows[j]))
    assert by_key == by_rule
    for s, r in zip(scores, rows):
        s2, r2 = oracle.unpack_key(oracle.pack_key(s, r))
        assert r2 == r and (s2 == s)
    assert oracle.pack_key(0.0, 3) == oracle.pack_key(-0.0, 3)
    assert oracle.pack_key(float("nan"), 0) < oracle.pack_key(-np.inf, 0xFFFFFFFE)


def test_merge_shards_equals_unsharded():
    rng = np.random.default_rng(1)
    rows = oracle.normalize_f64(rng.standard_normal((300, 16)).astype(np.float32))
    rows[250] = rows[10]
    q = oracle.normalize_f64(rng.standard_normal((4, 16)).astype(np.float32))
    s_all, i_all = oracle.exact_search(q, rows, 12)
    for world in (1, 2, 3, 8):
        bounds = oracle.shard_bounds(300, world)
        assert bounds[0][0] == 0 and bounds[-1][1] == 300
        parts = [oracle.exact_search(q, rows[lo:hi], 12) for lo, hi in bounds]
        s, i = oracle.merge_shards([p[0] for p in parts], [p[1] for p in parts], [lo for lo, _ in bounds], 12)
        assert np.array_equal(i, i_all) and np.array_equal(s, s_all)


def test_normalize_variants_agree():
    x = oracle.synthetic_rows(0, 500, 96, seed=3) * 5
    x[7] = 0
    a = oracle.normalize(x).numpy()
    b = oracle.normalize_f64(x)
    assert np.max(np.abs(a - b)) < 1e-7 and np.all(b[7] == 0)
    assert np.allclose(np.linalg.norm(b[np.arange(500) != 7].astype(np.float64), axis=1), 1.0, atol=1e-6)


def test_cos_sim_is_normalise_then_mm():
    a = oracle.synthetic_rows(0, 5, 32, seed=1)
    b = oracle.synthetic_rows(0, 9, 32, seed=2)
    s = oracle.cos_sim(a, b).numpy()
    want = (a / np.linalg.norm(a, axis=1, keepdims=True)) @ (b / np.linalg.norm(b, axis=1, keepdims=True)).T
    assert s.shape == (5, 9) and np.allclose(s, want, atol=1e-6)
    assert oracle.cos_sim(a[0], b).shape == (1, 9)


def test_synthetic_rows_are_position_independent():
    a = oracle.synthetic_rows(0, 64, 8, seed=0)
    b = oracle.synthetic_rows(10, 20, 8, seed=0)
    assert np.array_equal(a[10:30], b)


def test_ivf_oracle_full_probe_equals_exact():
    rng = np.random.default_rng(2)
    rows = oracle.normalize_f64(rng.standard_normal((400, 16)).astype(np.float32))
    cent = oracle.normalize_f64(rng.standard_normal((8, 16)).astype(np.float32))
    assign = oracle.ivf_assign(rows, cent)
    q = oracle.normalize_f64(rng.standard_normal((1, 16)).astype(np.float32))[0]
    s, i = oracle.ivf_search(q, rows, cent, assign, 10, nprobe=8)
    s2, i2 = oracle.exact_search(q, rows, 10)
    assert np.array_equal(i, i2[0])
    s3, i3 = oracle.ivf_search(q, rows, cent, assign, 10, nprobe=2)
    assert 0.0 < oracle.recall_at_k(i3, i2[0]) <= 1.0


def test_merge_of_shards_equals_unsharded_hypothesis():
    """Top-k of a union == top-k of the per-shard top-k's, for any contiguous sharding."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as hst

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(n=hst.integers(1, 400), d=hst.integers(1, 24), k=hst.integers(1, 30), world=hst.integers(1, 5),
           seed=hst.integers(0, 1000))
    def run(n, d, k, world, seed):
        rows = oracle.bf16_round(oracle.normalize_f64(oracle.synthetic_rows(0, n, d, seed=seed)))
        if n > 2:
            rows[n - 1] = rows[0]
        q = oracle.normalize_f64(oracle.synthetic_queries(2, d, seed=seed + 7))
        want_s, want_i = oracle.exact_search(q, rows, k)
        ss, rr, base = [], [], []
        for lo, hi in oracle.shard_bounds(n, world):
            s, i = oracle.exact_search(q, rows[lo:hi], k)
            ss.append(s)
            rr.append(i)
            base.append(lo)
        s, r = oracle.merge_shards(ss, rr, base, k)
        assert np.array_equal(r, want_i)
        assert np.allclose(s[r >= 0], want_s[want_i >= 0])

    run()
