"""K6 — the six evaluation metrics computed on the GPU from the batched top-k (ts_eval_rankings) against
(a) the numbers the reference's own functions produced (tests/golden/compare_embeddings_metrics.json),
(b) the oracle's fp64 restatement of compare_embeddings.py:95-371 on random rankings / judgements.
Tolerance 1e-12 (fp64 sums in the reference's order; exp2 / log2 may differ in the last bit)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import oracle
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu

TOL = 1e-12
ORACLE = {"precision": oracle.precision_at_k, "hit": oracle.hit_at_k, "mrr": oracle.mrr_at_k,
          "ndcg": oracle.ndcg_at_k, "err": oracle.err_at_k, "q_measure": oracle.q_measure_at_k}


@pytest.fixture(scope="module")
def ts():
    import theoremsearch_b200 as ts
    return ts


@pytest.fixture(scope="module")
def metrics(ts):
    from theoremsearch_b200 import metrics
    return metrics


def random_case(seed, nq=257, n_docs=400, width=12, unjudged=0.0, pad_rows=0.15):
    """A ranking [nq, width] (some rows -1 padded at the end) and graded judgements; every judged query has exactly
    one document of relevance 1 somewhere (the reference's `correct` document)."""
    rng = np.random.default_rng(seed)
    ranked = np.stack([rng.permutation(n_docs)[:width] for _ in range(nq)]).astype(np.int64)
    for q in np.flatnonzero(rng.random(nq) < pad_rows):
        ranked[q, rng.integers(1, width):] = -1
    qrels = {}
    for q in range(nq):
        if rng.random() < unjudged:
            continue
        n = int(rng.integers(1, 40))
        # half the judged docs come from the ranking itself so that the metrics are not all zero
        pool = np.concatenate([ranked[q][ranked[q] >= 0], rng.integers(0, n_docs, size=n)])
        docs = list(dict.fromkeys(int(d) for d in rng.permutation(pool)))[:n]
        rels = rng.choice([0.0, 0.5, 2.0, 3.0], size=len(docs)).tolist()
        rels[int(rng.integers(0, len(docs)))] = 1.0
        qrels[q] = dict(zip(docs, rels))
    return ranked, qrels


def test_golden_numbers_of_the_reference_run(ts, metrics):
    g = load_golden("compare_embeddings_metrics")
    docs = np.array(g["docs"], np.float32)
    queries = np.array(g["queries"], np.float32)
    qrels = {int(q): {int(d): v for d, v in rd.items()} for q, rd in g["qrels"].items()}
    index = ts.build_index(docs, dtype="f32", normalize=True)
    ranked = metrics.rank_concepts(torch.from_numpy(queries), index, 10, as_tensor=True)
    assert ranked.is_cuda and ranked.cpu().tolist() == g["ranked_top10"]
    table = metrics.JudgedTable(qrels, ranked.shape[0], ranked.device, max_k=10)
    launches = ts.kernel_launches()
    for k_str, want in g["metrics"].items():
        k = int(k_str)
        got = table.evaluate(ranked, {m: k for m in metrics.METRIC_ORDER})
        for name in want:
            assert got[name] == pytest.approx(want[name], abs=TOL), (k, name)
    assert ts.kernel_launches() - launches == 2 * len(g["metrics"])        # metrics + mean kernel per evaluation
    # the public functions route a CUDA ranking to the device as well (dict or table)
    for k_str, want in g["metrics"].items():
        k = int(k_str)
        assert metrics.precision_at_k(ranked, qrels, k) == pytest.approx(want["precision"], abs=TOL)
        assert metrics.hit_at_k(ranked, table, k) == pytest.approx(want["hit"], abs=TOL)
        assert metrics.mrr_at_k(ranked, qrels, k) == pytest.approx(want["mrr"], abs=TOL)
        assert metrics.ndcg_at_k(ranked, table, k) == pytest.approx(want["ndcg"], abs=TOL)
        assert metrics.err_at_k(ranked, qrels, k) == pytest.approx(want["err"], abs=TOL)
        assert metrics.q_measure_at_k(ranked, qrels, k) == pytest.approx(want["q_measure"], abs=TOL)
    rep = metrics.evaluate_rankings(ranked, table, 3)
    assert rep == pytest.approx(metrics.evaluate_rankings(ranked.cpu().numpy(), qrels, 3), abs=TOL)
    table.close()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_rankings_against_the_oracle(metrics, seed):
    ranked, qrels = random_case(seed)
    dev = torch.device("cuda", 0)
    r = torch.from_numpy(ranked).to(dev)
    table = metrics.JudgedTable(qrels, ranked.shape[0], dev, max_k=16)
    for k in (1, 3, 5, 10, 12):
        got, per_query = table.evaluate(r, {m: k for m in metrics.METRIC_ORDER}, per_query=True)
        for name, fn in ORACLE.items():
            assert got[name] == pytest.approx(fn(ranked, qrels, k), abs=TOL), (k, name)
        assert per_query.shape == (ranked.shape[0], 6)
        assert per_query.mean(dim=0).cpu().numpy() == pytest.approx([got[m] for m in metrics.METRIC_ORDER], abs=TOL)
        # per query: the oracle on a one-row ranking
        for q in (0, 17, 256):
            one = ranked[q:q + 1]
            assert float(per_query[q, 3]) == pytest.approx(oracle.ndcg_at_k(one, {0: qrels[q]}, k), abs=TOL)
            assert float(per_query[q, 4]) == pytest.approx(
                oracle.err_at_k(one, {0: qrels[q]}, k, max_rel=table.max_relevance), abs=TOL)
    # different cuts per metric in ONE launch, no cut for MRR (k=None: the whole ranking)
    got = table.evaluate(r, {"precision": 1, "hit": 5, "mrr": None, "ndcg": 10, "err": 3, "q_measure": 12})
    assert got["precision"] == pytest.approx(oracle.precision_at_k(ranked, qrels, 1), abs=TOL)
    assert got["hit"] == pytest.approx(oracle.hit_at_k(ranked, qrels, 5), abs=TOL)
    assert got["mrr"] == pytest.approx(oracle.mrr_at_k(ranked, qrels, None), abs=TOL)
    assert got["ndcg"] == pytest.approx(oracle.ndcg_at_k(ranked, qrels, 10), abs=TOL)
    assert got["err"] == pytest.approx(oracle.err_at_k(ranked, qrels, 3), abs=TOL)
    assert got["q_measure"] == pytest.approx(oracle.q_measure_at_k(ranked, qrels, 12), abs=TOL)
    # linear gain, explicit normaliser
    assert table.evaluate(r, {"ndcg": 5}, gain="linear")["ndcg"] == pytest.approx(
        oracle.ndcg_at_k(ranked, qrels, 5, gain="linear"), abs=TOL)
    for max_rel in (3.0, 4.0, 2.5):
        got = table.evaluate(r, {"err": 10, "q_measure": 10}, max_rel=max_rel)
        assert got["err"] == pytest.approx(oracle.err_at_k(ranked, qrels, 10, max_rel=max_rel), abs=TOL)
        assert got["q_measure"] == pytest.approx(oracle.q_measure_at_k(ranked, qrels, 10, max_rel=max_rel), abs=TOL)
    # the host arithmetic of the same module agrees too
    assert metrics.ndcg_at_k(ranked, qrels, 5) == pytest.approx(metrics.ndcg_at_k(r, table, 5), abs=TOL)
    assert metrics.q_measure_at_k(ranked, qrels, 5) == pytest.approx(metrics.q_measure_at_k(r, table, 5), abs=TOL)
    table.close()


def test_unjudged_queries_and_missing_correct_document(metrics):
    ranked, qrels = random_case(7, nq=130, unjudged=0.2)
    assert 0 < len(qrels) < 130
    dev = torch.device("cuda", 0)
    r = torch.from_numpy(ranked).to(dev)
    table = metrics.JudgedTable(qrels, 130, dev, max_k=12)
    got = table.evaluate(r, {"ndcg": 10, "err": 10, "q_measure": 10})     # an unjudged query scores 0 (:229,:281,:339)
    assert got["ndcg"] == pytest.approx(oracle.ndcg_at_k(ranked, qrels, 10), abs=TOL)
    assert got["err"] == pytest.approx(oracle.err_at_k(ranked, qrels, 10), abs=TOL)
    assert got["q_measure"] == pytest.approx(oracle.q_measure_at_k(ranked, qrels, 10), abs=TOL)
    with pytest.raises(StopIteration):                # compare_embeddings.py:111 `next(...)` on a query without a 1
        table.evaluate(r, {"precision": 1})
    with pytest.raises(StopIteration):
        metrics.evaluate_rankings(r, qrels, 3)
    table.close()
    # nothing relevant anywhere: ERR and Q-measure are 0 (compare_embeddings.py:276-279)
    zero = {q: {int(ranked[q, 0]): 0.0} for q in range(130)}
    t0 = metrics.JudgedTable(zero, 130, dev)
    assert t0.max_relevance == 0.0
    assert t0.evaluate(r, {"ndcg": 5, "err": 5, "q_measure": 5}) == {"ndcg": 0.0, "err": 0.0, "q_measure": 0.0}
    t0.close()


def test_metrics_of_a_real_batched_search(ts, metrics):
    """4096 queries x top-10 over 50k rows through K3, judged against planted duplicates: ids stay on the GPU."""
    from theoremsearch_b200 import synthetic
    dev = torch.device("cuda", 0)
    index = ts.TheoremIndex(256, 50_000, dtype="bf16", device=dev)
    synthetic.fill_index(index, 0, 50_000, seed=3)
    q = synthetic.make_queries(4096, 256, dev)
    _, ids = index.search(q, 10)
    host = ids.cpu().numpy()
    rng = np.random.default_rng(0)
    qrels = {}
    for i in range(4096):
        picks = [int(host[i, int(rng.integers(0, 10))]), int(rng.integers(0, 50_000)), int(rng.integers(0, 50_000))]
        qrels[i] = dict(zip(dict.fromkeys(picks), [1.0, 2.0, 0.5]))
    rep = metrics.evaluate_rankings(ids, qrels, 5)
    want = {"P@1": oracle.precision_at_k(host, qrels, 1), "H@5": oracle.hit_at_k(host, qrels, 5),
            "MRR@5": oracle.mrr_at_k(host, qrels, 5), "nDCG@5": oracle.ndcg_at_k(host, qrels, 5),
            "ERR@5": oracle.err_at_k(host, qrels, 5), "Q-measure@5": oracle.q_measure_at_k(host, qrels, 5)}
    assert rep == pytest.approx(want, abs=TOL)
    assert 0.0 < rep["H@5"] < 1.0 and rep["nDCG@5"] > 0.0


def test_bad_tables_and_cuts_are_refused(ts, metrics):
    lib = ts._lib.lib
    h = C.c_void_p()
    offsets = np.array([0, 2], np.int64)
    rels = np.array([1.0, 0.5], np.float64)
    twice = np.array([5, 5], np.int64)
    assert lib.ts_eval_create(C.byref(h), 0, 1, offsets.ctypes.data, twice.ctypes.data, rels.ctypes.data, 8) == -1
    assert "twice" in ts._lib.last_error()
    negative = np.array([5, -2], np.int64)
    assert lib.ts_eval_create(C.byref(h), 0, 1, offsets.ctypes.data, negative.ctypes.data, rels.ctypes.data, 8) == -1
    assert lib.ts_eval_create(C.byref(h), 0, 1, offsets.ctypes.data, twice.ctypes.data, rels.ctypes.data, 65) == -1
    # nDCG@k needs the k best judged relevances: a table cut at max_k refuses a larger k when a query judges more
    dev = torch.device("cuda", 0)
    qrels = {0: {d: 0.5 for d in range(20)}}
    qrels[0][3] = 1.0
    table = metrics.JudgedTable(qrels, 1, dev, max_k=4)
    r = torch.arange(12, device=dev, dtype=torch.int64)[None, :]
    assert table.evaluate(r, {"ndcg": 4})["ndcg"] == pytest.approx(oracle.ndcg_at_k(r.cpu().numpy(), qrels, 4), abs=TOL)
    with pytest.raises(ts.TheoremSearchError, match="max_k"):
        table.evaluate(r, {"ndcg": 8})
    assert table.evaluate(r, {"err": 8})["err"] == pytest.approx(oracle.err_at_k(r.cpu().numpy(), qrels, 8), abs=TOL)
    with pytest.raises(ts.TheoremSearchError):
        table.evaluate(r.cpu(), {"err": 8})
    with pytest.raises(ts.TheoremSearchError):
        table.evaluate(torch.zeros((2, 4), dtype=torch.int64, device=dev), {"err": 3})
    table.close()
