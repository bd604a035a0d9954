"""GPU tests for the rows either side of the search path: the showcase app's
library files as a resident index, pgvector-shaped streams, the latest-slogan join, save/load of the
quantised index (bit-identical results), and the six evaluation metrics from ONE batched top-k."""
import numpy as np
import pytest
import torch

from oracle import oracle
from tests.helpers import TableModel, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ts():
    import theoremsearch_b200 as ts
    assert torch.cuda.is_available()
    return ts


def test_library_files_to_index_and_search_theorems(ts, tmp_path):
    emb = oracle.synthetic_rows(0, 300, 64, seed=8)
    meta = [{"paper_title": f"P{i}", "type": "theorem", "content": f"c{i}"} for i in range(300)]
    ts.save_embedding_library(str(tmp_path), torch.from_numpy(emb), meta)
    index, data = ts.load_embedding_library(str(tmp_path))
    assert len(index) == 300 and data == meta
    q = emb[17] + 0.01 * oracle.synthetic_rows(0, 1, 64, seed=9)[0]
    model = TableModel({"find seventeen": q})
    hits = ts.search_theorems("find seventeen", model, data, index)           # test_app.py:67 call shape
    assert [h["index"] for h in hits][0] == 17 and hits[0]["theorem"] == meta[17]
    want = oracle.search_theorems_topk(torch.from_numpy(q), torch.from_numpy(oracle.bf16_round(oracle.normalize_f64(emb))), 5)
    assert [h["index"] for h in hits] == [int(i) for i in want[0]]


def test_stream_and_latest_slogan_index(ts):
    from theoremsearch_b200 import formats
    d = 32
    emb = oracle.synthetic_rows(0, 40, d, seed=3)
    # cursor-shaped batches: pgvector text rows, (ids, X) pairs and bare arrays mixed
    b1 = (np.arange(100, 110, dtype=np.int64), ["[" + ",".join(repr(float(v)) for v in r) + "]" for r in emb[:10]])
    b2 = (np.arange(110, 140, dtype=np.int64), emb[10:40])
    index = formats.index_from_embedding_stream([b1, b2], d, 40)
    assert len(index) == 40
    assert np.array_equal(index.get_rows().cpu().numpy(), oracle.bf16_round(oracle.normalize_f64(emb)))
    s, i = index.search(torch.from_numpy(emb[33]), 1)
    assert int(i[0, 0]) == 133
    # one embedding per theorem: the latest slogan wins, ids are theorem ids
    theorem = np.array([5, 6, 5, 7, 6, 5] + list(range(100, 134)), dtype=np.int64)
    slogan = np.arange(40, dtype=np.int64)
    ix2 = formats.index_from_slogan_table(theorem, slogan, emb)
    assert len(ix2) == 3 + 34
    s, i = ix2.search(torch.from_numpy(emb[5]), 1)       # slogan 5 is theorem 5's latest
    assert int(i[0, 0]) == 5
    s, i = ix2.search(torch.from_numpy(emb[0]), 3)       # slogan 0 (an old slogan of theorem 5) is not indexed
    kept = oracle.bf16_round(oracle.normalize_f64(emb[[5, 4, 3] + list(range(6, 40))]))
    ref_s, ref_i = oracle.exact_search(oracle.normalize_f64(emb[0:1]), kept, 3,
                                       ids=np.array([5, 6, 7] + list(range(100, 134))))
    assert i.cpu().numpy().tolist() == ref_i.tolist()


@pytest.mark.parametrize("dtype,d,with_ids,ivf", [("bf16", 1024, False, None), ("bf16", 100, True, "fp8"),
                                                   ("f32", 64, True, None), ("bf16", 768, False, "bf16")])
def test_save_load_quantised_index_bit_identical(ts, tmp_path, dtype, d, with_ids, ivf):
    n = 5000
    x = oracle.synthetic_rows(0, n, d, seed=13)
    ids = (np.arange(n, dtype=np.int64) * 7 + 3) if with_ids else None
    a = ts.build_index(x, ids=ids, dtype=dtype)
    if ivf:
        a.ivf_train(20, iters=3, seed=2)
        a.ivf_build(ivf)
    path = str(tmp_path / "corpus.tsidx")
    ts.save_index(a, path)
    b = ts.load_index(path)
    assert len(b) == n and b.dim == d and b.dtype == dtype
    assert torch.equal(a.get_rows(), b.get_rows())
    q = torch.from_numpy(oracle.synthetic_queries(6, d))
    for k in (10, 100):
        sa, ia = a.search(q, k)
        sb, ib = b.search(q, k)
        assert torch.equal(sa, sb) and torch.equal(ia, ib)
    sa, ia = a.search(q[:1], 10)         # single-query (K2) path too
    sb, ib = b.search(q[:1], 10)
    assert torch.equal(sa, sb) and torch.equal(ia, ib)
    if ivf:
        assert b.nlist == 20
        for u, v in zip(a.ivf_lists(), b.ivf_lists()):
            assert torch.equal(u, v)
        sa, ia = a.ivf_search(q, 10, nprobe=5, rescore_k=50)
        sb, ib = b.ivf_search(q, 10, nprobe=5, rescore_k=50)
        assert torch.equal(sa, sb) and torch.equal(ia, ib)
    with open(path, "r+b") as f:
        f.write(b"XXXXXXXX")
    with pytest.raises(ts.TheoremSearchError):
        ts.load_index(path)


def test_evaluate_retrieval_one_search_six_metrics(ts):
    """compare_embeddings.py:55-92 through the CUDA path == the reference's own numbers on the golden set."""
    from theoremsearch_b200 import metrics
    g = load_golden("compare_embeddings_metrics")
    docs = np.array(g["docs"], dtype=np.float32)
    queries = np.array(g["queries"], dtype=np.float32)
    qrels = {int(q): {int(dd): v for dd, v in rd.items()} for q, rd in g["qrels"].items()}
    table = {f"d{i}": docs[i] for i in range(len(docs))}
    table.update({f"q{i}": queries[i] for i in range(len(queries))})

    class BatchModel(TableModel):
        def encode(self, texts, convert_to_tensor=False, **kw):
            if isinstance(texts, str):
                return super().encode(texts, convert_to_tensor=convert_to_tensor, **kw)
            m = np.stack([np.asarray(self.table[t], dtype=np.float32) for t in texts])
            return torch.from_numpy(m) if convert_to_tensor else m

    model = BatchModel(table)
    theorems = [(f"d{i}", "p") for i in range(len(docs))]
    qs = [(f"q{i}", "p") for i in range(len(queries))]
    index = ts.build_index(docs, dtype="f32")          # fp32 storage: the reference's precision
    ranked = metrics.rank_concepts(torch.from_numpy(queries), index, 10)
    assert ranked.tolist() == g["ranked_top10"]
    before = ts.kernel_launches()
    for k in (3, 5):
        rep = metrics.evaluate_retrieval(model, theorems, qs, qrels, top_k_report=k, verbose=False)
        want = g["metrics"][str(k)]
        assert rep["P@1"] == pytest.approx(g["metrics"]["1"]["precision"], abs=1e-12)
        assert rep[f"H@{k}"] == pytest.approx(want["hit"], abs=1e-12)
        assert rep[f"MRR@{k}"] == pytest.approx(want["mrr"], abs=1e-12)
        assert rep[f"nDCG@{k}"] == pytest.approx(want["ndcg"], abs=1e-12)
        assert rep[f"ERR@{k}"] == pytest.approx(want["err"], abs=1e-12)
        assert rep[f"Q-measure@{k}"] == pytest.approx(want["q_measure"], abs=1e-12)
    assert ts.kernel_launches() > before
