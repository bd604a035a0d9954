"""GPU parity tests for the batched path (K3: tcgen05/TMEM GEMM + fused top-k epilogue, K3b
compaction), through the C ABI.  Oracle: fp64 scores of the SAME inputs the tensor cores see
(bf16 corpus rows, bf16-rounded normalised queries), (score desc, row asc) order."""
import numpy as np
import pytest
import torch

from oracle import oracle
from tests.test_gpu_exact import check_against_oracle, unit_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ts():
    import theoremsearch_b200 as ts
    return ts


def prepared_bf16(q_raw):
    return oracle.bf16_round(oracle.normalize_f64(q_raw))


@pytest.mark.parametrize("n,d,nq,k", [
    (5000, 1024, 4, 10), (5000, 1024, 100, 10), (5000, 1024, 300, 100), (20000, 1024, 64, 100),
    (4097, 768, 33, 10), (3000, 100, 16, 5), (2000, 8, 8, 3), (40000, 256, 257, 20), (1000, 1024, 20, 1000),
    (127, 1024, 5, 10), (129, 512, 512, 1),
])
def test_batched_matches_oracle(ts, n, d, nq, k):
    rows = unit_rows(n, d, seed=n + d + 1)
    index = ts.build_index(rows, dtype="bf16", normalize=False)
    q_raw = oracle.synthetic_queries(nq, d, seed=300 + nq)
    assert nq >= ts.get_tunable("batch.min_nq")
    s, i = index.search(torch.from_numpy(q_raw), k, normalize=True)
    check_against_oracle(ts, index, prepared_bf16(q_raw), k, s, i)


def test_batched_with_mask_and_ids(ts):
    n, d, nq, k = 9000, 1024, 40, 10
    rows = unit_rows(n, d, seed=17)
    ids = np.arange(n, dtype=np.int64) * 3 + 11
    index = ts.TheoremIndex(d, n)
    index.add(rows, ids=ids, normalize=False)
    q_raw = oracle.synthetic_queries(nq, d, seed=9)
    rng = np.random.default_rng(3)
    for frac in (0.3, 0.002):
        allow = rng.random(n) < frac
        mask = ts.pack_allow_mask(allow, index.device)
        s, i = index.search(torch.from_numpy(q_raw), k, allow_mask=mask)
        check_against_oracle(ts, index, prepared_bf16(q_raw), k, s, i, ids=ids, allow=allow)


def test_batched_duplicates_tie_rule(ts):
    n, d, nq = 6000, 1024, 8
    rows = unit_rows(n, d, seed=23)
    q_raw = oracle.synthetic_queries(nq, d, seed=5)
    qn = oracle.normalize_f64(q_raw)
    dup = [5000, 12, 3333, 640, 4097]
    for r in dup:
        rows[r] = qn[2]                     # identical rows, cosine 1 with query 2
    index = ts.build_index(rows, dtype="bf16", normalize=False)
    s, i = index.search(torch.from_numpy(q_raw), 5)
    assert i[2].tolist() == sorted(dup)
    assert len(set(s[2].tolist())) == 1


def test_batched_equals_looped_single_query_on_bf16_queries(ts):
    """Same inputs through K2 (one query at a time, fp32 FMA) and K3 (tensor cores): same ids."""
    n, d, nq, k = 30000, 1024, 24, 10
    rows = unit_rows(n, d, seed=29)
    index = ts.build_index(rows, dtype="bf16", normalize=False)
    q = torch.from_numpy(prepared_bf16(oracle.synthetic_queries(nq, d, seed=6)))
    s3, i3 = index.search(q, k, normalize=False)
    s2 = torch.empty_like(s3)
    i2 = torch.empty_like(i3)
    for j in range(nq):
        s2[j], i2[j] = (x[0] for x in index.search(q[j], k, normalize=False))
    assert torch.equal(i2, i3)
    assert torch.allclose(s2, s3, atol=2e-6)
