"""GPU parity tests for the batched path (K3: tcgen05/TMEM GEMM + fused top-k epilogue, K3b
compaction, K3c fp32 re-score + exactness certificate, K2 fix-up), through the C ABI.
Oracle: fp64 scores of the same inputs (bf16 corpus rows, fp32 normalised queries) — the batched
path must return exactly what the single-query path returns."""
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


def prepared(q_raw):
    return oracle.normalize_f64(q_raw)


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
    check_against_oracle(ts, index, prepared(q_raw), k, s, i)


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
        check_against_oracle(ts, index, prepared(q_raw), k, s, i, ids=ids, allow=allow)


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


@pytest.mark.parametrize("n,d,nq,k", [(30000, 1024, 24, 10), (30000, 768, 70, 100), (5000, 200, 9, 33)])
def test_batched_is_bitwise_the_single_query_path(ts, n, d, nq, k):
    """K3 (tensor cores + re-score) and K2 (one query at a time): identical ids AND identical score bits."""
    rows = unit_rows(n, d, seed=29)
    index = ts.build_index(rows, dtype="bf16", normalize=False)
    q = torch.from_numpy(oracle.synthetic_queries(nq, d, seed=6))
    s3, i3 = index.search(q, k)
    s2 = torch.empty_like(s3)
    i2 = torch.empty_like(i3)
    for j in range(nq):
        s2[j], i2[j] = (x[0] for x in index.search(q[j], k))
    assert torch.equal(i2, i3)
    assert torch.equal(s2, s3)


def test_random_data_is_certified_without_fixups(ts):
    rows = unit_rows(200000, 256, seed=31)
    index = ts.build_index(rows, dtype="bf16", normalize=False)
    q = oracle.synthetic_queries(300, 256, seed=8)
    s, i = index.search(torch.from_numpy(q), 10)
    assert ts.last_batched_fixups() == 0
    check_against_oracle(ts, index, prepared(q), 10, s, i)


def test_near_duplicate_cluster_fails_certificate_and_is_fixed_up(ts):
    """300 rows within 1e-4 of each other at the top of one query's ranking: the bf16-query GEMM
    cannot separate them, the certificate must refuse, and the K2 re-scan must make it exact."""
    n, d, nq, k = 20000, 1024, 16, 10
    rows = unit_rows(n, d, seed=37)
    q_raw = oracle.synthetic_queries(nq, d, seed=11)
    qn = oracle.normalize_f64(q_raw)
    rng = np.random.default_rng(0)
    cluster = rng.choice(n, size=300, replace=False)
    for r in cluster:
        v = 0.9 * qn[3] + 0.436 * oracle.normalize_f64(rng.standard_normal(d).astype(np.float32))[0] * 1.0
        rows[r] = oracle.normalize_f64(v + 1e-4 * rng.standard_normal(d).astype(np.float32))[0]
    # make the cluster REALLY tight in score: same direction, tiny jitter
    base = oracle.normalize_f64(0.9 * qn[3] + 0.1 * rows[cluster[0]])[0]
    for r in cluster:
        rows[r] = oracle.normalize_f64(base + 2e-5 * rng.standard_normal(d).astype(np.float32))[0]
    index = ts.build_index(rows, dtype="bf16", normalize=False)
    s, i = index.search(torch.from_numpy(q_raw), k)
    fix = ts.last_batched_fixups()
    assert fix >= 1
    check_against_oracle(ts, index, prepared(q_raw), k, s, i)
    s1, i1 = index.search(torch.from_numpy(q_raw[3]), k)
    assert torch.equal(i[3], i1[0]) and torch.equal(s[3], s1[0])


def test_ascending_order_overflows_buffers_and_is_fixed_up(ts):
    """Rows sorted by ascending score for query 0: every row beats the running threshold, the
    candidate buffer overflows, the query is flagged and re-scanned exactly."""
    n, d, nq, k = 40000, 256, 8, 10
    rows = unit_rows(n, d, seed=41)
    q_raw = oracle.synthetic_queries(nq, d, seed=12)
    qn = oracle.normalize_f64(q_raw)
    order = np.argsort(oracle.bf16_round(rows) @ qn[0])
    rows = rows[order]
    index = ts.build_index(rows, dtype="bf16", normalize=False)
    ts.set_tunable("batch.dense", 0)     # a corpus this small would take the dense path, which cannot overflow
    try:
        s, i = index.search(torch.from_numpy(q_raw), k)
        assert ts.last_batched_fixups() >= 1
    finally:
        ts.set_tunable("batch.dense", 1)
    check_against_oracle(ts, index, prepared(q_raw), k, s, i)
    s_d, i_d = index.search(torch.from_numpy(q_raw), k)   # the dense path on the same adversarial order
    assert torch.equal(s_d, s) and torch.equal(i_d, i)


@pytest.mark.parametrize("n,d,nq,k", [
    (5000, 1024, 100, 10), (20000, 1024, 64, 100), (4097, 768, 33, 10), (3000, 100, 16, 5),
    (40000, 256, 257, 20), (127, 1024, 5, 10), (129, 512, 600, 1), (70000, 1024, 1000, 10),
])
def test_cta_pair_kernel_matches_oracle(ts, n, d, nq, k):
    """Force the cta_group::2 (CTA pair) GEMM for every batch size, including odd tile counts."""
    old = ts.get_tunable("batch.pair_min_nq")
    old_pair = ts.get_tunable("batch.cta_pair")
    ts.set_tunable("batch.pair_min_nq", 1)
    ts.set_tunable("batch.cta_pair", 1)
    try:
        rows = unit_rows(n, d, seed=n + d + 2)
        index = ts.build_index(rows, dtype="bf16", normalize=False)
        q_raw = oracle.synthetic_queries(nq, d, seed=400 + nq)
        s, i = index.search(torch.from_numpy(q_raw), k)
        fix = ts.last_batched_fixups()
        check_against_oracle(ts, index, prepared(q_raw), k, s, i)
        assert fix <= max(2, nq // 50), f"{fix} of {nq} queries needed the K2 fix-up: the pair GEMM is mis-scoring"
    finally:
        ts.set_tunable("batch.pair_min_nq", old)
        ts.set_tunable("batch.cta_pair", old_pair)


def test_dense_small_corpus_path_equals_chunked_path(ts):
    """Small corpora (N <= 2^18 rows and nq * N * 4 B <= 1 GiB) take one dense GEMM pass + a select kernel instead of the chunked
    threshold filter; both feed the same re-score / certificate, so results are bit-identical."""
    x = oracle.synthetic_rows(0, 30000, 1024, seed=91)
    x[29999] = x[7]
    index = ts.build_index(x)
    q = torch.from_numpy(oracle.synthetic_queries(70, 1024, seed=92))
    allow = np.random.default_rng(1).random(30000) < 0.5
    mask = ts.pack_allow_mask(allow, index.device)
    for k in (10, 100, 300):
        s_d, i_d = index.search(q, k)
        s_dm, i_dm = index.search(q, k, allow_mask=mask)
        ts.set_tunable("batch.dense", 0)
        try:
            s_c, i_c = index.search(q, k)
            s_cm, i_cm = index.search(q, k, allow_mask=mask)
        finally:
            ts.set_tunable("batch.dense", 1)
        assert torch.equal(s_d, s_c) and torch.equal(i_d, i_c)
        assert torch.equal(s_dm, s_cm) and torch.equal(i_dm, i_cm)
    s10, i10 = index.search(q, 10)
    check_against_oracle(ts, index, oracle.normalize_f64(q.numpy()), 10, s10, i10)


# ------------------------------------------------------------------------------------------ fp32 corpora (TF32 GEMM)
@pytest.mark.parametrize("n,d,nq,k", [
    (5000, 1024, 4, 10), (20000, 1024, 73, 10), (4097, 768, 33, 100), (3000, 100, 16, 5), (2000, 8, 8, 3),
    (300_000, 256, 40, 10),      # > 2^18 rows: the chunked threshold path, not the dense small-corpus path
    (127, 1024, 5, 10), (1000, 1024, 20, 1000),
])
def test_fp32_corpus_batches_take_the_tf32_gemm_and_stay_exact(ts, n, d, nq, k):
    """An fp32-stored corpus is the reference's own precision (pg `vector(1024)`, rds_schema.sql:50-53;
    compare_embeddings.py:61 on fp32 tensors). Batches now read the rows ONCE through tcgen05 kind::tf32; the
    candidates are re-scored in fp32 in K2's summation order and certified against the TF32 truncation bound, so the
    result must be the single-query scan's bit for bit, and the fp64 oracle's within 1e-5."""
    rows = unit_rows(n, d, seed=n + d + 2)
    index = ts.build_index(rows, dtype="f32", normalize=False)
    q_raw = oracle.synthetic_queries(nq, d, seed=500 + nq)
    before = ts.kernel_launches()
    s, i = index.search(torch.from_numpy(q_raw), k, normalize=True)
    torch.cuda.synchronize()
    assert ts.kernel_launches() - before >= 6, "the GEMM path is a chain of kernels; the per-query scan is one launch"
    assert ts.last_batched_fixups() >= 0
    check_against_oracle(ts, index, prepared(q_raw), k, s, i)
    for j in sorted({0, nq // 2, nq - 1}):
        s1, i1 = index.search(torch.from_numpy(q_raw[j]), k, normalize=True)       # K2 on fp32 rows
        assert torch.equal(s[j], s1[0]) and torch.equal(i[j], i1[0])
    ts.set_tunable("batch.tf32", 0)                                                # round 1's path: nq K2 passes
    try:
        before = ts.kernel_launches()
        s0, i0 = index.search(torch.from_numpy(q_raw), k, normalize=True)
        assert ts.kernel_launches() - before <= 2
    finally:
        ts.set_tunable("batch.tf32", 1)
    assert torch.equal(s, s0) and torch.equal(i, i0)


def test_fp32_corpus_near_duplicates_fail_the_tf32_certificate_and_are_fixed_up(ts):
    """Rows that differ from each other by less than the TF32 truncation error cannot be ranked by the GEMM:
    the certificate must send those queries to the exact re-scan, and the result must still be exact."""
    d, n = 1024, 6000
    rows = unit_rows(n, d, seed=77)
    base = rows[100].copy()
    rng = np.random.default_rng(1)
    for r in range(200, 500):                         # 300 rows within ~1e-4 of each other
        v = base + 1e-4 * rng.standard_normal(d).astype(np.float32)
        rows[r] = v / np.linalg.norm(v)
    index = ts.build_index(rows, dtype="f32", normalize=False)
    q = np.stack([base, rows[300], unit_rows(1, d, seed=5)[0]])
    s, i = index.search(torch.from_numpy(q), 10, normalize=True)
    assert ts.last_batched_fixups() >= 2
    check_against_oracle(ts, index, prepared(q), 10, s, i)
