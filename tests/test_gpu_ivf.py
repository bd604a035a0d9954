"""GPU parity tests for the IVF-Flat path (K4a assignment GEMM, list build + e4m3 quantisation,
K4b list scan, K4c re-score), through the C ABI.  Oracle: oracle.ivf_assign / oracle.ivf_search /
oracle.quantize_fp8_e4m3 fed with the index's OWN centroids and stored rows ("same inputs").

Bars: list layout and e4m3 quantisation bit-exact; assignment equal to the fp64 arg-max except
where the two best centroids are closer than ASSIGN_EPS (fp32 tensor-core accumulation);
bf16-list search equal to the oracle's exact top-k over the probed lists (same tie window as the
exact path); fp8-list search: returned scores exact (re-scored) and recall vs the oracle's
probed-list top-k >= 0.99 (BASELINE.json reports IVF as recall@10)."""
import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-5
TIE_EPS = 2e-6
ASSIGN_EPS = 2e-5


@pytest.fixture(scope="module")
def ts():
    import theoremsearch_b200 as ts
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    # the list-major batch path (K4d) normally wants >= 4 x SMs lists; the small test indexes must reach it too
    ts.set_tunable("ivf.group_min_lists", 1)
    yield ts
    ts.set_tunable("ivf.group_min_lists", 0)


def clustered_rows(n, d, ncenters, sigma, seed):
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((ncenters, d)).astype(np.float32)
    centers /= np.linalg.norm(centers, axis=1, keepdims=True)
    which = rng.integers(0, ncenters, n)
    x = centers[which] + sigma * rng.standard_normal((n, d)).astype(np.float32) / np.sqrt(d)
    return x.astype(np.float32)


def built(ts, rows, nlist, list_dtype, iters=4, seed=3, ids=None):
    index = ts.build_index(rows, dtype="bf16", normalize=True, ids=ids)
    index.ivf_train(nlist, iters=iters, seed=seed)
    index.ivf_build(list_dtype)
    return index


def layout(index):
    off, rows = index.ivf_lists()
    return off.cpu().numpy(), rows.cpu().numpy()


def assignment_from_lists(off, rows, n):
    assign = np.empty(n, dtype=np.int64)
    for l in range(len(off) - 1):
        assign[rows[off[l]:off[l + 1]]] = l
    return assign


@pytest.mark.parametrize("n,d,nlist", [(5000, 1024, 37), (3000, 768, 256), (2000, 100, 300), (700, 64, 5)])
def test_assignment_and_list_layout(ts, n, d, nlist):
    x = clustered_rows(n, d, 50, 0.7, seed=n + d)
    index = built(ts, x, nlist, "bf16")
    assert index.nlist == nlist
    off, rows = layout(index)
    # layout: offsets ascending from 0 to n, rows a permutation, ascending inside each list
    assert off[0] == 0 and off[-1] == n and np.all(np.diff(off) >= 0)
    assert np.array_equal(np.sort(rows), np.arange(n))
    for l in range(nlist):
        seg = rows[off[l]:off[l + 1]]
        assert np.all(np.diff(seg) > 0)
    assert np.array_equal(index.ivf_list_sizes().cpu().numpy(), np.diff(off))
    # assignment: nearest centroid of the bf16 centroid table the GEMM reads
    stored = index.get_rows().cpu().numpy()
    cent = oracle.bf16_round(index.ivf_centroids().cpu().numpy())
    s = stored.astype(np.float64) @ cent.astype(np.float64).T
    want = np.argmax(s, axis=1)
    got = assignment_from_lists(off, rows, n)
    bad = np.nonzero(got != want)[0]
    for r in bad:   # only near-ties may differ
        assert s[r, want[r]] - s[r, got[r]] <= ASSIGN_EPS, (r, got[r], want[r], s[r, want[r]], s[r, got[r]])
    assert bad.size <= max(2, n // 200)
    # centroids are unit vectors (spherical k-means)
    cn = np.linalg.norm(index.ivf_centroids().cpu().numpy().astype(np.float64), axis=1)
    assert np.max(np.abs(cn - 1.0)) < 1e-2


@pytest.mark.parametrize("d", [1024, 768, 100])
def test_fp8_list_quantisation_bit_exact(ts, d):
    x = clustered_rows(3000, d, 20, 0.8, seed=d)
    x[17] = 0.0   # all-zero row: scale 1, all zeros
    index = built(ts, x, 16, "fp8")
    off, rows = layout(index)
    stored = index.get_rows().cpu().numpy()
    want, _ = oracle.quantize_fp8_e4m3(stored[rows])
    got = index.ivf_list_data().cpu().numpy()
    assert np.array_equal(got, want)
    # bf16 lists hold the stored rows verbatim
    index.ivf_build("bf16")
    off, rows = layout(index)
    assert np.array_equal(index.ivf_list_data().cpu().numpy(), stored[rows])


def oracle_probed_topk(index, q_prepared, k, nprobe, skip_eps=1e-6):
    """oracle.ivf_search per query on the index's own centroids/lists. Returns list of
    (scores, ids) or None where the probe set itself sits on a near-tie."""
    stored = index.get_rows().cpu().numpy()
    cent = oracle.bf16_round(index.ivf_centroids().cpu().numpy())
    off, rows = layout(index)
    assign = assignment_from_lists(off, rows, stored.shape[0])
    out = []
    for q in q_prepared:
        cs = np.sort(cent.astype(np.float64) @ q.astype(np.float64))[::-1]
        if nprobe < len(cs) and cs[nprobe - 1] - cs[nprobe] < skip_eps:
            out.append(None)
            continue
        out.append(oracle.ivf_search(q, stored, cent, assign, k, nprobe))
    return out, stored


@pytest.mark.parametrize("n,d,nlist,nprobe,k", [(6000, 1024, 64, 8, 10), (4000, 768, 40, 5, 20), (3000, 96, 100, 100, 5),
                                               (2500, 1024, 30, 1, 10), (5000, 256, 33, 7, 100)])
def test_bf16_lists_match_oracle_over_probed_lists(ts, n, d, nlist, nprobe, k):
    x = clustered_rows(n, d, 40, 0.9, seed=7 * n)
    index = built(ts, x, nlist, "bf16")
    q = oracle.normalize_f64(clustered_rows(12, d, 40, 0.9, seed=7 * n))   # same centres: realistic queries
    s, i = index.ivf_search(torch.from_numpy(q), k, nprobe=nprobe, rescore_k=k, normalize=False)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    refs, stored = oracle_probed_topk(index, q, k, nprobe)
    checked = 0
    for qi, ref in enumerate(refs):
        if ref is None:
            continue
        checked += 1
        rs, ri = ref
        valid = ri >= 0
        assert np.array_equal(i[qi] >= 0, valid)
        g, r = i[qi][valid], ri[valid]
        alls = stored.astype(np.float64) @ q[qi].astype(np.float64)
        if not np.array_equal(g, r):
            assert len(set(g.tolist())) == g.size
            assert np.all(np.abs(alls[g] - alls[r]) <= TIE_EPS), (qi, g, r)
        assert np.max(np.abs(s[qi][valid] - alls[g]), initial=0.0) <= SCORE_TOL
    assert checked >= 8


def test_probing_every_list_is_exact_search(ts):
    x = clustered_rows(8000, 1024, 30, 1.0, seed=5)
    index = built(ts, x, 48, "bf16")
    q = torch.from_numpy(oracle.synthetic_queries(9, 1024))
    s_e, i_e = index.search(q, 10)
    s_a, i_a = index.ivf_search(q, 10, nprobe=48, rescore_k=10)
    assert torch.equal(i_e, i_a)
    assert torch.equal(s_e, s_a)   # re-scored in K2's summation order: bit-identical


def test_fp8_lists_recall_and_exact_scores(ts):
    n, d, nlist, nprobe, k = 30000, 1024, 64, 8, 10
    x = clustered_rows(n, d, 200, 0.9, seed=21)
    index = built(ts, x, nlist, "fp8")
    q = oracle.normalize_f64(clustered_rows(40, d, 200, 0.9, seed=21))
    s, i = index.ivf_search(torch.from_numpy(q), k, nprobe=nprobe, rescore_k=100, normalize=False)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    refs, stored = oracle_probed_topk(index, q, k, nprobe)
    hits = tot = 0
    for qi, ref in enumerate(refs):
        alls = stored.astype(np.float64) @ q[qi].astype(np.float64)
        g = i[qi]
        assert np.all(g >= 0) and len(set(g.tolist())) == k
        assert np.max(np.abs(s[qi] - alls[g])) <= SCORE_TOL        # scores are the exact (re-scored) ones
        assert np.all(np.diff(s[qi]) <= 0)
        if ref is None:
            continue
        hits += len(set(g.tolist()) & set(ref[1].tolist()))
        tot += k
    assert hits / tot >= 0.99, hits / tot
    # and against the unrestricted exact search the probe budget decides: report, loosely bounded
    _, i_exact = index.search(torch.from_numpy(q), k, normalize=False)
    assert oracle.recall_at_k(i, i_exact.cpu().numpy()) >= 0.6


@pytest.mark.parametrize("list_dtype", ["bf16", "fp8"])
def test_batched_queries_equal_single_queries(ts, list_dtype):
    x = clustered_rows(20000, 1024, 100, 0.9, seed=33)
    index = built(ts, x, 128, list_dtype)
    q = torch.from_numpy(oracle.normalize_f64(clustered_rows(70, 1024, 100, 0.9, seed=33)))
    s_b, i_b = index.ivf_search(q, 10, nprobe=16, rescore_k=64)     # coarse step on the tcgen05 path
    for qi in range(0, 70, 7):
        s_1, i_1 = index.ivf_search(q[qi], 10, nprobe=16, rescore_k=64)   # coarse step on the scan path
        assert torch.equal(i_b[qi], i_1[0]) and torch.equal(s_b[qi], s_1[0])


def test_empty_lists_ids_and_keys(ts):
    base = clustered_rows(40, 128, 4, 0.2, seed=1)
    x = np.repeat(base, 25, axis=0)                  # 1000 rows, only 40 distinct -> many empty lists
    ids = np.arange(1000, dtype=np.int64) * 3 + 11
    index = built(ts, x, 200, "fp8", ids=ids)
    off, rows = layout(index)
    assert (np.diff(off) == 0).sum() >= 100
    q = torch.from_numpy(oracle.normalize_f64(base[:5]))
    s, i = index.ivf_search(q, 30, nprobe=200, rescore_k=200, normalize=False)
    s_e, i_e = index.search(q, 30, normalize=False)
    assert torch.equal(s, s_e) and torch.equal(i, i_e)       # duplicates: ties -> lower row, via caller ids
    keys = index.ivf_search_keys(q, 30, nprobe=200, rescore_k=200, normalize=False)
    s_m, i_m = ts.merge_topk(keys.unsqueeze(0), 30)
    assert torch.equal(s_m, s)
    assert torch.equal(i_m * 3 + 11, i)
    # fewer eligible rows than k: padding
    s2, i2 = index.ivf_search(q[:1], 50, nprobe=1, rescore_k=50, normalize=False)
    n_in_list = int((i2[0] >= 0).sum())
    assert 0 < n_in_list <= 50
    assert torch.all(torch.isneginf(s2[0][n_in_list:])) and torch.all(i2[0][n_in_list:] == -1)


def test_shared_centroids_and_state_errors(ts):
    x = clustered_rows(4000, 256, 30, 0.8, seed=9)
    a = built(ts, x, 32, "bf16")
    b = ts.build_index(x, dtype="bf16", normalize=True)
    with pytest.raises(ts.TheoremSearchError) as e:
        b.ivf_search(torch.zeros(1, 256), 5)
    assert e.value.code == -6
    with pytest.raises(ts.TheoremSearchError) as e:
        b.ivf_build("bf16")
    assert e.value.code == -6
    b.ivf_set_centroids(a.ivf_centroids())
    b.ivf_build("bf16")
    for u, v in zip(layout(a), layout(b)):
        assert np.array_equal(u, v)
    # explicit sample training
    c = ts.build_index(x, dtype="bf16", normalize=True, capacity=4010)
    c.ivf_train(32, sample=torch.from_numpy(x[::2]).cuda(), iters=3, seed=1)
    c.ivf_build("fp8")
    assert int(c.ivf_list_sizes().sum()) == 4000
    # adding rows keeps the lists valid (pgvector's ivfflat accepts inserts after its build): the new rows sit in
    # overflow lists and are found; the full story is tests/test_gpu_mutation.py
    c.add(x[:10])
    assert c.ivf_pending() == (10, 0) and int(c.ivf_list_sizes().sum()) == 4010
    s, i = c.ivf_search(torch.from_numpy(x[3:4]), 2, nprobe=32, rescore_k=16)
    assert i[0].tolist() == [3, 4003] and s[0, 0] == s[0, 1]      # the original row and its added duplicate
    with pytest.raises(ts.TheoremSearchError):                    # the packed layout is not exportable while mutated
        c.ivf_lists()


def test_fp8_exhaustive_scan_mode(ts):
    """north_star's 'optionally fp8-e4m3, rescored in fp32' scan: one list with every row in row order."""
    x = clustered_rows(20000, 1024, 50, 1.0, seed=77)
    index = ts.build_index(x)
    index.build_fp8_shadow()
    off, rows = layout(index)
    assert off.tolist() == [0, 20000] and np.array_equal(rows, np.arange(20000))
    q = torch.from_numpy(oracle.normalize_f64(clustered_rows(8, 1024, 50, 1.0, seed=77)))
    s_e, i_e = index.search(q, 10, normalize=False)
    s_f, i_f = index.search_fp8(q, 10, rescore_k=200, normalize=False)
    assert oracle.recall_at_k(i_f.cpu().numpy(), i_e.cpu().numpy()) >= 0.95
    se, ie, sf, i_f = s_e.cpu().numpy(), i_e.cpu().numpy(), s_f.cpu().numpy(), i_f.cpu().numpy()
    for qi in range(8):          # every returned row carries its exact score
        exact = dict(zip(ie[qi].tolist(), se[qi].tolist()))
        for r, sc in zip(i_f[qi].tolist(), sf[qi].tolist()):
            if r in exact:
                assert sc == exact[r]


def test_build_index_fp8_dtype_scans_e4m3_and_rescores_exactly(ts):
    """``build_index(..., dtype='bf16' or 'fp8')`` — the fp8 index answers ``cos_sim_topk`` from the
    e4m3 copy with exact re-scored scores; a k beyond the candidate budget takes the exact bf16 scan."""
    x = clustered_rows(20000, 1024, 50, 1.0, seed=78)
    exact = ts.build_index(x)
    idx8 = ts.build_index(x, dtype="fp8")
    assert idx8.scan_dtype == "fp8" and idx8.dtype == "bf16" and idx8.nlist == 1 and len(idx8) == 20000
    q = torch.from_numpy(oracle.normalize_f64(clustered_rows(8, 1024, 50, 1.0, seed=79)))
    s_e, i_e = ts.cos_sim_topk(q, exact, 10, normalize_queries=False)
    s_f, i_f = ts.cos_sim_topk(q, idx8, 10, normalize_queries=False)
    assert oracle.recall_at_k(i_f.cpu().numpy(), i_e.cpu().numpy()) >= 0.95
    se, ie, sf, jf = s_e.cpu().numpy(), i_e.cpu().numpy(), s_f.cpu().numpy(), i_f.cpu().numpy()
    for qi in range(8):
        exact_of = dict(zip(ie[qi].tolist(), se[qi].tolist()))
        assert all(sc == exact_of[r] for r, sc in zip(jf[qi].tolist(), sf[qi].tolist()) if r in exact_of)
    s1, i1 = ts.cos_sim_topk(q[0], idx8, 10, normalize_queries=False)
    assert i1.shape == (10,) and torch.equal(i1, i_f[0]) and torch.equal(s1, s_f[0])
    s_big, i_big = ts.cos_sim_topk(q, idx8, 200, normalize_queries=False)
    s_big_e, i_big_e = ts.cos_sim_topk(q, exact, 200, normalize_queries=False)
    assert torch.equal(i_big, i_big_e) and torch.equal(s_big, s_big_e)
    with pytest.raises(ts.TheoremSearchError):
        ts.build_index(x[:0], dtype="fp8")


@pytest.mark.parametrize("list_dtype", ["bf16", "fp8"])
def test_filtered_ivf_search(ts, list_dtype):
    """The SQL WHERE of streamlit_app.py:175-243 applied inside the list scan: exact top-k among the ELIGIBLE
    rows of the probed lists."""
    n, d, nlist, nprobe, k = 12000, 512, 48, 6, 10
    x = clustered_rows(n, d, 60, 0.9, seed=91)
    index = built(ts, x, nlist, list_dtype)
    rng = np.random.default_rng(5)
    allow = rng.random(n) < 0.3
    mask = ts.pack_allow_mask(allow, index.device)
    q = oracle.normalize_f64(clustered_rows(10, d, 60, 0.9, seed=91))
    s, i = index.ivf_search(torch.from_numpy(q), k, nprobe=nprobe, rescore_k=200, normalize=False, allow_mask=mask)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    stored = index.get_rows().cpu().numpy()
    cent = oracle.bf16_round(index.ivf_centroids().cpu().numpy())
    off, rows = layout(index)
    assign = assignment_from_lists(off, rows, n)
    hits = tot = 0
    for qi in range(q.shape[0]):
        cs = cent.astype(np.float64) @ q[qi].astype(np.float64)
        order = np.sort(cs)[::-1]
        if order[nprobe - 1] - order[nprobe] < 1e-6:
            continue
        probe = oracle.rank_desc(cs, nprobe)
        member = np.isin(assign, probe) & allow
        ref_s, ref_i = oracle.exact_search(q[qi], stored, k, allow=member)
        got = i[qi][i[qi] >= 0]
        assert np.all(allow[got]) and np.all(np.isin(assign[got], probe))
        alls = stored.astype(np.float64) @ q[qi].astype(np.float64)
        assert np.max(np.abs(s[qi][i[qi] >= 0] - alls[got]), initial=0.0) <= SCORE_TOL
        if list_dtype == "bf16":
            assert np.array_equal(i[qi], ref_i[0]) or np.all(np.abs(alls[got] - alls[ref_i[0][ref_i[0] >= 0]]) <= TIE_EPS)
        hits += len(set(got.tolist()) & set(ref_i[0][ref_i[0] >= 0].tolist()))
        tot += int((ref_i[0] >= 0).sum())
    assert tot > 0 and hits / tot >= 0.99
    # an empty filter returns padding only
    none = ts.pack_allow_mask(np.zeros(n, dtype=bool), index.device)
    s0, i0 = index.ivf_search(torch.from_numpy(q[:2]), k, nprobe=nprobe, rescore_k=50, normalize=False, allow_mask=none)
    assert torch.all(i0 == -1) and torch.all(torch.isneginf(s0))


def test_list_major_batch_path_and_its_fallback(ts):
    """K4d (batches >= 64 queries over e4m3 lists) against the per-query scan K4b: same top-k after the exact
    re-score.  Then a corpus so skewed that the batch's score buffer cannot hold it: the device-side flag must
    route the batch to K4b and the result must not change."""
    x = clustered_rows(30000, 1024, 80, 0.9, seed=123)
    index = built(ts, x, 96, "fp8")
    q = torch.from_numpy(oracle.normalize_f64(clustered_rows(200, 1024, 80, 0.9, seed=123)))
    s_g, i_g = index.ivf_search(q, 10, nprobe=12, rescore_k=100)          # list-major
    old = ts.get_tunable("ivf.group_min_nq")
    try:
        ts.set_tunable("ivf.group_min_nq", 0)
        index._ws = {}
        s_c, i_c = index.ivf_search(q, 10, nprobe=12, rescore_k=100)      # per-query scan
    finally:
        ts.set_tunable("ivf.group_min_nq", old)
        index._ws = {}
    assert oracle.recall_at_k(i_g.cpu().numpy(), i_c.cpu().numpy()) >= 0.995
    same = (i_g == i_c).all(dim=1)
    assert torch.equal(s_g[same], s_c[same])                                # exact (re-scored) scores either way
    # skewed: almost every row in one list, every query probes it -> nq * len >> 2 * nq * nprobe * avg_len
    rng = np.random.default_rng(3)
    base = rng.standard_normal(1024).astype(np.float32)
    big = base[None, :] + 0.05 * rng.standard_normal((24000, 1024)).astype(np.float32)
    rest = clustered_rows(1000, 1024, 60, 0.5, seed=9)
    xs = np.concatenate([rest, big])
    skew = built(ts, xs, 64, "fp8", iters=3)
    sizes = skew.ivf_list_sizes().cpu().numpy()
    assert sizes.max() > 20000
    qs = torch.from_numpy(oracle.normalize_f64(big[:128] + 0.01 * rng.standard_normal((128, 1024)).astype(np.float32)))
    s1, i1 = skew.ivf_search(qs, 10, nprobe=2, rescore_k=128)             # flag trips -> K4b does the work
    for j in (0, 17, 127):
        s0, i0 = skew.ivf_search(qs[j], 10, nprobe=2, rescore_k=128)      # single query: K4b directly
        assert torch.equal(i1[j], i0[0]) and torch.equal(s1[j], s0[0])


@pytest.mark.parametrize("d", [768, 100, 264])
def test_list_major_batch_path_other_dims(ts, d):
    x = clustered_rows(9000, d, 40, 0.9, seed=d)
    index = built(ts, x, 32, "fp8")
    q = torch.from_numpy(oracle.normalize_f64(clustered_rows(80, d, 40, 0.9, seed=d)))
    s_g, i_g = index.ivf_search(q, 10, nprobe=32, rescore_k=200)          # K4d, every list probed
    s_e, i_e = index.search(q, 10)                                         # exact
    assert oracle.recall_at_k(i_g.cpu().numpy(), i_e.cpu().numpy()) >= 0.99
    same = (i_g == i_e).all(dim=1)
    assert same.sum() >= 70 and torch.equal(s_g[same], s_e[same])


def test_few_lists_keep_the_per_query_scan(ts):
    """One-list index (the fp8 shadow): a batch must not be handed to the list-major kernel (one CTA per query
    group would scan the whole corpus); with the default policy it stays on K4b and matches single queries."""
    ts.set_tunable("ivf.group_min_lists", 0)
    try:
        x = clustered_rows(20000, 512, 30, 1.0, seed=4)
        index = ts.build_index(x)
        index.build_fp8_shadow()
        q = torch.from_numpy(oracle.normalize_f64(clustered_rows(32, 512, 30, 1.0, seed=4)))
        before = ts.kernel_launches()
        s_b, i_b = index.search_fp8(q, 10, rescore_k=64)
        launches = ts.kernel_launches() - before
        s_1, i_1 = index.search_fp8(q[5], 10, rescore_k=64)
        assert torch.equal(i_b[5], i_1[0]) and torch.equal(s_b[5], s_1[0])
        assert launches <= 30      # coarse (K3 chunks) + K4b + re-score: none of K4d's table/scan/select kernels
    finally:
        ts.set_tunable("ivf.group_min_lists", 1)


def test_filtered_search_through_the_list_major_path(ts):
    """The allow mask inside K4d (scores of disallowed rows become -inf and never reach the selection)."""
    n, d = 20000, 1024
    x = clustered_rows(n, d, 60, 0.9, seed=55)
    index = built(ts, x, 64, "fp8")
    rng = np.random.default_rng(8)
    allow = rng.random(n) < 0.25
    mask = ts.pack_allow_mask(allow, index.device)
    q = torch.from_numpy(oracle.normalize_f64(clustered_rows(96, d, 60, 0.9, seed=55)))
    s_g, i_g = index.ivf_search(q, 10, nprobe=8, rescore_k=128, allow_mask=mask)           # K4d (96 queries)
    got = i_g.cpu().numpy()
    assert np.all(allow[got[got >= 0]])
    for j in (0, 31, 95):
        s_1, i_1 = index.ivf_search(q[j], 10, nprobe=8, rescore_k=128, allow_mask=mask)   # K4b
        assert torch.equal(i_g[j], i_1[0]) and torch.equal(s_g[j], s_1[0])


def test_ivf_host_buffer_entry_point(ts):
    """ts_ivf_search_host: numpy in, numpy out, same result as the device-tensor call."""
    x = clustered_rows(15000, 1024, 40, 0.9, seed=66)
    index = built(ts, x, 64, "fp8")
    q = oracle.normalize_f64(clustered_rows(20, 1024, 40, 0.9, seed=66))
    s_d, i_d = index.ivf_search(torch.from_numpy(q), 10, nprobe=8, rescore_k=100, normalize=False)
    s_h, i_h = index.ivf_search_host(q, 10, nprobe=8, rescore_k=100, normalize=False)
    assert np.array_equal(i_h, i_d.cpu().numpy()) and np.array_equal(s_h, s_d.cpu().numpy())
    s_1, i_1 = index.ivf_search_host(q[3], 10, nprobe=8, rescore_k=100, normalize=False)
    assert np.array_equal(i_1[0], i_h[3])


@pytest.mark.parametrize("mode,d", [(3, 1024), (4, 1024), (3, 768), (3, 100), (4, 264), (1, 1024), (0, 1024)])
def test_list_major_scoring_variants_agree_with_the_per_query_scan(ts, mode, d):
    """K4d's scoring back ends — tcgen05 kind::f8f6f4 with 16 / 8 queries per group (modes 3 / 4: e4m3 rows straight
    into the tensor cores, two-term e4m3 queries), legacy mma.sync (1), CUDA cores (0) — must hand the exact
    re-score the same candidates the per-query scan (K4b) does: same top-10 after re-scoring, exact scores."""
    x = clustered_rows(40000, d, 90, 0.9, seed=7 + d)
    index = built(ts, x, 128, "fp8")
    q = torch.from_numpy(oracle.normalize_f64(clustered_rows(300, d, 90, 0.9, seed=7 + d)))
    old_mode, old_nq = ts.get_tunable("ivf.group_mma"), ts.get_tunable("ivf.group_min_nq")
    try:
        ts.set_tunable("ivf.group_mma", mode)
        index._ws = {}
        s_g, i_g = index.ivf_search(q, 10, nprobe=16, rescore_k=100)      # list-major, this back end
        ts.set_tunable("ivf.group_min_nq", 0)
        index._ws = {}
        s_c, i_c = index.ivf_search(q, 10, nprobe=16, rescore_k=100)      # per-query scan
    finally:
        ts.set_tunable("ivf.group_mma", old_mode)
        ts.set_tunable("ivf.group_min_nq", old_nq)
        index._ws = {}
    assert oracle.recall_at_k(i_g.cpu().numpy(), i_c.cpu().numpy()) >= 0.995
    same = (i_g == i_c).all(dim=1)
    assert same.sum() >= 290 and torch.equal(s_g[same], s_c[same])
    # every list probed: equal to the exact search up to e4m3 candidate selection
    s_e, i_e = index.search(q[:64], 10)
    ts.set_tunable("ivf.group_mma", mode)
    try:
        s_a, i_a = index.ivf_search(q[:64], 10, nprobe=128, rescore_k=200)
    finally:
        ts.set_tunable("ivf.group_mma", old_mode)
    assert oracle.recall_at_k(i_a.cpu().numpy(), i_e.cpu().numpy()) >= 0.99
