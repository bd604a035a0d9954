"""GPU parity tests for the exact path (K1 normalise/cast, K2 scan + fused top-k, K5 merge),
all through the C ABI (`theoremsearch_b200` is a ctypes veneer).  Oracle: oracle/oracle.py.

Bar: ids bit-exact against the fp64 oracle on the SAME quantised inputs, except inside
eps-windows where fp32 re-association may legitimately reorder near-ties (eps written below);
scores within 1e-5 of the fp64 value (north_star allows 1e-3 for bf16)."""
import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-5   # |fp32-accumulated score - fp64 oracle score|
TIE_EPS = 2e-6     # oracle scores closer than this may swap ranks


@pytest.fixture(scope="module")
def ts():
    import theoremsearch_b200 as ts
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return ts


def unit_rows(n, d, seed):
    x = oracle.synthetic_rows(0, n, d, seed=seed)
    return oracle.normalize_f64(x)


def check_against_oracle(ts, index, queries_prepared, k, got_scores, got_ids, ids=None, allow=None):
    """queries_prepared: exactly what the kernel dots with (fp32, already normalised)."""
    rows = index.get_rows().cpu().numpy()
    ref_s, ref_i = oracle.exact_search(queries_prepared, rows, k, ids=ids, allow=allow)
    all_s = oracle.scores_f64(queries_prepared, rows)
    got_scores = got_scores.cpu().numpy()
    got_ids = got_ids.cpu().numpy()
    id_to_row = None if ids is None else {int(v): i for i, v in enumerate(ids)}
    for qi in range(ref_i.shape[0]):
        valid = ref_i[qi] >= 0
        assert np.array_equal(got_ids[qi] >= 0, valid), f"query {qi}: padding differs"
        assert np.all(np.isneginf(got_scores[qi][~valid]))
        g = got_ids[qi][valid]
        r = ref_i[qi][valid]
        if not np.array_equal(g, r):
            grow = g if id_to_row is None else np.array([id_to_row[int(v)] for v in g])
            rrow = r if id_to_row is None else np.array([id_to_row[int(v)] for v in r])
            assert len(set(g.tolist())) == g.size, f"query {qi}: duplicate ids"
            assert np.all(np.abs(all_s[qi][grow] - all_s[qi][rrow]) <= TIE_EPS), \
                f"query {qi}: ids differ outside the tie window\n got {g}\n ref {r}"
            grow_s = all_s[qi][grow]
        else:
            grow_s = ref_s[qi][valid]
        assert np.max(np.abs(got_scores[qi][valid] - grow_s), initial=0.0) <= SCORE_TOL


# ------------------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("d", [1024, 768, 100, 7])
def test_k1_normalize_cast_bit_exact(ts, d):
    x = oracle.synthetic_rows(0, 3000, d, seed=11) * 3.0
    x[5] = 0.0                       # zero row -> eps path -> zeros
    x[6] *= 1e-20                    # tiny norm, still above eps after sqrt? (1e-20*sqrt(d) > 1e-12 is false)
    index = ts.build_index(x, dtype="bf16", normalize=True)
    got = index.get_rows().cpu().numpy()
    want = oracle.bf16_round(oracle.normalize_f64(x))
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    index32 = ts.build_index(x, dtype="f32", normalize=True)
    assert np.array_equal(index32.get_rows().cpu().numpy(), oracle.normalize_f64(x))
    raw = ts.build_index(x, dtype="f32", normalize=False)
    assert np.array_equal(raw.get_rows().cpu().numpy(), x)


def test_k1_matches_torch_normalize_within_ulp(ts):
    """F.normalize (what util.cos_sim runs, test_app.py:76) sums in fp32; K1 in fp64: the
    normalised values agree to 1 fp32 ulp-ish."""
    x = oracle.synthetic_rows(0, 2000, 1024, seed=12)
    got = ts.build_index(x, dtype="f32").get_rows().cpu().numpy()
    want = oracle.normalize(x).numpy()
    assert np.max(np.abs(got - want)) <= 2e-8 * 4


def test_k1_sources_device_and_dtypes(ts):
    x = torch.from_numpy(oracle.synthetic_rows(0, 500, 1024, seed=13)).cuda()
    a = ts.build_index(x, dtype="bf16").get_rows()
    b = ts.build_index(x.cpu().numpy(), dtype="bf16").get_rows()
    assert torch.equal(a, b)
    xb = x.to(torch.bfloat16)
    c = ts.build_index(xb, dtype="bf16").get_rows().cpu().numpy()
    want = oracle.bf16_round(oracle.normalize_f64(xb.float().cpu().numpy()))
    assert np.array_equal(c, want)


# ------------------------------------------------------------------------------------------- K2
def test_tiny_hand_checkable_ties_lower_id_first(ts):
    # D=4, N=8, exactly representable; rows 1,4 identical and rows 2,6 identical -> tie rule.
    rows = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1],
                     [0, 1, 0, 0], [-1, 0, 0, 0], [0, 0, 1, 0], [0.5, 0.5, 0.5, 0.5]], dtype=np.float32)
    q = np.array([0, 1, 1, 0], dtype=np.float32)
    index = ts.build_index(rows, dtype="bf16", normalize=False)
    s, i = ts.cos_sim_topk(torch.from_numpy(q), index, 8, normalize_queries=False)
    assert i.tolist() == [1, 2, 4, 6, 7, 0, 3, 5]
    assert s.tolist() == [1, 1, 1, 1, 1, 0, 0, 0]
    s, i = ts.cos_sim_topk(torch.from_numpy(q), index, 3, normalize_queries=False)
    assert i.tolist() == [1, 2, 4]


@pytest.mark.parametrize("n,d,k", [
    (5000, 1024, 10), (5000, 1024, 1), (4097, 1024, 5), (3000, 768, 10), (3000, 768, 100),
    (2500, 100, 10), (2000, 8, 20), (2000, 7, 3), (3001, 384, 32), (3001, 512, 33),
    (1500, 2048, 10), (1500, 1536, 50), (20000, 1024, 200), (20000, 1024, 1024), (9000, 256, 128),
])
def test_scan_matches_oracle(ts, n, d, k):
    rows = unit_rows(n, d, seed=n + d)
    index = ts.build_index(rows, dtype="bf16", normalize=False)
    q = oracle.normalize_f64(oracle.synthetic_queries(3, d, seed=77 + k))
    s, i = ts.cos_sim_topk(torch.from_numpy(q), index, k, normalize_queries=False)
    check_against_oracle(ts, index, q, k, s, i)


@pytest.mark.parametrize("d,k", [(1024, 10), (768, 64), (100, 10)])
def test_scan_f32_storage(ts, d, k):
    rows = unit_rows(4000, d, seed=5)
    index = ts.build_index(rows, dtype="f32", normalize=False)
    q = oracle.normalize_f64(oracle.synthetic_queries(2, d))
    s, i = ts.cos_sim_topk(torch.from_numpy(q), index, k, normalize_queries=False)
    check_against_oracle(ts, index, q, k, s, i)


def test_query_normalisation_on_device(ts):
    rows = unit_rows(3000, 1024, seed=3)
    index = ts.build_index(rows, dtype="bf16", normalize=False)
    q_raw = oracle.synthetic_queries(4, 1024) * 7.5
    s, i = ts.cos_sim_topk(torch.from_numpy(q_raw), index, 10, normalize_queries=True)
    check_against_oracle(ts, index, oracle.normalize_f64(q_raw), 10, s, i)
    # bf16 queries are accepted too
    qb = torch.from_numpy(q_raw).to(torch.bfloat16)
    s, i = ts.cos_sim_topk(qb, index, 10, normalize_queries=True)
    check_against_oracle(ts, index, oracle.normalize_f64(qb.float().numpy()), 10, s, i)


def test_planted_neighbours_exact_order(ts):
    """Planted rows at cosines 0.90, 0.89, ... to the query: order is deterministic in bf16."""
    d, n, k = 1024, 30000, 10
    rows = unit_rows(n, d, seed=21)
    q = oracle.normalize_f64(oracle.synthetic_queries(1, d, seed=5))[0]
    rng = np.random.default_rng(0)
    planted = rng.choice(n, size=k, replace=False)
    for j, r in enumerate(planted):
        c = 0.90 - 0.01 * j
        noise = rows[r] - (rows[r] @ q) * q
        noise /= np.linalg.norm(noise)
        rows[r] = c * q + np.sqrt(1 - c * c) * noise
    index = ts.build_index(rows, dtype="bf16", normalize=False)
    s, i = ts.cos_sim_topk(torch.from_numpy(q), index, k, normalize_queries=False)
    assert i.tolist() == planted.tolist()
    assert np.allclose(s.cpu().numpy(), 0.90 - 0.01 * np.arange(k), atol=1e-3)


def test_duplicate_rows_tie_rule_at_scale(ts):
    d, n = 1024, 10000
    rows = unit_rows(n, d, seed=31)
    q = oracle.normalize_f64(oracle.synthetic_queries(1, d, seed=9))[0]
    best = rows[123].copy()
    c = 0.8
    noise = best - (best @ q) * q
    noise /= np.linalg.norm(noise)
    best = (c * q + np.sqrt(1 - c * c) * noise).astype(np.float32)
    dup = [9000, 17, 4242, 123, 8000]
    for r in dup:
        rows[r] = best
    index = ts.build_index(rows, dtype="bf16", normalize=False)
    s, i = ts.cos_sim_topk(torch.from_numpy(q), index, 5, normalize_queries=False)
    assert i.tolist() == sorted(dup)
    assert len(set(s.tolist())) == 1


def test_zero_rows_and_zero_query(ts):
    d = 1024
    x = oracle.synthetic_rows(0, 1000, d, seed=41)
    x[10] = 0
    x[20] = 0
    index = ts.build_index(x, dtype="bf16", normalize=True)
    q = oracle.synthetic_queries(1, d)[0]
    s, i = ts.cos_sim_topk(torch.from_numpy(q), index, 1000)
    s = s.cpu().numpy()
    i = i.cpu().numpy()
    assert s[np.where(i == 10)[0][0]] == 0.0 and s[np.where(i == 20)[0][0]] == 0.0
    # zero query: every score is 0 -> pure id order
    s, i = ts.cos_sim_topk(torch.zeros(d), index, 7)
    assert i.tolist() == list(range(7)) and s.tolist() == [0.0] * 7


def test_k_larger_than_n_and_empty_index(ts):
    rows = unit_rows(6, 1024, seed=1)
    index = ts.build_index(rows, dtype="bf16", normalize=False)
    q = rows[2]
    s, i = ts.cos_sim_topk(torch.from_numpy(q), index, 10, normalize_queries=False)
    assert i[0].item() == 2 and sorted(i[:6].tolist()) == list(range(6))
    assert i[6:].tolist() == [-1] * 4 and torch.isneginf(s[6:]).all()
    empty = ts.TheoremIndex(1024, 16)
    s, i = empty.search(torch.from_numpy(q), 3)
    assert i.tolist() == [[-1, -1, -1]]


def test_caller_ids_and_incremental_add(ts):
    d = 768
    rows = unit_rows(5000, d, seed=51)
    ids = (np.arange(5000, dtype=np.int64) * 7 + 1_000_000_007)
    index = ts.TheoremIndex(d, 5000)
    index.add(rows[:1234], ids=ids[:1234], normalize=False)
    index.add(torch.from_numpy(rows[1234:]).cuda(), ids=torch.from_numpy(ids[1234:]), normalize=False)
    assert len(index) == 5000
    q = oracle.normalize_f64(oracle.synthetic_queries(2, d))
    s, i = index.search(torch.from_numpy(q), 10, normalize=False)
    check_against_oracle(ts, index, q, 10, s, i, ids=ids)
    index.add(rows[:1], ids=np.array([5], dtype=np.int64), normalize=False)  # past the initial reservation: it grows
    assert len(index) == 5001 and index.capacity >= 5001
    s2, i2 = index.search(torch.from_numpy(rows[:1]), 2, normalize=False)
    assert i2[0].tolist() == [int(ids[0]), 5] and s2[0, 0] == s2[0, 1]     # the duplicate, tie -> lower row first


def test_allow_mask_is_applied_before_topk(ts):
    d, n, k = 1024, 8000, 10
    rows = unit_rows(n, d, seed=61)
    index = ts.build_index(rows, dtype="bf16", normalize=False)
    rng = np.random.default_rng(1)
    q = oracle.normalize_f64(oracle.synthetic_queries(2, d))
    for frac in (0.5, 0.01, 0.0005):
        allow = rng.random(n) < frac
        mask = ts.pack_allow_mask(allow, index.device)
        s, i = index.search(torch.from_numpy(q), k, normalize=False, allow_mask=mask)
        check_against_oracle(ts, index, q, k, s, i, allow=allow)


def test_search_host_path_equals_device_path(ts):
    rows = unit_rows(6000, 1024, seed=71)
    index = ts.build_index(rows, dtype="bf16", normalize=False)
    q = oracle.synthetic_queries(3, 1024)
    s_d, i_d = index.search(torch.from_numpy(q), 10)
    s_h, i_h = index.search_host(q, 10, timing=True)
    assert np.array_equal(i_d.cpu().numpy(), i_h) and np.array_equal(s_d.cpu().numpy(), s_h)
    assert index.last_kernel_ms > 0


def test_reference_call_shape_search_theorems(ts):
    class Model:
        def __init__(self, table):
            self.table = table

        def encode(self, query, convert_to_tensor=True):
            return torch.from_numpy(self.table[query])

    d = 1024
    corpus = oracle.synthetic_rows(0, 500, d, seed=81)  # NOT normalised, like test_app.py:130
    theorems = [{"type": "theorem", "paper_url": f"u{i}", "content": f"c{i}", "global_context": ""} for i in range(500)]
    qv = corpus[77] + 0.05 * oracle.synthetic_queries(1, d)[0]
    hits = ts.search_theorems("a query", Model({"a query": qv}), theorems, torch.from_numpy(corpus))
    want_idx, want_s = oracle.search_theorems_topk(qv, corpus, 5)
    assert [h["index"] for h in hits] == want_idx.tolist()
    assert np.allclose([h["similarity"] for h in hits], want_s, atol=1e-3)  # bf16 corpus vs fp32 reference
    assert hits[0]["theorem"] is theorems[77]


# ------------------------------------------------------------------------------------------- K5
@pytest.mark.parametrize("nshards,nq,k", [(2, 1, 10), (8, 3, 10), (8, 5, 100), (4, 2, 1000), (3, 1, 33)])
def test_merge_topk_matches_oracle(ts, nshards, nq, k):
    d = 256
    n = 4000
    rows = unit_rows(n, d, seed=91)
    q = oracle.normalize_f64(oracle.synthetic_queries(nq, d))
    bounds = oracle.shard_bounds(n, nshards)
    keys, sh_s, sh_r = [], [], []
    for lo, hi in bounds:
        ix = ts.build_index(rows[lo:hi], dtype="bf16", normalize=False)
        keys.append(ix.search_keys(torch.from_numpy(q), k, normalize=False))
        s, i = ix.search(torch.from_numpy(q), k, normalize=False)
        sh_s.append(s.cpu().numpy())
        sh_r.append(i.cpu().numpy())
    gathered = torch.stack(keys)  # [nshards, nq, k]
    s, i = ts.merge_topk(gathered, k, shard_base=[lo for lo, _ in bounds])
    want_s, want_r = oracle.merge_shards(sh_s, sh_r, [lo for lo, _ in bounds], k)
    assert np.array_equal(i.cpu().numpy(), want_r)
    assert np.array_equal(s.cpu().numpy().astype(np.float64), want_s)
    # and the sharded result equals the unsharded one
    whole = ts.build_index(rows, dtype="bf16", normalize=False)
    s1, i1 = whole.search(torch.from_numpy(q), k, normalize=False)
    assert torch.equal(i1, i) and torch.equal(s1, s)


def _keys_tensor(scores, rows):
    """[L, Q, k] (score, local row) -> packed keys as K2 / K5 exchange them (row < 0 = empty slot)."""
    out = np.zeros(scores.shape, dtype=np.uint64)
    for idx in np.ndindex(*scores.shape):
        if rows[idx] >= 0:
            out[idx] = oracle.pack_key(float(scores[idx]), int(rows[idx]))
    return torch.from_numpy(out.view(np.int64)).cuda()


@pytest.mark.parametrize("case", ["random", "one_list_holds_everything", "staircase_overflows_the_prune_buffer",
                                   "few_non_empty_lists", "ties_across_lists"])
@pytest.mark.parametrize("nlists,k", [(148, 10), (64, 32), (17, 1), (256, 5)])
def test_pruned_merge_of_many_lists_equals_the_oracle(ts, case, nlists, k):
    """K5 with 16..256 lists and k <= 32 takes the pruned merge (only keys >= the k-th largest list head survive) —
    the same code K2's last CTA runs over its 148 per-CTA lists on every search. Adversarial layouts: every winner in
    one list, a staircase that leaves more than 256 survivors (the general merge must take over), almost-empty
    inputs (no k-th head exists), equal scores in different lists (lower global row first)."""
    rng = np.random.default_rng(nlists * 100 + k)
    nq = 3
    scores = np.full((nlists, nq, k), -np.inf, dtype=np.float32)
    rows = np.full((nlists, nq, k), -1, dtype=np.int64)
    base = np.arange(nlists, dtype=np.int64) * 1000
    for q in range(nq):
        for l in range(nlists):
            if case == "random":
                n = int(rng.integers(0, k + 1))
                sc = rng.standard_normal(n)
            elif case == "one_list_holds_everything":
                n = k
                sc = rng.standard_normal(n) + (100.0 if l == (7 * q + 3) % nlists else 0.0)
            elif case == "staircase_overflows_the_prune_buffer":
                n = k
                sc = l + rng.random(n) * 0.9                     # list l beats list l - 1 entirely
            elif case == "few_non_empty_lists":
                n = k if l in (1, nlists - 2) else 0
                sc = rng.standard_normal(n)
            else:                                                # ties: the same few score values everywhere
                n = k
                sc = rng.integers(0, 4, size=n).astype(np.float64)
            sc = np.sort(np.asarray(sc, dtype=np.float32))[::-1]
            rr = rng.permutation(900)[:n]
            if case == "ties_across_lists":                      # within a list: equal scores sorted by row ascending
                order = np.lexsort((rr, -sc))
                sc, rr = sc[order], rr[order]
            scores[l, q, :n], rows[l, q, :n] = sc, rr
    keys = _keys_tensor(scores, rows)
    s, i = ts.merge_topk(keys, k, shard_base=base.tolist())
    want_s, want_r = oracle.merge_shards(list(scores.astype(np.float64)), list(rows), base.tolist(), k)
    assert np.array_equal(i.cpu().numpy(), want_r)
    assert np.array_equal(s.cpu().numpy().astype(np.float64), want_s)


def test_errors_are_loud(ts):
    with pytest.raises(ts.TheoremSearchError):
        ts.TheoremIndex(4096, 10)          # dim too large
    index = ts.build_index(unit_rows(100, 64, seed=1), normalize=False)
    with pytest.raises(ts.TheoremSearchError):
        index.search(torch.zeros(64), 5000)  # k too large
    with pytest.raises(ts.TheoremSearchError):
        index.search(torch.zeros(65), 5)     # wrong dim


# ------------------------------------------------------------------------------------------- property tests
def test_random_shapes_equal_oracle_hypothesis(ts):
    """on random small shapes the CUDA path equals the oracle; the merge of shard results
    equals the unsharded result."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as hst

    @settings(max_examples=30, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(n=hst.integers(1, 2500), d=hst.integers(1, 260), k=hst.integers(1, 40), nq=hst.integers(1, 6),
           dtype=hst.sampled_from(["bf16", "f32"]), seed=hst.integers(0, 10_000), dup=hst.booleans())
    def run(n, d, k, nq, dtype, seed, dup):
        x = oracle.synthetic_rows(0, n, d, seed=seed)
        if dup and n > 3:
            x[n - 1] = x[1]                      # tie -> lower row
        index = ts.build_index(x, dtype=dtype)
        q = oracle.synthetic_queries(nq, d, seed=seed + 1)
        s, i = index.search(torch.from_numpy(q), k)
        check_against_oracle(ts, index, oracle.normalize_f64(q), k, s, i)
        # shard-merge: two masks that partition the rows
        cut = (n // 2 // 32) * 32
        if cut > 0:
            allow = np.zeros(n, dtype=bool)
            allow[:cut] = True
            m_lo = ts.pack_allow_mask(allow, index.device)
            m_hi = ts.pack_allow_mask(~allow, index.device)
            keys = torch.stack([index.search_keys(torch.from_numpy(q), k, allow_mask=m_lo),
                                index.search_keys(torch.from_numpy(q), k, allow_mask=m_hi)])
            ms, mi = ts.merge_topk(keys, k)
            assert torch.equal(ms, s) and torch.equal(mi, i)
        index.close()

    run()


def test_concurrent_host_threads_share_one_index(ts):
    """SURVEY §8b threading contract: searches on one index from several host threads (Streamlit: one script
    thread per session over a cached resource, streamlit_app.py:52-59) are safe — each thread gets its own
    ctx (stream, pinned staging, workspace) and its own device workspace."""
    import threading
    x = oracle.synthetic_rows(0, 60000, 512, seed=31)
    index = ts.build_index(x)
    qs = oracle.synthetic_queries(64, 512, seed=32)
    want_s, want_i = index.search_host(qs, 10)
    errors = []

    def worker(t):
        try:
            for rep in range(30):
                j = (t * 7 + rep) % 64
                s, i = index.search_host(qs[j], 10)                  # host path: per-thread ctx
                assert np.array_equal(i[0], want_i[j]) and np.array_equal(s[0], want_s[j])
                with torch.cuda.stream(streams[t]):                   # device path: per-thread, per-stream workspace
                    sd, idd = index.search(torch.from_numpy(qs[j]), 10)
                    streams[t].synchronize()
                assert np.array_equal(idd.cpu().numpy()[0], want_i[j])
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    streams = [torch.cuda.Stream() for _ in range(6)]
    threads = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors[:3]
    # contexts of exited threads are reclaimed (Streamlit: a fresh script thread per rerun) and a larger
    # context replaces this thread's smaller one: the table is bounded by live threads
    assert len(index._ctx) == 7
    s20, i20 = index.search_host(qs, 20)
    assert len(index._ctx) == 1
    assert np.array_equal(i20[:, :10], want_i)
