"""Index mutation as the reference's writer does it (ec2/generate_embeddings/__main__.py:84-101: upsert keyed by
slogan_id into a table without a fixed size): an upsert stream must leave the index EQUAL to one rebuilt from
scratch from the final table — exact search bit for bit, IVF search through overflow lists and tombstones equal to
the IVF search over freshly packed lists with the same centroids — and the row store must grow on demand."""
import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ts():
    import theoremsearch_b200 as ts
    assert torch.cuda.is_available()
    return ts


def _stream(rng, table, n_batches, batch, dim, id_space):
    """Random upsert batches over `id_space` ids (so existing ids recur), some ids repeated inside a batch."""
    for _ in range(n_batches):
        ids = rng.integers(0, id_space, size=batch)
        ids[batch // 2] = ids[0]                       # an in-batch duplicate: the later row must win
        rows = rng.standard_normal((batch, dim)).astype(np.float32)
        yield ids.astype(np.int64), rows


def _rebuild(ts, table, dim, dtype="bf16"):
    ids = np.fromiter(table.keys(), dtype=np.int64, count=len(table))
    rows = np.stack(list(table.values())) if table else np.zeros((0, dim), np.float32)
    return ts.build_index(rows, ids=ids, dtype=dtype), ids, rows


@pytest.mark.parametrize("dim,host_rows", [(256, False), (1024, True), (100, False)])
def test_upsert_stream_equals_rebuild_from_scratch(ts, dim, host_rows):
    rng = np.random.default_rng(dim)
    table = {}
    index = ts.TheoremIndex(dim, 64, dtype="bf16")              # far too small on purpose: it must grow
    first_ids = np.arange(1000, 1300, dtype=np.int64)
    first = rng.standard_normal((300, dim)).astype(np.float32)
    assert oracle.upsert_rows(table, first_ids, first) == 0
    assert index.upsert(first if host_rows else torch.from_numpy(first).cuda(), first_ids) == 0
    for ids, rows in _stream(rng, table, 12, 97, dim, 1600):
        want = oracle.upsert_rows(table, ids + 1000, rows)
        got = index.upsert(rows if host_rows else torch.from_numpy(rows).cuda(), ids + 1000)
        assert got == want
    assert len(index) == len(table) and index.capacity >= len(index)
    fresh, ids_all, rows_all = _rebuild(ts, table, dim)
    assert torch.equal(index.get_rows(), fresh.get_rows())      # same rows, same places
    q = torch.from_numpy(oracle.synthetic_queries(9, dim))
    for qq in (q[:1], q):                                       # scan path and tensor-core path
        s0, i0 = fresh.search(qq, 10)
        s1, i1 = index.search(qq, 10)
        assert torch.equal(s0, s1) and torch.equal(i0, i1)
    # and against the oracle on the final table
    stored = oracle.bf16_round(oracle.normalize_f64(rows_all))
    o_s, o_i = oracle.exact_search(oracle.normalize_f64(q.numpy()), stored, 10, ids=ids_all)
    s1, i1 = index.search(q, 10)
    assert np.array_equal(i1.cpu().numpy(), o_i)
    assert np.max(np.abs(s1.cpu().numpy() - o_s)) < 1e-5


def test_rows_without_ids_use_their_position_as_id(ts):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((50, 64)).astype(np.float32)
    index = ts.build_index(x)                                    # no ids: id == row
    y = rng.standard_normal((3, 64)).astype(np.float32)
    assert index.upsert(y, [5, 50, 7]) == 2                      # rows 5 and 7 replaced, id 50 == next row: appended
    assert len(index) == 51
    x2 = np.concatenate([x, y[1:2]])
    x2[5], x2[7] = y[0], y[2]
    fresh = ts.build_index(x2)
    assert torch.equal(index.get_rows(), fresh.get_rows())
    s0, i0 = fresh.search(torch.from_numpy(y), 5)
    s1, i1 = index.search(torch.from_numpy(y), 5)
    assert torch.equal(s0, s1) and torch.equal(i0, i1)
    assert i1[:, 0].tolist() == [5, 50, 7]
    # an id that is not the next row position switches the index to an explicit id table
    assert index.upsert(y[:1], [9000]) == 0
    s2, i2 = index.search(torch.from_numpy(y[:1]), 2)
    assert sorted(i2[0].tolist()) == [5, 9000] and s2[0, 0] == s2[0, 1]


@pytest.mark.parametrize("no_vmm", [0, 1])
def test_add_grows_capacity_and_keeps_results(ts, no_vmm):
    """The row store grows on demand. Default: virtual-memory mapped, grows in place (the data pointer never moves, no
    row is copied); `store.no_vmm = 1`: plain allocation, copied on growth. Same results either way."""
    from theoremsearch_b200._lib import lib
    x = oracle.synthetic_rows(0, 5000, 128, seed=11)
    ts.set_tunable("store.no_vmm", no_vmm)
    try:
        index = ts.TheoremIndex(128, 10, dtype="bf16")
    finally:
        ts.set_tunable("store.no_vmm", 0)
    assert int(lib.ts_index_grows_in_place(index.handle)) == 1 - no_vmm
    ptr0 = lib.ts_index_data(index.handle)
    for lo in range(0, 5000, 700):
        index.add(torch.from_numpy(x[lo:lo + 700]).cuda())
    assert len(index) == 5000 and index.capacity >= 5000
    if not no_vmm:
        assert lib.ts_index_data(index.handle) == ptr0        # grown in place
        index.reserve(3_000_000)                              # 768 MB more address space backed on demand
        assert lib.ts_index_data(index.handle) == ptr0 and index.capacity == 3_000_000
    fresh = ts.build_index(x)
    q = torch.from_numpy(oracle.synthetic_queries(4, 128))
    s0, i0 = fresh.search(q, 10)
    s1, i1 = index.search(q, 10)
    assert torch.equal(s0, s1) and torch.equal(i0, i1)
    sh, ih = index.search_host(q.numpy(), 10)                    # the host ctx re-sizes its workspace too
    assert np.array_equal(ih, i0.cpu().numpy())


@pytest.mark.parametrize("list_dtype", ["bf16", "fp8"])
def test_ivf_lists_stay_valid_across_upserts(ts, list_dtype):
    """Overflow lists + tombstones: (a) probing every list equals the exact search on the mutated corpus (bf16
    lists), (b) a realistic probe budget equals the same search over freshly packed lists with the same
    centroids, (c) re-pack clears the pending state without changing a result."""
    dim, n0, nlist = 256, 40_000, 64
    rng = np.random.default_rng(5)
    table = {}
    ids0 = np.arange(n0, dtype=np.int64) * 3
    x0 = oracle.synthetic_rows(0, n0, dim, seed=21)
    oracle.upsert_rows(table, ids0, x0)
    index = ts.build_index(x0, ids=ids0)
    index.ivf_train(nlist, n_sample=20_000, iters=4, seed=0)
    index.ivf_build(list_dtype)
    assert index.ivf_pending() == (0, 0)
    for ids, rows in _stream(rng, table, 6, 301, dim, 2 * n0):   # ~half replace (ids divisible by 3), half append
        oracle.upsert_rows(table, ids, rows)
        index.upsert(torch.from_numpy(rows).cuda(), ids)
    ovf, dead = index.ivf_pending()
    assert ovf > 0 and dead > 0 and ovf + dead < n0 // 10        # incremental state, no automatic re-pack yet
    fresh, ids_all, rows_all = _rebuild(ts, table, dim)
    fresh.ivf_set_centroids(index.ivf_centroids())
    fresh.ivf_build(list_dtype)
    q = torch.from_numpy(oracle.synthetic_queries(40, dim))
    if list_dtype == "bf16":
        s_e, i_e = index.search(q, 10)
        for qq, se, ie in ((q, s_e, i_e), (q[:1], s_e[:1], i_e[:1])):
            s_a, i_a = index.ivf_search(qq, 10, nprobe=nlist, rescore_k=10)
            assert torch.equal(se, s_a) and torch.equal(ie, i_a)
    for qq in (q, q[:1]):                                        # batch (per-query scan while mutated) and latency mode
        s_m, i_m = index.ivf_search(qq, 10, nprobe=8, rescore_k=100)
        s_f, i_f = fresh.ivf_search(qq, 10, nprobe=8, rescore_k=100)
        assert torch.equal(s_m, s_f) and torch.equal(i_m, i_f)
    if list_dtype == "fp8":
        # the list-major batch scan (K4d, tcgen05) over main + overflow lists with tombstones skipped in its epilogue
        ts.set_tunable("ivf.group_min_lists", 1)
        try:
            index._ws, fresh._ws = {}, {}
            before = ts.kernel_launches()
            s_g, i_g = index.ivf_search(q, 10, nprobe=8, rescore_k=100)
            assert ts.kernel_launches() - before >= 12               # coarse K3 chain + invert kernels + scan + select
            s_h, i_h = fresh.ivf_search(q, 10, nprobe=8, rescore_k=100)
            assert oracle.recall_at_k(i_g.cpu().numpy(), i_h.cpu().numpy()) >= 0.995
            same = (i_g == i_h).all(dim=1)
            assert same.sum() >= 38 and torch.equal(s_g[same], s_h[same])
            s_m, i_m = index.ivf_search(q, 10, nprobe=8, rescore_k=100)
            assert torch.equal(s_g, s_m) and torch.equal(i_g, i_m)   # deterministic
        finally:
            ts.set_tunable("ivf.group_min_lists", 0)
            index._ws, fresh._ws = {}, {}
    sizes = index.ivf_list_sizes()
    assert int(sizes.sum()) == len(index) + dead                 # tombstones still occupy their slots
    index.ivf_repack()
    assert index.ivf_pending() == (0, 0)
    assert int(index.ivf_list_sizes().sum()) == len(index)
    off, rows = index.ivf_lists()
    assert sorted(rows.cpu().tolist()) == list(range(len(index)))
    s_r, i_r = index.ivf_search(q, 10, nprobe=8, rescore_k=100)
    s_f, i_f = fresh.ivf_search(q, 10, nprobe=8, rescore_k=100)
    assert torch.equal(s_r, s_f) and torch.equal(i_r, i_f)


def test_heavy_churn_triggers_the_automatic_repack(ts):
    dim, n0 = 128, 20_000
    x0 = oracle.synthetic_rows(0, n0, dim, seed=31)
    index = ts.build_index(x0)
    index.ivf_train(32, n_sample=n0, iters=3, seed=0)
    index.ivf_build("fp8")
    more = oracle.synthetic_rows(n0, 1500, dim, seed=31)
    index.add(torch.from_numpy(more).cuda())
    assert index.ivf_pending() == (1500, 0)
    index.add(torch.from_numpy(oracle.synthetic_rows(n0 + 1500, 3000, dim, seed=31)).cuda())   # churn past the re-pack threshold
    assert index.ivf_pending() == (0, 0) and len(index) == n0 + 4500
    q = torch.from_numpy(oracle.synthetic_queries(3, dim))
    s_e, i_e = index.search(q, 5)
    s_a, i_a = index.ivf_search(q, 5, nprobe=32, rescore_k=200)
    assert torch.equal(i_e, i_a) and torch.equal(s_e, s_a)


def test_fp8_scan_copy_follows_added_rows(ts):
    """build_index(dtype='fp8') then add(): cos_sim_topk must keep answering, with the new rows."""
    x = oracle.synthetic_rows(0, 30_000, 256, seed=41)
    index = ts.build_index(x, dtype="fp8")
    extra = oracle.synthetic_rows(30_000, 2_000, 256, seed=41)
    index.add(torch.from_numpy(extra).cuda())
    assert index.fp8_scan_ready
    q = torch.from_numpy(extra[1234])                            # a new row as the query: it must come back first
    s, i = ts.cos_sim_topk(q, index, 10)
    assert int(i[0]) == 30_000 + 1234
    s_e, i_e = index.search(q, 10)
    assert float(s[0]) == float(s_e[0, 0])                       # re-scored exactly
    assert len(set(i.tolist()) & set(i_e[0].tolist())) >= 9      # e4m3 candidate selection: recall, not identity


# ------------------------------------------------------------------------------------------------ delete by id
def _same_search(ts, index, table, dim, k=10, nq=7):
    """Exact search over `index` == the oracle over the table's rows (as stored: normalised, bf16), ids and scores."""
    ids_all = np.fromiter(table.keys(), dtype=np.int64, count=len(table))
    rows_all = np.stack(list(table.values()))
    stored = oracle.bf16_round(oracle.normalize_f64(rows_all))
    q = oracle.synthetic_queries(nq, dim)
    o_s, o_i = oracle.exact_search(oracle.normalize_f64(q), stored, k, ids=ids_all)
    for qq, lo in ((torch.from_numpy(q), 0), (torch.from_numpy(q[2:3]), 2)):     # K3 and K2
        s, i = index.search(qq, k)
        assert np.array_equal(i.cpu().numpy(), o_i[lo:lo + qq.shape[0]])
        assert np.max(np.abs(s.cpu().numpy() - o_s[lo:lo + qq.shape[0]])) < 1e-5


@pytest.mark.parametrize("with_ids", [True, False])
def test_delete_by_id_equals_the_oracle_table(ts, with_ids):
    """ec2/parse_arxiv_papers/__main__.py:271-274 + ON DELETE CASCADE (rds_schema.sql:35,46,51): rows vanish by id;
    unknown and repeated ids match nothing; the store stays dense; deletes, upserts and adds interleave."""
    dim, n0 = 192, 3000
    rng = np.random.default_rng(17)
    x0 = rng.standard_normal((n0, dim)).astype(np.float32)
    ids0 = (np.arange(n0, dtype=np.int64) * 7 + 11) if with_ids else np.arange(n0, dtype=np.int64)
    table = {}
    oracle.upsert_rows(table, ids0, x0)
    index = ts.build_index(x0, ids=ids0 if with_ids else None)
    # scattered rows, a run at the very end (no row has to move for those), an unknown id and a repeated one
    victims = np.concatenate([rng.choice(ids0[:-50], size=400, replace=False), ids0[-20:], [10**12], ids0[-1:]])
    want = oracle.delete_rows(table, victims)
    got, moves = index.delete(victims, return_moves=True)
    assert got == want == 420 and len(index) == n0 - 420 == len(table)
    assert moves.shape[1] == 2 and len(moves) <= 400
    assert (moves[:, 0] >= len(index)).all() and (moves[:, 1] < len(index)).all()      # tail rows filled the holes
    assert len(set(moves[:, 1].tolist())) == len(moves)
    _same_search(ts, index, table, dim)
    gone = set(int(v) for v in victims)
    _, ids_now = index.search(torch.from_numpy(oracle.synthetic_queries(1, dim)), 1000)
    assert not (set(ids_now[0].tolist()) & gone) and len(set(ids_now[0].tolist())) == 1000
    assert index.delete(victims) == 0                                                  # already gone: matches nothing
    # interleave: upsert (replace + append, one of them a deleted id coming back), delete again, add
    back = int(victims[0])
    up_ids = np.array([int(ids0[5]) if int(ids0[5]) in table else int(next(iter(table))), back, 10**9 + 1], dtype=np.int64)
    up_rows = rng.standard_normal((3, dim)).astype(np.float32)
    assert index.upsert(up_rows, up_ids) == oracle.upsert_rows(table, up_ids, up_rows)
    assert index.delete([10**9 + 1, int(next(iter(table)))]) == oracle.delete_rows(table, [10**9 + 1, int(next(iter(table)))])
    _same_search(ts, index, table, dim)
    # everything goes, then the index fills again
    assert index.delete(list(table.keys())) == len(table)
    table.clear()
    assert len(index) == 0
    new_ids = np.arange(5000, 5300, dtype=np.int64)
    new_rows = rng.standard_normal((300, dim)).astype(np.float32)
    index.upsert(new_rows, new_ids)
    oracle.upsert_rows(table, new_ids, new_rows)
    _same_search(ts, index, table, dim)


@pytest.mark.parametrize("list_dtype", ["bf16", "fp8"])
def test_ivf_lists_stay_valid_across_deletes(ts, list_dtype):
    """Deleted rows' list entries are tombstoned, moved rows' entries renamed, overflow rows filtered: the IVF search
    over the mutated lists equals the one over freshly packed lists of the surviving table (same centroids)."""
    dim, n0, nlist = 256, 30_000, 48
    rng = np.random.default_rng(23)
    x0 = oracle.synthetic_rows(0, n0, dim, seed=51)
    ids0 = np.arange(n0, dtype=np.int64) * 2 + 1
    table = {}
    oracle.upsert_rows(table, ids0, x0)
    index = ts.build_index(x0, ids=ids0)
    index.ivf_train(nlist, n_sample=15_000, iters=4, seed=0)
    index.ivf_build(list_dtype)
    # some rows into the overflow first (replacements and appends), then deletes that hit main lists AND overflow
    up_ids = np.concatenate([rng.choice(ids0, size=300, replace=False), np.arange(10**6, 10**6 + 200)]).astype(np.int64)
    up_rows = rng.standard_normal((500, dim)).astype(np.float32)
    oracle.upsert_rows(table, up_ids, up_rows)
    index.upsert(torch.from_numpy(up_rows).cuda(), up_ids)
    ovf0, dead0 = index.ivf_pending()
    assert ovf0 == 500 and dead0 == 300
    victims = np.concatenate([rng.choice(ids0, size=600, replace=False), up_ids[:50], up_ids[-60:]])
    assert index.delete(victims) == oracle.delete_rows(table, victims)
    ovf1, dead1 = index.ivf_pending()
    assert ovf1 < ovf0 and dead1 > dead0 and len(index) == len(table)
    with pytest.raises(ts.TheoremSearchError):
        index.ivf_lists()                                       # not exportable until the re-pack
    fresh, ids_all, rows_all = _rebuild(ts, table, dim)
    fresh.ivf_set_centroids(index.ivf_centroids())
    fresh.ivf_build(list_dtype)
    q = torch.from_numpy(oracle.synthetic_queries(33, dim))
    s_e, i_e = index.search(q, 10)
    s_x, i_x = fresh.search(q, 10)
    assert torch.equal(s_e, s_x) and torch.equal(i_e, i_x)      # same table, different row order: no exact ties here
    if list_dtype == "bf16":
        for qq, se, ie in ((q, s_e, i_e), (q[:1], s_e[:1], i_e[:1])):
            s_a, i_a = index.ivf_search(qq, 10, nprobe=nlist, rescore_k=10)
            assert torch.equal(se, s_a) and torch.equal(ie, i_a)
    for qq in (q, q[:1]):
        s_m, i_m = index.ivf_search(qq, 10, nprobe=8, rescore_k=100)
        s_f, i_f = fresh.ivf_search(qq, 10, nprobe=8, rescore_k=100)
        assert torch.equal(s_m, s_f) and torch.equal(i_m, i_f)
    # a row that moved is still found under its own content
    _, moves = index.delete([int(ids_all[3])], return_moves=True)
    oracle.delete_rows(table, [int(ids_all[3])])
    if len(moves):
        moved_row = index.get_rows(int(moves[0, 1]), 1)
        s1, i1 = index.ivf_search(moved_row, 1, nprobe=8, rescore_k=20, normalize=False)
        s2, i2 = index.search(moved_row, 1, normalize=False)
        assert torch.equal(i1, i2) and torch.equal(s1, s2)
    index.ivf_repack()
    assert index.ivf_pending() == (0, 0)
    off, rows = index.ivf_lists()
    assert sorted(rows.cpu().tolist()) == list(range(len(index)))
    # heavy deletion goes through the automatic re-pack
    many = list(table.keys())[: len(table) // 5]
    assert index.delete(many) == oracle.delete_rows(table, many)
    assert index.ivf_pending() == (0, 0) and len(index) == len(table)
    fresh2, _, _ = _rebuild(ts, table, dim)
    s_a, i_a = index.ivf_search(q, 10, nprobe=nlist, rescore_k=100)
    s_b, i_b = fresh2.search(q, 10)
    assert oracle.recall_at_k(i_a.cpu().numpy(), i_b.cpu().numpy()) >= (1.0 if list_dtype == "bf16" else 0.97)


def test_fp8_scan_copy_follows_deletes(ts):
    x = oracle.synthetic_rows(0, 20_000, 256, seed=61)
    index = ts.build_index(x, dtype="fp8")
    index.delete(np.arange(100, 600))
    assert index.fp8_scan_ready and len(index) == 19_500
    s, i = ts.cos_sim_topk(torch.from_numpy(x[19_999]), index, 5)       # the last row moved into a hole: still id 19999
    assert int(i[0]) == 19_999
    s2, i2 = ts.cos_sim_topk(torch.from_numpy(x[300]), index, 5)        # a deleted row is not found any more
    assert 300 not in i2.tolist()


def test_rows_added_without_ids_after_a_delete_get_fresh_ids(ts):
    """No ids = the row's position; a delete shrinks the positions but ids are not handed out twice (SERIAL-like)."""
    rng = np.random.default_rng(9)
    x = rng.standard_normal((50, 64)).astype(np.float32)
    index = ts.build_index(x)
    assert index.delete([3, 10]) == 2 and len(index) == 48
    y = rng.standard_normal((4, 64)).astype(np.float32)
    index.add(torch.from_numpy(y).cuda())                       # device path
    index.add(y[:1] * 2.0 + 1.0)                                # host path
    assert len(index) == 53
    _, ids = index.search(torch.from_numpy(x[:1]), 53)
    got = sorted(ids[0].tolist())
    assert got == sorted(set(range(50)) - {3, 10}) + [50, 51, 52, 53, 54]
    s, i = index.search(torch.from_numpy(y), 1)
    assert i[:, 0].tolist() == [50, 51, 52, 53]
    s, i = index.search(torch.from_numpy(x[49:50]), 1)          # row 49 moved into a freed slot, still id 49
    assert int(i[0, 0]) == 49


@pytest.mark.parametrize("seed,dtype", [(0, "bf16"), (1, "bf16"), (2, "f32")])
def test_random_interleaving_of_upserts_deletes_and_adds(ts, seed, dtype):
    """A writer's life: random upsert / delete / add batches; after every few operations the index must equal the
    oracle's table — exact search (K2 and K3) and, on bf16 rows, the IVF search with every list probed (overflow
    lists, tombstones, renamed entries and the automatic re-pack all occur along the way)."""
    dim, nlist = 64, 8
    rng = np.random.default_rng(100 + seed)
    table = {}
    ids0 = np.arange(0, 4000, 2, dtype=np.int64)
    x0 = rng.standard_normal((len(ids0), dim)).astype(np.float32)
    oracle.upsert_rows(table, ids0, x0)
    index = ts.build_index(x0, ids=ids0, dtype=dtype)
    ivf = dtype == "bf16"
    if ivf:
        index.ivf_train(nlist, iters=3, seed=0)
        index.ivf_build("bf16")
    fresh_id = 10**6
    states = set()
    for step in range(36):
        op = rng.integers(0, 3)
        if op == 0:                                             # upsert: existing and new ids mixed
            ids = rng.integers(0, 6000, size=int(rng.integers(1, 260))).astype(np.int64)
            rows = rng.standard_normal((len(ids), dim)).astype(np.float32)
            assert index.upsert(rows if step % 2 else torch.from_numpy(rows).cuda(), ids) == oracle.upsert_rows(table, ids, rows)
        elif op == 1 and len(table) > 600:                      # delete: stored ids, unknown ids, repeats
            stored = np.fromiter(table.keys(), dtype=np.int64, count=len(table))
            ids = np.concatenate([rng.choice(stored, size=int(rng.integers(1, 300)), replace=False),
                                  rng.integers(7000, 8000, size=5)])
            ids = np.concatenate([ids, ids[:3]])
            assert index.delete(ids) == oracle.delete_rows(table, ids)
        else:                                                   # add: fresh ids
            n = int(rng.integers(1, 200))
            ids = np.arange(fresh_id, fresh_id + n, dtype=np.int64)
            fresh_id += n
            rows = rng.standard_normal((n, dim)).astype(np.float32)
            index.add(torch.from_numpy(rows).cuda(), ids=ids)
            oracle.upsert_rows(table, ids, rows)
        assert len(index) == len(table)
        if ivf:
            states.add(tuple(v > 0 for v in index.ivf_pending()))
        if step % 4 == 3 or step == 35:
            ids_all = np.fromiter(table.keys(), dtype=np.int64, count=len(table))
            rows_all = oracle.normalize_f64(np.stack(list(table.values())))
            stored_rows = oracle.bf16_round(rows_all) if dtype == "bf16" else rows_all.astype(np.float32).astype(np.float64)
            q = oracle.synthetic_queries(5, dim, seed=step)
            o_s, o_i = oracle.exact_search(oracle.normalize_f64(q), stored_rows, 10, ids=ids_all)
            qt = torch.from_numpy(q)
            s3, i3 = index.search(qt, 10)
            s2, i2 = index.search(qt[:1], 10)
            assert np.array_equal(i3.cpu().numpy(), o_i) and np.array_equal(i2.cpu().numpy(), o_i[:1]), step
            assert np.max(np.abs(s3.cpu().numpy() - o_s)) < 1e-5
            if ivf:
                s_a, i_a = index.ivf_search(qt, 10, nprobe=nlist, rescore_k=10)
                assert torch.equal(i_a, i3) and torch.equal(s_a, s3), step
                s_b, i_b = index.ivf_search(qt[:1], 10, nprobe=nlist, rescore_k=10)
                assert torch.equal(i_b, i3[:1]) and torch.equal(s_b, s3[:1]), step
    if ivf:
        assert (True, True) in states                            # overflow rows and tombstones were live together
