"""GPU tests of the sharded path on ONE GPU: K5 on gathered keys, and the device-initiated exchange
(`ts_search_sharded`) with world = 1 — the same kernels (slot store, flag, bounded wait, merge of the
world lists, rebasing to global rows) without a second GPU: the default two-kernel form (scan + exchange
kernel chained by programmatic dependent launch), its `independent` stream mode, the one-kernel form, the
host-buffer entry point and the time-out path.  The N-GPU run of the same checks is
`tests/run_sharded_multi_gpu.py` under torchrun; `bench.py --gpus N` runs a parity block as well."""
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


@pytest.mark.parametrize("k", [1, 10, 100])
def test_fused_exchange_world1_equals_plain_search(ts, k):
    from theoremsearch_b200.sharded import ShardedIndex
    x = oracle.synthetic_rows(0, 30000, 1024, seed=2)
    index = ts.build_index(x)
    sh = ShardedIndex(index, 30000).enable_peer_exchange(max_nq=3, max_k=128)
    q = torch.from_numpy(oracle.synthetic_queries(3, 1024))
    for nq in (1, 3):
        s0, i0 = index.search(q[:nq], k)
        for form in ({}, {"independent": True}, {"one_kernel": True}):
            for _ in range(3):      # both parities of the double-buffered slots, repeatedly
                s1, i1 = sh.search(q[:nq], k, **form)
                assert torch.equal(s0, s1) and torch.equal(i0, i1), form
    assert not sh.peer_exchange_error()
    # a batch falls back to the gather path (world 1: local keys -> K5) and still agrees
    qb = torch.from_numpy(oracle.synthetic_queries(16, 1024))
    s0, i0 = index.search(qb, min(k, 100))
    s1, i1 = sh.search(qb, min(k, 100))
    assert torch.equal(s0, s1) and torch.equal(i0, i1)
    sh.close()


def test_stream_of_independent_searches_overlaps_and_stays_exact(ts):
    """200 back-to-back searches with the `independent` promise (the scan of query n+1 overlaps the exchange
    kernel of query n, both reuse one workspace): every result equals the plain search of the same query."""
    from theoremsearch_b200.sharded import ShardedIndex
    x = oracle.synthetic_rows(0, 200_000, 256, seed=8)
    index = ts.build_index(x)
    sh = ShardedIndex(index, 200_000).enable_peer_exchange(max_nq=1, max_k=32)
    q = torch.from_numpy(oracle.synthetic_queries(200, 256)).cuda()
    torch.cuda.synchronize()
    outs = [sh.search(q[i:i + 1], 10, independent=True) for i in range(200)]
    torch.cuda.synchronize()
    ref_s, ref_i = index.search(q, 10)      # batched path, bitwise equal to the single-query path
    for i, (s, ids) in enumerate(outs):
        assert torch.equal(s[0], ref_s[i]) and torch.equal(ids[0], ref_i[i]), i
    # against the oracle as well
    stored = index.get_rows().cpu().numpy()
    o_s, o_i = oracle.exact_search(oracle.normalize_f64(q[:5].cpu().numpy()), stored, 10)
    for i in range(5):
        assert np.array_equal(outs[i][1][0].cpu().numpy(), o_i[i])
    assert not sh.peer_exchange_error()
    sh.close()


def test_single_gpu_stream_chain_equals_plain_search_with_caller_ids_and_mask(ts):
    """`TheoremIndex.search(..., independent=True)`: the sharded path's kernel chain on one GPU (a world of one)."""
    n, d = 60_000, 512
    x = oracle.synthetic_rows(0, n, d, seed=12)
    ids = np.arange(n, dtype=np.int64) * 5 + 7
    index = ts.build_index(x, ids=ids)
    q = torch.from_numpy(oracle.synthetic_queries(40, d)).cuda()
    allow = np.random.default_rng(0).random(n) < 0.4
    mask = ts.pack_allow_mask(allow, index.device)
    torch.cuda.synchronize()
    for k, m in ((10, None), (100, None), (10, mask)):
        outs = [index.search(q[i:i + 1], k, allow_mask=m, independent=True) for i in range(40)]
        torch.cuda.synchronize()
        for i, (s, idv) in enumerate(outs):
            s0, i0 = index.search(q[i:i + 1], k, allow_mask=m)
            assert torch.equal(s, s0) and torch.equal(idv, i0), (k, i)
    assert (outs[0][1] % 5 == 2).all()                       # caller ids, not row positions
    index.close()


def test_stream_chains_of_two_host_threads_do_not_interfere(ts):
    """Each (host thread, stream) chains its own searches through its own exchange handle."""
    import threading
    n, d = 40_000, 256
    index = ts.build_index(oracle.synthetic_rows(0, n, d, seed=14))
    q = torch.from_numpy(oracle.synthetic_queries(64, d)).cuda()
    want = [index.search(q[i:i + 1], 10) for i in range(64)]
    torch.cuda.synchronize()
    got, errors = {}, []

    def worker(t):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                outs = [index.search(q[i:i + 1], 10, independent=True) for i in range(t, 64, 2) for _ in range(3)]
                stream.synchronize()
            got[t] = outs
        except Exception as e:                                  # surfaced by the assert below
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    assert len(index._xchg1) == 2
    for t in range(2):
        for j, (s, i) in enumerate(got[t]):
            s0, i0 = want[t + 2 * (j // 3)]
            assert torch.equal(s, s0) and torch.equal(i, i0)
    index.close()


def test_fused_exchange_rebases_rows_and_maps_ids(ts):
    """A 'shard' that starts at global row 5000: fused results carry global rows / caller ids."""
    from theoremsearch_b200.sharded import ShardedIndex
    x = oracle.synthetic_rows(0, 2000, 256, seed=5)
    index = ts.build_index(x)
    sh = ShardedIndex(index, 2000)
    sh.lo, sh.hi = 5000, 7000                 # pretend this rank holds global rows [5000, 7000)
    sh.enable_peer_exchange(max_nq=2, max_k=32)
    qs = torch.from_numpy(oracle.synthetic_queries(2, 256))
    for j in range(2):                        # single queries: the device-initiated exchange path
        q = qs[j:j + 1]
        s0, i0 = index.search(q, 10)
        for form in ({}, {"one_kernel": True}):
            s1, i1 = sh.search(q, 10, **form)
            assert torch.equal(s0, s1) and torch.equal(i1, i0 + 5000)
            sh.id_map = (torch.arange(7000, dtype=torch.int64, device="cuda") * 2 + 1)
            s2, i2 = sh.search(q, 10, **form)
            assert torch.equal(i2, (i0 + 5000) * 2 + 1)
            sh.id_map = None
    sh.close()


def test_sharded_host_buffer_entry_point(ts):
    from theoremsearch_b200.sharded import ShardedIndex
    x = oracle.synthetic_rows(0, 50_000, 768, seed=6)
    index = ts.build_index(x)
    sh = ShardedIndex(index, 50_000).enable_peer_exchange(max_nq=2, max_k=32)
    q = oracle.synthetic_queries(6, 768)
    for i in range(6):
        s0, i0 = index.search_host(q[i], 10)
        s1, i1 = sh.search_host(q[i], 10)
        assert np.array_equal(s0, s1) and np.array_equal(i0, i1)
    s0, i0 = index.search_host(q[:2], 7)      # nq = 2 is below batch.min_nq only if that tunable was raised
    s1, i1 = sh.search_host(q[:2], 7)
    assert np.array_equal(s0, s1) and np.array_equal(i0, i1)
    sh.close()


@pytest.mark.parametrize("form", [{}, {"one_kernel": True}])
def test_peer_timeout_poisons_the_result_and_is_visible_without_polling(ts, form):
    """a rank that gives up waiting must not return a partial merge. The test hook makes the rank
    withhold its own flag, so the bounded wait expires: the result is (-inf, -1), the sticky flag is readable
    without a device synchronise, the next search raises PeerExchangeTimeout, resync() recovers."""
    from theoremsearch_b200 import _lib
    from theoremsearch_b200.sharded import PeerExchangeTimeout, ShardedIndex
    x = oracle.synthetic_rows(0, 5000, 256, seed=7)
    index = ts.build_index(x)
    sh = ShardedIndex(index, 5000).enable_peer_exchange(max_nq=1, max_k=32)
    q = torch.from_numpy(oracle.synthetic_queries(1, 256))
    s0, i0 = sh.search(q, 10, **form)
    sh.set_exchange_timeout(0.05)
    _lib.set_tunable("xchg.debug_no_flag", 1)
    try:
        s1, i1 = sh.search(q, 10, **form)
        torch.cuda.synchronize()
    finally:
        _lib.set_tunable("xchg.debug_no_flag", 0)
    assert torch.isinf(s1).all() and (s1 < 0).all() and (i1 == -1).all()
    assert sh.peer_exchange_error()
    with pytest.raises(PeerExchangeTimeout):
        sh.search(q, 10, **form)
    sh.resync()
    assert not sh.peer_exchange_error()
    s2, i2 = sh.search(q, 10, **form)
    assert torch.equal(s0, s2) and torch.equal(i0, i2)
    sh.close()
