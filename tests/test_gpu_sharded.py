"""GPU tests of the sharded path on ONE GPU: K5 on gathered keys, and the fused in-kernel exchange
(`ts_search_sharded`) with world = 1 — the same kernel path (slot store, flag, bounded wait, merge of
the world lists, rebasing to global rows) without a second GPU.  The two-GPU run of the same checks is
`tests/run_sharded_multi_gpu.py` under torchrun (needs N GPUs; the driver's scaling bench exercises it)."""
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
        for _ in range(3):      # both parities of the double-buffered slots, repeatedly
            s1, i1 = sh.search(q[:nq], k)
            assert torch.equal(s0, s1) and torch.equal(i0, i1)
    assert not sh.peer_exchange_error()
    # a batch falls back to the gather path (world 1: local keys -> K5) and still agrees
    qb = torch.from_numpy(oracle.synthetic_queries(16, 1024))
    s0, i0 = index.search(qb, min(k, 100))
    s1, i1 = sh.search(qb, min(k, 100))
    assert torch.equal(s0, s1) and torch.equal(i0, i1)
    sh.close()


def test_fused_exchange_rebases_rows_and_maps_ids(ts):
    """A 'shard' that starts at global row 5000: fused results carry global rows / caller ids."""
    from theoremsearch_b200.sharded import ShardedIndex
    x = oracle.synthetic_rows(0, 2000, 256, seed=5)
    index = ts.build_index(x)
    sh = ShardedIndex(index, 2000)
    sh.lo, sh.hi = 5000, 7000                 # pretend this rank holds global rows [5000, 7000)
    sh.enable_peer_exchange(max_nq=2, max_k=32)
    qs = torch.from_numpy(oracle.synthetic_queries(2, 256))
    for j in range(2):                        # single queries: the fused in-kernel exchange path
        q = qs[j:j + 1]
        s0, i0 = index.search(q, 10)
        s1, i1 = sh.search(q, 10)
        assert torch.equal(s0, s1) and torch.equal(i1, i0 + 5000)
        sh.id_map = (torch.arange(7000, dtype=torch.int64, device="cuda") * 2 + 1)
        s2, i2 = sh.search(q, 10)
        assert torch.equal(i2, (i0 + 5000) * 2 + 1)
        sh.id_map = None
    sh.close()
