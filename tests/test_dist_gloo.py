"""The N>1 path on CPU: world_size-2 (and 3) gloo process groups drive ShardedIndex's host logic
(shard bounds, all-gather layout, row rebasing).  The two CUDA stages (local scan -> packed keys,
K5 merge) are replaced by test-only oracle stand-ins; the real ones are covered by -m gpu tests."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle


def _keys_from(scores, rows):
    out = np.zeros(scores.shape, dtype=np.uint64)
    for a in range(scores.shape[0]):
        for b in range(scores.shape[1]):
            if rows[a, b] >= 0:
                out[a, b] = oracle.pack_key(float(scores[a, b]), int(rows[a, b]))
    return torch.from_numpy(out.view(np.int64))


def _oracle_merge(gathered, k, shard_base=None, id_map=None):
    g = gathered.numpy().view(np.uint64)
    nshards, nq, _ = g.shape
    ss, rr = [], []
    for sh in range(nshards):
        s = np.full((nq, k), -np.inf)
        r = np.full((nq, k), -1, dtype=np.int64)
        for a in range(nq):
            for b in range(k):
                if g[sh, a, b]:
                    s[a, b], r[a, b] = oracle.unpack_key(int(g[sh, a, b]))
        ss.append(s)
        rr.append(r)
    s, r = oracle.merge_shards(ss, rr, shard_base.tolist(), k)
    return torch.from_numpy(s), torch.from_numpy(r)


def _worker(rank, world, port, n, d, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from theoremsearch_b200.sharded import ShardedIndex, shard_bounds
    rows = oracle.bf16_round(oracle.normalize_f64(oracle.synthetic_rows(0, n, d, seed=0)))
    rows[n - 1] = rows[3]                                     # a cross-shard duplicate: tie -> lower GLOBAL row
    q = oracle.normalize_f64(oracle.synthetic_queries(3, d))
    lo, hi = shard_bounds(n, world)[rank]

    def local_search(queries, kk, normalize, allow_mask):
        s, i = oracle.exact_search(queries.numpy(), rows[lo:hi], kk)
        return _keys_from(s.astype(np.float32), i)

    sh = ShardedIndex(None, n, local_search=local_search, merge=_oracle_merge)
    assert (sh.lo, sh.hi) == (lo, hi) and sh.world == world
    s, i = sh.search(torch.from_numpy(q), k, normalize=False)
    np.save(os.path.join(out_dir, f"ids_{rank}.npy"), i.numpy())
    np.save(os.path.join(out_dir, f"scores_{rank}.npy"), s.numpy())
    # host-buffer entry point without a peer exchange: the gather + merge form, numpy in / numpy out
    s_h, i_h = sh.search_host(q.astype(np.float32), k, normalize=False)
    assert np.array_equal(i_h, i.numpy()) and np.array_equal(s_h, s.numpy())
    assert not sh.peer_exchange_error()
    sh.resync()                                               # a no-op without an exchange, must not dead-lock or raise
    sh.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_search_equals_unsharded(tmp_path, world):
    n, d, k = 1001, 64, 10
    port = 29600 + world + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, n, d, k, str(tmp_path)), nprocs=world, join=True)
    rows = oracle.bf16_round(oracle.normalize_f64(oracle.synthetic_rows(0, n, d, seed=0)))
    rows[n - 1] = rows[3]
    q = oracle.normalize_f64(oracle.synthetic_queries(3, d))
    want_s, want_i = oracle.exact_search(q, rows, k)
    for r in range(world):
        ids = np.load(tmp_path / f"ids_{r}.npy")
        sc = np.load(tmp_path / f"scores_{r}.npy")
        assert np.array_equal(ids, want_i), f"rank {r}"
        assert np.allclose(sc, want_s, atol=1e-6)


def test_shard_bounds_cover_and_balance():
    from theoremsearch_b200.sharded import shard_bounds
    for n in (0, 1, 7, 10_000_000, 100_000_001):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
            assert b == oracle.shard_bounds(n, w)


def _ivf_worker(rank, world, port, n, d, k, nlist, nprobe, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from theoremsearch_b200.sharded import ShardedIndex, shard_bounds
    rows = oracle.bf16_round(oracle.normalize_f64(oracle.synthetic_rows(0, n, d, seed=2)))
    cent = oracle.bf16_round(oracle.normalize_f64(oracle.synthetic_rows(0, nlist, d, seed=3)))   # ONE coarse quantiser
    q = oracle.normalize_f64(oracle.synthetic_queries(4, d))
    lo, hi = shard_bounds(n, world)[rank]
    local_rows = rows[lo:hi]
    local_assign = oracle.ivf_assign(local_rows, cent)          # every rank files its own slice of every list

    def local_ivf_search(queries, kk, npr, rescore_k, normalize, allow_mask):
        s = np.full((queries.shape[0], kk), -np.inf)
        i = np.full((queries.shape[0], kk), -1, dtype=np.int64)
        for a, qq in enumerate(queries.numpy()):
            ss, ii = oracle.ivf_search(qq, local_rows, cent, local_assign, kk, npr)
            s[a, :len(ss)], i[a, :len(ii)] = ss, ii
        return _keys_from(s.astype(np.float32), i)

    sh = ShardedIndex(None, n, local_search=None, merge=_oracle_merge, local_ivf_search=local_ivf_search)
    s, i = sh.ivf_search(torch.from_numpy(q), k, nprobe=nprobe, rescore_k=k, normalize=False)
    np.save(os.path.join(out_dir, f"ivf_ids_{rank}.npy"), i.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_ivf_equals_unsharded(tmp_path, world):
    """Row-sharded IVF with one shared coarse quantiser: every rank probes the same lists in its own
    slice; gather + merge must equal the unsharded IVF result."""
    n, d, k, nlist, nprobe = 1203, 32, 10, 12, 4
    port = 29700 + world + (os.getpid() % 200)
    mp.spawn(_ivf_worker, args=(world, port, n, d, k, nlist, nprobe, str(tmp_path)), nprocs=world, join=True)
    rows = oracle.bf16_round(oracle.normalize_f64(oracle.synthetic_rows(0, n, d, seed=2)))
    cent = oracle.bf16_round(oracle.normalize_f64(oracle.synthetic_rows(0, nlist, d, seed=3)))
    q = oracle.normalize_f64(oracle.synthetic_queries(4, d))
    assign = oracle.ivf_assign(rows, cent)
    for r in range(world):
        ids = np.load(tmp_path / f"ivf_ids_{r}.npy")
        for a in range(4):
            _, want = oracle.ivf_search(q[a], rows, cent, assign, k, nprobe)
            assert ids[a].tolist() == want.tolist(), (r, a)
