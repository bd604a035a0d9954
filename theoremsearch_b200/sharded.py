"""Row-sharded corpus across the GPUs of one box (one process per GPU, ``torch.distributed``).

New surface with no reference counterpart (the reference has no distributed code at all,
SURVEY §2); specified by BASELINE.json: GPU g holds rows [g*N/G, (g+1)*N/G), every GPU scans its
shard for the (replicated) queries, the k x 8-byte packed (score,row) keys of every GPU are exchanged
over NVLink — by the GPUs themselves for small batches (stores into CUDA-IPC mapped peer memory), by
ONE NCCL all-gather for large ones — and merged on every rank.  No corpus bytes cross NVLink.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib
from .index import _TORCH_TO_TS, TheoremIndex, _stream_ptr, merge_topk


class PeerExchangeTimeout(_lib.TheoremSearchError):
    """A rank waited longer than the exchange time-out for a peer's keys. The affected result holds
    score = -inf / id = -1 in every position; call ``ShardedIndex.resync()`` on every rank to recover."""


def shard_bounds(n_rows: int, world_size: int) -> list[tuple[int, int]]:
    """Contiguous shards [g*N/G, (g+1)*N/G)."""
    return [((g * n_rows) // world_size, ((g + 1) * n_rows) // world_size) for g in range(world_size)]


class ShardedIndex:
    """One rank's view of a row-sharded corpus.

    ``local`` holds this rank's rows; ``n_total`` is the global row count.  ``local_search`` and
    ``merge`` default to the CUDA path (``TheoremIndex.search_keys`` / K5); tests on the gloo
    backend inject CPU stand-ins for those two to exercise the host logic (bounds, gather
    layout, rebasing) without a GPU.
    """

    def __init__(self, local: Optional[TheoremIndex], n_total: int, group=None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None,
                 id_map: Optional[torch.Tensor] = None, local_ivf_search: Optional[Callable] = None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_total = int(n_total)
        self.bounds = shard_bounds(self.n_total, self.world)
        self.lo, self.hi = self.bounds[self.rank]
        self.local = local
        self._local_search = local_search or (lambda q, k, normalize, allow_mask:
                                              local.search_keys(q, k, normalize=normalize, allow_mask=allow_mask))
        self._merge = merge or merge_topk
        self._local_ivf_search = local_ivf_search or (
            lambda q, k, nprobe, rescore_k, normalize, allow_mask:
            local.ivf_search_keys(q, k, nprobe=nprobe, rescore_k=rescore_k, normalize=normalize, allow_mask=allow_mask))
        self.id_map = id_map
        self._base = None
        self._xchg = None
        self._xchg_limits = (0, 0)

    def shard_base(self, device) -> torch.Tensor:
        if self._base is None or self._base.device != torch.device(device):
            self._base = torch.tensor([lo for lo, _ in self.bounds], dtype=torch.int64, device=device)
        return self._base

    # -------------------------------------------------------------------------------- fused peer exchange
    def enable_peer_exchange(self, max_nq: int = 3, max_k: int = 256) -> "ShardedIndex":
        """Set up the device-initiated exchange (``ts_search_sharded``): every rank allocates a receive area,
        the CUDA-IPC handles are all-gathered once (host-side plumbing), peers are mapped.  Afterwards
        small-batch searches need no collective call at all: an exchange kernel chained to the scan by
        programmatic dependent launch stores the shard's k keys into the peers' memory over NVLink, waits
        for theirs and merges, while the next search's scan already streams the corpus."""
        if self._xchg is not None:
            return self
        dev = self.local.device
        # Every rank runs the SAME sequence of collectives whatever fails locally (a rank that raised early
        # while its peers wait in a collective would deadlock the job); the outcome is agreed on at the end.
        hb = int(lib.ts_xchg_handle_bytes())
        mine = np.zeros(hb, dtype=np.uint8)
        h = C.c_void_p()
        err = None
        try:
            check(lib.ts_xchg_create(C.byref(h), dev.index, self.world, self.rank, int(max_nq), int(max_k)))
            check(lib.ts_xchg_handle(h, mine.ctypes.data))
        except _lib.TheoremSearchError as e:
            err = e
        if self.world > 1:
            allh = torch.empty(self.world * hb, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, torch.from_numpy(mine).to(dev), group=self.group)
            host = np.ascontiguousarray(allh.cpu().numpy())
            if err is None:
                try:
                    check(lib.ts_xchg_connect(h, host.ctypes.data))
                except _lib.TheoremSearchError as e:
                    err = e
            ok = torch.tensor([0 if err is not None else 1], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)   # also: every area is mapped before any store
            if not bool(ok.item()):
                if h:
                    lib.ts_xchg_destroy(h)
                raise err if err is not None else _lib.TheoremSearchError(
                    -2, "peer exchange unavailable: another rank could not map the peer memory")
        else:
            if err is None:
                try:
                    check(lib.ts_xchg_connect(h, None))
                except _lib.TheoremSearchError as e:
                    err = e
            if err is not None:
                if h:
                    lib.ts_xchg_destroy(h)
                raise err
        self._xchg = h
        self._xchg_limits = (int(max_nq), int(max_k))
        return self

    def close(self) -> None:
        if self._xchg is not None:
            if self.world > 1 and dist.is_initialized():
                torch.cuda.synchronize()
                dist.barrier(group=self.group)      # nobody unmaps while a peer may still store
            lib.ts_xchg_destroy(self._xchg)
            self._xchg = None

    def _fused_ok(self, nq: int, k: int) -> bool:
        return (self._xchg is not None and nq <= self._xchg_limits[0] and k <= self._xchg_limits[1]
                and nq < _lib.get_tunable("batch.min_nq"))

    def _search_fused(self, queries, k: int, normalize: bool, allow_mask, independent: bool = False,
                      one_kernel: bool = False):
        ix = self.local
        q = ix._prep_queries(queries)
        nq = q.shape[0]
        scores = torch.empty((nq, k), dtype=torch.float32, device=ix.device)
        ids = torch.empty((nq, k), dtype=torch.int64, device=ix.device)
        ws = ix._workspace(nq, k)
        flags = (_lib.TS_SHARDED_INDEPENDENT if independent else 0) | (_lib.TS_SHARDED_ONE_KERNEL if one_kernel else 0)
        try:
            check(lib.ts_search_sharded(ix.handle, self._xchg, q.data_ptr(), _TORCH_TO_TS[q.dtype], nq, int(k),
                                        int(normalize), ix._mask_ptr(allow_mask), int(self.lo),
                                        self.id_map.data_ptr() if self.id_map is not None else None,
                                        scores.data_ptr(), ids.data_ptr(), ws.data_ptr(), ws.numel(), flags,
                                        _stream_ptr(ix.device)))
        except _lib.TheoremSearchError as e:
            if e.code == -6 and self.peer_exchange_error():
                raise PeerExchangeTimeout(e.code, _lib.last_error()) from None
            raise
        q.record_stream(torch.cuda.current_stream(ix.device))
        return scores, ids

    def search_host(self, queries: np.ndarray, k: int, normalize: bool = True,
                    allow_mask: Optional[torch.Tensor] = None):
        """End-to-end call with HOST buffers, the sharded counterpart of ``TheoremIndex.search_host``: numpy
        fp32 queries in, numpy (scores [nq, k], ids [nq, k]) out on every rank; staging, H2D, scan + exchange,
        D2H and the synchronise happen inside ``ts_search_sharded_host``."""
        q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32))
        if q.ndim == 1:
            q = q[None, :]
        nq = q.shape[0]
        if self.local is not None and q.shape[1] != self.local.dim:
            raise _lib.TheoremSearchError(-1, f"queries must be [nq, {self.local.dim}], got {q.shape}")
        if not self._fused_ok(nq, k):     # batches, or no peer exchange: the gather + merge form, copied back
            qt = torch.from_numpy(q)
            s, i = self.search(qt.to(self.local.device) if self.local is not None else qt, k, normalize, allow_mask)
            return s.cpu().numpy(), i.cpu().numpy()
        scores = np.empty((nq, k), dtype=np.float32)
        ids = np.empty((nq, k), dtype=np.int64)
        try:
            check(lib.ts_search_sharded_host(self.local.handle, self._xchg, q.ctypes.data, nq, int(k), int(normalize),
                                             self.local._mask_ptr(allow_mask), int(self.lo),
                                             self.id_map.data_ptr() if self.id_map is not None else None,
                                             scores.ctypes.data, ids.ctypes.data))
        except _lib.TheoremSearchError as e:
            if e.code == -6 and self.peer_exchange_error():
                raise PeerExchangeTimeout(e.code, _lib.last_error()) from None
            raise
        return scores, ids

    def peer_exchange_error(self) -> bool:
        """True once a search on this rank gave up waiting for a peer (its result was poisoned with -inf / -1).
        Reads a host-mapped flag: no device synchronisation, cheap enough to poll after every result."""
        return self._xchg is not None and int(lib.ts_xchg_error(self._xchg)) != 0

    def set_exchange_timeout(self, seconds: float) -> None:
        check(lib.ts_xchg_set_timeout_ms(self._xchg, int(seconds * 1000)))

    def resync(self) -> None:
        """Collective recovery after ``PeerExchangeTimeout`` or after the ranks' call sequences diverged (one rank
        raised between two searches): barrier, clear every rank's receive area / sequence counter / error flag,
        barrier. A rank that timed out stops pushing keys, so its peers time out on their next search too and
        every rank ends up here."""
        if self._xchg is None:
            return
        torch.cuda.synchronize(self.local.device)
        if self.world > 1:
            dist.barrier(group=self.group)
        check(lib.ts_xchg_reset(self._xchg))
        if self.world > 1:
            dist.barrier(group=self.group)

    def search(self, queries: torch.Tensor, k: int, normalize: bool = True,
               allow_mask: Optional[torch.Tensor] = None, independent: bool = False, one_kernel: bool = False):
        """Replicated queries [nq, D] -> global (scores [nq, k], ids [nq, k]) on every rank.

        With the peer exchange enabled a small batch is a scan kernel plus an exchange kernel chained by
        programmatic dependent launch (``ts_search_sharded``). ``independent=True`` promises that the queries /
        mask are not produced by the kernel enqueued just before this call on the current stream; the scan then
        starts streaming the corpus while the previous search's exchange kernel is still merging, so a stream
        of searches runs at the local scan rate. ``one_kernel=True`` selects the form where the scan kernel's
        last CTA does the exchange itself. Raises ``PeerExchangeTimeout`` if an EARLIER search timed out waiting
        for a peer (the flag is sticky; see ``resync``)."""
        nq = 1 if queries.dim() == 1 else queries.shape[0]
        if self._fused_ok(nq, k):
            return self._search_fused(queries, k, normalize, allow_mask, independent, one_kernel)
        keys = self._local_search(queries, k, normalize, allow_mask)          # [nq, k] packed keys
        if self.world == 1:
            gathered = keys.unsqueeze(0)
        else:
            flat = torch.empty(self.world * keys.numel(), dtype=keys.dtype, device=keys.device)
            dist.all_gather_into_tensor(flat, keys.contiguous().view(-1), group=self.group)
            gathered = flat.view((self.world,) + tuple(keys.shape))               # [G, nq, k], shard-major
        return self._merge(gathered, k, shard_base=self.shard_base(keys.device), id_map=self.id_map)

    # -------------------------------------------------------------------------------- IVF-Flat
    def ivf_train_build(self, nlist: int, n_sample: int = 0, iters: int = 10, seed: int = 0,
                        list_dtype: str = "fp8", src: int = 0) -> None:
        """One coarse quantiser for all shards: rank ``src`` trains on (a strided sample of) its own
        rows, the centroids are broadcast, every rank files its rows under them (every GPU
        holds all centroids and its slice of each list)."""
        dev = self.local.device
        if self.rank == src:
            self.local.ivf_train(nlist, n_sample=n_sample, iters=iters, seed=seed)
            cent = self.local.ivf_centroids()
        else:
            cent = torch.empty((nlist, self.local.dim), dtype=torch.float32, device=dev)
        if self.world > 1:
            dist.broadcast(cent, src=src, group=self.group)
        if self.rank != src:
            self.local.ivf_set_centroids(cent)
        self.local.ivf_build(list_dtype)

    def ivf_search(self, queries: torch.Tensor, k: int, nprobe: int = 32, rescore_k: int = 100,
                   normalize: bool = True, allow_mask: Optional[torch.Tensor] = None):
        """ANN over the sharded corpus: every rank probes the same ``nprobe`` lists (same centroids) in
        its own slice, re-scores its candidates exactly, and the packed keys are gathered and merged
        exactly like the exact path."""
        keys = self._local_ivf_search(queries, k, nprobe, rescore_k, normalize, allow_mask)
        if self.world == 1:
            gathered = keys.unsqueeze(0)
        else:
            flat = torch.empty(self.world * keys.numel(), dtype=keys.dtype, device=keys.device)
            dist.all_gather_into_tensor(flat, keys.contiguous().view(-1), group=self.group)
            gathered = flat.view((self.world,) + tuple(keys.shape))
        return self._merge(gathered, k, shard_base=self.shard_base(keys.device), id_map=self.id_map)
