"""``TheoremIndex`` — the corpus of theorem/slogan embeddings resident in one B200's HBM.

Host-side mirror of the reference's corpus containers: the ``embeddings_db`` tensor of
``test_app.py:129-130`` / ``app_showcase_model.py:52`` and the pgvector table
``theorem_embedding_qwen`` (``rds_schema.sql:50-53``).  All arithmetic happens in
``libtheoremsearch.so``; torch is used only for device buffers and streams.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import TS_BF16, TS_F16, TS_F32, check, lib

_TORCH_TO_TS = {torch.float32: TS_F32, torch.bfloat16: TS_BF16, torch.float16: TS_F16}
_NAME_TO_TS = {"bf16": TS_BF16, "bfloat16": TS_BF16, "f32": TS_F32, "fp32": TS_F32, "float32": TS_F32}


def _stream_ptr(device: torch.device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise _lib.TheoremSearchError(
            -2, "no CUDA device visible; theoremsearch_b200 has no CPU fallback (B200 / sm_100a only)")


def pack_allow_mask(allow: np.ndarray | torch.Tensor, device: torch.device | str | None = None) -> torch.Tensor:
    """Boolean row mask [N] -> the uint32 bitmask K2 consumes (bit r of word r//32), as an
    int32 torch tensor (same bits) on ``device``.  Mirrors the SQL WHERE of
    ``streamlit_app.py:175-243`` being applied before ``ORDER BY ... LIMIT``."""
    a = np.asarray(allow.cpu() if isinstance(allow, torch.Tensor) else allow).astype(bool)
    n = a.shape[0]
    words = (n + 31) // 32
    padded = np.zeros(words * 32, dtype=bool)
    padded[:n] = a
    bits = np.packbits(padded.reshape(words, 32), axis=1, bitorder="little").view(np.uint32).reshape(words)
    t = torch.from_numpy(bits.view(np.int32).copy())
    return t.to(device) if device is not None else t


class TheoremIndex:
    """Exact-search index over L2-normalised embeddings stored as bf16 (or fp32) rows."""

    def __init__(self, dim: int, capacity: int, dtype: str = "bf16", device: int | str | torch.device = 0):
        _require_cuda()
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise _lib.TheoremSearchError(-1, f"TheoremIndex lives on a CUDA device, not {dev}")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.dim = int(dim)
        self.dtype = dtype
        self.scan_dtype = dtype      # "fp8" once build_index(dtype="fp8") has attached the e4m3 scan copy
        h = C.c_void_p()
        check(lib.ts_index_create(C.byref(h), self.device.index, self.dim, _NAME_TO_TS[dtype], int(capacity)))
        self._h = h
        self._ws: dict[tuple, torch.Tensor] = {}
        self._ctx: dict[tuple, C.c_void_p] = {}
        self._ctx_timing: dict[int, bool] = {}
        self._lock = threading.Lock()

    # -------------------------------------------------------------------------------- lifecycle
    def close(self) -> None:
        for ctx in self._ctx.values():
            lib.ts_ctx_destroy(ctx)
        self._ctx.clear()
        self._ctx_timing.clear()
        for h, _k in getattr(self, "_xchg1", {}).values():
            lib.ts_xchg_destroy(h)
        self._xchg1 = {}
        if getattr(self, "_h", None):
            lib.ts_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(lib.ts_index_size(self._h))

    @property
    def capacity(self) -> int:
        return int(lib.ts_index_capacity(self._h))

    @property
    def row_bytes(self) -> int:
        return int(lib.ts_index_row_bytes(self._h))

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    # -------------------------------------------------------------------------------- building
    def add(self, rows, ids=None, normalize: bool = True) -> "TheoremIndex":
        """Append rows [n, dim]. CUDA tensors are consumed in place (K1 on the current stream);
        numpy arrays / CPU tensors go through the chunked host path. ``ids``: int64 [n]. The store grows when
        it is full; built IVF lists stay valid (new rows go to overflow lists)."""
        if isinstance(rows, torch.Tensor) and rows.is_cuda:
            if rows.device != self.device:
                raise _lib.TheoremSearchError(-1, f"rows on {rows.device}, index on {self.device}")
            if rows.dim() != 2 or rows.shape[1] != self.dim:
                raise _lib.TheoremSearchError(-1, f"rows must be [n, {self.dim}], got {tuple(rows.shape)}")
            if rows.dtype not in _TORCH_TO_TS:
                rows = rows.to(torch.float32)
            rows = rows.contiguous()
            id_ptr = None
            if ids is not None:
                ids = torch.as_tensor(ids, dtype=torch.int64, device=self.device).contiguous()
                id_ptr = ids.data_ptr()
            check(lib.ts_index_add(self._h, rows.data_ptr(), _TORCH_TO_TS[rows.dtype], rows.shape[0],
                                   int(normalize), id_ptr, _stream_ptr(self.device)))
            # the kernel reads `rows`/`ids` asynchronously: keep them alive until the stream drains
            torch.cuda.current_stream(self.device).synchronize()
            return self
        arr = rows.detach().cpu() if isinstance(rows, torch.Tensor) else rows
        if isinstance(arr, torch.Tensor):
            if arr.dtype not in _TORCH_TO_TS:
                arr = arr.to(torch.float32)
            code = _TORCH_TO_TS[arr.dtype]
            arr = arr.contiguous()
            n, d, ptr = arr.shape[0], arr.shape[1], arr.data_ptr()
        else:
            arr = np.ascontiguousarray(np.asarray(arr, dtype=np.float32))
            if arr.ndim != 2:
                raise _lib.TheoremSearchError(-1, f"rows must be 2-D, got shape {arr.shape}")
            code = TS_F32
            n, d, ptr = arr.shape[0], arr.shape[1], arr.ctypes.data
        if d != self.dim:
            raise _lib.TheoremSearchError(-1, f"rows must be [n, {self.dim}], got [{n}, {d}]")
        id_arr = None
        if ids is not None:
            id_arr = np.ascontiguousarray(np.asarray(ids, dtype=np.int64))
            if id_arr.shape != (n,):
                raise _lib.TheoremSearchError(-1, f"ids must be [{n}], got {id_arr.shape}")
        check(lib.ts_index_add_host(self._h, ptr, code, n, int(normalize),
                                    id_arr.ctypes.data if id_arr is not None else None))
        return self

    def reserve(self, capacity: int) -> "TheoremIndex":
        """Make room for ``capacity`` rows now (``add`` / ``upsert`` also grow the store on demand)."""
        check(lib.ts_index_reserve(self._h, int(capacity)))
        return self

    def upsert(self, rows, ids, normalize: bool = True) -> int:
        """``INSERT ... ON CONFLICT (slogan_id) DO UPDATE SET embedding = EXCLUDED.embedding`` — the reference
        writer's statement (ec2/generate_embeddings/__main__.py:84-101 through ec2/rds/upsert.py:29-52, executed
        row by row): a row whose id is already stored is replaced in place, any other is appended; a repeated id
        keeps its last occurrence. ``rows`` [n, dim]: CUDA tensor (consumed in place), CPU tensor or numpy;
        ``ids``: n int64 values (host). Built IVF lists stay valid. Returns the number of stored rows replaced."""
        id_arr = np.ascontiguousarray(np.asarray(ids.cpu() if isinstance(ids, torch.Tensor) else ids, dtype=np.int64))
        replaced = C.c_int64(0)
        if isinstance(rows, torch.Tensor) and rows.is_cuda:
            if rows.device != self.device:
                raise _lib.TheoremSearchError(-1, f"rows on {rows.device}, index on {self.device}")
            if rows.dim() != 2 or rows.shape[1] != self.dim:
                raise _lib.TheoremSearchError(-1, f"rows must be [n, {self.dim}], got {tuple(rows.shape)}")
            if rows.dtype not in _TORCH_TO_TS:
                rows = rows.to(torch.float32)
            rows = rows.contiguous()
            if id_arr.shape != (rows.shape[0],):
                raise _lib.TheoremSearchError(-1, f"ids must be [{rows.shape[0]}], got {id_arr.shape}")
            check(lib.ts_index_upsert(self._h, rows.data_ptr(), _TORCH_TO_TS[rows.dtype], rows.shape[0], int(normalize),
                                      id_arr.ctypes.data, C.byref(replaced), _stream_ptr(self.device)))
            return int(replaced.value)
        arr = rows.detach().cpu() if isinstance(rows, torch.Tensor) else rows
        if isinstance(arr, torch.Tensor):
            if arr.dtype not in _TORCH_TO_TS:
                arr = arr.to(torch.float32)
            code, arr = _TORCH_TO_TS[arr.dtype], arr.contiguous()
            n, d, ptr = arr.shape[0], arr.shape[1], arr.data_ptr()
        else:
            arr = np.ascontiguousarray(np.asarray(arr, dtype=np.float32))
            if arr.ndim != 2:
                raise _lib.TheoremSearchError(-1, f"rows must be 2-D, got shape {arr.shape}")
            code, (n, d), ptr = TS_F32, arr.shape, arr.ctypes.data
        if d != self.dim or id_arr.shape != (n,):
            raise _lib.TheoremSearchError(-1, f"rows must be [n, {self.dim}] with n ids, got {(n, d)} and {id_arr.shape}")
        check(lib.ts_index_upsert_host(self._h, ptr, code, n, int(normalize), id_arr.ctypes.data, C.byref(replaced)))
        return int(replaced.value)

    def delete(self, ids, return_moves: bool = False):
        """``DELETE`` by id — what the cascade of the reference's re-parse does to the corpus table
        (``DELETE FROM theorem WHERE paper_id = ANY(%s)``, ec2/parse_arxiv_papers/__main__.py:271-274;
        ``ON DELETE CASCADE`` down to theorem_embedding_qwen, rds_schema.sql:35,46,51). Ids that are not stored
        match nothing. The store stays dense: the last rows move into the freed slots. Returns the number of rows
        deleted; with ``return_moves=True`` also the relocations as an int64 [m, 2] array of (old row, new row)
        for callers that keep row-aligned side tables (``TheoremStore``, allow masks). Built IVF lists stay valid."""
        id_arr = np.ascontiguousarray(np.asarray(ids.cpu() if isinstance(ids, torch.Tensor) else ids, dtype=np.int64)).reshape(-1)
        n = int(id_arr.shape[0])
        deleted, moved = C.c_int64(0), C.c_int64(0)
        src = np.empty(max(n, 1), dtype=np.int64)
        dst = np.empty(max(n, 1), dtype=np.int64)
        check(lib.ts_index_delete(self._h, id_arr.ctypes.data, n, C.byref(deleted), src.ctypes.data, dst.ctypes.data,
                                  C.byref(moved), _stream_ptr(self.device)))
        if return_moves:
            m = int(moved.value)
            return int(deleted.value), np.stack([src[:m], dst[:m]], axis=1)
        return int(deleted.value)

    def get_rows(self, first: int = 0, n: Optional[int] = None) -> torch.Tensor:
        """Stored rows dequantised to fp32 (device tensor) — the oracle's 'same inputs'."""
        n = len(self) - first if n is None else n
        out = torch.empty((n, self.dim), dtype=torch.float32, device=self.device)
        check(lib.ts_index_get_rows(self._h, first, n, out.data_ptr(), _stream_ptr(self.device)))
        return out

    # -------------------------------------------------------------------------------- searching
    def _workspace(self, nq: int, k: int) -> torch.Tensor:
        need = int(lib.ts_workspace_bytes(self._h, nq, k))
        # one workspace per (host thread, stream): concurrent searches on one index must not share scratch
        # (Streamlit runs one script thread per session over a cached index, streamlit_app.py:52-59)
        key = (threading.get_ident(), _stream_ptr(self.device), nq, k)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            with self._lock:
                if len(self._ws) >= 16:     # a few shapes per thread may alternate (single queries + one batch size)
                    self._ws.pop(next(iter(self._ws)))
                self._ws[key] = ws
        return ws

    def _prep_queries(self, queries) -> torch.Tensor:
        q = queries if isinstance(queries, torch.Tensor) else torch.as_tensor(np.asarray(queries))
        if q.dim() == 1:
            q = q.unsqueeze(0)
        if q.dim() != 2 or q.shape[1] != self.dim:
            raise _lib.TheoremSearchError(-1, f"queries must be [nq, {self.dim}], got {tuple(q.shape)}")
        if q.dtype not in _TORCH_TO_TS:
            q = q.to(torch.float32)
        return q.to(self.device).contiguous()

    def _stream_exchange(self, k: int):
        """A world-of-one exchange handle: gives a single GPU the sharded path's kernel chain (scan + finishing kernel
        linked by programmatic dependent launch), so that back-to-back searches overlap their tails."""
        # one handle per (host thread, stream): the handle numbers its searches and chains them in that order
        table = self.__dict__.setdefault("_xchg1", {})
        key = (threading.get_ident(), _stream_ptr(self.device))
        h, cap = table.get(key, (None, 0))
        if h is None or cap < k:
            if h is not None:
                torch.cuda.current_stream(self.device).synchronize()
                lib.ts_xchg_destroy(h)
            h, cap = C.c_void_p(), max(int(k), 32)
            check(lib.ts_xchg_create(C.byref(h), self.device.index, 1, 0, 1, cap))
            check(lib.ts_xchg_connect(h, None))
            with self._lock:
                # threads come and go (one script thread per Streamlit rerun): drop the handles of threads that exited
                alive = {t.ident for t in threading.enumerate()}
                doomed = [table.pop(kk) for kk in [kk for kk in table if kk[0] not in alive]]
                table[key] = (h, cap)
            if doomed:
                torch.cuda.synchronize(self.device)
                for old_h, _ in doomed:
                    lib.ts_xchg_destroy(old_h)
        return h

    def search(self, queries, k: int, normalize: bool = True, allow_mask: Optional[torch.Tensor] = None,
               independent: bool = False):
        """Exact top-k on device tensors. Returns (scores float32 [nq, k], ids int64 [nq, k]),
        score descending, ties -> lower row; padded with (-inf, -1).

        ``independent=True`` (single queries): a promise that ``queries`` / ``allow_mask`` are not written by the
        kernel enqueued just before this call on the current stream (e.g. the queries were uploaded earlier). The
        search then runs as a scan kernel plus a finishing kernel chained by programmatic dependent launch, and the
        scan of the NEXT such search starts on every SM the moment this one's scan CTA retires: a stream of searches
        runs at the streaming rate of the scan, without launch gaps or merge tails between queries."""
        q = self._prep_queries(queries)
        nq = q.shape[0]
        if independent and nq == 1 and k <= _lib.TS_MAX_K and nq < _lib.get_tunable("batch.min_nq"):
            scores = torch.empty((1, k), dtype=torch.float32, device=self.device)
            ids = torch.empty((1, k), dtype=torch.int64, device=self.device)
            ws = self._workspace(1, k)
            check(lib.ts_search_sharded(self._h, self._stream_exchange(k), q.data_ptr(), _TORCH_TO_TS[q.dtype], 1, int(k),
                                        int(normalize), self._mask_ptr(allow_mask), 0, None, scores.data_ptr(),
                                        ids.data_ptr(), ws.data_ptr(), ws.numel(), _lib.TS_SHARDED_INDEPENDENT,
                                        _stream_ptr(self.device)))
            q.record_stream(torch.cuda.current_stream(self.device))
            return scores, ids
        scores = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        ids = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        ws = self._workspace(nq, k)
        mask_ptr = self._mask_ptr(allow_mask)
        check(lib.ts_search(self._h, q.data_ptr(), _TORCH_TO_TS[q.dtype], nq, int(k), int(normalize), mask_ptr,
                            scores.data_ptr(), ids.data_ptr(), ws.data_ptr(), ws.numel(),
                            _stream_ptr(self.device)))
        q.record_stream(torch.cuda.current_stream(self.device))
        return scores, ids

    def search_keys(self, queries, k: int, normalize: bool = True, allow_mask: Optional[torch.Tensor] = None):
        """This shard's candidates as packed 64-bit keys (int64 tensor holding the uint64 bits)
        [nq, k] — the all-gather payload of the sharded path."""
        q = self._prep_queries(queries)
        nq = q.shape[0]
        keys = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        ws = self._workspace(nq, k)
        check(lib.ts_search_keys(self._h, q.data_ptr(), _TORCH_TO_TS[q.dtype], nq, int(k), int(normalize),
                                 self._mask_ptr(allow_mask), keys.data_ptr(), ws.data_ptr(), ws.numel(),
                                 _stream_ptr(self.device)))
        q.record_stream(torch.cuda.current_stream(self.device))
        return keys

    def _mask_ptr(self, allow_mask):
        if allow_mask is None:
            return None
        words = (len(self) + 31) // 32
        if (not allow_mask.is_cuda or allow_mask.device != self.device or allow_mask.dtype != torch.int32
                or allow_mask.numel() < words or not allow_mask.is_contiguous()):
            raise _lib.TheoremSearchError(
                -1, f"allow_mask must be a contiguous int32 CUDA tensor of >= {words} words on {self.device} "
                    "(see pack_allow_mask)")
        return allow_mask.data_ptr()

    def _get_ctx(self, nq: int, k: int) -> C.c_void_p:
        """A ts_ctx (stream + pinned staging + workspace) of this host thread: one ctx per thread makes
        concurrent ``search_host`` calls on one index safe (include/theoremsearch.h, ts_ctx_create).
        Streamlit starts a fresh script thread per rerun over the cached index (streamlit_app.py:52-59),
        so contexts of threads that have exited are destroyed here, and a larger context replaces the
        smaller ones of the same thread: the table stays bounded by live threads."""
        tid = threading.get_ident()
        nq, k = max(int(nq), 1), max(int(k), 1)
        with self._lock:
            for (t, mq, mk), ctx in self._ctx.items():
                if t == tid and nq <= mq and k <= mk:
                    return ctx
            alive = {t.ident for t in threading.enumerate()}
            stale = [key for key in self._ctx
                     if key[0] not in alive or (key[0] == tid and key[1] <= nq and key[2] <= k)]
            doomed = [self._ctx.pop(key) for key in stale]
        for ctx in doomed:          # no live thread can be inside a call on these
            self._ctx_timing.pop(ctx.value, None)
            lib.ts_ctx_destroy(ctx)
        ctx = C.c_void_p()
        check(lib.ts_ctx_create(C.byref(ctx), self._h, nq, k))
        with self._lock:
            self._ctx[(tid, nq, k)] = ctx
        return ctx

    def search_host(self, queries: np.ndarray, k: int, normalize: bool = True,
                    allow_mask: Optional[torch.Tensor] = None, timing: bool = False):
        """End-to-end call with HOST buffers (what a UI callback does): numpy fp32 queries in,
        numpy (scores [nq, k], ids [nq, k]) out; H2D, search, D2H and the synchronise happen
        inside the C call (``ts_search_host``)."""
        q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32))
        if q.ndim == 1:
            q = q[None, :]
        if q.shape[1] != self.dim:
            raise _lib.TheoremSearchError(-1, f"queries must be [nq, {self.dim}], got {q.shape}")
        nq = q.shape[0]
        ctx = self._get_ctx(nq, k)
        if self._ctx_timing.get(ctx.value, False) != bool(timing):      # not an FFI call per query
            check(lib.ts_ctx_set_timing(ctx, 1 if timing else 0))
            self._ctx_timing[ctx.value] = bool(timing)
        scores = np.empty((nq, k), dtype=np.float32)
        ids = np.empty((nq, k), dtype=np.int64)
        check(lib.ts_search_host(ctx, q.ctypes.data, nq, int(k), int(normalize), self._mask_ptr(allow_mask),
                                 scores.ctypes.data, ids.ctypes.data))
        if timing:
            self.last_kernel_ms = float(lib.ts_ctx_last_kernel_ms(ctx))
        return scores, ids


    # -------------------------------------------------------------------------------- IVF-Flat (K4)
    def ivf_train(self, nlist: int, sample: Optional[torch.Tensor] = None, n_sample: int = 0, iters: int = 10,
                  seed: int = 0) -> "TheoremIndex":
        """Spherical k-means -> ``nlist`` centroids (pgvector ivfflat's training step).  ``sample``:
        CUDA fp32 [n, dim], or None to train on every (len/n_sample)-th stored row in place."""
        if int(nlist) != 1:
            self.scan_dtype = self.dtype     # real IVF lists replace the one-list e4m3 scan copy
        ptr, n = None, int(n_sample)
        if sample is not None:
            sample = sample.to(self.device, torch.float32).contiguous()
            if sample.dim() != 2 or sample.shape[1] != self.dim:
                raise _lib.TheoremSearchError(-1, f"sample must be [n, {self.dim}], got {tuple(sample.shape)}")
            ptr, n = sample.data_ptr(), sample.shape[0]
        check(lib.ts_ivf_train(self._h, ptr, n, int(nlist), int(iters), int(seed), _stream_ptr(self.device)))
        return self

    def ivf_set_centroids(self, centroids: torch.Tensor) -> "TheoremIndex":
        if int(centroids.shape[0]) != 1:
            self.scan_dtype = self.dtype
        c = centroids.to(self.device, torch.float32).contiguous()
        if c.dim() != 2 or c.shape[1] != self.dim:
            raise _lib.TheoremSearchError(-1, f"centroids must be [nlist, {self.dim}], got {tuple(c.shape)}")
        check(lib.ts_ivf_set_centroids(self._h, c.data_ptr(), c.shape[0], _stream_ptr(self.device)))
        torch.cuda.current_stream(self.device).synchronize()
        return self

    @property
    def nlist(self) -> int:
        return int(lib.ts_ivf_nlist(self._h))

    def ivf_centroids(self) -> torch.Tensor:
        out = torch.empty((self.nlist, self.dim), dtype=torch.float32, device=self.device)
        check(lib.ts_ivf_get_centroids(self._h, out.data_ptr(), _stream_ptr(self.device)))
        return out

    def ivf_build(self, list_dtype: str = "fp8") -> "TheoremIndex":
        """File every stored row under its nearest centroid; lists stored as e4m3 (+ per-row scale) or bf16."""
        code = {"fp8": _lib.TS_FP8_E4M3, "fp8_e4m3": _lib.TS_FP8_E4M3, "e4m3": _lib.TS_FP8_E4M3, "bf16": TS_BF16}[list_dtype]
        check(lib.ts_ivf_build(self._h, code, _stream_ptr(self.device)))
        self._ws = {}
        return self

    def ivf_pending(self) -> tuple[int, int]:
        """(rows filed in overflow lists, tombstoned list positions) since the last build / re-pack."""
        a, b = C.c_int64(0), C.c_int64(0)
        check(lib.ts_ivf_pending(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def ivf_repack(self) -> "TheoremIndex":
        """Re-file every row under the current centroids now (clears overflow lists and tombstones)."""
        check(lib.ts_ivf_repack(self._h, _stream_ptr(self.device)))
        return self

    def ivf_lists(self):
        """(offsets int64 [nlist+1], rows int64 [len]): list l holds corpus rows rows[offsets[l]:offsets[l+1]]."""
        off = torch.empty(self.nlist + 1, dtype=torch.int64, device=self.device)
        rows = torch.empty(len(self), dtype=torch.int64, device=self.device)
        check(lib.ts_ivf_get_lists(self._h, off.data_ptr(), rows.data_ptr(), _stream_ptr(self.device)))
        return off, rows

    def ivf_list_sizes(self) -> torch.Tensor:
        out = torch.empty(self.nlist, dtype=torch.int64, device=self.device)
        check(lib.ts_ivf_list_sizes(self._h, out.data_ptr(), _stream_ptr(self.device)))
        return out

    def ivf_list_data(self, first: int = 0, n: Optional[int] = None) -> torch.Tensor:
        """List rows at positions [first, first+n) dequantised to fp32 (what the list scan scores)."""
        n = len(self) - first if n is None else n
        out = torch.empty((n, self.dim), dtype=torch.float32, device=self.device)
        check(lib.ts_ivf_get_list_data(self._h, first, n, out.data_ptr(), _stream_ptr(self.device)))
        return out

    # -------------------------------------------------------------------------------- fp8 exhaustive scan
    def build_fp8_shadow(self) -> "TheoremIndex":
        """An e4m3 copy of the whole corpus (one inverted list holding every row, in row order) for
        BASELINE.json's "optionally fp8-e4m3, rescored in fp32" single-query scan: half the bytes per query."""
        if self.nlist > 1:
            raise _lib.TheoremSearchError(
                -6, f"build_fp8_shadow: the index holds {self.nlist} IVF lists; the e4m3 scan copy is the one-list "
                    "special case of the same structure and would replace them")
        self.ivf_train(1, n_sample=1, iters=0)
        return self.ivf_build("fp8")

    @property
    def fp8_scan_ready(self) -> bool:
        """True while the one-list e4m3 scan copy covers every stored row (``add`` invalidates it)."""
        return self.nlist == 1 and int(lib.ts_ivf_list_dtype(self._h)) == _lib.TS_FP8_E4M3

    def search_fp8(self, queries, k: int, rescore_k: int = 128, normalize: bool = True,
                   allow_mask: Optional[torch.Tensor] = None):
        """Exhaustive scan of the e4m3 shadow (K4b), exact re-score of the ``rescore_k`` best against the bf16
        rows (K4c).  Scores are exact for the returned rows; the id set is the exact top-k unless a true
        neighbour falls outside the e4m3 top-``rescore_k`` (reported as recall by ``bench_extra.py fp8-scan``)."""
        return self.ivf_search(queries, k, nprobe=1, rescore_k=rescore_k, normalize=normalize, allow_mask=allow_mask)

    def _ivf_workspace(self, nq: int, k: int, nprobe: int, rescore_k: int) -> torch.Tensor:
        need = int(lib.ts_ivf_workspace_bytes(self._h, nq, k, nprobe, rescore_k))
        key = ("ivf", threading.get_ident(), _stream_ptr(self.device), nq, k, nprobe, rescore_k)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(max(need, 256), dtype=torch.uint8, device=self.device)
            with self._lock:
                if len(self._ws) >= 16:
                    self._ws.pop(next(iter(self._ws)))
                self._ws[key] = ws
        return ws

    def ivf_search(self, queries, k: int, nprobe: int = 32, rescore_k: int = 100, normalize: bool = True,
                   allow_mask: Optional[torch.Tensor] = None):
        """ANN top-k: probe the ``nprobe`` nearest lists, keep ``rescore_k`` candidates by list-precision
        score, re-score them exactly.  Returns (scores [nq, k], ids [nq, k]) like :meth:`search`."""
        q = self._prep_queries(queries)
        nq = q.shape[0]
        scores = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        ids = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        ws = self._ivf_workspace(nq, k, nprobe, rescore_k)
        check(lib.ts_ivf_search(self._h, q.data_ptr(), _TORCH_TO_TS[q.dtype], nq, int(k), int(nprobe), int(rescore_k),
                                int(normalize), self._mask_ptr(allow_mask), scores.data_ptr(), ids.data_ptr(),
                                ws.data_ptr(), ws.numel(), _stream_ptr(self.device)))
        q.record_stream(torch.cuda.current_stream(self.device))
        return scores, ids

    def ivf_search_host(self, queries: np.ndarray, k: int, nprobe: int = 32, rescore_k: int = 100,
                        normalize: bool = True, allow_mask: Optional[torch.Tensor] = None):
        """Host buffers in / out like :meth:`search_host`, over the IVF path (``ts_ivf_search_host``)."""
        q = np.ascontiguousarray(np.asarray(queries, dtype=np.float32))
        if q.ndim == 1:
            q = q[None, :]
        if q.shape[1] != self.dim:
            raise _lib.TheoremSearchError(-1, f"queries must be [nq, {self.dim}], got {q.shape}")
        nq = q.shape[0]
        ctx = self._get_ctx(nq, k)
        scores = np.empty((nq, k), dtype=np.float32)
        ids = np.empty((nq, k), dtype=np.int64)
        check(lib.ts_ivf_search_host(ctx, q.ctypes.data, nq, int(k), int(nprobe), int(rescore_k), int(normalize),
                                     self._mask_ptr(allow_mask), scores.ctypes.data, ids.ctypes.data))
        return scores, ids

    def ivf_search_keys(self, queries, k: int, nprobe: int = 32, rescore_k: int = 100, normalize: bool = True,
                        allow_mask: Optional[torch.Tensor] = None):
        q = self._prep_queries(queries)
        nq = q.shape[0]
        keys = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        ws = self._ivf_workspace(nq, k, nprobe, rescore_k)
        check(lib.ts_ivf_search_keys(self._h, q.data_ptr(), _TORCH_TO_TS[q.dtype], nq, int(k), int(nprobe),
                                     int(rescore_k), int(normalize), self._mask_ptr(allow_mask), keys.data_ptr(),
                                     ws.data_ptr(), ws.numel(), _stream_ptr(self.device)))
        q.record_stream(torch.cuda.current_stream(self.device))
        return keys


def merge_topk(keys: torch.Tensor, k: int, shard_base: Optional[Sequence[int] | torch.Tensor] = None,
               id_map: Optional[torch.Tensor] = None):
    """K5 on gathered candidates: keys int64 [nshards, nq, k] (packed uint64 bits) ->
    (scores [nq, k], ids [nq, k]) with rows rebased by ``shard_base``."""
    _require_cuda()
    if keys.dim() != 3 or keys.dtype != torch.int64 or not keys.is_cuda:
        raise _lib.TheoremSearchError(-1, "keys must be an int64 CUDA tensor [nshards, nq, k]")
    keys = keys.contiguous()
    nshards, nq, kk = keys.shape
    if kk != k:
        raise _lib.TheoremSearchError(-1, f"keys last dim {kk} != k {k}")
    dev = keys.device
    base = None
    if shard_base is not None:
        base = torch.as_tensor(shard_base, dtype=torch.int64).to(dev).contiguous()
    scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(lib.ts_merge_topk(keys.data_ptr(), nshards, nq, k, base.data_ptr() if base is not None else None,
                                id_map.data_ptr() if id_map is not None else None, scores.data_ptr(),
                                ids.data_ptr(), _stream_ptr(dev)))
    return scores, ids
