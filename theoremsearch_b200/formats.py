"""Corpus on-disk / wire formats either side of the search path.

* ``load_embedding_library`` / ``save_embedding_library`` — the showcase app's pair of files,
  ``corpus_embeddings.pt`` (``torch.save`` of the fp32 [N, D] tensor) + ``theorems_data.pkl``
  (``app_create_embeddings.py:85-93`` writes them, ``app_showcase_model.py:41-58`` loads them).
* ``parse_pgvector_text`` / ``index_from_embedding_stream`` — rows of
  ``theorem_embedding_qwen(slogan_id, embedding vector(1024))`` as a server-side cursor yields them
  (``experiments/pca_plotting.py:75-89``: batches of ``np.float32[b, D]``; without pgvector's adapter
  psycopg2 returns the ``'[x,y,...]'`` text form).
* ``latest_slogan_per_theorem`` — the ``DISTINCT ON (theorem_id) ... ORDER BY theorem_id,
  slogan_id DESC`` CTE of ``streamlit_app.py:254-259`` plus the ``slogan_id -> theorem_id`` map
  (``rds_schema.sql:33-36``), so the index can be built over one embedding per theorem and return
  theorem ids.
* ``save_index`` / ``load_index`` — the QUANTISED index (bf16/fp32 rows exactly as stored, ids, IVF
  centroids), so a 20 GB corpus is not re-normalised and re-quantised on every process start.
"""
from __future__ import annotations

import json
import os
import pickle
import struct
from typing import Iterable, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from .api import build_index
from .index import TheoremIndex

EMBEDDINGS_FILE = "corpus_embeddings.pt"   # app_create_embeddings.py:86
METADATA_FILE = "theorems_data.pkl"        # app_create_embeddings.py:87
INDEX_MAGIC = b"TSIDX001"
_CHUNK_BYTES = 64 << 20


# ------------------------------------------------------------------------------ showcase library
def save_embedding_library(directory: str, corpus_embeddings, theorems_data: Sequence[dict]) -> None:
    """Write the two files exactly as ``create_embedding_library`` does (app_create_embeddings.py:84-93)."""
    os.makedirs(directory, exist_ok=True)
    emb = corpus_embeddings if isinstance(corpus_embeddings, torch.Tensor) else torch.as_tensor(np.asarray(corpus_embeddings))
    torch.save(emb, os.path.join(directory, EMBEDDINGS_FILE))
    with open(os.path.join(directory, METADATA_FILE), "wb") as f:
        pickle.dump(list(theorems_data), f)


def load_embedding_library(directory: str, device: int | str | torch.device = 0, dtype: str = "bf16",
                           as_index: bool = True):
    """``load_embedding_library`` (app_showcase_model.py:41-58): returns ``(embeddings, theorems_data)``
    or ``(None, None)`` when a file is missing, as the reference does.  With ``as_index`` the embeddings
    come back as a resident ``TheoremIndex`` (row i <-> ``theorems_data[i]``) instead of a CPU tensor."""
    embeddings_path = os.path.join(directory, EMBEDDINGS_FILE)
    data_path = os.path.join(directory, METADATA_FILE)
    if not os.path.exists(embeddings_path) or not os.path.exists(data_path):
        return None, None
    embeddings = torch.load(embeddings_path, map_location=torch.device("cpu"))
    with open(data_path, "rb") as f:
        theorems_data = pickle.load(f)
    if embeddings.dim() != 2 or embeddings.shape[0] != len(theorems_data):
        raise _lib.TheoremSearchError(-1, f"library mismatch: embeddings {tuple(embeddings.shape)} vs "
                                          f"{len(theorems_data)} metadata rows")
    if not as_index:
        return embeddings, theorems_data
    return build_index(embeddings, dtype=dtype, normalize=True, device=device), theorems_data


# ------------------------------------------------------------------------------ pgvector rows
def parse_pgvector_text(value) -> np.ndarray:
    """One ``vector`` column value -> float32 [D].  Accepts pgvector's text form ``'[0.1,0.2,...]'``,
    a list / ndarray (what the registered adapter returns), or bytes of the text form."""
    if isinstance(value, (bytes, bytearray)):
        value = value.decode("ascii")
    if isinstance(value, str):
        s = value.strip()
        if not (s.startswith("[") and s.endswith("]")):
            raise ValueError(f"not a pgvector literal: {value[:40]!r}")
        body = s[1:-1].strip()
        return np.array([float(t) for t in body.split(",")], dtype=np.float32) if body else np.zeros(0, np.float32)
    return np.asarray(value, dtype=np.float32)


def index_from_embedding_stream(batches: Iterable, dim: int, capacity: int, dtype: str = "bf16",
                                device: int | str | torch.device = 0, normalize: bool = True) -> TheoremIndex:
    """Consume a cursor-shaped stream (``stream_joined_embeddings``, experiments/pca_plotting.py:75-89)
    into an index.  Each item is ``X`` or ``(ids, X)`` / ``(X, anything)``: ``X`` = float32 [b, D] or a list
    of pgvector values; ``ids`` = int64 [b] (``slogan_id`` / ``theorem_id``)."""
    index = TheoremIndex(dim, capacity, dtype=dtype, device=device)
    for item in batches:
        ids = None
        x = item
        if isinstance(item, tuple):
            a, b = item[0], item[1]
            a_is_ids = isinstance(a, np.ndarray) and a.ndim == 1 and np.issubdtype(a.dtype, np.integer)
            (ids, x) = (a, b) if a_is_ids else (None, a)
        if not (isinstance(x, np.ndarray) and x.ndim == 2):
            x = np.vstack([parse_pgvector_text(v) for v in x]) if len(x) else np.zeros((0, dim), np.float32)
        if x.shape[0]:
            index.add(np.ascontiguousarray(x, dtype=np.float32), ids=ids, normalize=normalize)
    return index


def latest_slogan_per_theorem(theorem_ids, slogan_ids):
    """``SELECT DISTINCT ON (ts.theorem_id) ... ORDER BY ts.theorem_id, ts.slogan_id DESC``
    (streamlit_app.py:254-259): positions of the rows to keep (one per theorem: its highest slogan_id),
    ordered by theorem_id — and the kept (theorem_id, slogan_id) pairs."""
    t = np.asarray(theorem_ids, dtype=np.int64)
    s = np.asarray(slogan_ids, dtype=np.int64)
    if t.shape != s.shape or t.ndim != 1:
        raise ValueError("theorem_ids and slogan_ids must be 1-D arrays of equal length")
    order = np.lexsort((-s, t))                    # theorem_id asc, slogan_id desc
    first = np.ones(order.size, dtype=bool)
    first[1:] = t[order][1:] != t[order][:-1]
    keep = order[first]
    return keep, t[keep], s[keep]


def index_from_slogan_table(theorem_ids, slogan_ids, embeddings, dtype: str = "bf16",
                            device: int | str | torch.device = 0) -> TheoremIndex:
    """Index over the latest slogan embedding of every theorem, returning THEOREM ids: the join
    ``theorem_embedding_qwen e ON e.slogan_id = latest_slogan.slogan_id`` of streamlit_app.py:279."""
    keep, t_keep, _ = latest_slogan_per_theorem(theorem_ids, slogan_ids)
    emb = embeddings[torch.as_tensor(keep)] if isinstance(embeddings, torch.Tensor) else np.asarray(embeddings)[keep]
    return build_index(emb, ids=t_keep, dtype=dtype, normalize=True, device=device)


# ------------------------------------------------------------------------------ quantised index
def save_index(index: TheoremIndex, path: str) -> None:
    """[magic | u64 header length | JSON header | raw rows | ids | IVF centroids fp32]."""
    n = len(index)
    row_bytes = index.row_bytes
    has_ids = bool(lib.ts_index_has_ids(index.handle))
    nlist = index.nlist
    list_dtype = int(lib.ts_ivf_list_dtype(index.handle))
    header = {"version": 1, "dim": index.dim, "dtype": index.dtype, "rows": n, "row_bytes": row_bytes,
              "has_ids": has_ids, "nlist": nlist if nlist > 0 else 0,
              "ivf_list_dtype": {_lib.TS_BF16: "bf16", _lib.TS_FP8_E4M3: "fp8"}.get(list_dtype),
              "scan_dtype": getattr(index, "scan_dtype", index.dtype)}
    hj = json.dumps(header).encode()
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(INDEX_MAGIC)
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        chunk = max(1, _CHUNK_BYTES // max(row_bytes, 1))
        buf = np.empty(chunk * row_bytes, dtype=np.uint8)
        for first in range(0, n, chunk):
            m = min(chunk, n - first)
            check(lib.ts_index_read_raw_host(index.handle, first, m, buf.ctypes.data, None))
            f.write(memoryview(buf)[: m * row_bytes])
        if has_ids:
            ids = np.empty(n, dtype=np.int64)
            if n:
                check(lib.ts_index_read_raw_host(index.handle, 0, n, None, ids.ctypes.data))
            f.write(ids.tobytes())
        if header["nlist"]:
            f.write(index.ivf_centroids().cpu().numpy().astype(np.float32).tobytes())
    os.replace(tmp, path)


def load_index(path: str, device: int | str | torch.device = 0, capacity: Optional[int] = None) -> TheoremIndex:
    """Inverse of ``save_index``.  Rows are copied back verbatim (bit-identical scores); the IVF lists
    are rebuilt from the saved centroids (assignment is deterministic, a second per 40M rows)."""
    with open(path, "rb") as f:
        if f.read(len(INDEX_MAGIC)) != INDEX_MAGIC:
            raise _lib.TheoremSearchError(-1, f"{path}: not a theoremsearch index file")
        (hl,) = struct.unpack("<Q", f.read(8))
        h = json.loads(f.read(hl).decode())
        n, row_bytes = int(h["rows"]), int(h["row_bytes"])
        index = TheoremIndex(int(h["dim"]), max(capacity or n, n, 1), dtype=h["dtype"], device=device)
        if index.row_bytes != row_bytes:
            raise _lib.TheoremSearchError(-1, f"{path}: row_bytes {row_bytes} != this build's {index.row_bytes}")
        rows_off = f.tell()
        ids = None
        if h["has_ids"] and n:
            f.seek(rows_off + n * row_bytes)
            ids = np.frombuffer(f.read(n * 8), dtype=np.int64)
            f.seek(rows_off)
        chunk = max(1, _CHUNK_BYTES // max(row_bytes, 1))
        for first in range(0, n, chunk):
            m = min(chunk, n - first)
            buf = np.frombuffer(f.read(m * row_bytes), dtype=np.uint8)
            if buf.size != m * row_bytes:
                raise _lib.TheoremSearchError(-1, f"{path}: truncated")
            idp = np.ascontiguousarray(ids[first:first + m]) if ids is not None else None
            check(lib.ts_index_append_raw_host(index.handle, buf.ctypes.data, m,
                                               idp.ctypes.data if idp is not None else None))
        if h.get("nlist"):
            f.seek(rows_off + n * row_bytes + (n * 8 if h["has_ids"] else 0))
            cent = np.frombuffer(f.read(int(h["nlist"]) * int(h["dim"]) * 4), dtype=np.float32)
            cent = torch.from_numpy(cent.reshape(int(h["nlist"]), int(h["dim"])).copy())
            index.ivf_set_centroids(cent)
            if h.get("ivf_list_dtype"):
                index.ivf_build(h["ivf_list_dtype"])
        if h.get("scan_dtype") == "fp8" and index.fp8_scan_ready:
            index.scan_dtype = "fp8"
    return index
