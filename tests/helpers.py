"""Test-only stand-ins. ``OracleIndex`` answers the ``TheoremIndex.search_host`` contract from
the CPU oracle so host-side logic (filters, rows, re-rank, sharding) can be tested without a GPU.
It is NOT part of the product and lives under tests/ on purpose."""
import datetime
import json
import os

import numpy as np
import torch

from oracle import oracle


def unpack_allow_mask(mask, n):
    if mask is None:
        return None
    w = np.asarray(mask.cpu().numpy()).view(np.uint32)
    bits = np.unpackbits(w.view(np.uint8), bitorder="little")
    return bits[:n].astype(bool)


class OracleIndex:
    device = "cpu"

    def __init__(self, rows, quantize=None):
        rows = np.asarray(rows, dtype=np.float32)
        self.rows = oracle.bf16_round(rows) if quantize == "bf16" else rows

    def __len__(self):
        return self.rows.shape[0]

    def search_host(self, queries, k, normalize=True, allow_mask=None, timing=False):
        q = np.asarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if normalize:
            q = oracle.normalize_f64(q)
        s, i = oracle.exact_search(q, self.rows, k, allow=unpack_allow_mask(allow_mask, len(self)))
        return s.astype(np.float32), i


class TableModel:
    """model.encode stand-in: fixed vectors per query string (same contract the golden script used)."""

    def __init__(self, table):
        self.table = table

    def encode(self, text, convert_to_tensor=False, normalize_embeddings=False, convert_to_numpy=False, **kw):
        v = np.asarray(self.table[text], dtype=np.float32)
        if normalize_embeddings:
            v = oracle.normalize(v).numpy()[0]
        return torch.from_numpy(v) if convert_to_tensor else v


def load_golden(name):
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with open(os.path.join(here, name + ".json")) as f:
        return json.load(f)


def golden_store_rows(g):
    rows = []
    for r in g["rows"]:
        r = list(r)
        r[4] = datetime.datetime.fromisoformat(r[4]) if r[4] else None
        rows.append(tuple(r))
    return rows


BASE_FILTERS = {"sources": ["arXiv", "Stacks Project"], "authors": [], "tags": [], "year_range": None,
                "journal_status": "All", "types": [], "citation_range": (0, 10**9),
                "include_unknown_citations": True, "paper_filter": {"ids": set(), "titles": set()}}
