"""Seeded synthetic corpora / queries of the BASELINE.json shapes, generated ON DEVICE in fixed
1 Mi-row chunks (chunk c uses seed + c) so the content does not depend on how many GPUs share
the corpus.  Device-side generation is for benchmarks and full-size GPU tests;
small parity tests use the CPU generator in ``oracle/oracle.py`` and upload the rows."""
from __future__ import annotations

import torch

from .index import TheoremIndex

CHUNK_ROWS = 1 << 20
QUERY_SEED = 1_000_000


def fill_index(index: TheoremIndex, first_row: int, n_rows: int, seed: int = 0, sub_rows: int = 1 << 18) -> None:
    """Append global rows [first_row, first_row + n_rows) of the synthetic corpus to ``index``
    (raw N(0,1) rows; K1 normalises and quantises them)."""
    dev = index.device
    gen = torch.Generator(device=dev)
    pos = first_row
    end = first_row + n_rows
    while pos < end:
        c, off = divmod(pos, CHUNK_ROWS)
        take = min(end - pos, CHUNK_ROWS - off)
        gen.manual_seed(seed + c)
        # stream the chunk in sub-blocks from its start so content is position-independent
        done = 0
        while done < off + take:
            m = min(sub_rows, off + take - done)
            blk = torch.randn((m, index.dim), generator=gen, dtype=torch.float32, device=dev)
            lo = max(off - done, 0)
            if lo < m:
                index.add(blk[lo:], normalize=True)
            done += m
        pos += take


def make_queries(nq: int, dim: int, device, seed: int = QUERY_SEED) -> torch.Tensor:
    gen = torch.Generator(device=device).manual_seed(seed)
    return torch.randn((nq, dim), generator=gen, dtype=torch.float32, device=device)


def clustered_centers(n_centers: int, dim: int, device, seed: int = 0) -> torch.Tensor:
    gen = torch.Generator(device=device).manual_seed(seed)
    centers = torch.randn((n_centers, dim), generator=gen, dtype=torch.float32, device=device)
    return centers / centers.norm(dim=1, keepdim=True)


def fill_index_clustered(index: TheoremIndex, n_rows: int, n_centers: int, sigma: float = 1.0, seed: int = 0,
                         sub_rows: int = 1 << 18, first_row: int = 0) -> torch.Tensor:
    """Clustered synthetic corpus (i.i.d. Gaussian rows have no list structure, so IVF recall
    on them is a pessimistic bound): row = centre[c] + sigma * N(0, I/dim), c uniform.  Appends global rows
    [first_row, first_row + n_rows); like ``fill_index`` the content of a row depends only on its global
    position (1 Mi-row chunks, chunk c seeded with seed + 1 + c).  Returns the unit centres
    [n_centers, dim] so queries can be drawn from the same mixture."""
    dev = index.device
    centers = clustered_centers(n_centers, index.dim, dev, seed)
    gen = torch.Generator(device=dev)
    pos, end = first_row, first_row + n_rows
    while pos < end:
        c, off = divmod(pos, CHUNK_ROWS)
        take = min(end - pos, CHUNK_ROWS - off)
        gen.manual_seed(seed + 1 + c)
        done = 0
        while done < off + take:
            m = min(sub_rows, off + take - done)
            which = torch.randint(0, n_centers, (m,), generator=gen, device=dev)
            blk = torch.randn((m, index.dim), generator=gen, dtype=torch.float32, device=dev)
            lo = max(off - done, 0)
            if lo < m:
                blk = blk[lo:]
                blk.mul_(sigma / index.dim ** 0.5).add_(centers[which[lo:]])
                index.add(blk, normalize=True)
            done += m
        pos += take
    return centers


def make_clustered_queries(nq: int, centers: torch.Tensor, sigma: float = 1.0, seed: int = QUERY_SEED) -> torch.Tensor:
    dev = centers.device
    gen = torch.Generator(device=dev).manual_seed(seed)
    which = torch.randint(0, centers.shape[0], (nq,), generator=gen, device=dev)
    q = torch.randn((nq, centers.shape[1]), generator=gen, dtype=torch.float32, device=dev)
    return q * (sigma / centers.shape[1] ** 0.5) + centers[which]
