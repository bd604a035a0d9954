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
