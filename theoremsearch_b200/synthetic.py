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


class HierarchicalCorpus:
    """A synthetic corpus with the structure sentence embeddings have and i.i.d. Gaussian rows lack: a few
    hundred topic blobs, each an ANISOTROPIC low-rank cloud (a rank-``rank`` subspace with a decaying
    spectrum), ``n_leaves`` overlapping leaf clusters inside them (10x more leaves than IVF lists, so k-means
    cannot simply recover the generating centres), plus isotropic noise in all ``dim`` coordinates:

        row = normalize( blob[b] + (leaf_offset[l] + beta * w * spectrum) @ basis[b % n_bases]
                         + gamma * eps / sqrt(dim) ),     l uniform, b = l % n_blobs, w, eps ~ N(0, I)

    Nearest neighbours of a query are the rows closest to it INSIDE a blob's latent cloud, which k-means
    lists cut through — recall@10 rises with nprobe instead of being 1.0 at nprobe = 1 (planted centres) or
    ~0 (isotropic noise).  Like the other generators the content of a row depends only on its global
    position (1 Mi-row chunks, chunk c seeded with seed + 1 + c)."""

    def __init__(self, dim: int, n_leaves: int, device, seed: int = 0, n_blobs: int = 256, rank: int = 48,
                 alpha: float = 0.9, beta: float = 0.45, gamma: float = 0.6, n_bases: int = 16):
        self.dim, self.n_leaves, self.n_blobs, self.rank = dim, n_leaves, n_blobs, rank
        self.alpha, self.beta, self.gamma, self.seed = alpha, beta, gamma, seed
        self.device = torch.device(device)
        gen = torch.Generator(device=self.device).manual_seed(seed)
        r = lambda *shape: torch.randn(shape, generator=gen, dtype=torch.float32, device=self.device)
        blobs = r(n_blobs, dim)
        self.blobs = blobs / blobs.norm(dim=1, keepdim=True)
        self.bases = r(n_bases, rank, dim) / dim ** 0.5                       # rows of norm ~1
        self.spectrum = 1.0 / torch.sqrt(1.0 + torch.arange(rank, dtype=torch.float32, device=self.device))
        self.leaf_latent = alpha * r(n_leaves, rank) * self.spectrum          # leaf offsets in latent space

    def _rows(self, m: int, gen: torch.Generator) -> torch.Tensor:
        dev = self.device
        leaf = torch.randint(0, self.n_leaves, (m,), generator=gen, device=dev)
        w = torch.randn((m, self.rank), generator=gen, dtype=torch.float32, device=dev)
        eps = torch.randn((m, self.dim), generator=gen, dtype=torch.float32, device=dev)
        blob = leaf % self.n_blobs
        latent = self.leaf_latent[leaf] + self.beta * w * self.spectrum
        out = eps.mul_(self.gamma / self.dim ** 0.5)
        out += self.blobs[blob]
        which = blob % self.bases.shape[0]
        for b in range(self.bases.shape[0]):                                  # grouped latent -> dim projection
            sel = (which == b).nonzero(as_tuple=True)[0]
            if sel.numel():
                out.index_add_(0, sel, latent[sel] @ self.bases[b])
        return out

    def fill(self, index: TheoremIndex, n_rows: int, first_row: int = 0, sub_rows: int = 1 << 18) -> None:
        gen = torch.Generator(device=self.device)
        pos, end = first_row, first_row + n_rows
        while pos < end:
            c, off = divmod(pos, CHUNK_ROWS)
            take = min(end - pos, CHUNK_ROWS - off)
            gen.manual_seed(self.seed + 1 + c)
            done = 0
            while done < off + take:
                m = min(sub_rows, off + take - done)
                blk = self._rows(m, gen)
                lo = max(off - done, 0)
                if lo < m:
                    index.add(blk[lo:], normalize=True)
                done += m
            pos += take

    def rows(self, n: int, seed: int) -> torch.Tensor:
        return self._rows(n, torch.Generator(device=self.device).manual_seed(seed))

    def queries(self, nq: int, seed: int = QUERY_SEED) -> torch.Tensor:
        """Fresh draws from the same mixture (un-normalised; the search normalises)."""
        return self.rows(nq, seed)
