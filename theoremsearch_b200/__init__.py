"""theoremsearch_b200 — B200-native retrieval hot path of TheoremSearch.

Everything numeric runs in ``libtheoremsearch.so`` (hand-written sm_100a CUDA, C ABI in
``include/theoremsearch.h``); this package is the thin Python mirror of the reference's search
call shapes.  There is no CPU fallback."""
from ._lib import (LIB_PATH, LibraryNotBuilt, TheoremSearchError, get_tunable, kernel_launches,
                   last_batched_fixups, pack_key, set_tunable, unpack_key)
from .api import build_index, cos_sim_topk, search_theorems
from .index import TheoremIndex, merge_topk, pack_allow_mask
from . import formats, metrics, sharded, store
from .formats import load_embedding_library, load_index, save_embedding_library, save_index

__all__ = [
    "LIB_PATH", "LibraryNotBuilt", "TheoremSearchError", "TheoremIndex", "build_index", "cos_sim_topk",
    "search_theorems", "merge_topk", "pack_allow_mask", "kernel_launches", "set_tunable", "get_tunable",
    "pack_key", "unpack_key", "last_batched_fixups", "formats", "metrics", "sharded", "store", "load_embedding_library",
    "save_embedding_library", "load_index", "save_index",
]
