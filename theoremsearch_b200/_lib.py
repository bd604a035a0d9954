"""ctypes binding of ``libtheoremsearch.so`` (C ABI in ``include/theoremsearch.h``).

The library is the product path; there is no Python/CPU fallback.  If the shared object has
not been built, importing this module raises ``LibraryNotBuilt`` with the build command, and
every compute entry point raises ``TheoremSearchError`` when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtheoremsearch.so")

TS_F32, TS_BF16, TS_FP8_E4M3, TS_F16 = 0, 1, 2, 3
TS_MAX_K = 1024
TS_SHARDED_INDEPENDENT, TS_SHARDED_ONE_KERNEL = 1, 2
TS_MAX_DIM = 2048

TS_OK = 0
ERROR_NAMES = {
    -1: "TS_ERR_BAD_ARG",
    -2: "TS_ERR_CUDA",
    -3: "TS_ERR_OOM",
    -4: "TS_ERR_UNSUPPORTED",
    -5: "TS_ERR_CAPACITY",
    -6: "TS_ERR_STATE",
}


class LibraryNotBuilt(ImportError):
    pass


class TheoremSearchError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_u64 = C.c_uint64
_sz = C.c_size_t

# name -> (restype, argtypes); kept in one table so tests can check it against the header.
SIGNATURES = {
    "ts_abi_version": (_i, []),
    "ts_last_error": (C.c_char_p, []),
    "ts_kernel_launches": (_u64, []),
    "ts_index_create": (_i, [C.POINTER(_p), _i, _i, _i, _i64]),
    "ts_index_destroy": (None, [_p]),
    "ts_index_add": (_i, [_p, _p, _i, _i64, _i, _p, _p]),
    "ts_index_add_host": (_i, [_p, _p, _i, _i64, _i, _p]),
    "ts_index_reserve": (_i, [_p, _i64]),
    "ts_index_grows_in_place": (_i, [_p]),
    "ts_index_upsert": (_i, [_p, _p, _i, _i64, _i, _p, _p, _p]),
    "ts_index_upsert_host": (_i, [_p, _p, _i, _i64, _i, _p, _p]),
    "ts_index_delete": (_i, [_p, _p, _i64, _p, _p, _p, _p, _p]),
    "ts_ivf_pending": (_i, [_p, _p, _p]),
    "ts_ivf_repack": (_i, [_p, _p]),
    "ts_index_size": (_i64, [_p]),
    "ts_index_capacity": (_i64, [_p]),
    "ts_index_dim": (_i, [_p]),
    "ts_index_dtype": (_i, [_p]),
    "ts_index_device": (_i, [_p]),
    "ts_index_get_rows": (_i, [_p, _i64, _i64, _p, _p]),
    "ts_index_data": (_p, [_p]),
    "ts_index_row_bytes": (_sz, [_p]),
    "ts_index_has_ids": (_i, [_p]),
    "ts_index_read_raw_host": (_i, [_p, _i64, _i64, _p, _p]),
    "ts_index_append_raw_host": (_i, [_p, _p, _i64, _p]),
    "ts_ivf_list_dtype": (_i, [_p]),
    "ts_workspace_bytes": (_sz, [_p, _i, _i]),
    "ts_search": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "ts_search_keys": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "ts_merge_topk": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "ts_xchg_create": (_i, [C.POINTER(_p), _i, _i, _i, _i, _i]),
    "ts_xchg_destroy": (None, [_p]),
    "ts_xchg_handle_bytes": (_i, []),
    "ts_xchg_handle": (_i, [_p, _p]),
    "ts_xchg_connect": (_i, [_p, _p]),
    "ts_xchg_error": (_i, [_p]),
    "ts_xchg_set_timeout_ms": (_i, [_p, _i64]),
    "ts_xchg_reset": (_i, [_p]),
    "ts_xchg_seq": (C.c_uint32, [_p]),
    "ts_search_sharded": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _i64, _p, _p, _p, _p, _sz, _i, _p]),
    "ts_search_sharded_host": (_i, [_p, _p, _p, _i, _i, _i, _p, _i64, _p, _p, _p]),
    "ts_pack_key": (_u64, [C.c_float, C.c_uint32]),
    "ts_unpack_key": (None, [_u64, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]),
    "ts_ctx_create": (_i, [C.POINTER(_p), _p, _i, _i]),
    "ts_ctx_destroy": (None, [_p]),
    "ts_search_host": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "ts_ctx_set_timing": (_i, [_p, _i]),
    "ts_ctx_last_kernel_ms": (C.c_float, [_p]),
    "ts_ivf_train": (_i, [_p, _p, _i64, _i, _i, _u64, _p]),
    "ts_ivf_build": (_i, [_p, _i, _p]),
    "ts_ivf_workspace_bytes": (_sz, [_p, _i, _i, _i, _i]),
    "ts_ivf_search": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "ts_ivf_search_keys": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "ts_ivf_set_centroids": (_i, [_p, _p, _i, _p]),
    "ts_ivf_get_centroids": (_i, [_p, _p, _p]),
    "ts_ivf_get_lists": (_i, [_p, _p, _p, _p]),
    "ts_ivf_get_list_data": (_i, [_p, _i64, _i64, _p, _p]),
    "ts_ivf_search_host": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "ts_ivf_nlist": (_i, [_p]),
    "ts_ivf_list_sizes": (_i, [_p, _p, _p]),
    "ts_eval_create": (_i, [C.POINTER(_p), _i, _i, _p, _p, _p, _i]),
    "ts_eval_destroy": (None, [_p]),
    "ts_eval_num_queries": (_i, [_p]),
    "ts_eval_max_relevance": (C.c_double, [_p]),
    "ts_eval_first_query_without_correct_doc": (_i64, [_p]),
    "ts_eval_rankings": (_i, [_p, _p, _i64, _i, _p, _i, C.c_double, _p, _p, _p]),
    "ts_set_tunable": (_i, [C.c_char_p, _i]),
    "ts_get_tunable": (_i, [C.c_char_p, C.POINTER(_i)]),
    "ts_debug_last_batched_fixups": (_i, []),
    "ts_debug_ivf_timeline": (_i, [_p, _i]),
    "ts_debug_scan_timeline": (_i, [_p, _i, _i]),
}


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise LibraryNotBuilt(
            f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C theoremsearch_b200/csrc -j`. There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == the .so is stale: rebuild
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def last_error() -> str:
    return lib.ts_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != TS_OK:
        raise TheoremSearchError(rc, last_error())


def kernel_launches() -> int:
    return int(lib.ts_kernel_launches())


_tunables: dict[str, int] = {}      # values read or set through this module (every writer goes through set_tunable)


def set_tunable(name: str, value: int) -> None:
    check(lib.ts_set_tunable(name.encode(), int(value)))
    _tunables[name] = int(value)


def get_tunable(name: str) -> int:
    cached = _tunables.get(name)
    if cached is not None:          # on the per-query path (ShardedIndex.search): no FFI call for a constant
        return cached
    v = _i(0)
    check(lib.ts_get_tunable(name.encode(), C.byref(v)))
    _tunables[name] = v.value
    return v.value


def last_batched_fixups() -> int:
    return int(lib.ts_debug_last_batched_fixups())


def pack_key(score: float, row: int) -> int:
    return int(lib.ts_pack_key(C.c_float(score), C.c_uint32(row)))


def unpack_key(key: int) -> tuple[float, int]:
    s = C.c_float()
    r = C.c_uint32()
    lib.ts_unpack_key(_u64(key), C.byref(s), C.byref(r))
    return s.value, r.value
