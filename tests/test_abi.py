"""The C-ABI library loads and exports exactly what include/theoremsearch.h declares, and the
ctypes table agrees with the header. No compute calls (no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "theoremsearch.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = re.findall(r"TS_API\s+([^;{]+?)\s*\(([^;]*?)\)\s*;", src, flags=re.S)
    out = {}
    for head, args in decls:
        name = head.split()[-1].lstrip("*")
        args = " ".join(args.split())
        nargs = 0 if args in ("void", "") else args.count(",") + 1
        out[name] = nargs
    return out


def test_header_declares_the_expected_surface():
    fns = header_functions()
    for must in ("ts_index_create", "ts_index_add", "ts_search", "ts_merge_topk", "ts_workspace_bytes",
                 "ts_last_error", "ts_index_destroy", "ts_ivf_train", "ts_ivf_build", "ts_ivf_search",
                 "ts_search_host", "ts_search_keys"):
        assert must in fns


def test_library_exports_every_declared_symbol():
    import theoremsearch_b200 as ts
    lib = ctypes.CDLL(ts.LIB_PATH)
    fns = header_functions()
    for name in fns:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", ts.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    extra = {e for e in exported if e.startswith("ts_")} - set(fns)
    assert not extra, f"exported but undeclared: {extra}"
    assert not {e for e in exported if not e.startswith("ts_")}, "non-ts_ symbols leak from the library"


def test_ctypes_table_matches_header_arity():
    from theoremsearch_b200 import _lib
    fns = header_functions()
    assert set(_lib.SIGNATURES) == set(fns)
    for name, (_res, args) in _lib.SIGNATURES.items():
        assert len(args) == fns[name], f"{name}: ctypes has {len(args)} args, header {fns[name]}"


def test_no_gpu_calls_fail_loudly_not_silently():
    import torch

    import theoremsearch_b200 as ts
    assert ts._lib.lib.ts_abi_version() == 2
    if torch.cuda.is_available():
        pytest.skip("GPU present: the loud-failure path is for CPU-only hosts")
    with pytest.raises(ts.TheoremSearchError):
        ts.TheoremIndex(1024, 10)
    h = ctypes.c_void_p()
    rc = ts._lib.lib.ts_index_create(ctypes.byref(h), 0, 1024, ts._lib.TS_BF16, 10)
    assert rc == -2 and "no CPU fallback" in ts._lib.last_error()
    import numpy as np
    offsets, docs, rels = np.array([0, 1], np.int64), np.array([3], np.int64), np.array([1.0])
    rc = ts._lib.lib.ts_eval_create(ctypes.byref(h), 0, 1, offsets.ctypes.data, docs.ctypes.data, rels.ctypes.data, 8)
    assert rc == -2 and "no CPU fallback" in ts._lib.last_error()      # K6 has no host path behind the ABI either


def test_key_packing_matches_oracle():
    import numpy as np

    import theoremsearch_b200 as ts
    from oracle import oracle
    rng = np.random.default_rng(0)
    for s in list(rng.standard_normal(50).astype(np.float32)) + [0.0, -0.0, float("inf"), float("-inf"), float("nan")]:
        for r in (0, 1, 12345, 0xFFFFFFFE):
            assert ts.pack_key(float(s), r) == oracle.pack_key(float(s), r)
    assert ts.unpack_key(ts.pack_key(0.25, 77)) == (0.25, 77)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "theoremsearch_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
