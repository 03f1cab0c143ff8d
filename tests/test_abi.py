"""The C-ABI library loads on a CPU-only box and exports every symbol include/clv_b200.h declares;
layout / size queries (host-only code) agree with the oracle's tables.  No compute is called."""
import ctypes
import os
import re

import clvae_b200  # noqa: F401
from clvae_b200 import _lib
from oracle import clv_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "clv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(clv_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    L = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "missing export " + s
        assert s in _lib.PROTOTYPES, "no ctypes prototype for " + s
    assert L.clv_version() >= 100
    assert L.clv_error_string(-2).decode().startswith("unsupported")


def _c_kind(decl):
    """Coarse ctypes class of one C parameter declaration."""
    d = decl.strip()
    if "*" in d:
        return "ptr"
    t = d.rsplit(None, 1)[0] if " " in d else d
    return {"int32_t": "i32", "int": "i32", "int64_t": "i64", "uint64_t": "u64", "float": "f32",
            "double": "f64"}.get(t.replace("const ", "").strip(), t)


def test_ctypes_prototypes_match_the_header_signatures():
    """Every entry point: the number of parameters and the class of each (pointer / int32 / int64 /
    uint64 / float / double) in include/clv_b200.h equals the ctypes prototype of _lib.py (an
    argument added on one side only is silent memory corruption at call time)."""
    src = open(os.path.join(ROOT, "include", "clv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    kinds = {ctypes.c_int32: "i32", ctypes.c_int: "i32", ctypes.c_int64: "i64", ctypes.c_uint64: "u64",
             ctypes.c_float: "f32", ctypes.c_double: "f64", ctypes.c_void_p: "ptr", ctypes.c_char_p: "ptr"}
    checked = 0
    for m in re.finditer(r"\b(?:int|int64_t|const char\*)\s+(clv_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        name, params = m.group(1), m.group(2).strip()
        want = [] if params in ("", "void") else [_c_kind(p) for p in params.split(",")]
        restype, argtypes = _lib.PROTOTYPES[name]
        got = [kinds.get(a, "ptr" if hasattr(a, "contents") or "LP_" in getattr(a, "__name__", "") else str(a))
               for a in argtypes]
        assert got == want, (name, got, want)
        checked += 1
    assert checked >= 30


def test_param_layout_matches_oracle_tables():
    for (L_, C, Z, xp) in [(16, 10, 2, True), (16, 10, 2, False), (32, 12, 4, True)]:
        cfg = _lib.make_cfg(0, 200, L_, 88, 88, Z, C, xp)
        P, offs, rows, cols = _lib.param_layout(cfg)
        shapes = O.vrnn_param_shapes(L_, 88, 88, Z, C, xp)
        o = 0
        for i, (name, shp) in enumerate(shapes):
            assert offs[i] == o, name
            assert (rows[i], cols[i]) == (shp if len(shp) == 2 else (0, shp[0])), name
            n = 1
            for d in shp:
                n *= d
            o += n
        assert P == o
    cfg = _lib.make_cfg(0, 200, 16, 88, 88, 2, 10, True)
    assert _lib.param_layout(cfg)[0] == 266134
    cfg = _lib.make_cfg(1, 100, 1, 88, 88, 4, 2, True, Hc=88)
    P, offs, rows, cols = _lib.param_layout(cfg)
    assert P == 32922
    shapes = O.vae_param_shapes(88, 88, 4, 88, 2, True)
    assert [(r, c) for r, c in zip(rows, cols)] == [s if len(s) == 2 else (0, s[0]) for _, s in shapes]


def test_argument_validation_without_gpu():
    L = _lib.lib()
    cfg = _lib.make_cfg(0, 200, 16, 88, 64, 2, 10, True)            # H != 88
    assert L.clv_workspace_bytes(ctypes.byref(cfg)) == -2
    cfg = _lib.make_cfg(0, 200, 16, 88, 88, 2, 20, True)            # C > 16
    assert L.clv_workspace_bytes(ctypes.byref(cfg)) == -2
    cfg = _lib.make_cfg(0, 200, 16, 88, 88, 2, 10, True)
    assert L.clv_workspace_bytes(ctypes.byref(cfg)) > 0
    assert L.clv_workspace_offset(ctypes.byref(cfg), b"gates_e") > 0
    assert L.clv_workspace_offset(ctypes.byref(cfg), b"nope") == -1
    assert L.clv_train_step(ctypes.byref(cfg), None, None, None, None, None, None, None, None, None,
                            None, 0, None) == -1
    assert L.clv_gemm(None, None) == -1
