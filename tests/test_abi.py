"""CPU-side checks of the C-ABI boundary: libsqlx.so loads without a GPU and exports exactly what
include/sqlx.h declares; argument validation fails with a message instead of launching."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sqlx.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sqlx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import sqlx
    assert os.path.isfile(sqlx.LIB_PATH), "build libsqlx.so first (make / __graft_entry__.build())"
    handle = ctypes.CDLL(sqlx.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 25
    missing = [s for s in declared if not hasattr(handle, s)]
    assert not missing, missing
    # the ctypes prototype table covers the same set
    assert sorted(sqlx.exported_symbols()) == declared


def test_version_and_error_string():
    import sqlx
    L = sqlx.lib()
    assert L.sqlx_version() >= 100
    assert isinstance(L.sqlx_last_error(), bytes)


def test_argument_validation_without_gpu():
    import sqlx
    L = sqlx.lib()
    # NULL pointers / bad shapes are rejected before any CUDA call
    assert L.sqlx_ssim_fwd(None, None, 1, 3, 8, 8, 3, None, None) == -1
    assert b"NULL" in L.sqlx_last_error()
    assert L.sqlx_sql_pred_fwd(None, None, None, None, None, 1, 24, 8, 8, 64, None, None) == -1
    assert b"embedding dim" in L.sqlx_last_error()
    assert L.sqlx_sql_pred_fwd(None, None, None, None, None, 1, 32, 300, 8, 64, None, None) == -1
    assert b"query_nums" in L.sqlx_last_error()
    assert L.sqlx_sql_workspace_bytes(12, 32, 64, 64, 30720) > 0
    desc = sqlx._lib.PhotoDesc(2, 192, 640, 96, 320, 7, 3, 1, 0.85, 0.15, 1e-5, 1e-7)
    assert L.sqlx_photo_fwd(ctypes.byref(desc), None, None, None, None, None, None, None, None, None, None, None, None, 0,
                            None) == -1
    assert b"S=7" in L.sqlx_last_error()


def test_product_path_has_no_cpu_fallback():
    import torch
    import sqlx
    with pytest.raises(sqlx.SqlxError):
        sqlx.SSIM()(torch.rand(1, 3, 16, 16), torch.rand(1, 3, 16, 16))
    with pytest.raises(sqlx.SqlxError):
        sqlx.FullQueryLayer()(torch.rand(1, 32, 8, 8), torch.rand(1, 4, 32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "sfmnext-impl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(import|from)\s+oracle|oracle\.", txt), os.path.join(dirpath, f)
