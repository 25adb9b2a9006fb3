"""libb200unet.so loads on a CPU box and exports every symbol include/b200unet.h declares."""
import importlib
import os
import re

import pytest

from conftest import PKG, ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "b200unet.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2u_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = importlib.import_module(PKG + "._lib")
    if not os.path.exists(lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    l = lib.lib()
    names = _declared()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(l, n)]
    assert not missing, missing
    assert sorted(lib.SYMBOLS) == names
    assert l.b2u_version() == 100


def test_errors_are_reported_not_thrown():
    lib = importlib.import_module(PKG + "._lib")
    l = lib.lib()
    # argument validation happens before any CUDA call, so it is testable without a GPU
    rc = l.b2u_head_fwd(0, None, 8, 7, None, None, None, 10, None)
    assert rc == -1 and b"cin" in l.b2u_last_error()
    with pytest.raises(lib.B2UError):
        lib.check(rc, "head_fwd")


def test_engine_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    E = importlib.import_module(PKG + ".engine")
    G = importlib.import_module(PKG + ".graphs")
    lib = importlib.import_module(PKG + "._lib")
    with pytest.raises(lib.B2UError):
        E.Engine(G.unet(32, 1))
