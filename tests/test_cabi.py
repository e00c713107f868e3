"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/atspeed.h declares, argument validation returns error codes (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

from _common import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "atspeed.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(atspeed_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from atspeed_b200 import _lib, build
    build.build_library()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in atspeed.h but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes signature in _lib.SYMBOLS"
    assert sorted(_lib.SYMBOLS) == declared
    assert lib.atspeed_abi_version() == _lib.ABI_VERSION


def test_argument_errors_are_codes_not_crashes():
    from atspeed_b200 import _lib
    lib = _lib.load()
    cfg = _lib.Config(K=64, N=40, max_new_tokens=4, max_prompt=128, num_sms=148)   # K above the limit
    md = _lib.ModelDesc(vocab=100, hidden=64, n_layers=1, n_heads=4, head_dim=16, mlp=128, rms_eps=1e-6)
    n = C.c_size_t(0)
    rc = lib.atspeed_session_workspace_bytes(C.byref(md), None, C.byref(cfg), C.byref(n))
    assert rc < 0 and b"K=64" in lib.atspeed_last_error()
    cfg.K = 10
    assert lib.atspeed_session_workspace_bytes(C.byref(md), None, C.byref(cfg), C.byref(n)) == 0 and n.value > 0
    md.head_dim = 24
    assert lib.atspeed_session_workspace_bytes(C.byref(md), None, C.byref(cfg), C.byref(n)) < 0
    assert lib.atspeed_session_destroy(None) == 0
    with pytest.raises(_lib.AtSpeedError):
        _lib.check(-1)


def test_product_path_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under atspeed_b200/ may import it."""
    pkg = os.path.join(ROOT, "atspeed_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
            assert not re.search(r"""["']/root/reference""", src), fn   # cited in docstrings, never read
