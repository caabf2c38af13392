"""The C-ABI library loads without a GPU and exports exactly what include/hardnet_b200.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (REPO / "include" / "hardnet_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hn_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from hardnetnas_b200 import _lib
    if not _lib.LIB_PATH.exists():
        from hardnetnas_b200 import build
        build.build()
    return _lib.load()


def test_header_declares_symbols():
    syms = _declared_symbols()
    assert "hn_forward" in syms and "hn_create" in syms and "hn_last_error" in syms


def test_library_exports_every_declared_symbol(lib):
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/hardnet_b200.h but not exported"


def test_python_signatures_cover_header(lib):
    from hardnetnas_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_version_and_error_string(lib):
    assert lib.hn_version() >= 100
    assert isinstance(lib.hn_last_error(), bytes)


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    st = lib.hn_create(ctypes.byref(h), 0, 0)
    assert st != 0 and not h.value
    assert len(lib.hn_last_error()) > 0


def test_product_package_does_not_import_oracle():
    for py in (REPO / "hardnetnas_b200").rglob("*.py"):
        src = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{py} imports the oracle"
        assert "/root/reference" not in src, f"{py} reads the reference tree"
