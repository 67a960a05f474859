"""The C-ABI library loads on a machine without a GPU, exports every symbol include/umv.h declares, and
refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from unimedvl_b200 import build, _lib
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from unimedvl_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "umv.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(umv_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 29
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/umv.h but not exported by libumv.so"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.umv_abi_version() == 1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from unimedvl_b200 import _lib, config
    from unimedvl_b200.engine import Engine
    d = _lib.Dims(hidden=896, heads=7, kv_heads=1, inter=1536, layers=1, vocab=2048, max_tokens=64, max_seqs=2, kv_pages=4)
    h = C.c_void_p()
    assert lib.umv_create(C.byref(d), C.byref(h)) == _lib.ERR_CUDA
    assert b"no CPU fallback" in lib.umv_last_error()
    with pytest.raises(_lib.UmvError):
        Engine(config.tiny())


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "unimedvl_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_status_to_exception_mapping():
    from unimedvl_b200 import _lib
    assert _lib._EXC[_lib.ERR_INVALID] is ValueError
    assert _lib._EXC[_lib.ERR_UNSUPPORTED] is NotImplementedError
    assert _lib._EXC[_lib.ERR_STATE] is AssertionError
