"""CPU: the C-ABI library builds, loads, and exports every symbol include/gnan_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gnan_build", os.path.join(ROOT, "graph-neural-additive-networks---gnan_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def header_functions():
    src = open(os.path.join(ROOT, "include", "gnan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gnan_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in gnan_b200.h but not exported"


def test_python_binding_covers_header(lib_path):
    from gnan_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_functions()
    lib = _lib.load()
    assert lib.gnan_version() == 100
    assert lib.gnan_last_error() is not None


def test_argument_validation_without_gpu(lib_path):
    """Validation happens before any CUDA call, so error codes can be checked on a CPU-only box."""
    from gnan_b200 import _lib
    lib = _lib.load()
    p = _lib.MlpParams(3, 24, 2, 3, None, None, None, None, None, None)   # H=24 unsupported, wo NULL
    rc = lib.gnan_mlp_fwd(None, 4, 3, ctypes.byref(p), 0.0, 0, None, 0, None, None, 0, None)
    assert rc == 1 and b"wo" in lib.gnan_last_error()
    rc = lib.gnan_aggregate_rows_fwd(None, 4, 4, 16, None, 0, 5, 1, None, None, 1, None, None)
    assert rc == 1
    assert lib.gnan_mlp_workspace_bytes(0, ctypes.byref(p), 0, 0) == 0
