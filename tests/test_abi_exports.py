"""The C-ABI library loads on a CPU-only box and exports every symbol include/otters_b200.h declares."""
import ctypes
import os
import re

import pytest

from helpers import ROOT, ob


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "otters_b200.h")).read()
    return sorted(set(re.findall(r"OTTERS_API[^;]*?\b(otters_\w+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = ctypes.CDLL(os.path.join(ROOT, "otters_b200", "libotters_b200.so"))
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/otters_b200.h but not exported"


def test_python_binding_covers_header():
    from otters_b200 import _ffi

    assert sorted(_ffi.BOUND_SYMBOLS) == declared_symbols()


def test_version_and_error_string():
    from otters_b200 import _ffi

    assert b"sm_100a" in _ffi.otters_version()
    assert isinstance(_ffi.last_error(), str)


def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device the product must fail loudly instead of computing on the CPU."""
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(ob.OttersError) as ei:
        ob.Context(0)
    assert "no CPU fallback" in str(ei.value)
    store = ob.VecStore(2)
    store.add_vector([1.0, 0.0])
    with pytest.raises(ob.OttersError):
        store.query([1.0, 0.0], ob.Metric.Cosine).take(1).collect()


def test_product_does_not_import_oracle():
    """Nothing under otters_b200/ may reference oracle/ (the checker is never on the product path)."""
    pkg = os.path.join(ROOT, "otters_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower().replace("no cpu fallback", ""), f"{f} mentions the oracle"
