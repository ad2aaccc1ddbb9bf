"""The C-ABI shared library loads on a CPU-only host and exports every symbol include/cfr_b200.h declares
(no compute calls here: those need a GPU)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "cfr_b200.h")).read()
    return sorted(set(re.findall(r"CFR_API\s+[\w\s\*]+?\b(cfr_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from certifyingfacerecognition_b200 import _lib, build
    build.build()
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cfr_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_version_and_error_string_without_gpu():
    from certifyingfacerecognition_b200 import _lib
    lib = _lib.load()
    assert lib.cfr_version() == 100
    assert isinstance(lib.cfr_last_error(), bytes)
    assert lib.cfr_launch_count() == 0 or lib.cfr_launch_count() > 0


def test_conv_desc_layout_matches_header():
    """ctypes mirror of cfr_conv_desc must have the C struct's size (checked against a tiny C program's sizeof
    would need a compiler at test time; here: field order / count sanity + natural alignment)."""
    from certifyingfacerecognition_b200._lib import ConvDesc
    names = [f[0] for f in ConvDesc._fields_]
    assert names[:5] == ["inp", "N", "Hin", "Win", "Cin"]
    assert names[-2:] == ["stat_sum", "stat_sq"]
    assert ctypes.sizeof(ConvDesc) % 8 == 0


def test_no_cpu_fallback():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from certifyingfacerecognition_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine({}, {}, torch.zeros(5, 512), torch.zeros(4, 512))
