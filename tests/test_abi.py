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


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md maps every exported entry point to the reference interface it replaces; a symbol added to the header
    without a row there fails here.  Rows abbreviate families as `cfr_program_add_blur_act_stats`, `_finalize_stats`, ... or
    `cfr_matcher_create / _run / _destroy`, so a name counts as documented when the text after its family prefix appears."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = []
    for n in _declared_symbols():
        if n in text:
            continue
        stem = next((pre for pre in ("cfr_program_add", "cfr_program", "cfr_matcher", "cfr_sampler", "cfr_profile")
                     if n.startswith(pre + "_")), None)
        if stem is None or stem not in text or not re.search(r"[`/ ]" + re.escape(n[len(stem):]) + r"\b", text):
            missing.append(n)
    assert not missing, missing


def test_version_and_error_string_without_gpu():
    from certifyingfacerecognition_b200 import _lib
    lib = _lib.load()
    assert lib.cfr_version() == 100
    assert isinstance(lib.cfr_last_error(), bytes)
    assert lib.cfr_launch_count() == 0 or lib.cfr_launch_count() > 0


def test_conv_desc_layout_matches_header(tmp_path):
    """ctypes mirrors of cfr_conv_desc / cfr_sampler_desc have the C structs' size and field offsets (a tiny C program
    compiled against include/cfr_b200.h prints them)."""
    import subprocess
    from certifyingfacerecognition_b200._lib import ConvDesc, SamplerDesc
    names = [f[0] for f in ConvDesc._fields_]
    assert names[:5] == ["inp", "N", "Hin", "Win", "Cin"]
    assert names[-5:] == ["stat_sum", "stat_sq", "kSplit", "keepMap", "keepDim"]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "cfr_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(cfr_conv_desc), '
                   'offsetof(cfr_conv_desc, out), offsetof(cfr_conv_desc, stat_sum), offsetof(cfr_conv_desc, kSplit), '
                   'sizeof(cfr_sampler_desc), offsetof(cfr_sampler_desc, tail), offsetof(cfr_conv_desc, keepMap), '
                   'offsetof(cfr_sampler_desc, img_chunk_bytes)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert got == [ctypes.sizeof(ConvDesc), ConvDesc.out.offset, ConvDesc.stat_sum.offset, ConvDesc.kSplit.offset,
                   ctypes.sizeof(SamplerDesc), SamplerDesc.tail.offset, ConvDesc.keepMap.offset,
                   SamplerDesc.img_chunk_bytes.offset]


def test_no_cpu_fallback():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from certifyingfacerecognition_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine({}, {}, torch.zeros(5, 512), torch.zeros(4, 512))
