"""CPU-only checks of the drop-in boundary: the header is plain C, the library loads, exports every symbol that
include/dnz_gpu.h declares, and fails loudly (no CPU fallback) when no CUDA device is present."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dnz_gpu.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dnz_[a-z_0-9]+)\s*\(", src)))


def test_header_is_plain_c():
    subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", HEADER])


def test_library_exports_every_declared_symbol():
    import denormalized_b200 as d
    L = d.lib()
    names = declared_functions()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert set(d.capi.EXPORTS) <= set(names)
    # no torch / C++ types leak through the boundary: every exported dnz_ symbol is an unmangled C name
    out = subprocess.check_output(["nm", "-D", "--defined-only", d.library_path()], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported


def test_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import denormalized_b200 as d
    with pytest.raises(d.DnzError) as e:
        d.GpuStreamingWindow(d.canonical_schema(), "sensor_name", [("count", "reading", "count")], 1000)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)
    assert d.lib().dnz_device_count() == 0


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "denormalized_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cc")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "dnz_oracle" not in text, f
