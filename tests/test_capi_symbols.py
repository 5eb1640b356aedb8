"""The C-ABI library loads and exports every symbol include/fastsmc_b200.h declares (no compute calls: CPU box)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "fastsmc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fsmc_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from fastsmc_b200 import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    names = _declared()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fastsmc_b200.h but not exported"
    assert set(_native.EXPORTS) == set(names)


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point fails with an error code; nothing silently runs on the CPU."""
    from fastsmc_b200 import _native
    if _native.lib().fsmc_device_count() > 0:
        return
    try:
        _native.Context(0)
    except _native.FastSMCError as e:
        assert e.code in (-1, -2)
    else:
        raise AssertionError("Context creation must fail without a GPU")


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may touch oracle/."""
    for base, _, files in os.walk(os.path.join(ROOT, "fastsmc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("oracle/", "oracle/") or "pyoracle" not in text, f
                assert "libfastsmc_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f
