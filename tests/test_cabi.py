"""The C-ABI library builds, loads, and exports every symbol include/dreamer_b200.h declares (CPU only:
no compute entry point is called)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from dreamer_b200 import _lib
    return _lib


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dreamer_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(drm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    syms = declared_symbols()
    assert len(syms) >= 20
    handle = lib.load()
    for s in syms:
        assert hasattr(handle, s), f"{s} is declared in the header but not exported"
        assert s in lib.SIGNATURES, f"{s} has no ctypes signature in dreamer_b200/_lib.py"
    assert set(lib.SIGNATURES) == set(syms)


def test_abi_version_and_error_string(lib):
    h = lib.load()
    assert h.drm_abi_version() == 1
    assert isinstance(h.drm_last_error(), bytes)
    assert h.drm_launch_count() >= 0


def test_library_is_sm100a_tcgen05_tma():
    """SASS evidence that the shipped binary is the Blackwell-native path (B200_PROFILING.md table)."""
    so = os.path.join(ROOT, "dreamer_b200", "libdreamer_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True, check=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass  # no legacy mma.sync path


def test_no_cpu_fallback_without_cuda(lib):
    """Product wrappers refuse CPU tensors instead of silently computing elsewhere."""
    import torch
    from dreamer_b200 import ops
    with pytest.raises(RuntimeError):
        ops.categorical32(torch.zeros(4, 32), torch.zeros(4))
    with pytest.raises(RuntimeError):
        ops.lambda_return(torch.zeros(2, 3), torch.zeros(2, 3), torch.zeros(2, 4), 0.99, 0.95)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dreamer_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
