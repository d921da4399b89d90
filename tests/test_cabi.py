"""CPU-only: the C-ABI library builds for sm_100a, loads, exports every symbol include/spectre_b200.h declares,
and refuses to run without a GPU (no CPU fallback). No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from spectre_b200 import build
    return build.build()


def declared_symbols():
    with open(os.path.join(ROOT, "include", "spectre_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(libpath):
    syms = declared_symbols()
    assert len(syms) >= 40
    lib = ctypes.CDLL(libpath)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, "declared in include/spectre_b200.h but not exported: %s" % missing


def test_only_abi_symbols_are_exported(libpath):
    out = subprocess.check_output(["nm", "-D", "--defined-only", libpath], text=True)
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert exported and all(s.startswith("spb_") for s in exported), exported


def test_cubin_is_sm_100a(libpath):
    out = subprocess.run(["cuobjdump", "-lelf", libpath], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu(libpath):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    from spectre_b200 import halo2
    with pytest.raises(halo2.BackendError):
        halo2.Backend([0])


def test_product_never_imports_oracle():
    """The product package must not reference the oracle (test infrastructure)."""
    pkg = os.path.join(ROOT, "spectre_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    text = f.read()
                assert "halo2_oracle" not in text and "from oracle" not in text and "import oracle" not in text, fn
