"""The C-ABI library loads and exports every symbol include/*.h declares (no GPU needed)."""
import ctypes
import os
import re

from conftest import ROOT
from latticednaorigami_b200 import binding


def declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ldo_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(binding.LIB_PATH), "build the CUDA library first (__graft_entry__.build())"
    lib = ctypes.CDLL(binding.LIB_PATH)
    names = declared_symbols("ldo_b200.h") + declared_symbols("ldo_host.h")
    assert len(names) > 50
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"


def test_binding_lists_match_headers():
    assert sorted(binding.ENGINE_SYMBOLS) == declared_symbols("ldo_b200.h")
    assert sorted(binding.HOST_SYMBOLS) == declared_symbols("ldo_host.h")


def test_missing_library_fails_loudly(tmp_path):
    import pytest
    with pytest.raises(binding.LdoError, match="no CPU fallback"):
        binding.load(str(tmp_path / "libldo_b200.so"))


def test_product_does_not_reference_oracle():
    """The shipped package and C sources never import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "latticednaorigami_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_ref" not in text and "liboracle" not in text and "hostsim.so" not in text, f


def test_loader_refuses_a_non_cuda_build(hostsim_lib):
    """The package's loader accepts CUDA builds only: the host emulation used by this test-suite exports the
    whole ABI but carries another build marker and cannot be reached through binding.load / Simulation(lib_path=)."""
    import pytest

    from conftest import HOSTSIM_LIB
    with pytest.raises(binding.LdoError, match="not a CUDA build"):
        binding.load(HOSTSIM_LIB)
    with pytest.raises(binding.LdoError, match="not a CUDA build"):
        binding.Simulation("/nonexistent.inp", 1, 0, lib_path=HOSTSIM_LIB)
    assert binding.load().ldo_build_info().decode() == "cuda sm_100a"
