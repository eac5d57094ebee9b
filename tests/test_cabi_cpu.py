"""CPU-side checks of the boundary: the C-ABI library builds, loads, and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    import __graft_entry__
    __graft_entry__.build()
    from gloria_nlp_project_b200.build import LIB_PATH
    assert os.path.exists(LIB_PATH)
    return LIB_PATH


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gloria_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gloria_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(libpath):
    h = ctypes.CDLL(libpath)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/gloria_b200.h but not exported"


def test_ctypes_prototypes_cover_header(libpath):
    from gloria_nlp_project_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == declared_symbols()
    assert _lib.lib().gloria_b200_version() >= 100


def test_argument_validation_without_gpu(libpath):
    """Entry points validate arguments before touching the device: bad calls return an error code + message."""
    from gloria_nlp_project_b200 import _lib
    L = _lib.lib()
    rc = L.gloria_b200_local_sim_fwd_f32(None, None, None, 1, 1, 1, 1, 1, 1, 0, 4.0, 5.0, 0, 1e-8, None, None, None,
                                         None, 0, None)
    assert rc == 1 and b"null" in L.gloria_b200_last_error()
    rc = L.gloria_b200_ce_bidir_fwd(None, 4, 1.0, None, None, None, None)
    assert rc == 1
    assert L.gloria_b200_local_f32_workspace(48, 48, 768, 361, 97, 97, 0) > 0
    with pytest.raises(RuntimeError):
        _lib.check(rc, "ce_bidir_fwd")


def test_no_cpu_fallback():
    import torch
    from gloria_nlp_project_b200 import gloria_loss
    with pytest.raises(RuntimeError, match="CUDA"):
        gloria_loss.local_loss(torch.randn(2, 8, 2, 2), torch.randn(2, 8, 4), [3, 2])
    with pytest.raises(RuntimeError, match="CUDA"):
        gloria_loss.global_loss(torch.randn(2, 8), torch.randn(2, 8))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "gloria_nlp_project_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
