"""CPU-only: the C-ABI library builds for sm_100a, loads, and exports every symbol include/pdeop.h declares
(no compute calls without a GPU); the product refuses to run without CUDA."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    from mech_nn_discovery_pde_b200.build import build
    path = build()
    dll = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "pdeop.h")).read()
    declared = sorted(set(re.findall(r"\b(pdeop_[a-z_]+)\s*\(", hdr)))
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(dll, name), f"{name} declared in pdeop.h but not exported"
    dll.pdeop_backend_name.restype = ctypes.c_char_p
    assert dll.pdeop_backend_name().decode() == "cuda-sm100a"


def test_sass_is_sm100a():
    from mech_nn_discovery_pde_b200.build import build
    import subprocess
    path = build()
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback():
    from mech_nn_discovery_pde_b200 import MultigridLayer, _lib
    from oracle.cases import IV_LISTS
    with pytest.raises(_lib.PdeopError):
        MultigridLayer(bs=1, coord_dims=(16, 16), order=2, n_ind_dim=1, n_iv=1, n_grid=2,
                       init_index_mi_list=IV_LISTS["burgers"], n_iv_steps=1)
