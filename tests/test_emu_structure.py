"""CPU-only checks of the host-side logic (plan index tables, axis tables of K, wavefront Gauss-Seidel
ordering, V-cycle / FGMRES orchestration, gradient formulas, Python layers) by running the shared
per-element bodies through the test-only host emulator and comparing with the reference's golden vectors.
The CUDA kernels call the same bodies; their parity proper is tested on the GPU (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

from mech_nn_discovery_pde_b200 import _lib
from oracle.cases import IV_LISTS
from tests.emu.emu_lib import emu_library
from tests.helpers import GOLDEN, StageRunner, load_layer_case, rel, run_layer_case

import os

STAGE_CASES = ["mg_2d_16x16_g2", "mg_2d_32x32_g3", "mg_3d_8x16x16_g2_nodsf"]


@pytest.mark.parametrize("name", STAGE_CASES)
def test_stages_vs_reference(name):
    lib = emu_library()
    s = np.load(os.path.join(GOLDEN, f"stages_{name}.npz"))
    z, dims, steps = load_layer_case(name)
    B = int(z["bs"])
    sr = StageRunner(lib, "cpu", dims, IV_LISTS[str(z["iv_name"])], B, int(z["n_grid"]), bool(z["dsf"]), z["coeffs"],
                     steps)
    assert int(sr.info[3]) == 0
    assert rel(sr.stage(_lib.STAGE_ATB, 0, z["rhs"], z["iv_rhs"]), s["Atb"]) < 1e-13
    assert rel(sr.stage(_lib.STAGE_APPLY_K, 0, s["v"]), s["Kv"]) < 1e-13
    assert rel(sr.stage(_lib.STAGE_GS, 0, s["v"], s["x0"], count=1), s["gs1"]) < 1e-12
    gs3 = sr.stage(_lib.STAGE_GS, 0, s["v"], s["x0"], count=3)
    assert rel(gs3, s["gs3"]) < 1e-12
    # pipelined sweeps are order independent inside a step: identical bits either way
    lib.dll.pdeop_emu_set_gs_reverse(1)
    try:
        gs3r = sr.stage(_lib.STAGE_GS, 0, s["v"], s["x0"], count=3)
    finally:
        lib.dll.pdeop_emu_set_gs_reverse(0)
    assert np.array_equal(gs3, gs3r)
    assert rel(sr.stage(_lib.STAGE_RESTRICT, 0, s["v"], out_level=1), s["restrict"]) < 1e-13
    assert rel(sr.stage(_lib.STAGE_PROLONG, 1, s["vc"], out_level=0), s["prolong"]) < 1e-13
    assert rel(sr.stage(_lib.STAGE_VCYCLE, 0, s["v"]), s["vcycle"]) < 1e-9
    if "v1" in s.files:
        assert rel(sr.stage(_lib.STAGE_APPLY_K, 1, s["v1"]), s["K1v1"]) < 1e-13


LAYER_CASES = ["dense_1d_24", "dense_2d_8x10", "dense_2d_12x12_sine_uniform", "mg_2d_16x16_g2",
               "mg_2d_16x32_g2_nodsf"]


@pytest.mark.parametrize("name", LAYER_CASES)
def test_layer_vs_reference(name):
    lib = emu_library()
    z, out = run_layer_case(lib, "cpu", name)
    dense = str(z["kind"]) == "dense"
    tol_x, tol_g = (1e-8, 2e-7) if dense else (1e-8, 1e-8)
    assert rel(out["u"], z["u"]) < tol_x
    assert rel(out["d_coeffs"], z["d_coeffs"]) < tol_g
    assert rel(out["d_rhs"], z["d_rhs"]) < tol_g
    assert rel(out["d_iv_rhs"], z["d_iv_rhs"]) < tol_g
    for c in range(len(out["d_steps"])):
        assert rel(out["d_steps"][c], z[f"d_steps{c}"]) < max(tol_g, 1e-7)
    if not dense:
        info = z["info"]
        assert int(out["info_fwd"][0]) == int(info[0, 0]) and int(out["info_bwd"][0]) == int(info[1, 0])
        assert abs(out["info_fwd"][1] - info[0, 1]) <= 1e-6 * info[0, 1]
        assert abs(out["info_bwd"][1] - info[1, 1]) <= 1e-6 * info[1, 1]
