"""Pins the CPU oracle (oracle/pde_oracle.py) against vectors produced by the unmodified reference
(oracle/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import pde_oracle as O
from oracle.cases import IV_LISTS
from tests.helpers import golden_loss_w, rel_sampled, seeded_stage_vectors

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


STRUCTS = sorted(glob.glob(os.path.join(GOLDEN, "struct_*.npz")))
LAYERS = sorted(glob.glob(os.path.join(GOLDEN, "layer_*.npz")))
STAGES = sorted(glob.glob(os.path.join(GOLDEN, "stages_*.npz")))


@pytest.mark.parametrize("path", STRUCTS, ids=[os.path.basename(p) for p in STRUCTS])
def test_structure_and_values(path):
    z = np.load(path)
    dims = tuple(int(v) for v in z["dims"])
    st = O.build_structure(dims, IV_LISTS[str(z["iv_name"])])
    assert st.n_eq == int(z["n_eq"]) and st.n_init == int(z["n_init"]) and st.n_deriv == int(z["n_deriv"])
    # bit-exact index work
    assert np.array_equal(np.repeat(np.arange(st.n_eq), st.M), z["eq_row"])
    assert np.array_equal((st.eq_g[:, None] * st.M + np.arange(st.M)).reshape(-1), z["eq_col"])
    assert np.array_equal(st.init_var, z["init_col"])
    assert np.array_equal(st.d_row, z["d_row"])
    assert np.array_equal(st.d_col, z["d_col"])
    # per-call values with non-uniform steps
    off = 0
    steps = []
    for n in dims:
        steps.append(torch.tensor(z["steps"][:, off:off + n - 1]))
        off += n - 1
    dv = O.derivative_values(st, steps).numpy()
    assert dv.shape == z["dvals"].shape
    assert rel(dv, z["dvals"]) < 1e-13
    # known answer (lp_pde_central_diff.py:929-937,981-984): uniform-step filled values == hard-coded constants
    h = float(z["uniform_step"])
    dvu = O.derivative_values(st, [torch.full((2, n - 1), h, dtype=torch.float64) for n in dims]).numpy()
    assert rel(dvu, z["dvals_uniform"]) < 1e-13
    hs = float(z["static_step"])
    dvs = O.derivative_values(st, [torch.full((1, n - 1), hs, dtype=torch.float64) for n in dims]).numpy()[0]
    assert np.abs(dvs - z["d_static_val"]).max() < 1e-9


def _load_layer(path):
    z = np.load(path)
    dims = tuple(int(v) for v in z["dims"])
    steps = [z[f"steps{c}"] for c in range(len(dims))]
    return z, dims, steps


# fixtures whose oracle run takes minutes on the build container's CPU: PDEOP_SLOW=1 enables them
SLOW = {"layer_mg_3d_16x32x32_g3_nodsf.npz"}


@pytest.mark.parametrize("path", LAYERS, ids=[os.path.basename(p) for p in LAYERS])
def test_layer(path):
    if os.path.basename(path) in SLOW and not os.environ.get("PDEOP_SLOW"):
        pytest.skip("minutes of CPU; set PDEOP_SLOW=1 (result recorded in profiles/README.md)")
    z, dims, steps = _load_layer(path)
    iv = IV_LISTS[str(z["iv_name"])]
    B = int(z["bs"])
    G = int(np.prod(dims))
    M = 1 + 2 * len(dims)
    g_out = golden_loss_w(z, dims).reshape(B, G * M)
    if str(z["kind"]) == "dense":
        res = O.dense_layer(dims, iv, z["coeffs"], z["rhs"], z["iv_rhs"], steps, grad_out=g_out)
        tol_x, tol_g = 1e-8, 2e-7   # cond(AtA)~1e10: two LAPACK-based direct solves agree to ~1e-9
        if G * M >= 5000:
            tol_x, tol_g = 5e-8, 1e-6   # 32x32 sine layer (n = 5120): conditioning grows with the grid
    else:
        res = O.mg_layer(dims, iv, z["coeffs"], z["rhs"], z["iv_rhs"], steps, int(z["n_grid"]), bool(z["dsf"]),
                         grad_out=g_out)
        info = z["info"]
        assert res.info_fwd[0] == int(info[0, 0]) and res.info_bwd[0] == int(info[1, 0])
        assert abs(res.info_fwd[1] - info[0, 1]) <= 1e-6 * info[0, 1]
        assert abs(res.info_bwd[1] - info[1, 1]) <= 1e-6 * info[1, 1]
        tol_x, tol_g = 1e-8, 1e-8
    assert rel_sampled(res.x, z, "u") < tol_x
    assert rel_sampled(res.d_coeffs, z, "d_coeffs") < tol_g
    assert rel_sampled(res.d_rhs, z, "d_rhs") < tol_g
    if "d_rhs_fp32quirk" in z.files:
        assert rel(res.d_rhs, z["d_rhs_fp32quirk"]) < 1e-6      # shipped add_pad buffer is fp32 (lp_...:1634)
    assert rel(res.d_iv_rhs, z["d_iv_rhs"]) < tol_g
    for c in range(len(dims)):
        assert rel(res.d_steps[c], z[f"d_steps{c}"]) < max(tol_g, 1e-7)


STAGES_SEEDED = sorted(glob.glob(os.path.join(GOLDEN, "stagesS_*.npz")))


@pytest.mark.parametrize("path", STAGES_SEEDED, ids=[os.path.basename(p) for p in STAGES_SEEDED])
def test_mg_stages_seeded(path):
    """Three-level 3-D V-cycle, 5 Gauss-Seidel sweeps and the normal matvec of the reference's own GL default grid
    (inputs regenerated from the seed, outputs sampled; oracle/make_golden_r2.py:golden_stages_seeded)."""
    s = np.load(path)
    z, dims, steps = _load_layer(path.replace("stagesS_", "layer_"))
    iv = IV_LISTS[str(z["iv_name"])]
    mg = O.mg_setup(dims, iv, z["coeffs"], z["rhs"], z["iv_rhs"], steps, int(z["n_grid"]), bool(z["dsf"]))
    v, x0 = seeded_stage_vectors(s)
    assert rel_sampled(mg.Atb0, s, "Atb") < 1e-13
    assert rel_sampled(mg.K_list[0] @ v, s, "Kv") < 1e-13
    assert rel_sampled(O.smooth_gs(mg.L_list[0], mg.U_list[0], v, x0, 5), s, "gs5") < 1e-12
    assert rel_sampled(O.v_cycle_start(mg, v), s, "vcycle") < 1e-9


@pytest.mark.parametrize("path", STAGES, ids=[os.path.basename(p) for p in STAGES])
def test_mg_stages(path):
    s = np.load(path)
    z, dims, steps = _load_layer(path.replace("stages_", "layer_"))
    iv = IV_LISTS[str(z["iv_name"])]
    mg = O.mg_setup(dims, iv, z["coeffs"], z["rhs"], z["iv_rhs"], steps, int(z["n_grid"]), bool(z["dsf"]))
    assert rel(mg.Atb0, s["Atb"]) < 1e-13
    assert rel(mg.K_list[0] @ s["v"], s["Kv"]) < 1e-13
    assert rel(O.smooth_gs(mg.L_list[0], mg.U_list[0], s["v"], s["x0"], 1), s["gs1"]) < 1e-12
    assert rel(O.smooth_gs(mg.L_list[0], mg.U_list[0], s["v"], s["x0"], 3), s["gs3"]) < 1e-12
    assert rel(O.restrict(mg, 0, s["v"]), s["restrict"]) < 1e-13
    assert rel(O.prolong(mg, 1, s["vc"]), s["prolong"]) < 1e-13
    assert rel(O.v_cycle_start(mg, s["v"]), s["vcycle"]) < 1e-9
    if "v1" in s.files:
        assert rel(mg.K_list[1] @ s["v1"], s["K1v1"]) < 1e-13
    Kc = mg.K_list[-1]
    assert rel(np.einsum("bij,bj->bi", Kc, s["vcs"]), s["Kc_vcs"]) < 1e-13
