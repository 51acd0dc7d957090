import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from mech_nn_discovery_pde_b200 import _lib
from oracle.cases import IV_LISTS
from tests.helpers import GOLDEN, StageRunner, load_layer_case
lib = _lib.get_library()
for name in ["mg_2d_16x16_g2", "mg_3d_8x16x16_g2_nodsf"]:
    s = np.load(os.path.join(GOLDEN, f"stages_{name}.npz"))
    z, dims, steps = load_layer_case(name)
    sr = StageRunner(lib, "cuda:0", dims, IV_LISTS[str(z["iv_name"])], int(z["bs"]), int(z["n_grid"]), bool(z["dsf"]), z["coeffs"], steps)
    for cnt in (1, 3, 5):
        a = sr.stage(_lib.STAGE_GS, 0, s["v"], s["x0"], count=cnt)
        b = sr.stage(_lib.STAGE_GS, 0, s["v"], s["x0"], count=cnt, gs_variant=1)
        a2 = sr.stage(_lib.STAGE_GS, 0, s["v"], s["x0"], count=cnt)
        d = np.abs(a - b)
        print(name, "LD", os.environ.get("PDEOP_GS_LD"), "count", cnt, "ndiff", int((d > 0).sum()), "of", d.size, "max abs", d.max(),
              "max rel", (d / (np.abs(b) + 1e-300)).max(), "repeatable", np.array_equal(a, a2))
