"""ctypes binding of the C ABI in include/pdeop.h.

The product loads exactly one library: ``csrc/libpdeop.so`` (hand-written CUDA for sm_100a, built by
``build.py`` / ``__graft_entry__.build()``).  There is no CPU fallback: if the library is missing, was
not built for CUDA, or no CUDA device is present, ``get_library()`` raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpdeop.so")

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int

# enum pdeop_query
Q_NLEVELS, Q_G, Q_M, Q_N_EQ, Q_N_INIT, Q_NTOT, Q_FTOT, Q_PERSIST_BYTES, Q_SCRATCH_BYTES = range(9)
Q_DIM0 = 16
# enum pdeop_stage
STAGE_APPLY_K, STAGE_GS, STAGE_RESTRICT, STAGE_PROLONG, STAGE_VCYCLE, STAGE_COARSE_SOLVE, STAGE_ATB = range(7)


class PlanOpts(ctypes.Structure):
    """pdeop_plan_opts: creation-time options of a plan."""
    _fields_ = [("evolution", c_int), ("chain", c_int), ("gs_pipe", c_int), ("reserved", c_int * 5)]


class PcgCfg(ctypes.Structure):
    """pdeop_pcg_cfg: converged-mode knobs."""
    _fields_ = [("max_iter", c_int), ("rtol", ctypes.c_double), ("smoother", c_int), ("sweeps", c_int),
                ("jacobi_w", ctypes.c_double), ("power_iters", c_int), ("cheb_ratio", ctypes.c_double)]


class SolverCfg(ctypes.Structure):
    _fields_ = [("gs_pre", c_int), ("gs_post", c_int), ("mg_steps", c_int), ("max_iter", c_int),
                ("restart", c_int), ("atol", ctypes.c_double), ("gs_variant", c_int)]


EXPORTS = [
    "pdeop_plan_create", "pdeop_plan_destroy", "pdeop_plan_query", "pdeop_last_error", "pdeop_backend_name",
    "pdeop_mg_forward", "pdeop_mg_backward", "pdeop_dense_forward", "pdeop_dense_backward", "pdeop_mg_setup",
    "pdeop_stage", "pdeop_fgmres", "pdeop_plan_profile_enable", "pdeop_plan_profile_collect", "pdeop_launch_count",
    "pdeop_plan_set_tuning", "pdeop_plan_get_tuning", "pdeop_plan_create_ex", "pdeop_mg_forward_converged",
    "pdeop_mg_backward_converged",
]

TUNING_KEYS = {"gs_pipe": 0, "chain": 1}

PROFILE_CATEGORIES = ["gs_fine", "gs_coarse", "apply_fine", "apply_coarse", "transfer", "coarse_solve", "factor",
                      "krylov", "setup", "grads", "layout"]


class PdeopError(RuntimeError):
    pass


def _ptr(t):
    if t is None:
        return c_void_p(0)
    assert t.is_contiguous(), "pdeop: tensors crossing the C ABI must be contiguous"
    return c_void_p(t.data_ptr())


class PdeopLibrary:
    """Typed wrapper over one loaded libpdeop*.so."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise PdeopError(f"pdeop: native library not found at {path}; run `python -m "
                             f"mech_nn_discovery_pde_b200.build` (or __graft_entry__.build())")
        self.path = path
        self.dll = ctypes.CDLL(path)
        d = self.dll
        for name in EXPORTS:
            if not hasattr(d, name):
                raise PdeopError(f"pdeop: {path} does not export {name}")
        d.pdeop_last_error.restype = ctypes.c_char_p
        d.pdeop_backend_name.restype = ctypes.c_char_p
        d.pdeop_plan_create.argtypes = [c_int, ctypes.POINTER(c_int), c_int, c_int, c_int, c_int, c_int,
                                        ctypes.POINTER(c_int), ctypes.POINTER(c_void_p)]
        d.pdeop_plan_create_ex.argtypes = [c_int, ctypes.POINTER(c_int), c_int, c_int, c_int, c_int, c_int,
                                           ctypes.POINTER(c_int), ctypes.POINTER(PlanOpts), ctypes.POINTER(c_void_p)]
        d.pdeop_plan_destroy.argtypes = [c_void_p]
        d.pdeop_plan_destroy.restype = None
        d.pdeop_plan_query.argtypes = [c_void_p, c_int, c_int, ctypes.POINTER(ctypes.c_longlong)]
        PP = ctypes.POINTER(c_void_p)
        CFG = ctypes.POINTER(SolverCfg)
        d.pdeop_mg_forward.argtypes = [c_void_p, CFG] + [c_void_p] * 3 + [PP] * 3 + [c_void_p] * 5
        d.pdeop_mg_backward.argtypes = [c_void_p, CFG] + [c_void_p] * 16
        PCG = ctypes.POINTER(PcgCfg)
        d.pdeop_mg_forward_converged.argtypes = [c_void_p, PCG] + [c_void_p] * 3 + [PP] * 3 + [c_void_p] * 5
        d.pdeop_mg_backward_converged.argtypes = [c_void_p, PCG] + [c_void_p] * 16
        d.pdeop_dense_forward.argtypes = [c_void_p] + [c_void_p] * 11
        d.pdeop_dense_backward.argtypes = [c_void_p] + [c_void_p] * 16
        d.pdeop_mg_setup.argtypes = [c_void_p, c_void_p] + [PP] * 3 + [c_void_p] * 4
        d.pdeop_stage.argtypes = [c_void_p, CFG, c_int, c_int, c_int] + [c_void_p] * 6
        d.pdeop_fgmres.argtypes = [c_void_p, CFG, c_int] + [c_void_p] * 7
        d.pdeop_plan_profile_enable.argtypes = [c_void_p, c_int]
        d.pdeop_plan_profile_collect.argtypes = [c_void_p, ctypes.POINTER(ctypes.c_double),
                                                 ctypes.POINTER(ctypes.c_longlong), c_int]
        d.pdeop_plan_set_tuning.argtypes = [c_void_p, c_int, c_int]
        d.pdeop_plan_get_tuning.argtypes = [c_void_p, c_int, ctypes.POINTER(c_int)]
        d.pdeop_launch_count.restype = ctypes.c_longlong
        self.backend = d.pdeop_backend_name().decode()

    def profile_enable(self, plan, on=True):
        self.check(self.dll.pdeop_plan_profile_enable(plan, 1 if on else 0))

    def profile_collect(self, plan):
        """{category: (total_ms, kernel groups)} of one plan since profile_enable; waits for the recorded events."""
        n = len(PROFILE_CATEGORIES)
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_longlong * n)()
        self.dll.pdeop_plan_profile_collect(plan, ms, cnt, n)
        return {PROFILE_CATEGORIES[i]: (float(ms[i]), int(cnt[i])) for i in range(n)}

    def launch_count(self):
        return int(self.dll.pdeop_launch_count())

    def set_tuning(self, plan, key, value):
        """Kernel-variant switch of one plan (pdeop.h: pdeop_plan_set_tuning); key 'gs_pipe'."""
        self.check(self.dll.pdeop_plan_set_tuning(plan, TUNING_KEYS[key], int(value)))

    def get_tuning(self, plan, key):
        out = c_int()
        self.check(self.dll.pdeop_plan_get_tuning(plan, TUNING_KEYS[key], ctypes.byref(out)))
        return int(out.value)

    def check(self, rc):
        if rc != 0:
            raise PdeopError("pdeop: " + self.dll.pdeop_last_error().decode())

    def plan_create(self, dims, order, batch, n_grid, downsample_first, iv_desc, evolution=False, chain=None,
                    gs_pipe=None):
        d = len(dims)
        dims_a = (c_int * d)(*[int(v) for v in dims])
        n_iv = len(iv_desc[0]) if iv_desc else 0
        flat = [int(v) for lvl in iv_desc for spec in lvl for v in spec]
        iv_a = (c_int * max(len(flat), 1))(*flat)
        out = c_void_p()
        opts = PlanOpts()
        opts.evolution = int(bool(evolution))
        opts.chain = -1 if chain is None else int(bool(chain))
        opts.gs_pipe = -1 if gs_pipe is None else int(gs_pipe)
        self.check(self.dll.pdeop_plan_create_ex(d, dims_a, order, batch, n_grid, int(bool(downsample_first)), n_iv,
                                                 iv_a, ctypes.byref(opts), ctypes.byref(out)))
        return out

    def plan_destroy(self, plan):
        self.dll.pdeop_plan_destroy(plan)

    def query(self, plan, what, level=0):
        out = ctypes.c_longlong()
        self.check(self.dll.pdeop_plan_query(plan, what, level, ctypes.byref(out)))
        return int(out.value)


_LIB = None


def get_library():
    """The CUDA library, or an exception.  Never returns anything that computes on the CPU."""
    global _LIB
    if _LIB is None:
        if not torch.cuda.is_available():
            raise PdeopError("pdeop: no CUDA device available; this package has no CPU path")
        lib = PdeopLibrary(LIB_PATH)
        if not lib.backend.startswith("cuda"):
            raise PdeopError(f"pdeop: {LIB_PATH} reports backend '{lib.backend}', expected the CUDA build")
        _LIB = lib
    return _LIB


def current_stream_ptr(device):
    if device.type == "cuda":
        return c_void_p(torch.cuda.current_stream(device).cuda_stream)
    return c_void_p(0)
