"""Test-infrastructure stub: maps the handful of CuPy calls the reference's multigrid
makes (solver/multigrid.py:233-235,361,389,396,449,474,492) onto numpy so the reference
can run on CPU in the build container.  Not part of the product."""
import numpy as _np
ndarray = _np.ndarray
def asarray(x, *a, **k):
    try:
        import torch
        if isinstance(x, torch.Tensor):
            return x.detach().cpu().numpy()
    except ImportError:
        pass
    return _np.asarray(x, *a, **k)
zeros_like = _np.zeros_like
