"""mech_nn_discovery_pde_b200 -- B200-native (sm_100a) differentiable PDE-layer solve.

Drop-in for the PDE-layer path of alpz/mech-nn-discovery-pde: ``MultigridLayer``, ``PDEDenseLayer``
and ``PDEConfig`` keep the reference's names, constructor arguments, tensor shapes and dtypes; the
solve itself runs in hand-written CUDA kernels behind the C ABI of include/pdeop.h.  There is no CPU
path: constructing a layer without the CUDA library or without a GPU raises.
"""
from .config import PDEConfig
from .solver.multigrid import MultigridLayer, MultigridSolver
from .solver.pde_layer_dense import PDEDenseLayer

__all__ = ["PDEConfig", "MultigridLayer", "MultigridSolver", "PDEDenseLayer"]
