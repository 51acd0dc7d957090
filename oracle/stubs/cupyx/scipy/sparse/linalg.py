"""Test-infrastructure stub: cupyx.scipy.sparse.linalg -> scipy (reference solver/multigrid.py:404)."""
from scipy.sparse.linalg import *  # noqa
from scipy.sparse.linalg import spsolve_triangular  # noqa
