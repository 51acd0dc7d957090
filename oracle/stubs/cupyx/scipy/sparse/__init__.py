"""Test-infrastructure stub: cupyx.scipy.sparse -> scipy.sparse (reference solver/multigrid.py:237-239)."""
from scipy.sparse import *  # noqa
from scipy.sparse import coo_matrix, tril, triu  # noqa
