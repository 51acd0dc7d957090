"""Loads the test-only CPU emulator of the pdeop backend (tests/emu/pdeop_host.cpp).  TEST INFRASTRUCTURE."""
import ctypes

from mech_nn_discovery_pde_b200._lib import PdeopLibrary
from tests.emu.build import build

_EMU = None


def emu_library():
    global _EMU
    if _EMU is None:
        _EMU = PdeopLibrary(build())
        assert _EMU.backend == "host-emulator"
        _EMU.dll.pdeop_emu_set_gs_reverse.argtypes = [ctypes.c_int]
    return _EMU
