"""Builds the test-only CPU emulator tests/emu/libpdeop_emu.so (see pdeop_host.cpp).  TEST INFRASTRUCTURE."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = [os.path.join(HERE, "pdeop_host.cpp"), os.path.join(ROOT, "mech_nn_discovery_pde_b200", "csrc", "pdeop_solver.cpp")]
OUT = os.path.join(HERE, "libpdeop_emu.so")


def build(force=False):
    deps = SRC + [os.path.join(ROOT, "mech_nn_discovery_pde_b200", "csrc", f)
                  for f in ("pdeop_elem.h", "pdeop_common.h", "pdeop_backend.h", "pdeop_lstsq.h", "pdeop_gs_line.h")] + \
        [os.path.join(ROOT, "include", "pdeop.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    cmd = ["g++", "-O2", "-mfma", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-o", OUT] + SRC
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
