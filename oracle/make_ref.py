"""Builds oracle/_ref: a runnable copy of the UNMODIFIED reference's solver path  --  TEST INFRASTRUCTURE ONLY.

oracle/_ref/ is git-ignored build output (like the compiled .so files): it is produced in the build container,
where /root/reference exists, and travels to the GPU box with the repository snapshot so that
`bench.py --impl reference` can time the reference's own code on the box's host cores (oracle/run_ref.py).
Nothing is edited: the reference's Python files are copied byte for byte; CuPy / ipdb, which the image does not
have, are satisfied by the stub modules of oracle/stubs (numpy/scipy shims).

    python oracle/make_ref.py            # no-op with a message when /root/reference is absent
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PDEOP_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
# the hot path's modules only: solver/*, config.py, extras/* (logger + run-dir helpers imported by the QP modules)
WANTED = ["solver", "extras", "config.py"]


def build(force=False):
    if not os.path.isdir(os.path.join(REF, "solver")):
        print(f"make_ref: {REF} not present; keeping whatever oracle/_ref holds")
        return None
    stamp = os.path.join(OUT, ".complete")
    if os.path.exists(stamp) and not force:
        return OUT
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    for name in WANTED:
        src = os.path.join(REF, name)
        if os.path.isdir(src):
            shutil.copytree(src, os.path.join(OUT, name), ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "logs"))
        elif os.path.exists(src):
            shutil.copy2(src, os.path.join(OUT, name))
    stubs = os.path.join(HERE, "stubs")
    for name in os.listdir(stubs):
        src = os.path.join(stubs, name)
        if name == "__pycache__":
            continue
        if os.path.isdir(src):
            shutil.copytree(src, os.path.join(OUT, name), ignore=shutil.ignore_patterns("__pycache__"))
        else:
            shutil.copy2(src, os.path.join(OUT, name))
    open(stamp, "w").write("copied from " + REF + "\n")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
