"""Compile the C part of the oracle: ``oracle/tdho_oracle.c`` -> ``oracle/libtdho_oracle.so``.

TEST INFRASTRUCTURE ONLY (see ``oracle/qs_oracle.py``).  ``__graft_entry__.build()`` calls this so the
checker exists on the GPU box (the ``.so`` is git-ignored but travels with the snapshot); the tests
call it lazily when the library is missing or stale.  No fast-math: the oracle is deterministic.
"""

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.path.join(HERE, "tdho_oracle.c")
LIB_PATH = os.path.join(HERE, "libtdho_oracle.so")


def build(force=False):
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= os.path.getmtime(SOURCE):
        return LIB_PATH
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("gcc not found: the C oracle cannot be built")
    tmp = LIB_PATH + ".tmp"
    cmd = [gcc, "-O2", "-fopenmp", "-shared", "-fPIC", "-o", tmp, SOURCE, "-lm"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("gcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True))
