"""Build the plain-C oracle (oracle/rtvq_ref.c -> oracle/liboracle_c.so).  TEST INFRASTRUCTURE.

The reference is pure Python (no C/C++ sources under /root/reference), so there is
nothing to compile into ``oracle/_ref/``; the Python reference is instead imported in
the build container by ``tests/golden/make_golden.py`` to produce the golden fixtures.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rtvq_ref.c")
OUT = os.path.join(HERE, "liboracle_c.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
           "-Wall", "-Wextra", "-o", OUT, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
