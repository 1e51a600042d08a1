"""Compiles oracle/prophet_oracle.c (the plain-C restatement used as checker and CPU baseline)
into oracle/_build/libprophet_oracle.so with gcc.  Test infrastructure only."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libprophet_oracle.so")
SRC = os.path.join(HERE, "prophet_oracle.c")


def build(force: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-fno-fast-math", "-o", LIB, SRC, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("gcc failed building the C oracle")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
