#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY: compile oracle/rebvo_oracle.cpp (the CPU restatement) into oracle/_build/liboracle_port.so
with the reference's floating-point regime (x86-64, -O2, no FMA contraction, no -march)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rebvo_oracle.cpp")
OUT = os.path.join(HERE, "_build", "liboracle_port.so")


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) > os.path.getmtime(SRC):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-std=c++11", "-O2", "-m64", "-fPIC", "-shared", "-ffp-contract=off", "-o", OUT, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle port build failed:\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
