"""Builds the plain-C restatement oracle/unet_oracle.c into oracle/_build/liboracle.so (gcc).

TEST INFRASTRUCTURE ONLY.  The reference itself is pure Python on top of TensorFlow: there is
no C / C++ source under /root/reference to compile into oracle/_ref, and TensorFlow is not
installed or installable here, so `kind` of the CPU baseline is "port" (see DESIGN.md).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "unet_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-std=c99", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
