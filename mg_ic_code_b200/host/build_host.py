"""Builds the C++ host mirror + driver (g++, links only against the C ABI library libmgic_b200.so)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
EXE = os.path.join(PKG, "lib", "poisson_solver_b200")
SRC = ["VariableCoeffPoissonOperator.cpp", "VariableCoeffPoissonOperatorFactory.cpp", "poisson_solver_b200.cpp"]


def stale():
    if not os.path.exists(EXE):
        return True
    t = os.path.getmtime(EXE)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".H", ".cpp", ".py"))]
    deps += [os.path.join(ROOT, "include", "mgic.h"), os.path.join(PKG, "lib", "libmgic_b200.so")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False):
    if not force and not stale():
        return EXE
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + HERE] + \
          [os.path.join(HERE, s) for s in SRC] + ["-L" + os.path.join(PKG, "lib"), "-lmgic_b200", "-Wl,-rpath,$ORIGIN", "-o", EXE]
    subprocess.check_call(cmd)
    return EXE


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
