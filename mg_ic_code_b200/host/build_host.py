"""Builds the C++ host mirror + driver (g++, links only against the C ABI library libmgic_b200.so)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
EXE = os.path.join(PKG, "lib", "poisson_solver_b200")
SRC = ["VariableCoeffPoissonOperator.cpp", "VariableCoeffPoissonOperatorFactory.cpp", "poisson_solver_b200.cpp"]
# the reference's own, UNMODIFIED Main_PoissonSolver.cpp, compiled where it lies against host/dropin (headers with Chombo's and
# the reference's names that resolve to this host layer): main() and the nonlinear loop of the reference driving the CUDA
# library.  Only where the reference is present; the binary travels to the GPU box with the snapshot.
REFERENCE_MAIN = os.path.join(os.environ.get("MGIC_REFERENCE", "/root/reference"), "Main_PoissonSolver.cpp")
EXE_REFMAIN = os.path.join(PKG, "lib", "Main_PoissonSolver_b200")


def stale():
    if not os.path.exists(EXE):
        return True
    t = os.path.getmtime(EXE)
    if os.path.exists(REFERENCE_MAIN) and not os.path.exists(EXE_REFMAIN):
        return True
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".H", ".cpp", ".py"))]
    deps += [os.path.join(HERE, "dropin", f) for f in os.listdir(os.path.join(HERE, "dropin"))]
    deps += [os.path.join(ROOT, "include", "mgic.h"), os.path.join(PKG, "lib", "libmgic_b200.so")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False):
    if not force and not stale():
        return EXE
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    link = ["-L" + os.path.join(PKG, "lib"), "-lmgic_b200", "-Wl,-rpath,$ORIGIN"]
    cmd = [cxx, "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + HERE] + \
          [os.path.join(HERE, s) for s in SRC] + link + ["-o", EXE]
    subprocess.check_call(cmd)
    if os.path.exists(REFERENCE_MAIN):
        cmd = [cxx, "-O2", "-std=c++17", "-I" + os.path.join(HERE, "dropin"), "-I" + HERE, "-I" + os.path.join(ROOT, "include"),
               REFERENCE_MAIN] + [os.path.join(HERE, s) for s in SRC[:2]] + link + ["-o", EXE_REFMAIN]
        subprocess.check_call(cmd)
    return EXE


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
