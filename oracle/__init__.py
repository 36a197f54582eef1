"""CPU oracle package (TEST INFRASTRUCTURE ONLY -- see mgic_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this.  PARITY: pinned bit for bit to the reference's
own C++ compiled into oracle/_ref (pyref: source terms, parameters, ParseBC, operator
class, factory); the .ChF kernels' arithmetic and everything that is Chombo's are
restated and UNPINNED (no Fortran compiler, Chombo not vendored); see DESIGN.md.
"""
from .pyoracle import Oracle, OraclePatch, OrcParams, lib, build, default_params, FIELD, interp_homo  # noqa: F401
