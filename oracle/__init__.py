"""CPU oracle package (TEST INFRASTRUCTURE ONLY -- see mgic_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this.  PARITY: pinned bit for bit to the reference's
own C++ compiled into oracle/_ref (pyref: source terms, parameters, ParseBC, operator
class, factory) and to a mechanical translation of its .ChF kernels (chf2c.py; no
Fortran compiler); everything that is Chombo's is restated and UNPINNED (not vendored);
see DESIGN.md.
"""
from .pyoracle import Oracle, OraclePatch, OrcParams, lib, build, default_params, FIELD, interp_homo, use_all_host_cores, condition_box, output_box  # noqa: F401
