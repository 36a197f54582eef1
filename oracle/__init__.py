"""CPU oracle package (TEST INFRASTRUCTURE ONLY -- see mgic_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this.  PARITY UNPINNED: the reference ships
no golden vectors and cannot be built here; see DESIGN.md.
"""
from .pyoracle import Oracle, OraclePatch, OrcParams, lib, build, default_params, FIELD, interp_homo  # noqa: F401
