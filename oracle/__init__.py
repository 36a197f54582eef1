"""CPU oracle package (TEST INFRASTRUCTURE ONLY -- see mgic_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this.  PARITY: source terms pinned bit for bit to
the reference's own code (pyref / oracle/_ref); the operator path is UNPINNED (the
reference ships no golden vectors and that part cannot be built here); see DESIGN.md.
"""
from .pyoracle import Oracle, OraclePatch, OrcParams, lib, build, default_params, FIELD, interp_homo  # noqa: F401
