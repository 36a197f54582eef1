"""mg_ic_code_b200 -- B200-native multigrid hot path of MG_IC_code (VariableCoeffPoissonOperator V-cycle).

Layout: csrc/ (CUDA kernels + C ABI, built into lib/libmgic_b200.so), operator.py (host mirror of the
reference's operator / factory interface), params.py (params.txt reader).  See DESIGN.md.
"""
from ._capi import MgicError, MgicParams, lib, library_path  # noqa: F401
from .params import DEFAULTS, make_params, read_params  # noqa: F401
from .operator import (AMRHierarchy, Context, Grids, Hierarchy, LevelField, MultigridVars, VariableCoeffPoissonOperator,  # noqa: F401
                       VariableCoeffPoissonOperatorFactory, level_op_from_params, nl_solve)
