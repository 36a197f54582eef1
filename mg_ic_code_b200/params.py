"""params.txt reader: the keys getPoissonParameters (Source/PoissonParameters.cpp:26-131), ParseBC
(Source/SetBCs.cpp:45-56) and poissonSolve (Main_PoissonSolver.cpp:106-126) read, `key = value` with '#'
comments and command-line style overrides, like [Chombo] ParmParse."""
import ctypes as C

from ._capi import MgicParams, MgicError

# params.txt of the reference (params.txt:12-84)
DEFAULTS = dict(
    alpha=1.0, beta=-1.0, G_Newton=1.0, phi_amplitude=0.1, phi_wavelength=1.0,
    bh1_bare_mass=0.5, bh1_spin=0.1, bh1_momentum=0.05, bh1_offset=10.0,
    bh2_bare_mass=0.5, bh2_spin=0.1, bh2_momentum=-0.05, bh2_offset=-10.0,
    L=100.0, bc_value=0.0, tolerance=1.0e-10, N=(64, 64, 64), max_level=0, block_factor=8,
    max_grid_size=16, coefficient_average_type="harmonic", is_periodic=0, bc_lo=(0, 0, 0), bc_hi=(0, 0, 0),
    numMGsmooth=4, numMGIterations=2, preCondSolverDepth=-1, max_iterations=100, max_NL_iterations=6,
    verbosity=2,
)
_INT3 = ("N", "bc_lo", "bc_hi")
_INTS = ("max_level", "block_factor", "max_grid_size", "is_periodic", "numMGsmooth", "numMGIterations",
         "preCondSolverDepth", "max_iterations", "max_NL_iterations", "verbosity")
_AVG = {"arithmetic": 0, "harmonic": 1}


def parse_params_text(text):
    out = {}
    for line in text.splitlines():
        line = line.split("#", 1)[0].strip()
        if not line or "=" not in line:
            continue
        k, v = line.split("=", 1)
        out[k.strip()] = v.split()
    return out


def make_params(d=None, **over):
    """dict -> MgicParams.  Missing solver keys take the reference's in-code defaults
    (Main_PoissonSolver.cpp:106-126: numMGIterations 4, numMGsmooth 4, preCondSolverDepth -1, tolerance 1e-7,
    max_iterations 100, max_NL_iterations 4) only when read from a file; this helper starts from params.txt."""
    vals = dict(DEFAULTS)
    if d:
        vals.update(d)
    vals.update(over)
    p = MgicParams()
    for k, v in vals.items():
        if k in _INT3:
            v = [int(x) for x in v]
            if len(v) != 3:
                raise MgicError(f"{k} needs 3 values")
            setattr(p, k, (C.c_int * 3)(*v))
        elif k == "coefficient_average_type":
            if isinstance(v, str):
                if v not in _AVG:
                    raise MgicError("bad coefficient_average_type in input")  # PoissonParameters.cpp:106
                v = _AVG[v]
            p.coefficient_average_type = int(v)
        elif k in _INTS:
            setattr(p, k, int(v))
        elif hasattr(p, k):
            setattr(p, k, float(v))
    return p


# keys the reference reads with ParmParse::get / getarr, i.e. whose absence stops it (Source/PoissonParameters.cpp:30-126,
# Source/SetBCs.cpp:45,55-56) ...
REQUIRED = ("alpha", "beta", "G_Newton", "phi_amplitude", "phi_wavelength", "bh1_bare_mass", "bh2_bare_mass", "bh1_spin", "bh2_spin",
            "bh1_offset", "bh2_offset", "bh1_momentum", "bh2_momentum", "max_level", "N", "L", "refine_threshold", "block_factor",
            "max_grid_size", "fill_ratio", "buffer_size", "is_periodic", "bc_lo", "bc_hi", "bc_value")
# ... and the in-code defaults of the keys it reads with query (PoissonParameters.cpp:59-60, Main_PoissonSolver.cpp:106-126)
QUERY_DEFAULTS = dict(verbosity=3, numMGIterations=1, numMGsmooth=4, preCondSolverDepth=-1, tolerance=1.0e-7, max_iterations=10,
                      max_NL_iterations=4)


def read_params(path, overrides=(), strict=False):
    """Read a reference-format params.txt; `overrides` are 'key=value' strings (ParmParse CLI overrides,
    Main_PoissonSolver.cpp:272).  strict=True is the reference's behaviour: a missing required key is an error and the
    optional solver keys take the reference's in-code defaults; otherwise missing keys take params.txt's values (DEFAULTS)."""
    with open(path) as f:
        raw = parse_params_text(f.read())
    for o in overrides:
        k, v = o.split("=", 1)
        raw[k.strip()] = v.split()
    if strict:
        for k in REQUIRED:
            if k not in raw or not raw[k]:
                raise MgicError(f"ParmParse::get: {k} not found")   # MayDay::Error in the reference
        for k, v in QUERY_DEFAULTS.items():
            raw.setdefault(k, [str(v)])
    d = {}
    for k, v in raw.items():
        if k in _INT3:
            d[k] = [int(x) for x in v]
        elif k == "coefficient_average_type":
            d[k] = v[0]
        elif k in DEFAULTS:
            d[k] = v[0]
    if "coefficient_average_type" not in raw:
        d["coefficient_average_type"] = -1  # "bogus default": solver default (arithmetic) applies
    return make_params(d)
