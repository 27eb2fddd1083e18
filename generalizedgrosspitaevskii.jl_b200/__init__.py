"""B200-native backend for the Strang-splitting time step of GeneralizedGrossPitaevskii.jl.

Only what the hot path needs: `csrc/` (sm_100a kernels + the C ABI of include/ggp.h -> libggp.so),
`host.py` (the reference's user surface above the C ABI) and `julia/` (the ccall shim a Julia user
loads).  Import as `ggp_b200` (see ggp_b200.py at the repository root: the directory name contains
a dot and cannot be imported directly).
"""
from ._lib import GgpError, LIB_PATH, load  # noqa: F401
from .host import (  # noqa: F401
    GrossPitaevskiiProblem, SMatrix, SVector, StrangSplitting, StrangSplittingIterator, UnsupportedForm,
    abs2, additiveIdentity, direct_grid, init, multiplicativeIdentity, reciprocal_grid,
    resolve_fixed_timestepping, solve, solve_, step_,
)
from . import _lib as lib  # noqa: F401
from . import parallel  # noqa: F401
