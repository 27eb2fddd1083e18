# Module root of the B200 backend: the reference's module (src/GeneralizedGrossPitaevskii.jl:1-20) with its
# Strang-splitting algorithm replaced by the libggp.so plan.  Everything that is NOT on the hot path -- the problem
# type, the grids, resolve_fixed_timestepping, `solve`, get_exponential, the progress shims -- is the reference's OWN
# source, included from the reference checkout (never copied):
#
#     GGP_REFERENCE_SRC=/path/to/GeneralizedGrossPitaevskii.jl/src     (default: /root/reference/src)
#
# UNEXECUTED in the build image (no Julia toolchain there); `julia/test/runtests.jl` is the one-command check.
module GeneralizedGrossPitaevskii

using KernelAbstractions, FFTW, LinearAlgebra, Random, ProgressMeter
import CommonSolve: solve, init, step!, solve!

using Reexport
@reexport using StaticArrays

const REFERENCE_SRC = get(ENV, "GGP_REFERENCE_SRC", "/root/reference/src")
isfile(joinpath(REFERENCE_SRC, "problem.jl")) ||
    error("set GGP_REFERENCE_SRC to the src/ directory of a GeneralizedGrossPitaevskii.jl checkout (looked in $REFERENCE_SRC)")

using DispatchDoctor: @stable
@stable default_mode = "disable" begin
    include(joinpath(REFERENCE_SRC, "problem.jl"))               # GrossPitaevskiiProblem, direct_grid, reciprocal_grid
    include(joinpath(REFERENCE_SRC, "kernels.jl"))               # AdditiveIdentity / MultiplicativeIdentity, _cis (host tables)
    include(joinpath(REFERENCE_SRC, "misc.jl"))                  # get_exponential / grid_map! on HOST arrays, progress shims
    include(joinpath(REFERENCE_SRC, "fixed_time_stepping.jl"))   # resolve_fixed_timestepping, solve; solve! is specialised below
end
# the reference's src/strang_splitting.jl is NOT included: these three files replace it
include("libggp.jl")                # ccall layer (outside @stable: Ptr / Ref / dlsym)
include("registered_forms.jl")      # closure recognition (probing is intentionally dynamic)
include("strang_splitting.jl")      # StrangSplitting, init / step! / solve! on the plan handle

export GrossPitaevskiiProblem, solve, StrangSplitting

end
