# Thin ccall layer over libggp.so (include/ggp.h).  Nothing here computes: every grid-sized
# floating-point operation of the time loop happens inside the CUDA library.
#
# NOTE: written without a Julia toolchain in the build image (SURVEY: "No Julia anywhere in the
# loop"); kept deliberately small and mechanical so that it can be reviewed by eye.  The Python host
# mirror (../host.py) implements the same logic line for line and IS exercised by the test-suite.
using Libdl

const GGP_ABI_VERSION = UInt32(5)
const GGP_MAX_COMPONENTS = 4
const GGP_C64, GGP_C128 = Int32(0), Int32(1)
const GGP_TABLE_NONE, GGP_TABLE_SCALAR, GGP_TABLE_DIAG, GGP_TABLE_FULL, GGP_TABLE_SEP_AXES = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)
const GGP_PUMP_NONE, GGP_PUMP_SEPARABLE, GGP_PUMP_DENSE = Int32(0), Int32(1), Int32(2)

# Mirror of `struct ggp_desc`; field order and types must match include/ggp.h exactly
# (tests/test_host_logic.py checks the ctypes twin of this struct against the C header).
struct GgpDesc
    abi_version::UInt32
    struct_size::UInt32
    ndim::Int32
    ncomp::Int32
    n::NTuple{3,Int64}
    nbatch::Int64
    batch_offset::Int64
    precision::Int32
    table_precision::Int32
    device::Int32
    reserved0::Int32
    stream::Ptr{Cvoid}
    dt::Float64
    disp_kind::Int32
    pot_kind::Int32
    disp_table::Ptr{Cvoid}
    pot_table::Ptr{Cvoid}
    nl_kind::Int32
    nl_scalar::Int32
    nl_c::NTuple{4,Float64}      # [i][re/im]
    nl_g::NTuple{8,Float64}      # [i][j][re/im]
    pump_kind::Int32
    pump_ncomp::Int32
    pump_table::Ptr{Cvoid}
    pump_amp0::NTuple{2,Float64}
    noise_kind::Int32
    noise_real::Int32
    noise_eta::NTuple{4,Float64}
    seed::UInt64
    slab_nranks::Int32           # 3-D slab decomposition (0/1 = off), see include/ggp.h
    slab_rank::Int32
    noise_alpha::NTuple{8,Float64}   # GGP_NOISE_FIELD: alpha_ij [i][j][re/im]
    noise_profile::Ptr{Cvoid}        # GGP_NOISE_FIELD: n[1] ComplexF64 values P(point(k1)) or C_NULL (quirk Q2)
    disp_sep_tol::Float64            # 0 = library default; see include/ggp.h (separable dispersion)
    disp_axes::NTuple{3,Ptr{Cvoid}}  # GGP_TABLE_SEP_AXES (ABI 4): per-axis factors of a scalar exp_D, else C_NULL
    mixed_precision_tables::Int32    # 1: ComplexF32 fields with ComplexF64 tables in the reference (quirk Q6)
    reserved1::Int32
    # ABI 5 (generic plan: M > 2, SMatrix nonlinearities): flat (re, im) coefficient arrays, C_NULL when unused
    nl_c_ext::Ptr{Cvoid}             # c_i (M) / C_ij (M², [i][j] row-major)
    nl_g_ext::Ptr{Cvoid}             # g_ij (M², [i][j]) / g_ijk (M³, [i][j][k])
    noise_eta_ext::Ptr{Cvoid}        # η_i (M)
    noise_alpha_ext::Ptr{Cvoid}      # α_ij (M², [i][j])
end

const _lib = Ref{Ptr{Cvoid}}(C_NULL)

function _handle()
    if _lib[] == C_NULL
        path = get(ENV, "GGP_LIBRARY", joinpath(@__DIR__, "..", "..", "libggp.so"))
        _lib[] = Libdl.dlopen(path)          # throws if missing: there is no CPU fallback
    end
    _lib[]
end

_sym(name::Symbol) = Libdl.dlsym(_handle(), name)

function _check(rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall(_sym(:ggp_last_error), Cstring, ()))
    error("libggp error $rc: $msg")
end

function ggp_plan_create(desc::GgpDesc)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    _check(ccall(_sym(:ggp_plan_create), Cint, (Ref{GgpDesc}, Ref{Ptr{Cvoid}}), desc, out))
    out[]
end

ggp_plan_destroy(h::Ptr{Cvoid}) = (ccall(_sym(:ggp_plan_destroy), Cint, (Ptr{Cvoid},), h); nothing)

function ggp_set_state(h::Ptr{Cvoid}, u::NTuple{M,<:Array}) where {M}
    ptrs = Ptr{Cvoid}[pointer(x) for x in u]
    GC.@preserve u ptrs _check(ccall(_sym(:ggp_set_state), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), h, ptrs))
end

# dest: tuple of contiguous views/arrays (one per component)
function ggp_get_state(h::Ptr{Cvoid}, dest::NTuple{M,Any}) where {M}
    ptrs = Ptr{Cvoid}[Ptr{Cvoid}(pointer(x)) for x in dest]
    GC.@preserve dest ptrs _check(ccall(_sym(:ggp_get_state), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), h, ptrs))
end

# Streaming save (include/ggp.h: ggp_save_async / ggp_save_wait): device snapshot in stream order, PCIe transfer
# on a second stream while the next interval steps.  `dest` must stay alive until ggp_save_wait (solve! keeps
# iter.result alive); page-locking it (cudaHostRegister through CUDA.jl, or ggp_host_alloc-backed arrays via
# unsafe_wrap) makes the transfer truly asynchronous -- pageable memory still works, the copy then synchronises.
function ggp_save_async(h::Ptr{Cvoid}, dest::NTuple{M,Any}) where {M}
    ptrs = Ptr{Cvoid}[Ptr{Cvoid}(pointer(x)) for x in dest]
    GC.@preserve dest ptrs _check(ccall(_sym(:ggp_save_async), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), h, ptrs))
end
ggp_save_wait(h::Ptr{Cvoid}) = _check(ccall(_sym(:ggp_save_wait), Cint, (Ptr{Cvoid},), h))

# Ensemble observables summed over the plan's trajectories on the device (all-reduced over ranks when a
# communicator is attached): kind 0 density, 1 n(k) = Σ|fft(u)/N|², 2 norm, 3 Σ|F(m)|²|F(n)|² (1-D, the trajectory
# sum inside `G2` of examples/truncated_wigner.jl:143-154).  `out` must hold M·nspatial (0, 1), M (2) or M·N² (3) doubles.
const GGP_OBS_DENSITY, GGP_OBS_MOMENTUM, GGP_OBS_NORM, GGP_OBS_G2_MOMENTUM = Cint(0), Cint(1), Cint(2), Cint(3)
function ggp_observe!(out::Array{Float64}, h::Ptr{Cvoid}, kind::Cint)
    GC.@preserve out _check(ccall(_sym(:ggp_observe), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), h, kind, pointer(out)))
    out
end

# Windowed first-order coherence (test/windowed_ft.jl:31-49 `correlation`, before its division by length(sol)):
# w1, w2 = window values on the direct grid; returns an N×N×M ComplexF64 array laid out [j, i, c] in Julia's
# column-major order (the C side writes out[c][i][j]).
function ggp_observe_windowed(h::Ptr{Cvoid}, w1::Vector{ComplexF64}, w2::Vector{ComplexF64}, M::Integer)
    N = length(w1)
    out = Array{ComplexF64}(undef, N, N, M)
    GC.@preserve w1 w2 out _check(ccall(_sym(:ggp_observe_windowed), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        h, Ptr{Float64}(pointer(w1)), Ptr{Float64}(pointer(w2)), Ptr{Float64}(pointer(out))))
    out
end

# Checkpoint / resume (include/ggp.h): fields + Philox counter word + F_now amplitude as one byte vector
function ggp_checkpoint(h::Ptr{Cvoid})
    n = ccall(_sym(:ggp_checkpoint_bytes), Int64, (Ptr{Cvoid},), h)
    n < 0 && _check(Cint(n))
    blob = Vector{UInt8}(undef, n)
    GC.@preserve blob _check(ccall(_sym(:ggp_checkpoint_save), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, UInt64), h, pointer(blob), n))
    blob
end
function ggp_restore(h::Ptr{Cvoid}, blob::Vector{UInt8})
    GC.@preserve blob _check(ccall(_sym(:ggp_checkpoint_load), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, UInt64), h, pointer(blob), length(blob)))
end

# amps: 2 x nsteps ComplexF64 (column s = [a(t_s + dt/2), a(t_s + dt)]) or nothing (static / no pump)
function ggp_step(h::Ptr{Cvoid}, nsteps::Integer, amps::Union{Nothing,AbstractMatrix{ComplexF64}})
    if amps === nothing
        _check(ccall(_sym(:ggp_step), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Ptr{Cvoid}}), h, nsteps, C_NULL, C_NULL))
    else
        a = Matrix{ComplexF64}(amps)
        GC.@preserve a _check(ccall(_sym(:ggp_step), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Ptr{Cvoid}}),
            h, nsteps, Ptr{Float64}(pointer(a)), C_NULL))
    end
end

# GGP_PUMP_DENSE plans (ABI 4): profiles = 2*nsteps host arrays, the pump evaluated on the direct grid at the
# reference's half-step times (evaluate_pump!, src/misc.jl:34-42), each npoints*ncomp ComplexF64, point-major
function ggp_step_dense(h::Ptr{Cvoid}, nsteps::Integer, profiles::Vector{Vector{ComplexF64}}, noise=nothing)
    ptrs = Ptr{Cvoid}[pointer(x) for x in profiles]
    nptrs = noise === nothing ? Ptr{Cvoid}[] : Ptr{Cvoid}[pointer(x) for x in noise]
    pn = noise === nothing ? Ptr{Ptr{Cvoid}}(C_NULL) : pointer(nptrs)
    GC.@preserve profiles ptrs noise nptrs _check(ccall(_sym(:ggp_step_dense), Cint,
        (Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}), h, nsteps, ptrs, pn))
end

# Page-lock caller-owned memory (iter.result) so that ggp_save_async is a true asynchronous DMA (ABI 4)
ggp_host_register(x::Array) = _check(ccall(_sym(:ggp_host_register), Cint, (Ptr{Cvoid}, UInt64), pointer(x), sizeof(x)))
ggp_host_unregister(x::Array) = _check(ccall(_sym(:ggp_host_unregister), Cint, (Ptr{Cvoid},), pointer(x)))

# TEST MODE: feed the reference's own noise buffers (step, half-step, component order)
function ggp_step_with_noise(h::Ptr{Cvoid}, nsteps::Integer, amps, noise::Vector{<:Array})
    ptrs = Ptr{Cvoid}[pointer(x) for x in noise]
    a = amps === nothing ? nothing : Matrix{ComplexF64}(amps)
    pa = a === nothing ? Ptr{Float64}(C_NULL) : Ptr{Float64}(pointer(a))
    GC.@preserve noise ptrs a _check(ccall(_sym(:ggp_step), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Ptr{Cvoid}}),
        h, nsteps, pa, ptrs))
end
