# Drop-in replacement for the reference's src/strang_splitting.jl + src/kernels.jl and for the inner
# loop of src/fixed_time_stepping.jl: same `StrangSplitting`, same `init / step! / solve! / solve`
# signatures and return types; the time loop runs in libggp.so.
#
# What stays Julia (SURVEY §8b): resolve_fixed_timestepping (kept verbatim in fixed_time_stepping.jl),
# the exp tables through the reference's own get_exponential on HOST arrays (exact parity, arbitrary
# closures), `ts` accumulation in type T, ProgressMeter, result allocation, RNG seed derivation.

struct StrangSplitting <: FixedTimeSteppingAlgorithm end

mutable struct StrangSplittingIterator{PROB,T,PROG1,PROG2,R,A,TT} <: FixedTimeSteppingIterator
    prob::PROB
    dt::T
    ts::Vector{T}
    steps_per_save::Int
    save_start::Bool
    progress::PROG1
    given_progress::PROG2
    result::R
    handle::Ptr{Cvoid}
    amps::A                 # 2 × nsteps ComplexF64, or nothing (static pump / no pump / dense pump)
    step_index::Int
    dense_pump::Bool        # the pump does not separate as S(r)a(t): profiles are evaluated on the host per half-step
    pump_times::TT          # the reference's pump evaluation times (quirk Q1), 2 per step; nothing without a pump
    pinned::Bool            # iter.result is page-locked (ggp_host_register)
end

_kind(::MultiplicativeIdentity) = GGP_TABLE_NONE
_kind(::Array{<:Number}) = GGP_TABLE_SCALAR
_kind(::Array{<:SVector}) = GGP_TABLE_DIAG
_kind(::Array{<:SMatrix}) = GGP_TABLE_FULL
# host table -> ComplexF64 array-of-structs exactly as Julia lays it out (point-major, static entries column-major)
_flat(::MultiplicativeIdentity) = ComplexF64[]
_flat(t::Array{<:Number}) = ComplexF64.(vec(t))
_flat(t::Array{<:StaticArray}) = ComplexF64[x for s in vec(t) for x in s]
_ptr(v::Vector{ComplexF64}) = isempty(v) ? Ptr{Cvoid}(C_NULL) : Ptr{Cvoid}(pointer(v))

function init(prob::GrossPitaevskiiProblem{N,M}, ::StrangSplitting, tspan;
    dt, nsaves, show_progress=true, progress=nothing, save_start=true, workgroup_size=(), rng=nothing,
    device=-1, batch_offset=0) where {N,M}

    result = map(x -> stack(x for _ ∈ 1:nsaves+save_start), prob.u0)                  # reference :41-43
    dt, ts, steps_per_save = resolve_fixed_timestepping(dt, tspan, nsaves)            # :45
    _progress = _Progress(progress, steps_per_save * nsaves; enabled=show_progress)    # :46

    CT = eltype(first(prob.u0))
    CT <: Union{ComplexF32,ComplexF64} || error("fields must be ComplexF32 or ComplexF64")
    sz = size(first(prob.u0))
    nspatial = prod(sz[1:N]); nbatch = prod(sz[N+1:end])

    # tables: the reference's own get_exponential on host arrays (src/misc.jl:12-20)
    rg, dg = reciprocal_grid(prob), direct_grid(prob)
    host_u0 = map(Array, prob.u0)
    exp_D = get_exponential(prob.dispersion, host_u0, rg, prob.param, dt)             # :53 (full dt)
    exp_V = get_exponential(prob.potential, host_u0, dg, prob.param, dt / 2)          # :54
    # 1x1 wrappers (SVector{1}, SMatrix{1,1}) collapse to scalars; M-vectors / MxM matrices keep their kind
    dflat, vflat = _flat(exp_D), _flat(exp_V)
    dkind = (M == 1 && _kind(exp_D) != GGP_TABLE_NONE) ? GGP_TABLE_SCALAR : _kind(exp_D)
    vkind = (M == 1 && _kind(exp_V) != GGP_TABLE_NONE) ? GGP_TABLE_SCALAR : _kind(exp_V)

    # ComplexF32 problems: keep the separable-dispersion fast path on large grids (include/ggp.h: disp_sep_tol)
    sep_tol = (CT == ComplexF32 && dkind == GGP_TABLE_SCALAR && exp_D isa AbstractArray{<:Number}) ?
              separable_dispersion_tol(prob.dispersion, rg, prob.param, exp_D) : 0.0

    M ≤ GGP_MAX_COMPONENTS || error("the B200 backend is built for up to $GGP_MAX_COMPONENTS components")
    nl_kind = Int32(0); nl_scalar = Int32(0); nl_c = ntuple(_ -> 0.0, 4); nl_g = ntuple(_ -> 0.0, 8)
    nl_c_ext = Float64[]; nl_g_ext = Float64[]          # ABI 5: M > 2 and SMatrix nonlinearities (generic plan)
    if !(prob.nonlinearity isa AdditiveIdentity)
        kind, c, g = recognise_nonlinearity(prob.nonlinearity, prob.param, Val(M))
        nl_scalar = Int32(kind == :scalar)
        if kind == :matrix || M > 2
            nl_kind = Int32(kind == :matrix ? 2 : 1)
            # a Number-valued closure: one row (c, g_j); SVector: M rows; SMatrix: C_ij and g_ijk
            nl_c_ext = _reim_rowmajor(kind == :scalar ? c[1:1] : c)
            nl_g_ext = _reim_rowmajor(kind == :scalar ? g[1:1, :] : g)
        else
            nl_kind = Int32(1)
            nl_c = ntuple(k -> (i = (k - 1) ÷ 2 + 1; i > M ? 0.0 : (isodd(k) ? real(c[i]) : imag(c[i]))), 4)
            nl_g = ntuple(k -> begin
                    i = (k - 1) ÷ 4 + 1; j = ((k - 1) ÷ 2) % 2 + 1
                    (i > M || j > M) ? 0.0 : (isodd(k) ? real(g[i, j]) : imag(g[i, j]))
                end, 8)
        end
    end

    # pump amplitudes at the reference's times: t is incremented BEFORE step! (SURVEY Q1)
    nsteps = nsaves * steps_per_save
    pump_kind = Int32(0); pump_ncomp = Int32(0); sflat = ComplexF64[]; amp0 = (0.0, 0.0); amps = nothing
    dense_pump = false; pump_times = nothing
    if !(prob.pump isa AdditiveIdentity)
        t = ts[1]; times = Vector{typeof(t)}(undef, 2nsteps)
        for s in 1:nsteps
            t += dt; times[2s-1] = t + dt / 2; times[2s] = t + dt
        end
        S, ncomp, amp, dense, zero_pump = recognise_pump(prob.pump, prob, ts[1], times)
        pump_times = times
        if zero_pump
            # F = 0 at every scheduled time: no pump term
        elseif dense
            # not S(r)a(t): the reference's own procedure, profile by profile (evaluate_pump!, src/misc.jl:34-42)
            dense_pump = true
            pump_kind = GGP_PUMP_DENSE; pump_ncomp = Int32(ncomp)
            sflat = pump_on_grid(prob.pump, prob, ts[1])                              # primed at tspan[1], :58
        else
            pump_kind = GGP_PUMP_SEPARABLE; pump_ncomp = Int32(ncomp); sflat = vec(S) # point-major, then component
            a0 = amp(ts[1]); amp0 = (real(a0), imag(a0))                              # primed at tspan[1], :58
            A = ComplexF64[amp(times[k]) for k in 1:2nsteps]
            amps = all(==(a0), A) ? nothing : reshape(A, 2, nsteps)
        end
    end

    noise_kind = Int32(0); noise_real = Int32(0); eta = ntuple(_ -> 0.0, 4); seed = UInt64(0)
    alpha = ntuple(_ -> 0.0, 8); nprof = ComplexF64[]
    eta_ext = Float64[]; alpha_ext = Float64[]          # ABI 5: M > 2
    if !(prob.position_noise_func isa AdditiveIdentity)
        e, al, P = recognise_noise(prob.position_noise_func, prob)
        # the reference indexes ξ with the leading ndims(ξ) indices only (src/kernels.jl:24,27): a prototype without
        # the batch dims shares one noise field between trajectories; the device draws per (element, trajectory)
        (length(prob.noise_prototype) == M && all(x -> size(x) == sz, prob.noise_prototype)) ||
            error("noise_prototype must hold one array per component with the size of u0 (not a registered form)")
        allequal(map(x -> eltype(x) <: Real, prob.noise_prototype)) ||
            error("noise_prototype arrays must be all real or all complex")
        field = any(!iszero, al) || P !== nothing
        noise_kind = Int32(field ? 2 : 1); noise_real = Int32(eltype(first(prob.noise_prototype)) <: Real)
        P === nothing || (nprof = P)
        if M > 2
            eta_ext = _reim_rowmajor(collect(e)); alpha_ext = _reim_rowmajor(al)
        else
            alpha = ntuple(k -> begin
                    i = (k - 1) ÷ 4 + 1; j = ((k - 1) ÷ 2) % 2 + 1
                    (i > M || j > M) ? 0.0 : (isodd(k) ? real(al[i, j]) : imag(al[i, j]))
                end, 8)
            eta = ntuple(k -> (i = (k - 1) ÷ 2 + 1; i > M ? 0.0 : (isodd(k) ? real(e[i]) : imag(e[i]))), 4)
        end
        seed = rng === nothing ? rand(UInt64) : rand(rng, UInt64)
    end

    desc = GgpDesc(GGP_ABI_VERSION, UInt32(sizeof(GgpDesc)), Int32(N), Int32(M),
        ntuple(i -> i ≤ N ? Int64(sz[i]) : Int64(1), 3), Int64(nbatch), Int64(batch_offset),
        CT == ComplexF32 ? GGP_C64 : GGP_C128, GGP_C128, Int32(device), Int32(0), C_NULL, Float64(dt),
        dkind, vkind, _ptr(dflat), _ptr(vflat), nl_kind, nl_scalar, nl_c, nl_g,
        pump_kind, pump_ncomp, _ptr(sflat), amp0, noise_kind, noise_real, eta, seed, Int32(0), Int32(0),
        alpha, _ptr(nprof), sep_tol, (Ptr{Cvoid}(C_NULL), Ptr{Cvoid}(C_NULL), Ptr{Cvoid}(C_NULL)),
        Int32(CT == ComplexF32 && exp_D isa AbstractArray && real(eltype(eltype(exp_D))) == Float64), Int32(0),
        _fptr(nl_c_ext), _fptr(nl_g_ext), _fptr(eta_ext), _fptr(alpha_ext))
    handle = GC.@preserve dflat vflat sflat nprof nl_c_ext nl_g_ext eta_ext alpha_ext ggp_plan_create(desc)
    ggp_set_state(handle, host_u0)                                                    # u = copy.(prob.u0), :48

    # page-lock `result` so that the streaming saves of solve! are asynchronous DMA transfers
    pinned = false
    try
        foreach(ggp_host_register, result); pinned = true
    catch
        pinned = false                     # pageable memory still works: the copies then synchronise
    end
    iter = StrangSplittingIterator(prob, dt, ts, steps_per_save, save_start, _progress, progress, result,
        handle, amps, 0, dense_pump, pump_times, pinned)
    finalizer(iter) do it
        it.handle == C_NULL || ggp_plan_destroy(it.handle)
        it.handle = C_NULL
    end
    iter
end

# nsteps steps from the current position of the schedule: amplitudes (separable pump), profiles (dense pump) or nothing
function _advance!(iter::StrangSplittingIterator, nsteps::Int)
    i0 = iter.step_index
    if iter.dense_pump
        done = 0
        while done < nsteps                                   # a few steps at a time: 2k full-grid profiles on the host
            k = min(8, nsteps - done)
            profs = Vector{ComplexF64}[pump_on_grid(iter.prob.pump, iter.prob, iter.pump_times[2(i0+done+i)-1+h])
                                       for i in 1:k for h in 0:1]
            ggp_step_dense(iter.handle, k, profs)
            done += k
        end
    else
        a = iter.amps === nothing ? nothing : view(iter.amps, :, i0+1:i0+nsteps)
        ggp_step(iter.handle, nsteps, a)
    end
    iter.step_index += nsteps
    nothing
end

# step!(iter, t, dt): one Strang step on the device (reference :86-90).  t/dt are accepted for
# signature compatibility; the pump schedule was fixed at init.
step!(iter::StrangSplittingIterator, t, dt) = _advance!(iter, 1)

# solve!: the reference's loop (src/fixed_time_stepping.jl:26-54) with the inner `for _ in 1:steps_per_save`
# batched into one ggp_step call and `map(copy!, slice, iter.u)` replaced by one streaming save per interval
# (device snapshot + D2H on a second stream, overlapped with the next interval; one wait at the end).
function solve!(iter::StrangSplittingIterator)
    save_start, sps, dt, p, ts = iter.save_start, iter.steps_per_save, iter.dt, iter.progress, iter.ts
    nd = ndims(first(iter.result))
    t = ts[1]
    for n ∈ 1:size(first(iter.result), nd)-save_start
        _advance!(iter, sps)
        for _ ∈ 1:sps
            t += dt                                   # accumulated in T exactly like the reference (:44)
            _next!(p)
        end
        slices = map(x -> selectdim(x, nd, n + save_start), iter.result)   # contiguous: last dim
        ggp_save_async(iter.handle, slices)
        ts[n+1] = t
    end
    ggp_save_wait(iter.handle)
    if iter.pinned
        foreach(ggp_host_unregister, iter.result); iter.pinned = false
    end
    _finish!(p, iter.given_progress)
    ts[begin+1-save_start:end], iter.result
end

# Checkpoint / resume (SURVEY §8f N3; the reference has none).  `checkpoint(iter)` returns the bytes that let
# `restore!(init(prob, alg, tspan; same kwargs...), blob)` continue bit-identically (fields, Philox counter word,
# F_now amplitude, position in the pump schedule).
checkpoint(iter::StrangSplittingIterator) = vcat(ggp_checkpoint(iter.handle), reinterpret(UInt8, [Int64(iter.step_index)]))
function restore!(iter::StrangSplittingIterator, blob::Vector{UInt8})
    ggp_restore(iter.handle, blob[1:end-8])
    iter.step_index = Int(reinterpret(Int64, blob[end-7:end])[1])
    iter
end

# On-device ensemble observables (SURVEY §8f N1): `observe(iter, :momentum)` returns Σ_traj |fft(u_c)/N|² per component
# without downloading the ensemble; `:g2` the N×N trajectory sum of examples/truncated_wigner.jl:143-154 (1-D).
function observe(iter::StrangSplittingIterator, what::Symbol)
    N = length(iter.prob.lengths); M = length(iter.prob.u0)
    sz = size(first(iter.prob.u0))[1:N]
    if what === :norm
        return ggp_observe!(zeros(Float64, M), iter.handle, GGP_OBS_NORM)
    elseif what === :g2
        N == 1 || error("g2 is defined for 1-D ensembles")
        return ggp_observe!(zeros(Float64, sz[1], sz[1], M), iter.handle, GGP_OBS_G2_MOMENTUM)   # [n, m, c] (symmetric in m, n)
    end
    kind = what === :density ? GGP_OBS_DENSITY : what === :momentum ? GGP_OBS_MOMENTUM : error("unknown observable $what")
    ggp_observe!(zeros(Float64, sz..., M), iter.handle, kind)
end
