# Closure recognition: the user's Julia closures stay untouched; the ones that must run on the
# device are matched against the registered forms by probing (SURVEY §7 hard part 2, §8a).
# An unrecognised closure throws -- the B200 backend has no CPU fallback.

_aslist(x::Number, M) = (ntuple(_ -> ComplexF64(x), M), :scalar)
_aslist(x::SVector, M) = (ntuple(i -> ComplexF64(x[i]), M), :vector)
_aslist(x::SMatrix{1,1}, M) = (ntuple(_ -> ComplexF64(x[1]), M), :vector)
# M x M SMatrix (M > 1): the M² entries row-major [i][j]  (src/kernels.jl:22-25: cis(::SMatrix) is the matrix exponential)
_aslist(x::SMatrix{K,K}, M) where {K} = (ntuple(k -> ComplexF64(x[(k-1)÷K+1, (k-1)%K+1]), K * K), :matrix)

"""
Registered nonlinearity forms (host.py: recognise_nonlinearity), fitted by probing and verified on held-out samples:
  Number / SVector   G_i(u)  = c_i  + Σ_j g_ij |u_j|²    -> (:scalar | :vector, c::Vector (M), g::Matrix (M × M))
  SMatrix (M × M)    G_ij(u) = C_ij + Σ_k g_ijk |u_k|²   -> (:matrix, C::Matrix (M × M), g::Array (M × M × M))
"""
function recognise_nonlinearity(f, param, ::Val{M}) where {M}
    rng = Random.Xoshiro(0xC0FFEE)
    P = 4 * (M + 1) + 8
    probe() = SVector{M,ComplexF64}(ntuple(_ -> (0.2 + 1.8 * rand(rng)) * cis(2π * rand(rng)), M))
    us = [probe() for _ in 1:P]
    vals = [_aslist(f(u, param), M) for u in us]
    kind = vals[1][2]
    nrows = length(vals[1][1])                       # M, or M² for the matrix form
    A = [j == 0 ? 1.0 : abs2(us[p][j]) for p in 1:P, j in 0:M]
    c = zeros(ComplexF64, nrows); g = zeros(ComplexF64, nrows, M)
    for i in 1:nrows
        coef = A \ ComplexF64[v[1][i] for v in vals]
        c[i] = coef[1]; g[i, :] .= coef[2:end]
    end
    scale = max(maximum(abs, c), maximum(abs, g), floatmin(Float64))
    for _ in 1:16                                    # held-out verification (also catches phase dependence)
        u = probe(); v = _aslist(f(u, param), M)[1]
        for i in 1:nrows
            pred = c[i] + sum(g[i, j] * abs2(u[j]) for j in 1:M)
            abs(pred - v[i]) ≤ 1e-9 * max(scale, abs(v[i])) ||
                error("nonlinearity is not of a registered form (c_i + Σ_j g_ij |u_j|², or the same per entry of an SMatrix); no CPU fallback")
        end
    end
    c[abs.(c).<1e-13*scale] .= 0; g[abs.(g).<1e-13*scale] .= 0
    if kind == :matrix
        # rows are [i][j] row-major: C[i,j] = c[(i-1)M+j], G3[i,j,k] = g[(i-1)M+j, k]
        C = [c[(i-1)*M+j] for i in 1:M, j in 1:M]
        G3 = [g[(i-1)*M+j, k] for i in 1:M, j in 1:M, k in 1:M]
        return :matrix, C, G3
    end
    kind, c, g
end

# flat (re, im) Float64 arrays in the row-major [i][j]([k]) order of include/ggp.h (Julia arrays are column-major)
_reim_rowmajor(v::AbstractVector) = Float64[f(x) for x in v for f in (real, imag)]
_reim_rowmajor(m::AbstractMatrix) = Float64[f(m[i, j]) for i in axes(m, 1) for j in axes(m, 2) for f in (real, imag)]
_reim_rowmajor(a::AbstractArray{<:Any,3}) =
    Float64[f(a[i, j, k]) for i in axes(a, 1) for j in axes(a, 2) for k in axes(a, 3) for f in (real, imag)]
_fptr(v::Vector{Float64}) = isempty(v) ? Ptr{Cvoid}(C_NULL) : Ptr{Cvoid}(pointer(v))

"The pump closure on the whole direct grid at time t, point-major then component (grid_map!, src/misc.jl:34-37)"
function pump_on_grid(pump, prob, t)
    rs = direct_grid(prob)
    vals = [pump(r, prob.param, t) for r in Iterators.product(rs...)]
    v1 = first(vals)
    nc = v1 isa Number ? 1 : length(v1)
    out = Vector{ComplexF64}(undef, nc * length(vals))
    for (k, v) in enumerate(vals), c in 1:nc
        out[(k-1)*nc+c] = v isa Number ? v : v[c]
    end
    out
end

"""
F(r,t) = S(r)·a(t) or a dense pump (host.py: PumpModel).  Returns `(S (ncomp × npoints), ncomp, amp, dense, zero)`.
Decisions are taken over EVERY scheduled time, never a handful of samples: the closure is evaluated at K probe points
for all of `times`; `zero` needs F = 0 at all of them, separability needs F(r_k,t) = a(t)·S(r_k) at all of them (plus a
few full-grid checks).  `dense = true`: the caller evaluates `pump_on_grid` per half-step (GGP_PUMP_DENSE).
"""
function recognise_pump(pump, prob, t0, times)
    rs = direct_grid(prob)
    pts = collect(Iterators.product(rs...))
    M = length(prob.u0)
    ongrid(t) = begin
        flat = pump_on_grid(pump, prob, t)
        nc = length(flat) ÷ length(pts)
        reshape(flat, nc, length(pts))
    end
    sched = Any[t0; times]
    cand = Any[t0]
    isempty(times) || append!(cand, [times[1+(k*(length(times)-1))÷7] for k in 0:7])
    S, tref = ongrid(cand[1]), cand[1]
    for t in cand[2:end]
        F = ongrid(t)
        maximum(abs, F) > maximum(abs, S) && ((S, tref) = (F, t))
    end
    ncomp = size(S, 1)
    (ncomp == 1 || ncomp == M) || error("pump must return a Number or an SVector of length M")
    np = length(pts); K = 12
    mag = vec(maximum(abs, S; dims=1))
    order = sortperm(mag; rev=true)
    probe = unique(vcat([order[1+(k*(np-1))÷(K-1)] for k in 0:K-1], [1 + (k * (np - 1)) ÷ (K - 1) for k in 0:K-1]))
    at(t) = [(v = pump(pts[k], prob.param, t); ComplexF64(v isa Number ? v : v[c])) for c in 1:ncomp, k in probe]
    vals = [at(t) for t in sched]
    if maximum(v -> maximum(abs, v), vals) == 0 && maximum(abs, S) == 0
        return S, ncomp, t -> zero(ComplexF64), false, true
    end
    if maximum(abs, S) == 0                      # every full-grid candidate vanishes, but a probe saw the pump
        tref = sched[argmax(map(v -> maximum(abs, v), vals))]
        S = ongrid(tref)
    end
    idx = argmax(abs.(S)); cidx, pidx = Tuple(idx)
    rpt = pts[pidx]; ref = S[idx]
    amp(t) = begin
        v = pump(rpt, prob.param, t)
        ComplexF64(v isa Number ? v : v[cidx]) / ref
    end
    A = ComplexF64[amp(t) for t in sched]
    Sk = S[:, probe]
    scale = max(maximum(v -> maximum(abs, v), vals), maximum(abs, Sk) * maximum(abs, A), floatmin(Float64))
    for (i, v) in enumerate(vals)
        maximum(abs, v .- A[i] .* Sk) ≤ 1e-10 * scale || return S, ncomp, amp, true, false
    end
    for t in cand                                 # ... and on the full grid at the candidate times
        F = ongrid(t)
        maximum(abs, F .- amp(t) .* S) ≤ 1e-10 * max(maximum(abs, F), maximum(abs, S)) || return S, ncomp, amp, true, false
    end
    S, ncomp, amp, false, false
end

"""
Registered noise form  η_i(u, r) = P(r)·(e_i + Σ_j a_ij |u_j|)  (host.py: recognise_noise): constant amplitudes,
`α·abs(u[1])` and spatial profiles of docs/src/stochastic_simulations.md:62-86.  Returns `(e, a, P)`; `P` is
`nothing` or the n₁ profile values at the reference's points `(x[k], y[k], …)` — `build_field_at(grid, K)` indexes every
grid axis with `K[1]` (src/kernels.jl:27,41; quirk Q2).
"""
function recognise_noise(f, prob)
    M = length(prob.u0)
    rng = Random.Xoshiro(0xBEEF)
    rs = direct_grid(prob)
    n1 = length(rs[1])
    inbounds = all(g -> length(g) ≥ n1, rs)
    pts = inbounds ? [map(g -> g[k], rs) for k in 1:n1] : nothing
    ev(u, r) = ComplexF64.(collect(_aslist(f(SVector{M,ComplexF64}(u), r, prob.param), M)[1]))
    probe() = ntuple(_ -> (0.2 + 1.8 * rand(rng)) * cis(2π * rand(rng)), M)
    uref = probe()
    along = pts === nothing ? nothing : [ev(uref, r) for r in pts]
    k0 = along === nothing ? 0 : argmax(map(v -> maximum(abs, v), along))
    r0 = along === nothing ? map(g -> g[cld(length(g), 2)], rs) : pts[k0]
    U = [probe() for _ in 1:(4 * (M + 1) + 8)]
    A = [j == 0 ? 1.0 : abs(u[j]) for u in U, j in 0:M]
    Y = permutedims(reduce(hcat, [ev(u, r0) for u in U]))
    coef = A \ Y                                                  # (M+1) × M
    V = [probe() for _ in 1:12]
    pred = [j == 0 ? 1.0 : abs(u[j]) for u in V, j in 0:M] * coef
    truth = permutedims(reduce(hcat, [ev(u, r0) for u in V]))
    scale = max(maximum(abs, truth), maximum(abs, coef), floatmin(Float64))
    maximum(abs, pred .- truth) ≤ 1e-9 * scale ||
        error("noise amplitude is not of the registered form P(r)·(e_i + Σ_j a_ij |u_j|) (no CPU fallback)")
    coef[abs.(coef) .< 1e-13 * scale] .= 0
    e, a = coef[1, :], permutedims(coef[2:end, :])                 # a[i, j]
    P = nothing
    if pts === nothing
        maximum(abs, ev(uref, map(first, rs)) .- ev(uref, r0)) ≤ 1e-12 * scale ||
            error("position-dependent noise with n₁ longer than another axis: the reference's `point` is out of bounds")
    else
        ref = along[k0]; c = argmax(abs.(ref))
        if abs(ref[c]) > 0
            prof = ComplexF64[v[c] / ref[c] for v in along]
            if maximum(abs, prof .- 1) > 1e-12
                all(k -> maximum(abs, along[k] .- prof[k] .* ref) ≤ 1e-9 * scale, 1:n1) ||
                    error("noise amplitude does not separate as P(r) × (field part)")
                P = prof
            end
        end
    end
    e, a, P
end

"""
`disp_sep_tol` of include/ggp.h (host.py: separable_dispersion_tol): for a ComplexF32 problem whose scalar dispersion
is a sum over axes in Float64 arithmetic, the deviation of the table from the product of its own axis factors (its
rounding, eps32·|phase|) with a 25 % margin; otherwise 0 (library default).
"""
function separable_dispersion_tol(D, rg, param, table::AbstractArray{<:Number})
    d = length(rg); d < 2 && return 0.0
    rng = Random.Xoshiro(0x5E9)
    g = map(x -> Float64.(collect(x)), rg)
    for _ in 1:2048
        k = map(x -> x[rand(rng, 1:length(x))], g)
        z = map(x -> x[1], g)
        full = D(k, param); full isa Number || return 0.0
        acc = ComplexF64(full) + (d - 1) * ComplexF64(D(z, param))
        for a in 1:d
            acc -= ComplexF64(D(ntuple(b -> b == a ? k[b] : z[b], d), param))
        end
        abs(acc) ≤ 1e-12 * max(abs(full), floatmin(Float64)) || return 0.0
    end
    tab = reshape(ComplexF64.(table), :, size(table, d))          # (perp, line): the strided kernel runs along the last axis
    d0 = tab[1, 1]; (d0 == 0 || !isfinite(d0)) && return 0.0
    dev = maximum(abs, tab .- tab[:, 1] .* permutedims(tab[1, :]) ./ d0)
    1.25 * dev / maximum(abs, tab) + 1e-9
end
