# Closure recognition: the user's Julia closures stay untouched; the ones that must run on the
# device are matched against the registered forms by probing (SURVEY §7 hard part 2, §8a).
# An unrecognised closure throws -- the B200 backend has no CPU fallback.

_aslist(x::Number, M) = (ntuple(_ -> ComplexF64(x), M), true)
_aslist(x::SVector, M) = (ntuple(i -> ComplexF64(x[i]), M), false)
_aslist(x::SMatrix{1,1}, M) = (ntuple(_ -> ComplexF64(x[1]), M), false)

"G_i(u) = c_i + Σ_j g_ij |u_j|²  ->  (scalar::Bool, c::Vector{ComplexF64}, g::Matrix{ComplexF64})"
function recognise_nonlinearity(f, param, ::Val{M}) where {M}
    rng = Random.Xoshiro(0xC0FFEE)
    P = 4 * (M + 1) + 8
    probe() = SVector{M,ComplexF64}(ntuple(_ -> (0.2 + 1.8rand(rng)) * cis(2π * rand(rng)), M))
    us = [probe() for _ in 1:P]
    vals = [_aslist(f(u, param), M) for u in us]
    scalar = vals[1][2]
    A = [j == 0 ? 1.0 : abs2(us[p][j]) for p in 1:P, j in 0:M]
    c = zeros(ComplexF64, M); g = zeros(ComplexF64, M, M)
    for i in 1:M
        coef = A \ ComplexF64[v[1][i] for v in vals]
        c[i] = coef[1]; g[i, :] .= coef[2:end]
    end
    scale = max(maximum(abs, c), maximum(abs, g), floatmin(Float64))
    for _ in 1:16                                    # held-out verification (also catches phase dependence)
        u = probe(); v = _aslist(f(u, param), M)[1]
        for i in 1:M
            pred = c[i] + sum(g[i, j] * abs2(u[j]) for j in 1:M)
            abs(pred - v[i]) ≤ 1e-9 * max(scale, abs(v[i])) ||
                error("nonlinearity is not of the registered form c_i + Σ_j g_ij |u_j|² (no CPU fallback)")
        end
    end
    c[abs.(c).<1e-13*scale] .= 0; g[abs.(g).<1e-13*scale] .= 0
    scalar, c, g
end

"F(r,t) = S(r)·a(t): returns (S::Matrix{ComplexF64} (ncomp × npoints), ncomp, amp::Function)"
function recognise_pump(pump, prob, tspan, times)
    rs = direct_grid(prob)
    pts = Iterators.product(rs...)
    M = length(prob.u0)
    ongrid(t) = begin
        vals = [pump(r, prob.param, t) for r in pts]
        v1 = first(vals)
        nc = v1 isa Number ? 1 : length(v1)
        S = Matrix{ComplexF64}(undef, nc, length(vals))
        for (k, v) in enumerate(vals), c in 1:nc
            S[c, k] = v isa Number ? v : v[c]
        end
        S
    end
    cand = Any[first(tspan), last(tspan), (first(tspan) + last(tspan)) / 2]
    isempty(times) || append!(cand, (times[1], times[max(1, end ÷ 3)], times[max(1, 2end ÷ 3)]))
    S, tref = ongrid(cand[1]), cand[1]
    for t in cand[2:end]
        F = ongrid(t)
        maximum(abs, F) > maximum(abs, S) && ((S, tref) = (F, t))
    end
    ncomp = size(S, 1)
    (ncomp == 1 || ncomp == M) || error("pump must return a Number or an SVector of length M")
    if maximum(abs, S) == 0
        return S, ncomp, t -> zero(ComplexF64)
    end
    idx = argmax(abs.(S)); cidx, pidx = Tuple(idx)
    rpt = collect(pts)[pidx]; ref = S[idx]
    amp(t) = begin
        v = pump(rpt, prob.param, t)
        ComplexF64(v isa Number ? v : v[cidx]) / ref
    end
    for t in cand[1:3]                                # separability check on the full grid
        F = ongrid(t)
        maximum(abs, F .- amp(t) .* S) ≤ 1e-10 * max(maximum(abs, F), maximum(abs, S)) ||
            error("pump is not separable as S(r)·a(t) (no CPU fallback)")
    end
    S, ncomp, amp
end

"η_i = const per component"
function recognise_noise(f, prob)
    M = length(prob.u0)
    rng = Random.Xoshiro(0xBEEF)
    rs = direct_grid(prob)
    vals = map(1:4) do _
        u = SVector{M,ComplexF64}(ntuple(_ -> randn(rng, ComplexF64), M))
        r = map(g -> g[rand(rng, 1:length(g))], rs)
        collect(_aslist(f(u, r, prob.param), M)[1])
    end
    all(v -> maximum(abs, v .- vals[1]) ≤ 1e-12 * max(maximum(abs, vals[1]), floatmin(Float64)), vals) ||
        error("field- or position-dependent noise amplitudes are not a registered form yet (no CPU fallback)")
    vals[1]
end
