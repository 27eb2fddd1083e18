# Golden vectors from the REAL reference (marcsgil/GeneralizedGrossPitaevskii.jl, unmodified, CPU arrays = FFTW +
# KernelAbstractions CPU backend): pins the oracle (oracle/ggp_oracle.py) and the CUDA path to an actual run of the
# package instead of to the builder's reading of it (SURVEY §8c: "parity unpinned" until this has been run).
#
# NOT executed in the build image (no Julia there).  On any machine with Julia >= 1.10 and the reference checkout:
#
#   python tests/golden/make_golden.py --export-inputs          # writes tests/golden/ref_inputs_v1/*.npy  (NumPy only)
#   julia --project=/path/to/GeneralizedGrossPitaevskii.jl tests/golden/make_golden_ref.jl
#   git add tests/golden/ref_v1 && python -m pytest tests/test_golden_ref.py        # CPU: oracle vs reference
#   python -m pytest tests/test_golden_ref.py -m gpu                                # B200: CUDA path vs reference
#
# The cases are those of tests/golden/make_golden.py (same closures as tests/problems.py, i.e. the reference's own test
# and example closures); the initial fields -- and, for the stochastic cases, every noise array the reference draws, in
# its draw order (src/misc.jl:44-51) -- are read from the exported inputs, so both sides integrate identical data:
# the reference's `rng` keyword gets a replay generator whose `randn!` returns the recorded arrays.
#
# Output: tests/golden/ref_v1/<case>__ts.npy and <case>__u<c>.npy (final save), written so that NumPy reads the same
# (batch..., n_d, ..., n_1) C-order arrays the oracle produces (Julia's column-major (n_1, ..., n_d, batch...) memory).
using GeneralizedGrossPitaevskii, Random, LinearAlgebra

const HERE = @__DIR__
const INP = joinpath(HERE, "ref_inputs_v1")
const OUT = joinpath(HERE, "ref_v1")
isdir(INP) || error("run `python tests/golden/make_golden.py --export-inputs` first (looked for $INP)")
mkpath(OUT)

# ---- minimal .npy reader / writer (format 1.0, little-endian, C order) ---------------------------------------------
const DESCR = Dict("<c16" => ComplexF64, "<c8" => ComplexF32, "<f8" => Float64, "<f4" => Float32)
function read_npy(path)
    open(path) do io
        magic = read(io, 6); magic == UInt8[0x93, 0x4e, 0x55, 0x4d, 0x50, 0x59] || error("$path: not an .npy file")
        major = read(io, UInt8); read(io, UInt8)
        hlen = major == 1 ? Int(ltoh(read(io, UInt16))) : Int(ltoh(read(io, UInt32)))
        header = String(read(io, hlen))
        occursin("'fortran_order': False", header) || error("$path: expected C order")
        T = DESCR[match(r"'descr': '([^']+)'", header).captures[1]]
        shp = [parse(Int, m.match) for m in eachmatch(r"\d+", match(r"'shape': \(([^)]*)\)", header).captures[1])]
        data = Vector{T}(undef, prod(shp; init=1))
        read!(io, data)
        reshape(data, reverse(shp)...)          # NumPy (batch..., n_d, ..., n_1) C order == Julia (n_1, ..., n_d, batch...)
    end
end
function write_npy(path, a::Array{T}) where {T}
    descr = first(k for (k, v) in DESCR if v == T)
    shp = join(reverse(size(a)), ", ") * (ndims(a) == 1 ? "," : "")
    header = "{'descr': '$descr', 'fortran_order': False, 'shape': ($shp), }"
    pad = 64 - (10 + length(header) + 1) % 64
    header *= " "^(pad % 64) * "\n"
    open(path, "w") do io
        write(io, UInt8[0x93, 0x4e, 0x55, 0x4d, 0x50, 0x59, 0x01, 0x00]); write(io, htol(UInt16(length(header))))
        write(io, header); write(io, a)
    end
end

# ---- replay generator: hands the recorded noise arrays to the reference's sample_noise! (src/misc.jl:44-51) ---------
mutable struct ReplayRNG <: Random.AbstractRNG
    bufs::Vector{Any}
    idx::Int
end
function Random.randn!(r::ReplayRNG, x::AbstractArray)
    r.idx += 1
    copyto!(x, r.bufs[r.idx])
    x
end
function replay(case)
    bufs = Any[]
    k = 0
    while isfile(joinpath(INP, "$(case)__xi_$(k).npy"))
        push!(bufs, read_npy(joinpath(INP, "$(case)__xi_$(k).npy"))); k += 1
    end
    isempty(bufs) && error("no recorded noise for $case")
    ReplayRNG(bufs, 0)
end

u0_of(case, M) = ntuple(c -> read_npy(joinpath(INP, "$(case)__u0_$(c-1).npy")), M)
function save(case, ts, sol)
    write_npy(joinpath(OUT, "$(case)__ts.npy"), collect(ts))
    for (c, s) in enumerate(sol)
        nd = ndims(s)
        write_npy(joinpath(OUT, "$(case)__u$(c-1).npy"), Array(selectdim(s, nd, size(s, nd))))
    end
    println(case, ": ", map(size, sol), "  max|u1| = ", maximum(abs, selectdim(sol[1], ndims(sol[1]), size(sol[1], ndims(sol[1])))))
end
run(case, prob, tspan; kw...) = save(case, solve(prob, StrangSplitting(), tspan; show_progress=false, kw...)...)

# ---- the cases (tests/problems.py, file:line of the reference closures cited there) ---------------------------------
free_disp(ks, param) = sum(abs2, ks) / 2

function kerr(case, R, L, nsteps, d)                               # problems.kerr2d / kerr3d
    u0 = u0_of(case, 1)
    nl(u, p) = p.g * abs2(u[1])
    prob = GrossPitaevskiiProblem(u0, ntuple(_ -> R(L), d); dispersion=free_disp, nonlinearity=nl, param=(; g=R(1)))
    run(case, prob, (R(0), R(nsteps) * R(1e-3)); dt=R(1e-3), nsaves=1)
end
kerr("kerr2d_c128", Float64, 16.0, 8, 2)
kerr("kerr2d_c64", Float32, 16.0, 8, 2)
kerr("kerr3d_c128", Float64, 8.0, 4, 3)

let case = "quick_start_kerr"                                      # examples/quick_start.jl, N = 32
    nl(u, p) = p.g * abs2(u[1])
    prob = GrossPitaevskiiProblem(u0_of(case, 1), (8, 8); dispersion=free_disp, nonlinearity=nl, param=(; g=-6))
    run(case, prob, (0, 0.4); dt=0.01, nsaves=64)
end

function exciton_polariton(case; nsaves, tspan, dt, time_pump)     # test/exciton_polariton_test.jl:1-46
    ħ = 0.654; Ωr = 5.07 / 2ħ; γx = 0.0015 / ħ; γc = 0.07 / 0.6571 / ħ
    ωx = 1484.44 / ħ; ωc = 1482.76 / ħ; m = ħ^2 / (2 * 2e-1); ωp = ωc
    δx = ωp - ωx; δc = ωp - ωc; A = 2; w = 100; g = 1e-2 / ħ; L = 256
    param = (; ħ, m, δc, γc, δx, γx, Ωr, A, w, g, L, tmax=last(tspan))
    function dispersion(k, p)
        Dcc = p.ħ * sum(abs2, k) / 2p.m - p.δc - im * p.γc
        Dxx = -p.δx - im * p.γx
        @SMatrix [Dcc p.Ωr; p.Ωr Dxx]
    end
    nonlinearity(ψ, p) = @SVector [0, p.g * abs2(ψ[2])]
    function pump(r, p, t)
        amp = 1.0
        if time_pump                                              # examples/bistability.jl:71-74 envelope
            val = -t * (t - p.tmax) * 4 / p.tmax^2
            amp = val > 0 ? sqrt(val) : 0.0
        end
        @SVector [p.A * exp(-sum(x -> (x - p.L / 2)^2, r) / p.w^2) * amp, 0]
    end
    prob = GrossPitaevskiiProblem(u0_of(case, 2), (L, L); dispersion, nonlinearity, pump, param)
    run(case, prob, tspan; dt, nsaves)
end
exciton_polariton("exciton_polariton"; nsaves=4, tspan=(0, 1.5625), dt=1e-1, time_pump=false)
exciton_polariton("exciton_polariton_time_pump"; nsaves=2, tspan=(0, 1.6), dt=0.05, time_pump=true)

let case = "bistability"                                           # test/bistability_cycle.jl:1-55, n = 64
    ω₀ = 1483; g = 0.01; δ = 0.3; kz = 27; γ = 0.1; ωₚ = ω₀ + δ; L = 256; Imax = 0.6; width = 50
    tspan = (0, 25.78125)
    param = (; tmax=last(tspan), Imax, width, ωₚ, ω₀, kz, γ, g, L)
    dispersion(ks, p) = -im * p.γ / 2 + p.ω₀ * (1 + sum(abs2, ks) / 2p.kz^2) - p.ωₚ
    nonlinearity(ψ, p) = p.g * abs2(ψ[1])
    I(t, tmax, Imax) = (val = -Imax * t * (t - tmax) * 4 / tmax^2; val < 0 ? 0.0 : val)
    pump(x, p, t) = exp(-sum(abs2, x .- p.L / 2) / p.width^2) * √I(t, p.tmax, p.Imax)
    prob = GrossPitaevskiiProblem(u0_of(case, 1), (L,); dispersion, nonlinearity, pump, param)
    run(case, prob, tspan; dt=0.05, nsaves=4)
end

let case = "windowed_ft_noise"                                     # test/windowed_ft.jl:23-29,62-90, 8 trajectories
    L = 20; N = 64; dL = L / N; ħ = 0.6582; γ = 0.1 / ħ; m = ħ^2 / 2.5; δ₀ = 0.49 / ħ
    param = (; δ₀, m, γ, ħ, L, dL, N, dt=4)
    dispersion(ks, p) = -im * p.γ / 2 + p.ħ * sum(abs2, ks) / 2p.m - p.δ₀
    noise(ψ, r, p) = √(p.γ / 2 / p.dL)
    u0 = u0_of(case, 1)
    prob = GrossPitaevskiiProblem(u0, (L,); dispersion, param, position_noise_func=noise, noise_prototype=similar.(u0))
    run(case, prob, (0, 200); dt=4, nsaves=1, save_start=false, rng=replay(case))
end

let case = "truncated_wigner_2d_noise"                             # examples/truncated_wigner.jl:33-50,93-99 in 2-D
    ħ = 0.6582; γ = 0.047 / ħ; m = 1 / 6; g = 3e-4 / ħ; δ = 0.49 / ħ; A = 10; L = 512; N = 16
    vol = (L / N)^2
    param = (; ħ, m, δ, γ, g, A, L, dx=vol)
    dispersion(ks, p) = p.ħ * sum(abs2, ks) / 2p.m - p.δ - im * p.γ / 2
    pump(x, p, t) = p.A
    nonlinearity(ψ, p) = p.g * (abs2(ψ[1]) - 1 / p.dx)
    noise(ψ, xs, p) = √(p.γ / 2p.dx)
    u0 = u0_of(case, 1)
    prob = GrossPitaevskiiProblem(u0, (L, L); dispersion, nonlinearity, pump, param, position_noise_func=noise,
        noise_prototype=similar.(u0))
    run(case, prob, (0, 0.5); dt=0.05, nsaves=1, save_start=false, rng=replay(case))
end

function noise_forms(case, form, d, M)                             # docs/src/stochastic_simulations.md:62-86 (problems.noise_forms)
    param = (; α=0.3, β=0.7, σ2=30.0, c=0.2, g=0.5, γ=0.1)
    dispersion(ks, p) = sum(abs2, ks) / 2 - im * p.γ / 2
    nonlinearity(u, p) = p.g * abs2(u[1])
    prof(r, p) = p.β * exp(-sum(abs2, r) / p.σ2)
    η = if form == "field"
        (u, r, p) -> p.c + p.α * abs(u[1])
    elseif M == 1
        (u, r, p) -> prof(r, p) * (p.c + p.α * abs(u[1]))
    else
        (u, r, p) -> @SVector [prof(r, p) * (p.c + p.α * abs(u[2])), prof(r, p) * (2p.c + 0.5p.α * abs(u[1]))]
    end
    u0 = u0_of(case, M)
    prob = GrossPitaevskiiProblem(u0, ntuple(_ -> 10.0, d); dispersion, nonlinearity, param, position_noise_func=η,
        noise_prototype=similar.(u0))
    run(case, prob, (0, 0.5); dt=0.05, nsaves=2, save_start=true, rng=replay(case))
end
noise_forms("noise_field_1d", "field", 1, 1)
noise_forms("noise_profile_q2_2d_two_comp", "both", 2, 2)

function generic_case(case)                                        # problems.generic: the generic plan's shapes
    p = (; g=0.8, om=0.6, gamma=0.05, v=0.3)
    disp_scalar(ks, p) = sum(abs2, ks) / 2 - im * p.gamma / 2
    pump_static(rs, p, t) = 0.4 * exp(-sum(abs2, rs) / 8)
    dt = 0.01; tspan = (0.0, 6 * dt)
    prob = if case == "generic_np2_2d"                             # 96 x 80 points: no power of two on either axis
        nl(u, p) = p.g * (abs2(u[1]) - 0.1im)
        GrossPitaevskiiProblem(u0_of(case, 1), (6.0, 8.0); dispersion=disp_scalar, nonlinearity=nl, pump=pump_static, param=p)
    elseif case == "generic_rabi_vv"                               # SMatrix nonlinearity x SVector potential
        nl2(u, p) = @SMatrix [p.g*abs2(u[1])+0.2*abs2(u[2]) p.om; p.om p.g*abs2(u[2])-0.03im]
        pot(rs, p) = @SVector [p.v * sum(abs2, rs) / 10, 0.2]
        GrossPitaevskiiProblem(u0_of(case, 2), (6.0, 8.0); dispersion=disp_scalar, nonlinearity=nl2, potential=pot,
            pump=pump_static, param=p)
    else                                                           # three components, 3 x 3 matrix dispersion, 2 trajectories
        disp3(ks, p) = @SMatrix [sum(abs2, ks)/2 p.om 0; p.om sum(abs2, ks)/3-im*p.gamma 0.5*p.om; 0 0.5*p.om 0.1]
        nl3(u, p) = p.g * (abs2(u[1]) + abs2(u[3]))
        GrossPitaevskiiProblem(u0_of(case, 3), (6.0,); dispersion=disp3, nonlinearity=nl3, param=p)
    end
    run(case, prob, tspan; dt, nsaves=2)
end
foreach(generic_case, ("generic_np2_2d", "generic_rabi_vv", "generic_m3_matdisp"))

println("wrote ", OUT)
