# Times the UNTOUCHED reference (marcsgil/GeneralizedGrossPitaevskii.jl, CPU arrays: KernelAbstractions CPU backend +
# FFTW) on the headline workload of bench.py: C2 = 2-D scalar Kerr GPE 2048^2 ComplexF32, D = |k|^2/2, g = +1,
# dt = 1e-3 (SURVEY.md §8d).  NOT executed in the build image (no Julia there): shipped so that anyone with Julia can
# put the real reference next to `python bench.py --impl reference`, which times a CPU restatement instead.
#
#   JULIA_NUM_THREADS=16 julia --project=/path/to/GeneralizedGrossPitaevskii.jl bench/ref_cpu.jl [N] [nsteps]
#
# Prints one JSON line in bench.py's format (metric grid-point-steps/s).
using GeneralizedGrossPitaevskii, FFTW, Random

N = length(ARGS) ≥ 1 ? parse(Int, ARGS[1]) : 2048
nsteps = length(ARGS) ≥ 2 ? parse(Int, ARGS[2]) : 200
FFTW.set_num_threads(Threads.nthreads())

L = 64.0f0
rs = StepRangeLen(0.0f0, L / N, N)
rng = Random.MersenneTwister(1234)          # the synthetic field of bench.py uses NumPy's PCG64: same statistics, other draws
u0 = (ComplexF32[exp(-((x - L / 2)^2 + (y - L / 2)^2) / 16) * (1 + 0.1f0 * randn(rng, ComplexF32)) for x in rs, y in rs],)
dispersion(ks, param) = sum(abs2, ks) / 2
nonlinearity(u, param) = param.g * abs2(u[1])
prob = GrossPitaevskiiProblem(u0, (L, L); dispersion, nonlinearity, param=(; g=1.0f0))
dt = 1.0f-3

solve(prob, StrangSplitting(), (0.0f0, 3dt); dt, nsaves=1, show_progress=false)              # compile + FFTW planning
t = @elapsed solve(prob, StrangSplitting(), (0.0f0, nsteps * dt); dt, nsaves=1, show_progress=false)
value = N^2 * nsteps / t
println("{\"impl\": \"reference-julia\", \"metric\": \"grid-point-steps/s\", \"value\": $value, \"unit\": \"grid-point-steps/s\", ",
        "\"steps\": $nsteps, \"ms_per_step\": $(1e3 * t / nsteps), \"dtype\": \"c64\", \"data\": \"synthetic\", ",
        "\"config\": {\"workload\": \"C2: 2-D scalar Kerr GPE $(N)^2 ComplexF32\", \"threads\": $(Threads.nthreads())}}")
