# time_reference_cpu.jl — times the real reference path (NumericalEarth update_state!) on the CPU with all threads, for an
# out-of-band comparison with bench.py's cpu_baseline (which is the oracle port, not the Julia reference).  UNEXECUTED here
# (no Julia in the build image).        julia -t auto --project=/path/to/ClimaOcean.jl julia/time_reference_cpu.jl [Nx Ny Nz]
include(joinpath(@__DIR__, "synthetic_inputs.jl"))
using Oceananigans.TimeSteppers: update_state!
using Printf

function main(; Nx = 4320, Ny = 1800, Nz = 75, repeats = 5, flux_configuration = :default)
    model = synthetic_coupled_model(CPU(), Nx, Ny, Nz; latitude = (-75, 75), flux_configuration)
    model.clock.time = 1.37 * 3hours
    update_state!(model)                                    # compile + first touch
    best = minimum(@elapsed(update_state!(model)) for _ in 1:repeats)
    @printf("{\"impl\": \"julia-reference\", \"threads\": %d, \"cells\": %d, \"seconds\": %.6f, \"Mcells_per_s\": %.3f}\n",
            Threads.nthreads(), Nx * Ny, best, Nx * Ny / best / 1e6)
end

if abspath(PROGRAM_FILE) == @__FILE__
    length(ARGS) >= 3 ? main(Nx = parse(Int, ARGS[1]), Ny = parse(Int, ARGS[2]), Nz = parse(Int, ARGS[3])) : main()
end
