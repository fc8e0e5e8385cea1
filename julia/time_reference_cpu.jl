# time_reference_cpu.jl — UNEXECUTED.  Times the real reference path (NumericalEarth update_state!)
# on the CPU with all threads, for an out-of-band comparison with bench.py's cpu_baseline (which is
# the oracle port, not the Julia reference).   julia -t auto julia/time_reference_cpu.jl
using ClimaOcean, Oceananigans, Printf
function main(; Nx = 4320, Ny = 1800, Nz = 75, repeats = 5)
    grid = LatitudeLongitudeGrid(CPU(); size = (Nx, Ny, Nz), longitude = (0, 360), latitude = (-75, 75), z = (-5000, 0), halo = (7, 7, 7))
    ocean = ocean_simulation(grid)
    # atmosphere = synthetic PrescribedAtmosphere as in dump_reference.jl
    # model = OceanSeaIceModel(ocean; atmosphere)
    # best = minimum(@elapsed(NumericalEarth.EarthSystemModels.update_state!(model)) for _ in 1:repeats)
    # @printf("%d threads: %.3f s  →  %.2f Mcells/s\n", Threads.nthreads(), best, Nx * Ny / best / 1e6)
    error("construct the synthetic PrescribedAtmosphere for the installed NumericalEarth version first")
end
abspath(PROGRAM_FILE) == @__FILE__ && main()
