# dump_reference.jl — UNEXECUTED helper for someone with Julia + network access.
#
# Builds BASELINE config 1 (64×32×8 LatitudeLongitudeGrid, synthetic PrescribedAtmosphere) with the
# real ClimaOcean/NumericalEarth stack on CPU, fills every input with the synthetic pattern of
# climaocean.jl_b200/synth.py (SplitMix64 + triangle waves: integer/IEEE-basic-ops only, so the values
# are bit-identical across languages), calls the REAL update_state!, and writes every input and output
# array as raw little-endian Float64 under `tests/golden/julia_c1/<name>.bin` plus `manifest.txt`
# (name, size).  tests/test_julia_golden.py picks these up when present and compares the CUDA path
# against them at 1e-12 — that is the step that turns "parity unpinned" into a pinned parity claim.
using ClimaOcean, Oceananigans
using Oceananigans.Units

const SEED_BASE = 0xC0F10000
splitmix(x::UInt64) = (z = x; z = (z ⊻ (z >> 30)) * 0xBF58476D1CE4E5B9; z = (z ⊻ (z >> 27)) * 0x94D049BB133111EB; z ⊻ (z >> 31))
u01(field_id, k) = Float64(splitmix(UInt64(SEED_BASE + field_id) + (UInt64(k) + 1) * 0x9E3779B97F4A7C15) >> 11) * 2.0^-53
tri(m, N) = abs(2.0 * (mod(m, N) / N) - 1.0)
function pattern(field_id, lo, hi, i, j, Nxg, Nyg; level = 0, w = 0.5)   # zero-based global i (periodic), j
    I = mod(i, Nxg); J = j
    k = (J + 64) * Nxg + I + level * Nxg * (Nyg + 128)
    p1, p2 = 1 + field_id % 3, 1 + (field_id ÷ 3) % 2
    s = 0.5 * tri(I * p1 + (field_id * 37) % Nxg, Nxg) + 0.5 * tri((J + 64) * p2 + (field_id * 11) % Nyg, Nyg)
    return lo + (hi - lo) * (w * u01(field_id, k) + (1 - w) * s)
end

function main(outdir = joinpath(@__DIR__, "..", "tests", "golden", "julia_c1"))
    mkpath(outdir)
    grid = LatitudeLongitudeGrid(CPU(); size = (64, 32, 8), longitude = (0, 360), latitude = (-60, 60), z = (-5000, 0), halo = (7, 7, 7))
    ocean = ocean_simulation(grid)
    u, v = ocean.model.velocities.u, ocean.model.velocities.v
    T, S = ocean.model.tracers.T, ocean.model.tracers.S
    for (f, id, lo, hi) in ((u, 1, -1.0, 1.0), (v, 2, -1.0, 1.0), (T, 3, -1.8, 30.0), (S, 4, 30.0, 38.0))
        p = parent(f)
        for kk in axes(p, 3), jj in axes(p, 2), ii in axes(p, 1)
            p[ii, jj, kk] = pattern(id, lo, hi, ii - 1 - 7, jj - 1 - 7, 64, 32; level = clamp(kk - 1 - 7, 0, 7))
        end
    end
    # Synthetic PrescribedAtmosphere on 640×320, 8 levels 3 h apart (fill with `pattern`, ids 10–18), then:
    #   atmosphere = PrescribedAtmosphere(atmos_grid, times; ...)
    #   model = OceanSeaIceModel(ocean; atmosphere, radiation)
    #   model.clock.time = 1.37 * 3hours; update_state!(model)
    # and dump: parent(model.interfaces.net_fluxes.ocean.{u,v,T,S}), atmosphere_ocean_interface.fluxes.*, exchange state.
    error("fill in the PrescribedAtmosphere construction for the installed NumericalEarth version, then remove this line")
end

abspath(PROGRAM_FILE) == @__FILE__ && main()
