# dump_reference.jl — golden-vector dumper for someone with Julia + network access.  UNEXECUTED in this repository's build
# environment (no Julia toolchain; the arithmetic lives in the un-vendored NumericalEarth.jl, /root/reference/Project.toml:21,31-32).
#
# Builds BASELINE config 1 (64×32×8 LatitudeLongitudeGrid, synthetic PrescribedAtmosphere) with the real ClimaOcean /
# NumericalEarth stack on the CPU, fills every input with the synthetic patterns of climaocean.jl_b200/synth.py
# (julia/synthetic_inputs.jl: bit-identical across languages), runs the REAL update_state! at t = 1.37·3 h, and writes every
# output array as raw little-endian Float64 under tests/golden/julia_c1/<name>.bin plus manifest.txt (name nx ny).
# tests/test_julia_golden.py compares the CUDA path against them at 1e-12 when they are present — the step that turns
# "parity unpinned" into a pinned parity claim.
#
#     julia --project=/path/to/ClimaOcean.jl julia/dump_reference.jl [default|corrected|ncar] [outdir]
include(joinpath(@__DIR__, "synthetic_inputs.jl"))
using Oceananigans.TimeSteppers: update_state!

function dump(name, field, outdir, manifest)
    p = Array(parent(field))[:, :, 1]                      # (Nx + 2H, Ny + 2H): i fastest, as the C ABI sees it
    write(joinpath(outdir, name * ".bin"), htol.(Float64.(p)))
    println(manifest, name, " ", size(p, 1), " ", size(p, 2))
end

function main(flux_configuration = :default, outdir = joinpath(@__DIR__, "..", "tests", "golden", "julia_c1"))
    mkpath(outdir)
    model = synthetic_coupled_model(CPU(), 64, 32, 8; latitude = (-60, 60), flux_configuration)
    model.clock.time = 1.37 * 3hours                       # exercises the time weights (SURVEY §8d)
    update_state!(model)
    itf = model.interfaces
    open(joinpath(outdir, "manifest.txt"), "w") do manifest
        net = itf.net_fluxes.ocean                          # omip_diagnostics.jl:77-80
        dump("net_u", net.u, outdir, manifest); dump("net_v", net.v, outdir, manifest)
        dump("net_T", net.T, outdir, manifest); dump("net_S", net.S, outdir, manifest)
        ao = itf.atmosphere_ocean_interface.fluxes           # omip_diagnostics.jl:81-82
        dump("latent_heat", ao.latent_heat, outdir, manifest); dump("sensible_heat", ao.sensible_heat, outdir, manifest)
        dump("water_vapor", ao.water_vapor, outdir, manifest)
        dump("x_momentum", ao.x_momentum, outdir, manifest); dump("y_momentum", ao.y_momentum, outdir, manifest)
        x = itf.exchanger.exchange_atmosphere_state          # interpolated atmosphere on the ocean grid
        for n in (:u, :v, :T, :q, :p, :Qs, :Qℓ, :Mp)
            hasproperty(x, n) && dump("exchange_" * replace(String(n), "ℓ" => "l"), getproperty(x, n), outdir, manifest)
        end
    end
    @info "reference dump written" outdir flux_configuration
end

if abspath(PROGRAM_FILE) == @__FILE__
    cfg = length(ARGS) >= 1 ? Symbol(ARGS[1]) : :default
    length(ARGS) >= 2 ? main(cfg, ARGS[2]) : main(cfg)
end
