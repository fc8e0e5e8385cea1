# synthetic_inputs.jl — the synthetic inputs of SURVEY.md §8d in Julia, bit-identical to climaocean.jl_b200/synth.py
# (SplitMix64 + triangle waves: integer hashing and IEEE basic operations only, no libm).
# UNEXECUTED here (no Julia in the build image); shared by dump_reference.jl and time_reference_cpu.jl.
using ClimaOcean, Oceananigans
using Oceananigans.Units
using Oceananigans.Fields: interior
using Oceananigans.OutputReaders: FieldTimeSeries

const SEED_BASE = 0xC0F10000
@inline function splitmix(x::UInt64)
    z = x
    z = (z ⊻ (z >> 30)) * 0xBF58476D1CE4E5B9
    z = (z ⊻ (z >> 27)) * 0x94D049BB133111EB
    return z ⊻ (z >> 31)
end
u01(field_id, k) = Float64(splitmix(UInt64(SEED_BASE + field_id) + (UInt64(k) + 1) * 0x9E3779B97F4A7C15) >> 11) * 2.0^-53
tri(m, N) = abs(2.0 * (mod(m, N) / N) - 1.0)

"value of synthetic field `field_id` at zero-based global (i, j) (i periodic over Nxg), time / depth `level`"
function pattern(field_id, lo, hi, i, j, Nxg, Nyg; level = 0, w = 0.5)
    I = mod(i, Nxg); J = j
    k = (J + 64) * Nxg + I + level * Nxg * (Nyg + 128)
    p1, p2 = 1 + field_id % 3, 1 + (field_id ÷ 3) % 2
    s = 0.5 * tri(I * p1 + (field_id * 37) % Nxg, Nxg) + 0.5 * tri((J + 64) * p2 + (field_id * 11) % Nyg, Nyg)
    return lo + (hi - lo) * (w * u01(field_id, k) + (1 - w) * s)
end

const ATM_IDS    = (u = 10, v = 11, T = 12, q = 13, p = 14, Qs = 15, Ql = 16, rain = 17, snow = 18)
const ATM_RANGES = (u = (-25.0, 25.0), v = (-25.0, 25.0), T = (250.0, 305.0), q = (1e-4, 2e-2), p = (9.6e4, 1.04e5),
                    Qs = (0.0, 1000.0), Ql = (100.0, 450.0), rain = (0.0, 3e-4), snow = (0.0, 3e-4))

"fill the parent (halos included) of an ocean prognostic field; surface value relaxed with depth as in synth.ocean_state"
function fill_ocean_field!(f, id, lo, hi, Nx, Ny, Nz, H; is_T = false, is_velocity = false)
    p = parent(f)
    for kk in axes(p, 3), jj in axes(p, 2), ii in axes(p, 1)
        k = kk - 1 - H
        lev = clamp(k, 0, Nz - 1)
        depth = k >= Nz - 1 ? 0.0 : (Nz - 1 - max(k, 0)) / max(Nz - 1, 1)
        x = pattern(id, lo, hi, ii - 1 - H, jj - 1 - H, Nx, Ny; level = lev)
        is_T && (x = x - depth * (x + 1.0) * 0.9)
        is_velocity && (x = x * (1.0 - 0.8 * depth))
        p[ii, jj, kk] = x
    end
    return f
end

"ocean_simulation on a flat-bottom LatitudeLongitudeGrid with the synthetic u, v, T, S"
function synthetic_ocean(arch, Nx, Ny, Nz; latitude = (-60, 60), H = 7)
    grid = LatitudeLongitudeGrid(arch; size = (Nx, Ny, Nz), longitude = (0, 360), latitude, z = (-5000, 0), halo = (H, H, H))
    ocean = ocean_simulation(grid)
    u, v = ocean.model.velocities.u, ocean.model.velocities.v
    T, S = ocean.model.tracers.T, ocean.model.tracers.S
    fill_ocean_field!(u, 1, -1.0, 1.0, Nx, Ny, Nz, H; is_velocity = true)
    fill_ocean_field!(v, 2, -1.0, 1.0, Nx, Ny, Nz, H; is_velocity = true)
    fill_ocean_field!(T, 3, -1.8, 30.0, Nx, Ny, Nz, H; is_T = true)
    fill_ocean_field!(S, 4, 30.0, 38.0, Nx, Ny, Nz, H)
    return ocean
end

"PrescribedAtmosphere (+ downwelling radiation) on a regular 640×320 source grid, Nt levels 3 h apart"
function synthetic_atmosphere(arch; Nxa = 640, Nya = 320, Nt = 8, Δt = 3hours)
    agrid = LatitudeLongitudeGrid(arch; size = (Nxa, Nya), longitude = (0, 360), latitude = (-90, 90), topology = (Periodic, Bounded, Flat))
    times = collect(range(0.0, step = Float64(Δt), length = Nt))
    atmosphere = PrescribedAtmosphere(agrid, times; surface_layer_height = 10, boundary_layer_height = 512)
    targets = (u = atmosphere.velocities.u, v = atmosphere.velocities.v, T = atmosphere.tracers.T, q = atmosphere.tracers.q,
               p = atmosphere.pressure, rain = atmosphere.freshwater_flux.rain, snow = atmosphere.freshwater_flux.snow)
    # radiation lives in the atmosphere (ClimaOcean ≤ 0.8) or in a separate PrescribedRadiation (this snapshot, atmosphere.jl:39-44)
    radiation = nothing
    if hasproperty(atmosphere, :downwelling_radiation)
        targets = merge(targets, (Qs = atmosphere.downwelling_radiation.shortwave, Ql = atmosphere.downwelling_radiation.longwave))
    else
        radiation = PrescribedRadiation(agrid, times; ocean_surface = SurfaceRadiationProperties(0.06, 1.0))
        targets = merge(targets, (Qs = radiation.downwelling_shortwave, Ql = radiation.downwelling_longwave))
    end
    for name in keys(targets)
        fts = getproperty(targets, name)
        id = getproperty(ATM_IDS, name); lo, hi = getproperty(ATM_RANGES, name)
        for n in 1:Nt
            data = interior(fts[n])
            for j in 1:Nya, i in 1:Nxa
                data[i, j, 1] = pattern(id, lo, hi, i - 1, j - 1, Nxa, Nya; level = n - 1)
            end
        end
        Oceananigans.BoundaryConditions.fill_halo_regions!(fts)
    end
    return atmosphere, radiation
end

"OceanSeaIceModel(ocean; atmosphere[, radiation]) — constructor call sites: README.md:74-75, examples/one_degree_tripolar_ocean_sea_ice.jl:42-51"
function synthetic_coupled_model(arch, Nx, Ny, Nz; latitude = (-60, 60), flux_configuration = :default)
    ocean = synthetic_ocean(arch, Nx, Ny, Nz; latitude)
    atmosphere, radiation = synthetic_atmosphere(arch)
    if flux_configuration == :default
        model = isnothing(radiation) ? OceanSeaIceModel(ocean; atmosphere) : OceanSeaIceModel(ocean; atmosphere, radiation)
    else   # the OMIP presets of src/OMIPConfigurations/omip_simulation.jl:123-164 (sea_ice = nothing: ocean only)
        ao = flux_configuration == :corrected ? ClimaOcean.OMIPConfigurations.corrected_atmosphere_ocean_fluxes(Float64) :
                                                ClimaOcean.OMIPConfigurations.ncar_atmosphere_ocean_fluxes(Float64)
        interfaces = ComponentInterfaces(atmosphere, ocean; radiation, atmosphere_ocean_fluxes = ao)
        model = OceanSeaIceModel(ocean; atmosphere, interfaces)
    end
    return model
end
