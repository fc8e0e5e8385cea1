# CoFluxExt — Julia glue that routes ClimaOcean's surface-flux path through libcoflux.so.
#
# !!! UNEXECUTED !!!  No Julia toolchain exists in the build environment or on the GPU box this
# repository is developed on.  This file is the binding a maintainer would add; it has never been
# run.  Layouts are checked at load time against `coflux_sizeof`, so a drifted mirror fails loudly.
#
# Reference call sites this overrides (ClimaOcean v0.10.0 re-exports NumericalEarth,
# src/ClimaOcean.jl:31-42): `update_state!(::OceanSeaIceModel)` and, through it,
# interpolate_atmosphere_state!, compute_atmosphere_ocean_fluxes!, compute_sea_ice_ocean_fluxes!,
# compute_net_ocean_fluxes!  (SURVEY.md §3.2).
module CoFluxExt

using CUDA
using Oceananigans
using Oceananigans.Architectures: architecture, GPU
using NumericalEarth.EarthSystemModels: OceanSeaIceModel
import NumericalEarth.EarthSystemModels: update_state!

const libcoflux = get(ENV, "COFLUX_LIB", joinpath(@__DIR__, "..", "..", "..", "climaocean.jl_b200", "lib", "libcoflux.so"))

# ---- struct mirrors of include/coflux.h -------------------------------------------------------
struct CofluxArray
    ptr      :: Ptr{Cvoid}
    stride_i :: Int64
    stride_j :: Int64
    stride_k :: Int64
    stride_n :: Int64
    off_i    :: Int32
    off_j    :: Int32
    off_k    :: Int32
    reserved :: Int32
end
const NULL_ARRAY = CofluxArray(C_NULL, 0, 0, 0, 0, 0, 0, 0, 0)

"Descriptor of an Oceananigans Field's parent (halo-padded, column-major)."
function CofluxArray(f::Oceananigans.Fields.AbstractField)
    p  = parent(f)
    Hx, Hy, Hz = Oceananigans.Grids.halo_size(f.grid)
    sz = size(p)
    hz = sz[3] == 1 ? 0 : Hz                      # reduced (2-D) fields have a singleton k, no halo
    return CofluxArray(Ptr{Cvoid}(UInt(pointer(p))), 1, sz[1], sz[1] * sz[2], 0, Hx, Hy, hz, 0)
end

"Descriptor of a FieldTimeSeries window in memory (4-D parent, time last)."
function CofluxArray(fts::Oceananigans.OutputReaders.FieldTimeSeries)
    p  = parent(fts)
    Hx, Hy, _ = Oceananigans.Grids.halo_size(fts.grid)
    sz = size(p)
    return CofluxArray(Ptr{Cvoid}(UInt(pointer(p))), 1, sz[1], sz[1] * sz[2], sz[1] * sz[2] * sz[3], Hx, Hy, 0, 0)
end

# coflux_config is large; it is filled by the library (coflux_default_config +
# coflux_apply_flux_configuration) into an opaque, correctly sized buffer and then edited through
# the handful of scalars ClimaOcean exposes (ocean_minimum_salinity, velocity formulation ...).
struct CofluxConfigBuffer
    bytes :: Vector{UInt8}
end
function CofluxConfigBuffer(Nx, Ny, Nz, FT; flux_configuration = :default, velocity_formulation = :relative)
    n = ccall((:coflux_sizeof, libcoflux), Cint, (Cstring,), "config")
    n > 0 || error("libcoflux does not know struct coflux_config")
    buf = CofluxConfigBuffer(zeros(UInt8, n))
    dtype = FT === Float64 ? 64 : 32
    check(ccall((:coflux_default_config, libcoflux), Cint, (Ptr{UInt8}, Int32, Int32, Int32, Int32), buf.bytes, Nx, Ny, Nz, dtype))
    vel = velocity_formulation === :relative ? 0 :
          velocity_formulation === :wind     ? 1 :
          error("Unknown velocity_formulation: $velocity_formulation. Options: :relative, :wind")
    check(ccall((:coflux_apply_flux_configuration, libcoflux), Cint, (Ptr{UInt8}, Cstring, Int32), buf.bytes, String(flux_configuration), vel))
    return buf
end

check(status) = status == 0 ? nothing :
    error("coflux status $status: " * unsafe_string(ccall((:coflux_last_error, libcoflux), Cstring, ())))

mutable struct CofluxContext
    handle :: Ptr{Cvoid}
end
function CofluxContext(cfg::CofluxConfigBuffer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:coflux_create, libcoflux), Cint, (Ref{Ptr{Cvoid}}, Ptr{UInt8}), h, cfg.bytes))
    ctx = CofluxContext(h[])
    finalizer(c -> ccall((:coflux_destroy, libcoflux), Cint, (Ptr{Cvoid},), c.handle), ctx)
    return ctx
end

# ---- bundles (field order as in include/coflux.h) --------------------------------------------
struct AtmosSeries
    u::CofluxArray; v::CofluxArray; T::CofluxArray; q::CofluxArray; p::CofluxArray
    Qs::CofluxArray; Ql::CofluxArray; rain::CofluxArray; snow::CofluxArray
    times::Ptr{Float64}; Nt::Int32; time_indexing::Int32; cycle_period::Float64
    fi::CofluxArray; fj::CofluxArray; cos_theta::CofluxArray; sin_theta::CofluxArray
    ring_start::Int32; ring_capacity::Int32          # ABI 2: device ring buffer (coflux_forcing_window); 0, 0 = plain layout
end
struct LandSeries                                     # ABI 2: JRA55PrescribedLand (friver + licalvf → exchange Mp)
    rivers::CofluxArray; icebergs::CofluxArray
    times::Ptr{Float64}; Nt::Int32; time_indexing::Int32; cycle_period::Float64
    fi::CofluxArray; fj::CofluxArray
    ring_start::Int32; ring_capacity::Int32
end
struct ExchangeState;   u::CofluxArray; v::CofluxArray; T::CofluxArray; p::CofluxArray; q::CofluxArray; Qs::CofluxArray; Ql::CofluxArray; Mp::CofluxArray; end
struct OceanSurface;    u::CofluxArray; v::CofluxArray; T::CofluxArray; S::CofluxArray; mask::CofluxArray; end
struct InterfaceFluxes; latent_heat::CofluxArray; sensible_heat::CofluxArray; water_vapor::CofluxArray; x_momentum::CofluxArray; y_momentum::CofluxArray
                        interface_temperature::CofluxArray; friction_velocity::CofluxArray; temperature_scale::CofluxArray; humidity_scale::CofluxArray; iterations::CofluxArray; end
struct NetOceanFluxes;  u::CofluxArray; v::CofluxArray; T::CofluxArray; S::CofluxArray; upwelling_longwave::CofluxArray; downwelling_longwave::CofluxArray
                        downwelling_shortwave::CofluxArray; penetrating_shortwave::CofluxArray; end
struct SeaIceState;     thickness::CofluxArray; previous_thickness::CofluxArray; concentration::CofluxArray; salinity::CofluxArray
                        u::CofluxArray; v::CofluxArray; top_temperature::CofluxArray; snow_thickness::CofluxArray; albedo::CofluxArray; end
struct IceOceanFluxes;  frazil_heat::CofluxArray; interface_heat::CofluxArray; salt::CofluxArray; x_momentum::CofluxArray; y_momentum::CofluxArray; end
struct NetSeaIceFluxes; top_heat::CofluxArray; bottom_heat::CofluxArray; top_u::CofluxArray; top_v::CofluxArray; end        # ABI 2
struct FluxAverages                                   # ABI 2: WindowedTimeAverage accumulators (omip_diagnostics.jl:125-158)
    tau_x::CofluxArray; tau_y::CofluxArray; JT::CofluxArray; JS::CofluxArray; Qc::CofluxArray; Qv::CofluxArray
    JT_atmosphere_ocean::CofluxArray; JT_ice_ocean::CofluxArray; JS_ice_ocean::CofluxArray; JT_frazil::CofluxArray
    previous_interval::Float64; dt::Float64
end
struct UpdateInputs;    atmosphere::Ptr{AtmosSeries}; ocean::Ptr{OceanSurface}; sea_ice::Ptr{Cvoid}; ice_ocean::Ptr{Cvoid}; land::Ptr{LandSeries}; end
struct UpdateOutputs;   exchange::Ptr{ExchangeState}; atmosphere_ocean::Ptr{InterfaceFluxes}; net_ocean::Ptr{NetOceanFluxes}; end

function __init__()
    for (name, T) in (("array", CofluxArray), ("atmos_series", AtmosSeries), ("exchange_state", ExchangeState),
                      ("ocean_surface", OceanSurface), ("interface_fluxes", InterfaceFluxes),
                      ("net_ocean_fluxes", NetOceanFluxes), ("update_inputs", UpdateInputs), ("update_outputs", UpdateOutputs),
                      ("land_series", LandSeries), ("sea_ice_state", SeaIceState), ("ice_ocean_fluxes", IceOceanFluxes),
                      ("net_sea_ice_fluxes", NetSeaIceFluxes), ("flux_averages", FluxAverages))
        n = ccall((:coflux_sizeof, libcoflux), Cint, (Cstring,), name)
        n == sizeof(T) || error("CoFluxExt: layout of $name drifted (Julia $(sizeof(T)) B, library $n B)")
    end
end

# A per-model cache: context + the construction-time fractional indices (fi, fj) Fields.
const CONTEXTS = IdDict{Any, Any}()

"""
    update_state!(model::OceanSeaIceModel{<:GPU ...})

Ocean-only coupled model on a GPU: one `coflux_update_state` call (2 kernel launches) on CUDA.jl's
task-local stream replaces interpolate_atmosphere_state! + compute_atmosphere_ocean_fluxes! +
compute_net_ocean_fluxes!.  All arrays stay Julia-owned `Field`s; nothing is copied.
(The exact field paths into `model.interfaces` must be adapted to the installed NumericalEarth version.)
"""
function coflux_update_state!(model::OceanSeaIceModel)
    ocean, atmos, itf = model.ocean, model.atmosphere, model.interfaces
    grid = ocean.model.grid
    ctx, fi, fj = get!(CONTEXTS, model) do
        Nx, Ny, Nz = size(grid)
        cfg = CofluxConfigBuffer(Nx, Ny, Nz, eltype(grid))
        CofluxContext(cfg), itf.exchanger.regridder.i, itf.exchanger.regridder.j    # construction-time fractional indices
    end
    times = collect(Float64, atmos.times)
    u, v = ocean.model.velocities.u, ocean.model.velocities.v
    T, S = ocean.model.tracers.T, ocean.model.tracers.S
    series = Ref(AtmosSeries(CofluxArray(atmos.velocities.u), CofluxArray(atmos.velocities.v), CofluxArray(atmos.tracers.T),
                             CofluxArray(atmos.tracers.q), CofluxArray(atmos.pressure),
                             CofluxArray(model.radiation.downwelling_shortwave), CofluxArray(model.radiation.downwelling_longwave),
                             CofluxArray(atmos.freshwater_flux.rain), CofluxArray(atmos.freshwater_flux.snow),
                             pointer(times), length(times), 0, 0.0, CofluxArray(fi), CofluxArray(fj), NULL_ARRAY, NULL_ARRAY, Int32(0), Int32(0)))
    xs = itf.exchanger.exchange_atmosphere_state
    xch = Ref(ExchangeState(CofluxArray(xs.u), CofluxArray(xs.v), CofluxArray(xs.T), CofluxArray(xs.p), CofluxArray(xs.q),
                            CofluxArray(xs.Qs), CofluxArray(xs.Qℓ), CofluxArray(xs.Mp)))
    oc = Ref(OceanSurface(CofluxArray(u), CofluxArray(v), CofluxArray(T), CofluxArray(S), NULL_ARRAY))
    f = itf.atmosphere_ocean_interface.fluxes
    ao = Ref(InterfaceFluxes(CofluxArray(f.latent_heat), CofluxArray(f.sensible_heat), CofluxArray(f.water_vapor),
                             CofluxArray(f.x_momentum), CofluxArray(f.y_momentum),
                             CofluxArray(itf.atmosphere_ocean_interface.temperature), NULL_ARRAY, NULL_ARRAY, NULL_ARRAY, NULL_ARRAY))
    n = itf.net_fluxes.ocean
    net = Ref(NetOceanFluxes(CofluxArray(n.u), CofluxArray(n.v), CofluxArray(n.T), CofluxArray(n.S),
                             CofluxArray(f.upwelling_longwave), CofluxArray(f.downwelling_longwave), CofluxArray(f.downwelling_shortwave),
                             NULL_ARRAY))
    GC.@preserve times series xch oc ao net begin
        inp = Ref(UpdateInputs(Base.unsafe_convert(Ptr{AtmosSeries}, series), Base.unsafe_convert(Ptr{OceanSurface}, oc), C_NULL, C_NULL, C_NULL))
        out = Ref(UpdateOutputs(Base.unsafe_convert(Ptr{ExchangeState}, xch), Base.unsafe_convert(Ptr{InterfaceFluxes}, ao),
                                Base.unsafe_convert(Ptr{NetOceanFluxes}, net)))
        check(ccall((:coflux_update_state, libcoflux), Cint,
                    (Ptr{Cvoid}, Ref{UpdateInputs}, Ref{UpdateOutputs}, Float64, Ptr{Cvoid}),
                    ctx.handle, inp, out, Float64(model.clock.time), Ptr{Cvoid}(UInt(CUDA.stream().handle))))
    end
    return nothing
end

# Opt in per model:  CoFluxExt.enable!(model)  makes update_state!(model) take the coflux path.
const ENABLED = IdDict{Any, Bool}()
enable!(model)  = (ENABLED[model] = true; nothing)
disable!(model) = (delete!(ENABLED, model); nothing)

function update_state!(model::OceanSeaIceModel)
    if get(ENABLED, model, false) && architecture(model.ocean.model.grid) isa GPU && isnothing(model.sea_ice)
        return coflux_update_state!(model)
    end
    return invoke(update_state!, Tuple{Any}, model)     # the stock NumericalEarth method
end

# NormalizeSalinity (src/OMIPConfigurations/omip_simulation.jl:187-220) through coflux: the callable keeps its
# fields; only `compute!(mean_total)` + `parent(flux_field) .-= mean_total` are replaced.  UNEXECUTED like the rest.
struct SalinityNormalization; flux::CofluxArray; additional::CofluxArray; area::CofluxArray; mask::CofluxArray; end

function normalize_salinity!(n, sim; area::CofluxArray, mask::CofluxArray = NULL_ARRAY)
    model = sim.model.ocean.model
    ctx = CONTEXTS[sim.model][1]            # created by the first coflux_update_state!(model)
    if !isnothing(n.additional_fluxes)      # stock materialisation of the additional flux (omip_simulation.jl:210-216)
        grid = model.grid
        fields = merge(model.velocities, model.tracers)
        launch!(architecture(grid), grid, :xy, ClimaOcean.OMIPConfigurations._materialize_top_flux!,
                n.additional_buffer, n.additional_fluxes, grid, model.clock, fields)
    end
    add = isnothing(n.additional_buffer) ? NULL_ARRAY : CofluxArray(n.additional_buffer)
    norm = Ref(SalinityNormalization(CofluxArray(n.flux_field), add, area, mask))
    check(ccall((:coflux_normalize_salinity_flux, libcoflux), Cint, (Ptr{Cvoid}, Ref{SalinityNormalization}, Ptr{Cvoid}),
                ctx.handle, norm, Ptr{Cvoid}(UInt(CUDA.stream().handle))))
    # distributed: coflux_salinity_flux_sums → MPI.Allreduce!(sums, +, comm) → coflux_subtract_mean_flux
    return nothing
end


# compute_net_sea_ice_fluxes! (ABI 2) — what the coupled model hands to ClimaSeaIce after the flux solves.  UNEXECUTED.
function compute_net_sea_ice_fluxes!(ctx::CofluxContext, xch::Ref{ExchangeState}, oc::Ref{OceanSurface}, ice::Ref{SeaIceState},
                                     ai::Ref{InterfaceFluxes}, io::Ref{IceOceanFluxes}, net::Ref{NetSeaIceFluxes})
    check(ccall((:coflux_assemble_net_sea_ice_fluxes, libcoflux), Cint,
                (Ptr{Cvoid}, Ref{ExchangeState}, Ref{OceanSurface}, Ref{SeaIceState}, Ref{InterfaceFluxes}, Ref{IceOceanFluxes},
                 Ref{NetSeaIceFluxes}, Ptr{Cvoid}),
                ctx.handle, xch, oc, ice, ai, io, net, Ptr{Cvoid}(UInt(CUDA.stream().handle))))
end

# Time-averaged flux outputs (ABI 2): attach once per collection with the seconds already in the window and this Δt;
# the next update_state! / compute_sea_ice_ocean_fluxes! update the running means in their epilogues.  UNEXECUTED.
attach_flux_averages!(ctx::CofluxContext, avg::Ref{FluxAverages}) =
    check(ccall((:coflux_attach_flux_averages, libcoflux), Cint, (Ptr{Cvoid}, Ref{FluxAverages}), ctx.handle, avg))
detach_flux_averages!(ctx::CofluxContext) =
    check(ccall((:coflux_attach_flux_averages, libcoflux), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx.handle, C_NULL))

# Device forcing window (ABI 2): `time_indices_in_memory` levels of every series in a device ring, uploads on the
# window's own copy stream, ordered against the compute stream by events.  UNEXECUTED.
mutable struct ForcingWindow
    handle :: Ptr{Cvoid}
end
function ForcingWindow(ctx::CofluxContext, n_fields::Integer, plane_elements::Integer, capacity::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:coflux_forcing_window_create, libcoflux), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cvoid}, Int32, Int64, Int32),
                h, ctx.handle, n_fields, plane_elements, capacity))
    w = ForcingWindow(h[])
    finalizer(x -> ccall((:coflux_forcing_window_destroy, libcoflux), Cint, (Ptr{Cvoid},), x.handle), w)
    return w
end
upload!(w::ForcingWindow, level::Integer, host_planes::Vector{Ptr{Cvoid}}) =      # pinned host planes, one per field
    check(ccall((:coflux_forcing_window_upload, libcoflux), Cint, (Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}), w.handle, level, host_planes))
wait_levels!(w::ForcingWindow, first::Integer, last::Integer) =
    check(ccall((:coflux_forcing_window_wait, libcoflux), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}),
                w.handle, first, last, Ptr{Cvoid}(UInt(CUDA.stream().handle))))
release_levels!(w::ForcingWindow, first::Integer, last::Integer) =
    check(ccall((:coflux_forcing_window_release, libcoflux), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}),
                w.handle, first, last, Ptr{Cvoid}(UInt(CUDA.stream().handle))))
function field_pointer(w::ForcingWindow, field::Integer)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:coflux_forcing_window_field, libcoflux), Cint, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), w.handle, field, p))
    return p[]        # capacity × plane_elements elements; describe it with CofluxArray(...; stride_n = plane_elements) and ring_start / ring_capacity
end

end # module
