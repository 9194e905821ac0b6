# BreezeB200Ext.jl — the package extension a Breeze.jl maintainer would add to dispatch the hot paths to libbreeze_b200.so.
#
# NOT EXECUTED IN THIS PROJECT: Julia is not installed in the authoring container or on the GPU boxes (DESIGN.md §1). The same C ABI
# (include/breeze_b200.h, include/breeze_b200_compressible.h) is exercised from Python/ctypes by the parity tests. This file is the
# binding of INTEGRATION.md kept as source so that it can be dropped into ext/BreezeB200Ext/ of the reference repository.
#
# Hooks overridden (reference file:line):
#   time_step!(model::AtmosphereModel{<:AnelasticDynamics,…,<:SSPRungeKutta3}, Δt)     src/TimeSteppers/ssp_runge_kutta_3.jl:209
#   time_step!(model::CompressibleAcousticModel, Δt)                                   src/TimeSteppers/acoustic_runge_kutta_3.jl:264
# following the precedent of ext/BreezeReactantExt/Timesteppers.jl:6-19 (dispatch on the architecture type parameter).

module BreezeB200Ext

using Breeze, Oceananigans
using Breeze.AtmosphereModels: AtmosphereModel, prognostic_fields
using Breeze.AnelasticEquations: AnelasticDynamics
using Breeze.TimeSteppers: SSPRungeKutta3, AcousticRungeKutta3
using Breeze.CompressibleEquations: CompressibleDynamics, ThermalDivergenceDamping
import Oceananigans.TimeSteppers: time_step!, update_state!
import Oceananigans.Fields: set!

const lib = "libbreeze_b200"           # on LD_LIBRARY_PATH / via Libdl.dlopen

# mirrors `struct bz_config` of include/breeze_b200.h field for field
Base.@kwdef mutable struct BzConfig
    abi_version::Int32 = 1
    Nx::Int32 = 8; Ny::Int32 = 8; Nz::Int32 = 8
    topology_x::Int32 = 0; topology_y::Int32 = 0
    x0::Float64 = 0; x1::Float64 = 1; y0::Float64 = 0; y1::Float64 = 1; z0::Float64 = 0; z1::Float64 = 1
    surface_pressure::Float64 = 101325; potential_temperature::Float64 = 288; standard_pressure::Float64 = 1e5
    molar_gas_constant::Float64 = 8.314462618; gravitational_acceleration::Float64 = 9.81
    energy_reference_temperature::Float64 = 273.15; triple_point_temperature::Float64 = 273.16
    triple_point_pressure::Float64 = 611.657; dry_air_molar_mass::Float64 = 0.02897; dry_air_heat_capacity::Float64 = 1005
    vapor_molar_mass::Float64 = 0.018015; vapor_heat_capacity::Float64 = 1850
    liquid_reference_latent_heat::Float64 = 2500800; liquid_heat_capacity::Float64 = 4181
    ice_reference_latent_heat::Float64 = 2834000; ice_heat_capacity::Float64 = 2108
    advection_order::Int32 = 5; microphysics::Int32 = 0
    n_ranks::Int32 = 1; rank::Int32 = 0; device::Int32 = 0; reserved0::Int32 = 0
    nccl_unique_id::NTuple{128,UInt8} = ntuple(_ -> 0x00, 128)
    use_tma::Int32 = 0; z_chunks::Int32 = 0; formulation::Int32 = 0; reserved::NTuple{5,Int32} = ntuple(_ -> Int32(0), 5)
end

struct B200 <: Oceananigans.Architectures.AbstractArchitecture   # the architecture the hooks dispatch on
    device::Int
end

mutable struct B200Context
    handle::Ptr{Cvoid}
end

check(rc, ctx) = rc == 0 || error(unsafe_string(ccall((:bz_last_error, lib), Cstring, (Ptr{Cvoid},), ctx)))

# grid extents, topology and thermodynamic constants common to both paths
function grid_config(model; kw...)
    grid, c = model.grid, model.thermodynamic_constants
    TX, TY, _ = Oceananigans.Grids.topology(grid)
    return BzConfig(; Nx = grid.Nx, Ny = grid.Ny, Nz = grid.Nz,
                    topology_x = TX === Oceananigans.Grids.Flat ? 1 : 0, topology_y = TY === Oceananigans.Grids.Flat ? 1 : 0,
                    x0 = grid.xᶠᵃᵃ[1], x1 = grid.xᶠᵃᵃ[grid.Nx + 1], y0 = grid.yᵃᶠᵃ[1], y1 = grid.yᵃᶠᵃ[grid.Ny + 1],
                    z0 = grid.z.cᵃᵃᶠ[1], z1 = grid.z.cᵃᵃᶠ[grid.Nz + 1],
                    molar_gas_constant = c.molar_gas_constant, gravitational_acceleration = c.gravitational_acceleration,
                    dry_air_molar_mass = c.dry_air.molar_mass, dry_air_heat_capacity = c.dry_air.heat_capacity,
                    vapor_molar_mass = c.vapor.molar_mass, vapor_heat_capacity = c.vapor.heat_capacity,
                    device = model.grid.architecture.device, kw...)
end

function B200Context(model)
    ref = model.dynamics.reference_state
    cfg = grid_config(model; surface_pressure = ref.surface_pressure, potential_temperature = ref.potential_temperature,
                      standard_pressure = ref.standard_pressure,
                      formulation = model.formulation isa Breeze.StaticEnergyFormulation ? 1 : 0)
    # WENO(order = 5 | 7 | 9) (examples/dry_thermal_bubble.jl:24, examples/bomex.jl:204): 2 * buffer - 1 of the momentum scheme
    cfg.advection_order = 2 * Oceananigans.Advection.required_halo_size_x(model.advection.momentum) - 1
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:bz_create, lib), Cint, (Ref{BzConfig}, Ref{Ptr{Cvoid}}), cfg, h)
    rc == 0 || error(unsafe_string(ccall((:bz_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL)))
    ctx = B200Context(h[])
    finalizer(c -> ccall((:bz_destroy, lib), Cvoid, (Ptr{Cvoid},), c.handle), ctx)
    return ctx
end

const contexts = IdDict{Any,B200Context}()
context(model) = get!(() -> B200Context(model), contexts, model)

const B200AnelasticModel = AtmosphereModel{<:AnelasticDynamics, <:Any, <:B200, <:SSPRungeKutta3}

# set!(model; ...): let Breeze fill its host fields, then push the prognostics and run the projection on the device
function push_state!(model::B200AnelasticModel; enforce_mass_conservation = true)
    p = prognostic_fields(model)
    arrays = map(f -> Array(interior(f)), (p.ρu, p.ρv, p.ρw, p.ρθ, p.ρqᵛ))
    ctx = context(model)
    check(ccall((:bz_set_state, lib), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint),
                ctx.handle, arrays..., enforce_mass_conservation), ctx.handle)
end

# time_step!(model, Δt): src/TimeSteppers/ssp_runge_kutta_3.jl:209
function time_step!(model::B200AnelasticModel, Δt; callbacks = [])
    ctx = context(model)
    check(ccall((:bz_time_step, lib), Cint, (Ptr{Cvoid}, Cdouble), ctx.handle, Δt), ctx.handle)
    model.clock.time += Δt; model.clock.iteration += 1
    return nothing
end

# NaNChecker callback of run! (atmosphere_model.jl:561-572): a device-side reduction instead of copying ρu to the host
function state_is_finite(model::B200AnelasticModel)
    ctx = context(model); ok = Ref{Cint}(0)
    check(ccall((:bz_state_is_finite, lib), Cint, (Ptr{Cvoid}, Ref{Cint}), ctx.handle, ok), ctx.handle)
    return ok[] == 1
end

# one x-z / x-y / y-z slice for an output writer (axis 0: x = index, 1: y = index, 2: z = index; 0-based)
function pull_slice(model::B200AnelasticModel, id::Integer, axis::Integer, index::Integer, dims)
    a = Array{Float64}(undef, dims)
    ctx = context(model)
    check(ccall((:bz_get_slice, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float64}), ctx.handle, id, axis, index, a), ctx.handle)
    return a
end

# host-resident driver loop: buffers in and out every step without blocking (pinned arrays; bz_synchronize before touching them)
function step_with_host_buffers!(model::B200AnelasticModel, Δt, inputs::NTuple{5,Array{Float64,3}}, outputs::NTuple{5,Array{Float64,3}})
    ctx = context(model)
    check(ccall((:bz_set_state_async, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint),
                ctx.handle, inputs..., 0), ctx.handle)
    check(ccall((:bz_time_step, lib), Cint, (Ptr{Cvoid}, Cdouble), ctx.handle, Δt), ctx.handle)
    check(ccall((:bz_get_state_async, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                ctx.handle, outputs...), ctx.handle)
end

# pull a field back when an output writer / diagnostic needs it
function pull!(field, model::B200AnelasticModel, id::Integer)      # id = BZ_RHO_U … BZ_QL
    a = Array{Float64}(undef, size(interior(field)))
    ctx = context(model)
    check(ccall((:bz_get_field, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), ctx.handle, id, a), ctx.handle)
    copyto!(interior(field), a)
end

#####
##### Compressible split-explicit dynamics (include/breeze_b200_compressible.h)
#####

# mirrors `struct bzc_config`: the BzConfig above followed by the SplitExplicitTimeDiscretization / CompressibleDynamics knobs
Base.@kwdef mutable struct BzcConfig
    base::BzConfig = BzConfig()
    reference_state::Int32 = 1; substeps::Int32 = 0; damping::Int32 = 1; substep_distribution::Int32 = 0
    apply_first_substep_pressure_gradient::Int32 = 0; damp_vertical::Int32 = 0
    acoustic_cfl::Float64 = 0.5; forward_weight::Float64 = 0.65; damping_coefficient::Float64 = 0.1
    damping_length_scale::Float64 = 0; thermodynamic_tendency_factor::Float64 = 1; vertical_momentum_tendency_factor::Float64 = 1
    reserved::NTuple{8,Int32} = ntuple(_ -> Int32(0), 8)
end

const B200AcousticModel = AtmosphereModel{<:CompressibleDynamics, <:Any, <:B200, <:AcousticRungeKutta3}

function B200AcousticContext(model)
    td, dyn = model.dynamics.time_discretization, model.dynamics
    cfg = BzcConfig(base = grid_config(model; surface_pressure = dyn.surface_pressure, standard_pressure = dyn.standard_pressure),
                    reference_state = dyn.reference_state === nothing ? 0 : 1,
                    substeps = something(td.substeps, 0), acoustic_cfl = td.acoustic_cfl, forward_weight = td.forward_weight,
                    damping = td.damping isa ThermalDivergenceDamping ? 1 : 0,
                    damping_coefficient = td.damping isa ThermalDivergenceDamping ? td.damping.coefficient : 0.0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    ccall((:bzc_create, lib), Cint, (Ref{BzcConfig}, Ref{Ptr{Cvoid}}), cfg, h) == 0 ||
        error(unsafe_string(ccall((:bzc_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL)))
    if dyn.reference_state !== nothing      # θᵣ(z) of the ExnerReferenceState: θᵣ = pᵣ / (Rᵈ ρᵣ πᵣ) (reference_states.jl:611-672)
        ref = dyn.reference_state
        Rᵈ  = Breeze.dry_air_gas_constant(model.thermodynamic_constants)
        θr  = vec(Array(interior(ref.pressure))) ./ (Rᵈ .* vec(Array(interior(ref.density))) .* vec(Array(interior(ref.exner_function))))
        ccall((:bzc_set_reference_potential_temperature, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), h[], θr)
    end
    ctx = B200Context(h[])
    finalizer(c -> ccall((:bzc_destroy, lib), Cvoid, (Ptr{Cvoid},), c.handle), ctx)
    return ctx
end

const acoustic_contexts = IdDict{Any,B200Context}()
acoustic_context(model) = get!(() -> B200AcousticContext(model), acoustic_contexts, model)

# set!(model; …): Breeze fills its host fields, the prognostics (ρᵈ, ρu, ρv, ρw, ρθ, ρqᵛ) are pushed and update_state! runs on the device
function push_state!(model::B200AcousticModel)
    p = prognostic_fields(model)
    arrays = map(f -> Array(interior(f)), (model.dynamics.dry_density, p.ρu, p.ρv, p.ρw, p.ρθ, model.moisture_density))
    ctx = acoustic_context(model)
    ccall((:bzc_set_state, lib), Cint,
          (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), ctx.handle, arrays...) == 0 ||
        error(unsafe_string(ccall((:bzc_last_error, lib), Cstring, (Ptr{Cvoid},), ctx.handle)))
end

# time_step!(model::CompressibleAcousticModel, Δt): src/TimeSteppers/acoustic_runge_kutta_3.jl:264
function time_step!(model::B200AcousticModel, Δt; callbacks = [])
    ctx = acoustic_context(model)
    ccall((:bzc_time_step, lib), Cint, (Ptr{Cvoid}, Cdouble), ctx.handle, Δt) == 0 ||
        error(unsafe_string(ccall((:bzc_last_error, lib), Cstring, (Ptr{Cvoid},), ctx.handle)))
    model.clock.time += Δt; model.clock.iteration += 1
    return nothing
end

# Float32 models (`Oceananigans.defaults.FloatType = Float32`, examples/splitting_supercell.jl:86) bind the same entry points of
# libbreeze_b200_f32 under the prefix bzcf_ (include/breeze_b200_compressible_f32.h): Ptr{Float32} arrays, Cfloat scalars, the same BzcConfig.
#   ccall((:bzcf_set_state, lib32), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}), …)
#   ccall((:bzcf_time_step, lib32), Cint, (Ptr{Cvoid}, Cfloat), ctx.handle, Δt)

end # module
