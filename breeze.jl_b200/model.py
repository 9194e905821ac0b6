"""Host-side mirror of the Breeze.jl interface for the anelastic SSP-RK3 hot path.

Same names, argument meaning and error behaviour as the reference's user-facing constructors, so the
parity tests read like the reference's own tests (Julia's `set!`/`time_step!`/`run!` become `set_`,
`time_step_`, `run_` plus methods):

    grid      = RectilinearGrid(size=(256, 256), x=(-10e3, 10e3), z=(0, 10e3), topology=(Periodic, Flat, Bounded))
    reference = ReferenceState(grid, potential_temperature=300)
    model     = AtmosphereModel(grid, dynamics=AnelasticDynamics(reference), advection=WENO(order=5))
    set_(model, θ=lambda x, z: 300 + ...)
    simulation = Simulation(model, Δt=2, stop_iteration=100); run_(simulation)

Reference: README.md:67-76, src/AtmosphereModels/atmosphere_model.jl:114-314,
src/AtmosphereModels/set_atmosphere_model.jl:198-360, src/TimeSteppers/ssp_runge_kutta_3.jl:209-278.
All arithmetic happens behind the C ABI (abi.py); this file only marshals arguments.
"""
from __future__ import annotations

from dataclasses import dataclass, field as _dc_field
from typing import Callable, Optional, Sequence

import numpy as np

from . import abi

# --- topologies ------------------------------------------------------------------------------------


class Periodic: ...


class Flat: ...


class Bounded: ...


class B200:
    """The architecture: hand-written sm_100a kernels behind libbreeze_b200.so (one context per GPU)."""

    def __init__(self, device: int = 0, rank: int = 0, n_ranks: int = 1, nccl_unique_id: bytes | None = None,
                 use_tma: int = 0, z_chunks: int = 0, float_type: str = "Float64"):
        if float_type not in ("Float64", "Float32"):
            raise ValueError(f"float_type must be Float64 or Float32, got {float_type!r}")
        self.device, self.rank, self.n_ranks = device, rank, n_ranks
        self.nccl_unique_id = nccl_unique_id
        self.use_tma, self.z_chunks = use_tma, z_chunks
        self.float_type = float_type                      # RectilinearGrid(GPU(), Float32; ...) of the reference (benchmarking/README.md:74)

    def library(self) -> abi.Library:
        return abi.load_cuda_library_f32() if self.float_type == "Float32" else abi.load_cuda_library()


@dataclass
class RectilinearGrid:
    """RectilinearGrid(arch; size, x, y, z, topology) with uniform spacing (Oceananigans).

    `size`/extents list only the non-Flat dimensions, in (x, y, z) order, as in Oceananigans."""
    architecture: object = None
    size: Sequence[int] = (8, 8, 8)
    x: Optional[Sequence[float]] = None
    y: Optional[Sequence[float]] = None
    z: Sequence[float] = (0.0, 1.0)
    topology: Sequence[type] = (Periodic, Periodic, Bounded)

    def __post_init__(self):
        if self.architecture is None:
            self.architecture = B200()
        topo = tuple(self.topology)
        if topo[2] is not Bounded:
            raise ValueError("the anelastic hot path needs a Bounded z dimension")
        for t in topo[:2]:
            if t not in (Periodic, Flat):
                raise ValueError("x and y must be Periodic or Flat on this path (the FFT-based pressure solver)")
        size = (self.size,) if np.isscalar(self.size) else tuple(self.size)
        active = [d for d, t in enumerate(topo) if t is not Flat]
        if len(size) != len(active):
            raise ValueError(f"size={size} does not match the {len(active)} non-Flat dimensions of topology")
        full = [1, 1, 1]
        for n, d in zip(size, active):
            full[d] = int(n)
        self.Nx, self.Ny, self.Nz = full
        ext = [self.x, self.y, self.z]
        for d, t in enumerate(topo):
            if t is Flat:
                ext[d] = (0.0, 1.0)
            elif ext[d] is None:
                raise ValueError(f"extent for dimension {'xyz'[d]} is required")
        (self.x0, self.x1), (self.y0, self.y1), (self.z0, self.z1) = [tuple(map(float, e)) for e in ext]
        self.topology = topo
        self.Δx = (self.x1 - self.x0) / self.Nx if topo[0] is not Flat else 1.0
        self.Δy = (self.y1 - self.y0) / self.Ny if topo[1] is not Flat else 1.0
        self.Δz = (self.z1 - self.z0) / self.Nz

    # node coordinates (global)
    def xnodes(self, face=False):
        return self.x0 + (np.arange(self.Nx) + (0.0 if face else 0.5)) * self.Δx

    def ynodes(self, face=False):
        return self.y0 + (np.arange(self.Ny) + (0.0 if face else 0.5)) * self.Δy

    def znodes(self, face=False):
        n = self.Nz + 1 if face else self.Nz
        return self.z0 + (np.arange(n) + (0.0 if face else 0.5)) * self.Δz


@dataclass
class ThermodynamicConstants:
    """src/Thermodynamics/thermodynamics_constants.jl:182-212 (defaults identical)."""
    molar_gas_constant: float = 8.314462618
    gravitational_acceleration: float = 9.81
    energy_reference_temperature: float = 273.15
    triple_point_temperature: float = 273.16
    triple_point_pressure: float = 611.657
    dry_air_molar_mass: float = 0.02897
    dry_air_heat_capacity: float = 1005.0
    vapor_molar_mass: float = 0.018015
    vapor_heat_capacity: float = 1850.0
    liquid_reference_latent_heat: float = 2500800.0
    liquid_heat_capacity: float = 4181.0
    ice_reference_latent_heat: float = 2834000.0
    ice_heat_capacity: float = 2108.0


@dataclass
class ReferenceState:
    """ReferenceState(grid, constants; surface_pressure, potential_temperature, standard_pressure)
    (src/Thermodynamics/reference_states.jl:402-445). Profiles are evaluated by the library."""
    grid: RectilinearGrid
    constants: ThermodynamicConstants = _dc_field(default_factory=ThermodynamicConstants)
    surface_pressure: float = 101325.0
    potential_temperature: float = 288.0
    standard_pressure: float = 1e5
    density: Optional[np.ndarray] = None      # optional override, like `set!(reference_state.density, f)`


@dataclass
class AnelasticDynamics:
    reference_state: ReferenceState


@dataclass
class WENO:
    order: int = 5


@dataclass
class FPlane:
    """Oceananigans FPlane(f=...)"""
    f: float = 0.0


class SubsidenceForcing:
    """SubsidenceForcing(wˢ): wˢ a function of z or an array at the Nz+1 z-faces (src/Forcings/subsidence_forcing.jl)."""

    def __init__(self, wˢ):
        self.w = wˢ


class GeostrophicForcing:
    def __init__(self, velocity, direction):
        self.velocity, self.direction = velocity, direction


def geostrophic_forcings(uᵍ, vᵍ):
    """geostrophic_forcings(uᵍ, vᵍ) → forcings keyed `u` (uses vᵍ) and `v` (uses uᵍ) (src/Forcings/geostrophic_forcings.jl:128-132)."""
    return {"u": GeostrophicForcing(vᵍ, "x"), "v": GeostrophicForcing(uᵍ, "y")}


class Forcing:
    """Forcing(field): a prescribed horizontally uniform profile (function of z or array over the Nz centres)."""

    def __init__(self, profile):
        self.profile = profile


class FluxBoundaryCondition:
    """Bottom FluxBoundaryCondition with a constant value (examples/bomex.jl:77-84)."""

    def __init__(self, value):
        self.value = float(value)


class DragFluxBoundaryCondition:
    """Bottom momentum flux -ρ₀ u★² ρu / |ρ𝐮ₕ| (examples/bomex.jl:93-101); give the same object for ρu and ρv."""

    def __init__(self, rho0, u_star):
        self.rho0_ustar2 = float(rho0) * float(u_star) ** 2


class SaturationAdjustment:
    """SaturationAdjustment(equilibrium=WarmPhaseEquilibrium()) (src/Microphysics/saturation_adjustment.jl:23-55)."""

    def __init__(self, equilibrium="WarmPhaseEquilibrium"):
        if equilibrium not in ("WarmPhaseEquilibrium",):
            raise NotImplementedError("only the warm-phase equilibrium is on the path (BOMEX)")
        self.equilibrium = equilibrium


def _evaluate(value, grid: RectilinearGrid, xs, ys, zs, shape):
    """set!(field, value): a number, an array of the interior shape, or a function of the non-Flat coordinates."""
    if callable(value):
        Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij", sparse=True)      # broadcastable 1-D axes
        args = [a for a, t in zip((X, Y, Z), grid.topology) if t is not Flat]
        try:
            out = np.asarray(value(*args), dtype=np.float64)                # numpy-vectorised callable
            out = np.broadcast_to(out, shape)
        except Exception:
            full = [np.broadcast_to(a, shape) for a in args]
            out = np.vectorize(value, otypes=[float])(*full)                # scalar callable
        return np.ascontiguousarray(out, dtype=np.float64)
    arr = np.asarray(value, dtype=np.float64)
    return np.ascontiguousarray(np.broadcast_to(arr, shape)).copy()


class AtmosphereModel:
    """AtmosphereModel(grid; dynamics, advection, microphysics, thermodynamic_constants)
    for the configurations on the hot path: AnelasticDynamics, WENO(order=5), SSPRungeKutta3,
    LiquidIcePotentialTemperature formulation, closure = nothing.

    `dynamics = CompressibleDynamics(SplitExplicitTimeDiscretization(...))` dispatches to the AcousticRungeKutta3 model of
    breeze_b200.compressible (the reference selects the stepper from the dynamics type, dynamics_interface.jl:71)."""

    def __new__(cls, grid=None, dynamics=None, *args, **kw):
        from .compressible import CompressibleAtmosphereModel, CompressibleDynamics
        if isinstance(dynamics, CompressibleDynamics):
            kw.pop("timestepper", None)
            return CompressibleAtmosphereModel(grid, dynamics, *args, **kw)
        return super().__new__(cls)

    def __init__(self, grid: RectilinearGrid, dynamics: AnelasticDynamics | None = None, advection: WENO | None = None,
                 microphysics=None, thermodynamic_constants: ThermodynamicConstants | None = None,
                 timestepper: str = "SSPRungeKutta3", coriolis: FPlane | None = None, forcing: dict | None = None,
                 boundary_conditions: dict | None = None, formulation: str = "LiquidIcePotentialTemperature"):
        if timestepper != "SSPRungeKutta3":
            raise NotImplementedError("only :SSPRungeKutta3 is on the hot path")
        if formulation not in ("LiquidIcePotentialTemperature", "StaticEnergy"):
            raise ValueError(f"formulation must be :LiquidIcePotentialTemperature or :StaticEnergy, got {formulation!r}")
        if formulation == "StaticEnergy" and (microphysics is not None or forcing or boundary_conditions):
            raise NotImplementedError("StaticEnergyFormulation is on the path without microphysics / forcings only")
        self.formulation = formulation
        self.grid = grid
        self.architecture = grid.architecture
        if dynamics is None:
            dynamics = AnelasticDynamics(ReferenceState(grid, thermodynamic_constants or ThermodynamicConstants()))
        self.dynamics = dynamics
        ref = dynamics.reference_state
        self.thermodynamic_constants = thermodynamic_constants or ref.constants
        self.advection = advection or WENO(order=5)
        if self.advection.order not in (5, 7, 9):
            raise NotImplementedError("WENO(order = 5, 7 or 9) is on the hot path")
        self.microphysics = microphysics

        lib = self.architecture.library()
        cfg = lib.default_config_struct()
        cfg.Nx, cfg.Ny, cfg.Nz = grid.Nx, grid.Ny, grid.Nz
        cfg.topology_x = abi.BZ_FLAT if grid.topology[0] is Flat else abi.BZ_PERIODIC
        cfg.topology_y = abi.BZ_FLAT if grid.topology[1] is Flat else abi.BZ_PERIODIC
        cfg.x0, cfg.x1, cfg.y0, cfg.y1, cfg.z0, cfg.z1 = grid.x0, grid.x1, grid.y0, grid.y1, grid.z0, grid.z1
        cfg.surface_pressure = ref.surface_pressure
        cfg.potential_temperature = ref.potential_temperature
        cfg.standard_pressure = ref.standard_pressure
        for name in ThermodynamicConstants.__dataclass_fields__:
            setattr(cfg, name, getattr(self.thermodynamic_constants, name))
        cfg.advection_order = self.advection.order
        cfg.microphysics = (abi.BZ_MICROPHYSICS_NONE if microphysics is None
                            else abi.BZ_MICROPHYSICS_WARM_SATURATION_ADJUSTMENT)
        arch = self.architecture
        cfg.n_ranks, cfg.rank, cfg.device = getattr(arch, "n_ranks", 1), getattr(arch, "rank", 0), getattr(arch, "device", 0)
        cfg.use_tma, cfg.z_chunks = getattr(arch, "use_tma", 0), getattr(arch, "z_chunks", 0)
        cfg.formulation = abi.BZ_FORMULATION_STATIC_ENERGY if formulation == "StaticEnergy" else abi.BZ_FORMULATION_POTENTIAL_TEMPERATURE
        uid = getattr(arch, "nccl_unique_id", None)
        if uid is not None:
            for n, b in enumerate(uid[:128]):
                cfg.nccl_unique_id[n] = b
        self.context = abi.Context(lib, cfg)
        if ref.density is not None:
            self.context.set_reference_state(density=ref.density)
        # local x-slab of this rank
        self.Nx_local = self.context.Nx_local
        self.i0 = cfg.rank * self.Nx_local
        self.coriolis, self.forcing, self.boundary_conditions = coriolis, forcing, boundary_conditions
        if coriolis is not None or forcing or boundary_conditions:
            self._install_forcing()

    def _profile(self, value, face=False):
        z = self.grid.znodes(face=face)
        if callable(value):
            return np.array([value(zz) for zz in z], dtype=np.float64)
        a = np.asarray(value, dtype=np.float64)
        return np.full(len(z), float(a)) if a.ndim == 0 else a          # a constant (SubsidenceForcing(wˢ::Number), subsidence_forcing.jl)

    def _install_forcing(self):
        """forcing = (; u = (subsidence, geostrophic.u), θ = subsidence, qᵉ = (subsidence, Forcing(drying)), e = Forcing(cooling))
        and boundary_conditions = (ρθ=…, ρqᵉ=…, ρu=…, ρv=…) as in examples/bomex.jl:119-208."""
        kw = dict(coriolis_f=self.coriolis.f if self.coriolis else 0.0)
        alias = {"θ": "θ", "θˡⁱ": "θ", "qᵉ": "q", "qᵗ": "q", "qᵛ": "q", "q": "q", "u": "u", "v": "v", "e": "e"}
        subs_on, subs_w = [], None
        for key, items in (self.forcing or {}).items():
            name = alias.get(key)
            if name is None:
                raise ValueError(f"forcing key {key}: supply forcings under the specific names u, v, θ, qᵉ, e")
            for item in (items if isinstance(items, (tuple, list)) else (items,)):
                if isinstance(item, SubsidenceForcing):
                    if name == "e":
                        raise ValueError("SubsidenceForcing applies to u, v, θ, qᵉ")
                    w = self._profile(item.w, face=True)
                    if subs_w is not None and not np.array_equal(w, subs_w):
                        raise NotImplementedError("one subsidence profile per model")
                    subs_w = w
                    subs_on.append(name)
                elif isinstance(item, GeostrophicForcing):
                    f = kw["coriolis_f"]
                    if f == 0.0:
                        raise ValueError("geostrophic forcing needs coriolis=FPlane(f)")
                    kw["geostrophic_v" if item.direction == "x" else "geostrophic_u"] = self._profile(item.velocity)
                elif isinstance(item, Forcing):
                    if name == "q":
                        kw["q_tendency"] = self._profile(item.profile)
                    elif name == "e":
                        kw["e_tendency"] = self._profile(item.profile)
                    else:
                        raise NotImplementedError("Forcing(field) is on the path for qᵉ and e only")
                else:
                    raise NotImplementedError(f"forcing of type {type(item).__name__}")
        if subs_w is not None:
            kw.update(subsidence_w=subs_w, subsidence_on=tuple(subs_on))
        for key, bc in (self.boundary_conditions or {}).items():
            if isinstance(bc, DragFluxBoundaryCondition):
                kw["drag_rho_ustar2"] = bc.rho0_ustar2
            elif key in ("ρθ", "ρθˡⁱ"):
                kw["theta_flux"] = bc.value
            elif key in ("ρqᵉ", "ρqᵗ", "ρqᵛ", "ρq"):
                kw["q_flux"] = bc.value
            else:
                raise NotImplementedError(f"boundary condition on {key}")
        self.context.set_forcing(**kw)

    # --- coordinates of the local slab -----------------------------------------------------------
    def _coords(self, loc):
        g = self.grid
        xs = g.xnodes(face=(loc == "u"))[self.i0:self.i0 + self.Nx_local]
        ys = g.ynodes(face=(loc == "v"))
        zs = g.znodes(face=(loc == "w"))
        return xs, ys, zs

    def reference_profiles(self):
        return self.context.reference_state()

    # --- set! ------------------------------------------------------------------------------------
    def set(self, enforce_mass_conservation=True, **kw):
        """set!(model; θ, u, v, w, qᵗ, ρu, ρv, ρw, ρθ, ρqᵛ, ...) (set_atmosphere_model.jl:198-360)."""
        ctx, g = self.context, self.grid
        rho, p_ref, _ = ctx.reference_state()
        rho_c = rho[:, None, None]
        rho_f = np.empty(g.Nz + 1)
        rho_f[1:-1] = 0.5 * (rho[1:] + rho[:-1])
        rho_f[0], rho_f[-1] = rho[0], rho[-1]          # wall faces: only ever multiply w = 0
        rho_f = rho_f[:, None, None]
        args = dict(rho_u=None, rho_v=None, rho_w=None, rho_theta=None, rho_q=None)
        energy = self.formulation == "StaticEnergy"
        if energy:
            # set_thermodynamic_variable!(::StaticEnergyModel, …) (static_energy_tendency.jl:76-190): θ or T → e = cᵖᵐ T + g z,
            # T = Π θ with the moisture set in the same call (vapour only)
            allowed = {"θ", "θˡⁱ", "T", "e", "ρe"}
            given = [k for k in kw if k in allowed]
            if len(given) > 1:
                raise ValueError(f"set! one thermodynamic variable at a time, got {given}")
            if given and given[0] != "ρe":
                name = given[0]
                xs, ys, zs = self._coords("c")
                val = _evaluate(kw.pop(name), g, xs, ys, zs, ctx.shape(3))
                qname = next((k for k in ("qᵗ", "qᵛ", "qᵉ", "qt", "qv") if k in kw), None)
                q = _evaluate(kw[qname], g, xs, ys, zs, ctx.shape(3)) if qname else ctx.get_field("ρq") / rho_c
                c = self.thermodynamic_constants
                Rd, Rv = c.molar_gas_constant / c.dry_air_molar_mass, c.molar_gas_constant / c.vapor_molar_mass
                Rm, cpm = (1 - q) * Rd + q * Rv, (1 - q) * c.dry_air_heat_capacity + q * c.vapor_heat_capacity
                if name in ("θ", "θˡⁱ"):
                    val = (p_ref[:, None, None] / self.dynamics.reference_state.standard_pressure) ** (Rm / cpm) * val
                if name != "e":
                    val = cpm * val + c.gravitational_acceleration * zs[:, None, None]
                kw["ρθ"] = val * rho_c                       # the thermodynamic slot of the ABI carries ρe
            elif given:
                kw["ρθ"] = kw.pop("ρe")
        aliases = {"θ": "theta", "θˡⁱ": "theta", "qᵗ": "q", "qᵛ": "q", "qᵉ": "q", "ρθ": "rho_theta", "ρθˡⁱ": "rho_theta",
                   "ρqᵗ": "rho_q", "ρqᵛ": "rho_q", "ρqᵉ": "rho_q", "ρu": "rho_u", "ρv": "rho_v", "ρw": "rho_w",
                   "qt": "q", "qv": "q"}
        for name, value in kw.items():
            key = aliases.get(name, name)
            if key in ("u", "v", "w", "rho_u", "rho_v", "rho_w"):
                comp = key[-1]
                xs, ys, zs = self._coords(comp)
                fid = {"u": 0, "v": 1, "w": 2}[comp]
                arr = _evaluate(value, g, xs, ys, zs, ctx.shape(fid))
                if not key.startswith("rho_"):
                    arr = arr * (rho_f if comp == "w" else rho_c)      # set_velocity!: ρu = ℑ(ρ) u
                args["rho_" + comp] = arr
            elif key in ("theta", "q", "rho_theta", "rho_q"):
                xs, ys, zs = self._coords("c")
                arr = _evaluate(value, g, xs, ys, zs, ctx.shape(3))
                if not key.startswith("rho_"):
                    arr = arr * rho_c
                args["rho_theta" if key.endswith("theta") else "rho_q"] = arr
            else:
                raise ValueError(
                    f"Cannot set! {name} in AtmosphereModel because {name} is neither a prognostic variable, "
                    "a settable thermodynamic variable, nor a settable diagnostic variable!")
        ctx.set_state(enforce_mass_conservation=enforce_mass_conservation, **args)

    # --- fields ----------------------------------------------------------------------------------
    def field(self, name):
        """interior(field) as a numpy array shaped (Nz[+1], Ny, Nx): x fastest, as in Julia memory."""
        return self.context.get_field(name)

    def slice(self, name, **at):
        """view(field, :, j, :) etc. for slice output: model.slice("θ", y=64) → (Nz, Nx); indices are 0-based, x is rank-local."""
        (axis, index), = at.items()
        return self.context.get_slice(name, axis, index)

    @property
    def clock(self):
        t, it = self.context.clock()
        return {"time": t, "iteration": it}

    def time_step(self, Δt):
        self.context.time_step(Δt)


def enable_peer_memory(model: AtmosphereModel):
    """Multi-GPU: map the neighbouring ranks' device arenas (CUDA IPC) so that ghost cells and the FFT transposes become peer
    loads over NVLink inside the consuming kernels. Needs an initialised torch.distributed process group (one rank per GPU);
    a no-op on one rank."""
    ctx = model.context
    if ctx.cfg.n_ranks <= 1:
        return
    import torch
    import torch.distributed as dist
    mine = torch.tensor(list(ctx.ipc_export()), dtype=torch.uint8, device="cuda")
    parts = [torch.empty_like(mine) for _ in range(ctx.cfg.n_ranks)]
    dist.all_gather(parts, mine)
    ctx.ipc_attach(b"".join(bytes(p.cpu().tolist()) for p in parts))


def set_(model: AtmosphereModel, **kw):
    model.set(**kw)


def time_step_(model: AtmosphereModel, Δt):
    """time_step!(model, Δt) — one SSP-RK3 step (src/TimeSteppers/ssp_runge_kutta_3.jl:209)."""
    model.time_step(Δt)


def many_time_steps_(model: AtmosphereModel, Δt, N=100):
    """benchmarking/src/timestepping.jl:11-16"""
    model.context.time_steps(Δt, N)


class TimeStepWizard:
    """Oceananigans TimeStepWizard(cfl, max_change, min_change, max_Δt) driven by cell_advection_timescale
    (src/AtmosphereModels/cell_advection_timescale.jl:46-65)."""

    def __init__(self, cfl=0.2, max_change=1.1, min_change=0.5, max_Δt=np.inf, min_Δt=0.0):
        self.cfl, self.max_change, self.min_change, self.max_Δt, self.min_Δt = cfl, max_change, min_change, max_Δt, min_Δt

    def __call__(self, sim):
        τ = sim.model.context.cell_advection_timescale()
        new = self.cfl * τ
        new = min(new, self.max_change * sim.Δt)
        new = max(new, self.min_change * sim.Δt)
        sim.Δt = float(min(max(new, self.min_Δt), self.max_Δt))


class Simulation:
    """Simulation(model; Δt, stop_time, stop_iteration) with callbacks every `interval` iterations."""

    def __init__(self, model: AtmosphereModel, Δt, stop_time=np.inf, stop_iteration=np.inf):
        self.model, self.Δt, self.stop_time, self.stop_iteration = model, float(Δt), stop_time, stop_iteration
        self.callbacks: list[tuple[Callable, int]] = []

    def add_callback(self, f: Callable, interval: int = 1):
        self.callbacks.append((f, interval))


def conjure_time_step_wizard_(simulation: Simulation, cfl=0.7, interval=10, **kw):
    simulation.add_callback(TimeStepWizard(cfl=cfl, **kw), interval)


class NaNChecker:
    """Oceananigans NaNChecker as installed by default_nan_checker(::AtmosphereModel) (atmosphere_model.jl:561-572): every
    `interval` iterations a device-side reduction over the prognostic fields; stops the simulation on a NaN."""

    def __init__(self, erroring=False):
        self.erroring = erroring

    def __call__(self, sim):
        if not sim.model.context.state_is_finite():
            t, it = sim.model.context.clock()
            msg = f"time = {t}, iteration = {it}: NaN found in field ρu. Stopping simulation."
            if self.erroring:
                raise FloatingPointError(msg)
            print(msg)
            sim.running = False


def run_(simulation: Simulation, nan_check_interval: int = 100):
    """run!(simulation): the loop around time_step! (callbacks, wizard, NaN checker every 100 iterations as in Oceananigans)."""
    m = simulation.model
    simulation.running = True
    checker = NaNChecker()
    t, it = m.context.clock()
    while simulation.running and t < simulation.stop_time and it < simulation.stop_iteration:
        for f, interval in simulation.callbacks:
            if it % interval == 0:
                f(simulation)
        if nan_check_interval and it % nan_check_interval == 0:
            checker(simulation)
            if not simulation.running:
                break
        Δt = min(simulation.Δt, simulation.stop_time - t)
        m.time_step(Δt)
        t, it = m.context.clock()
    m.context.synchronize()
