"""Host mirror + ctypes binding of the compressible split-explicit path (include/breeze_b200_compressible.h).

Mirrors the reference interface for this path: `CompressibleDynamics(SplitExplicitTimeDiscretization(...); ...)`
(src/CompressibleEquations/compressible_dynamics.jl:114-177, time_discretizations.jl:540-588),
`AtmosphereModel(grid; dynamics = CompressibleDynamics(...))` → `AcousticRungeKutta3`
(src/TimeSteppers/acoustic_runge_kutta_3.jl:64-113), `set!(model; ρ, θ, u, v, w, ...)`, `time_step!(model, Δt)`.
Everything numerical runs behind the C ABI (prefix bzc_ in libbreeze_b200.so; the CPU oracle exports orcc_).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field as _dc_field
from typing import Callable, Optional

import numpy as np

from . import abi
from .abi import BreezeError, bz_config

BZC_REFERENCE_NONE, BZC_REFERENCE_EXNER = 0, 1
BZC_NO_DIVERGENCE_DAMPING, BZC_THERMAL_DIVERGENCE_DAMPING = 0, 1
BZC_PROPORTIONAL_SUBSTEPS, BZC_CONSTANT_SUBSTEP_SIZE, BZC_MONOLITHIC_FIRST_STAGE = 0, 1, 2
BZC_SPONGE_NONE, BZC_SPONGE_LINEAR_RAMP, BZC_SPONGE_CUBIC_RAMP, BZC_SPONGE_SIN2_RAMP = 0, 1, 2, 3

FIELD_IDS = {
    "ρ": 0, "ρᵈ": 0, "ρu": 1, "ρv": 2, "ρw": 3, "ρθ": 4, "u": 5, "v": 6, "w": 7, "θ": 8, "T": 9, "p": 10,
    "Gρ": 11, "Gρu": 12, "Gρv": 13, "Gρw": 14, "Gρθ": 15, "Gˢρw": 16, "Πᴸ": 17, "θᴸ": 18, "γRᵐᴸ": 19,
    "ρ′": 20, "ρθ′": 21, "ρu′": 22, "ρv′": 23, "ρw′": 24, "⟨u⟩": 25, "⟨v⟩": 26, "⟨w⟩": 27,
    "ρqᵛ": 28, "qᵛ": 29, "ρᵗ": 30, "Gρqᵛ": 31,          # ρᵗ: total density ρᵈ + ρqᵛ (dynamics.total_density)
}
Z_FACE_FIELDS = {3, 7, 16, 24, 27}
PROGNOSTIC = ("ρ", "ρu", "ρv", "ρw", "ρθ", "ρqᵛ")


class bzc_config(C.Structure):
    _fields_ = [
        ("base", bz_config),
        ("reference_state", C.c_int32), ("substeps", C.c_int32), ("damping", C.c_int32),
        ("substep_distribution", C.c_int32), ("apply_first_substep_pressure_gradient", C.c_int32), ("damp_vertical", C.c_int32),
        ("acoustic_cfl", C.c_double), ("forward_weight", C.c_double), ("damping_coefficient", C.c_double),
        ("damping_length_scale", C.c_double), ("thermodynamic_tendency_factor", C.c_double),
        ("vertical_momentum_tendency_factor", C.c_double),
        ("sponge", C.c_int32), ("reserved1", C.c_int32), ("sponge_damping_rate", C.c_double), ("sponge_depth", C.c_double),
        ("reserved", C.c_int32 * 2),
    ]


_dp, _vp = C.POINTER(C.c_double), C.c_void_p

def abi_symbols(real=C.c_double):
    """name -> (restype, argtypes) of every symbol include/breeze_b200_compressible.h declares. `real` is the library's field type: c_double
    for bzc_ (libbreeze_b200.so) and orcc_ (the oracle), c_float for bzcf_ (libbreeze_b200_f32.so), whose entry points are the same with every
    `double` array / scalar argument a `float` (bzc_config and the clock of get_clock stay double)."""
    rp = C.POINTER(real)
    return {
        "default_config": (None, [C.POINTER(bzc_config)]),
        "create": (C.c_int, [C.POINTER(bzc_config), C.POINTER(_vp)]),
        "destroy": (None, [_vp]),
        "last_error": (C.c_char_p, [_vp]),
        "set_reference_potential_temperature": (C.c_int, [_vp, rp]),
        "get_reference_state": (C.c_int, [_vp, rp, rp, rp]),
        "set_state": (C.c_int, [_vp, rp, rp, rp, rp, rp, rp]),
        "time_step": (C.c_int, [_vp, real]),
        "time_steps": (C.c_int, [_vp, real, C.c_int]),
        "compute_slow_tendencies": (C.c_int, [_vp]),
        "stage_substep_count_and_size": (C.c_int, [_vp, real, real, C.POINTER(C.c_int32), rp]),
        "acoustic_substep_loop": (C.c_int, [_vp, real, real]),
        "get_field": (C.c_int, [_vp, C.c_int, rp]),
        "get_state": (C.c_int, [_vp, rp, rp, rp, rp, rp, rp]),
        "get_clock": (C.c_int, [_vp, _dp, C.POINTER(C.c_int64)]),
        "synchronize": (C.c_int, [_vp]),
    }


def cuda_only_symbols(real=C.c_double):
    return {
        "profile_enable": (C.c_int, [_vp, C.c_int]),
        "profile_read": (C.c_int, [_vp, C.POINTER(real), C.POINTER(C.c_int64)]),
        "kernel_launch_count": (C.c_int64, [_vp]),
        "stream": (_vp, [_vp]),
        "device_bytes": (C.c_int64, [_vp]),
    }


ABI_SYMBOLS = abi_symbols()
CUDA_ONLY_SYMBOLS = cuda_only_symbols()


class CompressibleLibrary:
    """The compressible entry points of one loaded shared object (prefix bzc_ or orcc_)."""

    def __init__(self, lib: abi.Library):
        self.base, self.dll, self.cuda = lib, lib.dll, lib.cuda
        self.prefix = {"bz_": "bzc_", "bzf_": "bzcf_", "orc_": "orcc_"}.get(lib.prefix, lib.prefix[:-1] + "c_")
        self.real = getattr(lib, "creal", C.c_double)           # ctypes type of the library's reals
        self.dtype = getattr(lib, "real", np.float64)           # numpy type of its field arrays
        table = dict(abi_symbols(self.real))
        if lib.cuda:
            table.update(cuda_only_symbols(self.real))
        for name, (res, args) in table.items():
            fn = getattr(self.dll, self.prefix + name)       # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)


_LIBS: dict[int, CompressibleLibrary] = {}


def compressible_library(lib: abi.Library) -> CompressibleLibrary:
    if id(lib) not in _LIBS:
        _LIBS[id(lib)] = CompressibleLibrary(lib)
    return _LIBS[id(lib)]


def _as_dp(a):
    """pointer to a C-contiguous float64 or float32 array (ctypes checks it against the bound library's argument type)"""
    if a is None:
        return None
    assert a.dtype in (np.float64, np.float32) and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double if a.dtype == np.float64 else C.c_float))


class CompressibleContext:
    """Owns one bzc_ctx / orcc_ctx. Arrays are numpy float64 shaped (Nz[+1], Ny, Nx): x fastest, Julia's interior(field)."""

    def __init__(self, lib: CompressibleLibrary, cfg: bzc_config):
        self.lib, self.cfg, self.handle = lib, cfg, _vp()
        self.real = lib.dtype                                    # numpy dtype of every field array of this library
        rc = lib.create(C.byref(cfg), C.byref(self.handle))
        if rc != 0:
            msg = lib.last_error(None)
            raise BreezeError(f"{lib.prefix}create failed ({rc}): {msg.decode() if msg else ''}")
        b = cfg.base
        self.Nx = 1 if b.topology_x == abi.BZ_FLAT else b.Nx
        self.Ny = 1 if b.topology_y == abi.BZ_FLAT else b.Ny
        self.Nz = b.Nz

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.last_error(self.handle)
            raise BreezeError(f"{self.lib.prefix}{what} failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if self.handle:
            self.lib.destroy(self.handle)
            self.handle = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shape(self, fid: int):
        return (self.Nz + 1 if fid in Z_FACE_FIELDS else self.Nz, self.Ny, self.Nx)

    def set_reference_potential_temperature(self, theta_r):
        a = np.ascontiguousarray(theta_r, dtype=self.real)
        if a.shape != (self.Nz,):
            raise BreezeError(f"θᵣ profile: expected {self.Nz} values, got {a.shape}")
        self._check(self.lib.set_reference_potential_temperature(self.handle, _as_dp(a)), "set_reference_potential_temperature")

    def reference_state(self):
        p, rho, pi = (np.empty(self.Nz, dtype=self.real) for _ in range(3))
        self._check(self.lib.get_reference_state(self.handle, _as_dp(p), _as_dp(rho), _as_dp(pi)), "get_reference_state")
        return p, rho, pi

    def set_state(self, rho=None, rho_u=None, rho_v=None, rho_w=None, rho_theta=None, rho_qv=None):
        """`rho` is the DRY density ρᵈ; a context that never receives `rho_qv` stays dry."""
        arrs = []
        for fid, a in enumerate((rho, rho_u, rho_v, rho_w, rho_theta, rho_qv)):
            if a is None:
                arrs.append(None)
                continue
            a = np.ascontiguousarray(a, dtype=self.real)
            if a.shape != self.shape(fid if fid < 5 else 0):
                raise BreezeError(f"field {PROGNOSTIC[fid]}: expected shape {self.shape(fid if fid < 5 else 0)}, got {a.shape}")
            arrs.append(a)
        self._check(self.lib.set_state(self.handle, *[_as_dp(a) for a in arrs]), "set_state")

    def time_step(self, dt):
        self._check(self.lib.time_step(self.handle, float(dt)), "time_step")

    def time_steps(self, dt, n):
        self._check(self.lib.time_steps(self.handle, float(dt), int(n)), "time_steps")

    def compute_slow_tendencies(self):
        self._check(self.lib.compute_slow_tendencies(self.handle), "compute_slow_tendencies")

    def stage_substep_count_and_size(self, dt, beta):
        n, d = C.c_int32(), self.lib.real()
        self._check(self.lib.stage_substep_count_and_size(self.handle, float(dt), float(beta), C.byref(n), C.byref(d)),
                    "stage_substep_count_and_size")
        return n.value, d.value

    def acoustic_substep_loop(self, dt, beta):
        self._check(self.lib.acoustic_substep_loop(self.handle, float(dt), float(beta)), "acoustic_substep_loop")

    def get_field(self, name_or_id):
        fid = FIELD_IDS[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)
        out = np.empty(self.shape(fid), dtype=self.real)
        self._check(self.lib.get_field(self.handle, fid, _as_dp(out)), "get_field")
        return out

    def get_state(self, out=None):
        """The prognostics (ρᵈ, ρu, ρv, ρw, ρθ[, ρqᵛ]) in one call, into `out` (e.g. pinned buffers) when given."""
        if out is None:
            out = [np.empty(self.shape(f), dtype=self.real) for f in range(5)]
        ptrs = [_as_dp(a) for a in out] + [None] * (6 - len(out))
        self._check(self.lib.get_state(self.handle, *ptrs), "get_state")
        return out

    def clock(self):
        t, it = C.c_double(), C.c_int64()
        self._check(self.lib.get_clock(self.handle, C.byref(t), C.byref(it)), "get_clock")
        return t.value, it.value

    def synchronize(self):
        self._check(self.lib.synchronize(self.handle), "synchronize")

    # instrumentation (CUDA library only)
    def profile_enable(self, on=True):
        self._check(self.lib.profile_enable(self.handle, int(on)), "profile_enable")

    def profile_read(self):
        ms, n = np.zeros(8, dtype=self.real), np.zeros(8, dtype=np.int64)
        self._check(self.lib.profile_read(self.handle, _as_dp(ms), n.ctypes.data_as(C.POINTER(C.c_int64))), "profile_read")
        return ms, n

    def kernel_launch_count(self):
        return int(self.lib.kernel_launch_count(self.handle))

    def stream(self):
        return self.lib.stream(self.handle)

    def device_bytes(self):
        return int(self.lib.device_bytes(self.handle))


# ---- reference interface -------------------------------------------------------------------------------------------
@dataclass
class NoDivergenceDamping:
    pass


@dataclass
class ThermalDivergenceDamping:
    """time_discretizations.jl:229-247"""
    coefficient: float = 0.1
    length_scale: Optional[float] = None
    damp_vertical: bool = False


@dataclass
class LinearRamp:
    pass


@dataclass
class CubicRamp:
    pass


@dataclass
class Sin2Ramp:
    pass


@dataclass
class UpperSponge:
    """UpperSponge(damping_rate = 0.2, depth = 5e3, ramp = CubicRamp()) (time_discretizations.jl:440-520)."""
    damping_rate: float = 0.2
    depth: float = 5e3
    ramp: object = _dc_field(default_factory=CubicRamp)

    def __post_init__(self):
        if not isinstance(self.ramp, (LinearRamp, CubicRamp, Sin2Ramp)):
            raise ValueError("`ramp` must be an `<:AbstractRamp` (e.g. `CubicRamp()`, `Sin2Ramp()`, `LinearRamp()`)")


@dataclass
class ProportionalSubsteps:
    pass


@dataclass
class ConstantSubstepSize:
    pass


@dataclass
class MonolithicFirstStage:
    pass


@dataclass
class SplitExplicitTimeDiscretization:
    """time_discretizations.jl:540-588 (open boundaries are not on the path)."""
    substeps: Optional[int] = None
    acoustic_cfl: float = 0.5
    forward_weight: float = 0.65
    thermodynamic_tendency_factor: float = 1.0
    vertical_momentum_tendency_factor: float = 1.0
    apply_first_substep_pressure_gradient: bool = False
    damping: object = _dc_field(default_factory=ThermalDivergenceDamping)
    sponge: object = None
    substep_distribution: object = _dc_field(default_factory=ProportionalSubsteps)

    def __post_init__(self):
        if not isinstance(self.damping, (NoDivergenceDamping, ThermalDivergenceDamping)):
            raise ValueError("`damping` must be an `AcousticDampingStrategy`")
        if self.sponge is not None and not isinstance(self.sponge, UpperSponge):
            raise ValueError("`sponge` must be `nothing` or an `UpperSponge`")
        if not self.acoustic_cfl > 0:
            raise ValueError(f"`acoustic_cfl` must be positive (got {self.acoustic_cfl})")


@dataclass
class CompressibleDynamics:
    """CompressibleDynamics(time_discretization; standard_pressure, surface_pressure, reference_potential_temperature,
    reference_state) (compressible_dynamics.jl:114-177). `reference_potential_temperature`: a constant or a function θᵣ(z)."""
    time_discretization: SplitExplicitTimeDiscretization = _dc_field(default_factory=SplitExplicitTimeDiscretization)
    standard_pressure: float = 1e5
    surface_pressure: float = 101325.0
    reference_potential_temperature: object = None
    reference_state: object = "auto"

    def __post_init__(self):
        if self.reference_state not in ("auto", None):
            raise ValueError(f"`reference_state` must be `:auto` or `nothing`; received {self.reference_state!r}.")
        if self.reference_state is None and self.reference_potential_temperature is not None:
            raise ValueError("`reference_state = nothing` disables the reference state and is mutually exclusive with an "
                             "explicit reference profile")
        if not isinstance(self.time_discretization, SplitExplicitTimeDiscretization):
            raise NotImplementedError("only SplitExplicitTimeDiscretization is on the path")


class CompressibleAtmosphereModel:
    """AtmosphereModel(grid; dynamics = CompressibleDynamics(SplitExplicitTimeDiscretization(...)), advection = WENO(order=5))
    stepped by AcousticRungeKutta3 (dry or vapour-laden air, microphysics = nothing)."""

    def __init__(self, grid, dynamics: CompressibleDynamics, advection=None, thermodynamic_constants=None, microphysics=None):
        from .model import Flat, ThermodynamicConstants, WENO
        if microphysics is not None:
            raise NotImplementedError("the compressible path carries vapour only (microphysics = nothing)")
        self._moist = False
        self.grid, self.architecture, self.dynamics = grid, grid.architecture, dynamics
        self.thermodynamic_constants = thermodynamic_constants or ThermodynamicConstants()
        self.advection = advection or WENO(order=5)
        if self.advection.order not in (5, 7, 9):
            raise NotImplementedError("WENO(order = 5, 7 or 9) is on the hot path")
        lib = compressible_library(self.architecture.library())
        cfg = bzc_config()
        lib.default_config(C.byref(cfg))
        b, td = cfg.base, dynamics.time_discretization
        b.Nx, b.Ny, b.Nz = grid.Nx, grid.Ny, grid.Nz
        b.topology_x = abi.BZ_FLAT if grid.topology[0] is Flat else abi.BZ_PERIODIC
        b.topology_y = abi.BZ_FLAT if grid.topology[1] is Flat else abi.BZ_PERIODIC
        b.x0, b.x1, b.y0, b.y1, b.z0, b.z1 = grid.x0, grid.x1, grid.y0, grid.y1, grid.z0, grid.z1
        b.surface_pressure, b.standard_pressure = dynamics.surface_pressure, dynamics.standard_pressure
        θr = dynamics.reference_potential_temperature
        b.potential_temperature = 288.0 if (θr is None or callable(θr)) else float(θr)
        for name in ThermodynamicConstants.__dataclass_fields__:
            setattr(b, name, getattr(self.thermodynamic_constants, name))
        b.device = getattr(self.architecture, "device", 0)
        b.advection_order = self.advection.order
        cfg.reference_state = BZC_REFERENCE_NONE if dynamics.reference_state is None else BZC_REFERENCE_EXNER
        cfg.substeps = int(td.substeps or 0)
        cfg.acoustic_cfl, cfg.forward_weight = td.acoustic_cfl, td.forward_weight
        cfg.thermodynamic_tendency_factor = td.thermodynamic_tendency_factor
        cfg.vertical_momentum_tendency_factor = td.vertical_momentum_tendency_factor
        cfg.apply_first_substep_pressure_gradient = int(td.apply_first_substep_pressure_gradient)
        if isinstance(td.damping, ThermalDivergenceDamping):
            cfg.damping = BZC_THERMAL_DIVERGENCE_DAMPING
            cfg.damping_coefficient = td.damping.coefficient
            cfg.damping_length_scale = float(td.damping.length_scale or 0.0)
            cfg.damp_vertical = int(td.damping.damp_vertical)
        else:
            cfg.damping = BZC_NO_DIVERGENCE_DAMPING
        if td.sponge is not None:
            cfg.sponge = {LinearRamp: 1, CubicRamp: 2, Sin2Ramp: 3}[type(td.sponge.ramp)]
            cfg.sponge_damping_rate, cfg.sponge_depth = td.sponge.damping_rate, td.sponge.depth
        cfg.substep_distribution = {ProportionalSubsteps: 0, ConstantSubstepSize: 1, MonolithicFirstStage: 2}[type(td.substep_distribution)]
        self.context = CompressibleContext(lib, cfg)
        if callable(θr):
            self.context.set_reference_potential_temperature([θr(z) for z in grid.znodes()])

    # --- set! ------------------------------------------------------------------------------------
    def set(self, **kw):
        """set!(model; ρ, θ, u, v, w, ρu, ρv, ρw, ρθ) (set_atmosphere_model.jl:198-360): density first, then the thermodynamic
        variable and velocities weighted by it (ρθ = ρᵈ θ; set_velocity!: ρu = ℑ(ρᵈ) u with periodic / zero-gradient halos)."""
        from .model import Flat, _evaluate
        ctx, g = self.context, self.grid
        # Python NFKC-normalises identifiers, so a keyword written qᵛ arrives as "qv" (and ρᵈ as "ρd"): normalise every key alike
        import unicodedata
        nf = lambda t: unicodedata.normalize("NFKC", t)                     # noqa: E731
        canon = {nf(t): t for t in ("ρ", "ρᵈ", "θ", "u", "v", "w", "ρu", "ρv", "ρw", "ρθ", "qᵛ", "qᵗ", "qᵉ", "ρqᵛ", "ρqᵗ")}
        canon.update({nf("θˡⁱ"): "θ", nf("ρθˡⁱ"): "ρθ", "rho": "ρ", "theta": "θ"})
        kw = {canon.get(nf(k), k): v for k, v in kw.items()}
        for k in kw:
            if k not in ("ρ", "ρᵈ", "θ", "u", "v", "w", "ρu", "ρv", "ρw", "ρθ", "qᵛ", "qᵗ", "qᵉ", "ρqᵛ", "ρqᵗ"):
                raise ValueError(f"Cannot set! {k} in AtmosphereModel because {k} is neither a prognostic variable, "
                                 "a settable thermodynamic variable, nor a settable diagnostic variable!")
        xs, ys, zs, zf = g.xnodes(), g.ynodes(), g.znodes(), g.znodes(face=True)
        xf, yf = g.xnodes(face=True), g.ynodes(face=True)
        cshape, wshape = ctx.shape(0), ctx.shape(3)
        # establish_densities! (compressible_time_stepping.jl:89-137): `ρ` is the TOTAL density (ρᵈ = ρ − ρqᵛ), `ρᵈ` the dry one
        # (ρ = ρᵈ/(1 − qᵛ)); moisture as a mass fraction of the total density (qᵛ / qᵗ) or as a density (ρqᵛ)
        moist_given = [k for k in ("qᵛ", "qᵗ", "qᵉ", "ρqᵛ", "ρqᵗ") if k in kw]
        if len(moist_given) > 1:
            raise ValueError(f"set! one moisture variable at a time, got {moist_given}")
        total = _evaluate(kw["ρ"], g, xs, ys, zs, cshape) if "ρ" in kw else None
        dry = _evaluate(kw["ρᵈ"], g, xs, ys, zs, cshape) if "ρᵈ" in kw else None
        if total is not None and dry is not None:
            raise ValueError("set! either the total density ρ or the dry density ρᵈ")
        rqv = None
        if moist_given:
            mval = _evaluate(kw[moist_given[0]], g, xs, ys, zs, cshape)
            if moist_given[0].startswith("ρ"):
                rqv = mval
                if total is not None:
                    dry = total - rqv
            else:
                if total is not None:
                    rqv, dry = total * mval, total - total * mval
                else:
                    if dry is None:
                        dry = ctx.get_field("ρ")
                    tot = dry / (1 - mval)
                    rqv = tot * mval
        elif total is not None:
            dry = total - (ctx.get_field("ρqᵛ") if self._moist else 0.0)
        rho = dry if dry is not None else ctx.get_field("ρ")         # coupling density ρᵈ: ρθ = ρᵈ θ, ρu = ℑ(ρᵈ) u
        args = dict(rho=dry, rho_qv=rqv)
        if rqv is not None:
            self._moist = True
        for comp, axis, coords in (("u", 2, (xf, ys, zs)), ("v", 1, (xs, yf, zs)), ("w", 0, (xs, ys, zf))):
            if "ρ" + comp in kw:
                args["rho_" + comp] = _evaluate(kw["ρ" + comp], g, *coords, wshape if comp == "w" else cshape)
            elif comp in kw:
                val = _evaluate(kw[comp], g, *coords, wshape if comp == "w" else cshape)
                if comp == "w":
                    rf = np.empty(wshape)
                    rf[1:-1] = 0.5 * (rho[1:] + rho[:-1])
                    rf[0], rf[-1] = rho[0], rho[-1]
                    val = val * rf
                    val[0] = 0.0
                    val[-1] = 0.0
                else:
                    flat = g.topology[0 if comp == "u" else 1] is Flat
                    rf = rho if flat else 0.5 * (rho + np.roll(rho, 1, axis=axis))
                    val = val * rf
                args["rho_" + comp] = val
        if "ρθ" in kw:
            args["rho_theta"] = _evaluate(kw["ρθ"], g, xs, ys, zs, cshape)
        elif "θ" in kw:
            args["rho_theta"] = _evaluate(kw["θ"], g, xs, ys, zs, cshape) * rho
        ctx.set_state(**args)

    def field(self, name):
        return self.context.get_field(name)

    @property
    def clock(self):
        t, it = self.context.clock()
        return {"time": t, "iteration": it}

    def time_step(self, Δt):
        """time_step!(model::CompressibleAcousticModel, Δt) (acoustic_runge_kutta_3.jl:264)."""
        self.context.time_step(Δt)

    def reference_profiles(self):
        return self.context.reference_state()
