"""ctypes binding of the C ABI declared in include/breeze_b200.h.

`Library(path, prefix)` binds one shared object exporting that ABI. The product uses exactly one:
`load_cuda_library()` → breeze.jl_b200/csrc/libbreeze_b200.so (prefix ``bz_``). It raises if the CUDA
library is missing or does not load: there is no CPU fallback in this package. (The CPU oracle under
oracle/ exports the same ABI with prefix ``orc_``; only tests, smoke() and bench.py's CPU-baseline legs
bind it, through oracle/oracle_lib.py.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

BZ_ABI_VERSION = 1
BZ_PERIODIC, BZ_FLAT = 0, 1
BZ_MICROPHYSICS_NONE, BZ_MICROPHYSICS_WARM_SATURATION_ADJUSTMENT = 0, 1
BZ_FORMULATION_POTENTIAL_TEMPERATURE, BZ_FORMULATION_STATIC_ENERGY = 0, 1

FIELD_IDS = {
    "ρu": 0, "ρv": 1, "ρw": 2, "ρθ": 3, "ρqᵛ": 4, "ρqᵉ": 4, "ρq": 4,
    "u": 5, "v": 6, "w": 7, "θ": 8, "qᵛ": 9, "T": 10, "φ": 11, "qˡ": 12,
    "ρe": 3, "e": 8,          # StaticEnergyFormulation: the thermodynamic slots carry ρe / e
    # ASCII aliases
    "rho_u": 0, "rho_v": 1, "rho_w": 2, "rho_theta": 3, "rho_q": 4,
    "theta": 8, "qv": 9, "phi": 11, "ql": 12,
}
PROGNOSTIC = ("ρu", "ρv", "ρw", "ρθ", "ρq")
Z_FACE_FIELDS = {2, 7}


class bz_config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("Nx", C.c_int32), ("Ny", C.c_int32), ("Nz", C.c_int32),
        ("topology_x", C.c_int32), ("topology_y", C.c_int32),
        ("x0", C.c_double), ("x1", C.c_double), ("y0", C.c_double), ("y1", C.c_double),
        ("z0", C.c_double), ("z1", C.c_double),
        ("surface_pressure", C.c_double), ("potential_temperature", C.c_double),
        ("standard_pressure", C.c_double),
        ("molar_gas_constant", C.c_double), ("gravitational_acceleration", C.c_double),
        ("energy_reference_temperature", C.c_double), ("triple_point_temperature", C.c_double),
        ("triple_point_pressure", C.c_double), ("dry_air_molar_mass", C.c_double),
        ("dry_air_heat_capacity", C.c_double), ("vapor_molar_mass", C.c_double),
        ("vapor_heat_capacity", C.c_double), ("liquid_reference_latent_heat", C.c_double),
        ("liquid_heat_capacity", C.c_double), ("ice_reference_latent_heat", C.c_double),
        ("ice_heat_capacity", C.c_double),
        ("advection_order", C.c_int32), ("microphysics", C.c_int32),
        ("n_ranks", C.c_int32), ("rank", C.c_int32), ("device", C.c_int32), ("reserved0", C.c_int32),
        ("nccl_unique_id", C.c_uint8 * 128),
        ("use_tma", C.c_int32), ("z_chunks", C.c_int32), ("formulation", C.c_int32), ("reserved", C.c_int32 * 5),
    ]


class bz_forcing(C.Structure):
    _fields_ = [
        ("coriolis_f", C.c_double),
        ("subsidence_w", C.POINTER(C.c_double)), ("subsidence_mask", C.c_int32), ("reserved0", C.c_int32),
        ("geostrophic_u", C.POINTER(C.c_double)), ("geostrophic_v", C.POINTER(C.c_double)),
        ("q_tendency", C.POINTER(C.c_double)), ("e_tendency", C.POINTER(C.c_double)),
        ("theta_flux", C.c_double), ("q_flux", C.c_double), ("drag_rho_ustar2", C.c_double),
    ]


class BreezeError(RuntimeError):
    pass


_dp = C.POINTER(C.c_double)
_vp = C.c_void_p


def abi_symbols(real=C.c_double):
    """name -> (restype, argtypes) of every symbol include/breeze_b200.h declares. `real` is the library's field type: c_double for
    libbreeze_b200.so (bz_) and the oracle (orc_), c_float for the Float32 build libbreeze_b200_f32.so (bzf_), whose entry points are the
    same with every `double` array / scalar argument a `float` (the clock of bz_get_clock and the profiles of bz_forcing stay double)."""
    rp = C.POINTER(real)
    return {
        "default_config": (None, [C.POINTER(bz_config)]),
        "abi_version": (C.c_int, []),
        "create": (C.c_int, [C.POINTER(bz_config), C.POINTER(_vp)]),
        "destroy": (None, [_vp]),
        "last_error": (C.c_char_p, [_vp]),
        "get_reference_state": (C.c_int, [_vp, rp, rp, rp]),
        "set_reference_state": (C.c_int, [_vp, rp, rp, rp]),
        "set_state": (C.c_int, [_vp, rp, rp, rp, rp, rp, C.c_int]),
        "set_forcing": (C.c_int, [_vp, C.POINTER(bz_forcing)]),
        "time_step": (C.c_int, [_vp, real]),
        "time_steps": (C.c_int, [_vp, real, C.c_int]),
        "compute_tendencies": (C.c_int, [_vp]),
        "get_tendency": (C.c_int, [_vp, C.c_int, rp]),
        "pressure_correct": (C.c_int, [_vp, real]),
        "get_field": (C.c_int, [_vp, C.c_int, rp]),
        "get_state": (C.c_int, [_vp, rp, rp, rp, rp, rp]),
        "get_clock": (C.c_int, [_vp, _dp, C.POINTER(C.c_int64)]),
        "cell_advection_timescale": (C.c_int, [_vp, rp]),
        "max_abs_divergence": (C.c_int, [_vp, rp]),
        "state_is_finite": (C.c_int, [_vp, C.POINTER(C.c_int)]),
        "get_slice": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, rp]),
        "synchronize": (C.c_int, [_vp]),
    }


def cuda_only_symbols(real=C.c_double):
    """CUDA-library-only symbols (instrumentation, asynchronous marshalling, multi-GPU bootstrap); the oracle does not export them."""
    rp = C.POINTER(real)
    return {
        "set_state_async": (C.c_int, [_vp, rp, rp, rp, rp, rp, C.c_int]),
        "get_state_async": (C.c_int, [_vp, rp, rp, rp, rp, rp]),
        "profile_enable": (C.c_int, [_vp, C.c_int]),
        "profile_read": (C.c_int, [_vp, rp, C.POINTER(C.c_int64)]),
        "kernel_launch_count": (C.c_int64, [_vp]),
        "stream": (_vp, [_vp]),
        "device_bytes": (C.c_int64, [_vp]),
        "nccl_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
        "ipc_export": (C.c_int, [_vp, C.POINTER(C.c_uint8)]),
        "ipc_attach": (C.c_int, [_vp, C.POINTER(C.c_uint8)]),
    }


ABI_SYMBOLS = abi_symbols()
CUDA_ONLY_SYMBOLS = cuda_only_symbols()


def _as_dp(a):
    """pointer to a C-contiguous float64 or float32 array (ctypes checks it against the bound library's argument type)"""
    if a is None:
        return None
    assert a.dtype in (np.float64, np.float32) and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double if a.dtype == np.float64 else C.c_float))


class Library:
    """One loaded shared object exporting the ABI with a given prefix."""

    def __init__(self, path: str, prefix: str, cuda: bool, real=np.float64):
        if not os.path.exists(path):
            raise BreezeError(f"{path} is missing — build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.path, self.prefix, self.cuda = path, prefix, cuda
        self.real = np.dtype(real).type                   # the library's field / host-array type
        self.creal = C.c_double if self.real is np.float64 else C.c_float
        self.dll = C.CDLL(path, mode=C.RTLD_GLOBAL if not cuda else C.RTLD_LOCAL)
        table = abi_symbols(self.creal)
        if cuda:
            table.update(cuda_only_symbols(self.creal))
        for name, (res, args) in table.items():
            fn = getattr(self.dll, prefix + name)      # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)

    def default_config_struct(self) -> bz_config:
        cfg = bz_config()
        self.default_config(C.byref(cfg))
        return cfg


class Context:
    """Owns one bz_ctx / orc_ctx. Array arguments and results are numpy float64, shaped (Nz[+1], Ny, Nx)
    in C order — i.e. x fastest, the memory order of Julia's `interior(field)`."""

    def __init__(self, lib: Library, cfg: bz_config):
        self.lib = lib
        self.cfg = cfg
        self.real = getattr(lib, "real", np.float64)
        self.handle = _vp()
        rc = lib.create(C.byref(cfg), C.byref(self.handle))
        if rc != 0:
            msg = lib.last_error(None)
            raise BreezeError(f"{lib.prefix}create failed ({rc}): {msg.decode() if msg else ''}")
        self.Nx_local = cfg.Nx // max(1, cfg.n_ranks)
        self.Ny, self.Nz = cfg.Ny, cfg.Nz

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

    def shape(self, field_id: int):
        nz = self.Nz + 1 if field_id in Z_FACE_FIELDS else self.Nz
        return (nz, self.Ny, self.Nx_local)

    # --- reference state -------------------------------------------------------------------------
    def reference_state(self):
        rho, p, T = (np.empty(self.Nz, dtype=self.real) for _ in range(3))
        self._check(self.lib.get_reference_state(self.handle, _as_dp(rho), _as_dp(p), _as_dp(T)), "get_reference_state")
        return rho, p, T

    def set_reference_state(self, density=None, pressure=None, temperature=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=self.real) for a in (density, pressure, temperature)]
        self._check(self.lib.set_reference_state(self.handle, *[_as_dp(a) for a in arrs]), "set_reference_state")

    # --- state -----------------------------------------------------------------------------------
    def set_state(self, rho_u=None, rho_v=None, rho_w=None, rho_theta=None, rho_q=None, enforce_mass_conservation=True):
        arrs = []
        for fid, a in enumerate((rho_u, rho_v, rho_w, rho_theta, rho_q)):
            if a is None:
                arrs.append(None)
                continue
            a = np.ascontiguousarray(a, dtype=self.real)
            if a.shape != self.shape(fid):
                raise BreezeError(f"field {fid}: expected shape {self.shape(fid)}, got {a.shape}")
            arrs.append(a)
        self._check(self.lib.set_state(self.handle, *[_as_dp(a) for a in arrs], int(enforce_mass_conservation)), "set_state")

    def set_forcing(self, coriolis_f=0.0, subsidence_w=None, subsidence_on=("u", "v", "θ", "q"), geostrophic_u=None,
                    geostrophic_v=None, q_tendency=None, e_tendency=None, theta_flux=0.0, q_flux=0.0, drag_rho_ustar2=0.0):
        """Install the forcing / Coriolis / bottom-flux terms (bz_forcing). Profiles are arrays over the z levels."""
        F = bz_forcing()
        keep = []

        def ptr(a, n):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != (n,):
                raise BreezeError(f"forcing profile: expected {n} values, got {a.shape}")
            keep.append(a)
            return a.ctypes.data_as(_dp)

        F.coriolis_f = float(coriolis_f)
        F.subsidence_w = ptr(subsidence_w, self.Nz + 1)
        F.subsidence_mask = sum(1 << {"u": 0, "v": 1, "θ": 2, "q": 3}[n] for n in subsidence_on) if subsidence_w is not None else 0
        F.geostrophic_u, F.geostrophic_v = ptr(geostrophic_u, self.Nz), ptr(geostrophic_v, self.Nz)
        F.q_tendency, F.e_tendency = ptr(q_tendency, self.Nz), ptr(e_tendency, self.Nz)
        F.theta_flux, F.q_flux, F.drag_rho_ustar2 = float(theta_flux), float(q_flux), float(drag_rho_ustar2)
        self._check(self.lib.set_forcing(self.handle, C.byref(F)), "set_forcing")

    def clear_forcing(self):
        self._check(self.lib.set_forcing(self.handle, None), "set_forcing")

    def get_field(self, name_or_id):
        fid = FIELD_IDS[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)
        out = np.empty(self.shape(fid), dtype=self.real)
        self._check(self.lib.get_field(self.handle, fid, _as_dp(out)), "get_field")
        return out

    def get_state(self, out=None):
        if out is None:
            out = [np.empty(self.shape(f), dtype=self.real) for f in range(5)]
        self._check(self.lib.get_state(self.handle, *[_as_dp(a) for a in out]), "get_state")
        return out

    def set_state_async(self, arrays, enforce_mass_conservation=False):
        """bz_set_state_async: `arrays` = five C-contiguous float64 arrays (or None) that stay alive until synchronize()."""
        self._check(self.lib.set_state_async(self.handle, *[_as_dp(a) for a in arrays], int(enforce_mass_conservation)), "set_state_async")

    def get_state_async(self, out):
        self._check(self.lib.get_state_async(self.handle, *[_as_dp(a) for a in out]), "get_state_async")

    def get_tendency(self, name_or_id):
        fid = FIELD_IDS[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)
        out = np.empty(self.shape(fid), dtype=self.real)
        self._check(self.lib.get_tendency(self.handle, fid, _as_dp(out)), "get_tendency")
        return out

    # --- stepping --------------------------------------------------------------------------------
    def time_step(self, dt):
        self._check(self.lib.time_step(self.handle, float(dt)), "time_step")

    def time_steps(self, dt, n):
        self._check(self.lib.time_steps(self.handle, float(dt), int(n)), "time_steps")

    def compute_tendencies(self):
        self._check(self.lib.compute_tendencies(self.handle), "compute_tendencies")

    def pressure_correct(self, dt):
        self._check(self.lib.pressure_correct(self.handle, float(dt)), "pressure_correct")

    def synchronize(self):
        self._check(self.lib.synchronize(self.handle), "synchronize")

    def clock(self):
        t, it = C.c_double(), C.c_int64()
        self._check(self.lib.get_clock(self.handle, C.byref(t), C.byref(it)), "get_clock")
        return t.value, it.value

    def cell_advection_timescale(self):
        tau = self.lib.creal() if hasattr(self.lib, "creal") else C.c_double()
        self._check(self.lib.cell_advection_timescale(self.handle, C.byref(tau)), "cell_advection_timescale")
        return tau.value

    def max_abs_divergence(self):
        d = self.lib.creal() if hasattr(self.lib, "creal") else C.c_double()
        self._check(self.lib.max_abs_divergence(self.handle, C.byref(d)), "max_abs_divergence")
        return d.value

    def state_is_finite(self) -> bool:
        ok = C.c_int()
        self._check(self.lib.state_is_finite(self.handle, C.byref(ok)), "state_is_finite")
        return bool(ok.value)

    def get_slice(self, name_or_id, axis: str, index: int):
        """One 2-D slice of interior(field): axis "x" → (Nz[+1], Ny), "y" → (Nz[+1], Nx), "z" → (Ny, Nx)."""
        fid = FIELD_IDS[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)
        nz, ny, nx = self.shape(fid)
        ax = {"x": 0, "y": 1, "z": 2}[axis]
        out = np.empty({0: (nz, ny), 1: (nz, nx), 2: (ny, nx)}[ax], dtype=self.real)
        self._check(self.lib.get_slice(self.handle, fid, ax, int(index), _as_dp(out)), "get_slice")
        return out

    # --- instrumentation (CUDA library only) -----------------------------------------------------
    def profile_enable(self, on=True):
        self._check(self.lib.profile_enable(self.handle, int(on)), "profile_enable")

    def profile_read(self):
        ms = np.zeros(8, dtype=self.real)
        n = np.zeros(8, dtype=np.int64)
        self._check(self.lib.profile_read(self.handle, _as_dp(ms), n.ctypes.data_as(C.POINTER(C.c_int64))), "profile_read")
        return ms, n

    def ipc_export(self) -> bytes:
        buf = (C.c_uint8 * 64)()
        self._check(self.lib.ipc_export(self.handle, buf), "ipc_export")
        return bytes(buf)

    def ipc_attach(self, handles: bytes):
        n = max(1, self.cfg.n_ranks)
        if len(handles) != 64 * n:
            raise BreezeError(f"ipc_attach: expected {64 * n} bytes of handles, got {len(handles)}")
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        self._check(self.lib.ipc_attach(self.handle, buf), "ipc_attach")

    def kernel_launch_count(self):
        return int(self.lib.kernel_launch_count(self.handle))

    def stream(self):
        return self.lib.stream(self.handle)

    def device_bytes(self):
        return int(self.lib.device_bytes(self.handle))


_CUDA_LIB = None


def cuda_library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libbreeze_b200.so")


def nccl_unique_id() -> bytes:
    """Rank 0: a fresh ncclUniqueId to broadcast to the other ranks."""
    lib = load_cuda_library()
    buf = (C.c_uint8 * 128)()
    rc = lib.nccl_unique_id(buf)
    if rc != 0:
        raise BreezeError(f"bz_nccl_unique_id failed ({rc}): {lib.last_error(None).decode()}")
    return bytes(buf)


def load_cuda_library() -> Library:
    """The product's only compute backend. Fails loudly when the CUDA extension is missing."""
    global _CUDA_LIB
    if _CUDA_LIB is None:
        _CUDA_LIB = Library(cuda_library_path(), "bz_", cuda=True)
    return _CUDA_LIB


_CUDA_LIB_F32 = None


def cuda_library_path_f32() -> str:
    return os.environ.get("BZ_F32_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libbreeze_b200_f32.so")   # BZ_F32_LIB: A/B of variant builds


def load_cuda_library_f32() -> Library:
    """The Float32 build of the anelastic path (prefix bzf_, float32 host arrays): B200(float_type="Float32")."""
    global _CUDA_LIB_F32
    if _CUDA_LIB_F32 is None:
        _CUDA_LIB_F32 = Library(cuda_library_path_f32(), "bzf_", cuda=True, real=np.float32)
    return _CUDA_LIB_F32
