"""CPU-only checks of the boundary: the CUDA library loads, exports every symbol include/breeze_b200.h declares, refuses to
run without a GPU (no CPU fallback), and the host mirror validates its arguments like the reference does."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import breeze_b200 as bz
from breeze_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="breeze_b200.h", prefix="bz_"):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z_0-9]+)\s*\(", text)) - {prefix + "ctx", prefix + "config"})


def test_cuda_library_exports_every_declared_symbol():
    lib = abi.load_cuda_library()
    for name in _declared_symbols():
        assert hasattr(lib.dll, name), name
    assert lib.abi_version() == abi.BZ_ABI_VERSION


def test_cuda_library_exports_every_compressible_symbol():
    """include/breeze_b200_compressible.h: every declared bzc_* entry point is exported and bound by the host mirror."""
    lib = abi.load_cuda_library()
    declared = _declared_symbols("breeze_b200_compressible.h", "bzc_")
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib.dll, name), name
    clib = bz.compressible_library(lib)
    cfg = bz.bzc_config()
    clib.default_config(C.byref(cfg))
    assert cfg.forward_weight == 0.65 and cfg.acoustic_cfl == 0.5 and cfg.damping_coefficient == 0.1 and cfg.base.abi_version == 1
    assert C.sizeof(bz.bzc_config) == C.sizeof(abi.bz_config) + 4 * 6 + 8 * 6 + 4 * 8
    h = C.c_void_p()
    import torch
    if not torch.cuda.is_available():
        assert clib.create(C.byref(cfg), C.byref(h)) != 0
        assert b"no CPU fallback" in clib.last_error(None)


def test_oracle_exports_the_compressible_abi(oracle_arch):
    from breeze_b200 import compressible
    lib = oracle_arch.library()
    for name in compressible.ABI_SYMBOLS:
        assert hasattr(lib.dll, "orcc_" + name), name


def test_compressible_argument_validation(oracle_arch):
    with pytest.raises(ValueError):
        bz.CompressibleDynamics(reference_state="none")
    with pytest.raises(ValueError):
        bz.CompressibleDynamics(reference_state=None, reference_potential_temperature=300.0)
    with pytest.raises(ValueError):
        bz.SplitExplicitTimeDiscretization(acoustic_cfl=0)
    with pytest.raises(ValueError):
        bz.SplitExplicitTimeDiscretization(damping=(bz.ThermalDivergenceDamping(), bz.NoDivergenceDamping()))
    grid = bz.RectilinearGrid(oracle_arch, size=(8, 8, 8), x=(0, 1e3), y=(0, 1e3), z=(0, 1e3))
    m = bz.AtmosphereModel(grid, dynamics=bz.CompressibleDynamics())
    assert isinstance(m, bz.CompressibleAtmosphereModel)
    with pytest.raises(ValueError):
        m.set(qˡ=0.01)
    with pytest.raises(ValueError):
        m.set(ρ=1.0, ρᵈ=1.0)
    with pytest.raises(bz.BreezeError):
        m.context.set_state(rho=np.zeros((3, 3, 3)))


def test_oracle_exports_the_same_abi(oracle_arch):
    lib = oracle_arch.library()
    for name in abi.ABI_SYMBOLS:
        assert hasattr(lib.dll, "orc_" + name), name


def test_config_struct_layout_matches_c():
    # sizeof(bz_config) seen by ctypes must equal the C compiler's (the library fills it in bz_default_config)
    lib = abi.load_cuda_library()
    cfg = lib.default_config_struct()
    assert cfg.abi_version == 1 and cfg.advection_order == 5 and cfg.n_ranks == 1
    assert cfg.surface_pressure == 101325.0 and cfg.ice_heat_capacity == 2108.0
    assert C.sizeof(abi.bz_config) == 4 * 6 + 8 * 22 + 4 * 6 + 128 + 4 * 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = abi.load_cuda_library()
    cfg = lib.default_config_struct()
    h = C.c_void_p()
    assert lib.create(C.byref(cfg), C.byref(h)) != 0
    assert b"no CPU fallback" in lib.last_error(None)
    with pytest.raises(bz.BreezeError):
        grid = bz.RectilinearGrid(size=(8, 8, 8), x=(0, 1), y=(0, 1), z=(0, 1))
        bz.AtmosphereModel(grid)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "breeze.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle", text, re.M), f
                assert "liboracle" not in text, f


def test_grid_and_model_argument_validation(oracle_arch):
    with pytest.raises(ValueError):
        bz.RectilinearGrid(oracle_arch, size=(8, 8, 8), x=(0, 1), y=(0, 1), z=(0, 1), topology=(bz.Periodic, bz.Periodic, bz.Periodic))
    with pytest.raises(ValueError):
        bz.RectilinearGrid(oracle_arch, size=(8, 8), x=(0, 1), y=(0, 1), z=(0, 1))
    grid = bz.RectilinearGrid(oracle_arch, size=(8, 8, 8), x=(0, 1), y=(0, 1), z=(0, 1))
    with pytest.raises(NotImplementedError):
        bz.AtmosphereModel(grid, advection=bz.WENO(order=11))
    model = bz.AtmosphereModel(grid)
    with pytest.raises(ValueError):
        model.set(banana=1.0)
    with pytest.raises(bz.BreezeError):
        model.context.set_state(rho_u=np.zeros((3, 3, 3)))
    # model construction leaves θ = θ₀ (initialize_model_thermodynamics!)
    assert np.allclose(model.field("θ"), 288.0)
    assert model.field("ρw").shape == (9, 8, 8) and model.field("ρu").shape == (8, 8, 8)


def test_set_velocity_uses_face_density(oracle_arch):
    grid = bz.RectilinearGrid(oracle_arch, size=(8, 8, 8), x=(0, 1), y=(0, 1), z=(0, 4000.0))
    model = bz.AtmosphereModel(grid)
    model.set(u=2.0, enforce_mass_conservation=False)
    rho = model.reference_profiles()[0]
    assert np.allclose(model.field("ρu"), 2.0 * rho[:, None, None])
    assert np.allclose(model.field("u"), 2.0)


def test_time_step_wizard(oracle_arch):
    grid = bz.RectilinearGrid(oracle_arch, size=(8, 8, 8), x=(0, 800.0), y=(0, 800.0), z=(0, 800.0))
    model = bz.AtmosphereModel(grid)
    model.set(u=10.0)
    assert model.context.cell_advection_timescale() == pytest.approx(100.0 / 10.0, rel=1e-12)
    sim = bz.Simulation(model, Δt=1.0, stop_iteration=1)
    bz.conjure_time_step_wizard_(sim, cfl=0.5, interval=1)
    bz.run_(sim)
    assert sim.Δt == pytest.approx(1.1)                              # max_change limits the growth towards cfl·τ = 5


def test_nan_checker_and_slices(oracle_arch):
    """NaNChecker of run! (atmosphere_model.jl:561-572) and 2-D slice output, host logic + oracle side of the ABI."""
    grid = bz.RectilinearGrid(oracle_arch, size=(8, 6, 5), x=(0, 800.0), y=(0, 600.0), z=(0, 500.0))
    model = bz.AtmosphereModel(grid)
    rng = np.random.default_rng(0)
    model.set(u=rng.standard_normal((5, 6, 8)), θ=288 + rng.standard_normal((5, 6, 8)))
    full, w = model.field("θ"), model.field("ρw")
    assert np.array_equal(model.slice("θ", x=3), full[:, :, 3])
    assert np.array_equal(model.slice("θ", y=2), full[:, 2, :])
    assert np.array_equal(model.slice("θ", z=4), full[4])
    assert np.array_equal(model.slice("ρw", y=5), w[:, 5, :]) and model.slice("ρw", y=5).shape == (6, 8)
    with pytest.raises(bz.BreezeError):
        model.slice("θ", z=5)
    assert model.context.state_is_finite()
    sim = bz.Simulation(model, Δt=0.1, stop_iteration=3)
    bz.run_(sim, nan_check_interval=1)
    assert model.clock["iteration"] == 3
    bad = model.field("ρu")
    bad[1, 2, 3] = np.nan
    model.context.set_state(rho_u=bad, enforce_mass_conservation=False)
    assert not model.context.state_is_finite()
    sim = bz.Simulation(model, Δt=0.1, stop_iteration=10)
    bz.run_(sim, nan_check_interval=1)
    assert model.clock["iteration"] == 3                              # the checker stopped the run before another step


def test_float32_library_exports_the_same_entry_points():
    """csrc/libbreeze_b200_f32.so (make_f32.py): every entry point of include/breeze_b200.h under the prefix bzf_; loading binds them all."""
    from breeze_b200 import abi
    lib = abi.load_cuda_library_f32()
    assert lib.prefix == "bzf_" and lib.real is np.float32
    for name in list(abi.abi_symbols()) + list(abi.cuda_only_symbols()):
        assert hasattr(lib.dll, "bzf_" + name), name
    assert lib.abi_version() == abi.BZ_ABI_VERSION
    declared = _declared_symbols("breeze_b200_f32.h", "bzf_")           # include/breeze_b200_f32.h declares exactly what is exported
    assert set(declared) == {"bzf_" + n for n in list(abi.abi_symbols()) + list(abi.cuda_only_symbols())}
    generated = os.path.join(ROOT, "breeze.jl_b200", "csrc", "f32", "stage_kernel.cuh")
    code = [ln.split("//")[0] for ln in open(generated).read().split("\n")]
    assert not any(re.search(r"\bdouble\b", ln) for ln in code)


def test_float32_library_exports_the_compressible_entry_points():
    """The compressible path of the Float32 library (prefix bzcf_, include/breeze_b200_compressible_f32.h — the precision
    examples/splitting_supercell.jl:86 runs in): every declared symbol is exported, and the header is what make_f32.py --headers generates."""
    from breeze_b200 import abi, compressible
    lib = compressible.compressible_library(abi.load_cuda_library_f32())
    assert lib.prefix == "bzcf_" and lib.dtype is np.float32
    names = list(compressible.abi_symbols()) + list(compressible.cuda_only_symbols())
    for name in names:
        assert hasattr(lib.dll, "bzcf_" + name), name
    assert set(_declared_symbols("breeze_b200_compressible_f32.h", "bzcf_")) == {"bzcf_" + n for n in names}
    assert set(_declared_symbols("breeze_b200_compressible.h", "bzc_")) == {"bzc_" + n for n in names}
    generated = os.path.join(ROOT, "breeze.jl_b200", "csrc", "f32", "compressible.cuh")
    code = [ln.split("//")[0] for ln in open(generated).read().split("\n")]
    assert not any(re.search(r"\bdouble\b", ln) for ln in code)
