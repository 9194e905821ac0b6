"""The Float32 build of the compressible split-explicit path (libbreeze_b200_f32.so, prefix bzcf_; `B200(float_type="Float32")`) — the
precision the reference's own supercell example runs in (examples/splitting_supercell.jl:86). Compiled from the mechanically retyped copy
of the FP64 sources (breeze.jl_b200/make_f32.py), so these tests check that it is the same algorithm — agreement with the FP64 CPU oracle
to Float32 round-off, tolerances stated per test and measured values logged (BZ_PARITY_REPORT) — and the invariants that do not depend on
precision. Declared tolerances (relative to each field's max-norm):
  * reference state (pᵣ, ρᵣ, πᵣ: a discrete hydrostatic integration in Float32)                  : 2e-6   (measured 3e-7)
  * update_state! diagnostics                                                                    : 2e-6   (measured 2e-7)
  * five WS-RK3 steps of the warm bubble: ρ, ρθ, θ, p                                            : 5e-6   (measured 5e-7)
  * five steps: momentum and velocities (relative to the largest component)                      : 3e-4   (measured 3.4e-5)
    (the pressure-gradient force is a difference of pressures ≈ 1e5 Pa carrying 6e-3 Pa of Float32 round-off)
  * a resting hydrostatic atmosphere after 50 steps: max |w| < 1e-3 m/s                                    (measured 6e-5; the FP64 path holds 1e-10)
Measured values: profiles/r2z_parity_errors_compressible_f32.txt.
"""
import numpy as np
import pytest

from conftest import rel_err, report

pytestmark = pytest.mark.gpu

PROGNOSTIC = ["ρ", "ρu", "ρv", "ρw", "ρθ"]
TOL_REF, TOL_DIAG, TOL_THERMO, TOL_MOMENTUM = 2e-6, 2e-6, 5e-6, 3e-4


def _model(arch, size, flat_y=False, **td):
    import breeze_b200 as bz
    if flat_y:
        grid = bz.RectilinearGrid(arch, size=size, x=(-5e3, 5e3), z=(0, 10e3), topology=(bz.Periodic, bz.Flat, bz.Bounded))
    else:
        grid = bz.RectilinearGrid(arch, size=size, x=(-5e3, 5e3), y=(-5e3, 5e3), z=(0, 10e3))
    dyn = bz.CompressibleDynamics(bz.SplitExplicitTimeDiscretization(**td), reference_potential_temperature=300.0)
    return bz.AtmosphereModel(grid, dynamics=dyn)


def _pair(oracle_arch, size, flat_y=False, **td):
    import breeze_b200 as bz
    models = [_model(a, size, flat_y, **td) for a in (bz.B200(float_type="Float32"), oracle_arch)]
    g = models[0].grid
    _, rho_r, _ = models[1].reference_profiles()
    rho = np.broadcast_to(rho_r[:, None, None], (g.Nz, g.Ny, g.Nx)).copy()

    def theta(*xyz):
        x, z = xyz[0], xyz[-1]
        r2 = x ** 2 + (z - 3000.0) ** 2 + (xyz[1] ** 2 if len(xyz) == 3 else 0.0)
        return 300.0 + 2.0 * np.cos(np.pi / 2 * np.minimum(1.0, np.sqrt(r2) / 2000.0)) ** 2

    for m in models:
        m.set(ρ=rho, θ=theta, u=3.0, v=0.0 if flat_y else -2.0)
    return models


def test_library_and_reference_state(oracle_arch):
    gpu, cpu = _model(__import__("breeze_b200").B200(float_type="Float32"), (8, 8, 40)), _model(oracle_arch, (8, 8, 40))
    assert gpu.context.lib.prefix == "bzcf_" and gpu.field("ρθ").dtype == np.float32
    for a, b in zip(gpu.reference_profiles(), cpu.reference_profiles()):
        assert a.dtype == np.float32
        assert rel_err(a.astype(np.float64), b) < TOL_REF


@pytest.mark.parametrize("size,flat_y", [((32, 16, 24), False), ((64, 40), True)])
def test_update_state_matches_oracle(oracle_arch, size, flat_y):
    gpu, cpu = _pair(oracle_arch, size, flat_y)
    for name in PROGNOSTIC + ["u", "v", "w", "θ", "T", "p"]:
        assert rel_err(gpu.field(name).astype(np.float64), cpu.field(name)) < TOL_DIAG, name


@pytest.mark.parametrize("size,flat_y,dt", [((32, 32, 24), False, 2.0), ((64, 40), True, 1.0)])
def test_bubble_steps_match_oracle(oracle_arch, size, flat_y, dt):
    gpu, cpu = _pair(oracle_arch, size, flat_y, substeps=6)
    for m in (gpu, cpu):
        for _ in range(5):
            m.time_step(dt)
    assert gpu.clock == cpu.clock
    for name in ("ρ", "ρθ", "θ", "p"):
        assert rel_err(gpu.field(name).astype(np.float64), cpu.field(name)) < TOL_THERMO, name
    mom = max(np.abs(cpu.field(n)).max() for n in ("ρu", "ρv", "ρw"))
    for name in ("ρu", "ρv", "ρw"):
        err = np.abs(gpu.field(name) - cpu.field(name)).max() / mom
        report(err, name)
        assert err < TOL_MOMENTUM, name
    vel = max(np.abs(cpu.field(n)).max() for n in ("u", "w"))
    for name in ("u", "w"):
        err = np.abs(gpu.field(name) - cpu.field(name)).max() / vel
        report(err, name)
        assert err < TOL_MOMENTUM, name
    assert np.abs(gpu.field("w")).max() > 1e-3                # the bubble does move


def test_mass_conservation_walls_and_rest_state():
    """Discrete mass conservation to Float32 accumulation error, impenetrable walls exactly, and a resting hydrostatic atmosphere
    that stays at rest to Float32 round-off of the pressure-gradient / buoyancy balance (test/substepper_rest_state.jl, FP64 there: 1e-10)."""
    import breeze_b200 as bz
    arch = bz.B200(float_type="Float32")
    m = _model(arch, (32, 32, 24), substeps=6)
    _, rho_r, _ = m.reference_profiles()
    m.set(ρ=np.broadcast_to(rho_r[:, None, None], m.context.shape(0)).copy(),
          θ=lambda x, y, z: 300.0 + 2.0 * np.exp(-(x ** 2 + y ** 2 + (z - 3000.0) ** 2) / 1500.0 ** 2), u=3.0)
    M0 = m.field("ρ").astype(np.float64).sum()
    for _ in range(3):
        m.time_step(2.0)
    assert abs(m.field("ρ").astype(np.float64).sum() - M0) / M0 <= 1e-6
    rw = m.field("ρw")
    assert np.abs(rw[-1]).max() == 0 and np.abs(rw[0]).max() == 0
    rest = _model(arch, (8, 8, 32), substeps=6)
    _, rho_r, _ = rest.reference_profiles()
    rest.set(ρ=np.broadcast_to(rho_r[:, None, None], rest.context.shape(0)).copy(), θ=300.0)
    for _ in range(50):
        rest.time_step(2.0)
    w = float(np.abs(rest.field("w")).max())
    report(w, "rest-state max|w| after 50 steps")
    assert np.isfinite(w) and w < 1e-3


def test_config4_shape_in_float32():
    """BASELINE config 4's shape (256 x 256 x 64, 6 substeps) in the precision the shipped example uses: runs, stays finite, the warm
    perturbation rises."""
    import breeze_b200 as bz
    grid = bz.RectilinearGrid(bz.B200(float_type="Float32"), size=(256, 256, 64), x=(0, 168e3), y=(0, 168e3), z=(0, 20e3))
    dyn = bz.CompressibleDynamics(bz.SplitExplicitTimeDiscretization(substeps=6), reference_potential_temperature=300.0)
    m = bz.AtmosphereModel(grid, dynamics=dyn)
    _, rho, _ = m.reference_profiles()
    m.set(ρ=np.broadcast_to(rho[:, None, None], m.context.shape(0)).copy(),
          θ=lambda x, y, z: 300.0 + 3.0 * np.exp(-((x - 84e3) ** 2 + (y - 84e3) ** 2) / 10e3 ** 2 - (z - 1500.0) ** 2 / 1500.0 ** 2), u=10.0)
    for _ in range(3):
        m.time_step(6.0)
    m.context.synchronize()
    w = m.field("w")
    assert w.dtype == np.float32 and np.isfinite(w).all() and np.abs(w).max() > 1e-4
