"""The reference's own checks of the BOMEX-type forcings (test/geostrophic_subsidence_forcings.jl), restated through the host
mirror on the CPU oracle and — same assertions — through the C ABI on the GPU:

  * `:12-40`   GeostrophicForcing: uᵍ = -10, vᵍ = 0, f = 1e-4, one tiny step from rest ⇒ min ρv < 0 (Fρv = f ρᵣ uᵍ);
  * `:145-170` geostrophic + subsidence on u and v together ⇒ max ρv < 0;
  * `:91-129`  SubsidenceForcing on the moisture with θ = θ₀: a profile decreasing with height under sinking motion loses moisture;
  * `:284-329` constant-gradient profile ϕ = Γ z under uniform wˢ: Δ(ρϕ) = ρᵣ (-Δt wˢ Γ) at the lowest and the highest level
               (rtol 1e-3) for u, θ, qᵛ;
  * `:234-282` … and linearly in N over five steps (1e-3 of the expected change) for θ and qᵛ.

Differences of setup, none of them touching the assertions: the path carries WENO5 advection and a power-of-two horizontal
grid (the reference uses `advection = nothing` on 1 × 1 × 4 and 4 × 4 × 4 cells); the fields are horizontally uniform and at
rest, so advection contributes nothing.
"""
import numpy as np
import pytest


def _grid(arch, nz=4, lz=100.0):
    import breeze_b200 as bz
    return bz.RectilinearGrid(arch, size=(8, 8, nz), x=(0, 100.0), y=(0, 100.0), z=(0, lz))


def _model(arch, grid=None, **kw):
    import breeze_b200 as bz
    grid = grid or _grid(arch)
    return bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid)), advection=bz.WENO(order=5), **kw)


def _theta0(m):
    return m.dynamics.reference_state.potential_temperature


def check_geostrophic_smoke(arch):
    import breeze_b200 as bz
    m = _model(arch, coriolis=bz.FPlane(f=1e-4), forcing=bz.geostrophic_forcings(lambda z: -10.0, lambda z: 0.0))
    m.set(θ=_theta0(m))
    m.time_step(1e-6)
    assert m.field("ρv").min() < 0


def check_combined(arch):
    import breeze_b200 as bz
    geo = bz.geostrophic_forcings(lambda z: -10.0, lambda z: 0.0)
    sub = bz.SubsidenceForcing(lambda z: -0.01)
    m = _model(arch, coriolis=bz.FPlane(f=1e-4), forcing={"u": (sub, geo["u"]), "v": (sub, geo["v"])})
    m.set(θ=_theta0(m))
    m.time_step(1e-6)
    assert m.field("ρv").max() < 0


def check_moisture_subsidence(arch):
    import breeze_b200 as bz
    m = _model(arch, _grid(arch, nz=10, lz=1000.0), forcing={"qᵛ": bz.SubsidenceForcing(lambda z: -0.01)})
    m.set(θ=_theta0(m), qᵗ=lambda x, y, z: 0.015 - 1e-5 * z)
    before = m.field("ρq").sum()
    for _ in range(3):
        m.time_step(0.1)
    after = m.field("ρq").sum()
    assert np.isfinite(after) and after < before


def _gradient_case(arch, specific, n_steps):
    """ϕ = Γ z under uniform wˢ = 1; returns (actual change of ρϕ per level, expected ρᵣ N Δϕ)."""
    import breeze_b200 as bz
    gamma, dt, ws = 1e-2, 1e-2, 1.0
    m = _model(arch, _grid(arch, nz=4, lz=16.0), forcing={specific: bz.SubsidenceForcing(ws)})
    prof = lambda x, y, z: gamma * z
    if specific == "u":
        m.set(θ=_theta0(m), u=prof, enforce_mass_conservation=False)
        name = "ρu"
    elif specific == "θ":
        m.set(θ=prof)
        name = "ρθ"
    else:
        m.set(θ=_theta0(m), qᵗ=prof)
        name = "ρq"
    rho = m.reference_profiles()[0]
    f0 = m.field(name).copy()
    for _ in range(n_steps):
        m.time_step(dt)
    change = (m.field(name) - f0)[:, 2, 3]
    return change, n_steps * rho * (-dt * ws * gamma)


def check_gradient(arch, specific):
    change, expected = _gradient_case(arch, specific, 1)
    assert change[0] == pytest.approx(expected[0], rel=1e-3)
    assert change[-1] == pytest.approx(expected[-1], rel=1e-3)


def check_linear_accumulation(arch, specific):
    change, expected = _gradient_case(arch, specific, 5)
    assert np.max(np.abs(change - expected)) < 1e-3 * np.max(np.abs(expected))


# ---- CPU oracle -----------------------------------------------------------------------------------------------------
def test_oracle_geostrophic_smoke(oracle_arch):
    check_geostrophic_smoke(oracle_arch)


def test_oracle_combined_geostrophic_and_subsidence(oracle_arch):
    check_combined(oracle_arch)


def test_oracle_moisture_subsidence_dries_the_column(oracle_arch):
    check_moisture_subsidence(oracle_arch)


@pytest.mark.parametrize("specific", ["u", "θ", "qᵛ"])
def test_oracle_subsidence_constant_gradient(oracle_arch, specific):
    check_gradient(oracle_arch, specific)


@pytest.mark.parametrize("specific", ["θ", "qᵛ"])
def test_oracle_subsidence_linear_accumulation(oracle_arch, specific):
    check_linear_accumulation(oracle_arch, specific)


# ---- CUDA path --------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_geostrophic_and_combined():
    import breeze_b200 as bz
    check_geostrophic_smoke(bz.B200())
    check_combined(bz.B200())


@pytest.mark.gpu
def test_gpu_moisture_subsidence_dries_the_column():
    import breeze_b200 as bz
    check_moisture_subsidence(bz.B200())


@pytest.mark.gpu
@pytest.mark.parametrize("specific", ["u", "θ", "qᵛ"])
def test_gpu_subsidence_constant_gradient(specific):
    import breeze_b200 as bz
    check_gradient(bz.B200(), specific)
    if specific != "u":
        check_linear_accumulation(bz.B200(), specific)
