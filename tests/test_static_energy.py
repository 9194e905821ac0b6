"""StaticEnergyFormulation on the anelastic path: prognostic ρe, e = cᵖᵐ T + g z
(src/StaticEnergyFormulations/static_energy_tendency.jl:39-72, src/Thermodynamics/dynamic_states.jl:283-312).

CPU: the oracle against the reference's checks — setting θ or T gives the same temperature under both formulations
(test/set_atmosphere_model.jl:12-140), the e ↔ T round trip (test/unit_tests.jl:311-334) — and the physical cross-check that
both formulations evolve a warm bubble to the same state up to truncation error. GPU: CUDA vs oracle parity."""
import numpy as np
import pytest

from conftest import bubble_theta, rel_err, report

G, CPD = 9.81, 1005.0


def _model(arch, formulation, size=(16, 8, 12), **kw):
    import breeze_b200 as bz
    grid = bz.RectilinearGrid(arch, size=size, x=(-10e3, 10e3), y=(-5e3, 5e3), z=(0, 10e3))
    return bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)),
                              formulation=formulation, **kw)


def test_setting_theta_gives_the_same_temperature_in_both_formulations(oracle_arch):
    """test/set_atmosphere_model.jl:12-98"""
    a, b = _model(oracle_arch, "LiquidIcePotentialTemperature"), _model(oracle_arch, "StaticEnergy")
    for m in (a, b):
        m.set(θ=bubble_theta(), qᵗ=0.004)
    assert np.abs(a.field("T") - b.field("T")).max() < 1e-11
    z = b.grid.znodes()[:, None, None]
    cpm = (1 - 0.004) * CPD + 0.004 * 1850.0
    assert rel_err(b.field("e"), cpm * b.field("T") + G * z) < 1e-14          # with_temperature(::StaticEnergyState)
    assert np.array_equal(b.field("ρe"), b.context.get_field(3))


def test_setting_temperature_and_energy_round_trip(oracle_arch):
    """test/set_atmosphere_model.jl:100-140, test/unit_tests.jl:311-334: T → e → T."""
    m = _model(oracle_arch, "StaticEnergy")
    T0 = 280.0 + 10 * np.random.default_rng(0).random(m.context.shape(3))
    m.set(T=T0)
    assert np.abs(m.field("T") - T0).max() < 1e-11
    e = m.field("e")
    m.set(e=e + 1005.0)                                                       # +1 K of dry static energy
    assert np.abs(m.field("T") - (T0 + 1.0)).max() < 1e-10


def test_argument_validation(oracle_arch):
    import breeze_b200 as bz
    with pytest.raises(ValueError):
        _model(oracle_arch, "Enthalpy")
    with pytest.raises(NotImplementedError):
        _model(oracle_arch, "StaticEnergy", microphysics=bz.SaturationAdjustment())
    m = _model(oracle_arch, "StaticEnergy")
    with pytest.raises(ValueError):
        m.set(θ=300.0, T=280.0)


def test_formulations_agree_on_a_rising_bubble(oracle_arch):
    """ρe and ρθ are two prognostic choices for the same adiabatic dynamics: after 10 steps the temperature fields agree to
    truncation error (≪ the 2 K bubble amplitude) and the momenta to ~1e-5 of their scale."""
    a, b = _model(oracle_arch, "LiquidIcePotentialTemperature"), _model(oracle_arch, "StaticEnergy")
    for m in (a, b):
        m.set(θ=bubble_theta(), u=1.0, qᵗ=0.003)
        for _ in range(10):
            m.time_step(2.0)
    assert np.abs(a.field("T") - b.field("T")).max() < 1e-3
    assert rel_err(a.field("ρw"), b.field("ρw")) < 1e-4
    assert np.abs(a.field("ρw")).max() > 0.1


@pytest.mark.gpu
@pytest.mark.parametrize("size,flat_y", [((32, 16, 24), False), ((64, 40), True)])
@pytest.mark.parametrize("z_chunks", [0, 3])
def test_cuda_static_energy_matches_oracle(oracle_arch, size, flat_y, z_chunks):
    """Tendencies (1e-11, same-form smoothness indicators) and 5 steps (1e-8) of the CUDA path against the oracle."""
    import breeze_b200 as bz
    import oracle_lib
    rng = np.random.default_rng(11)
    models = []
    for arch in (bz.B200(z_chunks=z_chunks), oracle_arch):
        kw = dict(x=(-10e3, 10e3), z=(0, 10e3))
        kw.update(dict(topology=(bz.Periodic, bz.Flat, bz.Bounded)) if flat_y else dict(y=(-10e3, 10e3)))
        grid = bz.RectilinearGrid(arch, size=size, **kw)
        models.append(bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)),
                                         formulation="StaticEnergy"))
    g = models[0].grid
    shp_c, shp_w = (g.Nz, g.Ny, g.Nx), (g.Nz + 1, g.Ny, g.Nx)
    u, w, q = 3.0 * rng.standard_normal(shp_c), rng.standard_normal(shp_w), 0.01 * rng.random(shp_c)
    for m in models:
        m.set(θ=bubble_theta(), u=u, w=w, qᵗ=q)
    gpu, cpu = models
    for name in ("ρe", "T", "e"):
        assert rel_err(gpu.field(name), cpu.field(name)) < 1e-13, name
    oracle_lib.set_beta_form(1)
    try:
        for m in models:
            m.context.compute_tendencies()
        for f in range(5):
            assert rel_err(gpu.context.get_tendency(f), cpu.context.get_tendency(f)) < 1e-11, f
    finally:
        oracle_lib.set_beta_form(0)
    for m in models:                 # five steps against the oracle in the reference (quadratic) form; the state carries grid-scale noise
        for _ in range(5):
            m.time_step(1.0)
    mom = max(np.abs(cpu.field(f)).max() for f in ("ρu", "ρv", "ρw"))
    for name in ("ρu", "ρv", "ρw", "ρe", "ρq", "T"):
        scale = mom if name in ("ρu", "ρv", "ρw") else np.abs(cpu.field(name)).max()
        err = np.abs(gpu.field(name) - cpu.field(name)).max() / scale
        report(err, name)
        assert err < 2e-8, name
