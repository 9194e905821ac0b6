"""WENO(order = 7 / 9) on both CUDA paths against the CPU oracle: the compressible slow-tendency kernels (c_slow_tendencies<4 / 5>,
c_moisture_tendency<4 / 5>) and the anelastic high-order stage kernel (stage_hi_kernel<4 / 5>, csrc/stage_hi.cuh) — the scheme the
reference's shipped examples run (examples/dry_thermal_bubble.jl:24, examples/bomex.jl:204, examples/splitting_supercell.jl:279).

Tolerances: the device reconstructions evaluate the smoothness indicators in the first differences of the stencil; a single tendency
evaluation is compared with the oracle switched to the same (algebraically identical) form at 1e-11; every multi-step comparison runs
against the oracle's default VALUE-form indicators (how the reference stores them) at 1e-7: measured ≤ 1.9e-8 on a B200 (ten steps of
the shipped Δθ = 10 K bubble; profiles/r2c_parity_errors_weno_high_order.txt). The order-9 value forms carry coefficients up to 200, so
their cancellation noise is ≈ 200 |ψ|² eps ≈ 4e-9 against β + ε — the two ORACLE forms differ from each other by as much
(tests/test_oracle_weno_high_order.py::test_beta_forms_agree_to_cancellation_noise)."""
import numpy as np
import pytest

from conftest import bubble_theta, rel_err, report

pytestmark = pytest.mark.gpu

TOL_TENDENCY = 1e-11
TOL_STEPS = 1e-7
PROGNOSTIC = ["ρ", "ρu", "ρv", "ρw", "ρθ"]


def _model(arch, size, order, flat_y=False):
    import breeze_b200 as bz
    if flat_y:
        grid = bz.RectilinearGrid(arch, size=size, x=(-5e3, 5e3), z=(0, 10e3), topology=(bz.Periodic, bz.Flat, bz.Bounded))
    else:
        grid = bz.RectilinearGrid(arch, size=size, x=(-5e3, 5e3), y=(-5e3, 5e3), z=(0, 10e3))
    dyn = bz.CompressibleDynamics(bz.SplitExplicitTimeDiscretization(substeps=6), reference_potential_temperature=300.0)
    return bz.AtmosphereModel(grid, dynamics=dyn, advection=bz.WENO(order=order))


def _pair(oracle_arch, size, order, flat_y=False, seed=0):
    import breeze_b200 as bz
    rng = np.random.default_rng(seed)
    models = [_model(a, size, order, flat_y) for a in (bz.B200(), oracle_arch)]
    g = models[0].grid
    shp_c, shp_w = (g.Nz, g.Ny, g.Nx), (g.Nz + 1, g.Ny, g.Nx)
    _, rho_r, _ = models[1].reference_profiles()
    rho = rho_r[:, None, None] * (1 + 1e-3 * rng.standard_normal(shp_c))
    u = 3.0 + rng.standard_normal(shp_c)
    v = (-2.0 + rng.standard_normal(shp_c)) * (0.0 if flat_y else 1.0)
    w = 0.5 * rng.standard_normal(shp_w)

    def theta(*xyz):
        x, z = xyz[0], xyz[-1]
        r2 = x ** 2 + (z - 3000.0) ** 2 + (xyz[1] ** 2 if len(xyz) == 3 else 0.0)
        return 300.0 + 2.0 * np.cos(np.pi / 2 * np.minimum(1.0, np.sqrt(r2) / 2000.0)) ** 2

    for m in models:
        m.set(ρ=rho, θ=theta, u=u, v=v, w=w)
    return models


@pytest.mark.parametrize("order", [7, 9])
@pytest.mark.parametrize("size,flat_y", [((32, 16, 24), False), ((40, 12, 14), False), ((64, 20), True)])
def test_slow_tendencies_match_oracle(oracle_arch, order, size, flat_y):
    from oracle_lib import set_beta_form
    gpu, cpu = _pair(oracle_arch, size, order, flat_y)
    set_beta_form(1)
    try:
        for m in (gpu, cpu):
            m.context.compute_slow_tendencies()
        for name in ["Gρ", "Gρu", "Gρv", "Gρw", "Gρθ", "Gˢρw"]:
            assert rel_err(gpu.field(name), cpu.field(name)) < 1e-11, name
    finally:
        set_beta_form(0)


@pytest.mark.parametrize("order", [7, 9])
def test_five_steps_match_oracle(oracle_arch, order):
    from oracle_lib import set_beta_form
    gpu, cpu = _pair(oracle_arch, (32, 16, 24), order, seed=3)
    for m in (gpu, cpu):             # oracle in the reference (value-form) indicators; the state carries grid-scale noise
        for _ in range(5):
            m.time_step(1.0)
    for name in PROGNOSTIC:
        assert rel_err(gpu.field(name), cpu.field(name)) < TOL_STEPS, name


# ---- anelastic path: stage_hi_kernel -------------------------------------------------------------------------------------------------
APROG = ["ρu", "ρv", "ρw", "ρθ", "ρq"]


def _anelastic(arch, size, order, flat_y=False, formulation="LiquidIcePotentialTemperature", microphysics=None, **arch_kw):
    import breeze_b200 as bz
    if flat_y:
        grid = bz.RectilinearGrid(arch, size=size, x=(-10e3, 10e3), z=(0, 10e3), topology=(bz.Periodic, bz.Flat, bz.Bounded))
    else:
        grid = bz.RectilinearGrid(arch, size=size, x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
    return bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)),
                              advection=bz.WENO(order=order), formulation=formulation, microphysics=microphysics)


def _anelastic_pair(oracle_arch, size, order, flat_y=False, seed=0, moist=True, **kw):
    import breeze_b200 as bz
    rng = np.random.default_rng(seed)
    arch_kw = {k: kw.pop(k) for k in ("z_chunks",) if k in kw}
    models = [_anelastic(a, size, order, flat_y, **kw) for a in (bz.B200(**arch_kw), oracle_arch)]
    g = models[0].grid
    shp_c, shp_w = (g.Nz, g.Ny, g.Nx), (g.Nz + 1, g.Ny, g.Nx)
    u = 3.0 * rng.standard_normal(shp_c)
    v = 2.0 * rng.standard_normal(shp_c) * (0.0 if flat_y else 1.0)
    w = rng.standard_normal(shp_w)
    q = 0.01 * rng.random(shp_c)
    for m in models:
        kws = dict(θ=bubble_theta(), u=u, v=v, w=w)
        if moist:
            kws["qᵗ"] = q
        m.set(**kws)
    return models


@pytest.mark.parametrize("order", [7, 9])
@pytest.mark.parametrize("size,flat_y,z_chunks", [((32, 16, 24), False, 0), ((64, 8, 30), False, 3), ((64, 40), True, 0), ((16, 8, 17), False, 1)])
def test_anelastic_tendencies_match_oracle(oracle_arch, order, size, flat_y, z_chunks):
    from oracle_lib import set_beta_form
    gpu, cpu = _anelastic_pair(oracle_arch, size, order, flat_y, z_chunks=z_chunks)
    for name in APROG + ["φ", "u", "w", "θ"]:                        # set! incl. the projection on the wider-halo layout
        assert rel_err(gpu.field(name), cpu.field(name)) < 1e-9, name
    gpu.context.compute_tendencies()
    set_beta_form(1)
    try:
        cpu.context.compute_tendencies()
    finally:
        set_beta_form(0)
    for name in APROG:
        assert rel_err(gpu.context.get_tendency(name), cpu.context.get_tendency(name)) < TOL_TENDENCY, name


@pytest.mark.parametrize("order", [7, 9])
@pytest.mark.parametrize("size,flat_y", [((32, 32, 32), False), ((128, 64), True)])
def test_anelastic_ten_steps_match_oracle(oracle_arch, order, size, flat_y):
    gpu, cpu = _anelastic(__import__("breeze_b200").B200(), size, order, flat_y), _anelastic(oracle_arch, size, order, flat_y)
    for m in (gpu, cpu):
        m.set(θ=bubble_theta(), u=1.0)
    for _ in range(10):
        gpu.time_step(2.0)
        cpu.time_step(2.0)
    for name in APROG + ["u", "w", "θ", "T"]:
        assert rel_err(gpu.field(name), cpu.field(name)) < TOL_STEPS, name
    assert gpu.context.max_abs_divergence() < 1e-10


def test_shipped_dry_thermal_bubble_configuration(oracle_arch):
    """examples/dry_thermal_bubble.jl as shipped: 2-D 128 x 128, WENO(order = 9), formulation = :StaticEnergy; 10 steps against the oracle."""
    import breeze_b200 as bz
    gpu = _anelastic(bz.B200(), (128, 128), 9, flat_y=True, formulation="StaticEnergy")
    cpu = _anelastic(oracle_arch, (128, 128), 9, flat_y=True, formulation="StaticEnergy")
    for m in (gpu, cpu):
        m.set(θ=bubble_theta(dtheta=10.0))
    for _ in range(10):
        gpu.time_step(2.0)
        cpu.time_step(2.0)
    for name in ("ρu", "ρw", "ρe", "T", "w"):
        assert rel_err(gpu.field(name), cpu.field(name)) < TOL_STEPS, name
    assert np.abs(gpu.field("w")).max() > 0.05


@pytest.mark.parametrize("order", [7, 9])
def test_anelastic_static_energy_and_saturation_adjustment(oracle_arch, order):
    """The other instantiations of the high-order stage kernel: StaticEnergy in 3-D with moisture, and warm-phase saturation adjustment with cloud."""
    import breeze_b200 as bz
    from oracle_lib import set_beta_form
    gpu, cpu = _anelastic_pair(oracle_arch, (32, 16, 24), order, formulation="StaticEnergy", seed=2)
    pairs = [(gpu, cpu)]
    rng = np.random.default_rng(11)
    ms = [_anelastic(a, (32, 16, 24), order, microphysics=bz.SaturationAdjustment()) for a in (bz.B200(), oracle_arch)]
    z = ms[0].grid.znodes()[:, None, None]
    qt = 0.016 * np.exp(-z / 2500.0) * (1 + 0.5 * rng.random((24, 16, 32)))
    uu, ww = rng.standard_normal((24, 16, 32)), 0.5 * rng.standard_normal((25, 16, 32))
    for m in ms:
        m.set(θ=bubble_theta(dtheta=3.0), qᵗ=qt, u=uu, w=ww)
    assert (ms[1].field("qˡ") > 0).mean() > 0.02
    pairs.append(tuple(ms))
    for g, c in pairs:
        g.context.compute_tendencies()
        set_beta_form(1)
        try:
            c.context.compute_tendencies()
        finally:
            set_beta_form(0)
        for f in range(5):
            assert rel_err(g.context.get_tendency(f), c.context.get_tendency(f)) < TOL_TENDENCY, f
        for _ in range(3):
            g.time_step(1.0)
            c.time_step(1.0)
        mom = max(np.abs(c.field(n)).max() for n in ("ρu", "ρv", "ρw"))
        for n in ("ρu", "ρv", "ρw"):
            err = np.abs(g.field(n) - c.field(n)).max() / mom
            report(err, n)
            assert err < TOL_STEPS, n
        for n in ("ρθ", "ρq", "T"):
            assert rel_err(g.field(n), c.field(n)) < TOL_STEPS, n


def test_bomex_as_shipped_weno9(oracle_arch):
    """examples/bomex.jl:204 runs WENO(order = 9): the FORCED + saturation-adjustment instantiation on a reduced grid with cloud, 3 steps."""
    import breeze_b200 as bz
    from oracle_lib import set_beta_form
    gpu = bz.cases.bomex_model(bz.B200(), size=(32, 16, 30), extent=3200.0, order=9, cloud=True)
    cpu = bz.cases.bomex_model(oracle_arch, size=(32, 16, 30), extent=3200.0, order=9, cloud=True)
    gpu.context.compute_tendencies()
    set_beta_form(1)
    try:
        cpu.context.compute_tendencies()
    finally:
        set_beta_form(0)
    for name in APROG:
        assert rel_err(gpu.context.get_tendency(name), cpu.context.get_tendency(name)) < TOL_TENDENCY, name
    for _ in range(3):
        gpu.time_step(2.0)
        cpu.time_step(2.0)
    mom = max(np.abs(cpu.field(n)).max() for n in ("ρu", "ρv", "ρw"))
    for n in ("ρu", "ρv", "ρw"):
        err = np.abs(gpu.field(n) - cpu.field(n)).max() / mom
        report(err, n)
        assert err < TOL_STEPS, n
    for n in ("ρθ", "ρq", "T", "qˡ"):
        assert rel_err(gpu.field(n), cpu.field(n)) < TOL_STEPS, n
