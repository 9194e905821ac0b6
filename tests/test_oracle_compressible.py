"""CPU: the compressible (acoustic substepping) oracle pinned to the reference's own known-answer tests.

Each test cites the reference test it restates (paths relative to the reference repository):
  test/acoustic_substepping_components.jl — substep sequencing, the frozen horizontal pressure gradient, tridiagonal
                                            coefficients, adaptive substep counts
  test/substepper_rest_state.jl           — T1 discrete hydrostatic balance, T2 EoS/reference pressure, T3 slow vertical
                                            tendency at rest, T4 rest-atmosphere drift over a Δt sweep
  test/substepper_structural.jl           — S1 bottom tridiagonal row, S4 mass conservation, S5 top face
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import breeze_b200 as bz  # noqa: E402
from oracle_lib import CPUOracle, load_oracle_library  # noqa: E402

T0_REST, G_REST, CPD_REST = 250.0, 9.80665, 1005.0
RD = 8.314462618 / 0.02897
EPS = np.finfo(float).eps


def theta_isothermal(z):
    return T0_REST * np.exp(G_REST * z / (CPD_REST * T0_REST))


def rest_model(arch=None, Nx=8, Ny=8, Nz=32, Lz=10e3, Lh=100e3, **td):
    grid = bz.RectilinearGrid(arch or CPUOracle(), size=(Nx, Ny, Nz), x=(0, Lh), y=(0, Lh), z=(0, Lz))
    dyn = bz.CompressibleDynamics(bz.SplitExplicitTimeDiscretization(**td), reference_potential_temperature=theta_isothermal,
                                  surface_pressure=1e5, standard_pressure=1e5)
    return bz.AtmosphereModel(grid, dynamics=dyn)


def set_rest_state(m):
    """test/substepper_rest_state.jl:98-117: ρ ← ρ_ref, ρθ ← p_ref / (Rᵈ Π_ref), zero velocities."""
    p, rho, pi = m.reference_profiles()
    shape = m.context.shape(0)
    z3 = np.zeros(shape)
    m.context.set_state(rho=np.broadcast_to(rho[:, None, None], shape).copy(),
                        rho_theta=np.broadcast_to((p / (RD * pi))[:, None, None], shape).copy(),
                        rho_u=z3, rho_v=z3, rho_w=np.zeros(m.context.shape(3)))


def test_first_small_step_pressure_gradient_sequencing():
    """test/acoustic_substepping_components.jl:50-56"""
    f = load_oracle_library().dll.orcc_apply_horizontal_pressure_gradient_substep
    assert f(1, 1, 0) and not f(1, 2, 0) and f(2, 2, 0) and not f(1, 6, 0) and f(6, 6, 0)


def test_first_substep_retains_frozen_horizontal_pressure_gradient():
    """test/acoustic_substepping_components.jl:58-93: p = 2x + 3y, Δτ = 0.5 ⇒ ρu′ = -1, ρv′ = -1.5 exactly."""
    grid = bz.RectilinearGrid(CPUOracle(), size=(4, 4, 4), x=(0, 4), y=(0, 4), z=(0, 4))
    m = bz.AtmosphereModel(grid, dynamics=bz.CompressibleDynamics(reference_state=None))
    ctx = m.context
    X, Y = grid.xnodes()[None, None, :], grid.ynodes()[None, :, None]
    p = np.ascontiguousarray(np.broadcast_to(2 * X + 3 * Y, ctx.shape(0)))
    ones = np.ones(ctx.shape(0))
    dll = load_oracle_library().dll
    dp = C.POINTER(C.c_double)
    dll.orcc_test_explicit_horizontal_step.argtypes = [C.c_void_p, dp, dp, dp, dp, C.c_double, C.c_int]
    rc = dll.orcc_test_explicit_horizontal_step(ctx.handle, p.ctypes.data_as(dp), None, ones.ctypes.data_as(dp),
                                                ones.ctypes.data_as(dp), 0.5, 0)
    assert rc == 0
    assert ctx.get_field("ρu′")[1, 1, 1] == -1
    assert ctx.get_field("ρv′")[1, 1, 1] == -1.5


def test_acoustic_vertical_tridiagonal_coefficients():
    """test/acoustic_substepping_components.jl:95-166"""
    Nz, Lz = 5, 1000.0
    k1 = np.arange(1, Nz + 1)
    Pi, th, gR = 0.90 + 0.02 * k1, 280.0 + 3 * k1, 390.0 + 5 * k1
    dtm, dm, g, dz = 0.7, 0.03, 9.81, Lz / Nz
    Cc = lambda k: gR[k - 1] * Pi[k - 1]                                    # noqa: E731  (1-based k)
    thf = lambda k: th[0] if k == 1 else (th[Nz - 1] if k == Nz + 1 else (th[k - 1] + th[k - 2]) / 2)   # noqa: E731
    dll = load_oracle_library().dll
    dp = C.POINTER(C.c_double)
    dll.orcc_test_tridiagonal_coefficients.argtypes = [C.c_int, C.c_double, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_int, dp]

    def coeff(k):
        out = np.zeros(3)
        assert dll.orcc_test_tridiagonal_coefficients(Nz, dz, Pi.ctypes.data_as(dp), th.ctypes.data_as(dp), gR.ctypes.data_as(dp),
                                                      g, dtm, dm, k, out.ctypes.data_as(dp)) == 0
        return out

    assert coeff(1)[1] == 1 and coeff(1)[2] == 0                            # also substepper_structural.jl S1 (:87-110)
    for k in range(2, Nz + 1):
        lower, diag, upper = coeff(k)
        e_lower = -dtm ** 2 * Cc(k - 1) * thf(k - 1) / dz ** 2 + dtm ** 2 * g / (2 * dz) - dm / dz ** 2
        e_diag = 1 + dtm ** 2 * thf(k) * (Cc(k) + Cc(k - 1)) / dz ** 2 + 2 * dm / dz ** 2
        assert lower == pytest.approx(e_lower, rel=1e-13)
        assert diag == pytest.approx(e_diag, rel=1e-13)
        if k <= Nz - 1:
            e_upper = -dtm ** 2 * Cc(k) * thf(k + 1) / dz ** 2 - dtm ** 2 * g / (2 * dz) - dm / dz ** 2
            assert upper == pytest.approx(e_upper, rel=1e-13)


@pytest.mark.parametrize("cfl,size,topo", [(0.5, (100, 6, 10), None), (0.25, (100, 6, 10), None), (1.0, (100, 6, 10), None),
                                           (0.5, (100, 10), (bz.Periodic, bz.Flat, bz.Bounded))])
def test_compute_acoustic_substeps(cfl, size, topo):
    """test/acoustic_substepping_components.jl:269-316: N = ⌈Δt ℂᵃᶜ / (ν Δx)⌉ with ℂᵃᶜ = √(γᵈ Rᵈ 300)."""
    kw = dict(x=(0, 100e3), z=(0, 10e3))
    if topo is None:
        kw["y"] = (0, 6e3)
    else:
        kw["topology"] = topo
    grid = bz.RectilinearGrid(CPUOracle(), size=size, **kw)
    td = bz.SplitExplicitTimeDiscretization(acoustic_cfl=cfl)
    m = bz.AtmosphereModel(grid, dynamics=bz.CompressibleDynamics(td))
    gam = 1005.0 / (1005.0 - RD)
    expected = int(np.ceil(12 * np.sqrt(gam * RD * 300) / (cfl * 1000)))
    assert int(np.ceil(12 * np.sqrt(1.4 * 287.0 * 300) / (cfl * 1000))) == expected
    n_fwd, dtau = m.context.stage_substep_count_and_size(12.0, 1.0)
    n_bwd, _ = m.context.stage_substep_count_and_size(-12.0, 1.0)
    assert n_fwd == expected == n_bwd and dtau == pytest.approx(12.0 / expected)


def test_stage_substep_distributions():
    """acoustic_substepping.jl:476-508: ProportionalSubsteps ⌈βN⌉ with Δτ = βΔt/Nτ; ConstantSubstepSize: N rounded to 6."""
    grid = bz.RectilinearGrid(CPUOracle(), size=(8, 8, 8), x=(0, 8e3), y=(0, 8e3), z=(0, 8e3))
    m = bz.AtmosphereModel(grid, dynamics=bz.CompressibleDynamics(bz.SplitExplicitTimeDiscretization(substeps=6)))
    assert [m.context.stage_substep_count_and_size(6.0, b)[0] for b in (1 / 3, 1 / 2, 1)] == [2, 3, 6]
    assert m.context.stage_substep_count_and_size(6.0, 0.5)[1] == pytest.approx(1.0)
    m = bz.AtmosphereModel(grid, dynamics=bz.CompressibleDynamics(
        bz.SplitExplicitTimeDiscretization(substeps=8, substep_distribution=bz.ConstantSubstepSize())))
    assert [m.context.stage_substep_count_and_size(6.0, b) for b in (1 / 3, 1 / 2, 1)] == [(4, 0.5), (6, 0.5), (12, 0.5)]
    m = bz.AtmosphereModel(grid, dynamics=bz.CompressibleDynamics(
        bz.SplitExplicitTimeDiscretization(substeps=8, substep_distribution=bz.MonolithicFirstStage())))
    assert m.context.stage_substep_count_and_size(6.0, 1 / 3) == (1, 2.0)


def test_T1_reference_state_discrete_hydrostatic_balance():
    """test/substepper_rest_state.jl:159-171: max |δz p_ref + g ℑz ρ_ref| <= 1e-9."""
    m = rest_model(Nx=16, Ny=16, Nz=64, Lz=30e3)
    p, rho, _ = m.reference_profiles()
    res = (p[1:] - p[:-1]) / m.grid.Δz + 9.81 * (rho[1:] + rho[:-1]) / 2
    assert np.abs(res).max() <= 1e-9


def test_T2_T3_rest_state_pressure_and_slow_tendency():
    """test/substepper_rest_state.jl:183-215: EoS pressure within 100 ulp of p_ref, ρ exact; Gˢρw <= 1e-12 at rest."""
    m = rest_model(Nx=16, Ny=16, Nz=64, Lz=30e3)
    set_rest_state(m)
    p, rho, _ = m.reference_profiles()
    assert np.abs(m.field("p") - p[:, None, None]).max() <= 100 * EPS * p.max()
    assert np.abs(m.field("ρ") - rho[:, None, None]).max() == 0
    m.context.compute_slow_tendencies()
    assert np.abs(m.field("Gˢρw")).max() <= 1e-12


@pytest.mark.parametrize("dt,td", [(0.5, {}), (20.0, {}), (20.0, dict(forward_weight=0.55, damping=bz.NoDivergenceDamping()))])
def test_T4_rest_atmosphere_drift(dt, td):
    """test/substepper_rest_state.jl:263-303: max|w| <= 1e-10 m/s over 200 outer steps for Δt ∈ {0.5, 20} s."""
    m = rest_model(**td)
    set_rest_state(m)
    envelope = 0.0
    for n in range(1, 201):
        m.time_step(dt)
        if n % 10 == 0:
            w = np.abs(m.field("w")).max()
            assert np.isfinite(w)
            envelope = max(envelope, w)
    assert envelope <= 1e-10


def bubble_model(arch=None, size=(16, 16, 16), substeps=None):
    grid = bz.RectilinearGrid(arch or CPUOracle(), size=size, x=(-5e3, 5e3), y=(-5e3, 5e3), z=(0, 10e3))
    dyn = bz.CompressibleDynamics(bz.SplitExplicitTimeDiscretization(substeps=substeps), reference_potential_temperature=300.0)
    m = bz.AtmosphereModel(grid, dynamics=dyn)
    p, rho, pi = m.reference_profiles()

    def theta(x, y, z):
        r = np.sqrt(x ** 2 + y ** 2 + (z - 3000.0) ** 2)
        return 300.0 + 2.0 * np.cos(np.pi / 2 * np.minimum(1.0, r / 2000.0)) ** 2

    m.set(ρ=np.broadcast_to(rho[:, None, None], m.context.shape(0)).copy(), θ=theta, u=2.0, v=-1.0)
    return m


def test_S4_S5_mass_conservation_and_top_face():
    """test/substepper_structural.jl:211-247: total mass conserved to 1e-12 over a step; ρw = 0 on the top face."""
    m = bubble_model()
    M0 = m.field("ρ").sum()
    for _ in range(3):
        m.time_step(2.0)
    assert abs(m.field("ρ").sum() - M0) / M0 <= 1e-12
    rw = m.field("ρw")
    assert np.abs(rw[-1]).max() <= 1e-12 and np.abs(rw[0]).max() == 0


def test_warm_bubble_rises_and_stays_symmetric():
    """A warm bubble in a resting isentropic atmosphere accelerates upward; with u = v = 0 the solution keeps the x/y mirror
    symmetry of the initial condition (each WENO bias is mirrored) and the horizontal momentum sums to zero."""
    grid = bz.RectilinearGrid(CPUOracle(), size=(16, 16, 16), x=(-5e3, 5e3), y=(-5e3, 5e3), z=(0, 10e3))
    m = bz.AtmosphereModel(grid, dynamics=bz.CompressibleDynamics(reference_potential_temperature=300.0))
    _, rho, _ = m.reference_profiles()
    m.set(ρ=np.broadcast_to(rho[:, None, None], m.context.shape(0)).copy(),
          θ=lambda x, y, z: 300.0 + 2.0 * np.cos(np.pi / 2 * np.minimum(1.0, np.sqrt(x ** 2 + y ** 2 + (z - 3000.0) ** 2) / 2000.0)) ** 2)
    for _ in range(10):
        m.time_step(2.0)
    w = m.field("w")
    assert 0.05 < w.max() < 5.0 and np.isfinite(w).all()
    k, j, i = np.unravel_index(np.argmax(w), w.shape)
    assert 5 <= j <= 10 and 5 <= i <= 10
    th = m.field("θ")
    assert np.abs(th - th[:, :, ::-1]).max() < 1e-9 and np.abs(th - th[:, ::-1, :]).max() < 1e-9
    assert abs(m.field("ρu").sum()) < 1e-9 * np.abs(m.field("ρw")).sum()


# ---- moisture (vapour only, microphysics = nothing) --------------------------------------------------------------------------
def moist_bubble(arch=None, q0=0.01, size=(16, 16, 16), explicit_zero=False):
    grid = bz.RectilinearGrid(arch or CPUOracle(), size=size, x=(-5e3, 5e3), y=(-5e3, 5e3), z=(0, 10e3))
    m = bz.AtmosphereModel(grid, dynamics=bz.CompressibleDynamics(bz.SplitExplicitTimeDiscretization(substeps=6),
                                                                  reference_potential_temperature=300.0))
    _, rho, _ = m.reference_profiles()
    kw = dict(ρ=np.broadcast_to(rho[:, None, None], m.context.shape(0)).copy(), u=2.0, v=-1.0,
              θ=lambda x, y, z: 300.0 + 2.0 * np.cos(np.pi / 2 * np.minimum(1.0, np.sqrt(x ** 2 + y ** 2 + (z - 3000.0) ** 2) / 2000.0)) ** 2)
    if q0 or explicit_zero:
        kw["qᵛ"] = lambda x, y, z: q0 * np.exp(-z / 2500.0) * (1 + 0.2 * np.sin(2 * np.pi * x / 10e3)) + 0 * y
    m.set(**kw)
    return m


def test_set_splits_total_density_and_mixture_eos():
    """establish_densities! (compressible_time_stepping.jl:89-137): `ρ` given ⇒ ρᵈ = ρ − ρ qᵛ; `ρᵈ` given ⇒ ρ = ρᵈ / (1 − qᵛ);
    p = ρ Rᵐ T with the mixture gas constant and the TOTAL density (:215-235)."""
    m = moist_bubble(q0=0.012)
    rho_t, rho_d, rqv, qv = m.field("ρᵗ"), m.field("ρ"), m.field("ρqᵛ"), m.field("qᵛ")
    _, rho_ref, _ = m.reference_profiles()
    assert np.abs(rho_t - rho_ref[:, None, None]).max() < 1e-15 * rho_ref.max() * 4
    assert np.abs(rho_d + rqv - rho_t).max() < 1e-15 and np.abs(rqv / rho_t - qv).max() < 1e-17
    Rm = (1 - qv) * RD + qv * (8.314462618 / 0.018015)
    assert np.abs(m.field("p") - rho_t * Rm * m.field("T")).max() < 1e-9
    m2 = moist_bubble(q0=0.0)
    m2.set(ρᵈ=rho_d, qᵛ=qv)
    assert np.abs(m2.field("ρᵗ") - rho_t).max() < 1e-15 and np.abs(m2.field("ρqᵛ") - rqv).max() < 1e-17


def test_zero_moisture_reproduces_the_dry_path_bit_for_bit():
    """qᵛ ≡ 0 through the moist code path: γᵐRᵐ, the mixture EOS and ρ = ρᵈ + 0 collapse to the dry values exactly
    (acoustic_substepping.jl:411-413)."""
    dry, wet = moist_bubble(q0=0.0), moist_bubble(q0=0.0, explicit_zero=True)
    for m in (dry, wet):
        for _ in range(3):
            m.time_step(3.0)
    for name in ("ρ", "ρu", "ρv", "ρw", "ρθ", "T", "p"):
        assert np.array_equal(dry.field(name), wet.field(name)), name
    assert np.abs(wet.field("ρqᵛ")).max() == 0


def test_vapour_mass_is_conserved_and_moves_with_the_flow():
    m = moist_bubble(q0=0.012)
    M0, D0 = m.field("ρqᵛ").sum(), m.field("ρ").sum()
    q_start = m.field("ρqᵛ").copy()
    for _ in range(5):
        m.time_step(3.0)
    assert abs(m.field("ρqᵛ").sum() - M0) / M0 < 1e-13
    assert abs(m.field("ρ").sum() - D0) / D0 < 1e-13
    assert np.abs(m.field("ρqᵛ") - q_start).max() > 1e-6                     # u = 2 m/s carries the x-modulated vapour
    assert np.isfinite(m.field("w")).all() and m.field("qᵛ").min() > 0


# ---- UpperSponge --------------------------------------------------------------------------------------------------------------
def test_upper_sponge_coefficients():
    """test/acoustic_substepping_components.jl:476-499: LinearRamp, rate 0.2, depth 2000 on z ∈ [0, 8000], Nz = 8."""
    from breeze_b200 import compressible
    lib = load_oracle_library()
    cfg = compressible.bzc_config()
    lib.dll.orcc_default_config(C.byref(cfg))
    cfg.base.Nx = cfg.base.Ny = cfg.base.Nz = 8
    cfg.base.z0, cfg.base.z1 = 0.0, 8000.0
    cfg.sponge, cfg.sponge_damping_rate, cfg.sponge_depth = compressible.BZC_SPONGE_LINEAR_RAMP, 0.2, 2000.0
    diag, rhs = lib.dll.orcc_test_sponge_term_diag, lib.dll.orcc_test_sponge_rhs
    diag.restype = rhs.restype = C.c_double
    diag.argtypes = [C.POINTER(compressible.bzc_config), C.c_int, C.c_double]
    rhs.argtypes = [C.POINTER(compressible.bzc_config), C.c_int, C.c_double, C.c_double]
    dtm, dts = 3.0, 2.0
    assert diag(C.byref(cfg), 1, dtm) == 0
    assert diag(C.byref(cfg), 9, dtm) == pytest.approx(dtm * 0.2)
    assert rhs(C.byref(cfg), 9, dts, 4.0) == pytest.approx(dts * 0.2 * 4.0)
    assert diag(C.byref(cfg), 8, dtm) == pytest.approx(dtm * 0.2 * 0.5)          # half-way up the ramp
    cfg.sponge = compressible.BZC_SPONGE_NONE
    assert diag(C.byref(cfg), 9, dtm) == 0 and rhs(C.byref(cfg), 9, dts, 4.0) == 0
    cfg.sponge = compressible.BZC_SPONGE_CUBIC_RAMP
    assert diag(C.byref(cfg), 8, dtm) == pytest.approx(dtm * 0.2 * (0.25 * (3 - 1.0)))
    cfg.sponge = compressible.BZC_SPONGE_SIN2_RAMP
    assert diag(C.byref(cfg), 8, dtm) == pytest.approx(dtm * 0.2 * np.sin(np.pi / 4) ** 2)


def test_upper_sponge_damps_the_vertical_momentum_perturbation_below_the_lid():
    """The sponge is a Rayleigh term on the acoustic PERTURBATION (ρw)′ = ρw − ρwᴸ (acoustic_substepping.jl:584-590): within the layer it
    shrinks the change of ρw over a step; below the layer the solution is untouched to first order."""
    def run(sponge):
        grid = bz.RectilinearGrid(CPUOracle(), size=(16, 16, 16), x=(-5e3, 5e3), y=(-5e3, 5e3), z=(0, 10e3))
        td = bz.SplitExplicitTimeDiscretization(substeps=6, sponge=sponge)
        m = bz.AtmosphereModel(grid, dynamics=bz.CompressibleDynamics(td, reference_potential_temperature=300.0))
        _, rho, _ = m.reference_profiles()
        m.set(ρ=np.broadcast_to(rho[:, None, None], m.context.shape(0)).copy(), θ=300.0,
              w=lambda x, y, z: 0.5 * np.sin(np.pi * z / 10e3) * np.cos(2 * np.pi * x / 10e3) + 0 * y)
        start = m.field("ρw").copy()
        m.time_step(3.0)
        return m.field("ρw") - start
    free, damped = run(None), run(bz.UpperSponge(damping_rate=0.3, depth=4000.0, ramp=bz.Sin2Ramp()))
    low = slice(2, 6)
    for k in (12, 13, 14, 15):                                    # the ramp grows towards the lid: 3 … 12 % less change per step
        assert np.abs(damped[k]).max() < 0.98 * np.abs(free[k]).max(), k
    assert np.abs(damped[15]).max() < 0.92 * np.abs(free[15]).max()
    assert np.abs(damped[low] - free[low]).max() < 0.05 * np.abs(free[low]).max()
    with pytest.raises(ValueError):
        bz.SplitExplicitTimeDiscretization(sponge="strong")
    with pytest.raises(ValueError):
        bz.UpperSponge(ramp="cubic")
