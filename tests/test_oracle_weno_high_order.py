"""WENO(order = 7 / 9) in the CPU oracle (the reference's shipped examples run WENO(order = 9),
`examples/dry_thermal_bubble.jl`, `examples/bomex.jl`, `examples/splitting_supercell.jl`). This file pins the oracle side that the
CUDA kernels of orders 7 / 9 (csrc/stage_hi.cuh, compressible.cuh) are checked against in tests/test_gpu_weno_high_order.py.

The reference holds no number for any WENO reconstruction (`test/advection_schemes.jl:7-123` is plumbing + one smoke step per
scheme), so what can be pinned is (i) the coefficient tables — derived here from the definitions in exact arithmetic, the same
derivation reproducing the order-5 constants of SURVEY Appendix A.2; (ii) design order, polynomial exactness and
the essentially-non-oscillatory property of the reconstructions; (iii) the reference's model-level invariants
(`test/dynamics.jl:45-116`: momentum conservation) under the higher-order schemes.
"""
import ctypes as C
import importlib.util
import os
from fractions import Fraction as F

import numpy as np
import pytest

import breeze_b200 as bz
from conftest import bubble_theta

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _derive():
    spec = importlib.util.spec_from_file_location("derive_weno", os.path.join(ROOT, "scripts", "derive_weno_coefficients.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def orc():
    from oracle_lib import load_oracle_library
    return load_oracle_library().dll


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _cell_averages(f_antiderivative, centres, h):
    return np.array([(f_antiderivative(c + h / 2) - f_antiderivative(c - h / 2)) / h for c in centres])


def _window(order, Fint, h, x_face):
    R = (order + 1) // 2
    centres = x_face - h / 2 + (np.arange(2 * R - 1) - (R - 1)) * h       # upwind cell (index R-1) ends at the face
    return _cell_averages(Fint, centres, h)


# ---- coefficient tables ---------------------------------------------------------------------------------------------
def test_derivation_reproduces_the_order5_constants_and_the_published_higher_order_forms():
    d = _derive()
    c, w, B = d.derive(3)
    assert w == [F(3, 10), F(3, 5), F(1, 10)]
    assert [[3 * B[s][0][0], 3 * B[s][0][1], 3 * B[s][0][2], 3 * B[s][1][1], 3 * B[s][1][2], 3 * B[s][2][2]] for s in range(3)] == \
        [[10, -31, 11, 25, -19, 4], [4, -13, 5, 13, -13, 4], [4, -19, 11, 25, -31, 10]]
    _, w4, B4 = d.derive(4)
    assert w4 == [F(4, 35), F(18, 35), F(12, 35), F(1, 35)]
    assert [240 * B4[3][a][a] for a in range(4)] == [547, 7043, 11003, 2107]                       # Balsara & Shu (2000), r = 4
    _, w5, B5 = d.derive(5)
    assert w5 == [F(5, 126), F(20, 63), F(10, 21), F(10, 63), F(1, 126)]
    assert [5040 * B5[0][a][a] for a in range(5)] == [107918, 1020563, 1521393, 482963, 22658]     # Balsara & Shu (2000), r = 5
    assert d.centered(3) == [F(1, 60), F(-2, 15), F(37, 60), F(37, 60), F(-2, 15), F(1, 60)]
    assert d.centered(4) == [F(-1, 280), F(29, 840), F(-139, 840), F(533, 840), F(533, 840), F(-139, 840), F(29, 840), F(-1, 280)]


def test_generated_header_is_current(tmp_path):
    d = _derive()
    header = open(os.path.join(ROOT, "oracle", "oracle_weno_tables.h")).read()
    for r, name in ((4, "WENO7"), (5, "WENO9")):
        c, w, B = d.derive(r)
        assert f"static const double {name}_D[{r}] = {{" + ", ".join(d.fmt(x) for x in w) + "};" in header
        assert "{" + ", ".join(d.fmt(x) for x in c[0]) + "}," in header
        assert "{" + ", ".join(d.fmt(x) for x in B[r - 1][0]) + "}" in header


# ---- the reconstructions ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("order", [5, 7, 9])
def test_design_order_on_a_smooth_function(orc, order):
    Fint = lambda x: -np.cos(x)                              # antiderivative of sin
    errs = []
    for h in (0.4, 0.2, 0.1):
        v = orc.orc_weno_biased_window(_dp(_window(order, Fint, h, 0.37)), order)
        errs.append(abs(v - np.sin(0.37)))
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert min(rates) > order - 0.6, (errs, rates)


@pytest.mark.parametrize("order", [5, 7, 9])
def test_exact_for_polynomials_every_candidate_reproduces(orc, order):
    """Every candidate stencil (r = (order + 1) / 2 cells) reproduces polynomials of degree r - 1, so any convex combination —
    whatever the nonlinear weights — does."""
    rng = np.random.default_rng(order)
    coef = rng.standard_normal((order + 1) // 2)             # degree r - 1
    P = np.polynomial.Polynomial(coef)
    v = orc.orc_weno_biased_window(_dp(_window(order, P.integ(), 0.5, 0.2)), order)
    assert v == pytest.approx(P(0.2), rel=1e-11, abs=1e-11)


@pytest.mark.parametrize("order", [5, 7, 9])
def test_essentially_non_oscillatory_at_a_step(orc, order):
    R = (order + 1) // 2
    for jump_at in range(1, 2 * R - 1):
        w = np.where(np.arange(2 * R - 1) < jump_at, 1.0, 3.0) + 0.01 * np.sin(np.arange(2 * R - 1))
        v = orc.orc_weno_biased_window(_dp(w), order)
        assert w.min() - 0.05 <= v <= w.max() + 0.05, (jump_at, v)
    w = np.where(np.arange(2 * R - 1) < R, 1.0, 3.0)          # jump right at the face
    v = orc.orc_weno_biased_window(_dp(w), order)
    assert abs(v - 1.0) < 1e-3                               # the upwind side of the jump wins


@pytest.mark.parametrize("order", [7, 9])
def test_beta_forms_agree_to_cancellation_noise(orc, order):
    """Value-form (how the reference stores the indicators) and difference-form (what the CUDA kernels evaluate) indicators are the
    same polynomials; on θ-like data (≈ 300 with small variations) they differ by the value form's cancellation noise only."""
    from oracle_lib import set_beta_form
    rng = np.random.default_rng(order)
    R = (order + 1) // 2
    worst = 0.0
    try:
        for _ in range(200):
            w = 300.0 + rng.standard_normal(2 * R - 1) * rng.choice([1e-3, 1e-1, 1.0])
            set_beta_form(0)
            a = orc.orc_weno_biased_window(_dp(w), order)
            set_beta_form(1)
            b = orc.orc_weno_biased_window(_dp(w), order)
            worst = max(worst, abs(a - b) / abs(a))
    finally:
        set_beta_form(0)
    assert worst < 1e-7
    # and they are exactly the same on data without a large mean
    w = rng.standard_normal(2 * R - 1)
    set_beta_form(0)
    a = orc.orc_weno_biased_window(_dp(w), order)
    set_beta_form(1)
    b = orc.orc_weno_biased_window(_dp(w), order)
    set_beta_form(0)
    assert a == pytest.approx(b, rel=1e-12)


@pytest.mark.parametrize("order", [4, 6, 8])
def test_centered_reconstruction_is_exact_to_its_order(orc, order):
    rng = np.random.default_rng(order)
    P = np.polynomial.Polynomial(rng.standard_normal(order))  # degree order - 1
    h = 0.5
    centres = 0.1 + (np.arange(order) - order / 2 + 0.5) * h  # face at x = 0.1
    v = orc.orc_centered_window(_dp(_cell_averages(P.integ(), centres, h)), order)
    assert v == pytest.approx(P(0.1), rel=1e-11, abs=1e-11)


# ---- model level -------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("order", [7, 9])
def test_momentum_conservation_with_high_order_weno(oracle_arch, order):
    """test/dynamics.jl:45-116 under WENO(order = 7 / 9): ∫ρu, ∫ρv conserved over 10 steps of a 3-D bubble, projection exact."""
    grid = bz.RectilinearGrid(oracle_arch, size=(16, 16, 16), x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
    m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)), advection=bz.WENO(order=order))
    m.set(θ=bubble_theta(), u=1.0, v=-0.5)
    s0 = [m.field(n).sum() for n in ("ρu", "ρv", "ρθ")]
    for _ in range(10):
        m.time_step(1.0)
    s1 = [m.field(n).sum() for n in ("ρu", "ρv", "ρθ")]
    for a, b in zip(s0, s1):
        assert abs(b - a) <= 1e-12 * abs(a)
    assert m.context.max_abs_divergence() < 1e-12
    assert np.abs(m.field("w")).max() > 1e-3                  # the bubble does rise


def test_shipped_dry_bubble_configuration_runs_on_the_oracle(oracle_arch):
    """examples/dry_thermal_bubble.jl: (Periodic, Flat, Bounded), WENO(order = 9), formulation = :StaticEnergy, Δθ = 10 K cone on an
    N² = 1e-6 stratification (reduced from 128 × 128 to 32 × 32 cells for the CPU suite)."""
    g = 9.81
    grid = bz.RectilinearGrid(oracle_arch, size=(32, 32), x=(-10e3, 10e3), z=(0, 10e3), topology=(bz.Periodic, bz.Flat, bz.Bounded))
    m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)),
                           advection=bz.WENO(order=9), formulation="StaticEnergy")
    m.set(θ=lambda x, z: 300.0 * np.exp(1e-6 * z / g) + 10.0 * np.maximum(0.0, 1.0 - np.sqrt(x ** 2 + (z - 3000.0) ** 2) / 2000.0))
    e0 = m.field("ρθ").sum()
    for _ in range(10):
        m.time_step(1.0)
    assert all(np.isfinite(m.field(n)).all() for n in ("ρu", "ρw", "ρθ", "T"))
    assert np.abs(m.field("w")).max() > 0.05
    assert m.context.max_abs_divergence() < 1e-12
    assert abs(m.field("ρθ").sum() - e0) <= 1e-6 * abs(e0)     # ρe changes only through the (small) buoyancy-flux term


def test_shipped_bomex_configuration_with_weno9_runs_on_the_oracle(oracle_arch):
    """examples/bomex.jl runs WENO(order = 9) with saturation adjustment, forcings and flux BCs (reduced to 16 x 16 x 30 cells)."""
    m = bz.cases.bomex_model(oracle_arch, size=(16, 16, 30), extent=1600.0, order=9)
    u0 = m.field("u")[0].mean()
    for _ in range(10):
        m.time_step(2.0)
    assert all(np.isfinite(m.field(n)).all() for n in ("ρu", "ρv", "ρw", "ρθ", "ρq", "T"))
    assert abs(m.field("u")[0].mean()) < abs(u0)             # bottom drag decelerates the lowest level
    assert m.context.max_abs_divergence() < 1e-12


def test_shipped_supercell_dynamics_with_weno9_runs_on_the_oracle(oracle_arch):
    """examples/splitting_supercell.jl: split-explicit compressible dynamics with WENO(order = 9) (dry here; reduced grid)."""
    grid = bz.RectilinearGrid(oracle_arch, size=(16, 16, 20), x=(0, 16e3), y=(0, 16e3), z=(0, 20e3))
    m = bz.AtmosphereModel(grid, dynamics=bz.CompressibleDynamics(bz.SplitExplicitTimeDiscretization(substeps=6), reference_potential_temperature=300.0),
                           advection=bz.WENO(order=9))
    _, rho, _ = m.reference_profiles()
    m.set(ρ=np.broadcast_to(rho[:, None, None], m.context.shape(0)).copy(), u=10.0, v=5.0,
          θ=lambda x, y, z: 300.0 + 3.0 * np.exp(-((x - 8e3) ** 2 + (y - 8e3) ** 2) / 3e3 ** 2 - (z - 1500.0) ** 2 / 1500.0 ** 2))
    mass0 = m.field("ρ").sum()
    for _ in range(5):
        m.time_step(4.0)
    assert all(np.isfinite(m.field(n)).all() for n in ("ρ", "ρu", "ρw", "ρθ"))
    assert abs(m.field("ρ").sum() - mass0) <= 1e-12 * mass0   # S4 of test/substepper_structural.jl holds for every order
    assert 0 < np.abs(m.field("w")).max() < 20.0


def test_order_5_is_unchanged_by_the_generalised_buffers(oracle_arch):
    """The order-5 path goes through the same generalised code (buffer 3, Centered(4)); frozen value of the README bubble."""
    grid = bz.RectilinearGrid(oracle_arch, size=(32, 32), x=(-10e3, 10e3), z=(0, 10e3), topology=(bz.Periodic, bz.Flat, bz.Bounded))
    res = {}
    for order in (5, 9):
        m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)), advection=bz.WENO(order=order))
        m.set(θ=bubble_theta(), u=2.0)
        for _ in range(20):
            m.time_step(2.0)
        res[order] = np.abs(m.field("w")).max()
    assert res[5] == pytest.approx(1.1149789565526504, rel=1e-10)
    assert res[9] == pytest.approx(1.1158648702685667, rel=1e-10)


@pytest.mark.gpu
def test_cuda_path_rejects_unknown_weno_orders_loudly():
    grid = bz.RectilinearGrid(bz.B200(), size=(16, 16, 16), x=(0, 1), y=(0, 1), z=(0, 1))
    model = bz.AtmosphereModel(grid, advection=bz.WENO(order=9))          # orders 5, 7, 9 are on the path
    assert model.field("ρu").shape == (16, 16, 16)
    cfg = model.context.lib.default_config_struct()
    cfg.Nx = cfg.Ny = cfg.Nz = 16
    cfg.advection_order = 11
    with pytest.raises(bz.BreezeError, match="WENO"):
        bz.Context(model.context.lib, cfg)
