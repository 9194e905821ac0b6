"""The reference's short stability / cross-dynamics checks of the split-explicit path, restated on both hot-path families
(test/acoustic_substepping_stability.jl):

  * `:121-180`  SK94 inertia–gravity-wave case at the advection-limited Δt = 12 s (Ns = 8, divergence damping 0.10, 20 steps):
                no NaN, max |w| < 1 m/s, ρ > 0;
  * `:183-326`  tiny dry thermal bubble (16 × 16 cells, one Δt = 0.5 s): the split-explicit compressible response and the
                anelastic response have the same buoyant scale (max w within rtol 1.25) and the same updraft centroid
                (within 2 Δz). The reference's third arm, explicit compressible time stepping, is not part of the hot path
                 and is not built;
  * `:332-355`  a balanced atmosphere stays quiet: max |w| < sqrt(eps) after 10 steps of Δt = 12 s.

Each check runs on the CPU oracle (here) and through the C ABI on the GPU (`-m gpu`), with the reference's own thresholds.
"""
import numpy as np
import pytest

G = 9.81
RD = 8.314462618 / 0.02897
CPD = 1005.0
KAPPA = RD / CPD


def _igw_model(arch, Ns=8, kd=0.10, size=(100, 6, 10)):
    import breeze_b200 as bz
    Lx, Ly, Lz = 100e3, 6e3, 10e3
    grid = bz.RectilinearGrid(arch, size=size, x=(0, Lx), y=(0, Ly), z=(0, Lz))
    N2 = 0.01 ** 2

    def theta_bg(z):
        return 300.0 * np.exp(N2 * z / G)

    td = bz.SplitExplicitTimeDiscretization(substeps=Ns, damping=bz.ThermalDivergenceDamping(coefficient=kd))
    dyn = bz.CompressibleDynamics(td, surface_pressure=100000.0, reference_potential_temperature=theta_bg)
    m = bz.AtmosphereModel(grid, dynamics=dyn)
    _, rho_r, _ = m.reference_profiles()
    m.set(θ=lambda x, y, z: theta_bg(z) + 0.01 * np.sin(np.pi * z / Lz) / (1 + (x - Lx / 3) ** 2 / 5000.0 ** 2),
          u=20.0, ρ=np.broadcast_to(rho_r[:, None, None], m.context.shape(0)).copy())
    return m


def _check_igw(arch, size=(100, 6, 10)):
    m = _igw_model(arch, size=size)
    for _ in range(20):
        m.time_step(12.0)
    rho, rw = m.field("ρ"), m.field("ρw")
    assert np.all(np.isfinite(rho)) and np.all(np.isfinite(rw))
    w = m.field("w")
    assert np.max(np.abs(w)) < 1.0
    assert np.max(np.abs(w)) > 0.0            # the perturbation does set the wave off
    assert rho.min() > 0.0
    return float(np.max(np.abs(w)))


def _tiny_bubble(arch, kind):
    import breeze_b200 as bz
    grid = bz.RectilinearGrid(arch, size=(16, 16), x=(-8e3, 8e3), z=(0, 8e3), topology=(bz.Periodic, bz.Flat, bz.Bounded))
    p0 = pst = 100000.0
    th0 = 300.0

    def exner(z):
        return (p0 / pst) ** KAPPA - G * z / (CPD * th0)

    def theta(x, z):
        return th0 + 10.0 * np.maximum(0.0, 1.0 - np.sqrt(x ** 2 + (z - 3000.0) ** 2) / 2000.0)

    if kind == "anelastic":
        ref = bz.ReferenceState(grid, surface_pressure=p0, potential_temperature=th0, standard_pressure=pst)
        m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(ref), advection=bz.WENO(order=5))
        m.set(θ=theta)
    else:
        dyn = bz.CompressibleDynamics(bz.SplitExplicitTimeDiscretization(substeps=6), surface_pressure=p0, standard_pressure=pst,
                                      reference_potential_temperature=th0)
        m = bz.AtmosphereModel(grid, dynamics=dyn)
        m.set(θ=theta, ρ=lambda x, z: pst * exner(z) ** (1.0 / KAPPA) / (RD * theta(x, z) * exner(z)))
    return m


def _updraft(m):
    w = np.asarray(m.field("w"))                          # (Nz [+1], Ny, Nx), level k = bottom face of cell k
    g = m.grid
    pos = np.maximum(0.0, w)
    zf = g.z0 + np.arange(w.shape[0]) * (g.z1 - g.z0) / g.Nz
    return pos.max(), float((pos.sum(axis=(1, 2)) * zf).sum() / pos.sum())


def _check_tiny_bubble(arch):
    split, anel = _tiny_bubble(arch, "split_explicit"), _tiny_bubble(arch, "anelastic")
    split.time_step(0.5)
    anel.time_step(0.5)
    for m in (split, anel):
        assert np.all(np.isfinite(m.field("w")))
    (ws, zs), (wa, za) = _updraft(split), _updraft(anel)
    assert ws > 0 and wa > 0
    assert abs(ws - wa) <= 1.25 * max(ws, wa)             # isapprox(split.max_w, anelastic.max_w; rtol = 1.25)
    dz = 8e3 / 16
    assert abs(zs - za) <= 2 * dz
    return ws, wa, zs, za


def _check_quiet(arch):
    import breeze_b200 as bz
    grid = bz.RectilinearGrid(arch, size=(16, 8, 10), x=(0, 16e3), y=(0, 8e3), z=(0, 10e3))
    dyn = bz.CompressibleDynamics(bz.SplitExplicitTimeDiscretization(substeps=8), surface_pressure=100000.0,
                                  reference_potential_temperature=300.0)
    m = bz.AtmosphereModel(grid, dynamics=dyn)
    _, rho_r, _ = m.reference_profiles()
    m.set(θ=300.0, u=0.0, ρ=np.broadcast_to(rho_r[:, None, None], m.context.shape(0)).copy())
    for _ in range(10):
        m.time_step(12.0)
    assert np.max(np.abs(m.field("w"))) < np.sqrt(np.finfo(np.float64).eps)


# ---- CPU oracle -----------------------------------------------------------------------------------------------------
def test_oracle_igw_stays_bounded(oracle_arch):
    _check_igw(oracle_arch)


def test_oracle_tiny_bubble_split_explicit_matches_anelastic_scale(oracle_arch):
    ws, wa, zs, za = _check_tiny_bubble(oracle_arch)
    # the two responses are in fact much closer than the reference's wiring thresholds; keep a tighter regression band
    assert abs(zs - za) <= 600.0 and 0.2 < ws / wa < 5.0
    # frozen oracle values of this case (regression guard for the oracle itself; generated by this very test on the oracle of round 1)
    assert (ws, wa) == pytest.approx((0.11671592213357988, 0.049573037547965534), rel=1e-9)
    assert (zs, za) == pytest.approx((3004.8913660376206, 3430.053537683285), rel=1e-9)


def test_oracle_balanced_state_stays_quiet(oracle_arch):
    _check_quiet(oracle_arch)


# ---- CUDA path --------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("size", [(100, 6, 10), (128, 8, 10)])      # the reference's grid (not a multiple of the tile sizes), and a tiled one
def test_gpu_igw_stays_bounded(size, oracle_arch):
    import breeze_b200 as bz
    m = _check_igw(bz.B200(), size)
    o = _check_igw(oracle_arch, size)
    assert abs(m - o) <= 1e-6 * o                                    # max |w| after the 20 steps agrees with the oracle's


@pytest.mark.gpu
def test_gpu_tiny_bubble_split_explicit_matches_anelastic_scale(oracle_arch):
    import breeze_b200 as bz
    ws, wa, zs, za = _check_tiny_bubble(bz.B200())
    ows, owa, ozs, oza = _check_tiny_bubble(oracle_arch)
    # and the CUDA diagnostics agree with the oracle's
    assert abs(ws - ows) <= 1e-8 * ows and abs(wa - owa) <= 1e-8 * owa
    assert abs(zs - ozs) <= 1e-4 and abs(za - oza) <= 1e-4


@pytest.mark.gpu
def test_gpu_balanced_state_stays_quiet():
    import breeze_b200 as bz
    _check_quiet(bz.B200())
