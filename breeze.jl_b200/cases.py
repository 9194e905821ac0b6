"""Case definitions used by the tests and examples.

BOMEX (Siebesma et al. 2003, J. Atmos. Sci. 60, 1201-1219): the reference's examples/bomex.jl takes these piecewise-linear
profiles from AtmosphericProfilesLibrary.jl, which is not available here; they are restated from the paper.
"""
from __future__ import annotations

import numpy as np


def _pw(z, zs, vs):
    return float(np.interp(z, zs, vs))


def bomex_theta_liq_ice(z):
    return _pw(z, [0, 520, 1480, 2000, 3000], [298.7, 298.7, 302.4, 308.2, 311.85])


def bomex_q_tot(z):
    return 1e-3 * _pw(z, [0, 520, 1480, 2000, 3000], [17.0, 16.3, 10.7, 4.2, 3.0])


def bomex_u(z):
    return _pw(z, [0, 700, 3000], [-8.75, -8.75, -4.61])


def bomex_geostrophic_u(z):
    return -10.0 + 1.8e-3 * z


def bomex_geostrophic_v(z):
    return 0.0


def bomex_subsidence(z):
    return _pw(z, [0, 1500, 2100, 3000], [0.0, -0.65e-2, 0.0, 0.0])


def bomex_dqtdt(z):
    return _pw(z, [0, 300, 500, 3000], [-1.2e-8, -1.2e-8, 0.0, 0.0])


def bomex_dTdt(z):
    """Radiative cooling of θ (K/s): -2 K/day below 1500 m, linearly to 0 at 2500 m."""
    return _pw(z, [0, 1500, 2500, 3000], [-2.0 / 86400, -2.0 / 86400, 0.0, 0.0])


def bomex_model(arch, size=(64, 64, 75), extent=6400.0, seed=938, order=5, cloud=False):
    """examples/bomex.jl:42-243: grid, reference state, forcings, flux BCs, perturbed initial condition. The example runs
    WENO(order=9); order 5 is the default here.

    cloud=True seeds moist thermals in the cumulus layer (500-1500 m) that are super-saturated from the first step on, so that
    the saturation adjustment's secant iteration runs in a benchmark or parity test without the ≈ 30 simulated minutes of spin-up
    the case needs to form its first clouds."""
    import breeze_b200 as bz
    grid = bz.RectilinearGrid(arch, size=size, x=(0, extent), y=(0, extent), z=(0, 3000.0))
    constants = bz.ThermodynamicConstants()
    ref = bz.ReferenceState(grid, constants, surface_pressure=101500.0, potential_temperature=299.1)
    Rd = constants.molar_gas_constant / constants.dry_air_molar_mass
    rho0 = 101500.0 / (Rd * 299.1)                              # density(θ₀, p₀, q = 0) as in the example (:70)
    subsidence = bz.SubsidenceForcing(bomex_subsidence)
    geo = bz.geostrophic_forcings(bomex_geostrophic_u, bomex_geostrophic_v)
    forcing = {"u": (subsidence, geo["u"]), "v": (subsidence, geo["v"]), "θ": subsidence,
               "qᵉ": (subsidence, bz.Forcing(bomex_dqtdt)),
               "e": bz.Forcing(lambda z: constants.dry_air_heat_capacity * bomex_dTdt(z))}
    drag = bz.DragFluxBoundaryCondition(rho0, 0.28)
    bcs = {"ρθ": bz.FluxBoundaryCondition(rho0 * 8e-3), "ρqᵉ": bz.FluxBoundaryCondition(rho0 * 5.2e-5), "ρu": drag, "ρv": drag}
    model = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(ref), advection=bz.WENO(order=order),
                               microphysics=bz.SaturationAdjustment(), coriolis=bz.FPlane(f=3.76e-5), forcing=forcing,
                               boundary_conditions=bcs)
    rng = np.random.default_rng(seed)                            # the example seeds Julia's RNG (:26); not reproducible bit-wise
    z = grid.znodes()
    shape = (grid.Nz, grid.Ny, grid.Nx)
    pert = (z < 1600.0)[:, None, None]
    theta = np.array([bomex_theta_liq_ice(zz) for zz in z])[:, None, None] + 0.1 * (rng.random(shape) - 0.5) * pert
    qt = np.array([bomex_q_tot(zz) for zz in z])[:, None, None] + 2.5e-5 * (rng.random(shape) - 0.5) * pert
    if cloud:
        x, y = grid.xnodes()[None, None, :], grid.ynodes()[None, :, None]
        blobs = np.maximum(0.0, np.sin(8 * np.pi * x / extent) * np.sin(8 * np.pi * y / extent)) ** 2     # 4 x 4 cells, half of them moist
        layer = np.clip(1.0 - np.abs(z - 1000.0) / 500.0, 0.0, 1.0)[:, None, None]
        qt = qt + 5e-3 * blobs * layer
    u = np.broadcast_to(np.array([bomex_u(zz) for zz in z])[:, None, None], shape)
    i0 = getattr(model, "i0", 0)
    sl = slice(i0, i0 + model.Nx_local)
    model.set(θ=theta[:, :, sl], qᵗ=qt[:, :, sl], u=u[:, :, sl])
    return model
