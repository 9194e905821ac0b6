"""Generates the committed fixtures under tests/golden/.

1. reference_known_answers.json — every number the reference itself pins for this path (doctests / test files), copied with
   its file:line. The reference (Julia) cannot run in this project, so these are transcribed, not generated.
2. oracle_bubble_16x8x12.npz, oracle_bomex_16x8x12.npz, oracle_compressible_16x8x12.npz — the CPU oracle's state after a few
   steps of small seeded cases (anelastic bubble, BOMEX-type forcing, compressible WS-RK3 with acoustic substepping).
   They freeze the oracle (so an accidental change of the checker is caught on CPU) and give the GPU tests a committed target
   that does not depend on re-running the oracle.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

KNOWN = {
    "saturation_specific_humidity_liquid_288K_101325Pa": {"value": 0.010359995391195264, "source": "src/Thermodynamics/vapor_saturation.jl:58-74"},
    "saturation_specific_humidity_ice_288K_101325Pa": {"value": 0.011945100768555072, "source": "src/Thermodynamics/vapor_saturation.jl:76-81"},
    "saturation_specific_humidity_mixed40_288K_101325Pa": {"value": 0.01128386068542303, "source": "src/Thermodynamics/vapor_saturation.jl:83-91"},
    "pressure_balanced_density_1_300_303": {"value": 0.9900990099009901, "source": "src/Thermodynamics/reference_states.jl:140-151"},
    "analytic_column_phi": {"formula": "z^2/2 - z^3/3 - 1/12 (mean removed), rho_r = z, rho_w = z^2 - z^3, Nz = 48, rtol 1e-3",
                            "source": "test/anelastic_pressure_solver_analytic.jl:9-51"},
    "projection_divergence_bound": {"formula": "max|div| < Nx*Ny*Nz*eps for 32^3 random momentum, rho_r = 1",
                                    "source": "test/anelastic_pressure_solver_nonhydrostatic.jl:40-46"},
    "secant_sqrt2": {"value": 1.4142135624, "source": "src/Solvers.jl:225-241"},
    "explicit_horizontal_step_frozen_pgf": {"formula": "p = 2x + 3y, dtau = 0.5, perturbation PGF gated off => rho_u' = -1, rho_v' = -1.5 exactly",
                                            "source": "test/acoustic_substepping_components.jl:58-93"},
    "rest_state_contracts": {"formula": "hydrostatic residual <= 1e-9; |p - p_ref| <= 100 ulp; |Gs_rho_w| <= 1e-12; max|w| <= 1e-10 over 200 steps at dt in {0.5, 20}",
                             "source": "test/substepper_rest_state.jl:159-303"},
    "acoustic_substeps_dx1km_dt12": {"formula": "ceil(12*sqrt(1.4*287*300)/(0.5*1000))", "source": "test/acoustic_substepping_components.jl:269-316"},
}


def bubble_case(arch):
    from conftest import bubble_theta, make_bubble_model
    m = make_bubble_model(arch, (16, 8, 12))
    rng = np.random.default_rng(20261017)
    g = m.grid
    m.set(θ=bubble_theta(), u=2.0 + 0.5 * rng.standard_normal((g.Nz, g.Ny, g.Nx)), v=-1.0, qᵗ=0.004 * rng.random((g.Nz, g.Ny, g.Nx)))
    for _ in range(3):
        m.time_step(2.0)
    return m


def bomex_case(arch):
    import breeze_b200 as bz
    m = bz.cases.bomex_model(arch, size=(16, 8, 12), extent=1600.0, seed=7)
    for _ in range(3):
        m.time_step(2.0)
    return m


def compressible_case(arch):
    """Moving warm bubble, compressible dynamics, WS-RK3 with 6 acoustic substeps per step (default ω = 0.65, Klemp damping 0.1)."""
    import breeze_b200 as bz
    grid = bz.RectilinearGrid(arch, size=(16, 8, 12), x=(-5e3, 5e3), y=(-2.5e3, 2.5e3), z=(0, 10e3))
    dyn = bz.CompressibleDynamics(bz.SplitExplicitTimeDiscretization(substeps=6), reference_potential_temperature=300.0)
    m = bz.AtmosphereModel(grid, dynamics=dyn)
    _, rho, _ = m.reference_profiles()
    rng = np.random.default_rng(20261018)
    shape = m.context.shape(0)
    m.set(ρ=rho[:, None, None] * (1 + 1e-4 * rng.standard_normal(shape)), u=3.0 + 0.3 * rng.standard_normal(shape), v=-1.0,
          θ=lambda x, y, z: 300.0 + 2.0 * np.cos(np.pi / 2 * np.minimum(1.0, np.sqrt(x ** 2 + y ** 2 + (z - 3000.0) ** 2) / 2500.0)) ** 2)
    for _ in range(3):
        m.time_step(3.0)
    return m


FIELDS = ["ρu", "ρv", "ρw", "ρθ", "ρq", "T", "φ"]
COMPRESSIBLE_FIELDS = ["ρ", "ρu", "ρv", "ρw", "ρθ", "T", "p", "⟨w⟩"]

if __name__ == "__main__":
    from oracle_lib import CPUOracle
    json.dump(KNOWN, open(os.path.join(HERE, "reference_known_answers.json"), "w"), indent=1, ensure_ascii=False)
    for name, case in (("oracle_bubble_16x8x12", bubble_case), ("oracle_bomex_16x8x12", bomex_case)):
        m = case(CPUOracle())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **{f: m.field(f) for f in FIELDS})
        print(name, {f: float(np.abs(m.field(f)).max()) for f in FIELDS})
    m = compressible_case(CPUOracle())
    np.savez_compressed(os.path.join(HERE, "oracle_compressible_16x8x12.npz"), **{f: m.field(f) for f in COMPRESSIBLE_FIELDS})
    print("oracle_compressible_16x8x12", {f: float(np.abs(m.field(f)).max()) for f in COMPRESSIBLE_FIELDS})
