"""The Float32 build of the anelastic path (csrc/libbreeze_b200_f32.so, prefix bzf_; `B200(float_type="Float32")`) — the precision the
reference benchmarks in by default (benchmarking/README.md:74). The library is compiled from a mechanically retyped copy of the FP64
sources (breeze.jl_b200/make_f32.py), so these tests check (i) that it is the same algorithm — agreement with the FP64 CPU oracle to
Float32 round-off amplified by the WENO weights, with the tolerance stated per test and the measured value logged (BZ_PARITY_REPORT) —
and (ii) the invariants that do not depend on precision (TMA == plain staging bit for bit, discrete mass conservation to Float32
round-off, P-independent properties). Declared tolerances (relative to each field's max-norm):
  * set! + projection (pressure solve, velocities, θ, T)      : 2e-5   (the solve amplifies 6e-8 by the condition of the Poisson operator)
  * one tendency evaluation                                    : 5e-3   (θ ≈ 300 K: the WENO second differences of θ carry 300 · 6e-8 of noise)
  * five SSP-RK3 steps of the bubble, thermodynamic fields     : 1e-5
  * five steps, momentum (relative to the largest component)   : 2e-5
  * BOMEX with cloud / StaticEnergy bubble, five steps         : thermodynamic fields 1e-5, momentum 2e-5 of the largest component,
    cloud liquid 5e-3 of its maximum (measured 4e-7, 5e-7 / 4e-6, 2.4e-4: profiles/r2z_parity_errors_f32_forced_moist.txt)
"""
import numpy as np
import pytest

from conftest import bubble_theta, make_bubble_model, rel_err, report

pytestmark = pytest.mark.gpu

PROGNOSTIC = ["ρu", "ρv", "ρw", "ρθ", "ρq"]
TOL_HOOK, TOL_TENDENCY, TOL_THERMO, TOL_MOMENTUM = 2e-5, 5e-3, 1e-5, 2e-5      # measured 2e-6, 1.2e-3, 2e-6, 2e-6 (profiles/r2p_parity_errors_f32.txt)


def _pair(oracle_arch, size, flat_y=False, seed=0, order=5, **kw):
    import breeze_b200 as bz
    rng = np.random.default_rng(seed)
    models = []
    for arch in (bz.B200(float_type="Float32", **kw), oracle_arch):
        if flat_y:
            grid = bz.RectilinearGrid(arch, size=size, x=(-10e3, 10e3), z=(0, 10e3), topology=(bz.Periodic, bz.Flat, bz.Bounded))
        else:
            grid = bz.RectilinearGrid(arch, size=size, x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
        models.append(bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)),
                                         advection=bz.WENO(order=order)))
    g = models[0].grid
    shp_c, shp_w = (g.Nz, g.Ny, g.Nx), (g.Nz + 1, g.Ny, g.Nx)
    # smooth velocities (grid-scale noise would only measure the Float32 noise of the WENO weights)
    z, y, x = np.meshgrid(g.znodes(), g.ynodes(), g.xnodes(), indexing="ij")
    u = 3.0 * np.sin(2 * np.pi * x / 20e3) * np.cos(2 * np.pi * z / 10e3) + 1.0
    v = 2.0 * np.cos(2 * np.pi * x / 20e3) * (0.0 if flat_y else 1.0) + 0 * y
    zf = np.broadcast_to(g.znodes(face=True)[:, None, None], shp_w)
    w = 0.5 * np.sin(np.pi * zf / 10e3) * np.sin(2 * np.pi * np.broadcast_to(g.xnodes()[None, None, :], shp_w) / 20e3)
    q = 0.01 * np.exp(-z / 3000.0)
    for m in models:
        m.set(θ=bubble_theta(), u=u, v=v, w=w, qᵗ=q)
    return models


def test_float32_library_is_float32(oracle_arch):
    import breeze_b200 as bz
    gpu, _ = _pair(oracle_arch, (16, 8, 8))
    assert gpu.field("ρθ").dtype == np.float32 and gpu.context.lib.prefix == "bzf_"
    assert gpu.context.lib.path.endswith("libbreeze_b200_f32.so")
    # the compressible path of the same library (bzcf_*, tests/test_gpu_compressible_float32.py)
    grid = bz.RectilinearGrid(bz.B200(float_type="Float32"), size=(16, 8, 8), x=(0, 1e3), y=(0, 1e3), z=(0, 1e3))
    cm = bz.AtmosphereModel(grid, dynamics=bz.CompressibleDynamics(bz.SplitExplicitTimeDiscretization(substeps=6)))
    assert cm.context.lib.prefix == "bzcf_" and cm.field("ρθ").dtype == np.float32


@pytest.mark.parametrize("size,flat_y", [((32, 16, 24), False), ((64, 40), True), ((48, 24, 16), False), ((40, 56, 12), False)])
@pytest.mark.parametrize("use_tma", [1, 2])
def test_set_state_projection_matches_oracle(oracle_arch, size, flat_y, use_tma):
    gpu, cpu = _pair(oracle_arch, size, flat_y, use_tma=use_tma)
    for name in PROGNOSTIC + ["φ", "u", "w", "θ", "T"]:
        assert rel_err(gpu.field(name).astype(np.float64), cpu.field(name)) < TOL_HOOK, name
    scale = max(np.abs(cpu.field(n)).max() for n in ("ρu", "ρv", "ρw")) / min(gpu.grid.Δx, gpu.grid.Δz)
    assert gpu.context.max_abs_divergence() < 1e-5 * scale                    # Float32 round-off of the projected momentum


@pytest.mark.parametrize("size,flat_y,order", [((32, 16, 24), False, 5), ((64, 40), True, 5), ((32, 16, 24), False, 9)])
def test_tendencies_match_oracle(oracle_arch, size, flat_y, order):
    gpu, cpu = _pair(oracle_arch, size, flat_y, order=order)
    gpu.context.compute_tendencies()
    cpu.context.compute_tendencies()
    for name in PROGNOSTIC:
        assert rel_err(gpu.context.get_tendency(name).astype(np.float64), cpu.context.get_tendency(name)) < TOL_TENDENCY, name


@pytest.mark.parametrize("size,flat_y,order", [((32, 32, 32), False, 5), ((128, 64), True, 5), ((32, 32, 32), False, 9)])
def test_five_steps_match_oracle(oracle_arch, size, flat_y, order):
    gpu, cpu = _pair(oracle_arch, size, flat_y, order=order)
    for _ in range(5):
        gpu.time_step(2.0)
        cpu.time_step(2.0)
    for name in ("ρθ", "ρq", "θ", "T"):
        assert rel_err(gpu.field(name).astype(np.float64), cpu.field(name)) < TOL_THERMO, name
    mom = max(np.abs(cpu.field(n)).max() for n in ("ρu", "ρv", "ρw"))
    for name in ("ρu", "ρv", "ρw"):
        err = np.abs(gpu.field(name) - cpu.field(name)).max() / mom
        report(err, name)
        assert err < TOL_MOMENTUM, name
    assert gpu.clock == cpu.clock and gpu.context.state_is_finite()


def test_tma_and_plain_staging_are_bit_identical():
    import breeze_b200 as bz
    outs = []
    for mode in (1, 2):
        m = make_bubble_model(bz.B200(use_tma=mode, float_type="Float32"), (32, 16, 24))
        m.set(θ=bubble_theta(), u=2.0, v=-1.0)
        for _ in range(3):
            m.time_step(2.0)
        outs.append([m.field(n) for n in PROGNOSTIC])
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_float32_512_cubed_properties():
    """The metric's workload in Float32: finite, divergence at Float32 round-off, ∫ρθ conserved to Float32 accumulation error, bubble rises."""
    import breeze_b200 as bz
    m = make_bubble_model(bz.B200(float_type="Float32"), (512, 512, 512))
    m.set(θ=bubble_theta())
    s0 = m.field("ρθ").astype(np.float64).sum()
    for _ in range(3):
        m.time_step(0.5)
    assert m.context.state_is_finite()
    rt = m.field("ρθ").astype(np.float64)
    assert abs(rt.sum() - s0) < 1e-6 * abs(s0)
    assert np.abs(m.field("w")).max() > 1e-3
    assert m.context.max_abs_divergence() < 1e-6


def test_bomex_with_cloud_matches_oracle(oracle_arch):
    """Forced + moist instantiation of the Float32 stage kernel (FPlane, subsidence, geostrophic, drying / cooling, flux BCs, saturation
    adjustment with its secant branch active): five steps of the BOMEX case against the FP64 oracle."""
    import breeze_b200 as bz
    gpu = bz.cases.bomex_model(bz.B200(float_type="Float32"), size=(32, 16, 30), extent=3200.0, cloud=True)
    cpu = bz.cases.bomex_model(oracle_arch, size=(32, 16, 30), extent=3200.0, cloud=True)
    for _ in range(5):
        gpu.time_step(1.0)
        cpu.time_step(1.0)
    for name in ("ρθ", "ρq", "θ", "T"):
        assert rel_err(gpu.field(name).astype(np.float64), cpu.field(name)) < TOL_THERMO, name
    mom = max(np.abs(cpu.field(n)).max() for n in ("ρu", "ρv", "ρw"))
    for name in ("ρu", "ρv", "ρw"):
        err = np.abs(gpu.field(name) - cpu.field(name)).max() / mom
        report(err, name)
        assert err < TOL_MOMENTUM, name
    ql_c = cpu.field("qˡ")
    assert ql_c.max() > 1e-4                                              # the secant branch ran
    assert rel_err(gpu.field("qˡ").astype(np.float64), ql_c) < 5e-3
    assert gpu.context.state_is_finite()


def test_static_energy_bubble_matches_oracle(oracle_arch):
    """StaticEnergy instantiation (ρe prognostic, buoyancy-flux term) of the Float32 stage kernel, WENO5 and WENO9."""
    import breeze_b200 as bz
    for order in (5, 9):
        models = []
        for arch in (bz.B200(float_type="Float32"), oracle_arch):
            grid = bz.RectilinearGrid(arch, size=(32, 16, 24), x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
            m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)),
                                   advection=bz.WENO(order=order), formulation="StaticEnergy")
            m.set(θ=bubble_theta(), u=2.0, v=-1.0)
            models.append(m)
        gpu, cpu = models
        for _ in range(5):
            gpu.time_step(2.0)
            cpu.time_step(2.0)
        for name in ("ρθ", "T"):
            assert rel_err(gpu.field(name).astype(np.float64), cpu.field(name)) < TOL_THERMO, (order, name)
        mom = max(np.abs(cpu.field(n)).max() for n in ("ρu", "ρv", "ρw"))
        for name in ("ρu", "ρv", "ρw"):
            err = np.abs(gpu.field(name) - cpu.field(name)).max() / mom
            report(err, f"order {order} {name}")
            assert err < TOL_MOMENTUM, (order, name)
