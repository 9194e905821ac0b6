"""GPU parity tests: libbreeze_b200.so (CUDA, through the C ABI) against the CPU oracle on identical inputs.

Tolerances (FP64). The two sides evaluate the same scheme with different but algebraically identical arithmetic
(difference-form vs quadratic-form smoothness indicators, one fused reciprocal vs four divisions, FMA contraction,
FFT butterfly order), so agreement is to round-off amplified by the WENO weights' sensitivity:
  * one pressure solve / projection               : 1e-9 relative to the field's max-norm
  * one tendency evaluation vs the oracle evaluating the smoothness indicators in the SAME (difference) form
    as the kernels (oracle_lib.set_beta_form(1))  : 1e-11
  * one tendency evaluation vs the oracle in the reference's quadratic form: 1e-7. The quadratic forms
    β = ψ₁(C₁ψ₁ + C₂ψ₂ + C₃ψ₃) + … cancel |ψ|² down to |δψ|², so for θ ≈ 300 K they carry ≈ 300²·eps ≈ 1e-11 of
    noise against β + ε ≈ 1e-8…1e-5: the reference's own CPU-vs-GPU runs differ by this much
    (docs/src/reproducibility.md). The oracle's two forms differ from each other by the same 1e-9…1e-8.
  * N = 10 full SSP-RK3 steps of the bubble       : 1e-8 relative to the field's max-norm
Every MULTI-STEP comparison runs against the oracle in its default form, i.e. the reference's quadratic-form smoothness
indicators; the kernel-matching difference form (set_beta_form(1)) is used for single tendency evaluations at 1e-11 only.
States seeded with grid-scale random noise (BOMEX) put every WENO stencil in its most weight-sensitive regime, where the two
algebraically identical indicator forms differ by the quadratic form's cancellation noise (≈ |ψ|² eps against β + ε): those
comparisons state their own, looser tolerance (TOL_STEPS_NOISY), measured on a B200 (profiles/r2_parity_errors.txt).
"""
import numpy as np
import pytest

from conftest import bubble_theta, make_bubble_model, rel_err, report

pytestmark = pytest.mark.gpu

TOL_HOOK = 1e-9
TOL_SAME_FORM = 1e-11
TOL_REFERENCE_FORM = 1e-7
TOL_STEPS = 1e-8
TOL_STEPS_NOISY = 1e-8       # steps from a state with grid-scale random noise, reference-form oracle (measured 4e-10: profiles/r2a_parity_errors.txt)
PROGNOSTIC = ["ρu", "ρv", "ρw", "ρθ", "ρq"]


def _pair(oracle_arch, size, flat_y=False, seed=0, moist=False, microphysics=None, **arch_kw):
    import breeze_b200 as bz
    rng = np.random.default_rng(seed)
    models = [make_bubble_model(a, size, flat_y=flat_y, microphysics=microphysics) for a in (bz.B200(**arch_kw), oracle_arch)]
    g = models[0].grid
    shp_c, shp_w = (g.Nz, g.Ny, g.Nx), (g.Nz + 1, g.Ny, g.Nx)
    u = 3.0 * rng.standard_normal(shp_c)
    v = 2.0 * rng.standard_normal(shp_c) * (0.0 if flat_y else 1.0)
    w = 1.0 * rng.standard_normal(shp_w)
    q = 0.01 * rng.random(shp_c) if moist else None
    for m in models:
        kw = dict(θ=bubble_theta(), u=u, v=v, w=w)
        if moist:
            kw["qᵗ"] = q
        m.set(**kw)
    return models


@pytest.mark.parametrize("size,flat_y", [((32, 16, 24), False), ((64, 40), True), ((16, 8, 8), False)])
@pytest.mark.parametrize("use_tma", [1, 2])
def test_set_state_projection_matches_oracle(oracle_arch, size, flat_y, use_tma):
    gpu, cpu = _pair(oracle_arch, size, flat_y, use_tma=use_tma)
    for name in PROGNOSTIC + ["φ", "u", "v", "w", "θ", "T"]:
        assert rel_err(gpu.field(name), cpu.field(name)) < TOL_HOOK, name
    n = np.prod(size)
    assert gpu.context.max_abs_divergence() < n * 2.3e-16 * 50


@pytest.mark.parametrize("size,flat_y", [((32, 16, 24), False), ((64, 40), True), ((64, 32, 16), False)])
@pytest.mark.parametrize("use_tma", [1, 2])
@pytest.mark.parametrize("moist", [False, True])
@pytest.mark.parametrize("beta_form,tol", [(1, TOL_SAME_FORM), (0, TOL_REFERENCE_FORM)])
def test_tendencies_match_oracle(oracle_arch, size, flat_y, use_tma, moist, beta_form, tol):
    import oracle_lib
    gpu, cpu = _pair(oracle_arch, size, flat_y, use_tma=use_tma, moist=moist)
    gpu.context.compute_tendencies()
    oracle_lib.set_beta_form(beta_form)
    try:
        cpu.context.compute_tendencies()
    finally:
        oracle_lib.set_beta_form(0)
    for name in PROGNOSTIC:
        a, b = gpu.context.get_tendency(name), cpu.context.get_tendency(name)
        assert rel_err(a, b) < tol, name


@pytest.mark.parametrize("size,flat_y,z_chunks", [((32, 16, 24), False, 1), ((32, 16, 48), False, 3), ((64, 40), True, 2)])
def test_z_chunking_is_bit_identical(oracle_arch, size, flat_y, z_chunks):
    import breeze_b200 as bz
    outs = []
    for zc in (1, z_chunks):
        m = make_bubble_model(bz.B200(z_chunks=zc), size, flat_y=flat_y)
        rng = np.random.default_rng(3)
        g = m.grid
        m.set(θ=bubble_theta(), u=rng.standard_normal((g.Nz, g.Ny, g.Nx)), w=rng.standard_normal((g.Nz + 1, g.Ny, g.Nx)))
        m.context.compute_tendencies()
        outs.append([m.context.get_tendency(n) for n in PROGNOSTIC])
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("size,flat_y,dt", [((32, 32, 32), False, 2.0), ((128, 64), True, 2.0)])
@pytest.mark.parametrize("use_tma", [1, 2])
def test_ten_steps_match_oracle(oracle_arch, size, flat_y, dt, use_tma):
    import breeze_b200 as bz
    gpu = make_bubble_model(bz.B200(use_tma=use_tma), size, flat_y=flat_y)
    cpu = make_bubble_model(oracle_arch, size, flat_y=flat_y)
    for m in (gpu, cpu):
        m.set(θ=bubble_theta(), u=1.0)
    for _ in range(10):
        gpu.time_step(dt)
        cpu.time_step(dt)
    for name in PROGNOSTIC + ["u", "w", "θ", "T"]:
        assert rel_err(gpu.field(name), cpu.field(name)) < TOL_STEPS, name
    assert gpu.clock == cpu.clock


def test_tma_and_plain_staging_are_bit_identical(oracle_arch):
    import breeze_b200 as bz
    outs = []
    for mode in (1, 2):
        m = make_bubble_model(bz.B200(use_tma=mode), (32, 16, 24))
        m.set(θ=bubble_theta(), u=2.0, v=-1.0)
        for _ in range(3):
            m.time_step(2.0)
        outs.append([m.field(n) for n in PROGNOSTIC])
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_analytic_column_pressure_solve(oracle_arch):
    """test/anelastic_pressure_solver_analytic.jl:9-51 through the CUDA path."""
    import breeze_b200 as bz
    grid = bz.RectilinearGrid(bz.B200(), size=48, z=(0, 1), topology=(bz.Flat, bz.Flat, bz.Bounded))
    ref = bz.ReferenceState(grid, surface_pressure=101325, potential_temperature=288, density=grid.znodes())
    model = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(ref))
    model.set(ρw=lambda z: z ** 2 - z ** 3)
    phi = model.field("φ")[:, 0, 0]
    z = grid.znodes()
    exact = z ** 2 / 2 - z ** 3 / 3 - 1 / 12
    exact -= exact.mean()
    assert abs(phi.mean()) < 10 * 48 * 2.3e-16
    assert np.linalg.norm(phi - exact) / np.linalg.norm(exact) < 1e-3


def test_projection_is_divergence_free_rho_one(oracle_arch):
    """test/anelastic_pressure_solver_nonhydrostatic.jl:7-49: ρᵣ ≡ 1, random momentum, max|div| < N eps."""
    import breeze_b200 as bz
    N = 32
    rng = np.random.default_rng(1)
    grid = bz.RectilinearGrid(bz.B200(), size=(N, N, N), x=(0, 1), y=(0, 1), z=(0, 1))
    model = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, density=np.ones(N))))
    model.set(ρu=rng.random((N, N, N)), ρv=rng.random((N, N, N)), ρw=rng.random((N + 1, N, N)))
    assert model.context.max_abs_divergence() < N ** 3 * 2.220446049250313e-16


def _moist_pair(oracle_arch, size):
    import breeze_b200 as bz
    rng = np.random.default_rng(11)
    models = [make_bubble_model(a, size, microphysics=bz.SaturationAdjustment()) for a in (bz.B200(), oracle_arch)]
    g = models[0].grid
    shp = (g.Nz, g.Ny, g.Nx)
    z = g.znodes()[:, None, None]
    qt = 0.016 * np.exp(-z / 2500.0) * (1 + 0.5 * rng.random(shp))      # super-saturated in places → cloud
    u, w = rng.standard_normal(shp), 0.5 * rng.standard_normal((g.Nz + 1, g.Ny, g.Nx))
    for m in models:
        m.set(θ=bubble_theta(theta0=300.0, dtheta=3.0), qᵗ=qt, u=u, w=w)
    return models


def test_saturation_adjustment_diagnostics_and_tendencies(oracle_arch):
    """BASELINE config 3's thermodynamics: warm-phase SaturationAdjustment inside the stage kernel (buoyancy) and diagnostics."""
    import oracle_lib
    gpu, cpu = _moist_pair(oracle_arch, (32, 16, 24))
    ql = cpu.field("qˡ")
    assert (ql > 0).mean() > 0.02, "the test state must contain cloud"
    for name in ("T", "qᵛ", "qˡ", "θ"):
        assert rel_err(gpu.field(name), cpu.field(name)) < 1e-11, name
    gpu.context.compute_tendencies()
    oracle_lib.set_beta_form(1)
    try:
        cpu.context.compute_tendencies()
    finally:
        oracle_lib.set_beta_form(0)
    for name in PROGNOSTIC:
        assert rel_err(gpu.context.get_tendency(name), cpu.context.get_tendency(name)) < TOL_SAME_FORM, name


def test_saturation_adjustment_steps(oracle_arch):
    gpu, cpu = _moist_pair(oracle_arch, (32, 16, 24))
    for _ in range(5):
        gpu.time_step(1.0)
        cpu.time_step(1.0)
    for name in PROGNOSTIC + ["T", "qˡ"]:
        assert rel_err(gpu.field(name), cpu.field(name)) < TOL_STEPS, name


def test_bomex_forcings_match_oracle(oracle_arch):
    """BASELINE config 3 physics (Coriolis, geostrophic + subsidence forcing, drying, radiative cooling, bottom fluxes and drag,
    saturation adjustment) on a reduced BOMEX grid: one tendency evaluation and five steps against the oracle."""
    import breeze_b200 as bz
    import oracle_lib
    gpu = bz.cases.bomex_model(bz.B200(), size=(32, 16, 30), extent=3200.0)
    cpu = bz.cases.bomex_model(oracle_arch, size=(32, 16, 30), extent=3200.0)
    gpu.context.compute_tendencies()
    oracle_lib.set_beta_form(1)
    try:
        cpu.context.compute_tendencies()
    finally:
        oracle_lib.set_beta_form(0)
    for name in PROGNOSTIC:
        assert rel_err(gpu.context.get_tendency(name), cpu.context.get_tendency(name)) < TOL_SAME_FORM, name
    # Five steps against the oracle in the REFERENCE form (default). The initial state carries grid-scale random noise, where the
    # WENO weights amplify the quadratic-form indicators' cancellation noise; momentum components are compared on the common
    # momentum scale (ρv, ρw are small and fed by buoyancy).
    for _ in range(5):
        gpu.time_step(2.0)
        cpu.time_step(2.0)
    mom_scale = max(np.abs(cpu.field(n)).max() for n in ("ρu", "ρv", "ρw"))
    for name in ("ρu", "ρv", "ρw"):
        err = np.abs(gpu.field(name) - cpu.field(name)).max() / mom_scale
        report(err, name)
        assert err < TOL_STEPS_NOISY, (name, err)
    for name in ("ρθ", "ρq", "T", "qˡ"):
        assert rel_err(gpu.field(name), cpu.field(name)) < TOL_STEPS_NOISY, name


# ---- BASELINE configurations at (or near) their full sizes ------------------------------------------------------------

def test_config0_readme_quickstart_2d_256(oracle_arch):
    """BASELINE config 0: README quick-start, 2-D 256×256 (Periodic, Flat, Bounded), Δt = 2, 10 steps, against the oracle."""
    import breeze_b200 as bz
    gpu = make_bubble_model(bz.B200(), (256, 256), flat_y=True)
    cpu = make_bubble_model(oracle_arch, (256, 256), flat_y=True)
    for m in (gpu, cpu):
        m.set(θ=bubble_theta())
    for _ in range(10):
        gpu.time_step(2.0)
        cpu.time_step(2.0)
    for name in ("ρu", "ρw", "ρθ", "u", "w", "θ", "T"):
        assert rel_err(gpu.field(name), cpu.field(name)) < TOL_STEPS, name


def test_config1_bubble_3d_128_vs_oracle(oracle_arch):
    """BASELINE config 1 (3-D dry bubble, anelastic WENO5, FP64 match vs the CPU run) at 128³ — the size the 16-core oracle
    steps in seconds; 3 steps. The 256³ / 512³ runs are checked through size-independent properties below."""
    import breeze_b200 as bz
    gpu = make_bubble_model(bz.B200(), (128, 128, 128))
    cpu = make_bubble_model(oracle_arch, (128, 128, 128))
    for m in (gpu, cpu):
        m.set(θ=bubble_theta())
    for _ in range(3):
        gpu.time_step(0.5)
        cpu.time_step(0.5)
    for name in PROGNOSTIC + ["w", "θ"]:
        assert rel_err(gpu.field(name), cpu.field(name)) < TOL_STEPS, name


def test_config3_bomex_full_size_with_cloud_vs_oracle(oracle_arch):
    """BASELINE config 3 at its own size: BOMEX 128 x 128 x 75 (moist θ_li, warm-phase saturation adjustment, subsidence + geostrophic +
    Coriolis + prescribed drying / cooling, surface flux BCs), two steps from a state WITH CLOUD (seeded moist thermals: the secant branch
    of the saturation adjustment runs), against the oracle in the reference form."""
    import breeze_b200 as bz
    gpu = bz.cases.bomex_model(bz.B200(), size=(128, 128, 75), extent=12800.0, cloud=True)
    cpu = bz.cases.bomex_model(oracle_arch, size=(128, 128, 75), extent=12800.0, cloud=True)
    assert (cpu.field("qˡ") > 0).mean() > 0.01, "the state must contain cloud"
    for _ in range(2):
        gpu.time_step(1.0)
        cpu.time_step(1.0)
    assert (cpu.field("qˡ") > 0).mean() > 0.01
    mom_scale = max(np.abs(cpu.field(n)).max() for n in ("ρu", "ρv", "ρw"))
    for name in ("ρu", "ρv", "ρw"):
        err = np.abs(gpu.field(name) - cpu.field(name)).max() / mom_scale
        report(err, name)
        assert err < TOL_STEPS_NOISY, (name, err)
    for name in ("ρθ", "ρq", "T", "qᵛ", "qˡ"):
        assert rel_err(gpu.field(name), cpu.field(name)) < TOL_STEPS_NOISY, name


def test_config1_bubble_3d_256_vs_oracle(oracle_arch):
    """BASELINE config 1 at its own size (256³, 3-D dry bubble, anelastic WENO5, FP64 match vs the CPU run): two steps against the oracle
    (≈ 20 s of host work on 16 cores). bench.py repeats this comparison over four steps in every N = 1 run (checks.parity_256)."""
    import breeze_b200 as bz
    gpu = make_bubble_model(bz.B200(), (256, 256, 256))
    cpu = make_bubble_model(oracle_arch, (256, 256, 256))
    for m in (gpu, cpu):
        m.set(θ=bubble_theta())
    for _ in range(2):
        gpu.time_step(0.5)
        cpu.time_step(0.5)
    for name in PROGNOSTIC + ["w", "θ"]:
        assert rel_err(gpu.field(name), cpu.field(name)) < TOL_STEPS, name


def test_cell_advection_timescale_and_finite_check_match_oracle(oracle_arch):
    """cell_advection_timescale (src/AtmosphereModels/cell_advection_timescale.jl:46-65) and the NaN checker's reduction
    (atmosphere_model.jl:561-572) on the device against the oracle / numpy."""
    gpu, cpu = _pair(oracle_arch, (32, 16, 24))
    tg, tc = gpu.context.cell_advection_timescale(), cpu.context.cell_advection_timescale()
    report(abs(tg - tc) / tc, "tau")
    assert abs(tg - tc) < 1e-12 * tc
    for _ in range(3):
        gpu.time_step(1.0)
        cpu.time_step(1.0)
    tg, tc = gpu.context.cell_advection_timescale(), cpu.context.cell_advection_timescale()
    assert abs(tg - tc) < 1e-9 * tc
    assert gpu.context.state_is_finite()
    bad = gpu.field("ρθ")
    bad[3, 2, 1] = np.nan
    gpu.context.set_state(rho_theta=bad, enforce_mass_conservation=False)
    assert not gpu.context.state_is_finite()


@pytest.mark.parametrize("size,flat_y", [((24, 48, 16), False), ((96, 24, 12), False), ((192, 20), True), ((48, 64, 10), False),
                                         ((40, 56, 12), False), ((80, 112, 10), False), ((160, 24, 8), False), ((224, 20), True),
                                         ((320, 16), True), ((448, 40, 6), False), ((640, 8, 6), False), ((16, 896, 6), False),
                                         ((1280, 12), True), ((8, 1792, 4), False)])
def test_mixed_radix_grids_match_oracle(oracle_arch, size, flat_y):
    """Horizontal sizes 3 · 2^m (the reference benchmarks 768 x 768 x 256, .github/workflows/Benchmarks.yml:41), 5 · 2^m and 7 · 2^m
    (896^3: benchmarking/README.md:225-233): the in-house FFT's radix-3 / 5 / 7 passes, the projection and three full steps against the
    oracle (whose DFT is an independent mixed-radix implementation)."""
    gpu, cpu = _pair(oracle_arch, size, flat_y, seed=5, moist=True)
    for name in PROGNOSTIC + ["φ"]:
        assert rel_err(gpu.field(name), cpu.field(name)) < TOL_HOOK, name
    assert gpu.context.max_abs_divergence() < np.prod(size) * 2.3e-16 * 50
    for _ in range(3):
        gpu.time_step(1.0)
        cpu.time_step(1.0)
    for name in PROGNOSTIC:
        assert rel_err(gpu.field(name), cpu.field(name)) < TOL_STEPS_NOISY, name


def test_reference_benchmark_grid_768x768x256_properties():
    """`768x768x256` of the reference's benchmark matrix (Benchmarks.yml:41) on the CUDA path: divergence-free momentum after every step,
    conservation of ∫ρθ, and the x↔y symmetry of the centred bubble (ρv(k, j, i) = ρu(k, i, j)) through the radix-3 transforms."""
    import breeze_b200 as bz
    grid = bz.RectilinearGrid(bz.B200(), size=(768, 768, 256), x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
    m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)), advection=bz.WENO(order=5))
    m.set(θ=bubble_theta())
    s0 = m.field("ρθ").sum()
    for _ in range(2):
        m.time_step(0.5)
    assert m.context.max_abs_divergence() < 1e-11
    ru, rv, rt = m.field("ρu"), m.field("ρv"), m.field("ρθ")
    assert abs(rt.sum() - s0) < 1e-12 * abs(s0)
    assert np.abs(rv - np.swapaxes(ru, 1, 2)).max() < 1e-10 * max(np.abs(ru).max(), 1e-30)
    assert np.abs(m.field("w")).max() > 1e-4


@pytest.mark.parametrize("case", ["bubble3d", "bubble2d", "bomex", "weno9"])
def test_graph_replay_is_bit_identical_to_eager_steps(oracle_arch, case):
    """bz_time_step replays a captured CUDA graph from the third step with the same (buffer rotation, Δt) on; with per-kernel profiling
    enabled it launches every kernel eagerly. Eight steps both ways (incl. a change of Δt and back) must agree bit for bit."""
    import breeze_b200 as bz

    def make():
        if case == "bomex":
            return bz.cases.bomex_model(bz.B200(), size=(32, 16, 30), extent=3200.0, cloud=True)
        if case == "bubble2d":
            m = make_bubble_model(bz.B200(), (64, 40), flat_y=True)
        elif case == "weno9":
            grid = bz.RectilinearGrid(bz.B200(), size=(32, 16, 24), x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
            m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)), advection=bz.WENO(order=9))
        else:
            m = make_bubble_model(bz.B200(), (32, 16, 24))
        m.set(θ=bubble_theta(), u=2.0)
        return m

    a, b = make(), make()
    b.context.profile_enable(True)                       # eager launches
    launches0 = a.context.kernel_launch_count()
    for dt in (1.0, 1.0, 1.0, 1.0, 1.0, 0.5, 1.0, 1.0):
        a.time_step(dt)
        b.time_step(dt)
    b.context.profile_read()
    for name in PROGNOSTIC + ["φ"]:
        assert np.array_equal(a.field(name), b.field(name)), name
    assert a.clock == b.clock
    assert a.context.kernel_launch_count() - launches0 > 8 * 20      # replayed launches are still counted


def test_slices_match_full_fields(oracle_arch):
    gpu, _ = _pair(oracle_arch, (32, 16, 24), moist=True)
    for name in ("θ", "ρw", "T", "φ"):
        full = gpu.field(name)
        assert np.array_equal(gpu.slice(name, x=5), full[:, :, 5]), name
        assert np.array_equal(gpu.slice(name, y=7), full[:, 7, :]), name
        assert np.array_equal(gpu.slice(name, z=3), full[3]), name


@pytest.mark.parametrize("N", [256, 512])
def test_full_size_properties(N):
    """256³ (config 1) and 512³ (the metric's workload): properties that do not need the oracle —
    discrete divergence-free momentum after every step, conservation of ∫ρu, ∫ρv, ∫ρθ (flux form, closed domain in z), the
    symmetry of the centred bubble (v(x,y,z) = u(y,x,z) under the x↔y swap of a symmetric initial state), and bit-identical
    results from the TMA-staged and plain-load kernels."""
    import breeze_b200 as bz
    results = []
    for use_tma in (1, 2):
        m = make_bubble_model(bz.B200(use_tma=use_tma), (N, N, N))
        m.set(θ=bubble_theta(), u=0.0)
        s0 = m.field("ρθ").sum()
        for _ in range(2):
            m.time_step(0.5)
        div = m.context.max_abs_divergence()
        assert div < 1e-11, div
        ru, rv, rt = m.field("ρu"), m.field("ρv"), m.field("ρθ")
        assert abs(ru.sum()) < 1e-6 * np.abs(ru).sum() + 1e-9 and abs(rv.sum()) < 1e-6 * np.abs(rv).sum() + 1e-9
        assert abs(rt.sum() - s0) < 1e-12 * abs(s0)
        # x↔y symmetry of the bubble: ρv(k, j, i) == ρu(k, i, j)
        sym = np.abs(rv - np.swapaxes(ru, 1, 2)).max()
        assert sym < 1e-10 * max(np.abs(ru).max(), 1e-30), sym
        results.append((ru, rt))
        del m
    assert np.array_equal(results[0][0], results[1][0]) and np.array_equal(results[0][1], results[1][1])
