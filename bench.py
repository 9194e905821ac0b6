#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path: Mcell-updates/s of one anelastic SSP-RK3 time step (WENO5) on a dry
thermal bubble, the reference's `grid_points_per_second` (benchmarking/src/utils.jl:138-141; protocol :119-136).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size 512] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0). A "step" is one full time step (3 RK stages, 3 pressure solves) of the whole grid.
  value    device-timed (CUDA events on the context's stream, max over ranks), state resident in HBM
  e2e      same metric through the C ABI with HOST buffers: per step bz_set_state (pinned host → device) +
           bz_time_step + bz_get_state (device → pinned host), all inside the timed region
  roofline the fused stage kernel: algorithmic bytes (88 B/cell stage 1, 128 B/cell stages 2-3) ÷ its
           average duration, measured live with CUDA events around every launch, ÷ measured HBM peak
  cpu_baseline   the CPU oracle (restatement of the reference algorithm, C + OpenMP) timed on this box's host cores
--impl reference times that same CPU restatement as the reference arm (Julia is not installed anywhere in this project, so
Breeze's own CPU() run cannot execute).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the ONE JSON line; NCCL's version / debug lines go to stderr

STAGE_BYTES_PER_CELL = (88.0 + 128.0 + 128.0) / 3.0      # average over the three stage-kernel launches of a step
METRIC = "Mcell-updates/s"


def bubble(x, y, z):
    r = np.sqrt(x ** 2 + y ** 2 + (z - 2000.0) ** 2)
    return 300.0 + 2.0 * np.cos(np.pi / 2 * np.minimum(1.0, r / 2000.0)) ** 2


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [a.strip() for a in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """Host cores this process may run on. torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently make the
    CPU legs single-threaded (round-1 SCALE records): the oracle's thread count is therefore always set explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def oracle_library(threads=None):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_lib
    lib = oracle_lib.load_oracle_library()
    lib.dll.orc_set_num_threads(int(threads) if threads else host_threads())
    return oracle_lib, lib


def bubble_model(arch, size, order=5, formulation="LiquidIcePotentialTemperature"):
    import breeze_b200 as bz
    grid = bz.RectilinearGrid(arch, size=(size, size, size), x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
    m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)), advection=bz.WENO(order=order),
                           formulation=formulation)
    m.set(θ=bubble)
    return m


def run_oracle(size, steps, warmup, dt, threads=None, keep_model=False, budget_s=None):
    """Time the CPU oracle on `size`^3 cells of the same bubble on all host threads (or `threads`).
    Returns (Mcell/s, cores, seconds/step[, model]). budget_s: stop after that many seconds of timed steps (at least one)."""
    oracle_lib, lib = oracle_library(threads)
    try:
        cores = lib.dll.orc_num_threads()
        m = bubble_model(oracle_lib.CPUOracle(), size)
        for _ in range(warmup):
            m.time_step(dt)
        t0 = time.perf_counter()
        done = 0
        for _ in range(steps):
            m.time_step(dt)
            done += 1
            if budget_s is not None and time.perf_counter() - t0 > budget_s:
                break
        el = time.perf_counter() - t0
        out = (size ** 3 * done / el / 1e6, cores, el / done)
        return out + ((m, done) if keep_model else ())
    finally:
        lib.dll.orc_set_num_threads(host_threads())


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config 4: compressible split-explicit WS-RK3 (acoustic substepping), 256 x 256 x 64 supercell-shaped grid, dry
# ---------------------------------------------------------------------------------------------------------------------
COLUMN_BYTES_PER_CELL = 8.0 * 31          # acoustic_column: 20 reads + 11 writes per cell
HORIZONTAL_BYTES_PER_CELL = 8.0 * 11      # acoustic_horizontal: 9 reads + 2 writes


def supercell_model(arch, size, substeps):
    import breeze_b200 as bz
    grid = bz.RectilinearGrid(arch, size=size, x=(0, 168e3), y=(0, 168e3), z=(0, 20e3))
    dyn = bz.CompressibleDynamics(bz.SplitExplicitTimeDiscretization(substeps=substeps), reference_potential_temperature=300.0)
    m = bz.AtmosphereModel(grid, dynamics=dyn)
    _, rho, _ = m.reference_profiles()
    m.set(ρ=np.broadcast_to(rho[:, None, None], m.context.shape(0)).copy(),
          θ=lambda x, y, z: 300.0 + 3.0 * np.exp(-((x - 84e3) ** 2 + (y - 84e3) ** 2) / 10e3 ** 2 - (z - 1500.0) ** 2 / 1500.0 ** 2),
          u=10.0, v=5.0)
    return m


def run_oracle_compressible(size, steps, warmup, dt, substeps):
    oracle_lib, lib = oracle_library()
    cores = lib.dll.orc_num_threads()
    m = supercell_model(oracle_lib.CPUOracle(), size, substeps)
    for _ in range(warmup):
        m.time_step(dt)
    t0 = time.perf_counter()
    for _ in range(steps):
        m.time_step(dt)
    el = time.perf_counter() - t0
    return int(np.prod(size)) * steps / el / 1e6, cores, el / steps


def bench_compressible(args, steps, warmup, with_cpu=True, with_e2e=True):
    """One JSON-able dict for the compressible workload on one GPU (the path does not shard: replicas only)."""
    import torch
    size, dt, nsub = (256, 256, 64), 6.0, 6
    m = supercell_model(__import__("breeze_b200").B200(device=int(os.environ.get("LOCAL_RANK", "0"))), size, nsub)
    ctx = m.context
    cells = int(np.prod(size))
    ext_stream = torch.cuda.ExternalStream(ctx.stream())
    for _ in range(max(warmup, 3)):
        ctx.time_step(dt)
    ctx.synchronize()
    # timed region: K steps between two CUDA events on the launching stream, no per-kernel events inside
    n0 = ctx.kernel_launch_count()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(ext_stream)
    for _ in range(steps):
        ctx.time_step(dt)
    e1.record(ext_stream)
    ctx.synchronize(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    clocks = sampler.stop()
    launches = ctx.kernel_launch_count() - n0
    # second pass with per-kernel-family events (86 event records per step would otherwise sit inside a 6 ms step)
    ctx.profile_enable(True); ctx.profile_read()
    for _ in range(steps):
        ctx.time_step(dt)
    fam_ms, fam_n = ctx.profile_read()
    ctx.profile_enable(False)
    peak, peak_src = measured_peak()
    col_ms = fam_ms[3] / max(1, fam_n[3])
    hor_ms = fam_ms[2] / max(1, fam_n[2])
    achieved = COLUMN_BYTES_PER_CELL * cells / (col_ms * 1e-3) / 1e9
    names = ["slow_tendencies_weno5", "stage_setup", "acoustic_horizontal", "acoustic_column", "stage_end_update_state"]
    out = {
        "metric": METRIC, "value": cells / (ms * 1e-3) / 1e6, "unit": "Mcell-updates/s", "n_gpus": 1, "steps": steps, "warmup": max(warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "compressible split-explicit WS-RK3, supercell-shaped 256x256x64 (168 km x 168 km x 20 km), dry, WENO5, dt=6 s, "
                               "6 acoustic substeps per step (2+3+6 over the stages), forward_weight 0.65, thermal divergence damping 0.1",
                   "grid": list(size), "parallelism": "one GPU (replicas only)", "l2": "working set 1.3 GB, larger than L2",
                   "device_bytes": ctx.device_bytes()},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "kernel": "c_acoustic_column (predictors + tridiagonal solve + recovery, one thread per column)", "kernel_ms": col_ms,
                     "peak_source": peak_src, "bytes_per_cell": COLUMN_BYTES_PER_CELL,
                     "second_kernel": {"kernel": "c_acoustic_horizontal", "kernel_ms": hor_ms,
                                       "achieved": HORIZONTAL_BYTES_PER_CELL * cells / (hor_ms * 1e-3) / 1e9}},
        "breakdown_ms_per_step": {names[f]: round(fam_ms[f] / steps, 4) for f in range(5)},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if with_e2e:
        shapes = [ctx.shape(f) for f in range(5)]
        pin_in = [torch.empty(sh, dtype=torch.float64).pin_memory() for sh in shapes]
        pin_out = [torch.empty(sh, dtype=torch.float64).pin_memory() for sh in shapes]
        ctx.get_state([t.numpy() for t in pin_in])
        nbytes = sum(int(np.prod(sh)) * 8 for sh in shapes)
        t0 = time.perf_counter()
        k = 3
        for _ in range(k):
            ctx.set_state(*[t.numpy() for t in pin_in])
            ctx.time_step(dt)
            ctx.get_state([t.numpy() for t in pin_out])
            pin_in, pin_out = pin_out, pin_in
        ctx.synchronize()
        out["e2e"] = {"value": cells / ((time.perf_counter() - t0) / k) / 1e6, "unit": "Mcell-updates/s", "h2d_bytes_per_step": nbytes,
                      "d2h_bytes_per_step": nbytes, "steps": k}
    if with_cpu:
        cs = (64, 64, 64)
        v, cores, sps = run_oracle_compressible(cs, 3, 1, dt, nsub)              # same kernels and substep count per cell-step
        out["cpu_baseline"] = {"value": v, "unit": "Mcell-updates/s", "cores": cores, "kind": "port",
                               "sample": f"64x64x64 cells of the same case (same extents, dt, substeps), 3 steps after 1 warm-up, {sps:.2f} s/step "
                                         "(CPU restatement of the reference algorithm)"}
    return out


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config 3: BOMEX shallow-cumulus LES 128 x 128 x 75 (moist, warm-phase saturation adjustment, forcings, flux BCs)
# ---------------------------------------------------------------------------------------------------------------------
def bench_bomex(args, steps, warmup, with_cpu=True):
    import torch
    import breeze_b200 as bz
    size, extent, dt = (128, 128, 75), 12800.0, 1.0
    m = bz.cases.bomex_model(bz.B200(device=int(os.environ.get("LOCAL_RANK", "0"))), size=size, extent=extent, cloud=True)
    ctx = m.context
    cells = int(np.prod(size))
    ext_stream = torch.cuda.ExternalStream(ctx.stream())
    for _ in range(max(warmup, 3)):
        ctx.time_step(dt)
    ctx.synchronize()
    n0 = ctx.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(ext_stream)
    for _ in range(steps):
        ctx.time_step(dt)
    e1.record(ext_stream)
    ctx.synchronize(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = ctx.kernel_launch_count() - n0
    ctx.profile_enable(True); ctx.profile_read()          # second pass: per-kernel-family events kept out of the timed region
    for _ in range(steps):
        ctx.time_step(dt)
    fam_ms, fam_n = ctx.profile_read()
    ctx.profile_enable(False)
    peak, peak_src = measured_peak()
    stage_ms = fam_ms[0] / max(1, fam_n[0])
    achieved = STAGE_BYTES_PER_CELL * cells / (stage_ms * 1e-3) / 1e9
    out = {
        "metric": METRIC, "value": cells / (ms * 1e-3) / 1e6, "unit": "Mcell-updates/s", "n_gpus": 1, "steps": steps, "warmup": max(warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BOMEX shallow-cumulus LES 128x128x75 (12.8 km x 12.8 km x 3 km), anelastic, WENO5, SSP-RK3, warm-phase saturation "
                               "adjustment, subsidence + geostrophic + Coriolis + prescribed drying/cooling, surface flux BCs, dt=1 s",
                   "grid": list(size), "parallelism": "one GPU", "l2": "working set 0.2 GB, larger than L2", "device_bytes": ctx.device_bytes()},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "kernel": "stage_kernel<FORCED, saturation adjustment>", "kernel_ms": stage_ms, "peak_source": peak_src,
                     "bytes_per_cell": STAGE_BYTES_PER_CELL},
        "breakdown_ms_per_step": {n: round(fam_ms[f] / steps, 4) for f, n in enumerate(["stage_tendency_rk", "poisson_forward", "thomas", "poisson_inverse", "projection_halo_means"])},
        "gpu_launches": int(launches),
        "checks": {"max_abs_divergence": ctx.max_abs_divergence(), "max_cloud_liquid": float(m.field("qˡ").max()),
                   "cloudy_cell_fraction": float((m.field("qˡ") > 0).mean()),
                   "note": "moist thermals seeded in the cumulus layer (cases.bomex_model(cloud=True)): the secant branch of the saturation adjustment runs in the timed region"},
    }
    if not with_cpu:
        return out
    oracle_lib, _ = oracle_library()
    cm = bz.cases.bomex_model(oracle_lib.CPUOracle(), size=(64, 64, 75), extent=6400.0, cloud=True)
    cm.time_step(dt)
    t0 = time.perf_counter()
    for _ in range(2):
        cm.time_step(dt)
    sps = (time.perf_counter() - t0) / 2
    out["cpu_baseline"] = {"value": 64 * 64 * 75 / sps / 1e6, "unit": "Mcell-updates/s", "cores": oracle_lib.load_oracle_library().dll.orc_num_threads(),
                           "kind": "port", "sample": f"64x64x75 cells of the same case (the example's own size), 2 steps after 1 warm-up, {sps:.2f} s/step"}
    return out


# ---------------------------------------------------------------------------------------------------------------------
# The shipped examples' scheme: WENO(order = 9) + StaticEnergyFormulation (examples/dry_thermal_bubble.jl:24), stage_hi_kernel
# ---------------------------------------------------------------------------------------------------------------------
HI_BYTES_PER_CELL = STAGE_BYTES_PER_CELL + 80.0 + (80.0 + 40.0 + 40.0) / 3.0   # + specific-field pass (5 reads, 5 writes) + its 5 re-reads as stencil operands


def bench_shipped_bubble(args):
    import torch
    import breeze_b200 as bz
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    out = {}
    # (i) examples/dry_thermal_bubble.jl as shipped: 2-D 128 x 128 (Periodic, Flat, Bounded), WENO9, :StaticEnergy, Δθ = 10 K
    grid = bz.RectilinearGrid(bz.B200(device=dev), size=(128, 128), x=(-10e3, 10e3), z=(0, 10e3), topology=(bz.Periodic, bz.Flat, bz.Bounded))
    m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)), advection=bz.WENO(order=9),
                           formulation="StaticEnergy")
    m.set(θ=lambda x, z: 300.0 + 10.0 * np.maximum(0.0, 1.0 - np.sqrt(x ** 2 + (z - 3000.0) ** 2) / 2000.0))
    ctx = m.context
    for _ in range(5):
        ctx.time_step(2.0)
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        ctx.time_step(2.0)
    ctx.synchronize()
    ms2d = (time.perf_counter() - t0) / 200 * 1e3
    out["as_shipped_2d_128x128"] = {"ms_per_step": ms2d, "value": 128 * 128 / (ms2d * 1e-3) / 1e6, "unit": "Mcell-updates/s", "steps": 200,
                                    "note": "launch-bound: 16 384 cells; wall clock around 200 steps incl. launch overheads", "max_abs_w": float(np.abs(m.field("w")).max())}
    del m, ctx
    # (ii) the same scheme and formulation on the 3-D 256^3 bubble (BASELINE config 1's grid)
    N = 256
    m = bubble_model(bz.B200(device=dev), N, order=9, formulation="StaticEnergy")
    ctx = m.context
    ext_stream = torch.cuda.ExternalStream(ctx.stream())
    for _ in range(3):
        ctx.time_step(args.dt)
    ctx.profile_enable(True)
    for _ in range(5):
        ctx.time_step(args.dt)
    ctx.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.synchronize(); torch.cuda.synchronize()
    e0.record(ext_stream)
    for _ in range(5):
        ctx.time_step(args.dt)
    e1.record(ext_stream)
    ctx.synchronize(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fam_ms, fam_n = ctx.profile_read()
    ctx.profile_enable(False)
    peak, peak_src = measured_peak()
    stage_ms = fam_ms[0] / max(1, fam_n[0])                  # specific-field pass + stage_hi_kernel per stage
    achieved = HI_BYTES_PER_CELL * N ** 3 / (stage_ms * 1e-3) / 1e9
    out["weno9_static_energy_3d_256"] = {
        "value": N ** 3 / (ms * 1e-3) / 1e6, "unit": "Mcell-updates/s", "ms_per_step": ms, "steps": 5,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "kernel": "specific_fields_kernel + stage_hi_kernel<5> (fused WENO9 tendencies + RK update)", "kernel_ms": stage_ms,
                     "bytes_per_cell": HI_BYTES_PER_CELL, "peak_source": peak_src,
                     "note": "≈ 2700 FP64 instructions per cell: FP64-pipe bound (floor ≈ 2.4 ms per launch at 256^3), not HBM bound"},
        "breakdown_ms_per_step": {n: round(fam_ms[f] / 5, 4) for f, n in enumerate(["stage_tendency_rk", "poisson_forward", "thomas", "poisson_inverse", "projection_halo"])},
        "checks": {"max_abs_divergence": ctx.max_abs_divergence()}}
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Float32 build of the same path (libbreeze_b200_f32.so): the precision the reference benchmarks in by default (benchmarking/README.md:74)
# ---------------------------------------------------------------------------------------------------------------------
def bench_float32(args, N):
    import torch
    import breeze_b200 as bz
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    grid = bz.RectilinearGrid(bz.B200(device=dev, float_type="Float32"), size=(N, N, N), x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
    m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)), advection=bz.WENO(order=5))
    m.set(θ=bubble)
    ctx = m.context
    ext_stream = torch.cuda.ExternalStream(ctx.stream())
    for _ in range(3):
        ctx.time_step(args.dt)
    ctx.profile_enable(True)
    for _ in range(10):
        ctx.time_step(args.dt)
    ctx.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.synchronize(); torch.cuda.synchronize()
    e0.record(ext_stream)
    for _ in range(10):
        ctx.time_step(args.dt)
    e1.record(ext_stream)
    ctx.synchronize(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fam_ms, fam_n = ctx.profile_read()
    ctx.profile_enable(False)
    peak, peak_src = measured_peak()
    stage_ms = float(fam_ms[0]) / max(1, int(fam_n[0]))
    bpc = STAGE_BYTES_PER_CELL / 2
    achieved = bpc * N ** 3 / (stage_ms * 1e-3) / 1e9
    # the same 64^3 bubble stepped by the Float64 CPU oracle: the declared Float32 tolerance of tests/test_gpu_float32.py, measured here
    small = bubble_model(bz.B200(device=dev, float_type="Float32"), 64)
    oracle_lib, _ = oracle_library()
    ref = bubble_model(oracle_lib.CPUOracle(), 64)
    for _ in range(5):
        small.time_step(2.0); ref.time_step(2.0)
    errs = {}
    for n in ("ρθ", "ρu", "ρw"):
        b = ref.field(n); a = small.field(n).astype(np.float64)
        sc = max(float(np.abs(ref.field(k)).max()) for k in (("ρu", "ρv", "ρw") if n != "ρθ" else ("ρθ",)))
        errs[n] = float(np.abs(a - b).max() / sc)
    return {"value": N ** 3 / (ms * 1e-3) / 1e6, "unit": "Mcell-updates/s", "ms_per_step": ms, "steps": 10, "dtype": "f32",
            "workload": f"3-D dry thermal bubble {N}^3, anelastic, WENO5, SSP-RK3, Float32, dt={args.dt}",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "stage_kernel (Float32 build: two CTAs per SM)", "kernel_ms": stage_ms, "bytes_per_cell": bpc, "peak_source": peak_src},
            "breakdown_ms_per_step": {n: round(float(fam_ms[f]) / 10, 4) for f, n in enumerate(["stage_tendency_rk", "poisson_forward", "thomas", "poisson_inverse", "projection_halo"])},
            "checks": {"finite": ctx.state_is_finite(), "max_abs_divergence": ctx.max_abs_divergence(),
                       "vs_float64_oracle_64cubed_5_steps": {"max_rel": errs, "tol": 2e-5, "ok": bool(max(errs.values()) < 2e-5)}}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=512, help="global grid is size^3 (strong scaling: fixed as N grows)")
    ap.add_argument("--dt", type=float, default=0.5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--cpu-size", type=int, default=256, help="the CPU restatement runs cpu_size^3 cells of the same bubble (≈ 10-20 s of host work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--use-tma", type=int, default=0)
    ap.add_argument("--z-chunks", type=int, default=0)
    ap.add_argument("--no-peer-memory", action="store_true", help="multi-GPU: NCCL send/recv instead of CUDA-IPC peer loads")
    ap.add_argument("--no-ensemble", action="store_true", help="skip the two-member variant of the host-buffer leg (a second 512^3 context: +21 GB)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (1024^3 on 8 GPUs would pin 80 GB of host memory)")
    ap.add_argument("--workload", default="bubble", choices=["bubble", "supercell", "bomex"],
                    help="bubble: BASELINE metric workload (512^3 anelastic); supercell: BASELINE config 4 (compressible split-explicit); "
                         "bomex: BASELINE config 3 (moist LES with forcings)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"3-D dry thermal bubble {args.size}^3, anelastic, WENO5, SSP-RK3, FP64, dt={args.dt}"

    # ---------------------------------------------------------------- reference arm: the CPU restatement, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return
        steps, warm = max(1, args.steps), max(0, args.warmup)
        note = "CPU restatement of the reference algorithm (oracle/, C + OpenMP); Breeze's own CPU() run needs Julia, which is not installed"
        if args.workload == "supercell":
            cs = (64, 64, 64)
            v, cores, sps = run_oracle_compressible(cs, steps, warm, 6.0, 6)
            sample = f"64x64x64 cells of the same case per step ({steps} steps after {warm} warm-up, {sps:.2f} s/step), {cores} host threads"
            print(json.dumps({
                "impl": "reference", "metric": METRIC, "value": v, "unit": "Mcell-updates/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
                "ms_per_step": sps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "compressible split-explicit WS-RK3, supercell-shaped case, dry, WENO5, dt=6 s, 6 substeps per step; "
                                       "reference arm: bounded sample of 64x64x64 cells per step", "grid": list(cs), "note": note,
                           "same_config_note": "throughput metric on a bounded sample (64x64x64 of the 256x256x64 case): same kernels and substep count per cell-step"},
                "cpu_baseline": {"value": v, "unit": "Mcell-updates/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": v, "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
            return
        # bounded sample: cs^3 cells of the same bubble per step, sized so that steps + warm-up end within a few minutes
        cs = min(args.cpu_size, args.size)
        t0 = time.perf_counter()
        run_oracle(min(cs, 64), 1, 0, args.dt)                                  # library load + first-touch, and a rate estimate
        est = (time.perf_counter() - t0)
        while cs > 64 and est * (cs / 64.0) ** 3 * (steps + warm) > 240.0:
            cs //= 2
        v, cores, sps = run_oracle(cs, steps, warm, args.dt)
        sample = f"{cs}^3 cells of the same bubble per step ({steps} steps after {warm} warm-up, {sps:.2f} s/step), {cores} host threads"
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "Mcell-updates/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": sps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload + f"; reference arm: bounded sample of {cs}^3 cells of it per step", "grid": [cs, cs, cs], "note": note,
                       "same_config": cs == args.size,
                       "same_config_note": f"Mcell-updates/s is intensive: the CPU arm steps {cs}^3 cells of the same bubble (same extents, dt, scheme) "
                                           f"because a {args.size}^3 oracle step takes ~{(args.size / cs) ** 3 * sps:.0f} s on these {cores} threads"},
            "cpu_baseline": {"value": v, "unit": "Mcell-updates/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    if args.workload == "supercell":
        if rank == 0:
            print(json.dumps(bench_compressible(args, args.steps, args.warmup)))
        return
    if args.workload == "bomex":
        if rank == 0:
            print(json.dumps(bench_bomex(args, args.steps, args.warmup)))
        return

    # ---------------------------------------------------------------- our arm
    import torch
    import torch.distributed as dist
    import breeze_b200 as bz
    from breeze_b200 import abi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libbreeze_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    def new_uid():
        """A fresh ncclUniqueId from rank 0 (one per communicator: every multi-rank context needs its own)."""
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.tensor(list(abi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        return bytes(buf.cpu().tolist())

    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        uid = new_uid()
    if args.gpus != world:
        if rank == 0:
            sys.stderr.write(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}\n")
        args.gpus = world

    arch = bz.B200(device=local_rank, rank=rank, n_ranks=world, nccl_unique_id=uid, use_tma=args.use_tma, z_chunks=args.z_chunks)
    N = args.size
    grid = bz.RectilinearGrid(arch, size=(N, N, N), x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
    model = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)), advection=bz.WENO(order=5))
    if world > 1 and not args.no_peer_memory:
        bz.enable_peer_memory(model)
    model.set(θ=bubble)
    ctx = model.context
    cells = N ** 3
    cells_local = cells // world

    def barrier():
        ctx.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ext_stream = torch.cuda.ExternalStream(ctx.stream())

    # warm-up
    for _ in range(max(args.warmup, 3)):
        ctx.time_step(args.dt)
    barrier()

    # timed region: K steps, CUDA events on the launching stream, per-kernel-family events for the roofline. The library recycles
    # its profiling events: one untimed profiled pass of the same length fills the pool, so no event is created inside the timed region.
    ctx.profile_enable(True)
    for _ in range(args.steps):
        ctx.time_step(args.dt)
    ctx.profile_read()                                     # reset the accumulators, refill the pool
    launches0 = ctx.kernel_launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(ext_stream)
    for _ in range(args.steps):
        ctx.time_step(args.dt)
    e1.record(ext_stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.kernel_launch_count() - launches0
    fam_ms, fam_n = ctx.profile_read()
    ctx.profile_enable(False)
    ms_per_step = ms / args.steps
    value = cells / (ms_per_step * 1e-3) / 1e6

    # roofline of the dominant kernel (fused stage kernel = family 0), this rank's slab
    peak, peak_src = measured_peak()
    stage_ms = fam_ms[0] / max(1, fam_n[0])
    achieved = STAGE_BYTES_PER_CELL * cells_local / (stage_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "stage_kernel_traffic.json")
    if os.path.exists(tp):
        try:
            t = json.load(open(tp))
            if t.get("size") == N and t.get("n_gpus") == world:
                traffic = t.get("dram_bytes_per_launch")                  # from the committed ncu capture named in the file (commit stamped there)
                traffic_src = {k: t.get(k) for k in ("source", "commit", "launches_captured")}
        except Exception:
            pass
    families = ["stage_tendency_rk", "poisson_forward", "thomas", "poisson_inverse", "projection_halo", "exchange"]
    breakdown = {families[f]: round(fam_ms[f] / args.steps, 4) for f in range(len(families))}

    # size-independent sanity properties of the state after the timed steps (the projection leaves div(ρu) at round-off)
    checks = {"max_abs_divergence": ctx.max_abs_divergence(), "cell_advection_timescale_s": ctx.cell_advection_timescale()}

    # e2e: HOST buffers in, HOST buffers out, every step, through the C ABI
    e2e_steps = max(1, args.e2e_steps)
    if args.no_e2e:
        e2e_steps = 0
    shapes = [ctx.shape(f) for f in range(5)] if e2e_steps else []
    pinned_in = [torch.empty(s, dtype=torch.float64).pin_memory() for s in shapes]
    pinned_out = [torch.empty(s, dtype=torch.float64).pin_memory() for s in shapes]
    if e2e_steps:
        ctx.get_state([t.numpy() for t in pinned_in])
    h2d = sum(int(np.prod(s)) * 8 for s in shapes)
    d2h = h2d
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        # host buffers in, one step, host buffers out — through the asynchronous C-ABI calls: chunked strided copies on two copy streams,
        # so the download of a step and the upload of the next (from the very buffers that download fills) run full duplex over PCIe
        ctx.set_state_async([t.numpy() for t in pinned_in])
        ctx.time_step(args.dt)
        ctx.get_state_async([t.numpy() for t in pinned_out])
        pinned_in, pinned_out = pinned_out, pinned_in
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / max(1, e2e_steps))
    e2e_value = cells / e2e_s / 1e6 if e2e_steps else None

    # The same end-to-end pattern with TWO independent members (an ensemble pair, each with its own context and pinned buffers) stepped
    # alternately: one member's step then runs underneath the other member's copies, so the PCIe link stays busy both ways all the time.
    # Reported beside the dependent chain above, which remains the headline e2e value.
    e2e_pair = None
    if e2e_steps and world == 1 and not args.no_ensemble:
        try:
            second = bubble_model(bz.B200(device=local_rank, use_tma=args.use_tma), N)
            members = [(ctx, pinned_in, pinned_out),
                       (second.context, [torch.empty(s, dtype=torch.float64).pin_memory() for s in shapes],
                        [torch.empty(s, dtype=torch.float64).pin_memory() for s in shapes])]
            second.context.get_state([t.numpy() for t in members[1][1]])
            for c2, _, _ in members:
                c2.synchronize()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                for m_i, (c2, bin_, bout) in enumerate(members):
                    c2.set_state_async([t.numpy() for t in bin_])
                    c2.time_step(args.dt)
                    c2.get_state_async([t.numpy() for t in bout])
                    members[m_i] = (c2, bout, bin_)
            for c2, _, _ in members:
                c2.synchronize()
            torch.cuda.synchronize()
            pair_s = (time.perf_counter() - t0) / (2 * e2e_steps)
            e2e_pair = {"value": cells / pair_s / 1e6, "unit": "Mcell-updates/s", "members": 2, "steps_per_member": e2e_steps,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "note": "two independent 512^3 members on one GPU, stepped alternately; every member-step uploads its five inputs and downloads its five results"}
            del second, members
        except Exception as e:
            e2e_pair = {"error": str(e)}

    PROG = ["ρu", "ρv", "ρw", "ρθ", "ρq"]

    def rel_diff(a, b):
        sc = float(np.abs(b).max())
        d = float(np.abs(a - b).max())
        return d / sc if sc > 0 else d

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cs = min(args.cpu_size, N)
        v, cores, sps, cpu_model, done = run_oracle(cs, 3, 1, args.dt, keep_model=True)
        cpu = {"value": v, "unit": "Mcell-updates/s", "cores": cores, "kind": "port",
               "sample": f"{cs}^3 cells of the same bubble, {done} steps after 1 warm-up, {sps:.2f} s/step (CPU restatement of the reference algorithm)"}
        # parity at BASELINE config 1's own size: the CUDA path steps the same cs^3 bubble the CPU leg just stepped (same ICs, dt, step count)
        # and the prognostic fields are compared (FP64, relative to each field's max-norm; tolerance as in tests/test_gpu_parity.py)
        try:
            gm = bubble_model(bz.B200(device=local_rank, use_tma=args.use_tma), cs)
            for _ in range(done + 1):
                gm.time_step(args.dt)
            errs = {n: rel_diff(gm.field(n), cpu_model.field(n)) for n in PROG}
            tol = 1e-8
            checks[f"parity_{cs}"] = {"grid": [cs, cs, cs], "steps": done + 1, "max_rel": errs, "tol": tol, "ok": bool(max(errs.values()) < tol),
                                      "against": "CPU oracle, reference (quadratic) smoothness-indicator form, all five prognostics"}
            del gm
        except Exception as e:
            checks[f"parity_{cs}"] = {"error": str(e), "ok": False}
        del cpu_model
        try:                                                          # at one thread and at all cores
            c1 = min(128, cs)
            v1, _, sps1 = run_oracle(c1, 1, 1, args.dt, threads=1)
            cpu["single_thread"] = {"value": v1, "unit": "Mcell-updates/s", "cores": 1, "sample": f"{c1}^3 cells, 1 step after 1 warm-up, {sps1:.2f} s/step"}
        except Exception as e:                                        # never lose the headline line
            cpu["single_thread"] = {"error": str(e)}

    # multi-GPU parity carried by the scaling records: a 64 x 32 x 32 bubble with shear and moisture stepped on all ranks and on rank 0 alone
    if world > 1:
        try:
            def small(arch):
                g = bz.RectilinearGrid(arch, size=(64, 32, 32), x=(-10e3, 10e3), y=(-5e3, 5e3), z=(0, 10e3))
                m = bz.AtmosphereModel(g, dynamics=bz.AnelasticDynamics(bz.ReferenceState(g, potential_temperature=300)), advection=bz.WENO(order=5))
                if arch.n_ranks > 1 and not args.no_peer_memory:
                    bz.enable_peer_memory(m)
                m.set(θ=lambda x, y, z: 300 + 2 * np.cos(np.pi / 2 * np.minimum(1, np.sqrt((x - 3000) ** 2 + y ** 2 + (z - 2000) ** 2) / 2000)) ** 2,
                      u=lambda x, y, z: 5 + np.sin(2 * np.pi * x / 20e3) * np.cos(2 * np.pi * y / 10e3) + 0 * z,
                      v=lambda x, y, z: -2 + np.cos(2 * np.pi * x / 20e3) + 0 * y + 0 * z,
                      qᵗ=lambda x, y, z: 0.01 * np.exp(-z / 3000) * (1 + 0.1 * np.sin(2 * np.pi * x / 20e3)) + 0 * y)
                return m
            sm = small(bz.B200(device=local_rank, rank=rank, n_ranks=world, nccl_unique_id=new_uid()))
            for _ in range(3):
                sm.time_step(1.0)
            worst = {}
            ref1 = None
            if rank == 0:
                ref1 = small(bz.B200(device=local_rank))
                for _ in range(3):
                    ref1.time_step(1.0)
            for n in PROG + ["φ"]:
                mine = torch.from_numpy(sm.field(n)).cuda()
                parts = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(parts, mine)
                if rank == 0:
                    worst[n] = rel_diff(torch.cat(parts, dim=2).cpu().numpy(), ref1.field(n))
            if rank == 0:
                checks["vs_single_gpu"] = {"grid": [64, 32, 32], "steps": 3, "ranks": world, "max_rel": worst, "tol": 1e-11,
                                           "ok": bool(max(worst.values()) < 1e-11)}
            del sm, ref1
        except Exception as e:
            checks["vs_single_gpu"] = {"error": str(e), "ok": False}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "Mcell-updates/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "grid": [N, N, N], "parallelism": f"x-slabs x{world}" + ("" if world == 1 else (", NCCL send/recv" if args.no_peer_memory else ", CUDA-IPC peer loads + NCCL barrier")), "l2": "inputs larger than L2 (5 fields x %.2f GB per rank)" % (cells_local * 8 / 1e9),
                       "staging": "tma" if args.use_tma != 2 else "plain", "device_bytes": ctx.device_bytes()},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "stage_kernel (fused WENO5 tendencies + RK update)", "kernel_ms": stage_ms, "peak_source": peak_src,
                         "bytes_per_cell": STAGE_BYTES_PER_CELL,
                         "traffic_source": traffic_src},   # dram__bytes_read + write per launch from the committed ncu --set full capture of this kernel
            "breakdown_ms_per_step": breakdown,
            "e2e": {"value": e2e_value, "unit": "Mcell-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "bz_set_state_async + bz_time_step + bz_get_state_async per step (pinned host buffers, all five prognostics both ways), bz_synchronize at the end",
                    "two_independent_members": e2e_pair},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "checks": checks,
        }
        if cpu:
            out["cpu_baseline"] = cpu
        if world == 1:
            # the second hot-path family (BASELINE config 4) rides along as a sub-record; `--workload supercell` gives its full line
            try:
                del model, ctx
                c4 = bench_compressible(args, 10, 3, with_cpu=False, with_e2e=False)
                out["config4_compressible"] = {k: c4[k] for k in ("value", "unit", "ms_per_step", "roofline", "breakdown_ms_per_step", "gpu_launches")}
                out["config4_compressible"]["workload"] = c4["config"]["workload"]
                try:    # the precision examples/splitting_supercell.jl:86 sets: same case through the Float32 library (bzcf_*), device time of 10 steps
                    m32 = supercell_model(bz.B200(device=local_rank, float_type="Float32"), (256, 256, 64), 6)
                    for _ in range(3):
                        m32.time_step(6.0)
                    m32.context.synchronize()
                    st32 = torch.cuda.ExternalStream(m32.context.stream())
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st32)
                    for _ in range(10):
                        m32.time_step(6.0)
                    e1.record(st32)
                    m32.context.synchronize(); torch.cuda.synchronize()
                    ms32 = e0.elapsed_time(e1) / 10
                    out["config4_compressible"]["float32_mode"] = {"value": 256 * 256 * 64 / (ms32 * 1e-3) / 1e6, "unit": "Mcell-updates/s", "ms_per_step": ms32,
                                                                   "dtype": "f32", "finite": bool(np.isfinite(m32.field("w")).all())}
                    del m32
                except Exception as e:
                    out["config4_compressible"]["float32_mode"] = {"error": str(e)}
            except Exception as e:                       # never lose the headline line
                out["config4_compressible"] = {"error": str(e)}
            try:
                c3 = bench_bomex(args, 20, 3, with_cpu=False)
                out["config3_bomex"] = {k: c3[k] for k in ("value", "unit", "ms_per_step", "breakdown_ms_per_step", "gpu_launches", "checks")}
                out["config3_bomex"]["workload"] = c3["config"]["workload"]
            except Exception as e:
                out["config3_bomex"] = {"error": str(e)}
            try:
                out["float32_mode"] = bench_float32(args, N)
            except Exception as e:
                out["float32_mode"] = {"error": str(e)}
            try:
                out["shipped_scheme_weno9_static_energy"] = bench_shipped_bubble(args)
            except Exception as e:
                out["shipped_scheme_weno9_static_energy"] = {"error": str(e)}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and any(isinstance(v, dict) and v.get("ok") is False for v in checks.values()):
        sys.stderr.write("bench.py: a parity check FAILED (see checks in the JSON line)\n")
        raise SystemExit(1)


if __name__ == "__main__":
    main()
