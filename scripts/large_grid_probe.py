"""One large single-GPU grid beyond 512^3: python scripts/large_grid_probe.py SIZE [Float64|Float32] [steps]
Builds the dry bubble at SIZE^3 (host arrays by broadcasting; bails out when the host has too little free memory), steps it and prints the
per-family device times, the device memory in use, max |div(ρu)| and the conservation of ∫ρθ. Development tool (896^3 = 7 · 2^7 is the
largest Float32 case of the reference's memory table, benchmarking/README.md:225-233; in Float64 the reference lists it as not fitting an H200)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

N = int(sys.argv[1]); ftype = sys.argv[2] if len(sys.argv) > 2 else "Float64"; steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
need_gb = 6 * N ** 3 * 8 / 1e9
with open("/proc/meminfo") as f:
    avail_gb = [int(l.split()[1]) for l in f if l.startswith("MemAvailable")][0] / 1e6
print(f"host memory available {avail_gb:.0f} GB, needed ≈ {need_gb:.0f} GB", flush=True)
if avail_gb < 1.5 * need_gb:
    sys.exit("not enough host memory for the initial condition arrays")

import torch
import breeze_b200 as bz

FAMILIES = ["stage", "fwd_y", "thomas", "inv_y", "project_halo", "exchange", "f6", "f7"]
grid = bz.RectilinearGrid(bz.B200(float_type=ftype), size=(N, N, N), x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)), advection=bz.WENO(order=5))
x = grid.xnodes()[None, None, :]; y = grid.ynodes()[None, :, None]; z = grid.znodes()[:, None, None]
r = np.sqrt(x ** 2 + y ** 2 + (z - 2000.0) ** 2) / 2000.0
np.minimum(r, 1.0, out=r)
theta = 300.0 + 2.0 * np.cos(np.pi / 2 * r) ** 2
del r
m.set(θ=theta)
del theta
s0 = float(m.field("ρθ").astype(np.float64).sum())
free, total = torch.cuda.mem_get_info()
print(f"{ftype} {N}^3: device memory in use {(total - free) / 1e9:.1f} GB of {total / 1e9:.0f} GB", flush=True)
m.time_step(0.5)
m.context.synchronize()
m.context.profile_enable(True)
for _ in range(steps):
    m.time_step(0.5)
ms, cnt = m.context.profile_read()
tot = float(sum(ms)) / steps
per = {f: round(float(ms[i]) / steps, 2) for i, f in enumerate(FAMILIES) if cnt[i]}
s1 = float(m.field("ρθ").astype(np.float64).sum())
print(f"{ftype} {N}^3: step={tot:8.2f} ms  {N ** 3 / tot / 1e3:7.1f} Mcell-updates/s  {per}  finite={m.context.state_is_finite()} "
      f"max|div|={m.context.max_abs_divergence():.2e}  d(sum ρθ)/sum={abs(s1 - s0) / abs(s0):.1e}  max|w|={float(np.abs(m.field('w')).max()):.4f}", flush=True)
