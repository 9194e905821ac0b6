"""Tiny anelastic run for compute-sanitizer (racecheck / memcheck / synccheck of the stage kernel's shared-memory protocol):
    compute-sanitizer --tool racecheck --kernel-name kns=stage_kernel python scripts/sanitizer_run.py [use_tma]
2 SSP-RK3 steps of a 32 x 16 x 24 moving bubble with moisture, two z-chunks (exercises the replayed level and the record relay)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import breeze_b200 as bz

tma = int(sys.argv[1]) if len(sys.argv) > 1 else 1
case = sys.argv[2] if len(sys.argv) > 2 else "weno5"          # weno5 | weno9 | f32 | mixed_radix (48 x 24 grid: radix-3 pass of the transforms) | radix57 (40 x 56: radix-5 / 7 passes)
os.environ["BZ_GRAPHS"] = "0"                                 # every launch visible to the tool
size = (48, 24, 24) if case == "mixed_radix" else (40, 56, 24) if case == "radix57" else (32, 16, 24)
arch = bz.B200(use_tma=tma if case in ("weno5", "f32", "mixed_radix", "radix57") else 0, z_chunks=2, float_type="Float32" if case == "f32" else "Float64")
grid = bz.RectilinearGrid(arch, size=size, x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)), advection=bz.WENO(order=9 if case == "weno9" else 5))
m.set(θ=lambda x, y, z: 300 + 2 * np.cos(np.pi / 2 * np.minimum(1, np.sqrt(x ** 2 + y ** 2 + (z - 2000) ** 2) / 2000)) ** 2, u=3.0, v=-2.0,
      qᵗ=lambda x, y, z: 0.01 * np.exp(-z / 3000) + 0 * x + 0 * y)
for _ in range(2):
    m.time_step(2.0)
m.context.synchronize()
print("sanitizer_run done: checksum", float(np.abs(m.field("ρθ")).sum()), "launches", m.context.kernel_launch_count())
