"""Short run for ncu: python scripts/profile_run.py SIZE STEPS [use_tma [Float32|Float64]]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import breeze_b200 as bz
N = int(sys.argv[1]); steps = int(sys.argv[2]); tma = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ftype = sys.argv[4] if len(sys.argv) > 4 else "Float64"
grid = bz.RectilinearGrid(bz.B200(use_tma=tma, float_type=ftype), size=(N, N, N), x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)))
m.set(θ=lambda x, y, z: 300 + 2 * np.cos(np.pi / 2 * np.minimum(1, np.sqrt(x ** 2 + y ** 2 + (z - 2000) ** 2) / 2000)) ** 2)
for _ in range(steps):
    m.time_step(0.5)
m.context.synchronize()
print("done", m.context.kernel_launch_count())
