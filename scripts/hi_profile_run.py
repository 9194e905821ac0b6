"""Short WENO(order = 9) run for ncu: python scripts/hi_profile_run.py SIZE STEPS [order]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import breeze_b200 as bz
N = int(sys.argv[1]); steps = int(sys.argv[2]); order = int(sys.argv[3]) if len(sys.argv) > 3 else 9
grid = bz.RectilinearGrid(bz.B200(), size=(N, N, N), x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)), advection=bz.WENO(order=order))
m.set(θ=lambda x, y, z: 300 + 2 * np.cos(np.pi / 2 * np.minimum(1, np.sqrt(x ** 2 + y ** 2 + (z - 2000) ** 2) / 2000)) ** 2)
for _ in range(steps):
    m.time_step(0.5)
m.context.synchronize()
print("done", m.context.kernel_launch_count())
