"""Timing of the WENO(order = 7 / 9) anelastic step: python scripts/hi_order_bench.py [LIB.so ...] [--size 256] [--steps 3] [--orders 9,7,5]
Per library (default: the in-tree one) and order: per-kernel-family device times per step of a SIZE^3 dry bubble. Development tool."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import breeze_b200 as bz
from breeze_b200 import abi

FAMILIES = ["stage", "fwd_y", "thomas", "inv_y", "project_halo", "exchange", "f6", "f7"]


def bubble(x, y, z):
    return 300 + 2 * np.cos(np.pi / 2 * np.minimum(1, np.sqrt(x ** 2 + y ** 2 + (z - 2000) ** 2) / 2000)) ** 2


def main():
    libs = [a for a in sys.argv[1:] if a.endswith(".so")] or [abi.cuda_library_path()]
    size = int(sys.argv[sys.argv.index("--size") + 1]) if "--size" in sys.argv else 256
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 3
    orders = [int(o) for o in sys.argv[sys.argv.index("--orders") + 1].split(",")] if "--orders" in sys.argv else [9, 7, 5]
    form = "StaticEnergy" if "--static-energy" in sys.argv else "LiquidIcePotentialTemperature"
    for path in libs:
        abi._CUDA_LIB = abi.Library(os.path.abspath(path), "bz_", cuda=True)
        for order in orders:
            grid = bz.RectilinearGrid(bz.B200(), size=(size, size, size), x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
            m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)),
                                   advection=bz.WENO(order=order), formulation=form)
            m.set(θ=bubble)
            m.time_step(0.5)
            m.context.synchronize()
            m.context.profile_enable(True)
            for _ in range(steps):
                m.time_step(0.5)
            ms, n = m.context.profile_read()
            per = {f: round(ms[i] / steps, 3) for i, f in enumerate(FAMILIES) if n[i]}
            print(f"{os.path.basename(path):24s} order {order} {size}^3 {form}: step={sum(ms) / steps:8.3f} ms  stage/launch={ms[0] / max(1, n[0]):7.3f} ms  "
                  f"{size ** 3 * steps / sum(ms) / 1e3:7.1f} Mcell-updates/s  {per}", flush=True)
            del m


if __name__ == "__main__":
    main()
