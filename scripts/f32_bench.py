"""Float32 vs Float64 library timing: python scripts/f32_bench.py [--size 512] [--steps 5] [--orders 5,9]
Per precision and order: per-kernel-family device times per step of a SIZE^3 dry bubble. Development tool."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import breeze_b200 as bz

FAMILIES = ["stage", "fwd_y", "thomas", "inv_y", "project_halo", "exchange", "f6", "f7"]


def bubble(x, y, z):
    return 300 + 2 * np.cos(np.pi / 2 * np.minimum(1, np.sqrt(x ** 2 + y ** 2 + (z - 2000) ** 2) / 2000)) ** 2


def main():
    size = int(sys.argv[sys.argv.index("--size") + 1]) if "--size" in sys.argv else 512
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 5
    orders = [int(o) for o in sys.argv[sys.argv.index("--orders") + 1].split(",")] if "--orders" in sys.argv else [5]
    ftypes = sys.argv[sys.argv.index("--types") + 1].split(",") if "--types" in sys.argv else ["Float32", "Float64"]
    for order in orders:
        for ft in ftypes:
            n = size if order == 5 else min(size, 256)
            grid = bz.RectilinearGrid(bz.B200(float_type=ft), size=(n, n, n), x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
            m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)), advection=bz.WENO(order=order))
            m.set(θ=bubble)
            for _ in range(2):
                m.time_step(0.5)
            m.context.synchronize()
            m.context.profile_enable(True)
            for _ in range(steps):
                m.time_step(0.5)
            ms, cnt = m.context.profile_read()
            per = {f: round(float(ms[i]) / steps, 3) for i, f in enumerate(FAMILIES) if cnt[i]}
            tot = float(sum(ms)) / steps
            print(f"{ft} order {order} {n}^3: step={tot:8.3f} ms  stage/launch={float(ms[0]) / max(1, cnt[0]):7.3f} ms  {n ** 3 / tot / 1e3:7.1f} Mcell-updates/s  {per}  "
                  f"finite={m.context.state_is_finite()} max|w|={float(np.abs(m.field('w')).max()):.4f}", flush=True)
            del m


if __name__ == "__main__":
    main()
