"""Sweep of the Poisson-transform launch configurations (tuning hooks of setup_poisson):
    python scripts/fft_sweep.py [--size 512] [--steps 5]
For each configuration: per-kernel-family device times per step at SIZE^3 and the max relative difference of the prognostics
after 2 steps of a 64 x 32 x 48 case against the default configuration (the transforms are bit-identical across launch shapes
as long as the per-line arithmetic is unchanged). Development tool; not part of the product path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import breeze_b200 as bz

CONFIGS = [
    ("default (y: 4 lines, 256 thr, 64 regs; x: 80 regs)", {}),
    ("y 80-register build (256, 3) [default until round 2]", {"BZ_FFT_Y_MINB": "3"}),
    ("y wide tiles: 8 lines, 512 thr, 64 regs", {"BZ_FFT_LINES_Y": "8"}),
    ("x 64-register build (256, 4)", {"BZ_FFT_X_MINB": "4"}),
    ("x 4 lines per CTA", {"BZ_FFT_LINES_X": "4"}),
    ("x 4 lines + 64 regs", {"BZ_FFT_LINES_X": "4", "BZ_FFT_X_MINB": "4"}),
    ("y wide + x 64 regs", {"BZ_FFT_LINES_Y": "8", "BZ_FFT_X_MINB": "4"}),
    ("x 64 regs", {"BZ_FFT_X_MINB": "4"}),
]
KEYS = ["BZ_FFT_LINES_Y", "BZ_FFT_LINES_X", "BZ_FFT_Y_MINB", "BZ_FFT_X_MINB"]
FAMILIES = ["stage", "fwd_y+fft_x", "thomas", "fft_x+inv_y", "project_halo", "exchange", "f6", "f7"]


def bubble(x, y, z):
    return 300 + 2 * np.cos(np.pi / 2 * np.minimum(1, np.sqrt(x ** 2 + y ** 2 + (z - 2000) ** 2) / 2000)) ** 2


def model(size):
    grid = bz.RectilinearGrid(bz.B200(), size=size, x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
    return bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)))


def main():
    size = int(sys.argv[sys.argv.index("--size") + 1]) if "--size" in sys.argv else 512
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 5
    ref = None
    for name, env in CONFIGS:
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(env)
        m = model((64, 32, 48))
        m.set(θ=bubble, u=lambda x, y, z: 3 + np.sin(2 * np.pi * x / 20e3) + 0 * y + 0 * z, v=-2.0)
        for _ in range(2):
            m.time_step(2.0)
        fields = [m.field(n) for n in ("ρu", "ρv", "ρw", "ρθ")]
        if ref is None:
            ref, diff = fields, 0.0
        else:
            diff = max(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300) for a, b in zip(fields, ref))
        del m
        m = model((size, size, size))
        m.set(θ=bubble)
        for _ in range(2):
            m.time_step(0.5)
        m.context.synchronize()
        m.context.profile_enable(True)
        for _ in range(steps):
            m.time_step(0.5)
        m.context.synchronize()
        ms, n = m.context.profile_read()
        per = {f: round(ms[i] / steps, 3) for i, f in enumerate(FAMILIES) if n[i]}
        print(f"{name:48s} step={sum(ms) / steps:7.3f} ms  {per}  diff_vs_default={diff:.2e}", flush=True)
        del m


if __name__ == "__main__":
    main()
