"""A/B timing of library builds: python scripts/variant_bench.py LIB.so [LIB2.so ...] [--size 512] [--steps 5]

For each shared object: 512^3 dry bubble, per-kernel-family device times (bz_profile_read) per step, and the max
relative difference of the prognostics after 2 steps on a 64x32x48 grid against the FIRST library listed.
Development tool; not part of the product path."""
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


def model(size, **kw):
    grid = bz.RectilinearGrid(bz.B200(**kw), size=size, x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
    return bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)))


def main():
    args = [a for a in sys.argv[1:] if a.endswith(".so")]
    size = int(sys.argv[sys.argv.index("--size") + 1]) if "--size" in sys.argv else 512
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 5
    ref_fields = None
    for path in args:
        abi._CUDA_LIB = abi.Library(os.path.abspath(path), "bz_", cuda=True)
        m = model((64, 32, 48))
        m.set(θ=bubble, u=lambda x, y, z: 3 + np.sin(2 * np.pi * x / 20e3) + 0 * y + 0 * z, v=-2.0,
              qᵗ=lambda x, y, z: 0.01 * np.exp(-z / 3000) + 0 * x + 0 * y)
        for _ in range(2):
            m.time_step(2.0)
        fields = [m.field(n) for n in ("ρu", "ρv", "ρw", "ρθ", "ρq")]
        if ref_fields is None:
            ref_fields, diff = fields, 0.0
        else:
            diff = max(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300) for a, b in zip(fields, ref_fields))
        del m
        m = model((size, size, size), use_tma=1)
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
        print(f"{os.path.basename(path):28s} step={sum(ms) / steps:7.3f} ms  stage/launch={ms[0] / max(1, n[0]):6.3f} ms  {per}  "
              f"diff_vs_first={diff:.2e}", flush=True)
        del m


if __name__ == "__main__":
    main()
