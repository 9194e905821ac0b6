"""Device timing of the compressible split-explicit path at the BASELINE config-4 shape (256 x 256 x 64, WS-RK3, acoustic
substepping): python scripts/compressible_bench.py [Nx Ny Nz] [--steps K] [--substeps N] [--float32]

Prints ms per step, Mcell-updates/s, per-kernel-family device time, and the achieved HBM bandwidth of the two substep kernels
against their algorithmic bytes. Development / profiling tool (also the command profiled with ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import breeze_b200 as bz

FAMILIES = ["slow_tendencies", "stage_setup", "horizontal", "column", "stage_end"]
# algorithmic bytes per cell per launch (FP64): horizontal = read ρu′ ρv′ (ρθ)′ (ρθ)′ˢ⁻ θᴸ Cᴸ p Gρu Gρv, write ρu′ ρv′;
# column = up: read ρ′ (ρθ)′ (ρw)′ ρu′ ρv′ θᴸ Cᴸ Gρ Gρθ Gˢρw, write (ρθ)′ˢ⁻ ρ′★ (ρθ)′★ t (ρw)′;
#          down: read (ρw)′ t ρ′★ (ρθ)′★ θᴸ ρu′ ρv′ ⟨u⟩ ⟨v⟩ ⟨w⟩, write (ρw)′ ρ′ (ρθ)′ ⟨u⟩ ⟨v⟩ ⟨w⟩
BYTES_HORIZONTAL = 8 * (9 + 2)
BYTES_COLUMN = 8 * (10 + 5 + 10 + 6)


def main():
    args = [a for a in sys.argv[1:] if a.isdigit()]
    size = tuple(int(a) for a in args[:3]) if len(args) >= 3 and sys.argv[1].isdigit() else (256, 256, 64)
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 20
    nsub = int(sys.argv[sys.argv.index("--substeps") + 1]) if "--substeps" in sys.argv else 6
    if "--lib" in sys.argv:                                    # A/B of a variant build (scripts/build_variant.sh)
        from breeze_b200 import abi
        abi._CUDA_LIB = abi.Library(os.path.abspath(sys.argv[sys.argv.index("--lib") + 1]), "bz_", cuda=True)
        print("library:", sys.argv[sys.argv.index("--lib") + 1])
    ftype = "Float32" if "--float32" in sys.argv else "Float64"
    print("precision:", ftype)
    grid = bz.RectilinearGrid(bz.B200(float_type=ftype), size=size, x=(0, 168e3), y=(0, 168e3), z=(0, 20e3))
    dyn = bz.CompressibleDynamics(bz.SplitExplicitTimeDiscretization(substeps=nsub), reference_potential_temperature=300.0)
    m = bz.AtmosphereModel(grid, dynamics=dyn)
    _, rho, _ = m.reference_profiles()
    m.set(ρ=np.broadcast_to(rho[:, None, None], m.context.shape(0)).copy(),
          θ=lambda x, y, z: 300.0 + 3.0 * np.exp(-((x - 84e3) ** 2 + (y - 84e3) ** 2) / 10e3 ** 2 - (z - 1500.0) ** 2 / 1500.0 ** 2),
          u=10.0, v=5.0)
    ctx = m.context
    for _ in range(3):
        m.time_step(6.0)
    ctx.synchronize()
    ctx.profile_enable(True)
    ctx.profile_read()
    n0 = ctx.kernel_launch_count()
    import time
    t0 = time.perf_counter()
    for _ in range(steps):
        m.time_step(6.0)
    ctx.synchronize()
    wall = (time.perf_counter() - t0) / steps * 1e3
    ms, n = ctx.profile_read()
    cells = np.prod(size)
    total = ms.sum() / steps
    print(f"grid {size}  substeps/step {nsub} ({sum(ctx.stage_substep_count_and_size(6.0, b)[0] for b in (1/3, 1/2, 1))} per WS-RK3 step)")
    print(f"device {total:.3f} ms/step (wall incl. event overhead {wall:.3f})  {cells / total / 1e3:.1f} Mcell-updates/s  "
          f"{(ctx.kernel_launch_count() - n0) / steps:.0f} launches/step  {ctx.device_bytes() / 2**20:.0f} MiB")
    for i, f in enumerate(FAMILIES):
        if n[i]:
            line = f"  {f:16s} {ms[i] / steps:8.3f} ms/step  {n[i] / steps:5.1f} launches/step  {ms[i] / n[i] * 1e3:8.1f} us/launch"
            if f == "horizontal":
                line += f"  {BYTES_HORIZONTAL * cells / (ms[i] / n[i] * 1e-3) / 1e9:7.0f} GB/s algorithmic"
            if f == "column":
                line += f"  {BYTES_COLUMN * cells / (ms[i] / n[i] * 1e-3) / 1e9:7.0f} GB/s algorithmic"
            print(line)
    w = m.field("w")
    print(f"max|w| = {np.abs(w).max():.4f} m/s, finite = {np.isfinite(w).all()}")


if __name__ == "__main__":
    main()
