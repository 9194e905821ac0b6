"""Where the host-buffer leg spends its time: python scripts/e2e_probe.py [--size 512]
Times, for one 512^3 context and pinned host buffers: download only, upload only, the two overlapped (the e2e pattern of bench.py without
the step), and the full e2e step; plus plain contiguous cudaMemcpy of the same bytes through torch as the PCIe reference. Development tool."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import breeze_b200 as bz


def main():
    N = int(sys.argv[sys.argv.index("--size") + 1]) if "--size" in sys.argv else 512
    grid = bz.RectilinearGrid(bz.B200(), size=(N, N, N), x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
    m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)))
    m.set(θ=lambda x, y, z: 300 + 2 * np.cos(np.pi / 2 * np.minimum(1, np.sqrt(x ** 2 + y ** 2 + (z - 2000) ** 2) / 2000)) ** 2)
    ctx = m.context
    shapes = [ctx.shape(f) for f in range(5)]
    A = [torch.empty(s, dtype=torch.float64).pin_memory() for s in shapes]
    B = [torch.empty(s, dtype=torch.float64).pin_memory() for s in shapes]
    nbytes = sum(int(np.prod(s)) * 8 for s in shapes)
    ctx.get_state([t.numpy() for t in A])
    ctx.time_step(0.5); ctx.synchronize()

    def timed(label, fn, reps=3):
        fn(); ctx.synchronize(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        ctx.synchronize(); torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print(f"{label:58s} {dt * 1e3:8.2f} ms   {nbytes / dt / 1e9:6.1f} GB/s per direction", flush=True)

    timed("download (get_state_async, strided 3-D, chunked)", lambda: ctx.get_state_async([t.numpy() for t in B]))
    timed("upload (set_state_async, strided 3-D, chunked)", lambda: ctx.set_state_async([t.numpy() for t in A]))

    def both():
        ctx.get_state_async([t.numpy() for t in B])
        ctx.set_state_async([t.numpy() for t in B])
    timed("download then upload of the same buffers (full duplex)", both)

    def step():
        ctx.set_state_async([t.numpy() for t in A])
        ctx.time_step(0.5)
        ctx.get_state_async([t.numpy() for t in A])
    timed("e2e step (upload, step, download)", step)
    timed("step only", lambda: ctx.time_step(0.5))
    d = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
    h = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
    timed("torch contiguous H2D of the same bytes", lambda: d.copy_(h, non_blocking=True))
    timed("torch contiguous D2H of the same bytes", lambda: h.copy_(d, non_blocking=True))


if __name__ == "__main__":
    main()
