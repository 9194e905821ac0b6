"""torchrun --nproc-per-node P scripts/multi_gpu_check.py [SIZE] [STEPS]: P-GPU x-slab run vs the 1-GPU run of the same problem.

Prints `MULTI_GPU_OK max_rel=<..>` from rank 0 (agreement to FFT-reordering round-off is expected)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import breeze_b200 as bz
from breeze_b200 import abi

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
CASE = sys.argv[3] if len(sys.argv) > 3 else "bubble"      # "bomex": forcings, flux BCs, saturation adjustment, horizontal means across slabs
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    buf.copy_(torch.tensor(list(abi.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(buf, 0)
uid = bytes(buf.cpu().tolist())


P2P = os.environ.get("BZ_P2P", "1") == "1"


def build(arch):
    if CASE == "bomex":
        m = bz.cases.bomex_model(arch, size=(N, N // 2, 24), extent=100.0 * N, seed=11)
        if P2P:                            # set! ran through the NCCL exchange; the steps below use peer loads
            bz.enable_peer_memory(m)
        return m
    grid = bz.RectilinearGrid(arch, size=(N, N // 2, N // 2), x=(-10e3, 10e3), y=(-5e3, 5e3), z=(0, 10e3))
    m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)))
    if P2P:
        bz.enable_peer_memory(m)
    m.set(θ=lambda x, y, z: 300 + 2 * np.cos(np.pi / 2 * np.minimum(1, np.sqrt((x - 3000) ** 2 + y ** 2 + (z - 2000) ** 2) / 2000)) ** 2,
          u=lambda x, y, z: 5 + np.sin(2 * np.pi * x / 20e3) * np.cos(2 * np.pi * y / 10e3) + 0 * z,
          v=lambda x, y, z: -2 + np.cos(2 * np.pi * x / 20e3) + 0 * y + 0 * z,
          qᵗ=lambda x, y, z: 0.01 * np.exp(-z / 3000) * (1 + 0.1 * np.sin(2 * np.pi * x / 20e3)) + 0 * y)
    return m


m = build(bz.B200(device=local, rank=rank, n_ranks=world, nccl_unique_id=uid))
for _ in range(steps):
    m.time_step(1.0)
tau = m.context.cell_advection_timescale()
div = m.context.max_abs_divergence()
names = ["ρu", "ρv", "ρw", "ρθ", "ρq", "φ"]
mine = {n: torch.from_numpy(m.field(n)).cuda() for n in names}
gathered = {}
for n in names:
    parts = [torch.empty_like(mine[n]) for _ in range(world)]
    dist.all_gather(parts, mine[n])
    gathered[n] = torch.cat(parts, dim=2).cpu().numpy()
if rank == 0:
    ref = build(bz.B200(device=local))
    for _ in range(steps):
        ref.time_step(1.0)
    worst = 0.0
    for n in names:
        a, b = gathered[n], ref.field(n)
        worst = max(worst, np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    t1, d1 = ref.context.cell_advection_timescale(), ref.context.max_abs_divergence()
    ok = worst < 1e-11 and abs(tau - t1) < 1e-9 * t1
    print(f"MULTI_GPU_{'OK' if ok else 'FAIL'} p2p={int(P2P)} ranks={world} max_rel={worst:.3e} tau={tau:.6f}/{t1:.6f} div={div:.2e}/{d1:.2e}")
dist.barrier()
dist.destroy_process_group()
