"""The drop-in boundary from plain C: tests/c/abi_smoke.c is compiled with gcc against include/*.h and linked to
libbreeze_b200.so — no Python, no torch, plain pointers and sizes only.

CPU: it builds, links, and the library refuses to run without a GPU ("no CPU fallback") through the C error path.
GPU: the C driver's checksums equal the ones of the same case driven through the Python host mirror."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "breeze.jl_b200", "csrc")
EXE = os.path.join(ROOT, "tests", "c", "abi_smoke")


def _build():
    import breeze_b200 as bz
    bz.load_cuda_library()                                   # raises if the product library is missing
    src = os.path.join(ROOT, "tests", "c", "abi_smoke.c")
    cmd = ["gcc", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
           "-L", CSRC, "-lbreeze_b200", "-Wl,-rpath," + CSRC, "-lm"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return EXE


def _have_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.mark.parametrize("path,entry", [("anelastic", "bz_create"), ("compressible", "bzc_create")])
def test_c_driver_builds_and_fails_loudly_without_a_gpu(path, entry):
    exe = _build()
    if _have_gpu():
        pytest.skip("GPU present: covered by the gpu test")
    out = subprocess.run([exe, path, "16", "8", "12", "1"], capture_output=True, text=True)
    assert out.returncode == 3
    assert entry in out.stderr and "no CPU fallback" in out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["anelastic", "compressible"])
def test_c_driver_matches_the_python_host_mirror(path):
    import breeze_b200 as bz
    exe = _build()
    Nx, Ny, Nz, steps = 32, 16, 24, 3
    out = subprocess.run([exe, path, str(Nx), str(Ny), str(Nz), str(steps)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    sums = {int(l.split()[1]): float(l.split()[2]) for l in out.stdout.splitlines() if l.startswith("checksum")}
    bubble = lambda zc: (lambda x, y, z: 300.0 + 2.0 * np.cos(np.pi / 2 * np.minimum(1.0, np.sqrt(x ** 2 + y ** 2 + (z - zc) ** 2) / 2000.0)) ** 2)   # noqa: E731
    if path == "anelastic":
        grid = bz.RectilinearGrid(bz.B200(), size=(Nx, Ny, Nz), x=(-10e3, 10e3), y=(-10e3, 10e3), z=(0, 10e3))
        m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)))
        m.set(θ=bubble(2000.0))
        names = {2: "ρw", 3: "ρθ", 11: "φ"}
    else:
        grid = bz.RectilinearGrid(bz.B200(), size=(Nx, Ny, Nz), x=(-5e3, 5e3), y=(-5e3, 5e3), z=(0, 10e3))
        m = bz.AtmosphereModel(grid, dynamics=bz.CompressibleDynamics(bz.SplitExplicitTimeDiscretization(substeps=6),
                                                                      reference_potential_temperature=300.0))
        _, rho, _ = m.reference_profiles()
        m.set(ρ=np.broadcast_to(rho[:, None, None], m.context.shape(0)).copy(), θ=bubble(3000.0))
        names = {3: "ρw", 4: "ρθ", 10: "p"}
    for _ in range(steps):
        m.time_step(2.0)
    assert "clock 6 3" in out.stdout
    for fid, name in names.items():
        assert sums[fid] == pytest.approx(np.abs(m.field(name)).sum(), rel=1e-9), name
