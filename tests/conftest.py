import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_arch():
    from oracle_lib import CPUOracle, load_oracle_library
    load_oracle_library()
    return CPUOracle()


def bubble_theta(theta0=300.0, dtheta=2.0, zc=2000.0, r0=2000.0):
    """θ of the README quick-start bubble (README.md:67-76), 2-D (x, z) or 3-D (x, y, z)."""
    def f(*xyz):
        x, z = xyz[0], xyz[-1]
        r2 = x ** 2 + (z - zc) ** 2
        if len(xyz) == 3:
            r2 = r2 + xyz[1] ** 2
        return theta0 + dtheta * np.cos(np.pi / 2 * np.minimum(1.0, np.sqrt(r2) / r0)) ** 2
    return f


def make_bubble_model(arch, size, flat_y=False, theta0=300.0, microphysics=None, extent=10e3, **arch_kw):
    import breeze_b200 as bz
    if flat_y:
        grid = bz.RectilinearGrid(arch, size=size, x=(-extent, extent), z=(0, 10e3), topology=(bz.Periodic, bz.Flat, bz.Bounded))
    else:
        grid = bz.RectilinearGrid(arch, size=size, x=(-extent, extent), y=(-extent, extent), z=(0, 10e3))
    ref = bz.ReferenceState(grid, potential_temperature=theta0)
    return bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(ref), advection=bz.WENO(order=5), microphysics=microphysics)


def rel_err(a, b):
    """max |a - b| / max |b| (fields with an O(1)-or-larger scale), or absolute when b is identically zero.
    With BZ_PARITY_REPORT=<file> every measured value is appended to that file with the running test's name, so that the
    tolerances stated in the tests can be set (and re-checked) against what a B200 actually measures (profiles/*parity_errors*)."""
    s = np.max(np.abs(b))
    d = np.max(np.abs(a - b))
    e = d / s if s > 0 else d
    report(e)
    return e


def report(value, label=""):
    path = os.environ.get("BZ_PARITY_REPORT")
    if path:
        test = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]
        with open(path, "a") as f:
            f.write(f"{test}\t{label}\t{float(value):.3e}\n")
