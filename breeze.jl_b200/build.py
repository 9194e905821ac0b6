"""In-tree build of csrc/libbreeze_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libbreeze_b200.so")
SOURCES = ["api.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-Xcompiler", "-O2"]


def dependencies():
    """Every file the library is compiled from: the sources, all headers under csrc/ (incl. generated tables) and include/."""
    inc = os.path.join(HERE, "..", "include")
    return ([os.path.join(CSRC, s) for s in SOURCES] + sorted(glob.glob(os.path.join(CSRC, "*.cuh")))
            + sorted(glob.glob(os.path.join(inc, "*.h"))) + [os.path.abspath(__file__)])


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(f) > t for f in dependencies())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [NVCC] + FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT, "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
