"""In-tree build of csrc/libbreeze_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libbreeze_b200.so")
OUT_F32 = os.path.join(CSRC, "libbreeze_b200_f32.so")      # the Float32 build of the anelastic path (make_f32.py)
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
    t = min((os.path.getmtime(o) if os.path.exists(o) else 0.0) for o in (OUT, OUT_F32))
    return t == 0.0 or any(os.path.getmtime(f) > t for f in dependencies() + [os.path.join(HERE, "make_f32.py")])


def build(force=False, verbose=False):
    """Compiles both libraries (in parallel): libbreeze_b200.so from csrc/api.cu, libbreeze_b200_f32.so from the retyped copy csrc/f32/."""
    if not force and not needs_build():
        return OUT
    import importlib.util
    spec = importlib.util.spec_from_file_location("bz_make_f32", os.path.join(HERE, "make_f32.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    f32_src = mod.generate()
    jobs = [([NVCC] + FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT, "-ldl"], "build.log"),
            ([NVCC] + FLAGS + ["-DBZ_F32", f32_src, "-o", OUT_F32, "-ldl"], "build_f32.log")]
    procs = [(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), cmd, log) for cmd, log in jobs]
    failed = False
    for proc, cmd, logname in procs:
        out, _ = proc.communicate()
        with open(os.path.join(CSRC, logname), "w") as f:
            f.write(" ".join(cmd) + "\n" + out)
        if proc.returncode != 0:
            sys.stderr.write(out)
            failed = True
        elif verbose:
            print(out)
    if failed:
        raise RuntimeError("nvcc failed")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
