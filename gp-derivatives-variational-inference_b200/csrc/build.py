"""Compile every CUDA source of the DSVGP hot path for sm_100a into ONE in-tree shared library:

    gp-derivatives-variational-inference_b200/dsvgp_b200/libdsvgp_b200.so

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.  Object files are
cached under csrc/build/ keyed on source mtime.  Run:  python csrc/build.py [--force] [-v]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "dsvgp_b200", "libdsvgp_b200.so")
SOURCES = ["kdir.cu", "gemm.cu", "chol.cu", "misc.cu", "optim.cu", "data.cu", "tc_prep.cu", "trmm_tc.cu", "api.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _newer(a, b):
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "dsvgp_b200.h"))
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    objs, jobs = [], []
    for s in srcs:
        src, obj = os.path.join(HERE, s), os.path.join(bdir, s[:-3] + ".o")
        objs.append(obj)
        if force or _newer(src, obj) or any(_newer(h, obj) for h in headers):
            jobs.append([NVCC, *FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)

    with ThreadPoolExecutor(max_workers=min(6, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if jobs or not os.path.exists(OUT):
        run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, *objs])  # static cudart: no runtime dependency beyond libcuda
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
