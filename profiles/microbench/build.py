"""Builds profiles/microbench/librenv_microbench.so (FMA / Philox peak micro-benchmarks used by bench.py only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "renv_microbench.cu")
LIB_PATH = os.path.join(HERE, "librenv_microbench.so")


def build(force=False):
    deps = [SRC, os.path.join(HERE, "..", "..", "random_envs_b200", "csrc", "renv_philox.cuh")]
    if not force and os.path.isfile(LIB_PATH) and all(os.path.getmtime(d) <= os.path.getmtime(LIB_PATH) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get("NVCC") or ("/usr/local/cuda/bin/nvcc" if os.path.isfile("/usr/local/cuda/bin/nvcc") else "nvcc")
    cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--shared", "-Xcompiler",
           "-fPIC", "-cudart", "static", "-o", LIB_PATH, SRC]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
