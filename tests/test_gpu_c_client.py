"""The C ABI used from C: tests/c_client/abi_client.c links librenv_b200.so + the C oracle, no Python in the loop."""
import os
import shutil
import subprocess

import pytest

from random_envs_b200 import build as lib_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(out):
    lib_build.build()
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    libdir = os.path.join(ROOT, "random_envs_b200")
    cmd = [nvcc, "-O1", "-o", out, os.path.join(ROOT, "tests", "c_client", "abi_client.c"),
           os.path.join(ROOT, "oracle", "cartpole_oracle.c"), "-I", os.path.join(ROOT, "include"), "-L", libdir,
           "-lrenv_b200", "-Xlinker", "-rpath=" + libdir, "-lm"]
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


def test_c_client_compiles_against_the_header(tmp_path):
    """No GPU needed: include/renv.h is valid C and the library exports what the client links."""
    res = _compile(str(tmp_path / "abi_client"))
    assert res.returncode == 0, res.stdout


@pytest.mark.gpu
def test_c_client_runs_and_matches_the_oracle(tmp_path):
    exe = str(tmp_path / "abi_client")
    res = _compile(exe)
    assert res.returncode == 0, res.stdout
    run = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert run.returncode == 0 and "abi_client ok" in run.stdout, run.stdout
