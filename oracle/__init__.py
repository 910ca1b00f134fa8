"""CPU oracle for the RandomCartPole-v0 / DR-sampler hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``random_envs_b200`` (the product) may
import, link or execute anything from this package.  The only allowed callers
are ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs, and there only as the checker or as the timed CPU
baseline -- never as the thing shipped.

Parity status: the reference (gabrieletiboni/random-envs) ships NO tests and NO
golden vectors, so parity is pinned by *executing the
reference's own unmodified source* in the build container
(``oracle/reference_loader.py`` + ``oracle/make_golden.py``) and committing the
resulting vectors under ``tests/golden/``.  Every restatement in this package
(``cartpole_port.py``, ``dr_port.py``, ``cartpole_oracle.c``) is checked
bit-for-bit against those vectors by ``tests/test_oracle_*.py``.

Layout
------
gym_shim.py          ~60-line stand-in for gym==0.21.0 (absent from the image),
                     only used to import the reference source unmodified.
reference_loader.py  loads /root/reference/random_envs/{random_env,random_cartpole}.py
                     by path (build container only; the GPU box has no /root/reference).
make_golden.py       runs the real reference and writes tests/golden/*.npz|json.
cartpole_port.py     scalar pure-Python restatement of step/reset/set_task,
                     TimeLimit and the SyncVectorEnv auto-reset loop.
dr_port.py           restatement of RandomEnv.sample_task + exact target CDFs.
cartpole_oracle.c    plain-C restatement (same libm as CPython => bit-exact),
                     used for bulk parity at sizes the Python port cannot reach.
c_oracle.py          ctypes bindings + build recipe for cartpole_oracle.c.
cpu_bench.py         times the port on the host cores (bench.py cpu_baseline).
"""
