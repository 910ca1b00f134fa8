"""Recipe for ``oracle/_ref/``: the reference's own two hot-path source files, UNMODIFIED, where bench.py can run them.

TEST / BENCH INFRASTRUCTURE (see oracle/__init__.py).  ``/root/reference`` exists only in the build container; the GPU
box gets a snapshot of this repo.  ``oracle/_ref/`` is git-ignored (no reference source ever enters the history) but NOT
gpurun-ignored, so the copies made here travel to the GPU box like the built ``.so`` files do, and
``bench.py --impl reference`` / ``cpu_baseline`` time the REAL reference (``kind: "reference"``) instead of the port.

    python -m oracle.make_ref            # copies random_envs/{random_env,random_cartpole}.py + writes MANIFEST.json

Only these two files are needed: they import nothing but gym / numpy / math and are exec'd by
path under the gym shim (oracle/reference_loader.py); ``random_envs/__init__.py`` (which imports the MuJoCo package) is
deliberately NOT copied.  ``__graft_entry__.build()`` calls ``make()`` whenever the reference tree is mounted.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("RENV_REFERENCE_ROOT", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
FILES = ("random_envs/random_env.py", "random_envs/random_cartpole.py")


def source_available():
    return all(os.path.isfile(os.path.join(REF_SRC, f)) for f in FILES)


def make():
    """Copy the files if the reference tree is mounted; returns the manifest (or None when there is nothing to copy)."""
    if not source_available():
        return None
    manifest = {"source": REF_SRC, "files": {}}
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fh:
            manifest["files"][rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(REF_DST, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    return manifest


if __name__ == "__main__":
    print(json.dumps(make(), indent=1, sort_keys=True))
