#!/usr/bin/env python
"""Stage the UNMODIFIED reference (baler v1.4.0) into oracle/_ref/ so that it can travel to the GPU box.

TEST INFRASTRUCTURE - NOT PRODUCT CODE.  oracle/_ref/ is git-ignored (nothing of the reference enters the history) but
not gpurun-ignored, so the staged package ships with the snapshot like the built .so files.  bench.py's
`--impl reference` arm and the `cpu_baseline` leg run the staged package (oracle/ref_runner.py).

    python oracle/stage_ref.py            # needs /root/reference (the build container); a no-op elsewhere

The reference is a poetry project (build-backend poetry.core, python < 3.11.10): `pip install --no-index --target` fails
here (no poetry-core in the offline wheelhouse; recorded below), and the package is pure Python with no build step, so
the package directory is staged as it lies.  Two import shims are applied at RUN time by oracle/ref_runner.py, not to
the staged files: the matplotlib stub (oracle/_shims) and ReduceLROnPlateau without the removed `verbose` kwarg.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("BALER_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def stage(verbose=True):
    if not os.path.isdir(os.path.join(REF, "baler")):
        return os.path.isdir(os.path.join(DST, "baler"))
    note = None
    if os.environ.get("BALER_STAGE_TRY_PIP"):
        tmp = tempfile.mkdtemp()
        try:
            src = os.path.join(tmp, "src")
            shutil.copytree(REF, src)
            r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                                "--find-links", "/opt/wheelhouse", "--target", os.path.join(tmp, "t"), src],
                               capture_output=True, text=True)
            note = "pip rc=%d: %s" % (r.returncode, (r.stderr or r.stdout).strip().splitlines()[-1:] or "")
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(os.path.join(REF, "baler"), os.path.join(DST, "baler"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    # the shipped example the CFD arm runs on (60 x 50 x 50 float64, 1.2 MB)
    cfd = os.path.join(REF, "workspaces", "CFD_workspace", "data", "CFD_animation.npz")
    if os.path.exists(cfd):
        shutil.copy(cfd, os.path.join(DST, "CFD_animation.npz"))
    version = "unknown"
    try:
        for line in open(os.path.join(REF, "pyproject.toml")):
            if line.startswith("version"):
                version = line.split("=")[1].strip().strip('"')
                break
    except OSError:
        pass
    json.dump({"source": REF, "package": "baler", "version": version, "method": "package directory staged as it lies",
               "pip": note or "pip install --no-index --target fails: build-backend poetry.core is not installed "
                              "(ModuleNotFoundError: No module named 'poetry'), python_requires < 3.11.10",
               "shims_at_run_time": ["oracle/_shims/matplotlib", "ReduceLROnPlateau(verbose=...) kwarg dropped"]},
              open(os.path.join(DST, "STAGED.json"), "w"), indent=1)
    if verbose:
        print("staged %s %s -> %s" % (REF, version, DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
