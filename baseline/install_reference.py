"""Recipe for the CPU arm of bench.py: put the UNMODIFIED reference package (`fastgps`, pure Python) under baseline/_ref/.

    python baseline/install_reference.py [/root/reference]

1. tries the contract's pip install (`pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse
   --target baseline/_ref <copy of the reference>`); in this image that fails because the reference's build backend
   (`pdm-backend`, pyproject.toml) is not installed and there is no network;
2. falls back to what that wheel would contain: a byte-for-byte copy of the package directory `fastgps/` (no build step
   exists for it -- the package is pure Python), and records the outcome in baseline/_ref/INSTALL.json.
baseline/_ref/ is git-ignored (reference sources never enter the history) but travels to the GPU box with the snapshot.
The reference's third-party dependency `qmcpy` is absent from the image; bench.py runs the reference on the tests-only
stand-in under oracle/qmcpy_standin, and says so in its JSON line."""
import filecmp
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def install(src="/root/reference"):
    if not os.path.isdir(os.path.join(src, "fastgps")):
        return None
    if os.path.isdir(os.path.join(DEST, "fastgps")) and not filecmp.dircmp(os.path.join(src, "fastgps"), os.path.join(DEST, "fastgps")).diff_files:
        return DEST
    shutil.rmtree(DEST, ignore_errors=True)
    os.makedirs(DEST, exist_ok=True)
    outcome = {"source": src}
    with tempfile.TemporaryDirectory() as tmp:
        cp = os.path.join(tmp, "ref")
        shutil.copytree(src, cp)
        r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
                            "--target", DEST, cp], capture_output=True, text=True)
        outcome["pip_rc"] = r.returncode
        outcome["pip_tail"] = (r.stderr or r.stdout).strip().splitlines()[-1:] if r.returncode else []
    if not os.path.isdir(os.path.join(DEST, "fastgps")):
        shutil.copytree(os.path.join(src, "fastgps"), os.path.join(DEST, "fastgps"), ignore=shutil.ignore_patterns("__pycache__"))
        outcome["method"] = "copied the pure-Python package directory (pip could not build: backend missing)"
    else:
        outcome["method"] = "pip install --target"
    cmp = filecmp.dircmp(os.path.join(src, "fastgps"), os.path.join(DEST, "fastgps"))
    outcome["identical_to_source"] = not (cmp.diff_files or cmp.left_only)
    with open(os.path.join(DEST, "INSTALL.json"), "w") as fh:
        json.dump(outcome, fh, indent=1)
    return DEST


if __name__ == "__main__":
    print(install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
