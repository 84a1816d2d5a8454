"""Install the UNMODIFIED reference for the `bench.py --impl reference` arm.

    python baseline/install_ref.py            # in the authoring container (the only place /root/reference exists)

The reference is a flat directory of scripts without setup.py / pyproject (so `pip install --no-index --target
baseline/_ref /root/reference` fails: "neither 'setup.py' nor 'pyproject.toml' found" -- recorded in DESIGN.md); of its
14 files only models.py imports.  This script therefore installs what CAN be installed: a byte-for-byte
copy of models.py into the git-ignored baseline/_ref/ (it travels to the GPU box with the snapshot; no reference source
ever enters the repository history), and records its sha256 so the bench line can state which file it ran.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("IINS_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def main():
    src = os.path.join(REF, "models.py")
    if not os.path.exists(src):
        print(f"{src} not found: nothing installed (the reference arm falls back to the oracle port)")
        return 1
    os.makedirs(DST, exist_ok=True)
    pip = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--target",
                          os.path.join(DST, "_pip"), REF], capture_output=True, text=True)
    pip_note = "ok" if pip.returncode == 0 else (pip.stderr.strip().splitlines() or ["failed"])[-1][:200]
    shutil.copyfile(src, os.path.join(DST, "models.py"))
    with open(src, "rb") as f:
        sha = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "sha256": sha, "pip_install": pip_note}, f, indent=1)
    print(f"installed {src} -> {DST}/models.py (sha256 {sha[:16]}...); pip install: {pip_note}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
