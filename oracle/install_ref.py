"""TEST / BASELINE INFRASTRUCTURE — not part of the product path.

Stages the UNMODIFIED reference under baseline/_ref/ (git-ignored, travels to the GPU box with gpurun) so that
`bench.py --impl reference`, the `cpu_baseline` leg and the stock-PyTorch GPU probe can run the reference's own
modules where /root/reference does not exist.

    python oracle/install_ref.py

The reference has no setup.py / pyproject.toml, so the prescribed
`pip install --no-index --target baseline/_ref /root/reference` cannot work ("neither 'setup.py' nor 'pyproject.toml'
found"); the install is a plain tree copy of det3d/ (Python + the DCN extension sources) and configs/.  Nothing is
copied into tracked paths.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def install(force=False):
    """Returns the staged root, or None when /root/reference is absent (GPU box: the staged copy is already there)."""
    if not os.path.isdir(os.path.join(SRC, "det3d")):
        return DST if os.path.isdir(os.path.join(DST, "det3d", "models")) else None
    if os.path.isdir(os.path.join(DST, "det3d", "models")) and not force:
        return DST
    os.makedirs(DST, exist_ok=True)
    keep = (".py", ".cpp", ".cu", ".cuh", ".h")
    for sub in ("det3d", "configs"):
        for dp, dn, fn in os.walk(os.path.join(SRC, sub)):
            rel = os.path.relpath(dp, SRC)
            for f in fn:
                if f.endswith(keep):
                    os.makedirs(os.path.join(DST, rel), exist_ok=True)
                    shutil.copyfile(os.path.join(dp, f), os.path.join(DST, rel, f))
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
