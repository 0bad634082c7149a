#!/usr/bin/env python
"""Install the UNMODIFIED reference into baseline/_ref (git-ignored, NOT gpurun-ignored: it travels to
the GPU box with the snapshot, /root/reference does not).

  python baseline/install_reference.py [/root/reference]

1. ``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of the checkout>``
   — the base contract's offline install (the copy is because the checkout is read-only and setuptools
   writes build/ and *.egg-info next to setup.py; --no-deps because torch / numpy / pillow / pyyaml / tqdm
   are already in the image and matplotlib / albumentations are optional for the hot path).  This puts the
   reference's ``unet`` package there: what ``bench.py --impl reference`` imports.
2. The files pip does not package (setup.py excludes them): ``scripts/`` and ``configs/`` are copied next
   to it, byte for byte, so that the drop-in test can run the reference's own ``scripts/train.py`` on the
   B200 package (``python -m unet.dropin baseline/_ref/scripts/train.py``).
Nothing under baseline/_ref is edited, imported by the product path, or committed.
"""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def install(src: str = "/root/reference", force: bool = False) -> str:
    if not os.path.isfile(os.path.join(src, "setup.py")):
        raise RuntimeError(f"no reference checkout at {src}")
    marker = os.path.join(DEST, "unet", "models", "unet.py")
    if os.path.isfile(marker) and not force and filecmp.cmp(marker, os.path.join(src, "unet", "models", "unet.py"),
                                                            shallow=False):
        return DEST
    shutil.rmtree(DEST, ignore_errors=True)
    with tempfile.TemporaryDirectory() as tmp:
        copy = os.path.join(tmp, "reference")
        shutil.copytree(src, copy, ignore=shutil.ignore_patterns(".git", "runs", "__pycache__"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", DEST, copy]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"pip install of the reference failed:\n{r.stdout}\n{r.stderr}")
    for extra in ("scripts", "configs"):
        shutil.copytree(os.path.join(src, extra), os.path.join(DEST, extra), ignore=shutil.ignore_patterns("__pycache__"))
    # the installed package must be the checkout's, file for file
    for dirpath, _, files in os.walk(os.path.join(src, "unet")):
        for f in files:
            if f.endswith(".py"):
                a = os.path.join(dirpath, f)
                b = os.path.join(DEST, os.path.relpath(a, src))
                if not filecmp.cmp(a, b, shallow=False):
                    raise RuntimeError(f"installed file differs from the checkout: {b}")
    return DEST


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    print(install(args[0] if args else "/root/reference", force="--force" in sys.argv))
