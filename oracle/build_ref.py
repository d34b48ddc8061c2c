"""Install the UNMODIFIED reference into ``oracle/_ref`` (TEST INFRASTRUCTURE ONLY).

    python oracle/build_ref.py            # pip install --no-index --no-deps --target oracle/_ref <copy of /root/reference>

The reference is a pure-Python setuptools package (``/root/reference/pyproject.toml``); its sampler
(``src/diffusion_pde/sampling/sample.py:243-363``) and denoiser (``src/diffusion_pde/models/nets.py``) need only
torch + numpy.  ``/root/reference`` does not exist on the GPU box, so this recipe installs the package where the
snapshot carries it: ``oracle/_ref/`` is git-ignored (no reference source enters the history) but NOT
gpurun-ignored, so it travels to the box next to our own built ``.so``.  There it serves two purposes:

* the same-device parity oracle of ``tests/test_gpu_reference.py`` (the real ``JointSampler.sample`` on ``cuda:0``);
* the baselines of ``bench.py`` (``gpu_reference`` leg on the same B200, ``--impl reference`` / ``cpu_baseline`` on
  the host cores).

The product package never imports anything from here (``tests/test_abi_cpu.py::test_product_never_imports_the_oracle``).
``/root/reference`` is read-only and setuptools writes ``build/`` + ``*.egg-info`` into the source tree, so the
install runs from a scratch copy under the system temp directory.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("DPDE_REFERENCE_ROOT", "/root/reference")
TARGET = os.path.join(HERE, "_ref")


def installed() -> bool:
    return os.path.isfile(os.path.join(TARGET, "diffusion_pde", "sampling", "sample.py"))


def build_ref(force: bool = False) -> str | None:
    """Returns the install directory, or None when there is no reference to install from (the GPU box)."""
    if installed() and not force:
        return TARGET
    if not os.path.isfile(os.path.join(REFERENCE, "pyproject.toml")):
        return TARGET if installed() else None
    tmp = tempfile.mkdtemp(prefix="dpde_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REFERENCE, src, ignore=shutil.ignore_patterns(".git", "notebooks", "figures", "*.out"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--no-compile",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"pip install of the reference failed:\n{r.stdout}\n{r.stderr}")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if not installed():
        raise RuntimeError(f"reference installed but {TARGET}/diffusion_pde/sampling/sample.py is missing")
    return TARGET


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv))
