"""Import the unmodified reference (TEST INFRASTRUCTURE ONLY).

The reference's top-level ``diffusion_pde/__init__.py`` pulls h5py / matplotlib /
wandb, which are absent here; its ``sampling`` and ``models`` sub-packages need
only torch + numpy.  Registering an empty ``diffusion_pde`` package whose
``__path__`` points at the reference source lets those two import untouched
(SURVEY.md section 8c).

Where the source comes from, in this order:

1. ``$DPDE_REFERENCE_SRC`` when set;
2. ``/root/reference/src/diffusion_pde`` -- the build container only;
3. ``oracle/_ref/diffusion_pde`` -- the pip-installed copy made by ``oracle/build_ref.py`` (git-ignored, shipped to
   the GPU box with the snapshot), which is how the real reference runs next to our kernels on the B200.

Callers must check :func:`reference_available` first.
"""
from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.environ.get("DPDE_REFERENCE_SRC"), "/root/reference/src/diffusion_pde",
               os.path.join(_HERE, "_ref", "diffusion_pde")]


def reference_src() -> str | None:
    for c in _CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "sampling", "sample.py")):
            return c
    return None


REFERENCE_SRC = reference_src()


def reference_available() -> bool:
    return reference_src() is not None


def reference_kind() -> str:
    """"source" (the read-only tree of the build container), "_ref" (the installed copy) or "absent"."""
    src = reference_src()
    if src is None:
        return "absent"
    return "_ref" if os.path.abspath(src).startswith(os.path.join(_HERE, "_ref")) else "source"


def import_reference():
    """Return ``(sampling, pde_losses, models)`` modules of the unmodified reference."""
    src = reference_src()
    if src is None:
        raise ImportError("reference source not found: run `python oracle/build_ref.py` in the build container")
    if "diffusion_pde" not in sys.modules:
        pkg = types.ModuleType("diffusion_pde")
        pkg.__path__ = [src]
        sys.modules["diffusion_pde"] = pkg
    import diffusion_pde.sampling as sampling  # noqa: E402
    from diffusion_pde.sampling import pde_losses  # noqa: E402
    import diffusion_pde.models as models  # noqa: E402

    return sampling, pde_losses, models
