"""Import the unmodified reference (TEST INFRASTRUCTURE ONLY).

The reference's top-level ``diffusion_pde/__init__.py`` pulls h5py / matplotlib /
wandb, which are absent here; its ``sampling`` and ``models`` sub-packages need
only torch + numpy.  Registering an empty ``diffusion_pde`` package whose
``__path__`` points at the reference source lets those two import untouched
(SURVEY.md section 8c).  ``/root/reference`` exists only in the build container,
never on the GPU box: callers must check :func:`reference_available` first.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_SRC = os.environ.get("DPDE_REFERENCE_SRC", "/root/reference/src/diffusion_pde")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "sampling", "sample.py"))


def import_reference():
    """Return ``(sampling, pde_losses, models)`` modules of the unmodified reference."""
    if not reference_available():
        raise ImportError(f"reference source not found under {REFERENCE_SRC}")
    if "diffusion_pde" not in sys.modules:
        pkg = types.ModuleType("diffusion_pde")
        pkg.__path__ = [REFERENCE_SRC]
        sys.modules["diffusion_pde"] = pkg
    import diffusion_pde.sampling as sampling  # noqa: E402
    from diffusion_pde.sampling import pde_losses  # noqa: E402
    import diffusion_pde.models as models  # noqa: E402

    return sampling, pde_losses, models
