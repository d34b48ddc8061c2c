"""`oracle/build_ref.py` installs the unmodified reference into `oracle/_ref` (git-ignored, shipped to the GPU box);
`oracle/ref_import.py` finds it when `/root/reference` is absent."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isfile("/root/reference/pyproject.toml"), reason="reference tree not present (GPU box)")
def test_build_ref_installs_and_matches_the_source_tree():
    from oracle import build_ref

    target = build_ref.build_ref()
    assert target and build_ref.installed()
    for rel in ("sampling/sample.py", "sampling/pde_losses.py", "models/nets.py", "models/loss.py"):
        a = open(os.path.join("/root/reference/src/diffusion_pde", rel), "rb").read()
        b = open(os.path.join(target, "diffusion_pde", rel), "rb").read()
        assert a == b, rel                                                        # unmodified
    # it is never tracked: the history stays free of reference sources
    tracked = subprocess.run(["git", "ls-files", "oracle/_ref"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    assert tracked == ""
    ignore = open(os.path.join(ROOT, ".gitignore")).read()
    assert "oracle/_ref/" in ignore
    gpurunignore = os.path.join(ROOT, ".gpurunignore")
    assert not os.path.exists(gpurunignore) or "oracle/_ref" not in open(gpurunignore).read()


def test_ref_import_falls_back_to_the_installed_copy():
    """With the container's source tree hidden, the import recipe must pick up oracle/_ref (what the GPU box sees)."""
    if not os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "diffusion_pde", "sampling", "sample.py")):
        pytest.skip("oracle/_ref not installed")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import oracle.ref_import as ri\n"
            "ri._CANDIDATES[1] = '/nonexistent'\n"
            "assert ri.reference_kind() == '_ref', ri.reference_kind()\n"
            "S, PL, M = ri.import_reference()\n"
            "assert S.JointSampler.__module__ == 'diffusion_pde.sampling.sample' and '_ref' in S.__file__\n"
            "print('ok')\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env={**os.environ, "DPDE_REFERENCE_SRC": ""})
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
