"""CPU oracle for the physics-guided sampler hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (its ``cpu_baseline``
leg and the ``--impl reference`` arm) may import it, and there only as the checker
or as the timed CPU baseline.  The product package
(``dynamical_pde_diffusion_b200``) never imports this module and has no CPU
fallback.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified
reference from ``/root/reference/src`` (stub-package recipe, ``oracle/ref_import.py``)
and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this
restatement against those vectors, and ``tests/test_oracle_vs_reference.py``
checks it against the live reference whenever ``/root/reference`` is present.
The one exception is the scalar reduction of the LLG m x H_eff residual and the
uniaxial-anisotropy term (see ``guided_sampler_ref.llg_residual_loss``): the
reference never reduces that residual to a sampler loss and has K0 = 0, so those
two choices are "parity unpinned" and documented as such in DESIGN.md.
"""
