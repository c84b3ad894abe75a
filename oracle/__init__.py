"""CPU/torch-fp32 oracle for the Soft-IntroVAE hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package (``soft-intro-vae-for-3d-mri_b200/``); only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and there only as the checker or as the timed CPU
baseline -- never as the thing shipped.

Parity pin: the reference repository has no tests and no golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, generated in the authoring container by ``oracle/gen_golden.py``
(imports ``/root/reference`` read-only) and committed under ``tests/golden/``.
"""
