"""Test infrastructure only: CPU restatement ("port") of the reference pixel-contrast path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and there only
as the checker / the timed CPU baseline.  The shipped path (``contrastiveseg_b200``) never
imports this package and fails loudly when its CUDA library is missing.

Parity status: PINNED.  ``oracle.ref_port`` is checked (tests/test_oracle_vs_reference.py,
tests/test_golden.py) against the reference's own modules imported from /root/reference and
against golden vectors that those modules produced (tests/golden/make_golden.py).
"""
