"""TEST INFRASTRUCTURE ONLY — CPU oracle for the log-mel front end.

Nothing under ``oracle/`` is part of the product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker or the CPU baseline, never as
the thing shipped.  The product (``tal_asrd_b200``) never imports this package
and raises if its CUDA library is missing.

Parity status: the reference repository ships no tests and no golden vectors
for this path, so by the letter of the rules the oracle is
"parity unpinned" by reference-owned fixtures.  It IS pinned against outputs of
the reference itself: ``oracle/make_golden.py`` imports the unmodified
``tal.asr.models.LogMelSpec`` from ``/root/reference`` (in the build container,
where that tree exists) and freezes its outputs, plus a float64 twin of the
same module, into ``tests/golden/*.npz``; ``tests/test_oracle.py`` checks the
restatement here against those files on every CPU run.
"""
