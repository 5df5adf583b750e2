"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the STYLER hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker /
CPU baseline -- never as the thing shipped.  The product path (``styler_b200``) fails loudly when its
CUDA library is missing; it never falls back to this package.

Parity status: pinned.  The reference publishes no golden vectors, so the oracle
is pinned against the *reference itself* run in the dev container (``oracle/ref_shim.py`` imports
``/root/reference`` unmodified; ``oracle/make_golden.py`` dumps its outputs to ``tests/golden``), and
``tests/test_oracle_golden.py`` re-checks the restatement against those fixtures on every run.
The one third-party piece that is absent from ``/root/reference`` -- ``librosa.filters.mel`` v0.7.2 used by
``audio/stft.py:128`` -- is restated from its published (Slaney) algorithm in ``oracle/stft_oracle.py``
and pinned by the librosa doc-example values; that single table is "parity unpinned" by the reference.
"""
