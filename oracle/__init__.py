"""TEST INFRASTRUCTURE ONLY — CPU oracle for the AppleCiDEr hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and only as the checker / the CPU baseline.
The product package ``applecider_b200`` never imports this package and fails
loudly when its CUDA library is missing.

Contents
--------
ref_loader.py   imports the REAL reference modules from /root/reference/src
                (exists only in the build container) behind hyrax/astropy stubs.
models.py       plain-PyTorch CPU restatement ("port") of the reference models
                (HyraxBaselineCLS, SpectraNet, AstroMiNN incl. a restated timm
                ConvNeXt-T, fusion AppleCider).  Pinned bit-exactly against the
                real reference by tests/test_oracle_pinned.py (runs here) and by
                the committed fixtures tests/golden/*.npz (generated from the REAL
                reference by tests/golden/make_golden.py).
preprocess.py   numpy restatement of the array-level preprocessing (P1–P5).
weights.py      deterministic, name-keyed weight generator so that 28 M-parameter
                state_dicts never need to be committed.

Parity status: the reference ships NO golden vectors or known-answer tests for
this path — the oracle is pinned instead against outputs of the
reference itself executed in the build container (fixtures + generating script
committed under tests/golden/).
"""
