"""CPU oracle for the mastering hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / CPU baseline.  The
product path (``python-audio-mastering_b200/``) never imports this package and
fails loudly when the CUDA library is missing.

Contents
--------
``thirdparty.py``  restatement of the two un-vendored dependencies the reference
                   calls on the hot path (pydub 0.25.x ``AudioSegment`` /
                   ``effects.compress_dynamic_range``; pyloudnorm 0.1.x ``Meter``)
                   on top of the REAL stdlib ``audioop`` and REAL ``scipy.signal``.
``refload.py``     imports ``/root/reference/worker/audio_mastering_engine.py``
                   UNCHANGED with those restatements installed as ``sys.modules``
                   shims (authoring container only: the GPU box has no
                   ``/root/reference``).
``port.py``        self-contained numpy/scipy restatement of the reference chain
                   (travels to the GPU box); every function cites the reference
                   file:line it follows.  The per-frame compressor loop has a C
                   restatement (``compressor.c``) cross-checked against the
                   faithful Python/audioop loop.
``make_golden.py`` runs the UNCHANGED reference (via ``refload``) on small seeded
                   inputs and writes ``tests/golden/*.npz``.

Parity pinning
--------------
The reference ships no tests or golden vectors, and pydub /
pyloudnorm are unpinned and absent from this image, so parity at those two
boundaries is *unpinned by the reference*: "parity unpinned" for rows a10 / a13.
First-party arithmetic (ENG:117-227) IS pinned: ``tests/golden`` holds outputs of
the unchanged reference file executed in the authoring container, and
``port.py`` is checked against them bit-for-bit.
"""
