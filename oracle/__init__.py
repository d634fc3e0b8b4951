"""CPU oracle for the RSSM hot path of youngers2006/Dreamer.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  Nothing under ``dreamer_b200/`` imports it, and the product
path has no CPU fallback.

It restates, as plain functional fp32 PyTorch/numpy on the CPU, the algorithm of the
reference's hot path (each function cites the reference file:line it follows).  The
restatement is pinned against the reference's own modules run in the build container
(``oracle/make_golden.py`` imports ``/root/reference``, patches only the two RNG draw
sites to consume host-supplied uniforms/normals, and writes ``tests/golden/*.npz``);
``tests/test_oracle_golden.py`` replays those fixtures on every run.  The reference ships
no golden vectors or tests of its own, so these generated fixtures are
the parity anchor.
"""
