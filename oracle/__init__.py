"""CPU oracle for the per-chunk occ/nuc scoring path of NucleoATAC.

TEST INFRASTRUCTURE ONLY.  This package is a float64 numpy/scipy (plus one small
C file) restatement of the reference algorithm, function by function, each citing
the reference file:line it follows.  It is the checker for the CUDA path: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  Nothing under ``nucleoatac_b200/``
(the product) imports it, and the product has no CPU fallback.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks the oracle against
(a) every known-answer test the reference's own test-suite holds for this path
(tests/test_xcor.py, test_var.py, test_occupancy.py, test_chunkmat2d.py,
test_tracks.py, test_utils.py) and (b) the outputs the reference itself shipped
in ``example/example_results`` (12 significant digits), via the fixtures under
``tests/golden/`` that ``tests/golden/make_golden.py`` extracted in the build
container (the reference tree does not travel to the GPU box), and (c)
``tests/test_oracle_pyref.py`` against vectors made by RUNNING the reference's own
``OccChunk.process`` / ``NucChunk.process`` / ``ChunkMat2D.get(flip=True)`` on synthetic
chunks in the build container (``tests/golden/make_golden_pyref.py``: the reference's
Python-2 modules loaded through a compatibility loader): occupancy grids, peaks and
calls equal, every track equal to the last bit.
"""
