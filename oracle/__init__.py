"""CPU oracle for the voxelised multi-view pose path -- TEST INFRASTRUCTURE ONLY.

A CPU restatement of the reference's algorithm for the hot path:
numpy for the geometry / sampling / NMS / soft-argmax arithmetic, plain
``torch.nn.functional`` CPU calls for the dense convolutions (the reference's
own arithmetic lives in the un-vendored third-party dependency PyTorch --
``requirements.txt:3`` un-pinned, ``README.md:43`` pins 1.13.1; this image has
2.11.0 -- whose published operator semantics are restated here).  Every function
cites the reference ``file:line`` it follows.

Parity status: PINNED against golden vectors produced by running the unmodified
reference in the build container (``tests/golden/make_golden.py`` ->
``tests/golden/*.npz``, checked by ``tests/test_oracle_golden.py``).  The
reference itself ships no tests or golden vectors.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package, and only as the checker or
the timed CPU baseline.  Nothing under ``selfpose3d_b200/`` imports it; the
product path raises if its CUDA library is missing.
"""
