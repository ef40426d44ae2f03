"""CPU oracle for the per-pixel loss / Dice-scoring path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``ecologysemanticsegmentation_b200/`` imports this package.  The only
callers are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- always as the checker or the timed CPU
baseline, never as the product path.

Parity status: PINNED.  The reference has no tests or golden vectors of its own
(SURVEY.md section 4), so the pin is the reference itself, executed in the build
container through ``tests/golden/ref_loader.py``:

* ``oracle.torch_port`` is an op-order-faithful restatement of
  ``ecology_semantic_segmentation/loss_functions.py`` and ``loss_composite.py`` in
  torch eager ops.  ``tests/test_oracle_vs_reference.py`` asserts it is *bit-identical*
  to the unmodified reference on CPU (values and autograd gradients) whenever
  ``/root/reference`` is present, and ``tests/golden/*.json|npz`` (written by
  ``tests/golden/make_golden.py`` from the real reference) pin it everywhere else.
* ``oracle.closed_form`` is the float64 sufficient-statistics formulation the CUDA
  kernels implement (sums -> closed-form losses and gradients); it is checked against
  ``torch_port`` and the golden vectors.
* ``oracle.counts`` is the exact-integer thresholded Dice counter.
"""
