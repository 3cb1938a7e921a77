"""CPU oracle for the SLCL loss hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The
product package (``soft-labeled-contrastive-learning_b200/slcl``) never imports
this module and raises if its CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * prototype path, EMA class centres, pseudo labels, hard centroids,
    ContrastiveLoss, SupCon/LocalCon/BlockCon: PINNED -- the restatement in
    ``slcl_oracle.py`` is checked against outputs of the reference's own
    functions (``tests/golden/*.npz``, produced by ``oracle/make_golden.py``
    which imports ``/root/reference`` in the build container) and against
    the nine known-answer values of SURVEY.md section 8(c); where the
    reference tree is mounted, ``tests/test_oracle_live_reference.py`` also runs
    the reference's callables and the restatement side by side on randomised
    inputs (losses and gradients, 30 cases; seg losses and ISCL included).
  * soft-label centroid path: pinned against the reference function with the
    one-line repair described in SURVEY.md section 0 (the shipped function
    raises NameError).
  * reversed-Monte-Carlo partition sampler, class-balanced anchor sampler and
    the rectangular (anchors x contrast set) pixel-to-pixel loss: PARITY
    UNPINNED -- the reference fork contains no implementation of them; the
    oracle *defines* them (spec in SURVEY.md section 8(c)) and the square
    special case of the rectangular loss is pinned against ``SupConLoss``.
"""
