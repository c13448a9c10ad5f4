"""CPU oracle for the incremental-session path of feyzaakyurek/subspace-reg.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: it may be imported by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- as the checker or
the timed CPU baseline -- and by nothing else.  The product path (``subspace-reg_b200/``) never imports it and has no
CPU fallback.

What it is: a plain PyTorch-CPU (fp32) restatement of the reference algorithm, function by function, each citing the
reference file:line it follows.  The arithmetic of the reference lives in PyTorch itself (third-party, not vendored;
the reference pins no version for its primary environment -- setup.sh:3 -- and ``pytorch==1.7.0`` as the alternative,
setup.sh:9); the oracle calls the same library ops (conv2d, batch_norm, leaky_relu, max_pool2d, dropout, bernoulli,
linear, cross_entropy, norm, qr, SGD/Adam) in the same order, on torch 2.11 CPU.

Pinning: the reference repository holds NO tests, golden vectors or known-answer fixtures for this path
(SURVEY.md section 4 / 8c), so parity is pinned against outputs of the unmodified reference executed in the build
container: ``oracle/make_golden.py`` imports /root/reference (with the ``.cuda()`` shim of
``oracle/reference_harness.py``), records weight trajectories, loss terms, BatchNorm buffers, features and
predictions, and commits them under ``tests/golden/``; ``tests/test_oracle_pins.py`` holds the oracle to them.
"""
