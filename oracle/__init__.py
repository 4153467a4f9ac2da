"""oracle/ — TEST INFRASTRUCTURE ONLY.

A CPU restatement (plain PyTorch ops, numpy for the integer/fp32 scalar recipes) of the reference's
hot path: models/unet.py, models/unet_dann.py, utils/metrics.py and train_dann.py:22-49 of
fransiskusbudi/multimodal_segmentation_project.  Only tests/, __graft_entry__.smoke() and bench.py's
CPU-baseline / `--impl reference` legs may import this package; the product
(multimodal_segmentation_project_b200/) never does and has no CPU fallback.

Arithmetic dependency: every operation of the reference's path is a PyTorch library op
(requirements.txt:2 `torch>=2.0.0`, unpinned; this image: torch 2.11.0+cu128, oneDNN CPU kernels).

Pinning: the reference ships NO tests, golden vectors or fixtures for this path,
so by the reference's own tests parity is unpinned.  Instead the restatement is pinned against the
reference ITSELF: oracle/make_golden.py imports the unmodified modules from /root/reference in the
build container, runs them on seeded inputs and writes tests/golden/*.npz; tests/test_oracle_golden.py
checks this restatement against those vectors on every run (bit-exact for metrics and initial weights,
<=1e-6 relative for fp32 forward/backward).
"""
