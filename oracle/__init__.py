"""CPU oracle for the mridc unrolled-reconstruction inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mridc_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may.  It is the *checker*, never the thing measured as the product or shipped.

What it is: a functional (state_dict-driven, no ``nn.Module``) restatement in PyTorch-CPU fp32 of the
reference arithmetic, each function citing the reference ``file:line`` (relative to the upstream repo
root) that it follows.  The dense arithmetic itself lives in third-party libraries the reference pins
(``torch==1.12.0``: ``torch.fft`` and ``torch.nn.functional.conv2d``; here torch 2.11 runs them).

Pinning: ``oracle/make_golden.py`` imports the *unmodified* reference leaf modules from
``/root/reference`` (stub-import recipe in ``oracle/ref_import.py``), runs them and this restatement on
the same seeded inputs, asserts they agree, and writes the reference's outputs to ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` re-checks the restatement against those committed vectors everywhere
(the GPU box has no ``/root/reference``), and re-checks against the live reference when it is present.
The reference's own numeric known-answer tests (``tests/collections/reconstruction/test_fft.py``) are
restated in ``tests/test_fft_kat.py`` against numpy.  Parity status: PINNED for fft/utils/RIM/VarNet/U-Net
blocks and the model forwards (vs. the live reference, fp32 bit-level or <=1e-6); the SSIM/PSNR metrics
restate scikit-image (not vendored, no reference test touches them) -> "parity unpinned" for metrics only.
"""
