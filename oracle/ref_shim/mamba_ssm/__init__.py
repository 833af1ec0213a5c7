"""TEST INFRASTRUCTURE ONLY (oracle shim) -- not product code.

Stand-in for the un-vendored third-party dependency ``mamba-ssm==1.2.2``
(pinned at /root/reference/environment.yml:29) so that the reference's own
``src/network/CleanUMamba.py`` can be imported unchanged in the build container
(no GPU, no network, wheel absent).  Only the pure-PyTorch ("slow") path of the
dependency is restated, following SURVEY.md Appendix A: ``selective_scan_ref``,
``Mamba.forward`` with ``use_fast_path=False``, ``Mamba.step`` (einsum branch),
``Block.forward`` (non-fused add+norm), ``create_block``, ``_init_weights`` and
``InferenceParams``.  The CUDA / Triton entry points of the real package are
deliberately ``None`` here, exactly as the reference forces them in
/root/reference/src/examples/using_pruning_groups.py:26-27.

Used by ``oracle/make_golden.py`` and ``tests/test_oracle_pin.py`` only.
"""
__version__ = "1.2.2+oracle.shim"
