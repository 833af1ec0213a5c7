"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference from /root/reference (build container only).

The reference's hot-path module (/root/reference/src/network/CleanUMamba.py) cannot be imported as-is here:
  * it needs ``mamba_ssm`` (CleanUMamba.py:12,14) and ``torchinfo`` (:15) -> provided by ``oracle/ref_shim``;
  * ``src/util/util.py:220-227`` evaluates ``.cuda()`` in a default argument at import time -> on a GPU-less
    host ``nn.Module.cuda`` is made an identity for the duration of the import.
In the build container the reference is imported where it lies.  The GPU box has no /root/reference: there the loader falls
back to the byte-for-byte staged copy under the git-ignored ``baseline/_ref/`` (``oracle/stage_reference.py``, made by
``__graft_entry__.build()``), which only ``bench.py``'s CPU legs use (the reference arm / ``cpu_baseline``).
"""
import os
import sys

import torch

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
REFERENCE_ROOT = os.environ.get("CLEANUMAMBA_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isfile("/root/reference/src/network/CleanUMamba.py") else _STAGED)
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shim")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "network", "CleanUMamba.py"))


def import_reference():
    """Returns the reference's ``src.network.network`` module (exposes ``Net`` and ``CleanUMamba``)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not mounted at {REFERENCE_ROOT}")
    for p in (_SHIM, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "src.network.network" in sys.modules:
        return sys.modules["src.network.network"]
    orig_cuda = torch.nn.Module.cuda
    if not torch.cuda.is_available():
        torch.nn.Module.cuda = lambda self, device=None: self
    try:
        import src.network.network as refnet  # noqa: WPS433 (the unmodified reference)
    finally:
        torch.nn.Module.cuda = orig_cuda
    return refnet
