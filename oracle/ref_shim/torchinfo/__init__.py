"""Oracle shim: stub of the absent ``torchinfo`` package (imported at
/root/reference/src/network/CleanUMamba.py:15, used only by its self-test). TEST INFRASTRUCTURE ONLY."""


def summary(*args, **kwargs):
    return None
