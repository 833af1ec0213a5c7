"""Model factory -- drop-in for /root/reference/src/network/network.py:5-11 (``Net("CleanUMamba", net_config)``)."""
from .CleanUMamba import CleanUMamba


def Net(network, net_config):
    if network != "CleanUMamba":
        raise NotImplementedError(network)
    return CleanUMamba(**net_config)
