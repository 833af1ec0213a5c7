"""cleanumamba_b200 -- B200-native (sm_100a) forward / streaming path of CleanUMamba behind the reference's API.

    from cleanumamba_b200.network import Net
    net = Net("CleanUMamba", json_cfg["network_config"]).cuda()
    clean = net(noisy)            # (B, 1, L) fp32 on the GPU
"""
from .CleanUMamba import CleanUMamba  # noqa: F401
from .network import Net  # noqa: F401

__all__ = ["CleanUMamba", "Net"]
