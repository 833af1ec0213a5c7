"""GLU activation module -- parameter-free holder mirroring /root/reference/src/network/layers.py:6-41.

In the CUDA path the gate is fused into the 1x1-conv GEMM epilogue (CUM_EPI_GLU_*); this module only exists so
that ``encoder[i][3]`` / ``decoder[j][1]`` keep their place in the module tree (pruning code indexes them)."""
import torch.nn as nn

_ACTS = ("Sigmoid", "ReLU", "SiLU", "GELU")


class Activation(nn.Module):
    def __init__(self, activation: str = "Sigmoid", bypass_channels: int = 0) -> None:
        super().__init__()
        assert activation in _ACTS, f"activation={activation}"
        self.kind = activation
        self.bypass_channels = bypass_channels
        self.activation = getattr(nn, activation)()

    def extra_repr(self) -> str:
        return f"glu={self.kind}, bypass_channels={self.bypass_channels}"

    def forward(self, input):  # pragma: no cover - never on the product path
        raise RuntimeError("cleanumamba_b200.Activation is fused into the GEMM epilogue; call the model's forward()")
