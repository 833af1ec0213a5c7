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

    def forward(self, input):
        """Stand-alone call of the module (the model's forward never takes this path: the gate is a GEMM epilogue there)."""
        if self.bypass_channels != 0:
            raise NotImplementedError("cleanumamba_b200: bypass_channels != 0 is not on the shipped path")
        from . import ops
        return ops.glu_forward(input, self.kind)
