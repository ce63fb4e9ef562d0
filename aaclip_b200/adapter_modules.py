"""Parameter containers with the reference's names and state_dict keys (model/adapter_modules.py:6-26).

`SimpleAdapter.fc` is `Sequential(Linear(bias=False), LeakyReLU())` (key `fc.0.weight`);
`SimpleProj.fc` is the same when relu else a bare `Linear` (key `fc.weight`).  The arithmetic runs in the
CUDA engine (GEMM epilogue), so these modules only hold parameters; calling them raises.
"""
from torch import nn


class _HostOnly(nn.Module):
    def forward(self, x):  # pragma: no cover - guard
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container: its arithmetic runs inside the aaclip_b200 CUDA "
            "engine via AdaptedCLIP.forward / encode_text (no PyTorch fallback)."
        )


class SimpleAdapter(_HostOnly):
    def __init__(self, c_in, c_out=768):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(c_in, c_out, bias=False), nn.LeakyReLU())


class SimpleProj(_HostOnly):
    def __init__(self, c_in, c_out=768, relu=True):
        super().__init__()
        if relu:
            self.fc = nn.Sequential(nn.Linear(c_in, c_out, bias=False), nn.LeakyReLU())
        else:
            self.fc = nn.Linear(c_in, c_out, bias=False)
