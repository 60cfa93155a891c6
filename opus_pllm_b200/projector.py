"""CSTP projection and the modality-refinement ("switch") projector on the weight-streaming tcgen05 GEMM.

Mirrors, for inference:
  * `CSTPBase.protein_forward(x)` (cstp_v3/modelling.py:396-400): F.normalize + Linear(1280, 5120);
  * `build_switch_projector(model_args, n_tokens=8)` (multi_modality_v1/model/protein_mlp/builder.py:11-25): callable
    with `.load_state_dict({'0.weight','0.bias','2.weight','2.bias'})`, `.to()`, `__call__(x[B,5120]) -> [B, 8*H]`.
At eval batch sizes both are pure weight streaming (2.5 GB of bf16 weights per call), so the batch is the small operand
of the swap-AB GEMM and erf-GELU / bias live in its epilogue.
"""
from __future__ import annotations

import ctypes as C
import re

import torch

from . import _lib as L
from . import ops


class B200ProteinProjector:
    """`protein_projector` seam (opus_arch.py:68-69,120). `weight` [5120, 1280], `bias` [5120] (Lightning ckpt keys
    `protein_projection.linear.{weight,bias}`)."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor, device="cuda"):
        self.device = torch.device(device)
        self.weight = weight.detach().to(self.device, torch.bfloat16).contiguous()
        self.bias = bias.detach().to(self.device, torch.float32).contiguous()
        self.out_dim, self.in_dim = self.weight.shape

    def to(self, device):
        return self

    def eval(self):
        return self

    def parameters(self):
        return iter((self.weight, self.bias))

    def protein_forward_normalised(self, x_l2: torch.Tensor) -> torch.Tensor:
        return ops.gemm(x_l2, self.weight, epilogue=L.EPI_BF16, bias=self.bias)

    def protein_forward(self, protein_embeddings: torch.Tensor) -> torch.Tensor:
        x = protein_embeddings.to(self.device, torch.float32).contiguous()
        return self.protein_forward_normalised(ops.l2norm(x))


class IdentityProjector:
    """no CSTP checkpoint: the ESM embedding goes straight to the switch projector (opus_arch.py:70-80)."""

    def to(self, device):
        return self

    def protein_forward(self, x):
        return x

    def __call__(self, x):
        return x


class B200SwitchProjector:
    """`switch_projector` seam: 'linear' or 'mlp{N}x_gelu' (N = 2 is what OPUS-PLLM ships)."""

    def __init__(self, in_dim: int, hidden_dim: int, projector_type: str = "mlp2x_gelu", device="cuda"):
        self.device = torch.device(device)
        self.in_dim, self.hidden_dim = in_dim, hidden_dim
        if projector_type == "linear":
            self.depth = 1
        else:
            m = re.match(r"^mlp(\d+)x_gelu$", projector_type)
            if not m:
                raise NotImplementedError(f"unknown switch projector type {projector_type}")
            self.depth = int(m.group(1))
        self.weights: list[torch.Tensor] = []
        self.biases: list[torch.Tensor] = []

    def load_state_dict(self, sd: dict, strict: bool = True):
        keys = [("weight", "bias")] if self.depth == 1 else [(f"{2 * i}.weight", f"{2 * i}.bias") for i in range(self.depth)]
        self.weights, self.biases = [], []
        for wk, bk in keys:
            if wk not in sd or bk not in sd:
                raise KeyError(f"switch projector state dict misses {wk}/{bk}")
            self.weights.append(sd[wk].detach().to(self.device, torch.bfloat16).contiguous())
            self.biases.append(sd[bk].detach().to(self.device, torch.float32).contiguous())
        if self.weights[0].shape != (self.hidden_dim, self.in_dim):
            raise ValueError(f"switch projector: expected first weight {(self.hidden_dim, self.in_dim)}, "
                             f"got {tuple(self.weights[0].shape)}")
        return self

    def to(self, device):
        return self

    def eval(self):
        return self

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if not self.weights:
            raise L.OpusError("switch projector has no weights loaded")
        h = x.to(self.device, torch.bfloat16).contiguous()
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            last = i == len(self.weights) - 1
            h = ops.gemm(h, w, epilogue=L.EPI_BF16 if last else L.EPI_BF16_GELU, bias=b)
        return h


class FusedProjectors:
    """normalised ESM embedding -> 8 soft tokens in one C call (opus_projector_forward): used by generate()."""

    def __init__(self, protein_projector, switch: B200SwitchProjector):
        self.pp, self.sw = protein_projector, switch
        m = L.ProjectorModel()
        ident = not isinstance(protein_projector, B200ProteinProjector)
        m.in_dim = switch.in_dim if ident else protein_projector.in_dim
        m.cstp_dim = switch.in_dim
        m.hidden_dim = switch.hidden_dim
        if not ident:
            m.w_cstp, m.b_cstp = protein_projector.weight.data_ptr(), protein_projector.bias.data_ptr()
        m.w0, m.b0 = switch.weights[0].data_ptr(), switch.biases[0].data_ptr()
        if switch.depth == 2:
            m.w2, m.b2 = switch.weights[1].data_ptr(), switch.biases[1].data_ptr()
        elif switch.depth != 1:
            raise NotImplementedError("fused projector path supports 'linear' and 'mlp2x_gelu'")
        self._m = m

    def __call__(self, x_l2: torch.Tensor) -> torch.Tensor:
        n = x_l2.shape[0]
        dev = x_l2.device
        cstp = torch.empty((n, self._m.cstp_dim), dtype=torch.bfloat16, device=dev)
        h0 = torch.empty((n, self._m.hidden_dim), dtype=torch.bfloat16, device=dev)
        out = torch.empty((n, self._m.hidden_dim), dtype=torch.bfloat16, device=dev)
        rc = L.load().opus_projector_forward(C.byref(self._m), x_l2.data_ptr(), n, cstp.data_ptr(), h0.data_ptr(),
                                             out.data_ptr(), None, 0, torch.cuda.current_stream().cuda_stream)
        L.check(rc, "opus_projector_forward")
        return out
