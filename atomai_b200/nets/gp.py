"""
Deep-kernel-learning modules (atomai/nets/gp.py:14-60).

  * fcFeatureExtractor — the MLP (feat_dim -> 1000 -> 500 -> 50 -> embedim, ReLU) as an
    nn.Sequential with the reference's module names (`linear1`, `relu1`, ...), executed as a chain
    of native GEMM kernels with fused bias + ReLU epilogues;
  * DeepKernel — feature extractor followed by the dense RBF / Matern-2.5 Gram kernel
    (atomai_b200/csrc/gram.cu): K = os * k(||(f(x1) - f(x2)) / lengthscale||).  The reference never
    forms a dense Gram: it configures gpytorch's KISS-GP (GridInterpolationKernel) around the same
    ScaleKernel(RBFKernel(ard)) (nets/gp.py:41-46).  gpytorch is not vendored by the reference and
    not installed in this image; GPRegressionModel therefore needs gpytorch at call time and the
    dense kernel is exposed for the covariance evaluation itself.
"""
from typing import Optional

import torch

from .. import engine, ops
from ..engine import Tape


class fcFeatureExtractor(torch.nn.Sequential):
    """MLP feature extractor (atomai/nets/gp.py:14-26)."""
    def __init__(self, feat_dim, embedim, **kwargs):
        super(fcFeatureExtractor, self).__init__()
        hidden_dim = kwargs.get("hidden_dim")
        hidden_dim = [1000, 500, 50] if hidden_dim is None else list(hidden_dim)
        hidden_dim.append(embedim)
        self.add_module("linear1", torch.nn.Linear(feat_dim, hidden_dim[0]))
        for i, h in enumerate(hidden_dim[1:]):
            self.add_module('relu{}'.format(i+1), torch.nn.ReLU())
            self.add_module('linear{}'.format(i+2), torch.nn.Linear(hidden_dim[i], h))

    def _emit(self, tape: Tape, x):
        mods = list(self.children())
        for i, m in enumerate(mods):
            if isinstance(m, torch.nn.Linear):
                relu_next = i + 1 < len(mods) and isinstance(mods[i + 1], torch.nn.ReLU)
                x = tape.linear(x, m, "relu" if relu_next else None)
        return x

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dtype != torch.float32:
            raise NotImplementedError("the native feature extractor computes in fp32 "
                                      "(precision='single'); float64 is not implemented")
        b = x.shape[0]
        y = engine.run_multi(self, self._emit, (x.reshape(b, 1, 1, -1).contiguous(),))
        return y.reshape(b, -1)


def dense_gram(x1: torch.Tensor, x2: torch.Tensor, lengthscale: torch.Tensor,
               outputscale: float = 1.0, kind: str = "rbf") -> torch.Tensor:
    """K[i, j] = outputscale * k(||(x1_i - x2_j) / lengthscale||) with gpytorch's RBFKernel /
    MaternKernel(nu=2.5) + ScaleKernel definitions, as one fused tiled kernel (no n1 x n2 distance
    matrix in HBM, exponentiation in the epilogue).  Forward only."""
    assert x1.is_cuda and x1.dtype == torch.float32 and x1.dim() == 2 and x2.dim() == 2
    inv_ls = (1.0 / lengthscale.to(x1.device, torch.float32).reshape(-1)).contiguous()
    if inv_ls.numel() == 1:
        inv_ls = inv_ls.expand(x1.shape[1]).contiguous()
    out = torch.empty((x1.shape[0], x2.shape[0]), device=x1.device, dtype=torch.float32)
    ops.gram(x1.contiguous(), x2.contiguous(), inv_ls, float(outputscale),
             {"rbf": 0, "matern": 1}[kind], out)
    return out


class DeepKernel(torch.nn.Module):
    """k(x, x') = outputscale * RBF((f(x) - f(x')) / lengthscale) with f = fcFeatureExtractor."""
    def __init__(self, feat_dim: int, embedim: int = 2, kind: str = "rbf", **kwargs):
        super().__init__()
        self.feature_extractor = fcFeatureExtractor(feat_dim, embedim, **kwargs)
        self.raw_lengthscale = torch.nn.Parameter(torch.zeros(embedim))
        self.raw_outputscale = torch.nn.Parameter(torch.zeros(()))
        self.kind = kind

    @property
    def lengthscale(self):
        return torch.nn.functional.softplus(self.raw_lengthscale)

    @property
    def outputscale(self):
        return torch.nn.functional.softplus(self.raw_outputscale)

    def forward(self, x1: torch.Tensor, x2: Optional[torch.Tensor] = None) -> torch.Tensor:
        with torch.no_grad():
            z1 = self.feature_extractor(x1)
            z2 = z1 if x2 is None else self.feature_extractor(x2)
            return dense_gram(z1, z2, self.lengthscale, float(self.outputscale), self.kind)


def GPRegressionModel(*args, **kwargs):
    """DKL GPR module of the reference (atomai/nets/gp.py:29-60): needs gpytorch (ExactGP, KISS-GP),
    which this build does not vendor."""
    try:
        import gpytorch  # noqa: F401
    except ImportError as e:
        raise ImportError(
            "GPRegressionModel/dklGPR configure gpytorch's KISS-GP machinery (CG / Lanczos), which "
            "is third-party code outside the accelerated hot path; install gpytorch to use them. "
            "atomai_b200 provides the feature extractor (fcFeatureExtractor) and the dense "
            "deep-kernel Gram evaluation (DeepKernel / dense_gram) natively.") from e
    raise NotImplementedError("gpytorch-backed GPRegressionModel is not wired in this round")
