"""
ELBO terms for the VAEs with the reference's function names and maths
(atomai/losses_metrics/vi_losses.py:13-137).  The reconstruction term — the only part that touches
image-sized tensors (2*B*H*W*4 bytes) — is one fused, vectorised reduction kernel that also emits
d/dx_hat for the backward pass; the KL terms act on (B, latent) tensors and stay in torch.
"""
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from .. import ops


class _HalfSqErrFn(torch.autograd.Function):
    """sum_b 0.5 * sum_px (x_hat - x)^2  (a scalar); gradient only w.r.t. x_hat."""

    @staticmethod
    def forward(ctx, x, x_hat):
        if not x_hat.is_cuda:
            raise RuntimeError("atomai_b200 losses run on CUDA (sm_100a) only")
        xc = x.detach().float().contiguous()
        xh = x_hat.detach().float().contiguous()
        assert xc.numel() == xh.numel(), "input and reconstruction differ in size"
        acc = torch.zeros(1, device=xh.device, dtype=torch.float64)
        ops.sqerr_reduce(xc, xh, acc)
        ctx.save_for_backward(xc, xh)
        ctx.shape = x_hat.shape
        return acc.float().reshape(())

    @staticmethod
    def backward(ctx, g):
        xc, xh = ctx.saved_tensors
        d = torch.empty_like(xh)
        ops.sqerr_reduce(xc, xh, None, d, 1.0, g.detach().float().reshape(1).contiguous())
        return None, d.reshape(ctx.shape)


def reconstruction_loss(loss_type: str, in_dim: Tuple[int], x: torch.Tensor,
                        x_reconstr: torch.Tensor, logits: bool = True) -> torch.Tensor:
    """
    Batch-summed reconstruction loss.  NOTE: the reference returns the per-sample vector
    (vi_losses.py:13-37) and every caller immediately takes `.mean()`; here the fused kernel
    returns the scalar SUM over the batch, and `vae_loss`/`rvae_loss` divide by B.
    """
    if loss_type != "mse":
        raise NotImplementedError("the native path implements the 'mse' reconstruction loss "
                                  "(the reference's 'ce' branch needs numpy<2, SURVEY.md §0.10)")
    return _HalfSqErrFn.apply(x, x_reconstr)


def kld_normal(q_param: Tuple[torch.Tensor],
               p_param: Optional[Tuple[torch.Tensor]] = None) -> torch.Tensor:
    """KL divergence between two normal distributions, summed over the latent axis
    (vi_losses.py:40-57)."""
    mu_1, log_sd_1 = q_param
    sd_1 = torch.exp(log_sd_1)
    if p_param is None:
        kl = -log_sd_1 + 0.5 * sd_1**2 + 0.5 * mu_1**2 - 0.5
    else:
        mu_2, log_sd_2 = p_param
        sd_2 = torch.exp(log_sd_2)
        kl = (log_sd_2 - log_sd_1 + 0.5 * (sd_1**2 + (mu_1 - mu_2)**2) / sd_2**2 - 0.5)
    return torch.sum(kl, -1)


def kld_rot(phi_prior: float, phi_logsd: torch.Tensor) -> torch.Tensor:
    """KL divergence for the rotation latent variable (vi_losses.py:77-84)."""
    phi_sd = torch.exp(phi_logsd)
    return -phi_logsd + np.log(phi_prior) + phi_sd**2 / (2 * phi_prior**2) - 0.5


def infocapacity(kl_div: torch.Tensor, capacity: List[float], **kwargs) -> torch.Tensor:
    """Controlled capacity increase (vi_losses.py:224-251)."""
    num_iter = kwargs.get("num_iter", 0)
    cap_min, cap_max, cap_num_iter, cap_gamma = capacity
    cap_current = min((cap_max - cap_min) * num_iter / float(cap_num_iter) + cap_min, cap_max)
    return cap_gamma * torch.abs(cap_current - kl_div)


def vae_loss(recon_loss: str, in_dim: Tuple[int], x: torch.Tensor, x_reconstr: torch.Tensor,
             *args: torch.Tensor, **kwargs: List[float]) -> torch.Tensor:
    """ELBO of a plain VAE (vi_losses.py:87-108)."""
    if len(args) != 2:
        raise ValueError("Pass mean and SD values of encoded distribution as args")
    capacity = kwargs.get("capacity")
    likelihood = -reconstruction_loss(recon_loss, in_dim, x, x_reconstr) / x.size(0)
    kl_div = kld_normal(args).mean()
    if capacity is not None:
        kl_div = infocapacity(kl_div, capacity, num_iter=kwargs.get("num_iter", 0))
    return likelihood - kl_div


def rvae_loss(recon_loss: str, in_dim: Tuple[int], x: torch.Tensor, x_reconstr: torch.Tensor,
              *args: torch.Tensor, **kwargs: Union[List[float], float]) -> torch.Tensor:
    """ELBO of the rotationally invariant VAE (vi_losses.py:111-137): the angle's mean is dropped
    from the KL term, translations stay inside kld_normal."""
    if len(args) != 2:
        raise ValueError("Pass mean and SD values of encoded distribution as args")
    z_mean, z_logsd = args
    phi_prior = kwargs.get("phi_prior", 0.1)
    capacity = kwargs.get("capacity")
    phi_logsd = z_logsd[:, 0]
    z_mean, z_logsd = z_mean[:, 1:], z_logsd[:, 1:]
    likelihood = -reconstruction_loss(recon_loss, in_dim, x, x_reconstr) / x.size(0)
    kl_div = kld_normal([z_mean, z_logsd]).mean() + kld_rot(phi_prior, phi_logsd).mean()
    if capacity is not None:
        kl_div = infocapacity(kl_div, capacity, num_iter=kwargs.get("num_iter", 0))
    return likelihood - kl_div
