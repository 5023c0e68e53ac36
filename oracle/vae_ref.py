"""
TEST INFRASTRUCTURE — CPU restatement (plain torch fp32 functional ops) of the reference's VAE /
rVAE / ImSpec algorithm; state_dict-first like oracle/nets_ref.py.  Only tests/, smoke() and
bench.py's CPU legs may import it.  Parity status: PINNED by tests/test_oracle_vae.py against
goldens produced by the unmodified reference (tests/golden/make_golden_vae.py).
"""
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .nets_ref import conv_block, dilated_block, _conv

SD = Dict[str, torch.Tensor]


def conv_encoder(x, sd: SD, num_layers=2):
    """convEncoderNet.forward — atomai/nets/ed.py:277-289 (ConvBlock lrelu 0.1, no BN)."""
    h = x.unsqueeze(1) if x.ndim in (2, 3) else x.permute(0, -1, 1, 2)
    h = conv_block(h, sd, "conv", num_layers, False, 0.1)
    h = h.reshape(h.shape[0], -1)
    return F.linear(h, sd["fc11.weight"], sd["fc11.bias"]), F.linear(h, sd["fc12.weight"], sd["fc12.bias"])


def fc_encoder(x, sd: SD, num_layers=2):
    """fcEncoderNet.forward — atomai/nets/ed.py:334-343."""
    h = x.reshape(x.shape[0], -1)
    for i in range(num_layers):
        h = torch.tanh(F.linear(h, sd[f"dense.{2*i}.weight"], sd[f"dense.{2*i}.bias"]))
    return F.linear(h, sd["fc11.weight"], sd["fc11.bias"]), F.linear(h, sd["fc12.weight"], sd["fc12.bias"])


def conv_decoder(z, sd: SD, hw, hidden, num_layers=2):
    """convDecoderNet.forward — atomai/nets/ed.py:514-527 (single channel)."""
    h = F.linear(z, sd["fc_linear.weight"]).reshape(-1, hidden, *hw)
    h = conv_block(h, sd, "decoder", num_layers, False, 0.1)
    return _conv(h, sd, "conv_1x1").squeeze(1)


def fc_decoder(z, sd: SD, hw, num_layers=2):
    """fcDecoderNet.forward — atomai/nets/ed.py:569-580 (single channel)."""
    h = z
    for i in range(num_layers):
        h = torch.tanh(F.linear(h, sd[f"decoder.{2*i}.weight"], sd[f"decoder.{2*i}.bias"]))
    return F.linear(h, sd["out.weight"], sd["out.bias"]).reshape(-1, *hw)


def imcoordgrid(hw):
    """atomai/utils/coords.py:47-54."""
    xx = torch.linspace(-1, 1, hw[0])
    yy = torch.linspace(1, -1, hw[1])
    x0, x1 = torch.meshgrid(xx, yy, indexing="ij")
    return torch.stack((x0, x1), 0).reshape(2, -1).T


def r_decoder(z, phi, dx, sd: SD, hw, num_layers=2):
    """transform_coordinates + rDecoderNet.forward — atomai/utils/coords.py:57-83,
    atomai/nets/ed.py:626-642, 672-687 (skip=False, c=1)."""
    b = z.shape[0]
    grid = imcoordgrid(hw).expand(b, -1, -1)
    c, s = torch.cos(phi), torch.sin(phi)
    rot = torch.stack([torch.stack([c, s], 1), torch.stack([-s, c], 1)], 1)
    coord = torch.bmm(grid, rot)
    if dx is not None:
        coord = coord + dx.unsqueeze(1)
    h = F.linear(coord.reshape(-1, 2), sd["coord_latent.fc_coord.weight"],
                 sd["coord_latent.fc_coord.bias"]).reshape(b, hw[0] * hw[1], -1)
    h = torch.tanh(h + F.linear(z, sd["coord_latent.fc_latent.weight"]).unsqueeze(1))
    h = h.reshape(b * hw[0] * hw[1], -1)
    for i in range(num_layers):
        h = torch.tanh(F.linear(h, sd[f"fc_decoder.{2*i}.weight"], sd[f"fc_decoder.{2*i}.bias"]))
    return F.linear(h, sd["out.weight"], sd["out.bias"]).reshape(b, *hw)


def kld_normal(mu, log_sd):
    sd = torch.exp(log_sd)
    return torch.sum(-log_sd + 0.5 * sd**2 + 0.5 * mu**2 - 0.5, -1)


def vae_elbo(x, x_rec, z_mean, z_logsd):
    """vae_loss — atomai/losses_metrics/vi_losses.py:87-108 (mse)."""
    b = x.shape[0]
    rec = 0.5 * torch.sum((x_rec.reshape(b, -1) - x.reshape(b, -1))**2, 1)
    return -rec.mean() - kld_normal(z_mean, z_logsd).mean()


def rvae_elbo(x, x_rec, z_mean, z_logsd, phi_prior=0.1):
    """rvae_loss — atomai/losses_metrics/vi_losses.py:111-137 (mse)."""
    b = x.shape[0]
    rec = 0.5 * torch.sum((x_rec.reshape(b, -1) - x.reshape(b, -1))**2, 1)
    phi_logsd = z_logsd[:, 0]
    kl_rot = -phi_logsd + np.log(phi_prior) + torch.exp(phi_logsd)**2 / (2 * phi_prior**2) - 0.5
    return -rec.mean() - (kld_normal(z_mean[:, 1:], z_logsd[:, 1:]).mean() + kl_rot.mean())


def rvae_forward(x, eps, enc: SD, dec: SD, hw, conv_enc: bool, dx_prior=0.1):
    """rVAE.forward_compute_elbo — atomai/models/dgm/rvae.py:110-147 with injected noise."""
    z_mean, z_logsd = (conv_encoder if conv_enc else fc_encoder)(x, enc)
    z = z_mean + torch.exp(z_logsd) * eps
    x_rec = r_decoder(z[:, 3:], z[:, 0], z[:, 1:3] * dx_prior, dec, hw)
    return rvae_elbo(x, x_rec, z_mean, z_logsd), z_mean, z_logsd, x_rec


def vae_forward(x, eps, enc: SD, dec: SD, hw, conv: bool, hidden=128):
    """VAE.forward_compute_elbo — atomai/models/dgm/vae.py:661-687 with injected noise."""
    z_mean, z_logsd = (conv_encoder if conv else fc_encoder)(x, enc)
    z = z_mean + torch.exp(z_logsd) * eps
    x_rec = conv_decoder(z, dec, hw, hidden) if conv else fc_decoder(z, dec, hw)
    return vae_elbo(x, x_rec, z_mean, z_logsd), z_mean, z_logsd, x_rec


def signal_ed(x, sd: SD, out_len: int, nb_enc=3, nb_dec=4, nbf=64, training=False,
              new_stats: Optional[dict] = None):
    """SignalED.forward for im2spec — atomai/nets/ed.py:66-79, 144-157, 223-228."""
    kw = dict(training=training, new_stats=new_stats)
    h = conv_block(x, sd, "encoder.conv", nb_enc, True, 0.1, **kw)
    z = F.linear(h.reshape(h.shape[0], -1), sd["encoder.fc.weight"], sd["encoder.fc.bias"])
    h = F.linear(z, sd["decoder.fc.weight"], sd["decoder.fc.bias"]).reshape(-1, nbf, out_len)
    h = dilated_block(h, sd, "decoder.dilblock", list(range(1, nb_dec + 1)), True, 0.1, **kw)
    h = conv_block(h, sd, "decoder.conv", 1, True, 0.1, **kw)
    return _conv(h, sd, "decoder.out")
