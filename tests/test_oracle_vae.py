"""oracle/vae_ref.py against goldens produced by the unmodified reference (CPU)."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

import golden_utils as gu
from oracle import vae_ref

VAE_CASES = {
    "rvae_conv_32": dict(kind="rvae", conv=True, seed=700),
    "rvae_fc_32": dict(kind="rvae", conv=False, seed=710),
    "vae_conv_32": dict(kind="vae", conv=True, seed=720),
    "vae_fc_32": dict(kind="vae", conv=False, seed=730),
}
HW = (32, 32)
N = 8


def build_vae(name):
    """(model [atomai_b200, CPU], enc sd, dec sd, x, eps, golden)."""
    from atomai_b200.models import VAE, rVAE
    c = VAE_CASES[name]
    if c["kind"] == "rvae":
        m = rVAE(HW, latent_dim=2, conv_encoder=c["conv"])
    else:
        m = VAE(HW, latent_dim=2, conv_encoder=c["conv"], conv_decoder=c["conv"])
    sds = []
    for net, seed in ((m.encoder_net, c["seed"]), (m.decoder_net, c["seed"] + 1)):
        shapes = OrderedDict((k, tuple(v.shape)) for k, v in net.state_dict().items())
        sd = OrderedDict((k, torch.from_numpy(v)) for k, v in gu.fill_state_dict(shapes, seed).items())
        net.load_state_dict(sd)
        sds.append(sd)
    x = torch.from_numpy(gu.images(c["seed"] + 2, N, *HW))
    eps = torch.from_numpy(np.random.RandomState(c["seed"] + 3).randn(N, m.z_dim).astype(np.float32))
    return m, sds[0], sds[1], x, eps, gu.load(name + ".npz")


@pytest.mark.parametrize("name", list(VAE_CASES))
def test_vae_oracle_matches_reference(name):
    c = VAE_CASES[name]
    m, enc, dec, x, eps, gold = build_vae(name)
    enc = {k: v.clone().requires_grad_(True) for k, v in enc.items()}
    dec = {k: v.clone().requires_grad_(True) for k, v in dec.items()}
    fwd = vae_ref.rvae_forward if c["kind"] == "rvae" else vae_ref.vae_forward
    elbo, zm, zl, xr = fwd(x, eps, enc, dec, HW, c["conv"])
    np.testing.assert_allclose(zm.detach().numpy(), gold["z_mean"], atol=2e-5)
    np.testing.assert_allclose(zl.detach().numpy(), gold["z_logsd"], atol=2e-5)
    np.testing.assert_allclose(xr.detach().numpy(), gold["x_reconstr"], atol=5e-5)
    assert abs(elbo.item() - float(gold["elbo"])) < 1e-3 * abs(float(gold["elbo"]))
    (-elbo).backward()
    for pref, sd in (("encoder", enc), ("decoder", dec)):
        for k, p in sd.items():
            g = p.grad.numpy()
            ref = gold[f"grad/{pref}.{k}"]
            got = gu.sample_flat(g, 97) if g.size > 4096 else g
            gn = float(gold[f"gradnorm/{pref}.{k}"])
            assert np.abs(got.reshape(-1) - ref.reshape(-1)).max() <= 2e-3 * gn / np.sqrt(g.size) * 30 + 1e-6, k


def build_imspec():
    from atomai_b200.nets import init_imspec_model
    net, _ = init_imspec_model((16, 16), (32,), 4)
    shapes = OrderedDict((k, tuple(v.shape)) for k, v in net.state_dict().items())
    sd = OrderedDict((k, torch.from_numpy(v)) for k, v in gu.fill_state_dict(shapes, 740).items())
    net.load_state_dict(sd)
    x = torch.from_numpy(gu.images(741, 6, 16, 16))[:, None]
    y = torch.from_numpy(np.random.RandomState(742).rand(6, 1, 32).astype(np.float32))
    return net, sd, x, y, gu.load("imspec_16_32.npz")


def test_imspec_oracle_matches_reference():
    net, sd, x, y, gold = build_imspec()
    with torch.no_grad():
        pe = vae_ref.signal_ed(x, sd, 32)
    np.testing.assert_allclose(pe.numpy(), gold["pred_eval"], atol=5e-5)
    stats = {}
    with torch.no_grad():
        pt = vae_ref.signal_ed(x, sd, 32, training=True, new_stats=stats)
    np.testing.assert_allclose(pt.numpy(), gold["pred_train"], atol=1e-4)
    assert abs(float(torch.nn.functional.mse_loss(pt, y)) - float(gold["loss_train"])) < 1e-5
