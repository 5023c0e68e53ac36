"""Parity of the native VAE / rVAE / ImSpec paths against goldens produced by the unmodified
reference on CPU (tests/golden/make_golden_vae.py): latent means / log-sds, reconstructions, ELBO
and parameter gradients with injected reparameterisation noise.
Tolerances: fp32 math 1e-4 rel (outputs), 2e-3 (gradients, rel L2); tf32 math 2e-3 / 5e-2."""
import numpy as np
import pytest
import torch

import golden_utils as gu
from test_oracle_vae import HW, VAE_CASES, build_imspec, build_vae

pytestmark = pytest.mark.gpu


def rel(a, ref):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(a - ref).max() / (np.abs(ref).max() + 1e-30))


def grad_rel(nets, gold):
    te = tr = 0.0
    for pref, net in nets:
        for k, p in net.named_parameters():
            assert p.grad is not None, f"no gradient for {pref}.{k}"
            g = p.grad.detach().cpu().numpy()
            ref = gold[f"grad/{pref}.{k}"]
            got = gu.sample_flat(g, 97) if g.size > 4096 else g
            te += float(((got.reshape(-1) - ref.reshape(-1)).astype(np.float64) ** 2).sum())
            tr += float((ref.reshape(-1).astype(np.float64) ** 2).sum())
    return (te / tr) ** 0.5


@pytest.mark.parametrize("math", ["fp32", "tf32"])
@pytest.mark.parametrize("name", list(VAE_CASES))
def test_vae_forward_backward(cuda, name, math):
    import atomai_b200 as ab
    ab.set_math(math)
    m, enc, dec, x, eps, gold = build_vae(name)
    assert m.device == "cuda"
    x, eps = x.to(cuda), eps.to(cuda)
    tol = 1e-4 if math == "fp32" else 2e-3
    m.encoder_net.train(); m.decoder_net.train()
    m.dx_prior = 0.1
    elbo = m.forward_compute_elbo(x, eps=eps)
    (-elbo).backward()
    assert abs(elbo.item() - float(gold["elbo"])) <= tol * abs(float(gold["elbo"]))
    with torch.no_grad():
        zm, zl = m.encoder_net(x)
    assert rel(zm.cpu().numpy(), gold["z_mean"]) <= tol
    assert rel(zl.cpu().numpy(), gold["z_logsd"]) <= tol
    with torch.no_grad():
        z = zm + torch.exp(zl) * eps
        if VAE_CASES[name]["kind"] == "rvae":
            xr = m.decoder_net.decode(z[:, 3:], z[:, 0], z[:, 1:3] * 0.1)
        else:
            xr = m.decoder_net(z)
    assert xr.shape == (x.shape[0], *HW)
    assert rel(xr.cpu().numpy(), gold["x_reconstr"]) <= tol * 3
    g = grad_rel([("encoder", m.encoder_net), ("decoder", m.decoder_net)], gold)
    assert g <= (2e-3 if math == "fp32" else 5e-2), g


def test_rdecoder_reference_signature(cuda):
    """decoder_net(x_coord, z) with a materialised, transformed grid (rvae.py:140-145)."""
    from atomai_b200.utils.coords import imcoordgrid, transform_coordinates
    import atomai_b200 as ab
    ab.set_math("fp32")
    m, enc, dec, x, eps, gold = build_vae("rvae_conv_32")
    z = torch.from_numpy(gold["z_mean"]).to(cuda) + torch.exp(torch.from_numpy(gold["z_logsd"]).to(cuda)) * eps.to(cuda)
    grid = imcoordgrid(HW).to(cuda).expand(z.shape[0], -1, -1)
    xc = transform_coordinates(grid, z[:, 0], (z[:, 1:3] * 0.1).unsqueeze(1))
    with torch.no_grad():
        xr = m.decoder_net(xc, z[:, 3:])
    assert rel(xr.cpu().numpy(), gold["x_reconstr"]) <= 5e-4


@pytest.mark.parametrize("math", ["fp32", "tf32"])
def test_imspec_forward_backward(cuda, math):
    import atomai_b200 as ab
    from atomai_b200.losses_metrics import select_loss
    ab.set_math(math)
    net, sd, x, y, gold = build_imspec()
    net = net.to(cuda)
    x, y = x.to(cuda), y.to(cuda)
    tol = 1e-4 if math == "fp32" else 5e-3
    net.eval()
    with torch.no_grad():
        pe = net(x)
    assert pe.shape == (6, 1, 32)
    assert rel(pe.cpu().numpy(), gold["pred_eval"]) <= tol
    net.train(); net.zero_grad()
    pt = net(x)
    assert rel(pt.detach().cpu().numpy(), gold["pred_train"]) <= tol * 3
    loss = select_loss("mse")(pt, y)
    assert abs(loss.item() - float(gold["loss_train"])) <= tol * 3 * float(gold["loss_train"])
    loss.backward()
    g = grad_rel([("net", net)], gold)
    assert g <= (5e-3 if math == "fp32" else 2e-1), g


def test_rvae_fit_runs_and_improves(cuda, tmp_path):
    import atomai_b200 as ab
    from atomai_b200.models import rVAE
    ab.set_math("tf32")
    X = gu.images(5, 64, 32, 32)
    m = rVAE((32, 32), latent_dim=2, conv_encoder=True, numhidden_encoder=32, numhidden_decoder=32)
    m.fit(X, training_cycles=3, batch_size=16, filename=str(tmp_path / "rvae"))
    hist = m.loss_history["train_loss"]
    assert len(hist) == 3 and hist[-1] > hist[0]          # ELBO increases
    zm, zs = m.encode(X[:10])
    assert zm.shape == (10, 5) and zs.shape == (10, 5)
    assert m.decode(zm[:, 3:]).shape == (10, 32, 32)
    ck = torch.load(str(tmp_path / "rvae.tar"), weights_only=False)
    assert "encoder" in ck and "decoder" in ck and ck["coord"] == 3
