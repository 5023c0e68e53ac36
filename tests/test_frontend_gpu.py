"""SURVEY.md 8(f) rows on the GPU: Locator front end (probabilities + thresholded masks in one
kernel), sub-image gather against the reference's goldens, sliding-window encoding, and
training-mode dropout."""
import numpy as np
import pytest
import torch

import golden_utils as gu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nb_classes", [3, 1])
def test_prob_mask_kernel_matches_softmax_and_cv_threshold(cuda, nb_classes):
    """prob == torch softmax / sigmoid to fp32 rounding; mask is EXACTLY cv2.threshold(prob, t, 1,
    THRESH_BINARY) of the returned probabilities (atomai/utils/img.py:554-564)."""
    import cv2
    from atomai_b200 import ops
    torch.manual_seed(0)
    lg = (torch.randn(2, 96, 80, nb_classes, device=cuda) * 3).contiguous()
    prob = torch.empty_like(lg)
    mask = torch.empty(lg.shape, device=cuda, dtype=torch.uint8)
    ops.prob_mask(lg, 0 if nb_classes > 1 else 1, 0.5, prob, mask)
    ref = torch.softmax(lg.cpu(), -1) if nb_classes > 1 else torch.sigmoid(lg.cpu())
    assert float((prob.cpu() - ref).abs().max()) <= 2e-7
    p = prob.cpu().numpy()
    for i in range(2):
        for c in range(nb_classes):
            _, thr = cv2.threshold(p[i, :, :, c], 0.5, 1, cv2.THRESH_BINARY)
            assert np.array_equal(mask[i, :, :, c].cpu().numpy().astype(np.float32), thr)


def test_locator_on_gpu_masks_equals_reference_golden(cuda):
    """The reference's Locator golden crop (test/predictors/test_locator.py): coordinates from
    masks thresholded on the GPU are bit-identical to the reference's."""
    from atomai_b200 import ops
    from atomai_b200.predictors import Locator
    g = gu.load("locator_crop.npz")
    nn_out = torch.from_numpy(g["nn_output"]).to(cuda).contiguous()
    prob = torch.empty_like(nn_out)
    mask = torch.empty(nn_out.shape, device=cuda, dtype=torch.uint8)
    ops.prob_mask(nn_out, 3, 0.5, prob, mask)            # identity + threshold
    coords = Locator(0.5, 5).run(prob.cpu().numpy(), masks=mask.cpu().numpy())
    assert np.array_equal(coords[0], g["coordinates"])


def test_gather_kernel_matches_reference_goldens(cuda):
    """extract_subimages on a CUDA stack against tests/golden/subimages.npz (unmodified reference,
    make_golden_img.py): same windows, same order, same dropped (non-fitting / NaN) windows."""
    from atomai_b200.utils import extract_subimages
    g = gu.load("subimages.npz")
    stack, coords, single, single_xy = gu.subimage_inputs()
    dstack = torch.from_numpy(stack).to(cuda)
    for r in (7, 8, 1, 12):
        for cls, tag in ((0, f"dict_r{r}"), (1, f"dict_c1_r{r}")):
            s, c, f = extract_subimages(dstack, coords, r, coord_class=cls)
            assert np.array_equal(s.cpu().numpy(), g[tag + "/sub"].astype(np.float32), equal_nan=True)
            assert np.array_equal(c, g[tag + "/com"]) and np.array_equal(f, g[tag + "/frames"])
    s, c, f = extract_subimages(dstack, coords, 64)
    assert s == [] and c == [] and f == []


def test_encode_image_sliding_window(cuda):
    """BaseVAE.encode_image_ (vae.py:300-344): every pixel whose window fits is encoded; equals
    encoding the explicitly cropped windows."""
    import atomai_b200 as ab
    from atomai_b200.models import VAE
    ab.set_math("tf32x3")
    rs = np.random.RandomState(0)
    img = rs.rand(40, 36)
    vae = VAE((8, 8), latent_dim=2, seed=1)
    cropped, enc = vae.encode_image_(img, num_batches=3)
    assert cropped.shape == (33, 29) and enc.shape == (33, 29, 2)
    # window centred at (r, c) covers img[r-4:r+4, c-4:c+4]; first encoded centre is (4, 4)
    np.testing.assert_array_equal(cropped, img[4:37, 4:33])
    for (r, c) in ((4, 4), (20, 17), (36, 32)):
        z, _ = vae.encode(img[r - 4:r + 4, c - 4:c + 4].astype(np.float32))
        np.testing.assert_allclose(enc[r - 4, c - 4], z[0], rtol=1e-4, atol=1e-5)


def test_dropout_training_mode(cuda):
    """ConvBlock with nn.Dropout in train mode (atomai/nets/blocks.py:68-69): the keep-rate and the
    1/(1-p) scaling are right, BatchNorm sees the dropped activations, and the backward applies the
    same mask (checked against torch autograd with the mask read back from the forward)."""
    import atomai_b200 as ab
    from atomai_b200.nets import ConvBlock
    ab.set_math("fp32")
    torch.manual_seed(0)
    blk = ConvBlock(2, 1, 8, 16, batch_norm=False, dropout_=0.5).to(cuda).train()
    x = torch.randn(4, 8, 32, 32, device=cuda, requires_grad=True)
    torch.manual_seed(123)
    y = blk(x)
    conv = blk.block[0]
    ref_pre = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x, conv.weight, conv.bias,
                                                                        padding=1), 0.01)
    keep = (y != 0)
    assert abs(float(keep.float().mean()) - 0.5) < 0.02
    assert torch.allclose(y[keep], 2.0 * ref_pre[keep], rtol=1e-4, atol=1e-5)
    g = torch.randn_like(y)
    y.backward(g)
    gx, gw = x.grad.clone(), conv.weight.grad.clone()
    x.grad = None
    conv.weight.grad = None
    (ref_pre * keep.float() * 2.0 * g).sum().backward()
    assert torch.allclose(gx, x.grad, rtol=1e-3, atol=1e-4)
    assert torch.allclose(gw, conv.weight.grad, rtol=1e-3, atol=1e-3)
    # eval mode: identity
    blk.eval()
    with torch.no_grad():
        assert torch.allclose(blk(x), ref_pre, rtol=1e-4, atol=1e-5)
    # with BatchNorm + dropout the whole Unet trains
    from atomai_b200.models import Segmentor
    X = gu.images(3, 8, 32, 32)
    yl = gu.labels(4, 8, 32, 32, 3)
    m = Segmentor("Unet", nb_classes=3, nb_filters=8, dropout=True)
    m.fit(X, yl, X[:4], yl[:4], training_cycles=4, batch_size=4, plot_training_history=False,
          filename="/tmp/drop_model")
    assert all(np.isfinite(m.loss_acc["train_loss"]))


def test_ensemble_trainer_and_predictor(cuda, tmp_path):
    """EnsembleTrainer (from_scratch / from_baseline / swag) and EnsemblePredictor with the
    reference's call patterns (test/trainers/test_etrainer.py, test/predictors/test_epredictor.py):
    members differ, mean/var shapes, mean/var equal numpy over the member outputs."""
    import atomai_b200 as ab
    from atomai_b200.predictors import EnsemblePredictor
    from atomai_b200.trainers import EnsembleTrainer
    ab.set_math("tf32x3")
    X = gu.images(1, 16, 32, 32)[:, None]
    y = gu.labels(2, 16, 32, 32, 3)
    Xt, yt = X[:8], y[:8]
    et = EnsembleTrainer("Unet", nb_classes=3, nb_filters=8)
    et.compile_ensemble_trainer(training_cycles=3, batch_size=4, filename=str(tmp_path / "ens"),
                                plot_training_history=False)
    net, ens = et.train_ensemble_from_scratch(X, y, Xt, yt, n_models=3)
    assert sorted(ens) == [0, 1, 2]
    w0, w1 = ens[0]["c1.block.0.weight"], ens[1]["c1.block.0.weight"]
    assert not torch.equal(w0, w1)
    ck = torch.load(str(tmp_path / "ens_ensemble_metadict.tar"), weights_only=False)
    assert sorted(ck["weights"]) == [0, 1, 2] and ck["model_type"] == "seg"
    net2, ens2 = et.train_ensemble_from_baseline(X, y, Xt, yt, n_models=2, training_cycles_base=3,
                                                 training_cycles_ensemble=2)
    assert sorted(ens2)[:2] == [0, 1]
    p = EnsemblePredictor(net, ens, nb_classes=3, verbose=0)
    mean, var = p.predict(X[:5, 0], num_batches=2)
    assert mean.shape == (5, 32, 32, 3) and var.shape == mean.shape
    allp = p.ensemble_forward(p.preprocess(X[:5, 0]))
    np.testing.assert_allclose(mean, allp.mean(0).transpose(0, 2, 3, 1), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(var, allp.var(0).transpose(0, 2, 3, 1), rtol=1e-4, atol=1e-9)
    assert np.all(var >= 0) and float(var.max()) > 0


def test_gpu_augmentation(cuda, tmp_path):
    """GPU datatransform against the reference's semantics
    (atomai/transforms/imaug.py:109-358): flips act on image and labels alike, blur equals
    scipy.ndimage.gaussian_filter, gamma equals skimage's adjust_gamma formula, noise levels
    follow skimage.util.random_noise, output is min-max normalised; Segmentor.fit accepts the
    augmentation kwargs."""
    from scipy import ndimage
    from atomai_b200.transforms import datatransform
    X = torch.from_numpy(gu.images(7, 6, 48, 48)).to(cuda)
    y = torch.from_numpy(gu.labels(8, 6, 48, 48, 3)).to(cuda)
    Xn = ((X - X.min()) / (X.max() - X.min())).cpu().numpy()
    # rotation only: every output is one of the five flips / turns of its input, labels follow
    out, lab = datatransform(3, seed=3, rotation=True).run(X, y)
    for i in range(6):
        cands = {"v": lambda a: a[::-1], "h": lambda a: a[:, ::-1], "hv": lambda a: a[::-1, ::-1],
                 "ccw": lambda a: np.rot90(a, 1), "id": lambda a: a}
        hit = [k for k, f in cands.items()
               if np.allclose(out[i].cpu().numpy(), f(Xn[i]), atol=1e-6)
               and np.array_equal(lab[i].cpu().numpy(), f(y[i].cpu().numpy()))]
        assert hit, i
    # blur only (same parameter draw as the reference: np.random.seed(seed); randint per image)
    out, _ = datatransform(3, seed=5, blur=[20, 21]).run(X, None)
    ref = np.stack([ndimage.gaussian_filter(im.astype(np.float64), 20 * 5e-2) for im in Xn])
    ref = (ref - ref.min()) / np.ptp(ref)
    assert np.abs(out.cpu().numpy() - ref).max() < 2e-5
    # gamma only
    out, _ = datatransform(3, seed=5, contrast=[15, 16]).run(X, None)
    ref = Xn.astype(np.float64) ** 1.5
    ref = (ref - ref.min()) / np.ptp(ref)
    assert np.abs(out.cpu().numpy() - ref).max() < 2e-5
    # gaussian noise: variance 1e-4 * 40, clipped to [0, 1] before the final normalisation
    out, _ = datatransform(3, seed=5, gauss_noise=[40, 41]).run(X, None)
    d = out.cpu().numpy() - Xn
    inner = (Xn > 0.2) & (Xn < 0.8)
    assert abs(d[inner].std() / np.sqrt(40e-4) - 1) < 0.15 and abs(d[inner].mean()) < 0.02
    # salt & pepper: amount = 30e-3 of the pixels become 0 or 1
    out, _ = datatransform(3, seed=5, salt_and_pepper=[30, 31]).run(X, None)
    o = out.cpu().numpy()
    changed = np.abs(o - Xn) > 1e-6
    assert abs(changed.mean() - 0.03) < 0.008 and set(np.unique(o[changed])) <= {0.0, 1.0}
    # poisson + jitter + background run and stay finite / normalised
    out, lab = datatransform(3, seed=9, rotation=True, poisson_noise=True, jitter=[10, 20],
                             background=True, gauss_noise=True).run(X, y)
    o = out.cpu().numpy()
    assert np.isfinite(o).all() and abs(o.min()) < 1e-6 and abs(o.max() - 1) < 1e-6
    assert sorted(np.unique(lab.cpu().numpy())) == [0, 1, 2]
    with pytest.raises(NotImplementedError):
        datatransform(3, zoom=True)
    from atomai_b200.models import Segmentor
    Xs, ys = gu.images(3, 8, 32, 32), gu.labels(4, 8, 32, 32, 3)
    m = Segmentor("Unet", nb_classes=3, nb_filters=8)
    m.fit(Xs, ys, Xs[:4], ys[:4], training_cycles=3, batch_size=4, plot_training_history=False,
          filename=str(tmp_path / "aug"), rotation=True, gauss_noise=True, contrast=True)
    assert all(np.isfinite(m.loss_acc["train_loss"])) and m.augment_fn is not None
