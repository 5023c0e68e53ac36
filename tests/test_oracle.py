"""The oracle (oracle/nets_ref.py) against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only; this is what pins the oracle (prompt §③)."""
import sys
from collections import OrderedDict

import numpy as np
import pytest
import torch

import golden_utils as gu
from oracle import nets_ref

CASES = {
    "unet_default_3c": dict(model="Unet", nb_classes=3, seed=100, n=2, h=32, w=48, cfg={}),
    "unet_nearest_1c": dict(model="Unet", nb_classes=1, seed=200, n=2, h=32, w=32,
                            cfg=dict(upsampling="nearest", nb_filters=8)),
    "unet_dilated_3c": dict(model="Unet", nb_classes=3, seed=300, n=2, h=64, w=64,
                            cfg=dict(with_dilation=True)),
    "unet_nobn_3c": dict(model="Unet", nb_classes=3, seed=400, n=2, h=32, w=32,
                         cfg=dict(batch_norm=False, layers=[2, 2, 2, 2])),
    "dilnet_default_3c": dict(model="dilnet", nb_classes=3, seed=500, n=2, h=32, w=32, cfg={}),
    "unet_default_3c_128": dict(model="Unet", nb_classes=3, seed=600, n=4, h=128, w=128, cfg={}),
    # BASELINE.json configs[1] geometry: default 3-class Unet on 512x512 images (N = 2)
    "unet_default_3c_512": dict(model="Unet", nb_classes=3, seed=700, n=2, h=512, w=512, cfg={}),
    # ResBlock networks (atomai/nets/fcnn.py:229-376)
    "segresnet_default_3c": dict(model="SegResNet", nb_classes=3, seed=800, n=2, h=64, w=64, cfg={}),
    "segresnet_nobn_1c": dict(model="SegResNet", nb_classes=1, seed=810, n=2, h=32, w=32,
                              cfg=dict(batch_norm=False, nb_filters=16, upsampling="nearest")),
    "reshednet_3c": dict(model="ResHedNet", nb_classes=3, seed=820, n=2, h=64, w=64,
                         cfg=dict(nb_filters=16, layers=[2, 2, 2])),
}


def logits_view(arr, gold):
    """Full logits, or the strided sample the larger goldens store."""
    stride = int(gold["logit_stride"]) if "logit_stride" in gold.files else 1
    arr = np.asarray(arr)
    return arr if stride == 1 else gu.sample_flat(arr, stride)


def build_case(name):
    """(net [atomai_b200 module, CPU], state_dict tensors, cfg, x, y, golden)."""
    from atomai_b200.nets import init_fcnn_model
    c = CASES[name]
    net, meta = init_fcnn_model(c["model"], c["nb_classes"], **c["cfg"])
    shapes = OrderedDict((k, tuple(v.shape)) for k, v in net.state_dict().items())
    vals = gu.fill_state_dict(shapes, c["seed"])
    sd = OrderedDict((k, torch.from_numpy(v)) for k, v in vals.items())
    net.load_state_dict(sd)
    x = torch.from_numpy(gu.images(c["seed"] + 1, c["n"], c["h"], c["w"]))[:, None]
    if c["nb_classes"] > 2:
        y = torch.from_numpy(gu.labels(c["seed"] + 2, c["n"], c["h"], c["w"], c["nb_classes"]))
    else:
        y = torch.from_numpy((gu.labels(c["seed"] + 2, c["n"], c["h"], c["w"], 2) > 0)
                             .astype(np.float32))[:, None]
    cfg = dict(meta)
    return net, sd, cfg, x, y, gu.load(name + ".npz")


def oracle_forward(name, sd, cfg, x, training, new_stats=None):
    fn = {"Unet": nets_ref.unet_forward, "dilnet": nets_ref.dilnet_forward,
          "SegResNet": nets_ref.segresnet_forward, "ResHedNet": nets_ref.reshednet_forward}[
              CASES[name]["model"]]
    return fn(x, sd, cfg, training=training, new_stats=new_stats)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_goldens(name):
    _, sd, cfg, x, y, gold = build_case(name)
    with torch.no_grad():
        le = oracle_forward(name, sd, cfg, x, False)
    np.testing.assert_allclose(logits_view(le.numpy(), gold), gold["logits_eval"], rtol=0, atol=2e-5)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.dtype.is_floating_point and "running_" not in k}
    work = dict(sd)
    work.update(leaves)
    stats = {}
    lt = oracle_forward(name, work, cfg, x, True, stats)
    np.testing.assert_allclose(logits_view(lt.detach().numpy(), gold), gold["logits_train"], rtol=0,
                               atol=2e-5)
    loss = nets_ref.seg_loss(lt, y, cfg["nb_classes"])
    assert abs(float(loss) - float(gold["loss_train"])) < 1e-5
    loss.backward()
    for k, p in leaves.items():
        g = p.grad.numpy()
        ref = gold["grad/" + k]
        got = gu.sample_flat(g, 97) if g.size > 4096 else g
        scale = max(float(gold["gradnorm/" + k]) / np.sqrt(g.size), 1e-8)
        assert np.abs(got.reshape(-1) - ref.reshape(-1)).max() < 2e-3 * scale + 1e-7, k
    for k, v in stats.items():
        np.testing.assert_allclose(v.numpy(), gold["buf/" + k], rtol=1e-5, atol=1e-6)


def test_oracle_adam_steps_match_reference():
    name = "unet_default_3c"
    _, sd, cfg, x, y, gold = build_case(name)
    sd = OrderedDict((k, v.clone()) for k, v in sd.items())
    state = {}
    losses = [nets_ref.unet_train_step(x, y, sd, cfg, state)[0] for _ in range(3)]
    np.testing.assert_allclose(losses, gold["adam_losses"], rtol=2e-4)
    np.testing.assert_allclose(sd["px.weight"].numpy(), gold["adam_px_weight"], atol=2e-5)
    np.testing.assert_allclose(sd["c1.block.0.weight"].numpy(), gold["adam_c1_weight"], atol=2e-5)


def test_oracle_pretrained_bfo_crop():
    w = gu.load("bfo_weights.npz")
    sd = {k: torch.from_numpy(w[k]) for k in w.files}
    g = gu.load("bfo_crop.npz")
    cfg = dict(nb_classes=3)
    with torch.no_grad():
        logits = nets_ref.unet_forward(torch.from_numpy(g["image"])[None, None], sd, cfg)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=0, atol=5e-5)
