"""Module tree / state_dict contract (the reference's tests address parameters by name and inspect
child module types: test/trainers/test_trainer.py:232-352) and host-side plumbing.  CPU only."""
import warnings

import numpy as np
import pytest
import torch

import golden_utils as gu
from atomai_b200.nets import ConvBlock, DilatedBlock, Unet, UpsampleBlock, dilnet, init_fcnn_model
from atomai_b200.trainers import SegTrainer
from atomai_b200.utils import preproc


def test_default_unet_layout_matches_reference_checkpoint():
    net = Unet(nb_classes=3)
    sd = net.state_dict()
    w = gu.load("bfo_weights.npz")            # pretrained/bfo.tar of the reference
    assert list(sd.keys()) == list(w.files) or set(sd.keys()) == set(w.files)
    for k in w.files:
        assert tuple(sd[k].shape) == w[k].shape, k
    assert sum(p.numel() for p in net.parameters()) == 594067
    assert len(sd) == 99
    assert [n for n, _ in net.named_children()] == [
        "c1", "c2", "c3", "bn", "upsample_block1", "c4", "upsample_block2", "c5",
        "upsample_block3", "c6", "px"]
    assert isinstance(net.upsample_block1, UpsampleBlock) and isinstance(net.c4, ConvBlock)


@pytest.mark.parametrize("bn", [True, False])
def test_batchnorm_and_layer_counts(bn):
    t = SegTrainer("Unet", nb_classes=3, batch_norm=bn, layers=[2, 3, 3, 4])
    keys = list(t.net.state_dict().keys())
    assert any("running_mean" in k for k in keys) == bn
    nconv = sum(isinstance(m, torch.nn.Conv2d) for m in t.net.modules())
    assert nconv == 2 + 3 + 3 + 4 + 3 + 3 + 2 + 3 + 1
    assert t.meta_state_dict["layers"] == [2, 3, 3, 4]


def test_dilated_unet_and_dilnet_structure():
    net, meta = init_fcnn_model("Unet", 1, with_dilation=True, nb_filters=32, layers=[2, 3, 3, 3])
    assert isinstance(net.bn, DilatedBlock) and meta["with_dilation"]
    dil = [m.dilation[0] for m in net.bn.atrous_module if isinstance(m, torch.nn.Conv2d)]
    assert dil == [2, 4, 6]
    net2, meta2 = init_fcnn_model("dilnet", 3)
    assert isinstance(net2, dilnet) and meta2["nb_filters"] == 25
    with pytest.raises(NotImplementedError):
        init_fcnn_model("SegNet", 3)


def test_resnet_structure_and_state_dict_keys():
    """ResBlock networks: module tree / checkpoint keys of atomai/nets/fcnn.py:229-376."""
    net, meta = init_fcnn_model("SegResNet", 3)
    assert meta["nb_filters"] == 32 and meta["layers"] == [2, 2, 2] and meta["dropout"] is None
    keys = set(net.state_dict())
    for k in ("c1.block.0.weight", "c2.res_module.0.c0.weight", "c2.res_module.1.bn2.running_var",
              "bn.res_module.0.c1.bias", "upsample_block1.conv.weight", "c3.res_module.0.c0.weight",
              "c4.block.2.weight", "px.bias"):
        assert k in keys, k
    assert net.c3.res_module[0].c0.weight.shape == (64, 128, 1, 1)
    net2, meta2 = init_fcnn_model("ResHedNet", 1)
    assert meta2["nb_filters"] == 64 and meta2["layers"] == [3, 4, 5]
    keys2 = set(net2.state_dict())
    for k in ("net1.res_module.2.c2.weight", "net2.1.res_module.3.bn1.weight",
              "net3.1.res_module.4.c0.bias", "net1score.0.weight", "net3score.1.running_mean",
              "out.weight"):
        assert k in keys2, k
    assert net2.out.weight.shape == (1, 3, 1, 1)


def test_constructor_argument_errors_match_reference():
    with pytest.raises(AssertionError):
        ConvBlock(3, 1, 1, 8)
    with pytest.raises(NotImplementedError):
        UpsampleBlock(2, 8, 4, mode="bicubic")


def test_loss_selection_reprs():
    from atomai_b200.losses_metrics import select_loss
    assert str(select_loss("ce", 3)) == "CrossEntropyLoss()"
    assert str(select_loss("ce", 1)) == "BCEWithLogitsLoss()"
    assert str(select_loss("mse")) == "MSELoss()"
    with pytest.raises(ValueError):
        select_loss("ce")


def test_preproc_dims_types_and_chunking():
    rs = np.random.RandomState(0)
    X = rs.rand(20, 8, 8).astype(np.float64)
    y = rs.randint(0, 3, (20, 8, 8))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Xl, yl, Xtl, ytl, nc = preproc.preprocess_training_image_data(X, y, X[:8], y[:8], 4)
    assert nc == 3 and len(Xl) == 5 and len(Xtl) == 2
    assert Xl[0].shape == (4, 1, 8, 8) and Xl[0].dtype == torch.float32
    assert yl[0].shape == (4, 8, 8) and yl[0].dtype == torch.int64
    yb = (y > 0).astype(np.int64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        *_, nc1 = preproc.preprocess_training_image_data(X, yb, X[:8], yb[:8], 4)
    assert nc1 == 1
    with pytest.raises(AssertionError):
        preproc.num_classes_from_labels(np.array([1, 2]))
    assert preproc.get_array_memsize(np.zeros(10, np.float64)) == 40
    assert preproc.get_array_memsize(np.zeros(10, np.float32), "double") == 80
    sh = preproc.shard_batches([torch.arange(8).reshape(8, 1)], 1, 4)
    assert sh[0].flatten().tolist() == [2, 3]


def test_torch_format_image_matches_reference_normalisation():
    img = np.random.RandomState(1).rand(2, 16, 16) * 7 - 3
    t = preproc.torch_format_image(img)
    ref = torch.from_numpy(((img - img.min()) / np.ptp(img))[:, None]).float()
    assert torch.equal(t, ref)


def test_trainer_cpu_is_structural_only():
    t = SegTrainer("Unet", nb_classes=3)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            t.net(torch.zeros(1, 1, 16, 16))
