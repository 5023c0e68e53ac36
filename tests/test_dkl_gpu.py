"""DKL row: native fcFeatureExtractor (forward + backward) and the dense deep-kernel Gram against
the float64 numpy oracle (oracle/gram_ref.py; gpytorch parity is UNPINNED, see its header)."""
import numpy as np
import pytest
import torch

from oracle import gram_ref

pytestmark = pytest.mark.gpu


def test_feature_extractor_forward_backward(cuda):
    from atomai_b200.nets import fcFeatureExtractor
    torch.manual_seed(0)
    fe = fcFeatureExtractor(37, 2).to(cuda)
    assert [n for n, _ in fe.named_children()] == ["linear1", "relu1", "linear2", "relu2",
                                                   "linear3", "relu3", "linear4"]
    x = torch.randn(300, 37, device=cuda)
    y = fe(x)
    sd = {k: v.detach().cpu().numpy() for k, v in fe.state_dict().items()}
    ref = gram_ref.mlp(x.cpu().numpy(), sd, ["linear1", "linear2", "linear3", "linear4"])
    assert np.abs(y.detach().cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max()
    # backward vs torch autograd on the same weights (fp64 on CPU)
    g = torch.randn_like(y)
    y.backward(g)
    ref_mod = torch.nn.Sequential(*[m for m in fcFeatureExtractor(37, 2).children()]).double()
    ref_mod.load_state_dict({k: v.detach().cpu().double() for k, v in
                             zip(ref_mod.state_dict().keys(), fe.state_dict().values())})
    yr = ref_mod(x.cpu().double())
    yr.backward(g.cpu().double())
    for (k, p), pr in zip(fe.named_parameters(), ref_mod.parameters()):
        e = (p.grad.cpu().double() - pr.grad).norm() / (pr.grad.norm() + 1e-30)
        assert e <= 1e-4, (k, float(e))


@pytest.mark.parametrize("kind", ["rbf", "matern"])
def test_dense_gram_matches_oracle(cuda, kind):
    from atomai_b200.nets import dense_gram
    rs = np.random.RandomState(0)
    x1 = rs.randn(333, 128).astype(np.float32)
    x2 = rs.randn(257, 128).astype(np.float32)
    ls = (rs.rand(128) * 4 + 8).astype(np.float32)
    K = dense_gram(torch.from_numpy(x1).to(cuda), torch.from_numpy(x2).to(cuda),
                   torch.from_numpy(ls), 1.7, kind).cpu().numpy()
    ref = gram_ref.gram(x1, x2, ls, 1.7, kind)
    assert K.shape == (333, 257)
    assert np.abs(K - ref).max() <= 2e-5 * np.abs(ref).max()


def test_gram_properties_at_scale(cuda):
    """Size-independent properties on a larger problem: symmetry, unit diagonal, bounds."""
    from atomai_b200.nets import dense_gram
    x = torch.randn(4096, 128, device=cuda)
    ls = torch.full((128,), 128 ** 0.5)
    K = dense_gram(x, x, ls, 1.0, "rbf")
    assert torch.allclose(K, K.t(), atol=1e-6)
    assert torch.allclose(torch.diagonal(K), torch.ones(4096, device=cuda), atol=1e-5)
    assert float(K.min()) >= 0.0 and float(K.max()) <= 1.0 + 1e-6
