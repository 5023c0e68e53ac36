"""Host-side logic of the predictors (batching policy, pre-processing, argument checks) with a
stand-in torch module on the CPU — atomai/predictors/predictor.py:82-106, 301-395.  The network
itself is covered by the GPU parity tests; here only what happens around it."""
import numpy as np
import pytest
import torch

from atomai_b200.predictors import BasePredictor, ImSpecPredictor


class _Im2Spec(torch.nn.Module):
    """(n, 1, h, w) -> (n, 1, L): deterministic stand-in for a trained SignalED."""
    def __init__(self, h, w, L):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.w = torch.nn.Parameter(torch.randn(h * w, L, generator=g))

    def forward(self, x):
        return (x.flatten(1) @ self.w)[:, None, :]


@pytest.mark.parametrize("n,num_batches", [(23, 10), (7, 10), (40, 3), (1, 10)])
def test_batch_predict_covers_every_sample(n, num_batches):
    """n // num_batches samples per batch, the remainder in one extra call, never an empty batch."""
    net = _Im2Spec(8, 6, 5)
    x = torch.randn(n, 1, 8, 6, generator=torch.Generator().manual_seed(1))
    calls = []
    p = BasePredictor(net, use_gpu=False)
    fwd = p.forward_
    p.forward_ = lambda t: (calls.append(len(t)), fwd(t))[1]
    out = p.batch_predict(x, (n, 1, 5), num_batches)
    assert sum(calls) == n and min(calls) >= 1
    np.testing.assert_allclose(out.numpy(), net(x).detach().numpy(), rtol=1e-6, atol=1e-6)


def test_imspec_predictor_shapes_and_norm():
    net = _Im2Spec(8, 6, 5)
    rs = np.random.RandomState(0)
    imgs = rs.rand(11, 8, 6).astype(np.float32) * 7 + 2
    p = ImSpecPredictor(net, (5,), use_gpu=False, verbose=False)
    out = p.run(imgs, num_batches=4)
    assert out.shape == (11, 5) and out.dtype == np.float32
    normed = (imgs - imgs.min()) / np.ptp(imgs)                 # norm=True is the default
    np.testing.assert_allclose(out, net(torch.from_numpy(normed)[:, None]).detach().numpy()[:, 0],
                               rtol=1e-5, atol=1e-5)
    single = p.predict(imgs[0], norm=False)                     # a 2-D image gets a batch axis
    assert single.shape == (1, 5)
    assert ImSpecPredictor(net, 5).output_dim == (5,)           # int -> tuple
    with pytest.raises(ValueError):
        ImSpecPredictor(net, (1, 2, 3))
    # spec2im direction: 1-D input gets a batch axis and a channel axis
    q = ImSpecPredictor(net, (8, 6), verbose=False)
    assert tuple(q.preprocess(np.arange(12, dtype=np.float32), norm=False).shape) == (1, 1, 12)
