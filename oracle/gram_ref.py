"""
TEST INFRASTRUCTURE — float64 numpy restatement of the deep-kernel covariance:
gpytorch 1.9+ (setup.py:40 `gpytorch>=1.9.1`, not vendored, not installed here) defines
  RBFKernel:      k = exp(-0.5 * sum_d ((x_d - x'_d) / l_d)^2)
  MaternKernel:   nu=2.5: k = (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r), r = scaled distance
  ScaleKernel:    outputscale * k
as configured at atomai/nets/gp.py:41-46, 100-111.  PARITY UNPINNED: no gpytorch in this image and
the reference's own GP tests assert only shapes/types.
"""
import numpy as np


def gram(x1, x2, lengthscale, outputscale=1.0, kind="rbf"):
    a = np.asarray(x1, np.float64) / np.asarray(lengthscale, np.float64)
    b = np.asarray(x2, np.float64) / np.asarray(lengthscale, np.float64)
    d2 = np.maximum((a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2 * a @ b.T, 0.0)
    if kind == "rbf":
        return outputscale * np.exp(-0.5 * d2)
    r = np.sqrt(d2)
    return outputscale * (1 + np.sqrt(5) * r + 5.0 / 3.0 * d2) * np.exp(-np.sqrt(5) * r)


def mlp(x, sd, names):
    """fcFeatureExtractor forward (atomai/nets/gp.py:14-26) in float64."""
    h = np.asarray(x, np.float64)
    for i, n in enumerate(names):
        h = h @ sd[n + ".weight"].astype(np.float64).T + sd[n + ".bias"].astype(np.float64)
        if i + 1 < len(names):
            h = np.maximum(h, 0)
    return h
