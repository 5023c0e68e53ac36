"""
dklGPR — deep kernel learning (DKL)-based Gaussian process regression, drop-in for
atomai.models.dklGPR (atomai/models/dklgp/dklgpr.py:23-241): fit, fit_ensemble, predict, embed,
sample_from_posterior, thompson with the reference's argument conventions and output shapes.
"""
import warnings
from typing import List, Tuple, Type, Union

import numpy as np
import torch

from ...trainers.gptrainer import dklGPTrainer


def _batches(x: torch.Tensor, **kwargs):
    bs = kwargs.get("batch_size", len(x))
    for i in range(0, len(x), bs):
        yield x[i:i + bs]


class dklGPR(dklGPTrainer):
    """
    Deep kernel learning (DKL)-based Gaussian process regression (GPR)

    Args:
        indim: input feature dimension
        embedim: embedding dimension (determines dimensionality of kernel space)
        shared_embedding_space: use one embedding space for all target outputs
        **device, **precision, **seed

    Examples:

        >>> dklgp = dklGPR(X.shape[-1], embedim=2, precision="single")
        >>> dklgp.fit(X, y, training_cycles=100, lr=1e-2)
        >>> mean, var = dklgp.predict(X_test, batch_size=len(X_test))
        >>> samples = dklgp.sample_from_posterior(X_test, num_samples=1000)
    """
    def __init__(self, indim: int, embedim: int = 2, shared_embedding_space: bool = True,
                 **kwargs: Union[str, int]) -> None:
        args = (indim, embedim, shared_embedding_space)
        super(dklGPR, self).__init__(*args, **kwargs)

    def fit(self, X, y, training_cycles: int = 1, **kwargs) -> None:
        """Initializes and trains a deep kernel GP model (kwargs feature_extractor,
        freeze_weights, lr, print_loss) — dklgpr.py:70-93."""
        _ = self.run(X, y, training_cycles, **kwargs)

    def fit_ensemble(self, X, y, training_cycles: int = 1, n_models: int = 5, **kwargs) -> None:
        """Initializes and trains an ensemble of deep kernel GP models on the same scalar
        target, each with its own feature extractor — dklgpr.py:95-132."""
        if y.ndim == 1:
            y = y[None]
        if y.shape[0] > 1:
            raise NotImplementedError(
                "The ensemble training is currently supported only for scalar targets")
        y = y.repeat(n_models, 0) if isinstance(y, np.ndarray) else y.repeat(n_models, 1)
        if self.correlated_output:
            msg = ("Replacing a single shared embedding space with" +
                   " {} independent ones").format(n_models)
            warnings.warn(msg)
            self.correlated_output = False
        self.ensemble = True
        _ = self.run(X, y, training_cycles, **kwargs)

    def _models(self) -> List[torch.nn.Module]:
        return list(self.gp_model.models) if not self.correlated_output else [self.gp_model]

    def _compute_posterior(self, X: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Posterior mean / variance of the latent function at X ((n, d)): (q, n) each, q = number
        of outputs (shared embedding) or of independent models — dklgpr.py:134-156."""
        self.gp_model.eval()
        X = X.to(self.device)
        outs = [m.posterior(X) for m in self._models()]
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])

    def sample_from_posterior(self, X, num_samples: int = 1000) -> np.ndarray:
        """Samples from the (marginal) posterior at X: (num_samples, q, n) — dklgpr.py:158-172.
        Points are sampled independently from their marginals (the reference draws from the joint
        posterior through gpytorch's lazy covariance)."""
        X, _ = self.set_data(X)
        mean, var = self._compute_posterior(X)
        eps = torch.randn((num_samples,) + tuple(mean.shape), device=mean.device, dtype=mean.dtype)
        return (mean[None] + var.sqrt()[None] * eps).cpu().numpy()

    def thompson(self, X_cand, scalarize_func=None, maximize: bool = True):
        """Thompson sampling for selecting the next measurement point — dklgpr.py:174-192."""
        X_cand, _ = self.set_data(X_cand)
        mean, var = self._compute_posterior(X_cand)
        tsample = mean + var.sqrt() * torch.randn_like(mean)
        if tsample.ndim > 1 and scalarize_func is not None:
            tsample = scalarize_func(tsample).unsqueeze(0)
        idx = tsample.argmax(1) if maximize else tsample.argmin(1)
        return tsample.cpu().numpy(), idx.cpu().numpy()

    def predict(self, x_new, **kwargs) -> Tuple[np.ndarray]:
        """Prediction of mean and variance using the trained model (**batch_size) —
        dklgpr.py:202-217."""
        x_new, _ = self.set_data(x_new, device='cpu')
        means, vars_ = [], []
        for x in _batches(x_new, **kwargs):
            m, v = self._compute_posterior(x)
            means.append(m.cpu())
            vars_.append(v.cpu())
        return (torch.cat(means, 1).numpy().squeeze(), torch.cat(vars_, 1).numpy().squeeze())

    def _embed(self, x_new: torch.Tensor):
        self.gp_model.eval()
        with torch.no_grad():
            if self.correlated_output:
                embeded = self.gp_model.embed(x_new)
            else:
                embeded = torch.cat([m.embed(x_new)[..., None] for m in self.gp_model.models], -1)
        return embeded.cpu()

    def embed(self, x_new, **kwargs: int) -> np.ndarray:
        """Embeds the input data to a "latent" space using the trained feature extractor NN —
        dklgpr.py:231-241."""
        x_new, _ = self.set_data(x_new, device='cpu')
        embeded = torch.cat([self._embed(x.to(self.device)) for x in _batches(x_new, **kwargs)], 0)
        if not self.correlated_output and not self.ensemble:
            embeded = embeded.permute(-1, 0, 1)
        return embeded.numpy()
