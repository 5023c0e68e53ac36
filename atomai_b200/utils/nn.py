"""
Seeding, weight (re)initialisation / averaging and model introspection helpers with the reference's
names and semantics (atomai/utils/nn.py:59-81, 120-146, 195-249).  Host-side only.
"""
import copy
import subprocess
from typing import Dict, Tuple, Type

import numpy as np
import torch
from torch.nn import (BatchNorm1d, BatchNorm2d, Conv1d, Conv2d, ConvTranspose1d,
                      ConvTranspose2d, Linear)


def set_train_rng(seed: int = 1) -> None:
    """Seeds numpy + torch (+CUDA) — atomai/utils/nn.py:136-146."""
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def average_weights(ensemble: Dict[int, Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    """Parameter-wise mean over an ensemble of state_dicts (stochastic weight averaging,
    atomai/utils/nn.py:59-81); BatchNorm statistics are taken from member 0."""
    out = copy.deepcopy(ensemble[0])
    skip = ("mean", "var", "tracked")
    for name in out:
        if name.split('_')[-1] in skip:
            continue
        stack = [m[name] for m in ensemble.values() if name in m]
        out[name].copy_(sum(stack) / float(len(stack)))
    return out


def sample_weights(ensemble: Dict[int, Dict[str, torch.Tensor]],
                   n_samples: int = 30) -> Dict[int, Dict[str, torch.Tensor]]:
    """theta_i ~ N(mu_i, sigma_i) per trainable parameter, with mean / std taken over the given
    state_dicts (SWAG-like sampling, atomai/utils/nn.py:84-117); BatchNorm statistics are copied
    from member 0."""
    first = ensemble[min(ensemble.keys())]
    out = {i: copy.deepcopy(first) for i in range(n_samples)}
    for name, ref in first.items():
        if name.split('_')[-1] in ("mean", "var", "tracked") or ref.dtype != torch.float32:
            continue
        w_all = torch.stack([m[name] for m in ensemble.values() if name in m], 0)
        ndist = torch.distributions.Normal(w_all.mean(0), w_all.std(0))
        for i in range(n_samples):
            out[i][name].copy_(ndist.sample())
    return out


def gpu_usage_map(cuda_device: int):
    """[memory.used, memory.total] in MiB (atomai/utils/nn.py:120-133 shells out to nvidia-smi;
    so does this, falling back to torch's own counters when nvidia-smi is unavailable)."""
    try:
        result = subprocess.check_output(
            ['nvidia-smi', '--id=' + str(cuda_device),
             '--query-gpu=memory.used,memory.total', '--format=csv,nounits,noheader'],
            encoding='utf-8', timeout=10)
        return [int(y) for y in result.split(',')][0:2]
    except Exception:  # noqa
        free, total = torch.cuda.mem_get_info(cuda_device)
        return [int((total - free) / 2**20), int(total / 2**20)]


def weights_init(module) -> None:
    """Xavier-uniform weights, zero bias (atomai/utils/nn.py:238-242)."""
    if isinstance(module, (Conv1d, Conv2d, ConvTranspose1d, ConvTranspose2d, Linear)):
        torch.nn.init.xavier_uniform_(module.weight.data)
        if module.bias is not None:
            torch.nn.init.zeros_(module.bias)


def reset_bnorm(module) -> None:
    if isinstance(module, (BatchNorm1d, BatchNorm2d)):
        module.reset_running_stats()
        module.reset_parameters()


def mock_forward(model: Type[torch.nn.Module], dims: Tuple[int] = (1, 64, 64)) -> torch.Tensor:
    """Passes a dummy variable through a network (atomai/utils/nn.py:195-208)."""
    x = torch.randn(1, *dims)
    p = next(model.parameters())
    x = x.to(p.device)
    was_training = model.training
    model.eval()
    with torch.no_grad():
        out = model(x)
    model.train(was_training)
    return out


def get_nb_classes(model: Type[torch.nn.Module]) -> int:
    """Number of output channels of a fully convolutional NN (atomai/utils/nn.py:211-218)."""
    return mock_forward(model).shape[1]


def get_downsample_factor(model: Type[torch.nn.Module]) -> int:
    """max/min feature-map width over the network (atomai/utils/nn.py:221-228).  The reference
    measures it with forward hooks on the top-level children; the native graph never calls the
    children's forward, so it is derived from the pooling structure recorded by a probe tape."""
    from .. import engine
    return engine.probe_downsample_factor(model)


def dummy_optimizer() -> Type[torch.optim.Optimizer]:
    return torch.optim.Optimizer([torch.zeros(1)], dict())
