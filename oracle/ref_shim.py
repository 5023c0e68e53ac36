"""
TEST INFRASTRUCTURE — import shim that makes the UNMODIFIED reference (/root/reference, AtomAI
v0.8.1) importable in this container, where matplotlib / skimage / mendeleev / gpytorch are not
installed.  Used only by tests/golden/make_golden.py to generate the
committed golden vectors; nothing under atomai_b200/ imports it and it never runs on the GPU box
(/root/reference does not exist there).
"""
import sys
import types
from unittest.mock import MagicMock

import torch

REFERENCE_ROOT = "/root/reference"


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return MagicMock(name=f"{self.__name__}.{name}")


def _stub(name):
    m = _Stub(name)
    m.__path__ = []
    sys.modules[name] = m
    return m


def import_reference():
    """Returns the reference `atomai` package (CPU)."""
    if "atomai" in sys.modules and getattr(sys.modules["atomai"], "__file__", "").startswith(REFERENCE_ROOT):
        return sys.modules["atomai"]
    sys.path.insert(0, REFERENCE_ROOT)
    for n in ["matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.cm",
              "matplotlib.patches", "mpl_toolkits", "mpl_toolkits.axes_grid1", "skimage",
              "skimage.exposure", "skimage.util", "mendeleev", "progressbar"]:
        if n not in sys.modules:
            _stub(n)
    plt = sys.modules["matplotlib.pyplot"]
    plt.subplots = lambda *a, **k: (MagicMock(), MagicMock())
    sys.modules["matplotlib"].pyplot = plt
    if "gpytorch" not in sys.modules:
        g = _stub("gpytorch")

        class ExactGP(torch.nn.Module):
            pass
        for sub, attrs in {"models": {"ExactGP": ExactGP},
                           "distributions": {"MultivariateNormal": object},
                           "likelihoods": {"Likelihood": object, "GaussianLikelihood": object},
                           "kernels": {"Kernel": object}}.items():
            m = _stub(f"gpytorch.{sub}")
            for k, v in attrs.items():
                setattr(m, k, v)
            setattr(g, sub, m)
    import atomai  # noqa
    return atomai
