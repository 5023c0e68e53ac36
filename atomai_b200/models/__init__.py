from .dgm import VAE, BaseVAE, rVAE
from .dklgp import dklGPR
from .imspec import ImSpec
from .loaders import load_model
from .segmentor import Segmentor

__all__ = ["Segmentor", "ImSpec", "VAE", "rVAE", "BaseVAE", "dklGPR", "load_model"]
