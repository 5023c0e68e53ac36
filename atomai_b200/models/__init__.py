from .dgm import VAE, BaseVAE, rVAE
from .imspec import ImSpec
from .loaders import load_model
from .segmentor import Segmentor

__all__ = ["Segmentor", "ImSpec", "VAE", "rVAE", "BaseVAE", "load_model"]
