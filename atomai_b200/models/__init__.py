from .loaders import load_model
from .segmentor import Segmentor

__all__ = ["Segmentor", "load_model"]
