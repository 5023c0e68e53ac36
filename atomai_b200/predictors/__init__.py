from .epredictor import EnsemblePredictor, ensemble_locate
from .predictor import BasePredictor, ImSpecPredictor, Locator, SegPredictor

__all__ = ["BasePredictor", "SegPredictor", "ImSpecPredictor", "Locator", "EnsemblePredictor",
           "ensemble_locate"]
