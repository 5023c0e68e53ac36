from .predictor import BasePredictor, ImSpecPredictor, Locator, SegPredictor

__all__ = ["BasePredictor", "SegPredictor", "ImSpecPredictor", "Locator"]
