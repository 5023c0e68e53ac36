from .predictor import BasePredictor, Locator, SegPredictor

__all__ = ["BasePredictor", "SegPredictor", "Locator"]
