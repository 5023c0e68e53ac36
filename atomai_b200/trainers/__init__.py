from .trainer import BaseTrainer, ImSpecTrainer, SegTrainer

__all__ = ["BaseTrainer", "SegTrainer", "ImSpecTrainer"]
