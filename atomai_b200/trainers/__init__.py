from .etrainer import BaseEnsembleTrainer, EnsembleTrainer
from .gptrainer import GPTrainer, dklGPTrainer
from .trainer import BaseTrainer, ImSpecTrainer, SegTrainer

__all__ = ["BaseTrainer", "SegTrainer", "ImSpecTrainer", "GPTrainer", "dklGPTrainer",
           "BaseEnsembleTrainer", "EnsembleTrainer"]
