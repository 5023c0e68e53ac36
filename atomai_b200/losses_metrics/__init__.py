from .losses import BCEWithLogitsLoss, CrossEntropyLoss, MSELoss, select_loss

__all__ = ["select_loss", "CrossEntropyLoss", "BCEWithLogitsLoss", "MSELoss"]
