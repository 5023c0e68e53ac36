"""
Segmentor — user-facing semantic-segmentation model (train + predict), drop-in for
atomai.models.Segmentor (atomai/models/segmentor.py:16-207).
"""
from typing import Dict, Optional, Tuple, Type, Union

import numpy as np
import torch

from ..predictors import SegPredictor
from ..trainers import SegTrainer
from ..utils.nn import get_downsample_factor


def _seg_augmentor(nb_classes: int, **kwargs):
    """On-the-fly augmentor when augmentation kwargs are given, None otherwise
    (atomai/transforms/imaug.py:406-432) — here the GPU implementation of
    atomai_b200.transforms."""
    from ..transforms import seg_augmentor
    return seg_augmentor(nb_classes, **kwargs)


class Segmentor(SegTrainer):
    """
    Model for semantic segmentation-based analysis of images

    Args:
        model: 'Unet' (default), 'dilnet' or a custom fully convolutional torch module
        nb_classes: number of classes in the classification scheme
        **batch_norm, **dropout, **upsampling, **nb_filters, **with_dilation, **layers, **seed:
            as in atomai/models/segmentor.py:16-52

    Example:

    >>> model = Segmentor(nb_classes=3)
    >>> model.fit(images, labels, images_test, labels_test, training_cycles=500, swa=True)
    >>> nn_output, coordinates = model.predict(expdata)
    """
    def __init__(self, model: Type[Union[str, torch.nn.Module]] = "Unet", nb_classes: int = 1,
                 **kwargs) -> None:
        super(Segmentor, self).__init__(model, nb_classes, **kwargs)
        self.downsample_factor = None

    def fit(self, X_train, y_train, X_test=None, y_test=None, loss: str = 'ce',
            optimizer: Optional[Type[torch.optim.Optimizer]] = None,
            training_cycles: int = 1000, batch_size: int = 32, compute_accuracy: bool = False,
            full_epoch: bool = False, swa: bool = False, perturb_weights: bool = False,
            **kwargs):
        """
        Compiles a trainer and performs model training — arguments and kwargs (lr_scheduler,
        print_loss, accuracy_metrics, filename, plot_training_history, batch_seed, memory_alloc)
        as in atomai/models/segmentor.py:61-149.
        """
        self.compile_trainer((X_train, y_train, X_test, y_test), loss, optimizer,
                             training_cycles, batch_size, compute_accuracy, full_epoch, swa,
                             perturb_weights, **kwargs)
        self.augment_fn = _seg_augmentor(self.nb_classes, **kwargs)
        _ = self.run()

    def predict(self, imgdata: Union[np.ndarray, torch.Tensor], refine: bool = False,
                logits: bool = True, resize: Tuple[int, int] = None,
                compute_coords: bool = True, **kwargs) -> Tuple[np.ndarray, Dict[int, np.ndarray]]:
        """
        Apply (trained) model to new data: returns the semantically segmented image(s) and a
        dictionary of (atomic) coordinates — atomai/models/segmentor.py:151-200
        (**thresh, **d, **num_batches, **norm, **verbose).
        """
        if self.downsample_factor is None:
            self.downsample_factor = get_downsample_factor(self.net)
        use_gpu = self.device == 'cuda'
        return SegPredictor(self.net, refine, resize, use_gpu, logits, nb_classes=self.nb_classes,
                            downsampling=self.downsample_factor, **kwargs).run(
                                imgdata, compute_coords, **kwargs)

    def load_weights(self, filepath: str) -> None:
        """Loads a saved weights dictionary (atomai/models/segmentor.py:202-207)."""
        weight_dict = torch.load(filepath, map_location=self.device)
        self.net.load_state_dict(weight_dict)
