"""Model registry entry for the ISTVT path — mirror of `model_selection` / `TransferModel` in the reference's
`network/models.py:28-282` and `network/models_copy.py:26-283`.

Only the two names the ISTVT path uses are served:
  'xception'   -> TransferModel wrapping the Xception parameter tree (what XceptionVidTr.xcep is,
                  vivit.py:196; forward through the whole backbone is out of scope);
  'resnet_3d'  -> the ISTVT model itself (the registry key the reference reuses for it,
                  models.py:175-180, models_copy.py:173-175).
Any other name raises: the rest of the reference's model zoo is out of scope.
"""
from __future__ import annotations

import torch
from torch import nn

from .xception import Xception


class TransferModel(nn.Module):
    def __init__(self, modelchoice: str, num_out_classes: int = 2, dropout: float = 0.5, batch_size: int = 16):
        super().__init__()
        self.modelchoice = modelchoice
        if modelchoice == "xception":
            self.model = Xception(num_classes=1000)
            num_ftrs = self.model.last_linear.in_features
            if not dropout:
                self.model.last_linear = nn.Linear(num_ftrs, num_out_classes)
            else:
                self.model.last_linear = nn.Sequential(nn.Dropout(p=dropout), nn.Linear(num_ftrs, num_out_classes))
        elif modelchoice == "resnet_3d":
            from .vivit.vivit import XceptionVidTr
            self.model = XceptionVidTr()
        else:
            raise NotImplementedError(f"model '{modelchoice}' is outside the ISTVT hot path served by istvt_b200")

    def low_level_features(self, x: torch.Tensor) -> torch.Tensor:
        return self.model.low_level_features(x)

    def get_model(self) -> nn.Module:
        return self.model

    def forward(self, x):
        return self.model(x)


def model_selection(modelname: str, num_out_classes: int, dropout=None, batch_size: int = 16) -> nn.Module:
    """Same signature as the reference (`network/models.py:240-282`).

    'xception' returns the TransferModel wrapper (reference behaviour, note: it ignores `dropout` and uses
    the wrapper's default 0.5 exactly like models_copy.py:246-248); every other name returns the bare
    module (`.get_model()`, models.py:282).
    """
    if modelname == "xception":
        return TransferModel(modelchoice="xception", num_out_classes=num_out_classes)
    return TransferModel(modelchoice=modelname, num_out_classes=num_out_classes, batch_size=batch_size).get_model()
