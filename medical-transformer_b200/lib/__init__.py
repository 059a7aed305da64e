"""Drop-in replacement for the reference's `lib` package (reference lib/__init__.py).

Only the model path is provided: `lib.models.axialunet`, `lib.models.axialnet.{gated, MedT,
logo}` -- exactly what train.py:95-102 / test.py:90-97 call.  The reference's ImageNet
scaffolding (build_dataloader / build_model / build_optimizer / Metric, never called by
train.py or test.py) is out of scope."""
from . import models  # noqa: F401

__all__ = ["models"]
