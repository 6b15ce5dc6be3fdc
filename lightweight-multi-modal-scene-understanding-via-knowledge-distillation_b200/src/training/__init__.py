from .trainer import SegmentationMetrics, Trainer  # noqa: F401
from .optim import FlatAdamW  # noqa: F401
