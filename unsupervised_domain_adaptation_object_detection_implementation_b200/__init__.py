"""B200-native domain-adaptation hot path (GRL + domain classifiers + RoIAlign + instance head +
consistency loss) behind the reference's mmdet-style module surface.  See DESIGN.md."""
from . import _lib  # noqa: F401  (fails loudly when libda_b200.so is missing)
from . import functional, ops, da_heads, da_losses, roi_extractors, hotpath, dist, optim, peer, data, checkpoint  # noqa: F401
from . import registry, config, backbones, detection, detectors  # noqa: F401
from .registry import build_detector, MODELS  # noqa: F401
from .config import Config  # noqa: F401
from .functional import set_engine, get_engine  # noqa: F401

__version__ = "0.1.0"
