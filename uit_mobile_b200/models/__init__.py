"""Name-lookup surface of the reference's ``models`` package (models/__init__.py:1-2): callers do
``getattr(models, cfg['model'])(**model_args)`` (run.py:127, evaluate.py:34, inference.py:46)."""
from .uit import *  # noqa: F401,F403
from .uit import UITBase, PRETRAINED_CHECKPOINTS, uit_xs, uit_xxs, uit_xxxs  # noqa: F401
from .mobilenetv2 import MobileNetV2  # noqa: F401  (models/__init__.py:2)
