"""uit_mobile_b200 — B200-native (sm_100a) implementation of UIT_Mobile's batched inference hot path:
raw 16 kHz waveform -> log-mel -> UiT-XS/XXS/XXXS encoder -> 537 joint AudioSet + GSC scores.

``uit_mobile_b200.models`` mirrors the reference's ``models`` package; the arithmetic lives in ``libuitk.so``
(C ABI in ``include/uitk.h``, sources in ``uit_mobile_b200/csrc``)."""
from . import models  # noqa: F401
from .models import PRETRAINED_CHECKPOINTS, UITBase, uit_xs, uit_xxs, uit_xxxs  # noqa: F401

__version__ = "0.1.0"
