"""B200-native SSL head + EMA for taindp98/Endoscopy-Image-Classification.

The directory name carries a hyphen (it is the reference's name plus ``_b200``),
so import it through the shim at the repo root::

    import endoscopy_image_classification_b200 as eic
    from endoscopy_image_classification_b200.loss import consistency_loss
    from endoscopy_image_classification_b200.ema import ModelEMA

or make the reference's own module names resolve to this package
(``from loss import consistency_loss`` etc.) with :func:`install_as_reference_modules`.

Only the hot path lives here: ``loss`` (consistency_loss,
ce_loss, PolyLoss), ``ema`` (ModelEMA), ``comatch_head`` (the inline head of
``comatch.py:162-220``), the trainer shells ``fixmatch`` / ``comatch`` /
``semiformer`` that call them, ``utils`` (AttrDict / get_config / AverageMeter)
and ``csrc`` (the sm_100a kernels behind ``include/b200ssl.h``).
"""
import sys as _sys

__version__ = "0.1.0"

_REFERENCE_MODULES = ("loss", "ema", "utils", "fixmatch", "comatch", "semiformer", "optimizer", "lr_scheduler")


def install_as_reference_modules(names=_REFERENCE_MODULES) -> None:
    """Register this package's modules under the reference's flat module names so
    that reference-style driver code (``from loss import ce_loss, consistency_loss``,
    ``from ema import ModelEMA``, ``from comatch import CoMatch`` ...) picks up the
    B200 implementation unchanged."""
    import importlib
    for n in names:
        _sys.modules[n] = importlib.import_module(f"{__name__}.{n}")
