"""Rebind the reference's model classes to the B200 kernels without editing the reference tree.

    import sys; sys.path.insert(0, "/path/to/early-exit-transformer")     # the unmodified reference
    import eec.dropin; eec.dropin.install()
    import train                                                            # reference train.py: now builds eec models

`install()` imports the reference's ``models.model.early_exit`` and replaces ``Early_conformer`` and
``Splitformer`` (early_exit.py:565-634, :227-364) by the eec classes of the same constructor signature,
forward contract and state_dict layout; ``full_conformer`` (AED mode, :637-811) is replaced by ``eec.full_conformer``
(encoder half on the B200 kernels, torch.nn decoders as in the reference).  ``Early_zipformer`` stays the reference's own
(out of this path's scope).  `patch_ctc(train_module)` makes the ``nn.CTCLoss(blank=0, zero_infinity=True)`` the reference's ``train.py``
builds (train.py:259) resolve to ``eec.CTCLoss`` -- scoped to THAT module's ``nn`` name: ``torch.nn`` itself is never modified.

    import train; eec.dropin.patch_ctc(train); train.main()
"""
from __future__ import annotations

import importlib
import types


class _NnView(types.ModuleType):
    """A view of ``torch.nn`` whose ``CTCLoss`` builds ``eec.CTCLoss`` for the reference's configuration; every other attribute is torch's."""

    def __init__(self, nn):
        super().__init__("torch.nn")
        self.__dict__["_nn"] = nn

    def __getattr__(self, name):
        return getattr(self.__dict__["_nn"], name)

    def CTCLoss(self, blank=0, reduction="mean", zero_infinity=False):   # noqa: N802
        import eec
        if reduction == "mean" and zero_infinity:
            return eec.CTCLoss(blank=blank, reduction=reduction, zero_infinity=True)
        return self.__dict__["_nn"].CTCLoss(blank=blank, reduction=reduction, zero_infinity=zero_infinity)


def patch_ctc(module, name: str = "nn"):
    """Rebind `module.<name>` (the reference's `from torch import nn`, train.py:6) to a view of torch.nn whose CTCLoss is eec's.  Only the
    given module's namespace changes."""
    import torch
    if getattr(module, name, None) is not torch.nn and not isinstance(getattr(module, name, None), _NnView):
        raise AttributeError(f"{module.__name__}.{name} is not torch.nn")
    setattr(module, name, _NnView(torch.nn))
    return module


def install(precision: str | None = None):
    import eec

    ref = importlib.import_module("models.model.early_exit")
    ref.Early_conformer = eec.Early_conformer
    ref.Splitformer = eec.Splitformer
    ref.full_conformer = eec.full_conformer
    if precision is not None:
        import os
        os.environ["EEC_PRECISION"] = precision
    return ref
