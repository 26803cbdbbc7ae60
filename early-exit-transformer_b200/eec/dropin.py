"""Rebind the reference's model classes to the B200 kernels without editing the reference tree.

    import sys; sys.path.insert(0, "/path/to/early-exit-transformer")     # the unmodified reference
    import eec.dropin; eec.dropin.install()
    import train                                                            # reference train.py: now builds eec models

`install()` imports the reference's ``models.model.early_exit`` and replaces ``Early_conformer`` and
``Splitformer`` (early_exit.py:565-634, :227-364) by the eec classes of the same constructor signature,
forward contract and state_dict layout; ``full_conformer`` (AED mode, :637-811) is replaced by ``eec.full_conformer``
(encoder half on the B200 kernels, torch.nn decoders as in the reference).  ``Early_zipformer`` stays the reference's own
(out of this path's scope).  `install(ctc=True)` additionally makes ``torch.nn.CTCLoss`` calls
with the reference's configuration (blank=0, zero_infinity=True) resolve to ``eec.CTCLoss``.
"""
from __future__ import annotations

import importlib


def install(ctc: bool = False, precision: str | None = None):
    import eec

    ref = importlib.import_module("models.model.early_exit")
    ref.Early_conformer = eec.Early_conformer
    ref.Splitformer = eec.Splitformer
    ref.full_conformer = eec.full_conformer
    if precision is not None:
        import os
        os.environ["EEC_PRECISION"] = precision
    if ctc:
        import torch

        _orig = torch.nn.CTCLoss

        def _ctc(blank=0, reduction="mean", zero_infinity=False):
            if reduction == "mean" and zero_infinity:
                return eec.CTCLoss(blank=blank, reduction=reduction, zero_infinity=True)
            return _orig(blank=blank, reduction=reduction, zero_infinity=zero_infinity)

        torch.nn.CTCLoss = _ctc
    return ref
