"""Drop-in packages that shadow the reference's imports.

Put this directory FIRST on ``sys.path`` (``wcmc_b200.dropin.install()``) and the reference's
``train_kpcn.py`` runs unchanged against the B200 backend:

  sbmc      -> KPCN, modules.{ConvChain, Autoencoder, KernelApply}   (train_kpcn.py:28-33,
               support/networks.py:4-5)
  support   -> networks.PathNet, losses.*, interfaces.KPCNInterface, utils.crop_like
  ttools    -> modules.image_operators.crop_like                      (train_kpcn.py:34, unused)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def install():
    """Makes `import sbmc`, `import support...`, `import ttools...` resolve to the B200 backend."""
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    return HERE
