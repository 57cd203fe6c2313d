"""vit4hep_b200: B200-native (sm_100a) implementation of vit4hep's CFM 3D-ViT hot path.

Host-side mirror of the reference interface (`ViT`, the CFM wrappers) over the C ABI in
include/vit4hep_b200.h; see DESIGN.md and INTEGRATION.md.
"""
from .vit import ViT  # noqa: F401
from .cfm import (CFM, CaloChallengeCFM, CaloChallengeCFM_DS1, CaloGANCFM, CaloHadCFM, LEMURSCFM,  # noqa: F401
                  GraphedTrainStep, PatchGeometry)

from .optim import ExponentialMovingAverage, FusedAdamW  # noqa: F401
from .energy import ParallelTransformer  # noqa: F401
from .postprocess import FusedReverseTransforms  # noqa: F401
from .preprocess import FusedForwardTransforms, ShowerDataset  # noqa: F401

__version__ = "0.1.0"
