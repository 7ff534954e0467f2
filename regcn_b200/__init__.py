"""regcn_b200 -- B200 (sm_100a) implementation of RE-GCN's per-snapshot evolution and all-entity
scoring path behind the reference's own module signatures.  See DESIGN.md for the path and its boundary.

The kernels live in libregcn_b200.so (C ABI: include/regcn_b200.h); there is no CPU or eager fallback.
"""
from . import _lib  # noqa: F401
from .graph import SnapshotCache, SnapshotGraph, build_sub_graph, build_sub_graphs  # noqa: F401
from .layers import RGCNBlockLayer, UnionRGCNLayer  # noqa: F401
from .rrgcn import RecurrentRGCN, RGCNCell  # noqa: F401
from .decoder import ConvTransE, ConvTransR  # noqa: F401
from .hyperbolic_layers import (HyperbolicRGCNCell, HyperbolicUnionRGCNLayer, LorentzRGCNCell,  # noqa: F401
                                LorentzRGCNLayer)
from .hyperbolic_decoder import (HyperbolicAttH, HyperbolicAttHRel, HyperbolicConvTransE,  # noqa: F401
                                 HyperbolicConvTransR, HyperbolicMuRP, HyperbolicMuRPRel, HyperbolicRotH,
                                 HyperbolicRotHRel)
from .hyperbolic_model import HyperbolicRecurrentRGCN  # noqa: F401
from . import utils  # noqa: F401
from . import knowledge_graph  # noqa: F401  (on-disk TKG format reader)
from . import optim, train, train_hyp  # noqa: F401  (training step: get_loss with gradients + clipped Adam)
from .evaluate import test  # noqa: F401  (the reference's evaluation loop, src/main.py:33)
from .fit import fit_epoch  # noqa: F401  (the reference's training epoch, src/main.py:213-246)

__version__ = "0.1.0"
