"""geomloss_b200 — B200-native (sm_100a) engine for the Sinkhorn / kernel-MMD hot path of GeomLoss.

Public surface (drop-in for the point-cloud losses of jeanfeydy/geomloss):

    from geomloss_b200 import SamplesLoss
    L = SamplesLoss("sinkhorn", p=2, blur=.05)(x, y)      # x, y float32 CUDA tensors

Operator level (the reference's softmin seam):  ``geomloss_b200.ops.softmin`` / ``kernel_conv``.
C ABI: ``include/b200ot.h`` (``geomloss_b200/libb200ot.so``, built by ``__graft_entry__.build()``).
"""
from .samples_loss import SamplesLoss  # noqa: F401
from .sinkhorn_images import sinkhorn_divergence  # noqa: F401  (the reference exports the IMAGE routine here)
from . import ot  # noqa: F401  (geomloss.ot.solve_sample facade)
from .barycenter_images import ImagesBarycenter  # noqa: F401

__version__ = "0.1.0"
__all__ = ["SamplesLoss", "ImagesBarycenter", "sinkhorn_divergence", "ot"]
