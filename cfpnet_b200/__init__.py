"""cfpnet_b200 — B200-native CFP fusion + cross-zone propagation path.

Host-side mirror of the reference's ``src/models`` interface for this path
(``HistogramEncoder``, ``TransformerFusion``) over the C ABI of ``include/cfp.h``
(``libcfp.so``: hand-written sm_100a CUDA).  See DESIGN.md.
"""
from .config import args
from .encoder import HistogramEncoder
from .fusion import TransformerFusion
from .pipeline import FusionPath
from .geometry import ZoneGeometry, collate_patch_info, patch_info_from_rect_data, zone_geometry

__all__ = ["args", "HistogramEncoder", "TransformerFusion", "FusionPath", "ZoneGeometry", "collate_patch_info",
           "patch_info_from_rect_data", "zone_geometry"]
