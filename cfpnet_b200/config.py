"""Flags of the reference's global ``src.config.args`` that the fusion path reads
(``src/models/fusion.py:16,25,134,154``), with the values of the reference's
``..._10x_combine1.txt`` config.  The reference parses them from ``sys.argv`` at
import time; here they are a plain mutable namespace.  When the modules are
dropped into the reference ``use_reference_args`` binds this
namespace to the reference's own ``args`` object instead.
"""
from types import SimpleNamespace

args = SimpleNamespace(
    zone_sample_num=16,
    attention_layer=["hist2image", "combine1", "image", "hist2image", "combine1", "image"],
    change_embedding=True,
    no_skip_inside=False,
)


def use_reference_args(ref_args) -> None:
    """Read the path's flags from the reference's ``src.config.args``."""
    for k in vars(args):
        setattr(args, k, getattr(ref_args, k))
