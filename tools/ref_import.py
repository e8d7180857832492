"""Import shim for the *reference* CFPNet modules (container-only tooling).

Used ONLY by ``tools/make_golden.py`` to generate the committed fixtures under
``tests/golden/``.  Nothing in ``tests/``, ``bench.py`` or the product package
imports this file: ``/root/reference`` does not exist on the GPU box.

The reference needs two shims to import:
  * ``timm`` is not installed -> stub the three names ``convnext.py`` /
    ``encoder.py`` import;
  * ``src/config.py`` parses ``sys.argv`` at import time -> point it at one of
    the reference's own config files first.
"""
import sys
import types

import torch.nn as nn

REF_ROOT = "/root/reference"
CFG_COMBINE1 = (REF_ROOT + "/configs/train_deltar_change_embedding_no_clip_grad_"
                "hist_encoder_optimized_10x_combine1.txt")


def import_reference(cfg=CFG_COMBINE1):
    """Returns the reference's ``src`` package namespace pieces as a dict."""
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        registry = types.ModuleType("timm.models.registry")
        layers.trunc_normal_ = nn.init.trunc_normal_

        class DropPath(nn.Identity):
            def __init__(self, p=0.0):
                super().__init__()

        layers.DropPath = DropPath
        registry.register_model = lambda f: f
        timm.models, models.layers, models.registry = models, layers, registry
        sys.modules.update({"timm": timm, "timm.models": models,
                            "timm.models.layers": layers,
                            "timm.models.registry": registry})
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    saved = sys.argv
    sys.argv = ["ref", "@" + cfg]
    try:
        from src.config import args
        from src.models import fusion, encoder, transformer, convnext, attention
        from src.utils import dataloader
    finally:
        sys.argv = saved
    return dict(args=args, fusion=fusion, encoder=encoder, transformer=transformer,
                convnext=convnext, attention=attention, dataloader=dataloader)
