"""Run one layer entry point a few times (for ncu captures).  usage: prof_layer.py <lkpm|dapm|twins|d2i> <level> [B]"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from cfpnet_b200 import _lib, geometry, synth  # noqa: E402
import cfpnet_b200  # noqa: E402
from cfpnet_b200.config import args  # noqa: E402

name, level = sys.argv[1], int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
C, _, max_res, lk = synth.LEVELS[level]
args.attention_layer = list(synth.COMBINE1_LAYERS)
m = cfpnet_b200.TransformerFusion(C, list(max_res), large_kernel=lk, patch_size=640 // max_res[1])
m.load_state_dict(synth.synthetic_state_dict({k: v.shape for k, v in m.state_dict().items()}, level))
m = m.cuda().bfloat16().eval()
H, W = synth.level_hw("G416", level)
inp = synth.make_inputs("G416", B, levels=())
g = geometry.zone_geometry(inp["patch_info"], max_res[1], H, W)
cg = _lib.CfpGeom.from_geometry(g)
code = _lib.CFP_BF16
packed, pos, pos2, keep = m._cache.get(m, m._pack)
lib = _lib.load()
nbytes = lib.cfp_workspace_bytes(B, H, W, C, m.ws, m.large_kernel, code, ctypes.byref(cg))
work = torch.empty(nbytes, device="cuda", dtype=torch.uint8)
x = torch.randn(B, H * W, C, device="cuda").bfloat16()
feat1 = torch.randn(B, 64, 16, C, device="cuda").bfloat16()
mask = inp["mask"].cuda().to(torch.uint8)
st = _lib.stream_ptr()
for it in range(3):
    if name == "lkpm":
        _lib.call("cfp_lkpm_fwd", x.data_ptr(), B, H, W, C, ctypes.byref(packed[1][1]), work.data_ptr(), nbytes, code, st)
    elif name == "dapm":
        _lib.call("cfp_dapm_fwd", x.data_ptr(), B, H, W, C, ctypes.byref(cg), ctypes.byref(packed[1][0]), work.data_ptr(), nbytes, code, st)
    elif name == "twins":
        _lib.call("cfp_twins_fwd", x.data_ptr(), B, H, W, C, ctypes.byref(packed[2]), work.data_ptr(), nbytes, code, st)
    elif name == "d2i":
        _lib.call("cfp_d2i_fwd", x.data_ptr(), x.data_ptr(), feat1.data_ptr(), pos2, mask.data_ptr(), B, H, W, C, 16,
                  ctypes.byref(cg), ctypes.byref(packed[0]), 0, work.data_ptr(), nbytes, code, st)
torch.cuda.synchronize()
print("done")
