"""Debug: tensor-core depthwise conv (planar output read back from the workspace) vs torch conv2d (tooling)."""
import ctypes, sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from cfpnet_b200 import _lib, synth
import cfpnet_b200
from cfpnet_b200.config import args
from cfpnet_b200.packing import fold_bn

level = int(sys.argv[1]); B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
C, _, max_res, lk = synth.LEVELS[level]
args.attention_layer = list(synth.COMBINE1_LAYERS)
m = cfpnet_b200.TransformerFusion(C, list(max_res), large_kernel=lk, patch_size=640 // max_res[1])
m.load_state_dict(synth.synthetic_state_dict({k: v.shape for k, v in m.state_dict().items()}, level))
m = m.cuda().bfloat16().eval()
H, W = synth.level_hw("G416", level)
inp = synth.make_inputs("G416", B, levels=())
from cfpnet_b200 import geometry
g = geometry.zone_geometry(inp["patch_info"], max_res[1], H, W)
cg = _lib.CfpGeom.from_geometry(g)
code = _lib.CFP_BF16
packed, pos, pos2, keep = m._cache.get(m, m._pack)
lib = _lib.load()
nbytes = lib.cfp_workspace_bytes(B, H, W, C, m.ws, m.large_kernel, code, ctypes.byref(cg))
work = torch.zeros(nbytes, device="cuda", dtype=torch.uint8)
if len(sys.argv) > 3:
    work.fill_(0xFF)          # NaN bit patterns: stale workspace must not leak into results
x = torch.randn(B, H * W, C, device="cuda").bfloat16()
x0 = x.clone()
_lib.call("cfp_lkpm_fwd", x.data_ptr(), B, H, W, C, ctypes.byref(packed[1][1]), work.data_ptr(), nbytes, code, _lib.stream_ptr())
torch.cuda.synchronize()
WO = (W + 7) // 8 * 8                      # row pitch of the planar output
out_bytes = (B * C * H * WO * 2 + 255) // 256 * 256
end = lib.cfp_workspace_bytes(B, H, W, C, 0, m.large_kernel, code, None)     # cfp_lkpm_fwd's own layout (no zones, no sr)
y = work[end - out_bytes: end - out_bytes + B * C * H * WO * 2].view(torch.bfloat16).view(B, C, H, WO)[..., :W].float().cpu()
blk = m.layers[1].large_kernel_path
scale, shift = fold_bn(blk.bn1)
taps = (blk.dwconv2.weight.detach().float()[:, 0] * scale[:, None, None]).to(torch.bfloat16).float().cpu()
bias = (blk.dwconv2.bias.detach().float() * scale + shift).cpu()
xin = x0.float().cpu().view(B, H, W, C).permute(0, 3, 1, 2)
ref = torch.relu(torch.nn.functional.conv2d(xin.double(), taps.double()[:, None], padding=(lk - 1) // 2, groups=C) + bias.double()[None, :, None, None]).float()
err = (y - ref)
print("rel-L2", float(err.norm() / ref.norm()), "nan", int(torch.isnan(y).sum()))
rowerr = err.pow(2).sum(dim=(0, 1, 3)).sqrt() / ref.pow(2).sum(dim=(0, 1, 3)).sqrt().clamp_min(1e-9)
colerr = err.pow(2).sum(dim=(0, 1, 2)).sqrt() / ref.pow(2).sum(dim=(0, 1, 2)).sqrt().clamp_min(1e-9)
frerr = err.pow(2).sum(dim=(1, 2, 3)).sqrt() / ref.pow(2).sum(dim=(1, 2, 3)).sqrt()
print("per-frame", [round(float(v), 4) for v in frerr])
print("rows", [round(float(v), 3) for v in rowerr])
print("cols", [round(float(v), 3) for v in colerr])
