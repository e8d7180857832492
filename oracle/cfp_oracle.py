"""CPU oracle for the CFP fusion + cross-zone propagation path.

TEST INFRASTRUCTURE ONLY.  This file restates, in plain functional PyTorch on
the CPU, the algorithm of the reference's hot path (denyingmxd/CFPNet,
``src/models``).  It is imported only by ``tests/``, by
``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, always as the checker or the timed CPU baseline and
never by the product package ``cfpnet_b200`` (which fails loudly when its CUDA
library is missing).

Parity pinning: the reference has no tests or golden vectors of its own
(SURVEY.md §4), so this restatement is pinned against outputs of the reference
modules themselves, imported in the build container and stored under
``tests/golden/`` by ``tools/make_golden.py`` (committed).  ``tests/test_oracle_golden.py``
checks every function here against those fixtures.  TRAIN mode (BatchNorm batch
statistics, ``bn_stats=`` dict; gradients = autograd over this file) is pinned the
same way on one forward + backward of the reference modules in ``.train()`` mode
(``tools/make_golden_train.py`` -> ``tests/golden/train_*.npz``,
``tests/test_oracle_train_golden.py``).

Every function cites the reference file:line it follows (paths relative to the
reference root).  All functions are dtype-agnostic: run them in float32 for
the "reference fp32" comparator and in float64 for ground truth.

Weights are passed as a flat ``state_dict``-style mapping with the reference's
key names; ``sub(sd, "layers.0.")`` makes a prefixed view.
"""
from __future__ import annotations

import math
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

BN_EPS = 1e-5          # nn.BatchNorm{1,2}d default
LN_EPS = 1e-5          # nn.LayerNorm default      (transformer.py:38-39,133)
LKPM_LN_EPS = 1e-6     # convnext.py:31  LayerNorm(dim, eps=1e-6)
ATTN_EPS = 1e-6        # attention.py:11


class _Sub(Mapping):
    """Prefix view over a flat state dict."""

    def __init__(self, sd, prefix):
        self.sd, self.prefix = sd, prefix

    def __getitem__(self, k):
        return self.sd[self.prefix + k]

    @property
    def full_prefix(self):
        return getattr(self.sd, "full_prefix", "") + self.prefix

    def __iter__(self):
        n = len(self.prefix)
        return (k[n:] for k in self.sd if k.startswith(self.prefix))

    def __len__(self):
        return sum(1 for _ in self)


def sub(sd, prefix):
    return _Sub(sd, prefix)


# --------------------------------------------------------------------------
# a2: integer zone geometry
# --------------------------------------------------------------------------
def patch_info_from_rects(rect: Tensor) -> dict:
    """Integer patch geometry of one frame from its zone rectangles.

    Follows src/utils/dataloader.py:13-40.  ``rect`` is ``[Z,4]`` float rows
    ``[y0,x0,y1,x1]`` in input pixels.  Pads are relative to the hard-coded
    480x640 canvas (:20-23); start/end indices are truncated toward zero by the
    float->int32 cast (:27-28,31-32); patch size is a ceil (:29-30).
    """
    r = rect.to(torch.float32)
    zone_num = int(math.sqrt(r.shape[0]))
    y0, x0, y1, x1 = r[:, 0], r[:, 1], r[:, 2], r[:, 3]
    max_ph = int((y1 - y0).max().to(torch.int32))
    max_pw = int((x1 - x0).max().to(torch.int32))
    pad_h_px = int(max(float((-y0.clamp(max=0)).abs().max()),
                       float((y1.clamp(min=480) - 480).max())))
    pad_w_px = int(max(float((-x0.clamp(max=0)).abs().max()),
                       float((x1.clamp(min=640) - 640).max())))
    out = {}
    for cps in (4, 8, 16):
        out[cps] = {
            "pad_size": torch.tensor([math.ceil(pad_h_px / cps), math.ceil(pad_w_px / cps)],
                                     dtype=torch.int32),
            "patch_size": torch.tensor([math.ceil(max_ph / cps), math.ceil(max_pw / cps)],
                                       dtype=torch.int32),
            "index_wo_pad": torch.stack([(y0 / cps).min(), (x0 / cps).min(),
                                         (y1 / cps).max(), (x1 / cps).max()]).to(torch.int32),
        }
    out["zone_num"] = zone_num
    return out


def collate_patch_info(infos: Sequence[dict]) -> dict:
    """What torch's default_collate does to a list of per-frame patch_info
    dicts: leading batch dim, ``zone_num`` -> [B]."""
    out = {}
    for cps in (4, 8, 16):
        out[cps] = {k: torch.stack([i[cps][k] for i in infos]) for k in infos[0][cps]}
    out["zone_num"] = torch.tensor([i["zone_num"] for i in infos])
    return out


def zone_geometry(patch_info: dict, max_w: int, H: int, W: int) -> Dict[str, int]:
    """Per-level host integers of ``TransformerFusion.forward``.

    Follows src/models/fusion.py:41 (``conv_patch_size = 640 / max_resolution[1]``)
    and :67-84 (batch-wise max / min, pad shift, interpolate predicate).
    """
    cps = 640 / max_w
    info = patch_info[cps]           # float key hashes equal to the int key
    zn = int(patch_info["zone_num"][0])
    pad_h, pad_w = (int(v) for v in info["pad_size"].max(dim=0)[0])
    p1, p2 = (int(v) for v in info["patch_size"].max(dim=0)[0])
    sy0, sx0 = (int(v) for v in info["index_wo_pad"].min(dim=0)[0][0:2])
    ey0, ex0 = (int(v) for v in info["index_wo_pad"].max(dim=0)[0][2:4])
    g = dict(zone_num=zn, pad_h=pad_h, pad_w=pad_w, p1=p1, p2=p2,
             sy_wo=sy0, sx_wo=sx0, ey_wo=ey0, ex_wo=ex0,
             sy=sy0 + pad_h, ey=ey0 + pad_h, sx=sx0 + pad_w, ex=ex0 + pad_w)
    g["tzh"], g["tzw"] = g["ey"] - g["sy"], g["ex"] - g["sx"]
    g["interpolate"] = int(g["tzh"] != p1 * zn or g["tzw"] != p2 * zn)
    clip = lambda v, hi: min(max(v, 0), hi)
    # the in-image zone rectangle of fusion.py:104
    g["ry0"], g["ry1"] = clip(sy0, H), clip(ey0, H)
    g["rx0"], g["rx1"] = clip(sx0, W), clip(ex0, W)
    return g


# --------------------------------------------------------------------------
# a3: masks
# --------------------------------------------------------------------------
def zone_masks(g: Mapping[str, int], mask: Tensor, B: int, H: int, W: int, D: int
               ) -> Tuple[Tensor, Tensor, Tensor]:
    """(zone_mask [B,H*W,D], hist_mask [(B*Z),p1*p2,D], pad_mask [B*tzh*tzw*D]).

    Follows src/models/fusion.py:103-120.  All bool.
    """
    zn, p1, p2 = g["zone_num"], g["p1"], g["p2"]
    zm = torch.zeros(B, H, W, dtype=torch.bool)
    zm[:, g["ry0"]:g["ry1"], g["rx0"]:g["rx1"]] = True
    zone_mask = zm.reshape(B, H * W, 1).expand(B, H * W, D).contiguous()
    hist_mask = mask.reshape(B * zn * zn, 1, 1).expand(B * zn * zn, p1 * p2, D).contiguous()
    tzh, tzw = g["tzh"], g["tzw"]
    pm = torch.ones(B, tzh, tzw, D, dtype=torch.bool)
    if g["pad_h"] > 0 or g["pad_w"] > 0:
        top, left = max(-g["sy_wo"], 0), max(-g["sx_wo"], 0)
        bot, right = max(g["ey_wo"] - H, 0), max(g["ex_wo"] - W, 0)
        pm[:, :top] = False
        if bot > 0:
            pm[:, tzh - bot:] = False
        pm[:, :, :left] = False
        if right > 0:
            pm[:, :, tzw - right:] = False
    return zone_mask, hist_mask, pm.reshape(-1)


# --------------------------------------------------------------------------
# a1: histogram encoder
# --------------------------------------------------------------------------
BN_MOMENTUM = 0.1      # nn.BatchNorm{1,2}d default


def _bn(x: Tensor, p: Mapping, name: str, dim: int, bn_stats: Optional[dict] = None) -> Tensor:
    """nn.BatchNorm{1,2}d over dim ``dim``.  ``bn_stats is None``: eval mode (running statistics).  A dict: TRAIN mode
    (``model.train()``, train.py:75) - normalise with the batch mean / biased variance over every other dim and record
    the module's updated buffers under its full state_dict name: running = (1 - 0.1) * running + 0.1 * batch, with the
    UNBIASED batch variance, ``num_batches_tracked + 1`` (torch semantics; the reference uses the module defaults)."""
    shape = [1] * x.dim()
    shape[dim] = -1
    w, b = p[name + ".weight"].to(x.dtype).view(shape), p[name + ".bias"].to(x.dtype).view(shape)
    if bn_stats is None:
        inv = torch.rsqrt(p[name + ".running_var"].to(x.dtype) + BN_EPS)
        return (x - p[name + ".running_mean"].to(x.dtype).view(shape)) * (inv.view(shape) * w) + b
    dims = [d for d in range(x.dim()) if d != dim]
    n = x.numel() // x.shape[dim]
    mean = x.mean(dims)
    var = x.var(dims, unbiased=False)
    full = getattr(p, "full_prefix", "") + name
    with torch.no_grad():
        bn_stats[full + ".running_mean"] = (1 - BN_MOMENTUM) * p[name + ".running_mean"].to(x.dtype) + BN_MOMENTUM * mean
        bn_stats[full + ".running_var"] = ((1 - BN_MOMENTUM) * p[name + ".running_var"].to(x.dtype)
                                           + BN_MOMENTUM * var * (n / max(n - 1, 1)))
        bn_stats[full + ".num_batches_tracked"] = p[name + ".num_batches_tracked"] + 1
    return (x - mean.view(shape)) * (torch.rsqrt(var + BN_EPS).view(shape) * w) + b


def _bn_eval(x: Tensor, p: Mapping, name: str, dim: int) -> Tensor:
    return _bn(x, p, name, dim, None)


def pointnet_block(p: Mapping, x: Tensor, bn_stats: Optional[dict] = None) -> Tensor:
    """3 x [pointwise conv + BN(eval) + ReLU] on the last dim.

    Follows src/models/encoder.py:17-24 (Conv1d k=1 == per-sample linear).
    ``x``: [..., Cin] -> [..., Cout].
    """
    for i in (1, 2, 3):
        w = p[f"conv{i}.weight"].to(x.dtype)[:, :, 0]
        x = F.linear(x, w, p[f"conv{i}.bias"].to(x.dtype))
        x = torch.relu(_bn(x, p, f"bn{i}", x.dim() - 1, bn_stats))
    return x


def hist_encoder(sd: Mapping, hist: Tensor, bn_stats: Optional[dict] = None) -> List[Tensor]:
    """``hist``: [B,Z,N] zone depth samples -> 3 token tensors [B,Z,N,{32,64,128}].

    Follows src/models/encoder.py:45-50 (the three extractors are chained and
    all three outputs are returned) and deltar.py:40 (``unsqueeze(-1)``).
    """
    x = hist.unsqueeze(-1)
    outs = []
    for i in (1, 2, 3):
        x = pointnet_block(sub(sd, f"hist_extractor{i}.pointnet_encoder."), x, bn_stats)
        outs.append(x)
    return outs


# --------------------------------------------------------------------------
# linear attention + LoFTR layer (shared by a5, a6, a9)
# --------------------------------------------------------------------------
def linear_attention(q: Tensor, k: Tensor, v: Tensor, nhead: int) -> Tensor:
    """``q``: [n,L,C], ``k``/``v``: [n,S,C] -> message [n,L,C].

    Follows src/models/attention.py:31-49: feature map elu(x)+1 on q and k,
    values pre-divided by S, KV = sum_s K^T V per head, Z = 1/(Q.sum_s K + eps),
    message = Q KV Z S.
    """
    n, L, C = q.shape
    S = k.shape[1]
    d = C // nhead
    Q = F.elu(q).add(1).view(n, L, nhead, d)
    K = F.elu(k).add(1).view(n, S, nhead, d)
    V = (v / S).view(n, S, nhead, d)
    KV = torch.einsum("nshd,nshv->nhdv", K, V)
    Z = 1.0 / (torch.einsum("nlhd,nhd->nlh", Q, K.sum(dim=1)) + ATTN_EPS)
    out = torch.einsum("nlhd,nhdv,nlh->nlhv", Q, KV, Z) * S
    return out.reshape(n, L, C)


def loftr_layer(p: Mapping, x: Tensor, source: Tensor, nhead: int) -> Tensor:
    """Follows src/models/transformer.py:41-71 (no masks are ever passed on
    this path: fusion.py:143, transformer.py:107,149)."""
    dt = x.dtype
    q = F.linear(x, p["q_proj.weight"].to(dt))
    k = F.linear(source, p["k_proj.weight"].to(dt))
    v = F.linear(source, p["v_proj.weight"].to(dt))
    msg = linear_attention(q, k, v, nhead)
    msg = F.linear(msg, p["merge.weight"].to(dt))
    C = x.shape[-1]
    msg = F.layer_norm(msg, (C,), p["norm1.weight"].to(dt), p["norm1.bias"].to(dt), LN_EPS)
    h = torch.relu(F.linear(torch.cat([x, msg], dim=-1), p["mlp.0.weight"].to(dt)))
    msg = F.linear(h, p["mlp.2.weight"].to(dt))
    msg = F.layer_norm(msg, (C,), p["norm2.weight"].to(dt), p["norm2.bias"].to(dt), LN_EPS)
    return x + msg


# --------------------------------------------------------------------------
# a9: Twins "image" layer
# --------------------------------------------------------------------------
def twins_window_size(max_res: Sequence[int]) -> int:
    """fusion.py:28"""
    return math.ceil(math.sqrt(math.sqrt(max_res[0] * max_res[1])))


def lsa(p: Mapping, x: Tensor, H: int, W: int, ws: int) -> Tensor:
    """Locally-grouped self attention, 8 heads (transformer.py:78,89-116).
    Zero-padded cells take part as keys (K = elu(0)+1 = 1, V = 0)."""
    B, N, C = x.shape
    t = x.view(B, H, W, C)
    pb, pr = (ws - H % ws) % ws, (ws - W % ws) % ws
    t = F.pad(t, (0, 0, 0, pr, 0, pb))
    Hp, Wp = H + pb, W + pr
    nh, nw = Hp // ws, Wp // ws
    t = t.view(B, nh, ws, nw, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B * nh * nw, ws * ws, C)
    t = loftr_layer(sub(p, "encoder_layer."), t, t, 8)
    t = t.view(B, nh, nw, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
    return t[:, :H, :W].reshape(B, N, C)


def gsa(p: Mapping, x: Tensor, H: int, W: int, ws: int) -> Tensor:
    """Global sub-sampled attention, 8 heads (transformer.py:122,138-150):
    keys/values come from a stride-ws conv + LayerNorm of the whole map."""
    B, N, C = x.shape
    dt = x.dtype
    m = x.transpose(1, 2).reshape(B, C, H, W)
    s = F.conv2d(m, p["sr.weight"].to(dt), p["sr.bias"].to(dt), stride=ws)
    s = s.reshape(B, C, -1).transpose(1, 2)
    s = F.layer_norm(s, (C,), p["norm.weight"].to(dt), p["norm.bias"].to(dt), LN_EPS)
    return loftr_layer(sub(p, "encoder_layer."), x, s, 8)


def twins_layer(p: Mapping, x: Tensor, H: int, W: int, ws: int) -> Tensor:
    """transformer.py:160-165"""
    return gsa(sub(p, "gsa."), lsa(sub(p, "lga."), x, H, W, ws), H, W, ws)


# --------------------------------------------------------------------------
# a6: DAPM (direct-attention propagation)
# --------------------------------------------------------------------------
def dapm(p: Mapping, feat0: Tensor, g: Mapping[str, int], H: int, W: int, nhead: int = 4,
         bn_stats: Optional[dict] = None) -> Tensor:
    """Follows src/models/transformer.py:204-248.  Outside-zone tokens query
    inside-zone tokens (raster order), the message map is zero inside the zone,
    then conv3x3(2C->C) -> BN -> conv3x3(C->C) -> BN (no activation, :242) and a
    residual.  ``merge``/``mlp``/``norm*`` of this layer are never used."""
    B, N, C = feat0.shape
    dt = feat0.dtype
    inside = torch.zeros(H, W, dtype=torch.bool)
    inside[g["ry0"]:g["ry1"], g["rx0"]:g["rx1"]] = True
    inside = inside.reshape(-1)
    fin, fout = feat0[:, inside], feat0[:, ~inside]
    q = F.linear(fout, p["q_proj.weight"].to(dt))
    k = F.linear(fin, p["k_proj.weight"].to(dt))
    v = F.linear(fin, p["v_proj.weight"].to(dt))
    msg = linear_attention(q, k, v, nhead)
    tmp = torch.zeros_like(feat0)
    tmp[:, ~inside] = msg
    m = torch.cat([feat0, tmp], dim=2).transpose(1, 2).reshape(B, 2 * C, H, W)
    m = _bn(F.conv2d(m, p["conv1.weight"].to(dt), padding=1), p, "bn1", 1, bn_stats)
    m = _bn(F.conv2d(m, p["conv2.weight"].to(dt), padding=1), p, "bn2", 1, bn_stats)
    return m.reshape(B, C, N).transpose(1, 2) + feat0


# --------------------------------------------------------------------------
# a7: LKPM (large-kernel depthwise propagation)
# --------------------------------------------------------------------------
def lkpm(p: Mapping, feat0: Tensor, H: int, W: int, bn_stats: Optional[dict] = None) -> Tensor:
    """Follows src/models/convnext.py:42-58: depthwise kxk (+bias) -> BN ->
    ReLU -> channels-last LayerNorm(eps 1e-6) -> Linear C->4C -> GELU(erf) ->
    Linear 4C->C -> residual.  gamma is None (layer_scale_init_value=0, :28-36);
    DropPath is the identity; ``conv1`` is never used.  Token-major in/out."""
    B, N, C = feat0.shape
    dt = feat0.dtype
    w = p["dwconv2.weight"].to(dt)
    k = w.shape[-1]
    m = feat0.transpose(1, 2).reshape(B, C, H, W)
    y = F.conv2d(m, w, p["dwconv2.bias"].to(dt), padding=(k - 1) // 2, groups=C)
    y = torch.relu(_bn(y, p, "bn1", 1, bn_stats))
    y = y.reshape(B, C, N).transpose(1, 2)
    y = F.layer_norm(y, (C,), p["norm.weight"].to(dt), p["norm.bias"].to(dt), LKPM_LN_EPS)
    y = F.gelu(F.linear(y, p["pwconv1.weight"].to(dt), p["pwconv1.bias"].to(dt)))
    y = F.linear(y, p["pwconv2.weight"].to(dt), p["pwconv2.bias"].to(dt))
    return feat0 + y


def combine1(p: Mapping, feat0: Tensor, g, H: int, W: int, bn_stats: Optional[dict] = None) -> Tensor:
    """transformer.py:261-275: LKPM(DAPM(feat0))"""
    return lkpm(sub(p, "large_kernel_path."), dapm(sub(p, "transformer_path."), feat0, g, H, W, bn_stats=bn_stats), H, W,
                bn_stats)


# --------------------------------------------------------------------------
# a5: hist2image (D-to-image cross attention)
# --------------------------------------------------------------------------
def hist2image(p: Mapping, feat0: Tensor, emb: Tensor, zone_tokens: Tensor, mask: Tensor,
               g: Mapping[str, int], H: int, W: int, no_skip_inside: bool = False) -> Tensor:
    """Follows src/models/fusion.py:132-157.

    ``feat0`` [B,N,C] is updated (a new tensor is returned); ``emb`` [B,N,C] is
    the map the zone canvas is cut from (== feat0 when ``change_embedding``);
    ``zone_tokens`` [(B*Z),S,C] already carry positional_encodings2; ``mask``
    [B,Z] bool.  Canvas = zero-padded map cut to the zone rectangle, optionally
    resized (bilinear, align_corners) to [zn*p1, zn*p2], split into zone
    patches; each patch attends to its own zone's S tokens; rows of invalid
    zones are zeroed *including the residual*; resized back; the in-image part
    of the canvas is added onto the zone rectangle of feat0.
    """
    B, N, C = feat0.shape
    zn, p1, p2 = g["zone_num"], g["p1"], g["p2"]
    m = emb.transpose(1, 2).reshape(B, C, H, W)
    m = F.pad(m, (g["pad_w"], g["pad_w"], g["pad_h"], g["pad_h"]))
    canvas = m[:, :, g["sy"]:g["ey"], g["sx"]:g["ex"]]
    if canvas.shape[2] != g["tzh"] or canvas.shape[3] != g["tzw"]:
        raise ValueError("zone rectangle leaves the padded map (the reference fails here too)")
    if g["interpolate"]:
        canvas = F.interpolate(canvas, size=[zn * p1, zn * p2], mode="bilinear", align_corners=True)
    t = canvas.reshape(B, C, zn, p1, zn, p2).permute(0, 2, 4, 3, 5, 1).reshape(B * zn * zn, p1 * p2, C)
    t = loftr_layer(p, t, zone_tokens, 4)
    t = t * mask.reshape(B * zn * zn, 1, 1).to(t.dtype)
    t = t.view(B, zn, zn, p1, p2, C).permute(0, 5, 1, 3, 2, 4).reshape(B, C, zn * p1, zn * p2)
    if g["interpolate"]:
        t = F.interpolate(t, size=[g["tzh"], g["tzw"]], mode="bilinear", align_corners=True)
    # in-image part of the canvas <-> zone rectangle of the map (pad_mask / zone_mask)
    top, left = max(-g["sy_wo"], 0), max(-g["sx_wo"], 0)
    hh, ww = g["ry1"] - g["ry0"], g["rx1"] - g["rx0"]
    t = t[:, :, top:top + hh, left:left + ww].permute(0, 2, 3, 1)
    out = feat0.clone().view(B, H, W, C)
    if no_skip_inside:
        out[:, g["ry0"]:g["ry1"], g["rx0"]:g["rx1"]] = t
    else:
        out[:, g["ry0"]:g["ry1"], g["rx0"]:g["rx1"]] += t
    return out.view(B, N, C)


# --------------------------------------------------------------------------
# TransformerFusion
# --------------------------------------------------------------------------
def draw_posenc_offsets(max_res: Sequence[int], H: int, W: int) -> Tuple[int, int]:
    """fusion.py:87-91: the crop offset of the positional-encoding table is
    drawn from torch's default CPU generator, y first, and only for a dim that
    is smaller than the table."""
    oy = ox = 0
    if H < max_res[0]:
        oy = int(torch.randint(0, max_res[0] - H + 1, [1]))
    if W < max_res[1]:
        ox = int(torch.randint(0, max_res[1] - W + 1, [1]))
    return oy, ox


def transformer_fusion(sd: Mapping, layer_names: Sequence[str], max_res: Sequence[int],
                       x: Tensor, feat1: Tensor, mask: Tensor, patch_info: dict,
                       offsets: Optional[Tuple[int, int]] = None,
                       change_embedding: bool = True, no_skip_inside: bool = False,
                       bn_stats: Optional[dict] = None) -> Tensor:
    """Whole ``TransformerFusion.forward`` (src/models/fusion.py:52-188).

    ``x`` [B,C,H,W], ``feat1`` [B,Z,S,C], ``mask`` [B,Z] bool -> [B,C,H,W].
    ``offsets`` = (oy, ox) of the positional-encoding crop; drawn from the
    global CPU RNG exactly like the reference when None.  ``bn_stats``: None = eval
    mode; a dict = train mode (batch statistics, updated buffers recorded in it -
    see ``_bn``).  Gradients come from autograd over this restatement.
    """
    B, C, H, W = x.shape
    dt = x.dtype
    g = zone_geometry(patch_info, max_res[1], H, W)
    oy, ox = offsets if offsets is not None else draw_posenc_offsets(max_res, H, W)
    pos = sd["positional_encodings"].to(dt).view(max_res[0], max_res[1], C)[oy:oy + H, ox:ox + W]
    feat0 = (x.permute(0, 2, 3, 1) + pos).reshape(B, H * W, C)
    emb0 = feat0
    ztok = (feat1 + sd["positional_encodings2"].to(dt)).reshape(-1, feat1.shape[2], C)
    ws = twins_window_size(max_res)
    for i, name in enumerate(layer_names):
        p = sub(sd, f"layers.{i}.")
        if name == "image":
            feat0 = twins_layer(p, feat0, H, W, ws)
        elif name == "hist2image":
            feat0 = hist2image(p, feat0, feat0 if change_embedding else emb0, ztok, mask, g, H, W,
                               no_skip_inside)
        elif name == "combine1":
            feat0 = combine1(p, feat0, g, H, W, bn_stats)
        else:
            raise NotImplementedError(name)
    return feat0.view(B, H, W, C).permute(0, 3, 1, 2).contiguous()


# --------------------------------------------------------------------------
# whole hot path for one batch ("frame" = hist encoder + 3 fusion calls)
# --------------------------------------------------------------------------
LEVELS = (  # name, C, max_resolution, large_kernel  (decoder.py:82-94)
    ("cross_atten3", 128, (30, 40), 7),
    ("cross_atten2", 64, (60, 80), 15),
    ("cross_atten1", 32, (120, 160), 31),
)


def fusion_path(sds: Mapping[str, Mapping], layer_names, xs: Sequence[Tensor], hist: Tensor,
                mask: Tensor, patch_info: dict, offsets=None) -> List[Tensor]:
    """hist encoder + the three fusion calls in decoder order L3, L2, L1
    (decoder.py:111,116,121; deltar.py:40).  ``xs`` = decoder features at
    1/16, 1/8, 1/4; ``sds`` = {"hist_encoder": sd, "cross_atten3": sd, ...}."""
    f32, f64, f128 = hist_encoder(sds["hist_encoder"], hist.to(xs[0].dtype))
    toks = {128: f128, 64: f64, 32: f32}
    outs = []
    for li, (name, C, max_res, _) in enumerate(LEVELS):
        off = None if offsets is None else offsets[li]
        outs.append(transformer_fusion(sds[name], layer_names, max_res, xs[li], toks[C], mask,
                                       patch_info, off))
    return outs


# --------------------------------------------------------------------------
# the caller of the path: decoder shell + adaptive-bins head (test harness for the end-to-end depth check;
# SURVEY.md §8 f1/f2 - these ops are NOT part of the product, they wrap it the way the reference's Decoder / Deltar do)
# --------------------------------------------------------------------------
def _conv(p: Mapping, name: str, x: Tensor, padding: int) -> Tensor:
    b = p[name + ".bias"].to(x.dtype) if (name + ".bias") in p else None
    return F.conv2d(x, p[name + ".weight"].to(x.dtype), b, padding=padding)


def upsample_bn(p: Mapping, x: Tensor, concat_with: Tensor) -> Tensor:
    """``UpSampleBN.forward`` (src/models/decoder.py:40-58): bilinear(align_corners) resize to the skip map, concat,
    2 x [conv3x3 + BN(eval) + LeakyReLU(0.01)]."""
    up = F.interpolate(x, size=[concat_with.size(2), concat_with.size(3)], mode="bilinear", align_corners=True)
    f = torch.cat([up, concat_with], dim=1)
    f = F.leaky_relu(_bn_eval(_conv(p, "_net.0", f, 1), p, "_net.1", 1), 0.01)
    return F.leaky_relu(_bn_eval(_conv(p, "_net.3", f, 1), p, "_net.4", 1), 0.01)


def decoder_shell(p: Mapping, img_features: Sequence[Tensor], hist_features: Sequence[Tensor], fuse) -> Tensor:
    """``Decoder.forward`` (src/models/decoder.py:96-128).  ``fuse(name, x, depth_feat)`` runs one TransformerFusion
    call (``name`` in cross_atten3 / 2 / 1): the oracle's, or the CUDA module under test."""
    x0, x1, x2, x3, x4 = img_features
    f1, f2, f3 = hist_features
    d4 = _conv(p, "conv4", x4, 0)
    d3 = _conv(p, "conv3", upsample_bn(sub(p, "up1."), d4, x3), 0)
    d3 = torch.cat([d3, fuse("cross_atten3", d3, f3)], dim=1)
    d2 = _conv(p, "conv2", upsample_bn(sub(p, "up2."), d3, x2), 0)
    d2 = torch.cat([d2, fuse("cross_atten2", d2, f2)], dim=1)
    d1 = _conv(p, "conv1", upsample_bn(sub(p, "up3."), d2, x1), 0)
    d1 = torch.cat([d1, fuse("cross_atten1", d1, f1)], dim=1)
    d0 = upsample_bn(sub(p, "up4."), d1, x0)
    return _conv(p, "conv0", d0, 1)


def depth_tail(sd: Mapping, unet_out: Tensor, min_val: float, max_val: float) -> Tuple[Tensor, Tensor]:
    """``DepthRegression.forward`` with norm='linear' (decoder.py:21-37), the ``conv_out`` softmax over the bins and the
    bin-centre expectation (deltar.py:18-19, 50-61).  Returns (bin_edges [B, n_bins+1], pred [B,1,H,W])."""
    h = sub(sd, "depth_head.")
    ram = _conv(h, "conv3x3", unet_out, 1)
    r = F.conv2d(unet_out, h["conv1x1.weight"].to(unet_out.dtype)).mean([2, 3])
    r = F.leaky_relu(F.linear(r, h["regressor.0.weight"].to(r.dtype), h["regressor.0.bias"].to(r.dtype)), 0.01)
    r = F.leaky_relu(F.linear(r, h["regressor.2.weight"].to(r.dtype), h["regressor.2.bias"].to(r.dtype)), 0.01)
    y = torch.relu(F.linear(r, h["regressor.4.weight"].to(r.dtype), h["regressor.4.bias"].to(r.dtype))) + 0.1
    widths = (max_val - min_val) * (y / y.sum(dim=1, keepdim=True))
    widths = F.pad(widths, (1, 0), mode="constant", value=min_val)
    edges = torch.cumsum(widths, dim=1)
    centers = 0.5 * (edges[:, :-1] + edges[:, 1:])
    prob = torch.softmax(_conv(sub(sd, "conv_out."), "0", ram, 0), dim=1)
    pred = torch.sum(prob * centers.view(*centers.shape, 1, 1), dim=1, keepdim=True)
    return edges, pred


def abs_rel(pred: Tensor, gt: Tensor) -> float:
    """``compute_errors``' abs_rel (src/utils/metrics.py:10)."""
    return float(((pred.double() - gt.double()).abs() / gt.double()).mean())


# --------------------------------------------------------------------------
# f3 (input side, SURVEY.md section 8f - "next" row, oracle step only): zone distribution -> zone depth samples
# --------------------------------------------------------------------------
def sample_points_from_hist(hist_data: Tensor, mask: Tensor, zone_sample_num: int = 16, sample_uniform: bool = True) -> Tensor:
    """``hist_data`` [Z,2] (mu, sigma) per zone, ``mask`` [Z] bool -> [Z, zone_sample_num] float32 depth samples, zeros
    for invalid zones.  Follows src/utils/dataloader.py:65-81 (``sample_point_from_hist_parallel``): with
    ``sample_uniform`` an even grid over mu +- 3 sigma built as ``w * start + (1 - w) * end`` from two float32
    ``linspace`` ramps (``tensor_linspace``, :43-58 - the blend, not ``torch.linspace(start, end)``, fixes the rounding);
    otherwise the normal quantiles at ppf = arange(delta, 1, (1 - 2 delta) / (n - 1)), delta = 1e-3."""
    Z = mask.numel()
    fh = torch.zeros(Z, zone_sample_num, dtype=torch.float32)
    mu, sigma = hist_data[mask, 0], hist_data[mask, 1]
    if sample_uniform:
        start, end = mu - 3.0 * sigma, mu + 3.0 * sigma
        w0 = torch.linspace(1, 0, steps=zone_sample_num).to(start)
        w1 = torch.linspace(0, 1, steps=zone_sample_num).to(start)
        fh[mask] = (w0 * start.unsqueeze(-1) + w1 * end.unsqueeze(-1)).to(torch.float32)
    else:
        import numpy as np
        delta = 1e-3
        ppf = torch.Tensor(np.arange(delta, 1, (1 - 2 * delta) / (zone_sample_num - 1)).tolist()).unsqueeze(0)
        q = mu.unsqueeze(-1) + sigma.unsqueeze(-1) * math.sqrt(2.0) * torch.erfinv(2 * ppf - 1)      # Normal(mu, sigma).icdf
        fh[mask] = q.to(torch.float32)
    return fh


# --------------------------------------------------------------------------
# f4 (loss, SURVEY.md section 8f - "next" row, oracle step only)
# --------------------------------------------------------------------------
def silog_loss(pred: Tensor, target: Tensor, mask: Optional[Tensor] = None, interpolate: bool = True) -> Tensor:
    """Scale-invariant log loss of the training step (src/loss.py:9-19): the prediction is resized to the target
    (bilinear, align_corners), masked, g = log(pred) - log(target), 10 * sqrt(var(g) + 0.15 * mean(g)^2) with the
    UNBIASED variance (torch.var default)."""
    if interpolate:
        pred = F.interpolate(pred, target.shape[-2:], mode="bilinear", align_corners=True)
    if mask is not None:
        pred, target = pred[mask], target[mask]
    g = torch.log(pred) - torch.log(target)
    return 10 * torch.sqrt(torch.var(g) + 0.15 * torch.pow(torch.mean(g), 2))


# --------------------------------------------------------------------------
# f3 (input side): depth map -> per-zone (mu, sigma, validity)
# --------------------------------------------------------------------------
def zone_hist_params(dep: Tensor, sy: int, sx: int, ph: int, pw: int, zone_num: int, max_distance: float):
    """One frame of ``get_hist_parallel`` (src/utils/dataloader.py:84-134) with the random draws already made:
    ``dep`` [H,W] metres; zones of ph x pw px from (sy, sx).  Returns (fh [Z,2] float64 (mu, sigma), mask [Z] bool,
    hist [Z,bins] float32 - the surviving cluster's counts).  :103-106 histogram per zone (torch.histc, 4 cm bins),
    :110-111 bin 0 cleared and 20 subtracted, :112-118 strongest contiguous cluster, :120 bin centres through a float32
    tensor of the upper edges, :126-131 moments in float64."""
    import numpy as np
    range_margin = list(np.arange(0, max_distance + 1e-9, 0.04))
    bins = int(max_distance / 0.04)
    Z = zone_num * zone_num
    hist = torch.zeros(Z, bins, dtype=torch.float32)
    for z in range(Z):
        zy, zx = divmod(z, zone_num)
        patch = dep[sy + zy * ph: sy + (zy + 1) * ph, sx + zx * pw: sx + (zx + 1) * pw].contiguous().float()
        hist[z] = torch.histc(patch, bins=bins, min=0, max=max_distance)
    hist[:, 0] = 0
    hist = torch.clip(hist - 20, 0, None)
    for z in range(Z):
        row = hist[z].clone()
        best, lo, hi, i = -1.0, 0, 0, 0
        while i < bins:
            if row[i] == 0:
                i += 1
                continue
            j = i
            while j < bins and row[j] != 0:
                j += 1
            s = float(row[i:j].sum())
            if s > best:
                best, lo, hi = s, i, j
            i = j
        hist[z] = 0
        hist[z, lo:hi] = row[lo:hi]
    dist = ((torch.Tensor(range_margin[1:]).double() + torch.tensor(range_margin[:-1], dtype=torch.float64)) / 2).unsqueeze(0)
    n = torch.sum(hist, dim=1)
    mask = n > 0
    mu = torch.sum(dist * hist, dim=1) / (n + 1e-9)
    std = torch.sqrt(torch.sum(hist * torch.pow(dist - mu.unsqueeze(-1), 2), dim=1) / (n + 1e-9)) + 1e-9
    return torch.stack([mu, std], dim=1), mask, hist


# --------------------------------------------------------------------------
# f4 (metrics)
# --------------------------------------------------------------------------
def depth_metrics(gt: Tensor, pred: Tensor) -> dict:
    """``compute_errors`` (src/utils/metrics.py:4-24) on 1-d tensors of valid pixels, float64."""
    gt, pred = gt.double(), pred.double()
    thresh = torch.maximum(gt / pred, pred / gt)
    err = torch.log(pred) - torch.log(gt)
    return dict(a1=float((thresh < 1.25).double().mean()), a2=float((thresh < 1.25 ** 2).double().mean()),
                a3=float((thresh < 1.25 ** 3).double().mean()), abs_rel=float(((gt - pred).abs() / gt).mean()),
                rmse=float(((gt - pred) ** 2).mean().sqrt()), log_10=float((torch.log10(gt) - torch.log10(pred)).abs().mean()),
                rmse_log=float(((torch.log(gt) - torch.log(pred)) ** 2).mean().sqrt()),
                silog=float(((err ** 2).mean() - err.mean() ** 2).sqrt() * 100), sq_rel=float((((gt - pred) ** 2) / gt).mean()))
