"""The row providers' index arithmetic (csrc/providers.cuh, canvas_resize_add_kernel in k_loftr.cu) restated in numpy
and checked on the CPU against the reference's own tensor regrouping (pad / slice / interpolate / reshape / permute /
boolean indexing, fusion.py:132-157, transformer.py:94-116,215-234) on EVERY zone layout of
tests/golden/geometry_cases.json that the product serves (8x8 and 6x6 grids, pad and resize branches, three levels).
The GPU parity tests run the compiled code on 11 layouts; this pins the formulas themselves on ~190."""
import json
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from cfpnet_b200 import geometry, synth
from test_geometry_golden import CASES, grid_rects


def served_layouts():
    out = []
    for i, c in enumerate(CASES):
        if "geo" not in c:
            continue
        _, stride, max_res, _ = synth.LEVELS[c["level"]]
        H, W = c["img"][0] // stride, c["img"][1] // stride
        g = geometry.zone_geometry(geometry.collate_patch_info([geometry.patch_info_from_rect_data(grid_rects(c))]),
                                   max_res[1], H, W)
        try:
            geometry.check_geometry(g, H, W)
        except ValueError:
            continue
        out.append((i, g, H, W, max_res))
    return out


LAYOUTS = served_layouts()


def zone_patch_rows(g):
    """ZonePatchRows::locate for one frame: canvas cell (cy, cx) of every dense row, rows grouped per zone."""
    Z, P = g.zone_num ** 2, g.p1 * g.p2
    r = np.arange(Z * P)
    grp, l = r // P, r % P
    z = grp % Z
    zy, zx = z // g.zone_num, z % g.zone_num
    py, px = l // g.p2, l % g.p2
    return zy * g.p1 + py, zx * g.p2 + px


def reference_regroup(t, g):
    """fusion.py:136-142 on a [1,1,H,W] map -> [(Z), p1*p2]"""
    zn, p1, p2 = g.zone_num, g.p1, g.p2
    m = F.pad(t, (g.pad_w, g.pad_w, g.pad_h, g.pad_h))[:, :, g.sy:g.ey, g.sx:g.ex]
    if g.interpolate:
        m = F.interpolate(m, size=[zn * p1, zn * p2], mode="bilinear", align_corners=True)
    return m.reshape(1, 1, zn, p1, zn, p2).permute(0, 2, 4, 3, 5, 1).reshape(zn * zn, p1 * p2)


def test_enough_layouts():
    assert len(LAYOUTS) >= 150
    assert {g.zone_num for _, g, *_ in LAYOUTS} == {6, 8}
    assert {bool(g.interpolate) for _, g, *_ in LAYOUTS} == {False, True}
    assert any(g.pad_h > 0 or g.pad_w > 0 for _, g, *_ in LAYOUTS)


@pytest.mark.parametrize("k", range(len(LAYOUTS)))
def test_hist2image_gather_and_scatter(k):
    _, g, H, W, _ = LAYOUTS[k]
    Z, P = g.zone_num ** 2, g.p1 * g.p2
    cy, cx = zone_patch_rows(g)
    rect = np.zeros((H, W), dtype=bool)
    rect[g.ry0:g.ry1, g.rx0:g.rx1] = True
    if not g.interpolate:
        # gather: canvas_at(sy_wo + cy, sx_wo + cx), zero outside the map - against the reference's regroup of an index map
        y, x = g.sy_wo + cy, g.sx_wo + cx
        inimg = (y >= 0) & (y < H) & (x >= 0) & (x < W)
        tok = np.where(inimg, y * W + x, -1)
        idx = (torch.arange(H * W, dtype=torch.float64) + 1).view(1, 1, H, W)
        want = reference_regroup(idx, g).numpy().astype(np.int64) - 1
        assert np.array_equal(tok.reshape(Z, P), want)
        # scatter: every cell of the zone rectangle is written exactly once, nothing else is (pad_mask / zone_mask)
        hit = tok[tok >= 0]
        assert len(np.unique(hit)) == len(hit) and np.array_equal(np.sort(hit), np.flatnonzero(rect.reshape(-1)))
        return
    # resize branch, gather: ZonePatchRows::load4's bilinear blend (align_corners) in float32
    gen = torch.Generator().manual_seed(k)
    m = torch.randn(1, 1, H, W, generator=gen)
    want = reference_regroup(m, g).numpy()
    oh, ow = g.zone_num * g.p1, g.zone_num * g.p2
    mp = m[0, 0].numpy()

    def canvas_at(ty, tx):
        y, x = g.sy_wo + ty, g.sx_wo + tx
        ok = (y >= 0) & (y < H) & (x >= 0) & (x < W)
        return np.where(ok, mp[np.clip(y, 0, H - 1), np.clip(x, 0, W - 1)], np.float32(0))

    fy = (cy.astype(np.float32) * np.float32((g.tzh - 1) / (oh - 1))) if oh > 1 else np.zeros_like(cy, dtype=np.float32)
    fx = (cx.astype(np.float32) * np.float32((g.tzw - 1) / (ow - 1))) if ow > 1 else np.zeros_like(cx, dtype=np.float32)
    y0, x0 = fy.astype(np.int64), fx.astype(np.int64)
    y1, x1 = np.minimum(y0 + 1, g.tzh - 1), np.minimum(x0 + 1, g.tzw - 1)
    ly, lx = fy - y0, fx - x0
    got = ((1 - ly) * (1 - lx) * canvas_at(y0, x0) + (1 - ly) * lx * canvas_at(y0, x1)
           + ly * (1 - lx) * canvas_at(y1, x0) + ly * lx * canvas_at(y1, x1))
    assert np.abs(got.reshape(Z, P) - want).max() <= 2e-5 * max(1.0, np.abs(want).max())
    # resize branch, scatter: canvas_resize_add_kernel - resize [oh, ow] back to [tzh, tzw], in-image part onto the rectangle
    out = torch.randn(1, 1, oh, ow, generator=gen)
    back = F.interpolate(out, size=[g.tzh, g.tzw], mode="bilinear", align_corners=True)[0, 0].numpy()
    top, left = max(-g.sy_wo, 0), max(-g.sx_wo, 0)
    hh, ww = g.ry1 - g.ry0, g.rx1 - g.rx0
    want2 = back[top:top + hh, left:left + ww]
    yy, xx = np.meshgrid(np.arange(g.ry0, g.ry1), np.arange(g.rx0, g.rx1), indexing="ij")
    ty, tx = yy - g.sy_wo, xx - g.sx_wo
    assert ty.min() == top and tx.min() == left
    fy = ty.astype(np.float32) * np.float32((oh - 1) / (g.tzh - 1)) if g.tzh > 1 else np.zeros_like(ty, dtype=np.float32)
    fx = tx.astype(np.float32) * np.float32((ow - 1) / (g.tzw - 1)) if g.tzw > 1 else np.zeros_like(tx, dtype=np.float32)
    y0, x0 = fy.astype(np.int64), fx.astype(np.int64)
    y1, x1 = np.minimum(y0 + 1, oh - 1), np.minimum(x0 + 1, ow - 1)
    ly, lx = fy - y0, fx - x0
    o = out[0, 0].numpy()
    got2 = (1 - ly) * (1 - lx) * o[y0, x0] + (1 - ly) * lx * o[y0, x1] + ly * (1 - lx) * o[y1, x0] + ly * lx * o[y1, x1]
    assert np.abs(got2 - want2).max() <= 2e-5 * max(1.0, np.abs(want2).max())


@pytest.mark.parametrize("k", range(len(LAYOUTS)))
def test_dapm_inside_outside_rows(k):
    """InsideSrc / OutsideRows::locate enumerate the cells in the raster order of feat0[zone] / feat0[~zone]"""
    _, g, H, W, _ = LAYOUTS[k]
    rect = np.zeros((H, W), dtype=bool)
    rect[g.ry0:g.ry1, g.rx0:g.rx1] = True
    flat = rect.reshape(-1)
    rw = g.rx1 - g.rx0
    Ni = (g.ry1 - g.ry0) * rw
    i = np.arange(Ni)
    assert np.array_equal((g.ry0 + i // max(rw, 1)) * W + g.rx0 + i % max(rw, 1), np.flatnonzero(flat))
    No = H * W - Ni
    o = np.arange(No)
    top, mid, d_out = g.ry0 * W, (g.ry1 - g.ry0) * (W - rw), max(W - rw, 1)
    row, j = (o - top) // d_out, (o - top) % d_out
    n = np.where(o < top, o, np.where(o < top + mid, (g.ry0 + row) * W + np.where(j < g.rx0, j, j + rw),
                                     g.ry1 * W + (o - top - mid)))
    assert np.array_equal(n, np.flatnonzero(~flat))


@pytest.mark.parametrize("level", [1, 2, 3])
@pytest.mark.parametrize("geom", ["G416", "G480"])
def test_lsa_window_rows(level, geom):
    """WindowRows::locate against the reference's window regroup of the zero-padded map (transformer.py:94-104)"""
    _, _, max_res, _ = synth.LEVELS[level]
    H, W = synth.level_hw(geom, level)
    ws = math.ceil(math.sqrt(math.sqrt(max_res[0] * max_res[1])))
    nwy, nwx = -(-H // ws), -(-W // ws)
    r = np.arange(nwy * nwx * ws * ws)
    grp, l = r // (ws * ws), r % (ws * ws)
    wy, wx, iy, ix = grp // nwx, grp % nwx, l // ws, l % ws
    y, x = wy * ws + iy, wx * ws + ix
    tok = np.where((y < H) & (x < W), y * W + x, -1)
    idx = (torch.arange(H * W, dtype=torch.float64) + 1).view(1, H, W, 1)
    pb, pr = (ws - H % ws) % ws, (ws - W % ws) % ws
    t = F.pad(idx, (0, 0, 0, pr, 0, pb))
    want = t.view(1, nwy, ws, nwx, ws, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1).numpy().astype(np.int64) - 1
    assert np.array_equal(tok, want)


@pytest.mark.parametrize("H,W", [(26, 34), (52, 68), (104, 136), (30, 40), (120, 160), (7, 5)])
@pytest.mark.parametrize("T", [1, 2, 4])
def test_conv3x3_padded_raster_scheme(H, W, T):
    """csrc/k_conv_tc.cu restated: a CTA stages the zero-padded raster of R rows (+halo) with one slack cell in front,
    output cell o of the SAME raster reads tap (dy, dx) at staged cell o + dy*WP + dx, junk columns / rows are dropped
    (`live`).  Checked against F.conv2d (float64, 3 channels) for every CTA of the grid, incl. the staging bound."""
    Cc = 3
    gen = torch.Generator().manual_seed(H * 1000 + W + T)
    x = torch.randn(1, Cc, H, W, generator=gen, dtype=torch.float64)
    w = torch.randn(Cc, Cc, 3, 3, generator=gen, dtype=torch.float64)
    zy0, zy1, zx0, zx1 = H // 4, H // 2 + 1, W // 3, W // 2          # second source: cells inside the zone are never read
    want = F.conv2d(x, w, padding=1)[0].numpy()
    xz = x.clone()
    xz[:, :, zy0:zy1, zx0:zx1] = 0
    want_z = F.conv2d(xz, w, padding=1)[0].numpy()
    WP = W + 2
    R = (T * 128) // WP
    if R < 1:
        pytest.skip("map wider than the CTA's raster: the launcher refuses it (CFP_REQUIRE R >= 1)")
    cells = T * 128 + 2 * WP + 2
    while cells % 8 != 1:
        cells += 1
    xn, wn = x[0].numpy(), w.numpy()
    got = np.full((2, Cc, H, W), np.nan)
    for y0 in range(0, H, R):
        for src in range(2):
            ci = np.arange(cells)
            idx = ci - 1
            pr, px = idx // WP, idx % WP
            y, xx = y0 - 1 + pr, px - 1
            ok = (idx >= 0) & (pr < R + 2) & (y >= 0) & (y < H) & (xx >= 0) & (xx < W)
            if src == 1:
                ok &= ~((y >= zy0) & (y < zy1) & (xx >= zx0) & (xx < zx1))
            a = np.where(ok[None, :], xn[:, np.clip(y, 0, H - 1), np.clip(xx, 0, W - 1)], 0.0)       # [Cc, cells]
            o = np.arange(T * 128)
            acc = np.zeros((Cc, T * 128))
            for tap in range(9):
                at = o + (tap // 3) * WP + tap % 3
                assert at.max() < cells                                    # staged buffer covers every operand view
                acc += wn[:, :, tap // 3, tap % 3] @ a[:, at]
            r, pxo = o // WP, o % WP
            yy, xo = y0 + r, pxo - 1
            live = (r < R) & (yy < H) & (xo >= 0) & (xo < W)
            got[src][:, yy[live], xo[live]] = acc[:, live]
    assert not np.isnan(got).any()                                         # every output cell written by exactly one CTA row
    assert np.abs(got[0] - want).max() <= 1e-10 and np.abs(got[1] - want_z).max() <= 1e-10


@pytest.mark.parametrize("S,groups", [(16, 72), (16, 36), (36, 60), (81, 48), (144, 108), (20, 3), (35, 2), (88, 5),
                                      (576, 2), (2304, 1), (9216, 2), (1, 7)])
def test_kv_state_padded_runs(S, groups):
    """csrc/k_chain_tc.cu::kv_state_tc_kernel's row enumeration restated: group sizes padded to a multiple of 16 so that
    every 16-row MMA step belongs to one group; consecutive steps of a group inside a 128-row tile form a run that is
    flushed once - with a plain store when groups never straddle tiles (128 % S_pad == 0: no memset), with an atomic add
    otherwise.  The emulation must reproduce sum_s K_s^T V_s per group (float64) and, in the plain-store case, write
    every group exactly once."""
    d = 4
    gen = np.random.default_rng(S * 131 + groups)
    K, V = gen.standard_normal((groups, S, d)), gen.standard_normal((groups, S, d))
    want = np.einsum("gsd,gsv->gdv", K, V)
    S_pad = (S + 15) // 16 * 16
    complete = 128 % S_pad == 0
    total = groups * S_pad
    ntiles = (total + 127) // 128
    kv = np.zeros((groups, d, d)) if not complete else np.full((groups, d, d), np.nan)
    writes = np.zeros(groups, dtype=int)
    for tile in range(ntiles):
        row0 = tile * 128
        p = row0 + np.arange(128)
        g_of, s_of = p // S_pad, p % S_pad
        real = (p < total) & (s_of < S)
        Kt = np.where(real[:, None], K[np.minimum(g_of, groups - 1), np.minimum(s_of, S - 1)], 0.0)   # padding rows are zeros
        Vt = np.where(real[:, None], V[np.minimum(g_of, groups - 1), np.minimum(s_of, S - 1)], 0.0)
        ks = 0
        while ks < 8 and row0 + 16 * ks < total:
            g = (row0 + 16 * ks) // S_pad
            ke = ks + 1
            while ke < 8 and row0 + 16 * ke < total and (row0 + 16 * ke) // S_pad == g:
                ke += 1
            rows = slice(16 * ks, 16 * ke)
            assert set(g_of[rows][real[rows]]) <= {g}                       # a run never mixes groups
            acc = Kt[rows].T @ Vt[rows]
            if complete:
                kv[g] = acc
            else:
                kv[g] += acc
            writes[g] += 1
            ks = ke
    assert np.abs(kv - want).max() <= 1e-9 * max(1.0, np.abs(want).max())
    if complete:
        assert (writes == 1).all()


def test_query_chain_state_slots_cover_every_tile():
    """run_query_tc sizes the shared-memory copy of the attention state for kv_slots = (128 + rpg - 2) / rpg + 1 groups;
    a 128-row tile starting at any multiple of 128 must never touch more (the bulk copies would overrun the buffer)."""
    for rpg in list(range(1, 300)) + [576, 884, 3536, 14144]:
        kv_slots = (128 + rpg - 2) // rpg + 1
        rows = rpg * 37
        row0 = np.arange(0, rows, 128)
        last = np.minimum(row0 + 127, rows - 1)
        ng = last // rpg - row0 // rpg + 1
        assert ng.max() <= kv_slots, (rpg, int(ng.max()), kv_slots)
