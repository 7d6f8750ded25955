"""Host-only checks of the engine's tile grid and crop-atlas packing (csrc/engine.cu::layout_tiles) against
RealESRGANer.tile_process as restated in oracle/rrdbnet.py: every tile appears exactly once with the
oracle's crop box, paste box and offsets; inside an atlas image the crops do not touch (a zero gap column at the coarsest
conv resolution between neighbours), fit the group's canvas, and an image holds at most the mask table's rectangle count."""
import math
import os

import pytest

from ss4k_b200 import _lib as L
from ss4k_b200 import engine as E


def oracle_tiles(h, w, scale, tile, pad, pre_pad, x2):
    """(src_y, src_x, crop_h, crop_w, off_y, off_x, dst_y, dst_x, paste_h, paste_w) of every tile, RealESRGANer order"""
    hp, wp = h + pre_pad, w + pre_pad
    if x2:
        hp += hp % 2
        wp += wp % 2
    out = []
    for y in range(math.ceil(hp / tile)):
        for x in range(math.ceil(wp / tile)):
            sx, sy = x * tile, y * tile
            ex, ey = min(sx + tile, wp), min(sy + tile, hp)
            sxp, exp_ = max(sx - pad, 0), min(ex + pad, wp)
            syp, eyp = max(sy - pad, 0), min(ey + pad, hp)
            out.append((syp, sxp, eyp - syp, exp_ - sxp, (sy - syp) * scale, (sx - sxp) * scale, sy * scale, sx * scale,
                        (ey - sy) * scale, (ex - sx) * scale))
    return sorted(out)


CASES = [
    # arch, scale, h, w, tile, pad, pre_pad
    (L.ARCH_RRDB, 2, 1080, 1920, 512, 10, 0),     # BASELINE.json configs[3]: nine shape classes -> two atlases
    (L.ARCH_RRDB, 2, 1080, 1920, 492, 10, 0),
    (L.ARCH_RRDB, 4, 1080, 1920, 256, 10, 0),     # 40 tiles: the interior class is cut into several atlas images
    (L.ARCH_RRDB, 4, 1080, 1920, 236, 10, 0),
    (L.ARCH_RRDB, 2, 45, 71, 32, 6, 5),           # odd frame: x2 mod-2 pad on top of pre_pad
    (L.ARCH_SRVGG, 4, 98, 150, 40, 6, 3),
    (L.ARCH_RRDB, 2, 98, 150, 40, 7, 0),          # odd tile_pad: odd crops -> no atlases for the x2 net (one image per crop)
]


@pytest.mark.parametrize("arch,scale,h,w,tile,pad,pre_pad", CASES)
def test_tile_grid_and_atlases(arch, scale, h, w, tile, pad, pre_pad):
    cfg = E.make_cfg(0, arch, 2, h, w, scale=scale, tile=tile, tile_pad=pad)
    cfg.reserved[1] = pre_pad
    lay = E.tile_layout(cfg)
    x2 = arch == L.ARCH_RRDB and scale == 2
    tdiv = 2 if x2 else 1
    got = []
    for g in lay["groups"]:
        per_img = {}
        for b in g["boxes"]:
            got.append((b["src_y"], b["src_x"], b["crop_h"], b["crop_w"], b["off_y"], b["off_x"], b["dst_y"], b["dst_x"],
                        b["paste_h"], b["paste_w"]))
            assert 0 <= b["img"] < g["nimg"]
            assert b["crop_h"] <= g["hc"] and b["atlas_x"] + b["crop_w"] <= g["wc"]
            assert b["crop_h"] * 10 >= g["hc"] * 9 or len(g["boxes"]) == g["nimg"]      # similar heights share a canvas
            per_img.setdefault(b["img"], []).append((b["atlas_x"], b["crop_w"]))
        for spans in per_img.values():
            assert len(spans) <= lay["max_rects"]
            spans.sort()
            for (x0, w0), (x1, _) in zip(spans, spans[1:]):
                assert x1 >= x0 + w0 + tdiv and x1 % tdiv == 0            # a zero gap column at the coarsest resolution
    assert sorted(got) == oracle_tiles(h, w, scale, tile, pad, pre_pad, x2)
    odd = any(b["crop_h"] % tdiv or b["crop_w"] % tdiv for g in lay["groups"] for b in g["boxes"])
    if odd:
        assert all(len(g["boxes"]) == g["nimg"] for g in lay["groups"])   # exact classes
    n_tiles = len(got)
    n_img = sum(g["nimg"] for g in lay["groups"])
    if not odd and n_tiles > 1:
        assert n_img < n_tiles                                             # crops really share images


def test_cfg4_is_two_atlases():
    lay = E.tile_layout(E.make_cfg(0, L.ARCH_RRDB, 1, 1080, 1920, scale=2, tile=512, tile_pad=10))
    shapes = sorted((g["hc"], g["wc"], g["nimg"], len(g["boxes"])) for g in lay["groups"])
    assert shapes == [(66, 2 * 532 + 522 + 394 + 3 * 2, 1, 4), (532, 4 * 532 + 2 * 522 + 2 * 394 + 7 * 2, 1, 8)]


def test_exact_classes_switch(monkeypatch):
    monkeypatch.setenv("SS4K_TILE_EXACT_CLASSES", "1")
    lay = E.tile_layout(E.make_cfg(0, L.ARCH_RRDB, 1, 1080, 1920, scale=2, tile=512, tile_pad=10))
    assert len(lay["groups"]) == 9 and all(len(g["boxes"]) == g["nimg"] and all(b["atlas_x"] == 0 for b in g["boxes"]) for g in lay["groups"])
