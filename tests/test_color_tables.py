"""Colour-conversion tables: the generated header and the golden file agree, and --
where Pillow is importable -- the tables reproduce Pillow's RGB <-> YCbCr conversion (CPU only, no GPU needed)."""
import hashlib
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "color_tables.json")
HEADER = os.path.join(ROOT, "implementing-jpeg-compression_b200", "csrc", "jb_color_tables.h")


def _golden():
    with open(GOLDEN) as fh:
        return json.load(fh)


def _header_tables():
    text = open(HEADER).read()
    out = {}
    for name, body in re.findall(r"JB_COLOR_TABLE_QUAL int16_t (\w+)\[\d+\] = \{([^}]*)\}", text):
        out[name] = [int(v) for v in body.replace("\n", " ").split(",") if v.strip()]
    return out


def test_header_matches_golden_tables():
    g, h = _golden(), _header_tables()
    for n in ("y", "cb", "cr"):
        assert h["JB_RGB2YCC_" + n.upper()] == [v for t in g["forward"][n] for v in t]
    for n in ("r_cr", "g_cb", "g_cr", "b_cb"):
        assert h["JB_YCC2RGB_" + n.upper()] == g["inverse"][n]
    assert all(-32768 <= v < 32768 for t in h.values() for v in t)


def _apply_forward(g, rgb):
    r, gg, b = (rgb[..., k].astype(np.int64) for k in range(3))
    f = {k: [np.array(t) for t in v] for k, v in g["forward"].items()}
    return np.stack([(f[n][0][r] + f[n][1][gg] + f[n][2][b]) >> 6 for n in ("y", "cb", "cr")], axis=-1).astype(np.uint8)


def _apply_inverse(g, ycc):
    y, cb, cr = (ycc[..., k].astype(np.int64) for k in range(3))
    t = {k: np.array(v) for k, v in g["inverse"].items()}
    return np.stack([np.clip(y + (t["r_cr"][cr] >> 6), 0, 255), np.clip(y + ((t["g_cb"][cb] + t["g_cr"][cr]) >> 6), 0, 255),
                     np.clip(y + (t["b_cb"][cb] >> 6), 0, 255)], axis=-1).astype(np.uint8)


def all_colours():
    i = np.arange(256, dtype=np.uint8)
    a, b, c = np.meshgrid(i, i, i, indexing="ij")
    return np.stack([a.ravel(), b.ravel(), c.ravel()], axis=1).reshape(4096, 4096, 3)


def test_tables_reproduce_the_golden_hashes_on_all_inputs():
    g = _golden()
    src = all_colours()
    assert hashlib.sha256(_apply_forward(g, src).tobytes()).hexdigest() == g["sha256_rgb_to_ycbcr_all_2^24"]
    assert hashlib.sha256(_apply_inverse(g, src).tobytes()).hexdigest() == g["sha256_ycbcr_to_rgb_all_2^24"]


def test_tables_against_live_pillow():
    Image = pytest.importorskip("PIL.Image")
    g = _golden()
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (211, 173, 3), dtype=np.uint8)
    assert np.array_equal(_apply_forward(g, img), np.asarray(Image.fromarray(img, "RGB").convert("YCbCr")))
    assert np.array_equal(_apply_inverse(g, img), np.asarray(Image.fromarray(img, "YCbCr").convert("RGB")))
