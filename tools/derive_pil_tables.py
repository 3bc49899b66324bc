"""Derive integer tables that reproduce Pillow's RGB <-> YCbCr conversion exactly (SURVEY.md section 8(f) row 1;
reference call sites compress.py:9, decompress.py:10, pipeline/__init__.py:103,120-124).

Pillow converts with per-channel lookup tables and a 6-bit fixed point:  out = (T0[c0] + T1[c1] + T2[c2]) >> 6  (+ offsets,
then clipping for the inverse).  The tables here are fitted against Pillow itself on all 2^24 inputs (run in the dev
container, where Pillow is installed) and are functionally identical to Pillow's on every input; the script writes
  implementing-jpeg-compression_b200/csrc/jb_color_tables.h   (int16 tables for the CUDA kernels)
  tests/golden/color_tables.json                              (the same tables + SHA-256 of both full mappings)
"""
import hashlib
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
I = np.arange(256)


def all_triples():
    a, b, c = np.meshgrid(I.astype(np.uint8), I.astype(np.uint8), I.astype(np.uint8), indexing="ij")
    return np.stack([a.ravel(), b.ravel(), c.ravel()], axis=1).reshape(4096, 4096, 3)


def descend(tables, target, shift=6, sweeps=8):
    """Coordinate descent on three additive tables until (T0[a] + T1[b] + T2[c]) >> shift == target everywhere."""
    for _ in range(sweeps):
        v = (tables[0][:, None, None] + tables[1][None, :, None] + tables[2][None, None, :]) >> shift
        if not (v != target).any():
            return True
        for ax in range(3):
            for k in range(256):
                best = None
                for d in range(-2, 3):
                    cand = tables[ax][k] + d
                    if ax == 0:
                        e = (((cand + tables[1][:, None] + tables[2][None, :]) >> shift) != target[k]).sum()
                    elif ax == 1:
                        e = (((tables[0][:, None] + cand + tables[2][None, :]) >> shift) != target[:, k]).sum()
                    else:
                        e = (((tables[0][:, None] + tables[1][None, :] + cand) >> shift) != target[:, :, k]).sum()
                    if best is None or e < best[0] or (e == best[0] and abs(d) < abs(best[1])):
                        best = (e, d)
                tables[ax][k] += best[1]
    return False


def main():
    src = all_triples()
    ycc = np.asarray(Image.fromarray(src, "RGB").convert("YCbCr")).reshape(256, 256, 256, 3).astype(np.int64)
    rgb = np.asarray(Image.fromarray(src, "YCbCr").convert("RGB")).reshape(256, 256, 256, 3).astype(np.int64)
    rnd = lambda c: np.floor(c * 64 * I + 0.5).astype(np.int64)
    fwd = {}
    for name, ch, coefs, off in (("y", 0, (0.299, 0.587, 0.114), 0), ("cb", 1, (-0.16874, -0.33126, 0.5), 128),
                                 ("cr", 2, (0.5, -0.41869, -0.08131), 128)):
        t = [rnd(c) for c in coefs]
        if off:
            t[2] = t[2] + 2
        if not descend(t, ycc[..., ch] - off):
            sys.exit("forward fit failed for " + name)
        t[2] = t[2] + (off << 6)                      # fold the +128 in: out = (T0 + T1 + T2) >> 6
        fwd[name] = [x.tolist() for x in t]
    # inverse: r = clip(y + (R_CR[cr] >> 6)), b = clip(y + (B_CB[cb] >> 6)), g = clip(y + ((G_CB[cb] + G_CR[cr]) >> 6));
    # the shifted terms are observable directly wherever nothing clips
    def delta(channel, axis):
        d = np.zeros(256, dtype=np.int64)
        for v in range(256):
            for y in (128, 40, 215, 10, 245):
                idx = [y, 128, 128]
                idx[axis] = v
                out = rgb[idx[0], idx[1], idx[2], channel]
                if 0 < out < 255:
                    d[v] = out - y
                    break
            else:
                sys.exit("inverse probe failed")
        return d
    d_r, d_b = delta(0, 2), delta(2, 1)
    e = np.zeros((256, 256), dtype=np.int64)          # e[cb][cr] = (G_CB[cb] + G_CR[cr]) >> 6
    for y in (128, 60, 200, 20, 235, 5, 250):
        g = rgb[y, :, :, 1]
        ok = (g > 0) & (g < 255)
        e[ok] = (g - y)[ok]
    # e = (G_CB[cb] + G_CR[cr]) >> 6 determines the two tables only up to what the floor can see; any pair that
    # reproduces e is as good as Pillow's.  Row sums order the rows by (value mod 64): try every rotation of the
    # distinct levels as "residue 0", then the column table follows from interval intersection.
    s_row = e.sum(1)
    t_row = s_row - s_row.min()
    levels = np.unique(t_row % 256)
    g_cb = g_cr = None
    for rot in range(len(levels)):
        order = np.roll(levels, -rot)
        rank = {int(l): k for k, l in enumerate(order)}
        u = np.array([rank[int(t % 256)] for t in t_row])
        n_of = np.array([(int(t % 256) - int(order[0])) % 256 for t in t_row])
        x = 64 * ((t_row - n_of) // 256) + u
        lo = (64 * e - x[:, None]).max(0)
        hi = (64 * e - x[:, None] + 63).min(0)
        if (lo <= hi).all() and np.array_equal((x[:, None] + lo[None, :]) >> 6, e):
            k = x[128]
            g_cb, g_cr = x - k, lo + k
            break
    if g_cb is None:
        sys.exit("inverse green fit failed")
    inv = {"r_cr": (d_r << 6).tolist(), "b_cb": (d_b << 6).tolist(), "g_cb": g_cb.tolist(), "g_cr": g_cr.tolist()}
    # verify both directions on every input with the tables alone
    R, G, B = src[..., 0].astype(np.int64), src[..., 1].astype(np.int64), src[..., 2].astype(np.int64)
    f = {k: [np.array(x) for x in v] for k, v in fwd.items()}
    out = np.stack([(f[n][0][R] + f[n][1][G] + f[n][2][B]) >> 6 for n in ("y", "cb", "cr")], axis=-1).astype(np.uint8)
    assert np.array_equal(out.reshape(256, 256, 256, 3), ycc.astype(np.uint8)), "forward tables do not reproduce Pillow"
    Y, Cb, Cr = R, G, B
    iv = {k: np.array(v) for k, v in inv.items()}
    back = np.stack([np.clip(Y + (iv["r_cr"][Cr] >> 6), 0, 255), np.clip(Y + ((iv["g_cb"][Cb] + iv["g_cr"][Cr]) >> 6), 0, 255),
                     np.clip(Y + (iv["b_cb"][Cb] >> 6), 0, 255)], axis=-1).astype(np.uint8)
    assert np.array_equal(back.reshape(256, 256, 256, 3), rgb.astype(np.uint8)), "inverse tables do not reproduce Pillow"
    golden = {"pillow_version": Image.__version__ if hasattr(Image, "__version__") else "", "forward": fwd, "inverse": inv,
              "sha256_rgb_to_ycbcr_all_2^24": hashlib.sha256(ycc.astype(np.uint8).tobytes()).hexdigest(),
              "sha256_ycbcr_to_rgb_all_2^24": hashlib.sha256(rgb.astype(np.uint8).tobytes()).hexdigest()}
    import PIL
    golden["pillow_version"] = PIL.__version__
    with open(os.path.join(ROOT, "tests", "golden", "color_tables.json"), "w") as fh:
        json.dump(golden, fh)
    with open(os.path.join(ROOT, "implementing-jpeg-compression_b200", "csrc", "jb_color_tables.h"), "w") as fh:
        fh.write("// jb_color_tables.h -- generated by tools/derive_pil_tables.py; do not edit.\n"
                 "// Integer tables that reproduce Pillow %s's RGB <-> YCbCr conversion on all 2^24 inputs:\n"
                 "//   y/cb/cr = (T[0][r] + T[1][g] + T[2][b]) >> 6      (the +128 of cb, cr is folded into T[2])\n"
                 "//   r = clip(y + (R_CR[cr] >> 6)), g = clip(y + ((G_CB[cb] + G_CR[cr]) >> 6)), b = clip(y + (B_CB[cb] >> 6))\n"
                 "// JB_COLOR_TABLE_QUAL: storage of the tables (the CUDA file sets it to __constant__).\n"
                 "#pragma once\n#include <stdint.h>\n#ifndef JB_COLOR_TABLE_QUAL\n#define JB_COLOR_TABLE_QUAL static const\n#endif\n\n" % PIL.__version__)
        def emit(name, rows):
            fh.write("JB_COLOR_TABLE_QUAL int16_t %s[%d] = {\n" % (name, len(rows)))
            for i in range(0, len(rows), 16):
                fh.write("    " + ", ".join(str(int(v)) for v in rows[i:i + 16]) + ",\n")
            fh.write("};\n\n")
        for n in ("y", "cb", "cr"):
            flat = [v for tbl in fwd[n] for v in tbl]
            assert max(abs(v) for v in flat) < 32768
            emit("JB_RGB2YCC_%s" % n.upper(), flat)
        for n in ("r_cr", "g_cb", "g_cr", "b_cb"):
            assert max(abs(v) for v in inv[n]) < 32768
            emit("JB_YCC2RGB_%s" % n.upper(), inv[n])
    print("tables written; forward sha", golden["sha256_rgb_to_ycbcr_all_2^24"][:16], "inverse sha", golden["sha256_ycbcr_to_rgb_all_2^24"][:16])


if __name__ == "__main__":
    main()
