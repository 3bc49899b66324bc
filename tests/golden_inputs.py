"""Seeded, formulaic inputs shared by the golden generator and the tests.

Only the *outputs* of the reference are stored under ``tests/golden/``; the
inputs are rebuilt here from the recorded kind/seed/shape, identically in the
dev container (generator), in the CPU tests and on the GPU box.
"""
import numpy as np


def synth_plane(h, w, seed, phase=0.0):
    """Smooth + noise content of SURVEY.md section 8(d):
    clip(127 + 100 sin(x/37) cos(y/53) + N(0, 8^2), 0, 255)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    v = 127.0 + 100.0 * np.sin(x / 37.0 + phase) * np.cos(y / 53.0 + phase)
    v = v + rng.normal(0.0, 8.0, (h, w))
    return np.clip(np.round(v), 0, 255).astype(np.int64)


def make_input(case):
    kind = case["kind"]
    h, w = case["h"], case["w"]
    if kind == "synth":
        return synth_plane(h, w, case["seed"])
    if kind == "uniform":
        return np.random.default_rng(case["seed"]).integers(0, 256, (h, w)).astype(np.int64)
    if kind == "const":
        return np.full((h, w), case["value"], dtype=np.int64)
    if kind == "arange":
        return (np.arange(h * w).reshape(h, w) % 256).astype(np.int64)
    if kind == "checker":
        y, x = np.mgrid[0:h, 0:w]
        return (((x // case["cell"]) + (y // case["cell"])) % 2 * 255).astype(np.int64)
    if kind == "matrix":
        return np.array(case["values"], dtype=np.int64)
    raise ValueError(kind)


def _c(name, kind, h, w, bs, d, transform, qname, qparam=None, **extra):
    c = dict(name=name, kind=kind, h=h, w=w, bs=bs, d=d, transform=transform,
             qname=qname, qparam=qparam)
    c.update(extra)
    return c


CASES = [
    # the default CLI configuration on smooth+noise content, multiple-of-32 and ragged sizes
    _c("synth_96x128_defaults", "synth", 96, 128, 4, 8, "DCT", "qtable", seed=11),
    _c("synth_61x75_defaults_ragged", "synth", 61, 75, 4, 8, "DCT", "qtable", seed=12),
    _c("synth_200x328_defaults", "synth", 200, 328, 4, 8, "DCT", "qtable", seed=13),
    _c("uniform_64x96_defaults", "uniform", 64, 96, 4, 8, "DCT", "qtable", seed=14),
    _c("uniform_64x64_bs1_qtable", "uniform", 64, 64, 1, 8, "DCT", "qtable", seed=15),
    _c("const0_40x40_defaults", "const", 40, 40, 4, 8, "DCT", "qtable", value=0),
    _c("const255_40x40_defaults", "const", 40, 40, 4, 8, "DCT", "qtable", value=255),
    _c("checker_64x64_bs2", "checker", 64, 64, 2, 8, "DCT", "qtable", cell=3),
    # the DFT option (real part only survives)
    _c("synth_64x64_dft_qtable", "synth", 64, 64, 4, 8, "DFT", "qtable", seed=21),
    _c("uniform_50x70_dft_none_bs3", "uniform", 50, 70, 3, 8, "DFT", "none", seed=22),
    _c("synth_31x29_dft_divide7_d5", "synth", 31, 29, 3, 5, "DFT", "divide", 7, seed=23),
    # quantiser modes
    _c("uniform_50x70_none_bs3", "uniform", 50, 70, 3, 8, "DCT", "none", seed=31),
    _c("synth_120x130_divide1000_d24_bs5", "synth", 120, 130, 5, 24, "DCT", "divide", 1000, seed=32),
    _c("uniform_120x130_divide40_d24_bs5", "uniform", 120, 130, 5, 24, "DCT", "divide", 40, seed=33),
    _c("uniform_33x47_discard2_d4_bs2", "uniform", 33, 47, 2, 4, "DCT", "discard", 2, seed=34),
    _c("synth_40x40_discard3_bs2", "synth", 40, 40, 2, 8, "DCT", "discard", 3, seed=35),
    _c("uniform_17x19_divide129_d2_bs1", "uniform", 17, 19, 1, 2, "DCT", "divide", 129, seed=36),
    # degenerate block sizes (reference tests/integration_tests.py:50-66)
    _c("uniform_16x16_d1_bs1_none", "uniform", 16, 16, 1, 1, "DCT", "none", seed=41),
    _c("arange_8x8_bs1_d1", "arange", 8, 8, 1, 1, "DCT", "none"),
    _c("arange_2x3_bs1_d8", "arange", 2, 3, 1, 8, "DCT", "none"),
    _c("arange_8x16_bs3_d8_none", "arange", 8, 16, 3, 8, "DCT", "none"),
    _c("arange_8x16_bs3_d8_dft", "arange", 8, 16, 3, 8, "DFT", "none"),
    _c("arange_8x8_bs1_qtable", "arange", 8, 8, 1, 8, "DCT", "qtable"),
    _c("arange_8x8_bs1_none", "arange", 8, 8, 1, 8, "DCT", "none"),
    _c("matrix_4x4_divide129_d2", "matrix", 4, 4, 1, 2, "DCT", "divide", 129,
       values=[[220, 255, 123, 205], [255, 255, 112, 10], [15, 51, 83, 221], [239, 73, 62, 22]]),
    _c("const255_24x24_divide1000_d24", "const", 24, 24, 1, 24, "DCT", "divide", 1000, value=255),
    _c("const255_24x24_divide40_d24", "const", 24, 24, 1, 24, "DCT", "divide", 40, value=255),
    # amplitude too large for a 15-bit size field -> BadRleCodeError (util.py:170-171)
    _c("const255_24x24_none_d24_error", "const", 24, 24, 1, 24, "DCT", "none", value=255),
    # a single pixel and one-row / one-column planes
    _c("uniform_1x1_defaults", "uniform", 1, 1, 4, 8, "DCT", "qtable", seed=51),
    _c("uniform_1x100_defaults", "uniform", 1, 100, 4, 8, "DCT", "qtable", seed=52),
    _c("uniform_100x1_bs2_none", "uniform", 100, 1, 2, 8, "DCT", "none", seed=53),
    # larger blocks with odd sizes
    _c("synth_130x170_bs2_d16_divide20", "synth", 130, 170, 2, 16, "DCT", "divide", 20, seed=61),
    _c("uniform_70x90_bs1_d32_divide50", "uniform", 70, 90, 1, 32, "DCT", "divide", 50, seed=62),
    _c("synth_64x64_bs8_d3_none", "synth", 64, 64, 8, 3, "DCT", "none", seed=63),
]
