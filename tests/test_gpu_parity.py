"""Parity tests proper: the CUDA path, called through the C-ABI, against the oracle and the
committed outputs of the unmodified reference.  All need a GPU."""
import numpy as np
import pytest

import golden_io
from golden_inputs import make_input, synth_plane
from oracle import ref_port as rp
from parity import check_pixels, check_quantised

pytestmark = pytest.mark.gpu

# JB_FLAG_FORCE_GENERIC = 1, JB_FLAG_NO_TMA = 2 (specialised kernels with plain loads / stores),
# JB_FLAG_TILE_DECODER = 64 (8x8 decoder stores 4-block tiles -- by TMA, or with 2 by plain stores -- instead of chunk rows)
FLAG_SETS = [pytest.param(0, id="default"), pytest.param(1, id="generic"), pytest.param(2, id="no_tma"),
             pytest.param(64, id="tile_decoder"), pytest.param(66, id="tile_decoder_no_tma")]


@pytest.fixture(scope="module")
def jb():
    import jpeg_b200
    return jpeg_b200


def _cfgs(jb, case_or_tuple):
    if isinstance(case_or_tuple, dict):
        c = case_or_tuple
        h, w, bs, d, tr, qn, qp = c["h"], c["w"], c["bs"], c["d"], c["transform"], c["qname"], c.get("qparam")
    else:
        h, w, bs, d, tr, qn, qp = case_or_tuple
    kw = {"keep": qp} if qn == "discard" else {"divisor": qp} if qn == "divide" else {}
    cfg = jb.Configuration(width=w, height=h, block_size=bs, dct_size=d, transform=tr,
                           quantization=jb.QuantizationMethod(qn, **kw))
    return cfg, rp.OracleConfig(w, h, bs, d, tr, qn, qp)


def _check_forward(jb, planes, cfg, ocfg, flags, what):
    """Quantised-integer gate + bit-exact packing gate for a list of planes."""
    n = cfg.dct_size ** 2
    streams = jb.compress_bands(planes, cfg, flags=flags)
    coeffs = jb.stages.forward_coefficients(np.stack(planes).astype(np.uint8), cfg, flags=flags)
    total_ties = 0
    for i, a in enumerate(planes):
        zz = rp.quantised_zigzag(a, ocfg)
        ties = check_quantised(coeffs[i], zz, rp.prerounding_zigzag(a, ocfg), what="%s plane %d" % (what, i))
        total_ties += ties
        # packing is bit exact given the coefficients the GPU itself produced
        assert streams[i] == rp.pack_blocks(coeffs[i].reshape(-1, n)), what
        if ties == 0:
            assert streams[i] == rp.compress_band(a, ocfg), what
    return streams, total_ties


@pytest.mark.parametrize("flags", FLAG_SETS)
@pytest.mark.parametrize("case", golden_io.cases(), ids=golden_io.case_ids())
def test_golden_reference_cases(jb, case, flags):
    a = make_input(case)
    cfg, ocfg = _cfgs(jb, case)
    if case.get("error"):
        with pytest.raises(jb.BadRleCodeError) as ei:
            jb.compress_band(a, cfg, flags=flags)
        assert str(ei.value) == case["error_msg"]
        return
    streams, ties = _check_forward(jb, [a], cfg, ocfg, flags, case["name"])
    if ties == 0:
        assert streams[0] == golden_io.stream(case)
    # decode the REFERENCE's stream
    rec = jb.decompress_band(golden_io.stream(case), cfg, flags=flags)
    assert rec.dtype == np.int64 and rec.shape == (case["h"], case["w"])
    check_pixels(rec, golden_io.restored(case), a, what=case["name"])


GRID = [
    (96, 160, 4, 8, "DCT", "qtable", None), (97, 161, 4, 8, "DCT", "qtable", None),
    (33, 31, 4, 8, "DFT", "qtable", None), (128, 128, 4, 8, "DFT", "none", None),
    (64, 96, 4, 8, "DCT", "none", None), (64, 96, 4, 8, "DCT", "divide", 3),
    (64, 96, 4, 8, "DCT", "discard", 4), (75, 50, 2, 8, "DCT", "qtable", None),
    (40, 56, 1, 8, "DCT", "qtable", None), (240, 250, 5, 24, "DCT", "divide", 1000),
    (240, 250, 5, 24, "DCT", "divide", 40), (90, 70, 3, 5, "DFT", "divide", 7),
    (64, 64, 2, 16, "DCT", "divide", 20), (50, 60, 1, 32, "DCT", "divide", 50),
    (30, 30, 7, 3, "DCT", "none", None), (20, 20, 1, 1, "DCT", "none", None),
    (64, 64, 16, 2, "DFT", "none", None),
    # aligned planes whose last block column is half inside the image (the row stores of the 8x8 decoder)
    (40, 48, 4, 8, "DCT", "qtable", None), (72, 176, 4, 8, "DFT", "qtable", None),
]


@pytest.mark.parametrize("flags", FLAG_SETS)
@pytest.mark.parametrize("h,w,bs,d,tr,qn,qp", GRID)
def test_differential_grid(jb, h, w, bs, d, tr, qn, qp, flags):
    cfg, ocfg = _cfgs(jb, (h, w, bs, d, tr, qn, qp))
    rng = np.random.default_rng(h * 7 + w)
    planes = [synth_plane(h, w, 3, phase=0.5), rng.integers(0, 256, (h, w)).astype(np.int64),
              np.full((h, w), 200, dtype=np.int64)]
    what = "grid %s" % ((h, w, bs, d, tr, qn, qp),)
    streams, _ = _check_forward(jb, planes, cfg, ocfg, flags, what)
    rec = jb.decompress_bands(streams, cfg, flags=flags)
    for i, a in enumerate(planes):
        check_pixels(rec[i], rp.decompress_band(streams[i], ocfg), a, what=what)


def test_rle_stage_vectors_from_the_reference_tests(jb):
    # tests/RLE_tests.py:16-26,36-46,77-86,99-122 through the CUDA packer / unpacker
    a = np.zeros((1, 64), dtype=np.int64)
    a[0, :15] = [-15, 0, 0, 0, 3, 2, 0, 0, 0, 0, 120, 0, 0, 0, 0]
    s = jb.stages.pack_coefficients(a, 8)[0]
    assert s == rp.pack_tuples([(0, 5, -15), (3, 3, 3), (0, 3, 2), (4, 8, 120), (0, 0)])
    a = np.zeros((1, 64), dtype=np.int64)
    a[0, 1], a[0, 34] = 2, 5
    s = jb.stages.pack_coefficients(a, 8)[0]
    assert s == rp.pack_tuples([(1, 3, 2), (15, 0, 0), (15, 0, 0), (2, 4, 5), (0, 0)])
    blocks = np.zeros((3, 9), dtype=np.int64)
    blocks[0] = [21, 3, 0, 0, 0, 0, 2, 0, 0]
    blocks[1] = [0, 0, 0, 15, 0, 0, 0, 0, 9]
    s = jb.stages.pack_coefficients(blocks, 3)[0]
    assert s == rp.pack_tuples([(0, 6, 21), (0, 3, 3), (4, 3, 2), (0, 0), (3, 5, 15), (4, 5, 9), (0, 0), (0, 0)])
    assert jb.stages.unpack_streams([s], 3, 3)[0].tolist() == blocks.tolist()
    a = np.zeros((1, 64), dtype=np.int64)
    a[0, 4] = 2
    bits = "".join(format(x, "08b") for x in jb.stages.pack_coefficients(a, 8)[0])
    assert bits == "0100" + "0011" + "110" + "0" * 13
    for x in ([(15, 0, 0), (15, 0, 0), (0, 2, 1), (0, 0)],
              [(1, 2, -1), (0, 3, -2), (8, 3, -3), (8, 5, -15), (0, 0)],
              [(14, 4, 7), (0, 0), (0, 0), (15, 0, 0), (0, 2, 1), (0, 0)]):
        want = rp.tuples_to_blocks(x, 64)
        got = jb.stages.unpack_streams([rp.pack_tuples(x)], want.shape[0], 8)[0]
        assert got.tolist() == want.tolist()
    with pytest.raises(jb.BadRleCodeError):
        bad = np.zeros((2, 64), dtype=np.int64)
        bad[1, 3] = 20000
        jb.stages.pack_coefficients(bad, 8)


@pytest.mark.parametrize("n,density,bits", [(64, 0.25, 7), (64, 1.0, 14), (64, 0.0, 1), (576, 0.05, 10),
                                            (1024, 0.02, 12), (1, 0.6, 9), (25, 0.3, 5)])
def test_pack_unpack_random_coefficients(jb, n, density, bits):
    rng = np.random.default_rng(n + bits)
    d = int(round(n ** 0.5))
    zz = np.zeros((3, 700, n), dtype=np.int64)
    mask = rng.random(zz.shape) < density
    vals = rng.integers(1, 1 << bits, zz.shape) * (rng.integers(0, 2, zz.shape) * 2 - 1)
    zz[mask] = vals[mask]
    streams = jb.stages.pack_coefficients(zz, d)
    for i in range(3):
        assert streams[i] == rp.pack_blocks(zz[i])
    back = jb.stages.unpack_streams(streams, 700, d)
    assert np.array_equal(back, zz)


# framing variants of the decoder: 0 = the parallel tile walk, JB_FLAG_SERIAL_FRAMING = 8 the single-thread walk
FRAMING = [pytest.param(0, id="tile_walk"), pytest.param(8, id="serial_walk")]


@pytest.mark.parametrize("framing", FRAMING)
def test_framing_stream_lengths_around_the_walk_warp_boundaries(jb, framing):
    """The walk gives every stream whole warps of 32 tiles (256 bytes each here): streams whose tile counts sit
    on and around 32 and 64, a one-tile stream and an all-EOB stream in one batch, decoded block for block."""
    n_blocks, n = 400, 64
    rng = np.random.default_rng(2024)
    # bytes per block grow with the density; find densities whose streams land on the wanted tile counts
    wanted = [1, 2, 31, 32, 33, 40, 63, 64, 65, 90]
    planes, got = [], []
    for tiles in wanted:
        target = tiles * 256 - 100
        lo, hi = 0.0, 1.0
        for _ in range(30):                       # bisection on the density (same random field every time)
            mid = (lo + hi) / 2
            field = np.random.default_rng(tiles).random((n_blocks, n))
            vals = np.random.default_rng(tiles + 1).integers(1, 128, (n_blocks, n))
            zz = np.where(field < mid, vals, 0).astype(np.int64)
            size = len(rp.pack_blocks(zz))
            if size > target:
                hi = mid
            else:
                lo = mid
        planes.append(zz)
        got.append(-(-size // 256))
    planes.append(np.zeros((n_blocks, n), dtype=np.int64))             # every block a lone EOB
    zz = np.stack(planes)
    streams = [rp.pack_blocks(z) for z in zz]
    tiles = [-(-len(x) // 256) for x in streams]
    assert {32, 64} & set(tiles) and min(tiles) <= 2 and max(tiles) >= 80, tiles
    assert len({t % 32 for t in tiles}) >= 6, tiles
    back = jb.stages.unpack_streams(streams, n_blocks, 8, flags=framing)
    assert np.array_equal(back, zz)
    assert jb.stages.pack_coefficients(zz, 8) == streams


def test_serial_framing_fallback_matches(jb):
    """JB_FLAG_SERIAL_FRAMING (8) routes every stream through the single-thread fallback walk."""
    cfg, ocfg = _cfgs(jb, (200, 328, 4, 8, "DCT", "qtable", None))
    planes = [synth_plane(200, 328, 3 + i) for i in range(3)]
    streams = jb.compress_bands(planes, cfg)
    assert np.array_equal(jb.decompress_bands(streams, cfg, flags=8), jb.decompress_bands(streams, cfg))
    with pytest.raises(jb.BadStreamError):
        jb.decompress_bands([streams[0], streams[1][:-2], streams[2]], cfg, flags=8)


def test_one_long_stream_among_short_ones(jb):
    """A batch of short streams takes the one-launch framing path (one CTA per stream); a stream far
    longer than its peers (more tiles than that path holds) is walked serially and still decodes."""
    import torch
    h = w = 1536
    cfg, ocfg = _cfgs(jb, (h, w, 1, 8, "DCT", "none", None))
    rng = np.random.default_rng(5)
    n_zero = 60
    noisy = rng.integers(0, 256, (h, w), dtype=np.uint8)
    planes = torch.zeros((n_zero + 1, h, w), dtype=torch.uint8, device="cuda")
    planes[17] = torch.from_numpy(noisy).cuda()
    comp = jb.compress_planes(planes, cfg)
    off = comp.host_offsets()
    lens = np.diff(off)
    assert lens[17] > 4096 * 256 and lens.sum() / (n_zero + 1) < 1024 * 256
    out, status = jb.decompress_planes(comp.data, comp.offsets[:-1], comp.offsets[1:] - comp.offsets[:-1], cfg,
                                       n_zero + 1, in_bytes=int(lens.sum()))
    jb.check_status(status)
    assert int(status.cpu()[2].item()) == 1                  # exactly the long stream took the serial walk
    assert torch.equal(out, planes)                          # 'none' on 8-bit data with block_size 1 is lossless


@pytest.mark.parametrize("framing", FRAMING)
def test_framing_with_many_false_block_starts(jb, framing):
    """Streams full of 0x00 bytes inside amplitude fields: most candidate offsets are false."""
    rng = np.random.default_rng(77)
    zz = np.zeros((2, 3000, 64), dtype=np.int64)
    pos = rng.integers(0, 64, (2, 3000, 6))
    for p in range(2):
        for b in range(3000):
            zz[p, b, pos[p, b]] = rng.choice([256, 512, 1024, 2048, 4096, 8192, -256, -4096, 1, -1], 6)
    streams = [rp.pack_blocks(zz[p]) for p in range(2)]
    frac_zero = np.mean(np.frombuffer(streams[0], dtype=np.uint8) == 0)
    assert frac_zero > 0.25
    back = jb.stages.unpack_streams(streams, 3000, 8, flags=framing)
    assert np.array_equal(back, zz)


@pytest.mark.parametrize("framing", FRAMING)
def test_framing_windows_halo_and_stream_alignment(jb, framing):
    """Streams that start at any byte alignment, with dense (long) blocks next to sparse ones, lengths and offsets
    of every residue mod 4, and nonzero padding bits, which the reference's decoder skips without looking
    (rle_byte_stream.py:80-82)."""
    rng = np.random.default_rng(31)
    n_blocks, n = 900, 64
    planes = []
    for i in range(7):
        dens = [0.9, 0.05, 0.5, 1.0, 0.2, 0.7, 0.02][i]
        vals = rng.integers(1, 1 << [14, 3, 9, 12, 6, 14, 2][i], (n_blocks, n)) * (rng.integers(0, 2, (n_blocks, n)) * 2 - 1)
        planes.append(np.where(rng.random((n_blocks, n)) < dens, vals, 0).astype(np.int64))
    zz = np.stack(planes)
    streams = [rp.pack_blocks(z) for z in zz]
    assert max(len(x) for x in streams) > 3 * 16384 and len({len(x) % 4 for x in streams}) >= 3
    assert len({sum(len(x) for x in streams[:i]) % 4 for i in range(7)}) >= 3          # start alignments
    back = jb.stages.unpack_streams(streams, n_blocks, 8, flags=framing)
    assert np.array_equal(back, zz)
    # padding bits set to 1 behind EOBs that do not end on a byte boundary: the reference's reader skips them
    # unseen, so such a stream is valid; here it takes the serial walk
    small = zz[1][:40]
    parts, touched = [], 0
    for blk in small:
        enc = bytearray(rp.pack_blocks(blk[None]))
        nz = np.nonzero(blk)[0]
        bits, prev = 8, -1
        for pos in nz:
            bits += 8 * ((pos - prev - 1) // 15) + 8 + int(abs(int(blk[pos]))).bit_length() + 1
            prev = pos
        pad = len(enc) * 8 - bits
        assert 0 <= pad < 8
        if pad:
            enc[-1] |= (1 << pad) - 1
            touched += 1
        parts.append(bytes(enc))
    assert touched > 10
    got = jb.stages.unpack_streams([b"".join(parts)], 40, 8, flags=framing)
    assert np.array_equal(got[0], small)


@pytest.mark.parametrize("framing", FRAMING)
def test_zero_heavy_streams_every_byte_a_candidate(jb, framing):
    # an all-black plane packs to one 0x00 per block: every byte is a block start
    cfg, ocfg = _cfgs(jb, (600, 800, 1, 8, "DCT", "qtable", None))
    a = np.zeros((600, 800), dtype=np.int64)
    s = jb.compress_band(a, cfg)
    assert s == b"\0" * (75 * 100)
    assert np.array_equal(jb.decompress_band(s, cfg, flags=framing), a)


@pytest.mark.parametrize("framing", FRAMING)
def test_malformed_streams_are_rejected(jb, framing):
    cfg, ocfg = _cfgs(jb, (64, 64, 4, 8, "DCT", "qtable", None))
    a = synth_plane(64, 64, 1)
    good = jb.compress_band(a, cfg)
    assert np.array_equal(jb.decompress_band(good, cfg, flags=framing), jb.decompress_band(good, cfg))
    with pytest.raises(jb.BadStreamError):
        jb.decompress_band(good[:-1], cfg, flags=framing)              # truncated
    with pytest.raises(jb.BadStreamError):
        jb.decompress_band(good + b"\0", cfg, flags=framing)           # one block too many
    with pytest.raises(jb.BadStreamError):
        jb.decompress_band(b"\x50" + good[1:], cfg, flags=framing)     # (5, 0, 0) is not a code
    with pytest.raises(jb.BadStreamError):
        jb.decompress_band(b"", cfg, flags=framing)                    # nothing at all
    # an error in one call leaves the workspace usable: the next call on good data succeeds
    assert np.array_equal(jb.decompress_band(good, cfg, flags=framing), jb.decompress_band(good, cfg))
    with pytest.raises(jb.BadQuantizationError):
        jb.Configuration(width=8, height=8, dct_size=4, quantization=jb.QuantizationMethod("qtable"))
    with pytest.raises(jb.EmptyArrayError):
        jb.compress_band(np.zeros((0, 5)), cfg)
    with pytest.raises(jb.BadArrayShapeError):
        jb.compress_band(np.zeros(5), cfg)


@pytest.mark.parametrize("flags", FLAG_SETS)
@pytest.mark.parametrize("h,w,tr,qn", [(1080, 1920, "DCT", "qtable"), (2160, 3840, "DCT", "qtable"),
                                       (1024, 2048, "DFT", "qtable")])
def test_full_size_frames(jb, h, w, tr, qn, flags):
    """BASELINE.json configs 2 / 4 / 5 geometry (one frame each): all gates, plus the
    round trip through the container and the image facade."""
    cfg, ocfg = _cfgs(jb, (h, w, 4, 8, tr, qn, None))
    planes = [synth_plane(h, w, 1000 + b, phase=0.7 * b) for b in range(3)]
    streams, ties = _check_forward(jb, planes, cfg, ocfg, flags, "%dx%d %s" % (h, w, tr))
    rec = jb.decompress_bands(streams, cfg, flags=flags)
    for i in range(3):
        check_pixels(rec[i], rp.decompress_band(streams[i], ocfg), planes[i], what="full frame")
    blob = jb.file_format.generate_data(cfg, jb.CompressedData(*streams))
    cfg2, data = jb.file_format.read_data(blob)
    assert (cfg2.width, cfg2.height, cfg2.block_size, cfg2.dct_size, cfg2.transform) == (w, h, 4, 8, tr)
    assert [data.y, data.cb, data.cr] == streams


def test_large_block_config3(jb):
    """BASELINE.json config 3: 3840x2160, --block_size 5 --dct_size 24 --quantization divide --qdivisor 1000."""
    h, w = 2160, 3840
    cfg, ocfg = _cfgs(jb, (h, w, 5, 24, "DCT", "divide", 1000))
    planes = [synth_plane(h, w, 77)]
    streams, _ = _check_forward(jb, planes, cfg, ocfg, 0, "config3")
    rec = jb.decompress_bands(streams, cfg)
    check_pixels(rec[0], rp.decompress_band(streams[0], ocfg), planes[0], what="config3")


def test_image_facade_round_trip(jb):
    from PIL import Image
    rng = np.random.default_rng(9)
    rgb = np.clip(np.stack([synth_plane(120, 200, 5 + i) for i in range(3)], -1) + rng.integers(-3, 4, (120, 200, 3)),
                  0, 255).astype(np.uint8)
    im = Image.fromarray(rgb, "RGB").convert("YCbCr")
    cfg = jb.Configuration(width=im.width, height=im.height, block_size=4, dct_size=8, transform="DCT",
                           quantization=jb.QuantizationMethod("qtable"))
    blob = jb.Jpeg(cfg).compress(im)
    out = jb.Jpeg.decompress(blob)
    assert out.mode == "YCbCr" and out.size == im.size
    ocfg = rp.OracleConfig(im.width, im.height, 4, 8, "DCT", "qtable")
    bands = [np.asarray(b).astype(np.int64) for b in im.split()]
    cfg2, data = jb.file_format.read_data(blob)
    for got, band, stream in zip(out.split(), bands, (data.y, data.cb, data.cr)):
        check_pixels(np.asarray(got), rp.decompress_band(stream, ocfg), band, what="facade")


def test_batch_equals_per_plane_and_sharding_concatenates(jb):
    """Multi-GPU contract on one GPU: a batch equals plane-by-plane calls, and block-row
    bands compressed separately concatenate to the whole-image stream."""
    h, w = 272, 320
    cfg, _ = _cfgs(jb, (h, w, 4, 8, "DCT", "qtable", None))
    planes = [synth_plane(h, w, 40 + i) for i in range(6)]
    batch = jb.compress_bands(planes, cfg)
    assert batch == [jb.compress_band(p, cfg) for p in planes]
    whole = jb.compress_band(planes[0], cfg)
    parts = []
    for r0, r1 in jb.sharding.block_row_bands(h, 4, 8, 4):
        if r1 > r0:
            sub, _ = _cfgs(jb, (r1 - r0, w, 4, 8, "DCT", "qtable", None))
            parts.append([jb.compress_band(planes[0][r0:r1], sub)])
        else:
            parts.append([])
    assert jb.sharding.concat_band_streams(parts)[0] == whole


def test_pipelined_host_round_trip_equals_the_two_calls(jb):
    """BatchCodec.roundtrip_host (sub-batches pipelined on two streams) returns what compress_host
    followed by decompress_host returns, and its per-sub-batch streams concatenate to the same bytes."""
    import torch
    h, w, n = 136, 264, 10
    cfg, _ = _cfgs(jb, (h, w, 4, 8, "DCT", "qtable", None))
    planes = np.stack([synth_plane(h, w, 60 + i) for i in range(n)]).astype(np.uint8)
    hp = torch.from_numpy(planes).pin_memory()
    bc = jb.BatchCodec(cfg, n)
    hs, offs = bc.compress_host(hp)
    ref_streams = hs.numpy().tobytes()
    ref_dec = bc.decompress_host(hs, offs).clone()
    for n_sub in (1, 3, 10):
        dec, parts = bc.roundtrip_host(hp, n_sub)
        assert torch.equal(dec, ref_dec)
        assert [p[0] for p in parts] == [n * j // n_sub for j in range(n_sub)]
        assert b"".join(p[1].numpy().tobytes() for p in parts) == ref_streams


def test_colour_conversion_matches_pillow_on_every_colour(jb):
    """jb_rgb_to_ycbcr_planes / jb_ycbcr_planes_to_rgb against the golden hashes of Pillow's conversion of all
    2^24 inputs (tests/golden/color_tables.json, made by tools/derive_pil_tables.py), and against live Pillow
    on ragged, unaligned shapes where it is importable."""
    import hashlib
    import json
    import os
    import torch
    from test_color_tables import all_colours
    with open(os.path.join(os.path.dirname(__file__), "golden", "color_tables.json")) as fh:
        golden = json.load(fh)
    src = torch.from_numpy(all_colours()).cuda()
    planes = jb.rgb_to_ycbcr_planes(src)                              # [3, 4096, 4096]
    ycc = torch.stack([planes[0], planes[1], planes[2]], dim=-1).cpu().numpy()
    assert hashlib.sha256(ycc.tobytes()).hexdigest() == golden["sha256_rgb_to_ycbcr_all_2^24"]
    as_planes = src.permute(2, 0, 1).contiguous()                     # the same triples read as (y, cb, cr)
    rgb = jb.ycbcr_planes_to_rgb(as_planes)[0].cpu().numpy()
    assert hashlib.sha256(rgb.tobytes()).hexdigest() == golden["sha256_ycbcr_to_rgb_all_2^24"]
    try:
        from PIL import Image
    except ImportError:
        return
    rng = np.random.default_rng(11)
    for n, h, w in ((1, 37, 53), (3, 16, 16), (2, 5, 130)):
        img = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
        want = np.stack([np.asarray(Image.fromarray(img[i], "RGB").convert("YCbCr")) for i in range(n)])
        got = jb.rgb_to_ycbcr_planes(torch.from_numpy(img).cuda()).cpu().numpy().reshape(n, 3, h, w)
        assert np.array_equal(got, want.transpose(0, 3, 1, 2))
        back = jb.ycbcr_planes_to_rgb(torch.from_numpy(np.ascontiguousarray(img.transpose(0, 3, 1, 2)).reshape(3 * n, h, w)).cuda())
        want_rgb = np.stack([np.asarray(Image.fromarray(img[i], "YCbCr").convert("RGB")) for i in range(n)])
        assert np.array_equal(back.cpu().numpy(), want_rgb)


def test_rgb_front_ends_equal_the_pillow_flow(jb):
    """Jpeg.compress_rgb / decompress_rgb == the compress.py / decompress.py flows with PIL's conversions."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(21)
    h, w = 72, 104
    base = np.stack([synth_plane(h, w, 90 + i) for i in range(3)], axis=-1).astype(np.uint8)
    img = Image.fromarray(base, "RGB")
    cfg, _ = _cfgs(jb, (h, w, 4, 8, "DCT", "qtable", None))
    blob = jb.Jpeg(cfg).compress_rgb(img)
    assert blob == jb.Jpeg(cfg).compress(img.convert("YCbCr"))
    out = jb.Jpeg.decompress_rgb(blob)
    assert np.array_equal(np.asarray(out), np.asarray(jb.Jpeg.decompress(blob).convert("RGB")))


@pytest.mark.parametrize("tr", ["DCT", "DFT"])
@pytest.mark.parametrize("h,w,n", [(90, 1536, 5), (64, 1920, 7), (200, 3840, 2), (33, 2048, 3)])
def test_decoder_variants_agree_on_wide_planes(jb, h, w, n, tr):
    """Dense, wide planes (bottom crop, several planes per batch): the row-store decoder (default), the tile decoder
    (64) and the generic kernel (1) give the same pixels -- the fast variants bit for bit, the generic one and the
    oracle within 1 LSB."""
    import torch
    cfg, ocfg = _cfgs(jb, (h, w, 4, 8, tr, "qtable", None))
    planes = np.stack([synth_plane(h, w, 300 + i) for i in range(n)]).astype(np.uint8)
    comp = jb.compress_planes(torch.from_numpy(planes).cuda(), cfg)
    lens = comp.offsets[1:] - comp.offsets[:-1]
    outs = {}
    for flags in (0, 64, 1):
        out, status = jb.decompress_planes(comp.data, comp.offsets[:-1], lens, cfg, n, in_bytes=comp.total_bytes(), flags=flags)
        jb.check_status(status)
        outs[flags] = out.cpu().numpy()
    assert np.array_equal(outs[0], outs[64])
    assert np.abs(outs[0].astype(np.int64) - outs[1].astype(np.int64)).max() <= 1
    streams = comp.to_bytes_list()
    want = rp.decompress_band(streams[n - 1], ocfg)
    assert np.abs(outs[0][n - 1].astype(np.int64) - want).max() <= 1


def test_batched_containers_equal_generate_data(jb):
    """jb_pack_containers == file_format.generate_data per image (file_format.py:86-93), for stream lengths of
    every alignment, and the batched RGB front end == the per-image Pillow flow."""
    import torch
    h, w, n = 40, 72, 7
    cfg, _ = _cfgs(jb, (h, w, 4, 8, "DCT", "qtable", None))
    planes = np.stack([synth_plane(h, w, 500 + i) for i in range(3 * n)]).astype(np.uint8)
    planes[4] = 0                                       # very short streams too
    comp = jb.compress_planes(torch.from_numpy(planes).cuda(), cfg)
    streams = comp.to_bytes_list()
    out, offs, status = jb.pack_containers(comp, cfg)
    got = jb.containers_to_bytes(out, offs, status)
    want = [jb.file_format.generate_data(cfg, jb.CompressedData(*streams[3 * i:3 * i + 3])) for i in range(n)]
    assert got == want
    try:
        from PIL import Image
    except ImportError:
        return
    rng = np.random.default_rng(8)
    imgs = [Image.fromarray((planes[3 * i:3 * i + 3].transpose(1, 2, 0) ^ rng.integers(0, 8, (h, w, 3), dtype=np.uint8)), "RGB")
            for i in range(n)]
    batch = jb.compress_images_rgb(imgs, cfg)
    assert batch == [jb.Jpeg(cfg).compress(im.convert("YCbCr")) for im in imgs]


# ---- BASELINE.json config 5: long streams (gigapixel DFT plane, block-row-band sharding) ------------------------
def _big_plane(h, w, seed):
    """Smooth + noise content for planes of tens of megapixels without tens of seconds of rng: a 1024 x 2048
    synth_plane tiled, plus a cheap position-dependent perturbation so that no two tiles are equal."""
    base = synth_plane(1024, 2048, seed)
    a = np.tile(base, (-(-h // 1024), -(-w // 2048)))[:h, :w]
    yy = np.arange(h, dtype=np.int64)[:, None]
    xx = np.arange(w, dtype=np.int64)[None, :]
    return np.clip(a + (yy * 7 + xx * 13 + seed) % 5 - 2, 0, 255)


@pytest.mark.parametrize("h,w,n_planes,min_tiles", [(6144, 16384, 1, 4097), (4096, 4096, 3, 1025)])
def test_long_stream_framing_config5(jb, h, w, n_planes, min_tiles):
    """The decoder path config 5 takes: streams of thousands of walk tiles go through jb_frame_reach / link /
    scan / emit (csrc/jb_framing.cu) instead of the one-CTA-per-stream kernel every other test uses -- one plane
    long enough for the large reach instantiation (> 4096 tiles), and a 3-plane frame for the small one.  All
    gates against the oracle: quantised integers (strict 1e-5 fraction; the DFT ties are re-evaluated in
    pocketfft's order, so the streams are byte-identical), packing, pixels, and no serial fallback."""
    import ctypes
    import torch
    cfg, ocfg = _cfgs(jb, (h, w, 4, 8, "DFT", "qtable", None))
    planes = [_big_plane(h, w, 70 + i) for i in range(n_planes)]
    d_planes = torch.from_numpy(np.stack(planes).astype(np.uint8)).cuda()
    comp = jb.compress_planes(d_planes, cfg)
    streams = comp.to_bytes_list()
    total = comp.total_bytes()
    p = cfg.c_params()
    lib = jb._lib.load()
    assert lib.jb_decompress_framing_path(ctypes.byref(p), n_planes, total) == 1          # JB_FRAMING_CHAIN
    tile_bytes = 256
    while tile_bytes < 4096 and tile_bytes < 12.0 * total / (n_planes * (h // 32) * (w // 32)):
        tile_bytes *= 2
    assert min(len(s) for s in streams) // tile_bytes >= min_tiles - 1, [len(s) for s in streams]
    coeffs = jb.stages.forward_coefficients(d_planes.cpu().numpy(), cfg)
    n_ties = 0
    for i, a in enumerate(planes):
        pre = rp.prerounding_zigzag(a, ocfg)
        n_ties += check_quantised(coeffs[i], rp.quantised_zigzag(a, ocfg), pre, what="config5 plane %d" % i,
                                  strict_fraction=True)
        assert streams[i] == rp.pack_blocks(coeffs[i].reshape(-1, 64))
    assert n_ties == 0, "%d DFT tie mismatches: the fp64 re-evaluation left pocketfft's order" % n_ties
    for i, a in enumerate(planes):
        assert streams[i] == rp.compress_band(a, ocfg)
    lens = comp.offsets[1:] - comp.offsets[:-1]
    out, status = jb.decompress_planes(comp.data, comp.offsets[:-1], lens, cfg, n_planes, in_bytes=total)
    jb.check_status(status)
    assert int(status.cpu()[2].item()) == 0                    # nobody took the serial fallback walk
    rec = out.cpu().numpy()
    for i, a in enumerate(planes):
        check_pixels(rec[i], rp.decompress_band(streams[i], ocfg), a, what="config5 plane %d" % i)


def test_band_sharded_gigapixel_frame_concatenates(jb):
    """config 5's sharding on one GPU: the 8 block-row bands of a 3-plane 4096 x 4096 DFT frame, compressed as
    images of their own (what each of 8 ranks does), concatenate per colour plane to the whole-image streams,
    and every band decodes to the rows the whole-image decode gives."""
    import torch
    h = w = 4096
    cfg, ocfg = _cfgs(jb, (h, w, 4, 8, "DFT", "qtable", None))
    planes = np.stack([_big_plane(h, w, 80 + i) for i in range(3)]).astype(np.uint8)
    d_planes = torch.from_numpy(planes).cuda()
    whole = jb.compress_planes(d_planes, cfg).to_bytes_list()
    whole_dec = jb.decompress_bands(whole, cfg)
    parts = []
    for r0, r1 in jb.sharding.block_row_bands(h, 4, 8, 8):
        assert r1 - r0 == h // 8
        sub, _ = _cfgs(jb, (r1 - r0, w, 4, 8, "DFT", "qtable", None))
        band = jb.compress_planes(d_planes[:, r0:r1].contiguous(), sub).to_bytes_list()
        parts.append(band)
        assert np.array_equal(jb.decompress_bands(band, sub), whole_dec[:, r0:r1])
    assert jb.sharding.concat_band_streams(parts) == whole
    assert whole[1] == rp.compress_band(planes[1].astype(np.int64), ocfg)


def test_programmatic_dependent_launch_gives_the_same_results(jb):
    """JB_FLAG_PDL (128): the kernels of a call -- and of the calls queued behind each other -- are launched with
    programmatic stream serialization; their prologues overlap the predecessor's tail and they wait before touching
    its data.  Same streams and pixels as without it, also when calls follow each other back to back on one
    stream, through a CUDA graph, and with the table reuse that moves the table loads in front of the wait."""
    import torch
    h, w, n = 272, 480, 12
    cfg, ocfg = _cfgs(jb, (h, w, 4, 8, "DCT", "qtable", None))
    planes = torch.from_numpy(np.stack([synth_plane(h, w, 700 + i) for i in range(n)]).astype(np.uint8)).cuda()
    ref = jb.BatchCodec(cfg, n)
    comp = ref.compress_device(planes)
    want_streams = comp.to_bytes_list()
    want_pixels, st = ref.decompress_device(comp, comp.total_bytes())
    jb.check_status(st)
    want_pixels = want_pixels.clone()
    bc = jb.BatchCodec(cfg, n, flags=jb._lib.JB_FLAG_PDL)
    for rep in range(4):                                   # the first call builds tables, the later ones reuse them
        c2 = bc.compress_device(planes)
        out, st = bc.decompress_device(c2, comp.total_bytes())
    jb.check_status(st)
    assert c2.to_bytes_list() == want_streams
    assert torch.equal(out, want_pixels)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        bc.compress_device(planes)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        c3 = bc.compress_device(planes)
        out3, st3 = bc.decompress_device(c3, comp.total_bytes())
    for rep in range(3):
        g.replay()
    torch.cuda.synchronize()
    jb.check_status(st3)
    assert c3.to_bytes_list() == want_streams
    assert torch.equal(out3, want_pixels)


@pytest.mark.parametrize("h,w,bs,d,tr,qn,qp,pitch_extra", [
    (97, 161, 4, 8, "DCT", "qtable", None, 15), (97, 161, 4, 8, "DCT", "qtable", None, 2), (128, 256, 4, 8, "DFT", "qtable", None, 32),
    (240, 250, 5, 24, "DCT", "divide", 1000, 6), (360, 1024, 5, 24, "DCT", "divide", 40, 16), (50, 60, 1, 32, "DCT", "divide", 50, 4),
    (33, 31, 3, 5, "DFT", "divide", 7, 1), (64, 64, 2, 16, "DCT", "divide", 20, 0)])
def test_no_write_outside_the_callers_buffers(jb, h, w, bs, d, tr, qn, qp, pitch_extra):
    """The sanitizer substitute (compute-sanitizer is not available on the GPU pool): every buffer the library
    writes -- streams, offsets, status, workspaces, decoded planes -- sits inside a larger allocation pre-filled with
    0xA5, with 256-byte guard bands on both sides; the decoded planes are given a row pitch wider than the image, so
    that the bytes between the rows are guards too.  After compress + decompress all guards must be intact, and the
    part of the stream buffer behind the last stream byte as well."""
    import ctypes
    import torch
    cfg, ocfg = _cfgs(jb, (h, w, bs, d, tr, qn, qp))
    n = 3
    lib = jb._lib.load()
    p = cfg.c_params()
    G = 256

    def guarded(nbytes):
        buf = torch.full((nbytes + 2 * G,), 0xA5, dtype=torch.uint8, device="cuda")
        return buf, buf[G:G + nbytes]

    def intact(buf, nbytes):
        return bool((buf[:G] == 0xA5).all()) and bool((buf[G + nbytes:] == 0xA5).all())

    planes = torch.from_numpy(np.stack([synth_plane(h, w, 900 + i) for i in range(n)]).astype(np.uint8)).cuda()
    cap = int(lib.jb_max_stream_bytes(ctypes.byref(p), n))
    out_buf, out = guarded(cap)
    ws_bytes = int(lib.jb_compress_workspace_bytes(ctypes.byref(p), n))
    ws_buf, ws = guarded(ws_bytes)
    comp = jb.compress_planes(planes, cfg, out=out, ws=ws)
    total = comp.total_bytes()
    want = jb.compress_planes(planes, cfg).to_bytes_list()
    assert comp.to_bytes_list() == want
    assert intact(out_buf, cap) and intact(ws_buf, ws_bytes)
    assert bool((out[total:] == 0xA5).all()), "compress wrote behind the last stream byte"
    # decode into planes with a wider pitch, inside a guarded allocation
    pitch = w + pitch_extra
    dec_bytes = n * h * pitch
    dec_buf, dec_flat = guarded(dec_bytes)
    dec = torch.as_strided(dec_flat, (n, h, w), (h * pitch, pitch, 1))
    ws2_bytes = int(lib.jb_decompress_workspace_bytes(ctypes.byref(p), n, total))
    ws2_buf, ws2 = guarded(ws2_bytes)
    lens = comp.offsets[1:] - comp.offsets[:-1]
    got, status = jb.decompress_planes(comp.data, comp.offsets[:-1], lens, cfg, n, in_bytes=total, out=dec, ws=ws2)
    jb.check_status(status)
    assert np.array_equal(got.cpu().numpy(), jb.decompress_bands(want, cfg))
    assert intact(dec_buf, dec_bytes) and intact(ws2_buf, ws2_bytes)
    if pitch_extra:
        gaps = torch.as_strided(dec_flat, (n, h, pitch_extra), (h * pitch, pitch, 1), storage_offset=dec_flat.storage_offset() + w)
        assert bool((gaps == 0xA5).all()), "decompress wrote between the rows"


# ---- float64 stage hooks (jb_stage_float64): the reference's stage-by-stage intermediates ---------------------------
# pipeline/base.py:42-72 -- every stage hands its execute() result to the next one.  The hook computes these arrays
# with the arithmetic that decides rounding ties inside the fused kernels (csrc/jb_refine.cuh), so a bit-for-bit match
# with the oracle pins that arithmetic on EVERY coefficient, not only on the few that land near a tie.
F64_CASES = [
    # h, w, bs, d, transform, quantiser, parameter, bitwise
    (64, 96, 4, 8, "DCT", "qtable", None, True),
    (61, 75, 4, 8, "DCT", "qtable", None, True),           # ragged: both edge replications
    (97, 161, 3, 8, "DCT", "divide", 7, True),             # block_size not a power of two: the mean is a division
    (40, 56, 1, 8, "DCT", "none", None, True),
    (50, 70, 2, 4, "DCT", "discard", 3, True),
    (120, 240, 5, 24, "DCT", "divide", 1000, True),        # config 3's geometry
    (66, 130, 2, 16, "DCT", "divide", 10, True),
    (64, 64, 1, 32, "DCT", "divide", 64, True),            # (undivided, the DC term of a 32 x 32 block overflows the size field)
    (64, 96, 4, 8, "DFT", "qtable", None, True),           # pocketfft's radix-8 pass, operation for operation
    (61, 75, 2, 8, "DFT", "divide", 3, True),
    (48, 48, 2, 4, "DFT", "none", None, False),            # other DFT sizes: direct sums (not pocketfft's order)
]


@pytest.mark.parametrize("h,w,bs,d,tr,qn,qp,bitwise", F64_CASES)
def test_float64_stage_intermediates_equal_the_oracles(jb, h, w, bs, d, tr, qn, qp, bitwise):
    cfg, ocfg = _cfgs(jb, (h, w, bs, d, tr, qn, qp))
    planes = [synth_plane(h, w, seed=100 + k) for k in range(2)]
    planes[1][:, :] = np.random.default_rng(5).integers(0, 256, size=(h, w))          # noise: every coefficient busy
    stack = np.stack(planes).astype(np.uint8)
    samples = jb.stages.float64_stage(stack, cfg, "samples")
    coeff = jb.stages.float64_stage(stack, cfg, "transform")
    pre = jb.stages.float64_stage(stack, cfg, "prerounding")
    for i, a in enumerate(planes):
        x = rp.subsample(a.astype(np.int64), bs, d)                                    # stages 0-3
        assert samples[i].shape == x.shape
        assert np.array_equal(samples[i].view(np.int64), x.view(np.int64)), "box means differ"
        y = rp.forward_transform(x, d, tr)                                             # BasisChange.execute
        want_y = rp._blocks_view(np.real(y), d)
        want_v = rp._blocks_view(rp.quantize_prerounding(y, d, qn, qp), d)             # the value np.round receives
        if bitwise:
            # (-0.0 == 0.0: compare values, then the bit patterns of everything that is not a zero)
            for got, want, what in ((coeff[i], want_y, "transform"), (pre[i], want_v, "prerounding")):
                assert np.array_equal(got, want), what
                nz = want != 0
                assert np.array_equal(got[nz].view(np.int64), np.ascontiguousarray(want[nz]).view(np.int64)), what
        else:
            scale = np.abs(want_y).max() + 1.0
            assert np.abs(coeff[i] - want_y).max() <= 1e-12 * scale
            assert np.abs(pre[i] - want_v).max() <= 1e-12 * scale
        # and the integers of the product path are these values rounded, wherever they are not within an ulp of a tie
        q = jb.stages.forward_coefficients(stack[i], cfg)[0]
        zz = rp.to_zigzag(rp._from_blocks(np.round(pre[i])), d).reshape(q.shape)
        if qn == "discard":
            zz = rp.quantised_zigzag(a.astype(np.int64), ocfg).reshape(q.shape)
        assert np.array_equal(q.astype(np.int64), zz.astype(np.int64))


def test_float64_stage_rejects_bad_arguments(jb):
    cfg, _ = _cfgs(jb, (32, 32, 2, 8, "DCT", "none", None))
    with pytest.raises(KeyError):
        jb.stages.float64_stage(np.zeros((32, 32), np.uint8), cfg, "nonsense")


@pytest.mark.parametrize("h,w,bs,d,qn,qp,n_planes", [
    (2160, 3840, 5, 24, "divide", 1000, 30),      # config 3, ten frames: 17 280 blocks -> chunks of 8 (one plane: chunks of 2)
    (1000, 2000, 3, 16, "divide", 25, 4),         # 16 x 16 blocks, ragged edges: 3 528 blocks -> chunks of 2 both ways
    (1536, 2048, 2, 32, "divide", 64, 12),        # 32 x 32 blocks: 9 216 blocks -> chunks of 2
    (2048, 4096, 2, 16, "divide", 40, 3),         # 64 x 128 blocks x 3 planes = 24 576 blocks -> chunks of 8
])
def test_large_block_calls_of_either_chunk_size_agree(jb, h, w, bs, d, qn, qp, n_planes):
    """The large-block kernels deal chunks of 8 blocks in big calls and chunks of 2 in small ones
    (jb_call_chunk_blocks): a batch must give, plane for plane, the bytes and pixels of one-plane calls, and one plane
    the oracle's."""
    cfg, ocfg = _cfgs(jb, (h, w, bs, d, "DCT", qn, qp))
    base = [synth_plane(h, w, 300 + k) for k in range(2)]
    planes = [np.roll(base[k % 2], 17 * k, axis=1).astype(np.uint8) for k in range(n_planes)]
    batch = jb.compress_bands(planes, cfg)
    singles = [jb.compress_band(planes[k], cfg) for k in (0, 1, n_planes - 1)]
    assert [batch[0], batch[1], batch[n_planes - 1]] == singles
    rec = jb.decompress_bands(batch, cfg)
    for k in (0, n_planes - 1):
        assert np.array_equal(rec[k], jb.decompress_band(batch[k], cfg))
    coeffs = jb.stages.forward_coefficients(planes[0], cfg)[0]
    p64 = planes[0].astype(np.int64)
    check_quantised(coeffs, rp.quantised_zigzag(p64, ocfg), rp.prerounding_zigzag(p64, ocfg), what="chunks")
    assert batch[0] == rp.pack_blocks(coeffs.reshape(-1, d * d))
    check_pixels(rec[0], rp.decompress_band(batch[0], ocfg), p64, what="chunks")


def test_small_and_large_call_tails_agree_at_the_threshold(jb):
    """Calls of up to 2048 chunks end with one scan-and-gather kernel, larger ones with the scan and the gather
    kernels: 256 planes of 512 x 512 are exactly 2048 chunks, 257 planes one more plane's worth."""
    h = w = 512
    cfg, ocfg = _cfgs(jb, (h, w, 4, 8, "DCT", "qtable", None))
    base = [synth_plane(h, w, 500 + k).astype(np.uint8) for k in range(4)]
    planes = [np.roll(base[k % 4], 13 * k, axis=0) for k in range(257)]
    small = jb.compress_bands(planes[:256], cfg)          # 2048 chunks
    large = jb.compress_bands(planes, cfg)                # 2056 chunks
    assert large[:256] == small
    for k in (0, 255, 256):
        assert large[k] == jb.compress_band(planes[k], cfg)
    coeffs = jb.stages.forward_coefficients(planes[256], cfg)[0]
    p64 = planes[256].astype(np.int64)
    check_quantised(coeffs, rp.quantised_zigzag(p64, ocfg), rp.prerounding_zigzag(p64, ocfg), what="threshold")
    assert large[256] == rp.pack_blocks(coeffs.reshape(-1, 64))
    rec = jb.decompress_bands(large, cfg)
    check_pixels(rec[256], rp.decompress_band(large[256], ocfg), p64, what="threshold")
    assert np.array_equal(rec[:256], jb.decompress_bands(small, cfg))
