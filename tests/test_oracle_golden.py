"""Pin the CPU oracle (oracle/ref_port.py) against (1) every exact vector the
reference's own tests hold for this path and (2) committed outputs of the
unmodified reference (tests/golden/reference_cases.json)."""
import numpy as np
import pytest

import golden_io
from golden_inputs import make_input
from oracle import ref_port as rp


def _cfg(case):
    return rp.OracleConfig(case["w"], case["h"], case["bs"], case["d"], case["transform"],
                           case["qname"], case.get("qparam"))


@pytest.mark.parametrize("case", golden_io.cases(), ids=golden_io.case_ids())
def test_golden_stream_and_reconstruction(case):
    a = make_input(case)
    cfg = _cfg(case)
    if case.get("error"):
        with pytest.raises(rp.BadRleCodeError) as ei:
            rp.compress_band(a, cfg)
        assert str(ei.value) == case["error_msg"]
        return
    zz = rp.quantised_zigzag(a, cfg)
    assert np.array_equal(zz, golden_io.zigzag(case))
    data = rp.compress_band(a, cfg)
    assert data == golden_io.stream(case)
    rec = rp.decompress_band(golden_io.stream(case), cfg)
    assert rec.shape == (case["h"], case["w"])
    assert np.array_equal(rec, golden_io.restored(case))
    # per-block byte lengths add up to the stream
    assert int(rp.block_byte_lengths(zz).sum()) == len(data)


def test_survey_known_answer_streams():
    # SURVEY.md section 8(a) known-answer vectors (derived by running the reference)
    a = np.arange(64).reshape(8, 8)
    cfg = rp.OracleConfig(8, 8, 1, 8, "DCT", "qtable")
    zz = rp.quantised_zigzag(a, cfg)
    assert rp.rle_tuples(zz) == [(0, 8, 126), (0, 5, -9), (0, 8, -69), (3, 2, -1), (2, 4, -6),
                                 (10, 2, -1), (0, 0)]
    assert rp.compress_band(a, cfg).hex() == "08fe054842299248d44800"
    cfg = rp.OracleConfig(8, 8, 1, 8, "DCT", "none")
    assert rp.compress_band(a, cfg).hex() == "0cfe008670b6726ab28565368cd392c8c000"
    flat = np.full((8, 8), 255)
    assert rp.compress_band(flat, rp.OracleConfig(8, 8, 1, 8, "DCT", "qtable")).hex() == "0bff8000"
    assert rp.compress_band(flat, rp.OracleConfig(8, 8, 1, 8, "DCT", "none")).hex() == "0fff8000"
    flat24 = np.full((24, 24), 255)
    assert rp.compress_band(flat24, rp.OracleConfig(24, 24, 1, 24, "DCT", "divide", 1000)).hex() == "09c98000"
    assert rp.compress_band(flat24, rp.OracleConfig(24, 24, 1, 24, "DCT", "divide", 40)).hex() == "0df2c000"
    with pytest.raises(rp.BadRleCodeError) as ei:
        rp.compress_band(flat24, rp.OracleConfig(24, 24, 1, 24, "DCT", "none"))
    assert str(ei.value) == "(0, 19, 146880)"
    a = np.arange(128).reshape(8, 16)
    assert rp.compress_band(a, rp.OracleConfig(16, 8, 3, 8, "DCT", "none")).hex() == \
        "0ee40025be16f8c2d908bbd3b43214881aac15090880f994b036cfc000"
    dft = rp.compress_band(a, rp.OracleConfig(16, 8, 3, 8, "DFT", "none"))
    assert len(dft) == 34 and dft.hex().startswith("0ee400250416f1c2")
    m = np.array([[220, 255, 123, 205], [255, 255, 112, 10], [15, 51, 83, 221], [239, 73, 62, 22]])
    cfg = rp.OracleConfig(4, 4, 1, 2, "DCT", "divide", 129)
    s = rp.compress_band(m, cfg)
    assert s.hex() == "05c00003e258120003e0581204800003e04816048000"
    assert rp.decompress_band(s, cfg).tolist() == [[255, 255, 78, 207], [255, 255, 116, 0],
                                                   [32, 70, 32, 252], [252, 32, 70, 32]]


# ---- vectors from the reference's own unit tests -------------------------------------------

def test_zigzag_orders_reference_tests():
    # tests/zigzag_tests.py:11-32
    a = np.arange(16).reshape(4, 4)
    assert a.ravel()[rp.zigzag_flat_indices(4)].tolist() == \
        [0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15]
    a = np.arange(9).reshape(3, 3)
    assert a.ravel()[rp.zigzag_flat_indices(3)].tolist() == [0, 1, 3, 6, 4, 2, 5, 7, 8]
    # tests/zigzag_tests.py:61-73 (stage on 2x2 blocks)
    a = np.arange(16).reshape(4, 4)
    assert rp.to_zigzag(a, 2).tolist() == [[[0, 1, 4, 5], [2, 3, 6, 7]],
                                           [[8, 9, 12, 13], [10, 11, 14, 15]]]
    # tests/zigzag_tests.py:75-96 (restore, also complex)
    a = np.arange(32).reshape(4, 8)
    assert np.array_equal(rp.from_zigzag(rp.to_zigzag(a, 2), 2), a)
    assert np.array_equal(rp.from_zigzag(rp.to_zigzag(a * 2j, 2), 2), a * 2j)
    # N=8 is the JPEG natural order
    jpeg = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34,
            27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37,
            44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]
    assert rp.zigzag_flat_indices(8).tolist() == jpeg


def test_rle_tuples_reference_tests():
    # tests/RLE_tests.py:12-26
    a = np.array([[-15, 0, 0, 0, 3, 2, 0, 0, 0, 0, 120, 0, 0, 0, 0]])
    assert rp.rle_tuples(a) == [(0, 5, -15), (3, 3, 3), (0, 3, 2), (4, 8, 120), (0, 0)]
    # :36-46
    a = np.array([[0, 2] + [0] * 32 + [5] + [0] * 5])
    t = rp.rle_tuples(a)
    assert t == [(1, 3, 2), (15, 0, 0), (15, 0, 0), (2, 4, 5), (0, 0)]
    assert rp.tuples_to_blocks(t, a.shape[1]).tolist() == a.tolist()
    # :51-62
    assert rp.rle_tuples(np.zeros((1, 9))) == [(0, 0)]
    # :66-86
    blocks = np.zeros((3, 1, 9))
    blocks[0, 0] = [21, 3, 0, 0, 0, 0, 2, 0, 0]
    blocks[1, 0] = [0, 0, 0, 15, 0, 0, 0, 0, 9]
    expected = [(0, 6, 21), (0, 3, 3), (4, 3, 2), (0, 0), (3, 5, 15), (4, 5, 9), (0, 0), (0, 0)]
    assert rp.rle_tuples(blocks) == expected
    assert rp.tuples_to_blocks(expected, 9).tolist() == blocks.reshape(3, 9).tolist()


def _bits(b):
    return "".join(format(x, "08b") for x in b)


def test_bitstrings_reference_tests():
    # tests/RLE_tests.py:99-122
    assert _bits(rp.pack_tuples([(4, 3, 2), (0, 0)])) == "0100" + "0011" + "110" + "0" * 13
    assert _bits(rp.pack_tuples([(15, 0, 0), (0, 0)])) == "1111" + "0000" + "0" * 8
    # round trips :124-138,182-196
    for x in ([(15, 0, 0), (15, 0, 0), (0, 2, 1), (0, 0)],
              [(1, 2, -1), (0, 3, -2), (8, 3, -3), (8, 5, -15), (0, 0)],
              [(14, 4, 7), (0, 0)],
              [(14, 4, 7), (0, 0), (0, 0), (15, 0, 0), (0, 2, 1), (0, 0)]):
        blocks = rp.unpack_stream(rp.pack_tuples(x), 64)
        assert rp.rle_tuples(blocks) == x
        assert rp.pack_blocks(blocks) == rp.pack_tuples(x)


@pytest.mark.parametrize("bad", [[(15, 0, 1), (0, 0)], [(15, 0, -10), (0, 0)], [(16, 3, 3), (0, 0)],
                                 [(-1, 3, 3), (0, 0)], [(10, 16, 0), (0, 0)], [(4, -1, 0), (0, 0)],
                                 [(40, -18, 0), (0, 0)], [(12, 0, 0), (0, 0)]])
def test_bad_codes_reference_tests(bad):
    # tests/RLE_tests.py:140-180
    with pytest.raises(rp.BadRleCodeError):
        rp.pack_tuples(bad)


def test_subsample_and_padding_reference_tests():
    # tests/subsample_tests.py:11-32
    a = np.array([[1, 2, 2, 1], [3, 2, 8, 1], [0, 0, 2, 2], [0, 4, 2, 2]])
    assert rp.subsample(a, 2, 1).tolist() == [[2, 3], [1, 2]]
    assert rp.subsample(a, 4, 1).tolist() == [[2]]
    # tests/padding_tests.py:29-54, tests/util_tests.py:9-17
    assert rp._edge_pad(np.array([[20], [10]]), 3).tolist() == [[20, 20, 20], [10, 10, 10], [10, 10, 10]]
    assert rp._edge_pad(np.array([[20, 3], [10, 9]]), 2).tolist() == [[20, 3], [10, 9]]
    assert [rp.padded_size(s, 3) for s in (3, 4, 5, 6, 7)] == [3, 6, 6, 6, 9]
    with pytest.raises(rp.BadArrayShapeError):
        rp._edge_pad(np.array([32, 31]), 2)
    with pytest.raises(rp.BadArrayShapeError):
        rp._edge_pad(np.array([[[32]]]), 2)
    with pytest.raises(rp.EmptyArrayError):
        rp._edge_pad(np.array([[]]), 3)


def test_quantisers_reference_tests():
    # tests/quantization_tests.py:10-54
    assert rp.quantize(np.array([[3.4, 8.], [0, 0.6]]), 2, "none").tolist() == [[3, 8], [0, 1]]
    q = rp.quantize(np.arange(9).reshape(3, 3).astype(float), 3, "discard", 2)
    assert q.tolist() == [[0, 1, 0], [3, 4, 0], [0, 0, 0]]
    q = rp.quantize(np.array([[80., 24., 169.]]).reshape(1, 3), 1, "divide", 40)
    assert q.tolist() == [[2, 1, 4]]
    assert rp.dequantize(q, 1, "divide", 40).tolist() == [[80, 40, 160]]
    with pytest.raises(rp.BadQuantizationError):
        rp.OracleConfig(8, 8, 1, 4, "DCT", "qtable")
    with pytest.raises(rp.BadQuantizationError):
        rp.OracleConfig(8, 8, 1, 8, "DCT", "bogus")


def test_dct_round_trips_reference_tests():
    # tests/basis_change_tests.py:9-38 (round trip only) and
    # tests/integration_tests.py:50-66 (lossless on integers)
    for n in (2, 8):
        a = np.arange(n * n).reshape(n, n)
        y = rp.forward_transform(a.astype(float), n, "DCT")
        assert np.array_equal(rp.inverse_transform(y, n, "DCT"), a)
    a = np.arange(6).reshape(2, 3)
    cfg = rp.OracleConfig(3, 2, 1, 8, "DCT", "none")
    assert np.array_equal(rp.decompress_band(rp.compress_band(a, cfg), cfg), a)
    a = np.arange(64).reshape(8, 8)
    cfg = rp.OracleConfig(8, 8, 1, 1, "DCT", "none")
    assert np.array_equal(rp.decompress_band(rp.compress_band(a, cfg), cfg), a)
    # DC of the un-normalised DCT is the plain sum
    y = rp.forward_transform(np.full((8, 8), 255.0), 8, "DCT")
    assert y[0, 0] == 255 * 64
    # DFT keeps only the real part, which equals Cc X Cc - S X S
    x = np.random.default_rng(0).integers(0, 256, (8, 8)).astype(float)
    k = np.arange(8)
    cc = np.cos(2 * np.pi * np.outer(k, k) / 8)
    ss = np.sin(2 * np.pi * np.outer(k, k) / 8)
    assert np.allclose(np.real(rp.forward_transform(x, 8, "DFT")), cc @ x @ cc - ss @ x @ ss, atol=1e-9)
