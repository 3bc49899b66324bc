"""The per-thread bit-level device code (csrc/jb_bits.h: RLE packing, block extent parse,
block decode) compiled for the host and checked against the oracle -- no GPU needed."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import ref_port as rp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "implementing-jpeg-compression_b200", "csrc")
LIB = os.path.join(ROOT, "implementing-jpeg-compression_b200", "libjbhostcheck.so")


@pytest.fixture(scope="module")
def hc():
    src = os.path.join(CSRC, "host_check.cpp")
    hdr = os.path.join(CSRC, "jb_bits.h")
    if (not os.path.exists(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, src])
    lib = ctypes.CDLL(LIB)
    lib.hc_pack_blocks.restype = ctypes.c_longlong
    lib.hc_walk_stream.restype = ctypes.c_longlong
    return lib


def _pack(hc, zz):
    zz = np.ascontiguousarray(zz, dtype=np.int32)
    nb, n = zz.shape
    out = np.zeros(nb * ((23 * n + 15) // 8) + 8, dtype=np.uint8)
    lens = np.zeros(nb, dtype=np.uint32)
    bp, br = ctypes.c_int(-1), ctypes.c_int(0)
    total = hc.hc_pack_blocks(zz.ctypes.data_as(ctypes.c_void_p), nb, n, out.ctypes.data_as(ctypes.c_void_p),
                              lens.ctypes.data_as(ctypes.c_void_p), ctypes.byref(bp), ctypes.byref(br))
    return total, out, lens, bp.value, br.value


def _random_blocks(rng, nb, n, density, amp_bits):
    zz = np.zeros((nb, n), dtype=np.int64)
    mask = rng.random((nb, n)) < density
    mags = rng.integers(1, 1 << amp_bits, (nb, n))
    sign = rng.integers(0, 2, (nb, n)) * 2 - 1
    zz[mask] = (mags * sign)[mask]
    return zz


@pytest.mark.parametrize("n,density,amp_bits", [(64, 0.2, 6), (64, 0.02, 14), (64, 1.0, 14), (64, 0.0, 1),
                                                (576, 0.05, 10), (1, 0.5, 8), (4, 0.5, 3), (1024, 0.01, 12),
                                                (9, 0.3, 5)])
def test_pack_walk_decode_match_oracle(hc, n, density, amp_bits):
    rng = np.random.default_rng(n * 1000 + amp_bits)
    zz = _random_blocks(rng, 300, n, density, amp_bits)
    expect = rp.pack_blocks(zz)
    total, out, lens, bp, _ = _pack(hc, zz)
    assert bp == -1
    assert total == len(expect)
    assert out[:total].tobytes() == expect
    assert lens.tolist() == rp.block_byte_lengths(zz).tolist()
    # walking the stream with the extent parser finds exactly the block starts
    starts = np.zeros(zz.shape[0] + 4, dtype=np.uint32)
    k = hc.hc_walk_stream(out.ctypes.data_as(ctypes.c_void_p), total, n, starts.ctypes.data_as(ctypes.c_void_p),
                          starts.size)
    assert k == zz.shape[0]
    assert starts[:k].tolist() == (np.cumsum(lens) - lens).tolist()
    # every block ends with a zero byte: the framing kernels rely on it
    ends = np.cumsum(lens) - 1
    assert np.all(out[ends] == 0)
    # decode each block
    dec = np.zeros(n, dtype=np.int32)
    for b in range(0, zz.shape[0], 7):
        rc = hc.hc_decode_block(out.ctypes.data_as(ctypes.c_void_p), int(starts[b]), total, n,
                                dec.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0
        assert dec.tolist() == zz[b].tolist()


def test_reference_bit_strings(hc):
    # tests/RLE_tests.py:99-122 in the reference
    zz = np.zeros((1, 64), dtype=np.int64)
    zz[0, 4] = 2
    total, out, _, _, _ = _pack(hc, zz)
    bits = "".join(format(x, "08b") for x in out[:total])
    assert bits == "0100" + "0011" + "110" + "0" * 13
    zz = np.zeros((1, 64), dtype=np.int64)
    zz[0, 32] = 1                       # two zero chains then (2, 2, 1)
    total, out, _, _, _ = _pack(hc, zz)
    assert out[:total].tobytes() == rp.pack_tuples([(15, 0, 0), (15, 0, 0), (2, 2, 1), (0, 0)])


def test_amplitude_overflow_is_reported(hc):
    zz = np.zeros((3, 576), dtype=np.int64)
    zz[1, 0] = 146880
    total, _, _, bp, br = _pack(hc, zz)
    assert total == -2 and bp == 0 and br == 0
    zz = np.zeros((2, 64), dtype=np.int64)
    zz[0, 20] = -16384
    total, _, _, bp, br = _pack(hc, zz)
    assert total == -1 and bp == 20 and br == 20 % 15
    zz[0, 20] = -16383
    total, _, _, bp, _ = _pack(hc, zz)
    assert total > 0 and bp == -1


def test_extent_parser_rejects_malformed(hc):
    end = ctypes.c_uint32(0)
    def parse(data, start=0, n=64):
        buf = np.frombuffer(bytes(data), dtype=np.uint8).copy()
        return hc.hc_parse_extent(buf.ctypes.data_as(ctypes.c_void_p), start, len(buf), n, ctypes.byref(end))
    assert parse(b"\x00") == 0 and end.value == 1
    assert parse(b"\x50\x00") == 1                    # (5, 0, 0): util.py:176-177
    assert parse(b"\x11\x00") == 1                    # size 1: no magnitude bits
    assert parse(b"\xf0" * 5 + b"\x00") == 1          # 75 zeros > 64 coefficients
    assert parse(b"\xf0" * 4 + b"\x00") == 0 and end.value == 5
    assert parse(b"\x08\xfe") == 1                    # runs off the end, no EOB
    good = rp.pack_tuples([(0, 8, 126), (0, 5, -9), (0, 0)])
    assert parse(good) == 0 and end.value == len(good)
