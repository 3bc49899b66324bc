"""CPU-only tests: host logic, the C-ABI surface, the container, the sharding helpers."""
import ctypes
import os
import re

import numpy as np
import pytest

import golden_io
import jpeg_b200 as jb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = jb._lib.load()
    header = open(os.path.join(ROOT, "include", "jpegb200.h")).read()
    declared = set(re.findall(r"^(?:int|size_t|const char\*|unsigned long long)\s+(jb_[a-z_0-9]+)\s*\(", header, flags=re.M))
    assert declared == set(jb._lib.SYMBOLS), declared ^ set(jb._lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.jb_version() == 100
    assert jb._lib.strerror(-8).startswith("amplitude")


def test_geometry_matches_the_oracle():
    from oracle.ref_port import Geometry
    for (h, w, bs, d) in [(512, 512, 4, 8), (2160, 3840, 4, 8), (2160, 3840, 5, 24), (1080, 1920, 4, 8),
                          (16384, 16384, 4, 8), (1, 1, 4, 8), (61, 75, 3, 5)]:
        g = jb.geometry(jb.Configuration(width=w, height=h, block_size=bs, dct_size=d))
        o = Geometry(h, w, bs, d)
        assert (g["h1"], g["w1"], g["h2"], g["w2"], g["vb"], g["hb"], g["blocks_per_plane"]) == \
            (o.h1, o.w1, o.h2, o.w2, o.vb, o.hb, o.nblocks)
        assert g["max_block_bytes"] == (23 * d * d + 8 + 7) // 8
    # SURVEY.md section 8 table
    assert jb.geometry(jb.Configuration(3840, 2160, 4, 8))["blocks_per_plane"] == 8160
    assert jb.geometry(jb.Configuration(3840, 2160, 5, 24))["blocks_per_plane"] == 576
    assert jb.geometry(jb.Configuration(1920, 1080, 4, 8))["blocks_per_plane"] == 2040


def test_parameter_validation_without_a_gpu():
    lib = jb._lib.load()
    g = jb._lib.jb_geometry()

    def rc(**kw):
        base = dict(height=8, width=8, block_size=1, dct_size=8, transform=0, qmode=0, qparam=0, flags=0)
        base.update(kw)
        p = jb._lib.jb_params(**base)
        return lib.jb_geometry_of(ctypes.byref(p), ctypes.byref(g))

    assert rc() == 0
    assert rc(qmode=3, dct_size=4) == jb._lib.JB_ERR_BAD_QUANTIZATION
    assert rc(qmode=7) == jb._lib.JB_ERR_BAD_QUANTIZATION
    assert rc(height=0) == jb._lib.JB_ERR_EMPTY_ARRAY
    assert rc(dct_size=33) == jb._lib.JB_ERR_UNSUPPORTED
    assert rc(block_size=0) == jb._lib.JB_ERR_BAD_PARAM
    assert rc(transform=2) == jb._lib.JB_ERR_BAD_PARAM
    assert rc(qmode=2, qparam=0) == jb._lib.JB_ERR_BAD_QUANTIZATION
    p = jb._lib.jb_params(1080, 1920, 4, 8, 0, 3, 0, 0)
    assert lib.jb_max_stream_bytes(ctypes.byref(p), 3) == 3 * 2040 * 185
    assert lib.jb_compress_workspace_bytes(ctypes.byref(p), 3) > 0
    assert lib.jb_decompress_workspace_bytes(ctypes.byref(p), 3, 100000) > 0


def test_quantization_method_and_configuration_mirror_the_reference():
    # pipeline/__init__.py:13-64
    q = jb.QuantizationMethod("divide", divisor=93)
    assert q.to_json() == '{"divisor": 93, "quantization_scheme_name": "divide"}'
    q2 = jb.QuantizationMethod.from_json(q.to_json())
    assert (q2.name, q2.params) == ("divide", {"divisor": 93})
    assert jb.QuantizationMethod("qtable").to_json() == '{"quantization_scheme_name": "qtable"}'
    assert jb.QuantizationMethod("discard").param == 2 and jb.QuantizationMethod("divide").param == 40
    with pytest.raises(jb.BadQuantizationError):
        jb.QuantizationMethod("bogus")
    with pytest.raises(jb.BadQuantizationError):
        jb.QuantizationMethod("qtable", divisor=3)
    with pytest.raises(jb.BadQuantizationError):
        jb.Configuration(8, 8, dct_size=4, quantization=jb.QuantizationMethod("qtable"))
    c = jb.Configuration(width=10, height=20)
    assert (c.block_size, c.dct_size, c.transform, c.quantization.name) == (2, 8, "DCT", "none")
    with pytest.raises(UnboundLocalError):
        jb.Configuration(8, 8, transform="XYZ").c_params()


def test_container_is_byte_identical_to_the_reference():
    g = golden_io.load()
    cfg = jb.Configuration(512, 512, 4, 8, "DCT", jb.QuantizationMethod("qtable"))
    assert jb.file_format.create_header(cfg).hex() == g["header_512_defaults_hex"]
    cfg = jb.Configuration(320, 400, 44, 16, "DFT", jb.QuantizationMethod("divide", divisor=93))
    assert jb.file_format.create_header(cfg).hex() == g["header_divide_hex"]
    blob = golden_io.unz(g["container"]["blob"])
    config, data = jb.file_format.read_data(blob)
    assert jb.file_format.create_header(config).hex() == g["container"]["header_hex"]
    assert jb.file_format.generate_data(config, data) == blob
    # tests/file_format_tests.py:36-56 of the reference
    q = jb.QuantizationMethod("divide", divisor=93)
    config = jb.Configuration(width=320, height=400, block_size=44, dct_size=16, transform="DCT", quantization=q)
    d = jb.CompressedData(y=bytes([4, 8, 15, 16, 23, 42]), cb=bytes([1, 2, 3, 4, 5]), cr=bytes([10]))
    rc, rd = jb.file_format.read_data(jb.file_format.generate_data(config, d))
    assert (rc.dct_size, rd.y, rd.cb, rd.cr) == (16, d.y, d.cb, d.cr)
    assert rc.quantization.params == {"divisor": 93}


def test_sharding_helpers():
    s = jb.sharding
    assert [s.image_slice(1024, r, 8) for r in (0, 7)] == [(0, 128), (896, 1024)]
    cover = [s.image_slice(10, r, 4) for r in range(4)]
    assert cover == [(0, 3), (3, 6), (6, 8), (8, 10)]
    bands = s.block_row_bands(16384, 4, 8, 8)
    assert bands[0] == (0, 2048) and bands[-1] == (14336, 16384)
    bands = s.block_row_bands(1080, 4, 8, 8)           # 34 block rows, last one ragged
    assert bands[-1][1] == 1080 and all(b[0] % 32 == 0 for b in bands)
    assert sum(b[1] - b[0] for b in bands) == 1080
    bands = s.block_row_bands(40, 4, 8, 4)             # 2 block rows over 4 ranks
    assert bands == [(0, 32), (32, 40), (40, 40), (40, 40)]
    assert s.concat_band_streams([[b"a", b"b"], [], [b"c", b"d"]]) == [b"ac", b"bd"]


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = jb.Configuration(8, 8, 1, 8)
    with pytest.raises(jb.NativeLibraryError):
        jb.compress_band(np.zeros((8, 8), dtype=np.uint8), cfg)
    with pytest.raises(jb.NativeLibraryError):
        jb.decompress_band(b"\0", cfg)


def test_bind_host_to_device_is_a_no_op_without_nvml_or_gpu():
    """sharding.bind_host_to_device narrows the CPU affinity to the cores next to a GPU; where NVML or the device
    is missing it must change nothing and say so (bench.py calls it unconditionally)."""
    import os
    import torch
    before = os.sched_getaffinity(0)
    got = jb.sharding.bind_host_to_device(0)
    after = os.sched_getaffinity(0)
    if not torch.cuda.is_available():
        assert got is None and after == before
    else:
        assert got is None or (set(got) == after and after <= before)
        os.sched_setaffinity(0, before)
