"""Run inside a fresh interpreter by tests/test_reference_dropin.py (the reference's top-level module names --
``pipeline``, ``util``, ``file_format`` ... -- must not leak into the pytest process): the UNMODIFIED reference
side by side with the CUDA path.  Prints one JSON object.

1. cross-decode: streams / containers written by the CUDA path, decoded by the stock reference
   (pipeline.decompress_band, Jpeg.decompress -- decompress.py:5-10), and the reference's own streams decoded
   by the CUDA path;
2. the swap of INTEGRATION.md: integration/cuda_path.py installed over pipeline.compress_band /
   decompress_band, then the reference's own tests/integration_tests.py and its Jpeg facade run over it.
"""
import io
import json
import os
import sys
import unittest
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
warnings.simplefilter("ignore")

import numpy as np  # noqa: E402

from oracle import load_reference as lr  # noqa: E402
from golden_inputs import synth_plane  # noqa: E402


def main():
    res = {}
    ref = lr.load_reference()
    P = ref.pipeline
    stock_compress, stock_decompress, StockJpeg = P.compress_band, P.decompress_band, P.Jpeg
    import jpeg_b200 as jb
    from PIL import Image

    # ---- 1. cross-decode --------------------------------------------------------------------------------
    cross = []
    for (h, w, bs, d, tr, qn, qp) in [(96, 128, 4, 8, "DCT", "qtable", None), (61, 75, 4, 8, "DCT", "qtable", None),
                                      (64, 64, 4, 8, "DFT", "qtable", None), (120, 130, 5, 24, "DCT", "divide", 1000),
                                      (50, 70, 3, 8, "DCT", "none", None)]:
        kw = {"divisor": qp} if qn == "divide" else {}
        rcfg = P.Configuration(width=w, height=h, block_size=bs, dct_size=d, transform=tr,
                               quantization=P.QuantizationMethod(qn, **kw))
        gcfg = jb.Configuration(width=w, height=h, block_size=bs, dct_size=d, transform=tr,
                                quantization=jb.QuantizationMethod(qn, **kw))
        a = synth_plane(h, w, 3 * h + w)
        g_stream = jb.compress_band(a, gcfg)
        r_stream = stock_compress(a.copy(), rcfg)
        by_stock = np.asarray(stock_decompress(g_stream, rcfg)).reshape(h, w)        # GPU stream, stock decoder
        by_gpu = jb.decompress_band(g_stream, gcfg)
        r_by_gpu = jb.decompress_band(r_stream, gcfg)                                # stock stream, GPU decoder
        r_by_stock = np.asarray(stock_decompress(r_stream, rcfg)).reshape(h, w)
        cross.append({"case": [h, w, bs, d, tr, qn], "streams_equal": g_stream == r_stream,
                      "gpu_stream_stock_vs_gpu_decoder_maxdiff": int(np.abs(by_stock - by_gpu).max()),
                      "stock_stream_stock_vs_gpu_decoder_maxdiff": int(np.abs(r_by_stock - r_by_gpu).max())})
    res["cross_decode"] = cross

    # the container through the stock Jpeg.decompress (what decompress.py calls)
    rng = np.random.default_rng(4)
    rgb = np.clip(np.stack([synth_plane(72, 104, 50 + i) for i in range(3)], -1) + rng.integers(-3, 4, (72, 104, 3)),
                  0, 255).astype(np.uint8)
    im = Image.fromarray(rgb, "RGB").convert("YCbCr")
    gcfg = jb.Configuration(width=104, height=72, block_size=4, dct_size=8, transform="DCT",
                            quantization=jb.QuantizationMethod("qtable"))
    rcfg = P.Configuration(width=104, height=72, block_size=4, dct_size=8, transform="DCT",
                           quantization=P.QuantizationMethod("qtable"))
    blob_gpu = jb.Jpeg(gcfg).compress(im)
    blob_ref = StockJpeg(rcfg).compress(im)
    dec_stock = np.asarray(StockJpeg.decompress(blob_gpu)).astype(np.int64)
    dec_gpu = np.asarray(jb.Jpeg.decompress(blob_gpu)).astype(np.int64)
    res["container"] = {"bytes_equal": blob_gpu == blob_ref,
                        "stock_vs_gpu_decode_maxdiff": int(np.abs(dec_stock - dec_gpu).max())}

    # ---- 2. the swap ------------------------------------------------------------------------------------
    sys.path.insert(0, lr.REFERENCE_ROOT)                     # `util`, `pipeline` of the stub's imports
    sys.path.insert(0, os.path.join(ROOT, "integration"))
    import cuda_path
    cuda_path.install(P)
    os.chdir(os.path.join(lr.REFERENCE_ROOT, "tests"))
    sys.path.insert(0, ".")
    import integration_tests                                   # binds the swapped functions at import
    assert integration_tests.compress_band is cuda_path.compress_band
    suite = unittest.defaultTestLoader.loadTestsFromModule(integration_tests)
    out = io.StringIO()
    r = unittest.TextTestRunner(stream=out, verbosity=0).run(suite)
    res["integration_tests"] = {"ran": r.testsRun, "failures": len(r.failures), "errors": len(r.errors),
                                "log": out.getvalue()[-2000:] if (r.failures or r.errors) else ""}
    # the reference's own facade and container on the swapped functions (compress.py:16-17, decompress.py:9)
    blob_swapped = P.Jpeg(rcfg).compress(im)
    dec_swapped = np.asarray(P.Jpeg.decompress(blob_swapped)).astype(np.int64)
    res["swapped_facade"] = {"bytes_equal_stock": blob_swapped == blob_ref,
                             "decode_vs_stock_maxdiff": int(np.abs(dec_swapped - dec_stock).max())}
    # error behaviour through the swap: amplitude too large for the 15-bit size field (util.py:170-171)
    try:
        P.compress_band(np.full((24, 24), 255), P.Configuration(width=24, height=24, block_size=1, dct_size=24))
        res["bad_rle"] = "no error"
    except ref.util.BadRleCodeError as e:
        res["bad_rle"] = str(e)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
