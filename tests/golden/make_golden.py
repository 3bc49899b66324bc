"""Generate golden vectors by running the UNMODIFIED reference (dev container only).

Usage (from the repo root, in the container that has /root/reference):

    python tests/golden/make_golden.py

Writes ``tests/golden/reference_cases.json`` -- for a grid of seeded inputs and
codec configurations: the exact byte stream ``pipeline.compress_band`` produced,
the plane ``pipeline.decompress_band`` reconstructed from it, the per-block
zigzagged quantised integers (from running the reference's stages 0-6 and the
RLE stage's int cast), and a container produced by ``file_format.generate_data``.

The reference needs two import shims here (see ``oracle/load_reference.py``);
its arithmetic is untouched.  The inputs are regenerated from the recorded
seeds by ``tests/golden_inputs.py`` so that only outputs need to be stored.
"""
import base64
import json
import os
import sys
import warnings
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.load_reference import load_reference  # noqa: E402
from golden_inputs import CASES, make_input         # noqa: E402


def _b64z(raw):
    return base64.b64encode(zlib.compress(bytes(raw), 9)).decode("ascii")


def main():
    warnings.simplefilter("ignore")
    ref = load_reference()
    P = ref.pipeline
    out = {"generator": "tests/golden/make_golden.py",
           "reference": "X-rayLaser/Implementing-JPEG-compression (unmodified, "
                        "imported with the bitarray stand-in and numpy aliases)",
           "numpy": np.__version__, "cases": []}

    for case in CASES:
        a = make_input(case)
        h, w = a.shape
        qname, qparam = case["qname"], case.get("qparam")
        kw = {}
        if qname == "discard":
            kw = {"keep": qparam}
        elif qname == "divide":
            kw = {"divisor": qparam}
        cfg = P.Configuration(width=w, height=h, block_size=case["bs"],
                              dct_size=case["d"], transform=case["transform"],
                              quantization=P.QuantizationMethod(qname, **kw))
        rec = dict(case)
        try:
            # stages 0..6 one by one to capture the quantised, zigzagged ints
            x = a.copy()
            steps = list(P.step_classes)
            for cls in steps[:7]:
                x = cls(cfg).execute(x)
            zz = np.zeros(x.shape, dtype=np.int64)
            zz[:] = np.round(np.real(x))     # the int cast of RunLengthBlock.encode
            stream = P.compress_band(a.copy(), cfg)
            restored = P.decompress_band(stream, cfg)
            rec["stream_len"] = len(stream)
            rec["stream"] = _b64z(stream)
            rec["zigzag_shape"] = list(zz.shape)
            rec["zigzag_i32"] = _b64z(zz.astype("<i4").tobytes())
            rec["restored_u8"] = _b64z(np.asarray(restored).astype(np.uint8).tobytes())
            rec["restored_min"] = int(np.min(restored))
            rec["restored_max"] = int(np.max(restored))
        except ref.util.BadRleCodeError as e:
            rec["error"] = "BadRleCodeError"
            rec["error_msg"] = str(e)
        out["cases"].append(rec)
        print(case["name"], rec.get("stream_len"), rec.get("error"))

    # one container (header + three length-prefixed bands), file_format.py:67-93
    case = CASES[0]
    a = make_input(case)
    h, w = a.shape
    cfg = P.Configuration(width=w, height=h, block_size=case["bs"], dct_size=case["d"],
                          transform=case["transform"],
                          quantization=P.QuantizationMethod("qtable"))
    planes = [a, np.flipud(a).copy(), np.fliplr(a).copy()]
    bands = [P.compress_band(p.copy(), cfg) for p in planes]
    blob = ref.file_format.generate_data(cfg, P.CompressedData(*bands))
    out["container"] = {"case": case["name"], "planes": ["a", "flipud(a)", "fliplr(a)"],
                        "blob": _b64z(blob), "blob_len": len(blob),
                        "header_hex": ref.file_format.create_header(cfg).hex()}
    # header for the CLI defaults on 512x512
    cfg512 = P.Configuration(width=512, height=512, block_size=4, dct_size=8,
                             transform="DCT", quantization=P.QuantizationMethod("qtable"))
    out["header_512_defaults_hex"] = ref.file_format.create_header(cfg512).hex()
    cfgdiv = P.Configuration(width=320, height=400, block_size=44, dct_size=16,
                             transform="DFT",
                             quantization=P.QuantizationMethod("divide", divisor=93))
    out["header_divide_hex"] = ref.file_format.create_header(cfgdiv).hex()

    path = os.path.join(HERE, "reference_cases.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
