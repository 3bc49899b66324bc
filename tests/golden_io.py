"""Loader for tests/golden/reference_cases.json (outputs of the unmodified reference)."""
import base64
import json
import os
import zlib

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_cases.json")
_cache = None


def load():
    global _cache
    if _cache is None:
        with open(_PATH) as f:
            _cache = json.load(f)
    return _cache


def unz(s):
    return zlib.decompress(base64.b64decode(s))


def cases():
    return load()["cases"]


def case_ids():
    return [c["name"] for c in cases()]


def stream(case):
    return unz(case["stream"])


def zigzag(case):
    return np.frombuffer(unz(case["zigzag_i32"]), dtype="<i4").astype(np.int64).reshape(case["zigzag_shape"])


def restored(case):
    return np.frombuffer(unz(case["restored_u8"]), dtype=np.uint8).reshape(case["h"], case["w"]).astype(np.int64)
