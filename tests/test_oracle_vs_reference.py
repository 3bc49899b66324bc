"""Live differential test: oracle vs the UNMODIFIED reference imported from
/root/reference.  Runs only in the dev container (the GPU box has no reference
tree); the committed golden vectors carry the same evidence elsewhere."""
import warnings

import numpy as np
import pytest

from oracle import ref_port as rp
from oracle.load_reference import load_reference, reference_available
from golden_inputs import synth_plane
from parity import check_quantised

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")

GRID = [
    (64, 96, 4, 8, "DCT", "qtable", None), (61, 75, 4, 8, "DCT", "qtable", None),
    (50, 70, 3, 8, "DCT", "none", None), (50, 70, 3, 8, "DFT", "none", None),
    (64, 64, 4, 8, "DFT", "qtable", None), (120, 130, 5, 24, "DCT", "divide", 40),
    (33, 47, 2, 4, "DCT", "discard", 2), (17, 19, 1, 2, "DCT", "divide", 129),
    (31, 29, 3, 5, "DFT", "divide", 7), (16, 16, 1, 1, "DCT", "none", None),
]


def _ref_cfg(P, h, w, bs, d, tr, qn, qp):
    kw = {"keep": qp} if qn == "discard" else {"divisor": qp} if qn == "divide" else {}
    return P.Configuration(width=w, height=h, block_size=bs, dct_size=d, transform=tr,
                           quantization=P.QuantizationMethod(qn, **kw))


@pytest.mark.parametrize("h,w,bs,d,tr,qn,qp", GRID)
@pytest.mark.parametrize("kind", ["synth", "uniform"])
def test_streams_and_reconstruction_match_reference(h, w, bs, d, tr, qn, qp, kind):
    warnings.simplefilter("ignore")
    P = load_reference().pipeline
    a = synth_plane(h, w, 5) if kind == "synth" else \
        np.random.default_rng(h * w).integers(0, 256, (h, w)).astype(np.int64)
    rc = _ref_cfg(P, h, w, bs, d, tr, qn, qp)
    oc = rp.OracleConfig(w, h, bs, d, tr, qn, qp)
    rb = P.compress_band(a.copy(), rc)
    ob = rp.compress_band(a, oc)
    if ob != rb:
        # only a float64 rounding tie may separate the two (see tests/parity.py)
        n = d * d
        ties = check_quantised(rp.unpack_stream(rb, n), rp.unpack_stream(ob, n),
                               rp.prerounding_zigzag(a, oc), what="reference vs oracle")
        assert ties > 0
    assert np.array_equal(rp.decompress_band(rb, oc), P.decompress_band(rb, rc))


def test_reference_own_unit_tests_pass_under_the_shims():
    """Oracle health check: the reference's 45 tests."""
    import os, subprocess, sys
    from oracle import load_reference as lr
    code = ("import sys; sys.path.insert(0, %r); from oracle.load_reference import load_reference; "
            "load_reference(); sys.path.insert(0, %r); import unittest; "
            "import os; os.chdir(%r); sys.path.insert(0, '.'); "
            "s = unittest.defaultTestLoader.loadTestsFromName('tests'); "
            "r = unittest.TextTestRunner(verbosity=0).run(s); "
            "print('RAN', r.testsRun, len(r.failures), len(r.errors))") % (
        os.path.dirname(os.path.dirname(os.path.abspath(lr.__file__))), lr.REFERENCE_ROOT,
        os.path.join(lr.REFERENCE_ROOT, "tests"))
    out = subprocess.run([sys.executable, "-W", "ignore", "-c", code], capture_output=True, text=True)
    assert "RAN 45 0 0" in out.stdout, out.stdout + out.stderr
