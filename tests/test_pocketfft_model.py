"""The float64 operation order the CUDA tie re-evaluation follows for the DFT option (csrc/jb_refine.cuh,
jb_pf_pass8) is the one numpy's pocketfft runs: the numpy statement of it (oracle/pocketfft8.py) must equal
np.fft.fft / np.fft.fft2 to the last bit."""
import numpy as np

from oracle import pocketfft8 as pf
from oracle import ref_port as rp


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint64)


def _same_bits(a, b):
    # +0.0 and -0.0 round alike; everything else must be identical
    a = np.ascontiguousarray(a) + 0.0
    b = np.ascontiguousarray(b) + 0.0
    return np.array_equal(_bits(a.real.copy()), _bits(b.real.copy())) and \
        np.array_equal(_bits(a.imag.copy()), _bits(b.imag.copy()))


def test_pass8_equals_numpy_fft_bitwise_on_random_complex_vectors():
    rng = np.random.default_rng(1)
    z = rng.normal(size=(50000, 8)) + 1j * rng.normal(size=(50000, 8))
    assert _same_bits(pf.fft8(z), np.fft.fft(z, axis=-1))


def test_fft2_equals_numpy_fft2_bitwise_on_box_means():
    # the values the reference transforms: means of 4x4 boxes of 8-bit samples, k / 16
    rng = np.random.default_rng(2)
    x = rng.integers(0, 4081, size=(40000, 8, 8)).astype(np.float64) / 16.0
    want = np.fft.fft2(x, axes=(-2, -1))
    got = pf.fft2_8x8(x)
    assert _same_bits(got, want)
    # smooth content: many coefficients near zero and many exact ties after the table quantiser
    y, xx = np.mgrid[0:8, 0:8]
    smooth = np.round(127 + 100 * np.sin((xx + rng.integers(0, 99, (5000, 1, 1))) / 9.0)
                      * np.cos((y + rng.integers(0, 99, (5000, 1, 1))) / 13.0)) * 16
    smooth = (smooth + rng.integers(-40, 41, smooth.shape)).clip(0, 4080) / 16.0
    assert _same_bits(pf.fft2_8x8(smooth), np.fft.fft2(smooth, axes=(-2, -1)))


def test_model_reproduces_the_oracle_quantised_dft_coefficients():
    # through the oracle's own stages: the model's real part, quantised, is the oracle's integer everywhere
    from golden_inputs import synth_plane
    a = synth_plane(256, 320, 5)
    cfg = rp.OracleConfig(320, 256, 4, 8, "DFT", "qtable")
    x = rp.subsample(a, 4, 8)
    blocks = x.reshape(x.shape[0] // 8, 8, x.shape[1] // 8, 8).transpose(0, 2, 1, 3)
    v = pf.fft2_8x8(blocks).real * (1.0 / rp.JPEG_LUMA_TABLE)[None, None]
    q = np.round(v).astype(np.int64)
    zz = q.reshape(q.shape[0], q.shape[1], 64)[:, :, rp.zigzag_flat_indices(8)]
    assert np.array_equal(zz, rp.quantised_zigzag(a, cfg).reshape(zz.shape))
