"""Operation-by-operation model of the length-8 complex FFT numpy runs for the reference's "DFT" option.

TEST INFRASTRUCTURE ONLY (see oracle/ref_port.py).  The reference calls ``np.fft.fft2`` per 8x8 block
(pipeline/basis_change.py:20-25); numpy >= 2.0 executes it with the C++ pocketfft it bundles (not under
/root/reference: third-party, version = the numpy of this image, 2.3.x): a length-8 transform is a single
radix-8 pass (``cfftp::pass8``, forward, ido = l1 = 1) along the last axis and then along the first.  Where a
coefficient lands exactly on a rounding tie depends on the order of these float64 operations, so the CUDA
path re-evaluates near-tie coefficients with the very same sequence (csrc/jb_refine.cuh: jb_pf_pass8).  This
module states the sequence in numpy; tests/test_pocketfft_model.py pins it bitwise against ``np.fft.fft`` /
``np.fft.fft2``, which is what ties the CUDA code to the reference's arithmetic.
"""
import numpy as np

HSQT2 = np.float64(0.707106781186547524400844362104849)


def _rot90(a):          # forward ROTX90: a * (-i)
    return a[1], -a[0]


def pass8_forward(c):
    """c: sequence of 8 (re, im) pairs of float64 arrays; returns the 8 outputs, same layout."""
    add = lambda p, q: (p[0] + q[0], p[1] + q[1])
    sub = lambda p, q: (p[0] - q[0], p[1] - q[1])
    a1, a5 = add(c[1], c[5]), sub(c[1], c[5])
    a3, a7 = add(c[3], c[7]), sub(c[3], c[7])
    a1, a3 = add(a1, a3), sub(a1, a3)
    a3 = _rot90(a3)
    a7 = _rot90(a7)
    a5, a7 = add(a5, a7), sub(a5, a7)
    a5 = (HSQT2 * (a5[0] + a5[1]), HSQT2 * (a5[1] - a5[0]))            # ROTX45
    a7 = (HSQT2 * (a7[1] - a7[0]), HSQT2 * (-a7[0] - a7[1]))           # ROTX135
    a0, a4 = add(c[0], c[4]), sub(c[0], c[4])
    a2, a6 = add(c[2], c[6]), sub(c[2], c[6])
    ch = [None] * 8
    s = add(a0, a2); ch[0], ch[4] = add(s, a1), sub(s, a1)
    s = sub(a0, a2); ch[2], ch[6] = add(s, a3), sub(s, a3)
    a6 = _rot90(a6)
    s = add(a4, a6); ch[1], ch[5] = add(s, a5), sub(s, a5)
    s = sub(a4, a6); ch[3], ch[7] = add(s, a7), sub(s, a7)
    return ch


def fft8(z):
    """np.fft.fft along the last axis (length 8) of a complex128 array, via the model."""
    z = np.asarray(z, dtype=np.complex128)
    out = pass8_forward([(z[..., k].real.copy(), z[..., k].imag.copy()) for k in range(8)])
    res = np.empty(z.shape, dtype=np.complex128)
    for k in range(8):
        res[..., k] = out[k][0] + 1j * out[k][1]
    return res


def fft2_8x8(x):
    """np.fft.fft2 over the last two axes (8 x 8) of a real or complex array, via the model:
    last axis first, then the first (numpy/fft/_pocketfft.py: _raw_fftnd walks the axes in reverse)."""
    rows = fft8(np.asarray(x, dtype=np.complex128))
    return np.swapaxes(fft8(np.swapaxes(rows, -1, -2)), -1, -2)
