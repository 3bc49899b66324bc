"""pipeline/cuda_path.py -- the file a maintainer of the reference adds to run the per-block codec path on a B200.

Self-contained on purpose: ctypes + torch (buffer carrier only) + the reference's own exception types; nothing
from this repository's Python package.  It replaces the two functions that drive the nine pipeline stages,

    compress_band(a, config)   -> bytes      pipeline/__init__.py:71-76
    decompress_band(b, config) -> ndarray    pipeline/__init__.py:79-88

through the C ABI of libjpegb200.so (include/jpegb200.h).  To swap them in:

    import pipeline, cuda_path
    cuda_path.install(pipeline)        # Jpeg.compress / Jpeg.decompress and both CLIs now run on the GPU

``Configuration``, ``QuantizationMethod``, ``file_format`` and the CLI flags stay as they are; the streams are
the reference's byte for byte (given identical quantised integers), so files written either way decode either
way.  Do NOT register AlgorithmStep subclasses instead: the metaclass appends every subclass to step_classes
(pipeline/base.py:23-31) and the stock compress_band would then run them next to the originals.

tests/test_reference_dropin.py runs the reference's own tests/integration_tests.py over this file.
"""
import ctypes
import os

import numpy as np
import torch

from util import BadArrayShapeError, BadRleCodeError, EmptyArrayError

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("JPEGB200_LIB") or os.path.join(
    os.path.dirname(_HERE), "implementing-jpeg-compression_b200", "libjpegb200.so")
_lib = ctypes.CDLL(_LIB_PATH)


class _Params(ctypes.Structure):            # jb_params: mirror of pipeline.Configuration + flags
    _fields_ = [(n, ctypes.c_int32) for n in
                ("height", "width", "block_size", "dct_size", "transform", "qmode", "qparam", "flags")]


_P, _SZ, _I, _PP = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(_Params)
_lib.jb_max_stream_bytes.restype = _SZ
_lib.jb_max_stream_bytes.argtypes = [_PP, _I]
_lib.jb_compress_workspace_bytes.restype = _SZ
_lib.jb_compress_workspace_bytes.argtypes = [_PP, _I]
_lib.jb_decompress_workspace_bytes.restype = _SZ
_lib.jb_decompress_workspace_bytes.argtypes = [_PP, _I, _SZ]
_lib.jb_compress_planes.restype = _I
_lib.jb_compress_planes.argtypes = [_P, _SZ, _SZ, _I, _PP, _P, _SZ, _P, _P, _P, _SZ, _P]
_lib.jb_decompress_planes.restype = _I
_lib.jb_decompress_planes.argtypes = [_P, _SZ, _P, _P, _I, _PP, _P, _SZ, _SZ, _P, _P, _SZ, _P]
_lib.jb_strerror.restype = ctypes.c_char_p
_lib.jb_strerror.argtypes = [_I]

_QMODE = {"none": 0, "discard": 1, "divide": 2, "qtable": 3}          # JB_Q_*
_JB_ERR_BAD_QUANTIZATION, _JB_ERR_EMPTY_ARRAY, _JB_ERR_BAD_RLE_CODE, _JB_ERR_BAD_STREAM = -2, -3, -8, -9


class BadStreamError(ValueError):
    """The stream does not decode to the block count of the configuration (the stock decoder dies with an
    IndexError / reshape error somewhere inside BitDecoder in that case)."""


def _params(config):
    q = config.quantization
    qparam = 0
    if q.name == "discard":
        qparam = q.params.get("keep", 2)                 # quantizers.py:13
    elif q.name == "divide":
        qparam = q.params.get("divisor", 40)             # quantizers.py:24
    if q.name not in _QMODE:
        raise _bad_quantization()
    if config.transform not in ("DCT", "DFT"):
        raise UnboundLocalError("unknown transform %r" % (config.transform,))      # basis_change.py:26
    return _Params(int(config.height), int(config.width), int(config.block_size), int(config.dct_size),
                   {"DCT": 0, "DFT": 1}[config.transform], _QMODE[q.name], int(qparam), 0)


def _bad_quantization():
    from pipeline import BadQuantizationError
    return BadQuantizationError()


def _raise(code, status=None):
    """C-ABI return / status codes -> the exceptions the stock path raises."""
    if code == 0:
        return
    if code == _JB_ERR_BAD_QUANTIZATION:
        raise _bad_quantization()
    if code == _JB_ERR_EMPTY_ARRAY:
        raise EmptyArrayError()
    if code == _JB_ERR_BAD_RLE_CODE:
        msg = ""
        if status is not None:
            w = int(status[1]) & (2 ** 64 - 1)
            if w != 2 ** 64 - 1:
                amp = (w & 0xFFFFF) - (1 << 19)
                msg = "({}, {}, {})".format((w >> 20) & 15, abs(amp).bit_length() + 1, amp)     # util.py:163
        raise BadRleCodeError(msg)
    if code == _JB_ERR_BAD_STREAM:
        raise BadStreamError(_lib.jb_strerror(code).decode())
    raise RuntimeError("libjpegb200: %s (%d)" % (_lib.jb_strerror(code).decode(), code))


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def compress_band(a, config):
    """Replaces pipeline.compress_band (pipeline/__init__.py:71-76)."""
    a = np.asarray(a)
    if a.ndim != 2:
        raise BadArrayShapeError()                       # util.py:27-28
    if 0 in a.shape:
        raise EmptyArrayError()                          # util.py:30-31
    if a.min() < 0 or a.max() > 255:
        raise ValueError("the CUDA path takes 8-bit samples")
    p = _params(config)
    p.height, p.width = a.shape                          # the stock stages take the geometry from the array
    planes = torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint8)).cuda()
    cap = _lib.jb_max_stream_bytes(ctypes.byref(p), 1)
    if cap == 0:
        raise _bad_quantization()
    out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    off = torch.empty(2, dtype=torch.int64, device="cuda")
    status = torch.empty(4, dtype=torch.int64, device="cuda")
    ws = torch.empty(max(256, _lib.jb_compress_workspace_bytes(ctypes.byref(p), 1)), dtype=torch.uint8, device="cuda")
    _raise(_lib.jb_compress_planes(_ptr(planes), planes.numel(), planes.stride(0), 1, ctypes.byref(p),
                                   _ptr(out), cap, _ptr(off), _ptr(status), _ptr(ws), ws.numel(), _stream()))
    st = status.cpu()                                    # synchronises with the stream
    _raise(-int(st[0]), st)
    return out[: int(off[1])].cpu().numpy().tobytes()


def decompress_band(compression_result, config):
    """Replaces pipeline.decompress_band (pipeline/__init__.py:79-88): (height, width) integer array."""
    data = bytes(compression_result)
    p = _params(config)
    n = len(data)
    d_in = torch.from_numpy(np.frombuffer(data + b"\0" * 16, dtype=np.uint8).copy()).cuda()
    off = torch.zeros(1, dtype=torch.int64, device="cuda")
    length = torch.full((1,), n, dtype=torch.int64, device="cuda")
    out = torch.empty((int(config.height), int(config.width)), dtype=torch.uint8, device="cuda")
    status = torch.empty(4, dtype=torch.int64, device="cuda")
    need = _lib.jb_decompress_workspace_bytes(ctypes.byref(p), 1, n)
    if need == 0:
        raise _bad_quantization()
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    _raise(_lib.jb_decompress_planes(_ptr(d_in), n, _ptr(off), _ptr(length), 1, ctypes.byref(p), _ptr(out),
                                     out.numel(), out.stride(0), _ptr(status), _ptr(ws), ws.numel(), _stream()))
    st = status.cpu()
    _raise(-int(st[0]), st)
    return out.cpu().numpy().astype(np.int64)


def install(pipeline_module):
    """Swap the two band functions of an imported ``pipeline`` package for the CUDA path."""
    pipeline_module.compress_band = compress_band
    pipeline_module.decompress_band = decompress_band
