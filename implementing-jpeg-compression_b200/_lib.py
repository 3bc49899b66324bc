"""ctypes binding of libjpegb200.so (include/jpegb200.h).

The library is the product: if it is missing or cannot be loaded this module
raises -- there is deliberately no CPU fallback."""
import ctypes
import os

from .errors import NativeLibraryError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjpegb200.so")

JB_TRANSFORM_DCT, JB_TRANSFORM_DFT = 0, 1
JB_Q_NONE, JB_Q_DISCARD, JB_Q_DIVIDE, JB_Q_QTABLE = 0, 1, 2, 3
JB_OK = 0
JB_ERR_BAD_PARAM, JB_ERR_BAD_QUANTIZATION, JB_ERR_EMPTY_ARRAY, JB_ERR_UNSUPPORTED = -1, -2, -3, -4
JB_ERR_WORKSPACE, JB_ERR_OUT_CAPACITY, JB_ERR_CUDA, JB_ERR_BAD_RLE_CODE = -5, -6, -7, -8
JB_ERR_BAD_STREAM, JB_ERR_NO_DEVICE = -9, -10
JB_FLAG_FORCE_GENERIC, JB_FLAG_NO_TMA, JB_FLAG_NO_REFINE, JB_FLAG_SERIAL_FRAMING = 1, 2, 4, 8
JB_FLAG_REUSE_TABLES = 16
JB_FLAG_TILE_DECODER = 64
JB_FLAG_PDL = 128
JB_F64_SAMPLES, JB_F64_TRANSFORM, JB_F64_PREROUNDING = 0, 1, 2
JB_STATUS_WORDS = 4
JB_MAX_DCT_SIZE = 32
JB_MAX_BLOCK_SIZE = 255


class jb_params(ctypes.Structure):
    _fields_ = [("height", ctypes.c_int32), ("width", ctypes.c_int32),
                ("block_size", ctypes.c_int32), ("dct_size", ctypes.c_int32),
                ("transform", ctypes.c_int32), ("qmode", ctypes.c_int32),
                ("qparam", ctypes.c_int32), ("flags", ctypes.c_int32)]


class jb_geometry(ctypes.Structure):
    _fields_ = [("h1", ctypes.c_int32), ("w1", ctypes.c_int32), ("h2", ctypes.c_int32),
                ("w2", ctypes.c_int32), ("vb", ctypes.c_int32), ("hb", ctypes.c_int32),
                ("blocks_per_plane", ctypes.c_int32), ("max_block_bytes", ctypes.c_int32),
                ("chunks_per_plane", ctypes.c_int32), ("reserved", ctypes.c_int32)]


# every symbol include/jpegb200.h declares: name -> (restype, argtypes)
_P = ctypes.c_void_p
_SZ = ctypes.c_size_t
_I = ctypes.c_int
_PP = ctypes.POINTER(jb_params)
SYMBOLS = {
    "jb_version": (_I, []),
    "jb_strerror": (ctypes.c_char_p, [_I]),
    "jb_geometry_of": (_I, [_PP, ctypes.POINTER(jb_geometry)]),
    "jb_max_stream_bytes": (_SZ, [_PP, _I]),
    "jb_compress_workspace_bytes": (_SZ, [_PP, _I]),
    "jb_decompress_workspace_bytes": (_SZ, [_PP, _I, _SZ]),
    "jb_decompress_framing_path": (_I, [_PP, _I, _SZ]),
    "jb_compress_planes": (_I, [_P, _SZ, _SZ, _I, _PP, _P, _SZ, _P, _P, _P, _SZ, _P]),
    "jb_decompress_planes": (_I, [_P, _SZ, _P, _P, _I, _PP, _P, _SZ, _SZ, _P, _P, _SZ, _P]),
    "jb_stage_forward_coeffs": (_I, [_P, _SZ, _SZ, _I, _PP, _P, _P, _P, _SZ, _P]),
    "jb_stage_float64": (_I, [_P, _SZ, _SZ, _I, _PP, _I, _P, _P, _SZ, _P]),
    "jb_stage_pack": (_I, [_P, _I, _I, _I, _P, _SZ, _P, _P, _P, _SZ, _P]),
    "jb_stage_pack_workspace_bytes": (_SZ, [_I, _I, _I]),
    "jb_stage_unpack": (_I, [_P, _SZ, _P, _P, _I, _I, _I, _I, _P, _P, _P, _SZ, _P]),
    "jb_stage_unpack_workspace_bytes": (_SZ, [_I, _I, _I, _SZ]),
    "jb_stage_inverse_coeffs": (_I, [_P, _I, _PP, _P, _SZ, _SZ, _P, _P, _SZ, _P]),
    "jb_rgb_to_ycbcr_planes": (_I, [_P, _SZ, _SZ, _I, _I, _I, _P, _SZ, _SZ, _P]),
    "jb_ycbcr_planes_to_rgb": (_I, [_P, _SZ, _SZ, _I, _I, _I, _P, _SZ, _SZ, _P]),
    "jb_containers_max_bytes": (_SZ, [_I, _I, _SZ]),
    "jb_pack_containers": (_I, [_P, _P, _I, ctypes.c_char_p, _I, _P, _SZ, _P, _P, _P]),
    "jb_debug_kernel_events": (_I, [_P, _P, _P, _P]),
    "jb_debug_launch_count": (ctypes.c_ulonglong, []),
}

_lib = None


def load():
    """Load the shared library once; raise NativeLibraryError if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise NativeLibraryError(
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C implementing-jpeg-compression_b200/csrc`). There is no CPU fallback." % LIB_PATH)
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:
        raise NativeLibraryError("cannot load %s: %s" % (LIB_PATH, e))
    for name, (restype, argtypes) in SYMBOLS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise NativeLibraryError("%s does not export %s" % (LIB_PATH, name))
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def strerror(code):
    return load().jb_strerror(int(code)).decode("ascii")
