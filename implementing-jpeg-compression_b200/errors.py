"""Exception types of the codec path, named as in the reference
(util.py:92-97,232-233; pipeline/__init__.py:67-68) so that callers written
against the reference keep working."""


class BadArrayShapeError(Exception):
    pass


class EmptyArrayError(Exception):
    pass


class BadRleCodeError(Exception):
    pass


class BadQuantizationError(Exception):
    pass


class BadStreamError(ValueError):
    """A byte stream that does not decode to the block count of its geometry (the
    reference fails with a ValueError from ``reshape``, run_length_encoding.py:76-78)."""


class NativeLibraryError(RuntimeError):
    """libjpegb200.so is missing or a CUDA call failed.  There is no CPU fallback."""
