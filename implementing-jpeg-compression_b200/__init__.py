"""B200-native per-block codec path of X-rayLaser/Implementing-JPEG-compression.

The directory name carries a hyphen (it mirrors the reference repository's name), so
import it through the ``jpeg_b200`` alias module at the repository root:

    import jpeg_b200 as jb
    data = jb.compress_band(plane, jb.Configuration(width=w, height=h, block_size=4,
                                                    quantization=jb.QuantizationMethod('qtable')))
"""
from .errors import (BadArrayShapeError, BadQuantizationError, BadRleCodeError, BadStreamError,  # noqa: F401
                     EmptyArrayError, NativeLibraryError)
from .config import Configuration, QuantizationMethod  # noqa: F401
from . import file_format  # noqa: F401
from .file_format import CompressedData  # noqa: F401
from .codec import (Jpeg, CompressedPlanes, BatchCodec, compress_band, decompress_band, compress_bands,  # noqa: F401
                    decompress_bands, compress_planes, decompress_planes, check_status, geometry,
                    rgb_to_ycbcr_planes, ycbcr_planes_to_rgb, pack_containers, containers_to_bytes,
                    compress_images_rgb)
from . import stages, sharding  # noqa: F401

__all__ = ["Configuration", "QuantizationMethod", "Jpeg", "CompressedData", "CompressedPlanes", "BatchCodec",
           "compress_band", "decompress_band", "compress_bands", "decompress_bands", "compress_planes",
           "decompress_planes", "check_status", "geometry", "rgb_to_ycbcr_planes", "ycbcr_planes_to_rgb", "pack_containers", "containers_to_bytes", "compress_images_rgb", "file_format", "stages", "sharding",
           "BadArrayShapeError", "BadQuantizationError", "BadRleCodeError", "BadStreamError",
           "EmptyArrayError", "NativeLibraryError"]
