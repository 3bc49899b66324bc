"""The reference's container, byte for byte (file_format.py:67-111).

Header, little-endian: u16 header_length (= 15 + len(json)), u16 width,
u16 height, u16 block_size, u16 dct_size, 3 ASCII bytes transform, u16
len(json), the quantisation JSON.  Body: for Y, Cb, Cr a u32 length followed by
the band's stream."""
import struct

from .config import Configuration, QuantizationMethod


class CompressedData:
    """pipeline.CompressedData (pipeline/__init__.py:91-95)."""

    def __init__(self, y, cb, cr):
        self.y = y
        self.cb = cb
        self.cr = cr


_FIXED = struct.Struct("<HHHHH3sH")


def create_header(config):
    qjson = config.quantization.to_json().encode("ascii")
    transform = config.transform.encode("ascii")
    return _FIXED.pack(2 + 13 + len(qjson), config.width, config.height, config.block_size,
                       config.dct_size, transform, len(qjson)) + qjson


def get_header(bytestream):
    hlen, width, height, block_size, dct_size, transform, qlen = _FIXED.unpack_from(bytestream, 0)
    qjson = bytes(bytestream[_FIXED.size:_FIXED.size + qlen]).decode()
    return Configuration(width=width, height=height, block_size=block_size, dct_size=dct_size,
                         transform=transform.decode(), quantization=QuantizationMethod.from_json(qjson))


def generate_data(config, compressed_data):
    parts = [create_header(config)]
    for band in (compressed_data.y, compressed_data.cb, compressed_data.cr):
        parts.append(struct.pack("<L", len(band)))
        parts.append(bytes(band))
    return b"".join(parts)


def read_data(bytestream):
    config = get_header(bytestream)
    (hlen,) = struct.unpack_from("<H", bytestream, 0)
    pos = hlen
    bands = []
    for _ in range(3):
        (n,) = struct.unpack_from("<L", bytestream, pos)
        pos += 4
        bands.append(bytes(bytestream[pos:pos + n]))
        pos += n
    return config, CompressedData(*bands)
