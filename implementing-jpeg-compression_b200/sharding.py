"""Partitioning of codec work across the GPUs of one box.

Blocks are independent and every block is byte aligned in the stream
(rle_byte_stream.py:55-56), so any contiguous run of blocks in raster order is a
self-contained byte substring.  Two partitions follow, neither needs a collective:

* a batch of images: rank r takes a contiguous slice of the images
  (``image_slice``); results are gathered by rank order;
* one huge image: rank r takes a band of block rows of every colour plane
  (``block_row_bands``), compresses it as an image of its own, and the host
  concatenates, per colour plane, the sub-streams in rank order
  (``concat_band_streams``).  Only the last band sees the bottom edge replication.
"""
import math


def image_slice(n_images, rank, world_size):
    """Contiguous [start, stop) of images for `rank`; sizes differ by at most one."""
    base, extra = divmod(n_images, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def block_row_bands(height, block_size, dct_size, world_size):
    """Split `height` source rows into at most `world_size` bands whose boundaries fall on
    block-row boundaries (multiples of block_size * dct_size source rows).

    Returns a list of (row_start, row_stop) per rank; ranks beyond the number of block
    rows get an empty band (row_start == row_stop)."""
    rows_per_block_row = block_size * dct_size
    n_block_rows = math.ceil(math.ceil(height / block_size) / dct_size)
    bands = []
    for r in range(world_size):
        b0, b1 = image_slice(n_block_rows, r, world_size)
        bands.append((min(b0 * rows_per_block_row, height), min(b1 * rows_per_block_row, height)))
    return bands


def concat_band_streams(per_rank_streams):
    """per_rank_streams[r][c] = bytes of colour plane c produced by rank r (empty bands
    contribute nothing).  Returns one stream per colour plane for the whole image."""
    n_planes = max((len(s) for s in per_rank_streams if s), default=0)
    return [b"".join(s[c] for s in per_rank_streams if s) for c in range(n_planes)]
