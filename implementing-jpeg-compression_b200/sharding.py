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


def bind_host_to_device(device_index):
    """Pin the calling process to the CPU cores next to GPU `device_index` (NVML's ideal CPU affinity), so that
    the pinned staging buffers it allocates afterwards are first-touched on the GPU's own NUMA node.  With one
    process per GPU the host side of the PCIe links is the bound of the end-to-end path (6.5 GB each way per
    batch), and copies that cross the socket interconnect share its bandwidth between all ranks.

    Returns the sorted list of cores, or None when NVML or the affinity call is unavailable (nothing is changed
    then).  The device is looked up by UUID, so CUDA_VISIBLE_DEVICES remapping is honoured."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cores = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cores = sorted(c for c in cores if c in allowed)
        if not cores:
            return None
        os.sched_setaffinity(0, cores)
        return cores
    except Exception:
        return None
