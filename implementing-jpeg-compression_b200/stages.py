"""Stage-level hooks: the kernels cut at the quantised-coefficient boundary.

They exist for the parity tests, which compare each half of the path with the
oracle separately (SURVEY.md section 8b "per-stage execute/invert test hooks"):

* ``forward_coefficients``  stages 0-6 + the RLE stage's int cast -> (n, vb, hb, d*d) int16
* ``pack_coefficients``     stages 7-8 (RunLengthEncoding + RleBytestream)     -> bytes per plane
* ``unpack_streams``        stages 8-7 inverted                                 -> (n, blocks, d*d) int16
* ``inverse_coefficients``  stages 6-0 inverted                                 -> (n, H, W) uint8
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .codec import (_ptr, _raise_for_code, _require_cuda, _stream_ptr, _workspace, check_status,
                    CompressedPlanes, geometry)


def forward_coefficients(planes, config, flags=0):
    lib = _lib.load()
    dev = _require_cuda()
    planes = torch.as_tensor(np.ascontiguousarray(planes, dtype=np.uint8)).to(dev)
    if planes.dim() == 2:
        planes = planes.unsqueeze(0)
    n, h, w = planes.shape
    p = config.c_params(flags)
    g = geometry(config)
    d2 = config.dct_size ** 2
    coeffs = torch.zeros((n, g["vb"], g["hb"], d2), dtype=torch.int16, device=dev)
    ws_bytes = lib.jb_compress_workspace_bytes(ctypes.byref(p), n)
    ws = _workspace.get("fwd", ws_bytes, dev)
    status = torch.empty(_lib.JB_STATUS_WORDS, dtype=torch.int64, device=dev)
    rc = lib.jb_stage_forward_coeffs(_ptr(planes), h * planes.stride(1) if n == 1 else planes.stride(0),
                                     planes.stride(1), n, ctypes.byref(p), _ptr(coeffs), _ptr(status),
                                     _ptr(ws), ws.numel(), _stream_ptr())
    _raise_for_code(rc)
    check_status(status)
    return coeffs.cpu().numpy()


def float64_stage(planes, config, which, flags=0):
    """Float64 intermediates of the compress direction, as the reference's stages hand them on
    (pipeline/base.py:42-72).  which: 'samples' -- after Padding, SubSampling, DCTPadding: (n, vb*d, hb*d) box means;
    'transform' -- after BasisChange.execute: (n, vb, hb, d, d), real part; 'prerounding' -- after the quantiser's
    scaling, the value np.round receives: same shape.  Computed in the reference's operation order (the arithmetic
    that decides rounding ties inside the fused kernels, csrc/jb_refine.cuh)."""
    lib = _lib.load()
    dev = _require_cuda()
    code = {"samples": _lib.JB_F64_SAMPLES, "transform": _lib.JB_F64_TRANSFORM, "prerounding": _lib.JB_F64_PREROUNDING}[which]
    planes = torch.as_tensor(np.ascontiguousarray(planes, dtype=np.uint8)).to(dev)
    if planes.dim() == 2:
        planes = planes.unsqueeze(0)
    n, h, w = planes.shape
    p = config.c_params(flags)
    g = geometry(config)
    d = config.dct_size
    if code == _lib.JB_F64_SAMPLES:
        out = torch.zeros((n, g["vb"] * d, g["hb"] * d), dtype=torch.float64, device=dev)
    else:
        out = torch.zeros((n, g["vb"], g["hb"], d, d), dtype=torch.float64, device=dev)
    ws_bytes = lib.jb_compress_workspace_bytes(ctypes.byref(p), n)
    ws = _workspace.get("fwd", ws_bytes, dev)
    rc = lib.jb_stage_float64(_ptr(planes), h * planes.stride(1) if n == 1 else planes.stride(0), planes.stride(1), n,
                              ctypes.byref(p), code, _ptr(out), _ptr(ws), ws.numel(), _stream_ptr())
    _raise_for_code(rc)
    return out.cpu().numpy()


def pack_coefficients(coeffs, dct_size):
    """coeffs: integer array (n_planes, blocks, d*d) (or (blocks, d*d)) in zigzag order."""
    lib = _lib.load()
    dev = _require_cuda()
    c = np.ascontiguousarray(coeffs, dtype=np.int32)
    if c.ndim == 2:
        c = c[None]
    n, blocks, d2 = c.shape
    assert d2 == dct_size * dct_size
    ct = torch.from_numpy(c).to(dev)
    cap = n * blocks * ((23 * d2 + 15) // 8)
    out = torch.empty(cap, dtype=torch.uint8, device=dev)
    offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
    status = torch.empty(_lib.JB_STATUS_WORDS, dtype=torch.int64, device=dev)
    ws_bytes = lib.jb_stage_pack_workspace_bytes(n, blocks, dct_size)
    ws = _workspace.get("fwd", ws_bytes, dev)
    rc = lib.jb_stage_pack(_ptr(ct), n, blocks, dct_size, _ptr(out), out.numel(), _ptr(offsets), _ptr(status),
                           _ptr(ws), ws.numel(), _stream_ptr())
    _raise_for_code(rc)
    return CompressedPlanes(out, offsets, status, n).to_bytes_list()


def unpack_streams(streams, blocks_per_plane, dct_size, flags=0):
    lib = _lib.load()
    dev = _require_cuda()
    streams = [bytes(s) for s in streams]
    n = len(streams)
    lens = np.array([len(s) for s in streams], dtype=np.int64)
    offs = np.concatenate(([0], np.cumsum(lens)[:-1])).astype(np.int64)
    blob = b"".join(streams)
    data = torch.from_numpy(np.frombuffer(blob + b"\0" * 16, dtype=np.uint8).copy()).to(dev)
    d2 = dct_size * dct_size
    coeffs = torch.zeros((n, blocks_per_plane, d2), dtype=torch.int16, device=dev)
    status = torch.empty(_lib.JB_STATUS_WORDS, dtype=torch.int64, device=dev)
    ws_bytes = lib.jb_stage_unpack_workspace_bytes(n, blocks_per_plane, dct_size, len(blob))
    ws = _workspace.get("inv", ws_bytes, dev)
    d_offs, d_lens = torch.from_numpy(offs).to(dev), torch.from_numpy(lens).to(dev)
    rc = lib.jb_stage_unpack(_ptr(data), len(blob), _ptr(d_offs), _ptr(d_lens), n, blocks_per_plane, dct_size, int(flags),
                             _ptr(coeffs), _ptr(status), _ptr(ws), ws.numel(), _stream_ptr())
    _raise_for_code(rc)
    check_status(status)
    return coeffs.cpu().numpy()


def inverse_coefficients(coeffs, config, flags=0):
    lib = _lib.load()
    dev = _require_cuda()
    c = np.ascontiguousarray(coeffs, dtype=np.int16)
    g = geometry(config)
    d2 = config.dct_size ** 2
    c = c.reshape(-1, g["blocks_per_plane"], d2)
    n = c.shape[0]
    ct = torch.from_numpy(c).to(dev)
    p = config.c_params(flags)
    h, w = int(config.height), int(config.width)
    out = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    status = torch.empty(_lib.JB_STATUS_WORDS, dtype=torch.int64, device=dev)
    ws_bytes = lib.jb_decompress_workspace_bytes(ctypes.byref(p), n, 0)
    ws = _workspace.get("inv", ws_bytes, dev)
    rc = lib.jb_stage_inverse_coeffs(_ptr(ct), n, ctypes.byref(p), _ptr(out), out.stride(0), out.stride(1),
                                     _ptr(status), _ptr(ws), ws.numel(), _stream_ptr())
    _raise_for_code(rc)
    check_status(status)
    return out.cpu().numpy()
