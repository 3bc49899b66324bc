"""Host side of the B200 codec path: the reference's band/image API over libjpegb200.so.

Drop-in surface (same names and meaning as the reference):

* ``compress_band(a, config) -> bytes``            pipeline/__init__.py:71-76
* ``decompress_band(data, config) -> ndarray``     pipeline/__init__.py:79-88
* ``Jpeg(config).compress(image) -> bytes`` / ``Jpeg.decompress(bytes) -> image``
                                                   pipeline/__init__.py:98-124

plus the batched, device-resident calls the reference does not have
(``compress_planes`` / ``decompress_planes``), which is what the benchmark and a
production caller use.  PyTorch is only the buffer carrier (device memory, the
current stream); every byte of codec work happens in the CUDA kernels.
"""
import ctypes
import threading

import numpy as np
import torch

from . import _lib, file_format
from .config import Configuration, QuantizationMethod  # noqa: F401  (re-exported)
from .errors import (BadArrayShapeError, BadQuantizationError, BadRleCodeError, BadStreamError,
                     EmptyArrayError, NativeLibraryError)


def _raise_for_code(code, detail_word=None):
    """Map a C-ABI code to the reference's exception types."""
    if code == _lib.JB_OK:
        return
    if code == _lib.JB_ERR_BAD_QUANTIZATION:
        raise BadQuantizationError()
    if code == _lib.JB_ERR_EMPTY_ARRAY:
        raise EmptyArrayError()
    if code == _lib.JB_ERR_BAD_RLE_CODE:
        msg = ""
        if detail_word is not None and detail_word != 0xFFFFFFFFFFFFFFFF:
            run = (detail_word >> 20) & 15
            amp = (detail_word & 0xFFFFF) - (1 << 19)
            size = abs(amp).bit_length() + 1
            msg = "({}, {}, {})".format(run, size, amp)      # util.py:163
        raise BadRleCodeError(msg)
    if code == _lib.JB_ERR_BAD_STREAM:
        raise BadStreamError(_lib.strerror(code))
    if code == _lib.JB_ERR_UNSUPPORTED:
        raise NotImplementedError("libjpegb200: " + _lib.strerror(code))
    if code in (_lib.JB_ERR_BAD_PARAM, _lib.JB_ERR_OUT_CAPACITY, _lib.JB_ERR_WORKSPACE):
        raise ValueError("libjpegb200: " + _lib.strerror(code))
    raise NativeLibraryError("libjpegb200: %s (%d)" % (_lib.strerror(code), code))


def _require_cuda(device=None):
    if not torch.cuda.is_available():
        raise NativeLibraryError("no CUDA device: the codec path runs only on the GPU (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def geometry(config):
    """Derived sizes of one band as a dict (run_length_encoding.py:80-88)."""
    p = config.c_params()
    g = _lib.jb_geometry()
    _raise_for_code(_lib.load().jb_geometry_of(ctypes.byref(p), ctypes.byref(g)))
    return {k: getattr(g, k) for k, _ in g._fields_ if k != "reserved"}


class CompressedPlanes:
    """Result of ``compress_planes``: device buffer + device offsets; nothing has been
    copied to the host yet."""

    def __init__(self, data, offsets, status, n_planes):
        self.data = data            # uint8 [capacity] on device; streams concatenated in plane order
        self.offsets = offsets      # int64 [n_planes + 1] on device
        self.status = status        # int64 [4] on device
        self.n_planes = n_planes
        self._host_offsets = None

    def check(self):
        """Synchronise with the producing stream and raise if the device reported an error."""
        st = self.status.cpu()
        code = int(st[0].item())
        if code:
            _raise_for_code(-code, int(st[1].item()) & 0xFFFFFFFFFFFFFFFF)
        return self

    def host_offsets(self):
        if self._host_offsets is None:
            self.check()
            self._host_offsets = self.offsets.cpu().numpy().astype(np.int64)
        return self._host_offsets

    def total_bytes(self):
        return int(self.host_offsets()[-1])

    def to_bytes_list(self):
        """One ``bytes`` per plane, in plane order (what compress_band returns)."""
        off = self.host_offsets()
        total = int(off[-1])
        host = self.data[:total].cpu().numpy().tobytes() if total else b""
        return [host[int(off[i]):int(off[i + 1])] for i in range(self.n_planes)]


class _Workspace:
    """Grow-only cache of device scratch buffers (the library itself allocates nothing): one buffer per
    (direction, device, CUDA stream, host thread).  Calls are asynchronous on the current stream and use the
    workspace until they finish, so two streams -- or two threads -- must never share one; calls queued on the
    SAME stream run in order and may.  A buffer that is outgrown is handed back to the caching allocator, which
    keeps it alive for the work already queued on the stream it was allocated on (the one that used it)."""

    def __init__(self):
        self._bufs = {}
        self._lock = threading.Lock()

    def get(self, key, nbytes, device):
        k = (key, device.index, torch.cuda.current_stream(device).cuda_stream, threading.get_ident())
        with self._lock:
            t = self._bufs.get(k)
            if t is None or t.numel() < nbytes:
                with torch.cuda.device(device):
                    t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
                self._bufs[k] = t
            return t


_workspace = _Workspace()


def compress_planes(planes, config, flags=0, out=None, ws=None):
    """Compress a batch of equally sized planes that already live on the GPU.

    ``planes``: uint8 CUDA tensor [n, H, W] (or [H, W]); rows contiguous.  Returns a
    ``CompressedPlanes`` whose buffers stay on the device; the call is asynchronous on the
    current stream.  Replaces ``compress_band`` applied to each plane in turn.

    ``ws``: caller-owned scratch (``jb_compress_workspace_bytes``); by default a cached buffer private to the
    current (device, stream, thread).  A caller that passes its own must not use it from two streams at once."""
    lib = _lib.load()
    if planes.dim() == 2:
        planes = planes.unsqueeze(0)
    if planes.dim() != 3:
        raise BadArrayShapeError(tuple(planes.shape))
    if planes.dtype != torch.uint8 or not planes.is_cuda:
        raise TypeError("compress_planes expects a uint8 CUDA tensor")
    n, h, w = planes.shape
    if h == 0 or w == 0 or n == 0:
        raise EmptyArrayError()
    if planes.stride(2) != 1:
        planes = planes.contiguous()
    if (h, w) != (config.height, config.width):
        # the reference takes the geometry from the array; keep the two in agreement
        raise ValueError("plane is %dx%d but config says %dx%d" % (h, w, config.height, config.width))
    dev = planes.device
    p = config.c_params(flags)
    cap = lib.jb_max_stream_bytes(ctypes.byref(p), n)
    if cap == 0:
        g = _lib.jb_geometry()
        _raise_for_code(lib.jb_geometry_of(ctypes.byref(p), ctypes.byref(g)))
    ws_bytes = lib.jb_compress_workspace_bytes(ctypes.byref(p), n)
    with torch.cuda.device(dev):
        if ws is None:
            ws = _workspace.get("fwd", ws_bytes, dev)
        elif ws.numel() < ws_bytes:
            raise ValueError("workspace too small: %d < %d bytes" % (ws.numel(), ws_bytes))
        if out is None:
            out = torch.empty(cap, dtype=torch.uint8, device=dev)
        offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
        status = torch.empty(_lib.JB_STATUS_WORDS, dtype=torch.int64, device=dev)
        rc = lib.jb_compress_planes(_ptr(planes), planes.stride(0) if n > 1 else h * planes.stride(1),
                                    planes.stride(1), n, ctypes.byref(p), _ptr(out), out.numel(),
                                    _ptr(offsets), _ptr(status), _ptr(ws), ws.numel(), _stream_ptr())
    _raise_for_code(rc)
    return CompressedPlanes(out, offsets, status, n)


def decompress_planes(data, offsets, lengths, config, n_planes, in_bytes=None, flags=0, out=None, ws=None):
    """Decompress ``n_planes`` streams held in the device buffer ``data``.

    ``offsets`` / ``lengths``: int64 CUDA tensors [n_planes].  ``in_bytes``: host-known upper
    bound of the stream bytes in ``data`` (defaults to ``data.numel()``).  Returns
    ``(planes uint8 [n, H, W] on device, status int64 [4] on device)``; asynchronous."""
    lib = _lib.load()
    dev = data.device
    p = config.c_params(flags)
    if in_bytes is None:
        in_bytes = data.numel()
    ws_bytes = lib.jb_decompress_workspace_bytes(ctypes.byref(p), n_planes, int(in_bytes))
    if ws_bytes == 0:
        g = _lib.jb_geometry()
        _raise_for_code(lib.jb_geometry_of(ctypes.byref(p), ctypes.byref(g)))
    h, w = int(config.height), int(config.width)
    with torch.cuda.device(dev):
        if ws is None:
            ws = _workspace.get("inv", ws_bytes, dev)
        elif ws.numel() < ws_bytes:
            raise ValueError("workspace too small: %d < %d bytes" % (ws.numel(), ws_bytes))
        if out is None:
            out = torch.empty((n_planes, h, w), dtype=torch.uint8, device=dev)
        status = torch.empty(_lib.JB_STATUS_WORDS, dtype=torch.int64, device=dev)
        rc = lib.jb_decompress_planes(_ptr(data), int(in_bytes), _ptr(offsets), _ptr(lengths), n_planes,
                                      ctypes.byref(p), _ptr(out), out.stride(0), out.stride(1),
                                      _ptr(status), _ptr(ws), ws.numel(), _stream_ptr())
    _raise_for_code(rc)
    return out, status


def check_status(status):
    st = status.cpu()
    code = int(st[0].item())
    if code:
        _raise_for_code(-code, int(st[1].item()) & 0xFFFFFFFFFFFFFFFF)


# ---- host-buffer entry points (the reference's signatures) --------------------------------------

def _as_uint8_plane(a):
    a = np.asarray(a)
    if a.ndim != 2:
        raise BadArrayShapeError()                       # util.py:27-28
    if a.shape[0] == 0 or a.shape[1] == 0:
        raise EmptyArrayError()                          # util.py:30-31
    if a.dtype != np.uint8:
        if np.iscomplexobj(a) or not np.all(a == np.round(a)) or a.min() < 0 or a.max() > 255:
            raise ValueError("the CUDA codec path takes 8-bit samples (integers 0..255)")
        a = a.astype(np.uint8)
    return np.ascontiguousarray(a)


def compress_band(a, config, flags=0):
    """pipeline.compress_band (pipeline/__init__.py:71-76): 2-D integer array -> bytes."""
    dev = _require_cuda()
    a = _as_uint8_plane(a)
    planes = torch.from_numpy(a).to(dev, non_blocking=False).unsqueeze(0)
    res = compress_planes(planes, config, flags=flags)
    return res.to_bytes_list()[0]


def decompress_band(compression_result, config, flags=0):
    """pipeline.decompress_band (pipeline/__init__.py:79-88): bytes -> (height, width) int array."""
    planes = decompress_bands([compression_result], config, flags=flags)
    return planes[0].astype(np.int64)


def compress_bands(arrays, config, flags=0):
    """compress_band for several equally sized planes in one launch; list of bytes."""
    dev = _require_cuda()
    stack = np.stack([_as_uint8_plane(a) for a in arrays])
    planes = torch.from_numpy(stack).to(dev)
    return compress_planes(planes, config, flags=flags).to_bytes_list()


def decompress_bands(streams, config, flags=0):
    """decompress_band for several streams in one launch; uint8 array [n, H, W] on the host."""
    dev = _require_cuda()
    streams = [bytes(s) for s in streams]
    lens = np.array([len(s) for s in streams], dtype=np.int64)
    offs = np.concatenate(([0], np.cumsum(lens)[:-1])).astype(np.int64)
    blob = b"".join(streams)
    data = torch.from_numpy(np.frombuffer(blob + b"\0" * 16, dtype=np.uint8).copy()).to(dev)
    out, status = decompress_planes(data, torch.from_numpy(offs).to(dev), torch.from_numpy(lens).to(dev),
                                    config, len(streams), in_bytes=len(blob), flags=flags)
    check_status(status)
    return out.cpu().numpy()


def rgb_to_ycbcr_planes(rgb, out=None):
    """PIL's ``image.convert('YCbCr')`` on the device, bit-identical (compress.py:9).

    ``rgb``: uint8 CUDA tensor [n, H, W, 3] or [H, W, 3], interleaved.  Returns uint8 planes
    [3 n, H, W]: image i owns planes 3i (Y), 3i + 1 (Cb), 3i + 2 (Cr) -- the order
    ``compress_planes`` expects for a batch of images.  Asynchronous on the current stream."""
    lib = _lib.load()
    if rgb.dim() == 3:
        rgb = rgb.unsqueeze(0)
    if rgb.dim() != 4 or rgb.shape[3] != 3:
        raise BadArrayShapeError(tuple(rgb.shape))
    if rgb.dtype != torch.uint8 or not rgb.is_cuda:
        raise TypeError("rgb_to_ycbcr_planes expects a uint8 CUDA tensor")
    rgb = rgb.contiguous()
    n, h, w, _ = rgb.shape
    if n == 0 or h == 0 or w == 0:
        raise EmptyArrayError()
    with torch.cuda.device(rgb.device):
        if out is None:
            out = torch.empty((3 * n, h, w), dtype=torch.uint8, device=rgb.device)
        rc = lib.jb_rgb_to_ycbcr_planes(_ptr(rgb), rgb.stride(0), rgb.stride(1), n, h, w,
                                        _ptr(out), out.stride(0), out.stride(1), _stream_ptr())
    _raise_for_code(rc)
    return out


def ycbcr_planes_to_rgb(planes, out=None):
    """PIL's ``image.convert('RGB')`` of a YCbCr image on the device, bit-identical (decompress.py:10).
    ``planes``: uint8 CUDA tensor [3 n, H, W] as produced by ``decompress_planes``; returns [n, H, W, 3]."""
    lib = _lib.load()
    if planes.dim() != 3 or planes.shape[0] % 3 != 0:
        raise BadArrayShapeError(tuple(planes.shape))
    if planes.dtype != torch.uint8 or not planes.is_cuda:
        raise TypeError("ycbcr_planes_to_rgb expects a uint8 CUDA tensor")
    planes = planes.contiguous()
    n3, h, w = planes.shape
    if n3 == 0 or h == 0 or w == 0:
        raise EmptyArrayError()
    with torch.cuda.device(planes.device):
        if out is None:
            out = torch.empty((n3 // 3, h, w, 3), dtype=torch.uint8, device=planes.device)
        rc = lib.jb_ycbcr_planes_to_rgb(_ptr(planes), planes.stride(0), planes.stride(1), n3 // 3, h, w,
                                        _ptr(out), out.stride(0), out.stride(1), _stream_ptr())
    _raise_for_code(rc)
    return out


def pack_containers(comp, config):
    """file_format.generate_data for a whole batch on the device (file_format.py:86-93).

    ``comp``: the ``CompressedPlanes`` of 3 n planes (image i = planes 3i, 3i+1, 3i+2).  Returns
    ``(uint8 CUDA tensor with the n containers back to back, int64 CUDA tensor [n + 1] of offsets, status)``;
    asynchronous.  ``containers_to_bytes`` turns the pair into one ``bytes`` per image."""
    lib = _lib.load()
    if comp.n_planes % 3:
        raise BadArrayShapeError((comp.n_planes,))
    n = comp.n_planes // 3
    header = file_format.create_header(config)
    dev = comp.data.device
    cap = lib.jb_containers_max_bytes(n, len(header), comp.data.numel())
    with torch.cuda.device(dev):
        out = torch.empty(cap, dtype=torch.uint8, device=dev)
        offs = torch.empty(n + 1, dtype=torch.int64, device=dev)
        status = torch.zeros(_lib.JB_STATUS_WORDS, dtype=torch.int64, device=dev)
        rc = lib.jb_pack_containers(_ptr(comp.data), _ptr(comp.offsets), n, header, len(header), _ptr(out), out.numel(),
                                    _ptr(offs), _ptr(status), _stream_ptr())
    _raise_for_code(rc)
    return out, offs, status


def containers_to_bytes(out, offs, status=None):
    """One ``bytes`` per image from ``pack_containers``' result (one device -> host copy for the batch)."""
    if status is not None:
        check_status(status)
    o = offs.cpu().numpy()
    host = out[:int(o[-1])].cpu().numpy().tobytes()
    return [host[int(o[i]):int(o[i + 1])] for i in range(len(o) - 1)]


def compress_images_rgb(images, config):
    """``[Jpeg(config).compress(im.convert('YCbCr')) for im in images]`` (compress.py:9-17), byte-identical, as one
    batch: the RGB images cross to the GPU once through a pinned buffer (``np.asarray`` views, no Python lists --
    compare util.py:110-112), colour conversion, compression and container assembly happen on the device, and one
    copy brings every container back."""
    n = len(images)
    if n == 0:
        return []
    h, w = int(config.height), int(config.width)
    dev = _require_cuda()
    staging = torch.empty((n, h, w, 3), dtype=torch.uint8, pin_memory=True)
    view = staging.numpy()
    for i, im in enumerate(images):
        a = np.asarray(im if im.mode == "RGB" else im.convert("RGB"))
        if a.shape != (h, w, 3):
            raise ValueError("image %d is %s but config says %dx%d" % (i, a.shape, h, w))
        view[i] = a
    rgb = staging.to(dev, non_blocking=True)
    comp = compress_planes(rgb_to_ycbcr_planes(rgb), config)
    out, offs, status = pack_containers(comp, config)
    comp.check()
    return containers_to_bytes(out, offs, status)


class Jpeg:
    """pipeline.Jpeg (pipeline/__init__.py:98-124) on the CUDA path: the three bands of an
    image go through one batched launch each way; the container is the reference's."""

    def __init__(self, config):
        self.config = config

    def compress(self, image):
        bands = [np.asarray(b, dtype=np.uint8) for b in image.split()]   # y, cb, cr
        y, cb, cr = compress_bands(bands, self.config)
        return file_format.generate_data(self.config, file_format.CompressedData(y, cb, cr))

    @staticmethod
    def decompress(bytestream):
        from PIL import Image
        config, data = file_format.read_data(bytestream)
        planes = decompress_bands([data.y, data.cb, data.cr], config)
        # np.dstack(...).astype(np.uint8) -> Image.fromarray(mode='YCbCr'), :120-124
        return Image.merge("YCbCr", [Image.fromarray(np.ascontiguousarray(planes[i]), "L") for i in range(3)])

    # The two CLI flows with the colour conversion on the device as well:
    def compress_rgb(self, image):
        """``Jpeg(config).compress(image.convert('YCbCr'))`` of compress.py:9-17, byte-identical; the RGB image
        crosses to the GPU once, interleaved, and is converted there."""
        rgb = torch.from_numpy(np.array(image.convert("RGB"), dtype=np.uint8)).to(_require_cuda())
        comp = compress_planes(rgb_to_ycbcr_planes(rgb), self.config)
        y, cb, cr = comp.to_bytes_list()
        return file_format.generate_data(self.config, file_format.CompressedData(y, cb, cr))

    @staticmethod
    def decompress_rgb(bytestream):
        """``Jpeg.decompress(data).convert('RGB')`` of decompress.py:9-10, identical pixels."""
        from PIL import Image
        config, data = file_format.read_data(bytestream)
        dev = _require_cuda()
        streams = [data.y, data.cb, data.cr]
        blob = torch.from_numpy(np.frombuffer(b"".join(streams), dtype=np.uint8).copy()).to(dev)
        lens = torch.tensor([len(x) for x in streams], dtype=torch.int64)
        offs = (torch.cumsum(lens, 0) - lens).to(dev)
        planes, status = decompress_planes(blob, offs, lens.to(dev), config, 3, in_bytes=int(lens.sum()))
        rgb = ycbcr_planes_to_rgb(planes)
        check_status(status)
        return Image.fromarray(rgb[0].cpu().numpy(), "RGB")


class BatchCodec:
    """Repeated batched (de)compression of ``n_planes`` equally sized planes with buffers that
    are allocated once: device staging for the planes and the streams, pinned host buffers for
    the host-facing calls.  This is the call a service that compresses frame batches makes."""

    def __init__(self, config, n_planes, device=None, flags=0, pinned=True):
        lib = _lib.load()
        self.config, self.n_planes, self.flags = config, int(n_planes), flags
        self.device = _require_cuda(device)
        p = config.c_params(flags)
        g = _lib.jb_geometry()
        _raise_for_code(lib.jb_geometry_of(ctypes.byref(p), ctypes.byref(g)))
        self.h, self.w = int(config.height), int(config.width)
        self.cap = lib.jb_max_stream_bytes(ctypes.byref(p), self.n_planes)
        with torch.cuda.device(self.device):
            self.d_planes = torch.empty((self.n_planes, self.h, self.w), dtype=torch.uint8, device=self.device)
            self.d_streams = torch.empty(self.cap, dtype=torch.uint8, device=self.device)
            self.d_decoded = torch.empty((self.n_planes, self.h, self.w), dtype=torch.uint8, device=self.device)
            # private workspaces: tables and control block are set up by the first call and reused by the later
            # ones (JB_FLAG_REUSE_TABLES: same direction, parameters AND plane count, since the layout behind the
            # tables depends on it) -- hence one compress workspace per plane count (the sub-batches of
            # roundtrip_host come in at most two sizes); nothing else ever writes to them
            self._ws_fwd = {}
            self._fwd_workspace(self.n_planes)
            self._ws_inv = torch.empty(256, dtype=torch.uint8, device=self.device)      # sized by the first call
        self._tables_inv = False
        self.pinned = pinned
        self.h_streams = None
        self.h_decoded = None

    def _fwd_workspace(self, n_planes):
        """[workspace tensor, set-up done] for compress calls on n_planes planes."""
        ent = self._ws_fwd.get(n_planes)
        if ent is None:
            p = self.config.c_params(self.flags)
            need = _lib.load().jb_compress_workspace_bytes(ctypes.byref(p), int(n_planes))
            with torch.cuda.device(self.device):
                ent = [torch.empty(max(256, int(need)), dtype=torch.uint8, device=self.device), False]
            self._ws_fwd[n_planes] = ent
        return ent

    def _compress(self, planes, out):
        ent = self._fwd_workspace(int(planes.shape[0]))
        fl = self.flags | (_lib.JB_FLAG_REUSE_TABLES if ent[1] else 0)
        comp = compress_planes(planes, self.config, flags=fl, out=out, ws=ent[0])
        ent[1] = True
        return comp

    def _decompress(self, data, offsets, lengths, n_planes, in_bytes, out):
        p = self.config.c_params(self.flags)
        need = _lib.load().jb_decompress_workspace_bytes(ctypes.byref(p), int(n_planes), int(in_bytes))
        if self._ws_inv.numel() < need:                    # the layout depends on the stream bytes: grow, rebuild tables
            old = self._ws_inv
            with torch.cuda.device(self.device):
                self._ws_inv = torch.empty(int(need) + (int(need) >> 2), dtype=torch.uint8, device=self.device)
            # work queued on the current stream may still be using the old buffer: keep the allocator from
            # recycling it before that work is done
            old.record_stream(torch.cuda.current_stream(self.device))
            self._tables_inv = False
        fl = self.flags | (_lib.JB_FLAG_REUSE_TABLES if self._tables_inv else 0)
        res = decompress_planes(data, offsets, lengths, self.config, n_planes, in_bytes=in_bytes, flags=fl, out=out,
                                ws=self._ws_inv)
        self._tables_inv = True
        return res

    # -- device-resident --------------------------------------------------------------------
    def compress_device(self, planes=None):
        return self._compress(self.d_planes if planes is None else planes, self.d_streams)

    def capture_graphs(self):
        """Capture the two device-resident calls on this object's own buffers (``d_planes`` -> ``d_streams`` ->
        ``d_decoded``) into CUDA graphs, once; returns ``(compress_graph, decompress_graph, comp, status)``.

        For single frames the calls are launch bound: three kernels each way whose launches from Python cost more
        than they run.  Replaying ``compress_graph`` re-compresses whatever ``d_planes`` holds; ``decompress_graph``
        decodes the streams the last compress replay left (their byte count may not exceed the count of the
        capturing call -- it sizes the framing launch -- so capture with representative content)."""
        if getattr(self, "_graphs", None) is not None:
            return self._graphs
        with torch.cuda.device(self.device):
            comp = self.compress_device()                      # builds tables, sizes the decompress workspace
            total = comp.total_bytes()
            _, status = self.decompress_device(comp, total + (total >> 3) + 4096)
            check_status(status)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                      # warm the allocator pools of the capture stream
                c2 = self.compress_device()
                self.decompress_device(c2, total + (total >> 3) + 4096)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g_c, g_d = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_c):
                comp_g = self.compress_device()
            with torch.cuda.graph(g_d, pool=g_c.pool()):
                _, status_g = self.decompress_device(comp_g, total + (total >> 3) + 4096)
        self._graphs = (g_c, g_d, comp_g, status_g)
        return self._graphs

    def decompress_device(self, comp, total_bytes):
        lengths = comp.offsets[1:] - comp.offsets[:-1]
        return self._decompress(comp.data, comp.offsets[:-1], lengths, self.n_planes, int(total_bytes), self.d_decoded)

    # -- host buffers in, host buffers out ------------------------------------------------------
    def compress_host(self, h_planes):
        """h_planes: uint8 host tensor [n, H, W] (pinned for full copy speed).  Returns
        (host uint8 tensor holding the concatenated streams, int64 numpy offsets [n + 1])."""
        self.d_planes.copy_(h_planes, non_blocking=True)
        comp = self.compress_device()
        offsets = comp.host_offsets()                      # syncs, raises on a device-side error
        total = int(offsets[-1])
        if self.h_streams is None or self.h_streams.numel() < total:
            self.h_streams = torch.empty(max(total, 1), dtype=torch.uint8, pin_memory=self.pinned)
        out = self.h_streams[:total]
        out.copy_(comp.data[:total], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out, offsets

    def roundtrip_host(self, h_planes, n_sub=8):
        """compress_host followed by decompress_host, run as a two-stage pipeline over ``n_sub``
        sub-batches of planes: while sub-batch i is decompressed and copied back (device -> host),
        sub-batch i + 1 is copied in (host -> device) and compressed, so both directions of the
        host link are busy.  The streams of every sub-batch are copied to the host and read back
        from the host, exactly as in the two separate calls.

        Returns ``(decoded, parts)``: the uint8 host tensor [n, H, W] and, per sub-batch,
        ``(first_plane, host uint8 tensor with its streams, int64 numpy offsets)``."""
        lib = _lib.load()
        n_sub = max(1, min(int(n_sub), self.n_planes))
        bounds = [self.n_planes * j // n_sub for j in range(n_sub + 1)]
        p = self.config.c_params(self.flags)
        caps = [int(lib.jb_max_stream_bytes(ctypes.byref(p), bounds[j + 1] - bounds[j])) for j in range(n_sub)]
        cap0 = [0]
        for c in caps:
            cap0.append(cap0[-1] + ((c + 255) // 256) * 256)
        with torch.cuda.device(self.device):
            if getattr(self, "_rt", None) is None or self._rt["n_sub"] != n_sub:
                self._rt = {
                    "n_sub": n_sub, "sa": torch.cuda.Stream(), "sb": torch.cuda.Stream(),
                    "d_work": torch.empty(cap0[-1], dtype=torch.uint8, device=self.device),
                    "d_back": torch.empty(cap0[-1], dtype=torch.uint8, device=self.device),
                    "h_off": [torch.empty(bounds[j + 1] - bounds[j] + 1, dtype=torch.int64, pin_memory=self.pinned)
                              for j in range(n_sub)],
                    "h_st": [torch.empty(2 * _lib.JB_STATUS_WORDS, dtype=torch.int64, pin_memory=self.pinned)
                             for j in range(n_sub)],
                    "h_streams": None,
                }
            rt = self._rt
            # (allocated on the caller's stream, used on sa / sb: tell the caching allocator)
            rt["d_work"].record_stream(rt["sa"]); rt["d_work"].record_stream(rt["sb"]); rt["d_back"].record_stream(rt["sb"])
            if self.h_decoded is None:
                self.h_decoded = torch.empty((self.n_planes, self.h, self.w), dtype=torch.uint8, pin_memory=self.pinned)
            cur = torch.cuda.current_stream()
            sa, sb = rt["sa"], rt["sb"]
            sa.wait_stream(cur)
            sb.wait_stream(cur)
            nw = _lib.JB_STATUS_WORDS

            def stage_in(j):
                p0, p1 = bounds[j], bounds[j + 1]
                with torch.cuda.stream(sa):
                    self.d_planes[p0:p1].copy_(h_planes[p0:p1], non_blocking=True)
                    comp = self._compress(self.d_planes[p0:p1], rt["d_work"][cap0[j]:cap0[j] + caps[j]])
                    rt["h_off"][j].copy_(comp.offsets, non_blocking=True)
                    rt["h_st"][j][:nw].copy_(comp.status, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(sa)
                return ev

            def stage_out(j, ev):
                p0, p1 = bounds[j], bounds[j + 1]
                ev.synchronize()
                st = rt["h_st"][j]
                if int(st[0].item()):
                    _raise_for_code(-int(st[0].item()), int(st[1].item()) & 0xFFFFFFFFFFFFFFFF)
                offsets = rt["h_off"][j].numpy().copy()
                total = int(offsets[-1])
                if rt["h_streams"] is None:
                    rt["h_streams"] = torch.empty(cap0[-1], dtype=torch.uint8, pin_memory=self.pinned)
                h_part = rt["h_streams"][cap0[j]:cap0[j] + total]
                with torch.cuda.stream(sb):
                    h_part.copy_(rt["d_work"][cap0[j]:cap0[j] + total], non_blocking=True)        # streams -> host
                    d_in = rt["d_back"][cap0[j]:cap0[j] + caps[j]]
                    d_in[:total].copy_(h_part, non_blocking=True)                                 # host -> device
                    d_off = rt["h_off"][j].to(self.device, non_blocking=True)
                    out, status = self._decompress(d_in, d_off[:-1], d_off[1:] - d_off[:-1], p1 - p0, total,
                                                   self.d_decoded[p0:p1])
                    self.h_decoded[p0:p1].copy_(out, non_blocking=True)
                    st[nw:].copy_(status, non_blocking=True)
                return (p0, h_part, offsets)

            parts = []
            ev = stage_in(0)
            for j in range(n_sub):
                ev_next = stage_in(j + 1) if j + 1 < n_sub else None
                parts.append(stage_out(j, ev))
                ev = ev_next
            cur.wait_stream(sa)
            cur.wait_stream(sb)
            cur.synchronize()
            for j in range(n_sub):
                st = rt["h_st"][j]
                if int(st[nw].item()):
                    _raise_for_code(-int(st[nw].item()), int(st[nw + 1].item()) & 0xFFFFFFFFFFFFFFFF)
        return self.h_decoded, parts

    def decompress_host(self, h_streams, offsets):
        """Inverse of compress_host; returns a uint8 host tensor [n, H, W]."""
        total = int(offsets[-1])
        self.d_streams[:total].copy_(h_streams[:total], non_blocking=True)
        d_off = torch.from_numpy(np.ascontiguousarray(offsets, dtype=np.int64)).to(self.device, non_blocking=True)
        lengths = d_off[1:] - d_off[:-1]
        out, status = self._decompress(self.d_streams, d_off[:-1], lengths, self.n_planes, total, self.d_decoded)
        if self.h_decoded is None:
            self.h_decoded = torch.empty((self.n_planes, self.h, self.w), dtype=torch.uint8, pin_memory=self.pinned)
        self.h_decoded.copy_(out, non_blocking=True)
        check_status(status)                               # syncs
        torch.cuda.current_stream().synchronize()
        return self.h_decoded
