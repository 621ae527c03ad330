"""Host-side mirror of the reference's operator interface for the per-z-slice path.

Names, argument meaning and error behaviour follow crackle/codec.py (compress :689-733, decompress :616-687,
decompress_range :632-687) and src/fastcrackle.cpp (compress :163-210, decompress :84-129).  All compute goes
through the C-ABI (libcrackle_b200.so); numpy arrays / bytes in and out like the reference, plus torch CUDA tensors
for device-resident volumes (the benchmark's `value` path)."""
import ctypes
import threading
from typing import Optional, Tuple

import numpy as np

from . import _capi


class Context:
    """One per GPU: owns the CUDA stream and the reusable device workspace (ckl_ctx).

    A Context (like the ckl_ctx under it) is single-threaded: one call at a time.  The module-level compress /
    decompress functions serialise on a lock around their shared default context; use one Context per thread for
    concurrent callers.  torch CUDA tensors are ordered against torch's CURRENT stream: the context is bound to it for
    the call (and stays on it), so producers and consumers of the tensor need no extra synchronisation."""

    def __init__(self, device: int = 0):
        self._h = ctypes.c_void_p()
        L = _capi.lib()
        if L.ckl_device_count() <= 0:
            raise RuntimeError("crackle_b200: no CUDA device available (there is no CPU fallback)")
        rc = L.ckl_ctx_create(int(device), ctypes.byref(self._h))
        if rc:
            raise RuntimeError(f"crackle_b200: ckl_ctx_create failed ({rc})")
        self.device = int(device)

    def close(self):
        if self._h:
            _capi.lib().ckl_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            msg = _capi.lib().ckl_ctx_error(self._h).decode()
            if rc == 2:
                raise ValueError(msg)
            raise RuntimeError(msg)

    def bind_torch_stream(self):
        """run on torch's current CUDA stream of this device (what produced / will consume the caller's CUDA tensors)"""
        import torch
        sp = torch.cuda.current_stream(self.device).cuda_stream
        if getattr(self, "_bound_stream", None) != sp:
            self.set_stream(sp)

    # -- instrumentation -----------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        """Run on a caller-owned stream, e.g. torch.cuda.current_stream().cuda_stream (0 = legacy default stream);
        None restores the context's own stream."""
        if cuda_stream_ptr is None:
            self._check(_capi.lib().ckl_ctx_own_stream(self._h))
            self._bound_stream = None
        else:
            self._check(_capi.lib().ckl_ctx_set_stream(self._h, ctypes.c_void_p(int(cuda_stream_ptr))))
            self._bound_stream = int(cuda_stream_ptr)

    def set_chunks(self, chunks: int = 0):
        """z-chunk pipelining of compress / decompress: 0 = automatic (large volumes), 1 = off, K = always K chunks.
        Output is identical for every setting."""
        self._check(_capi.lib().ckl_ctx_set_chunks(self._h, int(chunks)))

    def prof_enable(self, on=True):
        _capi.lib().ckl_prof_enable(self._h, int(on))

    def prof_read(self) -> dict:
        """{stage: (total_ms, calls)} accumulated since prof_enable."""
        buf = ctypes.create_string_buffer(8192)
        _capi.lib().ckl_prof_read(self._h, buf, 8192)
        out = {}
        for item in buf.value.decode().split(";"):
            if item:
                k, v = item.split("=")
                ms, n = v.split(":")
                out[k] = (float(ms), int(n))
        return out

    # -- compress ------------------------------------------------------------------------------------------
    def compress_ptr(self, ptr, on_device, data_width, sx, sy, sz, fortran_order=True, markov_model_order=0) -> int:
        n = ctypes.c_uint64()
        self._check(_capi.lib().ckl_compress(self._h, ptr, int(on_device), int(data_width), sx, sy, sz,
                                             int(bool(fortran_order)), int(markov_model_order), ctypes.byref(n)))
        return n.value

    def result_bytes(self) -> bytes:
        n = ctypes.c_uint64()
        _capi.lib().ckl_result_device(self._h, ctypes.byref(n))
        buf = ctypes.create_string_buffer(n.value)
        self._check(_capi.lib().ckl_result_copy(self._h, buf, 0, n.value))
        return buf.raw

    def result_to(self, ptr, on_device, capacity):
        self._check(_capi.lib().ckl_result_copy(self._h, ptr, int(on_device), capacity))

    def result_device(self):
        n = ctypes.c_uint64()
        p = _capi.lib().ckl_result_device(self._h, ctypes.byref(n))
        return p, n.value

    def compress(self, labels, markov_model_order=0, fortran_order=None) -> bytes:
        """labels: numpy array (host) or torch CUDA tensor laid out (sz,sy,sx) C-contiguous (= Fortran x-fastest)."""
        ptr, on_dev, width, (sx, sy, sz), f_order, keep = _as_fortran_volume(labels)
        if fortran_order is not None:
            f_order = fortran_order
        if on_dev:
            self.bind_torch_stream()
        self.compress_ptr(ptr, on_dev, width, sx, sy, sz, f_order, markov_model_order)
        del keep
        return self.result_bytes()

    # -- decompress ----------------------------------------------------------------------------------------
    def decompress_into(self, bin_ptr, bin_on_device, nbytes, z_start, z_end, label, out_ptr, out_on_device, out_cap):
        self._check(_capi.lib().ckl_decompress(self._h, bin_ptr, int(bin_on_device), nbytes, int(z_start), int(z_end),
                                               int(label is not None), int(label or 0), out_ptr, int(out_on_device), out_cap))

    def decompress(self, binary, z_start=0, z_end=-1, label=None) -> np.ndarray:
        """fastcrackle.decompress(binary, z_start, z_end, parallel, label): returns a 1-D numpy array."""
        h = header(binary)
        szr = _clamp_range(h, z_start, z_end)
        buf = np.frombuffer(binary, dtype=np.uint8)
        dt = np.uint8 if label is not None else np.dtype(f"u{h['data_width']}")
        out = np.empty(h["sx"] * h["sy"] * szr, dtype=dt)
        self.decompress_into(buf.ctypes.data, 0, buf.size, z_start, z_end, label, out.ctypes.data, 0, out.nbytes)
        return out


    def voxel_connectivity_graph(self, binary, z_start=0, z_end=-1, connectivity=4) -> np.ndarray:
        """fastcrackle.voxel_connectivity_graph(binary, z_start, z_end, parallel, connectivity) (src/fastcrackle.cpp:538-565):
        uint8 array of shape (sx, sy, z_end - z_start), Fortran order; bits 00 -z +z -y +y -x +x, set = passable."""
        h = header(binary)
        szr = _clamp_range(h, z_start, z_end)
        buf = np.frombuffer(binary, dtype=np.uint8)
        out = np.empty((h["sx"], h["sy"], szr), dtype=np.uint8, order="F")
        self._check(_capi.lib().ckl_voxel_connectivity_graph(self._h, buf.ctypes.data, 0, buf.size, int(z_start), int(z_end),
                                                              int(connectivity), out.ctypes.data, 0, out.nbytes))
        return out

    def reencode(self, binary, markov_model_order: int) -> bytes:
        """ckl_reencode: the same stream with its crack codes re-coded at another markov order (no voxel decode)"""
        buf = np.frombuffer(binary, dtype=np.uint8)
        n = ctypes.c_uint64()
        self._check(_capi.lib().ckl_reencode(self._h, buf.ctypes.data, 0, buf.size, int(markov_model_order), ctypes.byref(n)))
        return self.result_bytes()

    def zstack(self, binaries) -> bytes:
        """ckl_zstack: flat-label, order-0 streams stacked along z on the device"""
        bufs = [np.frombuffer(b, dtype=np.uint8) for b in binaries]
        ptrs = (ctypes.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])
        sizes = (ctypes.c_uint64 * len(bufs))(*[b.size for b in bufs])
        n = ctypes.c_uint64()
        self._check(_capi.lib().ckl_zstack(self._h, len(bufs), ptrs, sizes, 0, ctypes.byref(n)))
        return self.result_bytes()

    def zslice(self, binary, z_start: int, z_end: int) -> bytes:
        """ckl_zslice: the stream of slices [z_start, z_end) with its own label table"""
        buf = np.frombuffer(binary, dtype=np.uint8)
        n = ctypes.c_uint64()
        self._check(_capi.lib().ckl_zslice(self._h, buf.ctypes.data, 0, buf.size, int(z_start), int(z_end), ctypes.byref(n)))
        return self.result_bytes()

    def label_stats(self, binary, z_start=0, z_end=-1):
        """ckl_label_stats: (labels u64[n], counts u64[n], sums u64[n,3], bbox u32[n,6]) for slices [z_start, z_end),
        indexed like the stream's sorted unique label table -- computed from the runs, no volume is painted."""
        buf = np.frombuffer(binary, dtype=np.uint8)
        n = num_labels(binary)
        labels, counts = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
        sums, bbox = np.zeros((n, 3), dtype=np.uint64), np.zeros((n, 6), dtype=np.uint32)
        nu = ctypes.c_uint64()
        self._check(_capi.lib().ckl_label_stats(self._h, buf.ctypes.data, 0, buf.size, int(z_start), int(z_end), labels.ctypes.data,
                                                counts.ctypes.data, sums.ctypes.data, bbox.ctypes.data, 0, n, ctypes.byref(nu)))
        return labels, counts, sums, bbox


def launch_count() -> int:
    return int(_capi.lib().ckl_launch_count())


_default: Optional[Context] = None
_default_lock = threading.RLock()      # the module-level API shares one context: one call at a time (ctypes drops the GIL)


def sync_count() -> int:
    """host-side waits on a compute stream issued by the library since load (ckl_sync_count)"""
    return int(_capi.lib().ckl_sync_count())


def default_context() -> Context:
    global _default
    with _default_lock:
        return _default_context_locked()


def _default_context_locked() -> Context:
    global _default
    if _default is None:
        dev = 0
        try:
            import torch
            if torch.cuda.is_available():
                dev = torch.cuda.current_device()
        except Exception:
            pass
        _default = Context(dev)
    return _default


def _as_fortran_volume(labels):
    """-> (pointer, on_device, itemsize, (sx,sy,sz), f_order flag, keepalive)"""
    if isinstance(labels, np.ndarray):
        if np.issubdtype(labels.dtype, np.signedinteger):
            raise TypeError("Signed integer data types are not currently supported.")   # codec.py:720-721
        if labels.dtype.kind not in "ub":
            raise TypeError("crackle_b200: labels must be an unsigned integer array")
        f_order = labels.flags.f_contiguous
        a = np.asfortranarray(labels)                                                   # codec.py:723-724
        s = list(a.shape) + [1, 1, 1]
        return a.ctypes.data, 0, a.dtype.itemsize, (s[0], s[1], s[2]), f_order, a
    import torch
    if isinstance(labels, torch.Tensor):
        if not labels.is_cuda or not labels.is_contiguous():
            raise TypeError("crackle_b200: torch input must be a contiguous CUDA tensor shaped (sz, sy, sx)")
        if labels.dtype not in (torch.uint8, torch.uint16, torch.uint32, torch.uint64):
            raise TypeError("Signed integer data types are not currently supported.")
        s = [1, 1, 1] + list(labels.shape)
        sz, sy, sx = s[-3], s[-2], s[-1]
        return labels.data_ptr(), 1, labels.element_size(), (sx, sy, sz), True, labels
    raise TypeError("crackle_b200: unsupported input type")


def header(binary) -> dict:
    buf = np.frombuffer(binary, dtype=np.uint8)
    info = _capi.HeaderInfo()
    err = ctypes.create_string_buffer(256)
    rc = _capi.lib().crackle_b200_header(buf.ctypes.data if buf.size else None, buf.size, ctypes.byref(info), err, 256)
    if rc:
        raise RuntimeError(err.value.decode())
    return {n: getattr(info, n) for n, _ in _capi.HeaderInfo._fields_}


def _clamp_range(h, z_start, z_end):
    sz = h["sz"]
    zs = max(min(z_start, sz - 1), 0)
    ze = sz if z_end < 0 else max(min(z_end, sz), 0)
    if zs >= ze:
        if h["sx"] * h["sy"] * sz == 0:
            return 0
        raise RuntimeError(f"crackle: Invalid range: {zs} - {ze}")
    return ze - zs


# ---- the reference's public operator interface ---------------------------------------------------------------
def compress(labels, allow_pins: int = 0, markov_model_order: int = 0, bgcolor: Optional[int] = None,
             parallel: int = 0) -> bytes:
    """crackle.compress (codec.py:689-733).  Flat labels only: allow_pins must be 0 on this path."""
    if allow_pins:
        raise NotImplementedError("crackle_b200: pin label formats are outside the flat-label hot path; "
                                  "use the reference for allow_pins != 0")
    with _default_lock:
        return default_context().compress(labels, markov_model_order)


def decompress_range(binary, z_start: Optional[int], z_end: Optional[int], parallel: int = 0,
                     label: Optional[int] = None) -> np.ndarray:
    """crackle.decompress_range (codec.py:632-687)."""
    h = header(binary)
    sx, sy, sz = h["sx"], h["sy"], h["sz"]
    z_start = 0 if z_start is None else int(z_start)
    z_end = sz if z_end is None else int(z_end)
    order = "F" if h["fortran_order"] else "C"
    dtype = np.dtype(f"u{h['data_width']}")
    if sx * sy * sz == 0:
        return np.zeros((0,), dtype=dtype).reshape((sx, sy, max(z_end - z_start, 0)), order=order)
    with _default_lock:
        out = default_context().decompress(binary, z_start, z_end, label)
    szr = out.size // (sx * sy)
    out = out.reshape((sx, sy, szr), order=order)
    if label is not None:
        return out.view(bool)
    if h["is_signed"]:
        out = out.view(np.dtype(f"i{h['data_width']}"))
    return out


def reencode(binary, markov_model_order: int, parallel: int = 0) -> bytes:
    """crackle.codec.reencode (codec.py:877-881)"""
    if header(binary)["markov_model_order"] == markov_model_order:
        return binary
    with _default_lock:
        return default_context().reencode(binary, markov_model_order)


def zstack(images) -> bytes:
    """crackle.zstack (operations.py:424-548): arrays or streams of equal width and height stacked along z into one
    stream, without decoding the streams.  Arrays are compressed first, streams re-coded to markov order 0 first."""
    binaries = []
    for img in images:
        if img is None:
            continue
        if isinstance(img, np.ndarray):
            b = compress(img)
        else:
            b = reencode(bytes(getattr(img, "binary", img)), 0)
        if header(b)["sx"] * header(b)["sy"] * header(b)["sz"] == 0:
            continue
        binaries.append(b)
    if len(binaries) == 1:
        return binaries[0]
    with _default_lock:
        return default_context().zstack(binaries)


def zsplit(binary, z: int):
    """crackle.zsplit (operations.py:617-640): (before, middle, after) streams around slice z; an empty side is b''."""
    sz = header(binary)["sz"]
    if z < 0 or z >= sz:
        raise ValueError(f"{z} is outside the range 0 to {sz}.")
    if sz == 1:
        return (b"", binary, b"")
    with _default_lock:
        ctx = default_context()
        return (ctx.zslice(binary, 0, z) if z > 0 else b"", ctx.zslice(binary, z, z + 1),
                ctx.zslice(binary, z + 1, sz) if z + 1 < sz else b"")


def zshatter(binary):
    """crackle.zshatter (operations.py:642-662): one stream per slice."""
    sz = header(binary)["sz"]
    with _default_lock:
        ctx = default_context()
        return [ctx.zslice(binary, z, z + 1) for z in range(sz)]


def _labels_section(binary):
    h = header(binary)
    buf = np.frombuffer(binary, dtype=np.uint8)
    hb = 24 if h["format_version"] == 0 else 29
    off = hb + 4 * (h["sz"] + (0 if h["format_version"] == 0 else 1))
    return h, buf[off: off + h["num_label_bytes"]]


def num_labels(binary) -> int:
    """crackle.num_labels (codec.py:82-96), flat label format"""
    h, lab = _labels_section(binary)
    if h["sx"] * h["sy"] * h["sz"] == 0:
        return 0
    return int.from_bytes(lab[:8].tobytes(), "little")


def labels(binary) -> np.ndarray:
    """crackle.labels (codec.py:17-80): the sorted unique labels of the stream"""
    h, lab = _labels_section(binary)
    if h["sx"] * h["sy"] * h["sz"] == 0:
        return np.zeros((0,), dtype=np.dtype(f"u{h['data_width']}"))
    n = int.from_bytes(lab[:8].tobytes(), "little")
    return np.frombuffer(lab, dtype=f"<u{h['stored_data_width']}", count=n, offset=8).astype(np.dtype(f"u{h['data_width']}"))


def contains(binary, label: int) -> bool:
    """crackle.contains (codec.py:98-132)"""
    u = labels(binary)
    if label < 0 or u.size == 0 or label > int(u[-1]):
        return False
    i = int(np.searchsorted(u, np.asarray(label, dtype=u.dtype)))
    return i < u.size and int(u[i]) == int(label)


def _stats_range(binary, label):
    if label is None:
        return 0, -1
    if not contains(binary, label):
        raise ValueError(f"Label {label} not contained in image.")
    return z_range_for_label(binary, label)


def voxel_counts(binary, label: Optional[int] = None, parallel: int = 0):
    """crackle.voxel_counts (codec.py:949-982 -> operations.hpp:321-371): {label: voxels} (an int when `label` is given)."""
    z_start, z_end = _stats_range(binary, label)
    if num_labels(binary) == 1:
        h = header(binary)
        vcts = {int(labels(binary)[0]): h["sx"] * h["sy"] * h["sz"]}
    else:
        with _default_lock:
            lab, cnt, _, _ = default_context().label_stats(binary, z_start, z_end)
        keep = cnt > 0
        vcts = dict(zip(lab[keep].tolist(), cnt[keep].tolist()))
    return vcts[label] if label is not None else vcts


def centroids(binary, label: Optional[int] = None, parallel: int = 0):
    """crackle.centroids (codec.py:984-1007 -> operations.hpp:421-491): {label: [x, y, z]} (float64, sums / count)."""
    z_start, z_end = _stats_range(binary, label)
    with _default_lock:
        lab, cnt, sums, _ = default_context().label_stats(binary, z_start, z_end)
    keep = cnt > 0
    cen = sums[keep].astype(np.float64) / cnt[keep].astype(np.float64)[:, None]
    out = {int(k): [float(v[0]), float(v[1]), float(v[2])] for k, v in zip(lab[keep], cen)}
    return out[label] if label is not None else out


def bounding_boxes(binary, label: Optional[int] = None, parallel: int = 0, no_slice_conversion: bool = False):
    """crackle.bounding_boxes (codec.py:1009-1065 -> operations.hpp:541-617): {label: [xmin,ymin,zmin,xmax,ymax,zmax]} or
    slices.  Like the reference, every label of the stream has an entry; one that is absent from the decoded z-range keeps
    the initial (2^32-1, 2^32-1, 2^32-1, 0, 0, 0)."""
    z_start, z_end = _stats_range(binary, label)
    if num_labels(binary) == 1:
        h = header(binary)
        boxes = {int(labels(binary)[0]): np.array([0, 0, 0, h["sx"], h["sy"], h["sz"]], dtype=np.uint32)}
    else:
        with _default_lock:
            lab, _, _, bbox = default_context().label_stats(binary, z_start, z_end)
        boxes = {int(k): bbox[i].copy() for i, k in enumerate(lab)}
    if no_slice_conversion:
        return boxes[label] if label is not None else boxes
    if label is not None:
        boxes = {label: boxes[label]}
    boxes = {k: (slice(int(b[0]), int(b[3]) + 1), slice(int(b[1]), int(b[4]) + 1), slice(int(b[2]), int(b[5]) + 1))
             for k, b in boxes.items()}
    return boxes[label] if label is not None else boxes


def voxel_connectivity_graph(binary, connectivity: int = 6, parallel: int = 0) -> np.ndarray:
    """crackle.voxel_connectivity_graph (operations.py:936-954): uint8 (sx, sy, sz) Fortran-order array,
    bitset (right hand side is LSB) 00-z+z-y+y-x+x."""
    if connectivity not in (4, 6):
        raise ValueError(f"Only 4 and 6 connected are supported. Got: {connectivity}")
    with _default_lock:
        return default_context().voxel_connectivity_graph(binary, 0, -1, connectivity)


def z_range_for_label(binary, label: int) -> Tuple[int, int]:
    """crackle.codec.z_range_for_label_flat (codec.py:464-520): the z-range whose slices can contain `label`, from the
    label table alone (no voxel work); (-1, -1) when the label does not occur."""
    h = header(binary)
    if h["label_format"] != 0:
        raise ValueError("Label format not supported.")
    buf = np.frombuffer(binary, dtype=np.uint8)
    hb = 24 if h["format_version"] == 0 else 29
    off = hb + 4 * (h["sz"] + (0 if h["format_version"] == 0 else 1))
    lab = buf[off: off + h["num_label_bytes"]]
    nu = int.from_bytes(lab[:8].tobytes(), "little")
    sw = h["stored_data_width"]
    uniq = np.frombuffer(lab, dtype=f"<u{sw}", count=nu, offset=8)
    if label < 0 or label > np.iinfo(uniq.dtype).max:
        return (-1, -1)
    idx = int(np.searchsorted(uniq, np.asarray(label, dtype=uniq.dtype)))
    if idx >= nu or int(uniq[idx]) != int(label):
        return (-1, -1)
    cw = 1 if h["sx"] * h["sy"] <= 0xFF else 2 if h["sx"] * h["sy"] <= 0xFFFF else 4 if h["sx"] * h["sy"] <= 0xFFFFFFFF else 8
    o = 8 + nu * sw
    per_grid = np.cumsum(np.frombuffer(lab, dtype=f"<u{cw}", count=h["sz"], offset=o))
    kw = 1 if nu <= 0xFF else 2 if nu <= 0xFFFF else 4 if nu <= 0xFFFFFFFF else 8
    o += cw * h["sz"]
    keys = np.frombuffer(lab, dtype=f"<u{kw}", count=(len(lab) - o) // kw, offset=o)
    hits = np.flatnonzero(keys == idx)                 # fastcrackle.index_range
    if hits.size == 0:
        return (-1, -1)
    min_cc, max_cc = int(hits[0]), int(hits[-1])
    z_start = int(np.searchsorted(per_grid, min_cc))
    z_end = int(np.searchsorted(per_grid, max_cc))
    if per_grid[z_start] == min_cc:
        z_start = min(z_start + 1, h["sz"] - 1)
    if per_grid[z_end] == max_cc:
        z_end = min(z_end + 1, h["sz"] - 1)
    return (z_start, z_end + 1)


def decompress_binary_image(binary, label: int, parallel: int = 0, crop: bool = True) -> np.ndarray:
    """crackle.codec.decompress_binary_image (codec.py:588-614): only the z-range that can hold `label` is decoded."""
    z_start, z_end = z_range_for_label(binary, label)
    h = header(binary)
    order = "F" if h["fortran_order"] else "C"
    if z_start == -1 and z_end == -1 and crop:
        return np.zeros([0, 0, 0], dtype=bool, order=order)
    if (z_start == 0 and z_end == h["sz"]) or crop:
        return decompress_range(binary, z_start, z_end, parallel, label).view(bool)
    image = np.zeros([h["sx"], h["sy"], h["sz"]], dtype=bool, order=order)
    if z_start == -1 and z_end == -1:
        return image
    image[:, :, z_start:z_end] = decompress_range(binary, z_start, z_end, parallel, label)
    return image


def decompress(binary, label: Optional[int] = None, parallel: int = 0, crop: bool = False) -> np.ndarray:
    """crackle.decompress (codec.py:616-630)."""
    if label is None:
        return decompress_range(binary, None, None, parallel)
    return decompress_binary_image(binary, label, parallel, crop=crop)
