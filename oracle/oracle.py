"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and loader for the compiled reference
(oracle/_ref/_fastcrackle_ref).  TEST INFRASTRUCTURE ONLY: imported by tests/, bench.py's cpu_baseline /
--impl reference legs and __graft_entry__.smoke() -- never by the product package crackle_b200."""
import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "crackle_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-fvisibility=hidden", "-shared", "-fPIC", "-o", so, src])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        u64, i64, vp = ctypes.c_uint64, ctypes.c_int64, ctypes.c_void_p
        L.ckl_oracle_compress.argtypes = [vp, ctypes.c_int, u64, u64, u64, ctypes.c_int, ctypes.c_int,
                                          ctypes.POINTER(vp), ctypes.POINTER(u64)]
        L.ckl_oracle_compress.restype = ctypes.c_int
        L.ckl_oracle_decompress.argtypes = [vp, u64, i64, i64, ctypes.c_int, u64, vp, u64]
        L.ckl_oracle_decompress.restype = ctypes.c_int
        L.ckl_oracle_slice_ccl.argtypes = [vp, ctypes.c_int, u64, u64, vp]
        L.ckl_oracle_slice_ccl.restype = u64
        L.ckl_oracle_slice_crack_code.argtypes = [vp, ctypes.c_int, u64, u64, ctypes.c_int,
                                                  ctypes.POINTER(vp), ctypes.POINTER(u64)]
        L.ckl_oracle_crc32c.argtypes = [vp, u64]
        L.ckl_oracle_crc32c.restype = ctypes.c_uint32
        L.ckl_oracle_crc8.argtypes = [vp, u64]
        L.ckl_oracle_crc8.restype = ctypes.c_uint8
        L.ckl_oracle_free.argtypes = [vp]
        _LIB = L
    return _LIB


def _shape3(a):
    s = list(a.shape) + [1, 1, 1]
    return s[0], s[1], s[2]


def compress(labels, markov_model_order=0, fortran_order=None):
    """Mirror of crackle.compress(labels, allow_pins=0, markov_model_order) (codec.py:689-733)."""
    if np.issubdtype(labels.dtype, np.signedinteger):
        raise TypeError("Signed integer data types are not currently supported.")
    f_order = labels.flags.f_contiguous if fortran_order is None else fortran_order
    a = np.asfortranarray(labels)
    sx, sy, sz = _shape3(a)
    out, n = ctypes.c_void_p(), ctypes.c_uint64()
    rc = lib().ckl_oracle_compress(a.ctypes.data, a.dtype.itemsize, sx, sy, sz, int(bool(f_order)),
                                   int(markov_model_order), ctypes.byref(out), ctypes.byref(n))
    if rc:
        raise RuntimeError(f"oracle compress failed: {rc}")
    b = ctypes.string_at(out.value, n.value)
    lib().ckl_oracle_free(out)
    return b


def header(binary):
    b = bytes(binary[:29])
    fmt = int.from_bytes(b[5:7], "little")
    return dict(version=b[4], data_width=1 << (fmt & 3), stored_width=1 << ((fmt >> 2) & 3), crack_format=(fmt >> 4) & 1,
                label_format=(fmt >> 5) & 3, fortran=(fmt >> 7) & 1, signed=(fmt >> 8) & 1, order=(fmt >> 9) & 15,
                sx=int.from_bytes(b[7:11], "little"), sy=int.from_bytes(b[11:15], "little"),
                sz=int.from_bytes(b[15:19], "little"), num_label_bytes=int.from_bytes(b[20:28], "little"))


def decompress(binary, z_start=0, z_end=-1, label=None):
    """Mirror of fastcrackle.decompress(binary, z_start, z_end, parallel, label) + the reshape of
    codec.py:672 (returns an (sx,sy,szr) array in the stream's memory order)."""
    h = header(binary)
    sz = h["sz"]
    zs = max(min(z_start, sz - 1), 0)
    ze = sz if z_end < 0 else max(min(z_end, sz), 0)
    if zs >= ze:
        raise RuntimeError(f"crackle: Invalid range: {zs} - {ze}")
    dt = np.uint8 if label is not None else np.dtype(f"u{h['data_width']}")
    out = np.zeros(h["sx"] * h["sy"] * (ze - zs), dtype=dt)
    buf = np.frombuffer(binary, dtype=np.uint8)
    rc = lib().ckl_oracle_decompress(buf.ctypes.data, buf.size, z_start, z_end, int(label is not None),
                                     int(label or 0), out.ctypes.data, out.nbytes)
    if rc:
        raise RuntimeError(f"oracle decompress failed: {rc}")
    return out.reshape((h["sx"], h["sy"], ze - zs), order="F" if h["fortran"] else "C")


def slice_ccl(slice2d):
    a = np.asfortranarray(slice2d)
    cc = np.zeros(a.shape, dtype=np.uint32, order="F")
    n = lib().ckl_oracle_slice_ccl(a.ctypes.data, a.dtype.itemsize, a.shape[0], a.shape[1], cc.ctypes.data)
    return cc, int(n)


def slice_crack_code(slice2d, permissible):
    a = np.asfortranarray(slice2d)
    out, n = ctypes.c_void_p(), ctypes.c_uint64()
    lib().ckl_oracle_slice_crack_code(a.ctypes.data, a.dtype.itemsize, a.shape[0], a.shape[1], int(permissible),
                                      ctypes.byref(out), ctypes.byref(n))
    b = ctypes.string_at(out.value, n.value)
    lib().ckl_oracle_free(out)
    return b


def crc32c(data):
    b = np.frombuffer(bytes(data), dtype=np.uint8)
    return int(lib().ckl_oracle_crc32c(b.ctypes.data if b.size else None, b.size))


def sections(binary):
    """Split a v1 .ckl stream into its sections so parity failures localise."""
    h = header(binary)
    sz = h["sz"]
    b = bytes(binary)
    if len(b) == 29:
        return dict(header=b, z_index=b"", labels=b"", model=b"", codes=[], labels_crc=b"", slice_crcs=b"")
    zi = b[29:29 + 4 * (sz + 1)]
    p = 29 + 4 * (sz + 1)
    lab = b[p:p + h["num_label_bytes"]]
    p += h["num_label_bytes"]
    mb = 0 if h["order"] == 0 else (4 ** h["order"] * 5 + 4) // 8
    model = b[p:p + mb]
    p += mb
    sizes = np.frombuffer(zi[:4 * sz], dtype="<u4")
    codes = []
    for s in sizes:
        codes.append(b[p:p + int(s)])
        p += int(s)
    return dict(header=b[:29], z_index=zi, labels=lab, model=model, codes=codes, labels_crc=b[p:p + 4],
                slice_crcs=b[p + 4:])


def ref_module():
    """The compiled, unmodified reference (None when oracle/_ref is absent)."""
    d = os.path.join(_HERE, "_ref")
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        import _fastcrackle_ref
        return _fastcrackle_ref
    except Exception:
        return None


def ref_compress(labels, markov_model_order=0, parallel=1):
    m = ref_module()
    f_order = labels.flags.f_contiguous
    return m.compress(np.asfortranarray(labels), False, f_order, markov_model_order, False, True, 0, parallel)


def ref_decompress(binary, z_start=0, z_end=-1, label=None, parallel=1):
    m = ref_module()
    h = header(binary)
    out = m.decompress(binary, z_start, z_end, parallel, label)
    szr = out.size // max(1, h["sx"] * h["sy"])
    return out.reshape((h["sx"], h["sy"], szr), order="F" if h["fortran"] else "C")
