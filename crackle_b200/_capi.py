"""ctypes binding of the C-ABI in include/crackle_b200.h (libcrackle_b200.so).

The product path: there is no CPU fallback.  If the library is missing it is built in-tree (nvcc); if no CUDA
device is present every compute entry point raises RuntimeError."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libcrackle_b200.so")
_lib = None

u64, i64, u32, vp, cint = ctypes.c_uint64, ctypes.c_int64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int


class HeaderInfo(ctypes.Structure):
    _fields_ = [(n, u32) for n in ("format_version", "label_format", "crack_format", "is_signed", "data_width",
                                   "stored_data_width", "fortran_order", "markov_model_order", "sx", "sy", "sz",
                                   "is_sorted")] + [("num_label_bytes", u64)]


class ShardSummary(ctypes.Structure):
    _fields_ = [("max_label", u64), ("pairs", u64), ("first_voxel", u64), ("last_voxel", u64), ("voxels", u64),
                ("reserved", u64 * 3)]


class ShardPieces(ctypes.Structure):
    _fields_ = [("keys_bytes", u64), ("codes_bytes", u64), ("sz_local", u64)]


class ShardCounts(ctypes.Structure):
    _fields_ = [(n, u64) for n in ("n_unique_local", "n_components", "n_codepoints", "codes_bytes_order0", "sz_local", "runs",
                                   "keys_bytes", "codes_bytes")]


class ShardBlock(ctypes.Structure):
    _fields_ = [(n, u64) for n in ("offset", "sz_local", "n_components", "keys_bytes", "codes_bytes")]


# every symbol include/crackle_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "crackle_b200_compress": (cint, [vp, cint, u64, u64, u64, cint, cint, ctypes.POINTER(vp), ctypes.POINTER(u64),
                                     ctypes.c_char_p, ctypes.c_size_t]),
    "crackle_b200_decompress": (cint, [vp, u64, i64, i64, cint, u64, vp, u64, ctypes.c_char_p, ctypes.c_size_t]),
    "crackle_b200_free": (None, [vp]),
    "crackle_b200_header": (cint, [vp, u64, ctypes.POINTER(HeaderInfo), ctypes.c_char_p, ctypes.c_size_t]),
    "crackle_b200_version": (ctypes.c_char_p, []),
    "ckl_ctx_create": (cint, [cint, ctypes.POINTER(vp)]),
    "ckl_ctx_destroy": (None, [vp]),
    "ckl_ctx_error": (ctypes.c_char_p, [vp]),
    "ckl_device_count": (cint, []),
    "ckl_compress": (cint, [vp, vp, cint, cint, u64, u64, u64, cint, cint, ctypes.POINTER(u64)]),
    "ckl_result_copy": (cint, [vp, vp, cint, u64]),
    "ckl_result_device": (vp, [vp, ctypes.POINTER(u64)]),
    "ckl_decompress": (cint, [vp, vp, cint, u64, i64, i64, cint, u64, vp, cint, u64]),
    "ckl_label_stats": (cint, [vp, vp, cint, u64, i64, i64, vp, vp, vp, vp, cint, u64, ctypes.POINTER(u64)]),
    "ckl_voxel_connectivity_graph": (cint, [vp, vp, cint, u64, i64, i64, cint, vp, cint, u64]),
    "ckl_reencode": (cint, [vp, vp, cint, u64, cint, ctypes.POINTER(u64)]),
    "ckl_zstack": (cint, [vp, cint, ctypes.POINTER(vp), ctypes.POINTER(u64), cint, ctypes.POINTER(u64)]),
    "ckl_zslice": (cint, [vp, vp, cint, u64, u64, u64, ctypes.POINTER(u64)]),
    "ckl_shard_begin": (cint, [vp, vp, cint, cint, u64, u64, u64, ctypes.POINTER(ShardSummary)]),
    "ckl_shard_encode": (cint, [vp, cint, cint, cint, ctypes.POINTER(u64), ctypes.POINTER(u64), ctypes.POINTER(u64)]),
    "ckl_shard_encode_async": (cint, [vp, cint, cint, cint, ctypes.POINTER(u64), ctypes.POINTER(u64)]),
    "ckl_shard_encode_wait": (cint, [vp, ctypes.POINTER(u64), ctypes.POINTER(u64)]),
    "ckl_shard_unique": (cint, [vp, vp, cint]),
    "ckl_shard_stats": (cint, [vp, vp, cint]),
    "ckl_shard_finish": (cint, [vp, vp, cint, u64, vp, cint, ctypes.POINTER(ShardPieces)]),
    "ckl_shard_model": (cint, [vp, vp, cint, u64, ctypes.POINTER(u64)]),
    "ckl_shard_fetch": (cint, [vp, vp, vp, vp, vp, vp, cint]),
    "ckl_shard_info": (cint, [vp, ctypes.POINTER(ShardCounts)]),
    "ckl_shard_pack": (cint, [vp, vp, u64, ctypes.POINTER(u64)]),
    "ckl_shard_assemble": (cint, [vp, vp, ctypes.POINTER(ShardBlock), cint, vp, cint, u64, cint, cint, cint, cint, cint, u64, u64,
                                  ctypes.POINTER(u64)]),
    "ckl_prof_enable": (cint, [vp, cint]),
    "ckl_prof_read": (cint, [vp, ctypes.c_char_p, ctypes.c_size_t]),
    "ckl_launch_count": (u64, []),
    "ckl_sync_count": (u64, []),
    "ckl_ctx_set_stream": (cint, [vp, vp]),
    "ckl_ctx_own_stream": (cint, [vp]),
    "ckl_ctx_set_chunks": (cint, [vp, cint]),
    "ckl_crc32c": (cint, [vp, vp, cint, u64, ctypes.POINTER(u32)]),
    "ckl_sort_unique_u64": (cint, [vp, vp, u64, cint, ctypes.POINTER(u64)]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIBPATH):
            from . import build
            build.build_lib()
        L = ctypes.CDLL(_LIBPATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
