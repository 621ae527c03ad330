"""INTEGRATION.md Option B, as a module: the name `fastcrackle` the reference's Python package imports
(crackle/codec.py:8, crackle/operations.py:9).  The hot path -- compress / decompress of flat-label streams, and the
compressed-domain statistics voxel_counts / centroids / bounding_boxes and voxel_connectivity_graph -- goes to the B200 library (the in-tree pybind11
module crackle_b200.fastcrackle and the C-ABI); every other name of src/fastcrackle.cpp:641-669, and the pin label formats
this path does not cover, come from the reference's own compiled module (oracle/_ref/_fastcrackle_ref).

Test infrastructure: with CKL_SHIM_BACKEND=ref everything is routed to the reference (validates the harness on a box
without a GPU)."""
import os

import numpy as np

import _fastcrackle_ref as _ref
from _fastcrackle_ref import *  # noqa: F401,F403  reencode_markov, remap, index_range, connected_components, compute_pins, ...
from _fastcrackle_ref import CppPin  # noqa: F401

_USE_GPU = os.environ.get("CKL_SHIM_BACKEND", "b200") != "ref"
CALLS = {"b200": 0, "ref": 0}
if _USE_GPU:
    import crackle_b200 as _cb
    from crackle_b200 import fastcrackle as _gpu
    __backend__ = _gpu.__backend__
else:
    __backend__ = "reference"


def _flat(binary) -> bool:
    b = bytes(binary[:7])
    return len(b) >= 7 and ((int.from_bytes(b[5:7], "little") >> 5) & 3) == 0


def compress(labels, allow_pins=False, fortran_order=True, markov_model_order=0, optimize_pins=False, auto_bgcolor=True,
             manual_bgcolor=0, parallel=1):
    if _USE_GPU and not allow_pins and labels.dtype.kind == "u" and labels.size > 0:
        CALLS["b200"] += 1
        return _gpu.compress(labels, False, fortran_order, markov_model_order, optimize_pins, auto_bgcolor, manual_bgcolor, parallel)
    CALLS["ref"] += 1
    return _ref.compress(labels, allow_pins, fortran_order, markov_model_order, optimize_pins, auto_bgcolor, manual_bgcolor, parallel)


def decompress(buffer, z_start=0, z_end=-1, parallel=1, label=None):
    if _USE_GPU and _flat(buffer):
        CALLS["b200"] += 1
        return _gpu.decompress(buffer, z_start, z_end, parallel, label)
    CALLS["ref"] += 1
    return _ref.decompress(buffer, z_start, z_end, parallel, label)


def _stats(binary, z_start, z_end):
    return _cb.default_context().label_stats(bytes(binary), z_start, z_end)


def voxel_counts(binary, z_start=-1, z_end=-1, parallel=1):
    if _USE_GPU and _flat(binary):
        CALLS["b200"] += 1
        lab, cnt, _, _ = _stats(binary, z_start, z_end)
        return {int(k): int(c) for k, c in zip(lab, cnt) if c}
    CALLS["ref"] += 1
    return _ref.voxel_counts(binary, z_start, z_end, parallel)


def centroids(binary, z_start=-1, z_end=-1, parallel=1):
    if _USE_GPU and _flat(binary):
        CALLS["b200"] += 1
        lab, cnt, sums, _ = _stats(binary, z_start, z_end)
        keep = cnt > 0
        cen = sums[keep].astype(np.float64) / cnt[keep].astype(np.float64)[:, None]
        return {int(k): np.array(v) for k, v in zip(lab[keep], cen)}
    CALLS["ref"] += 1
    return _ref.centroids(binary, z_start, z_end, parallel)


def bounding_boxes(binary, z_start=-1, z_end=-1, parallel=1):
    if _USE_GPU and _flat(binary):
        CALLS["b200"] += 1
        lab, _, _, bbox = _stats(binary, z_start, z_end)
        return {int(k): bbox[i].copy() for i, k in enumerate(lab)}
    CALLS["ref"] += 1
    return _ref.bounding_boxes(binary, z_start, z_end, parallel)


def voxel_connectivity_graph(buffer, z_start=0, z_end=-1, parallel=1, connectivity=4):
    if _USE_GPU and _flat(buffer):
        CALLS["b200"] += 1
        return _cb.default_context().voxel_connectivity_graph(bytes(buffer), z_start, z_end, connectivity)
    CALLS["ref"] += 1
    return _ref.voxel_connectivity_graph(buffer, z_start, z_end, parallel, connectivity)
