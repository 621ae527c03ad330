"""Format version 0 streams (24-byte header, uint32 num_label_bytes, no crc8, z-index without crc, no trailing crcs:
header.hpp:113-131, 168-184; crc checks skipped crackle.hpp:276, 566, 599).  The encoder only writes version 1, so the v0
streams are derived from the golden v1 streams by dropping the version-1 fields."""
import numpy as np
import pytest

from conftest import load_golden

NAMES = ["voronoi_u64_96x80x5"]


def to_v0(b):
    b = bytes(b)
    sz = int.from_bytes(b[15:19], "little")
    nlb = int.from_bytes(b[20:28], "little")
    head = b[:4] + bytes([0]) + b[5:20] + int(nlb).to_bytes(4, "little")
    assert len(head) == 24
    return head + b[29:29 + 4 * sz] + b[29 + 4 * sz + 4: len(b) - 4 - 4 * sz]


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("order", [0, 5])
def test_oracle_decodes_v0(name, order):
    from oracle import oracle as O
    g = load_golden(name)
    a = np.asfortranarray(g["input"])
    v0 = to_v0(g[f"ckl_order{order}"])
    assert np.array_equal(np.asarray(O.decompress(v0)).reshape(a.shape, order="F"), a)
    ref = O.ref_module()
    if ref is not None:
        assert np.array_equal(ref.decompress(v0, 0, -1, 0, None).reshape(a.shape, order="F"), a)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("order", [0, 5])
def test_gpu_decodes_v0(name, order):
    import crackle_b200 as cb
    g = load_golden(name)
    a = np.asfortranarray(g["input"])
    v0 = to_v0(g[f"ckl_order{order}"])
    h = cb.header(v0)
    assert h["format_version"] == 0 and (h["sx"], h["sy"], h["sz"]) == a.shape
    assert np.array_equal(cb.decompress(v0).reshape(a.shape, order="F"), a)
    assert np.array_equal(cb.decompress_range(v0, 1, 3), a[:, :, 1:3])
    lab = int(g["label"])
    assert np.array_equal(cb.decompress(v0, label=lab).reshape(a.shape, order="F").view(bool), g["mask"].view(bool))
