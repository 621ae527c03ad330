"""Compressed-domain statistics (voxel_counts / centroids / bounding_boxes, the consumers of
for_each_z_parallel, src/operations.hpp:89-182, :321-665) computed on the GPU from the runs, checked against the compiled
reference (oracle/_ref) and against the same statistics taken from the decoded voxels with numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _volumes():
    from crackle_b200 import synth
    rng = np.random.default_rng(3)
    return [synth.jittered_voronoi((97, 130, 9), 13, np.uint32, seed=4, id_bits=20),
            synth.jittered_voronoi((256, 192, 12), 20, np.uint64, seed=1, id_bits=40),
            synth.random_blobs((61, 47, 6), 9, np.uint16, seed=3),
            np.asfortranarray(rng.integers(0, 50, (40, 37, 5)).astype(np.uint8)),          # noise: PERMISSIBLE crack format
            np.asfortranarray(np.full((16, 12, 3), 7, dtype=np.uint32))]                    # a single label


def _numpy_stats(v, z0, z1):
    sub = v[:, :, z0:z1]
    X, Y, Z = np.meshgrid(np.arange(v.shape[0]), np.arange(v.shape[1]), np.arange(z0, z1), indexing="ij")
    out = {}
    for lab in np.unique(sub):
        m = sub == lab
        n = int(m.sum())
        out[int(lab)] = (n, [X[m].sum() / n, Y[m].sum() / n, Z[m].sum() / n],
                         [X[m].min(), Y[m].min(), Z[m].min(), X[m].max(), Y[m].max(), Z[m].max()])
    return out


@pytest.mark.parametrize("order", [0, 3])
def test_label_stats_match_reference_and_voxels(order):
    import crackle_b200 as cb
    from oracle import oracle as O
    ref = O.ref_module()
    ctx = cb.default_context()
    for v in _volumes():
        b = O.compress(v, order)
        sz = v.shape[2]
        for z0, z1 in ((0, -1), (1, sz - 1), (sz - 1, sz)):
            lab, cnt, sums, bbox = ctx.label_stats(b, z0, z1)
            a0, a1 = max(min(z0, sz - 1), 0), (sz if z1 < 0 else min(z1, sz))
            want = _numpy_stats(v, a0, a1)
            got = {int(k): (int(c), (s / max(int(c), 1)).tolist(), bb.tolist()) for k, c, s, bb in zip(lab, cnt, sums.astype(np.float64), bbox)}
            assert np.array_equal(lab, np.unique(v).astype(np.uint64))
            for k, (n, cen, bb) in want.items():
                assert got[k][0] == n and got[k][1] == cen and got[k][2] == bb, (v.shape, z0, z1, k)
            for k in set(got) - set(want):               # labels of the stream that do not occur in this z-range
                assert got[k][0] == 0 and got[k][2] == [0xFFFFFFFF] * 3 + [0] * 3
            if ref is not None and len(want) > 0:
                rv = ref.voxel_counts(b, z0, z1, 1)
                assert {int(k): int(c) for k, c in zip(lab, cnt) if c} == {int(k): int(c) for k, c in rv.items()}
                rc = ref.centroids(b, z0, z1, 1)
                assert {k: got[k][1] for k in want} == {int(k): list(c) for k, c in rc.items()}       # float64, bit-exact
                rb = ref.bounding_boxes(b, z0, z1, 1)
                assert {k: got[k][2] for k in got} == {int(k): [int(x) for x in c] for k, c in rb.items()}


def test_python_mirror_of_the_reference_functions():
    # crackle.voxel_counts / centroids / bounding_boxes (codec.py:949-1065): same return shapes, label= shortcut, errors
    import crackle_b200 as cb
    v = np.zeros((40, 30, 12), dtype=np.uint16, order="F")
    v[5:20, 5:20, 3:7] = 7
    v[22:30, 2:9, 6:11] = 9
    b = cb.compress(v)
    assert cb.num_labels(b) == 3 and cb.labels(b).tolist() == [0, 7, 9] and cb.contains(b, 9) and not cb.contains(b, 8)
    vc = cb.voxel_counts(b)
    assert vc == {0: int((v == 0).sum()), 7: 15 * 15 * 4, 9: 8 * 7 * 5}
    assert cb.voxel_counts(b, label=7) == 15 * 15 * 4
    assert cb.centroids(b, label=9) == [25.5, 5.0, 8.0]
    bb = cb.bounding_boxes(b)
    assert bb[7] == (slice(5, 20), slice(5, 20), slice(3, 7)) and bb[9] == (slice(22, 30), slice(2, 9), slice(6, 11))
    assert cb.bounding_boxes(b, label=9, no_slice_conversion=True).tolist() == [22, 2, 6, 29, 8, 10]
    with pytest.raises(ValueError, match="not contained"):
        cb.voxel_counts(b, label=8)
    one = cb.compress(np.full((8, 9, 4), 5, dtype=np.uint8, order="F"))
    assert cb.voxel_counts(one) == {5: 8 * 9 * 4}
    assert cb.bounding_boxes(one, no_slice_conversion=True)[5].tolist() == [0, 0, 0, 8, 9, 4]


def test_label_stats_at_bench_slab_size():
    # 1024x1024 uint64 slab of the bench volume: totals and a device-resident stream
    import torch
    import crackle_b200 as cb
    from crackle_b200 import synth
    t = synth.jittered_voronoi_torch((1024, 1024, 12), 24, np.uint64, seed=0, id_bits=40, sz_total=1024)
    ctx = cb.Context(0)
    b = ctx.compress(t, 0)
    lab, cnt, sums, bbox = ctx.label_stats(b)
    assert int(cnt.sum()) == t.numel()
    u, c = torch.unique(t.view(torch.int64), return_counts=True)
    assert np.array_equal(lab.view(np.int64), u.cpu().numpy()) and np.array_equal(cnt.view(np.int64), c.cpu().numpy())
    assert int(sums[:, 2].sum()) == 1024 * 1024 * sum(range(12))
    assert bbox[:, 3].max() == 1023 and bbox[:, 4].max() == 1023 and bbox[:, 5].max() == 11 and bbox[:, :3].min() == 0
