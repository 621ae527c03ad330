"""SURVEY 8(f)1, the last consumer of for_each_z_parallel: voxel_connectivity_graph (src/operations.hpp:667-826, binding
src/fastcrackle.cpp:538-565, Python crackle/operations.py:936-954) on the GPU -- crack decode -> planes -> one byte per voxel
(connectivity 6: + CCL, label keys and the z comparison) -- against the compiled reference and against the graph computed
from the voxels with numpy (the reference's own test compares with cc3d.voxel_connectivity_graph, automated_test.py:976-989)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def vcg_numpy(v, connectivity, permissible, z0=0, z1=None):
    """bits 00 -z +z -y +y -x +x, set = the neighbour holds the same label.  Edges on the image border keep the fill value of
    the reference's decoder (crackcodes.hpp:706-862): open for IMPERMISSIBLE streams, closed for PERMISSIBLE ones; with
    connectivity 6 (and more than one slice in the stream) the outer z faces of the decoded range are open."""
    sz = v.shape[2]
    s = v[:, :, z0:(sz if z1 is None else z1)]
    b = 0 if permissible else 1
    px = np.full(s.shape, b, np.uint8); px[:-1] = s[:-1] == s[1:]
    mx = np.full(s.shape, b, np.uint8); mx[1:] = s[1:] == s[:-1]
    py = np.full(s.shape, b, np.uint8); py[:, :-1] = s[:, :-1] == s[:, 1:]
    my = np.full(s.shape, b, np.uint8); my[:, 1:] = s[:, 1:] == s[:, :-1]
    out = px | (mx << 1) | (py << 2) | (my << 3)
    if connectivity == 6 and sz > 1:
        pz = np.ones(s.shape, np.uint8); pz[:, :, :-1] = s[:, :, :-1] == s[:, :, 1:]
        mz = np.ones(s.shape, np.uint8); mz[:, :, 1:] = s[:, :, 1:] == s[:, :, :-1]
        out = out | (pz << 4) | (mz << 5)
    return np.asfortranarray(out.astype(np.uint8))


def _volumes():
    from crackle_b200 import synth
    rng = np.random.default_rng(3)
    return [synth.jittered_voronoi((97, 130, 9), 13, np.uint32, seed=4, id_bits=20),       # sx % 4 != 0: byte stores
            synth.jittered_voronoi((256, 64, 5), 20, np.uint64, seed=1, id_bits=40),
            synth.random_blobs((61, 47, 6), 9, np.uint16, seed=3),
            synth.random_blobs((64, 33, 4), 40, np.uint16, seed=5),                        # word-aligned rows, many labels
            np.asfortranarray(rng.integers(0, 3, (40, 37, 5)).astype(np.uint8)),          # noise: PERMISSIBLE crack format
            np.asfortranarray(rng.integers(0, 2, (68, 20, 3)).astype(np.uint8)),          # PERMISSIBLE, sx % 4 == 0
            np.asfortranarray(np.full((16, 12, 3), 7, dtype=np.uint32)),                  # a single label
            synth.random_blobs((33, 29, 1), 5, np.uint8, seed=2),                         # sz == 1: connectivity 6 adds nothing
            np.ascontiguousarray(synth.random_blobs((33, 45, 7), 8, np.uint32, seed=1))]  # C-order stream: the graph stays x-fastest


@pytest.mark.parametrize("order", [0, 3])
def test_vcg_matches_reference_and_voxels(order):
    import crackle_b200 as cb
    from oracle import oracle as O
    ref = O.ref_module()
    ctx = cb.Context(0)
    for v in _volumes():
        b = O.compress(v, order)
        perm = O.header(b)["crack_format"] == 1
        sz = v.shape[2]
        for conn in (4, 6):
            got = ctx.voxel_connectivity_graph(b, 0, -1, conn)
            assert got.dtype == np.uint8 and got.shape == v.shape and got.flags.f_contiguous
            assert np.array_equal(got, vcg_numpy(v, conn, perm)), (v.shape, order, conn)
            if ref is not None:
                assert np.array_equal(got, np.asarray(ref.voxel_connectivity_graph(b, 0, -1, 1, conn))), (v.shape, order, conn)
            assert np.array_equal(cb.voxel_connectivity_graph(b, connectivity=conn), got)          # crackle.voxel_connectivity_graph
        if sz > 2:                                   # z-ranges: (fastcrackle.cpp:540) the 2-D bits of a sub-range ...
            got = ctx.voxel_connectivity_graph(b, 1, sz - 1, 4)
            assert np.array_equal(got, vcg_numpy(v, 4, perm, 1, sz - 1))
            if ref is not None:
                assert np.array_equal(got, np.asarray(ref.voxel_connectivity_graph(b, 1, sz - 1, 1, 4)))
            # ... and with connectivity 6 the outer faces of the RANGE are open (the reference indexes its last slice with the
            # stream's sz there -- a write past its buffer for sub-ranges, unreachable from its Python interface)
            assert np.array_equal(ctx.voxel_connectivity_graph(b, 1, sz - 1, 6), vcg_numpy(v, 6, perm, 1, sz - 1))
    ctx.close()


def test_vcg_device_stream_and_errors():
    import torch
    import crackle_b200 as cb
    from crackle_b200 import _capi, synth
    from oracle import oracle as O
    v = synth.jittered_voronoi((128, 96, 6), 16, np.uint64, seed=9, id_bits=40)
    b = O.compress(v, 0)
    ctx = cb.Context(0)
    ctx.bind_torch_stream()
    dstream = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
    dout = torch.empty(v.size, dtype=torch.uint8, device="cuda")
    rc = _capi.lib().ckl_voxel_connectivity_graph(ctx._h, dstream.data_ptr(), 1, dstream.numel(), 0, -1, 6, dout.data_ptr(), 1, dout.numel())
    assert rc == 0
    torch.cuda.synchronize()
    assert np.array_equal(dout.cpu().numpy().reshape(v.shape, order="F"), vcg_numpy(v, 6, False))
    with pytest.raises(ValueError, match="only connectivity 4 and 6 are currently supported"):
        ctx.voxel_connectivity_graph(b, 0, -1, 8)
    with pytest.raises(ValueError, match="Only 4 and 6 connected are supported"):
        cb.voxel_connectivity_graph(b, connectivity=26)
    with pytest.raises(RuntimeError, match="Invalid range"):
        ctx.voxel_connectivity_graph(b, 4, 2, 4)
    rc = _capi.lib().ckl_voxel_connectivity_graph(ctx._h, dstream.data_ptr(), 1, dstream.numel(), 0, -1, 4, dout.data_ptr(), 1, 10)
    assert rc != 0 and "output buffer too small" in _capi.lib().ckl_ctx_error(ctx._h).decode()
    ctx.close()
