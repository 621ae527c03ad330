"""SURVEY 8(f)4 / INTEGRATION.md Option B, exercised: the reference's OWN test file (automated_test.py) runs through the
reference's own Python package (`crackle`), with `fastcrackle` replaced by tests/shims/fastcrackle.py -- compress /
decompress / voxel_counts / centroids / bounding_boxes of flat-label streams on the B200 library, everything else on the
reference's compiled module.  The package and the test file are staged by oracle/build_ref.sh into oracle/_ref/refpkg (the
checker's directory: git-ignored, travels to the GPU box with the compiled reference); nothing is read from
/root/reference at run time.  Tests that need the two .cpso.gz fixtures (compresso decoder absent) skip themselves; three
tests that import the third-party cc3d package are deselected."""
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, "oracle", "_ref", "refpkg")
DESELECT = ["test_connected_components", "test_voxel_connectivity_graph", "test_contacts"]      # import cc3d (not installed)


def _run(backend, extra=()):
    env = dict(os.environ, CKL_SHIM_BACKEND=backend,
               PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests", "shims"), PKG, os.path.join(ROOT, "oracle", "_ref"), ROOT]))
    cmd = [sys.executable, "-m", "pytest", os.path.join(PKG, "automated_test.py"), "-q", "-p", "no:cacheprovider", "--no-header",
           "-c", os.devnull, "--rootdir", PKG]
    cmd += ["-k", "not (" + " or ".join(DESELECT) + ")"]
    return subprocess.run(cmd + list(extra), capture_output=True, text=True, timeout=1800, env=env, cwd=PKG)


def _need_pkg():
    if not os.path.exists(os.path.join(PKG, "automated_test.py")):
        pytest.skip("oracle/_ref/refpkg is not staged on this box (oracle/build_ref.sh needs /root/reference)")


def test_harness_runs_the_reference_suite_on_the_reference():
    # CPU: the shims (fastremap, google_crc32c, compresso) and the staged package are sound -- all on the reference backend
    _need_pkg()
    r = _run("ref")
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:]
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert int(re.search(r"(\d+) passed", tail).group(1)) >= 170, tail


@pytest.mark.gpu
def test_reference_suite_passes_through_the_b200_drop_in():
    _need_pkg()
    r = _run("b200")
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:]
    assert r.returncode == 0, r.stdout[-6000:] + r.stderr[-2000:]
    assert int(re.search(r"(\d+) passed", tail).group(1)) >= 170, tail


@pytest.mark.gpu
def test_drop_in_module_serves_the_hot_path_and_exports_every_reference_name():
    # the shim really routes to the GPU library, and exports all 16 names + CppPin of src/fastcrackle.cpp:641-669
    _need_pkg()
    code = (
        "import numpy as np, crackle, fastcrackle\n"
        "names = ['decompress','compress','reencode_markov','remap','index_range','connected_components','compute_pins','point_cloud',"
        "'voxel_counts','centroids','bounding_boxes','get_slice_vcg','voxel_connectivity_graph','contacts','array_equal',"
        "'mode_pooling_2x2x1','CppPin']\n"
        "assert all(hasattr(fastcrackle, n) for n in names)\n"
        "assert 'crackle_b200' in fastcrackle.__backend__\n"
        "v = np.random.default_rng(0).integers(0, 9, (64, 48, 5)).astype(np.uint32)\n"
        "b = crackle.compress(v)\n"
        "assert np.array_equal(crackle.decompress(b), v)\n"
        "arr = crackle.CrackleArray(b)\n"
        "assert np.array_equal(arr[:, :, 2], v[:, :, 2]) and np.array_equal(arr[3:9, 1:7, 1:4], v[3:9, 1:7, 1:4])\n"
        "assert crackle.voxel_counts(b) == {int(k): int(c) for k, c in zip(*np.unique(v, return_counts=True))}\n"
        "bp = crackle.compress(v, allow_pins=True)\n"
        "assert np.array_equal(crackle.decompress(bp), v)\n"
        "assert fastcrackle.CALLS['b200'] >= 4 and fastcrackle.CALLS['ref'] >= 1, fastcrackle.CALLS\n"
        "print('DROP-IN OK', fastcrackle.CALLS)\n")
    env = dict(os.environ, CKL_SHIM_BACKEND="b200",
               PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests", "shims"), PKG, os.path.join(ROOT, "oracle", "_ref"), ROOT]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "DROP-IN OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
