"""The host KD-tree builder (csrc/kdtree_host.cpp) must reproduce scikit-learn's
KDTree(leaf_size=1) arrays bit for bit -- idx_array is what fixes the kNN tie order."""
import ctypes as C

import numpy as np
import pytest
from sklearn.neighbors import KDTree

from detecting_cones_aoslo_b200 import _lib


def build_host(points):
    pts = np.ascontiguousarray(points, dtype=np.float64)
    n = pts.shape[0]
    nn = _lib.lib.aoslo_kdtree_num_nodes(n)
    idx = np.zeros(n, np.int32)
    ns = np.zeros(nn, np.int32)
    ne = np.zeros(nn, np.int32)
    nb = np.zeros((nn, 4), np.float64)
    st = _lib.lib.aoslo_kdtree_build_host(pts.ctypes.data, n, idx.ctypes.data, ns.ctypes.data, ne.ctypes.data,
                                          nb.ctypes.data, nn)
    assert st == 0
    return idx, ns, ne, nb


def lattice_points(rng, n, span):
    # dense integer lattice with duplicates removed: exact coordinate ties everywhere
    p = rng.integers(0, span, size=(n * 2, 2))
    p = np.unique(p, axis=0)
    rng.shuffle(p)
    return p[:n].astype(np.float64)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 16, 17, 31, 33, 100, 283, 1000, 5521, 8145, 31876])
def test_matches_sklearn_arrays(n):
    rng = np.random.default_rng(n)
    pts = lattice_points(rng, n, max(4, int(np.sqrt(n) * 2.5)))
    n = pts.shape[0]
    tree = KDTree(pts, leaf_size=1)
    data, idx_array, node_data, node_bounds = tree.get_arrays()
    idx, ns, ne, nb = build_host(pts)
    assert _lib.lib.aoslo_kdtree_num_nodes(n) == node_data.shape[0]
    np.testing.assert_array_equal(idx, idx_array)
    np.testing.assert_array_equal(ns, node_data['idx_start'])
    np.testing.assert_array_equal(ne, node_data['idx_end'])
    np.testing.assert_array_equal(nb[:, :2], node_bounds[0])
    np.testing.assert_array_equal(nb[:, 2:], node_bounds[1])


def test_matches_sklearn_continuous_and_duplicates():
    rng = np.random.default_rng(7)
    pts = rng.random((4001, 2)) * 100
    pts[100:120] = pts[0:20]            # coincident points
    tree = KDTree(pts, leaf_size=1)
    _, idx_array, node_data, node_bounds = tree.get_arrays()
    idx, ns, ne, nb = build_host(pts)
    np.testing.assert_array_equal(idx, idx_array)
    np.testing.assert_array_equal(nb[:, :2], node_bounds[0])
