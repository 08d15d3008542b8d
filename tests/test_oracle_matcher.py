"""kNN oracle (exact brute force) vs OpenCV's CPU BFMatcher — the reference's own NO_OPENCV_CUDA path
(/root/reference/src/FeatureManager.cpp:266-269) — and its edge cases."""
import numpy as np
import pytest

from bundletrack_b200 import synth
from oracle import matcher_oracle as mo


def test_knn_matches_opencv_bfmatcher():
    cv2 = pytest.importorskip("cv2")
    a, b, _, _ = synth.make_descriptors(1, 700, 900)
    i1, d1 = mo.knn(a, b)
    i2, d2 = mo.knn_cv2(a, b)
    assert np.array_equal(i1, i2)
    assert np.abs(d1 - d2).max() < 1e-6


def test_knn_planted_matches_are_nearest():
    a, b, ia, ib = synth.make_descriptors(2, 400, 500, match_frac=0.5, noise=0.02)
    idx, dist = mo.knn(b[ib], a, k=1)
    assert (idx[:, 0] == ia).mean() > 0.99
    assert (np.diff(mo.knn(a, b)[1], axis=1) >= 0).all()       # ascending


def test_knn_ties_go_to_lower_index_and_padding():
    a = np.zeros((3, 256), np.float32); a[:, 0] = 1
    b = np.zeros((4, 256), np.float32); b[:, 0] = [1, 1, 0.5, 1]   # rows 0,1,3 identical
    idx, dist = mo.knn(a, b, k=5)
    assert idx[0].tolist() == [0, 1, 3, 2, -1]
    assert np.isinf(dist[0, 4]) and dist[0, 0] == 0
    e_i, e_d = mo.knn(a[:0], b)
    assert e_i.shape == (0, 5)
    z_i, z_d = mo.knn(a, b[:0])
    assert (z_i == -1).all() and np.isinf(z_d).all()
