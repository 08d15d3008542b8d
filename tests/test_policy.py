"""CPU tests (no GPU needed: these entry points are host code inside the library) of the policy functions around the hot path
against the numpy restatement in oracle/policy_oracle.py and against properties the reference's functions have by construction."""
import numpy as np
import pytest

from bundletrack_b200 import policy, synth
from oracle import policy_oracle as po


def _poses(seed, n, max_deg=60.0):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        R = synth.so3_exp(ax * np.deg2rad(rng.uniform(0, max_deg)))
        out.append(synth.se3(R, rng.normal(0, 0.05, 3)).astype(np.float32))
    return np.stack(out)


def test_rotation_geodesic():
    P = _poses(0, 12)
    for a in range(0, 12, 3):
        for b in range(12):
            want = po.rotation_geodesic(P[a], P[b])
            if want > 0.05:            # acos is ill-conditioned near 0 (d angle = d trace / (2 sin angle)); identical poses are checked below
                assert abs(policy.rotation_geodesic(P[a], P[b]) - want) <= 1e-5
    assert policy.rotation_geodesic(P[3], P[3]) <= 1e-3            # acos near 1 is ill-conditioned in float, like the reference
    R = synth.so3_exp(np.array([0, 0, np.pi]))                     # 180 degrees: the clamp keeps acos defined
    assert abs(policy.rotation_geodesic(synth.se3(R, np.zeros(3)), np.eye(4)) - np.pi) <= 1e-3


def test_keyframe_check_rules():
    P = _poses(1, 6, max_deg=90)
    new = _poses(2, 1, max_deg=90)[0]
    for min_rot in (1.0, 10.0, 30.0, 80.0):
        assert policy.keyframe_check(new, 7, 500, P, 0, min_rot) == po.keyframe_check(new, 7, 500, P, 0, min_rot)
    assert policy.keyframe_check(new, 0, 0, P, 100, 180.0)         # frame 0 always
    assert not policy.keyframe_check(new, 7, 50, P, 100, 0.0)      # too few keypoints
    assert not policy.keyframe_check(P[2], 7, 500, P, 0, 10.0)     # identical to an existing keyframe
    assert policy.keyframe_check(new, 7, 500, P[:0], 0, 10.0)      # no keyframes yet


@pytest.mark.parametrize("seed,K,maxf", [(3, 8, 15), (4, 20, 15), (5, 40, 10), (6, 16, 16), (7, 16, 17), (8, 30, 2), (9, 5, 1)])
def test_select_keyframes_matches_oracle(seed, K, maxf):
    P = _poses(seed, K, max_deg=120)
    new = _poses(100 + seed, 1, max_deg=120)[0]
    got = policy.select_keyframes(new, P, maxf)
    want = po.select_keyframes(new, P, maxf)
    assert np.array_equal(got, want)
    assert len(got) == min(K, max(maxf - 1, 0)) or (K + 1 > maxf and len(got) == max(maxf - 1, 1))
    if K + 1 > maxf:
        assert got[0] == 0                                         # keyframe 0 anchors the model frame
    assert np.all(np.diff(got) > 0)


def test_rigid_transform():
    rng = np.random.default_rng(11)
    for n in (3, 5, 40, 500):
        a = rng.normal(0, 0.1, (n, 3)).astype(np.float32)
        T = synth.se3(synth.so3_exp(rng.normal(0, 1.0, 3)), rng.normal(0, 0.3, 3))
        b = (a @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
        got = policy.rigid_transform(a, b)
        assert np.abs(got - T).max() <= 2e-5
        noisy = b + rng.normal(0, 0.002, b.shape).astype(np.float32)
        assert np.abs(policy.rigid_transform(a, noisy) - po.rigid_transform(a, noisy)).max() <= 2e-5
    # a mirrored target has no proper rotation that fits: the closed form still returns a rotation (det +1), the best one
    a = rng.normal(0, 0.1, (30, 3)).astype(np.float32)
    m = a * np.array([1, 1, -1], np.float32)
    got = policy.rigid_transform(a, m)
    assert abs(np.linalg.det(got[:3, :3].astype(np.float64)) - 1) <= 1e-5
    assert np.abs(got - po.rigid_transform(a, m)).max() <= 1e-4
    # coplanar / collinear input is rank deficient but still yields an orthonormal frame or the identity, never NaN
    line = np.outer(np.linspace(0, 1, 10), [1, 2, 3]).astype(np.float32)
    out = policy.rigid_transform(line, line + np.float32(0.5))
    assert np.isfinite(out).all() and np.allclose(out[:3, :3].T @ out[:3, :3], np.eye(3), atol=1e-4)
    # exactly coplanar points (one coordinate identically 0): S has an exactly zero singular value; Eigen's JacobiSVD still returns an
    # orthogonal U and the reference recovers the rotation (Utils.cpp:192-207) - so must this closed form, every time
    for trial in range(200):
        flat = rng.normal(0, 0.1, (12, 3)).astype(np.float32); flat[:, 2] = 0
        T = synth.se3(synth.so3_exp(rng.normal(0, 1.0, 3)), rng.normal(0, 0.3, 3))
        got = policy.rigid_transform(flat, (flat @ T[:3, :3].T + T[:3, 3]).astype(np.float32))
        assert np.abs(got - T).max() <= 5e-5, trial
    from bundletrack_b200 import _lib
    with pytest.raises(_lib.BtError):
        policy.rigid_transform(a[:2], a[:2])


def test_lfnet_reply_parsing():
    rng = np.random.default_rng(21)
    n, dim = 137, 256
    kp = rng.uniform(0, 400, (n, 2)).astype(np.float32)
    desc = rng.normal(size=(n, dim)).astype(np.float32)
    parts = [np.array([n, dim], np.int32).tobytes(), kp.tobytes(), desc.tobytes()]
    for roi in ((100, 420, 60, 300), (0, 640, 0, 480), (311, 352, 200, 333)):
        got_k, got_d = policy.lfnet_parse_reply(parts, roi)
        assert np.array_equal(got_d, desc)
        assert np.abs(got_k - po.lfnet_keypoints_to_image(kp, roi)).max() <= 2e-4
        side = max(roi[1] - roi[0], roi[3] - roi[2])
        assert got_k[:, 0].min() >= roi[0] - 1e-3 and got_k[:, 0].max() <= roi[0] + side + 1e-3      # inside the padded square of the roi
    empty = [np.array([0, dim], np.int32).tobytes(), b"", b""]
    k0, d0 = policy.lfnet_parse_reply(empty, (0, 10, 0, 10))
    assert k0.shape == (0, 2) and d0.shape == (0, dim)
    from bundletrack_b200 import _lib
    with pytest.raises(_lib.BtError):
        policy.lfnet_parse_reply([parts[0], parts[1][:-4], parts[2]], (0, 640, 0, 480))             # truncated keypoint part
    with pytest.raises(_lib.BtError):
        policy.lfnet_parse_reply([parts[0], parts[1], parts[2] + b"0000"], (0, 640, 0, 480))         # descriptor part too long
    with pytest.raises(_lib.BtError):
        policy.lfnet_parse_reply(parts, (5, 5, 0, 10))                                              # empty roi


def test_map_point_tracks_match_the_restatement():
    """A random stream of frame pairs (inlier matches with repeated keypoints, overwritten observations, forgotten frames) through the
    library's map-point bookkeeping and through the Python restatement: same propagated matches in the same order, same statistics."""
    rng = np.random.default_rng(33)
    n_frames = 9
    kps = [np.round(rng.uniform(0, 640, (60, 2)), 1).astype(np.float32) for _ in range(n_frames)]     # keypoints recur across pairs
    got, want = policy.Tracks(), po.Tracks()
    for step in range(60):
        a = int(rng.integers(1, n_frames)); b = int(rng.integers(0, a))
        n = int(rng.integers(0, 25))
        ia, ib = rng.integers(0, 60, n), rng.integers(0, 60, n)
        uv = np.concatenate([kps[a][ia], kps[b][ib]], 1).astype(np.float32)
        ex_n = int(rng.integers(0, 10))
        existing = np.concatenate([kps[a][rng.integers(0, 60, ex_n)], kps[b][rng.integers(0, 60, ex_n)]], 1).astype(np.float32)
        pg, pw = got.propagate(a, b, existing), want.propagate(a, b, existing)
        assert np.array_equal(pg, pw), (step, a, b)
        inl = rng.random(n) < 0.7
        got.update_pair(a, b, uv, inl); want.update_pair(a, b, uv, inl)
        assert got.stats() == want.stats()
        if step % 17 == 16:
            f = int(rng.integers(0, n_frames))
            got.forget_frame(f); want.forget_frame(f)
            assert got.stats() == want.stats()
    assert got.stats()[0] > 50
    # the documented rules, one by one
    t = policy.Tracks()
    t.update_pair(1, 0, [[10, 10, 20, 20]])
    assert t.stats() == (1, 2)
    t.update_pair(2, 1, [[30, 30, 10, 10]])                      # frame 1's keypoint already belongs to the point: frame 2 joins it
    assert t.stats() == (1, 3)
    assert np.array_equal(t.propagate(2, 0, np.zeros((0, 4))), [[30, 30, 20, 20]])      # seen in 2 and 0 through the shared point
    assert len(t.propagate(2, 0, [[30, 30, 5, 5]])) == 0          # (uA, vA) already matched
    assert len(t.propagate(2, 0, [[7, 7, 20, 20]])) == 0          # (uB, vB) already matched
    t.forget_frame(0)
    assert len(t.propagate(2, 0, np.zeros((0, 4)))) == 0 and t.stats() == (1, 2)
    from bundletrack_b200 import _lib
    with pytest.raises(_lib.BtError):
        t.update_pair(0, 1, [[1, 1, 2, 2]])                       # A must be the newer frame
    assert np.array_equal(t.propagate(2, 1, np.zeros((0, 4))), [[30, 30, 10, 10]])
    with pytest.raises(_lib.BtError):
        t.propagate(2, 1, np.zeros((0, 4)), capacity=0)           # more matches than the caller made room for
    t.close(); got.close()


def test_ba_gate_and_pose_record(tmp_path):
    """Bundler::optimizeGPU's NO_BA gate (Bundler.cpp:343) and saveNewframeResult's pose file (Bundler.cpp:362-378): ob_in_cam =
    inverse(cur_in_model), 10 significant digits, columns right-aligned to the widest coefficient like Eigen's operator<<."""
    from bundletrack_b200 import synth
    assert not policy.ba_gate(10, 10) and policy.ba_gate(11, 10) and not policy.ba_gate(0, 0) and policy.ba_gate(1, 0)
    rng = np.random.default_rng(3)
    for _ in range(20):
        T = synth.se3(synth.so3_exp(rng.normal(0, 0.8, 3)), rng.normal(0, 0.5, 3)).astype(np.float32)
        txt = policy.pose_text(T)
        lines = txt.split("\n")
        assert len(lines) == 5 and lines[4] == "" and len({len(l) for l in lines[:4]}) == 1      # four rows of equal width, trailing newline
        got = np.array([[float(v) for v in l.split()] for l in lines[:4]])
        assert got.shape == (4, 4)
        want = np.linalg.inv(T.astype(np.float64))
        assert np.abs(got - want).max() <= 2e-6
        # every coefficient carries the 10 significant digits of the float it was printed from
        inv32 = np.array([[np.float32(v) for v in l.split()] for l in lines[:4]], np.float32)
        assert np.array_equal(inv32, got.astype(np.float32))
        p = tmp_path / "0001.txt"
        policy.save_pose_txt(str(p), T)
        assert p.read_text() == txt
    with pytest.raises(Exception):
        policy.pose_text(np.zeros((4, 4), np.float32))      # singular
