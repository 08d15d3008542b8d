"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's host-side policy around the hot path.  Only tests/ may import this.

  rotation_geodesic   Utils::rotationGeodesicDistance          /root/reference/src/Utils.cpp:42-47
  keyframe_check      Bundler::checkAndAddKeyframe             /root/reference/src/Bundler.cpp:185-221
  select_keyframes    Bundler::selectKeyFramesForBA            /root/reference/src/Bundler.cpp:224-274 ("greedy_rot")
  rigid_transform     Utils::solveRigidTransformBetweenPoints  /root/reference/src/Utils.cpp:180-214
The reference has no tests or golden vectors for these and its selection walks a std::set ordered by POINTER value, so its tie
breaks are not reproducible: parity unpinned by the reference; the order used here (new frame first, then keyframes by index; ties
to the lower index) is the documented contract of the C entry points."""
import numpy as np


def rotation_geodesic(A, B):
    R1, R2 = np.asarray(A, np.float32)[:3, :3], np.asarray(B, np.float32)[:3, :3]
    t = np.float32((np.trace(R1 @ R2.T) - 1) / 2.0)
    return float(np.arccos(np.clip(t, -1.0, 1.0)))


def keyframe_check(pose_new, frame_id, n_keypts, keyframe_poses, min_feat_num=0, min_rot_deg=10.0):
    if frame_id == 0:
        return True
    if n_keypts < min_feat_num:
        return False
    return all(np.degrees(rotation_geodesic(pose_new, k)) >= min_rot_deg for k in keyframe_poses)


def select_keyframes(pose_new, keyframe_poses, max_BA_frames=15):
    K = len(keyframe_poses)
    if K + 1 <= max_BA_frames:
        return np.arange(K, dtype=np.int32)
    chosen = [0]
    while len(chosen) + 1 < max_BA_frames:
        best, best_i = np.float32(np.finfo(np.float32).max), -1
        for i in range(K):
            if i in chosen:
                continue
            cum = np.float32(rotation_geodesic(keyframe_poses[i], pose_new))
            for j in sorted(chosen):
                cum = np.float32(cum + np.float32(rotation_geodesic(keyframe_poses[i], keyframe_poses[j])))
            if cum < best:
                best, best_i = cum, i
        chosen.append(best_i)
    return np.array(sorted(chosen), np.int32)


def rigid_transform(p1, p2):
    p1, p2 = np.asarray(p1, np.float64), np.asarray(p2, np.float64)
    m1, m2 = p1.mean(0), p2.mean(0)
    S = (p1 - m1).T @ (p2 - m2)
    U, _, Vt = np.linalg.svd(S)
    V = Vt.T
    R = V @ U.T
    T = np.eye(4)
    if not np.allclose(R.T @ R, np.eye(3), atol=1e-5):
        return T.astype(np.float32)
    if np.linalg.det(R) < 0:
        V[:, 2] = -V[:, 2]
        R = V @ U.T
    T[:3, :3] = R
    T[:3, 3] = m2 - R @ m1
    return T.astype(np.float32) if np.isfinite(T).all() else np.eye(4, dtype=np.float32)


def lfnet_keypoints_to_image(kpts, roi):
    """Lfnet::detectFeature (FeatureManager.cpp:811-907, rot_deg = 0): forward = scale(400/side) @ translate(-umin, -vmin); keypoints
    come back through its inverse."""
    umin, umax, vmin, vmax = roi
    side = max(vmax - vmin, umax - umin)
    T = np.eye(3, dtype=np.float32); T[0, 2] = -umin; T[1, 2] = -vmin
    S = np.eye(3, dtype=np.float32); S[0, 0] = np.float32(400) / np.float32(side); S[1, 1] = np.float32(400) / np.float32(side)
    back = np.linalg.inv((S @ T).astype(np.float64))
    p = np.concatenate([np.asarray(kpts, np.float64), np.ones((len(kpts), 1))], 1) @ back.T
    return p[:, :2].astype(np.float32)


class Tracks:
    """SiftManager::updateFramePairMapPoints / findCorresByMapPoints / forgetFrame (FeatureManager.cpp:448-520,163-169) with Python
    containers: `frame_map[frame]` is the std::map<(u,v), MapPoint> of the frame (walked in sorted key order), a map point is a dict
    frame -> (u, v)."""

    def __init__(self):
        self.points = []
        self.frame_map = {}

    def update_pair(self, fa, fb, uv, is_inlier=None):
        A, B = self.frame_map.setdefault(fa, {}), self.frame_map.setdefault(fb, {})
        for i, (uA, vA, uB, vB) in enumerate(np.asarray(uv, np.float32).reshape(-1, 4)):
            if is_inlier is not None and not is_inlier[i]:
                continue
            kA, kB = (float(uA), float(vA)), (float(uB), float(vB))
            if kA in A and kB in B:
                continue
            if kB not in B:
                mp = {fb: kB}
                self.points.append(mp)
                B[kB] = mp
            else:
                mp = B[kB]
            mp[fa] = kA
            A[kA] = mp

    def propagate(self, fa, fb, existing_uv):
        matches = [tuple(float(x) for x in r) for r in np.asarray(existing_uv, np.float32).reshape(-1, 4)]
        n0 = len(matches)
        for kA in sorted(self.frame_map.get(fa, {})):
            mp = self.frame_map[fa][kA]
            if fb not in mp:
                continue
            kB = mp[fb]
            if any((m[0], m[1]) == kA or (m[2], m[3]) == kB for m in matches):
                continue
            matches.append((kA[0], kA[1], kB[0], kB[1]))
        return np.array(matches[n0:], np.float32).reshape(-1, 4)

    def forget_frame(self, f):
        for mp in self.points:
            mp.pop(f, None)
        self.frame_map.pop(f, None)

    def stats(self):
        return len(self.points), sum(len(mp) for mp in self.points)
