// tracks.cu — map-point bookkeeping between the matcher and RANSAC (the part that is not a kernel): host code,
// no CUDA calls.  Restated from
//   SiftManager::updateFramePairMapPoints   /root/reference/src/FeatureManager.cpp:448-485
//   SiftManager::findCorresByMapPoints      /root/reference/src/FeatureManager.cpp:489-520
//   SiftManager::forgetFrame (map points)   /root/reference/src/FeatureManager.cpp:163-169
// A map point is the set of image observations {frame -> (u, v)} of one surface point; every frame keeps an ordered map
// (u, v) -> map point.  The reference keys MapPoint::_img_pt by shared_ptr<Frame>; frames are identified by their id here, which
// changes no result (the pointer-ordered map is only searched, never walked).  Propagated matches come out in (uA, vA) order, the
// order std::map<std::pair<float,float>, ...> walks them in the reference.
#include <map>
#include <utility>
#include <vector>
#include "bt_common.cuh"

struct bt_tracks {
	struct MapPoint { std::map<int, std::pair<float, float>> img_pt; };
	std::vector<MapPoint> points;                                              // _map_points_global
	std::map<int, std::map<std::pair<float, float>, int>> frame_map;           // Frame::_map_points, by frame id
};

extern "C" int bt_tracks_create(bt_tracks** out) {
	BT_REQUIRE(out, BT_ERR_INVALID_ARG, "bt_tracks_create: out is NULL");
	*out = new bt_tracks();
	return BT_OK;
}
extern "C" void bt_tracks_destroy(bt_tracks* t) { delete t; }

extern "C" int bt_tracks_update_pair(bt_tracks* t, int frame_a, int frame_b, const float* uv, const unsigned char* is_inlier, int n) {
	BT_REQUIRE(t && (n == 0 || uv) && n >= 0, BT_ERR_INVALID_ARG, "bt_tracks_update_pair: bad arguments");
	BT_REQUIRE(frame_a > frame_b, BT_ERR_INVALID_ARG, "bt_tracks_update_pair: frame A must be the newer frame (id %d <= %d)", frame_a, frame_b);
	auto& mapA = t->frame_map[frame_a];
	auto& mapB = t->frame_map[frame_b];
	for (int i = 0; i < n; i++) {
		if (is_inlier && !is_inlier[i]) continue;
		const std::pair<float, float> kA(uv[4 * i], uv[4 * i + 1]), kB(uv[4 * i + 2], uv[4 * i + 3]);
		const auto itB = mapB.find(kB);
		if (mapA.find(kA) != mapA.end() && itB != mapB.end()) continue;       // both ends already belong to map points
		int id;
		if (itB == mapB.end()) {                                               // new map point anchored at frame B
			id = (int)t->points.size();
			t->points.emplace_back();
			mapB[kB] = id;
			t->points[id].img_pt[frame_b] = kB;
		} else {
			id = itB->second;
		}
		t->points[id].img_pt[frame_a] = kA;
		mapA[kA] = id;
	}
	return BT_OK;
}

extern "C" int bt_tracks_propagate(bt_tracks* t, int frame_a, int frame_b, const float* existing_uv, int n_existing, float* out_uv, int capacity, int* n_out) {
	BT_REQUIRE(t && n_out && (n_existing == 0 || existing_uv) && n_existing >= 0 && capacity >= 0 && (capacity == 0 || out_uv), BT_ERR_INVALID_ARG,
	           "bt_tracks_propagate: bad arguments");
	BT_REQUIRE(frame_a > frame_b, BT_ERR_INVALID_ARG, "bt_tracks_propagate: frame A must be the newer frame (id %d <= %d)", frame_a, frame_b);
	*n_out = 0;
	const auto fa = t->frame_map.find(frame_a);
	if (fa == t->frame_map.end()) return BT_OK;
	int n = 0;
	for (const auto& h : fa->second) {                                         // (uA, vA) order
		const bt_tracks::MapPoint& mp = t->points[h.second];
		const auto ob = mp.img_pt.find(frame_b);
		if (ob == mp.img_pt.end()) continue;
		const float uA = h.first.first, vA = h.first.second, uB = ob->second.first, vB = ob->second.second;
		bool existed = false;                                                  // against the caller's matches AND the ones appended so far
		for (int i = 0; i < n_existing && !existed; i++)
			existed = (existing_uv[4 * i] == uA && existing_uv[4 * i + 1] == vA) || (existing_uv[4 * i + 2] == uB && existing_uv[4 * i + 3] == vB);
		for (int i = 0; i < n && !existed; i++)
			existed = (out_uv[4 * i] == uA && out_uv[4 * i + 1] == vA) || (out_uv[4 * i + 2] == uB && out_uv[4 * i + 3] == vB);
		if (existed) continue;
		BT_REQUIRE(n < capacity, BT_ERR_CAPACITY, "bt_tracks_propagate: more than %d propagated matches", capacity);
		out_uv[4 * n] = uA; out_uv[4 * n + 1] = vA; out_uv[4 * n + 2] = uB; out_uv[4 * n + 3] = vB;
		n++;
	}
	*n_out = n;
	return BT_OK;
}

extern "C" int bt_tracks_forget_frame(bt_tracks* t, int frame) {
	BT_REQUIRE(t, BT_ERR_INVALID_ARG, "bt_tracks_forget_frame: NULL");
	for (auto& mp : t->points) mp.img_pt.erase(frame);
	t->frame_map.erase(frame);          // the reference drops the Frame object (and its _map_points) with the frame
	return BT_OK;
}

extern "C" int bt_tracks_stats(const bt_tracks* t, int* n_points, int* n_observations) {
	BT_REQUIRE(t && n_points && n_observations, BT_ERR_INVALID_ARG, "bt_tracks_stats: NULL argument");
	*n_points = (int)t->points.size();
	int obs = 0;
	for (const auto& mp : t->points) obs += (int)mp.img_pt.size();
	*n_observations = obs;
	return BT_OK;
}
