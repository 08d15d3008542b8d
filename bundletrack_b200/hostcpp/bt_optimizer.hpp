// bt_optimizer.hpp — C++ host shim with the REFERENCE'S signatures on top of the C-ABI (include/bundletrack_b200.h).
//
// Drop-in for what BundleTrack's host code talks to on the hot path:
//   * OptimizerGpu::optimizeFrames      (/root/reference/src/cuda/LossGPU.h:50, called from Bundler::optimizeGPU,
//                                        /root/reference/src/Bundler.cpp:350-351)
//   * ransacMultiPairGPU                (/root/reference/src/cuda/cuda_ransac.h:50, called from
//                                        SiftManager::runRansacMultiPairGPU, /root/reference/src/FeatureManager.cpp:713)
//   * knnMatchBothDirections            (the two cv::cuda::DescriptorMatcher::knnMatch calls of SiftManager::findCorresbyNN,
//                                        /root/reference/src/FeatureManager.cpp:271-273)
//   * processDepthAndNormals            (Frame::processDepth + Frame::depthToCloudAndNormals, /root/reference/src/Frame.cpp:152-233)
// Header-only, no Eigen/yaml-cpp/OpenCV dependency of its own: the pose and intrinsics types are template parameters
// that only need operator()(row, col) (Eigen::Matrix4f / Matrix3f satisfy it), and the yml values arrive through
// BtSolverConfig, which the caller fills from the UNCHANGED config_*.yml keys (see INTEGRATION.md for the 6-line
// yaml-cpp version).  Errors are thrown as std::runtime_error with bt_last_error(); nothing exits or hangs.
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/bundletrack_b200.h"

struct BtSolverConfig {            // bundle.* / p2p.* of config_nocs.yml, same names
	int num_iter_outter = 7;
	int num_iter_inner = 5;
	float robust_delta = 0.005f;
	float image_downscale = 4.f;
	float p2p_max_dist = 0.02f;
	float p2p_max_normal_angle = 45.f;   // degrees, converted like CUDASolverBundling.cpp:94
	bt_solver_params to_params() const {
		bt_solver_params p;
		p.num_iter_outer = num_iter_outter; p.num_iter_inner = num_iter_inner; p.robust_delta = robust_delta; p.image_downscale = image_downscale;
		p.dense_dist_thresh = p2p_max_dist; p.dense_cos_normal_thresh = std::cos(p2p_max_normal_angle / 180.0 * M_PI);
		p.depth_min = 0.1f; p.depth_max = 9999.f; p.w_sparse = 1.f; p.w_dense = 1.f;
		return p;
	}
};

// Layout-compatible with the reference's EntryJ (SIFTImageManager.h:44-59); BundleTrack code can pass its own
// std::vector<EntryJ> through reinterpret_cast<const bt_entryj*>.
static_assert(sizeof(bt_entryj) == 32, "EntryJ layout");

class OptimizerGpu {
public:
	explicit OptimizerGpu(const BtSolverConfig& cfg, int device = 0, int max_frames = 15, int max_corr = 1 << 16, int H = 480, int W = 640)
	    : cfg_(cfg) {
		check(bt_ctx_create(&ctx_, device), "bt_ctx_create");
		bt_solver_limits lim = { 1, max_frames, max_corr, H, W, cfg.image_downscale };
		check(bt_solver_reserve(ctx_, &lim), "bt_solver_reserve");
	}
	~OptimizerGpu() { bt_ctx_destroy(ctx_); }
	OptimizerGpu(const OptimizerGpu&) = delete;
	OptimizerGpu& operator=(const OptimizerGpu&) = delete;

	// Same argument list as the reference.  n_match_per_pair and colors_gpu are unused there too (LossGPU.cu:53, SBA.cpp:28-32).
	// `poses` is in-out: every frame of the window is overwritten with its optimised cam->model pose.
	template <class EntryJT, class Uchar4T, class Float4T, class Mat4, class Alloc, class Mat3>
	void optimizeFrames(const std::vector<EntryJT>& global_corres, const std::vector<int>& /*n_match_per_pair*/, int n_frames, int H, int W,
	                    const std::vector<float*>& depths_gpu, const std::vector<Uchar4T*>& /*colors_gpu*/, const std::vector<Float4T*>& normals_gpu,
	                    std::vector<Mat4, Alloc>& poses, const Mat3& K, void* stream = nullptr) {
		static_assert(sizeof(EntryJT) == sizeof(bt_entryj), "EntryJ must be the 32-byte reference struct");
		std::vector<const float*> dptr(n_frames), nptr(n_frames);
		std::vector<float> flat(16 * (size_t)n_frames);
		for (int f = 0; f < n_frames; f++) {
			dptr[f] = depths_gpu[f];
			nptr[f] = reinterpret_cast<const float*>(normals_gpu[f]);
			for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) flat[16 * f + 4 * r + c] = poses[f](r, c);   // Eigen is column-major; the ABI is row-major
		}
		bt_window win{};
		win.n_frames = n_frames; win.H = H; win.W = W;
		win.n_corr = (int)global_corres.size();
		win.corr = reinterpret_cast<const bt_entryj*>(global_corres.data());
		win.depth_dev = dptr.data(); win.normal_dev = nptr.data();
		win.fx = K(0, 0); win.fy = K(1, 1); win.cx = K(0, 2); win.cy = K(1, 2);
		win.dense_pairs = nullptr; win.n_dense_pairs = 0; win.compat_flip = 1;
		const bt_solver_params prm = cfg_.to_params();
		check(bt_solve_windows(ctx_, 1, &win, &prm, flat.data(), stream), "bt_solve_windows");
		for (int f = 0; f < n_frames; f++)
			for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) poses[f](r, c) = flat[16 * f + 4 * r + c];
	}
	bt_ctx* ctx() { return ctx_; }

private:
	static void check(int rc, const char* what) {
		if (rc != BT_OK) throw std::runtime_error(std::string(what) + ": " + bt_last_error());
	}
	BtSolverConfig cfg_;
	bt_ctx* ctx_ = nullptr;
};

// ransacMultiPairGPU with the reference's argument list (device float4 arrays per pair, host result vectors).
template <class Float4T>
inline void ransacMultiPairGPU(bt_ctx* ctx, const std::vector<Float4T*>& ptsA, const std::vector<Float4T*>& ptsB, const std::vector<int>& n_pts,
                               int n_trials, float dist_thres, std::vector<std::vector<int>>& inlier_ids, void* stream = nullptr) {
	const int n = (int)ptsA.size();
	inlier_ids.assign(n, {});
	if (n == 0) return;
	int total = 0, maxp = 1;
	for (int v : n_pts) { total += v; if (v > maxp) maxp = v; }
	auto check = [](int rc, const char* what) { if (rc != BT_OK) throw std::runtime_error(std::string(what) + ": " + bt_last_error()); };
	check(bt_ransac_reserve(ctx, n, maxp, n_trials), "bt_ransac_reserve");
	std::vector<const float*> a(n), b(n);
	for (int p = 0; p < n; p++) { a[p] = reinterpret_cast<const float*>(ptsA[p]); b[p] = reinterpret_cast<const float*>(ptsB[p]); }
	void *d_ids = nullptr, *d_cnt = nullptr;
	check(bt_dev_alloc(&d_ids, sizeof(int32_t) * (size_t)(total > 0 ? total : 1)), "bt_dev_alloc");
	check(bt_dev_alloc(&d_cnt, sizeof(int32_t) * n), "bt_dev_alloc");
	check(bt_ransac_pairs(ctx, n, a.data(), b.data(), n_pts.data(), n_trials, dist_thres, 0, (int32_t*)d_ids, (int32_t*)d_cnt, stream), "bt_ransac_pairs");
	std::vector<int32_t> ids((size_t)(total > 0 ? total : 1)), cnt(n);
	check(bt_memcpy_d2h(cnt.data(), d_cnt, sizeof(int32_t) * n, stream), "bt_memcpy_d2h");
	check(bt_memcpy_d2h(ids.data(), d_ids, sizeof(int32_t) * ids.size(), stream), "bt_memcpy_d2h");
	int off = 0;
	for (int p = 0; p < n; p++) { inlier_ids[p].assign(ids.begin() + off, ids.begin() + off + cnt[p]); off += n_pts[p]; }
	bt_dev_free(d_ids); bt_dev_free(d_cnt);
}

// ---- Frame::processDepth + Frame::depthToCloudAndNormals (/root/reference/src/Frame.cpp:152-233) ----------------------------------
struct BtDepthConfig {            // depth_processing.* of config_nocs.yml, same names
	float erode_radius = 1, erode_diff = 0.001f, erode_ratio = 0.8f;
	int bf_radius = 2; float sigma_D = 2.f, sigma_R = 100000.f;
	bt_depth_params to_params() const { bt_depth_params p; p.erode_radius = (int)erode_radius; p.erode_diff = erode_diff; p.erode_ratio = erode_ratio;
		p.bf_radius = bf_radius; p.sigma_D = sigma_D; p.sigma_R = sigma_R; return p; }
};
// raw depth (device, metres) -> Frame::_depth_gpu, Frame::_normal_gpu and, if wanted, the camera-space point map that the
// reference copies to the host for its PCL cloud (xyz_gpu may be nullptr).  One fused kernel; depth_gpu != depth_raw_gpu.
template <class Float4T, class Mat3>
inline void processDepthAndNormals(bt_ctx* ctx, const float* depth_raw_gpu, float* depth_gpu, Float4T* normal_gpu, Float4T* xyz_gpu, int H, int W, const Mat3& K,
                                   const BtDepthConfig& cfg, void* stream = nullptr) {
	const bt_depth_params prm = cfg.to_params();
	const float* in[1] = { depth_raw_gpu }; float* out[1] = { depth_gpu };
	float* nrm[1] = { reinterpret_cast<float*>(normal_gpu) }; float* xyz[1] = { reinterpret_cast<float*>(xyz_gpu) };
	if (bt_frames_preprocess(ctx, 1, in, H, W, K(0, 0), K(1, 1), K(0, 2), K(1, 2), &prm, out, xyz_gpu ? xyz : nullptr, nrm, stream) != BT_OK)
		throw std::runtime_error(std::string("bt_frames_preprocess: ") + bt_last_error());
}

// ---- the two knnMatch calls of SiftManager::findCorresbyNN (/root/reference/src/FeatureManager.cpp:271-273) in one call --------
// desA/desB: device CV_32F descriptor matrices (GpuMat::data, rows, step).  DMatchT needs queryIdx, trainIdx, distance (cv::DMatch).
template <class DMatchT>
inline void knnMatchBothDirections(bt_ctx* ctx, const float* desA, int nA, size_t stepA, const float* desB, int nB, size_t stepB, int dim, int k,
                                   std::vector<std::vector<DMatchT>>& matchesAB, std::vector<std::vector<DMatchT>>& matchesBA, void* stream = nullptr) {
	auto check = [](int rc, const char* what) { if (rc != BT_OK) throw std::runtime_error(std::string(what) + ": " + bt_last_error()); };
	check(bt_matcher_reserve(ctx, 1, nA > nB ? nA : nB, dim), "bt_matcher_reserve");
	bt_desc_view A{}; A.dev = desA; A.n = nA; A.dim = dim; A.pitch_bytes = stepA;
	bt_desc_view B{}; B.dev = desB; B.n = nB; B.dim = dim; B.pitch_bytes = stepB;
	void *iab = nullptr, *dab = nullptr, *iba = nullptr, *dba = nullptr;
	const size_t ea = (size_t)(nA > 0 ? nA : 1) * k, eb = (size_t)(nB > 0 ? nB : 1) * k;
	check(bt_dev_alloc(&iab, 4 * ea), "bt_dev_alloc"); check(bt_dev_alloc(&dab, 4 * ea), "bt_dev_alloc");
	check(bt_dev_alloc(&iba, 4 * eb), "bt_dev_alloc"); check(bt_dev_alloc(&dba, 4 * eb), "bt_dev_alloc");
	check(bt_knn_match_pairs(ctx, 1, &A, &B, k, (int32_t*)iab, (float*)dab, (int32_t*)iba, (float*)dba, stream), "bt_knn_match_pairs");
	std::vector<int32_t> hiab(ea), hiba(eb); std::vector<float> hdab(ea), hdba(eb);
	check(bt_memcpy_d2h(hiab.data(), iab, 4 * ea, stream), "bt_memcpy_d2h"); check(bt_memcpy_d2h(hdab.data(), dab, 4 * ea, stream), "bt_memcpy_d2h");
	check(bt_memcpy_d2h(hiba.data(), iba, 4 * eb, stream), "bt_memcpy_d2h"); check(bt_memcpy_d2h(hdba.data(), dba, 4 * eb, stream), "bt_memcpy_d2h");
	auto fill = [k](int n, const std::vector<int32_t>& idx, const std::vector<float>& dist, std::vector<std::vector<DMatchT>>& out) {
		out.assign(n, {});
		for (int q = 0; q < n; q++)
			for (int j = 0; j < k; j++) {
				if (idx[(size_t)q * k + j] < 0) break;          // fewer than k train rows
				DMatchT m{}; m.queryIdx = q; m.trainIdx = idx[(size_t)q * k + j]; m.distance = dist[(size_t)q * k + j];
				out[q].push_back(m);
			}
	};
	fill(nA, hiab, hdab, matchesAB); fill(nB, hiba, hdba, matchesBA);
	bt_dev_free(iab); bt_dev_free(dab); bt_dev_free(iba); bt_dev_free(dba);
}
