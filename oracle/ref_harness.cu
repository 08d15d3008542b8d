/*
 * oracle/ref_harness.cu — TEST INFRASTRUCTURE ONLY ("Oracle B").
 *
 * The reference's device code for the hot path (Solver/SolverBundling.cu, SBA.cu, CUDAImageUtil.cu,
 * cuda_ransac.cu) is compiled VERBATIM from /root/reference by oracle/Makefile and linked with this file into
 * oracle/_ref/libbt_ref.so.  The reference's host glue (LossGPU.cu, SBA.cpp, CUDACache.cpp, CUDASolverBundling.cpp)
 * cannot be compiled here (mLib -> Eigen/Sparse, yaml-cpp, absent offline), so this file re-does ONLY that glue —
 * same allocation pattern, same call order — and hands the work to the reference's own extern "C" stubs.
 * Nothing in bundletrack_b200/ links or loads this.
 *
 * Each function cites the reference host code whose sequence it follows (paths relative to /root/reference/src/cuda).
 */
#include <cstdio>
#include <cstring>
#include <chrono>
#include <limits>
#include <map>
#include <vector>
#include <cuda_runtime.h>

#include "GlobalDefines.h"
#include "SolverBundlingParameters.h"
#include "SolverBundlingState.h"
#include "CUDAImageUtil.h"
#include "cuda_ransac.h"

class CUDATimer;
extern "C" void buildVariablesToCorrespondencesTableCUDA(EntryJ* d_correspondences, unsigned int numberOfCorrespondences, unsigned int maxNumCorrespondencesPerImage, int* d_variablesToCorrespondences, int* d_numEntriesPerRow, CUDATimer* timer);
extern "C" void solveBundlingStub(SolverInput& input, SolverState& state, SolverParameters& parameters, SolverStateAnalysis& analysis, float* convergenceAnalysis, CUDATimer* timer);
extern "C" void convertMatricesToPosesCU(const float4x4* d_transforms, unsigned int numTransforms, float3* d_rot, float3* d_trans, const int* d_validImages);
extern "C" void convertPosesToMatricesCU(const float3* d_rot, const float3* d_trans, unsigned int numImages, float4x4* d_transforms, const int* d_validImages);

struct ref_params { /* mirrors oracle_params / bt_solver_params */
	int num_iter_outer, num_iter_inner;
	float robust_delta, image_downscale, dense_dist_thresh, dense_cos_normal_thresh, depth_min, depth_max, w_sparse, w_dense;
};

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "ref_harness: %s -> %s\n", #x, cudaGetErrorString(e_)); return -100; } } while (0)

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

/* OptimizerGpu::optimizeFrames (LossGPU.cu:53-139) with SBA::SBA/init/align (SBA.cpp:17-130),
 * CUDASolverBundling ctor/solve/dtor (Solver/CUDASolverBundling.cpp:22-288) and CUDACache ctor/storeFrame/dtor
 * (CUDACache.cpp:14-88, CUDACacheUtil.h:11-39) inlined.  Outputs the ordered dense pair list the reference ended up
 * with and two wall-clock times: the whole call, and the m_solver->solve region (SBA.cpp:132-135). */
extern "C" int ref_optimize_frames(int N, int H, int W, const float* const* d_depths, const float4* const* d_normals,
                                   float fx, float fy, float cx, float cy, int n_corr, const EntryJ* h_corr,
                                   float* h_poses_inout, const ref_params* prm,
                                   unsigned int* h_pairs_out /*[N(N-1)/2][2]*/, int* n_pairs_out,
                                   double* t_outer_ms, double* t_solve_ms) {
	CK(cudaDeviceSynchronize());
	const double t0 = now_ms();
	const int Wd = (int)(W / prm->image_downscale), Hd = (int)(H / prm->image_downscale);   /* LossGPU.cu:55-57 */

	/* --- CUDACache::CUDACache + alloc (CUDACache.cpp:14-58) --- */
	float4 intr = make_float4(fx * (float)Wd / (float)W, fy * (float)Hd / (float)H,
	                          cx * (float)(Wd - 1) / (float)(W - 1), cy * (float)(Hd - 1) / (float)(H - 1));
	float4x4 Kinv; Kinv.setIdentity();       /* m_inputIntrinsicsInv of [[fx,0,cx,0],[0,fy,cy,0],[0,0,1,0],[0,0,0,1]] */
	Kinv(0,0) = 1.0f / fx; Kinv(1,1) = 1.0f / fy; Kinv(0,2) = -cx / fx; Kinv(1,2) = -cy / fy;
	std::vector<CUDACachedFrame> cache(N);
	for (auto& f : cache) f.alloc(Wd, Hd);
	CUDACachedFrame* d_cache; float *d_intensityHelper, *d_filterHelper; float4 *d_helperCamPos, *d_helperNormals;
	CK(cudaMalloc(&d_cache, sizeof(CUDACachedFrame) * N));
	CK(cudaMemcpy(d_cache, cache.data(), sizeof(CUDACachedFrame) * N, cudaMemcpyHostToDevice));
	CK(cudaMalloc(&d_intensityHelper, sizeof(float) * Wd * Hd));
	CK(cudaMalloc(&d_filterHelper, sizeof(float) * W * H));
	CK(cudaMalloc(&d_helperCamPos, sizeof(float4) * W * H));
	CK(cudaMalloc(&d_helperNormals, sizeof(float4) * W * H));
	for (int i = 0; i < N; i++) {             /* CUDACache::storeFrame (CUDACache.cpp:76-88) */
		CUDACachedFrame& frame = cache[i];
		CUDAImageUtil::convertDepthFloatToCameraSpaceFloat4(d_helperCamPos, d_depths[i], Kinv, W, H);
		CUDAImageUtil::resampleFloat4(frame.d_cameraposDownsampled, Wd, Hd, d_helperCamPos, W, H);
		CUDAImageUtil::resampleFloat4(frame.d_normalsDownsampled, Wd, Hd, d_normals[i], W, H);
		CUDAImageUtil::resampleFloat(frame.d_depthDownsampled, Wd, Hd, d_depths[i], W, H);
		CUDAImageUtil::countNumValidDepth(frame.d_num_valid_points, frame.d_depthDownsampled, Hd, Wd);
	}

	/* --- poses up (LossGPU.cu:84-98): row-major float4x4 --- */
	float4x4* d_transforms;
	CK(cudaMalloc(&d_transforms, sizeof(float4x4) * N));
	CK(cudaMemcpy(d_transforms, h_poses_inout, sizeof(float4x4) * N, cudaMemcpyHostToDevice));

	const unsigned maxNumResiduals = (unsigned)N * (N - 1) / 2 * Hd * Wd / 4 + n_corr;   /* LossGPU.cu:102 */
	std::map<int, int> per; int max_corr_per_image = 0;                                    /* :104-115 */
	for (int i = 0; i < n_corr; i++) { per[h_corr[i].imgIdx_i]++; per[h_corr[i].imgIdx_j]++; }
	for (auto& kv : per) max_corr_per_image = kv.second > max_corr_per_image ? kv.second : max_corr_per_image;

	/* --- SBA::init (SBA.cpp:55-78) + CUDASolverBundling ctor (CUDASolverBundling.cpp:22-135) --- */
	float3 *d_xRot, *d_xTrans;
	CK(cudaMalloc(&d_xRot, sizeof(float3) * N)); CK(cudaMalloc(&d_xTrans, sizeof(float3) * N));
	SolverState st; SolverStateAnalysis ex; memset(&st, 0, sizeof st); memset(&ex, 0, sizeof ex);
	int *d_variablesToCorrespondences, *d_numEntriesPerRow;
	const unsigned nv = N;
	const unsigned nred = (maxNumResiduals + 512 - 1) / 512;
	const unsigned maxPairs = nv * (nv - 1) / 2;
	struct A { void** p; size_t bytes; };
	std::vector<A> allocs = {
		{ (void**)&st.d_deltaRot, sizeof(float3) * nv }, { (void**)&st.d_deltaTrans, sizeof(float3) * nv },
		{ (void**)&st.d_rRot, sizeof(float3) * nv }, { (void**)&st.d_rTrans, sizeof(float3) * nv },
		{ (void**)&st.d_zRot, sizeof(float3) * nv }, { (void**)&st.d_zTrans, sizeof(float3) * nv },
		{ (void**)&st.d_pRot, sizeof(float3) * nv }, { (void**)&st.d_pTrans, sizeof(float3) * nv },
		{ (void**)&st.d_Jp, sizeof(float3) * maxNumResiduals },
		{ (void**)&st.d_Ap_XRot, sizeof(float3) * nv }, { (void**)&st.d_Ap_XTrans, sizeof(float3) * nv },
		{ (void**)&st.d_scanAlpha, sizeof(float) * 2 }, { (void**)&st.d_rDotzOld, sizeof(float) * nv },
		{ (void**)&st.d_precondionerRot, sizeof(float3) * nv }, { (void**)&st.d_precondionerTrans, sizeof(float3) * nv },
		{ (void**)&st.d_sumResidual, sizeof(float) },
		{ (void**)&ex.d_maxResidual, sizeof(float) * nred }, { (void**)&ex.d_maxResidualIndex, sizeof(int) * nred },
		{ (void**)&d_variablesToCorrespondences, sizeof(int) * nv * (size_t)max_corr_per_image }, { (void**)&d_numEntriesPerRow, sizeof(int) * nv },
		{ (void**)&st.d_countHighResidual, sizeof(int) },
		{ (void**)&st.d_denseJtJ, sizeof(float) * 36 * nv * nv }, { (void**)&st.d_denseJtr, sizeof(float) * 6 * nv },
		{ (void**)&st.d_denseCorrCounts, sizeof(float) * maxPairs }, { (void**)&st.d_denseOverlappingImages, sizeof(uint2) * maxPairs },
		{ (void**)&st.d_numDenseOverlappingImages, sizeof(int) },
		{ (void**)&st.d_corrCount, sizeof(int) }, { (void**)&st.d_corrCountColor, sizeof(int) }, { (void**)&st.d_sumResidualColor, sizeof(float) },
		{ (void**)&st.d_xTransforms, sizeof(float4x4) * nv }, { (void**)&st.d_xTransformInverses, sizeof(float4x4) * nv },
	};
	for (auto& a : allocs) CK(cudaMalloc(a.p, a.bytes > 0 ? a.bytes : 4));
	ex.h_maxResidual = new float[nred]; ex.h_maxResidualIndex = new int[nred];
	for (size_t k = 0; k + 2 < allocs.size(); k++) CK(cudaMemset(*allocs[k].p, -1, allocs[k].bytes));   /* 29 memset(-1): all but the two transform arrays */

	SolverParameters P; memset(&P, 0, sizeof P);                /* m_defaultParams (:92-99) + solve() (:193-220) */
	P.denseDistThresh = prm->dense_dist_thresh; P.denseNormalThresh = prm->dense_cos_normal_thresh;
	P.denseColorThresh = 0.1f; P.denseColorGradientMin = 0.005f;
	P.denseDepthMin = prm->depth_min; P.denseDepthMax = prm->depth_max; P.denseOverlapCheckSubsampleFactor = 1;
	P.nNonLinearIterations = prm->num_iter_outer; P.nLinIterations = prm->num_iter_inner;
	P.verifyOptDistThresh = 0.02f; P.verifyOptPercentThresh = 0.05f;
	P.highResidualThresh = std::numeric_limits<float>::infinity();
	P.robust_delta = prm->robust_delta;
	std::vector<float> wS(prm->num_iter_outer, prm->w_sparse), wD(prm->num_iter_outer, prm->w_dense), wC(prm->num_iter_outer, 0.0f); /* SBA.cpp:27-32 */
	P.weightSparse = wS[0]; P.weightDenseDepth = wD[0]; P.weightDenseColor = wC[0];
	P.useDense = (P.weightDenseDepth > 0 || P.weightDenseColor > 0);
	P.useDenseDepthAllPairwise = true;                          /* SBA.cpp:91 usePairwise = true */

	/* --- SBA::align (SBA.cpp:81-130) --- */
	int* d_validImages; EntryJ* d_corr;
	std::vector<int> valid(N, 1);
	CK(cudaMalloc(&d_validImages, N * sizeof(int)));
	CK(cudaMemcpy(d_validImages, valid.data(), sizeof(int) * N, cudaMemcpyHostToDevice));
	CK(cudaMalloc(&d_corr, sizeof(EntryJ) * (n_corr > 0 ? n_corr : 1)));
	CK(cudaMemcpy(d_corr, h_corr, sizeof(EntryJ) * n_corr, cudaMemcpyHostToDevice));
	convertMatricesToPosesCU(d_transforms, N, d_xRot, d_xTrans, d_validImages);

	const double t1 = now_ms();                                  /* SBA::alignCUDA timer start (SBA.cpp:132) */
	st.d_xRot = d_xRot; st.d_xTrans = d_xTrans;
	SolverInput in; memset(&in, 0, sizeof in);
	in.d_correspondences = d_corr; in.d_variablesToCorrespondences = d_variablesToCorrespondences; in.d_numEntriesPerRow = d_numEntriesPerRow;
	in.numberOfImages = N; in.numberOfCorrespondences = n_corr;
	in.maxNumberOfImages = N; in.maxCorrPerImage = max_corr_per_image; in.maxNumDenseImPairs = maxPairs;
	in.weightsSparse = wS.data(); in.weightsDenseDepth = wD.data(); in.weightsDenseColor = wC.data();
	in.d_validImages = d_validImages; in.d_cacheFrames = d_cache;
	in.denseDepthWidth = Wd; in.denseDepthHeight = Hd; in.intrinsics = intr;
	CK(cudaMemset(d_numEntriesPerRow, 0, sizeof(int) * N));      /* buildVariablesToCorrespondencesTable (:276-282) */
	if (n_corr > 0) buildVariablesToCorrespondencesTableCUDA(d_corr, n_corr, max_corr_per_image, d_variablesToCorrespondences, d_numEntriesPerRow, NULL);
	solveBundlingStub(in, st, P, ex, NULL, NULL);
	const double t2 = now_ms();                                  /* SBA.cpp:134 (no device sync there either) */

	convertPosesToMatricesCU(d_xRot, d_xTrans, N, d_transforms, d_validImages);
	int np = 0;
	CK(cudaMemcpy(&np, st.d_numDenseOverlappingImages, sizeof(int), cudaMemcpyDeviceToHost));
	if (n_pairs_out) *n_pairs_out = np;
	if (h_pairs_out && np > 0 && np <= (int)maxPairs) CK(cudaMemcpy(h_pairs_out, st.d_denseOverlappingImages, sizeof(uint2) * np, cudaMemcpyDeviceToHost));
	CK(cudaFree(d_validImages)); CK(cudaFree(d_corr));
	CK(cudaMemcpy(h_poses_inout, d_transforms, sizeof(float4x4) * N, cudaMemcpyDeviceToHost));   /* LossGPU.cu:120 */
	CK(cudaFree(d_transforms));
	/* ~SBA, ~CUDASolverBundling, ~CUDACache */
	CK(cudaFree(d_xRot)); CK(cudaFree(d_xTrans));
	for (auto& a : allocs) CK(cudaFree(*a.p));
	delete[] ex.h_maxResidual; delete[] ex.h_maxResidualIndex;
	for (auto& f : cache) f.free();
	CK(cudaFree(d_cache)); CK(cudaFree(d_intensityHelper)); CK(cudaFree(d_filterHelper)); CK(cudaFree(d_helperCamPos)); CK(cudaFree(d_helperNormals));
	CK(cudaDeviceSynchronize());
	const double t3 = now_ms();
	if (t_outer_ms) *t_outer_ms = t3 - t0;
	if (t_solve_ms) *t_solve_ms = t2 - t1;
	return 0;
}

/* SiftManager::runRansacMultiPairGPU's device call (FeatureManager.cpp:701-717): upload model-frame points,
 * ransacMultiPairGPU (cuda_ransac.cu:1228-1323), return inlier ids. */
extern "C" int ref_ransac_pairs(int n_pairs, const float* const* h_ptsA /*[n][4]*/, const float* const* h_ptsB, const int* n_pts,
                                int n_trials, float dist_thres, int* inlier_ids_out /*concatenated*/, int* n_inliers_out, double* t_ms) {
	CK(cudaDeviceSynchronize());
	const double t0 = now_ms();
	std::vector<float4*> A(n_pairs), B(n_pairs); std::vector<int> n(n_pts, n_pts + n_pairs);
	for (int p = 0; p < n_pairs; p++) {
		CK(cudaMalloc(&A[p], sizeof(float4) * n[p])); CK(cudaMalloc(&B[p], sizeof(float4) * n[p]));
		CK(cudaMemcpy(A[p], h_ptsA[p], sizeof(float4) * n[p], cudaMemcpyHostToDevice));
		CK(cudaMemcpy(B[p], h_ptsB[p], sizeof(float4) * n[p], cudaMemcpyHostToDevice));
	}
	std::vector<std::vector<int>> inl;
	ransacMultiPairGPU(A, B, n, n_trials, dist_thres, inl);
	for (int p = 0; p < n_pairs; p++) { CK(cudaFree(A[p])); CK(cudaFree(B[p])); }
	CK(cudaDeviceSynchronize());
	if (t_ms) *t_ms = now_ms() - t0;
	int o = 0;
	for (int p = 0; p < n_pairs; p++) {
		n_inliers_out[p] = (int)inl[p].size();
		for (int v : inl[p]) inlier_ids_out[o++] = v;
	}
	return 0;
}

/* Frame::Frame -> updateDepthGPU, processDepth, depthToCloudAndNormals (/root/reference/src/Frame.cpp:62-81,107-127,152-233)
 * with the reference's own kernels and its allocation pattern (cudaMalloc/cudaFree of the two temporaries per frame, the
 * host round trips of the maps).  K^-1 as Eigen::Matrix3f::inverse() forms it (cofactors times 1/det, Frame.cpp:189). */
extern "C" int ref_frame_preprocess(int H, int W, const float* h_depth_raw, float fx, float fy, float cx, float cy,
                                    int erode_radius, float erode_diff, float erode_ratio, int bf_radius, float sigma_D, float sigma_R,
                                    float* h_depth_out, float* h_xyz_out /*[H*W*4]*/, float* h_normal_out /*[H*W*4]*/, double* t_ms) {
	CK(cudaDeviceSynchronize());
	const double t0 = now_ms();
	const int n = H * W;
	float *d_depth, *d_tmp; float4 *d_normal, *d_xyz;
	CK(cudaMalloc(&d_depth, sizeof(float) * n)); CK(cudaMalloc(&d_normal, sizeof(float4) * n));      /* Frame ctor, Frame.cpp:68-70 */
	CK(cudaMemcpy(d_depth, h_depth_raw, sizeof(float) * n, cudaMemcpyHostToDevice));                 /* updateDepthGPU */
	CK(cudaMalloc(&d_tmp, sizeof(float) * n));                                                       /* processDepth */
	CUDAImageUtil::erodeDepthMap(d_tmp, d_depth, erode_radius, W, H, erode_diff, erode_ratio);
	CUDAImageUtil::gaussFilterDepthMap(d_depth, d_tmp, bf_radius, sigma_D, sigma_R, W, H);
	CUDAImageUtil::gaussFilterDepthMap(d_tmp, d_depth, bf_radius, sigma_D, sigma_R, W, H);
	{ float* t = d_depth; d_depth = d_tmp; d_tmp = t; }
	CK(cudaMemcpy(h_depth_out, d_depth, sizeof(float) * n, cudaMemcpyDeviceToHost));                 /* updateDepthCPU */
	CK(cudaFree(d_tmp));
	CK(cudaMalloc(&d_xyz, sizeof(float4) * n));                                                      /* depthToCloudAndNormals */
	float4x4 Kinv; Kinv.setIdentity();
	{
		const float det = fx * fy, invdet = 1.0f / det;
		Kinv(0,0) = fy * invdet; Kinv(0,1) = 0.0f; Kinv(0,2) = (0.0f * cy - cx * fy) * invdet;
		Kinv(1,0) = 0.0f; Kinv(1,1) = fx * invdet; Kinv(1,2) = -(fx * cy - cx * 0.0f) * invdet;
		Kinv(2,0) = 0.0f; Kinv(2,1) = 0.0f; Kinv(2,2) = 1.0f;
	}
	CUDAImageUtil::convertDepthFloatToCameraSpaceFloat4(d_xyz, d_depth, Kinv, W, H);
	CUDAImageUtil::computeNormals(d_normal, d_xyz, W, H);
	CK(cudaMemcpy(h_xyz_out, d_xyz, sizeof(float4) * n, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(h_normal_out, d_normal, sizeof(float4) * n, cudaMemcpyDeviceToHost));
	CK(cudaFree(d_xyz)); CK(cudaFree(d_depth)); CK(cudaFree(d_normal));
	CK(cudaDeviceSynchronize());
	if (t_ms) *t_ms = now_ms() - t0;
	return 0;
}
