/*
 * bundletrack_b200.h — C-ABI of the B200-native pose-graph optimizer + feature matcher that drops in behind
 * BundleTrack's C++ surface.  Plain pointers and sizes only; no C++/torch types.  Every entry point returns
 * BT_OK (0) or a negative bt_status — it never exits, throws or hangs (the reference does all three).
 *
 * The reference has no plugin/FFI layer; the "boundary" is the three C++ call sites where BundleTrack's host code
 * hands device work to something replaceable.  Each entry point below names the call it replaces.
 * INTEGRATION.md shows the reference-side shim a maintainer adds at each site.
 *
 * Threading: a bt_ctx may be used by one host thread at a time.  All work is enqueued on the cudaStream_t passed
 * in (as void*); entry points that return host data synchronise that stream before returning.
 */
#ifndef BUNDLETRACK_B200_H
#define BUNDLETRACK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#define BT_API __attribute__((visibility("default")))
#else
#define BT_API
#endif

typedef enum {
	BT_OK = 0,
	BT_ERR_INVALID_ARG = -1,
	BT_ERR_CAPACITY = -2,     /* more windows / frames / correspondences / features than the context was created for */
	BT_ERR_CUDA = -3,         /* a CUDA runtime call failed; bt_last_error() has the string */
	BT_ERR_NO_DEVICE = -4,    /* no CUDA device, or not an sm_100 part: there is NO CPU fallback */
	BT_ERR_UNSUPPORTED = -5
} bt_status;

typedef struct bt_ctx bt_ctx;

/* One 3D-3D feature correspondence.  Layout-identical to the reference's `struct EntryJ`
 * (/root/reference/src/cuda/SIFTImageManager.h:44-59): 32 bytes, invalid <=> imgIdx_i == 0xFFFFFFFF. */
typedef struct {
	uint32_t imgIdx_i;
	uint32_t imgIdx_j;
	float pos_i[3];   /* camera-frame point in frame i, metres */
	float pos_j[3];
} bt_entryj;

/* Solver parameters, filled from the UNCHANGED config_*.yml keys:
 *   num_iter_outer  <- bundle.num_iter_outter      (Solver/CUDASolverBundling.cpp:193)
 *   num_iter_inner  <- bundle.num_iter_inner       (:210)
 *   robust_delta    <- bundle.robust_delta         (:214)
 *   image_downscale <- bundle.image_downscale      (LossGPU.cu:55)
 *   dense_dist_thresh        <- p2p.max_dist                       (CUDASolverBundling.cpp:93)
 *   dense_cos_normal_thresh  <- cos(p2p.max_normal_angle * pi/180) (:94)
 *   depth_min / depth_max    =  0.1 / 9999 (hard-wired, :97-98)
 *   w_sparse / w_dense       =  1 / 1      (hard-wired, SBA.cpp:28-30)                                   */
typedef struct {
	int num_iter_outer;
	int num_iter_inner;
	float robust_delta;
	float image_downscale;
	float dense_dist_thresh;
	float dense_cos_normal_thresh;
	float depth_min;
	float depth_max;
	float w_sparse;
	float w_dense;
} bt_solver_params;

/* One tracking window = one OptimizerGpu::optimizeFrames call (/root/reference/src/cuda/LossGPU.cu:53). */
typedef struct {
	int n_frames;                     /* N, frame 0 is the gauge (never moves) */
	int H, W;                         /* full-resolution image size */
	int n_corr;
	const bt_entryj* corr;            /* HOST pointer, n_corr entries (std::vector<EntryJ>::data()).  Pageable memory is copied
	                                   * during the call.  PAGE-LOCKED memory (cudaHostAlloc / cudaHostRegister /
	                                   * bt_host_alloc_pinned) is read in place by the copy engine - windows whose buffers follow
	                                   * each other share one transfer - and must stay unchanged until bt_solve_windows returns
	                                   * (split API: until bt_solve_fetch). */
	const float* const* depth_dev;    /* HOST array of N DEVICE pointers: Frame::_depth_gpu, float[H*W], metres, 0 = invalid */
	const float* const* normal_dev;   /* HOST array of N DEVICE pointers: Frame::_normal_gpu, float4[H*W], (nx,ny,nz,0) */
	float fx, fy, cx, cy;             /* K at full resolution */
	/* Dense point-to-plane pair list as (target, source) frame indices.  NULL => every unordered pair once with
	 * target = the larger index, which is what the reference's FindImageImageCorr_Kernel produces when the
	 * per-frame d_num_valid_points allocations have ascending addresses.  n_dense_pairs == 0 with a
	 * non-NULL pointer disables the dense term for this window. */
	const uint32_t* dense_pairs;
	int n_dense_pairs;
	/* 1 => reproduce the reference's FlipJtJ behaviour: a pair's cross block survives only when target < source
	 * (SURVEY.md Q2).  0 => keep every cross block (the mathematically complete Gauss-Newton system). */
	int compat_flip;
	/* Optional: N slots of the context's frame cache (bt_frame_cache_store).  When non-NULL, depth_dev / normal_dev are not
	 * read: the quarter-resolution maps CUDACache::storeFrame (CUDACache.cpp:76-88) would rebuild for this call were built
	 * once per keyframe.  Results are bit-identical either way. */
	const int32_t* cache_slots;
	/* Optional: correspondences that are ALREADY ON THE DEVICE, as bt_match_pairs leaves them: no entry
	 * crosses the PCIe bus.  corr_dev: DEVICE EntryJ array holding n_blocks blocks back to back; block b has block_n[b]
	 * entries that all carry (imgIdx_i, imgIdx_j) = (block_i[b], block_j[b]) and starts at block_off[b] (entries), with
	 * block_off[b+1] == block_off[b] + block_n[b].  block_* are HOST arrays - bt_match_pairs' n_entry_out / entry_off_out read
	 * back (a few dozen ints) and the pair list the caller gave it.  When corr_dev is non-NULL, corr / n_corr are ignored. */
	const bt_entryj* corr_dev;
	/* With HOST correspondences (corr_dev == NULL) block_n / n_blocks are optional too: the caller states that `corr` is n_blocks blocks
	 * back to back, block b holding block_n[b] entries of ONE (imgIdx_i, imgIdx_j) pair (read from the block's first entry, or from
	 * block_i / block_j when given) - what Bundler::optimizeGPU's n_match_per_pair describes (Bundler.cpp:298-351).  The library then
	 * skips its own grouping pass over the entries.  block_off is not read for host correspondences. */
	int n_blocks;
	const int32_t* block_off;
	const int32_t* block_n;
	const uint32_t* block_i;
	const uint32_t* block_j;
} bt_window;

typedef struct {
	int max_windows;        /* windows per bt_solve_* call */
	int max_frames;         /* frames per window (<= 32) */
	int max_corr;           /* correspondences per window */
	int H, W;               /* largest full-resolution frame */
	float image_downscale;  /* bundle.image_downscale */
} bt_solver_limits;

BT_API const char* bt_last_error(void);
BT_API int bt_version(void);

/* Create / destroy.  `device` is a CUDA ordinal.  The context owns every scratch buffer, so the per-call path does
 * no cudaMalloc/cudaFree (the reference does ~110 per call). */
BT_API int bt_ctx_create(bt_ctx** out, int device);
BT_API void bt_ctx_destroy(bt_ctx* ctx);
BT_API int bt_solver_reserve(bt_ctx* ctx, const bt_solver_limits* lim);

/* Replaces OptimizerGpu::optimizeFrames (/root/reference/src/cuda/LossGPU.cu:53-139; caller
 * /root/reference/src/Bundler.cpp:350-351) for a BATCH of independent windows.
 * poses_inout: HOST, row-major 4x4 cam->model per frame, windows concatenated (sum of n_frames * 16 floats);
 * overwritten with the optimised poses of EVERY frame of the window, exactly like the reference's in-out vector.
 * Equivalent to bt_solve_stage + bt_solve_run + bt_solve_fetch. */
BT_API int bt_solve_windows(bt_ctx* ctx, int n_windows, const bt_window* windows, const bt_solver_params* params,
                     float* poses_inout, void* stream);

/* Streaming form: `begin` stages + launches a batch and queues the download of its poses, `end` waits for the OLDEST batch begun and
 * copies its poses (same layout as bt_solve_windows) to poses_out.  At most two batches may be in flight: the host side of batch k+1
 * then overlaps the GPU side of batch k.  The windows' host arrays (corr, pointer tables, block arrays) may be reused as soon as
 * `begin` returns unless they are page-locked and read in place (then: until the matching `end`). */
BT_API int bt_solve_windows_begin(bt_ctx* ctx, int n_windows, const bt_window* windows, const bt_solver_params* params,
                           const float* poses_in, void* stream);
BT_API int bt_solve_windows_end(bt_ctx* ctx, float* poses_out);

/* The same call split at the host<->device boundary, so a caller (and bench.py) can keep inputs resident:
 *   stage : host -> device copies of correspondences, poses, window tables (async on `stream`)
 *   run   : the kernels only (frame cache, plan, persistent GN/PCG solve); asynchronous
 *   fetch : device -> host copy of the poses + stream synchronise                                       */
BT_API int bt_solve_stage(bt_ctx* ctx, int n_windows, const bt_window* windows, const bt_solver_params* params,
                   const float* poses_in, void* stream);
BT_API int bt_solve_run(bt_ctx* ctx, void* stream);
BT_API int bt_solve_fetch(bt_ctx* ctx, float* poses_out, void* stream);

/* Introspection used by tests / bench (device pointers stay owned by the context). */
typedef struct {
	int n_windows;
	int n_tiles_total;         /* dense tiles per GN iteration over the batch */
	int n_kernel_launches;     /* kernels enqueued by the last bt_solve_run */
	long long n_src_pixels;    /* valid quarter-res source pixels summed over (window, pair) */
} bt_solve_stats;
BT_API int bt_solve_get_stats(bt_ctx* ctx, bt_solve_stats* out);
/* Copies the dense system (6N x 6N JtJ, row-major, reference ordering trans 0-2 / rot 3-5, and 6N Jtr) that the
 * LAST Gauss-Newton iteration of window `w` assembled; debugging/parity aid. */
/* Per-kernel device time of the last bt_solve_run (ms3 = {frame cache, plan, persistent solve}), CUDA events on the
 * launching stream; opt-in because it adds four event records per run. */
BT_API int bt_solve_enable_timing(bt_ctx* ctx, int on);
BT_API int bt_solve_get_timing(bt_ctx* ctx, float* ms3);
/* Host-side time of the last bt_solve_windows call, microseconds: us6 = {window tables + first upload, frame-preparation launch,
 * correspondence scan + staging + uploads, k_solve launch, pose download + wait for the GPU, whole call}. */
BT_API int bt_solve_get_host_timing(bt_ctx* ctx, double* us6);
/* Developer aid: in-kernel phase timestamps (clock64) of every tile and tail of the next runs; 12 int64 per record:
 * {kind<<32|cta, ids, t[10]}.  bt_solve_get_profile returns the number of records copied (>= 0) or a negative status. */
BT_API int bt_solve_enable_profile(bt_ctx* ctx, int max_records);
BT_API int bt_solve_get_profile(bt_ctx* ctx, long long* out, int max_records);
BT_API int bt_solve_enable_debug(bt_ctx* ctx, int on);
BT_API int bt_solve_debug_dense(bt_ctx* ctx, int w, float* JtJ_out, float* Jtr_out);
/* Number of dense correspondences each pair found in the last GN iteration of window `w` (gating parity aid). */
BT_API int bt_solve_debug_counts(bt_ctx* ctx, int w, int n_pairs, float* counts_out);

/* ------------------------------------------------------------------------------------------------------------
 * Matcher.  Replaces the two cv::cuda::DescriptorMatcher::knnMatch calls in SiftManager::findCorresbyNN
 * (/root/reference/src/FeatureManager.cpp:271-273) for a batch of frame pairs.
 * Distances are sqrt(sum d^2) like cv::NORM_L2, ascending; ties -> lower train index first.                  */
typedef struct {
	const float* dev;      /* DEVICE pointer, row-major n x dim (cv::cuda::GpuMat data), fp32 */
	int n;
	int dim;               /* 256 for LF-Net; must be a multiple of 32 */
	size_t pitch_bytes;    /* GpuMat::step; 0 => dim*4 */
} bt_desc_view;

BT_API int bt_matcher_reserve(bt_ctx* ctx, int max_pairs, int max_feats, int dim);
/* idxAB/distAB: DEVICE [sum_p nA_p * k] results for queries in A against train B; idxBA/distBA the reverse
 * direction.  ONE tensor-core contraction per pair serves both directions (rows of its accumulator tile are reduced for
 * A->B, columns for B->A).  Pair p's block starts at k * (sum of nA over pairs < p) (resp. nB).  Rows with fewer than k
 * candidates are padded with idx -1 / dist +inf. */
BT_API int bt_knn_match_pairs(bt_ctx* ctx, int n_pairs, const bt_desc_view* A, const bt_desc_view* B, int k,
                       int32_t* idxAB, float* distAB, int32_t* idxBA, float* distBA, void* stream);
/* Persistent descriptor pool.  Lfnet::detectFeature uploads a frame's descriptors once
 * (/root/reference/src/FeatureManager.cpp:876-908); bt_desc_pool_store converts them ONCE into pool slot `slot` (fp16 rows for the
 * tensor pass, fp32 copy for the exact stage, norms, measured rounding errors - the caller's buffer is not referenced afterwards) and
 * bt_knn_match_slots / bt_match_pairs_pool then name slots: no per-call conversion.  Slots stay valid until overwritten or until
 * bt_matcher_reserve / bt_desc_pool_reserve re-size the pool. */
BT_API int bt_desc_pool_reserve(bt_ctx* ctx, int n_slots);
BT_API int bt_desc_pool_store(bt_ctx* ctx, int slot, const bt_desc_view* desc, void* stream);
BT_API int bt_knn_match_slots(bt_ctx* ctx, int n_pairs, const int32_t* slotA, const int32_t* slotB, int k,
                       int32_t* idxAB, float* distAB, int32_t* idxBA, float* distBA, void* stream);
/* Device time of the last kNN call: ms4 = {descriptor conversion, tensor-core pass, candidate selection + exact re-rank, exact
 * fallback}, info3 = {tensor-pass units (128 x 256 tiles), query rows, rows that needed the exact fallback}. */
BT_API int bt_knn_enable_timing(bt_ctx* ctx, int on);
/* Test knob: send every n-th query row through the exact brute-force fallback (0 = off).  Results do not change. */
BT_API int bt_knn_debug_force_fallback(bt_ctx* ctx, int every_nth);
BT_API int bt_knn_get_timing(bt_ctx* ctx, float* ms4, int* info3);

/* ------------------------------------------------------------------------------------------------------------
 * RANSAC.  Replaces ransacMultiPairGPU (/root/reference/src/cuda/cuda_ransac.cu:1228-1323; caller
 * /root/reference/src/FeatureManager.cpp:713).  ptsA/ptsB: HOST arrays of n_pairs DEVICE pointers to float4[n]
 * model-frame points (w ignored).  seed 0 => the reference's sampling sequence (XORWOW, curand_init(0, trial, 0)).
 * Winner = max inlier count, ties -> lowest trial id (the reference's findBestTrial is racy).
 * inlier_ids_out: DEVICE int32 [sum n_pts] (pair p's list starts at sum of n_pts over pairs < p);
 * n_inliers_out: DEVICE int32 [n_pairs]. */
BT_API int bt_ransac_reserve(bt_ctx* ctx, int max_pairs, int max_pts, int max_trials);
BT_API int bt_ransac_pairs(bt_ctx* ctx, int n_pairs, const float* const* ptsA, const float* const* ptsB, const int* n_pts,
                    int n_trials, float dist_thresh, uint64_t seed, int32_t* inlier_ids_out, int32_t* n_inliers_out,
                    void* stream);
/* Parity aid: the three uniforms per trial the sampler used (same values as curand_uniform after
 * curand_init(seed, trial, 0)) and the winning trial id of every pair of the last call. */
BT_API int bt_ransac_debug(bt_ctx* ctx, float* u3_out, int n_trials, int32_t* best_trial_out, int n_pairs);

/* ------------------------------------------------------------------------------------------------------------
 * Geometric pruning + "mutual" union + the fused matcher pipeline.  Replaces SiftManager::pruneMatches and
 * collectMutualMatches (/root/reference/src/FeatureManager.cpp:290-368) and, fused, everything findCorres does
 * between the descriptors and the EntryJ loop of Bundler::optimizeGPU (/root/reference/src/Bundler.cpp:298-324). */
typedef struct {
	const float* kpts_dev;     /* DEVICE float2[n]: cv::KeyPoint::pt (x = column, y = row), Frame::_keypts */
	int n;
	const float* depth_dev;    /* DEVICE float[H*W]:  Frame::_depth_gpu */
	const float* normal_dev;   /* DEVICE float4[H*W]: Frame::_normal_gpu */
	float pose[16];            /* Frame::_pose_in_model, row-major cam->model */
	int frame_id;              /* Frame::_id; |idA - idB| == 1 selects the *_neighbor thresholds */
	int window_index;          /* index of this frame in the sorted BA window (EntryJ::imgIdx_*) */
} bt_match_frame;

typedef struct {               /* feature_corres.* of config_*.yml, angles already as cosines (FeatureManager.cpp:292-295) */
	float max_dist_no_neighbor, cos_max_normal_no_neighbor, max_dist_neighbor, cos_max_normal_neighbor;
} bt_prune_params;

typedef struct {               /* the reference's `Correspondence` (uA,vA,uB,vB,ptA_cam,ptB_cam), 40 bytes */
	float uA, vA, uB, vB;
	float ptA_cam[3];
	float ptB_cam[3];
} bt_correspondence;

BT_API int bt_pipeline_reserve(bt_ctx* ctx, int max_pairs, int max_feats, int dim, int max_trials);
/* A[p] is the NEWER frame of pair p (frameA->_id > frameB->_id, FeatureManager.cpp:175).  idxAB/idxBA: DEVICE kNN indices
 * as produced by bt_knn_match_pairs for the same pairs.  corr_out: DEVICE, pair p's block starts at sum_{q<p}(nA_q+nB_q)
 * and holds n_corr_out[p] entries: survivors of A->B in query order, then those of B->A (duplicates kept). */
BT_API int bt_prune_mutual_pairs(bt_ctx* ctx, int n_pairs, const bt_match_frame* A, const bt_match_frame* B, int H, int W,
                          float fx, float fy, float cx, float cy, const int32_t* idxAB, const int32_t* idxBA, int k,
                          const bt_prune_params* prm, bt_correspondence* corr_out, int32_t* n_corr_out, void* stream);
/* kNN -> prune -> mutual -> RANSAC -> EntryJ without leaving the device.  entry_out: DEVICE, pairs back to back
 * (entry_off_out[p], n_entry_out[p]; *total_out entries in all).  Pairs with <= 5 pruned matches or < 5 RANSAC
 * inliers contribute nothing (FeatureManager.cpp:575-579,233-241).  The list never exceeds entry_capacity: a pair that does not
 * fit is cut and (offset, count) describe what was written, so they can be handed on as bt_window::block_off / block_n as they
 * are; entry_capacity = sum over pairs of (nA + nB) can never truncate. */
BT_API int bt_match_pairs(bt_ctx* ctx, int n_pairs, const bt_match_frame* A, const bt_match_frame* B,
                   const bt_desc_view* dA, const bt_desc_view* dB, int H, int W, float fx, float fy, float cx, float cy,
                   const bt_prune_params* prune, int ransac_trials, float ransac_inlier_dist, uint64_t ransac_seed,
                   bt_entryj* entry_out, int entry_capacity, int32_t* n_entry_out, int32_t* entry_off_out,
                   int32_t* total_out, void* stream);
/* The same chain on descriptor sets stored with bt_desc_pool_store (slotA[p] / slotB[p] instead of descriptor views). */
BT_API int bt_match_pairs_pool(bt_ctx* ctx, int n_pairs, const bt_match_frame* A, const bt_match_frame* B,
                        const int32_t* slotA, const int32_t* slotB, int H, int W, float fx, float fy, float cx, float cy,
                        const bt_prune_params* prune, int ransac_trials, float ransac_inlier_dist, uint64_t ransac_seed,
                        bt_entryj* entry_out, int entry_capacity, int32_t* n_entry_out, int32_t* entry_off_out,
                        int32_t* total_out, void* stream);

/* The rest of SiftManager::findCorres (/root/reference/src/FeatureManager.cpp:173-242) around the same chain:
 *  - extra[p]: the matches findCorresByMapPoints (:489-520) would APPEND for a NON-neighbour pair - bt_tracks_propagate's output for
 *    (A[p].frame_id, B[p].frame_id) called with n_existing = 0; the device drops the ones whose (uA,vA) or (uB,vB) already occurs among
 *    the pair's mutual matches, looks the 3-D points up at round(u), round(v) and appends the rest before RANSAC.  Ignored for
 *    neighbour pairs (|idA - idB| == 1), as in the reference.
 *  - status_out[p] (DEVICE, optional): BT_PAIR_OK (>= 5 inliers emitted), BT_PAIR_EMPTY (matches cleared, frame status unchanged) or
 *    BT_PAIR_FAIL (a neighbour pair ended with fewer than 5 matches: the reference marks frame A Frame::FAIL, :233-241,282-286).
 *  - uv_out (DEVICE float4 per emitted entry, optional): (uA, vA, uB, vB) of the entry - what updateFramePairMapPoints (:448-485) is fed
 *    through bt_tracks_update_pair. */
typedef struct { const float* uv; int n; } bt_match_extra;   /* HOST: n x (uA, vA, uB, vB) */
enum { BT_PAIR_OK = 0, BT_PAIR_EMPTY = 1, BT_PAIR_FAIL = 2 };
BT_API int bt_match_pairs_ex(bt_ctx* ctx, int n_pairs, const bt_match_frame* A, const bt_match_frame* B,
                      const bt_desc_view* dA, const bt_desc_view* dB, const int32_t* slotA, const int32_t* slotB,
                      int H, int W, float fx, float fy, float cx, float cy, const bt_prune_params* prune,
                      int ransac_trials, float ransac_inlier_dist, uint64_t ransac_seed, const bt_match_extra* extra,
                      bt_entryj* entry_out, int entry_capacity, int32_t* n_entry_out, int32_t* entry_off_out, int32_t* total_out,
                      int32_t* status_out, float* uv_out, void* stream);

/* Per-pair result cache: SiftManager::_matches (findCorres returns at once for a pair it has seen, FeatureManager.cpp:176;
 * forgetFrame drops a frame's pairs, :131-148).  Entries stay on the device; frames are named by Frame::_id.
 *   put     stores the blocks a bt_match_pairs* call produced (host copies of its n_entry / entry_off / status arrays).
 *   has     1 / 0.
 *   gather  writes the cached blocks of the listed pairs back to back into dst_dev with imgIdx_i / imgIdx_j rewritten to the given
 *           window indices (a frame's index changes from window to window) and fills block_off / block_n for bt_window::corr_dev;
 *           pairs with no cached matches contribute empty blocks. */
BT_API int bt_match_cache_reserve(bt_ctx* ctx, int max_pairs_cached, int max_entries_per_pair);
BT_API int bt_match_cache_put(bt_ctx* ctx, int n_pairs, const int32_t* idA, const int32_t* idB, const bt_entryj* entry_dev,
                       const int32_t* n_entry_host, const int32_t* entry_off_host, const int32_t* status_host, void* stream);
BT_API int bt_match_cache_has(bt_ctx* ctx, int idA, int idB);
BT_API int bt_match_cache_status(bt_ctx* ctx, int idA, int idB, int* n_entry, int* status);
BT_API int bt_match_cache_gather(bt_ctx* ctx, int n_pairs, const int32_t* idA, const int32_t* idB, const uint32_t* win_i, const uint32_t* win_j,
                          bt_entryj* dst_dev, int capacity, int32_t* block_off_host, int32_t* block_n_host, void* stream);
BT_API int bt_match_cache_forget_frame(bt_ctx* ctx, int frame_id);

/* ---- frame front end: Frame::processDepth + Frame::depthToCloudAndNormals -------------------------
 * /root/reference/src/Frame.cpp:152-233 -> CUDAImageUtil::erodeDepthMap, gaussFilterDepthMap x2,
 * convertDepthFloatToCameraSpaceFloat4, computeNormals (src/cuda/CUDAImageUtil.cu:676-806,310-336,342-423), fused into
 * one kernel per batch of frames.  Parameters are config_*.yml "depth_processing" (erode.radius/diff/ratio,
 * bilateral_filter.radius/sigma_D/sigma_R). */
typedef struct {
	int erode_radius; float erode_diff, erode_ratio;
	int bf_radius; float sigma_D, sigma_R;
} bt_depth_params;
/* All map pointers are DEVICE memory; the pointer ARRAYS are host memory.  depth_raw/depth_out: H*W float (metres);
 * xyz_out (may be NULL: not wanted) and normal_out: H*W float4 = 4 floats per pixel ((x,y,z,1) / (nx,ny,nz,0), zeros
 * where undefined) - depth_out and normal_out are exactly Frame::_depth_gpu / _normal_gpu as the solver reads them.
 * Out of place: depth_out[f] must differ from depth_raw[f]. */
BT_API int bt_frames_preprocess(bt_ctx* ctx, int n_frames, const float* const* depth_raw_dev, int H, int W,
                         float fx, float fy, float cx, float cy, const bt_depth_params* prm,
                         float* const* depth_out_dev, float* const* xyz_out_dev, float* const* normal_out_dev, void* stream);

/* Frame cache (second half).  The reference re-creates a CUDACache and re-downsamples every keyframe of
 * the window in EVERY optimizeFrames call (/root/reference/src/cuda/LossGPU.cu:74-101, CUDACache.cpp:76-88); a keyframe
 * takes part in many windows, so its quarter-resolution point/normal texel map and compacted source list are built here
 * once, when the keyframe enters the pool, and referenced by slot from bt_window::cache_slots.  The cache belongs to the
 * solver of this context: call bt_solver_reserve first.  Slots stay valid until overwritten.  depth_min/depth_max:
 * bt_solver_params::depth_min/max (they define which pixels enter the source list). */
BT_API int bt_frame_cache_reserve(bt_ctx* ctx, int capacity, int H, int W, float image_downscale);
BT_API int bt_frame_cache_store(bt_ctx* ctx, int n_frames, const int32_t* slots, const float* const* depth_dev, const float* const* normal_dev,
                         int H, int W, float fx, float fy, float cx, float cy, float depth_min, float depth_max, void* stream);

/* ---- host-side policy around the path: plain host code, callable without a GPU -----------------------------
 * Poses are row-major 4x4 cam->model.  keyframe_poses: n_keyframes consecutive 4x4 matrices. */
/* Utils::rotationGeodesicDistance (/root/reference/src/Utils.cpp:42-47): angle between the two rotations, radians. */
BT_API float bt_rotation_geodesic(const float* pose_a, const float* pose_b);
/* Bundler::checkAndAddKeyframe (/root/reference/src/Bundler.cpp:185-221): 1 = the frame becomes a keyframe (frame 0 always; otherwise
 * enough keypoints and at least min_rot_deg degrees away from EVERY existing keyframe), 0 = not.  The caller applies the
 * reference's `_status == OTHER` test. */
BT_API int bt_keyframe_check(const float* pose_new, int frame_id, int n_keypts, const float* keyframe_poses, int n_keyframes,
                      int min_feat_num, float min_rot_deg);
/* Bundler::selectKeyFramesForBA, method "greedy_rot" (Bundler.cpp:224-274): indices (ascending) of the keyframes that join the new
 * frame in the window; at most max_BA_frames - 1 of them.  chosen_out needs room for n_keyframes entries. */
BT_API int bt_select_keyframes(const float* pose_new, const float* keyframe_poses, int n_keyframes, int max_BA_frames,
                        int32_t* chosen_out, int* n_chosen_out);
/* Utils::solveRigidTransformBetweenPoints (Utils.cpp:180-214), the closed form behind SiftManager::procrustesByCorrespondence
 * (FeatureManager.cpp:523-557): pose_out (row-major 4x4) maps pts1 onto pts2 (n x 3 each) in the least-squares sense; identity when
 * the fit degenerates, as in the reference. */
BT_API int bt_rigid_transform(const float* pts1, const float* pts2, int n, float* pose_out);

/* The gate in front of the optimizer in Bundler::optimizeGPU (/root/reference/src/Bundler.cpp:343-347): bundle adjustment runs only when
 * the new frame takes part in MORE than min_fm_edges_newframe correspondences of the window; otherwise the frame's status becomes
 * Frame::NO_BA and the poses stay as they are.  n_edges_newframe = number of EntryJ records with the new frame on either side
 * (bt_match_cache_gather / FindCorres report it).  Returns 1 = run bt_solve_windows, 0 = NO_BA. */
BT_API int bt_ba_gate(int n_edges_newframe, int min_fm_edges_newframe);
/* Bundler::saveNewframeResult's pose record (/root/reference/src/Bundler.cpp:362-378): ob_in_cam = inverse(cur_in_model) as text, the 4 x 4
 * matrix printed the way `ofstream << std::setprecision(10) << Eigen::Matrix4f` prints it (10 significant digits, columns right-aligned
 * to the widest coefficient, one row per line).  bt_pose_format writes at most `cap` bytes incl. the terminating 0 into `text` and
 * returns BT_ERR_CAPACITY when it does not fit (1024 is always enough); bt_pose_write_txt writes `path` (the caller creates
 * <debug_dir>/poses/ as the reference does). */
BT_API int bt_pose_format(const float* cur_in_model, char* text, int cap);
BT_API int bt_pose_write_txt(const char* path, const float* cur_in_model);

/* Lfnet::detectFeature's reply handling (/root/reference/src/FeatureManager.cpp:876-907; rot_deg = 0, the only value the tracker uses):
 * validates the three message parts of the LF-Net server's reply (int32 (n, dim) | float32 n x 2 keypoints in the 400 x 400 network
 * input | float32 n x dim descriptors) and maps the keypoints back to image pixels through the crop / pad-to-square / resize
 * transform of `roi` = Frame::_roi (umin, umax, vmin, vmax).  kpts_out: n x 2 floats (x, y).  The descriptor part is already the
 * row-major matrix bt_knn_match_pairs takes: upload it as it is (bt_memcpy_h2d). */
BT_API int bt_lfnet_parse_reply(const void* info, size_t info_bytes, const void* kpts, size_t kpts_bytes, size_t desc_bytes, const int* roi,
                         float* kpts_out, int kpts_capacity, int* n_out, int* dim_out);

/* Map-point bookkeeping between matching and RANSAC (host code, no GPU):
 * SiftManager::updateFramePairMapPoints / findCorresByMapPoints / the map-point part of forgetFrame
 * (/root/reference/src/FeatureManager.cpp:448-520,163-169).  Frames are named by Frame::_id; uv arrays hold (uA, vA, uB, vB) per
 * match, A = the newer frame.  bt_tracks_propagate returns the matches findCorresByMapPoints would APPEND for the pair (map points
 * seen in both frames whose (uA,vA) and (uB,vB) are not matched yet), in the reference's order; the caller looks the 3-D points up
 * at round(u), round(v) of its point maps exactly as the reference does and hands them to RANSAC with the rest. */
typedef struct bt_tracks bt_tracks;
BT_API int bt_tracks_create(bt_tracks** out);
BT_API void bt_tracks_destroy(bt_tracks* t);
BT_API int bt_tracks_update_pair(bt_tracks* t, int frame_a, int frame_b, const float* uv, const unsigned char* is_inlier, int n);
BT_API int bt_tracks_propagate(bt_tracks* t, int frame_a, int frame_b, const float* existing_uv, int n_existing,
                        float* out_uv, int capacity, int* n_out);
BT_API int bt_tracks_forget_frame(bt_tracks* t, int frame);
BT_API int bt_tracks_stats(const bt_tracks* t, int* n_points, int* n_observations);

/* Small device-memory helpers so non-CUDA hosts (ctypes, cgo, JNI) can drive the library without another runtime. */
BT_API int bt_dev_alloc(void** out, size_t bytes);
BT_API int bt_dev_free(void* p);
BT_API int bt_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes, void* stream);
BT_API int bt_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes, void* stream);
BT_API int bt_host_alloc_pinned(void** out, size_t bytes);
BT_API int bt_host_free_pinned(void* p);
BT_API int bt_stream_sync(void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BUNDLETRACK_B200_H */
