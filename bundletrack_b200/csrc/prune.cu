// prune.cu — geometric pruning of the kNN candidates, the "mutual" union, and the glue that turns RANSAC inliers into
// EntryJ records — everything SiftManager::findCorres does between the two knnMatch calls and Bundler::optimizeGPU's
// EntryJ loop, kept on the device:
//
//   k_prune_mutual   one CTA per frame pair.  Direction A->B then B->A: for every query feature the FIRST of its k
//                    neighbours passing the gates of SiftManager::pruneMatches (/root/reference/src/FeatureManager.cpp:
//                    290-336: both keypoints round()ed inside the image, both depths >= 0.1, model-frame distance <=
//                    max_dist_*, cos(angle between model-frame normals) >= cos(max_normal_*), thresholds chosen by
//                    |idA-idB|==1) is kept; survivors are appended in query order, A->B block first, duplicates NOT
//                    removed — exactly collectMutualMatches (:338-368, a union).  Also emits the
//                    model-frame points RANSAC consumes (runRansacMultiPairGPU :676-706).
//   k_ransac_table_from_counts / k_emit_entryj
//                    counts -> RANSAC pair table (<= 5 candidates => pair dropped, runRansacBetween :575-579);
//                    inliers -> EntryJ {imgIdx_i = older frame, imgIdx_j = newer frame, pos_i = ptB_cam, pos_j = ptA_cam}
//                    (Bundler::optimizeGPU, /root/reference/src/Bundler.cpp:308-317), < 5 inliers => pair dropped
//                    (FeatureManager.cpp:233-241,727-730); pairs written back to back, which is the grouping the
//                    solver's tail wants.
#include <map>
#include <vector>
#include "bt_common.cuh"

namespace bt {

struct RansacPair { const float4* A; const float4* B; int n; int out_off; };   // same layout as in ransac.cu
int ransac_run_device(bt_ctx* ctx, const RansacPair* d_pairs, int n_pairs, int n_trials, float dist_thresh, uint64_t seed,
                      int32_t* inlier_ids_out, int32_t* n_inliers_out, cudaStream_t stream);

struct PruneFrame { const float2* kpts; int n; int frame_id; const float* depth; const float4* normal; float T[12]; };
struct PrunePair { PruneFrame A, B; int idx_off_A, idx_off_B; int out_off; int win_idx_A, win_idx_B; int n_extra, extra_off; };   // extras: propagated (uA,vA,uB,vB) matches, non-neighbour pairs only
struct PruneCam { int H, W; float ifx, ify, icx, icy; };

// One candidate (query q of frame Q, train t of frame Tn).  Returns true when it passes pruneMatches' gates.
__device__ __forceinline__ bool prune_gate(const PruneFrame& Q, const PruneFrame& Tn, int q, int t, const PruneCam& cam, float max_dist, float cos_max,
                                           float (&ptQ)[3], float (&ptT)[3], float (&mQ)[3], float (&mT)[3]) {
	const float2 pq = __ldg(Q.kpts + q), pt = __ldg(Tn.kpts + t);
	const int uq = (int)roundf(pq.x), vq = (int)roundf(pq.y), ut = (int)roundf(pt.x), vt = (int)roundf(pt.y);   // std::round
	if (uq < 0 || vq < 0 || uq >= cam.W || vq >= cam.H || ut < 0 || vt < 0 || ut >= cam.W || vt >= cam.H) return false;   // Utils::isPixelInsideImage
	const float dq = __ldg(Q.depth + (size_t)vq * cam.W + uq), dt = __ldg(Tn.depth + (size_t)vt * cam.W + ut);
	if (dq < 0.1f || dt < 0.1f) return false;                    // organised cloud point: z < 0.1 (zeros when depth < 0.1)
	// Frame::depthToCloudAndNormals: K^-1 (u d, v d, d)   (/root/reference/src/Frame.cpp:199-233, CUDAImageUtil.cu:310-326)
	ptQ[0] = cam.ifx * ((float)uq * dq) + cam.icx * dq; ptQ[1] = cam.ify * ((float)vq * dq) + cam.icy * dq; ptQ[2] = dq;
	ptT[0] = cam.ifx * ((float)ut * dt) + cam.icx * dt; ptT[1] = cam.ify * ((float)vt * dt) + cam.icy * dt; ptT[2] = dt;
	const float4 nq = __ldg(Q.normal + (size_t)vq * cam.W + uq), nt = __ldg(Tn.normal + (size_t)vt * cam.W + ut);
	float wq[3], wt[3];
#pragma unroll
	for (int r = 0; r < 3; r++) {   // pcl::transformPointWithNormal
		mQ[r] = Q.T[r * 4] * ptQ[0] + Q.T[r * 4 + 1] * ptQ[1] + Q.T[r * 4 + 2] * ptQ[2] + Q.T[r * 4 + 3];
		mT[r] = Tn.T[r * 4] * ptT[0] + Tn.T[r * 4 + 1] * ptT[1] + Tn.T[r * 4 + 2] * ptT[2] + Tn.T[r * 4 + 3];
		wq[r] = Q.T[r * 4] * nq.x + Q.T[r * 4 + 1] * nq.y + Q.T[r * 4 + 2] * nq.z;
		wt[r] = Tn.T[r * 4] * nt.x + Tn.T[r * 4 + 1] * nt.y + Tn.T[r * 4 + 2] * nt.z;
	}
	const float dx = mQ[0] - mT[0], dy = mQ[1] - mT[1], dz = mQ[2] - mT[2];
	const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
	// n1.normalized().dot(n2.normalized()): a zero normal normalises to NaN, the `<` test is then false and the
	// candidate is NOT rejected (FeatureManager.cpp:326) — reproduced by the !(dot < cos) form below.
	const float lq = sqrtf(wq[0] * wq[0] + wq[1] * wq[1] + wq[2] * wq[2]), lt = sqrtf(wt[0] * wt[0] + wt[1] * wt[1] + wt[2] * wt[2]);
	const float dotn = (wq[0] / lq) * (wt[0] / lt) + (wq[1] / lq) * (wt[1] / lt) + (wq[2] / lq) * (wt[2] / lt);
	if (dist > max_dist || dotn < cos_max) return false;
	return true;
}

__global__ void __launch_bounds__(512) k_prune_mutual(const PrunePair* __restrict__ pairs, PruneCam cam, const int32_t* __restrict__ idxAB, const int32_t* __restrict__ idxBA,
                                                       int k, bt_prune_params prm, bt_correspondence* __restrict__ corr, float4* __restrict__ PA, float4* __restrict__ PB,
                                                       int32_t* __restrict__ n_corr) {
	__shared__ int s_warp[16];
	__shared__ int s_base;
	const PrunePair pp = pairs[blockIdx.x];
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const bool neighbor = abs(pp.A.frame_id - pp.B.frame_id) == 1;
	const float max_dist = neighbor ? prm.max_dist_neighbor : prm.max_dist_no_neighbor;
	const float cos_max = neighbor ? prm.cos_max_normal_neighbor : prm.cos_max_normal_no_neighbor;
	if (tid == 0) s_base = 0;
	__syncthreads();
	for (int dir = 0; dir < 2; dir++) {
		const PruneFrame& Q = dir == 0 ? pp.A : pp.B;
		const PruneFrame& Tn = dir == 0 ? pp.B : pp.A;
		const int32_t* idx = (dir == 0 ? idxAB + (size_t)pp.idx_off_A * k : idxBA + (size_t)pp.idx_off_B * k);
		for (int base = 0; base < Q.n; base += blockDim.x) {
			const int q = base + tid;
			bool keep = false;
			float ptQ[3], ptT[3], mQ[3], mT[3];
			int tsel = -1;
			if (q < Q.n && Tn.n > 0) {
				for (int c = 0; c < k; c++) {
					const int t = __ldg(idx + (size_t)q * k + c);
					if (t < 0 || t >= Tn.n) continue;
					if (prune_gate(Q, Tn, q, t, cam, max_dist, cos_max, ptQ, ptT, mQ, mT)) { keep = true; tsel = t; break; }
				}
			}
			const unsigned bal = __ballot_sync(0xffffffffu, keep);
			if (lane == 0) s_warp[wid] = __popc(bal);
			__syncthreads();
			int off = s_base;
			for (int w = 0; w < wid; w++) off += s_warp[w];
			if (keep) {
				const int o = pp.out_off + off + __popc(bal & ((1u << lane) - 1u));
				const float2 kq = __ldg(Q.kpts + q), kt = __ldg(Tn.kpts + tsel);
				bt_correspondence c;
				// Correspondence(uA, vA, uB, vB, ptA, ptB): A is always the pair's first (newer) frame
				if (dir == 0) { c.uA = kq.x; c.vA = kq.y; c.uB = kt.x; c.vB = kt.y; for (int r = 0; r < 3; r++) { c.ptA_cam[r] = ptQ[r]; c.ptB_cam[r] = ptT[r]; } PA[o] = make_float4(mQ[0], mQ[1], mQ[2], 1.f); PB[o] = make_float4(mT[0], mT[1], mT[2], 1.f); }
				else { c.uA = kt.x; c.vA = kt.y; c.uB = kq.x; c.vB = kq.y; for (int r = 0; r < 3; r++) { c.ptA_cam[r] = ptT[r]; c.ptB_cam[r] = ptQ[r]; } PA[o] = make_float4(mT[0], mT[1], mT[2], 1.f); PB[o] = make_float4(mQ[0], mQ[1], mQ[2], 1.f); }
				corr[o] = c;
			}
			__syncthreads();
			if (tid == 0) { int t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_warp[w]; s_base += t; }
			__syncthreads();
		}
	}
	if (tid == 0) n_corr[blockIdx.x] = s_base;
}

// findCorresByMapPoints (/root/reference/src/FeatureManager.cpp:489-520) on the device, between the mutual union and RANSAC, for the
// NON-neighbour pairs: the caller hands the (uA, vA, uB, vB) of the map points both frames observe (bt_tracks_propagate, in the
// reference's order); a candidate is dropped when an existing match of the pair has the same (uA, vA) or the same (uB, vB), the rest
// are appended in order with the organised-cloud points at round(u), round(v) (zeros where the depth is < 0.1, as the cloud holds them).
__global__ void __launch_bounds__(256) k_append_propagated(const PrunePair* __restrict__ pairs, PruneCam cam, const float4* __restrict__ extra_uv,
                                                            bt_correspondence* __restrict__ corr, float4* __restrict__ PA, float4* __restrict__ PB,
                                                            int32_t* __restrict__ n_corr, int32_t* __restrict__ n_nn) {
	__shared__ int s_warp[8];
	__shared__ int s_base;
	const PrunePair pp = pairs[blockIdx.x];
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int n0 = n_corr[blockIdx.x];
	if (tid == 0) { n_nn[blockIdx.x] = n0; s_base = n0; }
	__syncthreads();
	if (pp.n_extra <= 0 || abs(pp.A.frame_id - pp.B.frame_id) == 1) return;
	const bt_correspondence* ex = corr + pp.out_off;
	for (int base = 0; base < pp.n_extra; base += blockDim.x) {
		const int e = base + tid;
		bool keep = false;
		float4 uv = make_float4(0.f, 0.f, 0.f, 0.f);
		if (e < pp.n_extra) {
			uv = __ldg(extra_uv + pp.extra_off + e);
			keep = true;
			for (int i = 0; i < n0; i++) {      // only the matches of the two knnMatch directions are compared: map-point keys are unique per frame
				const bt_correspondence c = ex[i];
				if ((c.uA == uv.x && c.vA == uv.y) || (c.uB == uv.z && c.vB == uv.w)) { keep = false; break; }
			}
		}
		const unsigned bal = __ballot_sync(0xffffffffu, keep);
		if (lane == 0) s_warp[wid] = __popc(bal);
		__syncthreads();
		int off = s_base;
		for (int w = 0; w < wid; w++) off += s_warp[w];
		if (keep) {
			const int o = pp.out_off + off + __popc(bal & ((1u << lane) - 1u));
			bt_correspondence c;
			c.uA = uv.x; c.vA = uv.y; c.uB = uv.z; c.vB = uv.w;
			float pa[3] = { 0.f, 0.f, 0.f }, pb[3] = { 0.f, 0.f, 0.f };
			const int ua = (int)roundf(uv.x), va = (int)roundf(uv.y), ub = (int)roundf(uv.z), vb = (int)roundf(uv.w);
			if (ua >= 0 && va >= 0 && ua < cam.W && va < cam.H) { const float d = __ldg(pp.A.depth + (size_t)va * cam.W + ua); if (d >= 0.1f) { pa[0] = cam.ifx * ((float)ua * d) + cam.icx * d; pa[1] = cam.ify * ((float)va * d) + cam.icy * d; pa[2] = d; } }
			if (ub >= 0 && vb >= 0 && ub < cam.W && vb < cam.H) { const float d = __ldg(pp.B.depth + (size_t)vb * cam.W + ub); if (d >= 0.1f) { pb[0] = cam.ifx * ((float)ub * d) + cam.icx * d; pb[1] = cam.ify * ((float)vb * d) + cam.icy * d; pb[2] = d; } }
			float ma[3], mb[3];
#pragma unroll
			for (int r = 0; r < 3; r++) {
				c.ptA_cam[r] = pa[r]; c.ptB_cam[r] = pb[r];
				ma[r] = pp.A.T[r * 4] * pa[0] + pp.A.T[r * 4 + 1] * pa[1] + pp.A.T[r * 4 + 2] * pa[2] + pp.A.T[r * 4 + 3];
				mb[r] = pp.B.T[r * 4] * pb[0] + pp.B.T[r * 4 + 1] * pb[1] + pp.B.T[r * 4 + 2] * pb[2] + pp.B.T[r * 4 + 3];
			}
			corr[o] = c; PA[o] = make_float4(ma[0], ma[1], ma[2], 1.f); PB[o] = make_float4(mb[0], mb[1], mb[2], 1.f);
		}
		__syncthreads();
		if (tid == 0) { int t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_warp[w]; s_base += t; }
		__syncthreads();
	}
	if (tid == 0) n_corr[blockIdx.x] = s_base;
}

__global__ void k_ransac_table_from_counts(const PrunePair* __restrict__ pairs, int n_pairs, const int32_t* __restrict__ n_corr, const float4* PA, const float4* PB, RansacPair* out) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n_pairs) return;
	const int n = n_corr[p];
	RansacPair r; r.A = PA + pairs[p].out_off; r.B = PB + pairs[p].out_off;
	r.n = (n <= 5) ? 0 : n;                 // countInlierCorres <= 5 => matches cleared before RANSAC
	r.out_off = pairs[p].out_off;
	out[p] = r;
}

// one CTA: exclusive scan of the kept inlier counts over pairs (pairs stay contiguous in the EntryJ list)
__global__ void __launch_bounds__(1024) k_emit_entryj(const PrunePair* __restrict__ pairs, int n_pairs, const bt_correspondence* __restrict__ corr,
                                                       const int32_t* __restrict__ inlier_ids, const int32_t* __restrict__ n_inliers,
                                                       bt_entryj* __restrict__ entry_out, int32_t* __restrict__ n_entry_out, int32_t* __restrict__ entry_off_out, int32_t* total_out, int capacity,
                                                       int32_t* __restrict__ status_out) {
	__shared__ int s_scan[1024];
	__shared__ int s_carry;
	const int tid = threadIdx.x;
	if (tid == 0) s_carry = 0;
	__syncthreads();
	for (int base = 0; base < n_pairs; base += 1024) {
		const int p = base + tid;
		int cnt = 0;
		if (p < n_pairs) { cnt = n_inliers[p]; if (cnt < 5) cnt = 0; }     // fewer than 5 matches => cleared
		s_scan[tid] = cnt;
		__syncthreads();
		for (int o = 1; o < 1024; o <<= 1) { const int v = (tid >= o) ? s_scan[tid - o] : 0; __syncthreads(); s_scan[tid] += v; __syncthreads(); }
		int excl = s_scan[tid] - cnt + s_carry;
		// never describe more than the caller's buffer holds: a pair that does not fit is cut (or dropped), so the (offset, count)
		// arrays a caller hands on as bt_window::block_off / block_n always stay inside entry_out
		if (excl > capacity) excl = capacity;
		if (excl + cnt > capacity) cnt = capacity - excl;
		if (p < n_pairs) {
			n_entry_out[p] = cnt; entry_off_out[p] = excl;
			// SiftManager::findCorres' outcome for the pair (FeatureManager.cpp:186-241,282-286): a pair left with fewer than 5 matches is
			// cleared, and when the two frames are neighbours (|idA - idB| == 1) the newer frame is marked Frame::FAIL
			if (status_out) status_out[p] = cnt > 0 ? BT_PAIR_OK : (abs(pairs[p].A.frame_id - pairs[p].B.frame_id) == 1 ? BT_PAIR_FAIL : BT_PAIR_EMPTY);
		}
		__syncthreads();
		if (tid == 1023) s_carry += s_scan[1023];
		__syncthreads();
	}
	if (tid == 0) *total_out = min(s_carry, capacity);
}

// one CTA per pair: its EntryJ block (a lone CTA filling all pairs took 0.1 ms for 20 k entries)
__global__ void __launch_bounds__(256) k_emit_fill(const PrunePair* __restrict__ pairs, const bt_correspondence* __restrict__ corr, const int32_t* __restrict__ inlier_ids,
                                                    const int32_t* __restrict__ n_entry, const int32_t* __restrict__ entry_off, bt_entryj* __restrict__ entry_out, int capacity,
                                                    float4* __restrict__ uv_out) {
	const int p = blockIdx.x;
	const int cnt = n_entry[p], off = entry_off[p];
	const PrunePair pp = pairs[p];
	for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
		if (off + i >= capacity) break;
		const bt_correspondence c = corr[pp.out_off + inlier_ids[pp.out_off + i]];
		bt_entryj e;
		e.imgIdx_i = (uint32_t)pp.win_idx_B; e.imgIdx_j = (uint32_t)pp.win_idx_A;          // i = older frame B, j = newer frame A
		for (int r = 0; r < 3; r++) { e.pos_i[r] = c.ptB_cam[r]; e.pos_j[r] = c.ptA_cam[r]; }
		entry_out[off + i] = e;
		if (uv_out) uv_out[off + i] = make_float4(c.uA, c.vA, c.uB, c.vB);
	}
}

struct PruneState {
	int max_pairs = 0, max_feats = 0;
	DevBuf pairs, corr, PA, PB, n_corr, n_nn, rpairs, inl, n_inl, idxAB, distAB, idxBA, distBA, entry_off, total, extra;
	PinnedBuf h_pairs, h_extra;
	size_t extra_cap = 0;               // propagated matches (float4 each) one call may carry
	cudaEvent_t ev_up = nullptr;      // recorded behind the upload of h_pairs (the calls are asynchronous: see fill_pairs)
};
static PruneState* g_prune_of(bt_ctx* ctx);

}  // namespace bt

using namespace bt;

// The prune state is owned by the context like the solver / matcher / RANSAC / front-end states.
namespace bt {
static PruneState* g_prune_of(bt_ctx* ctx) { return ctx->prune; }
void prune_destroy(bt_ctx* ctx) {
	PruneState* s = ctx->prune;
	if (!s) return;
	DevBuf* bufs[] = { &s->pairs, &s->corr, &s->PA, &s->PB, &s->n_corr, &s->n_nn, &s->extra, &s->rpairs, &s->inl, &s->n_inl, &s->idxAB, &s->distAB, &s->idxBA, &s->distBA, &s->entry_off, &s->total };
	for (DevBuf* b : bufs) b->release();
	s->h_pairs.release(); s->h_extra.release();
	if (s->ev_up) cudaEventDestroy(s->ev_up);
	delete s;
	ctx->prune = nullptr;
}
}

extern "C" int bt_pipeline_reserve(bt_ctx* ctx, int max_pairs, int max_feats, int dim, int max_trials) {
	BT_REQUIRE(ctx && max_pairs > 0 && max_feats > 0, BT_ERR_INVALID_ARG, "bt_pipeline_reserve: bad arguments");
	int rc = bt_matcher_reserve(ctx, max_pairs, max_feats, dim);
	if (rc != BT_OK) return rc;
	rc = bt_ransac_reserve(ctx, max_pairs, 3 * max_feats, max_trials);
	if (rc != BT_OK) return rc;
	PruneState* s = g_prune_of(ctx);
	if (!s) { s = new PruneState(); ctx->prune = s; }
	s->max_pairs = max_pairs; s->max_feats = max_feats;
	const size_t cap = (size_t)3 * max_feats * max_pairs;      // nA + nB mutual candidates + up to max_feats propagated map-point matches per pair
	s->extra_cap = (size_t)max_feats * max_pairs;
#define RES(buf, bytes) if ((rc = s->buf.alloc(bytes)) != BT_OK) return rc
	RES(pairs, sizeof(PrunePair) * max_pairs);
	RES(corr, sizeof(bt_correspondence) * cap);
	RES(PA, sizeof(float4) * cap); RES(PB, sizeof(float4) * cap);
	RES(n_corr, sizeof(int32_t) * max_pairs); RES(n_nn, sizeof(int32_t) * max_pairs);
	RES(extra, sizeof(float4) * s->extra_cap);
	RES(rpairs, sizeof(RansacPair) * max_pairs);
	RES(inl, sizeof(int32_t) * cap); RES(n_inl, sizeof(int32_t) * max_pairs);
	RES(idxAB, sizeof(int32_t) * 8 * (size_t)max_feats * max_pairs); RES(distAB, sizeof(float) * 8 * (size_t)max_feats * max_pairs);
	RES(idxBA, sizeof(int32_t) * 8 * (size_t)max_feats * max_pairs); RES(distBA, sizeof(float) * 8 * (size_t)max_feats * max_pairs);
	RES(entry_off, sizeof(int32_t) * max_pairs); RES(total, 16);
#undef RES
	if ((rc = s->h_extra.alloc(sizeof(float4) * s->extra_cap)) != BT_OK) return rc;
	return s->h_pairs.alloc(sizeof(PrunePair) * max_pairs);
}

static int fill_pairs(PruneState* s, int n_pairs, const bt_match_frame* A, const bt_match_frame* B, cudaStream_t stream, const bt_match_extra* extra = nullptr) {
	// an earlier asynchronous call's upload may still be reading the pinned table: wait for it before rewriting the block
	if (!s->ev_up) BT_CUDA(cudaEventCreateWithFlags(&s->ev_up, cudaEventDisableTiming));
	else BT_CUDA(cudaEventSynchronize(s->ev_up));
	PrunePair* hp = s->h_pairs.as<PrunePair>();
	int offA = 0, offB = 0, out = 0, xoff = 0;
	float4* hx = s->h_extra.as<float4>();
	for (int p = 0; p < n_pairs; p++) {
		const bt_match_frame* fr[2] = { &A[p], &B[p] };
		PruneFrame* dst[2] = { &hp[p].A, &hp[p].B };
		for (int s2 = 0; s2 < 2; s2++) {
			BT_REQUIRE(fr[s2]->n >= 0 && fr[s2]->n <= s->max_feats, BT_ERR_CAPACITY, "pair %d: %d features > reserved %d", p, fr[s2]->n, s->max_feats);
			BT_REQUIRE(fr[s2]->n == 0 || (fr[s2]->kpts_dev && fr[s2]->depth_dev && fr[s2]->normal_dev), BT_ERR_INVALID_ARG, "pair %d: NULL frame pointer", p);
			dst[s2]->kpts = reinterpret_cast<const float2*>(fr[s2]->kpts_dev); dst[s2]->n = fr[s2]->n; dst[s2]->frame_id = fr[s2]->frame_id;
			dst[s2]->depth = fr[s2]->depth_dev; dst[s2]->normal = reinterpret_cast<const float4*>(fr[s2]->normal_dev);
			memcpy(dst[s2]->T, fr[s2]->pose, sizeof(float) * 12);
		}
		hp[p].idx_off_A = offA; hp[p].idx_off_B = offB; hp[p].out_off = out;
		hp[p].win_idx_A = A[p].window_index; hp[p].win_idx_B = B[p].window_index;
		const int nx = (extra && extra[p].n > 0 && abs(A[p].frame_id - B[p].frame_id) != 1) ? extra[p].n : 0;
		BT_REQUIRE(nx == 0 || extra[p].uv, BT_ERR_INVALID_ARG, "pair %d: NULL propagated-match array", p);
		BT_REQUIRE(nx <= s->max_feats && (size_t)(xoff + nx) <= s->extra_cap, BT_ERR_CAPACITY, "pair %d: %d propagated matches exceed the reserved room", p, nx);
		hp[p].n_extra = nx; hp[p].extra_off = xoff;
		if (nx) memcpy(hx + xoff, extra[p].uv, sizeof(float4) * (size_t)nx);
		xoff += nx;
		offA += A[p].n; offB += B[p].n; out += A[p].n + B[p].n + nx;
	}
	if (xoff) BT_CUDA(cudaMemcpyAsync(s->extra.p, hx, sizeof(float4) * (size_t)xoff, cudaMemcpyHostToDevice, stream));
	BT_CUDA(cudaMemcpyAsync(s->pairs.p, hp, sizeof(PrunePair) * n_pairs, cudaMemcpyHostToDevice, stream));
	BT_CUDA(cudaEventRecord(s->ev_up, stream));
	return BT_OK;
}

static int prune_mutual_impl(bt_ctx* ctx, int n_pairs, const bt_match_frame* A, const bt_match_frame* B, int H, int W, float fx, float fy, float cx, float cy,
                             const int32_t* idxAB, const int32_t* idxBA, int k, const bt_prune_params* prm,
                             bt_correspondence* corr_out, int32_t* n_corr_out, void* stream_, const bt_match_extra* extra) {
	PruneState* s = ctx ? g_prune_of(ctx) : nullptr;
	BT_REQUIRE(s, BT_ERR_INVALID_ARG, "bt_prune_mutual_pairs: call bt_pipeline_reserve first");
	BT_REQUIRE(A && B && idxAB && idxBA && prm && corr_out && n_corr_out && n_pairs > 0 && n_pairs <= s->max_pairs && k >= 1 && k <= 8, BT_ERR_INVALID_ARG, "bt_prune_mutual_pairs: bad argument");
	cudaStream_t stream = (cudaStream_t)stream_;
	BT_CUDA(cudaSetDevice(ctx->device));
	int rc = fill_pairs(s, n_pairs, A, B, stream, extra);
	if (rc != BT_OK) return rc;
	PruneCam cam; cam.H = H; cam.W = W; cam.ifx = 1.0f / fx; cam.ify = 1.0f / fy; cam.icx = -cx / fx; cam.icy = -cy / fy;
	k_prune_mutual<<<n_pairs, 512, 0, stream>>>(s->pairs.as<PrunePair>(), cam, idxAB, idxBA, k, *prm, corr_out, s->PA.as<float4>(), s->PB.as<float4>(), n_corr_out);
	// findCorresByMapPoints for the non-neighbour pairs; also records the matches the two knnMatch directions gave (n_nn)
	k_append_propagated<<<n_pairs, 256, 0, stream>>>(s->pairs.as<PrunePair>(), cam, s->extra.as<float4>(), corr_out, s->PA.as<float4>(), s->PB.as<float4>(), n_corr_out, s->n_nn.as<int32_t>());
	BT_CUDA(cudaGetLastError());
	return BT_OK;
}

extern "C" int bt_prune_mutual_pairs(bt_ctx* ctx, int n_pairs, const bt_match_frame* A, const bt_match_frame* B, int H, int W, float fx, float fy, float cx, float cy,
                                     const int32_t* idxAB, const int32_t* idxBA, int k, const bt_prune_params* prm,
                                     bt_correspondence* corr_out, int32_t* n_corr_out, void* stream_) {
	return prune_mutual_impl(ctx, n_pairs, A, B, H, W, fx, fy, cx, cy, idxAB, idxBA, k, prm, corr_out, n_corr_out, stream_, nullptr);
}

// The whole matcher half of the hot path for a batch of frame pairs, device-resident end to end:
// kNN (both directions) -> prune -> mutual union -> RANSAC -> EntryJ.  Pair p's entries are contiguous in entry_out.
static int match_pairs_impl(bt_ctx* ctx, int n_pairs, const bt_match_frame* A, const bt_match_frame* B, const bt_desc_view* dA, const bt_desc_view* dB,
                            const int32_t* slotA, const int32_t* slotB, int H, int W, float fx, float fy, float cx, float cy, const bt_prune_params* prune, int ransac_trials,
                            float ransac_inlier_dist, uint64_t ransac_seed, bt_entryj* entry_out, int entry_capacity, int32_t* n_entry_out, int32_t* entry_off_out, int32_t* total_out, void* stream_,
                            const bt_match_extra* extra = nullptr, int32_t* status_out = nullptr, float* uv_out = nullptr) {
	PruneState* s = ctx ? g_prune_of(ctx) : nullptr;
	BT_REQUIRE(s, BT_ERR_INVALID_ARG, "bt_match_pairs: call bt_pipeline_reserve first");
	BT_REQUIRE(A && B && ((dA && dB) || (slotA && slotB)) && prune && entry_out && n_entry_out && entry_off_out && total_out && n_pairs > 0 && n_pairs <= s->max_pairs, BT_ERR_INVALID_ARG, "bt_match_pairs: bad argument");
	cudaStream_t stream = (cudaStream_t)stream_;
	BT_CUDA(cudaSetDevice(ctx->device));
	const int k = 5;   // k_near, FeatureManager.cpp:264
	int rc = dA ? bt_knn_match_pairs(ctx, n_pairs, dA, dB, k, s->idxAB.as<int32_t>(), s->distAB.as<float>(), s->idxBA.as<int32_t>(), s->distBA.as<float>(), stream_)
	            : bt_knn_match_slots(ctx, n_pairs, slotA, slotB, k, s->idxAB.as<int32_t>(), s->distAB.as<float>(), s->idxBA.as<int32_t>(), s->distBA.as<float>(), stream_);
	if (rc != BT_OK) return rc;
	rc = prune_mutual_impl(ctx, n_pairs, A, B, H, W, fx, fy, cx, cy, s->idxAB.as<int32_t>(), s->idxBA.as<int32_t>(), k, prune, s->corr.as<bt_correspondence>(), s->n_corr.as<int32_t>(), stream_, extra);
	if (rc != BT_OK) return rc;
	k_ransac_table_from_counts<<<(n_pairs + 127) / 128, 128, 0, stream>>>(s->pairs.as<PrunePair>(), n_pairs, s->n_corr.as<int32_t>(), s->PA.as<float4>(), s->PB.as<float4>(), s->rpairs.as<RansacPair>());
	rc = ransac_run_device(ctx, s->rpairs.as<RansacPair>(), n_pairs, ransac_trials, ransac_inlier_dist, ransac_seed, s->inl.as<int32_t>(), s->n_inl.as<int32_t>(), stream);
	if (rc != BT_OK) return rc;
	k_emit_entryj<<<1, 1024, 0, stream>>>(s->pairs.as<PrunePair>(), n_pairs, s->corr.as<bt_correspondence>(), s->inl.as<int32_t>(), s->n_inl.as<int32_t>(), entry_out, n_entry_out, entry_off_out, total_out, entry_capacity, status_out);
	k_emit_fill<<<n_pairs, 256, 0, stream>>>(s->pairs.as<PrunePair>(), s->corr.as<bt_correspondence>(), s->inl.as<int32_t>(), n_entry_out, entry_off_out, entry_out, entry_capacity,
	                                        reinterpret_cast<float4*>(uv_out));
	BT_CUDA(cudaGetLastError());
	return BT_OK;
}

extern "C" int bt_match_pairs(bt_ctx* ctx, int n_pairs, const bt_match_frame* A, const bt_match_frame* B, const bt_desc_view* dA, const bt_desc_view* dB,
                              int H, int W, float fx, float fy, float cx, float cy, const bt_prune_params* prune, int ransac_trials, float ransac_inlier_dist,
                              uint64_t ransac_seed, bt_entryj* entry_out, int entry_capacity, int32_t* n_entry_out, int32_t* entry_off_out, int32_t* total_out, void* stream_) {
	BT_REQUIRE(dA && dB, BT_ERR_INVALID_ARG, "bt_match_pairs: NULL descriptor views");
	return match_pairs_impl(ctx, n_pairs, A, B, dA, dB, nullptr, nullptr, H, W, fx, fy, cx, cy, prune, ransac_trials, ransac_inlier_dist, ransac_seed, entry_out, entry_capacity, n_entry_out, entry_off_out, total_out, stream_);
}

extern "C" int bt_match_pairs_pool(bt_ctx* ctx, int n_pairs, const bt_match_frame* A, const bt_match_frame* B, const int32_t* slotA, const int32_t* slotB,
                                   int H, int W, float fx, float fy, float cx, float cy, const bt_prune_params* prune, int ransac_trials, float ransac_inlier_dist,
                                   uint64_t ransac_seed, bt_entryj* entry_out, int entry_capacity, int32_t* n_entry_out, int32_t* entry_off_out, int32_t* total_out, void* stream_) {
	BT_REQUIRE(slotA && slotB, BT_ERR_INVALID_ARG, "bt_match_pairs_pool: NULL slot arrays");
	return match_pairs_impl(ctx, n_pairs, A, B, nullptr, nullptr, slotA, slotB, H, W, fx, fy, cx, cy, prune, ransac_trials, ransac_inlier_dist, ransac_seed, entry_out, entry_capacity, n_entry_out, entry_off_out, total_out, stream_);
}

// bt_match_pairs / bt_match_pairs_pool with the rest of SiftManager::findCorres (FeatureManager.cpp:173-242): propagated map-point
// matches for the non-neighbour pairs, the pair status (OK / cleared / Frame::FAIL) and the (uA, vA, uB, vB) of every emitted entry
// (what updateFramePairMapPoints needs).  Exactly one of (dA, dB) and (slotA, slotB) is given.
extern "C" int bt_match_pairs_ex(bt_ctx* ctx, int n_pairs, const bt_match_frame* A, const bt_match_frame* B, const bt_desc_view* dA, const bt_desc_view* dB,
                                 const int32_t* slotA, const int32_t* slotB, int H, int W, float fx, float fy, float cx, float cy, const bt_prune_params* prune,
                                 int ransac_trials, float ransac_inlier_dist, uint64_t ransac_seed, const bt_match_extra* extra,
                                 bt_entryj* entry_out, int entry_capacity, int32_t* n_entry_out, int32_t* entry_off_out, int32_t* total_out, int32_t* status_out, float* uv_out, void* stream_) {
	BT_REQUIRE(((dA && dB) ? 1 : 0) + ((slotA && slotB) ? 1 : 0) == 1, BT_ERR_INVALID_ARG, "bt_match_pairs_ex: give either descriptor views or pool slots");
	return match_pairs_impl(ctx, n_pairs, A, B, dA, dB, slotA, slotB, H, W, fx, fy, cx, cy, prune, ransac_trials, ransac_inlier_dist, ransac_seed, entry_out, entry_capacity, n_entry_out, entry_off_out, total_out,
	                        stream_, extra, status_out, uv_out);
}

// ------------------------------------------------------------------------------------------------ per-pair result cache
// SiftManager::_matches (/root/reference/src/FeatureManager.cpp:176): a pair is matched once; later windows reuse its matches.  The
// EntryJ blocks stay on the device in fixed-size slots; the host keeps {slot, count, status} per (idA, idB).
namespace bt {
struct CacheMove { int src, dst, n; uint32_t wi, wj; };      // entries [src, src+n) -> [dst, dst+n); gather rewrites the frame indices
__global__ void __launch_bounds__(128) k_cache_move(const CacheMove* __restrict__ tab, const bt_entryj* __restrict__ from, bt_entryj* __restrict__ to, int rewrite) {
	const CacheMove m = tab[blockIdx.x];
	for (int i = threadIdx.x; i < m.n; i += blockDim.x) {
		bt_entryj e = from[m.src + i];
		if (rewrite) { e.imgIdx_i = m.wi; e.imgIdx_j = m.wj; }
		to[m.dst + i] = e;
	}
}
struct MatchCache {
	int max_pairs = 0, slot_entries = 0;
	DevBuf arena, tab;
	PinnedBuf h_tab[2];
	cudaEvent_t ev[2] = { nullptr, nullptr };
	int flip = 0;
	struct Rec { int slot, n, status; };
	std::map<std::pair<int, int>, Rec> recs;
	std::vector<int> free_slots;
};
void mcache_destroy(bt_ctx* ctx) {
	MatchCache* c = ctx->mcache;
	if (!c) return;
	c->arena.release(); c->tab.release(); c->h_tab[0].release(); c->h_tab[1].release();
	for (auto& e : c->ev) if (e) cudaEventDestroy(e);
	delete c;
	ctx->mcache = nullptr;
}
static int cache_upload(MatchCache* c, const std::vector<CacheMove>& mv, cudaStream_t stream) {
	const int fl = c->flip; c->flip ^= 1;
	if (!c->ev[fl]) BT_CUDA(cudaEventCreateWithFlags(&c->ev[fl], cudaEventDisableTiming));
	else BT_CUDA(cudaEventSynchronize(c->ev[fl]));
	memcpy(c->h_tab[fl].p, mv.data(), sizeof(CacheMove) * mv.size());
	BT_CUDA(cudaMemcpyAsync(c->tab.p, c->h_tab[fl].p, sizeof(CacheMove) * mv.size(), cudaMemcpyHostToDevice, stream));
	BT_CUDA(cudaEventRecord(c->ev[fl], stream));
	return BT_OK;
}
}  // namespace bt

extern "C" int bt_match_cache_reserve(bt_ctx* ctx, int max_pairs_cached, int max_entries_per_pair) {
	BT_REQUIRE(ctx && max_pairs_cached > 0 && max_entries_per_pair > 0, BT_ERR_INVALID_ARG, "bt_match_cache_reserve: bad arguments");
	BT_CUDA(cudaSetDevice(ctx->device));
	if (!ctx->mcache) ctx->mcache = new MatchCache();
	MatchCache* c = ctx->mcache;
	int rc;
	if ((rc = c->arena.alloc(sizeof(bt_entryj) * (size_t)max_pairs_cached * max_entries_per_pair)) != BT_OK) return rc;
	if ((rc = c->tab.alloc(sizeof(CacheMove) * (size_t)max_pairs_cached)) != BT_OK) return rc;
	for (auto& h : c->h_tab) if ((rc = h.alloc(sizeof(CacheMove) * (size_t)max_pairs_cached)) != BT_OK) return rc;
	c->max_pairs = max_pairs_cached; c->slot_entries = max_entries_per_pair;
	c->recs.clear(); c->free_slots.clear();
	for (int k = max_pairs_cached - 1; k >= 0; k--) c->free_slots.push_back(k);
	return BT_OK;
}

extern "C" int bt_match_cache_put(bt_ctx* ctx, int n_pairs, const int32_t* idA, const int32_t* idB, const bt_entryj* entry_dev,
                                  const int32_t* n_entry_host, const int32_t* entry_off_host, const int32_t* status_host, void* stream_) {
	BT_REQUIRE(ctx && ctx->mcache && idA && idB && n_entry_host && entry_off_host && n_pairs > 0, BT_ERR_INVALID_ARG, "bt_match_cache_put: call bt_match_cache_reserve first / NULL argument");
	MatchCache* c = ctx->mcache;
	BT_REQUIRE(n_pairs <= c->max_pairs, BT_ERR_CAPACITY, "bt_match_cache_put: %d pairs > table of %d", n_pairs, c->max_pairs);
	BT_CUDA(cudaSetDevice(ctx->device));
	std::vector<CacheMove> mv;
	for (int p = 0; p < n_pairs; p++) {
		BT_REQUIRE(idA[p] > idB[p], BT_ERR_INVALID_ARG, "bt_match_cache_put: pair %d: frame A must be the newer frame", p);
		BT_REQUIRE(n_entry_host[p] >= 0 && n_entry_host[p] <= c->slot_entries, BT_ERR_CAPACITY, "bt_match_cache_put: pair %d has %d entries > slot size %d", p, n_entry_host[p], c->slot_entries);
		const auto key = std::make_pair((int)idA[p], (int)idB[p]);
		auto it = c->recs.find(key);
		int slot;
		if (it != c->recs.end()) slot = it->second.slot;
		else { BT_REQUIRE(!c->free_slots.empty(), BT_ERR_CAPACITY, "bt_match_cache_put: all %d slots are in use", c->max_pairs); slot = c->free_slots.back(); c->free_slots.pop_back(); }
		c->recs[key] = MatchCache::Rec{ slot, n_entry_host[p], status_host ? status_host[p] : (n_entry_host[p] > 0 ? BT_PAIR_OK : BT_PAIR_EMPTY) };
		if (n_entry_host[p] > 0) { BT_REQUIRE(entry_dev, BT_ERR_INVALID_ARG, "bt_match_cache_put: NULL entry list"); mv.push_back(CacheMove{ entry_off_host[p], slot * c->slot_entries, n_entry_host[p], 0u, 0u }); }
	}
	if (mv.empty()) return BT_OK;
	cudaStream_t stream = (cudaStream_t)stream_;
	int rc = cache_upload(c, mv, stream);
	if (rc != BT_OK) return rc;
	k_cache_move<<<(unsigned)mv.size(), 128, 0, stream>>>(c->tab.as<CacheMove>(), entry_dev, c->arena.as<bt_entryj>(), 0);
	BT_CUDA(cudaGetLastError());
	return BT_OK;
}

extern "C" int bt_match_cache_has(bt_ctx* ctx, int idA, int idB) {
	if (!ctx || !ctx->mcache) return 0;
	return ctx->mcache->recs.count(std::make_pair(idA, idB)) ? 1 : 0;
}

extern "C" int bt_match_cache_status(bt_ctx* ctx, int idA, int idB, int* n_entry, int* status) {
	BT_REQUIRE(ctx && ctx->mcache, BT_ERR_INVALID_ARG, "bt_match_cache_status: call bt_match_cache_reserve first");
	const auto it = ctx->mcache->recs.find(std::make_pair(idA, idB));
	BT_REQUIRE(it != ctx->mcache->recs.end(), BT_ERR_INVALID_ARG, "bt_match_cache_status: pair (%d, %d) is not cached", idA, idB);
	if (n_entry) *n_entry = it->second.n;
	if (status) *status = it->second.status;
	return BT_OK;
}

extern "C" int bt_match_cache_gather(bt_ctx* ctx, int n_pairs, const int32_t* idA, const int32_t* idB, const uint32_t* win_i, const uint32_t* win_j,
                                     bt_entryj* dst_dev, int capacity, int32_t* block_off_host, int32_t* block_n_host, void* stream_) {
	BT_REQUIRE(ctx && ctx->mcache && idA && idB && win_i && win_j && dst_dev && block_off_host && block_n_host && n_pairs > 0, BT_ERR_INVALID_ARG, "bt_match_cache_gather: NULL argument");
	MatchCache* c = ctx->mcache;
	BT_REQUIRE(n_pairs <= c->max_pairs, BT_ERR_CAPACITY, "bt_match_cache_gather: %d pairs > table of %d", n_pairs, c->max_pairs);
	BT_CUDA(cudaSetDevice(ctx->device));
	std::vector<CacheMove> mv;
	int off = 0;
	for (int p = 0; p < n_pairs; p++) {
		const auto it = c->recs.find(std::make_pair((int)idA[p], (int)idB[p]));
		BT_REQUIRE(it != c->recs.end(), BT_ERR_INVALID_ARG, "bt_match_cache_gather: pair (%d, %d) is not cached", idA[p], idB[p]);
		const int n = it->second.n;
		BT_REQUIRE(off + n <= capacity, BT_ERR_CAPACITY, "bt_match_cache_gather: %d entries > capacity %d", off + n, capacity);
		block_off_host[p] = off; block_n_host[p] = n;
		if (n > 0) mv.push_back(CacheMove{ it->second.slot * c->slot_entries, off, n, win_i[p], win_j[p] });
		off += n;
	}
	if (mv.empty()) return BT_OK;
	cudaStream_t stream = (cudaStream_t)stream_;
	int rc = cache_upload(c, mv, stream);
	if (rc != BT_OK) return rc;
	k_cache_move<<<(unsigned)mv.size(), 128, 0, stream>>>(c->tab.as<CacheMove>(), c->arena.as<bt_entryj>(), dst_dev, 1);
	BT_CUDA(cudaGetLastError());
	return BT_OK;
}

extern "C" int bt_match_cache_forget_frame(bt_ctx* ctx, int frame_id) {
	BT_REQUIRE(ctx && ctx->mcache, BT_ERR_INVALID_ARG, "bt_match_cache_forget_frame: call bt_match_cache_reserve first");
	MatchCache* c = ctx->mcache;
	for (auto it = c->recs.begin(); it != c->recs.end();) {
		if (it->first.first == frame_id || it->first.second == frame_id) { c->free_slots.push_back(it->second.slot); it = c->recs.erase(it); }
		else ++it;
	}
	return BT_OK;
}
