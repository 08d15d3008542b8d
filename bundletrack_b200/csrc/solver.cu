// solver.cu — B200-native Gauss-Newton x PCG pose-graph solve for a BATCH of independent tracking windows.
//
// Replaces (behind the C-ABI in include/bundletrack_b200.h) the reference's OptimizerGpu::optimizeFrames
// (/root/reference/src/cuda/LossGPU.cu:53-139) -> CUDACache::storeFrame (CUDACache.cpp:76-88) -> SBA::align
// (SBA.cpp:81-139) -> CUDASolverBundling::solve (Solver/CUDASolverBundling.cpp:190-288) -> solveBundlingStub
// (Solver/SolverBundling.cu:931-1003): ~260 kernel launches, ~75 memsets, 7 blocking copies and ~110 cudaMalloc per
// window there; THREE launches per BATCH here, none of them allocating:
//
//   k_prep_count   one CTA per 1024 quarter-res pixels of a (window, frame): number of valid pixels of the block.
//   k_prep_frames  same grid: quarter-res cache (camera-space point + normal interleaved in one
//                  32-byte texel so a bilinear tap is a single sector) + an order-preserving COMPACTED list of the
//                  valid source pixels (the object mask leaves ~10 % of the image valid), + se(3) of the input pose.
//                  Each block's offset in the ordered list = the counts of the blocks before it.
//   (plan)         the LAST CTA of k_prep_frames to finish: per-window tile counts -> exclusive scan -> flat tile list.
//   k_solve        persistent: tiles are (GN iteration, window, pair, pixel chunk), handed out by a queue (the next tile's 48-byte
//                  record is fetched one tile ahead) or - when one iteration's tiles fit the grid - assigned statically (CTA c =
//                  tile c of every iteration).  A tile evaluates point-to-plane residuals/Jacobians for its chunk and reduces a 6x6
//                  system in the TARGET camera frame; every window also has two SPARSE tiles per iteration (the moment sums of half
//                  of its correspondence groups each: they only need the poses, so they run next to the dense tiles); the CTA that
//                  retires a window's last tile of an iteration runs that window's "tail": table loads as asynchronous copies next
//                  to the per-pair sums, assembly of the (6(N-1))^2 normal equations in shared memory (dense blocks + explicit
//                  sparse blocks from per-pair moment sums), the PCG steps with one thread per unknown, the pose update, and the
//                  release of the window's next iteration.  Tiles of iteration k+1 wait on a per-window flag, so all GN iterations
//                  of all windows flow through ONE launch with no host round trip.
//
// Maths restated from the reference (see oracle/solver_oracle.c for the statement-by-statement CPU version):
//   * dense residual/gates: findDenseCorr (Solver/SolverBundlingDenseUtil.h:78-110), bilinear taps
//     (Solver/ICPUtil.h:83-110, zeros blended - SURVEY.md Q6), Huber weight (Solver/SolverBundlingUtil.h:24-39).
//   * Jacobian rows: computeJacobianBlockRow_i/j (Solver/SolverBundlingEquationsLie.h:214-230) reduce analytically to
//     J_i = [N, q x N], J_j = -J_i with N = R_i n_tgt and q the source point in the model frame (derivation in
//     DESIGN.md), so one symmetric 6x6 + one 6-vector per pair carries all four blocks addToLocalSystem writes
//     (SolverBundlingDenseUtil.h:217-285).  FlipJtJ (SolverBundling.cu:49-58) erases cross blocks with
//     target > source; `compat_flip` reproduces that.
//   * sparse term: evalMinusJTFDevice / applyJDevice / applyJTDevice (SolverBundlingEquationsLie.h:60-211) with the
//     Huber weight on the gradient and the Jacobi preconditioner only; J^T J is formed explicitly
//     from per-pair moment sums instead of matrix-free, which changes summation order only.
//   * PCG recurrences and the 1e-6 guards: SolverBundling.cu:575-818; update x <- log(exp(delta) exp(x)):
//     Solver/LieDerivUtil.h:126-194,276-282.
#include <algorithm>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <chrono>
#include "bt_common.cuh"

namespace bt {

#ifndef BT_SOLVE_MIN_CTAS
#define BT_SOLVE_MIN_CTAS 2
#endif
static constexpr int kTileVals = 28;          // 21 (sym 6x6) + 6 (rhs) + 1 (#correspondences found)
static constexpr int kGrpVals = 44;           // sparse moment sums per pair group
static constexpr int kMaxFrames = 32;
static constexpr int kSparseTiles = 2;        // sparse tiles per window and GN iteration: each takes a contiguous share of the window's (i,j) groups (45 groups = 23 + 22: one round of
                                              // 32 eight-lane teams each instead of two; measured as the longest tile of a lone window when it was one tile)
static constexpr size_t kTailSmemMax = 224 * 1024;   // dynamic shared memory a tail may use (227 KB per CTA minus k_solve's static arrays)
// (a 128-thread CTA variant of k_solve was measured on B200: SLOWER, 0.55 vs 0.46 ms for 32 windows - dropped)
static constexpr float kEps = 0.000001f;      // FLOAT_EPSILON, /root/reference/src/cuda/SolverUtil.h:10

struct WinDesc {
	int n_frames, n_corr, n_groups, n_pairs;
	int corr_off, grp_off, pair_off, frame_off;
	int tile_off, n_tiles;            // per GN iteration; written by the tile plan (last CTA of k_prep_frames)
	int H, W, w, h;                   // full / quarter resolution
	float fx, fy, cx, cy;             // quarter-res intrinsics (CUDACache.cpp:20-24)
	float ifx, ify, icx, icy;         // inverse full-res intrinsics (m_inputIntrinsicsInv)
	float scaleW, scaleH;             // (W-1)/(w-1), (H-1)/(h-1)  (CUDAImageUtil.cu:57-58)
	int compat_flip;
	int mem_off;                      // per-frame membership CSR of this window (see SolveArgs::mem)
	int unique_blocks;                // 1: no two groups / no two dense pairs share an unordered frame pair => cross blocks need no atomics
	int pad[3];
};

// The correspondence-side fields of a window, uploaded AFTER the frame preparation has been launched (the host scans the
// correspondences while the GPU already works); window_tail reads them from here, the copies inside WinDesc are unused.
struct WinSparse { int n_corr, n_groups, corr_off, grp_off, mem_off, unique_blocks, pad0, pad1; };

// One unit of work of k_solve.  Everything the tile's prologue needs travels in the record (frame slots, map pointers, the window's
// tile count), so that claiming a tile costs ONE dependent load - and that one is issued a whole tile ahead (k_solve keeps the next
// record in shared memory).  pair == -2 => the window's sparse tile (moment sums of its correspondences).
struct __align__(16) Tile {
	int win; int pair; int start; int count;
	int tgt_slot; int src_slot;          // global frame slots (frame_off + target / source) of the pair
	int n_tiles_win; int pad;            // tiles of the window per GN iteration (dense + the sparse one)
	const float4* src; const float4* tex;   // compacted source list of the source frame, texel map of the target frame
};

struct SolveArgs {
	const WinDesc* wins;
	const WinSparse* wsp;
	int n_windows;
	// frame slots
	const float* const* depth_ptr;
	const float4* const* normal_ptr;
	const int* frame_win;
	float4* texel;        // [F][2*npix_max]  the context's own maps (frames prepared by this call)
	float4* src;          // [F][2*npix_max]
	int* nsrc;            // [F]
	float* grp_sums;      // [groups of the batch][kGrpVals] sparse moment sums of the current GN iteration (written by the window's sparse tile)
	const float4* const* texel_tab;   // [F] where each frame slot's texel map / source list lives: the arrays above, or a
	const float4* const* src_tab;     //     frame-cache slot built earlier by bt_frame_cache_store
	const int* const* nsrc_cached;    // [F] cached frame: address of its source count, else nullptr
	const float* pose_in; // [F][16]
	float* x;             // [F][6]  rot, trans
	float* T;             // [F][12] row-major 3x4 cam->model
	float* pose_out;      // [F][16]
	int npix_max;
	// correspondences, grouped by (i,j)
	const bt_entryj* corr;
	const int* grp_i; const int* grp_j; const int* grp_start;   // grp_start has n_groups+1 entries per window
	const uint2* pairs;   // (target, source)
	const int* pair_win;  // owning window of every entry of `pairs`
	const int* pair_src_slot;   // global frame slot (frame_off + source) of every entry of `pairs`
	// per-window CSR built on the host: for every frame f, first the correspondence groups touching f, then the dense
	// pairs touching f.  Layout at mem_off: fg_start[N+1], fp_start[N+1], items[2G] (g | role<<16), items[2P] (p | role<<16)
	const int* mem;
	// schedule
	Tile* tiles; int* n_tiles_total; int max_tiles;
	int* pair_tile0; int* pair_ntile;   // per (window, pair): first tile (relative to the window's tile_off) and tile count
	float* partial;       // [max_tiles][kTileVals]
	int* tiles_done;      // [n_windows]
	int* iter_done;       // [n_windows]
	int* queue;           // [1]
	long long* n_src_px;  // [1] stats
	int chunk;
	int grp_in_smem;      // 0: the tail reads the groups' moment sums from global memory (windows too big for shared memory)
	int grid_ctas;        // CTAs k_solve will be launched with (tile planning targets one wave for small batches)
	float d2max;          // largest float whose correctly rounded square root is <= prm.dense_dist_thresh (computed once on the host)
	bt_solver_params prm;
	float* dbg_JtJ; float* dbg_Jtr; int dbg_stride;   // optional dense-system dump (last GN iteration)
	float* dbg_cnt; int dbg_cnt_stride;               // optional per-pair #correspondences found (last GN iteration)
	long long* prof; int prof_cap;                    // optional phase timestamps: [0] = record count, then 12 x int64 per record
};

// ------------------------------------------------------------------------------------------------ device math
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// Rodrigues coefficients with the reference's Taylor branches (LieDerivUtil.h:46-70,150-194).
__device__ void so3_coeffs(float theta_sq, float& A, float& B, float& C) {
	if (theta_sq < 1e-8f) { A = 1.0f - 0.16666667f * theta_sq; B = 0.5f; C = 0.16666667f; }
	else if (theta_sq < 1e-6f) {
		C = 0.16666667f * (1.0f - 0.05f * theta_sq);
		A = 1.0f - theta_sq * C;
		B = 0.5f - 0.25f * 0.16666667f * theta_sq;
	} else {
		const float theta = sqrtf(theta_sq), inv = 1.0f / theta;
		A = sinf(theta) * inv;
		B = (1.0f - cosf(theta)) * (inv * inv);
		C = (1.0f - A) * (inv * inv);
	}
}
__device__ void so3_exp_AB(V3 w, float A, float B, float R[9]) {  // rodrigues_so3_exp, LieDerivUtil.h:17-44
	const float wx2 = w.x * w.x, wy2 = w.y * w.y, wz2 = w.z * w.z;
	R[0] = 1.0f - B * (wy2 + wz2); R[4] = 1.0f - B * (wx2 + wz2); R[8] = 1.0f - B * (wx2 + wy2);
	float a = A * w.z, b = B * (w.x * w.y); R[1] = b - a; R[3] = b + a;
	a = A * w.y; b = B * (w.x * w.z); R[2] = b + a; R[6] = b - a;
	a = A * w.x; b = B * (w.y * w.z); R[5] = b - a; R[7] = b + a;
}
// poseToMatrix (LieDerivUtil.h:150-194): T[12] = row-major 3x4 [R | t]
__device__ __forceinline__ void se3_exp(V3 rot, V3 trans, float T[12]) {
	const float theta_sq = dot(rot, rot);
	float A, B, C;
	so3_coeffs(theta_sq, A, B, C);
	const V3 cr = cross(rot, trans);
	V3 t;
	if (theta_sq < 1e-8f) t = trans + cr * 0.5f;
	else t = trans + cr * B + cross(rot, cr) * C;
	float R[9];
	so3_exp_AB(rot, A, B, R);
	T[0] = R[0]; T[1] = R[1]; T[2] = R[2]; T[3] = t.x;
	T[4] = R[3]; T[5] = R[4]; T[6] = R[5]; T[7] = t.y;
	T[8] = R[6]; T[9] = R[7]; T[10] = R[8]; T[11] = t.z;
}
// ln_rotation (LieDerivUtil.h:72-124)
__device__ V3 so3_log(const float R[9]) {
	V3 r = mk((R[7] - R[5]) * 0.5f, (R[2] - R[6]) * 0.5f, (R[3] - R[1]) * 0.5f);
	const float cos_angle = (R[0] + R[4] + R[8] - 1.0f) * 0.5f;
	const float s = sqrtf(dot(r, r));
	if (cos_angle > 0.70710678118654752440f) {
		if (s > 0.0f) r = r * (asinf(s) / s);
	} else if (cos_angle > -0.70710678118654752440f) {
		r = r * (acosf(cos_angle) / s);
	} else {
		const float angle = 3.14159265358979323846f - asinf(s);
		const float d0 = R[0] - cos_angle, d1 = R[4] - cos_angle, d2 = R[8] - cos_angle;
		V3 r2;
		if (fabsf(d0) > fabsf(d1) && fabsf(d0) > fabsf(d2)) r2 = mk(d0, (R[3] + R[1]) * 0.5f, (R[2] + R[6]) * 0.5f);
		else if (fabsf(d1) > fabsf(d2)) r2 = mk((R[3] + R[1]) * 0.5f, d1, (R[7] + R[5]) * 0.5f);
		else r2 = mk((R[2] + R[6]) * 0.5f, (R[7] + R[5]) * 0.5f, d2);
		if (dot(r2, r) < 0.0f) r2 = r2 * -1.0f;
		r = r2 * (angle / sqrtf(dot(r2, r2)));
	}
	return r;
}
// matrixToPose (LieDerivUtil.h:126-148); T = row-major 3x4
__device__ __forceinline__ void se3_log(const float T[12], V3& rot, V3& trans) {
	const float R[9] = { T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10] };
	const V3 t = mk(T[3], T[7], T[11]);
	rot = so3_log(R);
	const float theta = sqrtf(dot(rot, rot));
	float shtot = 0.5f;
	if (theta > 0.00001f) shtot = sinf(theta * 0.5f) / theta;
	const V3 half = rot * -0.5f;
	float A, B, C, Hm[9];
	so3_coeffs(dot(half, half), A, B, C);
	so3_exp_AB(half, A, B, Hm);
	V3 tr = mk(Hm[0] * t.x + Hm[1] * t.y + Hm[2] * t.z, Hm[3] * t.x + Hm[4] * t.y + Hm[5] * t.z, Hm[6] * t.x + Hm[7] * t.y + Hm[8] * t.z);
	if (theta > 0.001f) tr = tr - rot * (dot(t, rot) * (1.0f - 2.0f * shtot) / dot(rot, rot));
	else tr = tr - rot * (dot(t, rot) / 24.0f);
	trans = tr * (1.0f / (2.0f * shtot));
}
__device__ __forceinline__ float huber_w(float e, float delta) {  // rho.y of huberLoss, SolverBundlingUtil.h:24-39
	return (e <= delta * delta) ? 1.0f : delta / sqrtf(e);
}
// The dense pixel loop's quotients.  Default: correctly rounded (__frcp_rn, IEEE Huber weight).  -DBT_DENSE_FAST_RCP builds the variant
// with one MUFU reciprocal (1 ulp) per quotient: the reference itself is compiled with -use_fast_math (/root/reference/CMakeLists.txt:7),
// and against its live kernels over 200 windows the two variants have the same parity statistics; the variant is 6 % faster
// on the batch, but individual gate-sensitive windows move by ~1e-4 either way, so the shipped default keeps the arithmetic the
// committed parity windows were pinned with.
#ifndef BT_DENSE_FAST_RCP      // default: correctly rounded reciprocals / IEEE Huber weight; -DBT_DENSE_FAST_RCP builds the MUFU variant (A/B in DESIGN.md)
__device__ __forceinline__ float rcp_fast(float x) { return __frcp_rn(x); }
__device__ __forceinline__ float huber_w_fast(float e, float delta) { return huber_w(e, delta); }
#else
__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float huber_w_fast(float e, float delta) {
	float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e));
	return (e <= delta * delta) ? 1.0f : delta * r;
}
#endif
// 256-bit global load through the read-only path (sm_100: LDG.E.ENL2.256.CONSTANT): one instruction per 32-byte texel / source record
struct __align__(32) F8 { float4 lo, hi; };
__device__ __forceinline__ F8 ldg256(const float4* p) {
	F8 r;
	asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) : "l"(p));
	return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ int ld_acquire(const int* p) {
	int v;
	asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
// Polling load for the iteration flags: relaxed, i.e. WITHOUT the L1 invalidation (CCTL.IVALL) every ld.acquire.gpu carries - a CTA
// spinning on an acquire load flushed its SM's L1 five million times per launch (ncu), under the feet of the co-resident CTA's
// texel taps.  The waiter confirms with ONE acquire load once the relaxed load has seen the value.
__device__ __forceinline__ int ld_relaxed(const int* p) {
	int v;
	asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
// ticket with release (orders this thread's - and, through the preceding __syncwarp / barrier, its warp's / CTA's - earlier writes
// before it) and acquire (the CTA that draws the last ticket sees every other tile's writes) semantics in ONE instruction; a full
// __threadfence() by every writer lane in front of a relaxed atomic cost ~1 k cycles per tile
__device__ __forceinline__ int ticket_acq_rel(int* p) {
	int old;
	asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(old) : "l"(p) : "memory");
	return old;
}
// asynchronous global -> shared copy of one 48-byte tile record (no registers held while it is in flight)
__device__ __forceinline__ void tile_fetch_async(void* smem_dst, const void* gsrc) {
	const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n cp.async.cg.shared.global [%2], [%3], 16;\n cp.async.cg.shared.global [%4], [%5], 16;"
	             ::"r"(d), "l"(gsrc), "r"(d + 16), "l"((const char*)gsrc + 16), "r"(d + 32), "l"((const char*)gsrc + 32) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {      // 4-byte asynchronous copy (through L1: immutable tables only)
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_cg(void* smem_dst, const void* gsrc) {  // 16-byte asynchronous copy from L2 (data other CTAs wrote during this launch)
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void tile_fetch_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void st_release(int* p, int v) {
	asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Optional in-kernel phase profile (developer aid): thread 0 stamps clock64() at phase boundaries.
struct ProfRec { long long kind_cta, tile_win, t[10]; };
__device__ __forceinline__ void prof_emit(const SolveArgs& a, const ProfRec& r) {
	const unsigned long long k = atomicAdd((unsigned long long*)a.prof, 1ull);
	if ((int)k < a.prof_cap) {
		long long* o = a.prof + 1 + k * 12;
		o[0] = r.kind_cta; o[1] = r.tile_win;
		for (int i = 0; i < 10; i++) o[2 + i] = r.t[i];
	}
}
#ifdef BT_PROF_GLOBALTIMER      // (variant build: nanoseconds of the device-wide timer, comparable across SMs, instead of the SM's own cycle counter)
__device__ __forceinline__ long long prof_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (long long)t; }
#else
__device__ __forceinline__ long long prof_now() { return clock64(); }
#endif
#define PROF_T(i) do { if (a.prof && threadIdx.x == 0) prf.t[i] = prof_now(); } while (0)

// ------------------------------------------------------------------------------------------------ k_prep_frames
// CUDACache::storeFrame fused (convertDepthFloatToCameraSpaceFloat4 + 2x resampleFloat4 nearest, CUDAImageUtil.cu:
// 310-326,82-99) + compaction of the source list + matrixToPose/poseToMatrix of the incoming pose (SBA.cu:71-79).
__device__ void plan_body(const SolveArgs& a, WinDesc* wins_rw);

// One CTA per frame.  The compacted source list keeps pixel order (fixed summation order downstream), so the CTA needs
// an ordered prefix over its pixels; instead of one block-wide scan per 1024 pixels (19 x 3 barriers per frame, the loads
// of each round exposed) it runs two sweeps around ONE scan: sweep 1 counts the valid pixels of every (round, warp) from
// the depth map alone, the scan turns the counts into offsets, sweep 2 loads depth + normal, writes the texel map and the
// source records at their final position.  No barrier inside either sweep, so the loads of consecutive rounds overlap.
static constexpr int kPrepRounds = 64;      // rounds of 1024 pixels per scan (one scan covers 65536 quarter-res pixels)

struct FrameGeom { int W, H, w, h; float ifx, ify, icx, icy, scaleW, scaleH; };

// Builds one frame's quarter-resolution texel map and compacted source list (1024 threads); returns the source count.
__device__ int build_frame_maps(const FrameGeom& g, const float* __restrict__ depth, const float4* __restrict__ normal, float4* texel, float4* src,
                                float dmin, float dmax, int* s_cnt, int* s_wt) {
	const int npix = g.w * g.h;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	// nearest-neighbour sample position of quarter-res pixel idx in the full-res maps, or -1 (resampleFloat4, CUDAImageUtil.cu:82-99)
	auto sample = [&](int idx, unsigned& xi, unsigned& yi) -> bool {
		const int x = idx % g.w, y = idx / g.w;
		xi = (unsigned)((float)x * g.scaleW + 0.5f); yi = (unsigned)((float)y * g.scaleH + 0.5f);
		return xi < (unsigned)g.W && yi < (unsigned)g.H;
	};
	int base = 0;
	for (int sb = 0; sb < npix; sb += kPrepRounds * 1024) {
		const int rounds = min(kPrepRounds, (npix - sb + 1023) >> 10);
		// ---- sweep 1: valid counts per (round, warp)
#pragma unroll 4
		for (int r = 0; r < rounds; r++) {
			const int idx = sb + (r << 10) + tid;
			bool valid = false;
			if (idx < npix) {       // must agree bit for bit with sweep 2: z = 0 for a missing / too-near sample
				unsigned xi, yi;
				float z = 0.f;
				if (sample(idx, xi, yi)) { const float d = __ldg(depth + (size_t)yi * g.W + xi); if (d >= 0.1f) z = d; }
				valid = (z > dmin && z < dmax);
			}
			const unsigned bal = __ballot_sync(0xffffffffu, valid);
			if (lane == 0) s_cnt[r * 32 + wid] = __popc(bal);
		}
		__syncthreads();
		// ---- exclusive scan of the rounds*32 counts in (round, warp) order: two entries per thread
		const int n = rounds * 32;
		const int v0 = (2 * tid < n) ? s_cnt[2 * tid] : 0, v1 = (2 * tid + 1 < n) ? s_cnt[2 * tid + 1] : 0;
		int incl = v0 + v1;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
		if (lane == 31) s_wt[wid] = incl;
		__syncthreads();
		if (wid == 0) {
			int t = s_wt[lane];
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += u; }
			s_wt[lane] = t;
		}
		__syncthreads();
		const int excl = base + incl - (v0 + v1) + (wid ? s_wt[wid - 1] : 0);
		if (2 * tid < n) s_cnt[2 * tid] = excl;
		if (2 * tid + 1 < n) s_cnt[2 * tid + 1] = excl + v0;
		base += s_wt[31];
		__syncthreads();
		// ---- sweep 2: texel map + source records at their final offsets
#pragma unroll 2
		for (int r = 0; r < rounds; r++) {
			const int idx = sb + (r << 10) + tid;
			float4 cp = make_float4(0.f, 0.f, 0.f, 0.f), nr = make_float4(0.f, 0.f, 0.f, 0.f);
			bool valid = false;
			if (idx < npix) {
				unsigned xi, yi;
				if (sample(idx, xi, yi)) {
					const size_t sidx = (size_t)yi * g.W + xi;
					const float d = __ldg(depth + sidx);
					nr = __ldg(normal + sidx);
					if (d >= 0.1f) cp = make_float4(g.ifx * ((float)xi * d) + g.icx * d, g.ify * ((float)yi * d) + g.icy * d, d, 1.0f);
				}
				texel[2 * idx] = make_float4(cp.x, cp.y, cp.z, nr.x);      // 32-byte texel: point xyz + normal xyz (+ pad)
				texel[2 * idx + 1] = make_float4(nr.y, nr.z, 0.f, 0.f);
				valid = (cp.z > dmin && cp.z < dmax);
			}
			const unsigned bal = __ballot_sync(0xffffffffu, valid);
			if (valid) {
				const int o = s_cnt[r * 32 + wid] + __popc(bal & ((1u << lane) - 1u));
				src[2 * o] = make_float4(cp.x, cp.y, cp.z, nr.x);
				src[2 * o + 1] = make_float4(nr.y, nr.z, nr.w, 0.f);
			}
		}
		__syncthreads();
	}
	return base;
}

// Frame preparation, one CTA per 1024 quarter-res pixels of a frame (two launches): k_prep_count counts the valid pixels of
// every block, k_prep_frames turns the counts of the blocks before it into its offset in the frame's ordered source list and
// writes its slice of the texel map and of the list.  (One CTA per frame walked 19 rounds serially: 50 us for a lone window.)
__device__ __forceinline__ bool prep_sample(const WinDesc& wd, int idx, unsigned& xi, unsigned& yi) {
	// nearest-neighbour sample position of quarter-res pixel idx in the full-res maps (resampleFloat4, CUDAImageUtil.cu:82-99)
	const int x = idx % wd.w, y = idx / wd.w;
	xi = (unsigned)((float)x * wd.scaleW + 0.5f); yi = (unsigned)((float)y * wd.scaleH + 0.5f);
	return xi < (unsigned)wd.W && yi < (unsigned)wd.H;
}

__global__ void __launch_bounds__(1024, 2) k_prep_count(SolveArgs a, int* __restrict__ blk_cnt) {
	const int fs = blockIdx.y, b = blockIdx.x;
	if (a.nsrc_cached[fs] || !(a.prm.w_dense > 0.0f)) return;
	const WinDesc wd = a.wins[a.frame_win[fs]];
	const int npix = wd.w * wd.h, idx = b * 1024 + threadIdx.x;
	__shared__ int s_w[32];
	bool valid = false;
	if (idx < npix) {       // must agree bit for bit with k_prep_frames: z = 0 for a missing / too-near sample
		unsigned xi, yi;
		float z = 0.f;
		if (prep_sample(wd, idx, xi, yi)) { const float d = __ldg(a.depth_ptr[fs] + (size_t)yi * wd.W + xi); if (d >= 0.1f) z = d; }
		valid = (z > a.prm.depth_min && z < a.prm.depth_max);
	}
	const unsigned bal = __ballot_sync(0xffffffffu, valid);
	if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = __popc(bal);
	__syncthreads();
	if (threadIdx.x < 32) {
		int t = s_w[threadIdx.x];
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
		if (threadIdx.x == 0) blk_cnt[fs * gridDim.x + b] = t;
	}
}

__global__ void __launch_bounds__(1024, 2) k_prep_frames(SolveArgs a, WinDesc* wins_rw, int* prep_ticket, const int* __restrict__ blk_cnt) {
	const int fs = blockIdx.y, b = blockIdx.x;
	const WinDesc wd = a.wins[a.frame_win[fs]];
	__shared__ int s_w[32];
	__shared__ int s_base;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int* cached = a.nsrc_cached[fs];
	const int npix = wd.w * wd.h;
	if (cached) { if (b == 0 && tid == 0) a.nsrc[fs] = *cached; }      // maps and source list were built by bt_frame_cache_store
	else if (!(a.prm.w_dense > 0.0f)) { if (b == 0 && tid == 0) a.nsrc[fs] = 0; }
	else if (b * 1024 < npix) {
		if (wid == 0) {      // offset of this block = valid pixels of the blocks before it (at most a few dozen)
			int t = 0;
			for (int q = lane; q < b; q += 32) t += blk_cnt[fs * gridDim.x + q];
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
			if (lane == 0) s_base = t;
		}
		const float dmin = a.prm.depth_min, dmax = a.prm.depth_max;
		float4* texel = a.texel + (size_t)fs * 2 * a.npix_max;
		float4* src = a.src + (size_t)fs * 2 * a.npix_max;
		const int idx = b * 1024 + tid;
		float4 cp = make_float4(0.f, 0.f, 0.f, 0.f), nr = make_float4(0.f, 0.f, 0.f, 0.f);
		bool valid = false;
		if (idx < npix) {
			unsigned xi, yi;
			if (prep_sample(wd, idx, xi, yi)) {
				const size_t sidx = (size_t)yi * wd.W + xi;
				const float d = __ldg(a.depth_ptr[fs] + sidx);
				nr = __ldg(a.normal_ptr[fs] + sidx);
				if (d >= 0.1f) cp = make_float4(wd.ifx * ((float)xi * d) + wd.icx * d, wd.ify * ((float)yi * d) + wd.icy * d, d, 1.0f);
			}
			texel[2 * idx] = make_float4(cp.x, cp.y, cp.z, nr.x);      // 32-byte texel: point xyz + normal xyz (+ pad)
			texel[2 * idx + 1] = make_float4(nr.y, nr.z, 0.f, 0.f);
			valid = (cp.z > dmin && cp.z < dmax);
		}
		const unsigned bal = __ballot_sync(0xffffffffu, valid);
		if (lane == 0) s_w[wid] = __popc(bal);
		__syncthreads();
		int off = (lane < wid) ? s_w[lane] : 0;       // exclusive prefix over the warps of this block
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) off += __shfl_xor_sync(0xffffffffu, off, o);
		if (valid) {
			const int o = s_base + off + __popc(bal & ((1u << lane) - 1u));
			src[2 * o] = make_float4(cp.x, cp.y, cp.z, nr.x);
			src[2 * o + 1] = make_float4(nr.y, nr.z, nr.w, 0.f);
		}
		if ((b + 1) * 1024 >= npix && tid == 1023) a.nsrc[fs] = s_base + off + __popc(bal);     // last block of the frame, last warp: the total
	}
	if (b == 0 && tid == 1023) {      // pose -> (rot, trans) -> T of iteration 0 (one lane of the frame's first block)
		float Tm[12];
		const float* P = a.pose_in + (size_t)fs * 16;
		for (int k = 0; k < 12; k++) Tm[k] = P[k];
		V3 rot, trans;
		se3_log(Tm, rot, trans);
		float* x = a.x + (size_t)fs * 6;
		x[0] = rot.x; x[1] = rot.y; x[2] = rot.z; x[3] = trans.x; x[4] = trans.y; x[5] = trans.z;
		se3_exp(rot, trans, Tm);
		float* T = a.T + (size_t)fs * 12;
		for (int k = 0; k < 12; k++) T[k] = Tm[k];
	}
	// the last CTA to finish plans the dense tiles of the whole batch (saves a launch and its dependency gap)
	__shared__ int s_last;
	__threadfence();
	__syncthreads();
	if (tid == 0) { const int t = atomicAdd(prep_ticket, 1); s_last = (t == (int)(gridDim.x * gridDim.y) - 1); if (s_last) { *prep_ticket = 0; __threadfence(); } }
	__syncthreads();
	if (s_last) plan_body(a, wins_rw);
}

// bt_frame_cache_store: the same maps, built once per keyframe into a slot of the context's frame cache.
struct CacheStoreArgs {
	FrameGeom g;
	const float* const* depth; const float4* const* normal; const int* slots;
	float4* texel; float4* src; int* nsrc; size_t slot_stride;   // float4 per slot (2 * npix_max)
	float dmin, dmax;
};
__global__ void __launch_bounds__(1024, 2) k_frame_cache_store(CacheStoreArgs c) {
	__shared__ int s_cnt[kPrepRounds * 32];
	__shared__ int s_wt[32];
	const int slot = c.slots[blockIdx.x];
	const int n = build_frame_maps(c.g, c.depth[blockIdx.x], c.normal[blockIdx.x], c.texel + (size_t)slot * c.slot_stride, c.src + (size_t)slot * c.slot_stride,
	                               c.dmin, c.dmax, s_cnt, s_wt);
	if (threadIdx.x == 0) c.nsrc[slot] = n;
}

// The same store spread over one CTA per 1024 quarter-res pixels (count + build, like k_prep_count / k_prep_frames): a handful of new
// frames per step - one per tracked window - otherwise runs on a handful of CTAs walking 19 rounds each (50 us for 32 frames).
__device__ __forceinline__ bool geom_sample(const FrameGeom& g, int idx, unsigned& xi, unsigned& yi) {
	const int x = idx % g.w, y = idx / g.w;
	xi = (unsigned)((float)x * g.scaleW + 0.5f); yi = (unsigned)((float)y * g.scaleH + 0.5f);
	return xi < (unsigned)g.W && yi < (unsigned)g.H;
}
__global__ void __launch_bounds__(1024, 2) k_cache_count(CacheStoreArgs c, int* __restrict__ blk_cnt) {
	const int f = blockIdx.y, b = blockIdx.x;
	const int npix = c.g.w * c.g.h, idx = b * 1024 + threadIdx.x;
	__shared__ int s_w[32];
	bool valid = false;
	if (idx < npix) {       // must agree bit for bit with k_cache_build: z = 0 for a missing / too-near sample
		unsigned xi, yi;
		float z = 0.f;
		if (geom_sample(c.g, idx, xi, yi)) { const float d = __ldg(c.depth[f] + (size_t)yi * c.g.W + xi); if (d >= 0.1f) z = d; }
		valid = (z > c.dmin && z < c.dmax);
	}
	const unsigned bal = __ballot_sync(0xffffffffu, valid);
	if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = __popc(bal);
	__syncthreads();
	if (threadIdx.x < 32) {
		int t = s_w[threadIdx.x];
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
		if (threadIdx.x == 0) blk_cnt[f * gridDim.x + b] = t;
	}
}
__global__ void __launch_bounds__(1024, 2) k_cache_build(CacheStoreArgs c, const int* __restrict__ blk_cnt) {
	const int f = blockIdx.y, b = blockIdx.x;
	__shared__ int s_w[32];
	__shared__ int s_base;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int npix = c.g.w * c.g.h;
	if (b * 1024 >= npix) return;
	const int slot = c.slots[f];
	if (wid == 0) {      // offset of this block = valid pixels of the blocks before it
		int t = 0;
		for (int q = lane; q < b; q += 32) t += blk_cnt[f * gridDim.x + q];
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
		if (lane == 0) s_base = t;
	}
	float4* texel = c.texel + (size_t)slot * c.slot_stride;
	float4* src = c.src + (size_t)slot * c.slot_stride;
	const int idx = b * 1024 + tid;
	float4 cp = make_float4(0.f, 0.f, 0.f, 0.f), nr = make_float4(0.f, 0.f, 0.f, 0.f);
	bool valid = false;
	if (idx < npix) {
		unsigned xi, yi;
		if (geom_sample(c.g, idx, xi, yi)) {
			const size_t sidx = (size_t)yi * c.g.W + xi;
			const float d = __ldg(c.depth[f] + sidx);
			nr = __ldg(c.normal[f] + sidx);
			if (d >= 0.1f) cp = make_float4(c.g.ifx * ((float)xi * d) + c.g.icx * d, c.g.ify * ((float)yi * d) + c.g.icy * d, d, 1.0f);
		}
		texel[2 * idx] = make_float4(cp.x, cp.y, cp.z, nr.x);
		texel[2 * idx + 1] = make_float4(nr.y, nr.z, 0.f, 0.f);
		valid = (cp.z > c.dmin && cp.z < c.dmax);
	}
	const unsigned bal = __ballot_sync(0xffffffffu, valid);
	if (lane == 0) s_w[wid] = __popc(bal);
	__syncthreads();
	int off = (lane < wid) ? s_w[lane] : 0;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) off += __shfl_xor_sync(0xffffffffu, off, o);
	if (valid) {
		const int o = s_base + off + __popc(bal & ((1u << lane) - 1u));
		src[2 * o] = make_float4(cp.x, cp.y, cp.z, nr.x);
		src[2 * o + 1] = make_float4(nr.y, nr.z, nr.w, 0.f);
	}
	if ((b + 1) * 1024 >= npix && tid == 1023) c.nsrc[slot] = s_base + off + __popc(bal);
}

// ------------------------------------------------------------------------------------------------ tile plan
__device__ __forceinline__ int chunks_for(int n, int chunk, int& per) {
	if (n <= 0) { per = 0; return 0; }
	const int nch = (n + chunk - 1) / chunk;
	per = (((n + nch - 1) / nch) + 31) & ~31;   // balanced, warp-aligned
	return (n + per - 1) / per;
}
// Runs on the LAST CTA of k_prep_frames to finish (1024 threads): per-window tile counts -> scans -> flat tile list.
__device__ void plan_body(const SolveArgs& a, WinDesc* wins_rw) {
	__shared__ int s_cnt[1024];
	__shared__ int s_scan[1024];
	__shared__ int s_carry;
	__shared__ unsigned long long s_px;
	const int tid = threadIdx.x;
	ProfRec prf; prf.kind_cta = (3ll << 32) | blockIdx.x; prf.tile_win = a.n_windows;
	for (int i = 0; i < 10; i++) prf.t[i] = 0;
	PROF_T(0);
	if (tid == 0) { s_carry = 0; s_px = 0ull; *a.queue = 0; }
	__syncthreads();
	// Small batches: when one GN iteration has only slightly more tiles than k_solve has CTAs, the few CTAs that get two tiles
	// put a whole extra tile on the critical path of every iteration.  Grow the chunk until an iteration fits one wave.
	int chunk = a.chunk;
	if (a.n_windows <= 1024) {
		__shared__ int s_total;
		const WinDesc wl0 = a.wins[a.n_windows - 1];
		const int n_pairs_all = wl0.pair_off + wl0.n_pairs;
		// (a lone window's plan is on its critical path: the pairs' source counts - two dependent global loads - are read once, not per attempt)
		const bool one_each = n_pairs_all <= 1024;
		const int n_mine = (one_each && tid < n_pairs_all) ? a.nsrc[a.pair_src_slot[tid]] : 0;
		for (int attempt = 0; attempt < 4; attempt++) {
			if (tid == 0) s_total = 0;
			__syncthreads();
			int mine = 0;
			if (one_each) { int per; mine = chunks_for(n_mine, chunk, per); }
			else for (int q = tid; q < n_pairs_all; q += 1024) { int per; mine += chunks_for(a.nsrc[a.pair_src_slot[q]], chunk, per); }
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
			if ((tid & 31) == 0 && mine) atomicAdd(&s_total, mine);
			__syncthreads();
			const int total = s_total + kSparseTiles * a.n_windows;      // + the sparse tiles of every window
			__syncthreads();
			if (total <= a.grid_ctas || total > 2 * a.grid_ctas) break;
			chunk = (int)(((long long)chunk * total / a.grid_ctas + 63) / 32 * 32);     // ~(total/grid) x larger, rounded up to warps
		}
	}
	for (int base = 0; base < a.n_windows; base += 1024) {
		const int nw = min(1024, a.n_windows - base);
		s_cnt[tid] = 0;
		__syncthreads();
		// (1) chunk count of every (window, pair) of this block of windows, in parallel
		const int first_pair = a.wins[base].pair_off;
		const WinDesc wl = a.wins[base + nw - 1];
		const int end_pair = wl.pair_off + wl.n_pairs;
		unsigned long long px = 0;
		for (int q = first_pair + tid; q < end_pair; q += 1024) {
			const int lo = a.pair_win[q];
			const int n = a.nsrc[a.pair_src_slot[q]];
			int per;
			const int nch = chunks_for(n, chunk, per);
			a.pair_ntile[q] = nch;
			atomicAdd(&s_cnt[lo - base], nch);
			px += (unsigned long long)n;
		}
		// one 64-bit shared atomic per warp: 1024 threads hammering one 64-bit shared word (a CAS loop in hardware) cost 30 us here
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) px += __shfl_xor_sync(0xffffffffu, px, o);
		if ((tid & 31) == 0 && px) atomicAdd(&s_px, px);
		__syncthreads();
		PROF_T(1);
		// (2) exclusive scan of the per-window tile counts: the dense tiles plus ONE sparse tile (the window's moment sums), first in line
		int cnt = 0;
		if (tid < nw) { cnt = s_cnt[tid] + kSparseTiles; a.tiles_done[base + tid] = 0; a.iter_done[base + tid] = 0; }
		{   // inclusive scan over the 1024 per-window counts: warp shuffles + one pass over the 32 warp totals
			__shared__ int s_wtot[32];
			int incl = cnt;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += u; }
			if ((tid & 31) == 31) s_wtot[tid >> 5] = incl;
			__syncthreads();
			if (tid < 32) {
				int t = s_wtot[tid];
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, t, o); if (tid >= o) t += u; }
				s_wtot[tid] = t;
			}
			__syncthreads();
			s_scan[tid] = incl + ((tid >> 5) ? s_wtot[(tid >> 5) - 1] : 0);
			__syncthreads();
		}
		PROF_T(2);
		const int excl = s_scan[tid] - cnt + s_carry;
		const bool fits = (s_carry + s_scan[1023] <= a.max_tiles);
		// (3) per-window prefix over its pairs: one warp per window, lanes over pairs
		if (tid < nw) {
			wins_rw[base + tid].tile_off = excl;
			wins_rw[base + tid].n_tiles = cnt;
			s_cnt[tid] = excl;      // reuse: window -> tile_off for step (4)
		}
		__syncthreads();
		for (int wl = (tid >> 5); wl < nw; wl += 32) {
			const int w = base + wl, lane = tid & 31;
			const int poff = a.wins[w].pair_off, np = a.wins[w].n_pairs;
			int carry = 0;
			for (int p0 = 0; p0 < np; p0 += 32) {
				const int p = p0 + lane;
				const int v = (p < np) ? a.pair_ntile[poff + p] : 0;
				int incl = v;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
				if (p < np) a.pair_tile0[poff + p] = kSparseTiles + carry + incl - v;      // the first slots of the window are its sparse tiles
				carry += __shfl_sync(0xffffffffu, incl, 31);
			}
			if (lane < kSparseTiles && fits) {      // start / count of a sparse tile = its share: part `start` of `count`
				Tile tl; tl.win = w; tl.pair = -2; tl.start = lane; tl.count = kSparseTiles; tl.tgt_slot = tl.src_slot = 0; tl.n_tiles_win = wins_rw[w].n_tiles;
				tl.pad = 0; tl.src = tl.tex = nullptr; a.tiles[s_cnt[wl] + lane] = tl;
			}
		}
		__syncthreads();
		PROF_T(3);
		// (4) tile records, one (window, pair) per thread
		if (fits) {
			for (int q = first_pair + tid; q < end_pair; q += 1024) {
				const int lo = a.pair_win[q];
				const int n = a.nsrc[a.pair_src_slot[q]];
				const int pair_off_w = a.wins[lo].pair_off;
				int per;
				const int nch = chunks_for(n, chunk, per);
				int t = s_cnt[lo - base] + a.pair_tile0[q];
				Tile tl; tl.win = lo; tl.pair = q - pair_off_w; tl.pad = 0;
				tl.src_slot = a.pair_src_slot[q]; tl.tgt_slot = a.wins[lo].frame_off + (int)a.pairs[q].x;
				tl.n_tiles_win = wins_rw[lo].n_tiles;
				tl.src = a.src_tab[tl.src_slot]; tl.tex = a.texel_tab[tl.tgt_slot];
				for (int c = 0; c < nch; c++) { tl.start = c * per; tl.count = min(per, n - c * per); a.tiles[t++] = tl; }
			}
		}
		__syncthreads();
		PROF_T(4);
		if (tid == 0) s_carry += s_scan[1023];
		__syncthreads();
	}
	if (tid == 0) {
		*a.n_tiles_total = (s_carry <= a.max_tiles && (long long)s_carry * a.prm.num_iter_outer < 0x7fffff00ll) ? s_carry : -1;   // -1 => capacity error reported by the host
		*a.n_src_px = (long long)s_px;
	}
	PROF_T(5);
	if (a.prof && tid == 0) prof_emit(a, prf);
}

// ------------------------------------------------------------------------------------------------ tile (dense term)
// Evaluate one chunk of compacted source pixels of pair (tgt i, src j).  Accumulates, in the TARGET camera frame,
// S' = sum w g' g'^T (upper triangle, 21) and b' = sum w g' r (6) with g' = (n_tgt, p' x n_tgt); the tail maps them to
// the model frame with the 6x6 adjoint of T_i.
struct TileAcc { float v[kTileVals]; };

// Software-pipelined over a thread's pixels (stride kThreads): iteration i issues the source load of pixel i+2 while it
// evaluates pixel i (an explicit L1 prefetch of the taps of pixel i+1 was measured SLOWER on B200 and was removed).
template <int NT> __device__ __forceinline__ void tile_pixels(const SolveArgs& a, const WinDesc& wd, const float4* __restrict__ src,
                                            const float4* __restrict__ tex, const float* __restrict__ sM, int start, int count, TileAcc& acc) {
	const float m00 = sM[0], m01 = sM[1], m02 = sM[2], m03 = sM[3];
	const float m10 = sM[4], m11 = sM[5], m12 = sM[6], m13 = sM[7];
	const float m20 = sM[8], m21 = sM[9], m22 = sM[10], m23 = sM[11];
	const float fx = wd.fx, fy = wd.fy, cx = wd.cx, cy = wd.cy;
	const unsigned W = (unsigned)wd.w, Hh = (unsigned)wd.h;
	const float dmin = a.prm.depth_min, dmax = a.prm.depth_max, cos_t = a.prm.dense_cos_normal_thresh;
	const float delta = a.prm.robust_delta, wdense = a.prm.w_dense;
	// `sqrtf(d2) <= dist_t` evaluated as `d2 <= d2max`, d2max = the largest float whose correctly rounded square root is <= dist_t
	// (same truth value for every float, no square root in the loop; found once on the host, make_args)
	const float d2max = a.d2max;
	const float4* sp = src + 2 * (size_t)start;
	int k = threadIdx.x;
	F8 cc, nn, ff;      // current, next, next-next source records
	cc.lo = cc.hi = nn.lo = nn.hi = ff.lo = ff.hi = make_float4(0.f, 0.f, 0.f, 0.f);
	if (k < count) cc = ldg256(sp + 2 * k);
	if (k + NT < count) nn = ldg256(sp + 2 * (k + NT));
	for (; k < count; k += NT) {
		const int k2 = k + 2 * NT;
		if (k2 < count) ff = ldg256(sp + 2 * k2);
		const float4 s0 = cc.lo, s1 = cc.hi;
		cc = nn; nn = ff;
		const float px = s0.x, py = s0.y, pz = s0.z, nx = s0.w, ny = s1.x, nz = s1.y;
		// camPosSrcToTgt = transform * camPosSrc ; nrmj = transform(3x3) * n  (w of the normal is 0)
		const float tx = m00 * px + m01 * py + m02 * pz + m03;
		const float ty = m10 * px + m11 * py + m12 * pz + m13;
		const float tz = m20 * px + m21 * py + m22 * pz + m23;
		// cameraToDepth, CUDACameraUtil.h:9-14.  One reciprocal serves both quotients (the reference itself is built with
		// -use_fast_math: its divisions are rcp.approx too; either way the result differs from the exact quotient by an ulp or two)
		const float itz = rcp_fast(tz);
		const float sx = (tx * fx) * itz + cx, sy = (ty * fy) * itz + cy;
		const int ix = (int)roundf(sx), iy = (int)roundf(sy);
		if (!(ix >= 0 && iy >= 0 && ix < (int)W && iy < (int)Hh)) continue;
		// bilinearInterpolationFloat4 on camera-space points AND normals (same taps, same weights)
		const float fx0 = floorf(sx), fy0 = floorf(sy);
		const int x0 = (int)fx0, y0 = (int)fy0;
		const float al = sx - fx0, be = sy - fy0;
		float a0[3] = { 0.f, 0.f, 0.f }, b0[3] = { 0.f, 0.f, 0.f }, a1[3] = { 0.f, 0.f, 0.f }, b1[3] = { 0.f, 0.f, 0.f };
		float w0 = 0.f, w1 = 0.f;
		const bool inx0 = (unsigned)x0 < W, inx1 = (unsigned)(x0 + 1) < W, iny0 = (unsigned)y0 < Hh, iny1 = (unsigned)(y0 + 1) < Hh;
		if (iny0) {
			const float4* row = tex + 2 * ((size_t)y0 * W);
			if (inx0) { const F8 t8 = ldg256(row + 2 * x0); const float4 cp = t8.lo; const float2 nr = make_float2(t8.hi.x, t8.hi.y); const float wt = 1.0f - al;
				a0[0] += wt * cp.x; a0[1] += wt * cp.y; a0[2] += wt * cp.z; b0[0] += wt * cp.w; b0[1] += wt * nr.x; b0[2] += wt * nr.y; w0 += wt; }
			if (inx1) { const F8 t8 = ldg256(row + 2 * (x0 + 1)); const float4 cp = t8.lo; const float2 nr = make_float2(t8.hi.x, t8.hi.y); const float wt = al;
				a0[0] += wt * cp.x; a0[1] += wt * cp.y; a0[2] += wt * cp.z; b0[0] += wt * cp.w; b0[1] += wt * nr.x; b0[2] += wt * nr.y; w0 += wt; }
		}
		if (iny1) {
			const float4* row = tex + 2 * ((size_t)(y0 + 1) * W);
			if (inx0) { const F8 t8 = ldg256(row + 2 * x0); const float4 cp = t8.lo; const float2 nr = make_float2(t8.hi.x, t8.hi.y); const float wt = 1.0f - al;
				a1[0] += wt * cp.x; a1[1] += wt * cp.y; a1[2] += wt * cp.z; b1[0] += wt * cp.w; b1[1] += wt * nr.x; b1[2] += wt * nr.y; w1 += wt; }
			if (inx1) { const F8 t8 = ldg256(row + 2 * (x0 + 1)); const float4 cp = t8.lo; const float2 nr = make_float2(t8.hi.x, t8.hi.y); const float wt = al;
				a1[0] += wt * cp.x; a1[1] += wt * cp.y; a1[2] += wt * cp.z; b1[0] += wt * cp.w; b1[1] += wt * nr.x; b1[2] += wt * nr.y; w1 += wt; }
		}
		float ww = 0.f, cxs = 0.f, cys = 0.f, czs = 0.f, nxs = 0.f, nys = 0.f, nzs = 0.f;
		if (w0 > 0.f) { const float r = (1.0f - be) * rcp_fast(w0); cxs += r * a0[0]; cys += r * a0[1]; czs += r * a0[2]; nxs += r * b0[0]; nys += r * b0[1]; nzs += r * b0[2]; ww += (1.0f - be); }
		if (w1 > 0.f) { const float r = be * rcp_fast(w1); cxs += r * a1[0]; cys += r * a1[1]; czs += r * a1[2]; nxs += r * b1[0]; nys += r * b1[1]; nzs += r * b1[2]; ww += be; }
		if (!(ww > 0.f)) continue;
		const float rw = rcp_fast(ww);
		const float qx = cxs * rw, qy = cys * rw, qz = czs * rw;        // camPosTgt
		if (!(qz > dmin && qz < dmax)) continue;
		const float tnx = nxs * rw, tny = nys * rw, tnz = nzs * rw;     // normalTgt
		const float rnx = m00 * nx + m01 * ny + m02 * nz;
		const float rny = m10 * nx + m11 * ny + m12 * nz;
		const float rnz = m20 * nx + m21 * ny + m22 * nz;
		const float dx = tx - qx, dy = ty - qy, dz = tz - qz;
		const float dist2 = dx * dx + dy * dy + dz * dz;
		const float dn = rnx * tnx + rny * tny + rnz * tnz;
		if (!(dn >= cos_t && dist2 <= d2max)) continue;
		const float res = -(dx * tnx + dy * tny + dz * tnz);            // dot(camPosTgt - camPosSrcToTgt, normalTgt)
		const float wgt = wdense * huber_w_fast(res * res, delta);
		float g[6];
		g[0] = tnx; g[1] = tny; g[2] = tnz;
		g[3] = ty * tnz - tz * tny; g[4] = tz * tnx - tx * tnz; g[5] = tx * tny - ty * tnx;   // p' x n_tgt
		int e = 0;
#pragma unroll
		for (int r = 0; r < 6; r++) {
			const float wg = wgt * g[r];
#pragma unroll
			for (int c = r; c < 6; c++) acc.v[e++] += wg * g[c];
		}
#pragma unroll
		for (int r = 0; r < 6; r++) acc.v[21 + r] += wgt * g[r] * res;
		acc.v[27] += 1.0f;
	}
}

// ------------------------------------------------------------------------------------------------ tail (per window, per GN iteration)
__constant__ unsigned char c_sym_r[21] = { 0,0,0,0,0,0, 1,1,1,1,1, 2,2,2,2, 3,3,3, 4,4, 5 };
__constant__ unsigned char c_sym_c[21] = { 0,1,2,3,4,5, 1,2,3,4,5, 2,3,4,5, 3,4,5, 4,5, 5 };
__constant__ unsigned char c_sym_idx[36] = { 0,1,2,3,4,5, 1,6,7,8,9,10, 2,7,11,12,13,14, 3,8,12,15,16,17, 4,9,13,16,18,19, 5,10,14,17,19,20 };

struct TailSmem {
	float* T;      // [N][12]
	float* A;      // [dimp][ld]
	float* rhs; float* Minv; float* r; float* z; float* p; float* Ap; float* delta;   // [dimp] each
	float* pairW;    // [P][28] per-pair sums over the pair's tiles (model frame; 27 used + count)
	float* grp;      // [G][44]
	float* fS;       // [N][20] per-frame sparse sums: n, S(3), Q(6), grad rot(3), grad trans(3), Pw(3), Pn
	float* fD;       // [N][28] per-frame dense sums: S(21) + signed b(6)
	int* gi; int* gj; int* gstart;     // [G], [G], [G+1]
	int* pt; int* ps; int* pt0; int* pnt;   // [P] each: pair target, source, first tile, #tiles
	int* fg_start; int* fp_start; int* fg_items; int* fp_items;   // per-frame membership CSR ([N+1], [N+1], [2G], [2P])
	int dimp, ld;
};
static constexpr int kFS = 20, kFD = 28;
// (grp_in_smem = false: big windows - 30 frames, 435 groups - leave the 44 moment sums per group in global memory and read them
// through L2; with them the tail of such a window would need 279 KB of shared memory)
// Leading dimension of the system matrix: the PCG reads a whole row per THREAD with 16-byte loads, which is conflict-free when the rows
// of eight consecutive threads start in eight different 4-bank groups: ld = 4 (mod 8), ld >= dimp rounded up to a multiple of 4.
__host__ __device__ inline int tail_ld(int dimp) { int ld = (dimp + 3) & ~3; if ((ld & 7) != 4) ld += 4; return ld; }
__host__ __device__ inline size_t tail_smem_floats(int N, int P, int G, bool grp_in_smem = true) {
	const int dimp = 6 * (N - 1), ld = tail_ld(dimp), dimp4 = (dimp + 3) & ~3;
	return (size_t)N * 12 + (size_t)dimp * ld + 7 * (size_t)dimp4 + (size_t)P * kTileVals + (grp_in_smem ? (size_t)G * kGrpVals : 0) +
	       (size_t)N * (kFS + kFD) + (size_t)(3 * G + 1 + 4 * P) + (size_t)(2 * (N + 1) + 2 * G + 2 * P) + 16;
}
__device__ inline void tail_carve(float* base, int N, int P, int G, TailSmem& s, bool grp_in_smem) {
	s.dimp = 6 * (N - 1); s.ld = tail_ld(s.dimp);
	const int dimp4 = (s.dimp + 3) & ~3;      // every vector starts 16-byte aligned (N * 12 and dimp * ld are multiples of 4 floats)
	float* q = base;
	s.T = q; q += N * 12;
	s.A = q; q += s.dimp * s.ld;
	s.rhs = q; q += dimp4; s.Minv = q; q += dimp4; s.r = q; q += dimp4; s.z = q; q += dimp4;
	s.p = q; q += dimp4; s.Ap = q; q += dimp4; s.delta = q; q += dimp4;
	s.pairW = q; q += P * kTileVals;
	s.grp = q; if (grp_in_smem) q += G * kGrpVals;
	s.fS = q; q += N * kFS; s.fD = q; q += N * kFD;
	int* iq = reinterpret_cast<int*>(q);
	s.gi = iq; iq += G; s.gj = iq; iq += G; s.gstart = iq; iq += G + 1;
	s.pt = iq; iq += P; s.ps = iq; iq += P; s.pt0 = iq; iq += P; s.pnt = iq; iq += P;
	s.fg_start = iq; iq += N + 1; s.fp_start = iq; iq += N + 1; s.fg_items = iq; iq += 2 * G; s.fp_items = iq; iq += 2 * P;
}
// index of (r,c), r<=c, in the packed upper triangle of a symmetric 6x6
__device__ __forceinline__ int sym6(int r, int c) { if (r > c) { const int t = r; r = c; c = t; } return r * 6 - (r * (r - 1)) / 2 + (c - r); }
__device__ __forceinline__ float skew(const float* v, int a, int b) {   // [v]x (a,b)
	if (a == b) return 0.f;
	const int k = 3 - a - b;                                // the remaining axis
	const float s = ((b - a + 3) % 3 == 1) ? -1.f : 1.f;    // (0,1):-vz (1,2):-vx (2,0):-vy ; transposed: +
	return s * v[k];
}
__device__ __forceinline__ void unpack_sym(int e, int& r, int& c) {
	r = 0;
	while (e >= 6 - r) { e -= 6 - r; r++; }
	c = r + e;
}

// Sparse moment sums of one window for the CURRENT poses: 8 lanes per (i,j) group of correspondences, loads one iteration
// ahead; 44 sums per group go to a.grp_sums.  This is the window's SPARSE tile: first in the window's slice of the queue, it runs
// on whichever CTA claims it while the dense tiles are in flight, so the window's tail does not wait for it.
template <int NT> __device__ void sparse_sums(const SolveArgs& a, int w, int part, int nparts) {
	const WinDesc wd0 = a.wins[w];
	const WinSparse ws = a.wsp[w];
	const int G = ws.n_groups, tid = threadIdx.x;
	const int sub = tid >> 3, sl = tid & 7, nsub = NT >> 3;
	const int g_lo = (int)(((long long)G * part) / nparts), g_hi = (int)(((long long)G * (part + 1)) / nparts);      // this tile's share of the groups
	for (int g0 = g_lo; g0 < g_hi; g0 += nsub) {
		const int g = g0 + sub;
		float m[kGrpVals];
#pragma unroll
		for (int k = 0; k < kGrpVals; k++) m[k] = 0.f;
		if (g < g_hi) {
			const int c0 = a.grp_start[ws.grp_off + w + g], c1 = a.grp_start[ws.grp_off + w + g + 1];
			const float* Ti = a.T + (size_t)(wd0.frame_off + a.grp_i[ws.grp_off + g]) * 12; const float* Tj = a.T + (size_t)(wd0.frame_off + a.grp_j[ws.grp_off + g]) * 12;
			const float t00 = __ldcg(Ti + 0), t01 = __ldcg(Ti + 1), t02 = __ldcg(Ti + 2), t03 = __ldcg(Ti + 3), t10 = __ldcg(Ti + 4), t11 = __ldcg(Ti + 5), t12 = __ldcg(Ti + 6), t13 = __ldcg(Ti + 7),
			            t20 = __ldcg(Ti + 8), t21 = __ldcg(Ti + 9), t22 = __ldcg(Ti + 10), t23 = __ldcg(Ti + 11);
			const float u00 = __ldcg(Tj + 0), u01 = __ldcg(Tj + 1), u02 = __ldcg(Tj + 2), u03 = __ldcg(Tj + 3), u10 = __ldcg(Tj + 4), u11 = __ldcg(Tj + 5), u12 = __ldcg(Tj + 6), u13 = __ldcg(Tj + 7),
			            u20 = __ldcg(Tj + 8), u21 = __ldcg(Tj + 9), u22 = __ldcg(Tj + 10), u23 = __ldcg(Tj + 11);
			int c = c0 + sl;
			float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
			if (c < c1) { const float4* e4 = reinterpret_cast<const float4*>(a.corr + ws.corr_off + c); lo = __ldg(e4); hi = __ldg(e4 + 1); }
			while (c < c1) {
				const int cn = c + 8;
				float4 lo_n = lo, hi_n = hi;
				if (cn < c1) { const float4* e4 = reinterpret_cast<const float4*>(a.corr + ws.corr_off + cn); lo_n = __ldg(e4); hi_n = __ldg(e4 + 1); }
				const float pix = lo.z, piy = lo.w, piz = hi.x, pjx = hi.y, pjy = hi.z, pjz = hi.w;   // {i, j, pi.x, pi.y}, {pi.z, pj.x, pj.y, pj.z}
				const V3 q = mk(t00 * pix + t01 * piy + t02 * piz + t03, t10 * pix + t11 * piy + t12 * piz + t13, t20 * pix + t21 * piy + t22 * piz + t23);
				const V3 sp = mk(u00 * pjx + u01 * pjy + u02 * pjz + u03, u10 * pjx + u11 * pjy + u12 * pjz + u13, u20 * pjx + u21 * pjy + u22 * pjz + u23);
				const V3 rr = q - sp;
				const float rho = huber_w(dot(rr, rr), a.prm.robust_delta);
				m[0] += 1.f;
				m[1] += q.x; m[2] += q.y; m[3] += q.z;
				m[4] += sp.x; m[5] += sp.y; m[6] += sp.z;
				m[7] += q.x * q.x; m[8] += q.x * q.y; m[9] += q.x * q.z; m[10] += q.y * q.y; m[11] += q.y * q.z; m[12] += q.z * q.z;
				m[13] += sp.x * sp.x; m[14] += sp.x * sp.y; m[15] += sp.x * sp.z; m[16] += sp.y * sp.y; m[17] += sp.y * sp.z; m[18] += sp.z * sp.z;
				m[19] += sp.x * q.x; m[20] += sp.x * q.y; m[21] += sp.x * q.z;      // Qsq[a][b] = sum s_a q_b
				m[22] += sp.y * q.x; m[23] += sp.y * q.y; m[24] += sp.y * q.z;
				m[25] += sp.z * q.x; m[26] += sp.z * q.y; m[27] += sp.z * q.z;
				const V3 gq = cross(q, rr), gs = cross(sp, rr);
				m[28] += rho * gq.x; m[29] += rho * gq.y; m[30] += rho * gq.z;
				m[31] += rho * gs.x; m[32] += rho * gs.y; m[33] += rho * gs.z;
				m[34] += rho * rr.x; m[35] += rho * rr.y; m[36] += rho * rr.z;
				m[37] += rho * q.x * q.x; m[38] += rho * q.y * q.y; m[39] += rho * q.z * q.z;
				m[40] += rho * sp.x * sp.x; m[41] += rho * sp.y * sp.y; m[42] += rho * sp.z * sp.z;
				m[43] += rho;
				lo = lo_n; hi = hi_n; c = cn;
			}
		}
#pragma unroll
		for (int k = 0; k < kGrpVals; k++) {
			float v = m[k];
			v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 1);
			if (g < g_hi && sl == 0) __stcg(a.grp_sums + (size_t)(ws.grp_off + g) * kGrpVals + k, v);
		}
	}
}

template <int NT> __device__ void window_tail(const SolveArgs& a, const WinDesc& wd_in, int w, int it, float* smem_base, int* p2_claim) {
	WinDesc wd = wd_in;
	{ const WinSparse ws = a.wsp[w]; wd.n_corr = ws.n_corr; wd.n_groups = ws.n_groups; wd.corr_off = ws.corr_off; wd.grp_off = ws.grp_off; wd.mem_off = ws.mem_off; wd.unique_blocks = ws.unique_blocks; }
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	static_assert(NT >= 6 * (kMaxFrames - 1), "the PCG runs one thread per unknown");
	const int N = wd.n_frames, P = wd.n_pairs, G = wd.n_groups;
	TailSmem s;
	const bool gsm = a.grp_in_smem != 0;
	tail_carve(smem_base, N, P, G, s, gsm);
	const float* grp_g = a.grp_sums + (size_t)wd.grp_off * kGrpVals;      // the same sums in global memory (written with st.cg by the sparse tile)
	const int dimp = s.dimp, ld = s.ld;
	const float wS = a.prm.w_sparse;
	const bool use_dense = a.prm.w_dense > 0.f && P > 0;
	const bool dbg = a.dbg_JtJ && it == a.prm.num_iter_outer - 1;
	ProfRec prf; prf.kind_cta = (1ll << 32) | blockIdx.x; prf.tile_win = ((long long)it << 32) | w;
	PROF_T(0);

	// ---- P0 || P2.  P0: everything the later phases index repeatedly goes to shared memory once (poses, group / pair tables, membership
	//      CSR, the groups' moment sums): ALL of it as asynchronous copies issued back to back - one L2 round trip instead of one per
	//      table - while the same warps clear the system matrix.  P2: per-pair sums over the pair's tiles (already in the model frame; the
	//      tile epilogue applied X S' X^T): a pair's tiles are consecutive and summed in order (deterministic), 4 tiles x RI items per
	//      thread in flight.  The second half of the CTA starts on P2 at once (it needs nothing from P0: the tile ranges come straight from
	//      global memory), the first half joins when its copies are issued; work is handed out in warp-sized blocks by a shared counter.
	int& s_p2_claim = *p2_claim;
	constexpr int RI = 5;
	const int n_p2 = use_dense ? P * kTileVals : 0;
	auto p2_work = [&]() {
		for (;;) {
			int base = 0;
			if (lane == 0) base = atomicAdd(&s_p2_claim, 32 * RI);
			base = __shfl_sync(0xffffffffu, base, 0);
			if (base >= n_p2) break;
			float v[RI]; const float* src[RI]; int nt[RI]; int mx = 0;
#pragma unroll
			for (int r = 0; r < RI; r++) {
				const int k = base + lane + 32 * r;
				v[r] = 0.f; nt[r] = 0; src[r] = a.partial;
				if (k < n_p2) {
					const int p = k / kTileVals, e = k - p * kTileVals;
					nt[r] = __ldg(a.pair_ntile + wd.pair_off + p);
					src[r] = a.partial + (size_t)(wd.tile_off + __ldg(a.pair_tile0 + wd.pair_off + p)) * kTileVals + e;
					mx = max(mx, nt[r]);
				}
			}
			for (int c = 0; c < mx; c += 4) {      // four tiles x RI items in flight per thread; the adds keep the tile order (x + 0.f == x)
				float t[4][RI];
#pragma unroll
				for (int u = 0; u < 4; u++) {
#pragma unroll
					for (int r = 0; r < RI; r++) t[u][r] = (c + u < nt[r]) ? __ldcg(src[r] + (size_t)(c + u) * kTileVals) : 0.f;
				}
#pragma unroll
				for (int u = 0; u < 4; u++) {
#pragma unroll
					for (int r = 0; r < RI; r++) v[r] += t[u][r];
				}
			}
#pragma unroll
			for (int r = 0; r < RI; r++) {
				const int k = base + lane + 32 * r;
				if (k < n_p2) {
					s.pairW[k] = v[r];
					if (a.dbg_cnt && (k % kTileVals) == 27 && it == a.prm.num_iter_outer - 1) a.dbg_cnt[(size_t)w * a.dbg_cnt_stride + k / kTileVals] = v[r];
				}
			}
		}
	};
	if (tid >= NT / 2) p2_work();
	else {
		constexpr int H = NT / 2;
		{   // poses of this iteration: written by this window's previous tail with st.cg -> read from L2
			const float4* Tg = reinterpret_cast<const float4*>(a.T + (size_t)wd.frame_off * 12);
			for (int k = tid; k < N * 3; k += H) cp_async16_cg(reinterpret_cast<float4*>(s.T) + k, Tg + k);
		}
		if (gsm) {  // the sparse moment sums were computed by sparse_sums() while the window's dense tiles were still running (st.cg)
			const float4* Gg = reinterpret_cast<const float4*>(grp_g);
			for (int k = tid; k < G * (kGrpVals / 4); k += H) cp_async16_cg(reinterpret_cast<float4*>(s.grp) + k, Gg + k);
		}
		for (int k = tid; k < G; k += H) { cp_async4(s.gi + k, a.grp_i + wd.grp_off + k); cp_async4(s.gj + k, a.grp_j + wd.grp_off + k); }
		for (int k = tid; k <= G; k += H) cp_async4(s.gstart + k, a.grp_start + wd.grp_off + w + k);
		for (int k = tid; k < P; k += H) {
			const uint2* pr = a.pairs + wd.pair_off + k;
			cp_async4(s.pt + k, &pr->x); cp_async4(s.ps + k, &pr->y);
			cp_async4(s.pt0 + k, a.pair_tile0 + wd.pair_off + k); cp_async4(s.pnt + k, a.pair_ntile + wd.pair_off + k);
		}
		{
			const int* mem = a.mem + wd.mem_off;     // fg_start[N+1] fp_start[N+1] fg_items[2G] fp_items[2P], contiguous like the smem copy
			const int nmem = 2 * (N + 1) + 2 * G + 2 * P;
			for (int k = tid; k < nmem; k += H) cp_async4(s.fg_start + k, mem + k);
		}
		{
			float4* A4 = reinterpret_cast<float4*>(s.A);      // (dimp * ld is a multiple of four floats)
			for (int k = tid; k < (dimp * ld) >> 2; k += H) A4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
		}
		p2_work();
		asm volatile("cp.async.wait_all;" ::: "memory");
	}
	__syncthreads();
	if (tid == 0) s_p2_claim = 0;      // for this CTA's next tail (nobody touches the counter again before several barriers have passed)
	PROF_T(1);
	PROF_T(2);
	__syncthreads();
	PROF_T(3);
	// ---- P3a: per-frame gathers (no atomics, fixed order).  Every output entry e reads one moment from the q-side or the s-side of each
	//      group touching the frame (offset / sign chosen once, outside the loop).  The loop over the frame's groups is a chain of two
	//      dependent shared-memory loads per step (item -> moment): four steps are loaded at a time, the adds keep the list order.
	for (int k = tid; k < N * (kFS + kFD); k += NT) {
		const int f = k / (kFS + kFD), e = k - f * (kFS + kFD);
		float acc = 0.f;
		if (e < kFS) {
			// fS layout: 0 n | 1-3 S | 4-9 Q | 10-12 grad rot | 13-15 grad trans | 16-18 Pw | 19 Pn
			int oq, os; float sg = 1.f;
			if (e == 0) { oq = 0; os = 0; }
			else if (e < 4) { oq = e; os = 3 + e; }
			else if (e < 10) { oq = 3 + e; os = 9 + e; }
			else if (e < 13) { oq = 18 + e; os = 21 + e; sg = -1.f; }
			else if (e < 16) { oq = 21 + e; os = 21 + e; sg = -1.f; }
			else if (e < 19) { oq = 21 + e; os = 24 + e; }
			else { oq = 43; os = 43; }
			const int q1 = s.fg_start[f + 1];
			for (int q = s.fg_start[f]; q < q1; q += 4) {
				int item[4]; float mv[4];
#pragma unroll
				for (int u = 0; u < 4; u++) item[u] = s.fg_items[min(q + u, q1 - 1)];
#pragma unroll
				for (int u = 0; u < 4; u++) {
					const int mo = (item[u] & 0xffff) * kGrpVals + ((item[u] >> 16) ? os : oq);      // role 0: this frame is the group's i (q side); 1: j (s side)
					mv[u] = gsm ? s.grp[mo] : __ldcg(grp_g + mo);
				}
#pragma unroll
				for (int u = 0; u < 4; u++) if (q + u < q1) acc += (item[u] >> 16) ? sg * mv[u] : mv[u];
			}
			s.fS[f * kFS + e] = acc;
		} else {
			const int d = e - kFS;
			if (use_dense && d < 27) {
				const float sgs = (d >= 21) ? -1.f : 1.f;      // Jtr_i += b, Jtr_j -= b ; S adds to both diagonal blocks
				const int q1 = s.fp_start[f + 1];
				for (int q = s.fp_start[f]; q < q1; q += 4) {
					int item[4]; float v[4];
#pragma unroll
					for (int u = 0; u < 4; u++) item[u] = s.fp_items[min(q + u, q1 - 1)];
#pragma unroll
					for (int u = 0; u < 4; u++) v[u] = s.pairW[(item[u] & 0xffff) * kTileVals + d];
#pragma unroll
					for (int u = 0; u < 4; u++) if (q + u < q1) acc += (item[u] >> 16) ? sgs * v[u] : v[u];             // role 0: target, 1: source
				}
			}
			s.fD[f * kFD + d] = acc;
		}
	}
	__syncthreads();
	PROF_T(4);
	// ---- P3b: diagonal blocks, right-hand side, Jacobi preconditioner from the per-frame sums
	{
		const int per = 21 + 6 + 6;
		for (int k = tid; k < (N - 1) * per; k += NT) {
			const int f = 1 + k / per, e = k % per;
			const int base = (f - 1) * 6;
			const float* F = s.fS + f * kFS;
			const float* D = s.fD + f * kFD;
			if (e < 21) {
				int r, c;
				unpack_sym(e, r, c);
				float sp;
				if (r < 3 && c < 3) sp = (r == c) ? F[0] : 0.f;                                   // TT = n I
				else if (r < 3) sp = -skew(F + 1, r, c - 3);                                       // TR = -[S]x
				else { const int ra = r - 3, cb = c - 3; const float tr = F[4] + F[7] + F[9];
					const int qi = (ra == 0) ? cb : (ra == 1 ? 2 + cb : 5);                          // packed (ra,cb), ra<=cb
					sp = ((ra == cb) ? tr : 0.f) - F[4 + qi]; }
				const float dn = D[e];
				const float v = wS * sp + dn;
				s.A[(base + r) * ld + base + c] = v;
				s.A[(base + c) * ld + base + r] = v;
				if (dbg) { float* Dg = a.dbg_JtJ + (size_t)w * a.dbg_stride * a.dbg_stride; const int dim = 6 * N;
					Dg[(f * 6 + r) * dim + f * 6 + c] = dn; Dg[(f * 6 + c) * dim + f * 6 + r] = dn; }
			} else if (e < 27) {
				const int r = e - 21;
				const float gs = (r < 3) ? F[13 + r] : F[10 + (r - 3)];
				const float dn = D[21 + r];
				s.rhs[base + r] = -wS * gs - dn;
				if (dbg && a.dbg_Jtr) a.dbg_Jtr[(size_t)w * a.dbg_stride + f * 6 + r] = dn;
			} else {
				const int r = e - 27;
				float pc;
				if (r < 3) pc = F[19];
				else { const int ax = r - 3; pc = F[16 + (ax + 1) % 3] + F[16 + (ax + 2) % 3]; }
				s.Minv[base + r] = (pc > kEps) ? 1.0f / pc : 1.0f;
			}
		}
	}
	__syncthreads();
	PROF_T(5);
	// ---- P4a: sparse cross blocks (i,j): J_i^T J_j, both frames free.  One thread per (group, 3x3 sub-block): uniform code per
	//      thread, 9 entries each.  Sub-blocks: 0 = TT (-n I), 1 = TR ([sum s]x), 2 = RT (-[sum q]x), 3 = RR (-((q.s) I - s q^T)).
	const bool uniq = wd.unique_blocks != 0;
	for (int k = tid; k < G * 4; k += NT) {
		const int g = k >> 2, sb = k & 3;
		const int gi = s.gi[g], gj = s.gj[g];
		if (gi < 1 || gj < 1 || gi == gj) continue;
		float ml[28];
#pragma unroll
		for (int q = 0; q < 28; q++) ml[q] = gsm ? s.grp[g * kGrpVals + q] : __ldcg(grp_g + (size_t)g * kGrpVals + q);
		const float* m = ml;
		const int r0 = (sb >> 1) * 3, c0 = (sb & 1) * 3;
		const float tr = m[19] + m[23] + m[27];
#pragma unroll
		for (int a3 = 0; a3 < 3; a3++) {
#pragma unroll
			for (int b3 = 0; b3 < 3; b3++) {
				float v;
				if (sb == 0) v = (a3 == b3) ? -m[0] : 0.f;
				else if (sb == 1) v = skew(m + 4, a3, b3);
				else if (sb == 2) v = -skew(m + 1, a3, b3);
				else v = -(((a3 == b3) ? tr : 0.f) - m[19 + a3 * 3 + b3]);
				v *= wS;
				float* p1 = &s.A[((gi - 1) * 6 + r0 + a3) * ld + (gj - 1) * 6 + c0 + b3];
				float* p2 = &s.A[((gj - 1) * 6 + c0 + b3) * ld + (gi - 1) * 6 + r0 + a3];
				if (uniq) { *p1 = v; *p2 = v; } else { atomicAdd(p1, v); atomicAdd(p2, v); }
			}
		}
	}
	__syncthreads();
	// ---- P4b: dense cross blocks: -S for pairs whose cross block survives FlipJtJ (target < source) or all if !compat
	if (use_dense) {
		for (int k = tid; k < P * 6; k += NT) {
			const int p = k / 6, r = k - p * 6;
			const int ti = s.pt[p], sj = s.ps[p];
			if (ti < 1 || sj < 1) continue;
			if (wd.compat_flip && !(ti < sj)) continue;
#pragma unroll
			for (int c = 0; c < 6; c++) {
				const float v = -s.pairW[p * kTileVals + c_sym_idx[r * 6 + c]];
				float* p1 = &s.A[((sj - 1) * 6 + r) * ld + (ti - 1) * 6 + c];
				float* p2 = &s.A[((ti - 1) * 6 + c) * ld + (sj - 1) * 6 + r];
				if (uniq) { *p1 += v; *p2 += v; } else { atomicAdd(p1, v); atomicAdd(p2, v); }
				if (dbg) {
					float* Dg = a.dbg_JtJ + (size_t)w * a.dbg_stride * a.dbg_stride;
					const int dim = 6 * N;
					atomicAdd(&Dg[(sj * 6 + r) * dim + ti * 6 + c], v);
					atomicAdd(&Dg[(ti * 6 + c) * dim + sj * 6 + r], v);
				}
			}
		}
	}
	__syncthreads();
	PROF_T(6);

	// ---- PCG (PCGInit_Kernel1/2, PCGStep_Kernel*: SolverBundling.cu:575-818).  One THREAD per unknown: thread r reads row r of A
	//      with 16-byte loads (conflict-free, see tail_ld) against p in shared memory (broadcast) - no cross-lane reduction inside the
	//      matrix-vector product; only the two scalar products of a step are reduced (warp shuffle + one word per warp through shared
	//      memory, named barrier among the ceil(dimp / 32) warps that take part).  The first version spread every ROW over the lanes of
	//      a warp: 20 dependent shuffles per four rows, 2.2 k cycles per step for 54 unknowns; this one takes ~0.8 k.
	{
		__shared__ float s_dot[3][8];
		const int nwp = (dimp + 31) >> 5;      // warps that take part (<= 6)
		if (dimp <= 64) {
			// up to 11 frames (the shipped max_BA_frames is 10): ONE warp, two rows per lane - the scalar products are plain warp shuffles,
			// nothing crosses a barrier (8.1 k -> ~4 k cycles for the five steps of a 10-frame window)
			if (wid == 0) {
				const int r0 = lane, r1 = lane + 32;
				const bool act0 = r0 < dimp, act1 = r1 < dimp;
				const int n4 = (dimp + 3) >> 2;
				float rr0 = act0 ? s.rhs[r0] : 0.f, rr1 = act1 ? s.rhs[r1] : 0.f;
				const float mi0 = act0 ? s.Minv[r0] : 0.f, mi1 = act1 ? s.Minv[r1] : 0.f;
				float pp0 = mi0 * rr0, pp1 = mi1 * rr1, dl0 = 0.f, dl1 = 0.f;
				if (r0 < 4 * n4) s.p[r0] = pp0;      // (inactive rows hold zeros: the padding of p up to a multiple of four is zero, like A's columns there)
				if (r1 < 4 * n4) s.p[r1] = pp1;
				__syncwarp();
				float rz = warp_sum(rr0 * pp0 + rr1 * pp1);
				const float4* A0 = reinterpret_cast<const float4*>(s.A + (act0 ? r0 : 0) * ld);
				const float4* A1 = reinterpret_cast<const float4*>(s.A + (act1 ? r1 : 0) * ld);
				const float4* p4 = reinterpret_cast<const float4*>(s.p);
#pragma unroll 1
				for (int lin = 0; lin < a.prm.num_iter_inner; lin++) {
					float a00 = 0.f, a01 = 0.f, a02 = 0.f, a03 = 0.f, a10 = 0.f, a11 = 0.f, a12 = 0.f, a13 = 0.f;
#pragma unroll 4
					for (int c = 0; c < n4; c++) {
						const float4 pv = p4[c], u = A0[c], v = A1[c];
						a00 += u.x * pv.x; a01 += u.y * pv.y; a02 += u.z * pv.z; a03 += u.w * pv.w;
						a10 += v.x * pv.x; a11 += v.y * pv.y; a12 += v.z * pv.z; a13 += v.w * pv.w;
					}
					const float ap0 = act0 ? (a00 + a01) + (a02 + a03) : 0.f, ap1 = act1 ? (a10 + a11) + (a12 + a13) : 0.f;
					const float pAp = warp_sum(pp0 * ap0 + pp1 * ap1);
					const float alpha = (pAp > kEps) ? rz / pAp : 0.f;
					dl0 += alpha * pp0; dl1 += alpha * pp1;
					rr0 -= alpha * ap0; rr1 -= alpha * ap1;
					const float zz0 = mi0 * rr0, zz1 = mi1 * rr1;
					const float rz_new = warp_sum(zz0 * rr0 + zz1 * rr1);
					const float beta = (rz > kEps) ? rz_new / rz : 0.f;
					rz = rz_new;
					pp0 = zz0 + beta * pp0; pp1 = zz1 + beta * pp1;
					__syncwarp();      // every lane has read the old p
					if (act0) s.p[r0] = pp0;
					if (act1) s.p[r1] = pp1;
					__syncwarp();
				}
				if (act0) s.delta[r0] = dl0;
				if (act1) s.delta[r1] = dl1;
			}
		} else if (wid < nwp) {
			const int r = tid;
			const bool act = r < dimp;
			const int rc = act ? r : dimp - 1;       // idle lanes of the last warp read a valid row
			const int nthr = nwp * 32;
			const int n4 = (dimp + 3) >> 2;
			float rr = act ? s.rhs[r] : 0.f;
			const float mi = act ? s.Minv[r] : 0.f;
			float pp = mi * rr, dl = 0.f;
			if (r < 4 * n4) s.p[r] = act ? pp : 0.f;      // (the padding of p up to a multiple of four is zero, and so are A's columns there)
			{ const float v = warp_sum(rr * pp); if (lane == 0) s_dot[0][wid] = v; }
			asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
			float rz = 0.f;
			for (int q = 0; q < nwp; q++) rz += s_dot[0][q];
			const float4* Arow = reinterpret_cast<const float4*>(s.A + rc * ld);
			const float4* p4 = reinterpret_cast<const float4*>(s.p);
#pragma unroll 1
			for (int lin = 0; lin < a.prm.num_iter_inner; lin++) {
				float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
				for (int c = 0; c < n4; c++) { const float4 av = Arow[c], pv = p4[c]; a0 += av.x * pv.x; a1 += av.y * pv.y; a2 += av.z * pv.z; a3 += av.w * pv.w; }
				const float ap = act ? (a0 + a1) + (a2 + a3) : 0.f;
				{ const float v = warp_sum(pp * ap); if (lane == 0) s_dot[1][wid] = v; }
				asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
				float pAp = 0.f;
				for (int q = 0; q < nwp; q++) pAp += s_dot[1][q];
				const float alpha = (pAp > kEps) ? rz / pAp : 0.f;
				dl += alpha * pp;
				rr -= alpha * ap;
				const float zz = mi * rr;
				{ const float v = warp_sum(zz * rr); if (lane == 0) s_dot[2][wid] = v; }
				asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
				float rz_new = 0.f;
				for (int q = 0; q < nwp; q++) rz_new += s_dot[2][q];
				const float beta = (rz > kEps) ? rz_new / rz : 0.f;
				rz = rz_new;
				pp = zz + beta * pp;
				if (act) s.p[r] = pp;
				asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");      // p of the next step is complete (and s_dot[1] / [2] may be rewritten)
			}
			if (act) s.delta[r] = dl;
		}
	}
	__syncthreads();
	PROF_T(7);
	// ---- pose update: x <- log(exp(delta) * exp(x))  (computeLieUpdate), new T for the next iteration
	const bool last = (it == a.prm.num_iter_outer - 1);
	for (int f = tid; f < N; f += NT) {
		float* xg = a.x + (size_t)(wd.frame_off + f) * 6;
		V3 rot = mk(__ldcg(xg + 0), __ldcg(xg + 1), __ldcg(xg + 2)), trans = mk(__ldcg(xg + 3), __ldcg(xg + 4), __ldcg(xg + 5));
		float Tn[12];
		if (f >= 1) {
			const float* dl = s.delta + (f - 1) * 6;
			float U[12];
			se3_exp(mk(dl[3], dl[4], dl[5]), mk(dl[0], dl[1], dl[2]), U);
			const float* Cm = s.T + f * 12;    // exp(x) of this iteration, already in shared memory
			for (int r = 0; r < 3; r++) {
				for (int c = 0; c < 4; c++) {
					float v = U[r * 4 + 0] * Cm[0 * 4 + c] + U[r * 4 + 1] * Cm[1 * 4 + c] + U[r * 4 + 2] * Cm[2 * 4 + c];
					if (c == 3) v += U[r * 4 + 3];
					Tn[r * 4 + c] = v;
				}
			}
			se3_log(Tn, rot, trans);
			__stcg(xg + 0, rot.x); __stcg(xg + 1, rot.y); __stcg(xg + 2, rot.z); __stcg(xg + 3, trans.x); __stcg(xg + 4, trans.y); __stcg(xg + 5, trans.z);
		}
		se3_exp(rot, trans, Tn);
		float* Tg = a.T + (size_t)(wd.frame_off + f) * 12;
		for (int k = 0; k < 12; k++) __stcg(Tg + k, Tn[k]);
		if (last) {   // convertPosesToMatricesCU (SBA.cu:97-104)
			float* Po = a.pose_out + (size_t)(wd.frame_off + f) * 16;
			for (int k = 0; k < 12; k++) Po[k] = Tn[k];
			Po[12] = 0.f; Po[13] = 0.f; Po[14] = 0.f; Po[15] = 1.f;
		}
	}
	__syncthreads();
	PROF_T(8);
	if (a.prof && threadIdx.x == 0) { prf.t[9] = 0; prof_emit(a, prf); }
}

// ------------------------------------------------------------------------------------------------ k_solve
template <int NT, int MINB> __global__ void __launch_bounds__(NT, MINB) k_solve(SolveArgs a) {
	extern __shared__ __align__(16) float dyn_smem[];
	__shared__ int s_tile, s_it, s_idx;
	__shared__ Tile s_tl[2];      // the tile being processed and the one claimed for the next round (thread 0 fetches its record a tile ahead)
	__shared__ float s_M[12];
	__shared__ float s_X[36];
	__shared__ float s_red[kTileVals];
	__shared__ float s_part[(NT / 32)][kTileVals];
	__shared__ int s_is_last, s_p2_claim;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int total = *a.n_tiles_total;
	if (total <= 0) return;
	const int all = total * a.prm.num_iter_outer;      // (the tile plan reports an overflow when this does not fit 31 bits)
	// One iteration's tiles fit the grid (a single window, a handful of windows): STATIC assignment - CTA c runs tile c of every GN
	// iteration.  With the queue the CTAs claim one tile ahead, i.e. into the NEXT iteration, and then sit on the iteration flag while
	// the current iteration's last tiles wait for a CTA that finishes its first one: measured on the device-wide timer for one 10-frame
	// window, 213 of 296 CTAs had a tile of an iteration, 35 of them two in a row, and an iteration's tile phase took 12-13 us instead of 7.
	const bool fixed = total <= (int)gridDim.x;
	if (fixed && (int)blockIdx.x >= total) return;
	if (tid == 0) {
		const int t0 = fixed ? (int)blockIdx.x : atomicAdd(a.queue, 1);
		s_tile = t0; s_p2_claim = 0;
		if (t0 < all) { const int it0 = t0 / total; s_it = it0; s_idx = t0 - it0 * total; s_tl[0] = a.tiles[t0 - it0 * total]; }
	}
	int cur = 0;
	for (;;) {
		__syncthreads();
		if (s_tile >= all) break;
		const int it = s_it, tl_idx = s_idx;
		const Tile& tl = s_tl[cur];      // (read from shared memory where it is used: 12 registers less across the pixel loop)
		// claim the NEXT tile now: the atomic's round trip hides behind this tile.  Safe for the iteration dependencies:
		// a CTA only ever waits on tiles with a smaller index than the one it is processing, never on its look-ahead.
		int nxt = 0;
		if (tid == 0) nxt = fixed ? s_tile + total : atomicAdd(a.queue, 1);
		ProfRec prf; PROF_T(0);
		const WinDesc& wd = a.wins[tl.win];      // (fields are read where they are used: geometry in the pixel loop, the rest in the tail)
		if (it > 0) {   // this window's previous GN iteration must have published its poses
			if (tid == 0) { while (ld_relaxed(a.iter_done + tl.win) < it) __nanosleep(64); (void)ld_acquire(a.iter_done + tl.win); }
			__syncthreads();
		}
		PROF_T(1);
		if (tl.pair == -2) {     // the window's sparse tile: moment sums of all its (i,j) groups for the current poses (CTA-uniform branch)
			PROF_T(2);
			sparse_sums<NT>(a, tl.win, tl.start, tl.count);
			PROF_T(3);
			int itn = 0, idxn = 0;      // thread 0: the record of the tile claimed above (in flight across the barrier and the ticket)
			if (tid == 0 && nxt < all) { itn = nxt / total; idxn = nxt - itn * total; tile_fetch_async(&s_tl[cur ^ 1], a.tiles + idxn); }
			__syncthreads();      // (every thread's sums are written; the barrier orders them before thread 0's release below)
			if (tid == 0) {
				const int done = ticket_acq_rel(a.tiles_done + tl.win) + 1;
				s_is_last = (done == (it + 1) * tl.n_tiles_win);
				s_tile = nxt;
				if (nxt < all) { s_it = itn; s_idx = idxn; }
				tile_fetch_wait();
			}
			PROF_T(4);
			__syncthreads();
		} else {
		TileAcc acc;
#pragma unroll
		for (int k = 0; k < kTileVals; k++) acc.v[k] = 0.f;
		if (tl.pair >= 0) {
			if (tid < 12) {   // transform = T_i^-1 T_j (rigid inverse), row r, col c
				const float* Ti = a.T + (size_t)tl.tgt_slot * 12;
				const float* Tj = a.T + (size_t)tl.src_slot * 12;
				const int r = tid >> 2, c = tid & 3;
				float v = 0.f;
				for (int k = 0; k < 3; k++) {
					const float rik = __ldcg(Ti + k * 4 + r);          // R_i^T (r,k) = R_i(k,r)
					v += rik * ((c < 3) ? __ldcg(Tj + k * 4 + c) : (__ldcg(Tj + k * 4 + 3) - __ldcg(Ti + k * 4 + 3)));
				}
				s_M[tid] = v;
			} else if (tid >= 32 && tid < 68) {   // X_i = [[R,0],[[t]x R, R]] of the TARGET frame: maps the tile's sums to the model frame
				const float* Ti = a.T + (size_t)tl.tgt_slot * 12;
				const int e = tid - 32, r = e / 6, c = e - r * 6;
				float v;
				if (r < 3) v = (c < 3) ? __ldcg(Ti + r * 4 + c) : 0.f;
				else if (c >= 3) v = __ldcg(Ti + (r - 3) * 4 + (c - 3));
				else { const float t3[3] = { __ldcg(Ti + 3), __ldcg(Ti + 7), __ldcg(Ti + 11) }; v = 0.f; for (int kk = 0; kk < 3; kk++) v += skew(t3, r - 3, kk) * __ldcg(Ti + kk * 4 + c); }
				s_X[e] = v;
			}
			__syncthreads();
			PROF_T(2);
			tile_pixels<NT>(a, wd, tl.src, tl.tex, s_M, tl.start, tl.count, acc);
		}
		PROF_T(3);
		int itn = 0, idxn = 0;
		if (tid == 0 && nxt < all) { itn = nxt / total; idxn = nxt - itn * total; tile_fetch_async(&s_tl[cur ^ 1], a.tiles + idxn); }   // the next tile's record, in flight behind the reduction
		// block reduction of the 28 sums -> this tile's slot.  Warp level: a transposing butterfly (31 shuffles instead of
		// 28 x 5): after the five steps lane L holds the warp total of value L.
		{
			float v[32];
#pragma unroll
			for (int k = 0; k < 32; k++) v[k] = (k < kTileVals) ? acc.v[k] : 0.f;
#pragma unroll
			for (int half = 16; half >= 1; half >>= 1) {
				const bool up = (lane & half) != 0;
#pragma unroll
				for (int k = 0; k < half; k++) {
					const float send = up ? v[k] : v[k + half];
					const float keep = up ? v[k + half] : v[k];
					v[k] = keep + __shfl_xor_sync(0xffffffffu, send, half);
				}
			}
			if (lane < kTileVals) s_part[wid][lane] = v[0];
		}
		__syncthreads();
		if (wid == 0) {   // warp 0 finishes the tile: cross-warp sum, model-frame transform, store, ticket
			float red = 0.f;
			if (lane < kTileVals) {
#pragma unroll
				for (int k = 0; k < (NT / 32); k++) red += s_part[k][lane];
				s_red[lane] = red;
			}
			__syncwarp();
			if (lane < kTileVals) {   // S = X S' X^T, b = X b' (rows 0-2 of X have 3 non-zeros); entry 27 = #correspondences
				float out = red;
				if (tl.pair >= 0 && lane < 27) {
					out = 0.f;
					if (lane < 21) {
						const int r = c_sym_r[lane], c = c_sym_c[lane];
						float xr[6], xc[6];
#pragma unroll
						for (int u = 0; u < 6; u++) { xr[u] = s_X[r * 6 + u]; xc[u] = s_X[c * 6 + u]; }   // zero entries make the short rows exact
#pragma unroll
						for (int u = 0; u < 6; u++) {
							float tsum = 0.f;
#pragma unroll
							for (int v = 0; v < 6; v++) tsum += s_red[c_sym_idx[u * 6 + v]] * xc[v];
							out += xr[u] * tsum;
						}
					} else {
						const int r = lane - 21;
#pragma unroll
						for (int u = 0; u < 6; u++) out += s_X[r * 6 + u] * s_red[21 + u];
					}
				}
				__stcg(a.partial + (size_t)tl_idx * kTileVals + lane, out);
			}
			__syncwarp();             // orders the 28 lanes' stores before lane 0's release
			if (lane == 0) {
				const int done = ticket_acq_rel(a.tiles_done + tl.win) + 1;
				s_is_last = (done == (it + 1) * tl.n_tiles_win);
				s_tile = nxt;
				if (nxt < all) { s_it = itn; s_idx = idxn; }
				tile_fetch_wait();
			}
		}
		PROF_T(4);
		__syncthreads();
		}      // dense tile
		PROF_T(5);
		if (a.prof && tid == 0) { prf.kind_cta = blockIdx.x; prf.tile_win = ((long long)tl_idx << 32) | ((long long)it << 16) | tl.count; prf.t[6] = prf.t[7] = prf.t[8] = prf.t[9] = 0; prof_emit(a, prf); }
		if (s_is_last) {
			window_tail<NT>(a, wd, tl.win, it, dyn_smem, &s_p2_claim);
			__threadfence();
			__syncthreads();
			if (tid == 0) st_release(a.iter_done + tl.win, it + 1);
		}
		cur ^= 1;
	}
}

// ------------------------------------------------------------------------------------------------ host side
// Everything bt_solve_stage uploads lives in ONE pinned block mirrored by ONE device block (same offsets), so staging is a
// single H2D copy: the small tables first (carved for the worst case of this window count), the correspondences last so
// that the copy length follows the actual number of entries.
struct StageLayout {
	size_t wins, dp, np, fw, texp, srcp, nsrcp, pose, pairs, pwin, psrc, late, wsp, gi, gj, gs, mem, corr, total;   // [0,late): needed by k_prep_frames
};
static StageLayout stage_layout(int n_windows, size_t F, size_t C, int max_frames, size_t maxG, size_t maxP) {
	StageLayout L;
	size_t off = 0;
	auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
	const size_t nw = (size_t)n_windows;
	L.wins = carve(sizeof(WinDesc) * nw); L.dp = carve(sizeof(void*) * F); L.np = carve(sizeof(void*) * F); L.fw = carve(sizeof(int) * F);
	L.texp = carve(sizeof(void*) * F); L.srcp = carve(sizeof(void*) * F); L.nsrcp = carve(sizeof(void*) * F);
	L.pose = carve(sizeof(float) * 16 * F); L.pairs = carve(sizeof(uint2) * maxP * nw); L.pwin = carve(sizeof(int) * maxP * nw); L.psrc = carve(sizeof(int) * maxP * nw);
	L.late = off;
	L.wsp = carve(sizeof(WinSparse) * nw); L.gi = carve(sizeof(int) * maxG * nw); L.gj = carve(sizeof(int) * maxG * nw); L.gs = carve(sizeof(int) * (maxG + 1) * nw);
	L.mem = carve(sizeof(int) * (2 * ((size_t)max_frames + 1) + 2 * maxG + 2 * maxP) * nw);
	L.corr = carve(sizeof(bt_entryj) * C + 32);
	L.total = off;
	return L;
}

struct SolverState {
	bt_solver_limits lim{};
	int npix_max = 0, max_pairs = 0, max_frames_total = 0, max_tiles = 0, max_groups = 0;
	DevBuf blk_cnt, grp_sums, stage_buf[2], texel, src, nsrc, x, T, pose_out, tiles, scalars, partial, tiles_done, iter_done, pair_tile0, pair_ntile, dbgJ, dbgR, dbgC, prof;
	StageLayout layout{};
	// frame cache (bt_frame_cache_*): quarter-res maps of keyframes, built once, referenced by bt_window::cache_slots
	struct CacheMeta { bool valid = false; int H = 0, W = 0; float fx = 0, fy = 0, cx = 0, cy = 0, dmin = 0, dmax = 0; };
	DevBuf c_texel, c_src, c_nsrc, c_tables, c_blk;
	PinnedBuf c_htables[2];              // two pinned table blocks, used alternately: a store only waits for the upload of the store before last,
	cudaEvent_t c_ev[2] = { nullptr, nullptr };   // so back-to-back stores (one new frame per window per step) never stall the host behind the GPU
	int c_flip = 0;
	int c_capacity = 0, c_npix = 0;
	float c_downscale = 0.f;
	std::vector<CacheMeta> c_meta;
	int prof_cap = 0;
	PinnedBuf h_stage, h_poses;
	// streaming form (bt_solve_windows_begin / _end): two batches may be in flight; each has its own pinned pose block and completion event
	PinnedBuf h_pipe[2];
	cudaEvent_t ev_pipe[2] = { nullptr, nullptr };
	bool pipe_pending[2] = { false, false };
	int pipe_frames[2] = { 0, 0 };
	int pipe_head = 0, pipe_tail = 0;
	// last staged batch
	int n_windows = 0, frames_total = 0, smem_bytes = 0, chunk = 1024;
	bt_solver_params prm{};
	std::vector<int> frame_off;
	std::vector<int> n_frames;
	bool staged = false, debug = false, timing = false;
	double host_us[6] = { 0, 0, 0, 0, 0, 0 };   // host time of the last call: tables+early upload, prep launch, correspondence scan+staging, run (launch), fetch (copy + wait), total
	bool grp_in_smem = true;
	int force_chunk = 0;                 // tuning knob (BT_SOLVE_CHUNK): source pixels per dense tile
	int launches = 0;
	int attr_bytes = 0, occ = BT_SOLVE_MIN_CTAS, occ_smem = -1;
	bool prep_launched = false, any_uncached = true;
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t ev_prev = nullptr, ev_corr = nullptr, ev_h2d = nullptr, ev_early = nullptr;
	cudaEvent_t ev_stage_free[2] = { nullptr, nullptr };   // behind the k_solve that last read staging block 0 / 1
	int stage_cur = 0;
	cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
};

void solver_destroy(bt_ctx* ctx) {
	SolverState* s = ctx->solver;
	if (!s) return;
	DevBuf* bufs[] = { &s->blk_cnt, &s->grp_sums, &s->c_texel, &s->c_src, &s->c_nsrc, &s->c_tables, &s->c_blk, &s->stage_buf[0], &s->stage_buf[1], &s->texel, &s->src, &s->nsrc, &s->x, &s->T, &s->pose_out, &s->tiles, &s->scalars, &s->partial,
	                   &s->tiles_done, &s->iter_done, &s->pair_tile0, &s->pair_ntile, &s->dbgJ, &s->dbgR, &s->dbgC, &s->prof };
	for (DevBuf* b : bufs) b->release();
	s->h_stage.release(); s->h_poses.release(); s->h_pipe[0].release(); s->h_pipe[1].release();
	for (auto& e : s->ev_pipe) if (e) cudaEventDestroy(e); s->c_htables[0].release(); s->c_htables[1].release();
	for (auto& e : s->c_ev) if (e) cudaEventDestroy(e);
	for (auto& e : s->ev) if (e) cudaEventDestroy(e);
	for (cudaEvent_t e : { s->ev_prev, s->ev_corr, s->ev_h2d, s->ev_early, s->ev_stage_free[0], s->ev_stage_free[1] }) if (e) cudaEventDestroy(e);
	if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
	delete s;
	ctx->solver = nullptr;
}

static int reserve_impl(bt_ctx* ctx, const bt_solver_limits* lim) {
	if (!ctx->solver) ctx->solver = new SolverState();
	SolverState* s = ctx->solver;
	BT_REQUIRE(lim->max_windows > 0 && lim->max_frames >= 2 && lim->max_frames <= kMaxFrames && lim->max_corr >= 0 && lim->H > 0 && lim->W > 0 && lim->image_downscale >= 1.0f,
	           BT_ERR_INVALID_ARG, "bt_solver_reserve: bad limits (max_frames must be 2..%d)", kMaxFrames);
	s->lim = *lim;
	{ const char* e = getenv("BT_SOLVE_CHUNK"); const int v = e ? atoi(e) : 0; s->force_chunk = (v >= 256 && v <= 16384) ? (v + 255) / 256 * 256 : 0; }
	const int w = (int)(lim->W / lim->image_downscale), h = (int)(lim->H / lim->image_downscale);
	s->npix_max = w * h;
	const int F = lim->max_windows * lim->max_frames;
	s->max_frames_total = F;
	s->max_pairs = lim->max_frames * (lim->max_frames - 1);   // both directions allowed in custom lists
	s->max_groups = lim->max_frames * lim->max_frames;
	const int min_chunk = 256;   // worst case: every pixel valid at the smallest chunk the scheduler ever picks
	s->max_tiles = lim->max_windows * (lim->max_frames * (lim->max_frames - 1) / 2) * ((s->npix_max + min_chunk - 1) / min_chunk + 1) + kSparseTiles * lim->max_windows;
	int rc;
#define RES(buf, bytes) if ((rc = s->buf.alloc(bytes)) != BT_OK) return rc
	{   // worst case over every window count <= max_windows (the table offsets grow with the count, so the maximum is at max_windows)
		const StageLayout L = stage_layout(lim->max_windows, (size_t)F, (size_t)lim->max_corr * lim->max_windows, lim->max_frames, (size_t)s->max_groups, (size_t)s->max_pairs);
		RES(stage_buf[0], L.total); RES(stage_buf[1], L.total);
	}
	RES(texel, sizeof(float4) * 2 * (size_t)s->npix_max * F);
	RES(src, sizeof(float4) * 2 * (size_t)s->npix_max * F);
	RES(nsrc, sizeof(int) * F);
	RES(grp_sums, sizeof(float) * kGrpVals * (size_t)s->max_groups * lim->max_windows);
	RES(blk_cnt, sizeof(int) * (size_t)F * ((s->npix_max + 1023) / 1024));
	RES(x, sizeof(float) * 6 * F); RES(T, sizeof(float) * 12 * F); RES(pose_out, sizeof(float) * 16 * F);
	RES(pair_tile0, sizeof(int) * (size_t)s->max_pairs * lim->max_windows);
	RES(pair_ntile, sizeof(int) * (size_t)s->max_pairs * lim->max_windows);
	RES(tiles, sizeof(Tile) * (size_t)s->max_tiles);
	RES(partial, sizeof(float) * kTileVals * (size_t)s->max_tiles);
	RES(scalars, 64);
	BT_CUDA(cudaMemset(s->scalars.p, 0, 64));
	RES(tiles_done, sizeof(int) * lim->max_windows); RES(iter_done, sizeof(int) * lim->max_windows);
#undef RES
	return BT_OK;
}

}  // namespace bt

using namespace bt;

extern "C" int bt_solver_reserve(bt_ctx* ctx, const bt_solver_limits* lim) {
	BT_REQUIRE(ctx && lim, BT_ERR_INVALID_ARG, "bt_solver_reserve: NULL argument");
	BT_CUDA(cudaSetDevice(ctx->device));
	return reserve_impl(ctx, lim);
}

extern "C" int bt_solve_enable_debug(bt_ctx* ctx, int on) {
	BT_REQUIRE(ctx && ctx->solver, BT_ERR_INVALID_ARG, "bt_solve_enable_debug: call bt_solver_reserve first");
	ctx->solver->debug = on != 0;
	return BT_OK;
}

extern "C" int bt_solve_enable_timing(bt_ctx* ctx, int on) {
	BT_REQUIRE(ctx && ctx->solver, BT_ERR_INVALID_ARG, "bt_solve_enable_timing: call bt_solver_reserve first");
	SolverState* s = ctx->solver;
	if (on) for (auto& e : s->ev) if (!e) BT_CUDA(cudaEventCreate(&e));
	s->timing = on != 0;
	return BT_OK;
}

// Device time of the LAST bt_solve_run as (prep incl. the tile plan, gap, solve), from CUDA events recorded on the
// launching stream.  Synchronises on the last event.
extern "C" int bt_solve_get_timing(bt_ctx* ctx, float* ms3) {
	BT_REQUIRE(ctx && ctx->solver && ctx->solver->timing && ms3, BT_ERR_INVALID_ARG, "bt_solve_get_timing: timing not enabled");
	SolverState* s = ctx->solver;
	BT_CUDA(cudaEventSynchronize(s->ev[3]));
	for (int k = 0; k < 3; k++) BT_CUDA(cudaEventElapsedTime(ms3 + k, s->ev[k], s->ev[k + 1]));
	return BT_OK;
}

static SolveArgs make_args(bt_ctx* ctx);
static int launch_prep(bt_ctx* ctx, cudaStream_t stream);

// Staging order (what makes the fused call fast): everything the frame preparation needs (window descriptors, frame
// pointers, poses, pair tables - ~100 KB) is built and uploaded first and, in the fused call, k_prep_frames is launched
// right there; only then are the correspondences (the bulk: 32 B each) copied into pinned memory and uploaded in
// chunks on a private copy stream, overlapping the kernel.  k_solve waits on the last chunk's event.
static int stage_impl(bt_ctx* ctx, int n_windows, const bt_window* windows, const bt_solver_params* params,
                      const float* poses_in, void* stream_, bool early_prep) {
	BT_REQUIRE(ctx && windows && params && poses_in, BT_ERR_INVALID_ARG, "bt_solve_stage: NULL argument");
	BT_REQUIRE(ctx->solver, BT_ERR_INVALID_ARG, "bt_solve_stage: call bt_solver_reserve first");
	SolverState* s = ctx->solver;
	cudaStream_t stream = (cudaStream_t)stream_;
	BT_CUDA(cudaSetDevice(ctx->device));
	BT_REQUIRE(n_windows > 0 && n_windows <= s->lim.max_windows, BT_ERR_CAPACITY, "bt_solve_stage: %d windows > reserved %d", n_windows, s->lim.max_windows);
	BT_REQUIRE(params->num_iter_outer > 0 && params->num_iter_inner > 0 && params->image_downscale == s->lim.image_downscale, BT_ERR_INVALID_ARG,
	           "bt_solve_stage: iteration counts must be positive and image_downscale must equal the reserved value");
	s->staged = false; s->prep_launched = false;
	const auto t_h0 = std::chrono::steady_clock::now();
	auto us_since = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); };
	if (s->ev_h2d) BT_CUDA(cudaEventSynchronize(s->ev_h2d));     // the previous upload still reads the pinned block
	// ---- sizes
	size_t F = 0, C = 0;
	for (int w = 0; w < n_windows; w++) {
		const bt_window& bw = windows[w];
		BT_REQUIRE(bw.n_frames >= 2 && bw.n_frames <= s->lim.max_frames, BT_ERR_CAPACITY, "window %d: n_frames %d outside [2,%d]", w, bw.n_frames, s->lim.max_frames);
		long long nc = bw.n_corr;
		if (bw.corr_dev) {     // device-resident blocks: validate the host-side description
			BT_REQUIRE(bw.n_blocks >= 0 && (bw.n_blocks == 0 || (bw.block_off && bw.block_n && bw.block_i && bw.block_j)), BT_ERR_INVALID_ARG, "window %d: corr_dev given without block arrays", w);
			nc = 0;
			for (int b = 0; b < bw.n_blocks; b++) {
				BT_REQUIRE(bw.block_n[b] >= 0 && bw.block_off[b] >= 0 && (b == 0 || bw.block_off[b] == bw.block_off[b - 1] + bw.block_n[b - 1]), BT_ERR_INVALID_ARG,
				           "window %d: correspondence block %d is not back to back with its predecessor", w, b);
				BT_REQUIRE(bw.block_n[b] == 0 || (bw.block_i[b] < (uint32_t)bw.n_frames && bw.block_j[b] < (uint32_t)bw.n_frames), BT_ERR_INVALID_ARG,
				           "window %d: correspondence block %d references a frame outside [0,%d)", w, b, bw.n_frames);
				nc += bw.block_n[b];
			}
		}
		BT_REQUIRE(nc >= 0 && nc <= s->lim.max_corr, BT_ERR_CAPACITY, "window %d: n_corr %lld > reserved %d", w, nc, s->lim.max_corr);
		BT_REQUIRE(bw.H > 0 && bw.W > 0 && bw.H <= s->lim.H && bw.W <= s->lim.W, BT_ERR_CAPACITY, "window %d: image %dx%d larger than reserved", w, bw.W, bw.H);
		BT_REQUIRE((int)(bw.W / params->image_downscale) >= 2 && (int)(bw.H / params->image_downscale) >= 2, BT_ERR_INVALID_ARG, "window %d: image too small", w);
		BT_REQUIRE(bw.corr_dev || bw.n_corr == 0 || bw.corr, BT_ERR_INVALID_ARG, "window %d: corr is NULL", w);
		BT_REQUIRE(params->w_dense <= 0.f || bw.cache_slots || (bw.depth_dev && bw.normal_dev), BT_ERR_INVALID_ARG, "window %d: depth/normal pointers are NULL and no cache slots given", w);
		F += bw.n_frames; C += (size_t)nc;
	}
	// ---- host staging: one pinned block mirrored by one device block
	const size_t maxP = (size_t)s->max_pairs, maxG = (size_t)s->max_groups;
	const StageLayout L = stage_layout(n_windows, F, C, s->lim.max_frames, maxG, maxP);
	BT_REQUIRE(L.total <= s->stage_buf[0].bytes, BT_ERR_CAPACITY, "bt_solve_stage: staging block %zu > reserved %zu bytes", L.total, s->stage_buf[0].bytes);
	int rc = s->h_stage.alloc(L.total);
	if (rc != BT_OK) return rc;
	char* hb = s->h_stage.as<char>();
	WinDesc* hw = (WinDesc*)(hb + L.wins);
	const void** hdp = (const void**)(hb + L.dp); const void** hnp = (const void**)(hb + L.np);
	const void** htex = (const void**)(hb + L.texp); const void** hsrc = (const void**)(hb + L.srcp); const void** hnsrc = (const void**)(hb + L.nsrcp);
	int* hfw = (int*)(hb + L.fw); float* hpose = (float*)(hb + L.pose); bt_entryj* hcorr = (bt_entryj*)(hb + L.corr);
	WinSparse* hws = (WinSparse*)(hb + L.wsp);
	int* hgi = (int*)(hb + L.gi); int* hgj = (int*)(hb + L.gj); int* hgs = (int*)(hb + L.gs); uint2* hpairs = (uint2*)(hb + L.pairs);
	int* hpwin = (int*)(hb + L.pwin); int* hpsrc = (int*)(hb + L.psrc); int* hmem = (int*)(hb + L.mem);
	s->frame_off.assign(n_windows, 0); s->n_frames.assign(n_windows, 0);

	s->any_uncached = false;
	// ==== phase 1: what the frame preparation needs - geometry, frame tables, poses, dense pair tables (no look at the correspondences)
	size_t f_off = 0, p_off = 0;
	for (int w = 0; w < n_windows; w++) {
		const bt_window& bw = windows[w];
		const int N = bw.n_frames;
		WinDesc& d = hw[w];
		memset(&d, 0, sizeof d);
		d.n_frames = N; d.H = bw.H; d.W = bw.W;
		d.w = (int)(bw.W / params->image_downscale); d.h = (int)(bw.H / params->image_downscale);   // LossGPU.cu:56-57
		d.fx = bw.fx * ((float)d.w / (float)bw.W); d.fy = bw.fy * ((float)d.h / (float)bw.H);         // CUDACache.cpp:20-24
		d.cx = bw.cx * ((float)(d.w - 1) / (float)(bw.W - 1)); d.cy = bw.cy * ((float)(d.h - 1) / (float)(bw.H - 1));
		d.ifx = 1.0f / bw.fx; d.ify = 1.0f / bw.fy; d.icx = -bw.cx / bw.fx; d.icy = -bw.cy / bw.fy;
		d.scaleW = (float)(bw.W - 1) / (float)(d.w - 1); d.scaleH = (float)(bw.H - 1) / (float)(d.h - 1);
		d.compat_flip = bw.compat_flip;
		d.frame_off = (int)f_off; d.pair_off = (int)p_off;
		s->frame_off[w] = (int)f_off; s->n_frames[w] = N;
		for (int f = 0; f < N; f++) {
			const size_t fs = f_off + f;
			if (bw.cache_slots && params->w_dense > 0.f) {     // maps built earlier by bt_frame_cache_store
				const int cs = bw.cache_slots[f];
				BT_REQUIRE(cs >= 0 && cs < s->c_capacity && s->c_meta[cs].valid, BT_ERR_INVALID_ARG, "window %d frame %d: cache slot %d is empty or out of range", w, f, cs);
				const SolverState::CacheMeta& m = s->c_meta[cs];
				BT_REQUIRE(m.H == bw.H && m.W == bw.W && m.fx == bw.fx && m.fy == bw.fy && m.cx == bw.cx && m.cy == bw.cy && m.dmin == params->depth_min && m.dmax == params->depth_max
				               && s->c_downscale == params->image_downscale,
				           BT_ERR_INVALID_ARG, "window %d frame %d: cache slot %d was stored with another image size, K, depth range or downscale", w, f, cs);
				hdp[fs] = nullptr; hnp[fs] = nullptr;
				htex[fs] = s->c_texel.as<float4>() + (size_t)cs * 2 * s->c_npix; hsrc[fs] = s->c_src.as<float4>() + (size_t)cs * 2 * s->c_npix;
				hnsrc[fs] = s->c_nsrc.as<int>() + cs;
			} else {
				hdp[fs] = (params->w_dense > 0.f) ? bw.depth_dev[f] : nullptr;
				hnp[fs] = (params->w_dense > 0.f) ? bw.normal_dev[f] : nullptr;
				htex[fs] = s->texel.as<float4>() + fs * 2 * (size_t)s->npix_max; hsrc[fs] = s->src.as<float4>() + fs * 2 * (size_t)s->npix_max;
				hnsrc[fs] = nullptr;
				if (params->w_dense > 0.f) s->any_uncached = true;
			}
			hfw[fs] = w;
			memcpy(hpose + fs * 16, poses_in + fs * 16, sizeof(float) * 16);
		}
		int np = 0;
		if (params->w_dense > 0.f) {
			if (bw.dense_pairs) {
				BT_REQUIRE(bw.n_dense_pairs >= 0 && bw.n_dense_pairs <= s->max_pairs, BT_ERR_CAPACITY, "window %d: %d dense pairs > %d", w, bw.n_dense_pairs, s->max_pairs);
				for (int p = 0; p < bw.n_dense_pairs; p++) {
					const uint32_t ti = bw.dense_pairs[2 * p], sj = bw.dense_pairs[2 * p + 1];
					BT_REQUIRE(ti < (uint32_t)N && sj < (uint32_t)N && ti != sj, BT_ERR_INVALID_ARG, "window %d: dense pair %d = (%u,%u) invalid", w, p, ti, sj);
					hpairs[p_off + np++] = make_uint2(ti, sj);
				}
			} else {
				for (int i = 0; i < N; i++) for (int j = 0; j < i; j++) hpairs[p_off + np++] = make_uint2((unsigned)i, (unsigned)j);
			}
		}
		d.n_pairs = np;
		for (int p = 0; p < np; p++) { hpwin[p_off + p] = w; hpsrc[p_off + p] = (int)f_off + (int)hpairs[p_off + p].y; }
		f_off += N; p_off += np;
	}
	s->n_windows = n_windows; s->frames_total = (int)f_off; s->prm = *params;
	// chunk: aim for >= 4 tiles per SM-resident CTA over the batch, between 256 and 2048 source pixels
	{
		const long long est_px = (long long)p_off * (s->npix_max / 8);   // ~12 % valid
		long long c = est_px / ((long long)ctx->sm_count * 2 * 4);
		c = std::max(256LL, std::min(2048LL, c));
		s->chunk = s->force_chunk ? s->force_chunk : (int)((c + 255) / 256 * 256);
	}
	s->layout = L;
	if (!s->copy_stream) {
		BT_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
		BT_CUDA(cudaEventCreateWithFlags(&s->ev_prev, cudaEventDisableTiming));
		BT_CUDA(cudaEventCreateWithFlags(&s->ev_corr, cudaEventDisableTiming));
		BT_CUDA(cudaEventCreateWithFlags(&s->ev_h2d, cudaEventDisableTiming));
	}
	// Two device staging blocks, used alternately: every upload of this batch runs on the private copy stream and only waits for the
	// k_solve that last read THIS block (two batches ago) - so with the streaming form the 2 MB of batch k+1 cross PCIe while batch k is
	// still being solved.  (With one block the uploads queued behind the previous k_solve: ~50 us of every streaming step.)
	s->stage_cur ^= 1;
	DevBuf& stage_dev = s->stage_buf[s->stage_cur];
	if (!s->ev_stage_free[s->stage_cur]) BT_CUDA(cudaEventCreateWithFlags(&s->ev_stage_free[s->stage_cur], cudaEventDisableTiming));
	else BT_CUDA(cudaStreamWaitEvent(s->copy_stream, s->ev_stage_free[s->stage_cur], 0));
	{   // correspondence blocks produced on the device (bt_match_pairs on the caller's stream) are copied device -> device on the copy stream: order them
		bool any_dev = false;
		for (int w = 0; w < n_windows && !any_dev; w++) any_dev = windows[w].corr_dev != nullptr;
		if (any_dev) { BT_CUDA(cudaEventRecord(s->ev_prev, stream)); BT_CUDA(cudaStreamWaitEvent(s->copy_stream, s->ev_prev, 0)); }
	}
	// early tables first; the caller's stream (frame prep in the fused call) waits for them only
	BT_CUDA(cudaMemcpyAsync(stage_dev.p, hb, L.late, cudaMemcpyHostToDevice, s->copy_stream));
	if (!s->ev_early) BT_CUDA(cudaEventCreateWithFlags(&s->ev_early, cudaEventDisableTiming));
	BT_CUDA(cudaEventRecord(s->ev_early, s->copy_stream));
	BT_CUDA(cudaStreamWaitEvent(stream, s->ev_early, 0));
	s->host_us[0] = us_since(t_h0);
	if (early_prep) { if ((rc = launch_prep(ctx, stream)) != BT_OK) return rc; }
	s->host_us[1] = us_since(t_h0) - s->host_us[0];

	// ==== phase 2 (the GPU is already busy in the fused call): correspondences -> groups, membership CSR, pinned copy, chunked upload
	size_t c_off = 0, g_off = 0, m_off = 0, smem_need = 0, smem_lean = 0, sent = 0;
	const bt_entryj* run_src = nullptr; size_t run_dst = 0, run_n = 0;      // pending in-place transfer from page-locked caller memory
	std::vector<int> bin;
	std::vector<char> seen;
	p_off = 0;
	for (int w = 0; w < n_windows; w++) {
		const bt_window& bw = windows[w];
		const int N = bw.n_frames, np = hw[w].n_pairs;
		const size_t f_off_w = (size_t)hw[w].frame_off;
		(void)f_off_w;
		// correspondences: stable counting sort by (i,j) so each pair's entries are contiguous (Bundler::optimizeGPU
		// already emits them that way, /root/reference/src/Bundler.cpp:298-324); invalid entries are dropped.
		int n_valid = 0, ng = 0;
		bool grouped = true;    // fast path: keys (i*N+j) non-decreasing and no invalid entries => groups are runs, plain copy
		if (bw.corr_dev) {      // blocks produced on the device: every non-empty block is a group, the entries move device -> device
			for (int b = 0; b < bw.n_blocks; b++) {
				if (bw.block_n[b] == 0) continue;
				hgi[g_off + ng] = (int)bw.block_i[b]; hgj[g_off + ng] = (int)bw.block_j[b]; hgs[g_off + w + ng] = n_valid; ng++;
				n_valid += bw.block_n[b];
			}
			hgs[g_off + w + ng] = n_valid;
			if (n_valid) {
				if (c_off > sent) {     // flush what the host path has staged so far: the pinned block has a hole where this window's entries would be
					BT_CUDA(cudaMemcpyAsync(stage_dev.as<char>() + L.corr + sent * sizeof(bt_entryj), hcorr + sent, (c_off - sent) * sizeof(bt_entryj), cudaMemcpyHostToDevice, s->copy_stream));
				}
				BT_CUDA(cudaMemcpyAsync(stage_dev.as<char>() + L.corr + c_off * sizeof(bt_entryj), bw.corr_dev + bw.block_off[0], (size_t)n_valid * sizeof(bt_entryj),
				                        cudaMemcpyDeviceToDevice, s->copy_stream));
				sent = c_off + n_valid;
			}
		} else {
			static_assert(offsetof(bt_entryj, imgIdx_i) == 0 && offsetof(bt_entryj, imgIdx_j) == 4, "EntryJ header layout");
			if (bw.block_n && bw.n_blocks > 0) {
				// the caller names the blocks (Bundler::optimizeGPU emits the entries pair by pair and hands n_match_per_pair along,
				// /root/reference/src/Bundler.cpp:298-351): only the first entry of every block is read - no pass over the 64 KB of a window
				long long off = 0;
				for (int b = 0; b < bw.n_blocks; b++) {
					const int nb = bw.block_n[b];
					BT_REQUIRE(nb >= 0 && off + nb <= bw.n_corr, BT_ERR_INVALID_ARG, "window %d: correspondence block %d overruns n_corr", w, b);
					if (nb == 0) continue;
					const uint32_t ei = bw.block_i ? bw.block_i[b] : bw.corr[off].imgIdx_i, ej = bw.block_j ? bw.block_j[b] : bw.corr[off].imgIdx_j;
					BT_REQUIRE(ei < (uint32_t)N && ej < (uint32_t)N, BT_ERR_INVALID_ARG, "window %d: correspondence block %d references a frame outside [0,%d)", w, b, N);
					hgi[g_off + ng] = (int)ei; hgj[g_off + ng] = (int)ej; hgs[g_off + w + ng] = (int)off; ng++;
					off += nb;
				}
				BT_REQUIRE(off == bw.n_corr, BT_ERR_INVALID_ARG, "window %d: the correspondence blocks cover %lld of %d entries", w, off, bw.n_corr);
			} else {
			// one 8-byte compare per entry (imgIdx_i | imgIdx_j << 32, little endian); range and order are only checked where a run ends
			unsigned long long prev_raw = 0ull;
			bool have_prev = false;
			long long prev_key = -1;
			for (int c = 0; c < bw.n_corr; c++) {
				unsigned long long raw;
				memcpy(&raw, &bw.corr[c], sizeof raw);
				if (have_prev && raw == prev_raw) continue;
				const uint32_t ei = (uint32_t)raw, ej = (uint32_t)(raw >> 32);
				if (ei >= (uint32_t)N || ej >= (uint32_t)N) { grouped = false; break; }
				const long long key = (long long)ei * N + ej;
				if (key < prev_key) { grouped = false; break; }
				hgi[g_off + ng] = (int)ei; hgj[g_off + ng] = (int)ej; hgs[g_off + w + ng] = c; ng++;
				prev_raw = raw; prev_key = key; have_prev = true;
			}
			}
		}
		if (bw.corr_dev) {
			// nothing to copy on the host
		} else if (grouped) {
			n_valid = bw.n_corr;
			hgs[g_off + w + ng] = n_valid;
			if (n_valid) {
				// Caller's buffer page-locked (cudaHostAlloc / cudaHostRegister / bt_host_alloc_pinned)?  Then the copy engine reads it in
				// place - no staging copy - and windows whose buffers follow each other in memory share one transfer.
				cudaPointerAttributes pa;
				const bool pinned_src = cudaPointerGetAttributes(&pa, bw.corr) == cudaSuccess && pa.type == cudaMemoryTypeHost;
				if (pinned_src) {
					if (c_off > sent) {      // flush what was staged so far: the staging block has a hole where this window's entries would be
						BT_CUDA(cudaMemcpyAsync(stage_dev.as<char>() + L.corr + sent * sizeof(bt_entryj), hcorr + sent, (c_off - sent) * sizeof(bt_entryj), cudaMemcpyHostToDevice, s->copy_stream));
					}
					if (run_n && run_src + run_n == bw.corr && run_dst + run_n == c_off) run_n += (size_t)n_valid;      // extends the pending run
					else {
						if (run_n) BT_CUDA(cudaMemcpyAsync(stage_dev.as<char>() + L.corr + run_dst * sizeof(bt_entryj), run_src, run_n * sizeof(bt_entryj), cudaMemcpyHostToDevice, s->copy_stream));
						run_src = bw.corr; run_dst = c_off; run_n = (size_t)n_valid;
					}
					sent = c_off + n_valid;
				} else {
					memcpy(hcorr + c_off, bw.corr, sizeof(bt_entryj) * (size_t)n_valid);      // the window's entries are still in cache from the scan
				}
			}
		} else {
			ng = 0;
			bin.assign((size_t)N * N + 1, 0);
			for (int c = 0; c < bw.n_corr; c++) {
				const bt_entryj& e = bw.corr[c];
				if (e.imgIdx_i == 0xFFFFFFFFu) continue;
				BT_REQUIRE(e.imgIdx_i < (uint32_t)N && e.imgIdx_j < (uint32_t)N, BT_ERR_INVALID_ARG, "window %d: correspondence %d references frame outside [0,%d)", w, c, N);
				bin[(size_t)e.imgIdx_i * N + e.imgIdx_j + 1]++;
				n_valid++;
			}
			for (size_t b = 0; b < (size_t)N * N; b++) {
				if (bin[b + 1] > 0) { hgi[g_off + ng] = (int)(b / N); hgj[g_off + ng] = (int)(b % N); ng++; }
			}
			for (size_t b = 0; b < (size_t)N * N; b++) bin[b + 1] += bin[b];
			{   // group starts (relative to corr_off); stored at grp_off + w + g to leave room for the extra end entry
				int gg = 0;
				for (size_t b = 0; b < (size_t)N * N; b++) if (bin[b + 1] > bin[b]) hgs[g_off + w + gg++] = bin[b];
				hgs[g_off + w + ng] = n_valid;
			}
			for (int c = 0; c < bw.n_corr; c++) {
				const bt_entryj& e = bw.corr[c];
				if (e.imgIdx_i == 0xFFFFFFFFu) continue;
				hcorr[c_off + bin[(size_t)e.imgIdx_i * N + e.imgIdx_j]++] = e;
			}
		}
		WinSparse& ws = hws[w];
		memset(&ws, 0, sizeof ws);
		ws.n_corr = n_valid; ws.n_groups = ng; ws.corr_off = (int)c_off; ws.grp_off = (int)g_off; ws.mem_off = (int)m_off;
		{   // cross blocks are written without atomics when every unordered frame pair occurs at most once per table
			seen.assign((size_t)N * N, 0);
			bool uq = true;
			for (int g = 0; g < ng && uq; g++) { const int a2 = std::min(hgi[g_off + g], hgj[g_off + g]), b2 = std::max(hgi[g_off + g], hgj[g_off + g]); if (seen[(size_t)a2 * N + b2]) uq = false; seen[(size_t)a2 * N + b2] = 1; }
			std::fill(seen.begin(), seen.end(), 0);
			for (int p = 0; p < np && uq; p++) { const int a2 = (int)std::min(hpairs[p_off + p].x, hpairs[p_off + p].y), b2 = (int)std::max(hpairs[p_off + p].x, hpairs[p_off + p].y); if (seen[(size_t)a2 * N + b2]) uq = false; seen[(size_t)a2 * N + b2] = 1; }
			ws.unique_blocks = uq ? 1 : 0;
		}
		{   // per-frame membership CSR: groups then pairs touching each frame, in increasing index order (fixed summation order)
			int* fg_start = hmem + m_off; int* fp_start = fg_start + (N + 1); int* fg_items = fp_start + (N + 1); int* fp_items = fg_items + 2 * ng;
			int q = 0;
			for (int f = 0; f < N; f++) {
				fg_start[f] = q;
				for (int g = 0; g < ng; g++) {
					if (hgi[g_off + g] == f) fg_items[q++] = g;
					else if (hgj[g_off + g] == f) fg_items[q++] = g | (1 << 16);
				}
			}
			fg_start[N] = q;
			q = 0;
			for (int f = 0; f < N; f++) {
				fp_start[f] = q;
				for (int p = 0; p < np; p++) {
					if ((int)hpairs[p_off + p].x == f) fp_items[q++] = p;
					else if ((int)hpairs[p_off + p].y == f) fp_items[q++] = p | (1 << 16);
				}
			}
			fp_start[N] = q;
			m_off += (size_t)(2 * (N + 1) + 2 * ng + 2 * np);
		}
		smem_need = std::max(smem_need, tail_smem_floats(N, np, ng) * sizeof(float));
		smem_lean = std::max(smem_lean, tail_smem_floats(N, np, ng, false) * sizeof(float));
		c_off += n_valid; g_off += ng; p_off += np;
		// upload what has accumulated once it is worth a copy (>= 256 KB), and whatever is left after the last window
		if ((c_off - sent) * sizeof(bt_entryj) >= 256 * 1024 || (w + 1 == n_windows && c_off > sent)) {
			BT_CUDA(cudaMemcpyAsync(stage_dev.as<char>() + L.corr + sent * sizeof(bt_entryj), hcorr + sent, (c_off - sent) * sizeof(bt_entryj), cudaMemcpyHostToDevice, s->copy_stream));
			sent = c_off;
		}
	}
	if (run_n) BT_CUDA(cudaMemcpyAsync(stage_dev.as<char>() + L.corr + run_dst * sizeof(bt_entryj), run_src, run_n * sizeof(bt_entryj), cudaMemcpyHostToDevice, s->copy_stream));
	s->grp_in_smem = smem_need <= kTailSmemMax;
	if (!s->grp_in_smem) smem_need = smem_lean;
	BT_REQUIRE(smem_need <= kTailSmemMax, BT_ERR_CAPACITY, "bt_solve_stage: window needs %zu bytes of shared memory (> %d KB)", smem_need, (int)(kTailSmemMax / 1024));
	s->smem_bytes = (int)smem_need;
	BT_CUDA(cudaMemcpyAsync(stage_dev.as<char>() + L.late, hb + L.late, L.corr - L.late, cudaMemcpyHostToDevice, s->copy_stream));      // group tables, CSR, WinSparse
	BT_CUDA(cudaEventRecord(s->ev_corr, s->copy_stream));
	BT_CUDA(cudaEventRecord(s->ev_h2d, s->copy_stream));
	BT_CUDA(cudaStreamWaitEvent(stream, s->ev_corr, 0));     // everything later on the caller's stream sees the correspondences
	if (s->debug) {
		const int st = 6 * s->lim.max_frames;
		if ((rc = s->dbgJ.alloc(sizeof(float) * (size_t)st * st * s->lim.max_windows)) != BT_OK) return rc;
		if ((rc = s->dbgR.alloc(sizeof(float) * (size_t)st * s->lim.max_windows)) != BT_OK) return rc;
		if ((rc = s->dbgC.alloc(sizeof(float) * (size_t)s->max_pairs * s->lim.max_windows)) != BT_OK) return rc;
	}
	s->staged = true;
	s->host_us[2] = us_since(t_h0) - s->host_us[0] - s->host_us[1];
	return BT_OK;
}

extern "C" int bt_solve_stage(bt_ctx* ctx, int n_windows, const bt_window* windows, const bt_solver_params* params,
                              const float* poses_in, void* stream_) {
	return stage_impl(ctx, n_windows, windows, params, poses_in, stream_, false);
}

static int ensure_occupancy(bt_ctx* ctx) {
	SolverState* s = ctx->solver;
	if (s->attr_bytes == 0) {      // the attribute belongs to (function, device), not to the context: set it to the ceiling stage_impl enforces, once,
		s->attr_bytes = (int)kTailSmemMax;   // so that a second context on the same device can never lower what another context's next launch needs
		BT_CUDA(cudaFuncSetAttribute(k_solve<256, BT_SOLVE_MIN_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, s->attr_bytes));
	}
	if (s->occ_smem != s->smem_bytes) {
		int occ_q = 1;
		BT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_q, k_solve<256, BT_SOLVE_MIN_CTAS>, 256, s->smem_bytes));
		s->occ = std::max(occ_q, 1); s->occ_smem = s->smem_bytes;
	}
	return BT_OK;
}

static SolveArgs make_args(bt_ctx* ctx) {
	SolverState* s = ctx->solver;
	SolveArgs a;
	memset(&a, 0, sizeof a);
	char* sd = s->stage_buf[s->stage_cur].as<char>();
	const StageLayout& L = s->layout;
	a.wins = (WinDesc*)(sd + L.wins); a.wsp = (const WinSparse*)(sd + L.wsp); a.n_windows = s->n_windows;
	a.depth_ptr = (const float**)(sd + L.dp); a.normal_ptr = (const float4**)(sd + L.np); a.frame_win = (int*)(sd + L.fw);
	a.texel = s->texel.as<float4>(); a.src = s->src.as<float4>(); a.nsrc = s->nsrc.as<int>(); a.grp_sums = s->grp_sums.as<float>();
	a.texel_tab = (const float4* const*)(sd + L.texp); a.src_tab = (const float4* const*)(sd + L.srcp); a.nsrc_cached = (const int* const*)(sd + L.nsrcp);
	a.pose_in = (float*)(sd + L.pose); a.x = s->x.as<float>(); a.T = s->T.as<float>(); a.pose_out = s->pose_out.as<float>();
	a.npix_max = s->npix_max;
	a.corr = (bt_entryj*)(sd + L.corr); a.grp_i = (int*)(sd + L.gi); a.grp_j = (int*)(sd + L.gj); a.grp_start = (int*)(sd + L.gs);
	a.pairs = (uint2*)(sd + L.pairs); a.pair_win = (int*)(sd + L.pwin); a.pair_src_slot = (int*)(sd + L.psrc); a.mem = (int*)(sd + L.mem);
	a.tiles = s->tiles.as<Tile>(); a.max_tiles = s->max_tiles; a.partial = s->partial.as<float>();
	a.pair_tile0 = s->pair_tile0.as<int>(); a.pair_ntile = s->pair_ntile.as<int>();
	a.n_tiles_total = s->scalars.as<int>(); a.queue = s->scalars.as<int>() + 1; a.n_src_px = (long long*)(s->scalars.as<char>() + 16);
	a.tiles_done = s->tiles_done.as<int>(); a.iter_done = s->iter_done.as<int>();
	a.chunk = s->chunk; a.grp_in_smem = s->grp_in_smem ? 1 : 0; a.prm = s->prm; a.grid_ctas = ctx->sm_count * s->occ;
	{   // `sqrtf(d2) <= dist_t` <=> `d2 <= d2max` for every float d2 (sqrtf is correctly rounded and monotonic)
		const float th = s->prm.dense_dist_thresh;
		float d2max = th * th;
		auto nextf = [](float v, int d) { int32_t b; memcpy(&b, &v, 4); b += d; float r; memcpy(&r, &b, 4); return r; };
		if (th > 0.f) {
			while (d2max > 0.f && sqrtf(d2max) > th) d2max = nextf(d2max, -1);
			while (sqrtf(nextf(d2max, 1)) <= th) d2max = nextf(d2max, 1);
		} else d2max = (th == 0.f) ? 0.f : -1.f;
		a.d2max = d2max;
	}
	if (s->prof_cap > 0) { a.prof = s->prof.as<long long>(); a.prof_cap = s->prof_cap; }
	if (s->debug) { a.dbg_JtJ = s->dbgJ.as<float>(); a.dbg_Jtr = s->dbgR.as<float>(); a.dbg_stride = 6 * s->lim.max_frames; a.dbg_cnt = s->dbgC.as<float>(); a.dbg_cnt_stride = s->max_pairs; }
	return a;
}

static int launch_prep(bt_ctx* ctx, cudaStream_t stream) {
	SolverState* s = ctx->solver;
	// (in the fused call the tail's shared-memory size is not known yet: the tile plan then aims at the occupancy of the last batch)
	SolveArgs a = make_args(ctx);
	if (s->timing) BT_CUDA(cudaEventRecord(s->ev[0], stream));
	// frames whose maps come from the frame cache need one block each (source count + pose); a batch of cached frames only skips the count kernel
	const dim3 pgrid(s->any_uncached ? (unsigned)((s->npix_max + 1023) / 1024) : 1u, (unsigned)s->frames_total);
	if (s->any_uncached) k_prep_count<<<pgrid, 1024, 0, stream>>>(a, s->blk_cnt.as<int>());
	k_prep_frames<<<pgrid, 1024, 0, stream>>>(a, const_cast<WinDesc*>(a.wins), s->scalars.as<int>() + 8, s->blk_cnt.as<int>());
	BT_CUDA(cudaGetLastError());
	if (s->timing) BT_CUDA(cudaEventRecord(s->ev[1], stream));
	s->prep_launched = true;
	return BT_OK;
}

extern "C" int bt_solve_run(bt_ctx* ctx, void* stream_) {
	BT_REQUIRE(ctx && ctx->solver && ctx->solver->staged, BT_ERR_INVALID_ARG, "bt_solve_run: nothing staged");
	SolverState* s = ctx->solver;
	cudaStream_t stream = (cudaStream_t)stream_;
	BT_CUDA(cudaSetDevice(ctx->device));
	{ int rco = ensure_occupancy(ctx); if (rco != BT_OK) return rco; }
	SolveArgs a = make_args(ctx);
	if (s->debug) {
		BT_CUDA(cudaMemsetAsync(s->dbgJ.p, 0, s->dbgJ.bytes, stream));
		BT_CUDA(cudaMemsetAsync(s->dbgR.p, 0, s->dbgR.bytes, stream));
	}
	if (s->prof_cap > 0) BT_CUDA(cudaMemsetAsync(s->prof.p, 0, 8, stream));
	if (!s->prep_launched) { int rcp = launch_prep(ctx, stream); if (rcp != BT_OK) return rcp; }
	s->prep_launched = false;       // a second bt_solve_run on the same staged batch prepares again (poses restart from the staged input)
	if (s->timing) BT_CUDA(cudaEventRecord(s->ev[2], stream));
	const int occ = s->occ;
	const int grid = ctx->sm_count * occ;
	k_solve<256, BT_SOLVE_MIN_CTAS><<<grid, 256, s->smem_bytes, stream>>>(a);
	BT_CUDA(cudaGetLastError());
	if (s->ev_stage_free[s->stage_cur]) BT_CUDA(cudaEventRecord(s->ev_stage_free[s->stage_cur], stream));      // this batch's staging block may be overwritten after this point
	if (s->timing) BT_CUDA(cudaEventRecord(s->ev[3], stream));
	s->launches = s->any_uncached ? 3 : 2;
	return BT_OK;
}

extern "C" int bt_solve_fetch(bt_ctx* ctx, float* poses_out, void* stream_) {
	BT_REQUIRE(ctx && ctx->solver && ctx->solver->staged && poses_out, BT_ERR_INVALID_ARG, "bt_solve_fetch: nothing staged / NULL output");
	SolverState* s = ctx->solver;
	cudaStream_t stream = (cudaStream_t)stream_;
	const size_t bytes = sizeof(float) * 16 * (size_t)s->frames_total;
	int rc = s->h_poses.alloc(bytes + 64);
	if (rc != BT_OK) return rc;
	BT_CUDA(cudaMemcpyAsync(s->h_poses.p, s->pose_out.p, bytes, cudaMemcpyDeviceToHost, stream));
	BT_CUDA(cudaMemcpyAsync(s->h_poses.as<char>() + bytes, s->scalars.p, 32, cudaMemcpyDeviceToHost, stream));
	BT_CUDA(cudaStreamSynchronize(stream));
	const int total = *(int*)(s->h_poses.as<char>() + bytes);
	BT_REQUIRE(total >= 0, BT_ERR_CAPACITY, "bt_solve_fetch: tile list overflow (more dense tiles than reserved)");
	memcpy(poses_out, s->h_poses.p, bytes);
	return BT_OK;
}

extern "C" int bt_solve_windows(bt_ctx* ctx, int n_windows, const bt_window* windows, const bt_solver_params* params,
                                float* poses_inout, void* stream) {
	const auto t0 = std::chrono::steady_clock::now();
	auto us = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); };
	int rc = stage_impl(ctx, n_windows, windows, params, poses_inout, stream, true);
	if (rc != BT_OK) return rc;
	const double t1 = us();
	rc = bt_solve_run(ctx, stream);
	if (rc != BT_OK) return rc;
	const double t2 = us();
	rc = bt_solve_fetch(ctx, poses_inout, stream);
	SolverState* s = ctx->solver;
	s->host_us[3] = t2 - t1; s->host_us[4] = us() - t2; s->host_us[5] = us();
	return rc;
}

// Streaming form of bt_solve_windows: `begin` stages and launches a batch and queues the download of its poses, `end` waits for the
// OLDEST batch begun and hands its poses out.  Up to two batches may be in flight, so the host side of batch k+1 (tables, grouping,
// uploads, launches) overlaps the GPU side of batch k; everything is ordered on `stream`.
extern "C" int bt_solve_windows_begin(bt_ctx* ctx, int n_windows, const bt_window* windows, const bt_solver_params* params, const float* poses_in, void* stream_) {
	BT_REQUIRE(ctx && ctx->solver, BT_ERR_INVALID_ARG, "bt_solve_windows_begin: call bt_solver_reserve first");
	SolverState* s = ctx->solver;
	const int slot = s->pipe_head;
	BT_REQUIRE(!s->pipe_pending[slot], BT_ERR_INVALID_ARG, "bt_solve_windows_begin: two batches are already in flight - call bt_solve_windows_end first");
	int rc = stage_impl(ctx, n_windows, windows, params, poses_in, stream_, true);
	if (rc != BT_OK) return rc;
	rc = bt_solve_run(ctx, stream_);
	if (rc != BT_OK) return rc;
	cudaStream_t stream = (cudaStream_t)stream_;
	const size_t bytes = sizeof(float) * 16 * (size_t)s->frames_total;
	if ((rc = s->h_pipe[slot].alloc(bytes + 64)) != BT_OK) return rc;
	if (!s->ev_pipe[slot]) BT_CUDA(cudaEventCreateWithFlags(&s->ev_pipe[slot], cudaEventDisableTiming));
	BT_CUDA(cudaMemcpyAsync(s->h_pipe[slot].p, s->pose_out.p, bytes, cudaMemcpyDeviceToHost, stream));
	BT_CUDA(cudaMemcpyAsync(s->h_pipe[slot].as<char>() + bytes, s->scalars.p, 32, cudaMemcpyDeviceToHost, stream));
	BT_CUDA(cudaEventRecord(s->ev_pipe[slot], stream));
	s->pipe_pending[slot] = true; s->pipe_frames[slot] = s->frames_total; s->pipe_head ^= 1;
	return BT_OK;
}
extern "C" int bt_solve_windows_end(bt_ctx* ctx, float* poses_out) {
	BT_REQUIRE(ctx && ctx->solver && poses_out, BT_ERR_INVALID_ARG, "bt_solve_windows_end: NULL argument");
	SolverState* s = ctx->solver;
	const int slot = s->pipe_tail;
	BT_REQUIRE(s->pipe_pending[slot], BT_ERR_INVALID_ARG, "bt_solve_windows_end: no batch in flight");
	BT_CUDA(cudaEventSynchronize(s->ev_pipe[slot]));
	s->pipe_pending[slot] = false; s->pipe_tail ^= 1;
	const size_t bytes = sizeof(float) * 16 * (size_t)s->pipe_frames[slot];
	const int total = *(int*)(s->h_pipe[slot].as<char>() + bytes);
	BT_REQUIRE(total >= 0, BT_ERR_CAPACITY, "bt_solve_windows_end: tile list overflow (more dense tiles than reserved)");
	memcpy(poses_out, s->h_pipe[slot].p, bytes);
	return BT_OK;
}

// Host time of the last bt_solve_windows call, microseconds: {window tables + first upload, frame-prep launch, correspondence scan +
// staging + uploads, k_solve launch, pose download + wait for the GPU, whole call}.
extern "C" int bt_solve_get_host_timing(bt_ctx* ctx, double* us6) {
	BT_REQUIRE(ctx && ctx->solver && us6, BT_ERR_INVALID_ARG, "bt_solve_get_host_timing: NULL argument");
	for (int k = 0; k < 6; k++) us6[k] = ctx->solver->host_us[k];
	return BT_OK;
}

extern "C" int bt_frame_cache_reserve(bt_ctx* ctx, int capacity, int H, int W, float image_downscale) {
	BT_REQUIRE(ctx && ctx->solver, BT_ERR_INVALID_ARG, "bt_frame_cache_reserve: call bt_solver_reserve first");
	BT_REQUIRE(capacity > 0 && H > 0 && W > 0 && image_downscale >= 1.0f, BT_ERR_INVALID_ARG, "bt_frame_cache_reserve: bad arguments");
	SolverState* s = ctx->solver;
	BT_CUDA(cudaSetDevice(ctx->device));
	const int w = (int)(W / image_downscale), h = (int)(H / image_downscale);
	BT_REQUIRE(w >= 2 && h >= 2, BT_ERR_INVALID_ARG, "bt_frame_cache_reserve: image too small");
	int rc;
	if ((rc = s->c_texel.alloc(sizeof(float4) * 2 * (size_t)w * h * capacity)) != BT_OK) return rc;
	if ((rc = s->c_src.alloc(sizeof(float4) * 2 * (size_t)w * h * capacity)) != BT_OK) return rc;
	if ((rc = s->c_nsrc.alloc(sizeof(int) * (size_t)capacity)) != BT_OK) return rc;
	if ((rc = s->c_blk.alloc(sizeof(int) * (size_t)capacity * ((w * h + 1023) / 1024))) != BT_OK) return rc;
	s->c_capacity = capacity; s->c_npix = w * h; s->c_downscale = image_downscale;
	s->c_meta.assign(capacity, SolverState::CacheMeta());
	return BT_OK;
}

extern "C" int bt_frame_cache_store(bt_ctx* ctx, int n_frames, const int32_t* slots, const float* const* depth_dev, const float* const* normal_dev,
                                    int H, int W, float fx, float fy, float cx, float cy, float depth_min, float depth_max, void* stream_) {
	BT_REQUIRE(ctx && ctx->solver && slots && depth_dev && normal_dev, BT_ERR_INVALID_ARG, "bt_frame_cache_store: NULL argument");
	SolverState* s = ctx->solver;
	BT_REQUIRE(s->c_capacity > 0, BT_ERR_INVALID_ARG, "bt_frame_cache_store: call bt_frame_cache_reserve first");
	BT_REQUIRE(n_frames > 0 && n_frames <= s->c_capacity, BT_ERR_CAPACITY, "bt_frame_cache_store: %d frames > capacity %d", n_frames, s->c_capacity);
	const int w = (int)(W / s->c_downscale), h = (int)(H / s->c_downscale);
	BT_REQUIRE(H > 0 && W > 0 && w >= 2 && h >= 2 && w * h <= s->c_npix, BT_ERR_CAPACITY, "bt_frame_cache_store: image %dx%d does not fit the reserved cache", W, H);
	cudaStream_t stream = (cudaStream_t)stream_;
	BT_CUDA(cudaSetDevice(ctx->device));
	const size_t b_ptr = sizeof(void*) * (size_t)n_frames, bytes = 2 * b_ptr + sizeof(int) * (size_t)n_frames;
	int rc;
	if ((rc = s->c_tables.alloc(bytes)) != BT_OK) return rc;
	const int fl = s->c_flip; s->c_flip ^= 1;
	if (s->c_ev[fl]) BT_CUDA(cudaEventSynchronize(s->c_ev[fl]));
	if ((rc = s->c_htables[fl].alloc(bytes)) != BT_OK) return rc;
	if (!s->c_ev[fl]) BT_CUDA(cudaEventCreateWithFlags(&s->c_ev[fl], cudaEventDisableTiming));
	char* hb = s->c_htables[fl].as<char>();
	const void** hd = (const void**)hb; const void** hn = (const void**)(hb + b_ptr); int* hs = (int*)(hb + 2 * b_ptr);
	for (int f = 0; f < n_frames; f++) {
		BT_REQUIRE(slots[f] >= 0 && slots[f] < s->c_capacity, BT_ERR_INVALID_ARG, "bt_frame_cache_store: slot %d out of range [0,%d)", slots[f], s->c_capacity);
		BT_REQUIRE(depth_dev[f] && normal_dev[f], BT_ERR_INVALID_ARG, "bt_frame_cache_store: frame %d has a NULL map", f);
		for (int g = 0; g < f; g++) BT_REQUIRE(slots[g] != slots[f], BT_ERR_INVALID_ARG, "bt_frame_cache_store: slot %d listed twice", slots[f]);
		hd[f] = depth_dev[f]; hn[f] = normal_dev[f]; hs[f] = slots[f];
	}
	BT_CUDA(cudaMemcpyAsync(s->c_tables.p, hb, bytes, cudaMemcpyHostToDevice, stream));
	BT_CUDA(cudaEventRecord(s->c_ev[fl], stream));
	CacheStoreArgs c;
	c.g.W = W; c.g.H = H; c.g.w = w; c.g.h = h;                                                   // same expressions as bt_solve_stage (CUDACache.cpp:20-24)
	c.g.ifx = 1.0f / fx; c.g.ify = 1.0f / fy; c.g.icx = -cx / fx; c.g.icy = -cy / fy;
	c.g.scaleW = (float)(W - 1) / (float)(w - 1); c.g.scaleH = (float)(H - 1) / (float)(h - 1);
	char* db = s->c_tables.as<char>();
	c.depth = (const float* const*)db; c.normal = (const float4* const*)(db + b_ptr); c.slots = (const int*)(db + 2 * b_ptr);
	c.texel = s->c_texel.as<float4>(); c.src = s->c_src.as<float4>(); c.nsrc = s->c_nsrc.as<int>(); c.slot_stride = 2 * (size_t)s->c_npix;
	c.dmin = depth_min; c.dmax = depth_max;
	if (n_frames >= 2 * ctx->sm_count) k_frame_cache_store<<<n_frames, 1024, 0, stream>>>(c);      // many frames: one CTA per frame fills the machine
	else {
		const dim3 grid((unsigned)((w * h + 1023) / 1024), (unsigned)n_frames);
		k_cache_count<<<grid, 1024, 0, stream>>>(c, s->c_blk.as<int>());
		k_cache_build<<<grid, 1024, 0, stream>>>(c, s->c_blk.as<int>());
	}
	BT_CUDA(cudaGetLastError());
	for (int f = 0; f < n_frames; f++) {
		SolverState::CacheMeta& m = s->c_meta[slots[f]];
		m.valid = true; m.H = H; m.W = W; m.fx = fx; m.fy = fy; m.cx = cx; m.cy = cy; m.dmin = depth_min; m.dmax = depth_max;
	}
	return BT_OK;
}

extern "C" int bt_solve_get_stats(bt_ctx* ctx, bt_solve_stats* out) {
	BT_REQUIRE(ctx && ctx->solver && out, BT_ERR_INVALID_ARG, "bt_solve_get_stats: NULL argument");
	SolverState* s = ctx->solver;
	char h[32];
	BT_CUDA(cudaMemcpy(h, s->scalars.p, 32, cudaMemcpyDeviceToHost));
	out->n_windows = s->n_windows;
	out->n_tiles_total = *(int*)h;
	out->n_kernel_launches = s->launches;
	out->n_src_pixels = *(long long*)(h + 16);
	return BT_OK;
}

extern "C" int bt_solve_enable_profile(bt_ctx* ctx, int max_records) {
	BT_REQUIRE(ctx && ctx->solver && max_records >= 0, BT_ERR_INVALID_ARG, "bt_solve_enable_profile: call bt_solver_reserve first");
	SolverState* s = ctx->solver;
	if (max_records > 0) { int rc = s->prof.alloc(8 + (size_t)max_records * 96); if (rc != BT_OK) return rc; }
	s->prof_cap = max_records;
	return BT_OK;
}
extern "C" int bt_solve_get_profile(bt_ctx* ctx, long long* out, int max_records) {
	BT_REQUIRE(ctx && ctx->solver && ctx->solver->prof_cap > 0 && out, BT_ERR_INVALID_ARG, "bt_solve_get_profile: profiling not enabled");
	SolverState* s = ctx->solver;
	BT_CUDA(cudaDeviceSynchronize());
	long long n = 0;
	BT_CUDA(cudaMemcpy(&n, s->prof.p, 8, cudaMemcpyDeviceToHost));
	if (n > s->prof_cap) n = s->prof_cap;
	if (n > max_records) n = max_records;
	if (n > 0) BT_CUDA(cudaMemcpy(out, s->prof.as<long long>() + 1, (size_t)n * 96, cudaMemcpyDeviceToHost));
	return (int)n;
}

extern "C" int bt_solve_debug_counts(bt_ctx* ctx, int w, int n_pairs, float* counts_out) {
	BT_REQUIRE(ctx && ctx->solver && ctx->solver->debug && ctx->solver->staged && counts_out, BT_ERR_INVALID_ARG, "bt_solve_debug_counts: debug not enabled / nothing run");
	SolverState* s = ctx->solver;
	BT_REQUIRE(w >= 0 && w < s->n_windows && n_pairs >= 0 && n_pairs <= s->max_pairs, BT_ERR_INVALID_ARG, "bt_solve_debug_counts: bad index");
	BT_CUDA(cudaDeviceSynchronize());
	BT_CUDA(cudaMemcpy(counts_out, s->dbgC.as<float>() + (size_t)w * s->max_pairs, sizeof(float) * n_pairs, cudaMemcpyDeviceToHost));
	return BT_OK;
}

extern "C" int bt_solve_debug_dense(bt_ctx* ctx, int w, float* JtJ_out, float* Jtr_out) {
	BT_REQUIRE(ctx && ctx->solver && ctx->solver->debug && ctx->solver->staged, BT_ERR_INVALID_ARG, "bt_solve_debug_dense: debug not enabled / nothing run");
	SolverState* s = ctx->solver;
	BT_REQUIRE(w >= 0 && w < s->n_windows, BT_ERR_INVALID_ARG, "bt_solve_debug_dense: window %d out of range", w);
	const int st = 6 * s->lim.max_frames, dim = 6 * s->n_frames[w];
	BT_CUDA(cudaDeviceSynchronize());
	if (JtJ_out) BT_CUDA(cudaMemcpy(JtJ_out, s->dbgJ.as<float>() + (size_t)w * st * st, sizeof(float) * dim * dim, cudaMemcpyDeviceToHost));
	if (Jtr_out) BT_CUDA(cudaMemcpy(Jtr_out, s->dbgR.as<float>() + (size_t)w * st, sizeof(float) * dim, cudaMemcpyDeviceToHost));
	return BT_OK;
}
