// knn.cu — all-pairs L2 k-nearest-neighbour matching of 256-d LF-Net descriptors on the 5th-gen tensor cores.
//
// Replaces the two cv::cuda::DescriptorMatcher::knnMatch(…, k=5) calls of SiftManager::findCorresbyNN
// (/root/reference/src/FeatureManager.cpp:271-273) for a BATCH of frame pairs.  OpenCV's CUDA matcher materialises the nQ x nT
// fp32 distance matrix with a SIMT kernel, once per direction, and runs k row-min passes over it.  Here ONE
// tensor-core contraction per pair serves BOTH directions and the distance matrix is never written:
//
//   k_desc_prep    fp32 (pitched GpuMat rows) -> a pool slot: fp16 rows [rows padded to 256][256] for the tensor pass, an fp32
//                  copy for the exact stage, |x~|^2 of the ROUNDED rows (padding rows: zero data, norm +inf) and the measured
//                  rounding error |x~ - x| of every row.  Done once per descriptor set - once per FRAME with the persistent
//                  pool (bt_desc_pool_store at detectFeature time, FeatureManager.cpp:907), once per call otherwise.
//   k_knn_tc       persistent, warp-specialised tcgen05 kernel over the flattened list of (pair, 256-row B tile, 128-row A tile)
//                  units, cut into equal contiguous ranges (one per SM: no wave quantisation).  Warp 0 keeps the B tile resident
//                  in shared memory (128 KB, reloaded K-chunk by K-chunk while the last unit of the previous tile still runs)
//                  and streams A tiles through a 5-stage ring with TMA (SWIZZLE_128B, K-major); warp 1 issues
//                  tcgen05.mma.cta_group::1.kind::f16 (fp16 x fp16 -> fp32, M=128, N=256, K=16) into a double-buffered TMEM
//                  accumulator; eight epilogue warps read it with tcgen05.ld (thread = A row, registers = 32 B columns), form
//                  d2 = |a~|^2 + |b~|^2 - 2 a~.b~ (scaled into [0,1] by a per-pair power of two) and reduce it BOTH ways:
//                    A->B  min over each group of 4 consecutive B columns, in registers            -> G_rowT[B group][A row]
//                    B->A  min over each group of 4 consecutive A rows: two butterfly steps of
//                          packed-fp16 shuffles across the lanes that hold those rows             -> G_col [A row group][B row]
//                  both stored as fp16 (round to nearest; the bound below carries the half ulp).  No data-dependent branch, no
//                  candidate lists, ~5 issue slots per matrix element: the kernel is paced by the tensor pipe.
//   k_knn_select   per query (lane) over its column of group minima: exact k-th smallest key -> a rigorous threshold (measured
//                  fp16 rounding error of both rows + accumulation slack + storage rounding) -> the groups that can still hold
//                  one of the k nearest; each member of such a group is then checked against the OTHER direction's matrix (the
//                  minimum over the 4-row group that contains the query is a lower bound of the member's own distance: 8 bytes
//                  instead of re-scoring four 512-byte rows) -> candidate list, typically k + 0..1 rows.
//   k_knn_rerank   exact float64 sum of (a-b)^2 on the fp32 rows of the candidates, top k by (distance, index).
//   k_knn_exact    exact brute force for the rows whose candidate set overflowed (degenerate inputs: many equal distances).
// Result == exact brute-force kNN (float64 distances, ties -> lower train index), distances returned as float(sqrt(d2)) like
// cv::NORM_L2.
#include <algorithm>
#include <cuda.h>
#include <cuda_fp16.h>
#include <map>
#include <math.h>
#include <stdlib.h>
#include "bt_common.cuh"

namespace bt {

static constexpr int KD = 256;              // descriptor dimension handled by the tensor path
static constexpr int BM = 128;              // A rows per unit (UMMA M = TMEM lanes)
static constexpr int BN = 256;              // B rows per unit (UMMA N = TMEM columns)
static constexpr int BK = 64;               // K elements per smem chunk: 64 fp16 = 128 B = one SWIZZLE_128B atom row
static constexpr int KCH = KD / BK;         // 4 chunks
static constexpr int STAGES = 5;            // A-operand pipeline depth (16 KB each)
static constexpr int GRP = 4;               // rows / columns per group minimum
static constexpr int EPI_WARPS = 16;        // 4 per TMEM lane quarter, 64 accumulator columns each: 4 epilogue warps per scheduler hide the shuffle / TMEM-load latencies
static constexpr int KNN_THREADS = 32 * (2 + EPI_WARPS); // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue
static constexpr uint32_t SA_BYTES = BM * BK * 2;     // 16 KB per stage: one K-chunk of an A tile
static constexpr uint32_t SB_BYTES = BN * BK * 2;     // 32 KB per K-chunk of the resident B tile
static constexpr uint32_t SMEM_B = KCH * SB_BYTES;    // 128 KB
static constexpr uint32_t SMEM_A = STAGES * SA_BYTES; // 80 KB
static constexpr uint32_t SMEM_NORM = BN * 4;         // 1 KB
static constexpr uint32_t SMEM_BAR = 256;
static constexpr uint32_t SMEM_TOTAL = SMEM_B + SMEM_A + SMEM_NORM + SMEM_BAR + 1024;   // + alignment slack
// CTA-pair form (cta_group::2, M = 256): a CTA keeps HALF of the B tile (128 rows, 64 KB) and streams its own 128 A rows through a deeper ring
static constexpr int STAGES_P = 5;
static constexpr uint32_t SBH_BYTES = (BN / 2) * BK * 2;                 // 16 KB per K-chunk of the B half tile
static constexpr uint32_t SMEM_BH = KCH * SBH_BYTES;                     // 64 KB: one B half tile
static constexpr uint32_t SMEM_B_P = 2 * SMEM_BH;                        // 128 KB: DOUBLE-BUFFERED - the next B tile loads while the current one is used (its own producer warp)
static constexpr uint32_t SMEM_A_P = STAGES_P * SA_BYTES;                // 80 KB
static constexpr uint32_t SMEM_TOTAL_P = SMEM_B_P + SMEM_A_P + SMEM_NORM + SMEM_BAR + 1024;
static constexpr int KNN_THREADS_P = KNN_THREADS + 32;                   // + the B-tile producer warp
static constexpr int SEL_MAXG = 16;         // groups / candidates kept per query; more => exact fallback
static constexpr int SEL_MAXC = 16;

struct KnnPair {          // one non-empty frame pair: units (B tile tt, A tile qt), qt fastest
	int a_row0, b_row0;   // first pool row of the two descriptor sets
	int nA, nB;
	int n_qt, n_tt;
	int unit0;
	int nA_pad, nB_pad;   // n_qt * 128, n_tt * 256
	int setA, setB;       // pool slots (per-set maxima)
	int pad;
	long long g_row;      // offset (halves) of G_rowT [nB_pad/4][nA_pad]
	long long g_col;      // offset (halves) of G_col  [nA_pad/4][nB_pad]
};
struct PairConst { float s, inv_s, delta, emaxA, emaxB, pad0, pad1, pad2; };

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "WAIT_LOOP:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra WAIT_DONE;\n"
	    "bra WAIT_LOOP;\n"
	    "WAIT_DONE:\n"
	    "}\n" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
	             "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
	             : "memory");
}
// ---- CTA-pair (cta_group::2) variants: both CTAs of a pair issue their TMA loads against the LEADER's (rank 0) full barrier, the leader's
//      MMA thread commits to the same barrier offset in BOTH CTAs (multicast), the peer's epilogue arrives on the leader's barrier remotely
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
	asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
	             "l"(tmap), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)      // peer bit cleared: the transaction bytes land on CTA 0's barrier
	             : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
	asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((unsigned short)3) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "setp.ne.b32 p, %4, 0;\n"
	    "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
	    "}\n" ::"r"(d_tmem),
	    "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
	    : "memory");
}
__device__ __forceinline__ void mbar_arrive_rank0(uint64_t* bar) {      // arrive on the barrier at this offset in the cluster's CTA 0
	uint32_t remote;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(0u));
	asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");      // (default .release.cta: with .release.cluster the arrive queued behind the previous unit's global stores - 26 % of the kernel's stall samples)
}
__device__ __forceinline__ void cluster_sync_all() {
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "setp.ne.b32 p, %4, 0;\n"
	    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
	    "}\n" ::"r"(d_tmem),
	    "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
	    : "memory");
}
// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B, rows at 128-byte pitch, 8-row groups 1024 B apart.
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major) | [32,46) stride byte
//   offset >> 4 (= 1024 >> 4) | [46,48) descriptor version = 1 (sm_100) | [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
	uint64_t d = 0;
	d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
	d |= (uint64_t)(1024 >> 4) << 32;
	d |= (uint64_t)1 << 46;
	d |= (uint64_t)2 << 61;
	return d;
}
// UMMA instruction descriptor (kind::f16): c_format F32 (1) @4, a_format / b_format F16 (0) @7 / @10, a/b K-major (0) @15/@16,
// N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
	return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
#define TMEM_LD32(taddr, v)                                                                                                   \
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                     \
	             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                      \
	             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                      \
	             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),     \
	               "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),          \
	               "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),         \
	               "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                       \
	             : "r"(taddr))

// ------------------------------------------------------------------------------------------------ k_desc_prep
struct PrepSet { const float* src; size_t pitch_bytes; int n; int row0; int rows_padded; int slot; };

__global__ void __launch_bounds__(256) k_desc_prep(const PrepSet* sets, __half* pool_h, float* pool_f, float* norms, float* errn, int* slot_meta) {
	const PrepSet st = sets[blockIdx.y];
	const int warps_per_block = blockDim.x >> 5, lane = threadIdx.x & 31;
	for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < st.rows_padded; r += gridDim.x * warps_per_block) {
		__half* dst = pool_h + (size_t)(st.row0 + r) * KD;
		float4* dstf = reinterpret_cast<float4*>(pool_f + (size_t)(st.row0 + r) * KD);
		float ss = 0.f, ee = 0.f;
		if (r < st.n) {
			const float* src = (const float*)((const char*)st.src + (size_t)r * st.pitch_bytes);
			// each lane converts 8 consecutive floats (two float4 loads, one 16-byte store) and keeps the fp32 copy
			const float4 a = __ldg(reinterpret_cast<const float4*>(src) + lane * 2), b = __ldg(reinterpret_cast<const float4*>(src) + lane * 2 + 1);
			dstf[lane * 2] = a; dstf[lane * 2 + 1] = b;
			const float v[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
			__half h[8];
#pragma unroll
			for (int k = 0; k < 8; k++) {
				// fp16 range: values beyond +-65504 are clamped, halves that would be subnormal are flushed (whatever the tensor core does
				// with subnormals, it then sees a zero); either way the difference lands in the measured error norm of the row
				float c = fminf(fmaxf(v[k], -65504.f), 65504.f);
				if (fabsf(c) < 6.1035156e-5f) c = 0.f;
				h[k] = __float2half_rn(c);
				const float f = __half2float(h[k]);
				ss += f * f; const float dlt = v[k] - f; ee += dlt * dlt;
			}
			*reinterpret_cast<uint4*>(dst + lane * 8) = *reinterpret_cast<const uint4*>(h);
		} else {
			*reinterpret_cast<uint4*>(dst + lane * 8) = make_uint4(0u, 0u, 0u, 0u);
			dstf[lane * 2] = make_float4(0.f, 0.f, 0.f, 0.f); dstf[lane * 2 + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, o); ee += __shfl_xor_sync(0xffffffffu, ee, o); }
		if (lane == 0) {
			norms[st.row0 + r] = (r < st.n) ? ss : __int_as_float(0x7f800000);
			const float en = sqrtf(ee) * 1.001f + 1e-7f;     // |x~ - x|, rounded up: the ACTUAL fp16 rounding error of this row
			errn[st.row0 + r] = (r < st.n) ? en : 0.f;
			if (r < st.n) { atomicMax(slot_meta + 2 * st.slot, __float_as_int(ss)); atomicMax(slot_meta + 2 * st.slot + 1, __float_as_int(en)); }   // non-negative floats order like ints
		}
	}
}

// per pair: the power-of-two scale that maps d2 into [0,1] for the fp16 group minima, the slack for the fp32 accumulation of the
// contraction and of the norms, and the largest row rounding errors of the two sets
__global__ void k_knn_pairconst(const KnnPair* __restrict__ pairs, int n_pairs, const int* __restrict__ slot_meta, PairConst* __restrict__ out) {
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n_pairs) return;
	const KnnPair pr = pairs[p];
	const float m2A = __int_as_float(slot_meta[2 * pr.setA]), m2B = __int_as_float(slot_meta[2 * pr.setB]);
	const float mx = 4.0f * fmaxf(m2A, m2B);          // d2 <= (|a| + |b|)^2 <= 4 max |x|^2
	float s = 1.0f;
	if (mx > 0.f && mx < 1e37f) {
		int e = ((__float_as_int(mx) >> 23) & 255) - 127;          // mx = m * 2^e, 1 <= m < 2
		if (__float_as_int(mx) & 0x7fffff) e++;                    // 2^e >= mx
		e = min(max(e, -100), 100);
		s = __int_as_float((127 - e) << 23);
	}
	PairConst c;
	c.s = s; c.inv_s = 1.0f / s;
	// |computed d2 - exact d2 of the rounded rows|: 272 fp32 accumulation steps of the tensor pipe (<= 2^-22 relative to the sum of
	// |products| <= |a||b| each), the two norms (256 terms each) and the epilogue's fma + add; bounded with room to spare
	c.delta = 1.5e-4f * (m2A + m2B) + 1e-30f;
	c.emaxA = __int_as_float(slot_meta[2 * pr.setA + 1]); c.emaxB = __int_as_float(slot_meta[2 * pr.setB + 1]);
	c.pad0 = c.pad1 = c.pad2 = 0.f;
	out[p] = c;
}

// ------------------------------------------------------------------------------------------------ k_knn_tc
// position in the flattened unit list; every role of the CTA walks the same sequence
struct UnitWalk {
	const KnnPair* pairs; int p; int tt, qt; KnnPair pr;
	__device__ __forceinline__ void init(const KnnPair* pairs_, int n_pairs, int u) {
		pairs = pairs_;
		int lo = 0, hi = n_pairs - 1;
		while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (pairs[mid].unit0 <= u) lo = mid; else hi = mid - 1; }
		p = lo; pr = pairs[p];
		const int l = u - pr.unit0;
		tt = l / pr.n_qt; qt = l - tt * pr.n_qt;
	}
	__device__ __forceinline__ void next() {      // (reads one element past the table after the very last unit: the host pads the table by one entry)
		if (++qt == pr.n_qt) { qt = 0; if (++tt == pr.n_tt) { tt = 0; ++p; pr = pairs[p]; } }
	}
};

__device__ __forceinline__ uint32_t h2_as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ uint32_t hmin2_u32(uint32_t a, uint32_t b) {
	const __half2 r = __hmin2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
	return h2_as_u32(r);
}

// ---- the epilogue's view of the accumulator -----------------------------------------------------------------------------------
// tcgen05.ld.16x256b hands a thread the m16n8 accumulator fragment (probed on a B200, scripts/probe/tmem_layout.cu): register 4j + e of
// thread T = TMEM lane base + T/4 + 8 (e >> 1), column 8j + 2 (T % 4) + (e & 1).  Two such loads (lanes +0..15, +16..31) give a thread
// FOUR rows (T/4 + 8k, k = 0..3) x the column pairs (8j + 2m, 8j + 2m + 1), m = T % 4.  Both group minima the matcher needs then stay
// inside the thread when a "group" is four indices 8 apart in a block of 32 - rows {r, r+8, r+16, r+24}, columns {c, c+8, c+16, c+24}:
//   B->A  min over the thread's four rows, per column                      (3 packed-fp16 min per column pair)
//   A->B  min over four column pairs 8 apart, per row                      (3 packed-fp16 min per two groups)
// no shuffle, no select, no cross-half step (the first version grouped 4 CONSECUTIVE rows / columns of the 32x32b layout: 24 selects +
// 12 shuffles + 24 fp32 min per 32 columns, and the ALU pipe, not the tensor pipe, paced the kernel).
// In POSITION space pos(i) = (i & ~31) + 4 (i % 8) + (i % 32) / 8 such a group is four CONSECUTIVE positions; both group-minimum matrices
// are stored by position, and a thread's results land on 8 / 16 contiguous bytes.  k_knn_select works in position space and converts
// back with unpos() where it touches real rows.
__host__ __device__ __forceinline__ int knn_unpos(int p) { return (p & ~31) + 8 * (p & 3) + ((p & 31) >> 2); }

#define TMEM_LD_16x256b_x8(taddr, v)                                                                                          \
	asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 "                                                                    \
	             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                      \
	             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                      \
	             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),     \
	               "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),          \
	               "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),         \
	               "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                       \
	             : "r"(taddr))

// one 16-lane load (two of the thread's rows x 8 column pairs): d2 (scaled) = |a~|^2 + |b~|^2 - 2 a~.b~ with packed f32x2 arithmetic,
// clamped at zero and rounded to fp16 pairs (cvt.rn.relu: a rounding-level negative d2 becomes +0, so the halves order like unsigned integers)
__device__ __forceinline__ void epi_rows(const uint32_t (&v)[32], const unsigned long long (&nb)[8], float nAs0, float nAs1, unsigned long long m2s2,
                                         uint32_t (&h0)[8], uint32_t (&h1)[8]) {
	unsigned long long na0, na1;
	asm("mov.b64 %0, {%1, %1};" : "=l"(na0) : "f"(nAs0));
	asm("mov.b64 %0, {%1, %1};" : "=l"(na1) : "f"(nAs1));
#pragma unroll
	for (int j = 0; j < 8; j++) {
		unsigned long long x0, x1;
		float lo, hi;
		asm("mov.b64 %0, {%1, %2};" : "=l"(x0) : "r"(v[4 * j]), "r"(v[4 * j + 1]));
		asm("mov.b64 %0, {%1, %2};" : "=l"(x1) : "r"(v[4 * j + 2]), "r"(v[4 * j + 3]));
		asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x0) : "l"(m2s2), "l"(x0), "l"(nb[j]));
		asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x1) : "l"(m2s2), "l"(x1), "l"(nb[j]));
		asm("add.rn.f32x2 %0, %1, %2;" : "=l"(x0) : "l"(x0), "l"(na0));
		asm("add.rn.f32x2 %0, %1, %2;" : "=l"(x1) : "l"(x1), "l"(na1));
		asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x0));
		asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(h0[j]) : "f"(hi), "f"(lo));
		asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x1));
		asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(h1[j]) : "f"(hi), "f"(lo));
	}
}

// developer aid: cycles the MMA-issuing thread of every CTA spent waiting, by cause (read with bt_knn_debug_prof)
__device__ long long g_knn_prof[160][6];

// PAIR = false: one CTA per (128 A rows x 256 B rows) unit.  PAIR = true: a CLUSTER OF TWO CTAs (one TPC) per (256 x 256) unit -
// tcgen05.mma.cta_group::2, M = 256: CTA r holds A rows [128 r, 128 r + 128) of the unit and HALF of the B tile; the leader (rank 0)
// issues every MMA, each SM's tensor core accumulates its own 128 rows x 256 columns in its own TMEM, the B operand is read from shared
// memory once per pair instead of once per SM, and the freed 64 KB hold a SECOND B buffer filled by its own producer warp (no bubble at
// a tile switch).  Measured on B200 (45 pairs x 2000^2): correct on the first run, but NOT faster - 0.083-0.088 ms vs 0.081 ms.  The MMA
// thread's own clock (bt_knn_debug_prof) shows why: per unit it waits ~100 cycles for the epilogue, ~110-250 for B and ~320 for A tiles in
// BOTH forms, and issues for ~2500 - the pass is paced by the tensor pipe at the clock the chip sustains under this load (~1.5 GHz: the
// kernel's 125 k cycles take 81 us), not by shared-memory bandwidth.  Kept behind BT_KNN_CTA_PAIRS=1 (tested), not the default.
template <bool PAIR> __global__ void __launch_bounds__(PAIR ? KNN_THREADS_P : KNN_THREADS, 1)
k_knn_tc(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const KnnPair* __restrict__ pairs, int n_pairs, int total_units,
         const float* __restrict__ norms, const PairConst* __restrict__ pconst, __half* __restrict__ G) {
	constexpr int NST = PAIR ? STAGES_P : STAGES;
	constexpr uint32_t B_CHUNK = PAIR ? SBH_BYTES : SB_BYTES;
	constexpr uint32_t B_BYTES = PAIR ? SMEM_B_P : SMEM_B, A_BYTES = PAIR ? SMEM_A_P : SMEM_A;
	constexpr int UM = PAIR ? 2 * BM : BM;      // A rows per unit
	extern __shared__ uint8_t smem_raw[];
	uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1024-B alignment
	uint8_t* sB = smem;
	uint8_t* sA = smem + B_BYTES;
	float* sNorm = reinterpret_cast<float*>(smem + B_BYTES + A_BYTES);
	uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_BYTES + A_BYTES + SMEM_NORM);
	uint64_t* b_full = bars + 0;               // [KCH]  resident B tile, per K-chunk
	uint64_t* b_empty = bars + KCH;            // [KCH]
	uint64_t* full = bars + 2 * KCH;           // [NST]
	uint64_t* empty = full + NST;              // [NST]
	uint64_t* tm_full = empty + NST;           // [2]
	uint64_t* tm_empty = tm_full + 2;          // [2]
	uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tm_empty + 2);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint32_t cr = 0;      // rank of this CTA in its pair
	if constexpr (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cr));
	if (warp == 0 && lane == 0) {
		asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
		asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
		for (int s = 0; s < KCH; s++) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
		for (int s = 0; s < NST; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
		for (int b = 0; b < 2; b++) { mbar_init(tm_full + b, 1); mbar_init(tm_empty + b, PAIR ? 2 * EPI_WARPS : EPI_WARPS); }
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if constexpr (PAIR) cluster_sync_all();      // the peer's barriers exist before anything can arrive on them
	if (warp == 1) {   // TMEM: all 512 columns (two 256-column accumulators); this kernel runs 1 CTA / SM
		if constexpr (PAIR) {
			asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
			asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
		} else {
			asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
			asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
		}
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_ptr_smem;
	// this CTA's (pair's) contiguous slice of the unit list (balanced to +-1 unit)
	const int wid_ = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, nw_ = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
	const int u0 = (int)(((long long)wid_ * total_units) / nw_), u1 = (int)(((long long)(wid_ + 1) * total_units) / nw_);

	if (warp == 0) {
		// ================================================= TMA producer
		if (lane == 0 && u0 < u1) {
			uint32_t stage = 0, sphase = 0, bphase = 0;
			UnitWalk w; w.init(pairs, n_pairs, u0);
			for (int u = u0; u < u1; u++) {
				const bool newb = (u == u0) || (w.qt == 0);
				for (int kc = 0; kc < KCH; kc++) {
					if constexpr (!PAIR) {
						if (newb) {      // the previous B tile's last MMAs of this K-chunk have retired: reload the chunk while the others still run
							mbar_wait(b_empty + kc, bphase ^ 1);
							mbar_expect_tx(b_full + kc, SB_BYTES);
							tma_load_2d(sB + kc * SB_BYTES, &tmap_b, b_full + kc, kc * BK, w.pr.b_row0 + w.tt * BN);
						}
					}      // (pair: the B tiles have their own producer warp and two buffers - this thread never blocks behind a tile switch)
					mbar_wait(empty + stage, sphase ^ 1);
					if constexpr (PAIR) {
						if (cr == 0) mbar_expect_tx(full + stage, 2 * SA_BYTES);
						tma_load_2d_pair(sA + stage * SA_BYTES, &tmap_a, full + stage, kc * BK, w.pr.a_row0 + w.qt * UM + (int)cr * BM);
					} else {
						mbar_expect_tx(full + stage, SA_BYTES);
						tma_load_2d(sA + stage * SA_BYTES, &tmap_a, full + stage, kc * BK, w.pr.a_row0 + w.qt * BM);
					}
					if (++stage == NST) { stage = 0; sphase ^= 1; }
				}
				if (newb) bphase ^= 1;
				if (u + 1 < u1) w.next();
			}
		}
	} else if (warp == 1) {
		// ================================================= MMA issuer (one thread)
		if (lane == 0 && u0 < u1 && cr == 0) {      // (pair: the leader issues for both SMs)
			constexpr uint32_t idesc = umma_idesc_f16(UM, BN);
			uint32_t stage = 0, sphase = 0, bphase = 0, tcount = 0, btile = 0;
			UnitWalk w; w.init(pairs, n_pairs, u0);
			long long t_tm = 0, t_b = 0, t_a = 0; const long long t_begin = clock64();
			for (int u = u0; u < u1; u++, tcount++) {
				const bool newb = (u == u0) || (w.qt == 0);
				const bool lastb = (u + 1 == u1) || (w.qt == w.pr.n_qt - 1);
				const uint32_t buf = tcount & 1, tphase = (tcount >> 1) & 1;
				long long t0 = clock64();
				mbar_wait(tm_empty + buf, tphase ^ 1);      // epilogue drained this accumulator
				t_tm += clock64() - t0;
				tc_fence_after();
				const uint32_t d_tmem = tmem_base + buf * BN;
				if constexpr (PAIR) { if (newb) { t0 = clock64(); mbar_wait(b_full + (btile & 1), (btile >> 1) & 1); t_b += clock64() - t0; } }      // the whole half tile (both CTAs') has landed
				for (int kc = 0; kc < KCH; kc++) {
					t0 = clock64();
					if constexpr (!PAIR) { if (newb) mbar_wait(b_full + kc, bphase); }
					long long t1 = clock64();
					mbar_wait(full + stage, sphase);
					t_b += t1 - t0; t_a += clock64() - t1;
					tc_fence_after();
					const uint32_t a_base = smem_u32(sA + stage * SA_BYTES), b_base = smem_u32(sB + (PAIR ? (btile & 1) * SMEM_BH : 0u) + kc * B_CHUNK);
#pragma unroll
					for (int k = 0; k < BK / 16; k++) {      // UMMA K = 16 halves = 32 bytes inside the 128-byte swizzle row
						if constexpr (PAIR) tc_mma_f16_pair(d_tmem, umma_desc_k128(a_base + k * 32), umma_desc_k128(b_base + k * 32), idesc, (uint32_t)((kc | k) != 0));
						else tc_mma_f16(d_tmem, umma_desc_k128(a_base + k * 32), umma_desc_k128(b_base + k * 32), idesc, (uint32_t)((kc | k) != 0));
					}
					if constexpr (PAIR) { tc_commit_pair(empty + stage); if (lastb && kc == KCH - 1) tc_commit_pair(b_empty + (btile & 1)); }      // (the same barrier in both CTAs; the B buffer is free after the tile's last MMA)
					else { tc_commit(empty + stage);                // frees the A stage when these MMAs retire
					       if (lastb) tc_commit(b_empty + kc); }    // ... and this K-chunk of the B tile after its last use
					if (++stage == NST) { stage = 0; sphase ^= 1; }
				}
				if constexpr (PAIR) tc_commit_pair(tm_full + buf); else tc_commit(tm_full + buf);      // accumulator complete -> epilogue
				if (newb) bphase ^= 1;
				if (lastb) btile++;
				if (u + 1 < u1) w.next();
			}
			if (blockIdx.x < 160) { long long* o = g_knn_prof[blockIdx.x]; o[0] = t_tm; o[1] = t_b; o[2] = t_a; o[3] = clock64() - t_begin; o[4] = u1 - u0; o[5] = 0; }
		}
	} else if (warp == 2 + EPI_WARPS) {
		// ================================================= (pair only) B-tile producer: the next tile's half loads into the other buffer while the current one is in use
		if constexpr (PAIR) {
			if (lane == 0 && u0 < u1) {
				uint32_t btile = 0;
				UnitWalk w; w.init(pairs, n_pairs, u0);
				for (int u = u0; u < u1; u++) {
					const bool newb = (u == u0) || (w.qt == 0);
					if (newb) {
						const uint32_t bb = btile & 1, ph = (btile >> 1) & 1;
						mbar_wait(b_empty + bb, ph ^ 1);      // the tile that used this buffer two tiles ago has retired
						if (cr == 0) mbar_expect_tx(b_full + bb, 2 * SMEM_BH);      // both CTAs' halves land on the leader's barrier
						for (int kc = 0; kc < KCH; kc++)
							tma_load_2d_pair(sB + bb * SMEM_BH + kc * SBH_BYTES, &tmap_a, b_full + bb, kc * BK, w.pr.b_row0 + w.tt * BN + (int)cr * (BN / 2));
						btile++;
					}
					if (u + 1 < u1) w.next();
				}
			}
		}
	} else if (u0 < u1) {
		// ================================================= epilogue: 16 warps, lane quarter = warp % 4, column quarter = (warp-2) / 4
		const int ew = warp - 2;
		const int quarter = warp & 3;            // TMEM lanes [32*quarter, 32*quarter+32) are the only ones this warp may read
		const int cq = ew >> 2;                  // columns [64*cq, 64*cq+64) of the 256-column tile
		const int rq = lane >> 2, m = lane & 3;  // this thread: tile rows 32*quarter + rq + 8k (k = 0..3), column pairs 64*cq + 8j + 2m (+1), j = 0..7
		const int etid = threadIdx.x - 64;       // 0..511
		uint32_t tcount = 0;
		UnitWalk w; w.init(pairs, n_pairs, u0);
		int cur_p = -1;
		float s = 1.f;
		for (int u = u0; u < u1; u++, tcount++) {
			const bool newb = (u == u0) || (w.qt == 0);
			if (w.p != cur_p) { cur_p = w.p; s = pconst[cur_p].s; }
			const int qrow0 = w.qt * UM + (int)cr * BM;      // first A row (within the pair's set) of this CTA's 128-row tile
			const float* nAp = norms + w.pr.a_row0 + qrow0 + quarter * 32 + rq;      // (used after the accumulator wait: the loads have long landed)
			const float nA0 = __ldg(nAp) * s, nA1 = __ldg(nAp + 8) * s, nA2 = __ldg(nAp + 16) * s, nA3 = __ldg(nAp + 24) * s;
			if (newb) {      // |b~|^2 of the resident B tile (scaled) -> smem, one float per thread of the first 8 epilogue warps
				asm volatile("bar.sync 1, 512;" ::: "memory");       // every epilogue thread is done with the previous tile's norms
				if (etid < BN) sNorm[etid] = __ldg(norms + w.pr.b_row0 + w.tt * BN + etid) * s;
				asm volatile("bar.sync 1, 512;" ::: "memory");
			}
			unsigned long long m2s2;
			{ const float m2s = -2.0f * s; asm("mov.b64 %0, {%1, %1};" : "=l"(m2s2) : "f"(m2s)); }
			const uint32_t buf = tcount & 1, tphase = (tcount >> 1) & 1;
			unsigned long long nb[8];      // (|b~|^2 of columns 8j + 2m, 8j + 2m + 1), j = 0..7
			{
				const uint32_t nrm_s = smem_u32(sNorm + cq * 64 + 2 * m);
#pragma unroll
				for (int j = 0; j < 8; j++) asm volatile("ld.shared.b64 %0, [%1];" : "=l"(nb[j]) : "r"(nrm_s + 32u * j));
			}
			const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * BN + cq * 64;
			// A->B matrix [B group][A position]: this thread's four rows are positions 4 rq .. 4 rq + 3 of their block of 32; its groups are 2m, 2m + 1 of each block of 32 columns
			__half* rowp = G + w.pr.g_row + (size_t)(w.tt * (BN / GRP) + cq * 16 + 2 * m) * (size_t)w.pr.nA_pad + (size_t)(qrow0 + quarter * 32 + 4 * rq);
			// B->A matrix [A group][B position]: row group rq of this quarter; its eight columns of a block of 32 are positions 8m .. 8m + 7
			__half* colp = G + w.pr.g_col + (size_t)(qrow0 / GRP + quarter * 8 + rq) * (size_t)w.pr.nB_pad + (size_t)(w.tt * BN + cq * 64 + 8 * m);
			const size_t rstride = (size_t)w.pr.nA_pad;
			mbar_wait(tm_full + buf, tphase);
			tc_fence_after();
			// TMEM -> registers: the second half (rows +16, +24) is in flight while the first is converted
			uint32_t va[32], vb[32];
			uint32_t h0[8], h1[8], h2[8], h3[8];
			TMEM_LD_16x256b_x8(taddr0, va);
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			TMEM_LD_16x256b_x8(taddr0 + (16u << 16), vb);
			epi_rows(va, nb, nA0, nA1, m2s2, h0, h1);
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			tc_fence_before();
			__syncwarp();
			if (lane == 0) { if constexpr (PAIR) mbar_arrive_rank0(tm_empty + buf); else mbar_arrive(tm_empty + buf); }      // the accumulator is in registers: the next unit's MMAs may overwrite it
			epi_rows(vb, nb, nA2, nA3, m2s2, h2, h3);
#pragma unroll
			for (int jb = 0; jb < 2; jb++) {
				// ---- B->A: minimum over the thread's four rows, per column pair; transposed into position order and stored as 16 bytes
				uint32_t c[4];
#pragma unroll
				for (int jj = 0; jj < 4; jj++) { const int j = 4 * jb + jj; c[jj] = hmin2_u32(hmin2_u32(h0[j], h1[j]), hmin2_u32(h2[j], h3[j])); }
				*reinterpret_cast<uint4*>(colp + 32 * jb) = make_uint4(__byte_perm(c[0], c[1], 0x5410), __byte_perm(c[2], c[3], 0x5410), __byte_perm(c[0], c[1], 0x7632), __byte_perm(c[2], c[3], 0x7632));
				// ---- A->B: minimum over the four column pairs of this block of 32 columns, per row: (group 2m, group 2m + 1) x four rows
				const uint32_t r0 = hmin2_u32(hmin2_u32(h0[4 * jb], h0[4 * jb + 1]), hmin2_u32(h0[4 * jb + 2], h0[4 * jb + 3]));
				const uint32_t r1 = hmin2_u32(hmin2_u32(h1[4 * jb], h1[4 * jb + 1]), hmin2_u32(h1[4 * jb + 2], h1[4 * jb + 3]));
				const uint32_t r2 = hmin2_u32(hmin2_u32(h2[4 * jb], h2[4 * jb + 1]), hmin2_u32(h2[4 * jb + 2], h2[4 * jb + 3]));
				const uint32_t r3 = hmin2_u32(hmin2_u32(h3[4 * jb], h3[4 * jb + 1]), hmin2_u32(h3[4 * jb + 2], h3[4 * jb + 3]));
				__half* rp = rowp + (size_t)(8 * jb) * rstride;
				*reinterpret_cast<uint2*>(rp) = make_uint2(__byte_perm(r0, r1, 0x5410), __byte_perm(r2, r3, 0x5410));
				*reinterpret_cast<uint2*>(rp + rstride) = make_uint2(__byte_perm(r0, r1, 0x7632), __byte_perm(r2, r3, 0x7632));
			}
			if (u + 1 < u1) w.next();
		}
	}
	tc_fence_before();
	__syncthreads();
	if constexpr (PAIR) cluster_sync_all();      // neither CTA's shared memory / tensor memory goes away while the other still uses it
	if (warp == 1) {
		tc_fence_after();
		if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
		else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
	}
}

// ------------------------------------------------------------------------------------------------ select
struct SelJob {           // one per (pair, direction)
	long long g_off;      // this direction's matrix  [n_groups padded][q_stride]  (halves, offset into G)
	long long x_off;      // the other direction's    [queries / 4][x_stride]
	int q_stride, x_stride;
	int nq, nt, n_groups; // n_groups = ceil(nt / 4)
	int q_pool_row0, t_pool_row0;
	int pconst;           // index into the PairConst table, -1: empty pair (no tensor pass ran)
	int dir;              // 0: A->B outputs, 1: B->A outputs
	int out_off;          // rows (x k) into the idx/dist outputs of this direction
	int q_base;           // global query index of this job's row 0
	int blk0;             // first k_knn_select block of this job
	int pad0, pad1;
};
__device__ __forceinline__ int find_by_start(const int* __restrict__ start, int n, int v) {      // largest j with start[j] <= v
	int lo = 0, hi = n - 1;
	while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (start[mid] <= v) lo = mid; else hi = mid - 1; }
	return lo;
}
struct KnnOut { int32_t* idx[2]; float* dist[2]; };

// branch-free insertion into an ascending list that drops its largest entry; half2 = two queries at once (the keys are non-negative
// halves, +inf = empty)
template <int K> __device__ __forceinline__ void kmin_insert2(__half2 (&a)[K], __half2 x) {
	__half2 prev = a[0];
	a[0] = __hmin2(a[0], x);
#pragma unroll
	for (int j = 1; j < K; j++) { const __half2 cur = a[j]; a[j] = __hmin2(cur, __hmax2(prev, x)); prev = cur; }
}
__device__ __forceinline__ __half2 u32_as_h2(unsigned v) { return *reinterpret_cast<__half2*>(&v); }

// 64 queries per block: lane = two adjacent queries (the matrices are [group][query], so a warp reads 128 contiguous bytes per group
// and every min / max of the selection network serves both queries), 8 warps = 8 segments of the group axis.
template <int K> __global__ void __launch_bounds__(256) k_knn_select(const SelJob* __restrict__ jobs, const int* __restrict__ blk_start, int n_jobs,
                                                                     const __half* __restrict__ G, const float* __restrict__ errn, const PairConst* __restrict__ pconst,
                                                                     int k, int32_t* __restrict__ cand, int32_t* __restrict__ cand_cnt, int* __restrict__ qjob) {
	__shared__ unsigned s_top[8][K][32];
	__shared__ unsigned s_hmax[64];
	__shared__ int s_ng[64], s_nc[64];
	__shared__ int s_grp[64][SEL_MAXG];
	__shared__ int s_cand[64][SEL_MAXC];
	const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
	const int ji = find_by_start(blk_start, n_jobs, (int)blockIdx.x);
	const SelJob jb = jobs[ji];
	const int q0 = ((int)blockIdx.x - jb.blk0) * 64;
	const bool live = jb.pconst >= 0;
	// Everything here runs in POSITION space (see knn_unpos): the matrices are stored by position, a group is four consecutive positions;
	// real row numbers only appear where a row's data (error norm, bounds, outputs) is touched.
	const int qa = q0 + 2 * lane;                                   // my two query positions: qa, qa + 1 (the matrices are padded to whole tiles: both loads stay inside)
	const bool anyq = live && (knn_unpos(qa) < jb.nq || knn_unpos(qa + 1) < jb.nq);
	const unsigned* Gq = reinterpret_cast<const unsigned*>(G + jb.g_off + qa);
	const size_t gstride = (size_t)jb.q_stride >> 1;                // in 32-bit words
	const int per = (jb.n_groups + 7) >> 3;
	const int g0 = min(seg * per, jb.n_groups), g1 = min(g0 + per, jb.n_groups);
	if (threadIdx.x < 64) { s_ng[threadIdx.x] = 0; s_nc[threadIdx.x] = 0; }
	// ---- pass 1: the K smallest keys of my segment, for both queries
	const __half2 inf2 = u32_as_h2(0x7c007c00u);
	__half2 top[K];
#pragma unroll
	for (int j = 0; j < K; j++) top[j] = inf2;
	if (anyq) {
		int g = g0;
		for (; g + 8 <= g1; g += 8) {      // eight independent loads in flight per lane: the matrices come from L2 / DRAM and this pass is latency-bound
			unsigned kq[8];
#pragma unroll
			for (int u = 0; u < 8; u++) kq[u] = __ldg(Gq + (size_t)(g + u) * gstride);
#pragma unroll
			for (int u = 0; u < 8; u++) kmin_insert2<K>(top, u32_as_h2(kq[u]));
		}
		for (; g + 4 <= g1; g += 4) {
			const unsigned k0 = __ldg(Gq + (size_t)g * gstride), k1 = __ldg(Gq + (size_t)(g + 1) * gstride);
			const unsigned k2 = __ldg(Gq + (size_t)(g + 2) * gstride), k3 = __ldg(Gq + (size_t)(g + 3) * gstride);
			kmin_insert2<K>(top, u32_as_h2(k0)); kmin_insert2<K>(top, u32_as_h2(k1)); kmin_insert2<K>(top, u32_as_h2(k2)); kmin_insert2<K>(top, u32_as_h2(k3));
		}
		for (; g < g1; g++) kmin_insert2<K>(top, u32_as_h2(__ldg(Gq + (size_t)g * gstride)));
	}
#pragma unroll
	for (int j = 0; j < K; j++) s_top[seg][j][lane] = h2_as_u32(top[j]);
	__syncthreads();
	if (seg == 0) {      // merge: the k-th smallest key of each whole column, then its threshold
		__half2 all[K];
#pragma unroll
		for (int j = 0; j < K; j++) all[j] = inf2;
		for (int s2 = 0; s2 < 8; s2++) {
#pragma unroll
			for (int j = 0; j < K; j++) kmin_insert2<K>(all, u32_as_h2(s_top[s2][j][lane]));
		}
		const int kk = min(min(k, jb.nt), K);
		unsigned hk2 = 0x7c007c00u;
#pragma unroll
		for (int j = 0; j < K; j++) if (j == kk - 1) hk2 = h2_as_u32(all[j]);
#pragma unroll
		for (int hq = 0; hq < 2; hq++) {
			const int q = knn_unpos(qa + hq);      // the real query row
			const unsigned hk16 = hq ? (hk2 >> 16) : (hk2 & 0xffffu);
			unsigned hmax = 0u;
			if (live && q < jb.nq) {
				hmax = 0xffffu;
				if (kk > 0 && hk16 < 0x7c00u) {
					const PairConst pc = pconst[jb.pconst];
					const float hk = __half2float(__ushort_as_half((unsigned short)hk16));
					const float eq = __ldg(errn + jb.q_pool_row0 + q);
					const float e = (eq + (jb.dir == 0 ? pc.emaxB : pc.emaxA)) * 1.0001f + 1e-9f;
					// k groups have a member (their minimum) whose stored key is <= hk  =>  the exact k-th distance is <= T:
					//   stored = rn16(s * computed d2): |stored - s d2c| <= 2^-11 stored + 2^-25;  |d2c - d2(rounded rows)| <= delta;  | |a~-b~| - |a-b| | <= e
					const float D2hi = fmaf(hk, 1.0f + 0x1p-10f, 0x1p-22f) * pc.inv_s + pc.delta;
					const float T = sqrtf(D2hi) * (1.0f + 1e-6f) + e;
					// a row whose rounded-row distance exceeds R is farther than T in exact arithmetic; translate R back into a stored key
					const float R = (T + e) * (1.0f + 1e-6f);
					const float hl = fmaf(R * R * (1.0f + 2e-6f) + pc.delta, pc.s * (1.0f + 0x1p-10f), 0x1p-22f);
					if (hl < 65000.f) hmax = (unsigned)__half_as_ushort(__float2half_ru(hl));
				}
			}
			s_hmax[2 * lane + hq] = hmax;
		}
	}
	__syncthreads();
	// ---- pass 2: groups that can still hold one of the k nearest (keys order like their bit patterns)
	const unsigned hm0 = s_hmax[2 * lane], hm1 = s_hmax[2 * lane + 1];
	if (anyq) {
		for (int gb = g0; gb < g1; gb += 8) {      // (second sweep over the same column: eight loads in flight again)
			unsigned kq[8];
#pragma unroll
			for (int u = 0; u < 8; u++) kq[u] = __ldg(Gq + (size_t)min(gb + u, g1 - 1) * gstride);
#pragma unroll
			for (int u = 0; u < 8; u++) {
				const int g = gb + u;
				const unsigned key2 = kq[u];
				const bool p0 = g < g1 && (key2 & 0xffffu) <= hm0 && hm0 != 0u, p1 = g < g1 && (key2 >> 16) <= hm1 && hm1 != 0u;
				if (p0) { const int slot = atomicAdd(&s_ng[2 * lane], 1); if (slot < SEL_MAXG) s_grp[2 * lane][slot] = g; }
				if (p1) { const int slot = atomicAdd(&s_ng[2 * lane + 1], 1); if (slot < SEL_MAXG) s_grp[2 * lane + 1][slot] = g; }
			}
		}
	}
	__syncthreads();
	// ---- members: the other direction's matrix holds, for (my 4-row group, member column), a minimum that includes my own entry
	//      => a lower bound of the member's stored key; 8 bytes per group instead of four descriptor rows
	for (int w2 = threadIdx.x; w2 < 64 * SEL_MAXG; w2 += 256) {
		const int ql = w2 & 63, it = w2 >> 6, q = q0 + ql;      // q: query POSITION
		if (live && it < min(s_ng[ql], SEL_MAXG)) {      // (queries beyond nq have hmax = 0 and therefore no groups)
			const int g = s_grp[ql][it];
			const unsigned hmax = s_hmax[ql];
			const uint2 m = __ldg(reinterpret_cast<const uint2*>(G + jb.x_off + (size_t)(q >> 2) * jb.x_stride + (size_t)g * GRP));
			const unsigned key4[4] = { m.x & 0xffffu, m.x >> 16, m.y & 0xffffu, m.y >> 16 };
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const int t = knn_unpos(g * GRP + j);      // the member's real row
				if (t < jb.nt && key4[j] <= hmax) { const int c = atomicAdd(&s_nc[ql], 1); if (c < SEL_MAXC) s_cand[ql][c] = t; }
			}
		}
	}
	__syncthreads();
	for (int w2 = threadIdx.x; w2 < 64 * SEL_MAXC; w2 += 256) {
		const int ql = w2 & 63, c = w2 >> 6, q = knn_unpos(q0 + ql);      // back to the real query row for the outputs
		if (q < jb.nq) {
			const int ng = s_ng[ql], nc = s_nc[ql];
			const bool over = ng > SEL_MAXG || nc > SEL_MAXC;
			if (c == 0) { cand_cnt[jb.q_base + q] = over ? -1 : nc; qjob[jb.q_base + q] = ji; }
			if (!over && c < nc) cand[(size_t)(jb.q_base + q) * SEL_MAXC + c] = s_cand[ql][c];
		}
	}
}

// ------------------------------------------------------------------------------------------------ exact stage
__device__ __forceinline__ bool dist_less(double da, int ia, double db, int ib) {      // non-negative doubles order like their bit patterns
	const long long a = __double_as_longlong(da), b = __double_as_longlong(db);
	return a < b || (a == b && ia < ib);
}

// one warp per query row: exact float64 distance of every candidate on the fp32 copies of the rows (two rows in flight), top k by
// (distance, index).  The kernel is paced by the float -> double conversions (XU pipe) and the FP64 pipe; a variant with eight lanes per
// candidate and four candidates in flight (three shuffle steps instead of five) was measured SLOWER (0.30 vs 0.23 ms for 180 k rows): it
// converts the query's slice four times as often and idles on the last round's unused candidate slots.
__global__ void __launch_bounds__(256) k_knn_rerank(const SelJob* __restrict__ jobs, const int* __restrict__ qjob, int total_q,
                                                    const int32_t* __restrict__ cand, const int32_t* __restrict__ cand_cnt, const float* __restrict__ pool_f,
                                                    int k, KnnOut out, int* fallback_rows, int* fallback_count, int force_fallback) {
	const int lane = threadIdx.x & 31;
	const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (gw >= total_q) return;
	const SelJob jb = jobs[__ldg(qjob + gw)];      // (k_knn_select wrote the job of every query next to its candidate count: a binary search over the
	                                               //  job table per warp cost 66 instructions and 10 % of this kernel's stall samples)
	const int r = gw - jb.q_base;
	const int kk = min(k, jb.nt);
	const float INF = __int_as_float(0x7f800000);
	const int nc = (jb.pconst >= 0) ? cand_cnt[gw] : 0;
	bool ok = nc >= kk;      // fewer candidates than neighbours can only mean an overflow (-1) or a broken bound: recompute exactly
	double my_d = 1e300; int my_i = 0x7fffffff;      // lane c keeps candidate c
	if (ok && nc > 0) {
		const int myc = (lane < nc) ? cand[(size_t)gw * SEL_MAXC + lane] : 0;
		const float* qrow = pool_f + (size_t)(jb.q_pool_row0 + r) * KD;
		const float* tbase = pool_f + (size_t)jb.t_pool_row0 * KD;
		const float4 qa = __ldg(reinterpret_cast<const float4*>(qrow) + lane * 2), qb = __ldg(reinterpret_cast<const float4*>(qrow) + lane * 2 + 1);
		const double q0 = qa.x, q1 = qa.y, q2 = qa.z, q3 = qa.w, q4 = qb.x, q5 = qb.y, q6 = qb.z, q7 = qb.w;
		auto d2_exact = [&](const float4& b0, const float4& b1) -> double {
			double s = 0.0, d;
			d = q0 - (double)b0.x; s += d * d; d = q1 - (double)b0.y; s += d * d;
			d = q2 - (double)b0.z; s += d * d; d = q3 - (double)b0.w; s += d * d;
			d = q4 - (double)b1.x; s += d * d; d = q5 - (double)b1.y; s += d * d;
			d = q6 - (double)b1.z; s += d * d; d = q7 - (double)b1.w; s += d * d;
			return s;
		};
		for (int c = 0; c < nc; c += 2) {
			const int t0 = __shfl_sync(0xffffffffu, myc, c), t1 = __shfl_sync(0xffffffffu, myc, min(c + 1, nc - 1));
			const float4* r0 = reinterpret_cast<const float4*>(tbase + (size_t)t0 * KD) + lane * 2;
			const float4* r1 = reinterpret_cast<const float4*>(tbase + (size_t)t1 * KD) + lane * 2;
			const float4 x0 = __ldg(r0), x1 = __ldg(r0 + 1), y0 = __ldg(r1), y1 = __ldg(r1 + 1);
			double sa = d2_exact(x0, x1), sb = d2_exact(y0, y1);
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
			if (lane == c) { my_d = sa; my_i = t0; }
			if (lane == c + 1 && c + 1 < nc) { my_d = sb; my_i = t1; }
		}
	}
	// ---- rank by counting: lane c's output slot = number of candidates that order before it by (distance, index); the indices are
	//      distinct, so the ranks are a permutation (a sorted insertion of doubles cost five times the distance evaluation)
	int rank = 0;
	for (int j = 0; j < nc; j++) {
		const double dj = __shfl_sync(0xffffffffu, my_d, j);
		const int ij = __shfl_sync(0xffffffffu, my_i, j);
		rank += dist_less(dj, ij, my_d, my_i) ? 1 : 0;
	}
	int32_t* io = out.idx[jb.dir] + ((size_t)jb.out_off + r) * k;
	float* dd = out.dist[jb.dir] + ((size_t)jb.out_off + r) * k;
	if (ok && lane < nc && rank < kk) { io[rank] = my_i; dd[rank] = (float)sqrt(my_d); }
	if (lane < k && (!ok || lane >= kk)) { io[lane] = -1; dd[lane] = INF; }      // rows with fewer than k train rows are padded; a row that goes to the fallback is rewritten there
	if (force_fallback > 0 && gw % force_fallback == 0) ok = false;
	if (jb.nt == 0) ok = true;      // nothing to find
	if (lane == 0 && !ok) { const int slot = atomicAdd(fallback_count, 1); fallback_rows[slot] = gw; }
}

// exact brute force for the rows whose candidate set overflowed.  Few rows (the normal case: none or one): each row's train set is
// cut into FB_SEG segments, one CTA per (row, segment), the last CTA of a row (ticket) merges the FB_SEG partial lists.  Many rows:
// one CTA per row.  Every merge is a warp-parallel selection (k rounds of a shuffle arg-min), never a serial walk.
static constexpr int FB_SEG = 64, FB_SPLIT_ROWS = 256;

__device__ __forceinline__ double exact_d2(const float* __restrict__ a, const float* __restrict__ b, int lane) {
	// warp-cooperative sum over KD floats: lane handles 8 consecutive elements; fixed order => deterministic
	const float4 a0 = __ldg(reinterpret_cast<const float4*>(a) + lane * 2), a1 = __ldg(reinterpret_cast<const float4*>(a) + lane * 2 + 1);
	const float4 b0 = __ldg(reinterpret_cast<const float4*>(b) + lane * 2), b1 = __ldg(reinterpret_cast<const float4*>(b) + lane * 2 + 1);
	double s = 0.0, d;
	d = (double)a0.x - (double)b0.x; s += d * d; d = (double)a0.y - (double)b0.y; s += d * d;
	d = (double)a0.z - (double)b0.z; s += d * d; d = (double)a0.w - (double)b0.w; s += d * d;
	d = (double)a1.x - (double)b1.x; s += d * d; d = (double)a1.y - (double)b1.y; s += d * d;
	d = (double)a1.z - (double)b1.z; s += d * d; d = (double)a1.w - (double)b1.w; s += d * d;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	return s;
}
// Every lane brings two (distance, index) entries (index 0x7fffffff = empty; real indices are distinct); lane j < 8 returns the j-th
// smallest of the warp's 64 by (distance, index).
__device__ __forceinline__ void warp_select8(double d0, int i0, double d1, int i1, double& od, int& oi) {
	const int lane = threadIdx.x & 31;
	od = 1e300; oi = 0x7fffffff;
#pragma unroll 1
	for (int r = 0; r < 8; r++) {
		const bool first = dist_less(d0, i0, d1, i1);
		double wd = first ? d0 : d1; int wi = first ? i0 : i1;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			const double xd = __shfl_xor_sync(0xffffffffu, wd, o); const int xi = __shfl_xor_sync(0xffffffffu, wi, o);
			if (dist_less(xd, xi, wd, wi)) { wd = xd; wi = xi; }
		}
		if (lane == r) { od = wd; oi = wi; }
		if (wi == 0x7fffffff) break;                     // nothing left (uniform: every lane holds the same winner)
		if (i0 == wi) { d0 = 1e300; i0 = 0x7fffffff; } else if (i1 == wi) { d1 = 1e300; i1 = 0x7fffffff; }
	}
}

__global__ void __launch_bounds__(256) k_knn_exact(const SelJob* __restrict__ jobs, const int* __restrict__ q_start, int n_jobs, const float* __restrict__ pool_f,
                                                    const int* __restrict__ fallback_rows, const int* __restrict__ fallback_count, int k, KnnOut out,
                                                    double* part_d, int* part_i, int* tickets) {
	__shared__ double s_d[8][8];
	__shared__ int s_i[8][8];
	__shared__ int s_last;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int n = *fallback_count;
	const bool split = n <= FB_SPLIT_ROWS;
	const int S = split ? FB_SEG : 1;
	for (int w = blockIdx.x; w < n * S; w += gridDim.x) {
		const int f = w / S, seg = w - f * S;
		const int gw = fallback_rows[f];
		const int ji = find_by_start(q_start, n_jobs, gw);
		const SelJob jb = jobs[ji];
		const int r = gw - jb.q_base;
		const float* qrow = pool_f + (size_t)(jb.q_pool_row0 + r) * KD;
		const float* tbase = pool_f + (size_t)jb.t_pool_row0 * KD;
		const int len = (jb.nt + S - 1) / S, t0 = seg * len, t1 = min(jb.nt, t0 + len);
		// ---- every warp: sorted top-8 of its share of the segment (two rows in flight)
		double best_d[8]; int best_i[8];
#pragma unroll
		for (int j = 0; j < 8; j++) { best_d[j] = 1e300; best_i[j] = 0x7fffffff; }
		auto insert = [&](double cd, int ci) {
#pragma unroll
			for (int j = 0; j < 8; j++) {
				if (dist_less(cd, ci, best_d[j], best_i[j])) { const double td = best_d[j]; const int tix = best_i[j]; best_d[j] = cd; best_i[j] = ci; cd = td; ci = tix; }
			}
		};
		for (int ti = t0 + warp; ti < t1; ti += 16) {
			const int tj = ti + 8;
			const double da = exact_d2(qrow, tbase + (size_t)ti * KD, lane);
			const double db = exact_d2(qrow, tbase + (size_t)min(tj, t1 - 1) * KD, lane);
			insert(da, ti);
			if (tj < t1) insert(db, tj);
		}
		__syncthreads();      // (the previous round's readers of s_d are done)
		if (lane < 8) {
			double md = 1e300; int mi = 0x7fffffff;
#pragma unroll
			for (int j = 0; j < 8; j++) if (lane == j) { md = best_d[j]; mi = best_i[j]; }
			s_d[warp][lane] = md; s_i[warp][lane] = mi;
		}
		__syncthreads();
		int32_t* io = out.idx[jb.dir] + ((size_t)jb.out_off + r) * k;
		float* dd = out.dist[jb.dir] + ((size_t)jb.out_off + r) * k;
		if (warp == 0) {      // the CTA's 8 best of its 64 entries
			double od; int oi;
			warp_select8((&s_d[0][0])[2 * lane], (&s_i[0][0])[2 * lane], (&s_d[0][0])[2 * lane + 1], (&s_i[0][0])[2 * lane + 1], od, oi);
			if (lane < 8) {
				if (!split) { if (lane < k) { io[lane] = oi == 0x7fffffff ? -1 : oi; dd[lane] = oi == 0x7fffffff ? __int_as_float(0x7f800000) : (float)sqrt(od); } }
				else { part_d[((size_t)f * FB_SEG + seg) * 8 + lane] = od; part_i[((size_t)f * FB_SEG + seg) * 8 + lane] = oi; }
			}
			if (lane == 0) s_last = 0;
			__syncwarp();
			if (split) { __threadfence(); if (lane == 0) s_last = (atomicAdd(tickets + f, 1) == FB_SEG - 1); }
		}
		__syncthreads();
		if (split && s_last) {     // FB_SEG sorted lists (512 entries, two per thread) -> per-warp 8 best -> the row's k best
			__threadfence();
			const size_t base = (size_t)f * FB_SEG * 8 + 2 * threadIdx.x;
			double od; int oi;
			warp_select8(__ldcg(part_d + base), __ldcg(part_i + base), __ldcg(part_d + base + 1), __ldcg(part_i + base + 1), od, oi);
			__syncthreads();
			if (lane < 8) { s_d[warp][lane] = od; s_i[warp][lane] = oi; }
			__syncthreads();
			if (warp == 0) {
				warp_select8((&s_d[0][0])[2 * lane], (&s_i[0][0])[2 * lane], (&s_d[0][0])[2 * lane + 1], (&s_i[0][0])[2 * lane + 1], od, oi);
				if (lane < k) { io[lane] = oi == 0x7fffffff ? -1 : oi; dd[lane] = oi == 0x7fffffff ? __int_as_float(0x7f800000) : (float)sqrt(od); }
				if (lane == 0) tickets[f] = 0;
			}
		}
		__syncthreads();
	}
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct MatcherState {
	int max_pairs = 0, max_feats = 0, dim = 0;
	int slot_rows = 0;                      // rows per pool slot (max_feats padded to 256)
	int n_transient = 0, n_persist = 0;     // transient slots [0, n_transient) serve bt_knn_match_pairs; persistent ones follow
	std::vector<int> slot_n;                // persistent slots: number of rows stored, -1 = empty
	DevBuf pool_h, pool_f, norms, errn, slot_meta, sets, pairs, pconst, G, jobs, blk_start, q_start, cand, cand_cnt, qjob, fb_rows, fb_count, fb_part_d, fb_part_i, fb_tickets;
	PinnedBuf h_stage;
	cudaEvent_t ev_up = nullptr;            // behind the upload of h_stage: the next call waits on it before rewriting the pinned block
	PFN_encodeTiled encode = nullptr;
	CUtensorMap tmap_a, tmap_b;
	bool maps_ready = false;
	size_t g_halves_cap = 0;
	int max_jobs = 0, max_q_total = 0, max_blocks = 0;
	int last_units = 0, last_rows = 0;
	cudaEvent_t ev[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
	bool timing = false;
	bool cta_pairs = false;     // BT_KNN_CTA_PAIRS=1: tensor pass on clusters of two CTAs (cta_group::2, M = 256).  Measured on B200: no faster than one CTA per tile (0.083-0.088 vs 0.081 ms for 45 pairs x 2000^2) - see the kernel's header - so it is not the default
	int force_fallback = 0;     // test knob: every n-th query row is sent to the exact fallback regardless of its candidates
};

void matcher_destroy(bt_ctx* ctx) {
	MatcherState* m = ctx->matcher;
	if (!m) return;
	DevBuf* bufs[] = { &m->pool_h, &m->pool_f, &m->norms, &m->errn, &m->slot_meta, &m->sets, &m->pairs, &m->pconst, &m->G, &m->jobs, &m->blk_start, &m->q_start, &m->cand, &m->cand_cnt, &m->qjob,
	                   &m->fb_rows, &m->fb_count, &m->fb_part_d, &m->fb_part_i, &m->fb_tickets };
	for (DevBuf* b : bufs) b->release();
	m->h_stage.release();
	if (m->ev_up) cudaEventDestroy(m->ev_up);
	for (auto& e : m->ev) if (e) cudaEventDestroy(e);
	delete m;
	ctx->matcher = nullptr;
}

static int make_tmap(MatcherState* m, CUtensorMap* out, void* base, uint64_t rows, uint32_t box_rows) {
	const cuuint64_t gdim[2] = { (cuuint64_t)KD, rows };
	const cuuint64_t gstride[1] = { (cuuint64_t)KD * 2 };
	const cuuint32_t box[2] = { (cuuint32_t)BK, box_rows };
	const cuuint32_t estr[2] = { 1, 1 };
	const CUresult r = m->encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return BT_ERR_CUDA; }
	return BT_OK;
}

// (re)allocates the descriptor pool for n_transient + n_persist slots and rebuilds the tensor maps over it
static int pool_alloc(MatcherState* m) {
	const size_t rows = (size_t)(m->n_transient + m->n_persist) * m->slot_rows + BN;
	int rc;
	if ((rc = m->pool_h.alloc(rows * KD * 2)) != BT_OK) return rc;
	if ((rc = m->pool_f.alloc(rows * KD * 4)) != BT_OK) return rc;
	if ((rc = m->norms.alloc(rows * 4)) != BT_OK) return rc;
	if ((rc = m->errn.alloc(rows * 4)) != BT_OK) return rc;
	if ((rc = m->slot_meta.alloc(sizeof(int) * 2 * (size_t)(m->n_transient + m->n_persist))) != BT_OK) return rc;
	if ((rc = make_tmap(m, &m->tmap_a, m->pool_h.p, (uint64_t)rows, BM)) != BT_OK) return rc;
	if ((rc = make_tmap(m, &m->tmap_b, m->pool_h.p, (uint64_t)rows, BN)) != BT_OK) return rc;
	m->maps_ready = true;
	return BT_OK;
}

struct PoolSetRef { int slot; int n; };      // a descriptor set in the pool: rows [slot * slot_rows, +n)

static int launch_prep(bt_ctx* ctx, const std::vector<PrepSet>& sets, char* hb, size_t& hoff, cudaStream_t stream) {
	MatcherState* m = ctx->matcher;
	if (sets.empty()) return BT_OK;
	const size_t b_sets = sizeof(PrepSet) * sets.size();
	memcpy(hb + hoff, sets.data(), b_sets);
	BT_CUDA(cudaMemcpyAsync(m->sets.p, hb + hoff, b_sets, cudaMemcpyHostToDevice, stream));
	hoff += (b_sets + 255) & ~(size_t)255;
	int max_padded = 0;
	for (auto& s : sets) { max_padded = std::max(max_padded, s.rows_padded); BT_CUDA(cudaMemsetAsync(m->slot_meta.as<int>() + 2 * s.slot, 0, 2 * sizeof(int), stream)); }
	k_desc_prep<<<dim3((unsigned)std::max(1, std::min((max_padded + 7) / 8, 64)), (unsigned)sets.size()), 256, 0, stream>>>(
	    m->sets.as<PrepSet>(), m->pool_h.as<__half>(), m->pool_f.as<float>(), m->norms.as<float>(), m->errn.as<float>(), m->slot_meta.as<int>());
	return BT_OK;
}

// the matcher proper on descriptor sets that are already in the pool
static int knn_run(bt_ctx* ctx, int n_pairs, const PoolSetRef* A, const PoolSetRef* B, int k, int32_t* idxAB, float* distAB, int32_t* idxBA, float* distBA,
                   char* hb, size_t hoff, cudaStream_t stream) {
	MatcherState* m = ctx->matcher;
	const int UM = m->cta_pairs ? 2 * BM : BM;
	std::vector<KnnPair> pairs;
	std::vector<SelJob> jobs;
	std::vector<int> blk_start, q_start;
	long long units = 0, g_off = 0;
	int q_total = 0, blocks = 0;
	size_t off_dir[2] = { 0, 0 };
	for (int p = 0; p < n_pairs; p++) {
		const int nA = A[p].n, nB = B[p].n;
		const bool live = nA > 0 && nB > 0;
		KnnPair pr; memset(&pr, 0, sizeof pr);
		if (live) {
			pr.a_row0 = A[p].slot * m->slot_rows; pr.b_row0 = B[p].slot * m->slot_rows; pr.nA = nA; pr.nB = nB;
			pr.n_qt = (nA + UM - 1) / UM; pr.n_tt = (nB + BN - 1) / BN; pr.unit0 = (int)units;      // a unit = UM A rows x 256 B rows (UM = 256 with CTA pairs)
			pr.nA_pad = pr.n_qt * UM; pr.nB_pad = pr.n_tt * BN; pr.setA = A[p].slot; pr.setB = B[p].slot;
			pr.g_row = g_off; g_off += (long long)(pr.nB_pad / GRP) * pr.nA_pad;
			pr.g_col = g_off; g_off += (long long)(pr.nA_pad / GRP) * pr.nB_pad;
			units += (long long)pr.n_qt * pr.n_tt;
			pairs.push_back(pr);
		}
		for (int dir = 0; dir < 2; dir++) {
			SelJob jb; memset(&jb, 0, sizeof jb);
			jb.nq = dir == 0 ? nA : nB; jb.nt = dir == 0 ? nB : nA; jb.n_groups = (jb.nt + 31) / 32 * (32 / GRP);      // whole blocks of 32 positions: a block's valid rows are spread over all of its groups
			jb.q_pool_row0 = (dir == 0 ? A[p].slot : B[p].slot) * m->slot_rows; jb.t_pool_row0 = (dir == 0 ? B[p].slot : A[p].slot) * m->slot_rows;
			jb.pconst = live ? (int)pairs.size() - 1 : -1;
			if (live) {
				jb.g_off = dir == 0 ? pr.g_row : pr.g_col; jb.x_off = dir == 0 ? pr.g_col : pr.g_row;
				jb.q_stride = dir == 0 ? pr.nA_pad : pr.nB_pad; jb.x_stride = dir == 0 ? pr.nB_pad : pr.nA_pad;
			}
			jb.dir = dir; jb.out_off = (int)off_dir[dir]; jb.q_base = q_total; jb.blk0 = blocks;
			jobs.push_back(jb);
			blk_start.push_back(blocks); q_start.push_back(q_total);
			blocks += (jb.nq + 63) / 64; q_total += jb.nq; off_dir[dir] += (size_t)jb.nq;
		}
	}
	BT_REQUIRE((size_t)g_off <= m->g_halves_cap && q_total <= m->max_q_total && blocks <= m->max_blocks && units < (1ll << 31), BT_ERR_CAPACITY, "bt_knn_match_pairs: work list overflow");
	{ KnnPair guard; memset(&guard, 0, sizeof guard); guard.unit0 = (int)units; guard.n_qt = guard.n_tt = 1; pairs.push_back(guard); }   // UnitWalk::next() may read one entry past the last pair
	const int n_live = (int)pairs.size() - 1;
	// ---- upload tables (one pinned block; the caller holds the event guard)
	const size_t b_pairs = sizeof(KnnPair) * pairs.size(), b_jobs = sizeof(SelJob) * jobs.size(), b_int = sizeof(int) * jobs.size();
	char* h0 = hb + hoff;
	memcpy(h0, pairs.data(), b_pairs);
	memcpy(h0 + b_pairs, jobs.data(), b_jobs);
	memcpy(h0 + b_pairs + b_jobs, blk_start.data(), b_int);
	memcpy(h0 + b_pairs + b_jobs + b_int, q_start.data(), b_int);
	BT_CUDA(cudaMemcpyAsync(m->pairs.p, h0, b_pairs, cudaMemcpyHostToDevice, stream));
	BT_CUDA(cudaMemcpyAsync(m->jobs.p, h0 + b_pairs, b_jobs, cudaMemcpyHostToDevice, stream));
	BT_CUDA(cudaMemcpyAsync(m->blk_start.p, h0 + b_pairs + b_jobs, b_int, cudaMemcpyHostToDevice, stream));
	BT_CUDA(cudaMemcpyAsync(m->q_start.p, h0 + b_pairs + b_jobs + b_int, b_int, cudaMemcpyHostToDevice, stream));
	BT_CUDA(cudaEventRecord(m->ev_up, stream));
	BT_CUDA(cudaMemsetAsync(m->fb_count.p, 0, 16, stream));
	if (n_live > 0) k_knn_pairconst<<<(n_live + 127) / 128, 128, 0, stream>>>(m->pairs.as<KnnPair>(), n_live, m->slot_meta.as<int>(), m->pconst.as<PairConst>());
	if (m->timing) BT_CUDA(cudaEventRecord(m->ev[1], stream));      // [ev0, ev1) = descriptor conversion + table upload + per-pair constants; [ev1, ev2) = k_knn_tc alone
	// ---- kernels
	if (n_live > 0) {
		if (m->cta_pairs) {      // clusters of two CTAs (one TPC each): cta_group::2 MMAs, M = 256
			cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
			const int n_cl = (int)std::min<long long>(units, ctx->sm_count / 2);
			cfg.gridDim = dim3(2 * n_cl); cfg.blockDim = dim3(KNN_THREADS_P); cfg.dynamicSmemBytes = SMEM_TOTAL_P; cfg.stream = stream;
			cudaLaunchAttribute at[1];
			at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
			cfg.attrs = at; cfg.numAttrs = 1;
			BT_CUDA(cudaLaunchKernelEx(&cfg, k_knn_tc<true>, m->tmap_a, m->tmap_b, (const KnnPair*)m->pairs.as<KnnPair>(), n_live, (int)units, (const float*)m->norms.as<float>(),
			                           (const PairConst*)m->pconst.as<PairConst>(), m->G.as<__half>()));
		} else {
			const int grid = (int)std::min<long long>(units, ctx->sm_count);
			k_knn_tc<false><<<grid, KNN_THREADS, SMEM_TOTAL, stream>>>(m->tmap_a, m->tmap_b, m->pairs.as<KnnPair>(), n_live, (int)units, m->norms.as<float>(), m->pconst.as<PairConst>(), m->G.as<__half>());
		}
	}
	if (m->timing) BT_CUDA(cudaEventRecord(m->ev[2], stream));
	KnnOut out; out.idx[0] = idxAB; out.idx[1] = idxBA; out.dist[0] = distAB; out.dist[1] = distBA;
	if (q_total > 0) {
		if (k <= 5) k_knn_select<5><<<blocks, 256, 0, stream>>>(m->jobs.as<SelJob>(), m->blk_start.as<int>(), (int)jobs.size(), m->G.as<__half>(), m->errn.as<float>(), m->pconst.as<PairConst>(), k, m->cand.as<int32_t>(), m->cand_cnt.as<int32_t>(), m->qjob.as<int>());
		else k_knn_select<8><<<blocks, 256, 0, stream>>>(m->jobs.as<SelJob>(), m->blk_start.as<int>(), (int)jobs.size(), m->G.as<__half>(), m->errn.as<float>(), m->pconst.as<PairConst>(), k, m->cand.as<int32_t>(), m->cand_cnt.as<int32_t>(), m->qjob.as<int>());
		k_knn_rerank<<<(q_total + 7) / 8, 256, 0, stream>>>(m->jobs.as<SelJob>(), m->qjob.as<int>(), q_total, m->cand.as<int32_t>(), m->cand_cnt.as<int32_t>(), m->pool_f.as<float>(),
		                                                 k, out, m->fb_rows.as<int>(), m->fb_count.as<int>(), m->force_fallback);
		if (m->timing) BT_CUDA(cudaEventRecord(m->ev[3], stream));
		k_knn_exact<<<ctx->sm_count * 2, 256, 0, stream>>>(m->jobs.as<SelJob>(), m->q_start.as<int>(), (int)jobs.size(), m->pool_f.as<float>(), m->fb_rows.as<int>(), m->fb_count.as<int>(), k, out,
		                                              m->fb_part_d.as<double>(), m->fb_part_i.as<int>(), m->fb_tickets.as<int>());
	} else if (m->timing) BT_CUDA(cudaEventRecord(m->ev[3], stream));
	if (m->timing) BT_CUDA(cudaEventRecord(m->ev[4], stream));
	BT_CUDA(cudaGetLastError());
	m->last_units = (int)units; m->last_rows = q_total;
	return BT_OK;
}

}  // namespace bt

using namespace bt;

extern "C" int bt_matcher_reserve(bt_ctx* ctx, int max_pairs, int max_feats, int dim) {
	BT_REQUIRE(ctx, BT_ERR_INVALID_ARG, "bt_matcher_reserve: NULL ctx");
	BT_REQUIRE(dim == KD, BT_ERR_UNSUPPORTED, "bt_matcher_reserve: descriptor dim %d unsupported (the tensor path is built for %d)", dim, KD);
	BT_REQUIRE(max_pairs > 0 && max_feats > 0, BT_ERR_INVALID_ARG, "bt_matcher_reserve: bad limits");
	BT_CUDA(cudaSetDevice(ctx->device));
	if (!ctx->matcher) ctx->matcher = new MatcherState();
	MatcherState* m = ctx->matcher;
	if (!m->encode) {
		void* fn = nullptr;
		cudaDriverEntryPointQueryResult qres;
		BT_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
		BT_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, BT_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
		m->encode = (PFN_encodeTiled)fn;
	}
	const int slot_rows = (max_feats + BN - 1) / BN * BN;
	if (slot_rows != m->slot_rows || std::min(2 * max_pairs, 128) != m->n_transient) std::fill(m->slot_n.begin(), m->slot_n.end(), -1);     // the persistent slots move: what they held is dropped
	m->max_pairs = max_pairs; m->max_feats = max_feats; m->dim = dim; m->slot_rows = slot_rows;
	m->n_transient = std::min(2 * max_pairs, 128);       // distinct descriptor sets one bt_knn_match_pairs call may name (all pairs of K frames: K sets)
	int rc;
	if ((rc = pool_alloc(m)) != BT_OK) return rc;
	{ const char* e = getenv("BT_KNN_CTA_PAIRS"); m->cta_pairs = (e && atoi(e) != 0); }
	const int pad_a = (max_feats + 2 * BM - 1) / (2 * BM) * (2 * BM);      // A sets are padded to whole 256-row units of the CTA-pair kernel
	m->g_halves_cap = (size_t)max_pairs * ((size_t)(slot_rows / GRP) * pad_a + (size_t)(pad_a / GRP) * slot_rows);
	m->max_jobs = 2 * max_pairs;
	m->max_q_total = 2 * max_pairs * max_feats;
	m->max_blocks = 2 * max_pairs * ((max_feats + 63) / 64);
#define RES(buf, bytes) if ((rc = m->buf.alloc(bytes)) != BT_OK) return rc
	RES(sets, sizeof(PrepSet) * (size_t)std::max(m->n_transient, 1));
	RES(pairs, sizeof(KnnPair) * ((size_t)max_pairs + 1));
	RES(pconst, sizeof(PairConst) * (size_t)max_pairs);
	RES(G, m->g_halves_cap * 2 + 64);
	RES(jobs, sizeof(SelJob) * m->max_jobs);
	RES(blk_start, sizeof(int) * m->max_jobs); RES(q_start, sizeof(int) * m->max_jobs);
	RES(cand, sizeof(int32_t) * SEL_MAXC * (size_t)m->max_q_total);
	RES(cand_cnt, sizeof(int32_t) * (size_t)m->max_q_total);
	RES(qjob, sizeof(int) * (size_t)m->max_q_total);
	RES(fb_rows, sizeof(int) * (size_t)m->max_q_total);
	RES(fb_count, 16);
	RES(fb_part_d, sizeof(double) * (size_t)FB_SPLIT_ROWS * FB_SEG * 8); RES(fb_part_i, sizeof(int) * (size_t)FB_SPLIT_ROWS * FB_SEG * 8);
	RES(fb_tickets, sizeof(int) * (size_t)FB_SPLIT_ROWS);
	BT_CUDA(cudaMemset(m->fb_tickets.p, 0, sizeof(int) * (size_t)FB_SPLIT_ROWS));
#undef RES
	if ((rc = m->h_stage.alloc(sizeof(PrepSet) * 2 * max_pairs + sizeof(KnnPair) * ((size_t)max_pairs + 1) + (sizeof(SelJob) + 2 * sizeof(int)) * m->max_jobs + 4096)) != BT_OK) return rc;
	if (!m->ev_up) BT_CUDA(cudaEventCreateWithFlags(&m->ev_up, cudaEventDisableTiming));
	BT_CUDA(cudaFuncSetAttribute(k_knn_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL));
	BT_CUDA(cudaFuncSetAttribute(k_knn_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL_P));
	return BT_OK;
}

// Persistent descriptor pool: Lfnet::detectFeature uploads a frame's descriptors once (FeatureManager.cpp:907);
// bt_desc_pool_store converts them once into a slot, bt_knn_match_slots / bt_match_pairs_pool then name slots.
extern "C" int bt_desc_pool_reserve(bt_ctx* ctx, int n_slots) {
	BT_REQUIRE(ctx && ctx->matcher && ctx->matcher->maps_ready, BT_ERR_INVALID_ARG, "bt_desc_pool_reserve: call bt_matcher_reserve first");
	BT_REQUIRE(n_slots > 0, BT_ERR_INVALID_ARG, "bt_desc_pool_reserve: bad slot count");
	MatcherState* m = ctx->matcher;
	BT_CUDA(cudaSetDevice(ctx->device));
	if (n_slots <= m->n_persist) return BT_OK;
	BT_CUDA(cudaDeviceSynchronize());          // the pool moves: nothing may still be reading the old one
	m->n_persist = n_slots;
	m->slot_n.assign(n_slots, -1);             // (growing the pool re-allocates it: stored sets are dropped)
	return pool_alloc(m);
}

extern "C" int bt_desc_pool_store(bt_ctx* ctx, int slot, const bt_desc_view* desc, void* stream_) {
	BT_REQUIRE(ctx && ctx->matcher && ctx->matcher->maps_ready && desc, BT_ERR_INVALID_ARG, "bt_desc_pool_store: call bt_matcher_reserve / bt_desc_pool_reserve first");
	MatcherState* m = ctx->matcher;
	BT_REQUIRE(slot >= 0 && slot < m->n_persist, BT_ERR_INVALID_ARG, "bt_desc_pool_store: slot %d outside [0,%d)", slot, m->n_persist);
	BT_REQUIRE(desc->dim == KD, BT_ERR_UNSUPPORTED, "bt_desc_pool_store: descriptor dim %d != %d", desc->dim, KD);
	BT_REQUIRE(desc->n >= 0 && desc->n <= m->max_feats, BT_ERR_CAPACITY, "bt_desc_pool_store: %d features > reserved %d", desc->n, m->max_feats);
	BT_REQUIRE(desc->n == 0 || desc->dev, BT_ERR_INVALID_ARG, "bt_desc_pool_store: NULL descriptor pointer");
	BT_REQUIRE(((uintptr_t)desc->dev & 15) == 0 && ((desc->pitch_bytes ? desc->pitch_bytes : (size_t)KD * 4) & 15) == 0, BT_ERR_INVALID_ARG, "bt_desc_pool_store: descriptors must be 16-byte aligned");
	cudaStream_t stream = (cudaStream_t)stream_;
	BT_CUDA(cudaSetDevice(ctx->device));
	BT_CUDA(cudaEventSynchronize(m->ev_up));
	PrepSet s; s.src = desc->dev; s.pitch_bytes = desc->pitch_bytes ? desc->pitch_bytes : (size_t)KD * 4; s.n = desc->n;
	s.slot = m->n_transient + slot; s.row0 = s.slot * m->slot_rows; s.rows_padded = (std::max(desc->n, 1) + BN - 1) / BN * BN;
	size_t hoff = 0;
	int rc = launch_prep(ctx, std::vector<PrepSet>(1, s), m->h_stage.as<char>(), hoff, stream);
	if (rc != BT_OK) return rc;
	BT_CUDA(cudaEventRecord(m->ev_up, stream));
	BT_CUDA(cudaGetLastError());
	m->slot_n[slot] = desc->n;
	return BT_OK;
}

extern "C" int bt_knn_match_slots(bt_ctx* ctx, int n_pairs, const int32_t* slotA, const int32_t* slotB, int k,
                                  int32_t* idxAB, float* distAB, int32_t* idxBA, float* distBA, void* stream_) {
	BT_REQUIRE(ctx && ctx->matcher && ctx->matcher->maps_ready, BT_ERR_INVALID_ARG, "bt_knn_match_slots: call bt_matcher_reserve first");
	BT_REQUIRE(slotA && slotB && idxAB && distAB && idxBA && distBA, BT_ERR_INVALID_ARG, "bt_knn_match_slots: NULL argument");
	BT_REQUIRE(k >= 1 && k <= 8, BT_ERR_UNSUPPORTED, "bt_knn_match_slots: k=%d outside [1,8]", k);
	MatcherState* m = ctx->matcher;
	BT_CUDA(cudaSetDevice(ctx->device));
	BT_REQUIRE(n_pairs > 0 && n_pairs <= m->max_pairs, BT_ERR_CAPACITY, "bt_knn_match_slots: %d pairs > reserved %d", n_pairs, m->max_pairs);
	std::vector<PoolSetRef> A(n_pairs), B(n_pairs);
	for (int p = 0; p < n_pairs; p++) {
		for (int s2 = 0; s2 < 2; s2++) {
			const int sl = s2 == 0 ? slotA[p] : slotB[p];
			BT_REQUIRE(sl >= 0 && sl < m->n_persist && m->slot_n[sl] >= 0, BT_ERR_INVALID_ARG, "bt_knn_match_slots: pair %d names slot %d, which is empty or out of range", p, sl);
			(s2 == 0 ? A[p] : B[p]) = PoolSetRef{ m->n_transient + sl, m->slot_n[sl] };
		}
	}
	BT_CUDA(cudaEventSynchronize(m->ev_up));
	cudaStream_t stream = (cudaStream_t)stream_;
	if (m->timing) BT_CUDA(cudaEventRecord(m->ev[0], stream));
	return knn_run(ctx, n_pairs, A.data(), B.data(), k, idxAB, distAB, idxBA, distBA, m->h_stage.as<char>(), 0, stream);
}

extern "C" BT_API int bt_knn_debug_prof(long long* out960) {      // developer aid, not part of the public header
	return cudaMemcpyFromSymbol(out960, g_knn_prof, sizeof(long long) * 160 * 6) == cudaSuccess ? BT_OK : BT_ERR_CUDA;
}
extern "C" int bt_knn_debug_force_fallback(bt_ctx* ctx, int every_nth) {
	BT_REQUIRE(ctx && ctx->matcher, BT_ERR_INVALID_ARG, "bt_knn_debug_force_fallback: call bt_matcher_reserve first");
	ctx->matcher->force_fallback = every_nth > 0 ? every_nth : 0;
	return BT_OK;
}

extern "C" int bt_knn_enable_timing(bt_ctx* ctx, int on) {
	BT_REQUIRE(ctx && ctx->matcher, BT_ERR_INVALID_ARG, "bt_knn_enable_timing: call bt_matcher_reserve first");
	MatcherState* m = ctx->matcher;
	if (on) for (auto& e : m->ev) if (!e) BT_CUDA(cudaEventCreate(&e));
	m->timing = on != 0;
	return BT_OK;
}
// ms4 = {descriptor prep, tensor-core pass, select + exact re-rank, exact fallback}; info3 = {units, query rows, fallback rows}
extern "C" int bt_knn_get_timing(bt_ctx* ctx, float* ms4, int* info3) {
	BT_REQUIRE(ctx && ctx->matcher && ctx->matcher->timing && ms4 && info3, BT_ERR_INVALID_ARG, "bt_knn_get_timing: timing not enabled");
	MatcherState* m = ctx->matcher;
	BT_CUDA(cudaEventSynchronize(m->ev[4]));
	for (int k = 0; k < 4; k++) BT_CUDA(cudaEventElapsedTime(ms4 + k, m->ev[k], m->ev[k + 1]));
	int fb = 0;
	BT_CUDA(cudaMemcpy(&fb, m->fb_count.p, sizeof(int), cudaMemcpyDeviceToHost));
	info3[0] = m->last_units; info3[1] = m->last_rows; info3[2] = fb;
	return BT_OK;
}

extern "C" int bt_knn_match_pairs(bt_ctx* ctx, int n_pairs, const bt_desc_view* A, const bt_desc_view* B, int k,
                                  int32_t* idxAB, float* distAB, int32_t* idxBA, float* distBA, void* stream_) {
	BT_REQUIRE(ctx && ctx->matcher && ctx->matcher->maps_ready, BT_ERR_INVALID_ARG, "bt_knn_match_pairs: call bt_matcher_reserve first");
	BT_REQUIRE(A && B && idxAB && distAB && idxBA && distBA, BT_ERR_INVALID_ARG, "bt_knn_match_pairs: NULL argument");
	BT_REQUIRE(k >= 1 && k <= 8, BT_ERR_UNSUPPORTED, "bt_knn_match_pairs: k=%d outside [1,8]", k);
	MatcherState* m = ctx->matcher;
	cudaStream_t stream = (cudaStream_t)stream_;
	BT_CUDA(cudaSetDevice(ctx->device));
	BT_REQUIRE(n_pairs > 0 && n_pairs <= m->max_pairs, BT_ERR_CAPACITY, "bt_knn_match_pairs: %d pairs > reserved %d", n_pairs, m->max_pairs);
	for (int p = 0; p < n_pairs; p++) {
		for (const bt_desc_view* v : { &A[p], &B[p] }) {
			BT_REQUIRE(v->dim == KD, BT_ERR_UNSUPPORTED, "pair %d: descriptor dim %d != %d", p, v->dim, KD);
			BT_REQUIRE(v->n >= 0 && v->n <= m->max_feats, BT_ERR_CAPACITY, "pair %d: %d features > reserved %d", p, v->n, m->max_feats);
			BT_REQUIRE(v->n == 0 || v->dev, BT_ERR_INVALID_ARG, "pair %d: NULL descriptor pointer", p);
			BT_REQUIRE(((uintptr_t)v->dev & 15) == 0 && ((v->pitch_bytes ? v->pitch_bytes : (size_t)KD * 4) & 15) == 0, BT_ERR_INVALID_ARG, "pair %d: descriptors must be 16-byte aligned", p);
		}
	}
	// ---- unique descriptor sets (a keyframe appears in many pairs) -> transient pool slots, converted once per call
	std::map<std::pair<const float*, int>, int> slot_of;
	std::vector<PrepSet> sets;
	auto slot = [&](const bt_desc_view& v) -> int {
		const auto key = std::make_pair(v.dev, v.n);
		auto it = slot_of.find(key);
		if (it != slot_of.end()) return it->second;
		PrepSet s; s.src = v.dev; s.pitch_bytes = v.pitch_bytes ? v.pitch_bytes : (size_t)v.dim * 4; s.n = v.n; s.slot = (int)sets.size();
		s.row0 = s.slot * m->slot_rows; s.rows_padded = (std::max(v.n, 1) + BN - 1) / BN * BN;
		sets.push_back(s);
		slot_of[key] = s.slot;
		return s.slot;
	};
	std::vector<PoolSetRef> sa(n_pairs), sb(n_pairs);
	for (int p = 0; p < n_pairs; p++) { sa[p] = PoolSetRef{ slot(A[p]), A[p].n }; sb[p] = PoolSetRef{ slot(B[p]), B[p].n }; }
	BT_REQUIRE((int)sets.size() <= m->n_transient, BT_ERR_CAPACITY, "bt_knn_match_pairs: %d distinct descriptor sets in one call > %d (store them with bt_desc_pool_store and use bt_knn_match_slots)", (int)sets.size(), m->n_transient);
	// the entry points are asynchronous: an earlier call's table upload may still be reading the pinned block
	BT_CUDA(cudaEventSynchronize(m->ev_up));
	if (m->timing) BT_CUDA(cudaEventRecord(m->ev[0], stream));
	size_t hoff = 0;
	int rc = launch_prep(ctx, sets, m->h_stage.as<char>(), hoff, stream);
	if (rc != BT_OK) return rc;
	return knn_run(ctx, n_pairs, sa.data(), sb.data(), k, idxAB, distAB, idxBA, distBA, m->h_stage.as<char>(), hoff, stream);
}
