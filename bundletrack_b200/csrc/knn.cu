// knn.cu — all-pairs L2 k-nearest-neighbour matching of 256-d LF-Net descriptors on the 5th-gen tensor cores.
//
// Replaces the two cv::cuda::DescriptorMatcher::knnMatch(…, k=5) calls of SiftManager::findCorresbyNN
// (/root/reference/src/FeatureManager.cpp:271-273) for a BATCH of frame pairs, both directions.  OpenCV's CUDA matcher
// materialises the nQ x nT fp32 distance matrix with a SIMT kernel and runs k row-min passes over it;
// here the distance matrix is never written:
//
//   k_desc_prep   fp32 (pitched GpuMat rows) -> bf16 pool [rows padded to 256][D] + fp32 |x~|^2 of the ROUNDED rows
//                 (padding rows are zero with norm = +inf so they can never be selected).
//   k_knn_tc      persistent, warp-specialised tcgen05 kernel.  Work item = (pair, direction, 128-row query tile, train
//                 split).  Warp 0 streams operands with TMA (SWIZZLE_128B, K-major); warp 1 issues
//                 tcgen05.mma.cta_group::1.kind::f16 (bf16 x bf16 -> fp32, M=128, N=256, K=16 per instruction) into a
//                 double-buffered TMEM accumulator (2 x 256 columns); eight epilogue warps read the accumulator with
//                 tcgen05.ld (thread t of a lane quarter owns query row t), form key = |t~|^2 - 2 q~.t~, reduce every
//                 group of 4 train columns to its minimum and push (min key | group index in the low 11 mantissa bits)
//                 through a branch-free min/max insertion network that keeps the 12 smallest group minima per thread:
//                 no data-dependent branch anywhere in the epilogue.
//   k_knn_rerank  per query row, three filters with rigorous intervals: groups whose minimum could still reach the top k
//                 (measured bf16 rounding error of both rows) -> their 4 members re-scored on the bf16 pool rows
//                 (coalesced, fp32) -> the survivors re-scored EXACTLY (float64 sum of (a-b)^2 on the fp32 inputs); top k
//                 by (distance, index).  The same bound proves that no train row outside the kept groups can enter the
//                 top k; rows where the proof fails go to
//   k_knn_exact   exact brute force for those rows only (rare; train set split over 16 CTAs per row + ticketed merge).
//   (Sixteen epilogue warps with four shorter lists per row were measured: 5 % faster tensor pass, 17x more fallback rows.)
// Result == exact brute-force kNN (float64 distances, ties -> lower train index), distances returned as
// float(sqrt(d2)) like cv::NORM_L2.
#include <algorithm>
#include <cuda.h>
#include <cuda_bf16.h>
#include <map>
#include <math.h>
#include "bt_common.cuh"

namespace bt {

static constexpr int KD = 256;              // descriptor dimension handled by the tensor path
static constexpr int BM = 128;              // query rows per tile (UMMA M)
static constexpr int BN = 256;              // train rows per tile (UMMA N)
static constexpr int BK = 64;               // K elements per smem chunk: 64 bf16 = 128 B = one SWIZZLE_128B atom row
static constexpr int KCH = KD / BK;         // 4 chunks
static constexpr int STAGES = 4;            // train-operand pipeline depth (32 KB each)
static constexpr int GRP = 4;                // train columns per group: only each group's minimum key enters the list
static constexpr int KC = 12;               // group minima kept per epilogue thread (per column half)
static constexpr int NCAND = 2 * KC;        // list entries per (query row, split)
static constexpr int IDX_BITS = 11;         // group index (relative to the split) packed in the low mantissa bits
static constexpr int MAX_SPLIT_ROWS = GRP << IDX_BITS;   // 8192 train rows per split
static constexpr int KNN_THREADS = 32 * 10; // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
static constexpr uint32_t SQ_BYTES = BM * BK * 2;     // 16 KB per K-chunk of the query tile
static constexpr uint32_t SB_BYTES = BN * BK * 2;     // 32 KB per stage
static constexpr uint32_t SMEM_Q = KCH * SQ_BYTES;    // 64 KB
static constexpr uint32_t SMEM_B = STAGES * SB_BYTES; // 128 KB
static constexpr uint32_t SMEM_NORM = 2 * BN * 4;     // 2 KB
static constexpr uint32_t SMEM_BAR = 256;
static constexpr uint32_t SMEM_TOTAL = SMEM_Q + SMEM_B + SMEM_NORM + SMEM_BAR + 1024;   // + alignment slack

struct KnnItem {
	int q_row0;      // pool row of the first query row of this tile
	int q_valid;     // valid query rows in this tile (<= 128)
	int t_row0;      // pool row of the first train row of this split
	int t_tiles;     // number of 256-row train tiles in this split
	int cand_off;    // index (in rows) into the candidate buffer for this tile's first row
	int pad0, pad1, pad2;
};

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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
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
// UMMA instruction descriptor (kind::f16): c_format F32 (1) @4, a_format BF16 (1) @7, b_format BF16 (1) @10,
// a/b K-major (0) @15/@16, N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
	return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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
struct PrepSet { const float* src; size_t pitch_bytes; int n; int row0; int rows_padded; int pad; };

__global__ void __launch_bounds__(256) k_desc_prep(const PrepSet* sets, __nv_bfloat16* pool, float* norms, float* errn, int* set_maxnorm2, int* set_maxerr) {
	const PrepSet st = sets[blockIdx.y];
	const int warps_per_block = blockDim.x >> 5, lane = threadIdx.x & 31;
	for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < st.rows_padded; r += gridDim.x * warps_per_block) {
		__nv_bfloat16* dst = pool + (size_t)(st.row0 + r) * KD;
		float ss = 0.f, ee = 0.f;
		if (r < st.n) {
			const float* src = (const float*)((const char*)st.src + (size_t)r * st.pitch_bytes);
			// each lane converts 8 consecutive floats (two float4 loads, one 16-byte store)
			const float4 a = __ldg(reinterpret_cast<const float4*>(src) + lane * 2), b = __ldg(reinterpret_cast<const float4*>(src) + lane * 2 + 1);
			const float v[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
			__nv_bfloat16 h[8];
#pragma unroll
			for (int k = 0; k < 8; k++) { h[k] = __float2bfloat16_rn(v[k]); const float f = __bfloat162float(h[k]); ss += f * f; const float dlt = v[k] - f; ee += dlt * dlt; }
			*reinterpret_cast<uint4*>(dst + lane * 8) = *reinterpret_cast<const uint4*>(h);
		} else {
			*reinterpret_cast<uint4*>(dst + lane * 8) = make_uint4(0u, 0u, 0u, 0u);
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, o); ee += __shfl_xor_sync(0xffffffffu, ee, o); }
		if (lane == 0) {
			norms[st.row0 + r] = (r < st.n) ? ss : __int_as_float(0x7f800000);
			const float en = sqrtf(ee) * 1.001f + 1e-7f;     // |x~ - x|, rounded up: the ACTUAL bf16 rounding error of this row
			errn[st.row0 + r] = (r < st.n) ? en : 0.f;
			if (r < st.n) { atomicMax(set_maxnorm2 + blockIdx.y, __float_as_int(ss)); atomicMax(set_maxerr + blockIdx.y, __float_as_int(en)); }   // non-negative floats order like ints
		}
	}
}

// ------------------------------------------------------------------------------------------------ k_knn_tc
// Branch-free insertion of x into the ascending list a[0..KC-1], dropping the largest.
__device__ __forceinline__ void topk_insert(float (&a)[KC], float x) {
	float prev = a[0];
	a[0] = fminf(a[0], x);
#pragma unroll
	for (int j = 1; j < KC; j++) {
		const float cur = a[j];
		a[j] = fminf(cur, fmaxf(prev, x));
		prev = cur;
	}
}

__global__ void __launch_bounds__(KNN_THREADS, 1)
k_knn_tc(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_t, const KnnItem* __restrict__ items, int n_items,
         const float* __restrict__ norms, float* __restrict__ cand) {
	extern __shared__ uint8_t smem_raw[];
	uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1024-B alignment
	uint8_t* sQ = smem;
	uint8_t* sB = smem + SMEM_Q;
	float* sNorm = reinterpret_cast<float*>(smem + SMEM_Q + SMEM_B);
	uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_Q + SMEM_B + SMEM_NORM);
	uint64_t* q_full = bars + 0;
	uint64_t* q_empty = bars + 1;
	uint64_t* full = bars + 2;                 // [STAGES]
	uint64_t* empty = bars + 2 + STAGES;       // [STAGES]
	uint64_t* tm_full = bars + 2 + 2 * STAGES; // [2]
	uint64_t* tm_empty = tm_full + 2;          // [2]
	uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tm_empty + 2);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (warp == 0 && lane == 0) {
		asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
		asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_t) : "memory");
		mbar_init(q_full, 1); mbar_init(q_empty, 1);
		for (int s = 0; s < STAGES; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
		for (int b = 0; b < 2; b++) { mbar_init(tm_full + b, 1); mbar_init(tm_empty + b, 8); }
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 1) {   // TMEM: all 512 columns (two 256-column accumulators); this kernel runs 1 CTA / SM
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_ptr_smem;

	if (warp == 0) {
		// ================================================= TMA producer
		if (lane == 0) {
			uint32_t stage = 0, sphase = 0, qphase = 0;
			for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
				const KnnItem item = items[it];
				mbar_wait(q_empty, qphase ^ 1);        // previous item's MMAs no longer read the query tile
				mbar_expect_tx(q_full, SMEM_Q);
				for (int kc = 0; kc < KCH; kc++) tma_load_2d(sQ + kc * SQ_BYTES, &tmap_q, q_full, kc * BK, item.q_row0);
				qphase ^= 1;
				for (int t = 0; t < item.t_tiles; t++) {
					for (int kc = 0; kc < KCH; kc++) {
						mbar_wait(empty + stage, sphase ^ 1);
						mbar_expect_tx(full + stage, SB_BYTES);
						tma_load_2d(sB + stage * SB_BYTES, &tmap_t, full + stage, kc * BK, item.t_row0 + t * BN);
						if (++stage == STAGES) { stage = 0; sphase ^= 1; }
					}
				}
			}
		}
	} else if (warp == 1) {
		// ================================================= MMA issuer (one thread)
		if (lane == 0) {
			constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
			uint32_t stage = 0, sphase = 0, qphase = 0, tcount = 0;
			for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
				const KnnItem item = items[it];
				mbar_wait(q_full, qphase);
				qphase ^= 1;
				for (int t = 0; t < item.t_tiles; t++, tcount++) {
					const uint32_t buf = tcount & 1, tphase = (tcount >> 1) & 1;
					mbar_wait(tm_empty + buf, tphase ^ 1);      // epilogue drained this accumulator
					tc_fence_after();
					const uint32_t d_tmem = tmem_base + buf * BN;
					for (int kc = 0; kc < KCH; kc++) {
						mbar_wait(full + stage, sphase);
						tc_fence_after();
						const uint32_t a_base = smem_u32(sQ + kc * SQ_BYTES), b_base = smem_u32(sB + stage * SB_BYTES);
#pragma unroll
						for (int k = 0; k < BK / 16; k++) {      // UMMA K = 16 bf16 = 32 bytes inside the 128-byte swizzle row
							tc_mma_bf16(d_tmem, umma_desc_k128(a_base + k * 32), umma_desc_k128(b_base + k * 32), idesc, (uint32_t)((kc | k) != 0));
						}
						tc_commit(empty + stage);                // frees the stage when these MMAs retire
						if (++stage == STAGES) { stage = 0; sphase ^= 1; }
					}
					tc_commit(tm_full + buf);                    // accumulator complete -> epilogue
				}
				tc_commit(q_empty);                              // query tile may be overwritten
			}
		}
	} else {
		// ================================================= epilogue: 8 warps, lane quarter = warp % 4, column half = (warp-2) / 4
		const int ew = warp - 2;
		const int quarter = warp & 3;            // TMEM lanes [32*quarter, 32*quarter+32) are the only ones this warp may read
		const int half = ew >> 2;                // columns [128*half, 128*half+128) of each 256-column tile
		const int row = quarter * 32 + lane;     // query row inside the tile == TMEM lane
		const int etid = threadIdx.x - 64;       // 0..255
		uint32_t tcount = 0;
		for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
			const KnnItem item = items[it];
			float best[KC];
#pragma unroll
			for (int j = 0; j < KC; j++) best[j] = __int_as_float(0x7f800000);
			for (int t = 0; t < item.t_tiles; t++, tcount++) {
				const uint32_t buf = tcount & 1, tphase = (tcount >> 1) & 1;
				// |t~|^2 of this train tile -> smem (one float per epilogue thread).  The previous user of sNorm[buf] (two
				// tiles ago) finished before its tm_empty arrival, which the MMA warp waited for before this tile's MMAs;
				// the named barrier below orders this store against the reads.
				sNorm[buf * BN + etid] = __ldg(norms + item.t_row0 + t * BN + etid);
				asm volatile("bar.sync 1, 256;" ::: "memory");
				mbar_wait(tm_full + buf, tphase);
				tc_fence_after();
				const float* nrm = sNorm + buf * BN + half * 128;
				const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * BN + half * 128;
				const int col_base = t * BN + half * 128;   // train index relative to the split
				// TMEM -> registers, double-buffered: chunk c+1 is in flight while chunk c is scanned.  Within a 32-column
				// chunk every key is tested against the threshold the chunk STARTED with (no loop-carried dependency, the
				// compares schedule ahead of the rare inserts); inserting against a stale threshold is harmless because the
				// min/max network leaves the list untouched when the key is not better than its current 8th entry.
				auto scan = [&](const uint32_t (&v)[32], int c) {
#pragma unroll
					for (int g = 0; g < 32 / GRP; g++) {
						float m = fmaf(-2.0f, __uint_as_float(v[g * GRP]), nrm[c * 32 + g * GRP]);
#pragma unroll
						for (int j = 1; j < GRP; j++) m = fminf(m, fmaf(-2.0f, __uint_as_float(v[g * GRP + j]), nrm[c * 32 + g * GRP + j]));
						m = fminf(m, 3.0e38f);    // an all-padding group has key +inf: OR-ing index bits into +inf would make a NaN and corrupt the min/max network
						const uint32_t packed = (__float_as_uint(m) & ~((1u << IDX_BITS) - 1u)) | (uint32_t)((col_base + c * 32) / GRP + g);
						topk_insert(best, __uint_as_float(packed));     // branch-free: a key worse than the 12th entry falls out again
					}
				};
				uint32_t va[32], vb[32];
				TMEM_LD32(taddr0, va);
				asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
				TMEM_LD32(taddr0 + 32, vb);
				scan(va, 0);
				asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
				TMEM_LD32(taddr0 + 64, va);
				scan(vb, 1);
				asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
				TMEM_LD32(taddr0 + 96, vb);
				scan(va, 2);
				asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
				scan(vb, 3);
				tc_fence_before();
				__syncwarp();
				if (lane == 0) mbar_arrive(tm_empty + buf);
			}
			if (row < item.q_valid) {
				float4* out = reinterpret_cast<float4*>(cand + ((size_t)(item.cand_off + row) * NCAND + half * KC));
#pragma unroll
				for (int j = 0; j < KC / 4; j++) out[j] = make_float4(best[4 * j], best[4 * j + 1], best[4 * j + 2], best[4 * j + 3]);
			}
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 1) {
		tc_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
	}
}

// ------------------------------------------------------------------------------------------------ exact re-rank
struct RerankJob {        // one per (pair, direction)
	const float* q; size_t q_pitch; int nq; int q_pool_row0;
	const float* t; size_t t_pitch; int nt; int t_set;
	int n_splits; int split_rows;   // train rows per split (multiple of 256)
	int cand_off;                   // candidate rows of (split s, query row r) start at cand_off + s*cand_split_stride + r
	int cand_split_stride;
	int out_off;                    // rows (x k) into the idx/dist outputs of this direction
	int dir;                        // 0: A->B outputs, 1: B->A outputs
	int t_pool_row0;                // first row of the train set in the bf16 pool / norms / errn
};

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
__device__ __forceinline__ int find_job(const int* __restrict__ job_row_start, int n_jobs, int gw) {
	int lo = 0, hi = n_jobs - 1;
	while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (job_row_start[mid] <= gw) lo = mid; else hi = mid - 1; }
	return lo;
}
struct KnnOut { int32_t* idx[2]; float* dist[2]; };

// lane-private exact squared distance between two fp32 rows (float64 accumulation, fixed order)
__device__ __forceinline__ double exact_d2_lane(const float* __restrict__ a, const float* __restrict__ b) {
	double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
	for (int k = 0; k < KD / 4; k++) {
		const float4 x = __ldg(reinterpret_cast<const float4*>(a) + k), y = __ldg(reinterpret_cast<const float4*>(b) + k);
		double d;
		d = (double)x.x - (double)y.x; s0 += d * d; d = (double)x.y - (double)y.y; s1 += d * d;
		d = (double)x.z - (double)y.z; s0 += d * d; d = (double)x.w - (double)y.w; s1 += d * d;
	}
	return s0 + s1;
}

static constexpr int RR_MAXG = 24;            // groups expanded per row in shared memory; more => exact fallback
static constexpr int RR_MAXC = RR_MAXG * GRP; // = 96 candidate train rows

// one warp per query row.  Three filters of increasing cost, each with a rigorous interval so that no true neighbour is lost:
//   (1) group keys from k_knn_tc (min over GRP train rows, index packed into the mantissa)   -> groups that can reach the k-th
//   (2) the GRP members of those groups re-evaluated on the bf16 pool rows (coalesced 512 B per row, fp32 accumulation;
//       |d~ - d| <= errn[q] + errn[t], both measured by k_desc_prep)                           -> rows that can reach the k-th
//   (3) exact float64 distance on the caller's fp32 rows for the survivors (typically k + 1..3 rows)
// Traffic per query row: ~8 groups x 2 KB + ~7 x 1 KB instead of ~32 uncoalesced fp32 rows.
__global__ void __launch_bounds__(256, 3) k_knn_rerank(const RerankJob* __restrict__ jobs, const int* __restrict__ job_row_start, int n_jobs, int total_rows,
                                                     const float* __restrict__ cand, const __nv_bfloat16* __restrict__ pool, const float* __restrict__ norms,
                                                     const float* __restrict__ errn, const int* __restrict__ set_maxnorm2, const int* __restrict__ set_maxerr,
                                                     int k, KnnOut out, int* fallback_rows, int* fallback_count, int force_fallback) {
	__shared__ int s_g[8][RR_MAXG];
	__shared__ float s_lo[8][RR_MAXC];
	__shared__ float s_hi[8][RR_MAXC];
	__shared__ int s_sv[8][RR_MAXC];
	const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
	const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (gw >= total_rows) return;
	const int ji = find_job(job_row_start, n_jobs, gw);
	const RerankJob jb = jobs[ji];
	const int r = gw - job_row_start[ji];
	const float* qrow = (const float*)((const char*)jb.q + (size_t)r * jb.q_pitch);
	const int kk = min(k, jb.nt);
	const float INF = __int_as_float(0x7f800000);
	const int E = jb.n_splits * NCAND;
	// the query's bf16 row (8 consecutive elements per lane), issued early
	const uint4 qraw = __ldg(reinterpret_cast<const uint4*>(pool + (size_t)(jb.q_pool_row0 + r) * KD) + lane);
	// bound on | |q~ - t~| - |q - t| |: measured rounding-error norms of both rows (train side: per-set maximum) plus slack for
	// the fp32 accumulation in the tensor core; the index packing costs <= 2^-IDX_BITS relative on the key itself.
	const float qn2 = __ldg(norms + jb.q_pool_row0 + r);
	const float qerr = __ldg(errn + jb.q_pool_row0 + r);
	// All interval bounds are evaluated in fp32 and widened so that they stay conservative: the float64 square-root and compare
	// sequences they replaced were a third of this kernel's instructions; float64 is kept for the exact distances only.
	// eps is rounded up by 1e-5 relative, the radicand gets 4e-6 of absolute slack for its three roundings, sqrtf is correctly
	// rounded and the factor (1 +- 1e-6) covers it and the final add.
	const float qn = sqrtf(qn2), tn = sqrtf(__int_as_float(set_maxnorm2[jb.t_set]));
	const float eps = (qerr + __int_as_float(set_maxerr[jb.t_set]) + 1e-4f * (qn + tn) + 1e-6f) * 1.00001f;
	const float pk_rel = 1.0f / (float)(1 << IDX_BITS);
	auto d_lo = [&](float key) -> float { const float v = qn2 + key - fabsf(key) * pk_rel - 4e-6f; return sqrtf(fmaxf(v, 0.f)) * (1.0f - 1e-6f) - eps; };
	auto d_hi = [&](float key) -> float { const float v = qn2 + key + fabsf(key) * pk_rel + 4e-6f; return sqrtf(fmaxf(v, 0.f)) * (1.0f + 1e-6f) + eps; };
	// ---- pass 1 over the list entries: k-th smallest key (d_hi is monotone in the key, and every entry IS a real train row, so
	//      the exact k-th distance is <= D5 = d_hi(k-th key)); tau = smallest "list is full" threshold (rows of groups that
	//      never made a list have key >= tau)
	float up[8];
#pragma unroll
	for (int j = 0; j < 8; j++) up[j] = INF;
	float tau = INF, upk = INF;                     // upk mirrors up[kk-1] (no dynamic register indexing)
	for (int e0 = 0; e0 < E; e0 += 32) {
		const int e = e0 + lane;
		float v = INF;
		if (e < E) { const int s = e / NCAND, c = e - s * NCAND; v = __ldg(cand + ((size_t)(jb.cand_off + s * jb.cand_split_stride + r) * NCAND + c)); if ((c % KC) == KC - 1) tau = fminf(tau, v); }
		for (int round = 0; round < kk; round++) {       // merge the 32 lane values into the warp-uniform sorted list up[0..kk)
			float m = v;
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
			if (!(m < upk)) break;
			float cdv = m;
#pragma unroll
			for (int j = 0; j < 8; j++) { if (j < kk && cdv < up[j]) { const float t2 = up[j]; up[j] = cdv; cdv = t2; } if (j == kk - 1) upk = up[j]; }
			const unsigned who = __ballot_sync(0xffffffffu, v == m);
			if (lane == (int)(__ffs(who) - 1)) v = INF;   // retire ONE lane holding it
		}
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) tau = fminf(tau, __shfl_xor_sync(0xffffffffu, tau, o));
	const float key_k = upk;
	const float D5 = (kk > 0 && key_k < INF) ? d_hi(key_k) : INF;
	// ---- pass 2: every group whose lower bound can still reach D5
	int ng = 0;
	for (int e0 = 0; e0 < E; e0 += 32) {
		const int e = e0 + lane;
		bool take = false; int gbase = 0;
		if (e < E) {
			const int s = e / NCAND, c = e - s * NCAND;
			const float pk = __ldg(cand + ((size_t)(jb.cand_off + s * jb.cand_split_stride + r) * NCAND + c));
			if (pk < INF) { gbase = s * jb.split_rows + (int)(__float_as_uint(pk) & ((1u << IDX_BITS) - 1u)) * GRP; take = (gbase < jb.nt) && (d_lo(pk) <= D5); }
		}
		const unsigned bal = __ballot_sync(0xffffffffu, take);
		const int pos = ng + __popc(bal & ((1u << lane) - 1u));
		if (take && pos < RR_MAXG) s_g[wl][pos] = gbase;
		ng += __popc(bal);
	}
	const bool overflow = ng > RR_MAXG;
	__syncwarp();
	bool ok = !overflow;
	double best_d[8]; int best_i[8];
#pragma unroll
	for (int j = 0; j < 8; j++) { best_d[j] = 1e300; best_i[j] = 0x7fffffff; }
	if (!overflow) {
		// ---- filter 2: the members of the taken groups on the bf16 pool rows, two groups (8 rows, 4 KB) per step
		float qf[8];
		{
			const uint32_t w[4] = { qraw.x, qraw.y, qraw.z, qraw.w };
#pragma unroll
			for (int i = 0; i < 4; i++) { qf[2 * i] = __uint_as_float(w[i] << 16); qf[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
		}
		const __nv_bfloat16* tpool = pool + (size_t)jb.t_pool_row0 * KD;
		for (int g0 = 0; g0 < ng; g0 += 2) {
			const int gA = s_g[wl][g0], gB = (g0 + 1 < ng) ? s_g[wl][g0 + 1] : gA;
			const uint4* pa = reinterpret_cast<const uint4*>(tpool + (size_t)gA * KD) + lane;     // consecutive rows are 32 uint4 apart
			const uint4* pb = reinterpret_cast<const uint4*>(tpool + (size_t)gB * KD) + lane;
			uint4 w[8];
#pragma unroll
			for (int j = 0; j < 4; j++) { w[j] = __ldg(pa + j * 32); w[4 + j] = __ldg(pb + j * 32); }
			const int trow = ((lane & 4) ? gB : gA) + (lane & 3);                             // the row lane (L & 7) will own after the reduction
			const float tn2 = __ldg(norms + jb.t_pool_row0 + trow), terr = __ldg(errn + jb.t_pool_row0 + trow);
			float v[8];
#pragma unroll
			for (int j = 0; j < 8; j++) {
				const uint32_t x[4] = { w[j].x, w[j].y, w[j].z, w[j].w };
				float a0 = 0.f, a1 = 0.f;
#pragma unroll
				for (int i = 0; i < 4; i++) { a0 = fmaf(qf[2 * i], __uint_as_float(x[i] << 16), a0); a1 = fmaf(qf[2 * i + 1], __uint_as_float(x[i] & 0xffff0000u), a1); }
				v[j] = a0 + a1;
			}
			// transposing butterfly over lane bits 2,1,0, then plain sums over bits 3,4: lane L ends with the total of value (L & 7)
#pragma unroll
			for (int half = 4; half >= 1; half >>= 1) {
				const bool upper = (lane & half) != 0;
#pragma unroll
				for (int j = 0; j < half; j++) {
					const float send = upper ? v[j] : v[j + half];
					const float keep = upper ? v[j + half] : v[j];
					v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
				}
			}
			v[0] += __shfl_xor_sync(0xffffffffu, v[0], 8);
			v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
			const int gi = g0 + ((lane >> 2) & 1);
			if (lane < 8 && gi < ng) {
				float lo = INF, hi = INF;
				if (trow < jb.nt) {
					// fp32 evaluation of |q~ - t~|^2 from 256 exact products: absolute error <= 4e-5 (|q~|^2 + |t~|^2) (sum of 256 terms,
					// the two norms, the final combination); sqrtf is correctly rounded, the factors (1 +- 1e-6) cover it and the adds
					const float d2 = qn2 + tn2 - 2.0f * v[0];
					const float slack = 4.0e-5f * (qn2 + tn2) + 1.0e-12f;
					const float e = (qerr + terr) * 1.0001f + 1.0e-9f;
					hi = sqrtf(fmaxf(d2 + slack, 0.f)) * (1.0f + 1.0e-6f) + e;
					lo = sqrtf(fmaxf(d2 - slack, 0.f)) * (1.0f - 1.0e-6f) - e;
				}
				s_lo[wl][gi * GRP + (lane & 3)] = lo;
				s_hi[wl][gi * GRP + (lane & 3)] = hi;
			}
		}
		__syncwarp();
		const int nc = ng * GRP;
		// k-th smallest upper bound over the members (warp-uniform): kk rounds of warp-min, retiring ONE holder per round
		float h[3];
#pragma unroll
		for (int j = 0; j < 3; j++) { const int c = lane + 32 * j; h[j] = (c < nc) ? s_hi[wl][c] : INF; }
		float hk = INF;
		for (int round = 0; round < kk; round++) {
			const float mine = fminf(h[0], fminf(h[1], h[2]));
			float m = mine;
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
			hk = m;
			if (!(m < INF)) break;
			const unsigned who = __ballot_sync(0xffffffffu, mine == m);
			if (lane == (int)(__ffs(who) - 1)) { if (h[0] == m) h[0] = INF; else if (h[1] == m) h[1] = INF; else h[2] = INF; }
		}
		// survivors: members whose lower bound can still reach the k-th upper bound
		int nsv = 0;
#pragma unroll
		for (int j = 0; j < 3; j++) {
			const int c = lane + 32 * j;
			const bool keep = (c < nc) && (s_lo[wl][c] <= hk) && (s_lo[wl][c] < INF);
			const unsigned bal = __ballot_sync(0xffffffffu, keep);
			if (keep) s_sv[wl][nsv + __popc(bal & ((1u << lane) - 1u))] = s_g[wl][c / GRP] + (c % GRP);
			nsv += __popc(bal);
		}
		__syncwarp();
		// ---- filter 3: exact float64 distance (warp-cooperative, two rows in flight); sorted insert by (distance, index)
		const float4 qa = __ldg(reinterpret_cast<const float4*>(qrow) + lane * 2), qb = __ldg(reinterpret_cast<const float4*>(qrow) + lane * 2 + 1);
		auto d2_exact = [&](const float4& b0, const float4& b1) -> double {
			double s = 0.0, d;
			d = (double)qa.x - (double)b0.x; s += d * d; d = (double)qa.y - (double)b0.y; s += d * d;
			d = (double)qa.z - (double)b0.z; s += d * d; d = (double)qa.w - (double)b0.w; s += d * d;
			d = (double)qb.x - (double)b1.x; s += d * d; d = (double)qb.y - (double)b1.y; s += d * d;
			d = (double)qb.z - (double)b1.z; s += d * d; d = (double)qb.w - (double)b1.w; s += d * d;
			return s;
		};
		auto insert = [&](double cd, int ci) {      // non-negative doubles order like their bit patterns: integer compares
#pragma unroll
			for (int j = 0; j < 8; j++) {
				const long long a64 = __double_as_longlong(cd), b64 = __double_as_longlong(best_d[j]);
				const bool lt = (a64 < b64) || (a64 == b64 && ci < best_i[j]);
				if (lt) { const double td = best_d[j]; const int tix = best_i[j]; best_d[j] = cd; best_i[j] = ci; cd = td; ci = tix; }
			}
		};
		for (int c = 0; c < nsv; c += 2) {
			const int t0 = s_sv[wl][c], t1 = (c + 1 < nsv) ? s_sv[wl][c + 1] : t0;
			const float4* r0 = reinterpret_cast<const float4*>((const char*)jb.t + (size_t)t0 * jb.t_pitch) + lane * 2;
			const float4* r1 = reinterpret_cast<const float4*>((const char*)jb.t + (size_t)t1 * jb.t_pitch) + lane * 2;
			const float4 x0 = __ldg(r0), x1 = __ldg(r0 + 1), y0 = __ldg(r1), y1 = __ldg(r1 + 1);
			double sa = d2_exact(x0, x1), sb = d2_exact(y0, y1);
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
			insert(sa, t0);
			if (c + 1 < nsv) insert(sb, t1);
		}
	}
	// ---- one FP64 square root per output slot, lane j takes slot j (a serial loop on lane 0 would run k square-root sequences back to back)
	double mine_d = 1e300; int mine_i = 0x7fffffff;
#pragma unroll
	for (int j = 0; j < 8; j++) if (lane == j) { mine_d = best_d[j]; mine_i = best_i[j]; }
	const bool have = lane < kk && mine_i != 0x7fffffff;
	const float dist = have ? (float)sqrt(mine_d) : INF;
	if (!overflow) {
		// ---- proof for everything that never made a list: its key is >= tau
		const float dk = __shfl_sync(0xffffffffu, dist, max(kk - 1, 0));
		const int ik = __shfl_sync(0xffffffffu, mine_i, max(kk - 1, 0));
		if (kk > 0 && tau < INF) ok = (ik != 0x7fffffff) && (dk * (1.0f + 1e-6f) < d_lo(tau));
		if (kk > 0 && ik == 0x7fffffff) ok = false;      // fewer than k survivors can only mean a broken bound: recompute exactly
	}
	if (lane < k) {
		out.idx[jb.dir][((size_t)jb.out_off + r) * k + lane] = have ? mine_i : -1;
		out.dist[jb.dir][((size_t)jb.out_off + r) * k + lane] = dist;
	}
	if (force_fallback > 0 && gw % force_fallback == 0) ok = false;
	if (lane == 0 && !ok) { const int slot = atomicAdd(fallback_count, 1); fallback_rows[slot] = gw; }
}

// exact brute force for the rows the proof rejected.  Few rows (the normal case: ~1 per thousand): each row's train set is cut
// into FB_SEG segments, one CTA per (row, segment), the last CTA of a row (ticket) merges the FB_SEG partial lists - a lone
// CTA walking 2000 train rows took 0.17 ms, longer than the tensor pass of the whole batch.  Many rows: one CTA per row.
static constexpr int FB_SEG = 16, FB_SPLIT_ROWS = 1024;

__device__ __forceinline__ bool dist_less(double da, int ia, double db, int ib) {      // non-negative doubles order like their bit patterns
	const long long a = __double_as_longlong(da), b = __double_as_longlong(db);
	return a < b || (a == b && ia < ib);
}

__global__ void __launch_bounds__(256) k_knn_exact(const RerankJob* __restrict__ jobs, const int* __restrict__ job_row_start, int n_jobs,
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
		const int ji = find_job(job_row_start, n_jobs, gw);
		const RerankJob jb = jobs[ji];
		const int r = gw - job_row_start[ji];
		const float* qrow = (const float*)((const char*)jb.q + (size_t)r * jb.q_pitch);
		const int len = (jb.nt + S - 1) / S, t0 = seg * len, t1 = min(jb.nt, t0 + len);
		double best_d[8]; int best_i[8];
#pragma unroll
		for (int j = 0; j < 8; j++) { best_d[j] = 1e300; best_i[j] = 0x7fffffff; }
		for (int ti = t0 + warp; ti < t1; ti += 8) {
			const float* trow = (const float*)((const char*)jb.t + (size_t)ti * jb.t_pitch);
			double cd = exact_d2(qrow, trow, lane); int ci = ti;
#pragma unroll
			for (int j = 0; j < 8; j++) {
				if (dist_less(cd, ci, best_d[j], best_i[j])) { const double td = best_d[j]; const int tix = best_i[j]; best_d[j] = cd; best_i[j] = ci; cd = td; ci = tix; }
			}
		}
		__syncthreads();
		if (lane == 0) for (int j = 0; j < 8; j++) { s_d[warp][j] = best_d[j]; s_i[warp][j] = best_i[j]; }
		__syncthreads();
		int32_t* io = out.idx[jb.dir] + ((size_t)jb.out_off + r) * k;
		float* dd = out.dist[jb.dir] + ((size_t)jb.out_off + r) * k;
		if (threadIdx.x == 0) {   // 8-way merge of the per-warp sorted lists: the CTA's k best, sorted
			int ptr[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
			for (int j = 0; j < 8; j++) {
				int bw = -1;
				if (j < k) {
					for (int w2 = 0; w2 < 8; w2++) {
						if (ptr[w2] >= 8 || s_i[w2][ptr[w2]] == 0x7fffffff) continue;
						if (bw < 0 || dist_less(s_d[w2][ptr[w2]], s_i[w2][ptr[w2]], s_d[bw][ptr[bw]], s_i[bw][ptr[bw]])) bw = w2;
					}
				}
				const double md = bw < 0 ? 1e300 : s_d[bw][ptr[bw]];
				const int mi = bw < 0 ? 0x7fffffff : s_i[bw][ptr[bw]];
				if (bw >= 0) ptr[bw]++;
				if (!split) { if (j < k) { io[j] = bw < 0 ? -1 : mi; dd[j] = bw < 0 ? __int_as_float(0x7f800000) : (float)sqrt(md); } }
				else { part_d[((size_t)f * FB_SEG + seg) * 8 + j] = md; part_i[((size_t)f * FB_SEG + seg) * 8 + j] = mi; }
			}
			s_last = 0;
			if (split) { __threadfence(); s_last = (atomicAdd(tickets + f, 1) == FB_SEG - 1); }
		}
		__syncthreads();
		if (split && s_last) {     // FB_SEG sorted lists -> the row's k best (staged in shared memory: 128 dependent L2 reads cost 30 us)
			__shared__ double m_d[FB_SEG * 8];
			__shared__ int m_i[FB_SEG * 8];
			__threadfence();
			if (threadIdx.x < FB_SEG * 8) {
				m_d[threadIdx.x] = __ldcg(part_d + (size_t)f * FB_SEG * 8 + threadIdx.x);
				m_i[threadIdx.x] = __ldcg(part_i + (size_t)f * FB_SEG * 8 + threadIdx.x);
			}
			__syncthreads();
			if (threadIdx.x == 0) {
				tickets[f] = 0;
				int ptr[FB_SEG];
				for (int q = 0; q < FB_SEG; q++) ptr[q] = 0;
				for (int j = 0; j < k; j++) {
					int bq = -1;
					for (int q = 0; q < FB_SEG; q++) {
						if (ptr[q] >= 8 || m_i[q * 8 + ptr[q]] == 0x7fffffff) continue;
						if (bq < 0 || dist_less(m_d[q * 8 + ptr[q]], m_i[q * 8 + ptr[q]], m_d[bq * 8 + ptr[bq]], m_i[bq * 8 + ptr[bq]])) bq = q;
					}
					if (bq < 0) { io[j] = -1; dd[j] = __int_as_float(0x7f800000); }
					else { io[j] = m_i[bq * 8 + ptr[bq]]; dd[j] = (float)sqrt(m_d[bq * 8 + ptr[bq]]); ptr[bq]++; }
				}
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
	int pool_rows = 0;                      // capacity of the bf16 pool (rows)
	DevBuf pool, norms, errn, sets, set_max, set_err, items, cand, jobs, job_rows, fb_rows, fb_count, fb_part_d, fb_part_i, fb_tickets;
	PinnedBuf h_stage, h_small;
	PFN_encodeTiled encode = nullptr;
	CUtensorMap tmap_q, tmap_t;
	bool maps_ready = false;
	long long cand_rows_cap = 0;
	int max_items = 0, max_jobs = 0, max_rows_total = 0;
	int last_items = 0, last_rows = 0;
	cudaEvent_t ev[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
	bool timing = false;
	int force_fallback = 0;     // test knob: every n-th query row is sent to the exact fallback regardless of the proof
};

void matcher_destroy(bt_ctx* ctx) {
	MatcherState* m = ctx->matcher;
	if (!m) return;
	DevBuf* bufs[] = { &m->pool, &m->norms, &m->errn, &m->sets, &m->set_max, &m->set_err, &m->items, &m->cand, &m->jobs, &m->job_rows, &m->fb_rows, &m->fb_count, &m->fb_part_d, &m->fb_part_i, &m->fb_tickets };
	for (DevBuf* b : bufs) b->release();
	m->h_stage.release(); m->h_small.release();
	for (auto& e : m->ev) if (e) cudaEventDestroy(e);
	delete m;
	ctx->matcher = nullptr;
}

static int make_tmap(MatcherState* m, CUtensorMap* out, void* base, uint64_t rows, uint32_t box_rows) {
	const cuuint64_t gdim[2] = { (cuuint64_t)KD, rows };
	const cuuint64_t gstride[1] = { (cuuint64_t)KD * 2 };
	const cuuint32_t box[2] = { (cuuint32_t)BK, box_rows };
	const cuuint32_t estr[2] = { 1, 1 };
	const CUresult r = m->encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return BT_ERR_CUDA; }
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
	m->max_pairs = max_pairs; m->max_feats = max_feats; m->dim = dim;
	if (!m->encode) {
		void* fn = nullptr;
		cudaDriverEntryPointQueryResult qres;
		BT_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
		BT_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, BT_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
		m->encode = (PFN_encodeTiled)fn;
	}
	const int padded = (max_feats + BN - 1) / BN * BN;
	m->pool_rows = 2 * max_pairs * padded + BN;          // worst case: every pair uses two distinct descriptor sets
	const int qtiles = (max_feats + BM - 1) / BM;
	const int max_splits = 8;
	m->max_items = 2 * max_pairs * qtiles * max_splits;
	m->max_jobs = 2 * max_pairs;
	m->max_rows_total = 2 * max_pairs * max_feats;
	m->cand_rows_cap = (long long)2 * max_pairs * qtiles * BM * max_splits;
	int rc;
#define RES(buf, bytes) if ((rc = m->buf.alloc(bytes)) != BT_OK) return rc
	RES(pool, (size_t)m->pool_rows * KD * 2);
	RES(norms, (size_t)m->pool_rows * 4);
	RES(errn, (size_t)m->pool_rows * 4);
	RES(sets, sizeof(PrepSet) * 2 * max_pairs);
	RES(set_max, sizeof(int) * 2 * max_pairs);
	RES(set_err, sizeof(int) * 2 * max_pairs);
	RES(items, sizeof(KnnItem) * (size_t)m->max_items);
	RES(cand, sizeof(float) * NCAND * (size_t)m->cand_rows_cap);
	RES(jobs, sizeof(RerankJob) * m->max_jobs);
	RES(job_rows, sizeof(int) * (m->max_jobs + 1));
	RES(fb_rows, sizeof(int) * (size_t)m->max_rows_total);
	RES(fb_count, 16);
	RES(fb_part_d, sizeof(double) * (size_t)FB_SPLIT_ROWS * FB_SEG * 8); RES(fb_part_i, sizeof(int) * (size_t)FB_SPLIT_ROWS * FB_SEG * 8);
	RES(fb_tickets, sizeof(int) * (size_t)FB_SPLIT_ROWS);
	BT_CUDA(cudaMemset(m->fb_tickets.p, 0, sizeof(int) * (size_t)FB_SPLIT_ROWS));
#undef RES
	if ((rc = m->h_small.alloc(64)) != BT_OK) return rc;
	if ((rc = make_tmap(m, &m->tmap_q, m->pool.p, (uint64_t)m->pool_rows, BM)) != BT_OK) return rc;
	if ((rc = make_tmap(m, &m->tmap_t, m->pool.p, (uint64_t)m->pool_rows, BN)) != BT_OK) return rc;
	m->maps_ready = true;
	BT_CUDA(cudaFuncSetAttribute(k_knn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL));
	return BT_OK;
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
// ms4 = {descriptor prep, tensor-core pass, exact re-rank, exact fallback}; info3 = {work items, query rows, fallback rows}
extern "C" int bt_knn_get_timing(bt_ctx* ctx, float* ms4, int* info3) {
	BT_REQUIRE(ctx && ctx->matcher && ctx->matcher->timing && ms4 && info3, BT_ERR_INVALID_ARG, "bt_knn_get_timing: timing not enabled");
	MatcherState* m = ctx->matcher;
	BT_CUDA(cudaEventSynchronize(m->ev[4]));
	for (int k = 0; k < 4; k++) BT_CUDA(cudaEventElapsedTime(ms4 + k, m->ev[k], m->ev[k + 1]));
	int fb = 0;
	BT_CUDA(cudaMemcpy(&fb, m->fb_count.p, sizeof(int), cudaMemcpyDeviceToHost));
	info3[0] = m->last_items; info3[1] = m->last_rows; info3[2] = fb;
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
	// ---- unique descriptor sets (a keyframe appears in many pairs) -> pool slots
	std::map<std::pair<const float*, int>, int> slot_of;
	std::vector<PrepSet> sets;
	int pool_row = 0;
	auto slot = [&](const bt_desc_view& v) -> int {
		const auto key = std::make_pair(v.dev, v.n);
		auto it = slot_of.find(key);
		if (it != slot_of.end()) return it->second;
		PrepSet s; s.src = v.dev; s.pitch_bytes = v.pitch_bytes ? v.pitch_bytes : (size_t)v.dim * 4; s.n = v.n; s.row0 = pool_row;
		s.rows_padded = (std::max(v.n, 1) + BN - 1) / BN * BN; s.pad = 0;
		pool_row += s.rows_padded;
		sets.push_back(s);
		slot_of[key] = (int)sets.size() - 1;
		return (int)sets.size() - 1;
	};
	std::vector<int> sa(n_pairs), sb(n_pairs);
	for (int p = 0; p < n_pairs; p++) { sa[p] = slot(A[p]); sb[p] = slot(B[p]); }
	BT_REQUIRE(pool_row <= m->pool_rows, BT_ERR_CAPACITY, "bt_knn_match_pairs: descriptor pool overflow");
	// ---- work items + re-rank jobs.  Train splits: enough items to fill the machine; each split <= 8192 rows (13 index bits).
	long long base_items = 0;
	for (int p = 0; p < n_pairs; p++) base_items += (A[p].n + BM - 1) / BM + (B[p].n + BM - 1) / BM;
	std::vector<KnnItem> items;
	std::vector<RerankJob> jobs;
	std::vector<int> job_rows;
	long long cand_rows = 0; int total_rows = 0;
	size_t off_dir[2] = { 0, 0 };
	for (int p = 0; p < n_pairs; p++) {
		for (int dir = 0; dir < 2; dir++) {
			const bt_desc_view& Q = dir == 0 ? A[p] : B[p];
			const bt_desc_view& T = dir == 0 ? B[p] : A[p];
			const int qset = dir == 0 ? sa[p] : sb[p], tset = dir == 0 ? sb[p] : sa[p];
			const PrepSet& qs = sets[qset];
			const PrepSet& ts = sets[tset];
			const int qtiles = (Q.n + BM - 1) / BM, ttiles = ts.rows_padded / BN;
			// train splits: the epilogue is branch-free, so a split costs only the extra query-tile load; pick the smallest split
			// count whose item total fills whole waves of SMs to >= 90 % (or the best one found)
			int splits = 1;
			{
				double best_eff = 0.0;
				for (int sc = 1; sc <= 8; sc++) {
					const long long it_total = base_items * sc;
					const long long waves = (it_total + ctx->sm_count - 1) / ctx->sm_count;
					const double eff = (double)it_total / (double)(waves * ctx->sm_count);
					if (eff > best_eff + 1e-9) { best_eff = eff; splits = sc; }
					if (eff >= 0.9) { splits = sc; break; }
				}
			}
			splits = std::max(splits, (ts.rows_padded + MAX_SPLIT_ROWS - 1) / MAX_SPLIT_ROWS);
			splits = std::min(splits, std::max(ttiles, 1));
			BT_REQUIRE(splits <= 8, BT_ERR_CAPACITY, "pair %d: %d train rows need more than 8 splits", p, T.n);
			const int tiles_per_split = (ttiles + splits - 1) / splits;
			splits = (ttiles + tiles_per_split - 1) / std::max(tiles_per_split, 1);
			RerankJob jb; memset(&jb, 0, sizeof jb);
			jb.q = Q.dev; jb.q_pitch = qs.pitch_bytes; jb.nq = Q.n; jb.q_pool_row0 = qs.row0;
			jb.t = T.dev; jb.t_pitch = ts.pitch_bytes; jb.nt = T.n; jb.t_set = tset; jb.t_pool_row0 = ts.row0;
			jb.n_splits = (Q.n > 0 && T.n > 0) ? splits : 0; jb.split_rows = tiles_per_split * BN;
			jb.cand_off = (int)cand_rows; jb.cand_split_stride = qtiles * BM;
			jb.out_off = (int)off_dir[dir]; jb.dir = dir;
			if (Q.n > 0 && T.n > 0) {
				for (int s = 0; s < splits; s++) {
					for (int qt = 0; qt < qtiles; qt++) {
						KnnItem it; memset(&it, 0, sizeof it);
						it.q_row0 = qs.row0 + qt * BM; it.q_valid = std::min(BM, Q.n - qt * BM);
						it.t_row0 = ts.row0 + s * tiles_per_split * BN;
						it.t_tiles = std::min(tiles_per_split, ttiles - s * tiles_per_split);
						it.cand_off = (int)(cand_rows + (long long)s * qtiles * BM + (long long)qt * BM);
						items.push_back(it);
					}
				}
				cand_rows += (long long)splits * qtiles * BM;
			}
			jobs.push_back(jb);
			job_rows.push_back(total_rows);
			total_rows += Q.n;
			off_dir[dir] += Q.n;
		}
	}
	job_rows.push_back(total_rows);
	BT_REQUIRE((int)items.size() <= m->max_items && cand_rows <= m->cand_rows_cap && total_rows <= m->max_rows_total, BT_ERR_CAPACITY, "bt_knn_match_pairs: work list overflow");
	// ---- upload tables (one pinned block)
	const size_t b_sets = sizeof(PrepSet) * sets.size(), b_items = sizeof(KnnItem) * items.size(), b_jobs = sizeof(RerankJob) * jobs.size(), b_rows = sizeof(int) * job_rows.size();
	int rc = m->h_stage.alloc(b_sets + b_items + b_jobs + b_rows + 1024);
	if (rc != BT_OK) return rc;
	char* hb = m->h_stage.as<char>();
	memcpy(hb, sets.data(), b_sets);
	memcpy(hb + b_sets, items.data(), b_items);
	memcpy(hb + b_sets + b_items, jobs.data(), b_jobs);
	memcpy(hb + b_sets + b_items + b_jobs, job_rows.data(), b_rows);
	BT_CUDA(cudaMemcpyAsync(m->sets.p, hb, b_sets, cudaMemcpyHostToDevice, stream));
	if (b_items) BT_CUDA(cudaMemcpyAsync(m->items.p, hb + b_sets, b_items, cudaMemcpyHostToDevice, stream));
	BT_CUDA(cudaMemcpyAsync(m->jobs.p, hb + b_sets + b_items, b_jobs, cudaMemcpyHostToDevice, stream));
	BT_CUDA(cudaMemcpyAsync(m->job_rows.p, hb + b_sets + b_items + b_jobs, b_rows, cudaMemcpyHostToDevice, stream));
	BT_CUDA(cudaMemsetAsync(m->fb_count.p, 0, 16, stream));
	BT_CUDA(cudaMemsetAsync(m->set_max.p, 0, sizeof(int) * sets.size(), stream));
	BT_CUDA(cudaMemsetAsync(m->set_err.p, 0, sizeof(int) * sets.size(), stream));
	// ---- kernels
	if (m->timing) BT_CUDA(cudaEventRecord(m->ev[0], stream));
	int max_padded = 0;
	for (auto& s : sets) max_padded = std::max(max_padded, s.rows_padded);
	k_desc_prep<<<dim3((unsigned)std::max(1, std::min((max_padded + 7) / 8, 64)), (unsigned)sets.size()), 256, 0, stream>>>(
	    m->sets.as<PrepSet>(), m->pool.as<__nv_bfloat16>(), m->norms.as<float>(), m->errn.as<float>(), m->set_max.as<int>(), m->set_err.as<int>());
	if (m->timing) BT_CUDA(cudaEventRecord(m->ev[1], stream));
	if (!items.empty()) {
		const int grid = std::min((int)items.size(), ctx->sm_count);
		k_knn_tc<<<grid, KNN_THREADS, SMEM_TOTAL, stream>>>(m->tmap_q, m->tmap_t, m->items.as<KnnItem>(), (int)items.size(), m->norms.as<float>(), m->cand.as<float>());
	}
	if (m->timing) BT_CUDA(cudaEventRecord(m->ev[2], stream));
	KnnOut out; out.idx[0] = idxAB; out.idx[1] = idxBA; out.dist[0] = distAB; out.dist[1] = distBA;
	if (total_rows > 0) {
		k_knn_rerank<<<(total_rows + 7) / 8, 256, 0, stream>>>(m->jobs.as<RerankJob>(), m->job_rows.as<int>(), (int)jobs.size(), total_rows, m->cand.as<float>(),
		                                                     m->pool.as<__nv_bfloat16>(), m->norms.as<float>(), m->errn.as<float>(), m->set_max.as<int>(), m->set_err.as<int>(), k, out, m->fb_rows.as<int>(), m->fb_count.as<int>(), m->force_fallback);
		if (m->timing) BT_CUDA(cudaEventRecord(m->ev[3], stream));
		k_knn_exact<<<ctx->sm_count * 2, 256, 0, stream>>>(m->jobs.as<RerankJob>(), m->job_rows.as<int>(), (int)jobs.size(), m->fb_rows.as<int>(), m->fb_count.as<int>(), k, out,
		                                              m->fb_part_d.as<double>(), m->fb_part_i.as<int>(), m->fb_tickets.as<int>());
	} else if (m->timing) BT_CUDA(cudaEventRecord(m->ev[3], stream));
	if (m->timing) BT_CUDA(cudaEventRecord(m->ev[4], stream));
	BT_CUDA(cudaGetLastError());
	m->last_items = (int)items.size(); m->last_rows = total_rows;
	return BT_OK;
}
