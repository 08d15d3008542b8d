// Probe: which (lane, column) of tensor memory does register i of thread T receive from tcgen05.ld.16x256b.x8 ?  (developer aid)
// Writes value = (lane << 16) | column with tcgen05.st.32x32b (thread = lane, register = column), reads back with 16x256b.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(uint32_t* out) {
	__shared__ uint32_t s_addr;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_addr)), "r"(64u) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t base = s_addr;
	const int tl = warp * 32 + lane;      // TMEM lane of this thread for 32x32b
	for (int c0 = 0; c0 < 64; c0 += 8) {
		uint32_t v[8];
		for (int i = 0; i < 8; i++) v[i] = ((uint32_t)tl << 16) | (uint32_t)(c0 + i);
		const uint32_t ta = base + ((uint32_t)(warp * 32) << 16) + c0;
		asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(ta), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
	}
	asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	for (int half = 0; half < 2; half++) {
		uint32_t r[32];
		const uint32_t ta = base + ((uint32_t)(warp * 32 + half * 16) << 16);
		asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 "
		             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
		             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
		               "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
		               "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
		             : "r"(ta));
		asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
		for (int i = 0; i < 32; i++) out[((warp * 2 + half) * 32 + lane) * 32 + i] = r[i];
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64u) : "memory");
}
int main() {
	uint32_t* d; cudaMalloc(&d, 4 * 2 * 32 * 32 * 4);
	probe<<<1, 128>>>(d);
	cudaError_t e = cudaDeviceSynchronize();
	printf("status %s\n", cudaGetErrorString(e));
	static uint32_t h[4 * 2 * 32 * 32];
	cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
	int bad = 0;
	for (int w = 0; w < 4; w++) for (int half = 0; half < 2; half++) for (int T = 0; T < 32; T++) for (int i = 0; i < 32; i++) {
		const uint32_t v = h[((w * 2 + half) * 32 + T) * 32 + i];
		const int j = i >> 2, e2 = i & 3;
		const int row = w * 32 + half * 16 + T / 4 + 8 * (e2 >> 1), col = 8 * j + 2 * (T % 4) + (e2 & 1);
		if ((int)(v >> 16) != row || (int)(v & 0xffff) != col) { if (bad < 12) printf("w%d half%d T%d reg%d: got lane %u col %u, expected %d %d\n", w, half, T, i, v >> 16, v & 0xffff, row, col); bad++; }
	}
	printf("mismatches vs the m16n8-fragment hypothesis: %d\n", bad);
	for (int T = 0; T < 6; T++) { printf("T%d:", T); for (int i = 0; i < 8; i++) printf(" (%u,%u)", h[T * 32 + i] >> 16, h[T * 32 + i] & 0xffff); printf("\n"); }
	return 0;
}
