// tools/gather_bench.cu -- micro-benchmark: how fast can a B200 gather random 64 B "sides"
// from a multi-GB array, and which load flavour keeps DRAM traffic at the algorithmic 64 B?
// (measurement tool for DESIGN.md section 3, not part of the product)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

template <int MODE> __device__ __forceinline__ uint32_t load64(const uint8_t *p) {
	uint32_t acc = 0;
	if(MODE == 0) {            // plain 4 x LDG.128
		const uint4 *q = (const uint4 *)p;
		#pragma unroll
		for(int i = 0; i < 4; i++) { uint4 v = q[i]; acc += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w); }
	} else if(MODE == 1) {     // nc + L1::no_allocate
		#pragma unroll
		for(int i = 0; i < 4; i++) { uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + 16 * i)); acc += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w); }
	} else if(MODE == 2) {     // nc + L1::no_allocate + L2::64B
		#pragma unroll
		for(int i = 0; i < 4; i++) { uint4 v; asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + 16 * i)); acc += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w); }
	} else if(MODE == 3) {     // 2 x 256-bit loads
		#pragma unroll
		for(int i = 0; i < 2; i++) { uint64_t a, b, c, d; asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p + 32 * i)); acc += __popcll(a) + __popcll(b) + __popcll(c) + __popcll(d); }
	} else if(MODE == 4) {     // 2 x 256-bit loads, L2::64B
		#pragma unroll
		for(int i = 0; i < 2; i++) { uint64_t a, b, c, d; asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p + 32 * i)); acc += __popcll(a) + __popcll(b) + __popcll(c) + __popcll(d); }
	} else {                   // default caching 2 x 256-bit
		#pragma unroll
		for(int i = 0; i < 2; i++) { uint64_t a, b, c, d; asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p + 32 * i)); acc += __popcll(a) + __popcll(b) + __popcll(c) + __popcll(d); }
	}
	return acc;
}

// DEP = 1: each load's address depends on the previous load's data (an LF chain)
template <int MODE, int DEP>
__global__ void k_gather(const uint8_t *base, uint64_t nSides, int iters, uint32_t *out) {
	uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	uint64_t h = mix(t + 12345);
	uint32_t acc = 0;
	for(int i = 0; i < iters; i++) {
		uint64_t side = h % nSides;
		uint32_t v = load64<MODE>(base + side * 64);
		acc += v;
		h = mix(h + (DEP ? v : 0) + i);
	}
	out[t] = acc;
}

template <int MODE, int DEP> void run(const uint8_t *d, uint64_t nSides, uint32_t *out, int blocks, int threads, int iters) {
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	k_gather<MODE, DEP><<<blocks, threads>>>(d, nSides, 4, out);
	cudaDeviceSynchronize();
	cudaEventRecord(a);
	k_gather<MODE, DEP><<<blocks, threads>>>(d, nSides, iters, out);
	cudaEventRecord(b); cudaEventSynchronize(b);
	float ms; cudaEventElapsedTime(&ms, a, b);
	double bytes = (double)blocks * threads * iters * 64.0;
	printf("mode %d dep %d blocks %d x %d: %.3f ms  %.1f GB/s algorithmic (64 B per gather)\n", MODE, DEP, blocks, threads, ms, bytes / ms / 1e6);
}

int main(int argc, char **argv) {
	uint64_t bytes = 1ull << 30;
	if(argc > 1) bytes = (uint64_t)atoll(argv[1]) << 20;
	uint8_t *d; cudaMalloc(&d, bytes); cudaMemset(d, 0x5a, bytes);
	uint64_t nSides = bytes / 64;
	int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	uint32_t *out; cudaMalloc(&out, (size_t)sms * 2048 * 4 * 4);
	for(int occ = 1; occ <= 2; occ++) {
		int blocks = sms * 8 * occ, threads = 256, iters = 256;   // 2048 / 4096 threads per SM (2nd = two waves)
		run<0, 1>(d, nSides, out, blocks, threads, iters); run<1, 1>(d, nSides, out, blocks, threads, iters);
		run<2, 1>(d, nSides, out, blocks, threads, iters); run<3, 1>(d, nSides, out, blocks, threads, iters);
		run<4, 1>(d, nSides, out, blocks, threads, iters); run<5, 1>(d, nSides, out, blocks, threads, iters);
	}
	run<1, 0>(d, nSides, out, sms * 8, 256, 256); run<4, 0>(d, nSides, out, sms * 8, 256, 256);
	return 0;
}
