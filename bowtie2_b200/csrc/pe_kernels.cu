// pe_kernels.cu -- paired-end framing over arrays of anchors / pairs (one thread each).
#include "bt2g_internal.h"
#include "pe_device.cuh"

__global__ void k_frame_mate(bt2g_pe_policy pp, const bt2g_mate_anchor *anchors, uint64_t n, bt2g_mate_frame *out) {
	const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if(i >= n) return;
	bt2g_mate_frame f;
	pe_frame_anchor(pp, anchors[i], f);
	out[i] = f;
}

__global__ void k_pe_classify(bt2g_pe_policy pp, const int64_t *pairs, uint64_t n, int32_t *out) {
	const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if(i >= n) return;
	const int64_t *q = pairs + 6 * i;
	out[i] = pe_classify(pp, q[0], (uint64_t)q[1], q[2] != 0, q[3], (uint64_t)q[4], q[5] != 0);
}

void launch_frame_mate(const bt2g_pe_policy &pp, const bt2g_mate_anchor *anchors, uint64_t n, bt2g_mate_frame *out, cudaStream_t st) {
	if(n == 0) return;
	k_frame_mate<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(pp, anchors, n, out);
}

void launch_pe_classify(const bt2g_pe_policy &pp, const int64_t *pairs, uint64_t n, int32_t *out, cudaStream_t st) {
	if(n == 0) return;
	k_pe_classify<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(pp, pairs, n, out);
}
