// dp_ungapped.cu -- SwAligner::ungappedAlign (aligner_sw.cpp:286-487): one thread per problem.
// End-to-end scoring with no match bonus is "monotone" (Scoring::monotone): the running score and
// the N count only get worse, so the early exits of the reference (:392-396) equal a test on the
// totals.  Otherwise (local mode) the floor-at-zero scan of :399-431 is replayed as written.
#include "dp_ungapped_device.cuh"

template <typename OFF>
__global__ void k_ungapped(DevIndex<OFF> ix, bt2g_scoring sc, const uint8_t *seq, const uint8_t *qual, const uint64_t *roff,
                           const bt2g_ungapped_problem *probs, uint64_t n, bt2g_ungapped_result *out, uint8_t *mask, uint32_t stride) {
	const uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if(w >= n) return;
	const bt2g_ungapped_problem p = probs[w];
	const uint8_t *rs = seq + roff[p.read_idx], *rq = qual + roff[p.read_idx];
	const int len = (int)(roff[p.read_idx + 1] - roff[p.read_idx]);
	bt2g_ungapped_result r;
	ungapped_one<OFF>(ix, sc, rs, rq, len, p, r, mask ? mask + w * (uint64_t)stride : nullptr, stride);
	out[w] = r;
}

template <typename OFF>
void launch_ungapped(const DevIndex<OFF> &ix, const bt2g_scoring &sc, const uint8_t *seq, const uint8_t *qual, const uint64_t *roff,
                     const bt2g_ungapped_problem *probs, uint64_t n, bt2g_ungapped_result *out, uint8_t *mask, uint32_t stride, cudaStream_t st) {
	if(n == 0) return;
	k_ungapped<OFF><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(ix, sc, seq, qual, roff, probs, n, out, mask, stride);
}
template void launch_ungapped<uint32_t>(const DevIndex<uint32_t> &, const bt2g_scoring &, const uint8_t *, const uint8_t *, const uint64_t *, const bt2g_ungapped_problem *, uint64_t, bt2g_ungapped_result *, uint8_t *, uint32_t, cudaStream_t);
template void launch_ungapped<uint64_t>(const DevIndex<uint64_t> &, const bt2g_scoring &, const uint8_t *, const uint8_t *, const uint64_t *, const bt2g_ungapped_problem *, uint64_t, bt2g_ungapped_result *, uint8_t *, uint32_t, cudaStream_t);
