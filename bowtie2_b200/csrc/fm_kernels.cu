// fm_kernels.cu -- K1 (seed search) and K2 (offset resolution) kernels for sm_100a.
//
// Work decomposition (B200-first, not the reference's): the CPU hides DRAM latency by
// round-robining 8 seeds per thread with software prefetch (aligner_seed.cpp:621-627,
// :1877-2035).  Here every (read, strand, seed) search is its own thread, so a full SM holds
// 2048 independent dependent-load chains (each with two 64 B / 128 B side fetches in flight
// per step); latency is hidden by occupancy, not by hand interleaving.  Seeds of one read sit
// in adjacent lanes so the read bytes they share come from the same L1 lines.
#include "fm_device.cuh"

// ----------------------------------------------------------------------------------------
// primitive kernels (used by the parity tests and by the host policy as pure functions)
// ----------------------------------------------------------------------------------------
template <typename OFF>
__global__ void k_rank4(DevEbwt<OFF> e, const uint64_t *rows, uint64_t n, uint64_t *out) {
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if(i >= n) return;
	uint64_t r[4];
	rank4<OFF>(e, rows[i], r);
	out[4 * i + 0] = r[0]; out[4 * i + 1] = r[1]; out[4 * i + 2] = r[2]; out[4 * i + 3] = r[3];
}

template <typename OFF>
__global__ void k_maplf1(DevEbwt<OFF> e, const uint64_t *rows, const uint8_t *chars, uint64_t n, uint64_t *out) {
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if(i >= n) return;
	out[i] = maplf1<OFF>(e, rows[i], chars[i]);
}

template <typename OFF>
__global__ void k_ftab(DevEbwt<OFF> e, const uint64_t *idx, uint64_t n, uint64_t *out) {
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if(i >= n) return;
	out[2 * i] = ftab_hi<OFF>(e, idx[i]);
	out[2 * i + 1] = ftab_lo<OFF>(e, idx[i] + 1);
}

// ----------------------------------------------------------------------------------------
// K1': exact end-to-end sweep.  One thread per (read, strand).
// SeedAligner::exactSweep (aligner_seed.cpp:856-970): match the whole read right-to-left,
// restarting after every empty range (ftab jump when 10 clean bases remain, else fchr) and
// counting restarts up to mineMax = 2 (bt2_search.cpp:3520).
// ----------------------------------------------------------------------------------------
template <typename OFF>
__global__ void k_exact_sweep(DevIndex<OFF> ix, const uint8_t *seq, const uint64_t *roff, uint64_t nReads,
                              int nofw, int norc, uint8_t *mine, uint64_t *ee, unsigned long long *cnt) {
	uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if(t >= nReads * 2) return;
	uint64_t rd = t >> 1;
	int strand = (int)(t & 1);
	uint64_t *eo = ee + rd * 4 + strand * 2;
	if((strand == 0 && nofw) || (strand == 1 && norc)) { mine[t] = 0; eo[0] = eo[1] = 0; return; }
	const uint8_t *s = seq + roff[rd];
	const int len = (int)(roff[rd + 1] - roff[rd]);
	const DevEbwt<OFF> &e = ix.fw;
	const int ftabLen = e.ftabChars;
	const int mineMax = 2;
	uint64_t top = 0, bot = 0;
	int dep = 0, nedit = 0;
	unsigned nside = 0;
	bool doInit = true, done = false;
	while(dep < len && !done) {
		if(doInit) {
			int left = len - dep;
			bool doFtab = ftabLen > 1 && left >= ftabLen;
			uint64_t fi = 0;
			if(doFtab) {
				// ftab index = the last ftabLen chars of the unmatched prefix, left to right
				for(int i = 0; i < ftabLen; i++) {
					int c = read_char(s, len, strand, left - ftabLen + i);
					if(c > 3) { doFtab = false; break; }
					fi = (fi << 2) | (uint64_t)c;
				}
			}
			top = bot = 0;
			if(doFtab) {
				top = ftab_hi<OFF>(e, fi); bot = ftab_lo<OFF>(e, fi + 1);
				dep += ftabLen;
			} else {
				int c = read_char(s, len, strand, len - dep - 1);
				if(c < 4) { top = e.fchr[c]; bot = e.fchr[c + 1]; }
				dep++;
			}
			if(bot <= top) {
				nedit++;
				if(nedit >= mineMax) done = true;
				continue;
			}
			doInit = false;
		}
		if(dep < len) {
			int c = read_char(s, len, strand, len - dep - 1);
			if(c > 3) {
				top = bot = 0;
			} else if(bot - top > 1) {
				// two independent side fetches in flight
				uint64_t nt = rank1<OFF>(e, top, c);
				uint64_t nb = rank1<OFF>(e, bot, c);
				top = nt; bot = nb; nside += 2;
			} else {
				uint64_t nt = maplf1<OFF>(e, top, c); nside++;
				if(nt == BT2G_OFFMASK) { top = bot = 0; } else { top = nt; bot = nt + 1; }
			}
			if(bot <= top) {
				nedit++;
				if(nedit >= mineMax) done = true;
				doInit = true;
			}
			dep++;
		}
	}
	if(cnt) atomicAdd(cnt, (unsigned long long)nside);
	mine[t] = (uint8_t)nedit;
	if(!done && nedit == 0 && bot > top) { eo[0] = top; eo[1] = bot; } else { eo[0] = eo[1] = 0; }
}

// ----------------------------------------------------------------------------------------
// K1: exact multiseed search.  One thread per (read, strand, seed offset index).
// Seed::instantiate SEED_TYPE_EXACT (aligner_seed.cpp:252-259, N rejection :326-352);
// ftab start (CacheAndSeed :88-112, startSearchSeedBi :1672-1689, NDEBUG branch: mirror range
// is [ftabHi(bwi0), +width)); then seedlen-ftabChars steps of mapBiLFEx (range > 1,
// bt2_idx.h:2372-2413) or mapLF1 (range == 1, :2420) (searchSeedBi :1858-2037).
// ----------------------------------------------------------------------------------------
template <typename OFF>
__global__ void k_seed_search(DevIndex<OFF> ix, const uint8_t *seq, const uint64_t *roff, uint64_t nReads,
                              int seedLen, int maxSeeds, int nofw, int norc,
                              const int32_t *interval, const int32_t *offset,
                              uint64_t *out, int32_t *nseedsOut, unsigned long long *cnt) {
	uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	uint64_t perRead = 2ull * maxSeeds;
	if(t >= nReads * perRead) return;
	uint64_t rd = t / perRead;
	int rem = (int)(t - rd * perRead);
	int strand = rem / maxSeeds;
	int k = rem - strand * maxSeeds;
	uint64_t *o = out + t * 4;
	const uint8_t *s = seq + roff[rd];
	const int len = (int)(roff[rd + 1] - roff[rd]);
	const int per = interval[rd], off0 = offset[rd];
	// instantiateSeeds (aligner_seed.cpp:523-526)
	int nseeds = 1;
	if(len - off0 > seedLen) nseeds += (len - off0 - seedLen) / per;
	if(rem == 0 && nseedsOut) nseedsOut[rd] = nseeds;
	o[0] = o[1] = o[2] = o[3] = 0;
	if(k >= nseeds || (strand == 0 && nofw) || (strand == 1 && norc)) return;
	const int sl = seedLen < len ? seedLen : len;
	const int depth = k * per + off0;
	if(depth + sl > len) return;
	// seed char j (Watson orientation): fw: s[depth+j]; rc: comp(s[depth+sl-1-j])
	auto seedChar = [&](int j) -> int {
		if(strand == 0) return s[depth + j];
		int c = s[depth + sl - 1 - j];
		return c > 3 ? 4 : 3 - c;
	};
	for(int j = 0; j < sl; j++) if(seedChar(j) > 3) return;   // exact seeds cannot absorb an N
	const DevEbwt<OFF> &fw = ix.fw;
	const DevEbwt<OFF> &bw = ix.bw;
	const int ftabLen = fw.ftabChars;
	uint64_t topf, botf, topb = 0, botb = 0;
	int step;
	unsigned nside = 0;
	if(ftabLen > 1 && ftabLen <= sl) {
		uint64_t fwi = 0, bwi = 0;
		for(int i = 0; i < ftabLen; i++) {
			fwi = (fwi << 2) | (uint64_t)seedChar(sl - ftabLen + i);
			bwi = (bwi << 2) | (uint64_t)seedChar(sl - 1 - i);
		}
		topf = ftab_hi<OFF>(fw, fwi); botf = ftab_lo<OFF>(fw, fwi + 1);
		if(botf <= topf) return;
		if(bw.ebwt != nullptr) { topb = ftab_hi<OFF>(bw, bwi); botb = topb + (botf - topf); }
		step = ftabLen;
	} else {
		int c = seedChar(sl - 1);
		topf = topb = fw.fchr[c]; botf = botb = fw.fchr[c + 1];
		if(botf <= topf) return;
		step = 1;
	}
	for(; step < sl; step++) {
		int c = seedChar(sl - step - 1);
		if(botf - topf > 1) {
			uint64_t tt[4], bb[4];
			rank4<OFF>(fw, topf, tt);
			rank4<OFF>(fw, botf, bb);
			nside += 2;
			uint64_t w0 = bb[0] - tt[0], w1 = bb[1] - tt[1], w2 = bb[2] - tt[2];
			uint64_t tp = topb + (c > 0 ? w0 : 0) + (c > 1 ? w1 : 0) + (c > 2 ? w2 : 0);
			uint64_t nt = c == 0 ? tt[0] : (c == 1 ? tt[1] : (c == 2 ? tt[2] : tt[3]));
			uint64_t nb = c == 0 ? bb[0] : (c == 1 ? bb[1] : (c == 2 ? bb[2] : bb[3]));
			if(nb <= nt) break;
			topf = nt; botf = nb; topb = tp; botb = tp + (nb - nt);
		} else {
			uint64_t nt = maplf1<OFF>(fw, topf, c); nside++;
			if(nt == BT2G_OFFMASK) break;
			topf = nt; botf = nt + 1;
		}
	}
	if(cnt) atomicAdd(cnt, (unsigned long long)nside);
	if(step < sl) return;   // died before the last step
	o[0] = topf; o[1] = botf; o[2] = topb; o[3] = botb;
}

// ----------------------------------------------------------------------------------------
// K2: SA-offset resolution + joined->text translation.  One thread per BW row.
// ----------------------------------------------------------------------------------------
template <typename OFF>
__global__ void k_resolve(DevIndex<OFF> ix, const uint64_t *rows, const uint32_t *hitlen, uint64_t n,
                          int rejectStraddle, uint64_t *joined, uint64_t *tidx, uint64_t *textoff,
                          uint64_t *tlen, uint8_t *flags, unsigned long long *cnt) {
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if(i >= n) return;
	if(rows[i] == BT2G_OFFMASK) { if(flags) flags[i] = 4; return; }   // empty slot
	unsigned nside = 0;
	uint64_t off = get_offset<OFF>(ix, rows[i], nside);
	if(cnt) atomicAdd(cnt, (unsigned long long)nside);
	if(joined) joined[i] = off;
	if(tidx || textoff || tlen || flags) {
		uint64_t ti, to, tl; bool st;
		bool ok = joined_to_text<OFF>(ix, hitlen ? hitlen[i] : 1, off, rejectStraddle != 0, ti, to, tl, st);
		if(tidx) tidx[i] = ti;
		if(textoff) textoff[i] = to;
		if(tlen) tlen[i] = tl;
		if(flags) flags[i] = (uint8_t)((st ? 1 : 0) | (ok ? 0 : 2));
	}
}

// ----------------------------------------------------------------------------------------
// SwDriver::extend (aligner_sw_driver.cpp:299-484): how far each seed hit extends without an
// edit, to the left with the forward index and to the right with the mirror index (<= 255).
// Two threads per seed hit (one per direction) so both walks run concurrently.
// ----------------------------------------------------------------------------------------
template <typename OFF>
__global__ void k_extend(DevIndex<OFF> ix, const uint8_t *seq, const uint64_t *roff, uint64_t nReads,
                         int seedLen, int maxSeeds, const int32_t *interval, const int32_t *offset,
                         const uint64_t *ranges, uint8_t *out) {
	uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	const uint64_t perRead = 4ull * maxSeeds;     // [strand][seed][direction]
	if(t >= nReads * perRead) return;
	const uint64_t rd = t / perRead;
	int rem = (int)(t - rd * perRead);
	const int dir = rem & 1; rem >>= 1;
	const int strand = rem / maxSeeds, k = rem - strand * maxSeeds;
	out[t] = 0;
	const uint64_t *q = ranges + ((rd * 2 + strand) * maxSeeds + k) * 4;
	if(q[1] <= q[0]) return;
	const uint8_t *s = seq + roff[rd];
	const int len = (int)(roff[rd + 1] - roff[rd]);
	const int sl = seedLen < len ? seedLen : len;
	const int off = k * interval[rd] + offset[rd];
	const bool fw = strand == 0;
	uint32_t nl = 0, nr = 0;
	extend_hit<OFF>(ix, q, s, len, fw, off, sl, dir == 0, dir != 0, nl, nr);
	out[t] = (uint8_t)(dir == 0 ? nl : nr);
}

template <typename OFF>
void launch_extend(const DevIndex<OFF> &ix, const uint8_t *seq, const uint64_t *roff, uint64_t nReads, int seedLen, int maxSeeds,
                   const int32_t *interval, const int32_t *offset, const uint64_t *ranges, uint8_t *out, cudaStream_t st) {
	uint64_t n = nReads * 4ull * maxSeeds;
	if(n) k_extend<OFF><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(ix, seq, roff, nReads, seedLen, maxSeeds, interval, offset, ranges, out);
}
template void launch_extend<uint32_t>(const DevIndex<uint32_t> &, const uint8_t *, const uint64_t *, uint64_t, int, int, const int32_t *, const int32_t *, const uint64_t *, uint8_t *, cudaStream_t);
template void launch_extend<uint64_t>(const DevIndex<uint64_t> &, const uint8_t *, const uint64_t *, uint64_t, int, int, const int32_t *, const int32_t *, const uint64_t *, uint8_t *, cudaStream_t);

template <typename OFF>
__global__ void k_get_stretch(DevIndex<OFF> ix, const uint64_t *tidx, const int64_t *off, const int32_t *count,
                              uint64_t n, int stride, uint8_t *out) {
	uint64_t i = blockIdx.x;
	if(i >= n) return;
	for(int k = threadIdx.x; k < count[i]; k += blockDim.x) {
		out[i * (uint64_t)stride + k] = (uint8_t)ref_base<OFF>(ix, tidx[i], off[i] + k);
	}
}

// ----------------------------------------------------------------------------------------
// launchers (called from api.cu)
// ----------------------------------------------------------------------------------------
static inline unsigned gridFor(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

// Ebwt::mapLFRange (bt2_idx.h:2268-2305) = countBt2SideRange (:1804-1865) + countBt2SideRange2 (:2177-2239): the GroupWalk step
// of a range that is still several rows wide (group_walk.h:897).  One warp per range.  Lane 0 takes the four ranks at `top`
// (one side fetch, as rank4).  The rows are then walked side by side: the first WORDS lanes pull the side's BWT words with one
// coalesced load, every lane decodes the characters of its rows out of the word it gets by shuffle, writes them (32 consecutive
// bytes per warp store) and tallies them; the tallies are reduced over the warp.  The reference's four bool lists are
// masks[c][j] == (chars[j] == c); like the reference (:2209) the "$" row counts as an A in `in`, and not in `upto`.
template <typename OFF>
__global__ void k_maplf_range(DevEbwt<OFF> e, const uint64_t *tops, const uint64_t *nums, const uint64_t *rowOff, uint64_t n,
                              uint64_t *upto, uint64_t *in, uint8_t *chars) {
	constexpr uint32_t BL = SideGeom<OFF>::BWT_LEN, WORDS = SideGeom<OFF>::WORDS, SIDE_SZ = SideGeom<OFF>::SIDE_SZ;
	const uint32_t lane = threadIdx.x & 31;
	const uint64_t w = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
	if(w >= n) return;
	const uint64_t top = tops[w], num = nums[w];
	uint8_t *out = chars + rowOff[w];
	if(lane == 0) {
		uint64_t r[4];
		rank4<OFF>(e, top, r);
		upto[4 * w + 0] = r[0]; upto[4 * w + 1] = r[1]; upto[4 * w + 2] = r[2]; upto[4 * w + 3] = r[3];
	}
	uint64_t nA = 0, nC = 0, nG = 0, nT = 0, done = 0;
	uint64_t sideNum = top / BL;
	uint32_t charOff = (uint32_t)(top - sideNum * BL);
	while(done < num) {
		const uint64_t left = num - done;
		const uint32_t take = left < (uint64_t)(BL - charOff) ? (uint32_t)left : BL - charOff;
		uint64_t word = 0;
		if(lane < WORDS) word = __ldg((const unsigned long long *)(e.ebwt + sideNum * SIDE_SZ) + lane);
		for(uint32_t base = 0; base < take; base += 32) {
			const uint32_t k = base + lane, pos = charOff + k;
			const uint32_t src = (pos >> 5) < WORDS ? (pos >> 5) : WORDS - 1;
			const uint64_t ww = __shfl_sync(0xffffffffu, word, src);
			if(k < take) {
				const uint32_t c = (uint32_t)(ww >> ((pos & 31) * 2)) & 3;
				out[done + k] = (uint8_t)c;
				nA += c == 0; nC += c == 1; nG += c == 2; nT += c == 3;
			}
		}
		done += take;
		sideNum++; charOff = 0;                          // SideLocus::nextSide (:1857)
	}
#pragma unroll
	for(int d = 16; d > 0; d >>= 1) {
		nA += __shfl_xor_sync(0xffffffffu, nA, d); nC += __shfl_xor_sync(0xffffffffu, nC, d);
		nG += __shfl_xor_sync(0xffffffffu, nG, d); nT += __shfl_xor_sync(0xffffffffu, nT, d);
	}
	if(lane == 0) { in[4 * w + 0] = nA; in[4 * w + 1] = nC; in[4 * w + 2] = nG; in[4 * w + 3] = nT; }
}

template <typename OFF>
void launch_rank4(const DevEbwt<OFF> &e, const uint64_t *rows, uint64_t n, uint64_t *out, cudaStream_t st) {
	if(n) k_rank4<OFF><<<gridFor(n, 256), 256, 0, st>>>(e, rows, n, out);
}
template <typename OFF>
void launch_maplf1(const DevEbwt<OFF> &e, const uint64_t *rows, const uint8_t *chars, uint64_t n, uint64_t *out, cudaStream_t st) {
	if(n) k_maplf1<OFF><<<gridFor(n, 256), 256, 0, st>>>(e, rows, chars, n, out);
}
template <typename OFF>
void launch_maplf_range(const DevEbwt<OFF> &e, const uint64_t *tops, const uint64_t *nums, const uint64_t *rowOff, uint64_t n,
                        uint64_t *upto, uint64_t *in, uint8_t *chars, cudaStream_t st) {
	if(n) k_maplf_range<OFF><<<gridFor(n * 32, 256), 256, 0, st>>>(e, tops, nums, rowOff, n, upto, in, chars);
}
template <typename OFF>
void launch_ftab(const DevEbwt<OFF> &e, const uint64_t *idx, uint64_t n, uint64_t *out, cudaStream_t st) {
	if(n) k_ftab<OFF><<<gridFor(n, 256), 256, 0, st>>>(e, idx, n, out);
}
template <typename OFF>
void launch_exact_sweep(const DevIndex<OFF> &ix, const uint8_t *seq, const uint64_t *roff, uint64_t nReads,
                        int nofw, int norc, uint8_t *mine, uint64_t *ee, cudaStream_t st, unsigned long long *cnt) {
	if(nReads) k_exact_sweep<OFF><<<gridFor(nReads * 2, 128), 128, 0, st>>>(ix, seq, roff, nReads, nofw, norc, mine, ee, cnt);
}
template <typename OFF>
void launch_seed_search(const DevIndex<OFF> &ix, const uint8_t *seq, const uint64_t *roff, uint64_t nReads,
                        int seedLen, int maxSeeds, int nofw, int norc, const int32_t *interval,
                        const int32_t *offset, uint64_t *out, int32_t *nseeds, cudaStream_t st, unsigned long long *cnt) {
	uint64_t n = nReads * 2ull * maxSeeds;
	if(n) k_seed_search<OFF><<<gridFor(n, 128), 128, 0, st>>>(ix, seq, roff, nReads, seedLen, maxSeeds, nofw, norc,
	                                                        interval, offset, out, nseeds, cnt);
}
template <typename OFF>
void launch_resolve(const DevIndex<OFF> &ix, const uint64_t *rows, const uint32_t *hitlen, uint64_t n, int rej,
                    uint64_t *joined, uint64_t *tidx, uint64_t *textoff, uint64_t *tlen, uint8_t *flags, cudaStream_t st, unsigned long long *cnt) {
	if(n) k_resolve<OFF><<<gridFor(n, 128), 128, 0, st>>>(ix, rows, hitlen, n, rej, joined, tidx, textoff, tlen, flags, cnt);
}
template <typename OFF>
void launch_get_stretch(const DevIndex<OFF> &ix, const uint64_t *tidx, const int64_t *off, const int32_t *count,
                        uint64_t n, int stride, uint8_t *out, cudaStream_t st) {
	if(n) k_get_stretch<OFF><<<(unsigned)n, 64, 0, st>>>(ix, tidx, off, count, n, stride, out);
}

#define INSTANTIATE(OFF)                                                                                          \
	template void launch_rank4<OFF>(const DevEbwt<OFF> &, const uint64_t *, uint64_t, uint64_t *, cudaStream_t);  \
	template void launch_maplf1<OFF>(const DevEbwt<OFF> &, const uint64_t *, const uint8_t *, uint64_t, uint64_t *, cudaStream_t); \
	template void launch_maplf_range<OFF>(const DevEbwt<OFF> &, const uint64_t *, const uint64_t *, const uint64_t *, uint64_t, uint64_t *, uint64_t *, uint8_t *, cudaStream_t); \
	template void launch_ftab<OFF>(const DevEbwt<OFF> &, const uint64_t *, uint64_t, uint64_t *, cudaStream_t);   \
	template void launch_exact_sweep<OFF>(const DevIndex<OFF> &, const uint8_t *, const uint64_t *, uint64_t, int, int, uint8_t *, uint64_t *, cudaStream_t, unsigned long long *); \
	template void launch_seed_search<OFF>(const DevIndex<OFF> &, const uint8_t *, const uint64_t *, uint64_t, int, int, int, int, const int32_t *, const int32_t *, uint64_t *, int32_t *, cudaStream_t, unsigned long long *); \
	template void launch_resolve<OFF>(const DevIndex<OFF> &, const uint64_t *, const uint32_t *, uint64_t, int, uint64_t *, uint64_t *, uint64_t *, uint64_t *, uint8_t *, cudaStream_t, unsigned long long *); \
	template void launch_get_stretch<OFF>(const DevIndex<OFF> &, const uint64_t *, const int64_t *, const int32_t *, uint64_t, int, uint8_t *, cudaStream_t);
INSTANTIATE(uint32_t)
INSTANTIATE(uint64_t)
