// fm_onemm.cu -- K1'': end-to-end search with at most one mismatch
// (SeedAligner::oneMmSearch, aligner_seed.cpp:975-1325, called with repex=false, rep1mm=true from
// bt2_search.cpp:3709).  One thread per (read, strand, pass): pass 0 walks right-to-left in the
// forward index, pass 1 left-to-right in the mirror index; the half of the read nearest the start
// must match exactly, every position of the far half may take one of the other nucleotides, after
// which the rest must match exactly.  Both BW ranges (primary and the other index) are carried so
// that a hit found in the mirror index is reported as its forward-index range.
// All steps use ranks at top and bot (for a width-1 range this equals Ebwt::mapLF1, see fm_seed2.cu).
#include "fm_device.cuh"

struct OneMmCtx {
	const uint8_t *s, *q;
	int len, fw, ebwtfw;
	// seq = fw ? (ebwtfw ? patFw : patFwRev) : (ebwtfw ? patRc : patRcRev)   (aligner_seed.cpp:1036-1039)
	__device__ __forceinline__ int chr(int p) const {
		const int pp = ebwtfw ? p : len - 1 - p;           // index into patFw / patRc
		if(fw) return s[pp];
		const int c = s[len - 1 - pp];
		return c > 3 ? 4 : 3 - c;
	}
	// qual = fw ? (ebwtfw ? qual : qualRev) : (ebwtfw ? qualRev : qual)       (:1041-1043)
	__device__ __forceinline__ int qual(int p) const {
		const bool rev = fw ? !ebwtfw : ebwtfw;
		return q[rev ? len - 1 - p : p];
	}
};

template <typename OFF>
__device__ __forceinline__ void bi_step(const DevEbwt<OFF> &e, uint64_t top, uint64_t bot, uint64_t topp,
                                        uint64_t t[4], uint64_t b[4], uint64_t tp[4], uint64_t bp[4]) {
	rank4<OFF>(e, top, t);
	rank4<OFF>(e, bot, b);
	// mirror range by prefix sums of the widths (mapBiLFEx, bt2_idx.h:2404-2412)
	uint64_t acc = topp;
#pragma unroll
	for(int j = 0; j < 4; j++) { tp[j] = acc; acc += b[j] - t[j]; bp[j] = acc; }
}

// TEXT: once the range is ONE row the search continues in the joined text itself (ix.refBuf, the 2-bit packed reference = the
// joined text): every further LF step of a single row yields the text character next to the occurrence, so the rest of the read
// is compared with the text at the occurrence's joined offset (one SA lookup), and a hit is returned as that offset with
// BT2G_ROW_IS_OFFSET set in `top` (bot = top + 1) instead of its forward-index row -- the engine resolves rows to offsets anyway
// (csrc/xengine.cu: DevSvc::resolve).  The ABI entry point bt2g_one_mm keeps TEXT = false (rows).
template <typename OFF, bool TEXT>
__device__ __forceinline__ bool one_mm_text(const DevIndex<OFF> &ix, const OneMmCtx &c, uint64_t rowFw, int dep, int nea, int ns, int nceil,
                                            const bt2g_scoring &sc, int minsc, int maxHits, bt2g_mm_hit *out, int &nh) {
	const int len = c.len;
	const int64_t tlen = (int64_t)ix.fw.len;
	unsigned nside = 0;
	const int64_t P = (int64_t)get_offset<OFF>(ix, rowFw, nside);       // joined offset where the dep matched characters start
	// text position consumed by step d: leftwards from the occurrence in the forward-index pass, rightwards in the mirror pass
	const int64_t base = c.ebwtfw ? P + dep : P;
	auto tidx = [&](int d) -> int64_t { return c.ebwtfw ? base - 1 - d : base + d; };
	auto tchr = [&](int64_t b) -> int { return (int)((__ldg(ix.refBuf + (b >> 2)) >> ((b & 3) << 1)) & 3); };
	for(; dep < nea; dep++) {                                           // near half: exact
		const int64_t b = tidx(dep);
		if(b < 0 || b >= tlen || tchr(b) != c.chr(len - dep - 1)) return true;
	}
	for(; dep < len; dep++) {                                           // far half: one substitution allowed
		const int rdc = c.chr(len - dep - 1);
		const int quc = c.qual(len - dep - 1);
		if(rdc > 3 && nceil == 0) break;
		const int64_t b = tidx(dep);
		if(b < 0 || b >= tlen) break;                                   // the "$" row
		const int tc = tchr(b);
		if((ns == 0 || rdc > 3) && tc != rdc) {
			int depm = dep + 1;
			for(; depm < len; depm++) {
				const int rdcm = c.chr(len - depm - 1);
				if(rdcm > 3) break;
				const int64_t bm = tidx(depm);
				if(bm < 0 || bm >= tlen || tchr(bm) != rdcm) break;
			}
			if(depm == len) {
				int off5p = dep;
				if(c.fw == c.ebwtfw) off5p = len - off5p - 1;
				int qq = quc - 33; qq = qq < 0 ? 0 : (qq > 63 ? 63 : qq);
				const int pen = rdc > 3 ? -(int)sc.npen[qq] : -(int)sc.mmpen[qq];
				const int score = (len - 1) * sc.match_bonus + pen;
				bool valid = true;
				if(sc.local) {
					int lf = 0, lb = 0;
					for(int i = 0; i < len && valid; i++) {
						if(i == dep) { if(lf + pen <= 0) valid = false; lf += pen; } else lf += sc.match_bonus;
						if(len - i - 1 == dep) { if(lb + pen <= 0) valid = false; lb += pen; } else lb += sc.match_bonus;
					}
				}
				if(valid && score >= minsc) {
					if(nh < maxHits) {
						bt2g_mm_hit &h = out[nh];
						h.top = BT2G_ROW_IS_OFFSET | (uint64_t)(c.ebwtfw ? base - len : base); h.bot = h.top + 1;
						h.pos = off5p; h.chr = tc; h.qchr = rdc; h.score = score;
					}
					nh++;
				}
			}
		}
		if(rdc > 3 || tc != rdc) break;
		if(dep == len - 1) break;
	}
	return true;
}

template <typename OFF, bool TEXT>
__global__ void k_one_mm(DevIndex<OFF> ix, const uint8_t *seq, const uint8_t *qual, const uint64_t *roff, uint64_t nReads,
                         const int32_t *minsc, const uint8_t *strandMask, bt2g_scoring sc, int maxHits,
                         bt2g_mm_hit *hits, int32_t *counts, const uint32_t *sel, const uint32_t *nDev) {
	// sel != nullptr: request slot s searches read sel[s] (minsc / strandMask / hits / counts are indexed by slot);
	// nDev != nullptr: the number of slots is a device-side count
	uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if(nDev) nReads = *nDev;
	if(t >= nReads * 4) return;
	const uint64_t slot = t >> 2;
	const uint64_t rd = sel ? sel[slot] : slot;
	const int fwi = (int)((t >> 1) & 1), pass = (int)(t & 1);
	counts[t] = 0;
	if(!((strandMask[slot] >> fwi) & 1)) return;
	OneMmCtx c;
	c.s = seq + roff[rd]; c.q = qual + roff[rd];
	c.len = (int)(roff[rd + 1] - roff[rd]); c.fw = fwi == 0; c.ebwtfw = pass == 0;
	const int len = c.len;
	if(len < 2 || ix.bw.ebwt == nullptr) return;
	int ns = 0;
	for(int i = 0; i < len; i++) ns += c.s[i] > 3;
	if(ns > 1) return;                                          // :991-994
	const int nceil = (int)(sc.nceil_const + sc.nceil_linear * (double)len);    // SimpleFunc::f<int> (simple_func.h:89-111)
	const DevEbwt<OFF> &e = c.ebwtfw ? ix.fw : ix.bw;
	const DevEbwt<OFF> &ep = c.ebwtfw ? ix.bw : ix.fw;
	const int ftabLen = e.ftabChars;
	const int nea = c.ebwtfw ? (len >> 1) : ((len >> 1) + (len & 1));
	for(int dep = 0; dep < nea; dep++) if(c.chr(len - dep - 1) > 3) return;     // N in the near half
	uint64_t top, bot, topp, botp;
	int dep;
	if(ftabLen > 1 && ftabLen <= nea) {
		uint64_t fi = 0, fip = 0;
		for(int i = 0; i < ftabLen; i++) {
			fi = (fi << 2) | (uint64_t)c.chr(len - ftabLen + i);          // primary: left to right
			fip = (fip << 2) | (uint64_t)c.chr(len - 1 - i);              // other index: right to left
		}
		top = ftab_hi<OFF>(e, fi); bot = ftab_lo<OFF>(e, fi + 1);
		topp = ftab_hi<OFF>(ep, fip); botp = ftab_lo<OFF>(ep, fip + 1);
		if(bot <= top) return;
		dep = ftabLen;
	} else {
		const int ch = c.chr(len - 1);
		top = topp = e.fchr[ch]; bot = botp = e.fchr[ch + 1];
		if(bot <= top) return;
		dep = 1;
	}
	uint64_t tt[4], bb[4], tp[4], bp[4];
	int nh = 0;
	bt2g_mm_hit *out = hits + t * (uint64_t)maxHits;
	const bool text = TEXT && ix.refBuf != nullptr;
	// near half: exact
	for(; dep < nea; dep++) {
		if(text && bot - top == 1) { one_mm_text<OFF, TEXT>(ix, c, c.ebwtfw ? top : topp, dep, nea, ns, nceil, sc, minsc[slot], maxHits, out, nh); counts[t] = nh; return; }
		const int rdc = c.chr(len - dep - 1);
		bi_step<OFF>(e, top, bot, topp, tt, bb, tp, bp);
		const uint64_t nt = rdc == 0 ? tt[0] : (rdc == 1 ? tt[1] : (rdc == 2 ? tt[2] : tt[3]));
		const uint64_t nb = rdc == 0 ? bb[0] : (rdc == 1 ? bb[1] : (rdc == 2 ? bb[2] : bb[3]));
		if(nb <= nt) return;
		topp = rdc == 0 ? tp[0] : (rdc == 1 ? tp[1] : (rdc == 2 ? tp[2] : tp[3]));
		botp = topp + (nb - nt);
		top = nt; bot = nb;
	}
	// far half: one substitution allowed
	for(; dep < len; dep++) {
		if(text && bot - top == 1) { one_mm_text<OFF, TEXT>(ix, c, c.ebwtfw ? top : topp, dep, nea, ns, nceil, sc, minsc[slot], maxHits, out, nh); break; }
		const int rdc = c.chr(len - dep - 1);
		const int quc = c.qual(len - dep - 1);
		if(rdc > 3 && nceil == 0) break;
		if(bot - top == 1 && top == e.zOff) break;                  // mapLF1 returned -1: hit the "$" (:1150-1152)
		bi_step<OFF>(e, top, bot, topp, tt, bb, tp, bp);
		if(ns == 0 || rdc > 3) {
			for(int j = 0; j < 4; j++) {
				if(j == rdc || bb[j] == tt[j]) continue;
				// mismatch branch: the remainder must match exactly (:1176-1212)
				uint64_t topm = tt[j], botm = bb[j], topmp = tp[j], botmp = bp[j];
				int depm = dep + 1;
				for(; depm < len; depm++) {
					const int rdcm = c.chr(len - depm - 1);
					if(rdcm > 3) break;
					uint64_t tm[4], bm[4], tmp[4], bmp[4];
					bi_step<OFF>(e, topm, botm, topmp, tm, bm, tmp, bmp);
					const uint64_t nt = rdcm == 0 ? tm[0] : (rdcm == 1 ? tm[1] : (rdcm == 2 ? tm[2] : tm[3]));
					const uint64_t nb = rdcm == 0 ? bm[0] : (rdcm == 1 ? bm[1] : (rdcm == 2 ? bm[2] : bm[3]));
					if(nb <= nt) break;
					topmp = rdcm == 0 ? tmp[0] : (rdcm == 1 ? tmp[1] : (rdcm == 2 ? tmp[2] : tmp[3]));
					botmp = topmp + (nb - nt);
					topm = nt; botm = nb;
				}
				if(depm == len) {
					int off5p = dep;
					if(c.fw == c.ebwtfw) off5p = len - off5p - 1;
					int qq = quc - 33; qq = qq < 0 ? 0 : (qq > 63 ? 63 : qq);
					const int pen = rdc > 3 ? -(int)sc.npen[qq] : -(int)sc.mmpen[qq];     // sc.score(rdc, 1<<j, q)
					const int score = (len - 1) * sc.match_bonus + pen;
					bool valid = true;
					if(sc.local) {                                            // :1238-1262
						int lf = 0, lb = 0;
						for(int i = 0; i < len && valid; i++) {
							if(i == dep) { if(lf + pen <= 0) valid = false; lf += pen; } else lf += sc.match_bonus;
							if(len - i - 1 == dep) { if(lb + pen <= 0) valid = false; lb += pen; } else lb += sc.match_bonus;
						}
					}
					if(valid && score >= minsc[slot]) {
						if(nh < maxHits) {
							bt2g_mm_hit &h = out[nh];
							h.top = c.ebwtfw ? topm : topmp; h.bot = c.ebwtfw ? botm : botmp;
							h.pos = off5p; h.chr = j; h.qchr = rdc; h.score = score;
						}
						nh++;
					}
				}
			}
		}
		if(rdc > 3) break;
		const uint64_t nt = rdc == 0 ? tt[0] : (rdc == 1 ? tt[1] : (rdc == 2 ? tt[2] : tt[3]));
		const uint64_t nb = rdc == 0 ? bb[0] : (rdc == 1 ? bb[1] : (rdc == 2 ? bb[2] : bb[3]));
		if(nb <= nt) break;
		if(dep == len - 1) break;                                   // exact hit: not reported here (repex = false)
		topp = rdc == 0 ? tp[0] : (rdc == 1 ? tp[1] : (rdc == 2 ? tp[2] : tp[3]));
		botp = topp + (nb - nt);
		top = nt; bot = nb;
	}
	counts[t] = nh;
}

template <typename OFF>
void launch_one_mm(const DevIndex<OFF> &ix, const uint8_t *seq, const uint8_t *qual, const uint64_t *roff, uint64_t nReads,
                   const int32_t *minsc, const uint8_t *strandMask, const bt2g_scoring &sc, int maxHits, bt2g_mm_hit *hits,
                   int32_t *counts, cudaStream_t st) {
	const uint64_t n = nReads * 4;
	if(n) k_one_mm<OFF, false><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(ix, seq, qual, roff, nReads, minsc, strandMask, sc, maxHits, hits, counts, nullptr, nullptr);
}
// request-queue form: nSlots requests, slot s = read sel[s]
template <typename OFF>
void launch_one_mm_sel(const DevIndex<OFF> &ix, const uint8_t *seq, const uint8_t *qual, const uint64_t *roff, uint64_t nSlots, const uint32_t *sel,
                       const int32_t *minsc, const uint8_t *strandMask, const bt2g_scoring &sc, int maxHits, bt2g_mm_hit *hits,
                       int32_t *counts, cudaStream_t st, bool text) {
	const uint64_t n = nSlots * 4;
	if(!n) return;
	if(text) k_one_mm<OFF, true><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(ix, seq, qual, roff, nSlots, minsc, strandMask, sc, maxHits, hits, counts, sel, nullptr);
	else k_one_mm<OFF, false><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(ix, seq, qual, roff, nSlots, minsc, strandMask, sc, maxHits, hits, counts, sel, nullptr);
}
template void launch_one_mm_sel<uint32_t>(const DevIndex<uint32_t> &, const uint8_t *, const uint8_t *, const uint64_t *, uint64_t, const uint32_t *, const int32_t *, const uint8_t *, const bt2g_scoring &, int, bt2g_mm_hit *, int32_t *, cudaStream_t, bool);
template void launch_one_mm_sel<uint64_t>(const DevIndex<uint64_t> &, const uint8_t *, const uint8_t *, const uint64_t *, uint64_t, const uint32_t *, const int32_t *, const uint8_t *, const bt2g_scoring &, int, bt2g_mm_hit *, int32_t *, cudaStream_t, bool);
template void launch_one_mm<uint32_t>(const DevIndex<uint32_t> &, const uint8_t *, const uint8_t *, const uint64_t *, uint64_t, const int32_t *, const uint8_t *, const bt2g_scoring &, int, bt2g_mm_hit *, int32_t *, cudaStream_t);
template void launch_one_mm<uint64_t>(const DevIndex<uint64_t> &, const uint8_t *, const uint8_t *, const uint64_t *, uint64_t, const int32_t *, const uint8_t *, const bt2g_scoring &, int, bt2g_mm_hit *, int32_t *, cudaStream_t);
