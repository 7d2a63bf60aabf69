// dp_ungapped_device.cuh -- SwAligner::ungappedAlign (aligner_sw.cpp:286-487) as a device function
#pragma once
#include "fm_device.cuh"

// SwAligner::ungappedAlign for one problem (the body of k_ungapped; also called inline by the exact engine, csrc/xengine.cu).
// mask (optional): one byte per read row, 1 where the row carries an edit.
template <typename OFF>
__device__ __forceinline__ void ungapped_one(const DevIndex<OFF> &ix, const bt2g_scoring &sc, const uint8_t *rs, const uint8_t *rq, int len,
                                             const bt2g_ungapped_problem &p, bt2g_ungapped_result &r, uint8_t *m, uint32_t stride) {
	r.status = 0; r.score = 0; r.rowi = 0; r.rowf = 0; r.ns = 0; r.refns = 0; r.nedits = 0; r.pad = 0;
	if(len <= 0) return;
	const int nceil = (int)((double)sc.nceil_const + (double)sc.nceil_linear * (double)len);
	const int64_t rfi = p.refoff, rff = p.refoff + len, reflen = (int64_t)p.reflen;
	int64_t leftNs = 0, rightNs = 0;
	if(rfi < 0) { if(p.ohang) leftNs = -rfi; else return; }
	if(rff > reflen) { if(p.ohang) rightNs = rff - reflen; else return; }
	if(leftNs + rightNs > nceil) return;
	auto rdc = [&](int i) -> int { int c = p.fw ? rs[i] : rs[len - 1 - i]; return p.fw ? c : (c > 3 ? 4 : 3 - c); };
	auto qv = [&](int i) -> int { int q = (int)(p.fw ? rq[i] : rq[len - 1 - i]) - 33; return q < 0 ? 0 : (q > 63 ? 63 : q); };
	RefCursor<OFF> cur;
	auto rfc = [&](int i) -> int { return cur.get(ix, p.tidx, p.refoff + i); };      // off-end positions read as N
	auto cellsc = [&](int c, int f, int q) -> int { return (c > 3 || f > 3) ? -(int)sc.npen[q] : (c == f ? sc.match_bonus : -(int)sc.mmpen[q]); };
	int64_t score = 0;
	int ns = 0, rowi = 0, rowf = len - 1, rc = 1;
	if(sc.match_bonus == 0) {
		for(int i = 0; i < len; i++) {
			const int c = rdc(i), f = rfc(i);
			ns += (c > 3 || f > 3);
			score += cellsc(c, f, qv(i));
		}
		if(score < p.minsc || ns > nceil) rc = 0;
	} else {
		int64_t scoreMax = 0;
		int lastfloor = 0, sols = 0;
		rowi = -1;
		for(int i = 0; i < len; i++) {
			const int c = rdc(i), f = rfc(i);
			ns += (c > 3 || f > 3);
			score += cellsc(c, f, qv(i));
			if(score >= p.minsc && score >= scoreMax) {
				scoreMax = score; rowf = i;
				if(rowi != lastfloor) { rowi = lastfloor; sols++; }
			}
			if(score <= 0) { score = 0; lastfloor = i + 1; }
		}
		if(ns > nceil || scoreMax < p.minsc) rc = 0;
		else if(sols > 1) rc = -1;
		score = scoreMax;
	}
	r.status = rc;
	if(rc == 1) {
		int refns = 0, ned = 0;
		for(int i = 0; i < len && m && i < (int)stride; i++) m[i] = 0;
		for(int i = rowi; i <= rowf; i++) {
			const int f = rfc(i);
			if(f > 3 || rdc(i) != f) { ned++; refns += f > 3; if(m && i < (int)stride) m[i] = 1; }
		}
		r.score = (int32_t)score; r.rowi = rowi; r.rowf = rowf; r.ns = ns; r.refns = refns; r.nedits = ned;
	}
}

