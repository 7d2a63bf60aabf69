// oracle/ref_glue_dp.cpp -- TEST INFRASTRUCTURE ONLY (see ref_glue.cpp).
//
// Runs ONE dynamic-programming problem through the unmodified reference SwAligner exactly as
// SwDriver::extendSeeds does (aligner_sw_driver.cpp:1272-1376): DynProgFramer::frameSeedExtensionRect,
// SwAligner::initRead / initRef / align, then nextAlignment until exhausted.  Returns the
// candidate list (btncand_, aligner_sw.h:628) and every alignment produced, so the CUDA DP
// kernel can be checked cell-for-cell on candidates and edit-for-edit on backtraces.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <memory>
#include <iostream>
#include <sstream>
#include <algorithm>
#include <limits>
#include <map>
#include <set>
#include <fstream>
#include <thread>
#include <mutex>
#include <atomic>
#include <array>
#include <utility>
#include <stdexcept>

#define private public
#define protected public
#include "aligner_sw.h"
#include "pe.h"
#include "aligner_seed.h"
#include "random_source.h"
#include "unique.h"
#undef private
#undef protected
#include "bt2_idx.h"
#include "reference.h"
#include "dp_framer.h"
#include "scoring.h"
#include "simple_func.h"
#include "read.h"
#include "random_source.h"

struct RefHandleDp {   // must mirror RefHandle in ref_glue.cpp
	std::unique_ptr<Ebwt> fw;
	std::unique_ptr<Ebwt> bw;
	std::unique_ptr<BitPairReference> ref;
	std::unique_ptr<Scoring> sc_e2e;
	std::unique_ptr<Scoring> sc_loc;
};

extern "C" {

// DynProgFramer::frameSeedExtensionRect (dp_framer.cpp:81-129) with trimToRef = true
// (gReportOverhangs = false).  out9: refl, refr, refl_pretrim, refr_pretrim, triml, trimr,
// corel, corer, maxgap.  returns found.
int ref_frame_seed_rect(int64_t off, uint64_t rdlen, int64_t reflen, uint64_t maxrdgap, uint64_t maxrfgap,
                        int64_t maxns, uint64_t maxhalf, int64_t* out9) {
	DynProgFramer fr(true);
	DPRect r;
	bool found = fr.frameSeedExtensionRect(off, rdlen, reflen, maxrdgap, maxrfgap, maxns, maxhalf, r);
	out9[0] = r.refl; out9[1] = r.refr; out9[2] = r.refl_pretrim; out9[3] = r.refr_pretrim;
	out9[4] = (int64_t)r.triml; out9[5] = (int64_t)r.trimr; out9[6] = (int64_t)r.corel; out9[7] = (int64_t)r.corer;
	out9[8] = (int64_t)r.maxgap;
	return found ? 1 : 0;
}

// RandomSource (random_source.h:32-180): n draws after init(seed).  kind 0 nextU32, 1 nextU2, 2 nextBool,
// 3 nextU32 % mod (mod in arg), 4 nextFloat bits
void ref_rng_draws(uint32_t seed, const int32_t* kinds, const uint32_t* args, int n, uint32_t* out) {
	RandomSource r; r.init(seed);
	for(int i = 0; i < n; i++) {
		switch(kinds[i]) {
			case 0: out[i] = r.nextU32(); break;
			case 1: out[i] = r.nextU2(); break;
			case 2: out[i] = r.nextBool() ? 1u : 0u; break;
			case 3: out[i] = r.nextU32() % args[i]; break;
			default: { float f = r.nextFloat(); uint32_t u; memcpy(&u, &f, 4); out[i] = u; }
		}
	}
}

// SeedResults::rankSeedHits (aligner_seed.h:1019-1080) for given element counts per (strand, offset index);
// 0 = no hit.  out_off / out_fw receive the ranking.  Returns its length.
int ref_rank_seed_hits(uint32_t rndseed, int num_offs, const uint32_t* nelt_fw, const uint32_t* nelt_rc, int all,
                       uint32_t* out_off, int* out_fw) {
	SeedResults sr;
	Read rd; rd.init("r", "ACGT", "IIII");
	EList<uint32_t> off2; for(int i = 0; i < num_offs; i++) off2.push_back((uint32_t)i);
	sr.reset(rd, off2, (size_t)num_offs);
	for(int i = 0; i < num_offs; i++) {
		if(nelt_fw[i]) { sr.hitsFw_[i].init(0, 1, nelt_fw[i]); sr.nonzTot_++; }
		if(nelt_rc[i]) { sr.hitsRc_[i].init(0, 1, nelt_rc[i]); sr.nonzTot_++; }
	}
	RandomSource rnd; rnd.init(rndseed);
	sr.rankSeedHits(rnd, all != 0);
	int n = (int)sr.rankOffs_.size();
	for(int i = 0; i < n; i++) { out_off[i] = sr.rankOffs_[i]; out_fw[i] = sr.rankFws_[i] ? 1 : 0; }
	return n;
}

// BowtieMapq2::mapq (unique.h:170-392), the default MAPQ model.  Unpaired: best / secbest are the read's best and
// best-unchosen alignment scores; paired (ordlen > 0): the concordant-pair sums.  The summary object is filled
// directly (this glue is built with private members exposed, as for SwAligner).
int ref_mapq_v2(void* vh, int local, int64_t rdlen, int64_t ordlen, int64_t best, int has_sec, int64_t secbest) {
	RefHandleDp* h = (RefHandleDp*)vh;
	const Scoring& sc = local ? *h->sc_loc : *h->sc_e2e;
	SimpleFunc scoreMin;
	if(local) scoreMin.init(SIMPLE_FUNC_LOG, DEFAULT_MIN_CONST_LOCAL, DEFAULT_MIN_LINEAR_LOCAL);
	else      scoreMin.init(SIMPLE_FUNC_LINEAR, DEFAULT_MIN_CONST, DEFAULT_MIN_LINEAR);
	BowtieMapq2 mq(scoreMin, sc);
	AlnSetSumm s;
	s.reset();
	const bool paired = ordlen > 0;
	s.paired_ = paired;
	s.exhausted1_ = s.exhausted2_ = true;
	AlnScore b; b.score_ = best; b.ns_ = 0; b.gaps_ = 0; b.basesAligned_ = 0; b.edits_ = 0;
	AlnScore u; u.score_ = secbest; u.ns_ = 0; u.gaps_ = 0; u.basesAligned_ = 0; u.edits_ = 0;
	if(paired) { s.bestCScore_ = b; s.bestP1Score_ = b; if(has_sec) s.bestUnchosenCScore_ = u; }
	else       { s.bestUScore_ = b; if(has_sec) s.bestUnchosenUScore_ = u; }
	AlnFlags flags(paired ? ALN_FLAG_PAIR_CONCORD_MATE1 : ALN_FLAG_PAIR_UNPAIRED, true /*canMax*/, false, false, false, false, false, false,
	               false, true /*primary*/, paired, false, false, false);
	return (int)mq.mapq(s, flags, true, (size_t)rdlen, (size_t)ordlen, NULL);
}

// SwAligner::ungappedAlign (aligner_sw.cpp:286-487).  out8: score, refoff, trim5, trim3, ns, refns, nedits, -;
// edits (pos, chr, qchr, type) as for ref_dp.  returns the function's return value (0, -1, 1).
int ref_ungapped(void* vh, int local, const uint8_t* codes, const uint8_t* quals, int len, int fw,
                 uint64_t tidx, int64_t off, int64_t tlen, int ohang, int64_t minsc, int max_edits,
                 int64_t* out8, int32_t* edits) {
	RefHandleDp* h = (RefHandleDp*)vh;
	const Scoring& sc = local ? *h->sc_loc : *h->sc_e2e;
	static const char dna[] = "ACGTN";
	std::string s(len, 'N'), q(len, 'I');
	for(int i = 0; i < len; i++) { s[i] = dna[codes[i] > 4 ? 4 : codes[i]]; if(quals) q[i] = (char)quals[i]; }
	Read rd; rd.init("r", s.c_str(), q.c_str());
	SwAligner sw(NULL);
	SwResult res;
	Coord coord((TRefId)tidx, (TRefOff)off, fw != 0);
	for(int i = 0; i < 8; i++) out8[i] = 0;
	int al = sw.ungappedAlign(fw ? rd.patFw : rd.patRc, fw ? rd.qual : rd.qualRev, coord, *h->ref, (size_t)tlen, sc,
	                          ohang != 0, (TAlScore)minsc, res);
	if(al == 1) {
		const AlnRes& a = res.alres;
		out8[0] = a.score().score(); out8[1] = a.refoff();
		out8[2] = (int64_t)a.trimmed5p(true); out8[3] = (int64_t)a.trimmed3p(true);
		out8[4] = a.score().ns(); out8[5] = (int64_t)a.refNs();
		const EList<Edit>& ned = a.ned();
		out8[6] = (int64_t)ned.size();
		for(size_t k = 0; k < ned.size() && (int)k < max_edits; k++) {
			edits[4 * k] = (int32_t)ned[k].pos; edits[4 * k + 1] = ned[k].chr; edits[4 * k + 2] = ned[k].qchr; edits[4 * k + 3] = ned[k].type;
		}
	}
	return al;
}

// PairedEndPolicy::otherMate (pe.cpp:161-355) + DynProgFramer::frameFindMateRect (dp_framer.h:155-197,
// dp_framer.cpp:177-361) as chained in SwDriver::extendSeedsPaired (aligner_sw_driver.cpp:2226-2256).
// pol: PE_POLICY_FF=1, RR=2, FR=3, RF=4; flags bit0 flippingOk, bit1 dovetailOk, bit2 containOk,
// bit3 olapOk, bit4 expandToFit, bit5 local.  out15: oleft, ofw, oll, olr, orl, orr, then rect9.
// returns 0 none, 1 otherMate ok but rectangle entirely trimmed, 2 found.
int ref_frame_mate(int pol, uint64_t maxfrag, uint64_t minfrag, int flags,
                   int is1, int fw, int64_t off, int64_t maxalcols, uint64_t reflen, uint64_t len1, uint64_t len2,
                   uint64_t maxrdgap, uint64_t maxrfgap, int64_t maxns, uint64_t maxhalf, int64_t* out15) {
	PairedEndPolicy pe(pol, maxfrag, minfrag, (flags & 32) != 0, (flags & 1) != 0, (flags & 2) != 0,
	                   (flags & 4) != 0, (flags & 8) != 0, (flags & 16) != 0);
	bool oleft = false, ofw = false;
	int64_t oll = 0, olr = 0, orl = 0, orr = 0;
	for(int i = 0; i < 15; i++) out15[i] = 0;
	if(!pe.otherMate(is1 != 0, fw != 0, off, maxalcols, reflen, len1, len2, oleft, oll, olr, orl, orr, ofw)) return 0;
	out15[0] = oleft; out15[1] = ofw; out15[2] = oll; out15[3] = olr; out15[4] = orl; out15[5] = orr;
	DynProgFramer fr(true);
	DPRect r;
	const uint64_t orows = is1 ? len2 : len1;
	bool found = fr.frameFindMateRect(!oleft, oll, olr, orl, orr, orows, (int64_t)reflen, maxrdgap, maxrfgap, maxns, maxhalf, r);
	int64_t* o = out15 + 6;
	o[0] = r.refl; o[1] = r.refr; o[2] = r.refl_pretrim; o[3] = r.refr_pretrim;
	o[4] = (int64_t)r.triml; o[5] = (int64_t)r.trimr; o[6] = (int64_t)r.corel; o[7] = (int64_t)r.corer;
	o[8] = (int64_t)r.maxgap;
	return found ? 2 : 1;
}

// PairedEndPolicy::peClassifyPair (pe.cpp:37-137): PE_ALS_NORMAL=1, OVERLAP, CONTAIN, DOVETAIL, DISCORD=5
int ref_pe_classify(int pol, uint64_t maxfrag, uint64_t minfrag, int flags,
                    int64_t off1, uint64_t len1, int fw1, int64_t off2, uint64_t len2, int fw2) {
	PairedEndPolicy pe(pol, maxfrag, minfrag, (flags & 32) != 0, (flags & 1) != 0, (flags & 2) != 0,
	                   (flags & 4) != 0, (flags & 8) != 0, (flags & 16) != 0);
	return pe.peClassifyPair(off1, len1, fw1 != 0, off2, len2, fw2 != 0);
}

// Scoring::maxReadGaps / maxRefGaps (scoring.cpp:42,73), scoreMin, nCeil, perfectScore
void ref_score_params(void* vh, int local, int64_t minsc, uint64_t rdlen, int64_t* out4) {
	RefHandleDp* h = (RefHandleDp*)vh;
	const Scoring& sc = local ? *h->sc_loc : *h->sc_e2e;
	out4[0] = sc.maxReadGaps(minsc, rdlen);
	out4[1] = sc.maxRefGaps(minsc, rdlen);
	out4[2] = sc.perfectScore(rdlen);
	out4[3] = sc.nCeil.f<int>((double)rdlen);
}

// One DP problem.  rect9 as above.  Outputs:
//   summary[0]=found (align() return), [1]=best, [2]=ncand, [3]=naln
//   cands[3*i+{0,1,2}] = row, col, score     (sorted as btncand_ is after align())
//   alns: per alignment 8 int64: score, ns, gaps, refoff, trim5(soft), trim3(soft), nedits, fw
//   edits: per edit 4 int32: pos, chr, qchr, type; concatenated, alignment by alignment
int ref_dp(void* vh, int local, const uint8_t* codes, const uint8_t* quals, int len, int fw,
           uint64_t tidx, int64_t tlen, const int64_t* rect9, int64_t minsc, uint32_t rndseed,
           int max_cands, int max_alns, int max_edits,
           int64_t* summary, int64_t* cands, int64_t* alns, int32_t* edits) {
	RefHandleDp* h = (RefHandleDp*)vh;
	const Scoring& sc = local ? *h->sc_loc : *h->sc_e2e;
	static const char dna[] = "ACGTN";
	std::string s(len, 'N'), q(len, 'I');
	for(int i = 0; i < len; i++) { s[i] = dna[codes[i] > 4 ? 4 : codes[i]]; if(quals) q[i] = (char)quals[i]; }
	Read rd; rd.init("r", s.c_str(), q.c_str());
	DPRect rect;
	rect.refl = rect9[0]; rect.refr = rect9[1]; rect.refl_pretrim = rect9[2]; rect.refr_pretrim = rect9[3];
	rect.triml = (size_t)rect9[4]; rect.trimr = (size_t)rect9[5]; rect.corel = (size_t)rect9[6];
	rect.corer = (size_t)rect9[7]; rect.maxgap = (size_t)rect9[8];
	SwAligner sw(NULL);
	RandomSource rnd; rnd.init(rndseed);
	sw.initRead(rd.patFw, rd.patRc, rd.qual, rd.qualRev, 0, rd.length(), sc);
	size_t nsLeft = 0;
	sw.initRef(fw != 0, (TRefId)tidx, rect, *h->ref, (TRefOff)tlen, sc, (TAlScore)minsc,
	           true /*enable8*/, 2000 /*cminlen*/, 4 /*cpow2*/, false /*doTri*/, true /*extend*/, 0, nsLeft);
	TAlScore best = std::numeric_limits<TAlScore>::min();
	bool found = sw.align(best);
	summary[0] = found ? 1 : 0; summary[1] = best; summary[2] = 0; summary[3] = 0;
	if(!found) return 0;
	int nc = (int)sw.btncand_.size();
	summary[2] = nc;
	for(int i = 0; i < nc && i < max_cands; i++) {
		cands[3 * i] = (int64_t)sw.btncand_[i].row; cands[3 * i + 1] = (int64_t)sw.btncand_[i].col;
		cands[3 * i + 2] = sw.btncand_[i].score;
	}
	int naln = 0, ned = 0;
	SwResult res;
	while(!sw.done()) {
		res.reset();
		sw.nextAlignment(res, (TAlScore)minsc, rnd);
		if(res.empty()) break;
		if(naln < max_alns) {
			const AlnRes& a = res.alres;
			int64_t* o = alns + 8 * naln;
			o[0] = a.score().score(); o[1] = a.score().ns(); o[2] = a.score().gaps();
			o[3] = a.refoff(); o[4] = (int64_t)a.trimmed5p(true); o[5] = (int64_t)a.trimmed3p(true);
			o[6] = (int64_t)a.ned().size(); o[7] = a.fw() ? 1 : 0;
			for(size_t k = 0; k < a.ned().size(); k++) {
				if(ned < max_edits) {
					edits[4 * ned] = (int32_t)a.ned()[k].pos; edits[4 * ned + 1] = a.ned()[k].chr;
					edits[4 * ned + 2] = a.ned()[k].qchr; edits[4 * ned + 3] = a.ned()[k].type;
				}
				ned++;
			}
		}
		naln++;
	}
	summary[3] = naln;
	return ned;
}

} // extern "C"
