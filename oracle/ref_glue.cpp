// oracle/ref_glue.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" surface over the UNMODIFIED reference classes (compiled from
// /root/reference by oracle/Makefile into oracle/_ref/libbt2ref_{s,l}.so).  This file is
// OUR code: it contains no reference source, it only #includes the reference's headers
// and calls its public member functions, so that tests can ask the real bowtie2
// implementation "what does Ebwt::countBt2SideEx / SeedAligner::searchAllSeeds /
// SwAligner::align return for these inputs?".  It is the "reference" arm of the oracle;
// oracle/bt2_oracle.c is the plain-C restatement that is checked against it.
//
// Nothing in the product (bowtie2_b200/, include/) may link or load this.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <memory>
#include <iostream>
#include <sstream>

#include "bt2_idx.h"
#include "reference.h"
#include "aligner_seed.h"
#include "aligner_cache.h"
#include "aligner_sw.h"
#include "aligner_sw_driver.h"
#include "dp_framer.h"
#include "scoring.h"
#include "simple_func.h"
#include "read.h"
#include "random_source.h"
#include "sstring.h"

extern "C" int bowtie(int argc, const char **argv);

namespace {

struct RefHandle {
	std::unique_ptr<Ebwt> fw;
	std::unique_ptr<Ebwt> bw;
	std::unique_ptr<BitPairReference> ref;
	// scoring: end-to-end defaults or local defaults (scoring.h:28-84, bt2_search.cpp:5040)
	std::unique_ptr<Scoring> sc_e2e;
	std::unique_ptr<Scoring> sc_loc;
};

Scoring* makeScoring(bool local) {
	SimpleFunc scoreMin, nCeil;
	if(local) scoreMin.init(SIMPLE_FUNC_LOG, DEFAULT_MIN_CONST_LOCAL, DEFAULT_MIN_LINEAR_LOCAL);
	else      scoreMin.init(SIMPLE_FUNC_LINEAR, DEFAULT_MIN_CONST, DEFAULT_MIN_LINEAR);
	nCeil.init(SIMPLE_FUNC_LINEAR, 0.0f, std::numeric_limits<double>::max(), 0.0f, 0.15f);
	return new Scoring(
		local ? DEFAULT_MATCH_BONUS_LOCAL : DEFAULT_MATCH_BONUS,
		DEFAULT_MM_PENALTY_TYPE, DEFAULT_MM_PENALTY_MAX, DEFAULT_MM_PENALTY_MIN,
		scoreMin, nCeil, DEFAULT_N_PENALTY_TYPE, DEFAULT_N_PENALTY, DEFAULT_N_CAT_PAIR,
		DEFAULT_READ_GAP_CONST, DEFAULT_REF_GAP_CONST,
		DEFAULT_READ_GAP_LINEAR, DEFAULT_REF_GAP_LINEAR, 4 /* gGapBarrier, bt2_search.cpp:354 */);
}

void fillRead(Read& rd, const uint8_t* codes, const uint8_t* quals, int len, const char* name) {
	static const char dna[] = "ACGTN";
	std::string s(len, 'N'), q(len, 'I');
	for(int i = 0; i < len; i++) {
		s[i] = dna[codes[i] > 4 ? 4 : codes[i]];
		if(quals != NULL) q[i] = (char)quals[i];
	}
	rd.init(name, s.c_str(), q.c_str());
}

} // namespace

extern "C" {

int ref_off_size(void) { return (int)sizeof(TIndexOffU); }

void* ref_open(const char* base, int load_mirror, int load_ref) {
	try {
		RefHandle* h = new RefHandle();
		std::string b(base);
		// same argument pattern as bt2_search.cpp:4986-5005 / 5155-5173
		h->fw.reset(new Ebwt(b, 0, -1, true, -1, 0, false, false, false,
		                     true, true, true, true, false, false, false, false));
		h->fw->loadIntoMemory(0, -1, true, true, true, true, false);
		if(load_mirror) {
			h->bw.reset(new Ebwt(b + ".rev", 0, 1, false, -1, 0, false, false, false,
			                     true, true, true, true, false, false, false, false));
			h->bw->loadIntoMemory(0, 1, false, true, false, true, false);
		}
		if(load_ref) {
			h->ref.reset(new BitPairReference(b, false, false, NULL, NULL, false,
			                                  false, false, false, false, false));
			if(!h->ref->loaded()) { delete h; return NULL; }
		}
		h->sc_e2e.reset(makeScoring(false));
		h->sc_loc.reset(makeScoring(true));
		return h;
	} catch(...) {
		return NULL;
	}
}

void ref_close(void* vh) { delete (RefHandle*)vh; }

// ---- header scalars -------------------------------------------------------------------
// which: 0 len, 1 bwtLen, 2 lineRate, 3 offRate, 4 ftabChars, 5 numSides, 6 sideSz,
//        7 sideBwtSz, 8 zOff, 9 nPat, 10 nFrag, 11 offsLen, 12 ftabLen, 13 eftabLen, 14 ebwtTotLen
uint64_t ref_scalar(void* vh, int mirror, int which) {
	RefHandle* h = (RefHandle*)vh;
	const Ebwt& e = mirror ? *h->bw : *h->fw;
	const EbwtParams& p = e.eh();
	switch(which) {
		case 0: return p.len();       case 1: return p.bwtLen();
		case 2: return p.lineRate();  case 3: return p.offRate();
		case 4: return p.ftabChars(); case 5: return p.numSides();
		case 6: return p.sideSz();    case 7: return p.sideBwtSz();
		case 8: return e.zOff();      case 9: return e.nPat();
		case 10: return e.nFrag();    case 11: return p.offsLen();
		case 12: return p.ftabLen();  case 13: return p.eftabLen();
		case 14: return p.ebwtTotLen();
	}
	return 0;
}

uint64_t ref_fchr(void* vh, int mirror, int c) {
	RefHandle* h = (RefHandle*)vh;
	return (mirror ? *h->bw : *h->fw).fchr()[c];
}

uint64_t ref_plen(void* vh, uint64_t t) { return ((RefHandle*)vh)->fw->plen()[t]; }

// ---- FM primitives ---------------------------------------------------------------------
// Ebwt::countBt2SideEx (bt2_idx.h:1887) via SideLocus::initFromRow (:369)
void ref_rank4(void* vh, int mirror, uint64_t row, uint64_t* out4) {
	RefHandle* h = (RefHandle*)vh;
	const Ebwt& e = mirror ? *h->bw : *h->fw;
	SideLocus l; l.initFromRow((TIndexOffU)row, e.eh(), e.ebwt());
	TIndexOffU a[4] = {0, 0, 0, 0};
	e.countBt2SideEx(l, a);
	for(int i = 0; i < 4; i++) out4[i] = a[i];
}

// Ebwt::mapLF(l, c) == countBt2Side (bt2_idx.h:2344,1758)
uint64_t ref_rank1(void* vh, int mirror, uint64_t row, int c) {
	RefHandle* h = (RefHandle*)vh;
	const Ebwt& e = mirror ? *h->bw : *h->fw;
	SideLocus l; l.initFromRow((TIndexOffU)row, e.eh(), e.ebwt());
	return e.mapLF(l, c);
}

int ref_rowL(void* vh, int mirror, uint64_t row) {
	RefHandle* h = (RefHandle*)vh;
	const Ebwt& e = mirror ? *h->bw : *h->fw;
	return e.rowL((TIndexOffU)row);
}

// Ebwt::mapLF1(row, l, c) (bt2_idx.h:2420); returns all-ones (as u64) on failure
uint64_t ref_maplf1(void* vh, int mirror, uint64_t row, int c) {
	RefHandle* h = (RefHandle*)vh;
	const Ebwt& e = mirror ? *h->bw : *h->fw;
	SideLocus l; l.initFromRow((TIndexOffU)row, e.eh(), e.ebwt());
	TIndexOffU r = e.mapLF1((TIndexOffU)row, l, c);
	return r == (TIndexOffU)OFF_MASK ? ~(uint64_t)0 : (uint64_t)r;
}

// Ebwt::mapLFRange(ltop, lbot, num, cntsUpto, cntsIn, masks) (bt2_idx.h:2268), called as GWState::advance does
// (group_walk.h:897); the four bool lists come back as one character per row
void ref_maplf_range(void* vh, int mirror, uint64_t top, uint64_t num, uint64_t* upto, uint64_t* in, uint8_t* chars) {
	RefHandle* h = (RefHandle*)vh;
	const Ebwt& e = mirror ? *h->bw : *h->fw;
	SideLocus tloc, bloc;
	SideLocus::initFromTopBot((TIndexOffU)top, (TIndexOffU)(top + num), e.eh(), e.ebwt(), tloc, bloc);
	TIndexOffU u[4] = {0, 0, 0, 0}, n[4] = {0, 0, 0, 0};
	EList<bool> masks[4];
	e.mapLFRange(tloc, bloc, (TIndexOffU)num, u, n, masks);
	for(int c = 0; c < 4; c++) { upto[c] = u[c]; in[c] = n[c]; }
	for(uint64_t j = 0; j < num; j++) {
		int cc = 255;
		for(int c = 0; c < 4; c++) if(masks[c][j]) cc = (cc == 255) ? c : 254;   // exactly one list may hold row j
		chars[j] = (uint8_t)cc;
	}
}

// Ebwt::ftabLoHi(i, top, bot) (bt2_idx.h:1476)
void ref_ftab_lohi(void* vh, int mirror, uint64_t i, uint64_t* top, uint64_t* bot) {
	RefHandle* h = (RefHandle*)vh;
	const Ebwt& e = mirror ? *h->bw : *h->fw;
	TIndexOffU t, b;
	e.ftabLoHi((TIndexOffU)i, t, b);
	*top = t; *bot = b;
}

// Ebwt::getOffset(row) (bt2_idx.cpp:150)
uint64_t ref_get_offset(void* vh, uint64_t row) {
	return ((RefHandle*)vh)->fw->getOffset((TIndexOffU)row);
}

// Ebwt::joinedToTextOff (bt2_idx.cpp:54). returns 0 if tidx==OFF_MASK (rejected straddle)
int ref_joined_to_text(void* vh, uint64_t qlen, uint64_t off, int rejectStraddle,
                       uint64_t* tidx, uint64_t* textoff, uint64_t* tlen, int* straddled) {
	RefHandle* h = (RefHandle*)vh;
	TIndexOffU ti = 0, to = 0, tl = 0; bool st = false;
	h->fw->joinedToTextOff((TIndexOffU)qlen, (TIndexOffU)off, ti, to, tl, rejectStraddle != 0, st);
	*straddled = st ? 1 : 0;
	if(ti == (TIndexOffU)OFF_MASK) { *tidx = ~(uint64_t)0; *textoff = 0; *tlen = 0; return 0; }
	*tidx = ti; *textoff = to; *tlen = tl;
	return 1;
}

// BitPairReference::getStretch (reference.cpp:420): codes 0..3, 4 = N. off may be negative /
// beyond the end only as far as the caller pads (we pad like SwAligner::initRef, aligner_sw.cpp:200-245)
int ref_get_stretch(void* vh, uint64_t tidx, int64_t off, int64_t count, uint8_t* out) {
	RefHandle* h = (RefHandle*)vh;
	const BitPairReference& r = *h->ref;
	int64_t tlen = (int64_t)r.approxLen(tidx);
	for(int64_t i = 0; i < count; i++) {
		int64_t p = off + i;
		if(p < 0 || p >= tlen) out[i] = 4;
		else out[i] = (uint8_t)r.getBase(tidx, (size_t)p);
	}
	return 0;
}

// ---- seed search -----------------------------------------------------------------------
// SeedAligner::exactSweep (aligner_seed.cpp:856), mineMax=2, repex=true as called from
// bt2_search.cpp:3514.  out: mine[2], and exact end-to-end ranges top/bot for fw and rc
// (0,0 if none).  returns nelt.
uint64_t ref_exact_sweep(void* vh, const uint8_t* codes, int len, int nofw, int norc,
                         uint64_t* mine2, uint64_t* topbot4) {
	RefHandle* h = (RefHandle*)vh;
	Read rd; fillRead(rd, codes, NULL, len, "r");
	SeedAligner al;
	SeedResults shs;
	SeedSearchMetrics sdm;
	shs.clear();
	shs.nextRead(rd);
	size_t mfw = 0, mrc = 0;
	size_t nelt = al.exactSweep(*h->fw, rd, *h->sc_e2e, nofw != 0, norc != 0, 2, mfw, mrc, true, shs, sdm);
	mine2[0] = mfw; mine2[1] = mrc;
	EEHit f = shs.exactFwEEHit(), r = shs.exactRcEEHit();
	topbot4[0] = f.top; topbot4[1] = f.bot; topbot4[2] = r.top; topbot4[3] = r.bot;
	return nelt;
}

// SeedAligner::instantiateSeeds + searchAllSeeds (aligner_seed.cpp:498,597) for exact seeds
// (multiseedMms = 0), as driven by bt2_search.cpp:3919-3965.
// out_ranges: [2 strands][max_seeds][4] = topf,botf,topb,botb (all zero if no hit).
// returns number of seed offsets (nseeds), or -1 if more than max_seeds.
int ref_seed_search(void* vh, const uint8_t* codes, const uint8_t* quals, int len,
                    int seedlen, int interval, int offset, int nofw, int norc,
                    int max_seeds, uint64_t* out_ranges) {
	RefHandle* h = (RefHandle*)vh;
	Read rd; fillRead(rd, codes, quals, len, "r");
	AlignmentCache scCurrent(8 * 1024 * 1024, false);
	AlignmentCacheIface ca(&scCurrent, NULL, NULL);
	ca.nextRead();
	SeedAligner al;
	SeedResults shs;
	SeedSearchMetrics sdm;
	PerReadMetrics prm;
	shs.clear();
	shs.nextRead(rd);
	EList<Seed> seeds;
	Constraint gc = Constraint::penaltyFuncBased(h->sc_e2e->scoreMin);
	Seed::mmSeeds(0, seedlen, seeds, gc);
	std::pair<int,int> instFw, instRc;
	std::pair<int,int> inst = al.instantiateSeeds(seeds, (size_t)offset, interval, rd, *h->sc_e2e,
	                                              nofw != 0, norc != 0, ca, shs, sdm, instFw, instRc);
	int nseeds = (int)shs.numOffs();
	if(nseeds > max_seeds) return -1;
	memset(out_ranges, 0, sizeof(uint64_t) * 2 * max_seeds * 4);
	if(inst.first + inst.second == 0) return nseeds;
	al.searchAllSeeds(seeds, h->fw.get(), h->bw.get(), rd, *h->sc_e2e,
	                  std::numeric_limits<size_t>::max(), ca, shs, sdm, prm);
	EList<SATuple, 16> satups;
	for(int fwi = 0; fwi < 2; fwi++) {
		bool fw = (fwi == 0);
		for(int i = 0; i < nseeds; i++) {
			const QVal& qv = shs.hitsAtOffIdx(fw, (size_t)i);
			if(!qv.valid() || qv.empty()) continue;
			satups.clear();
			size_t nrange = 0, nelt = 0;
			ca.queryQval(qv, satups, nrange, nelt);
			if(satups.size() == 0) continue;
			// exact seeds have exactly one reference substring per seed
			uint64_t* o = out_ranges + ((size_t)fwi * max_seeds + i) * 4;
			o[0] = satups[0].topf; o[1] = satups[0].topf + satups[0].size();
			o[2] = satups[0].topb; o[3] = satups[0].topb + satups[0].size();
		}
	}
	return nseeds;
}

// SeedAligner::oneMmSearch (aligner_seed.cpp:975-1325) as called at bt2_search.cpp:3709.
// out: per hit 6 x int64: top, bot, pos, chr, qchr, score (in the order the reference stored
// them in SeedResults::mm1Hit_), plus fw flag in out_fw.  returns the number of hits.
int ref_one_mm(void* vh, int local, const uint8_t* codes, const uint8_t* quals, int len, int64_t minsc,
               int nofw, int norc, int max_hits, int64_t* out, int* out_fw) {
	RefHandle* h = (RefHandle*)vh;
	const Scoring& sc = local ? *h->sc_loc : *h->sc_e2e;
	Read rd; fillRead(rd, codes, quals, len, "r");
	SeedAligner al;
	SeedResults shs;
	SeedSearchMetrics sdm;
	shs.clear();
	shs.nextRead(rd);
	al.oneMmSearch(h->fw.get(), h->bw.get(), rd, sc, minsc, nofw != 0, norc != 0, local != 0, false, true, shs, sdm);
	const EList<EEHit>& hs = shs.mm1EEHits();
	int n = 0;
	for(size_t i = 0; i < hs.size(); i++) {
		if(n < max_hits) {
			int64_t* o = out + 6 * n;
			o[0] = hs[i].top; o[1] = hs[i].bot; o[2] = hs[i].e1.pos; o[3] = hs[i].e1.chr; o[4] = hs[i].e1.qchr; o[5] = hs[i].score;
			out_fw[n] = hs[i].fw ? 1 : 0;
		}
		n++;
	}
	return n;
}

// whole program entry, for completeness (bt2_search.cpp:5230)
int ref_bowtie_main(int argc, const char** argv) { return bowtie(argc, argv); }

} // extern "C"
