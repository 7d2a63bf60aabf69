// pe_device.cuh -- paired-end window / rectangle arithmetic, usable from kernels and from host code.
//   pe_other_mate   : PairedEndPolicy::otherMate      (pe.cpp:161-355, pePolicyMateDir pe.h:130-164)
//   pe_classify     : PairedEndPolicy::peClassifyPair (pe.cpp:37-137)
//   frame_mate_rect : DynProgFramer::frameFindMateAnchor{Left,Right}Rect (dp_framer.cpp:177-361)
#pragma once
#include "../../include/bt2g.h"

#define BT2G_HD __host__ __device__ __forceinline__

BT2G_HD int64_t pe_min64(int64_t a, int64_t b) { return a < b ? a : b; }
BT2G_HD int64_t pe_max64(int64_t a, int64_t b) { return a > b ? a : b; }

BT2G_HD bool pe_other_mate(const bt2g_pe_policy &pp, bool is1, bool fw, int64_t off, int64_t maxalcols,
                           uint64_t len1, uint64_t len2, bool &oleft, int64_t &oll, int64_t &olr, int64_t &orl,
                           int64_t &orr, bool &ofw) {
	switch(pp.pol) {
		case 1: oleft = (is1 != fw); ofw = fw; break;     // FF
		case 2: oleft = (is1 == fw); ofw = fw; break;     // RR
		case 3: oleft = !fw; ofw = !fw; break;            // FR
		default: oleft = fw; ofw = !fw; break;            // RF
	}
	const int64_t alen = (int64_t)(is1 ? len1 : len2);
	uint64_t maxfragU = pp.maxfrag, minfragU = pp.minfrag < 1 ? 1 : pp.minfrag;
	const bool expand = (pp.flags & BT2G_PE_EXPAND_TO_FIT) != 0;
	if(len1 > maxfragU && expand) maxfragU = len1;
	if(len2 > maxfragU && expand) maxfragU = len2;
	if(!expand && (len1 > maxfragU || len2 > maxfragU)) return false;
	const int64_t maxfrag = (int64_t)maxfragU, minfrag = (int64_t)minfragU;
	const bool olapOk = (pp.flags & BT2G_PE_OLAP_OK) != 0, dovetailOk = (pp.flags & BT2G_PE_DOVETAIL_OK) != 0;
	const bool flippingOk = (pp.flags & BT2G_PE_FLIPPING_OK) != 0;
	if(oleft) {
		oll = off + alen - maxfrag;
		olr = off + alen - minfrag;
		orl = oll;
		orr = off + maxfrag - 1;
		if(!olapOk) {
			orr = pe_min64(orr, off - 1);
			if(orr < olr) olr = orr;
		} else if(!dovetailOk) {
			orr = pe_min64(orr, off + alen - 1);
		} else if(!flippingOk && maxalcols != -1) {
			orr = pe_min64(orr, off + alen - 1 + (maxalcols - 1));
		}
	} else {
		orr = off + (maxfrag - 1);
		orl = off + (minfrag - 1);
		oll = off + alen - maxfrag;
		olr = orr;
		if(!olapOk) {
			oll = pe_max64(oll, off + alen);
			if(oll > orl) orl = oll;
		} else if(!dovetailOk) {
			oll = pe_max64(oll, off);
		} else if(!flippingOk && maxalcols != -1) {
			oll = pe_max64(oll, off - maxalcols + 1);
		}
	}
	return true;
}

BT2G_HD int pe_classify(const bt2g_pe_policy &pp, int64_t off1, uint64_t len1, bool fw1, int64_t off2, uint64_t len2, bool fw2) {
	uint64_t maxfrag = pp.maxfrag;
	const bool expand = (pp.flags & BT2G_PE_EXPAND_TO_FIT) != 0;
	if(len1 > maxfrag && expand) maxfrag = len1;
	if(len2 > maxfrag && expand) maxfrag = len2;
	const uint64_t minfrag = pp.minfrag < 1 ? 1 : pp.minfrag;
	bool oneLeft;
	if(pp.pol == 1 || pp.pol == 2) {
		if(fw1 != fw2) return 5;
		oneLeft = pp.pol == 1 ? fw1 : !fw1;
	} else {
		if(fw1 == fw2) return 5;
		oneLeft = pp.pol == 3 ? fw1 : !fw1;
	}
	const int64_t fraglo = pe_min64(off1, off2);
	const int64_t fraghi = pe_max64(off1 + (int64_t)len1, off2 + (int64_t)len2);
	const uint64_t frag = (uint64_t)(fraghi - fraglo);
	if(frag > maxfrag || frag < minfrag) return 5;
	const int64_t lo1 = off1, hi1 = off1 + (int64_t)len1 - 1, lo2 = off2, hi2 = off2 + (int64_t)len2 - 1;
	const bool containment = (lo1 >= lo2 && hi1 <= hi2) || (lo2 >= lo1 && hi2 <= hi1);
	int type = 1;
	bool olap = false;
	if((lo1 <= lo2 && hi1 >= lo2) || (lo1 <= hi2 && hi1 >= hi2) || containment) {
		olap = true;
		if(!(pp.flags & BT2G_PE_OLAP_OK)) return 5;
		type = 2;
	}
	if(!olap) {
		if((oneLeft && lo2 < lo1) || (!oneLeft && lo1 < lo2)) return 5;
	}
	if(containment) {
		if(!(pp.flags & BT2G_PE_CONTAIN_OK)) return 5;
		type = 3;
	}
	if((oneLeft && (hi1 > hi2 || lo2 < lo1)) || (!oneLeft && (hi2 > hi1 || lo1 < lo2))) {
		if(!(pp.flags & BT2G_PE_DOVETAIL_OK)) return 5;
		type = 4;
	}
	return type;
}

// trimToRef = true (gReportOverhangs = false): maxns is zeroed (dp_framer.cpp:222-226)
BT2G_HD bool frame_mate_rect(bool anchorLeft, int64_t ll, int64_t lr, int64_t rl, int64_t rr, int64_t rdlen, int64_t reflen,
                             uint64_t maxrdgap, uint64_t maxrfgap, uint64_t maxhalf, bt2g_mate_frame &f) {
	uint64_t maxgapU = maxrdgap > maxrfgap ? maxrdgap : maxrfgap;
	if(maxhalf > maxgapU) maxgapU = maxhalf;
	const int64_t maxgap = (int64_t)maxgapU;
	int64_t refl, refr;
	if(anchorLeft) { refl = (rl - (rdlen - 1)) - maxgap; refr = rr + maxgap; }
	else           { refl = ll - maxgap; refr = (lr + (rdlen - 1)) + maxgap; }
	(void)lr; (void)ll;
	int64_t triml = 0, trimr = 0;
	if(refr >= reflen) trimr = refr - (reflen - 1);
	if(refl < 0) triml = -refl;
	const int64_t width = refr - refl + 1;
	f.refl_pretrim = refl; f.refr_pretrim = refr;
	f.refl = refl + triml; f.refr = refr - trimr;
	f.triml = triml; f.trimr = trimr;
	f.maxgap = maxgap; f.corel = maxgap; f.corer = width - maxgap - 1;
	return f.refr >= f.refl;
}

BT2G_HD void pe_frame_anchor(const bt2g_pe_policy &pp, const bt2g_mate_anchor &a, bt2g_mate_frame &f) {
	bool oleft = false, ofw = false;
	int64_t oll = 0, olr = 0, orl = 0, orr = 0;
	f.status = 0; f.oleft = 0; f.ofw = 0; f.pad[0] = f.pad[1] = 0;
	f.oll = f.olr = f.orl = f.orr = 0;
	f.refl = f.refr = f.refl_pretrim = f.refr_pretrim = f.triml = f.trimr = f.corel = f.corer = f.maxgap = 0;
	if(!pe_other_mate(pp, a.is1 != 0, a.fw != 0, a.off, a.maxalcols, a.len1, a.len2, oleft, oll, olr, orl, orr, ofw)) return;
	f.oleft = oleft; f.ofw = ofw; f.oll = oll; f.olr = olr; f.orl = orl; f.orr = orr;
	const int64_t orows = a.is1 ? a.len2 : a.len1;
	// the reference passes the gap counts as size_t: negative values wrap (dp_framer.cpp:197-198)
	const bool ok = frame_mate_rect(!oleft, oll, olr, orl, orr, orows, (int64_t)a.reflen, (uint64_t)(int64_t)a.maxrdgap,
	                                (uint64_t)(int64_t)a.maxrfgap, (uint64_t)(int64_t)a.maxhalf, f);
	f.status = ok ? 2 : 1;
}
