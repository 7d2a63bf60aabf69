// mapq_device.cuh -- BowtieMapq2::mapq (unique.h:170-392), the reference's default MAPQ model, for a primary
// alignment.  best / secbest: the read's best and best-unchosen scores (pairs: the concordant sums), scMin /
// scPer: the minimum valid and the perfect score (pairs: summed over both mates).  The thresholds are float
// literals widened to double exactly as in the reference ("diff * (double)0.8f").
#pragma once

__host__ __device__ __forceinline__ int mapq_v2(long long best, bool hasSec, long long secbest, long long scMin, long long scPer, bool monotone) {
	long long diff = scPer - scMin; if(diff < 1) diff = 1;
	const double d = (double)diff;
	const long long bestOver = best - scMin;
	const double bo = (double)bestOver;
	if(!hasSec) {
		if(monotone) {
			if(bo >= d * (double)0.8f) return 42;
			if(bo >= d * (double)0.7f) return 40;
			if(bo >= d * (double)0.6f) return 24;
			if(bo >= d * (double)0.5f) return 23;
			if(bo >= d * (double)0.4f) return 8;
			if(bo >= d * (double)0.3f) return 3;
			return 0;
		}
		if(bo >= d * (double)0.8f) return 44;
		if(bo >= d * (double)0.7f) return 42;
		if(bo >= d * (double)0.6f) return 41;
		if(bo >= d * (double)0.5f) return 36;
		if(bo >= d * (double)0.4f) return 28;
		if(bo >= d * (double)0.3f) return 24;
		return 22;
	}
	long long ab = best < 0 ? -best : best, as = secbest < 0 ? -secbest : secbest;
	long long bd = ab - as; if(bd < 0) bd = -bd;
	const double bdd = (double)bd;
	const bool top = bestOver == diff;
	if(monotone) {
		if(bdd >= d * (double)0.9f) return top ? 39 : 33;
		if(bdd >= d * (double)0.8f) return top ? 38 : 27;
		if(bdd >= d * (double)0.7f) return top ? 37 : 26;
		if(bdd >= d * (double)0.6f) return top ? 36 : 22;
		if(bdd >= d * (double)0.5f) return top ? 35 : (bo >= d * (double)0.84f ? 25 : (bo >= d * (double)0.68f ? 16 : 5));
		if(bdd >= d * (double)0.4f) return top ? 34 : (bo >= d * (double)0.84f ? 21 : (bo >= d * (double)0.68f ? 14 : 4));
		if(bdd >= d * (double)0.3f) return top ? 32 : (bo >= d * (double)0.88f ? 18 : (bo >= d * (double)0.67f ? 15 : 3));
		if(bdd >= d * (double)0.2f) return top ? 31 : (bo >= d * (double)0.88f ? 17 : (bo >= d * (double)0.67f ? 11 : 0));
		if(bdd >= d * (double)0.1f) return top ? 30 : (bo >= d * (double)0.88f ? 12 : (bo >= d * (double)0.67f ? 7 : 0));
		if(bd > 0) return bo >= d * (double)0.67f ? 6 : 2;
		return bo >= d * (double)0.67f ? 1 : 0;
	}
	if(bdd >= d * (double)0.9f) return 40;
	if(bdd >= d * (double)0.8f) return 39;
	if(bdd >= d * (double)0.7f) return 38;
	if(bdd >= d * (double)0.6f) return 37;
	if(bdd >= d * (double)0.5f) return top ? 35 : (bo >= d * (double)0.5f ? 25 : 20);
	if(bdd >= d * (double)0.4f) return top ? 34 : (bo >= d * (double)0.5f ? 21 : 19);
	if(bdd >= d * (double)0.3f) return top ? 33 : (bo >= d * (double)0.5f ? 18 : 16);
	if(bdd >= d * (double)0.2f) return top ? 32 : (bo >= d * (double)0.5f ? 17 : 12);
	if(bdd >= d * (double)0.1f) return top ? 31 : (bo >= d * (double)0.5f ? 14 : 9);
	if(bd > 0) return bo >= d * (double)0.5f ? 11 : 2;
	return bo >= d * (double)0.5f ? 1 : 0;
}
