// dp_device.cuh -- launch descriptor shared by api.cu and dp_kernels.cu (DP types: include/bt2g.h)
#pragma once
#include "bt2g_internal.h"
#include <cstdlib>

// the s16x2 kernel needs every reachable score within +-DPX_LIMIT (dp_kernels.cu): end-to-end mode,
// minimum score >= -8000 and perfect score <= 8000; BT2G_DP_PACKED=0 in the environment disables it
static inline bool dp_packed_ok(const bt2g_scoring &sc, int64_t minMinsc, int maxLen) {
	const char *e = getenv("BT2G_DP_PACKED");
	if(e && e[0] == '0') return false;
	return !sc.local && minMinsc >= -8000 && (int64_t)sc.match_bonus * maxLen <= 8000 && sc.match_bonus >= 0;
}

struct DpLaunch {
	const uint8_t  *seq, *qual;
	const uint64_t *roff;
	const bt2g_dp_problem *probs;
	uint64_t        n;
	const uint32_t *nDev;         // optional: problem count produced on the device
	uint64_t        numSlots;     // persistent warp slots (multiple of 4)
	uint8_t        *codes;        // workspace: n * codeStride bytes
	int32_t        *lastH;
	uint64_t       *rawKeys;      // local mode: per-slot candidate keys (score<<32 | row<<16 | col)
	int             maxRaw;        // workspace: n * maxCol ints (e2e last-row scores)
	uint64_t        codeStride;
	int             maxCol;
	int             maxCands, maxAlns, maxOps;
	int             packed = 0;   // e2e: two problems per warp as s16x2 pairs (codes workspace: 2 * codeStride per slot)
	bt2g_dp_summary *summ;
	bt2g_dp_cand    *cands;
	bt2g_dp_aln     *alns;
	uint8_t         *ops;
};
