// dp_device.cuh -- launch descriptor shared by api.cu and dp_kernels.cu (DP types: include/bt2g.h)
#pragma once
#include "bt2g_internal.h"

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
	bt2g_dp_summary *summ;
	bt2g_dp_cand    *cands;
	bt2g_dp_aln     *alns;
	uint8_t         *ops;
};
