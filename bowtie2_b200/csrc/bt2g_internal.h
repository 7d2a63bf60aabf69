// bt2g_internal.h -- shared declarations of libbt2g.so (not part of the public ABI).
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/bt2g.h"

// ---- device-side view of one FM index (forward or mirror) --------------------------------
// Layout facts restated from the reference: a "side" is sideSz = 16*OFF_SIZE bytes = 48*OFF_SIZE
// bytes of 2-bit BWT (LSB-first pairs) followed by four OFF occurrence counts A,C,G,T that
// hold the counts BEFORE this side (bt2_idx.h:133-167, :1753-1756, :1904-1917).
template <typename OFF>
struct DevEbwt {
	const uint8_t *ebwt;
	const OFF     *ftab;
	const OFF     *eftab;
	uint64_t       len;
	uint64_t       zOff;
	uint64_t       zSide;   // zOff / sideBwtLen
	uint32_t       zChar;   // zOff % sideBwtLen
	uint64_t       fchr[5];
	int            ftabChars;
};

// engine-internal hit encoding: a BW "row" with this bit set is already the joined-text offset of the hit (1-mismatch search of
// a unique occurrence, fm_onemm.cu: one_mm_text); never crosses the C ABI
#define BT2G_ROW_IS_OFFSET (1ull << 63)

template <typename OFF>
struct DevIndex {
	DevEbwt<OFF>   fw, bw;
	const OFF     *offs;
	int            offRate;
	const OFF     *rstarts;  // 3*nFrag
	uint64_t       nFrag;
	const OFF     *plen;
	uint64_t       nPat;
	// packed reference
	const OFF     *recOff, *recLen;
	const uint64_t *recCumOff;    // [nRecs] position within its reference where record i's N-run starts
	const uint64_t *recCumUnamb;  // [nRecs] unambiguous bases preceding record i in ref_buf
	const uint64_t *refRecOffs;   // [nRefs+1] first record of each reference
	const uint64_t *refLens;      // [nRefs]
	const uint8_t  *refBuf;
	uint64_t       nRecs, nRefs;
	// optional extended seed table (bt2g_build_seed_table): for every K-mer the state of the bidirectional
	// search after its K characters: 3 OFF per entry = topf, botf, topb (all 0 = empty range)
	const OFF     *ktab;
	int            ktabChars;
	// SA sample in effect: the index's own offs[] / offRate, or the denser one of bt2g_build_dense_sa
	const OFF     *saOffs;
	int            saRate;
	// SwDriver::extend of a unique seed hit by comparing the read with the packed reference (fm_device.cuh: extend_one_text)
	// instead of walking the index; 0 = always walk (bt2g_set_extend_mode)
	int            extText;
};

struct DevArray {
	void    *ptr = nullptr;
	uint64_t bytes = 0;
	bool     owned = false;
};

struct bt2g_ctx {
	int device = 0;
	std::string err;
	bool loaded = false;
	bt2g_index_info info{};
	DevArray arr[BT2G_N_INDEX_ARRAYS];
	// derived reference tables (always owned)
	DevArray recCumOff, recCumUnamb, refRecOffs, refLens;
	uint64_t nRefs = 0;
	DevArray ktab; int ktabChars = 0;
	DevArray denseSa; int denseRate = -1;
	cudaStream_t stream = nullptr;
	bt2g_scoring scoring{};
	int extendText = 1;            // unique seed hits are extended against the packed reference (bt2g_set_extend_mode)
	int dpModeCap = 3;             // highest end-to-end DP kernel generation the launchers may pick (dp_device.cuh)
	// scratch buffers (grown on demand)
	std::vector<DevArray> scratch;
};

#define BT2G_CUDA_TRY(ctx, expr)                                                              \
	do {                                                                                      \
		cudaError_t e_ = (expr);                                                              \
		if(e_ != cudaSuccess) {                                                               \
			(ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                  \
			return -2;                                                                        \
		}                                                                                     \
	} while(0)

// host index reader (index_host.cpp)
struct HostIndex {
	bt2g_index_host d{};
	std::vector<uint8_t> plen, rstarts, ebwt_fw, ebwt_bw, ftab_fw, eftab_fw, ftab_bw, eftab_bw, offs;
	std::vector<uint8_t> rec_off, rec_len, rec_first, ref_buf;
	std::vector<std::string> names;                 // reference names stored in the .1 file
};
int bt2g_read_index_files(const char *basename, HostIndex &out, std::string &err);
// offRateOverride < 0: none (Ebwt's _overrideOffRate, bt2_io.cpp:217-230)
int bt2g_read_index_files_ex(const char *basename, int offRateOverride, HostIndex &out, std::string &err);

template <typename OFF> DevIndex<OFF> bt2g_dev_index(const bt2g_ctx *ctx);
