// dp_kernels.cu -- K3: seed-extension dynamic programming (fill + candidate gather + backtrace)
// for sm_100a.
//
// The reference fills the whole rdlen x (refr-refl+1) rectangle with a striped (Farrar) SSE
// kernel that stores E, F and H for every cell and then repairs the vertical-gap dependency
// in a "lazy F" fix-up loop (aligner_swsse_ee_u8.cpp:775-1146; the fix-up loop dominates its
// run time).  This kernel is a different design:
//   * persistent warps, one DP problem at a time per warp; lane k owns R consecutive read rows;
//     at step t lane k computes column t-k, so the 32 lanes sweep an anti-diagonal wavefront and
//     the vertical (F) and diagonal dependencies cross lanes through one __shfl_up per step --
//     no fix-up loop;
//   * exact 32-bit arithmetic with DPX max/add-max (no saturating 8/16-bit lanes, so the
//     reference's u8 -> i16 fallback (aligner_sw.cpp:518,569-605) has no equivalent here;
//     the two paths produce identical scores by the reference's own debug asserts :622-677);
//   * instead of spilling E/F/H (6 bytes/cell) the fill emits ONE byte per cell: the move the
//     reference's backtrace would take from the H, E and F states of that cell under its fixed
//     preference order diag > ref-gap open > ref-gap extend > read-gap open > read-gap extend
//     (aligner_swsse_ee_u8.cpp:1509-1520, E: :1376-1380, F: :1434-1438).  Bytes are laid out
//     wavefront-major ([step][lane][R]) so every step's store is one coalesced line per warp,
//     in a per-warp-slot workspace that is reused problem after problem and stays in L2;
//   * the backtrace replays SwAligner::nextAlignment (aligner_sw.cpp:737-1146) for ALL
//     candidates in their sorted order, marking visited cells in bit 7 of the move byte.  In
//     the reference a backtrace that reaches an already reported-through cell unwinds its whole
//     branch stack (every stacked cell is itself already marked, :1336-1340,:1561-1585) and
//     fails, so the remaining-option masks never matter and the walk is a pure function of the
//     move bytes + visited bits.  The walk is warp-cooperative: the 32 lanes prefetch the next
//     32 cells of the current diagonal in one load, every lane then steps through them in
//     lock-step (uniform control flow), so the latency chain is one L2 round trip per diagonal
//     run instead of one per cell.
#include "fm_device.cuh"
#include "dp_device.cuh"

#define DP_NEG (-(1 << 28))

__device__ __forceinline__ int dp_max(int a, int b) { return a > b ? a : b; }

// ---- SwAligner::nextAlignment over every candidate (aligner_sw.cpp:737-1146), shared by the
// end-to-end and local kernels.  Warp-cooperative (see the header comment).
template <int R>
__device__ __forceinline__ void dp_backtrace_all(const DpLaunch &L, const bt2g_scoring &sc, const bt2g_dp_problem &p, uint64_t w,
                                                 const uint8_t *rs, const uint8_t *rq, int rdlen, const uint8_t *refw,
                                                 uint8_t *codes, bt2g_dp_cand *cands, int ncand, bt2g_dp_summary *summ,
                                                 int lane, bool local) {
	const int rdgapo = sc.rdgap_const + sc.rdgap_linear, rdgape = sc.rdgap_linear;
	const int rfgapo = sc.rfgap_const + sc.rfgap_linear, rfgape = sc.rfgap_linear;
	int SQ = rdlen >> 4; if(SQ == 0) SQ = 1;                 // aligner_sw.cpp:754-755
	bt2g_dp_aln *alns = L.alns + w * (uint64_t)L.maxAlns;
	uint8_t *ops = L.ops + w * (uint64_t)L.maxAlns * L.maxOps;
	int naln = 0, flags = 0;
	auto cell = [&](int rr, int cc) -> uint8_t * { int k = rr / R; return codes + ((size_t)(cc + k) * 32 + k) * R + (rr - k * R); };
	for(int ci = 0; ci < ncand; ci++) {
		int row = cands[ci].row, col = cands[ci].col;      // same address in every lane: one broadcast load
		const int startRow = row;
		if(local) {
			// start cell already reported through (aligner_sw.cpp:771-789) is checked before domination
			const bool vis = (*cell(row, col) & 0x80) != 0;
			if(vis) { if(lane == 0) cands[ci].fate = BT2G_CAND_FILT_START; continue; }
			// domination by an already attempted candidate (:946-971): within SQ rows and columns
			bool dom = false;
			for(int k = lane; k < ci; k += 32) {
				const int f = cands[k].fate;
				if(f == BT2G_CAND_FAILED || f == BT2G_CAND_SUCCEEDED) {
					int dr = cands[k].row - row, dc = cands[k].col - col;
					dr = dr < 0 ? -dr : dr; dc = dc < 0 ? -dc : dc;
					if(dr <= SQ && dc <= SQ) dom = true;
				}
			}
			if(__any_sync(0xffffffffu, dom)) { if(lane == 0) cands[ci].fate = BT2G_CAND_FILT_DOMINATED; __syncwarp(); continue; }
		}
		// backtraceNucleotidesEnd2EndSseU8 (aligner_swsse_ee_u8.cpp:1283-1877); all lanes run the
		// same scalar walk, lane k additionally owns cell k of the current diagonal prefetch
		uint8_t *o = ops + (size_t)naln * L.maxOps;
		const bool room = naln < L.maxAlns;
		int nops = 0, score = 0, ns = 0, gaps = 0, ct = 0;   // ct: 0 H, 1 E, 2 F
		bool fail = false, core = false, opOverflow = false, done = false, first = true, filtStart = false;
		const int origCol = col;
		int trimBeg = 0;
		while(!done && !fail) {
			// prefetch the diagonal (row-k, col-k), k = lane
			const int rk = row - lane, ck = col - lane;
			uint8_t *cp = (rk >= 0 && ck >= 0) ? cell(rk, ck) : nullptr;
			const uint8_t mine = cp ? *cp : 0xff;
			uint32_t consumed = 0;
			for(int k = 0; k < 32; k++) {
				const uint8_t code = (uint8_t)__shfl_sync(0xffffffffu, (int)mine, k);
				if(code & 0x80) {
					// start cell already reported through -> BT_CAND_FATE_FILT_START (:771-789);
					// anywhere else the backtrace fails
					if(first) filtStart = true;
					fail = true; break;
				}
				first = false;
				consumed |= 1u << k;
				{
					int diagi = col - row + p.triml;
					if(diagi >= p.corel && diagi <= p.corer) core = true;
				}
				if(row == 0) { done = true; break; }
				int mv;   // 1 diag, 2 refopen, 3 rfext, 4 rdopen, 5 rdext
				if(ct == 0) { mv = code & 7; if(mv == 0) { trimBeg = row; done = true; break; } }
				else if(ct == 1) { int e = (code >> 3) & 3; if(e == 0) { fail = true; break; } mv = e == 1 ? 4 : 5; }
				else { int f = (code >> 5) & 3; if(f == 0) { fail = true; break; } mv = f == 1 ? 2 : 3; }
				const int pos = p.fw ? row : rdlen - 1 - row;
				int c = rs[pos]; c = p.fw ? c : (c > 3 ? 4 : 3 - c);
				int q = (int)rq[pos] - 33; q = q < 0 ? 0 : (q > 63 ? 63 : q);
				const int refc = refw[col];
				uint8_t op;
				bool stay = false;
				if(mv == 1) {
					if(c > 3 || refc > 3) { score -= sc.npen[q]; ns++; op = BT2G_OP_MM; }
					else if(c == refc) { score += sc.match_bonus; op = BT2G_OP_MATCH; }
					else { score -= sc.mmpen[q]; op = BT2G_OP_MM; }
					op |= (uint8_t)(refc << 2);
					row--; col--; ct = 0; stay = true;
				} else if(mv == 2 || mv == 3) {
					score -= (mv == 2) ? rfgapo : rfgape; gaps++;
					op = BT2G_OP_REFGAP;
					row--; ct = (mv == 2) ? 0 : 2;
				} else {
					score -= (mv == 4) ? rdgapo : rdgape; gaps++;
					op = BT2G_OP_READGAP | (uint8_t)(refc << 2);
					col--; ct = (mv == 4) ? 0 : 1;
				}
				if(room && lane == 0) { if(nops < L.maxOps) o[nops] = op; }
				if(nops >= L.maxOps) opOverflow = true;
				nops++;
				if(!stay) break;            // left the prefetched diagonal
			}
			if(cp && ((consumed >> lane) & 1u)) *cp = mine | 0x80;     // setReportedThrough (:1555)
			__syncwarp();
		}
		if(filtStart) { if(lane == 0) cands[ci].fate = BT2G_CAND_FILT_START; continue; }
		if(!fail) {
			// the alignment's first cell (row, col) (:1797-1813)
			const int pos = p.fw ? row : rdlen - 1 - row;
			int c = rs[pos]; c = p.fw ? c : (c > 3 ? 4 : 3 - c);
			int q = (int)rq[pos] - 33; q = q < 0 ? 0 : (q > 63 ? 63 : q);
			const int refc = refw[col];
			uint8_t op;
			if(c > 3 || refc > 3) { score -= sc.npen[q]; ns++; op = BT2G_OP_MM; }
			else if(c == refc) { score += sc.match_bonus; op = BT2G_OP_MATCH; }
			else { score -= sc.mmpen[q]; op = BT2G_OP_MM; }
			op |= (uint8_t)(refc << 2);
			if(!core) fail = true;                   // core-diagonal rejection (:1764-1795)
			else if(ns > p.nceil) fail = true;       // N ceiling (:1813-1818)
			else {
				if(room && lane == 0) { if(nops < L.maxOps) o[nops] = op; }
				if(nops >= L.maxOps) opOverflow = true;
				nops++;
			}
		}
		if(fail) { if(lane == 0) cands[ci].fate = BT2G_CAND_FAILED; continue; }
		if(lane == 0) cands[ci].fate = BT2G_CAND_SUCCEEDED;
		if(room) {
			if(lane == 0) {
				bt2g_dp_aln &a = alns[naln];
				a.cand_idx = ci; a.score = score; a.ns = ns; a.gaps = gaps; a.col0 = col; a.row0 = row;
				a.trim_beg = trimBeg; a.trim_end = rdlen - 1 - startRow; a.nops = nops;
				int refns = 0;
				for(int k = col; k <= origCol; k++) refns += refw[k] > 3;
				a.refns = refns;
			}
			if(opOverflow) flags |= BT2G_DP_FLAG_OPS_OVERFLOW;
		} else {
			flags |= BT2G_DP_FLAG_ALN_OVERFLOW;
		}
		naln++;
	}
	if(lane == 0) { summ->naln = naln; summ->flags |= flags; }
}

// R = rows per lane (rdlen <= 32*R)
template <typename OFF, int R>
__global__ void __launch_bounds__(128) k_dp_e2e(DevIndex<OFF> ix, bt2g_scoring sc, DpLaunch L) {
	extern __shared__ uint8_t smem[];
	const int warpInBlock = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint64_t slot = blockIdx.x * (uint64_t)(blockDim.x >> 5) + warpInBlock;
	const uint64_t nSlots = (uint64_t)gridDim.x * (blockDim.x >> 5);
	const uint64_t nProb = L.nDev ? (uint64_t)*L.nDev : L.n;
	// per-warp shared memory: last-row scores (ints) then the reference window (bytes)
	const size_t perWarp = ((size_t)L.maxCol * 5 + 16 + 15) & ~(size_t)15;
	int32_t *lastH = reinterpret_cast<int32_t *>(smem + (size_t)warpInBlock * perWarp);
	uint8_t *refw = reinterpret_cast<uint8_t *>(lastH + L.maxCol);
	// persistent warps: the move-byte workspace belongs to the warp SLOT, not to the problem, so
	// it is (#SMs x resident warps) x codeStride bytes and stays L2-resident across problems
	uint8_t *codes = L.codes + slot * L.codeStride;
	const int rdgapo = sc.rdgap_const + sc.rdgap_linear, rdgape = sc.rdgap_linear;
	const int rfgapo = sc.rfgap_const + sc.rfgap_linear, rfgape = sc.rfgap_linear;

	for(uint64_t w = slot; w < nProb; w += nSlots) {
		const bt2g_dp_problem p = L.probs[w];
		const uint8_t *rs = L.seq + L.roff[p.read_idx];
		const uint8_t *rq = L.qual + L.roff[p.read_idx];
		const int rdlen = (int)(L.roff[p.read_idx + 1] - L.roff[p.read_idx]);
		const int ncol = (int)(p.refr - p.refl + 1);
		bt2g_dp_summary *summ = L.summ + w;
		__syncwarp();
		if(ncol <= 0 || ncol > L.maxCol || rdlen > 32 * R || rdlen <= 0) {
			if(lane == 0) { summ->found = 0; summ->best = DP_NEG; summ->ncand = 0; summ->naln = 0; summ->flags = BT2G_DP_FLAG_BADSHAPE; }
			continue;
		}
		// reference window (SwAligner::initRef, aligner_sw.cpp:155-271): codes 0..3, 4 = N / off-end
		for(int k = lane; k < ncol; k += 32) refw[k] = (uint8_t)ref_base<OFF>(ix, p.tidx, p.refl + k);
		__syncwarp();

		// per-row constants (buildQueryProfileEnd2EndSseU8, aligner_swsse_ee_u8.cpp:75-142)
		int rc[R], mmp[R], npn[R];
		bool bar[R];
#pragma unroll
		for(int r = 0; r < R; r++) {
			int i = lane * R + r;
			if(i < rdlen) {
				int pos = p.fw ? i : rdlen - 1 - i;
				int c = rs[pos];
				rc[r] = p.fw ? c : (c > 3 ? 4 : 3 - c);
				int q = (int)rq[pos] - 33;
				q = q < 0 ? 0 : (q > 63 ? 63 : q);
				npn[r] = sc.npen[q];
				// a read N never "equals" a reference base and is charged the N penalty (Scoring::score, scoring.h:241-251)
				mmp[r] = rc[r] > 3 ? npn[r] : sc.mmpen[q];
				if(rc[r] > 3) rc[r] = 5;
				bar[r] = (i < sc.gapbar) || (rdlen - 1 - i < sc.gapbar);
			} else { rc[r] = 5; mmp[r] = 0; npn[r] = 0; bar[r] = true; }
		}
		const int lastLane = (rdlen - 1) / R, lastR = (rdlen - 1) % R;

		// Branch-free move codes.  With E = max(E'-rdgape, H'-rdgapo) and F = max(F^-rfgape, H^-rfgapo),
		// "which term is the max (open wins ties)" IS the E/F move, and the H move is
		//   diag if H == Hd, else the F move if H == F, else the E move
		// -- the same choice the reference makes from its five equality tests in preference order
		// (aligner_swsse_ee_u8.cpp:1468-1520): H==F with F==open implies H==H^-rfgapo (bit 0), H==F with
		// F==extend only implies bit 2, and likewise for E (bits 1, 3).  Cells below the minimum score
		// may get arbitrary codes; no backtrace visits them (scores are monotone along a path).
		int Hleft[R], Earr[R], Esel[R];
#pragma unroll
		for(int r = 0; r < R; r++) { Hleft[r] = DP_NEG; Earr[r] = DP_NEG; Esel[r] = 1; }
		int botH = DP_NEG, botF = DP_NEG, prevInH = DP_NEG;
		const int nsteps = ncol + lastLane;   // lanes beyond lastLane hold no rows
		for(int t = 0; t < nsteps; t++) {
			int inH = __shfl_up_sync(0xffffffffu, botH, 1);
			int inF = __shfl_up_sync(0xffffffffu, botF, 1);
			if(lane == 0) { inH = DP_NEG; inF = DP_NEG; }
			const int j = t - lane;
			if(j >= 0 && j < ncol && lane <= lastLane) {
				const int refc = refw[j];
				const bool refN = refc > 3;
				// H[i0-1][j-1]: row -1 is the free start row of end-to-end mode (vhilsw, :853,923-927)
				int diag = (lane == 0) ? 0 : prevInH;
				int upH = inH, upF = inF;
				uint32_t packed[(R + 3) / 4];
#pragma unroll
				for(int q4 = 0; q4 < (R + 3) / 4; q4++) packed[q4] = 0;
#pragma unroll
				for(int r = 0; r < R; r++) {
					// F[i][j] = max(F[i-1][j]-rfgape, H[i-1][j]-rfgapo), vetoed in gap-barrier rows (:944-945,:983-985)
					const int fo = upH - rfgapo, fe = upF - rfgape;
					const int fsel = fo >= fe ? 1 : 2;
					int F = bar[r] ? DP_NEG : dp_max(fo, fe);
					int s = (rc[r] == refc) ? sc.match_bonus : -mmp[r];
					s = refN ? -npn[r] : s;
					const int Hd = diag + s;
					const int E = Earr[r];
					const int H = __vimax3_s32(Hd, E, F);
					const int hsel = (H == Hd) ? 1 : ((H == F) ? 1 + fsel : (bar[r] ? 0 : 3 + Esel[r]));
					const uint32_t code = (uint32_t)(hsel | (Esel[r] << 3) | (fsel << 5));
					packed[r >> 2] |= code << ((r & 3) * 8);
					// E[i][j+1] = max(E[i][j]-rdgape, H[i][j]-rdgapo [vetoed in barrier rows]) (:966-969)
					const int eo = bar[r] ? DP_NEG : H - rdgapo, ee = E - rdgape;
					Esel[r] = eo >= ee ? 1 : 2;
					diag = Hleft[r]; Hleft[r] = H; Earr[r] = dp_max(eo, ee);
					upH = H; upF = F;
				}
				if(lane == lastLane) {
					int hl = Hleft[0];
#pragma unroll
					for(int r = 1; r < R; r++) hl = (lastR == r) ? Hleft[r] : hl;
					lastH[j] = hl;
				}
				botH = upH; botF = upF;
				prevInH = inH;
				uint8_t *dst = codes + ((size_t)t * 32 + lane) * R;
				if(R == 4) *reinterpret_cast<uint32_t *>(dst) = packed[0];
				else if(R == 8) *reinterpret_cast<uint2 *>(dst) = make_uint2(packed[0], packed[1]);
				else {
#pragma unroll
					for(int q4 = 0; q4 < (R + 3) / 4; q4++) reinterpret_cast<uint32_t *>(dst)[q4] = packed[q4];
				}
			} else if(j >= ncol) {
				botH = DP_NEG; botF = DP_NEG;
			}
		}
		__syncwarp();
		// best = max of the last row (aligner_swsse_ee_u8.cpp:1095-1100)
		int best = DP_NEG;
		for(int k = lane; k < ncol; k += 32) best = dp_max(best, lastH[k]);
#pragma unroll
		for(int o = 16; o > 0; o >>= 1) best = dp_max(best, __shfl_xor_sync(0xffffffffu, best, o));

		// ---- SwAligner::align tail (aligner_sw.cpp:679-729) + gatherCellsNucleotidesEnd2End (:1176-1208)
		if(lane == 0) { summ->best = best; summ->flags = 0; summ->naln = 0; summ->ncand = 0; summ->found = 0; }
		if(best < p.minsc) continue;
		bt2g_dp_cand *cands = L.cands + w * (uint64_t)L.maxCands;
		// rank of each candidate under DpBtCandidate::operator< (score desc, row equal, col desc;
		// aligner_sw_nuc.h:149-157) computed directly: every lane ranks the columns it owns
		int totalCand = 0;
		for(int j0 = 0; j0 < ncol; j0 += 32) {
			const int j = j0 + lane;
			const int s = j < ncol ? lastH[j] : DP_NEG;
			const bool isC = j < ncol && s >= p.minsc;
			if(isC) {
				int rank = 0;
				for(int k = 0; k < ncol; k++) {
					const int sk = lastH[k];
					rank += (sk >= p.minsc) && (sk > s || (sk == s && k > j));
				}
				if(rank < L.maxCands) { cands[rank].score = s; cands[rank].col = j; cands[rank].row = rdlen - 1; cands[rank].fate = 0; }
			}
			totalCand += __popc(__ballot_sync(0xffffffffu, isC));
		}
		const int ncand = totalCand < L.maxCands ? totalCand : L.maxCands;
		if(lane == 0) {
			summ->ncand = totalCand; summ->found = totalCand > 0;
			if(totalCand > L.maxCands) summ->flags |= BT2G_DP_FLAG_CAND_OVERFLOW;
		}
		__syncwarp();

		dp_backtrace_all<R>(L, sc, p, w, rs, rq, rdlen, refw, codes, cands, ncand, summ, lane, false);
	} // persistent loop over problems
}

// ----------------------------------------------------------------------------------------
template <typename OFF>
int launch_dp_e2e(const DevIndex<OFF> &ix, const bt2g_scoring &sc, const DpLaunch &L, int maxRdLen, cudaStream_t st) {
	if(L.n == 0) return 0;
	const int warpsPerBlock = 4;
	const size_t perWarp = ((size_t)L.maxCol * 5 + 16 + 15) & ~(size_t)15;
	size_t smem = (size_t)warpsPerBlock * perWarp;
	unsigned grid = (unsigned)(L.numSlots / warpsPerBlock);
	if(maxRdLen <= 128) {
		if(smem > 48 * 1024) cudaFuncSetAttribute(k_dp_e2e<OFF, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		k_dp_e2e<OFF, 4><<<grid, warpsPerBlock * 32, smem, st>>>(ix, sc, L);
	} else if(maxRdLen <= 256) {
		if(smem > 48 * 1024) cudaFuncSetAttribute(k_dp_e2e<OFF, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		k_dp_e2e<OFF, 8><<<grid, warpsPerBlock * 32, smem, st>>>(ix, sc, L);
	} else if(maxRdLen <= 512) {
		if(smem > 48 * 1024) cudaFuncSetAttribute(k_dp_e2e<OFF, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		k_dp_e2e<OFF, 16><<<grid, warpsPerBlock * 32, smem, st>>>(ix, sc, L);
	} else {
		return -1;
	}
	return 0;
}
template int launch_dp_e2e<uint32_t>(const DevIndex<uint32_t> &, const bt2g_scoring &, const DpLaunch &, int, cudaStream_t);
template int launch_dp_e2e<uint64_t>(const DevIndex<uint64_t> &, const bt2g_scoring &, const DpLaunch &, int, cudaStream_t);

// ----------------------------------------------------------------------------------------
// Local mode (alignNucleotidesLocalSseU8/I16, aligner_swsse_loc_i16.cpp:938-1367; gather
// :1420-1535; backtrace :1615-2218).  Same wavefront; differences from end-to-end:
//   * every score is floored at 0 (the reference stores score-0x8000 and lets signed saturation
//     clamp, :1004-1017,1107), row -1 and column -1 are 0;
//   * a move is legal only from a source cell whose score is > 0 (floorsc = 0, :1683-1840), so the
//     move codes are computed from the explicit equality tests, and "no legal move" ends the
//     alignment (soft trimming);
//   * candidates are all cells with score >= minsc in rows >= minrow whose own base matches and
//     whose diagonal successor does not (:1497-1518), collected during the fill.
// The u8 -> i16 rerun of the reference (aligner_sw.cpp:569-605) has no counterpart: scores are exact.
// ----------------------------------------------------------------------------------------
template <typename OFF, int R>
__global__ void __launch_bounds__(128) k_dp_local(DevIndex<OFF> ix, bt2g_scoring sc, DpLaunch L) {
	extern __shared__ uint8_t smem[];
	const int warpInBlock = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint64_t slot = blockIdx.x * (uint64_t)(blockDim.x >> 5) + warpInBlock;
	const uint64_t nSlots = (uint64_t)gridDim.x * (blockDim.x >> 5);
	const uint64_t nProb = L.nDev ? (uint64_t)*L.nDev : L.n;
	const size_t perWarp = ((size_t)L.maxCol * 5 + 16 + 15) & ~(size_t)15;
	int32_t *wsm = reinterpret_cast<int32_t *>(smem + (size_t)warpInBlock * perWarp);   // [0] = raw candidate counter
	uint8_t *refw = reinterpret_cast<uint8_t *>(wsm + L.maxCol);
	uint8_t *codes = L.codes + slot * L.codeStride;
	uint64_t *raw = L.rawKeys + slot * (uint64_t)L.maxRaw;
	const int rdgapo = sc.rdgap_const + sc.rdgap_linear, rdgape = sc.rdgap_linear;
	const int rfgapo = sc.rfgap_const + sc.rfgap_linear, rfgape = sc.rfgap_linear;
	const int bonus = sc.match_bonus;

	for(uint64_t w = slot; w < nProb; w += nSlots) {
		const bt2g_dp_problem p = L.probs[w];
		const uint8_t *rs = L.seq + L.roff[p.read_idx];
		const uint8_t *rq = L.qual + L.roff[p.read_idx];
		const int rdlen = (int)(L.roff[p.read_idx + 1] - L.roff[p.read_idx]);
		const int ncol = (int)(p.refr - p.refl + 1);
		bt2g_dp_summary *summ = L.summ + w;
		__syncwarp();
		if(ncol <= 0 || ncol + 1 > L.maxCol || rdlen > 32 * R || rdlen <= 0 || bonus <= 0) {
			if(lane == 0) { summ->found = 0; summ->best = DP_NEG; summ->ncand = 0; summ->naln = 0; summ->flags = BT2G_DP_FLAG_BADSHAPE; }
			continue;
		}
		// reference window plus the one extra character initRef captures (aligner_sw.cpp:170-173)
		for(int k = lane; k <= ncol; k += 32) refw[k] = (uint8_t)ref_base<OFF>(ix, p.tidx, p.refl + k);
		if(lane == 0) wsm[0] = 0;
		__syncwarp();

		int rc[R + 1], mmp[R], npn[R];
		bool bar[R];
#pragma unroll
		for(int r = 0; r <= R; r++) {
			int i = lane * R + r;
			int c = 6;                                   // beyond the read: matches nothing
			if(i < rdlen) {
				int pos = p.fw ? i : rdlen - 1 - i;
				c = rs[pos];
				c = p.fw ? c : (c > 3 ? 4 : 3 - c);
				if(r < R) {
					int q = (int)rq[pos] - 33;
					q = q < 0 ? 0 : (q > 63 ? 63 : q);
					npn[r] = sc.npen[q];
					mmp[r] = c > 3 ? npn[r] : sc.mmpen[q];
					bar[r] = (i < sc.gapbar) || (rdlen - 1 - i < sc.gapbar);
				}
			} else if(r < R) { mmp[r] = 0; npn[r] = 0; bar[r] = true; }
			rc[r] = c;                                   // raw code 0..4 (N = 4 matches a reference N in the gather test)
		}
		const int lastLane = (rdlen - 1) / R;
		const int minrow = (int)(((int64_t)p.minsc + bonus - 1) / bonus) - 1;   // aligner_swsse_loc_i16.cpp:1437

		int Hleft[R], Earr[R], Eprev[R];
#pragma unroll
		for(int r = 0; r < R; r++) { Hleft[r] = 0; Earr[r] = 0; Eprev[r] = 0; }
		int botH = 0, botF = 0, prevInH = 0, lmax = 0;
		const int nsteps = ncol + lastLane;
		for(int t = 0; t < nsteps; t++) {
			int inH = __shfl_up_sync(0xffffffffu, botH, 1);
			int inF = __shfl_up_sync(0xffffffffu, botF, 1);
			if(lane == 0) { inH = 0; inF = 0; }
			const int j = t - lane;
			if(j >= 0 && j < ncol && lane <= lastLane) {
				const int refc = refw[j], refn = refw[j + 1];
				const bool refN = refc > 3;
				int diag = (lane == 0) ? 0 : prevInH;
				int upH = inH, upF = inF;
				uint32_t packed[(R + 3) / 4];
#pragma unroll
				for(int q4 = 0; q4 < (R + 3) / 4; q4++) packed[q4] = 0;
#pragma unroll
				for(int r = 0; r < R; r++) {
					const int i = lane * R + r;
					const int fo = upH - rfgapo, fe = upF - rfgape;
					int F = bar[r] ? 0 : dp_max(dp_max(fo, fe), 0);
					const int fsel = (upH > 0 && fo == F) ? 1 : ((upF > 0 && fe == F) ? 2 : 0);
					int s = (rc[r] == refc && !refN) ? bonus : -mmp[r];
					s = (refN || rc[r] > 3) ? -npn[r] : s;
					const int Hd = diag + s;
					const int E = Earr[r];
					const int H = dp_max(__vimax3_s32(Hd, E, F), 0);
					int hsel = 0;
					if(diag > 0 && H == Hd) hsel = 1;
					else if(!bar[r]) {
						if(upH > 0 && H == fo) hsel = 2;
						else if(upF > 0 && H == fe) hsel = 3;
						else if(Hleft[r] > 0 && H == Hleft[r] - rdgapo) hsel = 4;
						else if(Eprev[r] > 0 && H == Eprev[r] - rdgape) hsel = 5;
					}
					const int esel = (Hleft[r] > 0 && Hleft[r] - rdgapo == E) ? 1 : ((Eprev[r] > 0 && Eprev[r] - rdgape == E) ? 2 : 0);
					const uint32_t code = (uint32_t)(hsel | (esel << 3) | (fsel << 5));
					packed[r >> 2] |= code << ((r & 3) * 8);
					const int eo = bar[r] ? 0 : H - rdgapo, ee = E - rdgape;
					// candidate cell (gatherCellsNucleotidesLocalSseI16, :1485-1518)
					if(i < rdlen && i >= minrow && H >= p.minsc && rc[r] == refc && !(i < rdlen - 1 && rc[r + 1] == refn)) {
						const int pos = atomicAdd(&wsm[0], 1);
						if(pos < L.maxRaw) raw[pos] = ((uint64_t)(uint32_t)H << 32) | ((uint64_t)i << 16) | (uint64_t)j;
					}
					lmax = dp_max(lmax, i < rdlen ? H : 0);
					diag = Hleft[r]; Hleft[r] = H; Eprev[r] = E; Earr[r] = dp_max(dp_max(eo, ee), 0);
					upH = H; upF = F;
				}
				botH = upH; botF = upF;
				prevInH = inH;
				uint8_t *dst = codes + ((size_t)t * 32 + lane) * R;
				if(R == 4) *reinterpret_cast<uint32_t *>(dst) = packed[0];
				else if(R == 8) *reinterpret_cast<uint2 *>(dst) = make_uint2(packed[0], packed[1]);
				else {
#pragma unroll
					for(int q4 = 0; q4 < (R + 3) / 4; q4++) reinterpret_cast<uint32_t *>(dst)[q4] = packed[q4];
				}
			} else if(j >= ncol) {
				botH = 0; botF = 0;
			}
		}
		__syncwarp();
		int best = lmax;
#pragma unroll
		for(int o = 16; o > 0; o >>= 1) best = dp_max(best, __shfl_xor_sync(0xffffffffu, best, o));
		if(lane == 0) { summ->best = best; summ->flags = 0; summ->naln = 0; summ->ncand = 0; summ->found = 0; }
		if(best < p.minsc) continue;
		// sort: DpBtCandidate::operator< = score desc, row desc, col desc = key desc
		const int nrawAll = wsm[0];
		const int nraw = nrawAll < L.maxRaw ? nrawAll : L.maxRaw;
		bt2g_dp_cand *cands = L.cands + w * (uint64_t)L.maxCands;
		for(int a0 = 0; a0 < nraw; a0 += 32) {
			const int a = a0 + lane;
			if(a < nraw) {
				const uint64_t key = raw[a];
				int rank = 0;
				for(int k = 0; k < nraw; k++) rank += raw[k] > key;
				if(rank < L.maxCands) {
					cands[rank].score = (int32_t)(key >> 32); cands[rank].row = (int32_t)((key >> 16) & 0xffff);
					cands[rank].col = (int32_t)(key & 0xffff); cands[rank].fate = 0;
				}
			}
		}
		const int ncand = nraw < L.maxCands ? nraw : L.maxCands;
		if(lane == 0) {
			summ->ncand = nrawAll; summ->found = nrawAll > 0;
			if(nrawAll > L.maxCands || nrawAll > L.maxRaw) summ->flags |= BT2G_DP_FLAG_CAND_OVERFLOW;
		}
		__syncwarp();
		dp_backtrace_all<R>(L, sc, p, w, rs, rq, rdlen, refw, codes, cands, ncand, summ, lane, true);
	}
}

template <typename OFF>
int launch_dp_local(const DevIndex<OFF> &ix, const bt2g_scoring &sc, const DpLaunch &L, int maxRdLen, cudaStream_t st) {
	if(L.n == 0) return 0;
	const int warpsPerBlock = 4;
	const size_t perWarp = ((size_t)L.maxCol * 5 + 16 + 15) & ~(size_t)15;
	size_t smem = (size_t)warpsPerBlock * perWarp;
	unsigned grid = (unsigned)(L.numSlots / warpsPerBlock);
	if(maxRdLen <= 128) {
		if(smem > 48 * 1024) cudaFuncSetAttribute(k_dp_local<OFF, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		k_dp_local<OFF, 4><<<grid, warpsPerBlock * 32, smem, st>>>(ix, sc, L);
	} else if(maxRdLen <= 256) {
		if(smem > 48 * 1024) cudaFuncSetAttribute(k_dp_local<OFF, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		k_dp_local<OFF, 8><<<grid, warpsPerBlock * 32, smem, st>>>(ix, sc, L);
	} else if(maxRdLen <= 512) {
		if(smem > 48 * 1024) cudaFuncSetAttribute(k_dp_local<OFF, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		k_dp_local<OFF, 16><<<grid, warpsPerBlock * 32, smem, st>>>(ix, sc, L);
	} else {
		return -1;
	}
	return 0;
}
template int launch_dp_local<uint32_t>(const DevIndex<uint32_t> &, const bt2g_scoring &, const DpLaunch &, int, cudaStream_t);
template int launch_dp_local<uint64_t>(const DevIndex<uint64_t> &, const bt2g_scoring &, const DpLaunch &, int, cudaStream_t);
