// sam_host.cpp -- SAM records from pipeline results (host only; no CUDA).
//
// The reporting tail of the reference for one alignment per read: AlnSinkSam::appendMate (aln_sink.cpp:1889-2060)
// with StackedAln (aligner_result.cpp:520-880: stacked alignment, leftAlign(false), CIGAR, MD:Z) and
// SamConfig::printAlignedOptFlags (sam.cpp:121-330: AS, XS, XN, XM, XO, XG, NM, MD, YS, YT in that order).
// Inputs are what bt2g_pipeline_run_*_host returns: bt2g_read_result + the op string of the reported alignment
// (ops list the alignment from its last read position back to its first, in reference orientation; every op
// that consumes a reference character carries its code in bits 2..4).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include <memory>
#include <mutex>
#include <atomic>
#include "../../include/bt2g.h"

namespace {

// Scratch blocks that outlive a call: a batch's text is hundreds of megabytes, and memory fresh from the allocator costs a page
// fault per 4 KB (and an munmap with its TLB shootdown when it is freed, felt by every other host thread of the process).  Blocks
// go back to the pool instead; the pool keeps at most POOL_BYTES.
struct Block {
	std::unique_ptr<uint8_t[]> p; size_t cap = 0;
	uint8_t *get() const { return p.get(); }
};
class BlockPool {
	static constexpr size_t POOL_BYTES = 6ull << 30, POOL_BLOCKS = 96;
	std::mutex m; std::vector<Block> idle; size_t held = 0;
public:
	Block take(size_t n) {
		{
			std::lock_guard<std::mutex> g(m);
			size_t best = idle.size();
			for(size_t k = 0; k < idle.size(); k++) if(idle[k].cap >= n && (best == idle.size() || idle[k].cap < idle[best].cap)) best = k;
			if(best != idle.size() && idle[best].cap <= 4 * n + (1u << 20)) {          // (not a huge block for a small request)
				Block b = std::move(idle[best]);
				idle.erase(idle.begin() + (long)best);
				held -= b.cap;
				return b;
			}
		}
		Block b; b.cap = (n + (1u << 16)) & ~(size_t)0xffff; b.p.reset(new uint8_t[b.cap]);   // uninitialised
		return b;
	}
	void give(Block &&b) {
		if(!b.p) return;
		std::lock_guard<std::mutex> g(m);
		if(held + b.cap > POOL_BYTES || idle.size() >= POOL_BLOCKS) return;          // freed
		held += b.cap;
		idle.push_back(std::move(b));
	}
};
BlockPool &pool() { static BlockPool p; return p; }

// text under construction in a pooled block
struct OutBuf {
	Block blk; size_t n = 0;
	~OutBuf() { pool().give(std::move(blk)); }
	char *room(size_t extra) {                                   // at least `extra` writable bytes at the returned cursor
		if(n + extra > blk.cap) {
			Block nb = pool().take(std::max(n + extra, blk.cap + blk.cap / 2));
			if(n) memcpy(nb.get(), blk.get(), n);
			pool().give(std::move(blk));
			blk = std::move(nb);
		}
		return (char *)blk.get() + n;
	}
};

inline char *putInt(char *p, long long v) {
	unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
	if(v < 0) *p++ = '-';
	if(u < 10) { *p++ = (char)('0' + u); return p; }
	char b[24]; int k = 24;
	do { b[--k] = (char)('0' + u % 10); u /= 10; } while(u);
	memcpy(p, b + k, (size_t)(24 - k));
	return p + (24 - k);
}
inline char *putStr(char *p, const char *s) { const size_t l = strlen(s); memcpy(p, s, l); return p + l; }
#define PUT_LIT(p, lit) (memcpy((p), (lit), sizeof(lit) - 1), (p) + (sizeof(lit) - 1))

struct Stacked {
	std::vector<char> ref, rel, rd;          // reference char / relation (= X I D) / read char per alignment column
	void leftAlign() {                       // StackedAln::leftAlign(false) (aligner_result.cpp:629-675)
		const size_t ln = ref.size();
		for(size_t i = 0; i < ln; i++) {
			const char r = rel[i];
			if(r == '=' || r == 'X') continue;
			size_t glen = 1;
			for(size_t j = i + 1; j < ln; j++) { if(rel[j] != r) break; glen++; }
			if(i == 0) { i += glen - 1; continue; }
			size_t l = i - 1, rr = l + glen;
			std::vector<char> &gp = (r == 'I') ? ref : rd;
			const std::vector<char> &ngp = (r == 'I') ? rd : ref;
			while(l > 0 && ngp[l] == ngp[rr]) {
				if(rel[l] == 'X') break;
				std::swap(gp[l], gp[rr]);
				std::swap(rel[l], rel[rr]);
				l--; rr--;
			}
			i += glen - 1;
		}
	}
};

void appendInt(std::string &o, long long v) {
	char b[24]; int k = 24;
	unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
	do { b[--k] = (char)('0' + u % 10); u /= 10; } while(u);
	if(v < 0) b[--k] = '-';
	o.append(b + k, (size_t)(24 - k));
}

} // namespace

static int formatRange(const bt2g_sam_opts *opt, const bt2g_reads *reads, const bt2g_read_result *res, const uint8_t *ops,
                       uint32_t maxOps, const bt2g_pair_result *pairs, uint64_t i0, uint64_t i1, OutBuf &dst) {
	// (a local buffer object: the callers' OutBufs sit side by side in one vector, and a cursor updated per record there would
	// bounce its cache line between the formatting threads)
	OutBuf o;
	std::swap(o.blk, dst.blk);
	struct Back { OutBuf &o, &dst; ~Back() { std::swap(o.blk, dst.blk); dst.n = o.n; o.n = 0; } } back{o, dst};
	static const char dna[] = "ACGTN";
	static const char comp[] = "TGCAN";
	// code -> character, any code above 4 reads as N
	static const struct Chars { char fw[256], rc[256]; Chars() { for(int c = 0; c < 256; c++) { fw[c] = "ACGTN"[c > 4 ? 4 : c]; rc[c] = "TGCAN"[c > 4 ? 4 : c]; } } } chars;
	const bool paired = pairs != nullptr;
	const bool xeq = (opt->flags & BT2G_SAM_XEQ) != 0, noUnal = (opt->flags & BT2G_SAM_NO_UNAL) != 0;
	const bool noDiscord = (opt->flags & BT2G_SAM_NO_DISCORDANT) != 0;
	o.n = 0;
	if(i1 > i0) o.room((size_t)(reads->off[i1] - reads->off[i0]) * 2 + (size_t)(i1 - i0) * 160);      // SEQ + QUAL + the usual fields
	size_t maxRef = 1;                                                 // longest RNAME ("*" when there is none)
	for(uint64_t t = 0; opt->ref_names && t < opt->n_refs; t++) if(opt->ref_names[t]) { const size_t l = strlen(opt->ref_names[t]); if(l > maxRef) maxRef = l; }
	const size_t rgLen = (opt->rg_optflag && opt->rg_optflag[0]) ? strlen(opt->rg_optflag) + 1 : 0;
	std::string cigar, mdz;
	std::vector<char> held; size_t heldN = 0;
	uint64_t nTrunc = 0;
	Stacked st;                                                        // buffers reused from record to record
	for(uint64_t i = i0; i < i1; i++) {
		const bt2g_read_result &r = res[i];
		const uint8_t *seq = reads->seq + reads->off[i];
		const uint8_t *qual = reads->qual ? reads->qual + reads->off[i] : nullptr;
		const int len = (int)(reads->off[i + 1] - reads->off[i]);
		const bool aligned = (r.found & 0xff) != 0;
		const bool fw = r.fw != 0;
		// ---- mate bookkeeping
		const bt2g_read_result *m = paired ? &res[i ^ 1ull] : nullptr;
		const bool mateAligned = m && (m->found & 0xff) != 0;
		const bt2g_pair_result *pr = paired ? &pairs[i >> 1] : nullptr;
		const bool concordant = pr && pr->pair_type == 1;
		// discordant = no concordant pair and exactly one alignment for each mate (AlnSinkWrap::getReport,
		// aln_sink.cpp:240-256); with more, the mates are reported as unpaired alignments of a paired read
		const bool discordant = pr && !noDiscord && pr->pair_type == 2 && r.score2 == INT32_MIN && m->score2 == INT32_MIN;
		const bool asPair = concordant || discordant;
		int flag = (r.found & 0x100) ? 256 : 0;                   // caller-marked secondary alignment (-k / -a records)
		if(paired) {
			flag |= 1;
			if(concordant) flag |= 2;
			if(!aligned) flag |= 4;
			if(!mateAligned) flag |= 8;
			if(aligned && !fw) flag |= 16;
			if(mateAligned && m->fw == 0) flag |= 32;
			flag |= (i & 1) ? 128 : 64;
		} else {
			if(!aligned) flag |= 4; else if(!fw) flag |= 16;
		}
		// ---- stacked alignment, CIGAR, MD:Z, edit counts
		cigar.clear(); mdz.clear();
		long long refExtent = 0;
		int nmm = 0, ngo = 0, ngx = 0, nedits = 0;
		if(aligned) {
			const int nops = (r.found & 0xff) == 2 ? 0 : (r.nops > (int)maxOps ? (int)maxOps : r.nops);   // never read past the ops row
			if((r.found & 0xff) != 2 && r.nops > (int)maxOps) nTrunc++;                                   // (its CIGAR / MD:Z lost their head)
			const uint8_t *op = ops ? ops + i * (uint64_t)maxOps : nullptr;
			bool gapless = true;
			if((r.found & 0xff) != 2) {
				if(!op) return -1;
				// (8 ops per step: bit 1 of an op's type is set exactly for the two gap kinds)
				int k = 0;
				for(; k + 8 <= nops; k += 8) { uint64_t w; memcpy(&w, op + k, 8); if(w & 0x0202020202020202ull) { gapless = false; break; } }
				for(; gapless && k < nops; k++) if((op[k] & 3) >= BT2G_OP_REFGAP) { gapless = false; break; }
			}
			if(gapless) {
				// no gaps: nothing to left-align; CIGAR is one M run and MD:Z a scan of the mismatches
				const int nrow = (r.found & 0xff) == 2 ? len : nops;
				if(r.trim_left > 0) { appendInt(cigar, r.trim_left); cigar += 'S'; }
				if(!xeq || nops == 0) { appendInt(cigar, nrow); cigar += xeq ? '=' : 'M'; }
				else {
					// --xeq: runs of = / X (StackedAln::buildCigar(true))
					for(int k = nops - 1; k >= 0;) {
						const bool mm = (op[k] & 3) == BT2G_OP_MM;
						int run = 1;
						while(k - run >= 0 && (((op[k - run] & 3) == BT2G_OP_MM) == mm)) run++;
						appendInt(cigar, run); cigar += mm ? 'X' : '=';
						k -= run;
					}
				}
				if(r.trim_right > 0) { appendInt(cigar, r.trim_right); cigar += 'S'; }
				int run = 0; bool mmLast = false, first = true;
				for(int k = nops - 1; k >= 0; k--) {
					// a stretch of matches (type 0 in both low bits), eight ops at a time
					while(k >= 7) { uint64_t w; memcpy(&w, op + k - 7, 8); if(w & 0x0303030303030303ull) break; run += 8; k -= 8; }
					if(k < 0) break;
					if((op[k] & 3) == BT2G_OP_MM) {
						if(run > 0) { appendInt(mdz, run); first = false; mmLast = false; run = 0; }
						if(mmLast || first) mdz += '0';
						const int rc = (op[k] >> 2) & 7;
						mdz += dna[rc > 4 ? 4 : rc]; first = false; mmLast = true;
						nmm++; nedits++;
					} else run++;
				}
				if((r.found & 0xff) == 2) run = len;
				if(run > 0) { appendInt(mdz, run); mmLast = false; }
				if(mmLast) mdz += '0';
				refExtent = nrow;
			} else {
			st.ref.clear(); st.rel.clear(); st.rd.clear();
			// read characters in reference orientation
			// (rows past the read -- a result row that does not belong to this read -- read as N instead of reading past the batch)
			auto rdch = [&](int row) -> char { if(row < 0 || row >= len) return 'N'; return fw ? dna[seq[row] > 4 ? 4 : seq[row]] : comp[seq[len - 1 - row] > 4 ? 4 : seq[len - 1 - row]]; };
			int row = r.trim_left;
			if((r.found & 0xff) == 2) {
				for(int k = 0; k < len; k++) { st.ref.push_back(rdch(k)); st.rel.push_back('='); st.rd.push_back(rdch(k)); }
			} else {
				if(!op) return -1;
				int prevGap = -1;                                  // 2 read gap, 1 ref gap run tracking for XO/XG
				for(int k = nops - 1; k >= 0; k--) {
					const int typ = op[k] & 3, rc = (op[k] >> 2) & 7;
					if(typ == BT2G_OP_MATCH) { st.ref.push_back(dna[rc > 4 ? 4 : rc]); st.rel.push_back('='); st.rd.push_back(rdch(row)); row++; prevGap = -1; }
					else if(typ == BT2G_OP_MM) { st.ref.push_back(dna[rc > 4 ? 4 : rc]); st.rel.push_back('X'); st.rd.push_back(rdch(row)); row++; nmm++; nedits++; prevGap = -1; }
					else if(typ == BT2G_OP_REFGAP) { st.ref.push_back('-'); st.rel.push_back('I'); st.rd.push_back(rdch(row)); row++; nedits++; ngx++; if(prevGap != 1) ngo++; prevGap = 1; }
					else { st.ref.push_back(dna[rc > 4 ? 4 : rc]); st.rel.push_back('D'); st.rd.push_back('-'); nedits++; ngx++; if(prevGap != 2) ngo++; prevGap = 2; }
				}
			}
			st.leftAlign();
			// CIGAR (StackedAln::buildCigar, xeq = false)
			if(r.trim_left > 0) { appendInt(cigar, r.trim_left); cigar += 'S'; }
			const size_t ln = st.rel.size();
			for(size_t k = 0; k < ln;) {
				char c = st.rel[k]; if(!xeq && (c == 'X' || c == '=')) c = 'M';
				size_t run = 1;
				while(k + run < ln) { char c2 = st.rel[k + run]; if(!xeq && (c2 == 'X' || c2 == '=')) c2 = 'M'; if(c2 != c) break; run++; }
				appendInt(cigar, (long long)run); cigar += c;
				k += run;
			}
			if(r.trim_right > 0) { appendInt(cigar, r.trim_right); cigar += 'S'; }
			// MD:Z (StackedAln::buildMdz + writeMdz)
			bool mmLast = false, gapLast = false, first = true;
			for(size_t k = 0; k < ln; k++) {
				const char c = st.rel[k];
				if(c == '=') {
					size_t run = 1, nins = 0;
					for(; k + run < ln; run++) { if(st.rel[k + run] == '=') {} else if(st.rel[k + run] == 'I') nins++; else break; }
					k += run - 1;
					if(run - nins > 0) { appendInt(mdz, (long long)(run - nins)); first = false; mmLast = false; gapLast = false; }
				} else if(c == 'X') {
					if(gapLast || mmLast || first) mdz += '0';
					mdz += st.ref[k]; first = false; mmLast = true; gapLast = false;
				} else if(c == 'D') {
					if(mmLast || first) mdz += '0';
					if(!gapLast) mdz += '^';
					mdz += st.ref[k]; first = false; mmLast = false; gapLast = true;
				}
			}
			if(mmLast || gapLast) mdz += '0';
			for(size_t k = 0; k < ln; k++) refExtent += st.rel[k] != 'I';
			}
		}
		// ---- the record, written in place: one capacity check, then plain stores
		if(noUnal && !aligned) continue;                              // --no-unal (AlnSinkSam::appendMate, aln_sink.cpp:1905)
		if(r.found & 0x200) continue;                                 // present only as its mate's mate (paired -k / -a entries)
		const char *nm = (opt->read_names && opt->read_names[i]) ? opt->read_names[i] : nullptr;
		size_t nml = 0;
		if(nm) {
			// QNAME = the name up to the first whitespace (sam.h printReadName), without a /1 or /2 mate suffix
			while(nm[nml] && nm[nml] != ' ' && nm[nml] != '\t') nml++;
			if(paired && nml >= 2 && nm[nml - 2] == '/' && (nm[nml - 1] == '1' || nm[nml - 1] == '2')) nml -= 2;
		}
		const size_t bound = nml + 2 * maxRef + cigar.size() + mdz.size() + 2 * (size_t)len + rgLen + 512;
		// a pair with only mate 2 aligned is printed aligned mate first (AlnSinkWrap::finishRead reports the
		// unpaired alignment of mate 2, then the unaligned mate 1, aln_sink.cpp:930-1010): mate 1's record waits
		// (not when mate 2's row is there only as mate context, paired -k / -a entries: nothing of this pair follows to wait for)
		const bool hold = paired && (i & 1) == 0 && !aligned && mateAligned && !(m->found & 0x200);
		if(hold && held.size() < bound) held.resize(bound);
		char *const rec = hold ? held.data() : o.room(bound + heldN);
		char *p = rec;
		if(nm) { memcpy(p, nm, nml); p += nml; }
		else { *p++ = 'r'; p = putInt(p, (long long)(paired ? (i >> 1) : i)); }
		*p++ = '\t'; p = putInt(p, flag); *p++ = '\t';
		auto refName = [&](uint64_t t) -> const char * { return (opt->ref_names && t < opt->n_refs && opt->ref_names[t]) ? opt->ref_names[t] : "*"; };
		if(aligned) {
			p = putStr(p, refName(r.tidx)); *p++ = '\t'; p = putInt(p, r.refoff + 1); *p++ = '\t'; p = putInt(p, r.mapq); *p++ = '\t';
			memcpy(p, cigar.data(), cigar.size()); p += cigar.size();
		}
		else if(paired && mateAligned) { p = putStr(p, refName(m->tidx)); *p++ = '\t'; p = putInt(p, m->refoff + 1); p = PUT_LIT(p, "\t0\t*"); }   // unaligned mate takes its mate's coordinates
		else p = PUT_LIT(p, "*\t0\t0\t*");
		*p++ = '\t';
		// RNEXT PNEXT TLEN
		if(paired && mateAligned) {
			if(aligned && m->tidx == r.tidx) *p++ = '='; else if(!aligned) *p++ = '='; else p = putStr(p, refName(m->tidx));
			*p++ = '\t'; p = putInt(p, m->refoff + 1); *p++ = '\t';
			long long tlen = 0;
			if(aligned && asPair && m->tidx == r.tidx) {
				// AlnRes::setFragmentLength (aligner_result.h:1311-1343); extents include soft-trimmed ends
				long long st0 = r.refoff - r.trim_left, en0 = r.refoff + refExtent - 1 + r.trim_right;
				// the mate's extent needs its own ops: every op but a reference gap (type 2 = bit 1 set, bit 0 clear) consumes a reference character
				long long mExt = 0;
				const int mlen = (int)(reads->off[(i ^ 1ull) + 1] - reads->off[i ^ 1ull]);
				if((m->found & 0xff) == 2) mExt = mlen;
				else if(ops) {
					const uint8_t *mo = ops + (i ^ 1ull) * (uint64_t)maxOps;
					const int mn = m->nops < (int)maxOps ? m->nops : (int)maxOps;
					int k = 0, gaps = 0;
					for(; k + 8 <= mn; k += 8) { uint64_t w; memcpy(&w, mo + k, 8); gaps += __builtin_popcountll((w >> 1) & ~w & 0x0101010101010101ull); }
					for(; k < mn; k++) gaps += (mo[k] & 3) == BT2G_OP_REFGAP;
					mExt = (mn > 0 ? mn : 0) - gaps;
				}
				long long st1 = m->refoff - m->trim_left, en1 = m->refoff + mExt - 1 + m->trim_right;
				bool up;
				const bool mfw = m->fw != 0, mate1 = (i & 1) == 0;
				if(st0 == st1) up = (fw && mfw && mate1) || (fw && !mfw);
				else up = st0 < st1;
				tlen = 1 + (en0 > en1 ? en0 : en1) - (st0 < st1 ? st0 : st1);
				if(!up) tlen = -tlen;
			}
			p = putInt(p, tlen);
		} else if(paired && aligned) { p = PUT_LIT(p, "=\t"); p = putInt(p, r.refoff + 1); p = PUT_LIT(p, "\t0"); }       // mate unaligned: points at this mate
		else p = PUT_LIT(p, "*\t0\t0");
		*p++ = '\t';
		// SEQ QUAL (reverse-complemented / reversed for the reverse strand)
		const bool rev = aligned && !fw;
		if(len == 0) p = PUT_LIT(p, "*\t*");                         // an empty (fully trimmed) read
		else {
			char *d = p;
			if(rev) for(int k = 0; k < len; k++) d[k] = chars.rc[seq[len - 1 - k]];
			else    for(int k = 0; k < len; k++) d[k] = chars.fw[seq[k]];
			d[len] = '\t';
			char *e = d + len + 1;
			if(!qual) { e[0] = '*'; p = e + 1; }
			else {
				if(rev) {
					int k = 0;
					for(; k + 8 <= len; k += 8) { uint64_t w; memcpy(&w, qual + len - 8 - k, 8); w = __builtin_bswap64(w); memcpy(e + k, &w, 8); }
					for(; k < len; k++) e[k] = (char)qual[len - 1 - k];
				}
				else memcpy(e, qual, (size_t)len);
				p = e + len;
			}
		}
		// optional fields
		if(aligned) {
			p = PUT_LIT(p, "\tAS:i:"); p = putInt(p, r.score);
			// XS:i of a paired read is the best unchosen PAIRED score of this mate (sam.cpp:146-158): never set for
			// mates reported as unpaired alignments
			if(r.score2 != INT32_MIN && (!paired || concordant)) { p = PUT_LIT(p, "\tXS:i:"); p = putInt(p, r.score2); }
			p = PUT_LIT(p, "\tXN:i:"); p = putInt(p, r.pad);
			p = PUT_LIT(p, "\tXM:i:"); p = putInt(p, nmm);
			p = PUT_LIT(p, "\tXO:i:"); p = putInt(p, ngo);
			p = PUT_LIT(p, "\tXG:i:"); p = putInt(p, ngx);
			p = PUT_LIT(p, "\tNM:i:"); p = putInt(p, nedits);
			p = PUT_LIT(p, "\tMD:Z:"); memcpy(p, mdz.data(), mdz.size()); p += mdz.size();
			// YS:i only for mates reported as a pair (summ.paired(), sam.cpp:250)
			if(paired && mateAligned && asPair) { p = PUT_LIT(p, "\tYS:i:"); p = putInt(p, m->score); }
		}
		p = PUT_LIT(p, "\tYT:Z:");
		{ const char *yt = !paired ? "UU" : (concordant ? "CP" : (discordant ? "DP" : "UP")); p[0] = yt[0]; p[1] = yt[1]; p += 2; }
		if(!aligned) {
			// YF:Z: why the read was filtered out (sam.cpp:331-345; filters at bt2_search.cpp:3405-3431)
			int ns = 0;
			for(int k = 0; k < len; k++) ns += seq[k] > 3;
			const double nc = (opt->nceil_const == 0.0 && opt->nceil_linear == 0.0) ? 0.0 : opt->nceil_const;
			const double nl = (opt->nceil_const == 0.0 && opt->nceil_linear == 0.0) ? (double)0.15f : opt->nceil_linear;
			int nceil = (int)(nc + nl * (double)len); if(nceil > len) nceil = len;
			if(len < 2) p = PUT_LIT(p, "\tYF:Z:LN");
			else if(ns > nceil) p = PUT_LIT(p, "\tYF:Z:NS");
			else if(len <= opt->sc_filter_maxlen) p = PUT_LIT(p, "\tYF:Z:SC");      // perfect score below the minimum (scoreFilter)
		}
		if(rgLen) { *p++ = '\t'; memcpy(p, opt->rg_optflag, rgLen - 1); p += rgLen - 1; }     // RG:Z:<id> (sam.cpp:384-387)
		*p++ = '\n';
		if(hold) { heldN = (size_t)(p - rec); continue; }
		if(heldN) { memcpy(p, held.data(), heldN); p += heldN; heldN = 0; }
		o.n += (size_t)(p - rec);
	}
	return nTrunc ? 1 : 0;
}

extern "C" int bt2g_sam_format(const bt2g_sam_opts *opt, const bt2g_reads *reads, const bt2g_read_result *res, const uint8_t *ops,
                               uint32_t maxOps, const bt2g_pair_result *pairs, char *out, uint64_t cap, uint64_t *written) {
	if(!opt || !reads || !res || !written) return -1;
	const uint64_t n = reads->n_reads;
	const bool paired = pairs != nullptr;
	if(paired && (n & 1)) return -1;
	// records are independent (pairs stay together): format contiguous ranges on opt->threads host threads
	int T = opt->threads > 0 ? opt->threads : 1;
	const uint64_t units = paired ? n / 2 : n;
	if((uint64_t)T > units) T = units ? (int)units : 1;
	std::vector<OutBuf> parts((size_t)T);
	std::vector<int> rcs((size_t)T, 0);
	auto work = [&](int t) {
		const uint64_t u0 = units * (uint64_t)t / T, u1 = units * (uint64_t)(t + 1) / T;
		rcs[t] = formatRange(opt, reads, res, ops, maxOps, pairs, paired ? 2 * u0 : u0, paired ? 2 * u1 : u1, parts[t]);
	};
	if(T == 1) work(0);
	else {
		std::vector<std::thread> th;
		for(int t = 0; t < T; t++) th.emplace_back(work, t);
		for(auto &x : th) x.join();
	}
	uint64_t total = 0;
	bool truncated = false;
	for(int t = 0; t < T; t++) { if(rcs[t] < 0) return rcs[t]; truncated = truncated || rcs[t] > 0; total += parts[t].n; }
	*written = total;
	if(total > cap || (!out && total)) return -3;                  // buffer too small: *written holds the size needed
	std::vector<uint64_t> at((size_t)T + 1, 0);
	for(int t = 0; t < T; t++) at[t + 1] = at[t] + parts[t].n;
	if(T == 1) { if(parts[0].n) memcpy(out, parts[0].blk.get(), parts[0].n); }
	else {
		std::vector<std::thread> th;
		for(int t = 0; t < T; t++) th.emplace_back([&, t]() { if(parts[t].n) memcpy(out + at[t], parts[t].blk.get(), parts[t].n); });
		for(auto &x : th) x.join();
	}
	return truncated ? 1 : 0;                                      // 1: a record had more edit ops than max_ops (text complete, that CIGAR is not)
}

// ---- FASTQ text -> bt2g_reads buffers (host) ----------------------------------------------------------------
// FastqPatternSource::parse (pat.cpp:1130-1245) for plain 4-line records without trimming: the name is the header
// line after '@'; sequence characters are letters ('.' = N), A/C/G/T in either case map to 0..3 and every other
// letter to 4 (alphabet.cpp asc2dna); the '+' line is skipped; qualities are kept as raw Phred+33 bytes and must
// number exactly the bases (tooFewQualities / tooManyQualities, pat.cpp:1226-1232).
// recStart (optional, maxReads + 1 entries): text offset at which each parsed record starts, then the end offset
static int fastqParseCore(const char *text, uint64_t len, uint64_t maxReads, uint64_t maxBases, uint8_t *seq, uint8_t *qual,
                          uint64_t *off, char *names, uint32_t nameStride, uint64_t *nReads, uint64_t *consumed, uint64_t *recStarts) {
	if(!text || !seq || !qual || !off || !nReads || !consumed) return -1;
	uint64_t cur = 0, n = 0, nb = 0;
	off[0] = 0;
	*nReads = 0; *consumed = 0;
	// base codes of the sequence characters (alphabet.cpp asc2dna): letters -> 0..4, '.' -> 4, anything else 0xff (not a sequence character)
	static const struct Lut { uint8_t v[256]; Lut() { for(int c = 0; c < 256; c++) { const int u = c & ~0x20; v[c] = ((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z')) ? (uint8_t)(u == 'A' ? 0 : u == 'C' ? 1 : u == 'G' ? 2 : u == 'T' ? 3 : 4) : (c == '.' ? 4 : 0xff); } } } lut;
	while(cur < len && n < maxReads) {
		while(cur < len && (text[cur] == '\n' || text[cur] == '\r')) cur++;
		if(cur >= len) break;
		const uint64_t recStart = cur;
		if(recStarts) recStarts[n] = recStart;
		if(text[cur] != '@') return -4;                           // not a FASTQ record
		{
			// fast path for the usual record: four '\n'-terminated lines, one sequence line of sequence characters only, as many
			// qualities; everything else (\r line ends, wrapped sequences, odd characters, truncation, limits) takes the general code below
			const char *p = text + cur, *end = text + len;
			const char *nl1 = (const char *)memchr(p, '\n', (size_t)(end - p));
			const char *nl2 = nl1 ? (const char *)memchr(nl1 + 1, '\n', (size_t)(end - nl1 - 1)) : nullptr;
			if(nl2 && nl2 + 1 < end && nl2[1] == '+' && nl1[-1] != '\r' && nl2 > nl1 + 1 && nl2[-1] != '\r') {
				const char *nl3 = (const char *)memchr(nl2 + 1, '\n', (size_t)(end - nl2 - 1));
				const char *nl4 = nl3 ? (const char *)memchr(nl3 + 1, '\n', (size_t)(end - nl3 - 1)) : nullptr;
				const uint64_t L = (uint64_t)(nl2 - nl1 - 1);
				if(nl4 && (uint64_t)(nl4 - nl3 - 1) == L && nb + L <= maxBases && nl4[-1] != '\r' && nl3[-1] != '\r' && !memchr(nl3 + 1, ' ', (size_t)L)) {
					const unsigned char *sq = (const unsigned char *)nl1 + 1;
					uint8_t *d = seq + nb;
					uint8_t bad = 0;
					for(uint64_t k = 0; k < L; k++) { const uint8_t v = lut.v[sq[k]]; d[k] = v; bad |= v; }
					if(!(bad & 0x80)) {
						memcpy(qual + nb, nl3 + 1, (size_t)L);
						if(names && nameStride) {
							uint64_t l = (uint64_t)(nl1 - p - 1);
							if(l >= nameStride) l = nameStride - 1;
							memcpy(names + n * (uint64_t)nameStride, p + 1, l);
							memset(names + n * (uint64_t)nameStride + l, 0, nameStride - l);   // the whole row is defined
						}
						nb += L; n++; off[n] = nb;
						cur = (uint64_t)(nl4 + 1 - text);
						while(cur < len && (text[cur] == '\n' || text[cur] == '\r')) cur++;     // blank lines between records (as the general path below)
						continue;
					}
				}
			}
		}
		cur++;
		const uint64_t nameBeg = cur;
		while(cur < len && text[cur] != '\n' && text[cur] != '\r') cur++;
		const uint64_t nameEnd = cur;
		while(cur < len && (text[cur] == '\n' || text[cur] == '\r')) cur++;
		// sequence up to the '+' line
		const uint64_t b0 = nb;
		bool full = false;
		while(cur < len && text[cur] != '+') {
			int c = (unsigned char)text[cur++];
			if(c == '.') c = 'N';
			if((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z')) {
				if(nb >= maxBases) { full = true; break; }
				const int u = c & ~0x20;
				seq[nb++] = (uint8_t)(u == 'A' ? 0 : u == 'C' ? 1 : u == 'G' ? 2 : u == 'T' ? 3 : 4);
			}
		}
		if(full) { nb = b0; cur = recStart; break; }              // out of room: stop before this record
		if(cur >= len) { nb = b0; cur = recStart; break; }        // truncated record: leave it for the next call
		while(cur < len && text[cur] != '\n' && text[cur] != '\r') cur++;   // the '+' line
		// exactly ONE line terminator: an empty quality line (an empty read, which the reference reports as unaligned, YF:Z:LN)
		// must not be skipped
		if(cur < len && text[cur] == '\r') cur++;
		if(cur < len && text[cur] == '\n') cur++;
		const uint64_t nbases = nb - b0;
		uint64_t nq = 0;
		while(cur < len && text[cur] != '\n' && text[cur] != '\r') {
			const unsigned char c = (unsigned char)text[cur++];
			if(c == ' ') return -5;                                // integer qualities are not supported
			if(nq < nbases) qual[b0 + nq] = c;
			nq++;
		}
		if(nq < nbases) {
			if(cur >= len) { nb = b0; cur = recStart; break; }    // truncated quality line
			return -6;                                             // fewer qualities than bases
		}
		if(nq > nbases) return -7;                                 // more qualities than bases
		if(names && nameStride) {
			uint64_t l = nameEnd - nameBeg;
			if(l >= nameStride) l = nameStride - 1;
			memcpy(names + n * (uint64_t)nameStride, text + nameBeg, l);
			memset(names + n * (uint64_t)nameStride + l, 0, nameStride - l);
		}
		n++;
		off[n] = nb;
		while(cur < len && (text[cur] == '\n' || text[cur] == '\r')) cur++;
	}
	*nReads = n; *consumed = cur;
	if(recStarts) recStarts[n] = cur;
	return 0;
}

extern "C" int bt2g_fastq_parse(const char *text, uint64_t len, uint64_t maxReads, uint64_t maxBases, uint8_t *seq, uint8_t *qual,
                                uint64_t *off, char *names, uint32_t nameStride, uint64_t *nReads, uint64_t *consumed) {
	return fastqParseCore(text, len, maxReads, maxBases, seq, qual, off, names, nameStride, nReads, consumed, nullptr);
}

// Multi-threaded variant: the text is cut at record boundaries (a line starting with '@' whose second-next line starts with
// '+'), the pieces are parsed concurrently into pooled scratch and copied to their places concurrently.  Same outputs and error
// codes as bt2g_fastq_parse for 4-line FASTQ (multi-line sequences fall back to the serial parser).
namespace {
struct Piece {
	Block blk; uint8_t *seq = nullptr, *qual = nullptr; uint64_t *off = nullptr, *starts = nullptr; char *names = nullptr;
	uint64_t t0 = 0, tl = 0, n = 0, used = 0, base = 0, take = 0; int rc = 0;
};
// cut points of a FASTQ text for `pieces` concurrent parsers (a line starting with '@' whose second-next line starts with '+')
void cutFastq(const char *text, uint64_t len, int pieces, std::vector<uint64_t> &cuts) {
	auto lineEnd = [&](uint64_t p) { const char *q = (const char *)memchr(text + p, '\n', (size_t)(len - p)); return q ? (uint64_t)(q - text) + 1 : len; };
	cuts.assign(1, 0);
	if(len >= (1u << 20)) for(int k = 1; k < pieces; k++) {
		uint64_t p = len / (uint64_t)pieces * (uint64_t)k;
		if(p <= cuts.back()) continue;
		p = lineEnd(p);
		bool found = false;
		for(int tries = 0; tries < 8 && p < len; tries++) {
			if(text[p] == '@') {
				const uint64_t l1 = lineEnd(p), l2 = lineEnd(l1);
				if(l2 < len && text[l2] == '+') { found = true; break; }
			}
			p = lineEnd(p);
		}
		if(found && p > cuts.back()) cuts.push_back(p);
	}
	cuts.push_back(len);
}
void parsePiece(Piece &q, const char *text, uint32_t nameStride, bool wantNames) {
	uint64_t nl = 0;
	for(const char *p = text + q.t0, *e = p + q.tl; (p = (const char *)memchr(p, '\n', (size_t)(e - p))) != nullptr; p++) nl++;
	const uint64_t cap = nl / 4 + 2;                               // >= records: each has at least four lines
	const uint64_t data = (q.tl + 15) & ~15ull, idx = (cap + 1) * 8;
	q.blk = pool().take((size_t)(2 * data + 2 * idx + (wantNames ? cap * (uint64_t)nameStride : 0) + 64));
	q.seq = q.blk.get(); q.qual = q.seq + data;
	q.off = (uint64_t *)(q.qual + data); q.starts = q.off + cap + 1;
	q.names = wantNames ? (char *)(q.starts + cap + 1) : nullptr;
	q.rc = fastqParseCore(text + q.t0, q.tl, cap, q.tl, q.seq, q.qual, q.off, q.names, nameStride, &q.n, &q.used, q.starts);
}
} // namespace

extern "C" int bt2g_fastq_parse_mt(const char *text, uint64_t len, uint64_t maxReads, uint64_t maxBases, uint8_t *seq, uint8_t *qual,
                                   uint64_t *off, char *names, uint32_t nameStride, uint64_t *nReads, uint64_t *consumed, int threads) {
	if(threads <= 1 || len < (1u << 20))
		return fastqParseCore(text, len, maxReads, maxBases, seq, qual, off, names, nameStride, nReads, consumed, nullptr);
	if(!text || !seq || !qual || !off || !nReads || !consumed) return -1;
	std::vector<uint64_t> cuts;
	cutFastq(text, len, threads, cuts);
	const size_t nc = cuts.size() - 1;
	if(nc < 2) return fastqParseCore(text, len, maxReads, maxBases, seq, qual, off, names, nameStride, nReads, consumed, nullptr);
	const bool wantNames = names && nameStride;
	std::vector<Piece> pc(nc);
	struct GiveBack { std::vector<Piece> &pc; ~GiveBack() { for(auto &q : pc) pool().give(std::move(q.blk)); } } giveBack{pc};
	{
		std::vector<std::thread> th;
		for(size_t k = 0; k < nc; k++) { pc[k].t0 = cuts[k]; pc[k].tl = cuts[k + 1] - cuts[k]; th.emplace_back([&, k]() { parsePiece(pc[k], text, nameStride, wantNames); }); }
		for(auto &t : th) t.join();
	}
	// ---- what each piece contributes while the limits allow; a piece that did not end on its boundary (a truncated last record) ends the parse
	uint64_t n = 0, nb = 0, cur = 0;
	std::vector<uint64_t> nb0(nc, 0);
	off[0] = 0;
	*nReads = 0; *consumed = 0;
	for(size_t k = 0; k < nc; k++) {
		Piece &q = pc[k];
		if(q.rc) return q.rc;
		uint64_t take = q.n;
		if(n + take > maxReads) take = maxReads - n;
		while(take > 0 && nb + q.off[take] > maxBases) take--;
		q.base = n; q.take = take; nb0[k] = nb;
		n += take; nb += q.off[take];
		cur = q.t0 + (take == q.n ? q.used : q.starts[take]);
		if(take < q.n || q.used < q.tl) break;                     // limit reached, or a truncated record at the end of the piece
	}
	{
		std::vector<std::thread> th;
		for(size_t k = 0; k < nc; k++) if(pc[k].take) th.emplace_back([&, k]() {
			const Piece &q = pc[k];
			const uint64_t bases = q.off[q.take];
			memcpy(seq + nb0[k], q.seq, (size_t)bases);
			memcpy(qual + nb0[k], q.qual, (size_t)bases);
			for(uint64_t i = 1; i <= q.take; i++) off[q.base + i] = nb0[k] + q.off[i];
			if(wantNames) for(uint64_t j = 0; j < q.take; j++) {
				const char *src = q.names + j * (uint64_t)nameStride;
				char *dst = names + (q.base + j) * (uint64_t)nameStride;
				const size_t nl = strnlen(src, nameStride - 1);
				memcpy(dst, src, nl);
				memset(dst + nl, 0, nameStride - nl);
			}
		});
		for(auto &t : th) t.join();
	}
	*nReads = n; *consumed = cur;
	return 0;
}

// ---- two mate files -> one interleaved batch ------------------------------------------------------------------------
// DualPatternComposer::nextBatch (pat.cpp:222-300) hands the aligner mate 1 and mate 2 of pair i together; here pair i
// becomes reads 2i and 2i + 1 of the batch, the layout every paired entry point of this library takes.  Both texts are cut at
// record boundaries and their pieces parsed concurrently (fastqParseCore, into pooled scratch); the records then go straight to
// their interleaved places, one copy task per piece.
extern "C" int bt2g_fastq_parse_pairs_mt(const char *text1, uint64_t len1, const char *text2, uint64_t len2, uint64_t maxPairs, uint64_t maxBases,
                                         uint8_t *seq, uint8_t *qual, uint64_t *off, char *names, uint32_t nameStride, uint64_t *nPairs,
                                         uint64_t *consumed1, uint64_t *consumed2, int threads) {
	if(!text1 || !text2 || !seq || !qual || !off || !nPairs || !consumed1 || !consumed2) return -1;
	const int T = threads > 1 ? threads : 1;
	const bool wantNames = names && nameStride;
	const char *text[2] = {text1, text2};
	const uint64_t len[2] = {len1, len2};
	std::vector<uint64_t> cuts[2];
	std::vector<Piece> pc[2];
	std::vector<std::pair<int, size_t>> tasks;
	for(int f = 0; f < 2; f++) {
		cutFastq(text[f], len[f], T > 1 ? (T + 1) / 2 : 1, cuts[f]);
		pc[f].resize(cuts[f].size() - 1);
		for(size_t k = 0; k < pc[f].size(); k++) { pc[f][k].t0 = cuts[f][k]; pc[f][k].tl = cuts[f][k + 1] - cuts[f][k]; tasks.emplace_back(f, k); }
	}
	auto runTasks = [&](auto &&fn) {
		std::atomic<size_t> next{0};
		auto loop = [&]() { for(size_t t; (t = next.fetch_add(1)) < tasks.size();) fn(tasks[t].first, tasks[t].second); };
		const int nt = (int)std::min<size_t>((size_t)T, tasks.size());
		if(nt <= 1) { loop(); return; }
		std::vector<std::thread> th;
		for(int t = 1; t < nt; t++) th.emplace_back(loop);
		loop();
		for(auto &x : th) x.join();
	};
	struct GiveBack { std::vector<Piece> *pc; ~GiveBack() { for(int f = 0; f < 2; f++) for(auto &q : pc[f]) pool().give(std::move(q.blk)); } } giveBack{pc};
	runTasks([&](int f, size_t k) { parsePiece(pc[f][k], text[f], nameStride, wantNames); });
	// ---- records usable from each file, in order: a piece that did not end on its boundary (a truncated last record) ends them
	uint64_t avail[2] = {0, 0}; size_t usedPieces[2] = {0, 0};
	*nPairs = 0; *consumed1 = 0; *consumed2 = 0;
	off[0] = 0;
	for(int f = 0; f < 2; f++)
		for(size_t k = 0; k < pc[f].size(); k++) {
			Piece &q = pc[f][k];
			if(q.rc) return q.rc;
			q.base = avail[f]; avail[f] += q.n; usedPieces[f] = k + 1;
			if(q.used < q.tl) break;
		}
	uint64_t n = std::min(std::min(avail[0], avail[1]), maxPairs);
	// ---- offsets of the interleaved batch (serial: two additions per pair); the base limit may shorten it
	{
		size_t k[2] = {0, 0}; uint64_t nb = 0, i = 0;
		for(; i < n; i++) {
			uint64_t l[2];
			for(int f = 0; f < 2; f++) {
				while(i >= pc[f][k[f]].base + pc[f][k[f]].n) k[f]++;
				const Piece &q = pc[f][k[f]];
				const uint64_t j = i - q.base;
				l[f] = q.off[j + 1] - q.off[j];
			}
			if(nb + l[0] + l[1] > maxBases) break;
			off[2 * i + 1] = nb + l[0]; nb += l[0] + l[1]; off[2 * i + 2] = nb;
		}
		n = i;
	}
	for(int f = 0; f < 2; f++)
		for(size_t k = 0; k < usedPieces[f]; k++) { Piece &q = pc[f][k]; q.take = n > q.base ? std::min(q.n, n - q.base) : 0; }
	// ---- every piece's records to their places
	runTasks([&](int f, size_t k) {
		const Piece &q = pc[f][k];
		for(uint64_t j = 0; j < q.take; j++) {
			const uint64_t r = 2 * (q.base + j) + (uint64_t)f, l = q.off[j + 1] - q.off[j];
			memcpy(seq + off[r], q.seq + q.off[j], (size_t)l);
			memcpy(qual + off[r], q.qual + q.off[j], (size_t)l);
			if(wantNames) {
				const char *src = q.names + j * (uint64_t)nameStride;
				char *dst = names + r * (uint64_t)nameStride;
				const size_t nl = strnlen(src, nameStride - 1);
				memcpy(dst, src, nl);
				memset(dst + nl, 0, nameStride - nl);                // whole row defined: the scratch it came from is not
			}
		}
	});
	// ---- first unparsed byte of each text
	uint64_t *consumed[2] = {consumed1, consumed2};
	for(int f = 0; f < 2; f++) {
		uint64_t c = 0;
		for(size_t k = 0; k < usedPieces[f]; k++) {
			const Piece &q = pc[f][k];
			if(q.take == q.n) c = q.t0 + q.used;
			else { c = q.t0 + q.starts[q.take]; break; }
		}
		*consumed[f] = c;
	}
	*nPairs = n;
	return 0;
}


// ---- host entry points over the __host__ __device__ policy arithmetic the kernels use -------------------------
// (same source as the device code: mapq_device.cuh, pe_device.cuh; lets the CPU test suite pin it against the reference)
#define __host__
#define __device__
#define __forceinline__ inline
#include "mapq_device.cuh"
#include "pe_device.cuh"

extern "C" int bt2g_mapq(int64_t best, int hasSecbest, int64_t secbest, int64_t scMin, int64_t scPerfect, int monotone) {
	return mapq_v2(best, hasSecbest != 0, secbest, scMin, scPerfect, monotone != 0);
}

extern "C" int bt2g_frame_mate_host(const bt2g_pe_policy *pol, const bt2g_mate_anchor *anchors, uint64_t n, bt2g_mate_frame *out) {
	if(!pol || !anchors || !out || pol->pol < 1 || pol->pol > 4) return -1;
	for(uint64_t i = 0; i < n; i++) pe_frame_anchor(*pol, anchors[i], out[i]);
	return 0;
}

extern "C" int bt2g_pe_classify_host(const bt2g_pe_policy *pol, const int64_t *pairs, uint64_t n, int32_t *out) {
	if(!pol || !pairs || !out || pol->pol < 1 || pol->pol > 4) return -1;
	for(uint64_t i = 0; i < n; i++) {
		const int64_t *q = pairs + 6 * i;
		out[i] = pe_classify(*pol, q[0], (uint64_t)q[1], q[2] != 0, q[3], (uint64_t)q[4], q[5] != 0);
	}
	return 0;
}

// ---- SAM header (SamConfig::printHeader, sam.cpp:54-111) -----------------------------------------
extern "C" int bt2g_sam_header_rg(const char *const *names, const uint64_t *lens, uint64_t n, const char *rgLine, const char *pgCl,
                                  char *out, uint64_t cap, uint64_t *written);
extern "C" int bt2g_sam_header(const char *const *names, const uint64_t *lens, uint64_t n, const char *pgCl,
                               char *out, uint64_t cap, uint64_t *written) {
	return bt2g_sam_header_rg(names, lens, n, nullptr, pgCl, out, cap, written);
}

extern "C" int bt2g_sam_header_rg(const char *const *names, const uint64_t *lens, uint64_t n, const char *rgLine, const char *pgCl,
                                  char *out, uint64_t cap, uint64_t *written) {
	if(!written || (n && (!names || !lens))) return -1;
	std::string o = "@HD\tVN:1.5\tSO:unsorted\tGO:query\n";
	for(uint64_t i = 0; i < n; i++) {
		o += "@SQ\tSN:";
		for(const char *c = names[i]; c && *c && !isspace((unsigned char)*c); c++) o += *c;     // printRefName: up to the first whitespace
		o += "\tLN:"; appendInt(o, (long long)lens[i]); o += '\n';
	}
	if(rgLine && rgLine[0]) { o += "@RG\t"; o += rgLine; o += '\n'; }
	if(pgCl) { o += "@PG\tID:bowtie2\tPN:bowtie2\tVN:2.5.5\tCL:\""; o += pgCl; o += "\"\n"; }
	*written = o.size();
	if(!out || cap < o.size()) return -3;
	memcpy(out, o.data(), o.size());
	return 0;
}

// ---- alignment summary (ReportingMetrics updates of AlnSinkWrap::finishRead, aln_sink.cpp:708-1046;
//      text of AlnSink::printAlSumm, aln_sink.cpp:349-528) ---------------------------------------------
extern "C" int bt2g_align_counts_add(bt2g_align_counts *c, const bt2g_read_result *res, uint64_t nReads, const bt2g_pair_result *pairs) {
	return bt2g_align_counts_add_ex(c, res, nReads, pairs, 0);
}

extern "C" int bt2g_align_counts_add_ex(bt2g_align_counts *c, const bt2g_read_result *res, uint64_t nReads, const bt2g_pair_result *pairs, uint32_t flags) {
	if(!c || (nReads && !res)) return -1;
	const bool noDiscord = (flags & BT2G_SAM_NO_DISCORDANT) != 0;
	auto aligned = [](const bt2g_read_result &r) { return (r.found & 0xff) != 0; };
	auto multi = [](const bt2g_read_result &r) { return r.score2 != INT32_MIN; };
	if(!pairs) {
		for(uint64_t i = 0; i < nReads; i++) {
			c->nread++; c->nunpaired++;
			if(!aligned(res[i])) c->nunp_0++;
			else if(multi(res[i])) c->nunp_gt1++;
			else c->nunp_uni1++;
		}
		return 0;
	}
	if(nReads & 1) return -1;
	for(uint64_t p = 0; p < nReads / 2; p++) {
		const bt2g_read_result &a = res[2 * p], &b = res[2 * p + 1];
		const bt2g_pair_result &pr = pairs[p];
		c->nread++; c->npaired++;
		if(pr.pair_type == 1) {
			// ">1" in the reference = a second concordant PAIR was found; the pipeline keeps one pair per read, so this
			// takes "both mates have a second alignment" as the indicator (an approximation)
			if(multi(a) && multi(b)) c->nconcord_gt1++; else c->nconcord_uni1++;
			continue;
		}
		c->nconcord_0++;
		if(!noDiscord && pr.pair_type == 2 && !multi(a) && !multi(b)) { c->ndiscord++; continue; }
		for(const bt2g_read_result *m : {&a, &b}) {
			if(!aligned(*m)) c->nunp_0_0++;
			else if(multi(*m)) c->nunp_0_gt1++;
			else c->nunp_0_uni1++;
		}
	}
	return 0;
}

extern "C" int bt2g_align_summary(const bt2g_align_counts *c, int discord, int mixed, char *out, uint64_t cap, uint64_t *written) {
	if(!c || !written) return -1;
	std::string o;
	char buf[64];
	auto num = [&](uint64_t v) { appendInt(o, (long long)v); };
	auto pct = [&](uint64_t nu, uint64_t de) {
		double p = 0.0;
		if(de != 0) p = 100.0 * (double)nu / (double)de;
		snprintf(buf, sizeof(buf), "%.2f%%", p);
		o += buf;
	};
	auto line = [&](const char *indent, uint64_t v, uint64_t de, const char *tail) { o += indent; num(v); o += " ("; pct(v, de); o += ") "; o += tail; o += '\n'; };
	if(c->nread > 0) { num(c->nread); o += " reads; of these:\n"; }
	else { num(c->nread); o += " reads\n"; }
	if(c->npaired > 0) {
		line("  ", c->npaired, c->nread, "were paired; of these:");
		line("    ", c->nconcord_0, c->npaired, "aligned concordantly 0 times");
		line("    ", c->nconcord_uni1, c->npaired, "aligned concordantly exactly 1 time");
		line("    ", c->nconcord_gt1, c->npaired, "aligned concordantly >1 times");
		if(discord) {
			o += "    ----\n    "; num(c->nconcord_0); o += " pairs aligned concordantly 0 times; of these:\n";
			line("      ", c->ndiscord, c->nconcord_0, "aligned discordantly 1 time");
		}
		const uint64_t nc0 = c->nconcord_0 - c->ndiscord;
		if(mixed) {
			o += "    ----\n    "; num(nc0); o += " pairs aligned 0 times concordantly or discordantly; of these:\n";
			o += "      "; num(nc0 * 2); o += " mates make up the pairs; of these:\n";
			line("        ", c->nunp_0_0, nc0 * 2, "aligned 0 times");
			line("        ", c->nunp_0_uni1, nc0 * 2, "aligned exactly 1 time");
			line("        ", c->nunp_0_gt1, nc0 * 2, "aligned >1 times");
		}
	}
	if(c->nunpaired > 0) {
		line("  ", c->nunpaired, c->nread, "were unpaired; of these:");
		line("    ", c->nunp_0, c->nunpaired, "aligned 0 times");
		line("    ", c->nunp_uni1, c->nunpaired, "aligned exactly 1 time");
		line("    ", c->nunp_gt1, c->nunpaired, "aligned >1 times");
	}
	const uint64_t cand = c->nunpaired + c->npaired * 2;
	const uint64_t al = (c->nconcord_uni1 + c->nconcord_gt1) * 2 + c->ndiscord * 2 + c->nunp_0_uni1 + c->nunp_0_gt1 + c->nunp_uni1 + c->nunp_gt1;
	pct(al, cand); o += " overall alignment rate\n";
	*written = o.size();
	if(!out || cap < o.size()) return -3;
	memcpy(out, o.data(), o.size());
	return 0;
}
