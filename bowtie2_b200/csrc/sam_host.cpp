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
#include "../../include/bt2g.h"

namespace {

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
                       uint32_t maxOps, const bt2g_pair_result *pairs, uint64_t i0, uint64_t i1, std::string &o) {
	static const char dna[] = "ACGTN";
	static const char comp[] = "TGCAN";
	const bool paired = pairs != nullptr;
	const bool xeq = (opt->flags & BT2G_SAM_XEQ) != 0, noUnal = (opt->flags & BT2G_SAM_NO_UNAL) != 0;
	o.reserve((size_t)(i1 - i0) * 400);
	std::string cigar, mdz, line, held;
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
		const bool discordant = pr && pr->pair_type == 2 && r.score2 == INT32_MIN && m->score2 == INT32_MIN;
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
			auto rdch = [&](int row) -> char { return fw ? dna[seq[row] > 4 ? 4 : seq[row]] : comp[seq[len - 1 - row] > 4 ? 4 : seq[len - 1 - row]]; };
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
		// ---- the record
		line.clear();
		if(opt->read_names && opt->read_names[i]) {
			// QNAME = the name up to the first whitespace (sam.h printReadName), without a /1 or /2 mate suffix
			const char *nm = opt->read_names[i];
			size_t l = 0;
			while(nm[l] && nm[l] != ' ' && nm[l] != '\t') l++;
			if(paired && l >= 2 && nm[l - 2] == '/' && (nm[l - 1] == '1' || nm[l - 1] == '2')) l -= 2;
			line.append(nm, l);
		}
		else { line += 'r'; appendInt(line, (long long)(paired ? (i >> 1) : i)); }
		line += '\t'; appendInt(line, flag); line += '\t';
		auto refName = [&](uint64_t t) -> const char * { return (opt->ref_names && t < opt->n_refs && opt->ref_names[t]) ? opt->ref_names[t] : "*"; };
		if(aligned) { line += refName(r.tidx); line += '\t'; appendInt(line, r.refoff + 1); line += '\t'; appendInt(line, r.mapq); line += '\t'; line += cigar; }
		else if(paired && mateAligned) { line += refName(m->tidx); line += '\t'; appendInt(line, m->refoff + 1); line += "\t0\t*"; }   // unaligned mate takes its mate's coordinates
		else line += "*\t0\t0\t*";
		line += '\t';
		// RNEXT PNEXT TLEN
		if(paired && mateAligned) {
			if(aligned && m->tidx == r.tidx) line += '='; else if(!aligned) line += '='; else line += refName(m->tidx);
			line += '\t'; appendInt(line, m->refoff + 1); line += '\t';
			long long tlen = 0;
			if(aligned && asPair && m->tidx == r.tidx) {
				// AlnRes::setFragmentLength (aligner_result.h:1311-1343); extents include soft-trimmed ends
				long long st0 = r.refoff - r.trim_left, en0 = r.refoff + refExtent - 1 + r.trim_right;
				// the mate's extent needs its own ops
				long long mExt = 0;
				const int mlen = (int)(reads->off[(i ^ 1ull) + 1] - reads->off[i ^ 1ull]);
				if((m->found & 0xff) == 2) mExt = mlen;
				else if(ops) { const uint8_t *mo = ops + (i ^ 1ull) * (uint64_t)maxOps; for(int k = 0; k < m->nops && k < (int)maxOps; k++) mExt += (mo[k] & 3) != BT2G_OP_REFGAP; }
				long long st1 = m->refoff - m->trim_left, en1 = m->refoff + mExt - 1 + m->trim_right;
				bool up;
				const bool mfw = m->fw != 0, mate1 = (i & 1) == 0;
				if(st0 == st1) up = (fw && mfw && mate1) || (fw && !mfw);
				else up = st0 < st1;
				tlen = 1 + (en0 > en1 ? en0 : en1) - (st0 < st1 ? st0 : st1);
				if(!up) tlen = -tlen;
			}
			appendInt(line, tlen);
		} else if(paired && aligned) { line += "=\t"; appendInt(line, r.refoff + 1); line += "\t0"; }       // mate unaligned: points at this mate
		else line += "*\t0\t0";
		line += '\t';
		// SEQ QUAL (reverse-complemented / reversed for the reverse strand)
		const bool rev = aligned && !fw;
		if(len == 0) line += "*\t*";                               // an empty (fully trimmed) read
		else {
			const size_t at = line.size();
			line.resize(at + (size_t)len + 1 + (qual ? (size_t)len : 1));
			char *d = &line[at];
			if(rev) for(int k = 0; k < len; k++) { const int c = seq[len - 1 - k]; d[k] = comp[c > 4 ? 4 : c]; }
			else    for(int k = 0; k < len; k++) { const int c = seq[k]; d[k] = dna[c > 4 ? 4 : c]; }
			d[len] = '\t';
			char *e = d + len + 1;
			if(!qual) e[0] = '*';
			else if(rev) for(int k = 0; k < len; k++) e[k] = (char)qual[len - 1 - k];
			else memcpy(e, qual, (size_t)len);
		}
		// optional fields
		if(aligned) {
			line += "\tAS:i:"; appendInt(line, r.score);
			// XS:i of a paired read is the best unchosen PAIRED score of this mate (sam.cpp:146-158): never set for
			// mates reported as unpaired alignments
			if(r.score2 != INT32_MIN && (!paired || concordant)) { line += "\tXS:i:"; appendInt(line, r.score2); }
			line += "\tXN:i:"; appendInt(line, r.pad);
			line += "\tXM:i:"; appendInt(line, nmm);
			line += "\tXO:i:"; appendInt(line, ngo);
			line += "\tXG:i:"; appendInt(line, ngx);
			line += "\tNM:i:"; appendInt(line, nedits);
			line += "\tMD:Z:"; line += mdz;
			// YS:i only for mates reported as a pair (summ.paired(), sam.cpp:250)
			if(paired && mateAligned && asPair) { line += "\tYS:i:"; appendInt(line, m->score); }
		}
		line += "\tYT:Z:";
		line += !paired ? "UU" : (concordant ? "CP" : (discordant ? "DP" : "UP"));
		if(!aligned) {
			// YF:Z: why the read was filtered out (sam.cpp:331-345; filters at bt2_search.cpp:3405-3431)
			int ns = 0;
			for(int k = 0; k < len; k++) ns += seq[k] > 3;
			const double nc = (opt->nceil_const == 0.0 && opt->nceil_linear == 0.0) ? 0.0 : opt->nceil_const;
			const double nl = (opt->nceil_const == 0.0 && opt->nceil_linear == 0.0) ? (double)0.15f : opt->nceil_linear;
			int nceil = (int)(nc + nl * (double)len); if(nceil > len) nceil = len;
			if(len < 2) line += "\tYF:Z:LN";
			else if(ns > nceil) line += "\tYF:Z:NS";
			else if(len <= opt->sc_filter_maxlen) line += "\tYF:Z:SC";      // perfect score below the minimum (scoreFilter)
		}
		if(opt->rg_optflag && opt->rg_optflag[0]) { line += '\t'; line += opt->rg_optflag; }     // RG:Z:<id> (sam.cpp:384-387)
		line += '\n';
		if(noUnal && !aligned) continue;                              // --no-unal (AlnSinkSam::appendMate, aln_sink.cpp:1905)
		if(r.found & 0x200) continue;                                 // present only as its mate's mate (paired -k / -a entries)
		// a pair with only mate 2 aligned is printed aligned mate first (AlnSinkWrap::finishRead reports the
		// unpaired alignment of mate 2, then the unaligned mate 1, aln_sink.cpp:930-1010)
		if(paired && (i & 1) == 0 && !aligned && mateAligned) { held = line; continue; }
		o += line;
		if(!held.empty()) { o += held; held.clear(); }
	}
	return 0;
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
	std::vector<std::string> parts((size_t)T);
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
	for(int t = 0; t < T; t++) { if(rcs[t]) return rcs[t]; total += parts[t].size(); }
	*written = total;
	if(!out || total > cap) return -3;                             // buffer too small: *written holds the size needed
	std::vector<uint64_t> at((size_t)T + 1, 0);
	for(int t = 0; t < T; t++) at[t + 1] = at[t] + parts[t].size();
	if(T == 1) memcpy(out, parts[0].data(), parts[0].size());
	else {
		std::vector<std::thread> th;
		for(int t = 0; t < T; t++) th.emplace_back([&, t]() { memcpy(out + at[t], parts[t].data(), parts[t].size()); });
		for(auto &x : th) x.join();
	}
	return 0;
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
							names[n * (uint64_t)nameStride + l] = 0;
						}
						nb += L; n++; off[n] = nb;
						cur = (uint64_t)(nl4 + 1 - text);
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
			names[n * (uint64_t)nameStride + l] = 0;
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
// '+'), the pieces are parsed concurrently into private buffers and concatenated in order.  Same outputs and error codes as
// bt2g_fastq_parse for 4-line FASTQ (multi-line sequences fall back to the serial parser).
extern "C" int bt2g_fastq_parse_mt(const char *text, uint64_t len, uint64_t maxReads, uint64_t maxBases, uint8_t *seq, uint8_t *qual,
                                   uint64_t *off, char *names, uint32_t nameStride, uint64_t *nReads, uint64_t *consumed, int threads) {
	if(threads <= 1 || len < (1u << 20))
		return fastqParseCore(text, len, maxReads, maxBases, seq, qual, off, names, nameStride, nReads, consumed, nullptr);
	if(!text || !seq || !qual || !off || !nReads || !consumed) return -1;
	// ---- cut points
	auto lineEnd = [&](uint64_t p) { while(p < len && text[p] != '\n') p++; return p < len ? p + 1 : len; };
	std::vector<uint64_t> cuts{0};
	for(int k = 1; k < threads; k++) {
		uint64_t p = len / (uint64_t)threads * (uint64_t)k;
		if(p <= cuts.back()) continue;
		p = lineEnd(p);                                            // first full line after the target
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
	const size_t nc = cuts.size() - 1;
	if(nc < 2) return fastqParseCore(text, len, maxReads, maxBases, seq, qual, off, names, nameStride, nReads, consumed, nullptr);
	// private buffers are left uninitialised (new[]): zero-filling them would cost more than the parse
	struct Piece { std::unique_ptr<uint8_t[]> seq, qual; std::unique_ptr<uint64_t[]> off, starts; std::unique_ptr<char[]> names;
	               uint64_t n = 0, used = 0; int rc = 0; };
	std::vector<Piece> pc(nc);
	std::vector<std::thread> th;
	for(size_t k = 0; k < nc; k++) th.emplace_back([&, k]() {
		Piece &q = pc[k];
		const uint64_t l = cuts[k + 1] - cuts[k];
		uint64_t nl = 0;
		for(const char *p = text + cuts[k], *e = p + l; (p = (const char *)memchr(p, '\n', (size_t)(e - p))) != nullptr; p++) nl++;
		const uint64_t cap = nl / 4 + 2;                           // >= records: each has at least four lines
		q.seq.reset(new uint8_t[l]); q.qual.reset(new uint8_t[l]); q.off.reset(new uint64_t[cap + 1]); q.starts.reset(new uint64_t[cap + 1]);
		if(names && nameStride) q.names.reset(new char[cap * (uint64_t)nameStride]());      // zeroed rows (copied whole)
		q.rc = fastqParseCore(text + cuts[k], l, cap, l, q.seq.get(), q.qual.get(), q.off.get(), q.names.get(),
		                      nameStride, &q.n, &q.used, q.starts.get());
	});
	for(auto &t : th) t.join();
	// ---- concatenate while the limits allow; a piece that did not end on its boundary (a truncated last record) ends the parse
	uint64_t n = 0, nb = 0, cur = 0;
	off[0] = 0;
	for(size_t k = 0; k < nc; k++) {
		Piece &q = pc[k];
		if(q.rc) { if(n == 0 || true) { *nReads = 0; *consumed = 0; return q.rc; } }
		uint64_t take = q.n;
		if(n + take > maxReads) take = maxReads - n;
		while(take > 0 && nb + q.off[take] > maxBases) take--;
		const uint64_t bases = q.off[take];
		memcpy(seq + nb, q.seq.get(), bases);
		memcpy(qual + nb, q.qual.get(), bases);
		for(uint64_t i = 1; i <= take; i++) off[n + i] = nb + q.off[i];
		if(names && nameStride && take) memcpy(names + n * (uint64_t)nameStride, q.names.get(), take * (uint64_t)nameStride);
		n += take; nb += bases;
		cur = cuts[k] + (take == q.n ? q.used : q.starts[take]);
		if(take < q.n || cuts[k] + q.used < cuts[k + 1]) break;   // limit reached, or a truncated record at the end of the piece
	}
	*nReads = n; *consumed = cur;
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
	if(!c || (nReads && !res)) return -1;
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
		if(pr.pair_type == 2 && !multi(a) && !multi(b)) { c->ndiscord++; continue; }
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
