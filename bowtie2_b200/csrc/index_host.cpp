// index_host.cpp -- bit-exact reader of bowtie2 index files (.bt2 / .bt2l).
//
// Replaces, for the alignment path, Ebwt::readIntoMemory (bt2_io.cpp:131-616) and the
// BitPairReference constructor (reference.cpp:30-260).  Field order follows the on-disk format
// (SURVEY.md Appendix A): .1 = i32 endian sentinel, OFF len, i32 lineRate, i32 linesPerSide,
// i32 offRate, i32 ftabChars, i32 flags, OFF nPat, OFF plen[nPat], OFF nFrag,
// OFF rstarts[3*nFrag], u8 ebwt[numSides*sideSz], OFF zOff, OFF fchr[5], OFF ftab[4^fc+1],
// OFF eftab[2*fc]; .2 = i32 sentinel, OFF offs[offsLen]; .3 = i32 sentinel, OFF nrecs,
// {OFF off, OFF len, u8 first}[nrecs]; .4 = raw 2-bit bases.
// Files whose sentinel reads 1<<24 are byte-swapped field by field exactly as the reference does
// (bt2_io.cpp:134-147 and every readU/readI call after it: header scalars, plen, rstarts, zOff, fchr, ftab,
// eftab, offs; NOT the ebwt[] sides, which the reference copies raw, bt2_io.cpp:320-386; reference.cpp:105-150
// for the .3 records).  An --offrate override larger than the stored offRate keeps every 2^diff-th SA sample
// (bt2_io.cpp:217-230, 545-573); a smaller one is ignored, as there.  The reference names that follow eftab in
// the .1 file ('\n'-separated, '\0'-terminated, bt2_io.cpp:496-511 / readEbwtRefnames :623-700) are kept.
#include "bt2g_internal.h"
#include <cstring>
#include <sys/stat.h>

namespace {

struct FileReader {
	FILE *f = nullptr;
	std::string path;
	bool swap = false;                                    // file written with the other byte order
	~FileReader() { if(f) fclose(f); }
	bool open(const std::string &p) { path = p; f = fopen(p.c_str(), "rb"); return f != nullptr; }
	bool read(void *dst, size_t n) { return fread(dst, 1, n, f) == n; }
	bool readI32(int32_t &v) {
		uint32_t x;
		if(!read(&x, 4)) return false;
		v = (int32_t)(swap ? __builtin_bswap32(x) : x);
		return true;
	}
	// first word of every index file: 1, or 1 with the bytes reversed (then everything after it is swapped)
	bool readSentinel() {
		uint32_t x;
		if(!read(&x, 4)) return false;
		if(x == 1) { swap = false; return true; }
		if(x == 0x01000000u) { swap = true; return true; }
		return false;
	}
	bool readOff(int offSize, uint64_t &v) {
		if(offSize == 4) { uint32_t x; if(!read(&x, 4)) return false; v = swap ? __builtin_bswap32(x) : x; return true; }
		if(!read(&v, 8)) return false;
		if(swap) v = __builtin_bswap64(v);
		return true;
	}
	bool readVec(std::vector<uint8_t> &v, uint64_t bytes) {
		v.resize(bytes);
		return bytes == 0 || read(v.data(), bytes);
	}
	// array of OFF-sized integers, brought to host byte order
	bool readOffVec(std::vector<uint8_t> &v, uint64_t count, int offSize) {
		if(!readVec(v, count * offSize)) return false;
		if(swap) {
			if(offSize == 4) { uint32_t *p = (uint32_t *)v.data(); for(uint64_t i = 0; i < count; i++) p[i] = __builtin_bswap32(p[i]); }
			else { uint64_t *p = (uint64_t *)v.data(); for(uint64_t i = 0; i < count; i++) p[i] = __builtin_bswap64(p[i]); }
		}
		return true;
	}
	bool skip(uint64_t bytes) { return fseeko(f, (off_t)bytes, SEEK_CUR) == 0; }
};

bool fileExists(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0; }

struct EbwtHeader {
	uint64_t len = 0, nPat = 0, nFrag = 0, zOff = 0, fchr[5] = {0, 0, 0, 0, 0};
	int32_t lineRate = 0, offRate = 0, ftabChars = 0, flags = 0;
};

// Reads one <base>.1.<ext>; rstarts kept only when wantRstarts (the mirror index's copy is
// never used, bt2_search.cpp:4845-4853).
int readEbwt1(const std::string &path, int offSize, bool wantRstarts, EbwtHeader &h,
              std::vector<uint8_t> &plen, std::vector<uint8_t> &rstarts, std::vector<uint8_t> &ebwt,
              std::vector<uint8_t> &ftab, std::vector<uint8_t> &eftab, std::vector<std::string> *names, std::string &err) {
	FileReader r;
	if(!r.open(path)) { err = "cannot open " + path; return -1; }
	int32_t linesPerSide = 0;
	if(!r.readSentinel()) { err = "bad endian sentinel in " + path; return -1; }
	bool ok = r.readOff(offSize, h.len) && r.readI32(h.lineRate) && r.readI32(linesPerSide) &&
	          r.readI32(h.offRate) && r.readI32(h.ftabChars) && r.readI32(h.flags) &&
	          r.readOff(offSize, h.nPat);
	if(!ok) { err = "short header: " + path; return -1; }
	if(h.lineRate != (offSize == 4 ? 6 : 7)) { err = "unexpected lineRate in " + path; return -1; }
	if(!r.readOffVec(plen, h.nPat, offSize) || !r.readOff(offSize, h.nFrag)) { err = "short plen: " + path; return -1; }
	if(wantRstarts) {
		if(!r.readOffVec(rstarts, h.nFrag * 3, offSize)) { err = "short rstarts: " + path; return -1; }
	} else {
		if(!r.skip(h.nFrag * 3 * offSize)) { err = "seek failed: " + path; return -1; }
	}
	// EbwtParams::init arithmetic (bt2_idx.h:133-167)
	uint64_t sideSz = 1ull << h.lineRate, sideBwtSz = sideSz - 4ull * offSize;
	uint64_t bwtSz = h.len / 4 + 1;
	uint64_t numSides = (bwtSz + sideBwtSz - 1) / sideBwtSz;
	if(!r.readVec(ebwt, numSides * sideSz)) { err = "short ebwt: " + path; return -1; }
	if(!r.readOff(offSize, h.zOff)) { err = "short zOff: " + path; return -1; }
	for(int i = 0; i < 5; i++) if(!r.readOff(offSize, h.fchr[i])) { err = "short fchr: " + path; return -1; }
	uint64_t ftabLen = (1ull << (2 * h.ftabChars)) + 1;
	if(!r.readOffVec(ftab, ftabLen, offSize) || !r.readOffVec(eftab, 2ull * h.ftabChars, offSize)) {
		err = "short ftab: " + path; return -1;
	}
	if(names) {
		// one name per line up to the terminating NUL (or the end of the file)
		names->clear();
		std::string cur;
		int c;
		while((c = fgetc(r.f)) != EOF && c != 0) {
			if(c == '\n') { names->push_back(cur); cur.clear(); }
			else cur.push_back((char)c);
		}
		if(!cur.empty()) names->push_back(cur);
	}
	return 0;
}

} // namespace

int bt2g_read_index_files(const char *basename, HostIndex &out, std::string &err) {
	return bt2g_read_index_files_ex(basename, -1, out, err);
}

int bt2g_read_index_files_ex(const char *basename, int offRateOverride, HostIndex &out, std::string &err) {
	std::string base(basename), ext = "bt2";
	int offSize = 4;
	if(!fileExists(base + ".1.bt2")) {
		if(!fileExists(base + ".1.bt2l")) { err = "no index at " + base + ".1.bt2[l]"; return -1; }
		ext = "bt2l"; offSize = 8;
	}
	EbwtHeader hf, hb;
	std::vector<uint8_t> dummy;
	if(readEbwt1(base + ".1." + ext, offSize, true, hf, out.plen, out.rstarts, out.ebwt_fw,
	             out.ftab_fw, out.eftab_fw, &out.names, err)) return -1;
	bool haveBw = fileExists(base + ".rev.1." + ext);
	if(haveBw) {
		std::vector<uint8_t> plenBw;
		if(readEbwt1(base + ".rev.1." + ext, offSize, false, hb, plenBw, dummy, out.ebwt_bw,
		             out.ftab_bw, out.eftab_bw, nullptr, err)) return -1;
		if(hb.len != hf.len || hb.ftabChars != hf.ftabChars) { err = "mirror index does not match forward index"; return -1; }
	}
	// .2: SA sample
	{
		FileReader r;
		std::string p = base + ".2." + ext;
		if(!r.open(p)) { err = "cannot open " + p; return -1; }
		if(!r.readSentinel()) { err = "bad sentinel in " + p; return -1; }
		uint64_t offsLen = (hf.len + 1 + (1ull << hf.offRate) - 1) >> hf.offRate;
		if(!r.readOffVec(out.offs, offsLen, offSize)) { err = "short offs: " + p; return -1; }
		if(offRateOverride > hf.offRate) {
			// keep every 2^diff-th sample: offs'[k] = offs[k << diff] (bt2_io.cpp:545-573)
			const int diff = offRateOverride - hf.offRate;
			if(diff >= 32) { err = "offrate override too large"; return -1; }
			uint64_t sampled = offsLen >> diff;
			if(offsLen & ((1ull << diff) - 1)) sampled++;
			if(offSize == 4) { uint32_t *o = (uint32_t *)out.offs.data(); for(uint64_t k = 0; k < sampled; k++) o[k] = o[k << diff]; }
			else { uint64_t *o = (uint64_t *)out.offs.data(); for(uint64_t k = 0; k < sampled; k++) o[k] = o[k << diff]; }
			out.offs.resize(sampled * offSize);
			out.offs.shrink_to_fit();
			hf.offRate = offRateOverride;
		}
	}
	// .3/.4: packed reference
	uint64_t nRecs = 0;
	if(fileExists(base + ".3." + ext)) {
		FileReader r;
		std::string p = base + ".3." + ext;
		if(!r.open(p)) { err = "cannot open " + p; return -1; }
		if(!r.readSentinel()) { err = "bad sentinel in " + p; return -1; }
		if(!r.readOff(offSize, nRecs)) { err = "short .3"; return -1; }
		out.rec_off.resize(nRecs * offSize); out.rec_len.resize(nRecs * offSize); out.rec_first.resize(nRecs);
		uint64_t cumsz = 0;
		for(uint64_t i = 0; i < nRecs; i++) {
			uint64_t o, l; uint8_t first;
			if(!r.readOff(offSize, o) || !r.readOff(offSize, l) || !r.read(&first, 1)) { err = "short record in " + p; return -1; }
			if(offSize == 4) { ((uint32_t *)out.rec_off.data())[i] = (uint32_t)o; ((uint32_t *)out.rec_len.data())[i] = (uint32_t)l; }
			else { ((uint64_t *)out.rec_off.data())[i] = o; ((uint64_t *)out.rec_len.data())[i] = l; }
			out.rec_first[i] = first ? 1 : 0;
			cumsz += l;
		}
		FileReader r4;
		std::string p4 = base + ".4." + ext;
		if(!r4.open(p4)) { err = "cannot open " + p4; return -1; }
		if(!r4.readVec(out.ref_buf, (cumsz + 3) >> 2)) { err = "short " + p4; return -1; }
	}
	bt2g_index_host &d = out.d;
	d.off_size = offSize; d.line_rate = hf.lineRate; d.off_rate = hf.offRate; d.ftab_chars = hf.ftabChars;
	d.len = hf.len; d.n_pat = hf.nPat; d.n_frag = hf.nFrag; d.z_off_fw = hf.zOff; d.z_off_bw = haveBw ? hb.zOff : 0;
	for(int i = 0; i < 5; i++) d.fchr[i] = hf.fchr[i];
	d.plen = out.plen.data(); d.rstarts = out.rstarts.data();
	d.ebwt_fw = out.ebwt_fw.data(); d.ebwt_bw = haveBw ? out.ebwt_bw.data() : nullptr;
	d.ftab_fw = out.ftab_fw.data(); d.eftab_fw = out.eftab_fw.data();
	d.ftab_bw = haveBw ? out.ftab_bw.data() : nullptr; d.eftab_bw = haveBw ? out.eftab_bw.data() : nullptr;
	d.offs = out.offs.data();
	d.n_recs = nRecs;
	d.rec_off = nRecs ? out.rec_off.data() : nullptr; d.rec_len = nRecs ? out.rec_len.data() : nullptr;
	d.rec_first = nRecs ? out.rec_first.data() : nullptr; d.ref_buf = nRecs ? out.ref_buf.data() : nullptr;
	return 0;
}

// ---- host-side view of an index on disk (no GPU involved) ---------------------------------------
// Used to read an index once on the loading rank (the arrays are then broadcast to the other GPUs,
// bowtie2_b200/dist.py), to print the SAM header, and by the CPU test-suite.
struct bt2g_index_file {
	HostIndex h;
	std::vector<const char *> namePtrs;
	std::vector<uint64_t> lens;
};

extern "C" int bt2g_index_file_open(const char *basename, int offrateOverride, bt2g_index_file **out, char *err, uint32_t errCap) {
	if(!basename || !out) return -1;
	bt2g_index_file *f = new bt2g_index_file();
	std::string e;
	if(bt2g_read_index_files_ex(basename, offrateOverride, f->h, e)) {
		if(err && errCap) { snprintf(err, errCap, "%s", e.c_str()); }
		delete f;
		return -1;
	}
	for(const std::string &s : f->h.names) f->namePtrs.push_back(s.c_str());
	const bt2g_index_host &d = f->h.d;
	for(uint64_t i = 0; i < d.n_pat; i++)
		f->lens.push_back(d.off_size == 4 ? (uint64_t)((const uint32_t *)d.plen)[i] : ((const uint64_t *)d.plen)[i]);
	*out = f;
	return 0;
}

extern "C" const bt2g_index_host *bt2g_index_file_desc(const bt2g_index_file *f) { return f ? &f->h.d : nullptr; }
extern "C" uint64_t bt2g_index_file_n_refs(const bt2g_index_file *f) { return f ? f->namePtrs.size() : 0; }
extern "C" const char *const *bt2g_index_file_ref_names(const bt2g_index_file *f) { return f ? f->namePtrs.data() : nullptr; }
extern "C" const uint64_t *bt2g_index_file_ref_lens(const bt2g_index_file *f) { return f ? f->lens.data() : nullptr; }
extern "C" void bt2g_index_file_close(bt2g_index_file *f) { delete f; }

extern "C" int bt2g_load_index_files_ex(bt2g_ctx *ctx, const char *basename, int offrateOverride) {
	if(!ctx || !basename) return -1;
	HostIndex h;
	if(bt2g_read_index_files_ex(basename, offrateOverride, h, ctx->err)) return -1;
	return bt2g_load_index_host(ctx, &h.d);
}
