// index_host.cpp -- bit-exact reader of bowtie2 index files (.bt2 / .bt2l).
//
// Replaces, for the alignment path, Ebwt::readIntoMemory (bt2_io.cpp:131-616) and the
// BitPairReference constructor (reference.cpp:30-260).  Field order follows the on-disk format
// (SURVEY.md Appendix A): .1 = i32 endian sentinel, OFF len, i32 lineRate, i32 linesPerSide,
// i32 offRate, i32 ftabChars, i32 flags, OFF nPat, OFF plen[nPat], OFF nFrag,
// OFF rstarts[3*nFrag], u8 ebwt[numSides*sideSz], OFF zOff, OFF fchr[5], OFF ftab[4^fc+1],
// OFF eftab[2*fc]; .2 = i32 sentinel, OFF offs[offsLen]; .3 = i32 sentinel, OFF nrecs,
// {OFF off, OFF len, u8 first}[nrecs]; .4 = raw 2-bit bases.
// Only little-endian files are accepted (the reference byte-swaps big-endian ones,
// bt2_io.cpp:134-147; nothing produces those on the platforms this library runs on).
#include "bt2g_internal.h"
#include <cstring>
#include <sys/stat.h>

namespace {

struct FileReader {
	FILE *f = nullptr;
	std::string path;
	~FileReader() { if(f) fclose(f); }
	bool open(const std::string &p) { path = p; f = fopen(p.c_str(), "rb"); return f != nullptr; }
	bool read(void *dst, size_t n) { return fread(dst, 1, n, f) == n; }
	bool readI32(int32_t &v) { return read(&v, 4); }
	bool readOff(int offSize, uint64_t &v) {
		if(offSize == 4) { uint32_t x; if(!read(&x, 4)) return false; v = x; return true; }
		return read(&v, 8);
	}
	bool readVec(std::vector<uint8_t> &v, uint64_t bytes) {
		v.resize(bytes);
		return bytes == 0 || read(v.data(), bytes);
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
              std::vector<uint8_t> &ftab, std::vector<uint8_t> &eftab, std::string &err) {
	FileReader r;
	if(!r.open(path)) { err = "cannot open " + path; return -1; }
	int32_t one = 0, linesPerSide = 0;
	if(!r.readI32(one)) { err = "short read: " + path; return -1; }
	if(one != 1) { err = "unsupported endianness in " + path; return -1; }
	bool ok = r.readOff(offSize, h.len) && r.readI32(h.lineRate) && r.readI32(linesPerSide) &&
	          r.readI32(h.offRate) && r.readI32(h.ftabChars) && r.readI32(h.flags) &&
	          r.readOff(offSize, h.nPat);
	if(!ok) { err = "short header: " + path; return -1; }
	if(h.lineRate != (offSize == 4 ? 6 : 7)) { err = "unexpected lineRate in " + path; return -1; }
	if(!r.readVec(plen, h.nPat * offSize) || !r.readOff(offSize, h.nFrag)) { err = "short plen: " + path; return -1; }
	if(wantRstarts) {
		if(!r.readVec(rstarts, h.nFrag * 3 * offSize)) { err = "short rstarts: " + path; return -1; }
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
	if(!r.readVec(ftab, ftabLen * offSize) || !r.readVec(eftab, 2ull * h.ftabChars * offSize)) {
		err = "short ftab: " + path; return -1;
	}
	return 0;
}

} // namespace

int bt2g_read_index_files(const char *basename, HostIndex &out, std::string &err) {
	std::string base(basename), ext = "bt2";
	int offSize = 4;
	if(!fileExists(base + ".1.bt2")) {
		if(!fileExists(base + ".1.bt2l")) { err = "no index at " + base + ".1.bt2[l]"; return -1; }
		ext = "bt2l"; offSize = 8;
	}
	EbwtHeader hf, hb;
	std::vector<uint8_t> dummy;
	if(readEbwt1(base + ".1." + ext, offSize, true, hf, out.plen, out.rstarts, out.ebwt_fw,
	             out.ftab_fw, out.eftab_fw, err)) return -1;
	bool haveBw = fileExists(base + ".rev.1." + ext);
	if(haveBw) {
		std::vector<uint8_t> plenBw;
		if(readEbwt1(base + ".rev.1." + ext, offSize, false, hb, plenBw, dummy, out.ebwt_bw,
		             out.ftab_bw, out.eftab_bw, err)) return -1;
		if(hb.len != hf.len || hb.ftabChars != hf.ftabChars) { err = "mirror index does not match forward index"; return -1; }
	}
	// .2: SA sample
	{
		FileReader r;
		std::string p = base + ".2." + ext;
		if(!r.open(p)) { err = "cannot open " + p; return -1; }
		int32_t one = 0;
		if(!r.readI32(one) || one != 1) { err = "bad sentinel in " + p; return -1; }
		uint64_t offsLen = (hf.len + 1 + (1ull << hf.offRate) - 1) >> hf.offRate;
		if(!r.readVec(out.offs, offsLen * offSize)) { err = "short offs: " + p; return -1; }
	}
	// .3/.4: packed reference
	uint64_t nRecs = 0;
	if(fileExists(base + ".3." + ext)) {
		FileReader r;
		std::string p = base + ".3." + ext;
		if(!r.open(p)) { err = "cannot open " + p; return -1; }
		int32_t one = 0;
		if(!r.readI32(one) || one != 1) { err = "bad sentinel in " + p; return -1; }
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
