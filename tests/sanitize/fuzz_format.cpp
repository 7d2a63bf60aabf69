// Memory-safety fuzz of bt2g_sam_format (tests/test_host_sanitizers.py builds it with -fsanitize=address,undefined together with
// csrc/sam_host.cpp): exact-size copies of real result arrays with randomly damaged fields, narrow op rows, long names, every flag,
// 1..9 threads; the size query, the formatted bytes and the one-thread bytes must agree.  argv: seed, dump file (header of six
// uint64: reads, max_ops, name stride, bases, sizeof read_result, sizeof pair_result; then seq, qual, off, res, ops, pairs, names).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <random>
#include "bt2g.h"
static std::vector<char> slurp(const char *p) { FILE *f = fopen(p, "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); std::vector<char> b(n); if(fread(b.data(), 1, n, f) != (size_t)n) abort(); fclose(f); return b; }
int main(int argc, char **argv) {
	std::mt19937_64 rng(argc > 1 ? atoi(argv[1]) : 1);
	auto rnd = [&](uint64_t n) { return n ? rng() % n : 0; };
	auto d = slurp(argc > 2 ? argv[2] : "data.bin");
	const uint64_t *h = (const uint64_t *)d.data();
	uint64_t nrAll = h[0], maxOps = h[1], ns = h[2], nb = h[3], rsz = h[4], psz = h[5];
	const char *p = d.data() + 48;
	const uint8_t *seq = (const uint8_t *)p; p += nb; const uint8_t *qual = (const uint8_t *)p; p += nb;
	const uint64_t *off = (const uint64_t *)p; p += 8 * (nrAll + 1);
	const bt2g_read_result *res = (const bt2g_read_result *)p; p += rsz * nrAll;
	const uint8_t *ops = (const uint8_t *)p; p += nrAll * maxOps;
	const bt2g_pair_result *pairs = (const bt2g_pair_result *)p; p += psz * (nrAll / 2);
	const char *names = p;
	const uint64_t nr = nrAll;
	const char *rn[1] = {"gi|9626243|ref|NC_001416.1|"};
	long bad = 0;
	for(int it = 0; it < 200; it++) {
		// exact-size copies (so that any read past an array is caught), randomly damaged result fields
		const uint64_t n = 2 * (1 + rnd(nr / 2 - 1));
		std::vector<uint8_t> s(seq, seq + off[n]), q(qual, qual + off[n]);
		std::vector<uint64_t> o(off, off + n + 1);
		std::vector<bt2g_read_result> r(res, res + n);
		const uint32_t mo = rnd(3) ? (uint32_t)maxOps : 20 + (uint32_t)rnd(maxOps);          // narrower rows than the alignments need
		std::vector<uint8_t> op((size_t)n * mo);
		for(uint64_t i = 0; i < n; i++) memcpy(op.data() + i * mo, ops + i * maxOps, mo < maxOps ? mo : maxOps);
		std::vector<bt2g_pair_result> pr(pairs, pairs + n / 2);
		std::vector<std::string> nm(n);
		std::vector<const char *> np_(n);
		for(uint64_t i = 0; i < n; i++) { nm[i] = std::string(names + i * ns); if(rnd(20) == 0) nm[i] += std::string(rnd(400), 'y'); np_[i] = nm[i].c_str(); }
		for(int k = 0; k < 30; k++) {
			bt2g_read_result &x = r[rnd(n)];
			switch(rnd(7)) {
			case 0: x.found ^= 0x100; break;
			case 1: x.found ^= 0x200; break;
			case 2: x.nops = (int32_t)rnd(2 * maxOps); break;
			case 3: x.found = 0; break;
			case 4: x.tidx = rnd(3); break;
			case 5: x.trim_left = (int32_t)rnd(10); break;
			case 6: x.score2 = (int32_t)rnd(100) - 50; break;
			}
		}
		for(int k = 0; k < 10; k++) pr[rnd(n / 2)].pair_type = (int32_t)rnd(4);
		bt2g_sam_opts opt; memset(&opt, 0, sizeof(opt)); opt.ref_names = rn; opt.n_refs = 1; opt.read_names = rnd(5) ? np_.data() : nullptr; opt.threads = 1 + (int)rnd(9);
		opt.flags = (uint32_t)rnd(8); opt.sc_filter_maxlen = (int32_t)rnd(40); if(rnd(3) == 0) opt.rg_optflag = "RG:Z:grp1";
		bt2g_reads rd; rd.n_reads = n; rd.seq = s.data(); rd.qual = rnd(6) ? q.data() : nullptr; rd.off = o.data();
		const bool paired = rnd(4) != 0;
		uint64_t w = 0;
		int rc = bt2g_sam_format(&opt, &rd, r.data(), op.data(), mo, paired ? pr.data() : nullptr, nullptr, 0, &w);
		if(rc != -3 && !(rc >= 0 && w == 0)) { printf("size query rc=%d\n", rc); bad++; continue; }
		std::vector<char> out(w);
		uint64_t w2 = 0;
		rc = bt2g_sam_format(&opt, &rd, r.data(), op.data(), mo, paired ? pr.data() : nullptr, out.data(), out.size(), &w2);
		if(rc < 0 || w2 != w) { printf("format rc=%d w=%lu w2=%lu\n", rc, (unsigned long)w, (unsigned long)w2); bad++; continue; }
		// one thread gives the same bytes
		opt.threads = 1;
		std::vector<char> out1(w);
		rc = bt2g_sam_format(&opt, &rd, r.data(), op.data(), mo, paired ? pr.data() : nullptr, out1.data(), out1.size(), &w2);
		if(rc < 0 || w2 != w || memcmp(out.data(), out1.data(), w)) { printf("threads differ it=%d\n", it); bad++; }
	}
	printf("200 iterations, %ld inconsistencies\n", bad);
	return bad != 0;
}
