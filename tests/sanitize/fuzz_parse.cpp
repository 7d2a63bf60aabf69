// Memory-safety and consistency fuzz of the FASTQ parsers (tests/test_host_sanitizers.py builds it with -fsanitize=address,undefined
// together with csrc/sam_host.cpp): random mate texts (blank lines, CRLF, empty reads, long names, truncated tails, damaged bytes),
// random limits, 1..9 threads; bt2g_fastq_parse_pairs_mt against bt2g_fastq_parse per file, bt2g_fastq_parse_mt against bt2g_fastq_parse.
// argv: seed, iterations.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <random>
#include "bt2g.h"
int main(int argc, char **argv) {
	std::mt19937_64 rng(argc > 1 ? atoi(argv[1]) : 1);
	int iters = argc > 2 ? atoi(argv[2]) : 300;
	auto rnd = [&](uint64_t n) { return n ? rng() % n : 0; };
	long bad = 0;
	for(int it = 0; it < iters; it++) {
		// two mate texts, sometimes big enough for the cutter (> 1 MiB)
		const bool big = rnd(4) == 0;
		const int nrec = big ? 9000 + (int)rnd(4000) : (int)rnd(60);
		std::string t[2];
		for(int f = 0; f < 2; f++) {
			const int n = nrec - (rnd(5) == 0 ? (int)rnd(3) : 0);
			for(int i = 0; i < n; i++) {
				int L = (int)rnd(big ? 160 : 40);
				if(rnd(30) == 0) L = 0;
				t[f] += "@r" + std::to_string(i) + (rnd(3) ? " comment/" + std::to_string(f + 1) : "");
				if(rnd(50) == 0) t[f] += std::string(300, 'x');          // a name longer than the stride
				t[f] += rnd(40) == 0 ? "\r\n" : "\n";
				for(int k = 0; k < L; k++) t[f] += "ACGTNacgtn.R"[rnd(12)];
				t[f] += "\n+\n";
				for(int k = 0; k < L; k++) { char c = (char)(33 + rnd(60)); if(k == 0 && rnd(8) == 0) c = '@'; t[f] += c; }
				t[f] += "\n";
				if(rnd(200) == 0) t[f] += "\n";
			}
			// corruptions
			if(rnd(6) == 0 && !t[f].empty()) t[f].resize(t[f].size() - rnd(std::min<size_t>(t[f].size(), 50)));     // truncated tail
			if(rnd(25) == 0 && t[f].size() > 10) t[f][rnd(t[f].size())] = "@+\n X"[rnd(5)];                          // a damaged byte
		}
		const int threads = 1 + (int)rnd(9);
		const uint32_t stride = rnd(3) ? 32 : 8 + (uint32_t)rnd(100);
		const uint64_t capPairs = rnd(4) == 0 ? rnd(nrec + 2) : (uint64_t)nrec + 5;
		const uint64_t capBases = rnd(6) == 0 ? rnd(t[0].size() + t[1].size() + 1) : t[0].size() + t[1].size();
		std::vector<uint8_t> seq(capBases + 1), qual(capBases + 1);
		std::vector<uint64_t> off(2 * capPairs + 1);
		std::vector<char> names((size_t)2 * capPairs * stride + 1);
		uint64_t n = 0, u1 = 0, u2 = 0;
		int rc = bt2g_fastq_parse_pairs_mt(t[0].data(), t[0].size(), t[1].data(), t[1].size(), capPairs, capBases, seq.data(), qual.data(), off.data(),
		                                   rnd(5) ? names.data() : nullptr, stride, &n, &u1, &u2, threads);
		if(rc == 0) {
			if(n > capPairs || off[2 * n] > capBases || u1 > t[0].size() || u2 > t[1].size()) { printf("LIMITS it=%d\n", it); bad++; }
			// the same through the serial single-file parser: record i of each file
			std::vector<uint8_t> s1(t[0].size() + 1), q1(t[0].size() + 1); std::vector<uint64_t> o1(n + 2); uint64_t n1 = 0, c1 = 0;
			int rc1 = bt2g_fastq_parse(t[0].data(), u1, n + 1, t[0].size(), s1.data(), q1.data(), o1.data(), nullptr, 0, &n1, &c1);
			if(rc1 == 0 && n1 != n) { printf("COUNT it=%d n=%lu serial=%lu\n", it, (unsigned long)n, (unsigned long)n1); bad++; }
			if(rc1 == 0 && n1 == n) for(uint64_t i = 0; i < n; i++) {
				const uint64_t l = o1[i + 1] - o1[i];
				if(off[2 * i + 1] - off[2 * i] != l || memcmp(seq.data() + off[2 * i], s1.data() + o1[i], l) || memcmp(qual.data() + off[2 * i], q1.data() + o1[i], l)) { printf("DATA it=%d i=%lu\n", it, (unsigned long)i); bad++; break; }
			}
		}
		// the single-file mt parser against the serial one
		{
			const uint64_t maxReads = rnd(3) == 0 ? rnd(nrec + 2) : (uint64_t)nrec + 5, maxBases = rnd(5) == 0 ? rnd(t[0].size() + 1) : t[0].size();
			std::vector<uint8_t> sa(maxBases + 1), qa(maxBases + 1), sb(maxBases + 1), qb(maxBases + 1);
			std::vector<uint64_t> oa(maxReads + 2), ob(maxReads + 2);
			std::vector<char> na((size_t)maxReads * stride + 1), nb_((size_t)maxReads * stride + 1);
			uint64_t ra = 0, rb = 0, ca = 0, cb = 0;
			int r1 = bt2g_fastq_parse(t[0].data(), t[0].size(), maxReads, maxBases, sa.data(), qa.data(), oa.data(), na.data(), stride, &ra, &ca);
			int r2 = bt2g_fastq_parse_mt(t[0].data(), t[0].size(), maxReads, maxBases, sb.data(), qb.data(), ob.data(), nb_.data(), stride, &rb, &cb, threads);
			if(r1 == 0 && r2 == 0 && (ra != rb || ca != cb || memcmp(sa.data(), sb.data(), oa[ra]) || memcmp(oa.data(), ob.data(), 8 * (ra + 1)) || memcmp(na.data(), nb_.data(), (size_t)ra * stride))) { printf("MT it=%d ra=%lu rb=%lu ca=%lu cb=%lu\n", it, (unsigned long)ra, (unsigned long)rb, (unsigned long)ca, (unsigned long)cb); bad++; }
			if((r1 == 0) != (r2 == 0) && !big) { /* an error in a later piece surfaces in the mt parser even when the serial one stopped earlier at a limit */ }
		}
	}
	printf("%d iterations, %ld inconsistencies\n", iters, bad);
	return bad != 0;
}
