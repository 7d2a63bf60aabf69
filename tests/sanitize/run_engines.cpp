// The host builds of the search policy -- csrc/xengine.cuh (the state machine k_xe_step runs on the device, here through
// csrc/xengine_host.cpp) and csrc/policy_engine.cpp (the coroutine engine) -- under AddressSanitizer + UndefinedBehaviorSanitizer over the
// C oracle's entry-point table: a fixed-capacity array of the unit state written out of bounds would be silent memory corruption on the
// device.  tests/test_host_sanitizers.py builds this with the three sources and feeds it dumps of tests/parity_fuzz.py cases.
// argv: index base, dump file.  Dump: six uint64 (reads, bases, name stride, sizeof params, local, off_size), scoring override (8 int32 +
// 2 double, match_bonus < 0 = none), the bt2g_policy_params bytes, seq, qual, off, names, then the expected result rows.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "bt2g.h"
extern "C" {
struct bt2o_index;
bt2o_index *bt2o_open(const char *base, int load_mirror, int load_ref);
void *bt2o_policy_table(bt2o_index *ix, int local, int off_size, bt2g_policy_backend *be);
void bt2o_policy_table_scoring(void *tv, int match_bonus, int mmp_max, int mmp_min, int n_pen, int rdgap_const, int rdgap_linear, int rfgap_const, int rfgap_linear);
void bt2o_policy_table_nceil(void *tv, double nceil_const, double nceil_linear);
}
static std::vector<char> slurp(const char *p) { FILE *f = fopen(p, "rb"); if(!f) { perror(p); exit(2); } fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); std::vector<char> b(n); if(fread(b.data(), 1, n, f) != (size_t)n) abort(); fclose(f); return b; }
int main(int argc, char **argv) {
	if(argc < 3) return 2;
	auto d = slurp(argv[2]);
	const uint64_t *h = (const uint64_t *)d.data();
	const uint64_t n = h[0], nb = h[1], ns = h[2], psz = h[3], local = h[4], offSize = h[5];
	if(psz != sizeof(bt2g_policy_params)) { printf("params size %lu != %zu\n", (unsigned long)psz, sizeof(bt2g_policy_params)); return 2; }
	const char *p = d.data() + 48;
	int32_t sci[8]; memcpy(sci, p, 32); p += 32;
	double scd[2]; memcpy(scd, p, 16); p += 16;
	bt2g_policy_params prm; memcpy(&prm, p, sizeof(prm)); p += sizeof(prm);
	const uint8_t *seq = (const uint8_t *)p; p += nb; const uint8_t *qual = (const uint8_t *)p; p += nb;
	std::vector<uint64_t> off(n + 1); memcpy(off.data(), p, 8 * (n + 1)); p += 8 * (n + 1);
	const char *names = p; p += n * ns;
	const bt2g_read_result *want = (const bt2g_read_result *)p;
	std::vector<const char *> np_(n); for(uint64_t i = 0; i < n; i++) np_[i] = names + i * ns;
	bt2o_index *ix = bt2o_open(argv[1], 1, 1);
	if(!ix) { printf("cannot open %s\n", argv[1]); return 2; }
	bt2g_policy_backend be; memset(&be, 0, sizeof(be));
	void *t = bt2o_policy_table(ix, (int)local, (int)offSize, &be);
	if(sci[0] >= 0) bt2o_policy_table_scoring(t, sci[0], sci[1], sci[2], sci[3], sci[4], sci[5], sci[6], sci[7]);
	if(scd[1] >= 0) bt2o_policy_table_nceil(t, scd[0], scd[1]);
	bt2g_reads rd; rd.n_reads = n; rd.seq = seq; rd.qual = qual; rd.off = off.data();
	uint32_t maxLen = 1; for(uint64_t i = 0; i < n; i++) if(off[i + 1] - off[i] > maxLen) maxLen = (uint32_t)(off[i + 1] - off[i]);
	const uint32_t maxOps = 4 * maxLen + 64;
	int bad = 0;
	for(int eng = 0; eng < 2; eng++) {
		std::vector<bt2g_read_result> res(n); std::vector<uint8_t> ops((size_t)n * maxOps); std::vector<bt2g_pair_result> pairs(n / 2 + 1);
		uint64_t stats[8] = {0};
		const int rc = eng == 0 ? bt2g_xengine_align_host(&be, &prm, &rd, np_.data(), res.data(), ops.data(), maxOps, prm.paired ? pairs.data() : nullptr, stats)
		                        : bt2g_policy_align(&be, &prm, &rd, np_.data(), res.data(), ops.data(), maxOps, prm.paired ? pairs.data() : nullptr, stats);
		if(rc) { printf("engine %d failed (%d)\n", eng, rc); bad++; continue; }
		for(uint64_t i = 0; i < n; i++)
			if(res[i].found != want[i].found || res[i].score != want[i].score || res[i].refoff != want[i].refoff || res[i].tidx != want[i].tidx || res[i].mapq != want[i].mapq || res[i].nops != want[i].nops) { printf("engine %d read %lu differs from the uninstrumented run\n", eng, (unsigned long)i); bad++; break; }
		printf("engine %d: %lu reads, stats %lu %lu %lu\n", eng, (unsigned long)n, (unsigned long)stats[0], (unsigned long)stats[1], (unsigned long)stats[2]);
	}
	return bad != 0;
}
