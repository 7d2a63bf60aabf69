// policy_engine.cpp -- the reference's sequential, RNG-driven search policy (multiseedSearchWorker, bt2_search.cpp:3101-4250;
// SwDriver::extendSeeds / extendSeedsPaired, aligner_sw_driver.cpp:921-2615; AlnSinkWrap / ReportingState, aln_sink.cpp) as
// C++20 coroutines, scheduled in WAVES: every read (pair) in flight is a coroutine that suspends on one hot-path request at a
// time; the scheduler groups the pending requests by primitive and answers each group with ONE call of the matching entry point
// of include/bt2g.h (through the bt2g_policy_backend table, so that the control flow can be pinned on the CPU against a backend
// that answers from the oracle).  It is the compiled twin of bowtie2_b200/policy_engine.py + policy_waves.py, which is the
// specification (pinned byte-for-byte against the reference program's SAM); results go into the pipeline's result arrays.
// Scope: the primary alignment per read / pair (the secondary records of -k / -a are produced by the Python engine only).
// Host code; built with g++ -std=c++20.
#include <algorithm>
#include <array>
#include <cmath>
#include <coroutine>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <unordered_set>
#include <utility>
#include <vector>
#include "../../include/bt2g.h"
#define __host__
#define __device__
#define __forceinline__ inline
#include "mapq_device.cuh"
#include "pe_device.cuh"

namespace {

constexpr int64_t MIN_I64 = std::numeric_limits<int64_t>::min();
constexpr int64_t BIG = (int64_t)1 << 62;
enum { EXHAUSTED = 1, FULFILLED, PERFECT, SOFT_LIMIT, HARD_LIMIT };

// ---------------------------------------------------------------------------------------------- small pieces
struct Func {                                     // SimpleFunc (simple_func.h:40-125)
	int type; double C, L;
	double f(double x) const {
		const double g = type == 1 ? 0.0 : type == 2 ? x : type == 3 ? std::sqrt(x) : std::log(x);
		return C + L * g;
	}
	long long fi(double x) const { return (long long)f(x); }
};

struct Rng {                                      // RandomSource (random_source.h:32-180)
	uint32_t last = 0; int lastOff = 30;
	void init(uint32_t s) { last = s; lastOff = 30; }
	uint32_t u32() {
		last = 1664525u * last + 1013904223u;
		uint32_t ret = last >> 16;
		last = 1664525u * last + 1013904223u;
		ret ^= last;
		lastOff = 0;
		return ret;
	}
	uint64_t u64() { const uint64_t hi = u32(); return (hi << 32) | u32(); }
	int boolean() { if(lastOff > 31) u32(); const int r = (last >> lastOff) & 1; lastOff++; return r; }
	double flt() { return (double)((float)u32() / (float)0xffffffff); }
};

template <typename T> void shufflePortion(std::vector<T> &v, size_t begin, size_t num, Rng &rnd) {
	if(num < 2) return;
	size_t left = num;
	for(size_t i = begin; i < begin + num - 1; i++) {
		const size_t r = (size_t)(rnd.u64() % left);
		if(r > 0) std::swap(v[i], v[i + r]);
		left--;
	}
}
template <typename T, typename K> void shuffleEqualStreaks(std::vector<T> &v, K key, Rng &rnd) {
	size_t streak = 0;
	for(size_t i = 1; i < v.size(); i++) {
		if(key(v[i]) == key(v[i - 1])) { if(streak == 0) streak = 1; streak++; }
		else { if(streak > 1) shufflePortion(v, i - streak, streak, rnd); streak = 0; }
	}
	if(streak > 1) shufflePortion(v, v.size() - streak, streak, rnd);
}

struct Random1toN {                               // random_util.h:32-160
	size_t n = 0, cur = 0, thresh = 0; bool swaplist = false, converted = false;
	std::vector<size_t> list, seen;
	void init(size_t n_, bool all) {
		n = n_; cur = 0; converted = false; swaplist = n < 128 || all; list.clear(); seen.clear();
		thresh = std::max<size_t>(16, (size_t)(0.10f * (float)n));
	}
	bool inited() const { return n > 0; }
	bool done() const { return inited() && cur >= n; }
	void setDone() { cur = n; }
	size_t next(Rng &rnd) {
		if(cur == 0 && !converted) {
			if(n == 1) { cur = 1; return 0; }
			if(swaplist) { list.resize(n); for(size_t i = 0; i < n; i++) list[i] = i; }
		}
		if(swaplist) {
			const size_t r = cur + (rnd.u32() % (uint32_t)(n - cur));
			if(r != cur) std::swap(list[cur], list[r]);
			return list[cur++];
		}
		const size_t seenSz = seen.size();
		size_t rn;
		for(;;) {
			rn = rnd.u32() % (uint32_t)n;
			bool again = false;
			for(size_t i = 0; i < seenSz; i++) if(seen[i] == rn) { again = true; break; }
			if(!again) break;
		}
		seen.push_back(rn);
		cur++;
		if(seen.size() >= thresh && cur < n) {
			std::vector<size_t> s = seen;
			std::sort(s.begin(), s.end());
			list.clear();
			size_t k = 0;
			for(size_t j = 0; j < n; j++) { if(k < s.size() && s[k] == j) k++; else list.push_back(j); }
			seen.clear(); cur = 0; n = list.size(); converted = true; swaplist = true;
		}
		return rn;
	}
};

struct IntervalSet {                              // seenDiags (EIvalMergeListBinned): membership only
	struct Iv { int64_t tidx; bool fw; int64_t a, b; };
	std::vector<Iv> iv;
	void add(int64_t tidx, bool fw, int64_t off, int64_t len) { iv.push_back({tidx, fw, off, off + len}); }
	bool present(int64_t tidx, bool fw, int64_t off) const {
		for(const Iv &x : iv) if(x.tidx == tidx && x.fw == fw && x.a <= off && off < x.b) return true;
		return false;
	}
};

struct Edit { int pos; int chr, qchr; int type; };    // type 1 read gap, 2 ref gap, 3 mismatch; chr / qchr = nucleotide codes, 5 = '-'

struct Aln {
	int64_t tidx = 0, refoff = 0; bool fw = true; int64_t score = 0; int rdlen = 0;
	std::vector<Edit> edits; int ns = 0, refns = 0, trim5 = 0, trim3 = 0;
	int ext() const { return rdlen - trim5 - trim3; }
	int trimLeft() const { return fw ? trim5 : trim3; }
	int refExtent() const { int e = ext(); for(const Edit &x : edits) e += (x.type == 1) - (x.type == 2); return e; }
	std::vector<Edit> leftToRight() const {            // AlnRes::invertEdits (edit.cpp:50-78)
		if(fw) return edits;
		std::vector<Edit> out;
		const int e = ext();
		for(auto it = edits.rbegin(); it != edits.rend(); ++it) out.push_back({e - it->pos - (it->type == 1 ? 0 : 1), it->chr, it->qchr, it->type});
		return out;
	}
};

struct RedundantAlns {                            // aligner_result.cpp:929-1030
	struct Key { int64_t tidx; bool fw; int64_t col; int row; bool operator==(const Key &o) const { return tidx == o.tidx && fw == o.fw && col == o.col && row == o.row; } };
	struct H { size_t operator()(const Key &k) const { return (size_t)(k.tidx * 1000003 + k.col * 31 + k.row * 2 + (k.fw ? 1 : 0)); } };
	std::unordered_set<Key, H> cells;
	template <typename F> void walk(const Aln &a, F f) const {
		const std::vector<Edit> ned = a.leftToRight();
		int64_t left = a.refoff;
		size_t k = 0;
		const int start = a.trimLeft(), n = start + a.ext();
		for(int i = start; i < n; i++) {
			int64_t diff = 1, right = left + 1;
			while(k < ned.size() && ned[k].pos == i) { if(ned[k].type == 2) diff = 0; k++; }
			if(i < n - 1) { size_t k2 = k; while(k2 < ned.size() && ned[k2].pos == i + 1) { if(ned[k2].type == 1) right++; k2++; } }
			for(int64_t j = left; j < right; j++) if(!f(Key{a.tidx, a.fw, j, i})) return;
			left = right + diff - 1;
		}
	}
	bool overlap(const Aln &a) const { bool o = false; walk(a, [&](const Key &k) { if(cells.count(k)) { o = true; return false; } return true; }); return o; }
	void add(const Aln &a) { walk(a, [&](const Key &k) { cells.insert(k); return true; }); }
};

// ---------------------------------------------------------------------------------------------- requests
enum ReqKind { RQ_EXACT_SWEEP, RQ_ONE_MM, RQ_SEED_SEARCH, RQ_EXTEND, RQ_RESOLVE, RQ_UNGAPPED, RQ_DP, RQ_NKINDS };

struct DpOut {
	bool found = false; int64_t best = 0;
	std::vector<std::pair<int64_t, int>> attempts;      // (candidate score, alignment index or -1)
	std::vector<Aln> alns;
	size_t cursor = 0; bool u8 = true;
};

struct Req {
	int kind = 0;
	int read = 0;                                        // index of the read (mate) in the input batch
	// inputs
	int nofw = 0, norc = 0, L = 0, interval = 1, offset = 0, fw = 1, rdoff = 0, seedlen = 0;
	int64_t minsc = 0, tidx = 0, refoff = 0, tlen = 0;
	uint64_t row = 0, rng4[4] = {0, 0, 0, 0};
	int qlen = 0, reject = 0;
	bt2g_dp_problem prob{};
	// outputs
	uint64_t nelt = 0; int mined[2] = {0, 0}; uint64_t tb[4] = {0, 0, 0, 0};
	struct MmHit { uint64_t top, bot; int pos, chr, qchr, score; bool fw; };
	std::vector<MmHit> mm;
	std::vector<uint64_t> ranges; int nseeds = 0;        // [2][nseeds][4]
	int nlex = 0, nrex = 0;
	uint64_t joined = 0; bool ok = false, straddled = false; int64_t rtidx = 0, rtoff = 0, rtlen = 0;
	int ugStatus = 0; Aln ugAln;
	DpOut dp;
};

struct ReadCtx;                                          // the per-unit state: which request it is blocked on, where to resume

// ---------------------------------------------------------------------------------------------- coroutine plumbing
template <typename T> struct Task {
	struct promise_type {
		T value{};
		std::coroutine_handle<> cont;
		Task get_return_object() { return Task{std::coroutine_handle<promise_type>::from_promise(*this)}; }
		std::suspend_always initial_suspend() noexcept { return {}; }
		struct Final {
			bool await_ready() noexcept { return false; }
			std::coroutine_handle<> await_suspend(std::coroutine_handle<promise_type> h) noexcept {
				return h.promise().cont ? h.promise().cont : std::noop_coroutine();
			}
			void await_resume() noexcept {}
		};
		Final final_suspend() noexcept { return {}; }
		void return_value(T v) { value = std::move(v); }
		void unhandled_exception() { std::terminate(); }
	};
	std::coroutine_handle<promise_type> h;
	explicit Task(std::coroutine_handle<promise_type> hh) : h(hh) {}
	Task(Task &&o) noexcept : h(o.h) { o.h = nullptr; }
	Task(const Task &) = delete;
	~Task() { if(h) h.destroy(); }
	// awaiting a sub-task: run it now, come back here when it finishes
	bool await_ready() const noexcept { return false; }
	std::coroutine_handle<> await_suspend(std::coroutine_handle<> parent) noexcept { h.promise().cont = parent; return h; }
	T await_resume() { return std::move(h.promise().value); }
};

struct Pending { Req *req = nullptr; std::coroutine_handle<> leaf; };

struct AwaitReq {                                        // co_await AwaitReq{pending slot, request}: suspend until the wave answers it
	Pending *slot; Req *req;
	bool await_ready() const noexcept { return false; }
	void await_suspend(std::coroutine_handle<> h) noexcept { slot->req = req; slot->leaf = h; }
	void await_resume() const noexcept {}
};


// ---------------------------------------------------------------------------------------------- engine
struct Params {                                       // bt2g_policy_params, resolved
	bool local, paired, mmode, all, nofw, norc, discord, mixed;
	int seedLen, seedRounds, streak, maxIters, maxUg, maxDp, maxMateStreak;
	int64_t khits, mhits;
	Func ival, smin, nceil;
	uint32_t seed;
	int matchBonus, mmpMax, mmpMin, nPen, rdgConst, rdgLin, rfgConst, rfgLin;
	bt2g_pe_policy pe;
	int offSize;
	// Scoring arithmetic (scoring.h / scoring.cpp)
	int64_t perfect(int len) const { return (int64_t)len * matchBonus; }
	int64_t minScore(int len) const { const long long m = smin.fi((double)len); return local ? std::max<long long>(m, 0) : std::min<long long>(m, 0); }
	int nCeilRaw(int len) const { return (int)nceil.fi((double)len); }
	int nCeil(int len) const { return std::min(nCeilRaw(len), len); }
	int maxReadGaps(int64_t minsc, int len) const {
		int64_t sc = perfect(len); bool first = true; int num = 0;
		while(sc >= minsc) { sc -= first ? rdgConst + rdgLin : rdgLin; first = false; num++; }
		return num - 1;
	}
	int maxRefGaps(int64_t minsc, int len) const {
		int64_t sc = perfect(len); bool first = true; int num = 0;
		while(sc >= minsc) { sc -= matchBonus; sc -= first ? rfgConst + rfgLin : rfgLin; first = false; num++; }
		return num - 1;
	}
	int mmPenalty(int q) const { const int ii = std::min(std::max(q, 0), 40); const float frac = (float)ii / 40.0f; return mmpMin + (int)(frac * (float)(mmpMax - mmpMin)); }
	int seedInterval(int len, bool both) const { long long v = ival.fi((double)len); if(both) v = (long long)((double)v * 1.2 + 0.5); return (int)std::max<long long>(v, 1); }
};

struct DPRect { int64_t refl, refr, reflPre, refrPre, triml, trimr, corel, corer, maxgap; bool trimmedAway() const { return refr < refl; } };

static bool frameSeedRect(int64_t off, int rdlen, int64_t reflen, int maxrdgap, int maxrfgap, int maxhalf, DPRect &r) {
	// DynProgFramer::frameSeedExtensionRect (dp_framer.cpp:81-129); the gap counts are size_t there: negative wraps to huge
	const uint64_t a = (uint64_t)(int64_t)maxrdgap, b = (uint64_t)(int64_t)maxrfgap;
	const int64_t maxgap = (int64_t)std::min<uint64_t>(std::max(a, b), (uint64_t)maxhalf);
	const int64_t refl = off - 2 * maxgap, refr = off + (rdlen - 1) + 2 * maxgap;
	int64_t triml = 0, trimr = 0;
	if(refr >= reflen) trimr = refr - (reflen - 1);
	if(refl < 0) triml = -refl;
	r = DPRect{refl + triml, refr - trimr, refl, refr, triml, trimr, maxgap, maxgap + 2 * maxgap, maxgap};
	return !r.trimmedAway();
}

struct SatPos { uint64_t topf = 0, topb = 0; int64_t size = 0; int keyLen = 0; bool fw = true; int offidx = 0, rdoff = 0, seedlen = 0, nlex = 0, nrex = 0; int64_t origSize = 0; };
struct EEHit { uint64_t top = 0, bot = 0; bool fw = true; int64_t score = 0; bool hasEdit = false; Edit edit{0, 0, 0, 3};
	int ns() const { return hasEdit && (edit.chr == 'N' || edit.qchr == 'N'); } int refns() const { return hasEdit && edit.chr == 'N'; } };
struct SatEntry { SatPos sp; int ee = -1; Random1toN rands; };        // ee = index into the current hit list, or -1

struct SeedHits { std::vector<uint64_t> hits; int n = 0, interval = 1, offset = 0, seedlen = 0; int64_t nonz = 0, nelt = 0;
	std::vector<int64_t> nfw, nrc; std::vector<std::pair<int, bool>> ranks;
	const uint64_t *at(bool fw, int i) const { return hits.data() + ((size_t)(fw ? 0 : 1) * n + i) * 4; } };

struct Mate {
	const uint8_t *codes = nullptr, *quals = nullptr; std::string name; int idx = 0;
	int rdlen = 0, nceil = 0; int64_t minsc = 0, perfect = 0; bool filt = true;
	std::vector<EEHit> mm1, ee;
	std::vector<std::array<int64_t, 3>> exRanges[2];                    // [fw?0:1]: (p5, len, size)
	IntervalSet seen;
	SeedHits sh; bool hasSh = false;
};

struct UnpairedSink {
	int64_t khits, mhits; bool mmode; std::vector<Aln> alns; bool done = false, exitM = false, exitK = false; int64_t best = MIN_I64, best2 = MIN_I64;
	bool report(const Aln &a) {
		alns.push_back(a);
		if(!done) {
			if(!mmode && (int64_t)alns.size() >= khits) done = exitK = true;
			else if(mmode && (int64_t)alns.size() > mhits) done = exitM = true;
		}
		if(a.score > best) { best2 = best; best = a.score; } else if(a.score > best2) best2 = a.score;
		return done;
	}
};

struct PairedSink {
	int64_t khits, mhits; bool mmode;
	std::vector<Aln> rs1, rs2, rs1u, rs2u;
	bool doneConcord = false, doneDiscord = false, doneUnp[2] = {false, false}, exitConcordM = false, exitConcordK = false, done = false;
	bool exitUnpK[2] = {false, false};
	int64_t nconcord = 0, nunp[2] = {0, 0}, bestPair = MIN_I64, best2Pair = MIN_I64;
	void updateDone() { done = doneUnp[0] && doneUnp[1] && doneDiscord && doneConcord; }
	bool report(const Aln *a1, const Aln *a2) {
		if(a1 && a2) {
			nconcord++;
			if(!mmode && nconcord >= khits) doneConcord = exitConcordK = true;
			else if(mmode && nconcord > mhits) doneConcord = exitConcordM = true;
			doneDiscord = true;
			if(doneConcord && !exitConcordM) doneUnp[0] = doneUnp[1] = true;
			updateDone();
			rs1.push_back(*a1); rs2.push_back(*a2);
			const int64_t sc = a1->score + a2->score;
			if(sc > bestPair) { best2Pair = bestPair; bestPair = sc; } else if(sc > best2Pair) best2Pair = sc;
		} else {
			const int m = a1 ? 0 : 1;
			const Aln &a = a1 ? *a1 : *a2;
			nunp[m]++;
			if(!doneUnp[m]) {
				if(!mmode && nunp[m] >= khits) { doneUnp[m] = true; exitUnpK[m] = true; updateDone(); }
				else if(mmode && nunp[m] > mhits) { doneUnp[m] = true; updateDone(); }
			}
			if(nunp[m] > 1) doneDiscord = true;
			(m == 0 ? rs1u : rs2u).push_back(a);
		}
		return done;
	}
	bool doneWithMate(bool mate1) const {
		const int m = mate1 ? 0 : 1;
		if(!doneUnp[m] || !doneConcord) return false;
		if(!doneDiscord && nunp[m] == 0) return false;
		return true;
	}
};

struct Result { bool aligned = false; Aln aln; bool hasXs = false; int64_t xs = 0; int mapq = 0; std::vector<Aln> secondary; };
// secPairs (-k / -a): the further selected concordant pairs in report order; Result::secondary of a mate: its further unpaired alignments
struct PairOut { int pairType = 0; Result m[2]; int kind = 5; int64_t scoreSum = 0, fraglen = 0; std::vector<std::pair<Aln, Aln>> secPairs; };

struct Engine {
	const Params &P; Pending *slot;
	Rng rnd; RedundantAlns red, redMate[2]; Mate m[2]; Mate *cur = nullptr;
	UnpairedSink usink{1, 50, true}; PairedSink psink{1, 50, true};
	int64_t nIters = 0, nDps = 0, nUgs = 0, nMateDps = 0; int64_t streakCur = 0;
	Engine(const Params &p, Pending *s) : P(p), slot(s) {}

	static uint32_t genRandSeed(const uint8_t *codes, const uint8_t *quals, int len, const std::string &name, uint32_t seed) {
		uint32_t rseed = (seed + 101u) * 59u * 61u * 67u * 71u * 73u * 79u * 83u;
		for(int i = 0; i < len; i++) rseed ^= (uint32_t)codes[i] << ((i & 15) << 1);
		for(int i = 0; i < len; i++) rseed ^= (uint32_t)quals[i] << ((i & 3) << 3);
		for(size_t i = 0; i < name.size(); i++) { if(name[i] == '/') break; rseed ^= (uint32_t)(unsigned char)name[i] << ((i & 3) << 3); }
		return rseed;
	}
	int64_t mapq(int64_t best, bool hasSec, int64_t sec, int64_t scMin, int64_t perfect) const {
		if(!P.mmode && !hasSec) return 255;
		return mapq_v2(best, hasSec, sec, scMin, perfect, !P.local);
	}
	void rankSeedHits(SeedHits &sh) {                               // SeedResults::rankSeedHits (aligner_seed.h:1019-1080)
		const int num = sh.n;
		sh.ranks.clear();
		if(P.all) {
			for(int i = 1; i < num; i++) for(int f = 0; f < 2; f++) if((f == 0 ? sh.nfw : sh.nrc)[i] > 0) sh.ranks.push_back({i, f == 0});
			if(num && sh.nfw[0] > 0) sh.ranks.push_back({0, true});
			if(num && sh.nrc[0] > 0) sh.ranks.push_back({0, false});
			return;
		}
		std::vector<char> sfw(num, 0), src(num, 0);
		while((int64_t)sh.ranks.size() < sh.nonz) {
			int64_t minsz = 0xffffffffll; int minidx = 0; bool minfw = true;
			const int rb = rnd.boolean();
			for(int fwi = 0; fwi < 2; fwi++) {
				const bool fw = fwi == (rb ? 1 : 0);
				const std::vector<int64_t> &rrs = fw ? sh.nfw : sh.nrc;
				const std::vector<char> &srt = fw ? sfw : src;
				int i = (int)(rnd.u32() % (uint32_t)num);
				for(int t = 0; t < num; t++) {
					if(rrs[i] > 0 && !srt[i] && rrs[i] < minsz) { minsz = rrs[i]; minidx = i; minfw = fw; }
					if(++i == num) i = 0;
				}
			}
			(minfw ? sfw : src)[minidx] = 1;
			sh.ranks.push_back({minidx, minfw});
		}
	}
	static void fillSeedHits(SeedHits &sh, const Req &r, int interval, int offset, int seedlen) {
		sh.hits = r.ranges; sh.n = r.nseeds; sh.interval = interval; sh.offset = offset; sh.seedlen = seedlen;
		sh.nfw.assign(sh.n, 0); sh.nrc.assign(sh.n, 0); sh.nonz = 0; sh.nelt = 0;
		for(int f = 0; f < 2; f++) for(int i = 0; i < sh.n; i++) {
			const uint64_t *h = sh.at(f == 0, i);
			const int64_t sz = h[1] > h[0] ? (int64_t)(h[1] - h[0]) : 0;
			(f == 0 ? sh.nfw : sh.nrc)[i] = sz;
			if(sz > 0) { sh.nonz++; sh.nelt += sz; }
		}
	}

	// ---- eeSaTups (aligner_sw_driver.cpp:66-290)
	void eeAdd(std::vector<SatEntry> &out, const std::vector<EEHit> &hits, int hi, int64_t &nelt, int64_t maxelt, bool &done) {
		const EEHit &hit = hits[hi];
		uint64_t tops[2] = {hit.top, 0}, bots[2] = {hit.bot, 0};
		const int64_t width = (int64_t)(hit.bot - hit.top);
		if(width <= 0) return;
		if(nelt + width > maxelt) {
			const int64_t trim = (nelt + width) - maxelt;
			const uint64_t rn = (P.offSize == 4 ? (uint64_t)rnd.u32() : rnd.u64()) % (uint64_t)width;
			const int64_t newwidth = width - trim;
			if(hit.top + rn + newwidth > hit.bot) { tops[0] = hit.top + rn; bots[0] = hit.bot; tops[1] = hit.top; bots[1] = hit.top + newwidth - (bots[0] - tops[0]); }
			else { tops[0] = hit.top + rn; bots[0] = tops[0] + newwidth; }
		}
		for(int i = 0; i < 2; i++) {
			if(done || bots[i] <= tops[i]) break;
			const int64_t w = (int64_t)(bots[i] - tops[i]);
			SatEntry e; e.sp.topf = tops[i]; e.sp.size = w; e.sp.keyLen = cur->rdlen; e.sp.fw = hit.fw; e.sp.seedlen = cur->rdlen; e.sp.origSize = w; e.ee = hi;
			e.rands.init((size_t)w, P.all);
			out.push_back(std::move(e));
			nelt += w;
			if(nelt >= maxelt) done = true;
		}
	}
	// hits = exact hits (strand order drawn here) or the mate's sorted 1-mismatch hits; returns the entry list and the hit list it indexes
	void eeSaTups(std::vector<EEHit> &eeExact, std::vector<SatEntry> &out, std::vector<EEHit> &hitList) {
		out.clear(); hitList.clear();
		int64_t nelt = 0; bool done = false;
		int64_t tot = 0, fwsz = 0;
		for(const EEHit &h : eeExact) { tot += (int64_t)(h.bot - h.top); if(h.fw) fwsz += (int64_t)(h.bot - h.top); }
		if(tot > 0) {
			const uint64_t rn = (P.offSize == 4 ? (uint64_t)rnd.u32() : rnd.u64()) % (uint64_t)tot;
			const bool fwFirst = !((int64_t)rn >= fwsz);
			for(int fwi = 0; fwi < 2 && !done; fwi++) {
				const bool fw = (fwi == 0) == fwFirst;
				for(const EEHit &h : eeExact) if(h.fw == fw) { hitList.push_back(h); eeAdd(out, hitList, (int)hitList.size() - 1, nelt, P.maxIters, done); break; }
			}
		}
		if(!done && !cur->mm1.empty()) {
			std::stable_sort(cur->mm1.begin(), cur->mm1.end(), [](const EEHit &a, const EEHit &b) { return a.score > b.score; });
			shuffleEqualStreaks(cur->mm1, [](const EEHit &h) { return h.score; }, rnd);
			for(const EEHit &h : cur->mm1) { if(done) break; hitList.push_back(h); eeAdd(out, hitList, (int)hitList.size() - 1, nelt, P.maxIters, done); }
		}
	}

	// ---- prioritizeSATupsRands (aligner_sw_driver.cpp:490-725)
	Task<int64_t> prioritize(SeedHits &sh, std::vector<SatEntry> &out) {
		out.clear();
		std::vector<SatPos> sats;
		int64_t nelt = 0;
		for(auto [offidx, fw] : sh.ranks) {
			const uint64_t *h = sh.at(fw, offidx);
			const int64_t sz = (int64_t)(h[1] - h[0]);
			const int rdoff = sh.offset + offidx * sh.interval, seedlen = sh.seedlen;
			nelt += sz;
			auto &rng = cur->exRanges[fw ? 0 : 1];
			bool skip = false;
			for(auto &x : rng) if(x[0] <= rdoff && x[0] + x[1] >= rdoff + seedlen && sz <= x[2]) { skip = true; break; }
			if(skip) { nelt -= sz; continue; }
			SatPos sp; sp.topf = h[0]; sp.topb = h[2]; sp.size = sz; sp.keyLen = seedlen; sp.fw = fw; sp.offidx = offidx; sp.rdoff = rdoff; sp.seedlen = seedlen; sp.origSize = sz;
			Req rq; rq.kind = RQ_EXTEND; rq.read = cur->idx; rq.fw = fw; rq.rdoff = rdoff; rq.seedlen = seedlen;
			for(int k = 0; k < 4; k++) rq.rng4[k] = h[k];
			co_await AwaitReq{slot, &rq};
			sp.nlex = rq.nlex; sp.nrex = rq.nrex;
			if(sp.nlex > 0 || sp.nrex > 0) rng.push_back({(int64_t)rdoff - (fw ? sp.nlex : sp.nrex), (int64_t)seedlen + sp.nlex + sp.nrex, sz});
			sats.push_back(sp);
		}
		const int nsm = 5;
		size_t nsmall = 0;
		for(const SatPos &s : sats) nsmall += s.size <= nsm;
		std::sort(sats.begin(), sats.end(), [](const SatPos &a, const SatPos &b) {
			if(a.size != b.size) return a.size < b.size;
			if(a.topf != b.topf) return a.topf < b.topf;
			if(a.offidx != b.offidx) return a.offidx < b.offidx;
			if(a.rdoff != b.rdoff) return a.rdoff < b.rdoff;
			if(a.seedlen != b.seedlen) return a.seedlen < b.seedlen;
			return a.fw && !b.fw;
		});
		int64_t added = 0;
		size_t j = 0;
		while(j < nsmall && added < P.maxIters) {
			SatEntry e; e.sp = sats[j]; e.rands.init((size_t)sats[j].size, P.all);
			out.push_back(std::move(e));
			added += sats[j].size; j++;
		}
		if(added >= P.maxIters || nsmall == sats.size()) co_return added;
		// RowSampler (aligner_sw_driver.h:179-256)
		const size_t nl = sats.size() - nsmall;
		std::vector<double> masses(nl); std::vector<char> elim(nl, 0); double mass = 0.0;
		for(size_t i = 0; i < nl; i++) {
			const SatPos &s = sats[nsmall + i];
			double num = (double)(s.nlex + s.nrex + 1); num *= num;
			double den = (double)s.size; den *= den;
			masses[i] = num / den; mass += masses[i];
		}
		std::vector<Random1toN> rands2(sats.size());
		while(added < P.maxIters && added < nelt) {
			const double rd = rnd.flt() * mass;
			double sofar = 0.0; size_t pick = 0, last = 0;
			bool got = false;
			for(size_t i = 0; i < nl; i++) if(!elim[i]) { last = i; sofar += masses[i]; if(rd < sofar) { pick = i; got = true; break; } }
			if(!got) pick = last;
			const size_t ri = pick + nsmall;
			if(!rands2[ri].inited()) rands2[ri].init((size_t)sats[ri].size, P.all);
			const size_t r = rands2[ri].next(rnd);
			if(rands2[ri].done()) { elim[pick] = 1; mass -= masses[pick]; }
			SatEntry e; e.sp = sats[ri]; e.sp.topf = sats[ri].topf + r; e.sp.topb = 0; e.sp.size = 1;
			e.rands.init(1, P.all);
			out.push_back(std::move(e));
			added++;
		}
		co_return added;
	}

	bool dpU8(const DpOut &dp, int64_t minsc, const Mate &mt) const {
		if(!P.local) return minsc >= -254;
		int bias = P.nPen;
		for(int i = 0; i < mt.rdlen; i++) bias = std::max(bias, P.mmPenalty((int)mt.quals[i] - 33));
		return dp.best + bias < 255;
	}
	// SwAligner::nextAlignment over the DP's attempt list: candidates below minsc are skipped, every attempt reseeds the RNG
	bool nextAlignment(DpOut &dp, int64_t minsc, Aln &out) {
		while(dp.cursor < dp.attempts.size()) {
			const auto [candScore, ai] = dp.attempts[dp.cursor++];
			if(candScore < minsc) continue;
			const uint32_t reseed = rnd.u32() + 1u;
			rnd.init(dp.u8 ? reseed + 1u : reseed);
			if(ai >= 0) { out = dp.alns[ai]; return true; }
		}
		return false;
	}
	Aln eeAln(const EEHit &h, int64_t tidx, int64_t refoff, bool fw, int rdlen) const {
		Aln a; a.tidx = tidx; a.refoff = refoff; a.fw = fw; a.score = h.score; a.rdlen = rdlen; a.ns = h.ns(); a.refns = h.refns();
		if(h.hasEdit) a.edits.push_back(h.edit);
		return a;
	}
	void tightenUnpaired() {
		if(!(P.mmode && usink.best2 != MIN_I64)) return;
		const int64_t bot = usink.best2 + ((usink.best - usink.best2) * 3) / 4;      // tighten == 3
		if(bot >= cur->minsc) { cur->minsc = bot; if(cur->minsc < cur->perfect) cur->minsc++; }
	}

	Task<int> extendSeeds(SeedHits *sh, std::vector<EEHit> eeExact);
	Task<Result> readSteps(int idx, const uint8_t *codes, const uint8_t *quals, int len, std::string name);
	Task<int> extendSeedsPaired(int ai, SeedHits *sh, std::vector<EEHit> eeExact);
	Task<PairOut> pairSteps(int idx1, const uint8_t *c1, const uint8_t *q1, int l1, std::string n1, const uint8_t *c2, const uint8_t *q2, int l2, std::string n2);
	Result finishRead();
	PairOut finishPair();
};


// ---------------------------------------------------------------------------------------------- extendSeeds (unpaired)
Task<int> Engine::extendSeeds(SeedHits *sh, std::vector<EEHit> eeExact) {
	Mate &c = *cur;
	const int rdlen = c.rdlen;
	const int64_t nonz = sh ? sh->nonz : 0;
	bool eeMode = !eeExact.empty() || !c.mm1.empty(), firstEe = true, firstExtend = true;
	int64_t nUgFail = 0, nDpFail = 0, neltLeft = 0;
	std::vector<SatEntry> satpos; std::vector<EEHit> hitList;
	for(;;) {
		if(eeMode) { if(firstEe) { firstEe = false; eeSaTups(eeExact, satpos, hitList); } else eeMode = false; }
		if(!eeMode) {
			if(nonz == 0) co_return EXHAUSTED;
			if(c.minsc == c.perfect) co_return PERFECT;
			if(firstExtend) { neltLeft = co_await prioritize(*sh, satpos); firstExtend = false; }
			if(neltLeft == 0) break;
		}
		for(size_t si = 0; si < satpos.size(); si++) {
			SatEntry &se = satpos[si];
			const EEHit *eh = eeMode ? &hitList[se.ee] : nullptr;
			if(eeMode && eh->score < c.minsc) co_return PERFECT;
			const bool isSmall = se.sp.size < 5, fw = se.sp.fw;
			int rdoff = se.sp.rdoff;
			if(!fw) rdoff = rdlen - rdoff - se.sp.seedlen;
			bool first = true;
			while(!se.rands.done() && (first || isSmall || eeMode)) {
				if(c.minsc == c.perfect) { if(!eeMode || eh->score < c.perfect) co_return PERFECT; }
				else if(eeMode && eh->score < c.minsc) break;
				if(nDps >= P.maxDp || nUgs >= P.maxUg || nIters >= P.maxIters) co_return HARD_LIMIT;
				nIters++; first = false;
				const size_t elt = se.rands.next(rnd);
				Req rq; rq.kind = RQ_RESOLVE; rq.row = se.sp.topf + elt; rq.qlen = se.sp.keyLen; rq.reject = eeMode;
				co_await AwaitReq{slot, &rq};
				if(!eeMode) neltLeft--;
				if(!rq.ok) continue;
				const int64_t tidx = rq.rtidx, toff = rq.rtoff, tlen = rq.rtlen, refoff = toff - rdoff;
				if(c.seen.present(tidx, fw, refoff)) continue;
				int readGaps = 0, refGaps = 0; bool ungapped = false;
				if(!eeMode) { readGaps = P.maxReadGaps(c.minsc, rdlen); refGaps = P.maxRefGaps(c.minsc, rdlen); ungapped = readGaps == 0 && refGaps == 0; }
				int state = 0; Aln fixed; Req dq;
				if(eeMode) { fixed = eeAln(*eh, tidx, refoff, fw, rdlen); state = 1; c.seen.add(tidx, fw, refoff, 1); }
				else if(ungapped) {
					Req uq; uq.kind = RQ_UNGAPPED; uq.read = c.idx; uq.fw = fw; uq.tidx = tidx; uq.refoff = refoff; uq.tlen = tlen; uq.minsc = c.minsc;
					co_await AwaitReq{slot, &uq};
					c.seen.add(tidx, fw, refoff, 1);
					nUgs++;
					if(uq.ugStatus == 0) { if(++nUgFail >= P.streak) co_return SOFT_LIMIT; continue; }
					else if(uq.ugStatus == -1) { if(++nUgFail >= P.streak) co_return SOFT_LIMIT; }
					else { nUgFail = 0; fixed = uq.ugAln; state = 2; }
				}
				if(state == 0) {
					DPRect rect;
					const bool found = frameSeedRect(refoff, rdlen, tlen, readGaps, refGaps, 15, rect);
					c.seen.add(tidx, fw, refoff, 1);
					if(!found) continue;
					c.seen.add(tidx, fw, rect.reflPre + rect.corel, rect.corer - rect.corel + 1);
					dq.kind = RQ_DP; dq.read = c.idx; dq.fw = fw; dq.tidx = tidx; dq.minsc = c.minsc;
					dq.prob = bt2g_dp_problem{}; dq.prob.fw = fw; dq.prob.tidx = (uint64_t)tidx; dq.prob.refl = rect.refl; dq.prob.refr = rect.refr;
					dq.prob.triml = (int32_t)rect.triml; dq.prob.corel = (int32_t)rect.corel; dq.prob.corer = (int32_t)rect.corer;
					dq.prob.minsc = (int32_t)c.minsc; dq.prob.nceil = P.nCeilRaw(rdlen);
					co_await AwaitReq{slot, &dq};
					nDps++;
					if(!dq.dp.found) { if(++nDpFail >= P.streak) co_return SOFT_LIMIT; continue; }
					nDpFail = 0;
					dq.dp.cursor = 0; dq.dp.u8 = dpU8(dq.dp, c.minsc, c);
				}
				bool firstInner = true;
				for(;;) {
					Aln a;
					if(state != 0) { if(!firstInner) break; a = fixed; }
					else if(!nextAlignment(dq.dp, c.minsc, a)) break;
					firstInner = false;
					if(red.overlap(a)) continue;
					red.add(a);
					if(usink.report(a)) co_return FULFILLED;
					tightenUnpaired();
				}
			}
		}
	}
	co_return EXHAUSTED;
}

Task<Result> Engine::readSteps(int idx, const uint8_t *codes, const uint8_t *quals, int len, std::string name) {
	Result none;
	int ns = 0;
	for(int i = 0; i < len; i++) ns += codes[i] > 3;
	if(len < 2 || ns > P.nCeil(len) || P.perfect(len) < P.minScore(len)) co_return none;
	Mate &c = m[0]; cur = &c;
	c.codes = codes; c.quals = quals; c.name = name; c.idx = idx; c.rdlen = len; c.minsc = P.minScore(len); c.perfect = P.perfect(len); c.nceil = P.nCeil(len);
	rnd.init(genRandSeed(codes, quals, len, name, P.seed));
	const int interval = P.seedInterval(len, false);
	usink = UnpairedSink{P.khits, P.mhits, P.mmode};
	bool done = false;
	auto after = [&](int ret, bool checkPerfect) {
		if(ret == FULFILLED) { if(usink.done) done = true; }
		else if(ret == PERFECT || ret == HARD_LIMIT) done = true;
		if(checkPerfect && !done && c.minsc == c.perfect) done = true;
	};
	Req sw; sw.kind = RQ_EXACT_SWEEP; sw.read = idx; sw.nofw = P.nofw; sw.norc = P.norc;
	co_await AwaitReq{slot, &sw};
	if(sw.nelt > 0) {
		std::vector<EEHit> ee;
		if(sw.tb[1] > sw.tb[0]) ee.push_back(EEHit{sw.tb[0], sw.tb[1], true, c.perfect});
		if(sw.tb[3] > sw.tb[2]) ee.push_back(EEHit{sw.tb[2], sw.tb[3], false, c.perfect});
		after(co_await extendSeeds(nullptr, ee), true);
	}
	if(!done) {
		const bool yfw = sw.mined[0] <= 1 && !P.nofw, yrc = sw.mined[1] <= 1 && !P.norc;
		if(yfw || yrc) {
			Req mq; mq.kind = RQ_ONE_MM; mq.read = idx; mq.minsc = c.minsc; mq.nofw = !yfw; mq.norc = !yrc;
			co_await AwaitReq{slot, &mq};
			c.mm1.clear();
			for(const Req::MmHit &h : mq.mm) { EEHit e{h.top, h.bot, h.fw, h.score, true, Edit{h.pos, h.chr, h.qchr, 3}}; c.mm1.push_back(e); }
			if(!c.mm1.empty() && !usink.done) { const int ret = co_await extendSeeds(nullptr, {}); c.mm1.clear(); after(ret, true); }
			else if(!c.mm1.empty()) done = true;
		}
	}
	const int nrounds = std::min(P.seedRounds, interval), L = P.seedLen;
	for(int roundi = 0; roundi < P.seedRounds; roundi++) {
		if(done || usink.done) { done = true; break; }
		if(roundi >= nrounds || interval <= roundi) continue;
		const int offset = (interval * roundi) / nrounds;
		if(offset > 0 && std::min(L, len) + offset > len) continue;
		Req sq; sq.kind = RQ_SEED_SEARCH; sq.read = idx; sq.L = std::min(L, len); sq.interval = interval; sq.offset = offset; sq.nofw = P.nofw; sq.norc = P.norc;
		co_await AwaitReq{slot, &sq};
		fillSeedHits(c.sh, sq, interval, offset, std::min(L, len));
		if(c.sh.nonz == 0) { done = true; break; }
		rankSeedHits(c.sh);
		after(co_await extendSeeds(&c.sh, {}), false);
		if(!done && c.sh.nelt / c.sh.nonz < 300) done = true;
	}
	co_return finishRead();
}

Result Engine::finishRead() {
	Result r;
	std::vector<Aln> &alns = usink.alns;
	if(alns.empty()) return r;
	std::vector<std::pair<int64_t, int>> buf;
	for(size_t i = 0; i < alns.size(); i++) buf.push_back({alns[i].score, (int)i});
	std::sort(buf.begin(), buf.end(), [](auto &a, auto &b) { return a > b; });          // score desc, index desc within ties
	shuffleEqualStreaks(buf, [](const std::pair<int64_t, int> &t) { return t.first; }, rnd);
	r.aligned = true; r.aln = alns[buf[0].second];
	r.hasXs = buf.size() > 1; r.xs = r.hasXs ? buf[1].first : 0;
	r.mapq = (int)mapq(r.aln.score, r.hasXs, r.xs, P.minScore(cur->rdlen), cur->perfect);
	// ReportingState::getReport (aln_sink.cpp:300-330): after a -k short circuit khits alignments, else min(found, khits)
	const int64_t num = usink.exitK ? P.khits : std::min<int64_t>((int64_t)alns.size(), P.khits);
	for(int64_t i = 1; i < std::min<int64_t>(num, (int64_t)buf.size()); i++) r.secondary.push_back(alns[buf[(size_t)i].second]);
	return r;
}


// ---------------------------------------------------------------------------------------------- pairs
static bool frameMateRect(bool anchorLeft, int64_t ll, int64_t lr, int64_t rl, int64_t rr, int rdlen, int64_t reflen, int maxrdgap, int maxrfgap,
                          int maxhalf, DPRect &r) {
	// DynProgFramer::frameFindMateRect (dp_framer.cpp:177-361): maxgap = max(gaps, maxhalf)
	const uint64_t a = (uint64_t)(int64_t)maxrdgap, b = (uint64_t)(int64_t)maxrfgap;
	const int64_t maxgap = (int64_t)std::max<uint64_t>(std::max(a, b), (uint64_t)maxhalf);
	int64_t refl, refr;
	if(anchorLeft) { refl = (rl - (rdlen - 1)) - maxgap; refr = rr + maxgap; }
	else { refl = ll - maxgap; refr = (lr + (rdlen - 1)) + maxgap; }
	int64_t triml = 0, trimr = 0;
	if(refr >= reflen) trimr = refr - (reflen - 1);
	if(refl < 0) triml = -refl;
	const int64_t width = refr - refl + 1;
	r = DPRect{refl + triml, refr - trimr, refl, refr, triml, trimr, maxgap, width - maxgap - 1, maxgap};
	return !r.trimmedAway();
}

Task<int> Engine::extendSeedsPaired(int ai, SeedHits *sh, std::vector<EEHit> eeExact) {
	const bool anchor1 = ai == 0;
	Mate &c = m[ai], &o = m[ai ^ 1]; cur = &c;
	const int rdlen = c.rdlen, ordlen = o.rdlen;
	const bool oppFilt = !o.filt;
	const int64_t operfect = o.perfect, bestPairScore = c.perfect + operfect;
	auto tightened = [&]() {
		int64_t ps = psink.best2Pair + ((psink.bestPair - psink.best2Pair) * 3) / 4;     // tighten == 3
		if(ps < bestPairScore) ps++;
		return ps;
	};
	const bool canTighten = P.mmode;
	if(canTighten && psink.best2Pair != MIN_I64) { const int64_t nc = tightened() - operfect; if(nc > c.minsc) c.minsc = nc; }
	const int64_t nonz = sh ? sh->nonz : 0;
	bool eeMode = !eeExact.empty() || !c.mm1.empty(), firstEe = true, firstExtend = true, swMateImmediately = true;
	int64_t nEeFail = 0, nUgFail = 0, nDpFail = 0, neltLeft = 0;
	const int64_t streak = streakCur;
	std::vector<SatEntry> satpos; std::vector<EEHit> hitList; std::vector<int64_t> mateStreaks;
	for(;;) {
		if(eeMode) { if(firstEe) { firstEe = false; eeSaTups(eeExact, satpos, hitList); mateStreaks.assign(satpos.size(), 0); } else eeMode = false; }
		if(!eeMode) {
			if(nonz == 0) co_return EXHAUSTED;
			if(P.mmode && c.minsc == c.perfect) co_return PERFECT;
			if(firstExtend) { neltLeft = co_await prioritize(*sh, satpos); firstExtend = false; mateStreaks.assign(satpos.size(), 0); }
			if(neltLeft == 0) break;
		}
		for(size_t si = 0; si < satpos.size(); si++) {
			SatEntry &se = satpos[si];
			const EEHit *eh = eeMode ? &hitList[se.ee] : nullptr;
			if(eeMode && eh->score < c.minsc) co_return PERFECT;
			const bool isSmall = se.sp.size < 5, fw = se.sp.fw;
			int rdoff = se.sp.rdoff;
			if(!fw) rdoff = rdlen - rdoff - se.sp.seedlen;
			bool first = true;
			while(!se.rands.done() && (first || isSmall || eeMode)) {
				if(c.minsc == c.perfect) { if(!eeMode || eh->score < c.perfect) co_return PERFECT; }
				else if(eeMode && eh->score < c.minsc) break;
				if(nDps >= P.maxDp || nMateDps >= P.maxDp || nUgs >= P.maxUg || nIters >= P.maxIters) co_return HARD_LIMIT;
				if(eeMode && nEeFail >= streak) co_return SOFT_LIMIT;
				if(!eeMode && (nDpFail >= streak || nUgFail >= streak)) co_return SOFT_LIMIT;
				if(mateStreaks[si] >= P.maxMateStreak) { se.rands.setDone(); break; }
				nIters++; first = false;
				const size_t elt = se.rands.next(rnd);
				Req rq; rq.kind = RQ_RESOLVE; rq.row = se.sp.topf + elt; rq.qlen = se.sp.keyLen; rq.reject = eeMode;
				co_await AwaitReq{slot, &rq};
				neltLeft--;
				if(!rq.ok) continue;
				const int64_t tidx = rq.rtidx, toff = rq.rtoff, tlen = rq.rtlen, refoff = toff - rdoff;
				if(c.seen.present(tidx, fw, refoff)) continue;
				int readGaps = 0, refGaps = 0; bool ungapped = false;
				if(!eeMode) { readGaps = P.maxReadGaps(c.minsc, rdlen); refGaps = P.maxRefGaps(c.minsc, rdlen); ungapped = readGaps == 0 && refGaps == 0; }
				int state = 0; Aln fixed; Req dq;
				if(eeMode) { fixed = eeAln(*eh, tidx, refoff, fw, rdlen); state = 1; c.seen.add(tidx, fw, refoff, 1); nEeFail++; }
				else if(ungapped) {
					Req uq; uq.kind = RQ_UNGAPPED; uq.read = c.idx; uq.fw = fw; uq.tidx = tidx; uq.refoff = refoff; uq.tlen = tlen; uq.minsc = c.minsc;
					co_await AwaitReq{slot, &uq};
					c.seen.add(tidx, fw, refoff, 1);
					nUgs++; nUgFail++;
					if(uq.ugStatus == 0) continue;
					if(uq.ugStatus == 1) { fixed = uq.ugAln; state = 2; }
				}
				if(state == 0) {
					DPRect rect;
					const bool found = frameSeedRect(refoff, rdlen, tlen, readGaps, refGaps, 15, rect);
					c.seen.add(tidx, fw, refoff, 1);
					if(!found) continue;
					c.seen.add(tidx, fw, rect.reflPre + rect.corel, rect.corer - rect.corel + 1);
					dq.kind = RQ_DP; dq.read = c.idx;
					dq.prob = bt2g_dp_problem{}; dq.prob.fw = fw; dq.prob.tidx = (uint64_t)tidx; dq.prob.refl = rect.refl; dq.prob.refr = rect.refr;
					dq.prob.triml = (int32_t)rect.triml; dq.prob.corel = (int32_t)rect.corel; dq.prob.corer = (int32_t)rect.corer;
					dq.prob.minsc = (int32_t)c.minsc; dq.prob.nceil = P.nCeilRaw(rdlen);
					co_await AwaitReq{slot, &dq};
					nDps++; nDpFail++;
					if(!dq.dp.found) continue;
					dq.dp.cursor = 0; dq.dp.u8 = dpU8(dq.dp, c.minsc, c);
				}
				bool firstInner = true, foundConcordant = false;
				for(;;) {
					Aln a;
					if(state != 0) { if(!firstInner) break; a = fixed; }
					else if(!nextAlignment(dq.dp, c.minsc, a)) break;
					firstInner = false;
					if(red.overlap(a)) continue;
					red.add(a);
					if(psink.doneWithMate(!anchor1) && !psink.doneWithMate(anchor1)) swMateImmediately = false;
					if(swMateImmediately) {
						bool foundMate = !oppFilt;
						int64_t ominscCur = o.minsc;
						Req oq; bool oleft = false, ofw = false; int64_t oll = 0, olr = 0, orl = 0, orr = 0; int ordgaps = 0, orfgaps = 0;
						if(foundMate) {
							if(canTighten && psink.best2Pair != MIN_I64) { const int64_t nc = tightened() - a.score; if(nc > ominscCur) ominscCur = nc; }
							ordgaps = P.maxReadGaps(ominscCur, ordlen); orfgaps = P.maxRefGaps(ominscCur, ordlen);
							foundMate = pe_other_mate(P.pe, anchor1, fw, a.refoff, (int64_t)ordlen + ordgaps, (uint64_t)(anchor1 ? rdlen : ordlen),
							                          (uint64_t)(anchor1 ? ordlen : rdlen), oleft, oll, olr, orl, orr, ofw);
						}
						DPRect orect{};
						if(foundMate) foundMate = frameMateRect(!oleft, oll, olr, orl, orr, ordlen, tlen, ordgaps, orfgaps, 15, orect);
						if(foundMate) {
							oq.kind = RQ_DP; oq.read = o.idx;
							oq.prob = bt2g_dp_problem{}; oq.prob.fw = ofw; oq.prob.tidx = (uint64_t)tidx; oq.prob.refl = orect.refl; oq.prob.refr = orect.refr;
							oq.prob.triml = (int32_t)orect.triml; oq.prob.corel = (int32_t)orect.corel; oq.prob.corer = (int32_t)orect.corer;
							oq.prob.minsc = (int32_t)ominscCur; oq.prob.nceil = P.nCeilRaw(ordlen);
							co_await AwaitReq{slot, &oq};
							nMateDps++;
							foundMate = oq.dp.found;
							if(foundMate) { oq.dp.cursor = 0; oq.dp.u8 = dpU8(oq.dp, ominscCur, o); }
						}
						bool didAnchor = false, brk = false;
						for(;;) {
							Aln oa; bool haveOa = false;
							if(foundMate) { haveOa = nextAlignment(oq.dp, ominscCur, oa); foundMate = haveOa; }
							int64_t oext = 0;
							if(foundMate) {
								if(!red.overlap(oa)) red.add(oa);
								oext = oa.refExtent();
								if(oa.refoff < 0 || oa.refoff + oext > tlen) foundMate = false;
							}
							int pairCl = 5;
							if(foundMate) {
								const int64_t aext = a.refExtent();
								const Aln &a1 = anchor1 ? a : oa, &a2 = anchor1 ? oa : a;
								pairCl = pe_classify(P.pe, a1.refoff, (uint64_t)(anchor1 ? aext : oext), a1.fw, a2.refoff, (uint64_t)(anchor1 ? oext : aext), a2.fw);
							}
							if(psink.doneConcord) foundMate = false;
							if(foundMate) {
								bool doneUnpaired = false;
								if(!anchor1 || !didAnchor) {
									if(anchor1) didAnchor = true;
									const Aln &r1 = anchor1 ? a : oa;
									if(!redMate[0].overlap(r1)) { redMate[0].add(r1); if(psink.report(&r1, nullptr)) doneUnpaired = true; }
								}
								if(anchor1 || !didAnchor) {
									if(!anchor1) didAnchor = true;
									const Aln &r2 = anchor1 ? oa : a;
									if(!redMate[1].overlap(r2)) { redMate[1].add(r2); if(psink.report(nullptr, &r2)) doneUnpaired = true; }
								}
								bool donePaired = false;
								if(pairCl != 5) {
									foundConcordant = true;
									if(psink.report(anchor1 ? &a : &oa, anchor1 ? &oa : &a)) donePaired = true;
									else if(canTighten && psink.best2Pair != MIN_I64) {
										const int64_t nc = tightened() - operfect;
										if(nc > c.minsc) { c.minsc = nc; if(c.minsc > a.score) brk = true; }
									}
								}
								if(brk) break;
								if(donePaired || doneUnpaired) co_return FULFILLED;
								if(psink.doneWithMate(anchor1)) co_return FULFILLED;
							} else if((P.mixed || P.discord) && !didAnchor) {
								didAnchor = true;
								if(!psink.doneUnp[anchor1 ? 0 : 1]) {
									RedundantAlns &rm = redMate[anchor1 ? 0 : 1];
									if(!rm.overlap(a)) { rm.add(a); if(psink.report(anchor1 ? &a : nullptr, anchor1 ? nullptr : &a)) co_return FULFILLED; }
								}
								if(psink.doneWithMate(anchor1)) co_return FULFILLED;
							}
							if(!haveOa) break;
						}
					} else if(P.mixed || P.discord) {
						if(!psink.doneUnp[anchor1 ? 0 : 1]) {
							RedundantAlns &rm = redMate[anchor1 ? 0 : 1];
							if(!rm.overlap(a)) { rm.add(a); if(psink.report(anchor1 ? &a : nullptr, anchor1 ? nullptr : &a)) co_return FULFILLED; }
						}
						if(psink.doneWithMate(anchor1)) co_return FULFILLED;
					}
				}
				if(foundConcordant) { mateStreaks[si] = 0; if(state == 2) nUgFail = 0; else if(state == 1) nEeFail = 0; else nDpFail = 0; }
				else mateStreaks[si]++;
			}
		}
	}
	co_return EXHAUSTED;
}

Task<PairOut> Engine::pairSteps(int idx1, const uint8_t *c1, const uint8_t *q1, int l1, std::string n1, const uint8_t *c2, const uint8_t *q2, int l2, std::string n2) {
	const uint8_t *cs[2] = {c1, c2}, *qs[2] = {q1, q2}; const int ls[2] = {l1, l2}; const std::string ns[2] = {n1, n2};
	for(int k = 0; k < 2; k++) {
		Mate &c = m[k];
		c.codes = cs[k]; c.quals = qs[k]; c.name = ns[k]; c.idx = idx1 + k; c.rdlen = ls[k];
		c.minsc = ls[k] ? P.minScore(ls[k]) : 0; c.perfect = P.perfect(ls[k]); c.nceil = ls[k] ? P.nCeil(ls[k]) : 0;
		int nn = 0; for(int i = 0; i < ls[k]; i++) nn += cs[k][i] > 3;
		c.filt = !(ls[k] < 2 || nn > P.nCeil(ls[k]) || P.perfect(ls[k]) < P.minScore(ls[k]));
	}
	const bool both = m[0].filt && m[1].filt;
	const uint32_t s1 = genRandSeed(c1, q1, l1, n1, P.seed), s2 = genRandSeed(c2, q2, l2, n2, P.seed);
	rnd.init(both ? (s1 ^ s2) : s1);
	int interval[2];
	for(int k = 0; k < 2; k++) interval[k] = ls[k] ? P.seedInterval(ls[k], both) : 1;
	int64_t streak = P.streak; int nroundsAll = P.seedRounds;
	if(both) { streak = (streak + 1) / 2; nroundsAll = (nroundsAll + 1) / 2; }
	streakCur = streak;
	psink = PairedSink{P.khits, P.mhits, P.mmode};
	psink.doneDiscord = !P.discord; psink.doneUnp[0] = psink.doneUnp[1] = !P.mixed;
	const bool m1fw = P.pe.pol == 1 || P.pe.pol == 3, m2fw = P.pe.pol == 1 || P.pe.pol == 4;
	const bool nofw[2] = {m1fw ? P.nofw : P.norc, m2fw ? P.nofw : P.norc}, norc[2] = {m1fw ? P.norc : P.nofw, m2fw ? P.norc : P.nofw};
	bool done[2] = {!m[0].filt, !m[1].filt};
	int matemap[2] = {0, 1}; int64_t nelt[2] = {0, 0}; int mined[2][2] = {{0, 0}, {0, 0}};
	auto after = [&](int ret, int mate) {
		if(ret == FULFILLED) { if(psink.doneWithMate(mate == 0)) done[mate] = true; if(psink.doneWithMate(mate == 1)) done[mate ^ 1] = true; }
		else if(ret == PERFECT || ret == HARD_LIMIT) done[mate] = true;
	};
	// ---- exact end-to-end
	for(int mi = 0; mi < 2; mi++) {
		const int mate = matemap[mi]; Mate &c = m[mate];
		if(!c.filt || done[mate] || psink.doneWithMate(mate == 0)) continue;
		Req sw; sw.kind = RQ_EXACT_SWEEP; sw.read = c.idx; sw.nofw = nofw[mate]; sw.norc = norc[mate];
		co_await AwaitReq{slot, &sw};
		nelt[mate] = (int64_t)sw.nelt; mined[mate][0] = sw.mined[0]; mined[mate][1] = sw.mined[1];
		c.ee.clear();
		if(sw.tb[1] > sw.tb[0]) c.ee.push_back(EEHit{sw.tb[0], sw.tb[1], true, c.perfect});
		if(sw.tb[3] > sw.tb[2]) c.ee.push_back(EEHit{sw.tb[2], sw.tb[3], false, c.perfect});
	}
	if(nelt[0] > 0 && nelt[1] > 0 && nelt[0] > nelt[1]) { matemap[0] = 1; matemap[1] = 0; } else { matemap[0] = 0; matemap[1] = 1; }
	for(int mi = 0; mi < 2; mi++) {
		const int mate = matemap[mi]; Mate &c = m[mate];
		if(nelt[mate] == 0) { c.ee.clear(); continue; }
		if(psink.doneWithMate(mate == 0)) { c.ee.clear(); done[mate] = true; continue; }
		std::vector<EEHit> ee = c.ee;
		const int ret = co_await extendSeedsPaired(mate, nullptr, ee);
		c.ee.clear();
		after(ret, mate);
		if(!done[mate] && c.minsc == c.perfect) done[mate] = true;
	}
	// ---- 1-mismatch end-to-end
	for(int mi = 0; mi < 2; mi++) {
		const int mate = matemap[mi]; Mate &c = m[mate];
		if(!c.filt || done[mate]) { c.mm1.clear(); nelt[mate] = 0; continue; }
		nelt[mate] = 0;
		const bool yfw = mined[mate][0] <= 1 && !nofw[mate], yrc = mined[mate][1] <= 1 && !norc[mate];
		if(yfw || yrc) {
			Req mq; mq.kind = RQ_ONE_MM; mq.read = c.idx; mq.minsc = c.minsc; mq.nofw = !yfw; mq.norc = !yrc;
			co_await AwaitReq{slot, &mq};
			c.mm1.clear();
			for(const Req::MmHit &h : mq.mm) { c.mm1.push_back(EEHit{h.top, h.bot, h.fw, h.score, true, Edit{h.pos, h.chr, h.qchr, 3}}); nelt[mate] += (int64_t)(h.bot - h.top); }
		}
	}
	if(nelt[0] > 0 && nelt[1] > 0 && nelt[0] > nelt[1]) { matemap[0] = 1; matemap[1] = 0; } else { matemap[0] = 0; matemap[1] = 1; }
	for(int mi = 0; mi < 2; mi++) {
		const int mate = matemap[mi]; Mate &c = m[mate];
		if(nelt[mate] == 0) continue;
		if(psink.doneWithMate(mate == 0)) { done[mate] = true; continue; }
		const int ret = co_await extendSeedsPaired(mate, nullptr, {});
		c.mm1.clear();
		after(ret, mate);
		if(!done[mate] && c.minsc == c.perfect) done[mate] = true;
	}
	// ---- seed rounds
	const int nrounds[2] = {std::min(nroundsAll, interval[0]), std::min(nroundsAll, interval[1])};
	const int L = P.seedLen;
	for(int roundi = 0; roundi < P.seedRounds; roundi++) {
		m[0].hasSh = m[1].hasSh = false;
		for(int mi = 0; mi < 2; mi++) {
			const int mate = matemap[mi]; Mate &c = m[mate];
			if(done[mate] || psink.doneWithMate(mate == 0)) { done[mate] = true; continue; }
			if(roundi >= nrounds[mate] || interval[mate] <= roundi) continue;
			const int offset = (interval[mate] * roundi) / nrounds[mate];
			if(offset > 0 && std::min(L, c.rdlen) + offset > c.rdlen) continue;
			Req sq; sq.kind = RQ_SEED_SEARCH; sq.read = c.idx; sq.L = std::min(L, c.rdlen); sq.interval = interval[mate]; sq.offset = offset; sq.nofw = nofw[mate]; sq.norc = norc[mate];
			co_await AwaitReq{slot, &sq};
			fillSeedHits(c.sh, sq, interval[mate], offset, std::min(L, c.rdlen));
			if(c.sh.nonz == 0) { done[mate] = true; break; }
			c.hasSh = true;
		}
		double uniq[2] = {0.0, 0.0};
		for(int k = 0; k < 2; k++) if(m[k].hasSh) {
			for(int64_t x : m[k].sh.nfw) if(x > 0) uniq[k] += 1.0 / (double)(x * x);
			for(int64_t x : m[k].sh.nrc) if(x > 0) uniq[k] += 1.0 / (double)(x * x);
		}
		if(m[0].hasSh && m[1].hasSh && uniq[1] > uniq[0]) { matemap[0] = 1; matemap[1] = 0; } else { matemap[0] = 0; matemap[1] = 1; }
		for(int mi = 0; mi < 2; mi++) {
			const int mate = matemap[mi]; Mate &c = m[mate];
			if(done[mate] || psink.doneWithMate(mate == 0)) { done[mate] = true; continue; }
			if(!c.hasSh) continue;
			cur = &c;
			rankSeedHits(c.sh);
			after(co_await extendSeedsPaired(mate, &c.sh, {}), mate);
		}
		for(int k = 0; k < 2; k++) if(!done[k] && m[k].hasSh && m[k].sh.nelt / m[k].sh.nonz < 300) done[k] = true;
	}
	co_return finishPair();
}

PairOut Engine::finishPair() {
	PairOut po;
	const int64_t mn[2] = {m[0].rdlen ? P.minScore(m[0].rdlen) : 0, m[1].rdlen ? P.minScore(m[1].rdlen) : 0};
	// selectByScore over pairs / unpaired lists (aln_sink.cpp:1477-1628)
	auto select = [&](const std::vector<Aln> &r1, const std::vector<Aln> *r2, std::vector<std::pair<int64_t, int>> &buf) {
		buf.clear();
		for(size_t i = 0; i < r1.size(); i++) buf.push_back({r1[i].score + (r2 ? (*r2)[i].score : 0), (int)i});
		std::sort(buf.begin(), buf.end(), [](auto &a, auto &b) { return a > b; });
		shuffleEqualStreaks(buf, [](const std::pair<int64_t, int> &t) { return t.first; }, rnd);
	};
	std::vector<std::pair<int64_t, int>> buf;
	auto unchosenP = [&](const std::vector<Aln> &rsu, const Aln &chosen, bool &has, int64_t &best) {
		has = false; best = 0;
		for(const Aln &a : rsu) { if(a.tidx == chosen.tidx && a.refoff == chosen.refoff && a.fw == chosen.fw) continue; if(!has || a.score > best) { has = true; best = a.score; } }
	};
	if(psink.nconcord > 0) {
		select(psink.rs1, &psink.rs2, buf);
		const Aln &a1 = psink.rs1[buf[0].second], &a2 = psink.rs2[buf[0].second];
		const bool hasC = buf.size() > 1;
		const int mq = (int)mapq(a1.score + a2.score, hasC, hasC ? buf[1].first : 0, mn[0] + mn[1], m[0].perfect + m[1].perfect);
		for(int k = 0; k < 2; k++) {
			Result &r = po.m[k]; r.aligned = true; r.aln = k == 0 ? a1 : a2; r.mapq = mq;
			unchosenP(k == 0 ? psink.rs1u : psink.rs2u, r.aln, r.hasXs, r.xs);
		}
		po.pairType = 1;
		po.scoreSum = a1.score + a2.score;
		po.kind = pe_classify(P.pe, a1.refoff, (uint64_t)a1.refExtent(), a1.fw, a2.refoff, (uint64_t)a2.refExtent(), a2.fw);
		{   // fragment length (pe.cpp:89-92): the span of the two alignments, soft-trimmed ends included
			const int64_t s1 = a1.refoff - a1.trimLeft(), e1 = a1.refoff + a1.refExtent() + (a1.rdlen - a1.ext() - a1.trimLeft());
			const int64_t s2 = a2.refoff - a2.trimLeft(), e2 = a2.refoff + a2.refExtent() + (a2.rdlen - a2.ext() - a2.trimLeft());
			po.fraglen = std::max(e1, e2) - std::min(s1, s2);
		}
		{   // ReportingState::getReport (aln_sink.cpp:300-330): after a -k short circuit khits pairs, else min(found, khits)
			const int64_t num = psink.exitConcordK ? P.khits : std::min<int64_t>(psink.nconcord, P.khits);
			for(int64_t j = 1; j < num && j < (int64_t)buf.size(); j++) po.secPairs.push_back({psink.rs1[buf[j].second], psink.rs2[buf[j].second]});
		}
		return po;
	}
	if(!psink.doneDiscord && psink.nunp[0] == 1 && psink.nunp[1] == 1) {
		const Aln &a1 = psink.rs1u[0], &a2 = psink.rs2u[0];
		const int mq = (int)mapq(a1.score + a2.score, false, 0, mn[0] + mn[1], m[0].perfect + m[1].perfect);
		for(int k = 0; k < 2; k++) { Result &r = po.m[k]; r.aligned = true; r.aln = k == 0 ? a1 : a2; r.mapq = mq; }
		po.pairType = 2;
		return po;
	}
	int nal = 0;
	for(int k = 0; k < 2; k++) {
		const std::vector<Aln> &rsu = k == 0 ? psink.rs1u : psink.rs2u;
		if(rsu.empty() || !P.mixed) continue;
		select(rsu, nullptr, buf);
		Result &r = po.m[k]; r.aligned = true; r.aln = rsu[buf[0].second];
		r.hasXs = buf.size() > 1; r.xs = r.hasXs ? rsu[buf[1].second].score : 0;
		r.mapq = (int)mapq(r.aln.score, r.hasXs, r.xs, mn[k], m[k].perfect);
		{
			const int64_t num = psink.exitUnpK[k] ? P.khits : std::min<int64_t>((int64_t)rsu.size(), P.khits);
			for(int64_t j = 1; j < num && j < (int64_t)buf.size(); j++) r.secondary.push_back(rsu[buf[j].second]);
		}
		nal++;
	}
	po.pairType = nal == 2 ? 2 : (nal == 1 ? 3 : 0);
	return po;
}


// ---------------------------------------------------------------------------------------------- waves
struct Batch {                                        // a sub-batch of reads for one entry-point call
	std::vector<uint8_t> seq, qual; std::vector<uint64_t> off{0}; bt2g_reads rb{};
	void add(const bt2g_reads *all, int read) {
		const uint64_t a = all->off[read], b = all->off[read + 1];
		seq.insert(seq.end(), all->seq + a, all->seq + b);
		if(all->qual) qual.insert(qual.end(), all->qual + a, all->qual + b); else qual.insert(qual.end(), b - a, (uint8_t)'I');
		off.push_back(seq.size());
	}
	const bt2g_reads *get() { rb.n_reads = off.size() - 1; rb.seq = seq.data(); rb.qual = qual.data(); rb.off = off.data(); return &rb; }
};

static const char DNA5[] = "ACGTN";

// device op string -> the reference's Edit list (lib.py: ops_to_edits)
static void opsToEdits(const uint8_t *ops, int nops, const uint8_t *codes, int rdlen, bool fw, int row0, int trimEnd, std::vector<Edit> &out) {
	out.clear();
	auto rd = [&](int row) { const int c = fw ? codes[row] : (codes[rdlen - 1 - row] > 3 ? 4 : 3 - codes[rdlen - 1 - row]); return (int)DNA5[c > 4 ? 4 : c]; };
	int row = row0;
	for(int k = nops - 1; k >= 0; k--) {
		const int typ = ops[k] & 3, refc = (ops[k] >> 2) & 7;
		if(typ == BT2G_OP_MATCH) row++;
		else if(typ == BT2G_OP_MM) { out.push_back({row - row0, (int)DNA5[refc > 4 ? 4 : refc], rd(row), 3}); row++; }
		else if(typ == BT2G_OP_REFGAP) { out.push_back({row - row0, (int)'-', rd(row), 2}); row++; }
		else out.push_back({row - row0, (int)DNA5[refc > 4 ? 4 : refc], (int)'-', 1});
	}
	if(!fw) {
		const int ext = rdlen - row0 - trimEnd;
		std::reverse(out.begin(), out.end());
		for(Edit &e : out) e.pos = ext - e.pos - (e.type == 1 ? 0 : 1);
	}
}

// alignment -> device op string (policy_engine.py: aln_to_ops); returns nops
static int alnToOps(const Aln &a, const uint8_t *codes, uint8_t *ops, uint32_t maxOps) {
	auto code = [](int ch) { return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 4; };
	const std::vector<Edit> ed = a.leftToRight();
	std::vector<uint8_t> fwd;
	size_t k = 0;
	const int row0 = a.trimLeft(), ext = a.ext(), rdlen = a.rdlen;
	for(int rel = 0; rel < ext; rel++) {
		while(k < ed.size() && ed[k].pos == rel && ed[k].type == 1) { fwd.push_back((uint8_t)(BT2G_OP_READGAP | (code(ed[k].chr) << 2))); k++; }
		if(k < ed.size() && ed[k].pos == rel) { fwd.push_back(ed[k].type == 2 ? (uint8_t)BT2G_OP_REFGAP : (uint8_t)(BT2G_OP_MM | (code(ed[k].chr) << 2))); k++; }
		else {
			const int row = row0 + rel;
			const int c = a.fw ? codes[row] : (codes[rdlen - 1 - row] > 3 ? 4 : 3 - codes[rdlen - 1 - row]);
			fwd.push_back((uint8_t)(BT2G_OP_MATCH | (c << 2)));
		}
	}
	const int n = (int)fwd.size();
	for(int i = 0; i < n && (uint32_t)i < maxOps; i++) ops[i] = fwd[n - 1 - i];
	return n;
}

struct Scheduler {
	const bt2g_policy_backend &be; const Params &P; const bt2g_reads *reads;
	uint64_t nWaves = 0, nCalls = 0, nRequests = 0;
	int rc = 0;
	Scheduler(const bt2g_policy_backend &b, const Params &p, const bt2g_reads *r) : be(b), P(p), reads(r) {}
	int rdlen(int read) const { return (int)(reads->off[read + 1] - reads->off[read]); }
	const uint8_t *codes(int read) const { return reads->seq + reads->off[read]; }

	void answer(int kind, std::vector<Req *> &rq) {
		nRequests += rq.size();
		switch(kind) {
		case RQ_EXACT_SWEEP: {
			for(int fl = 0; fl < 4; fl++) {
				std::vector<Req *> g; for(Req *r : rq) if((r->nofw ? 1 : 0) + (r->norc ? 2 : 0) == fl) g.push_back(r);
				if(g.empty()) continue;
				Batch b; for(Req *r : g) b.add(reads, r->read);
				std::vector<uint8_t> mine(2 * g.size()); std::vector<uint64_t> ee(4 * g.size());
				rc |= be.exact_sweep(be.ctx, b.get(), fl & 1, (fl >> 1) & 1, mine.data(), ee.data()); nCalls++;
				for(size_t k = 0; k < g.size(); k++) {
					Req *r = g[k];
					for(int j = 0; j < 4; j++) r->tb[j] = ee[4 * k + j];
					r->mined[0] = mine[2 * k]; r->mined[1] = mine[2 * k + 1];
					r->nelt = (r->tb[1] > r->tb[0] ? r->tb[1] - r->tb[0] : 0) + (r->tb[3] > r->tb[2] ? r->tb[3] - r->tb[2] : 0);
				}
			}
			break; }
		case RQ_ONE_MM: {
			Batch b; for(Req *r : rq) b.add(reads, r->read);
			const int MH = 64;
			std::vector<int32_t> minsc(rq.size()), counts(4 * rq.size()); std::vector<uint8_t> mask(rq.size());
			std::vector<bt2g_mm_hit> hits(rq.size() * 4 * (size_t)MH);
			for(size_t k = 0; k < rq.size(); k++) { minsc[k] = (int32_t)rq[k]->minsc; mask[k] = (uint8_t)((rq[k]->nofw ? 0 : 1) | (rq[k]->norc ? 0 : 2)); }
			rc |= be.one_mm(be.ctx, b.get(), minsc.data(), mask.data(), MH, hits.data(), counts.data()); nCalls++;
			for(size_t k = 0; k < rq.size(); k++) {
				rq[k]->mm.clear();
				for(int task = 0; task < 4; task++) for(int j = 0; j < counts[4 * k + task]; j++) {
					const bt2g_mm_hit &h = hits[(k * 4 + task) * (size_t)MH + j];
					rq[k]->mm.push_back({h.top, h.bot, h.pos, (int)DNA5[h.chr > 4 ? 4 : h.chr], (int)DNA5[h.qchr > 4 ? 4 : h.qchr], h.score, task < 2});
				}
			}
			break; }
		case RQ_SEED_SEARCH: {
			std::vector<char> used(rq.size(), 0);
			for(size_t s0 = 0; s0 < rq.size(); s0++) {
				if(used[s0]) continue;
				std::vector<Req *> g;
				for(size_t k = s0; k < rq.size(); k++) if(!used[k] && rq[k]->L == rq[s0]->L && rq[k]->nofw == rq[s0]->nofw && rq[k]->norc == rq[s0]->norc) { used[k] = 1; g.push_back(rq[k]); }
				Batch b; int nsMax = 1;
				std::vector<int32_t> iv(g.size()), of(g.size());
				for(size_t k = 0; k < g.size(); k++) {
					b.add(reads, g[k]->read); iv[k] = g[k]->interval; of[k] = g[k]->offset;
					const int len = rdlen(g[k]->read);
					int n = 1; if(len - g[k]->offset > g[k]->L) n += (len - g[k]->offset - g[k]->L) / g[k]->interval;
					nsMax = std::max(nsMax, n);
				}
				nsMax += 2;
				bt2g_seed_plan plan{g[0]->L, nsMax, g[0]->nofw, g[0]->norc, iv.data(), of.data()};
				std::vector<uint64_t> out(g.size() * 2 * (size_t)nsMax * 4); std::vector<int32_t> ns(g.size());
				rc |= be.seed_search(be.ctx, b.get(), &plan, out.data(), ns.data()); nCalls++;
				for(size_t k = 0; k < g.size(); k++) {
					Req *r = g[k]; r->nseeds = ns[k]; r->ranges.assign((size_t)2 * ns[k] * 4, 0);
					for(int st = 0; st < 2; st++) for(int i = 0; i < ns[k]; i++) for(int j = 0; j < 4; j++)
						r->ranges[((size_t)st * ns[k] + i) * 4 + j] = out[((k * 2 + st) * (size_t)nsMax + i) * 4 + j];
				}
			}
			break; }
		case RQ_EXTEND: {
			std::vector<char> used(rq.size(), 0);
			for(size_t s0 = 0; s0 < rq.size(); s0++) {
				if(used[s0]) continue;
				std::vector<Req *> g;
				for(size_t k = s0; k < rq.size(); k++) if(!used[k] && rq[k]->seedlen == rq[s0]->seedlen) { used[k] = 1; g.push_back(rq[k]); }
				Batch b; std::vector<int32_t> iv(g.size()), of(g.size()); std::vector<uint64_t> ranges(g.size() * 2 * 4, 0);
				for(size_t k = 0; k < g.size(); k++) {
					b.add(reads, g[k]->read); iv[k] = std::max(1, rdlen(g[k]->read)); of[k] = g[k]->rdoff;
					for(int j = 0; j < 4; j++) ranges[(k * 2 + (g[k]->fw ? 0 : 1)) * 4 + j] = g[k]->rng4[j];
				}
				bt2g_seed_plan plan{g[0]->seedlen, 1, 0, 0, iv.data(), of.data()};
				std::vector<uint8_t> out(g.size() * 2 * 2, 0);
				rc |= be.extend_exact(be.ctx, b.get(), &plan, ranges.data(), out.data()); nCalls++;
				for(size_t k = 0; k < g.size(); k++) { const size_t o = (k * 2 + (g[k]->fw ? 0 : 1)) * 2; g[k]->nlex = out[o]; g[k]->nrex = out[o + 1]; }
			}
			break; }
		case RQ_RESOLVE: {
			for(int rej = 0; rej < 2; rej++) {
				std::vector<Req *> g; for(Req *r : rq) if((r->reject ? 1 : 0) == rej) g.push_back(r);
				if(g.empty()) continue;
				const size_t n = g.size();
				std::vector<uint64_t> rows(n), joined(n), tidx(n), toff(n), tlen(n); std::vector<uint32_t> hl(n); std::vector<uint8_t> fl(n);
				for(size_t k = 0; k < n; k++) { rows[k] = g[k]->row; hl[k] = (uint32_t)g[k]->qlen; }
				rc |= be.resolve(be.ctx, rows.data(), hl.data(), n, rej, joined.data(), tidx.data(), toff.data(), tlen.data(), fl.data()); nCalls++;
				for(size_t k = 0; k < n; k++) {
					Req *r = g[k]; r->joined = joined[k]; r->ok = !((fl[k] >> 1) & 1); r->straddled = fl[k] & 1;
					r->rtidx = (int64_t)tidx[k]; r->rtoff = (int64_t)toff[k]; r->rtlen = (int64_t)tlen[k];
				}
			}
			break; }
		case RQ_UNGAPPED: {
			Batch b; std::vector<bt2g_ungapped_problem> probs(rq.size()); int maxLen = 1;
			for(size_t k = 0; k < rq.size(); k++) {
				b.add(reads, rq[k]->read); maxLen = std::max(maxLen, rdlen(rq[k]->read));
				bt2g_ungapped_problem &p = probs[k]; p = bt2g_ungapped_problem{};
				p.read_idx = (uint32_t)k; p.fw = rq[k]->fw; p.tidx = (uint64_t)rq[k]->tidx; p.refoff = rq[k]->refoff; p.reflen = (uint64_t)rq[k]->tlen;
				p.minsc = (int32_t)rq[k]->minsc; p.ohang = 0;
			}
			std::vector<bt2g_ungapped_result> out(rq.size()); std::vector<uint8_t> mask(rq.size() * (size_t)maxLen, 0);
			rc |= be.ungapped(be.ctx, b.get(), probs.data(), rq.size(), out.data(), mask.data(), (uint32_t)maxLen); nCalls++;
			std::vector<size_t> hit; for(size_t k = 0; k < rq.size(); k++) { rq[k]->ugStatus = out[k].status; if(out[k].status == 1) hit.push_back(k); }
			if(!hit.empty()) {
				std::vector<uint64_t> ti(hit.size()); std::vector<int64_t> of(hit.size()); std::vector<int32_t> cnt(hit.size());
				for(size_t j = 0; j < hit.size(); j++) { ti[j] = (uint64_t)rq[hit[j]]->tidx; of[j] = rq[hit[j]]->refoff; cnt[j] = rdlen(rq[hit[j]]->read); }
				std::vector<uint8_t> st(hit.size() * (size_t)maxLen, 4);
				rc |= be.get_stretch(be.ctx, ti.data(), of.data(), cnt.data(), hit.size(), maxLen, st.data()); nCalls++;
				for(size_t j = 0; j < hit.size(); j++) {
					Req *r = rq[hit[j]]; const size_t k = hit[j];
					const int len = rdlen(r->read); const uint8_t *cd = codes(r->read); const bool fw = r->fw;
					const int rowi = out[k].rowi, rowf = out[k].rowf, ext = rowf - rowi + 1;
					Aln &a = r->ugAln; a = Aln{};
					for(int i = rowi; i <= rowf; i++) if(mask[k * (size_t)maxLen + i]) {
						const int c = fw ? cd[i] : (cd[len - 1 - i] > 3 ? 4 : 3 - cd[len - 1 - i]);
						const int refc = st[j * (size_t)maxLen + i];
						const int rel = i - rowi;
						a.edits.push_back({fw ? rel : ext - 1 - rel, (int)DNA5[refc > 4 ? 4 : refc], (int)DNA5[c > 4 ? 4 : c], 3});
					}
					if(!fw) std::reverse(a.edits.begin(), a.edits.end());
					const int tl = rowi, tr = len - 1 - rowf;
					a.tidx = r->tidx; a.refoff = r->refoff + rowi; a.fw = fw; a.score = out[k].score; a.rdlen = len; a.ns = out[k].ns; a.refns = out[k].refns;
					a.trim5 = fw ? tl : tr; a.trim3 = fw ? tr : tl;
				}
			}
			break; }
		case RQ_DP: {
			const size_t CH = 4096;
			for(size_t c0 = 0; c0 < rq.size(); c0 += CH) {
				const size_t n = std::min(CH, rq.size() - c0);
				Batch b; std::vector<bt2g_dp_problem> probs(n); int maxLen = 1;
				for(size_t k = 0; k < n; k++) { Req *r = rq[c0 + k]; b.add(reads, r->read); probs[k] = r->prob; probs[k].read_idx = (uint32_t)k; maxLen = std::max(maxLen, rdlen(r->read)); }
				const int32_t maxCands = P.local ? 16384 : 1024, maxAlns = 16, maxOps = maxLen + 80;
				std::vector<bt2g_dp_summary> summ(n); std::vector<bt2g_dp_cand> cands(n * (size_t)maxCands); std::vector<bt2g_dp_aln> alns(n * (size_t)maxAlns);
				std::vector<uint8_t> ops(n * (size_t)maxAlns * maxOps);
				rc |= be.dp_extend(be.ctx, b.get(), probs.data(), n, maxCands, maxAlns, maxOps, summ.data(), cands.data(), alns.data(), ops.data()); nCalls++;
				for(size_t k = 0; k < n; k++) {
					Req *r = rq[c0 + k];
					if(summ[k].flags) {                                   // rare: more alignments / candidates / edit ops than the batch buffers hold
						// alone, with buffers grown until everything fits (cheap gaps and a high match bonus give alignments with far more
						// ops than rows; repeats give hundreds of candidates worth a backtrace)
						Batch b1; b1.add(reads, r->read); bt2g_dp_problem p1 = r->prob; p1.read_idx = 0;
						int32_t mc = 65536, ma = 128, mo = rdlen(r->read) + 80;
						std::vector<bt2g_dp_summary> s1(1); std::vector<bt2g_dp_cand> c1; std::vector<bt2g_dp_aln> a1; std::vector<uint8_t> o1;
						for(int attempt = 0; attempt < 6; attempt++) {
							c1.assign((size_t)mc, bt2g_dp_cand{}); a1.assign((size_t)ma, bt2g_dp_aln{}); o1.assign((size_t)ma * mo, 0);
							rc |= be.dp_extend(be.ctx, b1.get(), &p1, 1, mc, ma, mo, s1.data(), c1.data(), a1.data(), o1.data()); nCalls++;
							const uint32_t fl = s1[0].flags;
							if(!fl) break;
							if(fl & BT2G_DP_FLAG_CAND_OVERFLOW) mc = std::max(2 * mc, s1[0].ncand + 1);
							if(fl & BT2G_DP_FLAG_ALN_OVERFLOW) ma = std::max(2 * ma, s1[0].naln + 1);
							if(fl & BT2G_DP_FLAG_OPS_OVERFLOW) {
								int need = 2 * mo;
								for(int q = 0; q < s1[0].naln && q < (int)a1.size(); q++) need = std::max(need, a1[q].nops + 16);
								mo = need;
							}
						}
						if(s1[0].flags) { rc |= 1; if(getenv("BT2G_PE_DEBUG")) fprintf(stderr, "dp retry overflow: flags=%d ncand=%d naln=%d len=%d\n", s1[0].flags, s1[0].ncand, s1[0].naln, rdlen(r->read)); }
						fillDp(r, s1[0], c1.data(), a1.data(), o1.data(), mo);
					} else fillDp(r, summ[k], cands.data() + k * (size_t)maxCands, alns.data() + k * (size_t)maxAlns, ops.data() + k * (size_t)maxAlns * maxOps, maxOps);
				}
			}
			break; }
		}
	}
	void fillDp(Req *r, const bt2g_dp_summary &s, const bt2g_dp_cand *cands, const bt2g_dp_aln *alns, const uint8_t *ops, int maxOps) {
		DpOut &d = r->dp; d = DpOut{};
		d.found = s.found != 0; d.best = s.best;
		if(!d.found) return;
		const int len = rdlen(r->read); const bool fw = r->prob.fw != 0;
		std::vector<int> byCand((size_t)s.ncand, -1);
		for(int k = 0; k < s.naln; k++) {
			const bt2g_dp_aln &al = alns[k];
			Aln a; a.tidx = (int64_t)r->prob.tidx; a.refoff = r->prob.refl + al.col0; a.fw = fw; a.score = al.score; a.rdlen = len; a.ns = al.ns; a.refns = al.refns;
			a.trim5 = fw ? al.trim_beg : al.trim_end; a.trim3 = fw ? al.trim_end : al.trim_beg;
			opsToEdits(ops + (size_t)k * maxOps, al.nops, codes(r->read), len, fw, al.row0, al.trim_end, a.edits);
			d.alns.push_back(std::move(a));
			if(al.cand_idx >= 0 && al.cand_idx < s.ncand) byCand[al.cand_idx] = k;
		}
		for(int ci = 0; ci < s.ncand; ci++) {
			if(cands[ci].fate == BT2G_CAND_SUCCEEDED) d.attempts.push_back({cands[ci].score, byCand[ci]});
			else if(cands[ci].fate == BT2G_CAND_FAILED) d.attempts.push_back({cands[ci].score, -1});
		}
	}
};

} // namespace

// ---------------------------------------------------------------------------------------------- C ABI
extern "C" void bt2g_policy_backend_gpu(bt2g_ctx *ctx, bt2g_policy_backend *be) {
	be->ctx = ctx;
	be->exact_sweep = (decltype(be->exact_sweep))bt2g_exact_sweep;
	be->seed_search = (decltype(be->seed_search))bt2g_seed_search;
	be->one_mm = (decltype(be->one_mm))bt2g_one_mm;
	be->extend_exact = (decltype(be->extend_exact))bt2g_extend_exact;
	be->resolve = (decltype(be->resolve))bt2g_resolve;
	be->get_stretch = (decltype(be->get_stretch))bt2g_get_stretch;
	be->ungapped = (decltype(be->ungapped))bt2g_ungapped;
	be->dp_extend = (decltype(be->dp_extend))bt2g_dp_extend;
	bt2g_index_info inf{};
	be->off_size = bt2g_index_info_get(ctx, &inf) == 0 ? inf.off_size : 4;
	be->reserved = 0;
}

static void fillResult(const Result &r, const uint8_t *codes, bt2g_read_result &out, uint8_t *ops, uint32_t maxOps) {
	out = bt2g_read_result{};
	out.score2 = INT32_MIN;
	if(!r.aligned) return;
	const Aln &a = r.aln;
	int nops = alnToOps(a, codes, ops, maxOps);
	// (rows are maxOps wide; nops keeps the true count, as the device engine's x_fill_result does: a reader clamps, and
	// bt2g_sam_format reports such records -- their op string lost its tail, the caller's max_ops was too small)
	out.found = (a.edits.empty() && a.ext() == a.rdlen) ? 2 : 1;
	out.score = (int32_t)a.score; if(r.hasXs) out.score2 = (int32_t)r.xs;
	out.fw = a.fw; out.tidx = (uint64_t)a.tidx; out.refoff = a.refoff; out.nops = nops;
	out.trim_left = a.trimLeft(); out.trim_right = a.rdlen - a.ext() - a.trimLeft();
	out.mapq = r.mapq; out.pad = a.refns;
}

static int policyAlign(const bt2g_policy_backend *be, const bt2g_policy_params *pp, const bt2g_reads *reads, const char *const *names,
                       bt2g_read_result *res, uint8_t *ops, uint32_t maxOps, bt2g_pair_result *pairs, uint64_t *stats,
                       uint32_t maxPerRead, uint32_t *nReported) {
	if(!be || !pp || !reads || !res || !ops || (pp->paired && (!pairs || (reads->n_reads & 1)))) return -1;
	bool truncated = false;
	Params P{};
	P.local = pp->local; P.paired = pp->paired; P.all = pp->all_hits; P.mmode = pp->mmode; P.nofw = pp->nofw; P.norc = pp->norc;
	P.discord = pp->discord; P.mixed = pp->mixed;
	P.seedLen = pp->seed_len; P.seedRounds = pp->seed_rounds; P.streak = pp->dp_fail_streak;
	P.maxIters = 400; P.maxUg = 300; P.maxDp = 300; P.maxMateStreak = 10;
	P.khits = P.all ? BIG : (pp->khits > 0 ? pp->khits : 1);
	P.mhits = P.mmode ? (pp->mhits > 0 ? pp->mhits : 50) : BIG;
	if(P.all) { P.maxIters = P.maxUg = P.maxDp = P.streak = P.maxMateStreak = (int)(BIG >> 33); }
	else if(P.khits > 1) { const int k1 = (int)(P.khits - 1); P.streak += k1 * 10; P.maxMateStreak += k1 * 10; P.maxIters += k1 * 20; P.maxUg += k1 * 20; P.maxDp += k1 * 20; }
	P.ival = Func{pp->ival_type, pp->ival_const, pp->ival_coeff};
	P.smin = Func{pp->smin_type, pp->smin_const, pp->smin_coeff};
	P.nceil = Func{2, pp->nceil_const, pp->nceil_coeff};
	P.seed = pp->seed;
	P.matchBonus = pp->match_bonus; P.mmpMax = pp->mmp_max; P.mmpMin = pp->mmp_min; P.nPen = pp->n_pen;
	P.rdgConst = pp->rdgap_const; P.rdgLin = pp->rdgap_linear; P.rfgConst = pp->rfgap_const; P.rfgLin = pp->rfgap_linear;
	P.pe = pp->pe; P.offSize = be->off_size;
	Scheduler S(*be, P, reads);
	const size_t units = P.paired ? reads->n_reads / 2 : reads->n_reads;
	const size_t maxIn = pp->max_inflight > 0 ? (size_t)pp->max_inflight : 65536;
	struct Unit { std::unique_ptr<Engine> eng; std::unique_ptr<Task<Result>> tr; std::unique_ptr<Task<PairOut>> tp; Pending pend; size_t id; };
	std::vector<std::unique_ptr<Unit>> active;
	size_t next = 0;
	auto nameOf = [&](size_t i) { return names && names[i] ? std::string(names[i]) : std::string(); };
	auto finish = [&](Unit &u) {
		if(P.paired) {
			PairOut po = std::move(u.tp->h.promise().value);
			// entry e of pair i: rows 2 * (i * maxPerRead + e) + {0, 1} and pairs[i * maxPerRead + e].  Entry 0 carries the primaries of both
			// mates; further entries (-k / -a, AlnSink::reportHits, aln_sink.h:640-735): the other concordant pairs, or every record of mate 1
			// and then of mate 2, each beside the opposite mate's primary as its mate.  found bit 8 = secondary (FLAG 256, MAPQ 255),
			// bit 9 = present only as its mate's mate: bt2g_sam_format does not print it
			const size_t e0 = u.id * (size_t)maxPerRead;
			size_t n = 0;
			auto entry = [&](const Result &r1, const Result &r2, int mark1, int mark2) {
				if(n >= maxPerRead) { truncated = truncated || maxPerRead > 1; return; }
				const size_t e = e0 + n;
				pairs[e] = bt2g_pair_result{}; pairs[e].pair_type = po.pairType; pairs[e].kind = po.kind;
				pairs[e].score_sum = (int32_t)po.scoreSum; pairs[e].fraglen = po.fraglen;
				const Result *rr[2] = {&r1, &r2}; const int mk[2] = {mark1, mark2};
				for(int k = 0; k < 2; k++) {
					fillResult(*rr[k], S.codes((int)(2 * u.id + k)), res[2 * e + k], ops + (2 * e + k) * (size_t)maxOps, maxOps);
					res[2 * e + k].found |= mk[k];
				}
				n++;
			};
			auto secOf = [&](const Result &prim, const Aln &a) { Result s2; s2.aligned = true; s2.aln = a; s2.hasXs = prim.hasXs; s2.xs = prim.xs; s2.mapq = 255; return s2; };
			const Result &m0 = po.m[0], &m1 = po.m[1];
			if(po.pairType == 1) {
				entry(m0, m1, 0, 0);
				for(auto &pr2 : po.secPairs) entry(secOf(m0, pr2.first), secOf(m1, pr2.second), 0x100, 0x100);
			} else if(m0.secondary.empty() && m1.secondary.empty()) {
				entry(m0, m1, 0, 0);
			} else {
				// unpaired alignments of a pair: ALL of mate 1's records, then all of mate 2's, an unaligned mate's record last
				// (AlnSinkWrap::finishRead reports rs1u, then rs2u, then the unaligned mates: aln_sink.cpp:930-1010)
				if(m0.aligned) { entry(m0, m1, 0, 0x200); for(const Aln &a : m0.secondary) entry(secOf(m0, a), m1, 0x100, 0x200); }
				if(m1.aligned) { entry(m0, m1, 0x200, 0); for(const Aln &a : m1.secondary) entry(m0, secOf(m1, a), 0x200, 0x100); }
				if(!m0.aligned) entry(m0, m1, 0, 0x200);
				if(!m1.aligned) entry(m0, m1, 0x200, 0);
			}
			if(nReported) nReported[u.id] = (uint32_t)n;
		} else {
			// one row per reported alignment: the primary, then the secondaries (found bit 8 -> FLAG 256, MAPQ 255, the read's XS:i)
			Result &r = u.tr->h.promise().value;
			const size_t row0 = u.id * (size_t)maxPerRead;
			fillResult(r, S.codes((int)u.id), res[row0], ops + row0 * (size_t)maxOps, maxOps);
			size_t n = r.aligned ? 1 : 0;
			for(size_t j = 0; r.aligned && j < r.secondary.size(); j++) {
				if(n >= maxPerRead) { truncated = truncated || maxPerRead > 1; break; }
				Result s2; s2.aligned = true; s2.aln = r.secondary[j]; s2.hasXs = r.hasXs; s2.xs = r.xs; s2.mapq = 255;
				fillResult(s2, S.codes((int)u.id), res[row0 + n], ops + (row0 + n) * (size_t)maxOps, maxOps);
				res[row0 + n].found |= 0x100;
				n++;
			}
			if(nReported) nReported[u.id] = (uint32_t)n;
		}
	};
	auto isDone = [&](Unit &u) { return P.paired ? u.tp->h.done() : u.tr->h.done(); };
	while(next < units || !active.empty()) {
		while(next < units && active.size() < maxIn) {
			auto u = std::make_unique<Unit>();
			u->id = next; u->eng = std::make_unique<Engine>(P, &u->pend);
			if(P.paired) {
				const int i1 = (int)(2 * next), i2 = i1 + 1;
				u->tp = std::make_unique<Task<PairOut>>(u->eng->pairSteps(i1, S.codes(i1), reads->qual + reads->off[i1], S.rdlen(i1), nameOf(i1),
				                                                         S.codes(i2), reads->qual + reads->off[i2], S.rdlen(i2), nameOf(i2)));
				u->tp->h.resume();
			} else {
				const int i = (int)next;
				u->tr = std::make_unique<Task<Result>>(u->eng->readSteps(i, S.codes(i), reads->qual + reads->off[i], S.rdlen(i), nameOf(i)));
				u->tr->h.resume();
			}
			next++;
			if(isDone(*u)) finish(*u); else active.push_back(std::move(u));
		}
		if(active.empty()) continue;
		S.nWaves++;
		std::vector<Req *> groups[RQ_NKINDS];
		for(auto &u : active) groups[u->pend.req->kind].push_back(u->pend.req);
		for(int k = 0; k < RQ_NKINDS; k++) if(!groups[k].empty()) S.answer(k, groups[k]);
		if(S.rc) return -2;
		// resume every unit (they are independent: own engine, own RNG), on the host threads the caller allows
		{
			const size_t T = std::min<size_t>(std::max(1, pp->host_threads), std::max<size_t>(1, active.size() / 64));
			auto run = [&](size_t t) { for(size_t i = t; i < active.size(); i += T) active[i]->pend.leaf.resume(); };
			if(T == 1) run(0);
			else { std::vector<std::thread> th; for(size_t t = 0; t < T; t++) th.emplace_back(run, t); for(auto &x : th) x.join(); }
		}
		std::vector<std::unique_ptr<Unit>> still;
		for(auto &u : active) { if(isDone(*u)) finish(*u); else still.push_back(std::move(u)); }
		active.swap(still);
	}
	if(stats) { stats[0] = S.nWaves; stats[1] = S.nCalls; stats[2] = S.nRequests; }
	return truncated ? 1 : 0;
}

extern "C" int bt2g_policy_align(const bt2g_policy_backend *be, const bt2g_policy_params *pp, const bt2g_reads *reads, const char *const *names,
                                 bt2g_read_result *res, uint8_t *ops, uint32_t maxOps, bt2g_pair_result *pairs, uint64_t *stats) {
	return policyAlign(be, pp, reads, names, res, ops, maxOps, pairs, stats, 1, nullptr);
}

// paired -k N / -a: up to maxPerPair entries per pair (see `finish` above); res / ops hold 2 * n_pairs * maxPerPair rows, pairs
// n_pairs * maxPerPair records, nEntries[n_pairs] the entries used.  Returns 1 when some pair had more entries than maxPerPair.
extern "C" int bt2g_policy_align_pairs_k(const bt2g_policy_backend *be, const bt2g_policy_params *pp, const bt2g_reads *reads, const char *const *names,
                                         uint32_t maxPerPair, bt2g_read_result *res, uint8_t *ops, uint32_t maxOps, bt2g_pair_result *pairs,
                                         uint32_t *nEntries, uint64_t *stats) {
	if(!pp || !pp->paired || maxPerPair == 0 || !nEntries) return -1;
	return policyAlign(be, pp, reads, names, res, ops, maxOps, pairs, stats, maxPerPair, nEntries);
}

extern "C" int bt2g_policy_align_k(const bt2g_policy_backend *be, const bt2g_policy_params *pp, const bt2g_reads *reads, const char *const *names,
                                   uint32_t maxPerRead, bt2g_read_result *res, uint8_t *ops, uint32_t maxOps, uint32_t *nReported, uint64_t *stats) {
	if(!pp || pp->paired || maxPerRead == 0 || !nReported) return -1;
	return policyAlign(be, pp, reads, names, res, ops, maxOps, nullptr, stats, maxPerRead, nReported);
}
