// bt2g_stream_run (csrc/stream_host.cpp) under sanitizers: tests/test_host_sanitizers.py builds this file with csrc/stream_host.cpp and
// csrc/sam_host.cpp, once with -fsanitize=address,undefined and once with -fsanitize=thread.  The engines are synthetic (every read
// unaligned, after a random delay, so that blocks finish out of order); the run's SAM text must equal the text of the same blocks
// formatted one after the other by bt2g_fastq_parse[_pairs]_mt + bt2g_sam_format on the calling thread, for random block cuts, engine
// counts and thread counts; a failing engine / reader / writer must end the run with its code and no hang.  argv: seed, iterations.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>
#include "bt2g.h"

struct Eng { int id; int failAt; std::atomic<int> *calls; };

static int fakeAlign(void *engine, const bt2g_reads *reads, const char *, uint32_t, bt2g_read_result *res, uint8_t *ops, uint32_t maxOps,
                     bt2g_pair_result *pairs, uint64_t *) {
	Eng *e = (Eng *)engine;
	int c = e->calls->fetch_add(1);
	if(e->failAt >= 0 && c == e->failAt) return -77;
	std::this_thread::sleep_for(std::chrono::microseconds((reads->n_reads * 7919u + (unsigned)e->id * 104729u) % 3000u));
	memset(res, 0, sizeof(bt2g_read_result) * reads->n_reads);
	memset(ops, 0, (size_t)maxOps * reads->n_reads);
	if(pairs) memset(pairs, 0, sizeof(bt2g_pair_result) * (reads->n_reads / 2));
	return 0;
}

struct IO {
	const std::vector<std::pair<std::string, std::string>> *blocks;
	size_t next = 0;
	std::string out;
	int failRead = -1, failWrite = -1, writes = 0;
	void *bytes = nullptr;
};
static int nextBlock(void *u, const char **t1, uint64_t *l1, const char **t2, uint64_t *l2) {
	IO *io = (IO *)u;
	if((int)io->next == io->failRead) return -5;
	if(io->next >= io->blocks->size()) return 0;
	const auto &b = (*io->blocks)[io->next++];
	*t1 = b.first.data(); *l1 = b.first.size(); *t2 = b.second.data(); *l2 = b.second.size();
	return 1;
}
// the same blocks as two byte streams (the read callback), handed out in random-sized pieces
struct Bytes { std::string text[2]; size_t at[2] = {0, 0}; std::mt19937_64 *rng; IO *io; };
static int64_t readBytes(void *u, int mate, char *dst, uint64_t cap) {
	Bytes *b = (Bytes *)((IO *)u)->bytes;
	size_t left = b->text[mate].size() - b->at[mate];
	size_t n = std::min<size_t>(std::min<size_t>(left, cap), 1 + (*b->rng)() % 5000);
	memcpy(dst, b->text[mate].data() + b->at[mate], n);
	b->at[mate] += n;
	return (int64_t)n;
}
static int writeOut(void *u, const char *s, uint64_t n) {
	IO *io = (IO *)u;
	if(io->writes++ == io->failWrite) return 9;
	io->out.append(s, n);
	return 0;
}

int main(int argc, char **argv) {
	std::mt19937_64 rng(argc > 1 ? atoi(argv[1]) : 1);
	int iters = argc > 2 ? atoi(argv[2]) : 40;
	auto rnd = [&](uint64_t n) { return n ? rng() % n : 0; };
	long bad = 0, streamed = 0, solos = 0;
	const char *refNames[1] = {"chr1"};
	for(int it = 0; it < iters; it++) {
		const bool paired = rnd(2);
		const int nblocks = (int)rnd(9), E = 1 + (int)rnd(4);
		const uint64_t maxUnits = 1 + rnd(300);
		const uint32_t stride = 16 + (uint32_t)rnd(40), maxLen = 200, maxOps = 264;
		const bool withSolos = paired && rnd(3) == 0;             // some pairs have an empty mate 2: they go through the solo engine
		std::vector<std::pair<std::string, std::string>> blocks;
		uint64_t id = 0;
		for(int b = 0; b < nblocks; b++) {
			const uint64_t n = rnd(6) == 0 ? 0 : 1 + rnd(maxUnits);
			std::string t[2];
			for(uint64_t i = 0; i < n; i++, id++)
				for(int f = 0; f < (paired ? 2 : 1); f++) {
					int L = 1 + (int)rnd(maxLen);
					if(withSolos && f == 1 && rnd(7) == 0) L = 0;
					t[f] += "@read" + std::to_string(id) + "/" + std::to_string(f + 1) + "\n";
					for(int k = 0; k < L; k++) t[f] += "ACGTN"[rnd(5)];
					t[f] += "\n+\n";
					for(int k = 0; k < L; k++) t[f] += (char)(34 + rnd(40));
					t[f] += "\n";
				}
			blocks.emplace_back(t[0], t[1]);
		}
		// the expected text: block after block on this thread
		bt2g_sam_opts opt;
		memset(&opt, 0, sizeof(opt));
		opt.ref_names = refNames; opt.n_refs = 1;
		std::string want;
		uint64_t wantReads = 0, soloCount = 0;
		for(auto &b : blocks) {
			const uint64_t cap = b.first.size() + b.second.size() + 1, mr = maxUnits * (paired ? 2 : 1);
			std::vector<uint8_t> seq(cap), qual(cap);
			std::vector<uint64_t> off(mr + 1);
			std::vector<char> names(mr * stride);
			uint64_t n = 0, c1 = 0, c2 = 0;
			int rc = paired ? bt2g_fastq_parse_pairs_mt(b.first.data(), b.first.size(), b.second.data(), b.second.size(), maxUnits, cap, seq.data(), qual.data(),
			                                            off.data(), names.data(), stride, &n, &c1, &c2, 1)
			                : bt2g_fastq_parse_mt(b.first.data(), b.first.size(), mr, cap, seq.data(), qual.data(), off.data(), names.data(), stride, &n, &c1, 1);
			if(rc) { printf("iteration %d: reference parse failed %d\n", it, rc); bad++; break; }
			if(paired) n *= 2;
			if(!n) continue;
			std::vector<bt2g_read_result> res(n);
			memset(res.data(), 0, sizeof(bt2g_read_result) * n);
			std::vector<uint8_t> ops((size_t)n * maxOps, 0);
			std::vector<bt2g_pair_result> pr(n / 2 + 1);
			memset(pr.data(), 0, sizeof(bt2g_pair_result) * pr.size());
			std::vector<const char *> np(n);
			for(uint64_t i = 0; i < n; i++) np[i] = names.data() + i * stride;
			bt2g_reads rd{n, seq.data(), qual.data(), off.data()};
			bt2g_sam_opts o = opt;
			o.read_names = np.data();
			std::vector<char> out(off[n] * 2 + n * 400 + 4096);
			uint64_t need = 0;
			if(withSolos) {
				// pair by pair: an ordinary pair as two records, a pair with an empty mate 2 as ONE unpaired record of its mate 1
				for(uint64_t i = 0; i < n; i += 2) {
					const bool solo = off[i + 2] == off[i + 1];
					bt2g_reads one{solo ? 1ull : 2ull, seq.data(), qual.data(), off.data() + i};
					o.read_names = np.data() + i;
					rc = bt2g_sam_format(&o, &one, res.data(), ops.data(), maxOps, solo ? nullptr : pr.data(), out.data(), out.size(), &need);
					if(rc) break;
					want.append(out.data(), need);
					wantReads += solo ? 1 : 2;
					soloCount += solo;
				}
				if(rc) { printf("iteration %d: reference format failed %d\n", it, rc); bad++; break; }
				continue;
			}
			rc = bt2g_sam_format(&o, &rd, res.data(), ops.data(), maxOps, paired ? pr.data() : nullptr, out.data(), out.size(), &need);
			if(rc) { printf("iteration %d: reference format failed %d\n", it, rc); bad++; break; }
			want.append(out.data(), need);
			wantReads += n;
		}
		// the run, sometimes with a failing stage
		const int mode = (int)rnd(5);                       // 0..1 clean, 2 engine fails, 3 reader fails, 4 writer fails
		std::atomic<int> calls{0};
		std::vector<Eng> engs(E);
		std::vector<void *> handles(E);
		for(int j = 0; j < E; j++) { engs[j] = Eng{j, mode == 2 && nblocks ? (int)rnd(nblocks) : -1, &calls}; handles[j] = &engs[j]; }
		IO io;
		io.blocks = &blocks;
		if(mode == 3) io.failRead = (int)rnd(nblocks + 1);
		if(mode == 4) io.failWrite = (int)rnd(nblocks + 1);
		bt2g_stream_io sio{&io, nextBlock, writeOut, nullptr};
		Bytes bytes;
		const bool asBytes = mode <= 1 && rnd(2);              // clean runs: half of them through the read callback (the library cuts the blocks)
		if(asBytes) {
			streamed++;
			for(auto &bl : blocks) { bytes.text[0] += bl.first; bytes.text[1] += bl.second; }
			if(rnd(2) && !bytes.text[0].empty()) bytes.text[0].pop_back();              // a last record without a final newline
			if(rnd(3) == 0) bytes.text[paired ? 1 : 0] += "\n\n";                          // blank lines at the end
			bytes.rng = &rng; io.bytes = &bytes;
			sio.next_block = nullptr; sio.read = readBytes;
		}
		bt2g_stream_params sp;
		memset(&sp, 0, sizeof(sp));
		sp.paired = paired; sp.parse_threads = 1 + (int)rnd(4); sp.format_threads = 1 + (int)rnd(4); sp.depth = (int)rnd(4);
		sp.max_units = maxUnits; sp.max_len = maxLen; sp.max_ops = maxOps; sp.name_stride = stride;
		Eng soloEng{99, -1, &calls};
		if(withSolos) { sp.solo_engine = &soloEng; sp.solo_max_units = 1 + rnd(5); }
		sp.chunk_bytes = asBytes ? 50 + rnd(200000) : 0;        // (small chunks: records cut by the chunk end, chunks that grow)
		bt2g_align_counts cnt;
		memset(&cnt, 0, sizeof(cnt));
		uint64_t nReads = 0;
		char err[256];
		int rc = bt2g_stream_run(fakeAlign, handles.data(), E, &sp, &opt, &sio, &cnt, &nReads, err, sizeof(err));
		bool ok;
		solos += (long)soloCount;
		if(mode <= 1) ok = rc == 0 && io.out == want && nReads == wantReads && cnt.nread == (withSolos ? (wantReads + soloCount) / 2 : wantReads / (paired ? 2 : 1));
		else if(rc == 0) ok = io.out == want;               // the failure point was never reached (fewer non-empty blocks)
		else ok = (mode == 2 && rc == -77) || (mode == 3 && rc == -20) || (mode == 4 && rc == -25);
		if(ok && rc != 0) ok = want.compare(0, io.out.size(), io.out) == 0;   // what was written before the failure is a prefix, in order
		if(!ok) { printf("iteration %d: mode %d rc %d (%s): %zu bytes written, %zu expected\n", it, mode, rc, err, io.out.size(), want.size()); bad++; }
	}
	printf("%d iterations (%ld through the read callback, %ld solo reads), %ld inconsistencies\n", iters, streamed, solos, bad);
	return bad ? 1 : 0;
}
