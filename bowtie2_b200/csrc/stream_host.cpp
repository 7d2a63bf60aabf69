// stream_host.cpp -- FASTQ text in -> SAM text out around the engines, the whole batch loop in C++ (host code, no GPU work here).
//
// The batch loop of multiseedSearchWorker (bt2_search.cpp:3253-4254) with its reader (PatternComposer / PatternSourcePerThread,
// pat.cpp:222-300, :1050+) and its sink (AlnSinkWrap::finishRead -> AlnSinkSam::appendMate, aln_sink.cpp:643, :1889; the ordered
// output of --reorder, OutputQueue, outq.cpp) as four overlapped stages on host threads:
//
//     reader callback + parse (bt2g_fastq_parse_pairs_mt / bt2g_fastq_parse_mt)  ->  E aligner threads, one per engine
//     (bt2g_xengine_align: H2D, the device waves, D2H)  ->  format (bt2g_sam_format) + alignment counts + writer callback, in input order
//
// Every block in flight owns one slot of reused host buffers (parsed reads, names, results, edit ops); the SAM text of a block is
// written into one reused buffer and handed to the writer before the next block is formatted.  bowtie2_b200/stream.py is the same
// loop on Python threads.  Pairs with an empty mate 2 go through an unpaired engine of the same run (bt2g_stream_params.solo_engine), as the
// reference aligns their mate 1 with the unpaired policy (bt2_search.cpp:3326).
#include "../../include/bt2g.h"
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

// a reusable buffer WITHOUT value-initialisation (a std::vector would zero-fill hundreds of MB per slot before the parser / the engine
// overwrites them); need() may drop the contents, grow() keeps them
template <typename T>
struct Raw {
	T *p = nullptr;
	size_t cap = 0;
	Raw() = default;
	Raw(const Raw &) = delete;
	Raw &operator=(const Raw &) = delete;
	~Raw() { free(p); }
	T *data() { return p; }
	const T *data() const { return p; }
	T &operator[](size_t i) { return p[i]; }
	const T &operator[](size_t i) const { return p[i]; }
	size_t size() const { return cap; }
	bool need(size_t n) { if(n > cap) { free(p); p = (T *)malloc(n * sizeof(T)); cap = p ? n : 0; } return p != nullptr || n == 0; }
	bool grow(size_t n) { if(n > cap) { T *q = (T *)realloc(p, n * sizeof(T)); if(!q) return false; p = q; cap = n; } return true; }
};

struct Slot {
	Raw<uint8_t> seq, qual, ops;
	Raw<uint64_t> off;
	Raw<char> names;
	Raw<const char *> namePtrs;
	Raw<bt2g_read_result> res;
	Raw<bt2g_pair_result> pairs;
	uint64_t nReads = 0;
	// pairs of this block whose mate 2 is empty, and their mate 1s as a batch of unpaired reads with its results
	std::vector<uint64_t> soloIdx, soloOff;
	std::vector<uint8_t> soloSeq, soloQual, soloOps;
	std::vector<char> soloNames;
	std::vector<const char *> soloNamePtrs;
	std::vector<bt2g_read_result> soloRes;
};

struct Item { uint64_t k; int slot; };

// blocking queue; after abort() or close() + drained, pop() returns false
template <typename T>
struct Chan {
	std::mutex m;
	std::condition_variable cv;
	std::deque<T> q;
	bool closed = false;
	const std::atomic<bool> *failed = nullptr;
	void push(const T &v) { { std::lock_guard<std::mutex> l(m); q.push_back(v); } cv.notify_one(); }
	void close() { { std::lock_guard<std::mutex> l(m); closed = true; } cv.notify_all(); }
	bool pop(T &v) {
		std::unique_lock<std::mutex> l(m);
		cv.wait(l, [&] { return !q.empty() || closed || failed->load(); });
		if(failed->load() || q.empty()) return false;
		v = q.front(); q.pop_front();
		return true;
	}
};

struct Run {
	bt2g_stream_align_fn align;
	void *const *engines;
	int nEngines;
	bt2g_stream_params sp;
	const bt2g_sam_opts *opt;
	const bt2g_stream_io *io;
	bt2g_align_counts *counts;
	std::unique_ptr<Slot[]> slots;
	size_t nSlots = 0;
	Chan<int> freeSlots;
	Chan<Item> parsed, done;
	std::atomic<bool> failed{false};
	std::atomic<int> alignersLeft{0};
	std::mutex errM;
	int rc = 0;
	std::string err;
	bool cutOps = false;
	uint64_t nReads = 0;

	void fail(int code, const std::string &what) {
		{
			std::lock_guard<std::mutex> l(errM);
			if(rc == 0) { rc = code; err = what; }
		}
		failed.store(true);
		freeSlots.close(); parsed.close(); done.close();
	}

	// parse one block into slot si; n = reads parsed, c1 / c2 = bytes used of each text.  false = failed (fail() called).
	bool parseBlock(uint64_t k, int si, const char *t1, uint64_t l1, const char *t2, uint64_t l2, uint64_t &n, uint64_t &c1, uint64_t &c2) {
		const int mates = sp.paired ? 2 : 1;
		Slot &s = slots[si];
		const uint64_t maxReads = sp.max_units * mates, maxBases = l1 + l2 + 1;
		if(!s.seq.need(maxBases) || !s.qual.need(maxBases) || !s.off.need(maxReads + 1) || !s.names.need(maxReads * (uint64_t)sp.name_stride)) {
			fail(-12, "out of host memory"); return false;
		}
		n = c1 = c2 = 0;
		int prc;
		if(sp.paired) {
			prc = bt2g_fastq_parse_pairs_mt(t1, l1, t2, l2, sp.max_units, maxBases, s.seq.data(), s.qual.data(), s.off.data(), s.names.data(),
			                                sp.name_stride, &n, &c1, &c2, sp.parse_threads);
			n *= 2;
		} else {
			prc = bt2g_fastq_parse_mt(t1, l1, maxReads, maxBases, s.seq.data(), s.qual.data(), s.off.data(), s.names.data(), sp.name_stride, &n, &c1,
			                          sp.parse_threads);
			c2 = l2;
		}
		if(prc != 0) { fail(prc, "FASTQ parse failed in block " + std::to_string(k)); return false; }
		s.nReads = n;
		return true;
	}

	// the checks on a parsed block and its hand-over to the engines
	bool submit(uint64_t k, int si) {
		Slot &s = slots[si];
		const uint64_t n = s.nReads;
		s.soloIdx.clear();
		if(sp.paired) {
			for(uint64_t i = 0; i + 1 < n; i += 2) if(s.off[i + 2] == s.off[i + 1]) s.soloIdx.push_back(i / 2);
			if(!s.soloIdx.empty() && !sp.solo_engine) {
				fail(-23, "a pair with an empty mate 2 is an unpaired read for the reference (bt2_search.cpp:3326): the run needs a solo_engine");
				return false;
			}
		}
		for(uint64_t i = 0; i < n; i++) {
			if(s.off[i + 1] - s.off[i] > sp.max_len) { fail(-24, "a read is longer than the engines' max_len"); return false; }
			// the parser keeps name_stride - 1 bytes of a header line: a row filled to its last byte may have lost its tail (the name feeds
			// the read's random seed and the QNAME, so that is an error, not a silent cut)
			if(sp.name_stride >= 2 && s.names[i * (uint64_t)sp.name_stride + sp.name_stride - 2] != 0) {
				fail(-27, "a read name is longer than name_stride - 2 bytes");
				return false;
			}
		}
		if(!s.res.need(n) || !s.ops.need(n * (uint64_t)sp.max_ops) || !s.pairs.need(n / 2 + 1) || !s.namePtrs.need(n)) { fail(-12, "out of host memory"); return false; }
		for(uint64_t i = 0; i < n; i++) s.namePtrs[i] = s.names.data() + i * (uint64_t)sp.name_stride;
		parsed.push(Item{k, si});
		return true;
	}

	// blocks of whole records from next_block
	void readerBlocks() {
		for(uint64_t k = 0;; k++) {
			int si;
			if(!freeSlots.pop(si)) break;                        // (back-pressure: at most slots.size() blocks in flight)
			const char *t1 = nullptr, *t2 = nullptr;
			uint64_t l1 = 0, l2 = 0;
			int r = io->next_block(io->user, &t1, &l1, &t2, &l2);
			if(r < 0) { fail(-20, "the reader callback failed (" + std::to_string(r) + ")"); break; }
			if(r == 0) break;
			if(sp.paired && t2 == nullptr) { fail(-21, "paired run: the reader gave no mate-2 text"); break; }
			uint64_t n, c1, c2;
			if(!parseBlock(k, si, t1, l1, t2, l2, n, c1, c2)) break;
			if(c1 != l1 || c2 != l2) {
				fail(-22, "block " + std::to_string(k) + " does not hold whole records (or more than max_units, or mate files of different length): " +
				              std::to_string(l1 - c1) + " and " + std::to_string(l2 - c2) + " bytes left over");
				break;
			}
			if(!submit(k, si)) break;
		}
		parsed.close();
	}

	static bool blank(const Raw<char> &b, uint64_t have) {
		for(uint64_t i = 0; i < have; i++) if(b[i] != '\n' && b[i] != '\r' && b[i] != ' ' && b[i] != '\t') return false;
		return true;
	}

	// a byte stream per mate file from `read` (fread / gzread behind it): this thread keeps one text buffer per file, parses up to max_units
	// records from their fronts and moves what is left (the next records, a record cut by the chunk end) to the front for the next block --
	// the mate files are read in step by RECORD, so their lines need not have equal lengths (DualPatternComposer::nextBatch, pat.cpp:222-300)
	void readerStream() {
		const int nf = sp.paired ? 2 : 1;
		Raw<char> buf[2];
		uint64_t have[2] = {0, 0}, chunk = sp.chunk_bytes ? sp.chunk_bytes : (32ull << 20);
		bool eof[2] = {false, false};
		uint64_t k = 0;
		int si = -1;
		for(;;) {
			if(si < 0 && !freeSlots.pop(si)) break;
			for(int f = 0; f < nf; f++) {
				while(!eof[f] && have[f] < chunk) {
					if(!buf[f].grow(chunk + 1)) { fail(-12, "out of host memory"); parsed.close(); return; }
					int64_t r = io->read(io->user, f, buf[f].data() + have[f], chunk - have[f]);
					if(r < 0 || (uint64_t)r > chunk - have[f]) { fail(-20, "the read callback failed (" + std::to_string(r) + ")"); parsed.close(); return; }
					if(r == 0) {
						eof[f] = true;
						if(have[f] && buf[f][have[f] - 1] != '\n') buf[f][have[f]++] = '\n';      // a last record without a final newline
					}
					have[f] += (uint64_t)r;
				}
			}
			const bool done1 = eof[0] && blank(buf[0], have[0]), done2 = nf == 2 ? (eof[1] && blank(buf[1], have[1])) : true;
			if(done1 && done2) break;                                // end of the input
			if(nf == 2 && (done1 || done2)) {
				// DualPatternComposer::nextBatch (pat.cpp:256-290)
				fail(-26, std::string("Error, fewer reads in file specified with -") + (done1 ? "1" : "2") + " than in file specified with -" + (done1 ? "2" : "1"));
				break;
			}
			uint64_t n, c1, c2;
			if(!parseBlock(k, si, buf[0].data(), have[0], nf == 2 ? buf[1].data() : nullptr, nf == 2 ? have[1] : 0, n, c1, c2)) break;
			if(n == 0) {
				if(eof[0] && (nf == 1 || eof[1])) { fail(-22, "truncated FASTQ record at the end of the input"); break; }
				chunk *= 2;                                          // a record longer than the chunk: read more
				continue;
			}
			if(!submit(k, si)) break;
			k++; si = -1;
			// the text kept per file follows the records: a little more than max_units records, so that the blocks are full and little text is
			// left to move to the front
			const uint64_t units = sp.paired ? n / 2 : n, want = (std::max(c1, c2) / units + 1) * sp.max_units;
			chunk = want + want / 16 + 4096;
			memmove(buf[0].data(), buf[0].data() + c1, have[0] - c1); have[0] -= c1;
			if(nf == 2) { memmove(buf[1].data(), buf[1].data() + c2, have[1] - c2); have[1] -= c2; }
		}
		parsed.close();
	}

	void reader() { if(io->read) readerStream(); else readerBlocks(); }

	// `paired = !read_b().empty()` (bt2_search.cpp:3326): mate 1 of a pair whose mate 2 is empty goes through the UNPAIRED policy and leaves
	// one record.  One unpaired engine, shared by the aligner threads (such pairs are rare).
	std::mutex soloM;
	bool alignSolos(Slot &s, uint64_t k) {
		const uint64_t m = s.soloIdx.size(), stride = sp.name_stride;
		s.soloOff.assign(m + 1, 0);
		for(uint64_t j = 0; j < m; j++) { const uint64_t i = 2 * s.soloIdx[j]; s.soloOff[j + 1] = s.soloOff[j] + (s.off[i + 1] - s.off[i]); }
		s.soloSeq.resize(s.soloOff[m] + 1); s.soloQual.resize(s.soloOff[m] + 1); s.soloNames.resize(m * stride); s.soloNamePtrs.resize(m);
		s.soloRes.resize(m); s.soloOps.assign(m * (uint64_t)sp.max_ops, 0);
		for(uint64_t j = 0; j < m; j++) {
			const uint64_t i = 2 * s.soloIdx[j], l = s.off[i + 1] - s.off[i];
			memcpy(s.soloSeq.data() + s.soloOff[j], s.seq.data() + s.off[i], l);
			memcpy(s.soloQual.data() + s.soloOff[j], s.qual.data() + s.off[i], l);
			memcpy(s.soloNames.data() + j * stride, s.names.data() + i * stride, stride);
			s.soloNamePtrs[j] = s.soloNames.data() + j * stride;
		}
		const uint64_t cap = sp.solo_max_units ? sp.solo_max_units : sp.max_units;
		std::vector<uint64_t> off;
		std::lock_guard<std::mutex> l(soloM);
		for(uint64_t a = 0; a < m; a += cap) {
			const uint64_t b = std::min(m, a + cap);
			off.assign(s.soloOff.begin() + a, s.soloOff.begin() + b + 1);
			for(auto &o : off) o -= s.soloOff[a];                    // (a batch of its own: offsets from 0)
			bt2g_reads rd;
			rd.n_reads = b - a; rd.seq = s.soloSeq.data() + s.soloOff[a]; rd.qual = s.soloQual.data() + s.soloOff[a]; rd.off = off.data();
			int r = align(sp.solo_engine, &rd, s.soloNames.data() + a * stride, sp.name_stride, s.soloRes.data() + a, s.soloOps.data() + a * (uint64_t)sp.max_ops,
			              sp.max_ops, nullptr, nullptr);
			if(r != 0) { fail(r, "the solo engine failed on block " + std::to_string(k)); return false; }
		}
		return true;
	}

	void aligner(int j) {
		Item it;
		while(parsed.pop(it)) {
			Slot &s = slots[it.slot];
			bt2g_reads rd;
			rd.n_reads = s.nReads; rd.seq = s.seq.data(); rd.qual = s.qual.data(); rd.off = s.off.data();
			int r = s.nReads == 0 ? 0 : align(engines[j], &rd, s.names.data(), sp.name_stride, s.res.data(), s.ops.data(), sp.max_ops,
			                                 sp.paired ? s.pairs.data() : nullptr, nullptr);
			if(r != 0) { fail(r, "engine " + std::to_string(j) + " failed on block " + std::to_string(it.k)); break; }
			if(!s.soloIdx.empty() && !alignSolos(s, it.k)) break;
			done.push(it);
		}
		if(alignersLeft.fetch_sub(1) == 1) done.close();
	}

	// format one run of reads, add it to the counts, hand it to the writer callback
	bool emit(Raw<char> &out, const bt2g_reads &rd, const char *const *names, const bt2g_read_result *res, const uint8_t *ops,
	          const bt2g_pair_result *pr, uint64_t block) {
		bt2g_sam_opts o = *opt;
		o.read_names = names;
		o.threads = sp.format_threads;
		// one formatting pass in the common case (SEQ + QUAL + ~260 bytes of fields per record); -3 reports the size needed
		uint64_t cap = (rd.off[rd.n_reads] - rd.off[0]) * 2 + rd.n_reads * 260 + 4096, need = 0;
		if(!out.need(cap)) { fail(-12, "out of host memory"); return false; }
		int r = bt2g_sam_format(&o, &rd, res, ops, sp.max_ops, pr, out.data(), out.size(), &need);
		if(r == -3) {
			if(!out.need(need)) { fail(-12, "out of host memory"); return false; }
			r = bt2g_sam_format(&o, &rd, res, ops, sp.max_ops, pr, out.data(), out.size(), &need);
		}
		if(r < 0) { fail(r, "bt2g_sam_format failed on block " + std::to_string(block)); return false; }
		if(r == 1) cutOps = true;
		if(counts) bt2g_align_counts_add_ex(counts, res, rd.n_reads, pr, sp.count_flags);
		int w = io->write(io->user, out.data(), need);
		if(w != 0) { fail(-25, "the writer callback failed (" + std::to_string(w) + ")"); return false; }
		return true;
	}

	void writer() {
		std::map<uint64_t, int> pending;
		Raw<char> out;
		uint64_t next = 0;
		Item it;
		while(done.pop(it)) {
			pending[it.k] = it.slot;
			for(auto p = pending.find(next); p != pending.end(); p = pending.find(next)) {
				const int si = p->second;
				pending.erase(p);
				Slot &s = slots[si];
				if(s.nReads) {
					// the block as runs of ordinary pairs with the solo reads between them, in input order (one run = the block without solos)
					const uint64_t units = sp.paired ? s.nReads / 2 : s.nReads, per = sp.paired ? 2 : 1, nSolo = s.soloIdx.size();
					uint64_t prev = 0;
					for(uint64_t j = 0; j <= nSolo; j++) {
						const uint64_t end = j < nSolo ? s.soloIdx[j] : units;
						if(end > prev) {
							bt2g_reads rd;
							rd.n_reads = (end - prev) * per; rd.seq = s.seq.data(); rd.qual = s.qual.data(); rd.off = s.off.data() + prev * per;
							if(!emit(out, rd, s.namePtrs.data() + prev * per, s.res.data() + prev * per, s.ops.data() + prev * per * (uint64_t)sp.max_ops,
							         sp.paired ? s.pairs.data() + prev : nullptr, next)) return;
						}
						if(j < nSolo) {
							bt2g_reads rd;
							rd.n_reads = 1; rd.seq = s.soloSeq.data(); rd.qual = s.soloQual.data(); rd.off = s.soloOff.data() + j;
							if(!emit(out, rd, s.soloNamePtrs.data() + j, s.soloRes.data() + j, s.soloOps.data() + j * (uint64_t)sp.max_ops, nullptr, next)) return;
							prev = end + 1;
						}
					}
					nReads += s.nReads - nSolo;                      // (a solo pair leaves one record)
				}
				next++;
				freeSlots.push(si);
			}
		}
	}
};

} // namespace

extern "C" int bt2g_stream_run(bt2g_stream_align_fn align, void *const *engines, int32_t n_engines, const bt2g_stream_params *sp,
                               const bt2g_sam_opts *opt, const bt2g_stream_io *io, bt2g_align_counts *counts, uint64_t *n_reads,
                               char *err, uint32_t err_cap) {
	if(err && err_cap) err[0] = 0;
	if(!align || !engines || n_engines < 1 || !sp || !opt || !io || (!io->next_block && !io->read) || !io->write || sp->max_units == 0 || sp->name_stride == 0) {
		if(err && err_cap) snprintf(err, err_cap, "bt2g_stream_run: bad arguments");
		return -1;
	}
	Run R;
	R.align = align; R.engines = engines; R.nEngines = n_engines; R.sp = *sp; R.opt = opt; R.io = io; R.counts = counts;
	if(R.sp.parse_threads < 1) R.sp.parse_threads = 1;
	if(R.sp.format_threads < 1) R.sp.format_threads = 1;
	const int depth = sp->depth > 0 ? sp->depth : 2;
	R.nSlots = (size_t)(depth + n_engines + 1);
	R.slots = std::make_unique<Slot[]>(R.nSlots);
	R.freeSlots.failed = R.parsed.failed = R.done.failed = &R.failed;
	for(size_t i = 0; i < R.nSlots; i++) R.freeSlots.push((int)i);
	R.alignersLeft.store(n_engines);
	std::vector<std::thread> th;
	th.emplace_back([&] { R.reader(); });
	for(int j = 0; j < n_engines; j++) th.emplace_back([&R, j] { R.aligner(j); });
	th.emplace_back([&] { R.writer(); });
	for(auto &t : th) t.join();
	if(n_reads) *n_reads = R.nReads;
	if(R.rc != 0) {
		if(err && err_cap) snprintf(err, err_cap, "%s", R.err.c_str());
		return R.rc;
	}
	return R.cutOps ? 1 : 0;
}
