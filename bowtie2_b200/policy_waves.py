"""Wave scheduler for the exact search policy: many reads' step generators (policy_engine.read_steps / pair_steps) advance
together; at every wave each read is blocked on exactly one primitive request, the requests are grouped by primitive and
answered by ONE batched backend call per group.  This is the execution shape the GPU wants (large batches per kernel, the
sequential, RNG-driven control flow kept per read) and the blueprint for the device-side version, where the per-read
state machine is itself a kernel and the request lists are device queues.

The batched backend answers `batch(name, requests, ids)` -> list of results, `requests` being the argument tuples the
engine yielded and `ids` the read (or pair) numbers they belong to.  `ItemwiseBatch` adapts any per-item backend."""
from collections import defaultdict


class ItemwiseBatch:
    """a per-item backend (one call per request) behind the batched interface"""

    def __init__(self, backend):
        self.b = backend
        self.off_size = backend.off_size

    def batch(self, name, requests, ids):
        f = getattr(self.b, name)
        return [f(*a) for a in requests]


class WaveScheduler:
    def __init__(self, batched_backend, make_engine, max_inflight=1 << 16):
        """make_engine() -> a fresh PolicyEngine / PairedPolicyEngine with backend=None (one per read in flight)"""
        self.bb = batched_backend
        self.make_engine = make_engine
        self.max_inflight = max_inflight
        self.n_waves = 0
        self.n_calls = defaultdict(int)          # batched backend calls per primitive
        self.n_requests = defaultdict(int)       # requests per primitive

    def _run(self, make_gen, n):
        results = [None] * n
        next_item = 0
        active = {}                               # id -> (generator, pending request)

        def start(i):
            eng = self.make_engine()
            eng.off_size = self.bb.off_size
            g = make_gen(eng, i)
            try:
                active[i] = (g, next(g))
            except StopIteration as e:
                results[i] = e.value

        while next_item < n or active:
            while next_item < n and len(active) < self.max_inflight:
                start(next_item)
                next_item += 1
            if not active:
                continue
            groups = defaultdict(list)
            for i, (g, req) in active.items():
                groups[req[0]].append(i)
            self.n_waves += 1
            for name, ids in groups.items():
                answers = self.bb.batch(name, [active[i][1][1] for i in ids], ids)
                self.n_calls[name] += 1
                self.n_requests[name] += len(ids)
                for i, ans in zip(ids, answers):
                    g = active[i][0]
                    try:
                        active[i] = (g, g.send(ans))
                    except StopIteration as e:
                        results[i] = e.value
                        del active[i]
        return results

    def run_reads(self, reads, quals, names):
        return self._run(lambda eng, i: eng.read_steps(reads[i], quals[i], names[i]), len(reads))

    def run_pairs(self, reads, quals, names):
        """reads / quals / names interleaved: mate 1 of pair i at 2i, mate 2 at 2i + 1"""
        return self._run(lambda eng, i: eng.pair_steps(reads[2 * i], quals[2 * i], names[2 * i], reads[2 * i + 1], quals[2 * i + 1],
                                                       names[2 * i + 1]), len(reads) // 2)
