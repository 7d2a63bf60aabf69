"""FASTQ text in -> SAM text out around the device engine, with the host stages overlapped on threads:

    parse (bt2g_fastq_parse_mt / bt2g_fastq_parse_pairs_mt: the two mate files straight into one interleaved batch)
      ||  align (bt2g_xengine_align, one host thread per engine)  ||  format (bt2g_sam_format)

Every stage is one call into libbt2g.so per batch (ctypes releases the GIL), so the threads run concurrently; batches leave in
input order.  This is the batch loop of multiseedSearchWorker (bt2_search.cpp:3253-4254) with its reader
(PatternSourcePerThread, pat.cpp) and its sink (AlnSinkSam, aln_sink.cpp:1889) -- host plumbing: nothing here computes
alignments."""
import inspect
import queue
import threading

import numpy as np

from .lib import HostBuffers, NameTable, ReadBatch, XEngine, fastq_parse, fastq_parse_pairs, load_library, sam_format


def interleave_uniform(b1: ReadBatch, b2: ReadBatch, n1: NameTable, n2: NameTable):
    """mate 1 of pair i -> read 2i, mate 2 -> read 2i + 1; fast path for batches whose reads all have one length per file"""
    n = b1.n
    if n != b2.n:
        raise ValueError(f"mate files differ in length within a batch ({b1.n} vs {b2.n} records)")
    l1, l2 = b1.lengths(), b2.lengths()
    if n and (l1 == l1[0]).all() and (l2 == l2[0]).all() and l1[0] == l2[0]:
        L = int(l1[0])
        seq = np.empty((n, 2, L), dtype=np.uint8)
        qual = np.empty((n, 2, L), dtype=np.uint8)
        seq[:, 0], seq[:, 1] = b1.seq[:n * L].reshape(n, L), b2.seq[:n * L].reshape(n, L)
        qual[:, 0], qual[:, 1] = b1.qual[:n * L].reshape(n, L), b2.qual[:n * L].reshape(n, L)
        batch = ReadBatch(seq.reshape(-1), np.arange(0, (2 * n + 1) * L, L, dtype=np.uint64), qual.reshape(-1))
    else:
        from .align import interleave
        batch = interleave(b1, b2)
    rows = np.empty((2 * n, n1.rows.shape[1]), dtype=np.uint8)
    rows[0::2], rows[1::2] = n1.rows, n2.rows
    return batch, NameTable(rows)


class TextAligner:
    """engines: list of XEngine (all created with the same parameters); ref_names: @SQ names in index order.
    Every batch in flight owns one set of reused host buffers (lib.HostBuffers: parsed reads, names, results), and the SAM text
    of a batch is handed to the sink as a memoryview of one reused output buffer: the sink must consume it (write it) before it
    returns.  No per-batch allocation is left on the steady-state path."""

    def __init__(self, engines, ref_names, paired, local=False, parse_threads=4, format_threads=8, name_stride=32, depth=2, no_discordant=False):
        self.engines, self.ref_names, self.paired, self.local = list(engines), list(ref_names), paired, local
        self.parse_threads, self.format_threads, self.name_stride, self.depth = parse_threads, format_threads, name_stride, depth
        self.no_discordant = no_discordant                       # the engines' --no-discordant, for the record formatter
        self.lib = load_library()
        self._slots = [HostBuffers() for _ in range(depth + len(self.engines) + 1)]
        # (engine stand-ins of the CPU tests may not take reusable result buffers)
        self._reuse = [("out" in inspect.signature(e.align).parameters) for e in self.engines]
        self._out = HostBuffers()

    def _parse(self, item, slot):
        t1, t2 = item
        if not self.paired:
            b1, n1, used1 = fastq_parse(self.lib, t1, name_stride=self.name_stride, threads=self.parse_threads, out=slot)
            if used1 != len(t1):
                raise ValueError(f"FASTQ text of a batch ends inside a record (byte {used1} of {len(t1)})")
            return b1, n1
        batch, names, used1, used2 = fastq_parse_pairs(self.lib, t1, t2, name_stride=self.name_stride, threads=self.parse_threads, out=slot)
        if used1 != len(t1) or used2 != len(t2):
            raise ValueError(f"mate files differ in length within a batch ({batch.n // 2} whole pairs; {len(t1) - used1} and {len(t2) - used2} bytes left over)")
        return batch, names

    def run(self, items, sink):
        """items: iterable of (mate-1 FASTQ text, mate-2 FASTQ text or None), each at most one engine batch; sink(view) is called once
        per item, in input order, with the SAM text as a memoryview that is valid until the sink returns.  Returns the number of
        records written."""
        q_free, q_parsed, q_done = queue.Queue(), queue.Queue(), queue.Queue()
        for slot in self._slots:
            q_free.put(slot)
        errs, total = [], [0]
        END = object()

        def parser():
            try:
                for k, item in enumerate(items):
                    slot = q_free.get()                          # (back-pressure: at most len(slots) batches in flight)
                    if errs:
                        break
                    q_parsed.put((k, slot, *self._parse(item, slot)))
            except Exception as e:
                errs.append(e)
            for _ in self.engines:
                q_parsed.put(END)

        def aligner(eng, reuse):
            try:
                while True:
                    w = q_parsed.get()
                    if w is END:
                        break
                    k, slot, batch, names = w
                    res, ops, pairs, _ = eng.align(batch, names, out=slot) if reuse else eng.align(batch, names)
                    q_done.put((k, slot, batch, names, res, ops, pairs))
            except Exception as e:
                errs.append(e)
                q_free.put(HostBuffers())                        # never leave the parser waiting
            q_done.put(END)

        def formatter():
            try:
                pending, nxt, ended = {}, 0, 0
                while ended < len(self.engines):
                    w = q_done.get()
                    if w is END:
                        ended += 1
                        continue
                    pending[w[0]] = w[1:]
                    while nxt in pending:
                        slot, batch, names, res, ops, pairs = pending.pop(nxt)
                        txt = sam_format(self.lib, batch, res, ops, self.ref_names, read_names=names, pairs=pairs, threads=self.format_threads,
                                         local=self.local, as_bytes="view", out=self._out, no_discordant=self.no_discordant)
                        sink(txt)
                        total[0] += batch.n
                        nxt += 1
                        del batch, names, res, ops, pairs
                        q_free.put(slot)
            except Exception as e:
                errs.append(e)
                for _ in range(len(self._slots)):
                    q_free.put(HostBuffers())

        th = [threading.Thread(target=parser)] + [threading.Thread(target=aligner, args=(e, r)) for e, r in zip(self.engines, self._reuse)] + [threading.Thread(target=formatter)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]
        return total[0]
