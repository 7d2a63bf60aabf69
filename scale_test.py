import sys, time, torch, os
sys.path.insert(0, '.')
os.environ["BT2G_VERBOSE"] = "1"
from bench import make_genome_gpu
from bowtie2_b200.index_build import build_index
dev = torch.device("cuda", 0)
for mbp in [float(x) for x in sys.argv[1:]]:
    t0 = time.time()
    contigs = make_genome_gpu(torch, dev, 24, int(mbp * 1e6 / 24))
    torch.cuda.synchronize()
    print(f"genome {mbp} Mbp: {time.time()-t0:.1f}s", file=sys.stderr, flush=True)
    t0 = time.time()
    b = build_index(contigs)
    torch.cuda.synchronize()
    print(f"index {mbp} Mbp: {time.time()-t0:.1f}s; peak mem {torch.cuda.max_memory_allocated()/1e9:.1f} GB", file=sys.stderr, flush=True)
    del b, contigs
    torch.cuda.empty_cache()
